#!/usr/bin/env python3
"""Forward MDCT (960-point long block) on the GPU: device time per launch (CUDA events on the library stream), algorithmic GB/s
(4320 B in + 3840 B out per transform) against MEASURED_PEAKS.json.  OPUSB200_MDCT_SMEM_FFT=1 selects the general shared-memory kernel.
Usage: tools/mdct_bench.py [--n 131072] [--reps 20]"""
import argparse, ctypes, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import opus_b200

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=131072)   # 65 536 stereo frames
ap.add_argument('--reps', type=int, default=20)
a = ap.parse_args()
ops = opus_b200.Ops(0)
L = ops._L
stream = torch.cuda.ExternalStream(L.opusb200_ctx_cuda_stream(ops._h))
dev = torch.device('cuda', 0)
x = torch.randn((a.n, 1080), dtype=torch.float32, device=dev)
o = torch.empty((a.n, 960), dtype=torch.float32, device=dev)
p = lambda t: ctypes.c_void_p(t.data_ptr())
fn = lambda: L.opusb200_mdct_forward(ops._h, p(x), p(o), a.n, 3, 0, 1)
with torch.cuda.stream(stream):
    for _ in range(3):
        fn()
    stream.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); stream.synchronize()
        ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
peak = 6650.0
pp = os.path.join(ROOT, 'MEASURED_PEAKS.json')
if os.path.exists(pp):
    peak = float(json.load(open(pp))['hbm_gbs'])
gbs = a.n * 8160 / (ms * 1e-3) / 1e9
print(json.dumps({'kernel': 'general (smem FFT)' if os.environ.get('OPUSB200_MDCT_SMEM_FFT') else 'ob_mdct_forward_tma', 'transforms': a.n, 'ms_per_launch': ms,
                  'min_ms': float(min(ts)), 'algorithmic_bytes': a.n * 8160, 'achieved_GBs': gbs, 'peak_GBs': peak, 'frac': gbs / peak,
                  'working_set_MB': a.n * 8160 / 1e6, 'note': 'working set exceeds the 126 MB L2'}))
