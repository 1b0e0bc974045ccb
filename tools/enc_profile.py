"""Per-kernel timing of the batched CELT encoder (CUDA events around each launch are not available through the C ABI, so this
times whole encode_device calls for several complexities and signal types; use under ncu for the per-kernel split)."""
import argparse, ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import opus_b200, util
from opus_b200 import capi

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=65536)
ap.add_argument('--reps', type=int, default=5)
ap.add_argument('--cx', type=int, nargs='+', default=[0, 4, 5, 10])
ap.add_argument('--sig', default='harmonic')
a = ap.parse_args()
L = capi.lib()
dev = torch.device('cuda:0')
n, base, nf = a.n, 32, 4
gen = {'harmonic': util.sig_harmonic, 'music': util.sig_music, 'white': util.sig_white}[a.sig]
xb = np.stack([gen(s, 960 * nf, 2).astype(np.float32).reshape(nf, 1920) / 32768.0 for s in range(base)])
frames = [torch.from_numpy(np.tile(xb[:, f], ((n + base - 1) // base, 1))[:n].copy()).to(dev) for f in range(nf)]
pk = torch.empty((n, 160), dtype=torch.uint8, device=dev)
nbd = torch.empty(n, dtype=torch.int32, device=dev)
dptr = lambda t: ctypes.c_void_p(t.data_ptr())
for cx in a.cx:
    enc = opus_b200.EncoderBatch(n, channels=2, complexity=cx)
    st = torch.cuda.ExternalStream(enc.cuda_stream)
    with torch.cuda.stream(st):
        for f in range(nf):   # warm state
            L.opusb200_encode_device(enc._h, dptr(frames[f]), 960, dptr(pk), 160, 160, dptr(nbd))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for r in range(a.reps):
            L.opusb200_encode_device(enc._h, dptr(frames[r % nf]), 960, dptr(pk), 160, 160, dptr(nbd))
        e1.record(st)
        e1.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    print('complexity %d  %s  n=%d  %.3f ms/frame-batch  %.2f M frames/s' % (cx, a.sig, n, ms, n / ms / 1e3))
    enc.close()
