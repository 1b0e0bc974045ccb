#!/usr/bin/env python
"""Device-resident timing of the SILK frame analysis kernels of round 2 (gain quantiser, NLSF quantiser, LPC estimation, the whole prediction analysis,
noise shaping analysis) and of the LPCNet feature extraction / DRED encoder, with the reference's own function timed beside each SILK kernel on one
host core (oracle/_ref when present).  One JSON line per op.      python tools/bench_silk_analysis.py [--n 65536] [--reps 5]"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import opus_b200  # noqa: E402
import util  # noqa: E402
from bench_ops import time_gpu, time_cpu, dptr  # noqa: E402
from test_silk_gains_oracle_vs_ref import Gains, random_frame as gains_frame  # noqa: E402
from test_silk_nlsf_oracle_vs_ref import Nlsf, random_frame as nlsf_frame  # noqa: E402
from test_silk_find_lpc_oracle_vs_ref import FindLpc, lpc_input, random_case as lpc_case  # noqa: E402
from test_silk_pred_coefs_oracle_vs_ref import PredCoefs, random_case as pc_case, HIST  # noqa: E402
from test_silk_noise_shape_oracle_vs_ref import NoiseShape, frame_input, random_case as ns_case  # noqa: E402


def tile_records(recs, n):
    raw = np.frombuffer(bytes(recs), np.uint8).reshape(len(recs), -1)
    return torch.from_numpy(np.tile(raw, ((n + len(recs) - 1) // len(recs), 1))[:n].copy()).cuda()


def tile_rows(a, n):
    return torch.from_numpy(np.tile(a, ((n + a.shape[0] - 1) // a.shape[0], 1))[:n].copy()).cuda()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=65536)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--only', default='', help="'vad_pitch' = only the voice activity detector and the pitch analysis")
    a = ap.parse_args()
    n, base = a.n, 128
    ops = opus_b200.Ops(0)
    L = ops._L
    stream = torch.cuda.ExternalStream(L.opusb200_ctx_cuda_stream(ops._h))
    R = util.load_ref('c') if util.ref_available('c') else None
    rng = np.random.default_rng(0)

    def emit(op, ms, cpu_s):
        rec = {'op': op, 'n': n, 'unit': 'stream-frames (20 ms, 16 kHz)', 'ms': round(ms, 4), 'units_per_s': n / (ms * 1e-3), 'cpu_note': 'one call through ctypes (adds about a microsecond)',
               'cpu_ref_units_per_s_1core': (1.0 / cpu_s) if cpu_s else None}
        if cpu_s:
            rec['speedup_vs_1core'] = rec['units_per_s'] * cpu_s
        print(json.dumps(rec), flush=True)


    # ---- voice activity detector and pitch analysis ----
    from test_silk_vad_oracle_vs_ref import Vad, vad_signal
    from test_silk_pitch_oracle_vs_ref import Pitch, pitch_input, random_case as pitch_case
    v = (Vad * base)()
    ops.silk_vad_init(v)
    pcm = np.stack([vad_signal(rng, 16, 320, 2)[1] for _ in range(base)])
    d, dx = tile_records(v, n), tile_rows(pcm, n)
    ms = time_gpu(stream, lambda: L.opusb200_silk_encode_do_vad(ops._h, dptr(d), dptr(dx), 320, None, n, 320, 16, 1), a.reps)
    one = Vad.from_buffer_copy(bytes(v[3]))
    emit('silk_encode_do_VAD_FLP (silk_VAD_GetSA_Q8)', ms,
         time_cpu(lambda: R.refh_silk_encode_do_vad(ctypes.byref(one), pcm[3].ctypes.data_as(ctypes.c_void_p), 320, 16, 1)) if R else None)
    rec = {'op': 'silk_encode_do_VAD_FLP roofline', 'bytes_per_stream_frame': 2 * 320 + 2 * ctypes.sizeof(Vad), 'achieved_GBps': (2 * 320 + 2 * ctypes.sizeof(Vad)) * n / (ms * 1e-3) / 1e9}
    print(json.dumps(rec), flush=True)
    pt = (Pitch * base)()
    X_ = np.stack([pitch_input(rng, 16, 4, s % 3) for s in range(base)])
    for s in range(base):
        pt[s] = pitch_case(rng, 16, 1, (0, 0, np.float32(0)))
    buf_len = (22 + 5 * 4) * 16
    d, dx = tile_records(pt, n), tile_rows(X_, n)
    dres = torch.empty((n, buf_len), dtype=torch.float32, device='cuda')
    ms = time_gpu(stream, lambda: L.opusb200_silk_find_pitch_lags(ops._h, dptr(d), dptr(dres), buf_len, dptr(dx), X_.shape[1], 320, n, 16, 4, 16, 2, 1), a.reps)
    rr = np.zeros(buf_len, np.float32)
    emit('silk_find_pitch_lags_FLP (complexity 2)', ms,
         time_cpu(lambda: R.refh_silk_find_pitch_lags(ctypes.byref(Pitch.from_buffer_copy(bytes(pt[3]))), rr.ctypes.data_as(ctypes.c_void_p),
                                                      X_[3, 320:].ctypes.data_as(ctypes.c_void_p), 16, 4, 16, 2)) if R else None)
    if a.only == 'vad_pitch':
        return

    g = (Gains * base)()
    for s in range(base):
        g[s] = gains_frame(rng, 10, 1)
    d = tile_records(g, n)
    ms = time_gpu(stream, lambda: L.opusb200_silk_process_gains(ops._h, dptr(d), n, 4, 80, 4, 1), a.reps)
    emit('silk_process_gains', ms, time_cpu(lambda: R.refh_silk_process_gains(ctypes.byref(Gains.from_buffer_copy(bytes(g[3]))), 4, 80, 4)) if R else None)

    q = (Nlsf * base)()
    for s in range(base):
        q[s] = nlsf_frame(rng, 1)
        q[s].NLSF_MSVQ_Survivors = 16
    d = tile_records(q, n)
    ms = time_gpu(stream, lambda: L.opusb200_silk_process_nlsfs(ops._h, dptr(d), n, 16, 4, 1), a.reps)
    emit('silk_process_NLSFs (16 survivors)', ms, time_cpu(lambda: R.refh_silk_process_NLSFs(ctypes.byref(Nlsf.from_buffer_copy(bytes(q[3]))), 1, 4)) if R else None)

    f = (FindLpc * base)()
    x = np.stack([lpc_input(rng, 16, 4, 80, s % 3) for s in range(base)])
    for s in range(base):
        f[s] = lpc_case(rng, 16, 1)
        f[s].useInterpolatedNLSFs = 1
    d, dx = tile_records(f, n), tile_rows(x, n)
    ms = time_gpu(stream, lambda: L.opusb200_silk_find_lpc(ops._h, dptr(d), dptr(dx), n, 16, 4, 80, 1), a.reps)
    emit('silk_find_LPC_FLP (interpolation search)', ms,
         time_cpu(lambda: R.refh_silk_find_LPC(ctypes.byref(FindLpc.from_buffer_copy(bytes(f[3]))), x[3].ctypes.data_as(ctypes.c_void_p), 16, 4, 80)) if R else None)

    pc, Rr, Xx = (PredCoefs * base)(), [], []
    for s in range(base):
        pc[s], res, xx = pc_case(rng, 16, 4, 80, 16, 1, [0] * 16, 0)
        pc[s].NLSF_MSVQ_Survivors = 16
        Rr.append(res); Xx.append(xx)
    Rr, Xx = np.stack(Rr), np.stack(Xx)
    d, dr, dx = tile_records(pc, n), tile_rows(Rr, n), tile_rows(Xx, n)
    ms = time_gpu(stream, lambda: L.opusb200_silk_find_pred_coefs(ops._h, dptr(d), dptr(dr), dptr(dx), n, Xx.shape[1], HIST, 16, 4, 1), a.reps)
    emit('silk_find_pred_coefs_FLP (60 % voiced)', ms,
         time_cpu(lambda: R.refh_silk_find_pred_coefs(ctypes.byref(PredCoefs.from_buffer_copy(bytes(pc[3]))), Rr[3, HIST:].ctypes.data_as(ctypes.c_void_p),
                                                      Xx[3, HIST:].ctypes.data_as(ctypes.c_void_p), 16, 4, 80)) if R else None)

    ns, Rr, Xx = (NoiseShape * base)(), [], []
    for s in range(base):
        res, xx, la = frame_input(rng, 16, 4, s % 3)
        ns[s] = ns_case(rng, 16, (np.float32(0), np.float32(0)))
        ns[s].warping_Q16 = 16 * 983
        Rr.append(res); Xx.append(xx)
    Rr, Xx = np.stack(Rr), np.stack(Xx)
    d, dr, dx = tile_records(ns, n), tile_rows(Rr, n), tile_rows(Xx, n)
    ms = time_gpu(stream, lambda: L.opusb200_silk_noise_shape_analysis(ops._h, dptr(d), dptr(dr), Rr.shape[1], dptr(dx), Xx.shape[1], 80, n, 16, 4, 24, 1), a.reps)
    emit('silk_noise_shape_analysis_FLP (order 24, warped)', ms,
         time_cpu(lambda: R.refh_silk_noise_shape_analysis(ctypes.byref(NoiseShape.from_buffer_copy(bytes(ns[3]))), Rr[3].ctypes.data_as(ctypes.c_void_p),
                                                           Xx[3, 80:].ctypes.data_as(ctypes.c_void_p), 16, 4, 24)) if R else None)

    # ---- LPCNet feature extraction and the DRED encoder (weights: the seeded random models of oracle/_ref/libfargan_ref.so) ----
    import fargan_util as F
    from opus_b200 import nnet
    RF = F.load_ref()
    if RF is not None and hasattr(RF, 'refh_dred_stats'):
        nn = min(n, 8192)
        renc = F.RefLpcnetEnc(RF)
        lin, _ = renc.linear()
        cv = renc.conv()
        enc = nnet.LpcnetEncoder(ops, lin, cv, nn)
        pcm = tile_rows(np.stack([np.round(F.speech_like(s, 1, 9000.0)).astype(np.int16) for s in range(64)]), nn)
        feat = torch.empty((nn, 36), dtype=torch.float32, device='cuda')
        ms = time_gpu(stream, lambda: L.opusb200_lpcnet_compute_single_frame_features(enc._h, dptr(pcm), 1, dptr(feat), 1), a.reps)
        ref_s = time_cpu(lambda: renc.features(F.speech_like(5, 1, 9000.0)))
        print(json.dumps({'op': 'lpcnet_compute_single_frame_features', 'n': nn, 'unit': 'stream-frames (10 ms, 16 kHz)', 'ms': round(ms, 4),
                          'units_per_s': nn / (ms * 1e-3), 'cpu_ref_units_per_s_1core': 1.0 / ref_s, 'speedup_vs_1core': nn / (ms * 1e-3) * ref_s}), flush=True)
        layers, _ = F.RefDredEnc(RF).layers()
        stats, keep = F.dred_stats(RF)
        de = nnet.DredEncoder(ops, layers, nn)
        de.init_frontend(48000, 2, lin, cv)
        de.load_stats(stats)
        x48 = tile_rows(np.stack([F.dred_input(s, 48000, 2, 960) for s in range(64)]), nn)
        act = torch.ones((nn, 416), dtype=torch.uint8, device='cuda')
        pay = torch.empty((nn, 256), dtype=torch.uint8, device='cuda')
        nb = torch.empty(nn, dtype=torch.int32, device='cuda')

        def packet():
            L.opusb200_dred_compute_latents(de._h, dptr(x48), 960, 0, 1)
            L.opusb200_dred_encode_silk_frame(de._h, dptr(pay), 256, 10, 200, 6, 3, 15, dptr(act), dptr(nb), 1)
        ms = time_gpu(stream, packet, a.reps)
        front = F.RefDredFront(RF, 48000, 2)
        xin = F.dred_input(3, 48000, 2, 960)
        b200 = np.zeros(1000, np.uint8)
        a1 = np.ones(416, np.uint8)

        def ref_packet():
            front.compute_latents(xin, 960)
            RF.refh_dredenc_encode_silk_frame(F.vp(front.st), F.vp(b200), 10, 200, 6, 3, 15, F.vp(a1))
        ref_s = time_cpu(ref_packet)
        print(json.dumps({'op': 'dred_compute_latents + dred_encode_silk_frame (20 ms, 48 kHz stereo in)', 'n': nn, 'unit': 'stream-packets', 'ms': round(ms, 4),
                          'units_per_s': nn / (ms * 1e-3), 'cpu_ref_units_per_s_1core': 1.0 / ref_s, 'speedup_vs_1core': nn / (ms * 1e-3) * ref_s}), flush=True)


if __name__ == '__main__':
    main()
