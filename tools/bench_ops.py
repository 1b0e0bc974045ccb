#!/usr/bin/env python
"""Device-resident timing of the batched kernel-level ops (the §8 rows outside the decoder), with the reference's own
function timed beside each on one host core (oracle/_ref, when present).  One JSON line per op -> stdout.

    python tools/bench_ops.py [--n 65536] [--reps 5]

Inputs are resident in HBM; times are CUDA events on the library's own stream (torch.cuda.ExternalStream)."""
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
import opus_b200  # noqa: E402
from opus_b200 import capi  # noqa: E402
import util  # noqa: E402


def dptr(t):
    return ctypes.c_void_p(t.data_ptr())


def time_gpu(stream, fn, reps, warmup=3):
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            fn()
        stream.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def time_cpu(fn, min_s=1.0):
    fn()
    n, t0 = 0, time.perf_counter()
    while True:
        fn()
        n += 1
        dt = time.perf_counter() - t0
        if dt > min_s:
            return dt / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=65536)
    ap.add_argument('--reps', type=int, default=5)
    a = ap.parse_args()
    n = a.n
    ops = opus_b200.Ops(0)
    L = ops._L
    stream = torch.cuda.ExternalStream(L.opusb200_ctx_cuda_stream(ops._h))
    R = util.load_ref('simd') if util.ref_available('simd') else (util.load_ref('c') if util.ref_available('c') else None)
    RC = util.load_ref('c') if util.ref_available('c') else None
    dev = torch.device('cuda:0')
    rng = np.random.default_rng(0)
    out = []

    def emit(op, unit_name, units, ms, bytes_per_unit, cpu_s_per_unit=None, note=''):
        rec = {'op': op, 'n': units, 'unit': unit_name, 'ms': round(ms, 4), 'units_per_s': units / (ms * 1e-3),
               'algorithmic_GBps': units * bytes_per_unit / (ms * 1e-3) / 1e9,
               'cpu_ref_units_per_s_1core': (1.0 / cpu_s_per_unit) if cpu_s_per_unit else None, 'note': note}
        if cpu_s_per_unit:
            rec['speedup_vs_1core'] = rec['units_per_s'] * cpu_s_per_unit
        out.append(rec)
        print(json.dumps(rec), flush=True)

    # ---- SILK NSQ (delayed decision, 4 states, 16 kHz, 20 ms) and plain ----
    for name, c in (('silk_nsq_del_dec_4st_wb', (16, 4, 16, 24, 15729, 4)), ('silk_nsq_del_dec_2st_wb', (16, 4, 16, 16, 0, 2)),
                    ('silk_nsq_plain_wb', (16, 4, 16, 16, 0, 1))):
        cfg = util.nsq_cfg(*c)
        base = 256
        xs, fs = [], []
        for s in range(base):
            x, fr = util.make_nsq_frames(s, cfg, 1)
            xs.append(x[0])
            fs.append(fr[0])
        x = np.stack(xs)
        fr = np.stack(fs)
        reps_n = (n + base - 1) // base
        xd = torch.from_numpy(np.tile(x, (reps_n, 1))[:n].copy()).to(dev)
        frd = torch.from_numpy(np.tile(fr, reps_n)[:n].copy().view(np.uint8)).to(dev)
        pd = torch.empty((n, x.shape[1]), dtype=torch.int8, device=dev)
        q = ops.silk_nsq(n, *c)
        fn = lambda: L.opusb200_silk_nsq_process(q._h, dptr(frd), dptr(xd), dptr(pd), 1)
        ms = time_gpu(stream, fn, a.reps)
        cpu = None
        if R is not None:
            st = np.zeros(R.refh_nsq_state_size(), np.uint8)
            R.refh_nsq_state_init.restype = None
            R.refh_silk_nsq.restype = None
            R.refh_nsq_state_init(st.ctypes.data_as(ctypes.c_void_p))
            pul = np.zeros(x.shape[1], np.int8)
            state = {'i': 0}

            def one():
                i = state['i'] % base
                state['i'] += 1
                f1 = fr[i:i + 1].copy()
                R.refh_silk_nsq(cfg.ctypes.data_as(ctypes.c_void_p), st.ctypes.data_as(ctypes.c_void_p),
                                f1.ctypes.data_as(ctypes.c_void_p), x[i].ctypes.data_as(ctypes.c_void_p),
                                pul.ctypes.data_as(ctypes.c_void_p))
            cpu = time_cpu(one)
        # algorithmic bytes / frame: x16 in + pulses out + frame params + state read/write (xq, sLTP_shp carried: 2*(1280+2560))
        emit(name, 'frames(20ms)', n, ms, 320 * 2 + 320 + 400 + 2 * (1280 + 2560), cpu, 'cfg=%s' % (c,))
        q.close()

    # ---- dnn: FARGAN-sized int8 GRU step (272 -> 160; dnn/fargan.c sig_net gru1) on the tensor cores, and its pieces ----
    from opus_b200 import nnet as nnet_mod
    M, N = 272, 160
    Wi, Wr = util.Layer(1, M, 3 * N, 'i8'), util.Layer(2, N, 3 * N, 'i8', diag=True)

    def upl(Lh):
        return nnet_mod.LinearLayer(ops, Lh.nb_in, Lh.nb_out, bias=Lh.bias, subias=Lh.subias, weights=Lh.weights, float_weights=Lh.float_weights,
                                    weights_idx=Lh.idx, diag=Lh.diag, scale=Lh.scale, su_bias=True)
    dWi, dWr = upl(Wi), upl(Wr)
    xs = torch.from_numpy(np.clip(rng.standard_normal((n, M)) * 0.4, -1, 1).astype(np.float32)).to(dev)
    sts = torch.zeros((n, N), dtype=torch.float32, device=dev)
    outl = torch.empty((n, 3 * N), dtype=torch.float32, device=dev)
    fn = lambda: L.opusb200_nnet_compute_linear(ops._h, dWi._h, dptr(outl), dptr(xs), n, 1)
    ms = time_gpu(stream, fn, a.reps)
    cpu = None
    if RC is not None and hasattr(RC, 'refh_nnet_linear'):
        RC.refh_nnet_linear.restype = None
        x1 = np.clip(rng.standard_normal(M) * 0.4, -1, 1).astype(np.float32)
        o1 = np.zeros(3 * N, np.float32)
        vp = lambda arr: arr.ctypes.data_as(ctypes.c_void_p) if arr is not None else None
        cpu = time_cpu(lambda: RC.refh_nnet_linear(vp(o1), vp(x1), vp(Wi.bias), vp(Wi.subias), vp(Wi.weights), None, None, None, vp(Wi.scale), M, 3 * N), 0.5)
    rec_l = {'int8_TOPS': 2.0 * n * M * 3 * N / (ms * 1e-3) / 1e12}
    emit('nnet_linear_i8_272x480', 'stream-layers', n, ms, M * 4 + 3 * N * 4, cpu, 'tcgen05.mma kind::i8, TMEM accumulators (quantiser + GEMM kernels); %.2f TOPS' % rec_l['int8_TOPS'])
    fn = lambda: L.opusb200_nnet_compute_gru(ops._h, dWi._h, dWr._h, dptr(sts), dptr(xs), n, 1)
    ms = time_gpu(stream, fn, a.reps)
    emit('nnet_gru_272_160', 'stream-steps', n, ms, M * 4 + 2 * N * 4, None, 'two int8 linears + gate kernel; %.2f TOPS' % (2.0 * n * (M + N) * 3 * N / (ms * 1e-3) / 1e12))
    dWi.close(); dWr.close()

    # ---- FARGAN vocoder (BASELINE config 5: 32768 streams, 10 ms frames): one fargan_synthesize step for the whole batch ----
    import fargan_util as FG
    FR = FG.load_ref()
    if FR is not None:
        from opus_b200 import nnet as nn_mod
        nf = min(n, 32768)
        ref = FG.RefFargan(FR)
        layers, _ = ref.layers()
        fg = nn_mod.Fargan(ops, layers, nf)
        pcm0, feat0, feats = FG.make_inputs(5, 4)
        fg.cont(np.tile(pcm0, (nf, 1)), np.tile(feat0, (nf, 1)))
        d_feat = torch.from_numpy(np.tile(feats[0], (nf, 1)).astype(np.float32)).to(dev)
        d_out = torch.empty((nf, 160), dtype=torch.float32, device=dev)
        l0 = int(L.opusb200_ctx_launch_count(ops._h)) if hasattr(L, 'opusb200_ctx_launch_count') else 0
        fn = lambda: L.opusb200_fargan_synthesize(fg._h, dptr(d_out), dptr(d_feat), 1)
        ms = time_gpu(stream, fn, a.reps)
        ref.cont(pcm0, feat0)
        cpu = time_cpu(lambda: ref.synth(feats[0]), 1.0)
        macs = 4 * (328 * 192 + 192 * 192 + 192 * 4 + 272 * 480 + 160 * 480 + 160 * 160 + 240 * 384 + 128 * 384 + 128 * 128 + 208 * 384 + 128 * 384 + 128 * 128
                    + 688 * 128 + 128 * 128 + 128 * 40 + 80) + 32 * 64 + 192 * 128 + 128 * 320
        emit('fargan_synthesize_10ms', 'frames(10ms)', nf, ms, 20 * 4 + 160 * 4, cpu,
             'layer-by-layer graph, int8 layers on tcgen05; %.1f TOPS; CPU column = reference fargan.c, generic C branch, 1 core' % (2.0 * nf * macs / (ms * 1e-3) / 1e12))
        fg.close()

    # ---- DRED decoder network: DRED_rdovae_decode_all, 25 latents (= 1 s of redundancy: 100 feature frames) per stream ----
    if FR is not None and hasattr(FR, 'refh_dred_decode_all'):
        nd, nbl = min(n, 32768), 25
        rd = FG.RefDred(FR)
        dl, _ = rd.layers()
        dd = nn_mod.DredDecoder(ops, dl, nd)
        st0, lat0 = FG.make_dred_inputs(3, nbl, rd.latent, rd.state)
        d_st = torch.from_numpy(np.tile(st0, (nd, 1))).to(dev)
        d_lat = torch.from_numpy(np.tile(lat0[None], (nd, 1, 1))).to(dev)
        d_ft = torch.empty((nd, nbl, 80), dtype=torch.float32, device=dev)
        fn = lambda: L.opusb200_dred_decode_all(dd._h, dptr(d_ft), dptr(d_st), dptr(d_lat), nbl, 1)
        ms = time_gpu(stream, fn, max(2, a.reps // 2))
        cpu = time_cpu(lambda: rd.decode_all(st0, lat0), 1.0)
        emit('dred_decode_all_25_latents', 'streams x 1 s of DRED', nd, ms, 4 * (rd.state + nbl * (rd.latent + 1) + nbl * 80), cpu,
             'layer-by-layer graph over block-sparse int8 layers (dp4a) + small dense ones (tcgen05); CPU column = reference dred_rdovae_dec.c, generic C branch, 1 core')
        dd.close()

    # ---- SILK Burg (16 kHz: 4 x (80+16) samples, order 16) ----
    subfr, nb, D = 96, 4, 16
    t = np.arange(subfr * nb)
    base = 512
    xb = (3000 * np.sin(rng.uniform(0.01, 0.5, (base, 1)) * t) + 300 * rng.standard_normal((base, t.size))).astype(np.float32)
    xd = torch.from_numpy(np.tile(xb, ((n + base - 1) // base, 1))[:n].copy()).to(dev)
    Ad = torch.empty((n, D), dtype=torch.float32, device=dev)
    nd = torch.empty(n, dtype=torch.float32, device=dev)
    fn = lambda: L.opusb200_silk_burg_modified(ops._h, dptr(Ad), dptr(nd), dptr(xd), ctypes.c_float(1e-4), n, subfr, nb, D, 1)
    ms = time_gpu(stream, fn, a.reps)
    cpu = None
    if R is not None:
        R.silk_burg_modified_FLP.restype = ctypes.c_float
        R.silk_burg_modified_FLP.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int]
        Aw = np.zeros(D, np.float32)
        arch = R.opus_select_arch() if hasattr(R, 'opus_select_arch') else 0
        cpu = time_cpu(lambda: R.silk_burg_modified_FLP(Aw.ctypes.data_as(ctypes.c_void_p), xb[0].ctypes.data_as(ctypes.c_void_p),
                                                        1e-4, subfr, nb, D, arch))
    emit('silk_burg_modified', 'calls(384 smp, order 16)', n, ms, 384 * 4 + D * 4 + 4, cpu)

    # ---- SILK LPC analysis filter (320 samples, order 16) ----
    ln, d = 320, 16
    xi = rng.integers(-8000, 8000, size=(n, ln)).astype(np.int16)
    Bq = rng.integers(-3000, 3000, size=(n, d)).astype(np.int16)
    xd = torch.from_numpy(xi).to(dev)
    Bd = torch.from_numpy(Bq).to(dev)
    od = torch.empty((n, ln), dtype=torch.int16, device=dev)
    fn = lambda: L.opusb200_silk_lpc_analysis_filter(ops._h, dptr(od), dptr(xd), dptr(Bd), n, ln, d, 1)
    ms = time_gpu(stream, fn, a.reps)
    cpu = None
    if R is not None:
        o1 = np.zeros(ln, np.int16)
        R.silk_LPC_analysis_filter.restype = None
        cpu = time_cpu(lambda: R.silk_LPC_analysis_filter(o1.ctypes.data_as(ctypes.c_void_p), xi[0].ctypes.data_as(ctypes.c_void_p),
                                                          Bq[0].ctypes.data_as(ctypes.c_void_p), ln, d, 0), 0.5)
    emit('silk_lpc_analysis_filter', 'calls(320 smp, order 16)', n, ms, ln * 4 + d * 2, cpu)

    # ---- SILK resampler 48 -> 16 kHz, 20 ms ----
    xi = rng.integers(-8000, 8000, size=(n, 960)).astype(np.int16)
    xd = torch.from_numpy(xi).to(dev)
    od = torch.empty((n, 320), dtype=torch.int16, device=dev)
    rs = ops.silk_resampler(n, 48000, 16000)
    fn = lambda: L.opusb200_silk_resampler_process(rs._h, dptr(xd), dptr(od), 960, 1)
    ms = time_gpu(stream, fn, a.reps)
    cpu = None
    if R is not None:
        S = np.zeros(R.refh_resampler_size(), np.uint8)
        R.silk_resampler_init(S.ctypes.data_as(ctypes.c_void_p), 48000, 16000, 1)
        o1 = np.zeros(320, np.int16)
        cpu = time_cpu(lambda: R.silk_resampler(S.ctypes.data_as(ctypes.c_void_p), o1.ctypes.data_as(ctypes.c_void_p),
                                                xi[0].ctypes.data_as(ctypes.c_void_p), 960), 0.5)
    emit('silk_resampler_48_16', 'frames(20ms)', n, ms, 960 * 2 + 320 * 2, cpu)

    # ---- CELT encoder (complexity 0, CBR 64 kb/s = 160 bytes): analysis kernel + frame-coder kernel ----
    base = 64
    xb = np.stack([util.sig_music(s, 960, 2).astype(np.float32).reshape(-1) / 32768.0 for s in range(base)])
    pcm = torch.from_numpy(np.tile(xb, ((n + base - 1) // base, 1))[:n].copy()).to(dev)
    pk = torch.empty((n, 160), dtype=torch.uint8, device=dev)
    nbd = torch.empty(n, dtype=torch.int32, device=dev)
    for cx in (0, 4, 5, 10):
        enc = opus_b200.EncoderBatch(n, channels=2, complexity=cx)
        estream = torch.cuda.ExternalStream(enc.cuda_stream)
        fn = lambda: L.opusb200_encode_device(enc._h, dptr(pcm), 960, dptr(pk), 160, 160, dptr(nbd))
        ms = time_gpu(estream, fn, a.reps)
        cpu = None
        if R is not None:
            rs = np.zeros(R.refh_celt_encoder_size(2) + 64, np.uint8)
            vp = lambda arr: arr.ctypes.data_as(ctypes.c_void_p)
            R.refh_celt_encoder_init(vp(rs), 2, cx, -1000, 0, 24)
            o1 = np.zeros(1275, np.uint8)
            rr = ctypes.c_uint32(0)
            cpu = time_cpu(lambda: R.refh_celt_encode(vp(rs), vp(xb[0]), 960, vp(o1), 160, ctypes.byref(rr)), 1.0)
        emit('celt_encode_c%d_cbr64k_stereo_20ms' % cx, 'frames', n, ms, 1920 * 4 + 160, cpu, 'whole encoder: ob_celt_enc_analysis + ob_celt_enc_frame')
        enc.close()

    # ---- CELT encoder front end, 48 kHz stereo 20 ms: pre-emphasis -> MDCT -> band energies -> normalised bands (fused) ----
    base = 64
    xb = np.stack([util.sig_music(s, 960, 2).astype(np.float32).reshape(-1) / 32768.0 for s in range(base)])
    pcm = torch.from_numpy(np.tile(xb, ((n + base - 1) // base, 1))[:n].copy()).to(dev)
    Xd = torch.empty((n, 1920), dtype=torch.float32, device=dev)
    Ed = torch.empty((n, 42), dtype=torch.float32, device=dev)
    Ld = torch.empty((n, 42), dtype=torch.float32, device=dev)
    an = ops.celt_analysis(n, 2)
    fn = lambda: L.opusb200_celt_analysis_process(an._h, dptr(pcm), None, dptr(Xd), dptr(Ed), dptr(Ld), 960, 1)
    ms = time_gpu(stream, fn, a.reps)
    cpu = None
    if R is not None:
        mem, hist = np.zeros(2, np.float32), np.zeros(240, np.float32)
        X1, E1, L1 = np.zeros(1920, np.float32), np.zeros(42, np.float32), np.zeros(42, np.float32)
        R.refh_celt_enc_analysis.restype = None
        vp = lambda arr: arr.ctypes.data_as(ctypes.c_void_p)
        cpu = time_cpu(lambda: R.refh_celt_enc_analysis(vp(mem), vp(hist), vp(xb[0]), 2, 3, 0, vp(X1), vp(E1), vp(L1)), 0.5)
    emit('celt_enc_analysis_stereo_20ms', 'frames', n, ms, 1920 * 4 + 1920 * 4 + 84 * 4 + 2 * 488, cpu, 'algorithmic bytes: pcm in + X out + energies + state r/w')
    an.close()

    # ---- CELT forward MDCT, LM=3 long block ----
    xi = rng.standard_normal((n, 1080)).astype(np.float32)
    xd = torch.from_numpy(xi).to(dev)
    od = torch.empty((n, 960), dtype=torch.float32, device=dev)
    fn = lambda: L.opusb200_mdct_forward(ops._h, dptr(xd), dptr(od), n, 3, 0, 1)
    ms = time_gpu(stream, fn, a.reps)
    cpu = None
    if R is not None:
        o1 = np.zeros(960, np.float32)
        R.refh_mdct_forward.restype = None
        cpu = time_cpu(lambda: R.refh_mdct_forward(xi[0].ctypes.data_as(ctypes.c_void_p), o1.ctypes.data_as(ctypes.c_void_p), 3, 0), 0.5)
    emit('celt_mdct_forward_960', 'transforms', n, ms, 1080 * 4 + 960 * 4, cpu)

    # ---- CELT pitch xcorr (len 840, max_pitch 332: the 48 kHz/20 ms encoder pitch search after /2 decimation uses 240/~334) ----
    ln, mp = 240, 332
    xi = rng.standard_normal((n, ln)).astype(np.float32)
    yi = rng.standard_normal((n, ln + mp)).astype(np.float32)
    xd, yd = torch.from_numpy(xi).to(dev), torch.from_numpy(yi).to(dev)
    od = torch.empty((n, mp), dtype=torch.float32, device=dev)
    fn = lambda: L.opusb200_pitch_xcorr(ops._h, dptr(xd), dptr(yd), dptr(od), n, ln, mp, 1)
    ms = time_gpu(stream, fn, a.reps)
    cpu = None
    if R is not None:
        o1 = np.zeros(mp, np.float32)
        R.refh_pitch_xcorr.restype = None
        cpu = time_cpu(lambda: R.refh_pitch_xcorr(xi[0].ctypes.data_as(ctypes.c_void_p), yi[0].ctypes.data_as(ctypes.c_void_p),
                                                  o1.ctypes.data_as(ctypes.c_void_p), ln, mp), 0.5)
    emit('celt_pitch_xcorr_240x332', 'calls', n, ms, (ln + ln + mp + mp) * 4, cpu)

    # ---- CELT PVQ search N=16 K=8 ----
    N, K = 16, 8
    xi = rng.standard_normal((n * 8, N)).astype(np.float32)
    xd = torch.from_numpy(xi).to(dev)
    iy = torch.empty((n * 8, N), dtype=torch.int32, device=dev)
    yy = torch.empty(n * 8, dtype=torch.float32, device=dev)
    Kd = torch.full((n * 8,), K, dtype=torch.int32, device=dev)
    fn = lambda: L.opusb200_pvq_search(ops._h, dptr(xd), dptr(iy), dptr(yy), dptr(Kd), n * 8, N, 1)
    ms = time_gpu(stream, fn, a.reps)
    cpu = None
    if R is not None:
        x1 = xi[0].copy()
        i1 = np.zeros(N, np.int32)
        R.refh_pvq_search.restype = ctypes.c_float
        def one():
            x1[:] = xi[0]
            R.refh_pvq_search(x1.ctypes.data_as(ctypes.c_void_p), i1.ctypes.data_as(ctypes.c_void_p), K, N)
        cpu = time_cpu(one, 0.5)
    emit('celt_pvq_search_N16_K8', 'vectors', n * 8, ms, N * 8 + 8, cpu)

    # ---- round 2: the decoder-side resampler (16 -> 48 kHz: 2x all-pass + FIR interpolation), 20 ms ----
    rs = ops.silk_resampler(n, 16000, 48000, False)
    xi = torch.from_numpy(rng.integers(-30000, 30000, size=(n, 320)).astype(np.int16)).to(dev)
    yo = torch.empty((n, 960), dtype=torch.int16, device=dev)
    fn = lambda: L.opusb200_silk_resampler_process(rs._h, dptr(xi), dptr(yo), 320, 1)
    ms = time_gpu(stream, fn, a.reps)
    cpu = None
    if RC is not None:
        st_ = ctypes.create_string_buffer(RC.refh_resampler_size())
        RC.silk_resampler_init(st_, 16000, 48000, 0)
        x1 = xi[0].cpu().numpy().copy()
        y1 = np.zeros(960, np.int16)
        cpu = time_cpu(lambda: RC.silk_resampler(st_, y1.ctypes.data_as(ctypes.c_void_p), x1.ctypes.data_as(ctypes.c_void_p), 320), 0.5)
    emit('silk_resampler_16k_to_48k_20ms', 'frames', n, ms, 320 * 2 + 960 * 2 + 2 * 184, cpu)
    rs.close()

    # ---- round 2: the bitstream of a SILK frame (silk_encode_indices + silk_encode_pulses, WB 20 ms) on the device range-encoder object ----
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from test_silk_coder_oracle_vs_ref import SilkIndices, random_indices, make_pulses
    base = 512
    idx = (SilkIndices * base)()
    r2 = np.random.default_rng(5)
    for s_ in range(base):
        idx[s_] = random_indices(r2, 16, 4, 0)
    P = make_pulses(r2, base, 320, 1)
    reps_n = (n + base - 1) // base
    idx_np = np.frombuffer(idx, dtype=np.uint8).reshape(base, ctypes.sizeof(SilkIndices))
    d_idx = torch.from_numpy(np.tile(idx_np, (reps_n, 1))[:n].copy()).to(dev)
    d_P = torch.from_numpy(np.tile(P, (reps_n, 1))[:n].copy()).to(dev)
    d_st = torch.from_numpy(np.tile(np.array([idx[s_].signalType for s_ in range(base)], np.int32), reps_n)[:n].copy()).to(dev)
    d_qo = torch.from_numpy(np.tile(np.array([idx[s_].quantOffsetType for s_ in range(base)], np.int32), reps_n)[:n].copy()).to(dev)
    d_cc = torch.zeros(n, dtype=torch.int32, device=dev)
    re = ops.range_encoder(n, 608)

    def silk_frame():
        L.opusb200_range_encoder_init(re._h, 600)
        L.opusb200_silk_encode_indices(re._h, 16, 4, dptr(d_idx), dptr(d_cc), 1)
        L.opusb200_silk_encode_pulses(re._h, dptr(d_st), dptr(d_qo), dptr(d_P), P.shape[1], 320, 1)
    ms = time_gpu(stream, silk_frame, a.reps)
    cpu = None
    if RC is not None:
        rc = ctypes.create_string_buffer(RC.refh_silk_coder_size())
        prev = (ctypes.c_int * 2)(0, 0)
        q1 = P[0].copy()

        def one_frame():
            RC.refh_silk_coder_init(rc, 600)
            RC.refh_silk_encode_indices(rc, 16, 4, ctypes.byref(idx[0]), 0, prev)
            RC.refh_silk_encode_pulses(rc, int(idx[0].signalType), int(idx[0].quantOffsetType), q1.ctypes.data_as(ctypes.c_void_p), 320)
        cpu = time_cpu(one_frame, 0.5)
    emit('silk_frame_bitstream_wb_20ms', 'frames', n, ms, 36 + 320 + 100, cpu, 'indices + pulses through ec_enc; algorithmic bytes: indices + pulses in, ~100 B out')
    re.close()

    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'bench_ops.json'), 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
