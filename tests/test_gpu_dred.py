"""Parity of the batched CUDA DRED decoder network (opusb200_dred_decode_all) with the oracle restatement of dnn/dred_rdovae_dec.c,
itself bit-identical to the reference's own file (tests/test_dred_oracle_vs_ref.py).  Weights: the seeded random block-sparse
model inside oracle/_ref/libfargan_ref.so.  Bar: bit-identical feature frames."""
import numpy as np
import pytest

import fargan_util as F
import util

pytestmark = pytest.mark.gpu
R = F.load_ref()


@pytest.mark.skipif(R is None or not hasattr(R, 'refh_dred_decode_all'), reason='reference DRED build missing')
@pytest.mark.parametrize('n,nb', [(2, 3), (131, 12)])
def test_dred_decode_all_matches_oracle(n, nb):
    import opus_b200
    from opus_b200 import nnet
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    ref = F.RefDred(R)
    layers, shapes = ref.layers()
    ins = [F.make_dred_inputs(40 + s, nb, ref.latent, ref.state) for s in range(min(n, 5))]
    ops = opus_b200.Ops()
    dd = nnet.DredDecoder(ops, layers, n)
    state = np.stack([ins[s % len(ins)][0] for s in range(n)])
    lat = np.stack([ins[s % len(ins)][1] for s in range(n)])
    for rep in range(4):     # call 1 runs eagerly, call 2 is captured into a CUDA graph, later calls replay it
        got = dd.decode_all(state, lat)
    for s in range(len(ins)):   # (checked on the last, replayed call)
        want = np.zeros((nb, 80), np.float32)
        O.orc_dred_decode_all(layers, F.vp(want), F.vp(ins[s][0]), F.vp(ins[s][1]), nb, ref.latent)
        for r in range(s, n, len(ins)):
            assert np.array_equal(got[r], want), (r, float(np.abs(got[r] - want).max()))
    dd.close()


@pytest.mark.skipif(R is None or not hasattr(R, 'refh_dred_encode_dframe'), reason='oracle/_ref/libfargan_ref.so without the DRED encoder')
@pytest.mark.parametrize('n', [4, 150])
def test_dred_encoder_matches_oracle(n):
    """opusb200_dred_encode_dframe (the DRED encoder network, dnn/dred_rdovae_enc.c) for n streams x 12 consecutive double frames with carried
    state, against the oracle (pinned bit-exact on the reference's file): int8 layers are exact, the float layers keep the reference's order;
    tanh / sigmoid go through the shared approximations -> bit-identical expected, 1e-6 allowed"""
    import opus_b200
    from opus_b200 import nnet
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    ref = F.RefDredEnc(R)
    layers, _ = ref.layers()
    ops = opus_b200.Ops()
    enc = nnet.DredEncoder(ops, layers, n)
    kinds = min(n, 5)
    sts = []
    for s in range(kinds):
        st = np.zeros(O.orc_dred_enc_size() + 64, np.uint8)
        O.orc_dred_enc_init(F.vp(st), layers)
        sts.append(st)
    rng = np.random.default_rng(n)
    exact = 0
    for t in range(12):
        xk = (0.6 * rng.standard_normal((kinds, 40))).astype(np.float32)
        x = np.stack([xk[s % kinds] for s in range(n)])
        lat, stt = enc.encode_dframe(x)
        for s in range(kinds):
            wl, ws = np.zeros(25, np.float32), np.zeros(50, np.float32)
            O.orc_dred_encode_dframe(F.vp(sts[s]), F.vp(wl), F.vp(ws), F.vp(np.ascontiguousarray(xk[s])))
            for r in range(s, n, kinds):
                assert np.abs(lat[r] - wl).max() <= 1e-6 and np.abs(stt[r] - ws).max() <= 1e-6, (t, r)
            exact += int(np.array_equal(lat[s], wl) and np.array_equal(stt[s], ws))
    print('bit-identical frames: %d of %d' % (exact, 12 * kinds))
    assert exact >= 12 * kinds - 2
    enc.close()


@pytest.mark.skipif(R is None or not hasattr(R, 'refh_dredenc_size'), reason='oracle/_ref/libfargan_ref.so without dnn/dred_encoder.c')
@pytest.mark.parametrize('Fs,channels,frame_ms,n', [(48000, 2, 20, 70), (16000, 1, 20, 9), (24000, 1, 10, 9), (8000, 2, 40, 5), (12000, 1, 20, 5)])
def test_dred_compute_latents_matches_oracle(Fs, channels, frame_ms, n):
    """dred_compute_latents for a batch (16 kHz conversion kernel, feature extraction, RDOVAE encoder, ring buffers) against the oracle's front end,
    which is pinned bit for bit on dnn/dred_encoder.c.  The conversion and the buffers are exact; features and latents carry the libm note of
    tests/test_gpu_lpcnet_enc.py (almost always bit-identical; an int8 activation step otherwise): latents within 0.05, nine in ten double frames exact."""
    import conftest
    import opus_b200
    from opus_b200 import nnet
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    layers, _ = F.RefDredEnc(R).layers()
    renc = F.RefLpcnetEnc(R)
    lin, _ = renc.linear()
    cv = renc.conv()
    ops = opus_b200.Ops()
    enc = nnet.DredEncoder(ops, layers, n)
    enc.init_frontend(Fs, channels, lin, cv)
    kinds = min(n, 5)
    fronts = [F.OrcDredFront(O, R, Fs, channels) for _ in range(kinds)]
    fs = int(Fs * frame_ms / 1000)
    nfr = 12
    x = np.stack([F.dred_input(40 + s, Fs, channels, fs * nfr) for s in range(kinds)])
    exact = total = 0
    worst = 0.0
    for t in range(nfr):
        if t == 8:
            enc.reset()
            for f in fronts:
                f.reset()
        frame = np.ascontiguousarray(x[:, t * fs * channels:(t + 1) * fs * channels])
        enc.compute_latents(frame[np.arange(n) % kinds], fs)
        lat, st, info = enc.get_latents()
        for s in range(kinds):
            fronts[s].compute_latents(frame[s], fs)
            w_lat, w_st, w_ib, w_rm, w_ints = fronts[s].get()
            assert list(info) == [w_ints[0], w_ints[1], w_ints[2], w_ints[3]], (t, list(info), list(w_ints))
            for r in range(s, n, kinds):
                assert np.array_equal(lat[r], lat[s]) and np.array_equal(st[r], st[s]), (t, r)
            d = max(float(np.abs(lat[s] - w_lat).max()), float(np.abs(st[s] - w_st).max()))
            worst = max(worst, d)
            assert d < 0.05, (t, s, d)
            total += 1
            exact += int(np.array_equal(lat[s], w_lat) and np.array_equal(st[s], w_st))
    print('bit-identical buffers: %d of %d, worst %.3g' % (exact, total, worst))
    conftest.note_err('dred_compute_latents', worst)
    assert exact >= 0.9 * total
    enc.close()


@pytest.mark.skipif(R is None or not hasattr(R, 'refh_dred_stats'), reason='oracle/_ref/libfargan_ref.so without dnn/dred_encoder.c')
@pytest.mark.parametrize('n', [6, 200])
def test_dred_encode_silk_frame_matches_oracle(n):
    """the whole DRED encoder: dred_compute_latents + dred_encode_silk_frame per 20 ms packet, every stream with its own audio and voice-activity
    history, against the oracle (pinned bit for bit on dnn/dred_encoder.c).  Bar: payload sizes and bytes identical."""
    import opus_b200
    from opus_b200 import nnet
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    layers, _ = F.RefDredEnc(R).layers()
    renc = F.RefLpcnetEnc(R)
    lin, _ = renc.linear()
    cv = renc.conv()
    stats, keep = F.dred_stats(R)
    ops = opus_b200.Ops()
    enc = nnet.DredEncoder(ops, layers, n)
    enc.init_frontend(16000, 1, lin, cv)
    enc.load_stats(stats)
    kinds, nfr, fs = min(n, 6), 18, 320
    fronts = [F.OrcDredFront(O, R, 16000, 1) for _ in range(kinds)]
    x = np.stack([F.dred_input(150 + s, 16000, 1, fs * nfr) for s in range(kinds)])
    rng = np.random.default_rng(4)
    nonzero = 0
    for t in range(nfr):
        frame = np.ascontiguousarray(x[:, t * fs:(t + 1) * fs])
        enc.compute_latents(frame[np.arange(n) % kinds], fs)
        act = np.stack([F.dred_activity(t * 10 + s, ['all', 'bursts', 'late', 'none', 'bursts', 'all'][(s + t) % 6]) for s in range(kinds)])
        q0, dQ = int(rng.integers(0, 14)), int(rng.integers(0, 8))
        qmax = int(rng.integers(q0 + 1, 16)) if dQ else q0
        max_chunks, max_bytes = int(rng.integers(1, 27)), int(rng.choice([8, 40, 200, 600]))
        buf, nb = enc.encode_silk_frame(act[np.arange(n) % kinds], max_chunks, max_bytes, q0, dQ, qmax)
        for s in range(kinds):
            fronts[s].compute_latents(frame[s], fs)
            want = np.zeros(600, np.uint8)
            r = O.orc_dredenc_encode_silk_frame(F.vp(fronts[s].st), stats, F.vp(want), max_chunks, max_bytes, q0, dQ, qmax, F.vp(act[s]))
            for q in range(s, n, kinds):
                assert nb[q] == r, (t, q, int(nb[q]), r)
                assert np.array_equal(buf[q, :r], want[:r]), (t, q)
            nonzero += int(r > 0)
    assert nonzero > kinds * nfr // 4
    enc.close()
