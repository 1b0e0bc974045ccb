"""Pins oracle/dred_oracle.c against the reference's own dnn/dred_rdovae_dec.c (DRED_rdovae_decode_all) compiled with a seeded random,
block-sparsified model (oracle/gen_dred_weights.py).  Same build flags as the FARGAN pin (generic C branch of dnn/vec.h).
Bar: bit-identical feature frames."""
import numpy as np
import pytest

import fargan_util as F
import util

R = F.load_ref()
pytestmark = pytest.mark.skipif(R is None or not hasattr(R, 'refh_dred_decode_all'), reason='oracle/_ref/libfargan_ref.so without the DRED decoder')


@pytest.mark.parametrize('seed,nb', [(1, 1), (2, 7), (3, 26), (4, 50)])
def test_dred_decode_all_matches_reference(seed, nb):
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    ref = F.RefDred(R)
    layers, shapes = ref.layers()
    assert shapes[0][:2] == (ref.latent + 1, 96) and shapes[7][:2] == (576, 80) and shapes[13][:2] == (128, 320)
    assert shapes[14][2] and shapes[8][2] and shapes[7][2]          # input-sparse GRU matrices, sparse conv_dense / output layers
    state, latents = F.make_dred_inputs(seed, nb, ref.latent, ref.state)
    want = ref.decode_all(state, latents)
    got = np.zeros_like(want)
    O.orc_dred_decode_all(layers, F.vp(got), F.vp(state), F.vp(latents), nb, ref.latent)
    assert np.array_equal(got, want), float(np.abs(got - want).max())
    assert float(np.abs(want).max()) > 1e-3


@pytest.mark.skipif(R is None or not hasattr(R, 'refh_dred_encode_dframe'), reason='oracle/_ref/libfargan_ref.so without the DRED encoder')
@pytest.mark.parametrize('seed', [1, 2, 3])
def test_dred_encode_dframe_matches_reference(seed):
    """dred_rdovae_encode_dframe (dnn/dred_rdovae_enc.c:47): 30 consecutive double feature frames with the carried GRU / conv state: latents and
    initial state bit-identical"""
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    ref = F.RefDredEnc(R)
    layers, shapes = ref.layers()
    assert shapes[0][:2] == (40, 64) and shapes[1][:2] == (544, 32) and shapes[3][:2] == (128, 56) and shapes[9][:2] == (64, 96) and shapes[17][:2] == (448, 96)
    assert shapes[9][2] and shapes[4][2]          # input-sparse GRU matrices, sparse conv_dense layers
    st = np.zeros(O.orc_dred_enc_size() + 64, np.uint8)
    O.orc_dred_enc_init(F.vp(st), layers)
    rng = np.random.default_rng(seed)
    for t in range(30):
        x = (0.6 * rng.standard_normal(40)).astype(np.float32)
        wl, ws = ref.encode_dframe(x)
        gl, gs = np.zeros(25, np.float32), np.zeros(50, np.float32)
        O.orc_dred_encode_dframe(F.vp(st), F.vp(gl), F.vp(gs), F.vp(x))
        assert np.array_equal(gl, wl) and np.array_equal(gs, ws), (t, float(np.abs(gl - wl).max()), float(np.abs(gs - ws).max()))
    assert float(np.abs(wl).max()) > 1e-3


@pytest.mark.skipif(R is None or not hasattr(R, 'refh_dredenc_size'), reason='oracle/_ref/libfargan_ref.so without dnn/dred_encoder.c')
@pytest.mark.parametrize('Fs,channels,frame_ms,extra', [(48000, 2, 20, 0), (48000, 1, 10, 0), (16000, 1, 20, 0), (24000, 2, 20, 120), (12000, 1, 40, 0), (8000, 1, 20, 0),
                                                       (48000, 1, 2.5, 0)])
def test_dred_compute_latents_matches_reference(Fs, channels, frame_ms, extra):
    """dred_compute_latents over a run of frames (16 kHz conversion + feature extraction + RDOVAE encoder + ring buffers): everything bit for bit"""
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    ref = F.RefDredFront(R, Fs, channels)
    orc = F.OrcDredFront(O, R, Fs, channels)
    fs = int(Fs * frame_ms / 1000)
    nfr = 14 if frame_ms >= 10 else 60
    x = F.dred_input(3 + channels + Fs // 1000, Fs, channels, fs * nfr)
    x[:4] = [1.5, -1.5, 0.99999, -1.0][:4]   # the int16 conversion saturates
    seen = 0
    for t in range(nfr):
        frame = x[t * fs * channels:(t + 1) * fs * channels]
        ref.compute_latents(frame, fs, extra)
        orc.compute_latents(frame, fs, extra)
        a, b = ref.get(), orc.get()
        for u, v in zip(a, b):
            assert np.array_equal(u, v), t
        seen = max(seen, int(a[4][3]))
    assert seen >= 2 and np.abs(a[0][:25]).max() > 1e-3
    orc.reset()
    R.refh_dredenc_reset(F.vp(ref.st))
    ref.compute_latents(x[:fs * channels], fs, extra)
    orc.compute_latents(x[:fs * channels], fs, extra)
    for u, v in zip(ref.get(), orc.get()):
        assert np.array_equal(u, v)


@pytest.mark.skipif(R is None or not hasattr(R, 'refh_dred_stats'), reason='oracle/_ref/libfargan_ref.so without dnn/dred_encoder.c')
@pytest.mark.parametrize('seed', [1, 2, 3])
def test_dred_encode_silk_frame_matches_reference(seed):
    """dred_encode_silk_frame after every 20 ms frame of a run: returned size, payload bytes and the carried last_extra_dred_offset, bit for bit,
    over activity patterns (all active / bursts / silence now / none), quantiser settings and byte budgets (incl. too small ones -> 0)"""
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    Fs, ch, fs = 16000, 1, 320
    ref = F.RefDredFront(R, Fs, ch)
    orc = F.OrcDredFront(O, R, Fs, ch)
    stats, _ = F.dred_stats(R)
    rng = np.random.default_rng(seed)
    nfr = 30
    x = F.dred_input(60 + seed, Fs, ch, fs * nfr)
    sizes = []
    for t in range(nfr):
        frame = x[t * fs:(t + 1) * fs]
        ref.compute_latents(frame, fs)
        orc.compute_latents(frame, fs)
        kind = ['all', 'bursts', 'late', 'none'][int(rng.integers(0, 4))] if t > 2 else 'all'
        act = F.dred_activity(seed * 100 + t, kind)
        q0 = int(rng.integers(0, 14))
        dQ = int(rng.integers(0, 8))
        qmax = int(rng.integers(q0 + 1, 16)) if dQ else q0
        max_chunks = int(rng.integers(1, 27))
        max_bytes = int(rng.choice([6, 20, 60, 200, 1000]))
        a, b = np.zeros(1000, np.uint8), np.zeros(1000, np.uint8)
        ra = R.refh_dredenc_encode_silk_frame(F.vp(ref.st), F.vp(a), max_chunks, max_bytes, q0, dQ, qmax, F.vp(act.copy()))
        rb = O.orc_dredenc_encode_silk_frame(F.vp(orc.st), stats, F.vp(b), max_chunks, max_bytes, q0, dQ, qmax, F.vp(act))
        assert ra == rb, (t, ra, rb)
        assert np.array_equal(a[:ra], b[:rb]), t
        assert ref.get()[4][4] == orc.get()[4][4], t
        sizes.append(ra)
    assert max(sizes) > 30 and min(sizes) == 0
