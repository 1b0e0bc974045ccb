"""The device source of the LPCNet feature extraction (opus_b200/csrc/lpcnet_enc.cuh) compiled for the host-side warp emulator and compared with the
oracle (which is pinned bit for bit on the reference's own lpcnet_enc.c / freq.c / pitchdnn.c / burg.c).  Same compiler, same libm: bit for bit.
The pitch network between the two warp kernels is run with the oracle's layer functions here; on the GPU it runs on the nnet kernels
(tests/test_gpu_lpcnet_enc.py)."""
import ctypes

import numpy as np
import pytest

import emu_util
import fargan_util as fu
import util

L = fu.load_ref()
pytestmark = pytest.mark.skipif(L is None, reason='oracle/_ref/libfargan_ref.so not built (needs /root/reference): it is the weight source')


def _f(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize('is16', [0, 1])
def test_frame_features_emulated(is16):
    E = emu_util.load_emu_ops()
    O = util.load_oracle()
    E.emu_lpcnet_enc_state_size.restype = ctypes.c_int
    n, nfr = 6, 14
    ref = fu.RefLpcnetEnc(L)
    orcs = [fu.OrcLpcnetEnc(O, ref) for _ in range(n)]
    nets = [fu.OrcLpcnetEnc(O, ref) for _ in range(n)]   # only their pitch-network half is used: the layers between the two emulated kernels
    sig = np.stack([fu.speech_like(40 + s, nfr, [8000.0, 30000.0, 50.0][s % 3]) for s in range(n)])
    if is16:
        sig = np.round(sig).astype(np.int16)
    state = np.zeros(n * E.emu_lpcnet_enc_state_size(), np.uint8)
    for t in range(nfr):
        pcm = np.ascontiguousarray(sig[:, t * 160:(t + 1) * 160])
        iff, cin, out, feat = np.zeros((n, 88), np.float32), np.zeros((n, 226), np.float32), np.zeros((n, 192), np.float32), np.zeros((n, 36), np.float32)
        E.emu_lpcnet_enc_analysis(_f(state), _f(pcm), is16, _f(iff), _f(cin), n)
        assert np.all(cin[:, 0] == 0) and np.all(cin[:, 225] == 0)
        for s in range(n):
            xc = np.ascontiguousarray(cin[s, 1:225])
            O.orc_pitchdnn_net(fu.vp(nets[s].st), _f(out[s]), _f(iff[s]), _f(xc))
        E.emu_lpcnet_enc_finish(_f(state), _f(out), _f(feat), n)
        for s in range(n):
            want = orcs[s].features(pcm[s].astype(np.float32))
            w_if, w_xc, w_lpc, w_misc = orcs[s].get()
            assert np.array_equal(iff[s], w_if), (t, s)
            assert np.array_equal(cin[s, 1:225], w_xc), (t, s)
            assert np.array_equal(feat[s], want), (t, s, np.abs(feat[s] - want).max())


def test_burg_cepstrum_emulated():
    E = emu_util.load_emu_ops()
    O = util.load_oracle()
    n = 9
    x = np.stack([fu.speech_like(70 + s, 1, [8000.0, 30000.0, 3.0][s % 3]) for s in range(n)])
    x[8] = 0
    got = np.zeros((n, 36), np.float32)
    E.emu_lpcnet_burg(_f(got), _f(x), n)
    for s in range(n):
        want = np.zeros(36, np.float32)
        O.orc_burg_cepstral_analysis(_f(want), _f(x[s]))
        assert np.array_equal(got[s], want), (s, np.abs(got[s] - want).max())


@pytest.mark.parametrize('Fs,channels', [(48000, 2), (48000, 1), (24000, 1), (16000, 2), (12000, 1), (8000, 2)])
def test_dred_convert_16k_emulated(Fs, channels):
    """dred_convert_to_16k on the emulator against the oracle's front end (pinned on dnn/dred_encoder.c): the 16 kHz input buffer and the filter state
    after every 20 ms frame, bit for bit"""
    E = emu_util.load_emu_ops()
    O = util.load_oracle()
    n, nfr, fs = 5, 6, Fs // 50
    fronts = [fu.OrcDredFront(O, L, Fs, channels) for _ in range(n)]
    x = np.stack([fu.dred_input(90 + s, Fs, channels, fs * nfr) for s in range(n)])
    x[0, :3] = [1.7, -1.7, 0.999999]
    inbuf, mem = np.zeros((n, 640), np.float32), np.zeros((n, 9), np.float32)
    fill = 11
    for t in range(nfr):
        frame = np.ascontiguousarray(x[:, t * fs * channels:(t + 1) * fs * channels])
        E.emu_dred_convert_16k(_f(frame), _f(inbuf), _f(mem), n, Fs, channels, fs, 320, fill, 0, fs * channels)
        fill += 320
        got = inbuf[:, :fill].copy()
        inbuf[:, :fill - 320] = inbuf[:, 320:fill]   # what dred_compute_latents does after the double frame was processed
        fill -= 320
        for s in range(n):
            # the oracle consumed the same frame: compare its buffer BEFORE the shift through a replay of the conversion only
            fronts[s].compute_latents(frame[s], fs)
            lat, st, ib, rm, ints = fronts[s].get()
            assert ints[0] == fill
            assert np.array_equal(ib[:fill], inbuf[s, :fill]), (t, s)
            assert np.array_equal(rm, mem[s]), (t, s)
        assert np.all(np.isfinite(got))


@pytest.mark.parametrize('seed', [1, 2])
def test_dred_encode_silk_frame_emulated(seed):
    """dred_encode_silk_frame on the emulator (thread per stream, head-indexed rings) against the oracle (pinned on dnn/dred_encoder.c): sizes, payload
    bytes and the carried last_extra_dred_offset, for streams with different activity histories"""
    E = emu_util.load_emu_ops()
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    n, nfr, fs = 6, 16, 320
    fronts = [fu.OrcDredFront(O, L, 16000, 1) for _ in range(n)]
    stats, keep = fu.dred_stats(L)
    rng = np.random.default_rng(seed)
    x = np.stack([fu.dred_input(120 + seed * 10 + s, 16000, 1, fs * nfr) for s in range(n)])
    last_extra = np.zeros(n, np.int32)
    head = 0
    lat_ring, st_ring = np.zeros((n, 104, 25), np.float32), np.zeros((n, 104, 50), np.float32)
    nonzero = 0
    for t in range(nfr):
        for s in range(n):
            fronts[s].compute_latents(x[s, t * fs:(t + 1) * fs], fs)
        lat, st, ib, rm, ints = fronts[0].get()
        head = (head + 103) % 104     # one double frame per 20 ms at 16 kHz: the device's ring moves its head, the reference moves the data
        for s in range(n):
            l_s, s_s = fronts[s].get()[:2]
            lat_ring[s] = np.roll(l_s.reshape(104, 25), head, 0)
            st_ring[s] = np.roll(s_s.reshape(104, 50), head, 0)
        act = np.stack([fu.dred_activity(seed * 1000 + t * 10 + s, ['all', 'bursts', 'late', 'none', 'bursts', 'all'][(s + t) % 6]) for s in range(n)])
        q0, dQ = int(rng.integers(0, 14)), int(rng.integers(0, 8))
        qmax = int(rng.integers(q0 + 1, 16)) if dQ else q0
        max_chunks, max_bytes = int(rng.integers(1, 27)), int(rng.choice([8, 40, 200, 600]))
        buf, nb = np.zeros((n, 600), np.uint8), np.zeros(n, np.int32)
        iv = np.array([n, 600, head, 25, 50, ints[3], ints[2], ints[1], max_chunks, max_bytes, q0, dQ, qmax], np.int32)
        E.emu_dred_encode_silk_frame(_f(lat_ring), _f(st_ring), stats, _f(act), _f(last_extra), _f(buf), _f(nb), _f(iv))
        for s in range(n):
            want = np.zeros(600, np.uint8)
            r = O.orc_dredenc_encode_silk_frame(fu.vp(fronts[s].st), stats, fu.vp(want), max_chunks, max_bytes, q0, dQ, qmax, fu.vp(act[s]))
            assert r == nb[s], (t, s, r, nb[s])
            assert np.array_equal(buf[s, :r], want[:r]), (t, s)
            assert fronts[s].get()[4][4] == last_extra[s], (t, s)
            nonzero += int(r > 0)
    assert nonzero > n * nfr // 4
