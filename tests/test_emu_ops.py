"""Encoder-side op kernel SOURCES on the host SIMT emulator vs the oracle restatement: bit-exact."""
import ctypes

import numpy as np
import pytest

import emu_util
import util


def _f(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize('LM,transient', [(3, 0), (3, 1), (2, 1), (1, 0), (0, 0)])
def test_emu_mdct_forward(LM, transient):
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    rng = np.random.default_rng(5)
    N = 120 << LM
    x = (rng.standard_normal((3, N + 120)) * 1000).astype(np.float32)
    got = np.zeros((3, N), np.float32)
    E.emu_mdct_forward(_f(x), _f(got), 3, LM, transient)
    for i in range(3):
        want = np.zeros(N, np.float32)
        O.orc_mdct_forward(_f(x[i]), _f(want), LM, transient)
        assert np.array_equal(got[i], want)


@pytest.mark.parametrize('n', [1, 2, 5])
def test_emu_mdct_forward_register_fft(n):
    """celt_mdct_fast.cuh (long block, FFT in registers, two transforms per warp) is bit-identical to the oracle's clt_mdct_forward"""
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    rng = np.random.default_rng(50 + n)
    x = (rng.standard_normal((n, 1080)) * 1000).astype(np.float32)
    got = np.zeros((n, 960), np.float32)
    E.emu_mdct_forward_fast(_f(x), _f(got), n)
    for i in range(n):
        want = np.zeros(960, np.float32)
        O.orc_mdct_forward(_f(x[i]), _f(want), 3, 0)
        assert np.array_equal(got[i], want), (i, int((got[i] != want).sum()))


def test_emu_pitch_xcorr():
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    rng = np.random.default_rng(6)
    n, length, mp = 3, 240, 244
    x = rng.standard_normal((n, length)).astype(np.float32)
    y = rng.standard_normal((n, length + mp)).astype(np.float32)
    got = np.zeros((n, mp), np.float32)
    E.emu_pitch_xcorr(_f(x), _f(y), _f(got), n, length, mp)
    for i in range(n):
        want = np.zeros(mp, np.float32)
        O.orc_pitch_xcorr(_f(x[i]), _f(y[i]), _f(want), length, mp)
        assert np.array_equal(got[i], want)


@pytest.mark.parametrize('N', [2, 8, 22, 48, 176])
def test_emu_rotation_and_search(N):
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    O.orc_pvq_search.restype = ctypes.c_float
    rng = np.random.default_rng(N)
    n = 40
    X = rng.standard_normal((n, N)).astype(np.float32)
    X /= np.sqrt((X.astype(np.float64) ** 2).sum(1))[:, None].astype(np.float32)
    K = rng.integers(1, 129, size=n).astype(np.int32)
    for d in (1, -1):
        for stride in (1, 2):
            if N % stride:
                continue
            got = X.copy()
            E.emu_exp_rotation(_f(got), _f(K), n, N, d, stride, 2)
            for i in range(n):
                w = X[i].copy()
                O.orc_exp_rotation(_f(w), N, d, stride, int(K[i]), 2)
                assert np.array_equal(got[i], w)
    gx, giy, gyy = X.copy(), np.zeros((n, N), np.int32), np.zeros(n, np.float32)
    E.emu_pvq_search(_f(gx), _f(giy), _f(gyy), _f(K), n, N)
    for i in range(n):
        w, wi = X[i].copy(), np.zeros(N, np.int32)
        wy = O.orc_pvq_search(_f(w), _f(wi), int(K[i]), N)
        assert np.array_equal(giy[i], wi) and gyy[i] == wy and np.array_equal(gx[i], w)


def test_emu_silk_lpc_and_resampler():
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    rng = np.random.default_rng(9)
    n, length, d = 40, 320, 16
    x = rng.integers(-20000, 20000, size=(n, length)).astype(np.int16)
    B = rng.integers(-4000, 4000, size=(n, d)).astype(np.int16)
    got = np.zeros_like(x)
    E.emu_silk_lpc(_f(got), _f(x), _f(B), n, length, d)
    want = np.zeros(length, np.int16)
    for i in range(n):
        O.orc_silk_lpc_analysis_filter(_f(want), _f(x[i]), _f(B[i]), length, d)
        assert np.array_equal(got[i], want)
    # resampler 48 -> 16 kHz, 3 frames of 20 ms, state carried
    st = np.zeros(n * E.emu_silk_resampler_state_size(), np.uint8)
    os_ = [ctypes.create_string_buffer(O.orc_resampler_size()) for _ in range(n)]
    for s in os_:
        assert O.orc_resampler_init(s, 48000, 16000) == 0
    for fr in range(3):
        xin = rng.integers(-30000, 30000, size=(n, 960)).astype(np.int16)
        got = np.zeros((n, 320), np.int16)
        E.emu_silk_resampler_48_16(_f(st), _f(xin), _f(got), n, 960)
        w = np.zeros(320, np.int16)
        for i in range(n):
            O.orc_resampler_process(os_[i], _f(w), _f(xin[i]), 960)
            assert np.array_equal(got[i], w), (fr, i)


@pytest.mark.parametrize('fs_in,fs_out,for_enc', [(48000, 12000, 1), (16000, 12000, 1), (24000, 16000, 1), (12000, 8000, 1), (16000, 16000, 1), (8000, 16000, 0),
                                                  (8000, 48000, 0), (12000, 48000, 0), (16000, 24000, 0), (12000, 16000, 0), (16000, 48000, 0), (8000, 12000, 0)])
def test_emu_silk_resampler_rate_pairs(fs_in, fs_out, for_enc):
    """the device resampler on the emulator for the four resampler functions (copy, 2x all-pass, up-sampling + FIR interpolation, down-FIR with
    1 / 2 / 3 phases), three 20 ms frames with carried state, bit-exact against the oracle"""
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    rng = np.random.default_rng(fs_in // 1000 + fs_out // 100)
    n, n_in, n_out = 33, fs_in // 50, fs_out // 50
    st = np.zeros(n * E.emu_silk_resampler_state_size(), np.uint8)
    os_ = [ctypes.create_string_buffer(O.orc_resampler_size()) for _ in range(n)]
    for s in os_:
        assert O.orc_resampler_init2(s, fs_in, fs_out, for_enc) == 0
    for fr in range(3):
        xin = rng.integers(-32768, 32768, size=(n, n_in)).astype(np.int16)
        got = np.zeros((n, n_out), np.int16)
        assert E.emu_silk_resampler(_f(st), _f(xin), _f(got), n, n_in, fs_in, fs_out, for_enc) == 0
        w = np.zeros(n_out, np.int16)
        for i in range(n):
            O.orc_resampler_process(os_[i], _f(w), _f(xin[i]), n_in)
            assert np.array_equal(got[i], w), (fr, i)


def test_emu_silk_burg():
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    O.orc_silk_burg_modified.restype = ctypes.c_float
    O.orc_silk_burg_modified.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    E.emu_silk_burg.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    rng = np.random.default_rng(3)
    n, subfr, nb, D = 40, 96, 4, 16
    t = np.arange(subfr * nb)
    x = np.stack([(3000 * np.sin(0.03 * (i + 1) * t) + (3 if i % 2 else 500) * rng.standard_normal(t.size)) for i in range(n)]).astype(np.float32)
    for mig in (1e-4, 0.5):
        A = np.zeros((n, D), np.float32)
        nrg = np.zeros(n, np.float32)
        E.emu_silk_burg(_f(A), _f(nrg), _f(x), mig, n, subfr, nb, D)
        w = np.zeros(D, np.float32)
        for i in range(n):
            r = O.orc_silk_burg_modified(_f(w), _f(x[i]), mig, subfr, nb, D)
            assert r == nrg[i] and np.array_equal(A[i], w)


@pytest.mark.parametrize('cc,lm', [(2, 3), (1, 3), (2, 2), (2, 0)])
def test_emu_enc_analysis(cc, lm):
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    N = 120 << lm
    n, nf = 5, 3
    x = np.stack([util.sig_music(50 + s, N * nf, cc).astype(np.float32).reshape(nf, N * cc) / 32768.0 for s in range(n)])
    sb = np.array([0, 1, 0, 1, 1], np.int32) if lm else np.zeros(n, np.int32)
    st_e = np.zeros(n * E.emu_enc_analysis_state_size(), np.uint8)
    st_o = [np.zeros(O.orc_enc_analysis_state_size(), np.uint8) for _ in range(n)]
    for f in range(nf):
        pcm = np.ascontiguousarray(x[:, f])
        X, En, Lg = np.zeros((n, cc * N), np.float32), np.zeros((n, cc * 21), np.float32), np.zeros((n, cc * 21), np.float32)
        E.emu_celt_enc_analysis(_f(st_e), _f(pcm), _f(sb), _f(X), _f(En), _f(Lg), n, cc, lm)
        for s in range(n):
            Xo, Eo, Lo = np.zeros(cc * N, np.float32), np.zeros(cc * 21, np.float32), np.zeros(cc * 21, np.float32)
            O.orc_celt_enc_analysis(_f(st_o[s]), _f(pcm[s]), cc, lm, int(sb[s]), _f(Xo), _f(Eo), _f(Lo))
            assert np.array_equal(En[s].view(np.uint32), Eo.view(np.uint32)), (f, s)
            assert np.array_equal(Lg[s].view(np.uint32), Lo.view(np.uint32))
            assert np.array_equal(X[s].view(np.uint32), Xo.view(np.uint32))


@pytest.mark.parametrize('frame_length,kind', [(320, 1), (320, 2), (160, 0), (120, 1), (240, 3), (80, 2)])
def test_emu_silk_encode_pulses(frame_length, kind):
    """the device excitation coder (silk_coder.cuh) on the emulator against the oracle: buffer bytes, ec_tell_frac, final range; symbols in front"""
    from test_silk_coder_oracle_vs_ref import make_pulses
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    O.orc_silk_coder_done.restype = ctypes.c_uint32
    O.orc_silk_coder_tell_frac.restype = ctypes.c_uint32
    rng = np.random.default_rng(7 * frame_length + kind)
    n, storage, stride = 40, (1275 if kind == 2 else 300), 1280
    P = make_pulses(rng, n, frame_length, kind)
    st = rng.integers(0, 3, n).astype(np.int32)
    qo = rng.integers(0, 2, n).astype(np.int32)
    icdf = np.array([200, 120, 30, 0], np.uint8)
    syms = [rng.integers(-1, 4, n).astype(np.int32) for _ in range(3)]   # -1: nothing for that stream
    state = np.zeros(n * E.emu_silk_coder_state_size(), np.uint8)
    buf = np.zeros((n, stride), np.uint8)
    tf, fr = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    E.emu_silk_coder(0, _f(state), _f(buf), n, stride, storage, None, None, None, 0, 0, None, 0, 0, None, None)
    for sy in syms:
        E.emu_silk_coder(1, _f(state), _f(buf), n, stride, storage, _f(sy), None, None, 0, 0, _f(icdf), 4, 8, None, None)
    E.emu_silk_coder(2, _f(state), _f(buf), n, stride, storage, _f(st), _f(qo), _f(P), P.shape[1], frame_length, None, 0, 0, None, None)
    E.emu_silk_coder(3, _f(state), _f(buf), n, stride, storage, None, None, None, 0, 0, None, 0, 0, _f(tf), _f(fr))
    for s in range(n):
        oc = ctypes.create_string_buffer(O.orc_silk_coder_size())
        O.orc_silk_coder_init(oc, storage)
        for sy in syms:
            if sy[s] >= 0:
                O.orc_silk_coder_icdf(oc, int(sy[s]), _f(icdf), 8)
        q = P[s].copy()
        O.orc_silk_encode_pulses(oc, int(st[s]), int(qo[s]), _f(q), frame_length)
        assert O.orc_silk_coder_tell_frac(oc) == tf[s], s
        want = np.zeros(storage, np.uint8)
        assert O.orc_silk_coder_done(oc, _f(want)) == fr[s], s
        assert np.array_equal(buf[s, :storage], want), s


@pytest.mark.parametrize('fs_kHz,nb_subfr', [(16, 4), (12, 2), (8, 4)])
def test_emu_silk_frame_bitstream(fs_kHz, nb_subfr):
    """three consecutive SILK frames per stream on the emulator: silk_encode_indices (with the carried pitch-lag state) + silk_encode_pulses in one
    range coder per frame, against the oracle"""
    from test_silk_coder_oracle_vs_ref import SilkIndices, random_indices, make_pulses
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    O.orc_silk_coder_done.restype = ctypes.c_uint32
    O.orc_silk_coder_tell_frac.restype = ctypes.c_uint32
    rng = np.random.default_rng(fs_kHz + nb_subfr)
    n, storage, stride, fl = 36, 600, 608, 5 * nb_subfr * fs_kHz
    state = np.zeros(n * E.emu_silk_coder_state_size(), np.uint8)
    prev = np.zeros((n, 2), np.int32)
    oprev = [(ctypes.c_int * 2)(0, 0) for _ in range(n)]
    for frame in range(3):
        cond = np.array([[0, 2, 2][frame] if s % 3 else [1, 0, 2][frame] for s in range(n)], np.int32)
        idx = (SilkIndices * n)()
        for s in range(n):
            idx[s] = random_indices(rng, fs_kHz, nb_subfr, int(cond[s]))
            if frame and s % 2:
                idx[s].signalType = 2
        P = make_pulses(rng, n, fl, frame)
        st = np.array([idx[s].signalType for s in range(n)], np.int32)
        qo = np.array([idx[s].quantOffsetType for s in range(n)], np.int32)
        buf = np.zeros((n, stride), np.uint8)
        tf, fr = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        E.emu_silk_coder(0, _f(state), _f(buf), n, stride, storage, None, None, None, 0, 0, None, 0, 0, None, None)
        E.emu_silk_encode_indices(_f(state), _f(buf), n, stride, storage, fs_kHz, nb_subfr, ctypes.cast(idx, ctypes.c_void_p), _f(cond), _f(prev))
        E.emu_silk_coder(2, _f(state), _f(buf), n, stride, storage, _f(st), _f(qo), _f(P), P.shape[1], fl, None, 0, 0, None, None)
        E.emu_silk_coder(3, _f(state), _f(buf), n, stride, storage, None, None, None, 0, 0, None, 0, 0, _f(tf), _f(fr))
        for s in range(n):
            oc = ctypes.create_string_buffer(O.orc_silk_coder_size())
            O.orc_silk_coder_init(oc, storage)
            O.orc_silk_encode_indices(oc, fs_kHz, nb_subfr, ctypes.byref(idx[s]), int(cond[s]), oprev[s])
            q = P[s].copy()
            O.orc_silk_encode_pulses(oc, int(st[s]), int(qo[s]), _f(q), fl)
            assert O.orc_silk_coder_tell_frac(oc) == tf[s], (frame, s)
            want = np.zeros(storage, np.uint8)
            assert O.orc_silk_coder_done(oc, _f(want)) == fr[s] and np.array_equal(buf[s, :storage], want), (frame, s)
            assert list(oprev[s]) == list(prev[s]), (frame, s)


def test_silk_process_gains_emulated():
    """silk_process_gains_FLP + silk_gains_quant: the device source on the emulator against the oracle (pinned on the reference's functions), streams with
    carried LastGainIndex over several frames"""
    from test_silk_gains_oracle_vs_ref import Gains, random_frame, fields
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    rng = np.random.default_rng(12)
    n = 70
    for nb_subfr, subfr_length, nstates in [(4, 80, 4), (2, 60, 1)]:
        last = [10] * n
        for frame in range(8):
            recs = (Gains * n)()
            want = []
            for s in range(n):
                recs[s] = random_frame(rng, last[s], frame)
                w = Gains.from_buffer_copy(bytes(recs[s]))
                O.orc_silk_process_gains(ctypes.byref(w), nb_subfr, subfr_length, nstates)
                want.append(w)
            E.emu_silk_process_gains(ctypes.cast(recs, ctypes.c_void_p), n, nb_subfr, subfr_length, nstates)
            for s in range(n):
                fa, fb = fields(recs[s]), fields(want[s])
                assert np.array_equal(np.array(fa[0], np.float32).view(np.uint32), np.array(fb[0], np.float32).view(np.uint32)), (frame, s)
                assert fa[1:] == fb[1:], (frame, s, fa, fb)
                last[s] = recs[s].LastGainIndex


@pytest.mark.parametrize('wb,nb_subfr', [(1, 4), (0, 2)])
def test_silk_process_nlsfs_emulated(wb, nb_subfr):
    """silk_process_NLSFs: the device source on the emulator against the oracle (pinned on the reference): indices, quantised NLSFs, both coefficient sets"""
    from test_silk_nlsf_oracle_vs_ref import Nlsf, random_frame, same
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    rng = np.random.default_rng(30 + wb)
    n, order = 40, 16 if wb else 10
    prev = [None] * n
    for frame in range(6):
        recs = (Nlsf * n)()
        want = []
        for s in range(n):
            recs[s] = random_frame(rng, wb, prev[s])
            w = Nlsf.from_buffer_copy(bytes(recs[s]))
            O.orc_silk_process_NLSFs(ctypes.byref(w), wb, nb_subfr)
            want.append(w)
        E.emu_silk_process_nlsfs(ctypes.cast(recs, ctypes.c_void_p), n, wb, nb_subfr)
        for s in range(n):
            assert same(recs[s], want[s], order), (frame, s, list(recs[s].NLSFIndices), list(want[s].NLSFIndices))
            prev[s] = list(want[s].NLSF_Q15)


@pytest.mark.parametrize('order,nb_subfr,subfr_length', [(16, 4, 80), (10, 4, 60), (16, 2, 80)])
def test_silk_find_lpc_emulated(order, nb_subfr, subfr_length):
    """silk_find_LPC_FLP (Burg, A2NLSF, the interpolation search): the device source on the emulator against the oracle (pinned on the reference)"""
    from test_silk_find_lpc_oracle_vs_ref import FindLpc, lpc_input, random_case
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    rng = np.random.default_rng(50 + order + nb_subfr)
    n = 36
    x = np.stack([lpc_input(rng, order, nb_subfr, subfr_length, s % 4) for s in range(n)])
    recs = (FindLpc * n)()
    want = []
    for s in range(n):
        recs[s] = random_case(rng, order, s % 7)
        w = FindLpc.from_buffer_copy(bytes(recs[s]))
        O.orc_silk_find_LPC(ctypes.byref(w), x[s].ctypes.data_as(ctypes.c_void_p), order, nb_subfr, subfr_length)
        want.append(w)
    E.emu_silk_find_lpc(ctypes.cast(recs, ctypes.c_void_p), x.ctypes.data_as(ctypes.c_void_p), n, order, nb_subfr, subfr_length)
    for s in range(n):
        assert list(recs[s].NLSF_Q15)[:order] == list(want[s].NLSF_Q15)[:order] and recs[s].NLSFInterpCoef_Q2 == want[s].NLSFInterpCoef_Q2, s


@pytest.mark.parametrize('fs_kHz,nb_subfr', [(16, 4), (8, 4), (12, 2)])
def test_silk_find_pred_coefs_emulated(fs_kHz, nb_subfr):
    """silk_find_pred_coefs_FLP (the whole prediction analysis of a SILK frame): the device source on the emulator against the oracle (pinned on the
    reference), streams that carry sum_log_gain and the previous NLSFs over several frames"""
    from test_silk_pred_coefs_oracle_vs_ref import PredCoefs, random_case, outputs, HIST
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    order, subfr_length = (16 if fs_kHz == 16 else 10), 5 * fs_kHz
    rng = np.random.default_rng(70 + fs_kHz)
    n = 20
    prev, slg = [[0] * 16 for _ in range(n)], [0] * n
    for frame in range(4):
        recs = (PredCoefs * n)()
        want, R_, X_ = [], [], []
        for s in range(n):
            recs[s], res, x = random_case(rng, order, nb_subfr, subfr_length, fs_kHz, frame, prev[s], slg[s])
            w = PredCoefs.from_buffer_copy(bytes(recs[s]))
            O.orc_silk_find_pred_coefs(ctypes.byref(w), res[HIST:].ctypes.data_as(ctypes.c_void_p), x[HIST:].ctypes.data_as(ctypes.c_void_p), order, nb_subfr, subfr_length)
            want.append(w); R_.append(res); X_.append(x)
        R_, X_ = np.stack(R_), np.stack(X_)
        E.emu_silk_find_pred_coefs(ctypes.cast(recs, ctypes.c_void_p), R_.ctypes.data_as(ctypes.c_void_p), X_.ctypes.data_as(ctypes.c_void_p), n, R_.shape[1], HIST, order,
                                   nb_subfr, subfr_length)
        for s in range(n):
            oa, ob = outputs(recs[s], order, nb_subfr), outputs(want[s], order, nb_subfr)
            assert oa == ob, (frame, s, [i for i in range(len(oa)) if oa[i] != ob[i]])
            prev[s], slg[s] = list(want[s].prev_NLSFq_Q15), want[s].sum_log_gain_Q7


@pytest.mark.parametrize('fs_kHz,nb_subfr,order', [(16, 4, 24), (8, 4, 12), (12, 2, 14)])
def test_silk_noise_shape_emulated(fs_kHz, nb_subfr, order):
    """silk_noise_shape_analysis_FLP: the device source on the emulator against the oracle (pinned on the reference), smoother states carried over frames"""
    from test_silk_noise_shape_oracle_vs_ref import NoiseShape, frame_input, random_case, outputs
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    rng = np.random.default_rng(90 + fs_kHz)
    n = 24
    state = [(np.float32(0), np.float32(0))] * n
    for frame in range(4):
        recs = (NoiseShape * n)()
        want, R_, X_ = [], [], []
        for s in range(n):
            res, x, la = frame_input(rng, fs_kHz, nb_subfr, (s + frame) % 3)
            recs[s] = random_case(rng, fs_kHz, state[s])
            w = NoiseShape.from_buffer_copy(bytes(recs[s]))
            O.orc_silk_noise_shape_analysis(ctypes.byref(w), res.ctypes.data_as(ctypes.c_void_p), x[la:].ctypes.data_as(ctypes.c_void_p), fs_kHz, nb_subfr, order)
            want.append(w); R_.append(res); X_.append(x)
        R_, X_ = np.stack(R_), np.stack(X_)
        E.emu_silk_noise_shape(ctypes.cast(recs, ctypes.c_void_p), R_.ctypes.data_as(ctypes.c_void_p), R_.shape[1], X_.ctypes.data_as(ctypes.c_void_p), X_.shape[1],
                               5 * fs_kHz, n, fs_kHz, nb_subfr, order)
        for s in range(n):
            oa, ob = outputs(recs[s], nb_subfr, order), outputs(want[s], nb_subfr, order)
            assert oa == ob, (frame, s, [i for i in range(len(oa)) if oa[i] != ob[i]])
            state[s] = (want[s].HarmShapeGain_smth, want[s].Tilt_smth)


@pytest.mark.parametrize('fs_kHz,nb_subfr,order,complexity', [(16, 4, 16, 2), (8, 4, 10, 1), (12, 4, 12, 1), (16, 2, 16, 0), (12, 2, 12, 2)])
def test_silk_find_pitch_lags_emulated(fs_kHz, nb_subfr, order, complexity):
    """silk_find_pitch_lags_FLP + silk_pitch_analysis_core_FLP: the device source on the emulator against the oracle (pinned on the reference): residual,
    voicing decision, lags, indices, LTPCorr, predGain; streams carry the previous lag / type / correlation over frames"""
    from test_silk_pitch_oracle_vs_ref import Pitch, pitch_input, random_case, outputs
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    rng = np.random.default_rng(100 + fs_kHz + nb_subfr)
    n, hist, buf_len = 20, 20 * fs_kHz, (22 + 5 * nb_subfr) * fs_kHz
    prev = [(0, 0, np.float32(0))] * n
    voiced = 0
    for frame in range(4):
        recs = (Pitch * n)()
        X_ = np.stack([pitch_input(rng, fs_kHz, nb_subfr, (s + frame // 2) % 3) for s in range(n)])
        want, wres = [], []
        for s in range(n):
            recs[s] = random_case(rng, fs_kHz, frame, prev[s])
            w = Pitch.from_buffer_copy(bytes(recs[s]))
            r = np.zeros(buf_len, np.float32)
            O.orc_silk_find_pitch_lags(ctypes.byref(w), r.ctypes.data_as(ctypes.c_void_p), X_[s, hist:].ctypes.data_as(ctypes.c_void_p), fs_kHz, nb_subfr, order, complexity)
            want.append(w); wres.append(r)
        res = np.zeros((n, buf_len), np.float32)
        E.emu_silk_find_pitch_lags(ctypes.cast(recs, ctypes.c_void_p), res.ctypes.data_as(ctypes.c_void_p), buf_len, X_.ctypes.data_as(ctypes.c_void_p), X_.shape[1], hist, n,
                                   fs_kHz, nb_subfr, order, complexity)
        for s in range(n):
            assert np.array_equal(res[s].view(np.uint32), wres[s].view(np.uint32)), (frame, s)
            assert outputs(recs[s]) == outputs(want[s]), (frame, s, outputs(recs[s]), outputs(want[s]))
            voiced += int(want[s].signalType == 2)
            prev[s] = (want[s].signalType, want[s].pitchL[nb_subfr - 1] if want[s].signalType == 2 else 0, want[s].LTPCorr)
    assert voiced > 5


@pytest.mark.parametrize('fs_kHz,frame_ms,pad', [(16, 20, 0), (12, 20, 3), (8, 20, 8), (16, 10, 1), (12, 10, 0), (8, 10, 5)])
def test_silk_encode_do_vad_emulated(fs_kHz, frame_ms, pad):
    """silk_encode_do_VAD_FLP (silk_VAD_GetSA_Q8 streamed through the three filter banks in registers) + silk_HP_variable_cutoff: the device source on
    the emulator against the oracle (pinned on the reference), 40 frames with the whole record carried; pad shifts the streams' rows off 16-byte alignment
    so that both the vector and the scalar load paths run"""
    from test_silk_vad_oracle_vs_ref import drive_vad
    E = emu_util.load_emu_ops()

    def init(recs):
        O = util.load_oracle()
        for s in range(len(recs)):
            O.orc_silk_vad_init(ctypes.byref(recs[s]))

    def do_vad(recs, pcm, activity, frame_length):
        E.emu_silk_encode_do_vad(ctypes.cast(recs, ctypes.c_void_p), pcm.ctypes.data_as(ctypes.c_void_p), pcm.shape[1], activity.ctypes.data_as(ctypes.c_void_p), len(recs),
                                 frame_length, fs_kHz)

    def hp(recs):
        E.emu_silk_hp_variable_cutoff(ctypes.cast(recs, ctypes.c_void_p), len(recs), fs_kHz)

    same, total, active = drive_vad(init, do_vad, hp, 40, fs_kHz, frame_ms, 40, 500 + fs_kHz + frame_ms, pad)
    assert same == total and 50 < active < total - 50
