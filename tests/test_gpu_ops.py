"""Encoder-side batched kernel operations on the GPU, through the C ABI, vs the oracle (and the
compiled reference where it travelled): bit-exact."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _f(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope='module')
def ops():
    import opus_b200
    o = opus_b200.Ops()
    yield o
    o.close()


@pytest.mark.parametrize('n', [1, 2, 77, 4001, 40000])
def test_mdct_forward_long_block_kernel(ops, n):
    """the long-block kernel (ob_mdct_forward_tma: FFT in registers, inputs and tables staged by cp.async.bulk, persistent CTAs, two transforms
    per warp) against the oracle, and bit for bit against the general shared-memory kernel on every item"""
    O = util.load_oracle()
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((n, 1080)) * rng.choice([1.0, 300.0, 32768.0], size=(n, 1))).astype(np.float32)
    got = ops.mdct_forward(x, 3, 0)
    want = np.zeros(960, np.float32)
    for i in sorted(set(list(range(0, n, max(1, n // 40))) + [n - 1])):
        O.orc_mdct_forward(_f(x[i]), _f(want), 3, 0)
        assert np.array_equal(got[i], want), i
    import subprocess, sys, os, tempfile
    with tempfile.TemporaryDirectory() as d:   # the general kernel in a fresh process
        np.save(os.path.join(d, 'x.npy'), x)
        code = ("import sys, numpy as np; sys.path.insert(0, %r); import opus_b200; o = opus_b200.Ops(); "
                "np.save(%r, o.mdct_forward(np.load(%r), 3, 0))" % (util.ROOT, os.path.join(d, 'y.npy'), os.path.join(d, 'x.npy')))
        subprocess.check_call([sys.executable, '-c', code], env=dict(os.environ, OPUSB200_MDCT_SMEM_FFT='1'))
        assert np.array_equal(got, np.load(os.path.join(d, 'y.npy')))


@pytest.mark.parametrize('LM,transient', [(3, 0), (3, 1), (2, 0), (2, 1), (1, 0), (1, 1), (0, 0)])
def test_mdct_forward(ops, LM, transient):
    O = util.load_oracle()
    rng = np.random.default_rng(LM + 10 * transient)
    N = 120 << LM
    n = 300
    x = (rng.standard_normal((n, N + 120)) * rng.choice([1.0, 300.0, 32768.0], size=(n, 1))).astype(np.float32)
    got = ops.mdct_forward(x, LM, transient)
    want = np.zeros(N, np.float32)
    for i in range(0, n, 7):
        O.orc_mdct_forward(_f(x[i]), _f(want), LM, transient)
        assert np.array_equal(got[i], want), i
    if util.ref_available('c'):
        R = util.load_ref('c')
        R.refh_mdct_forward(_f(x[3]), _f(want), LM, transient)
        assert np.array_equal(got[3], want)
    # property at scale (reference unit test celt/tests/test_unit_mdct.c): SNR vs a float64 MDCT
    if not transient:
        k = np.arange(N)
        nn = np.arange(2 * N)
        # the MDCT input is N+overlap samples; the implied 2N window has (N-overlap)/2 zeros on each side
        win = np.sin(.5 * np.pi * np.sin(.5 * np.pi * (np.arange(120) + .5) / 120) ** 2)
        full = np.zeros(2 * N)
        z = (N - 120) // 2
        xin = x[0].astype(np.float64)
        full[z:z + N + 120] = xin
        w = np.ones(2 * N)
        w[:z] = 0
        w[z:z + 120] = win
        w[z + N:z + N + 120] = win[::-1]
        w[z + N + 120:] = 0
        ref = np.array([(full * w * np.cos(2 * np.pi / (2 * N) * (nn + .5 + N / 2) * (kk + .5))).sum() for kk in k]) / (N // 2) / 2
        # scale conventions differ by a constant factor; fit it
        g = (ref * got[0]).sum() / (ref * ref).sum()
        snr = 10 * np.log10((got[0].astype(np.float64) ** 2).sum() / ((got[0] - g * ref) ** 2).sum())
        assert snr > 120, snr


@pytest.mark.parametrize('length,max_pitch', [(240, 244), (480, 8), (1024, 64)])
def test_pitch_xcorr(ops, length, max_pitch):
    O = util.load_oracle()
    rng = np.random.default_rng(length)
    n = 200
    x = rng.standard_normal((n, length)).astype(np.float32)
    y = rng.standard_normal((n, length + max_pitch)).astype(np.float32)
    got = ops.pitch_xcorr(x, y, max_pitch)
    want = np.zeros(max_pitch, np.float32)
    for i in range(0, n, 11):
        O.orc_pitch_xcorr(_f(x[i]), _f(y[i]), _f(want), length, max_pitch)
        assert np.array_equal(got[i], want)


@pytest.mark.parametrize('N', [2, 4, 8, 16, 22, 32, 48, 88, 144, 176])
def test_rotation_and_search(ops, N):
    O = util.load_oracle()
    O.orc_pvq_search.restype = ctypes.c_float
    rng = np.random.default_rng(N)
    n = 500
    X = rng.standard_normal((n, N)).astype(np.float32)
    X /= np.sqrt((X.astype(np.float64) ** 2).sum(1))[:, None].astype(np.float32)
    X[0] = 0
    X[1] = 1e-20
    K = rng.integers(1, 129, size=n).astype(np.int32)
    for d in (1, -1):
        for stride in (1, 2, 8):
            if N % stride:
                continue
            got = ops.exp_rotation(X, K, d, stride, 3)
            for i in range(0, n, 13):
                w = X[i].copy()
                O.orc_exp_rotation(_f(w), N, d, stride, int(K[i]), 3)
                assert np.array_equal(got[i], w)
            # invertibility (celt/tests/test_unit_rotation.c)
            back = ops.exp_rotation(got, K, -d, stride, 3)
            assert np.abs(back[2:] - X[2:]).max() < 1e-5
    gx, giy, gyy = ops.pvq_search(X, K)
    assert (np.abs(giy).sum(1) == K).all()
    for i in range(0, n, 3):
        w, wi = X[i].copy(), np.zeros(N, np.int32)
        wy = O.orc_pvq_search(_f(w), _f(wi), int(K[i]), N)
        assert np.array_equal(giy[i], wi) and gyy[i] == wy and np.array_equal(gx[i], w), i


@pytest.mark.parametrize('d,length', [(16, 320), (10, 160), (24, 200)])
def test_silk_lpc_analysis_filter(ops, d, length):
    O = util.load_oracle()
    rng = np.random.default_rng(d)
    n = 3000
    x = rng.integers(-32768, 32768, size=(n, length)).astype(np.int16)
    B = rng.integers(-32768, 32768, size=(n, d)).astype(np.int16)
    got = ops.silk_lpc_analysis_filter(x, B)
    want = np.zeros(length, np.int16)
    for i in range(0, n, 97):
        O.orc_silk_lpc_analysis_filter(_f(want), _f(x[i]), _f(B[i]), length, d)
        assert np.array_equal(got[i], want)


@pytest.mark.parametrize('fs_in,fs_out', [(48000, 16000), (48000, 8000), (48000, 12000), (24000, 12000), (16000, 8000)])
def test_silk_resampler(ops, fs_in, fs_out):
    O = util.load_oracle()
    rng = np.random.default_rng(fs_in + fs_out)
    n = 1000
    rs = ops.silk_resampler(n, fs_in, fs_out)
    check = list(range(0, n, 83))
    orc = {i: ctypes.create_string_buffer(O.orc_resampler_size()) for i in check}
    for s in orc.values():
        assert O.orc_resampler_init(s, fs_in, fs_out) == 0
    n_in, n_out = fs_in // 50, fs_out // 50
    for fr in range(6):
        x = rng.integers(-32768, 32768, size=(n, n_in)).astype(np.int16)
        got = rs.process(x)
        w = np.zeros(n_out, np.int16)
        for i in check:
            O.orc_resampler_process(orc[i], _f(w), _f(x[i]), n_in)
            assert np.array_equal(got[i], w), (fr, i)
    rs.close()


@pytest.mark.parametrize('fs_in,fs_out,for_enc', [(i, o, 1) for i in (8000, 12000, 16000, 24000, 48000) for o in (8000, 12000, 16000)] +
                         [(i, o, 0) for i in (8000, 12000, 16000) for o in (8000, 12000, 16000, 24000, 48000)])
def test_silk_resampler_every_rate_pair(ops, fs_in, fs_out, for_enc):
    """all 30 (rate pair, direction) combinations silk_resampler_init accepts, 4 frames of 20 ms with carried state, bit-exact vs the oracle"""
    O = util.load_oracle()
    rng = np.random.default_rng(fs_in + fs_out + for_enc)
    n = 333
    rs = ops.silk_resampler(n, fs_in, fs_out, bool(for_enc))
    check = list(range(0, n, 37))
    orc = {i: ctypes.create_string_buffer(O.orc_resampler_size()) for i in check}
    for s in orc.values():
        assert O.orc_resampler_init2(s, fs_in, fs_out, for_enc) == 0
    n_in, n_out = fs_in // 50, fs_out // 50
    for fr in range(4):
        x = rng.integers(-32768, 32768, size=(n, n_in)).astype(np.int16)
        got = rs.process(x)
        assert got.shape == (n, n_out)
        w = np.zeros(n_out, np.int16)
        for i in check:
            O.orc_resampler_process(orc[i], _f(w), _f(x[i]), n_in)
            assert np.array_equal(got[i], w), (fr, i)
    rs.close()


def test_silk_resampler_rejects_other_pairs(ops):
    import opus_b200
    for a, b, e in [(48000, 24000, 1), (24000, 48000, 0), (44100, 16000, 1), (16000, 44100, 0)]:
        with pytest.raises(opus_b200.capi.OpusB200Error):
            ops.silk_resampler(4, a, b, bool(e))


def test_silk_burg_and_autocorr(ops):
    O = util.load_oracle()
    O.orc_silk_burg_modified.restype = ctypes.c_float
    O.orc_silk_burg_modified.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    rng = np.random.default_rng(17)
    n, subfr, nb, D = 2000, 96, 4, 16
    t = np.arange(subfr * nb)
    f = rng.uniform(0.01, 0.5, size=(n, 1))
    x = (3000 * np.sin(f * t[None, :]) + rng.choice([3.0, 300.0, 3000.0], size=(n, 1)) * rng.standard_normal((n, t.size))).astype(np.float32)
    for mig in (1e-4, 0.3):
        A, nrg = ops.silk_burg_modified(x, mig, subfr, nb, D)
        w = np.zeros(D, np.float32)
        for i in range(0, n, 53):
            r = O.orc_silk_burg_modified(_f(w), _f(x[i]), mig, subfr, nb, D)
            assert r == nrg[i] and np.array_equal(A[i], w), i
    ac = ops.silk_autocorrelation(x, D + 1)
    w = np.zeros(D + 1, np.float32)
    for i in range(0, n, 101):
        O.orc_silk_autocorrelation(_f(w), _f(x[i]), x.shape[1], D + 1)
        assert np.array_equal(ac[i], w)


@pytest.mark.parametrize('c', [(16, 4, 16, 24, 15729, 4), (16, 4, 16, 16, 0, 2), (16, 4, 16, 20, 15729, 3), (12, 4, 10, 14, 11796, 2),
                               (8, 2, 10, 16, 0, 4), (16, 4, 16, 16, 15729, 1), (16, 4, 16, 16, 0, 1), (8, 4, 10, 12, 0, 1)])
def test_silk_nsq(ops, c):
    """silk_NSQ / silk_NSQ_del_dec: pulses, seed and carried state bit-exact against the pinned oracle over 3 frames"""
    from test_emu_nsq import make_batch, oracle_run
    O = util.load_oracle()
    for nm in ['orc_nsq_state_init', 'orc_silk_nsq', 'orc_nsq_state_export']:
        getattr(O, nm).restype = None
    cfg = util.nsq_cfg(*c)
    n, nf = 203, 3
    x, frames = make_batch(100 + c[0] + c[3] + c[5], cfg, n, nf)
    po, so, eo = oracle_run(O, cfg, x, frames)
    q = ops.silk_nsq(n, *c)
    for f in range(nf):
        pulses, fr = q.process(frames[:, f], x[:, f])
        assert np.array_equal(pulses, po[:, f]), (c, f)
        assert np.array_equal(fr['Seed'], so[:, f])
    assert np.array_equal(q.state(), eo)
    q.close()


@pytest.mark.parametrize('cc,lm', [(2, 3), (1, 3), (2, 2), (2, 1), (1, 0)])
def test_celt_enc_analysis(ops, cc, lm):
    """fused encoder front end vs the pinned oracle over 4 consecutive frames: energies and normalised bands bit-exact
    (log energies go through the device's fp64 log: allowed to differ by one float ulp)"""
    O = util.load_oracle()
    N = 120 << lm
    n, nf = 300, 4
    x = np.stack([util.sig_music(900 + s, N * nf, cc).astype(np.float32).reshape(nf, N * cc) / 32768.0 for s in range(24)])
    x = np.tile(x, (n // 24 + 1, 1, 1))[:n]
    sb = (np.arange(n) % 3 == 1).astype(np.int32) if lm else np.zeros(n, np.int32)
    an = ops.celt_analysis(n, cc)
    st_o = [np.zeros(O.orc_enc_analysis_state_size(), np.uint8) for _ in range(n)]
    for f in range(nf):
        pcm = np.ascontiguousarray(x[:, f])
        X, En, Lg = an.process(pcm, sb)
        for s in range(0, n, 7):
            Xo, Eo, Lo = np.zeros(cc * N, np.float32), np.zeros(cc * 21, np.float32), np.zeros(cc * 21, np.float32)
            # the oracle must follow every frame of the streams it checks
            O.orc_celt_enc_analysis(_f(st_o[s]), _f(pcm[s]), cc, lm, int(sb[s]), _f(Xo), _f(Eo), _f(Lo))
            assert np.array_equal(En[s].view(np.uint32), Eo.view(np.uint32)), (f, s)
            assert np.array_equal(X[s].view(np.uint32), Xo.view(np.uint32))
            assert np.abs(Lg[s].view(np.int32).astype(np.int64) - Lo.view(np.int32).astype(np.int64)).max() <= 1
    an.close()


@pytest.mark.parametrize('frame_length,kind', [(320, 1), (320, 2), (160, 0), (120, 1), (240, 3), (80, 2)])
def test_silk_encode_pulses(ops, frame_length, kind):
    """silk_encode_pulses on the batched range-encoder object (state resident on the device between calls), symbols coded in front of the pulses:
    buffer bytes, ec_tell_frac and final range against the oracle (pinned on the reference's silk_encode_pulses + ec_enc)"""
    from test_silk_coder_oracle_vs_ref import make_pulses
    O = util.load_oracle()
    O.orc_silk_coder_done.restype = ctypes.c_uint32
    O.orc_silk_coder_tell_frac.restype = ctypes.c_uint32
    rng = np.random.default_rng(11 * frame_length + kind)
    n, storage = 600, (1275 if kind == 2 else 300)
    P = make_pulses(rng, n, frame_length, kind)
    st = rng.integers(0, 3, n).astype(np.int32)
    qo = rng.integers(0, 2, n).astype(np.int32)
    icdf = np.array([200, 120, 30, 0], np.uint8)
    syms = [rng.integers(-1, 4, n).astype(np.int32) for _ in range(3)]
    re = ops.range_encoder(n, 1275)
    re.init(storage)
    for sy in syms:
        re.icdf(sy, icdf, 8)
    re.encode_pulses(st, qo, P, frame_length)
    buf, tf, fr, er = re.done()
    assert (er == 0).all()
    for s in range(0, n, 7):
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
        assert np.array_equal(buf[s], want), s
    re.close()


@pytest.mark.parametrize('fs_kHz,nb_subfr', [(16, 4), (16, 2), (12, 4), (8, 4), (8, 2)])
def test_silk_frame_bitstream(ops, fs_kHz, nb_subfr):
    """the bitstream of a SILK frame on the device, three consecutive frames per stream: silk_encode_indices (independent / no-LTP-scaling /
    conditional coding, the pitch-lag delta state carried inside the object) + silk_encode_pulses in one range coder; bytes, ec_tell_frac and final
    range against the oracle (pinned on the reference's own functions)"""
    from test_silk_coder_oracle_vs_ref import SilkIndices, random_indices, make_pulses
    O = util.load_oracle()
    O.orc_silk_coder_done.restype = ctypes.c_uint32
    O.orc_silk_coder_tell_frac.restype = ctypes.c_uint32
    rng = np.random.default_rng(100 + fs_kHz + nb_subfr)
    n, storage, fl = 500, 600, 5 * nb_subfr * fs_kHz
    check = list(range(0, n, 9))
    oprev = {s: (ctypes.c_int * 2)(0, 0) for s in check}
    re = ops.range_encoder(n, 608)
    re.reset_silk_state()
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
        re.init(storage)
        re.encode_indices(fs_kHz, nb_subfr, idx, cond)
        re.encode_pulses(st, qo, P, fl)
        buf, tf, fr, er = re.done()
        assert (er == 0).all()
        for s in check:
            oc = ctypes.create_string_buffer(O.orc_silk_coder_size())
            O.orc_silk_coder_init(oc, storage)
            O.orc_silk_encode_indices(oc, fs_kHz, nb_subfr, ctypes.byref(idx[s]), int(cond[s]), oprev[s])
            q = P[s].copy()
            O.orc_silk_encode_pulses(oc, int(st[s]), int(qo[s]), _f(q), fl)
            assert O.orc_silk_coder_tell_frac(oc) == tf[s], (frame, s)
            want = np.zeros(storage, np.uint8)
            assert O.orc_silk_coder_done(oc, _f(want)) == fr[s] and np.array_equal(buf[s], want), (frame, s)
    re.close()


@pytest.mark.gpu
def test_silk_process_gains():
    """opusb200_silk_process_gains against the oracle (pinned bit for bit on silk_process_gains_FLP / silk_gains_quant): indices, quantised gains,
    LastGainIndex, quantOffsetType exact; Lambda and the float gains bit-identical (the one libm call per gain is exp / pow / sqrt of a double)"""
    import opus_b200
    from test_silk_gains_oracle_vs_ref import Gains, random_frame, fields
    O = util.load_oracle()
    ops = opus_b200.Ops()
    rng = np.random.default_rng(13)
    n = 300
    exact = total = 0
    for nb_subfr, subfr_length, nstates in [(4, 80, 4), (2, 60, 1)]:
        last = [10] * n
        for frame in range(6):
            recs = (Gains * n)()
            want = []
            for s in range(n):
                recs[s] = random_frame(rng, last[s], frame)
                w = Gains.from_buffer_copy(bytes(recs[s]))
                O.orc_silk_process_gains(ctypes.byref(w), nb_subfr, subfr_length, nstates)
                want.append(w)
            ops.silk_process_gains(recs, nb_subfr, subfr_length, nstates)
            for s in range(n):
                fa, fb = fields(recs[s]), fields(want[s])
                total += 1
                if fa[1:6] != fb[1:6]:
                    # a last-bit difference of the device's pow / exp can move a gain across a quantiser boundary: allow it rarely, and only by one step
                    assert max(abs(a - b) for a, b in zip(fa[2], fb[2])) <= 1, (frame, s, fa, fb)
                else:
                    exact += int(np.array_equal(np.array(fa[0], np.float32).view(np.uint32), np.array(fb[0], np.float32).view(np.uint32)) and fa[6] == fb[6])
                last[s] = want[s].LastGainIndex
                recs[s].LastGainIndex = want[s].LastGainIndex
    print('bit-identical records: %d of %d' % (exact, total))
    assert exact >= 0.995 * total


@pytest.mark.gpu
@pytest.mark.parametrize('fs_kHz,nb_subfr', [(16, 4), (8, 4), (12, 2)])
def test_silk_process_nlsfs(fs_kHz, nb_subfr):
    """opusb200_silk_process_nlsfs against the oracle (pinned bit for bit on silk_process_NLSFs): NLSF indices, quantised NLSFs and the prediction
    coefficients of both half-frames identical, for streams that carry their previous quantised NLSFs from frame to frame"""
    import opus_b200
    from test_silk_nlsf_oracle_vs_ref import Nlsf, random_frame, same
    O = util.load_oracle()
    ops = opus_b200.Ops()
    wb = int(fs_kHz == 16)
    rng = np.random.default_rng(40 + fs_kHz)
    n, order = 300, 16 if wb else 10
    prev = [None] * n
    for frame in range(5):
        recs = (Nlsf * n)()
        want = []
        for s in range(n):
            recs[s] = random_frame(rng, wb, prev[s])
            w = Nlsf.from_buffer_copy(bytes(recs[s]))
            O.orc_silk_process_NLSFs(ctypes.byref(w), wb, nb_subfr)
            want.append(w)
        ops.silk_process_nlsfs(recs, fs_kHz, nb_subfr)
        for s in range(n):
            assert same(recs[s], want[s], order), (frame, s)
            prev[s] = list(want[s].NLSF_Q15)


@pytest.mark.gpu
@pytest.mark.parametrize('order,nb_subfr,subfr_length', [(16, 4, 80), (10, 4, 60), (16, 2, 80)])
def test_silk_find_lpc(order, nb_subfr, subfr_length):
    """opusb200_silk_find_lpc against the oracle (pinned bit for bit on silk_find_LPC_FLP): NLSFs and the interpolation factor.  The Burg recursion and the
    residual energies are double precision with one sqrt on a clamped path: bit-identical on all but a rare stream, which may move an NLSF by an LSB"""
    import opus_b200
    from test_silk_find_lpc_oracle_vs_ref import FindLpc, lpc_input, random_case
    O = util.load_oracle()
    ops = opus_b200.Ops()
    rng = np.random.default_rng(60 + order + nb_subfr)
    n = 200
    x = np.stack([lpc_input(rng, order, nb_subfr, subfr_length, s % 4) for s in range(n)])
    recs = (FindLpc * n)()
    want = []
    for s in range(n):
        recs[s] = random_case(rng, order, s % 7)
        w = FindLpc.from_buffer_copy(bytes(recs[s]))
        O.orc_silk_find_LPC(ctypes.byref(w), x[s].ctypes.data_as(ctypes.c_void_p), order, nb_subfr, subfr_length)
        want.append(w)
    ops.silk_find_lpc(recs, x, order, nb_subfr, subfr_length)
    exact = 0
    for s in range(n):
        a, b = np.array(list(recs[s].NLSF_Q15)[:order]), np.array(list(want[s].NLSF_Q15)[:order])
        same_k = recs[s].NLSFInterpCoef_Q2 == want[s].NLSFInterpCoef_Q2
        exact += int(same_k and np.array_equal(a, b))
        if same_k:
            assert np.abs(a - b).max() <= 2, (s, a, b)
    print('bit-identical streams: %d of %d' % (exact, n))
    assert exact >= 0.97 * n


@pytest.mark.gpu
@pytest.mark.parametrize('fs_kHz,nb_subfr', [(16, 4), (8, 4), (12, 2)])
def test_silk_find_pred_coefs(fs_kHz, nb_subfr):
    """opusb200_silk_find_pred_coefs against the oracle (pinned bit for bit on silk_find_pred_coefs_FLP).  Its float parts are sums of products in double
    precision, one sqrt in the clamped Burg path and one pow for minInvGain: the device's pow can differ from glibc in the last bit, which can move an
    NLSF by an LSB on a rare stream.  Bar: every output identical on at least 97 % of the streams; on the others the LTP indices still identical."""
    import opus_b200
    from test_silk_pred_coefs_oracle_vs_ref import PredCoefs, random_case, outputs, HIST
    O = util.load_oracle()
    ops = opus_b200.Ops()
    order, subfr_length = (16 if fs_kHz == 16 else 10), 5 * fs_kHz
    rng = np.random.default_rng(80 + fs_kHz)
    n = 150
    prev, slg = [[0] * 16 for _ in range(n)], [0] * n
    exact = total = 0
    for frame in range(3):
        recs = (PredCoefs * n)()
        want, R_, X_ = [], [], []
        for s in range(n):
            recs[s], res, x = random_case(rng, order, nb_subfr, subfr_length, fs_kHz, frame, prev[s], slg[s])
            w = PredCoefs.from_buffer_copy(bytes(recs[s]))
            O.orc_silk_find_pred_coefs(ctypes.byref(w), res[HIST:].ctypes.data_as(ctypes.c_void_p), x[HIST:].ctypes.data_as(ctypes.c_void_p), order, nb_subfr, subfr_length)
            want.append(w); R_.append(res); X_.append(x)
        ops.silk_find_pred_coefs(recs, np.stack(R_), np.stack(X_), HIST, fs_kHz, nb_subfr)
        for s in range(n):
            oa, ob = outputs(recs[s], order, nb_subfr), outputs(want[s], order, nb_subfr)
            total += 1
            exact += int(oa == ob)
            assert oa[6:9] == ob[6:9] and oa[0] == ob[0], (frame, s)
            prev[s], slg[s] = list(want[s].prev_NLSFq_Q15), want[s].sum_log_gain_Q7
    print('bit-identical records: %d of %d' % (exact, total))
    assert exact >= 0.97 * total


@pytest.mark.gpu
@pytest.mark.parametrize('fs_kHz,nb_subfr,order', [(16, 4, 24), (8, 4, 12), (12, 2, 14)])
def test_silk_noise_shape_analysis(fs_kHz, nb_subfr, order):
    """opusb200_silk_noise_shape_analysis against the oracle (pinned bit for bit on silk_noise_shape_analysis_FLP).  The libm calls on the path are exp,
    log10, sqrt and pow of doubles: the device's can differ from glibc in the last bit of a float now and then.  Bar: shaping coefficients and gains within
    1e-5 relative on every stream, the integer decision identical, and most records bit-identical."""
    import opus_b200
    from test_silk_noise_shape_oracle_vs_ref import NoiseShape, frame_input, random_case, outputs
    O = util.load_oracle()
    ops = opus_b200.Ops()
    rng = np.random.default_rng(95 + fs_kHz)
    n = 120
    state = [(np.float32(0), np.float32(0))] * n
    exact = total = 0
    for frame in range(3):
        recs = (NoiseShape * n)()
        want, R_, X_ = [], [], []
        for s in range(n):
            res, x, la = frame_input(rng, fs_kHz, nb_subfr, (s + frame) % 3)
            recs[s] = random_case(rng, fs_kHz, state[s])
            w = NoiseShape.from_buffer_copy(bytes(recs[s]))
            O.orc_silk_noise_shape_analysis(ctypes.byref(w), res.ctypes.data_as(ctypes.c_void_p), x[la:].ctypes.data_as(ctypes.c_void_p), fs_kHz, nb_subfr, order)
            want.append(w); R_.append(res); X_.append(x)
        ops.silk_noise_shape_analysis(recs, np.stack(R_), np.stack(X_), 5 * fs_kHz, fs_kHz, nb_subfr, order)
        for s in range(n):
            oa, ob = outputs(recs[s], nb_subfr, order), outputs(want[s], nb_subfr, order)
            total += 1
            exact += int(oa == ob)
            assert recs[s].quantOffsetType == want[s].quantOffsetType or want[s].signalType != 2, (frame, s)
            for k in range(nb_subfr):
                a, b = np.array(recs[s].AR[k * 24:k * 24 + order]), np.array(want[s].AR[k * 24:k * 24 + order])
                assert np.abs(a - b).max() <= 1e-5 * max(1.0, np.abs(b).max()), (frame, s, k)
                assert abs(recs[s].Gains[k] - want[s].Gains[k]) <= 1e-5 * abs(want[s].Gains[k]), (frame, s, k)
            state[s] = (want[s].HarmShapeGain_smth, want[s].Tilt_smth)
    print('bit-identical records: %d of %d' % (exact, total))
    assert exact >= 0.9 * total


@pytest.mark.gpu
@pytest.mark.parametrize('fs_kHz,nb_subfr,order,complexity', [(16, 4, 16, 2), (8, 4, 10, 1), (12, 4, 12, 1), (16, 2, 16, 0)])
def test_silk_find_pitch_lags(fs_kHz, nb_subfr, order, complexity):
    """opusb200_silk_find_pitch_lags against the oracle (pinned bit for bit on silk_find_pitch_lags_FLP / silk_pitch_analysis_core_FLP).  The float path has one
    libm call family (log10 for the lag biases): a last-bit difference can flip a near-tie between two candidate lags on a rare frame.  Bar: the residual
    within 1e-6 relative everywhere, voicing decision / lags / indices identical on at least 97 % of the stream-frames."""
    import opus_b200
    from test_silk_pitch_oracle_vs_ref import Pitch, pitch_input, random_case, outputs
    O = util.load_oracle()
    ops = opus_b200.Ops()
    rng = np.random.default_rng(110 + fs_kHz + nb_subfr)
    n, hist, buf_len = 120, 20 * fs_kHz, (22 + 5 * nb_subfr) * fs_kHz
    prev = [(0, 0, np.float32(0))] * n
    exact = total = voiced = 0
    for frame in range(3):
        recs = (Pitch * n)()
        X_ = np.stack([pitch_input(rng, fs_kHz, nb_subfr, (s + frame) % 3) for s in range(n)])
        want, wres = [], []
        for s in range(n):
            recs[s] = random_case(rng, fs_kHz, frame, prev[s])
            w = Pitch.from_buffer_copy(bytes(recs[s]))
            r = np.zeros(buf_len, np.float32)
            O.orc_silk_find_pitch_lags(ctypes.byref(w), r.ctypes.data_as(ctypes.c_void_p), X_[s, hist:].ctypes.data_as(ctypes.c_void_p), fs_kHz, nb_subfr, order, complexity)
            want.append(w); wres.append(r)
        res = ops.silk_find_pitch_lags(recs, X_, hist, fs_kHz, nb_subfr, order, complexity)
        for s in range(n):
            assert np.abs(res[s] - wres[s]).max() <= 1e-6 * max(1.0, np.abs(wres[s]).max()), (frame, s)
            total += 1
            exact += int(outputs(recs[s]) == outputs(want[s]) and np.array_equal(res[s].view(np.uint32), wres[s].view(np.uint32)))
            voiced += int(want[s].signalType == 2)
            prev[s] = (want[s].signalType, want[s].pitchL[nb_subfr - 1] if want[s].signalType == 2 else 0, want[s].LTPCorr)
    print('bit-identical stream-frames: %d of %d (%d voiced)' % (exact, total, voiced))
    assert exact >= 0.97 * total and voiced > 20


@pytest.mark.parametrize('fs_kHz,frame_ms,pad', [(16, 20, 0), (12, 20, 3), (8, 10, 8), (16, 10, 1)])
def test_silk_encode_do_vad(fs_kHz, frame_ms, pad):
    """opusb200_silk_vad_init / _encode_do_vad / _hp_variable_cutoff against the oracle (pinned bit for bit on silk_VAD_Init, silk_encode_do_VAD_FLP with
    silk_VAD_GetSA_Q8 inside, silk_HP_variable_cutoff): 300 streams x 30 frames with the whole record carried (filter-bank states, noise levels, smoothed
    ratios, DTX counters, the cutoff smoother); pad moves the rows off 16-byte alignment so both load paths run"""
    import opus_b200
    from test_silk_vad_oracle_vs_ref import drive_vad
    ops = opus_b200.Ops()
    try:
        same, total, active = drive_vad(ops.silk_vad_init, lambda recs, pcm, act, fl: ops.silk_encode_do_vad(recs, pcm, fs_kHz, act, fl),
                                        lambda recs: ops.silk_hp_variable_cutoff(recs, fs_kHz), 300, fs_kHz, frame_ms, 30, 900 + fs_kHz + frame_ms, pad)
    finally:
        ops.close()
    print('bit-identical stream-frames: %d of %d (%d active)' % (same, total, active))
    assert same == total and 300 < active < total - 300
