"""Encoder-side kernel operations: oracle restatement (oracle/celt_enc_ops_oracle.c) vs the
reference's own functions (oracle/_ref), bit for bit."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built (reference absent)')


def _f(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize('LM,transient', [(3, 0), (3, 1), (2, 0), (2, 1), (1, 0), (1, 1), (0, 0)])
def test_mdct_forward(LM, transient):
    R, O = util.load_ref('c'), util.load_oracle()
    rng = np.random.default_rng(LM * 2 + transient)
    N = 120 << LM
    for scale in (1.0, 32768.0, 1e-3):
        x = (rng.standard_normal(N + 120) * scale).astype(np.float32)
        a = np.zeros(N, np.float32)
        b = np.zeros(N, np.float32)
        R.refh_mdct_forward(_f(x), _f(a), LM, transient)
        O.orc_mdct_forward(_f(x), _f(b), LM, transient)
        assert np.array_equal(a, b)


@pytest.mark.parametrize('length,max_pitch', [(240, 244), (480, 5), (1024, 120), (60, 17), (8, 4)])
def test_pitch_xcorr(length, max_pitch):
    R, O = util.load_ref('c'), util.load_oracle()
    rng = np.random.default_rng(length)
    x = rng.standard_normal(length).astype(np.float32)
    y = rng.standard_normal(length + max_pitch).astype(np.float32)
    a = np.zeros(max_pitch, np.float32)
    b = np.zeros(max_pitch, np.float32)
    R.refh_pitch_xcorr(_f(x), _f(y), _f(a), length, max_pitch)
    O.orc_pitch_xcorr(_f(x), _f(y), _f(b), length, max_pitch)
    assert np.array_equal(a, b)


def _unit(rng, n):
    x = rng.standard_normal(n).astype(np.float32)
    return (x / np.sqrt((x.astype(np.float64) ** 2).sum())).astype(np.float32)


@pytest.mark.parametrize('N', [2, 3, 4, 8, 11, 16, 22, 32, 44, 48, 64, 88, 96, 144, 176])
def test_exp_rotation_and_pvq_search(N):
    R, O = util.load_ref('c'), util.load_oracle()
    R.refh_pvq_search.restype = ctypes.c_float
    O.orc_pvq_search.restype = ctypes.c_float
    rng = np.random.default_rng(N)
    for K in (1, 2, 3, 5, 8, 13, 21, 40, 64, 128):
        for spread in (0, 1, 2, 3):
            for B in (1, 2, 4, 8):
                if N % B:
                    continue
                x = _unit(rng, N)
                for d in (1, -1):
                    a, b = x.copy(), x.copy()
                    R.refh_exp_rotation(_f(a), N, d, B, K, spread)
                    O.orc_exp_rotation(_f(b), N, d, B, K, spread)
                    assert np.array_equal(a, b), (K, spread, B, d)
        # search on the rotated vector (what alg_quant does), plus a few degenerate inputs
        for x in (_unit(rng, N), np.zeros(N, np.float32), np.full(N, 1e-20, np.float32)):
            a, b = x.copy(), x.copy()
            ia, ib = np.zeros(N, np.int32), np.zeros(N, np.int32)
            ya = R.refh_pvq_search(_f(a), _f(ia), K, N)
            yb = O.orc_pvq_search(_f(b), _f(ib), K, N)
            assert ya == yb and np.array_equal(ia, ib) and np.array_equal(a, b), K
            assert np.abs(ia).sum() == K


@pytest.mark.parametrize('cc', [1, 2])
@pytest.mark.parametrize('lm,short', [(3, 0), (3, 1), (2, 0), (2, 1), (1, 1), (0, 0)])
def test_enc_analysis_front_end(cc, lm, short):
    """pre-emphasis -> MDCT -> band energies -> log energies -> normalised bands, over consecutive frames (state carried)"""
    R, O = util.load_ref('c'), util.load_oracle()
    N = 120 << lm
    x = util.sig_music(7 + cc + lm, N * 6, cc).astype(np.float32) / 32768.0 if hasattr(util, 'sig_music') else None
    x = np.ascontiguousarray(x.reshape(6, N * cc))
    mem = np.zeros(2, np.float32)
    hist = np.zeros(2 * 120, np.float32)
    st = np.zeros(O.orc_enc_analysis_state_size(), np.uint8)
    for f in range(6):
        Xa, Ea, La = np.zeros(cc * N, np.float32), np.zeros(cc * 21, np.float32), np.zeros(cc * 21, np.float32)
        Xb, Eb, Lb = np.zeros_like(Xa), np.zeros_like(Ea), np.zeros_like(La)
        R.refh_celt_enc_analysis(_f(mem), _f(hist), _f(x[f]), cc, lm, short, _f(Xa), _f(Ea), _f(La))
        O.orc_celt_enc_analysis(_f(st), _f(x[f]), cc, lm, short, _f(Xb), _f(Eb), _f(Lb))
        assert np.array_equal(Ea.view(np.uint32), Eb.view(np.uint32)), f
        assert np.array_equal(La.view(np.uint32), Lb.view(np.uint32))
        assert np.array_equal(Xa.view(np.uint32), Xb.view(np.uint32))
        assert Ea.max() > 1
