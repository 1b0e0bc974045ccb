"""Pin oracle/lpcnet_enc_oracle.c (feature extraction: dnn/lpcnet_enc.c, freq.c, pitchdnn.c, burg.c) to the reference's own files compiled
unmodified with the seeded random PitchDNN of oracle/gen_pitchdnn_weights.py (oracle/_ref/libfargan_ref.so).  Both sides are the generic C build
with the same libm, so the comparison is bit for bit."""
import ctypes

import numpy as np
import pytest

import fargan_util as fu
import util

L = fu.load_ref()
pytestmark = pytest.mark.skipif(L is None, reason='oracle/_ref/libfargan_ref.so not built (needs /root/reference)')


def _oracle():
    O = util.load_oracle()
    O.orc_lpc_from_cepstrum.restype = ctypes.c_float
    return O


def test_fft320_against_numpy():
    O = _oracle()
    rng = np.random.default_rng(5)
    x = rng.standard_normal((320, 2)).astype(np.float32)
    y = np.zeros((320, 2), np.float32)
    O.orc_lpcnet_fft320(fu.vp(y), fu.vp(x))
    want = np.fft.fft(x[:, 0].astype(np.float64) + 1j * x[:, 1]) / 320
    assert np.abs(y[:, 0] + 1j * y[:, 1] - want).max() < 2e-7


@pytest.mark.parametrize('seed', [1, 2, 3])
def test_burg_cepstral_analysis(seed):
    O = _oracle()
    x = fu.speech_like(seed, 1)
    a, b = np.zeros(36, np.float32), np.zeros(36, np.float32)
    L.refh_burg_cepstral_analysis(fu.vp(a), fu.vp(x))
    O.orc_burg_cepstral_analysis(fu.vp(b), fu.vp(x))
    assert np.array_equal(a, b)
    assert np.all(np.isfinite(a))


@pytest.mark.parametrize('seed', [4, 5])
def test_lpc_from_cepstrum(seed):
    O = _oracle()
    L.refh_lpc_from_cepstrum.restype = ctypes.c_float
    rng = np.random.default_rng(seed)
    c = (rng.standard_normal(18) * np.array([3] + [1] * 17)).astype(np.float32)
    a, b = np.zeros(16, np.float32), np.zeros(16, np.float32)
    ea = L.refh_lpc_from_cepstrum(fu.vp(a), fu.vp(c))
    eb = O.orc_lpc_from_cepstrum(fu.vp(b), fu.vp(c))
    assert np.array_equal(a, b) and ea == eb


@pytest.mark.parametrize('seed', [10, 11])
def test_pitchdnn(seed):
    O = _oracle()
    ref = fu.RefLpcnetEnc(L)
    orc = fu.OrcLpcnetEnc(O, ref)
    rng = np.random.default_rng(seed)
    pd_off = 0
    for t in range(6):
        iff = rng.uniform(-1, 1, 88).astype(np.float32)
        xc = rng.uniform(-1, 1, 224).astype(np.float32)
        a = L.refh_pitchdnn_compute(fu.vp(ref.st), fu.vp(iff), fu.vp(xc))
        b = O.orc_pitchdnn_compute(fu.vp(orc.st[pd_off:]), fu.vp(iff), fu.vp(xc))
        assert a == b, (t, a, b)


@pytest.mark.parametrize('seed,scale', [(20, 8000.0), (21, 8000.0), (22, 30000.0), (23, 20.0), (24, 0.0)])
def test_frame_features(seed, scale):
    O = _oracle()
    ref = fu.RefLpcnetEnc(L)
    orc = fu.OrcLpcnetEnc(O, ref)
    nfr = 40
    x = fu.speech_like(seed, nfr, scale if scale else 1.0) * (1.0 if scale else 0.0)
    pitches = []
    for t in range(nfr):
        a = ref.features(x[t * 160:(t + 1) * 160])
        b = orc.features(x[t * 160:(t + 1) * 160])
        assert np.array_equal(a, b), (t, np.abs(a - b).max())
        for u, v in zip(ref.get(), orc.get()):
            assert np.array_equal(u, v), t
        pitches.append(a[18])
    assert np.all(np.isfinite(pitches))
    # reset returns to the initial state with the model kept
    orc.reset()
    ref2 = fu.RefLpcnetEnc(L)
    assert np.array_equal(ref2.features(x[:160]), orc.features(x[:160]))
