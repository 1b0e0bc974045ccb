"""Pin the LPC estimation of a SILK frame (oracle/silk_nlsf_oracle.c: orc_silk_find_LPC = silk_find_LPC_FLP with Burg, silk_A2NLSF, the interpolation
search through silk_NLSF2A + the float LPC analysis filter) and orc_silk_A2NLSF alone to the reference's functions of oracle/_ref, bit for bit."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built (needs /root/reference)')


class FindLpc(ctypes.Structure):   # OrcSilkFindLpc == RefhSilkFindLpc == OpusB200SilkFindLpc
    _fields_ = [('NLSF_Q15', ctypes.c_int16 * 16), ('prev_NLSFq_Q15', ctypes.c_int16 * 16), ('minInvGain', ctypes.c_float), ('useInterpolatedNLSFs', ctypes.c_int32),
                ('first_frame_after_reset', ctypes.c_int32), ('NLSFInterpCoef_Q2', ctypes.c_int32)]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def lpc_input(rng, order, nb_subfr, subfr_length, kind):
    """the pre-whitened frame silk_find_pred_coefs_FLP hands over: nb_subfr blocks of order + subfr_length samples"""
    n = nb_subfr * (subfr_length + order)
    t = np.arange(n)
    if kind == 0:       # resonant: an AR(2) colouring of noise whose pole moves inside the frame (interpolation pays)
        x = rng.standard_normal(n)
        y = np.zeros(n)
        for i in range(2, n):
            r = 0.9 + 0.08 * i / n
            w = 0.3 + 1.5 * i / n
            y[i] = x[i] + 2 * r * np.cos(w) * y[i - 1] - r * r * y[i - 2]
        x = y
    elif kind == 1:     # harmonic + noise
        x = sum(np.sin(2 * np.pi * f * t / 16000.0 * (1 + h)) for h, f in enumerate(rng.uniform(100, 900, 4))) + 0.05 * rng.standard_normal(n)
    elif kind == 2:     # nearly white
        x = rng.standard_normal(n)
    else:               # almost silent with a DC step (ill-conditioned: the bandwidth-expansion retries of A2NLSF)
        x = 1e-3 * rng.standard_normal(n) + (t > n // 2) * 0.5
    return (x * 10.0 ** rng.uniform(0, 3.5)).astype(np.float32)


def random_case(rng, order, frame):
    f = FindLpc()
    prev = np.sort(rng.uniform(0.03, 0.97, order)) * 32768
    for i in range(order):
        f.prev_NLSFq_Q15[i] = int(prev[i])
    f.minInvGain = np.float32(10.0 ** rng.uniform(-4, -1))
    f.useInterpolatedNLSFs = int(rng.integers(0, 2))
    f.first_frame_after_reset = int(frame == 0)
    return f


@pytest.mark.parametrize('d', [10, 16])
def test_a2nlsf(d):
    R, O = util.load_ref('c'), util.load_oracle()
    rng = np.random.default_rng(d)
    for k in range(400):
        if k % 4 == 3:   # garbage polynomials: the retries and the white-spectrum fall-back
            a = (rng.standard_normal(d) * 65536 * rng.uniform(0.1, 4)).astype(np.int32)
        else:
            roots = rng.uniform(0.5, 0.995, d // 2) * np.exp(1j * np.sort(rng.uniform(0.05, 3.0, d // 2)))
            poly = np.real(np.poly(np.concatenate([roots, np.conj(roots)])))
            a = np.round(-poly[1:] * 65536).astype(np.int32)
        a1, a2 = a.copy(), a.copy()
        n1, n2 = np.zeros(d, np.int16), np.zeros(d, np.int16)
        R.refh_silk_A2NLSF(_p(n1), _p(a1), d)
        O.orc_silk_A2NLSF(_p(n2), _p(a2), d)
        assert np.array_equal(n1, n2) and np.array_equal(a1, a2), k


@pytest.mark.parametrize('order,nb_subfr,subfr_length', [(16, 4, 80), (10, 4, 40), (10, 4, 60), (16, 2, 80)])
def test_find_lpc(order, nb_subfr, subfr_length):
    R, O = util.load_ref('c'), util.load_oracle()
    assert O.orc_silk_find_lpc_size() == ctypes.sizeof(FindLpc)
    rng = np.random.default_rng(order * 10 + nb_subfr)
    interp = 0
    for k in range(120):
        x = lpc_input(rng, order, nb_subfr, subfr_length, k % 4)
        a = random_case(rng, order, k % 7)
        b = FindLpc.from_buffer_copy(bytes(a))
        R.refh_silk_find_LPC(ctypes.byref(a), _p(x), order, nb_subfr, subfr_length)
        O.orc_silk_find_LPC(ctypes.byref(b), _p(x), order, nb_subfr, subfr_length)
        assert list(a.NLSF_Q15)[:order] == list(b.NLSF_Q15)[:order] and a.NLSFInterpCoef_Q2 == b.NLSFInterpCoef_Q2, (k, a.NLSFInterpCoef_Q2, b.NLSFInterpCoef_Q2)
        interp += int(a.NLSFInterpCoef_Q2 < 4)
    assert (interp > 0) == (nb_subfr == 4)
