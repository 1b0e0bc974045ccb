"""Pin the NLSF processing of a SILK frame (oracle/silk_nlsf_oracle.c: silk_process_NLSFs with NLSF_encode / VQ / del_dec_quant / decode / stabilize /
NLSF2A / LPC_fit / LPC_inverse_pred_gain) to the reference's functions of oracle/_ref: indices, quantised NLSFs and both sets of prediction
coefficients bit for bit; NB/MB and WB codebooks, 10 and 20 ms, with and without interpolation, survivors 2..32."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built (needs /root/reference)')


class Nlsf(ctypes.Structure):   # OrcSilkNlsf == RefhSilkNlsf == OpusB200SilkNlsf
    _fields_ = [('NLSF_Q15', ctypes.c_int16 * 16), ('prev_NLSFq_Q15', ctypes.c_int16 * 16), ('PredCoef_Q12', (ctypes.c_int16 * 16) * 2),
                ('NLSFIndices', ctypes.c_int8 * 17), ('NLSFInterpCoef_Q2', ctypes.c_int8), ('signalType', ctypes.c_int8), ('useInterpolatedNLSFs', ctypes.c_int8),
                ('speech_activity_Q8', ctypes.c_int32), ('NLSF_MSVQ_Survivors', ctypes.c_int32)]


def random_nlsf(rng, order, kind):
    """an NLSF vector in Q15: sorted and spread like real ones / clustered (small spacings: the stabiliser and the unstable-filter loop work) /
    unsorted garbage (the stabiliser's fall-back)"""
    if kind == 'spread':
        v = np.sort(rng.uniform(0.02, 0.98, order)) * 32768
    elif kind == 'cluster':
        c = rng.uniform(0.1, 0.9, 3)
        v = np.sort(np.clip(c[rng.integers(0, 3, order)] + rng.normal(0, 0.004, order), 0.001, 0.999)) * 32768
    else:
        v = rng.uniform(0, 1, order) * 32768
    return np.clip(v, 0, 32767).astype(np.int16)


def random_frame(rng, wb, prev=None):
    order = 16 if wb else 10
    f = Nlsf()
    kind = ['spread', 'spread', 'cluster', 'garbage'][int(rng.integers(0, 4))]
    x = random_nlsf(rng, order, kind)
    for i in range(order):
        f.NLSF_Q15[i] = int(x[i])
        f.prev_NLSFq_Q15[i] = int(prev[i]) if prev is not None else int(np.sort(random_nlsf(rng, order, 'spread'))[i])
    f.useInterpolatedNLSFs = int(rng.integers(0, 2))
    f.NLSFInterpCoef_Q2 = int(rng.integers(0, 5)) if f.useInterpolatedNLSFs else 4
    f.signalType = int(rng.integers(0, 3))
    f.speech_activity_Q8 = int(rng.integers(0, 257))
    f.NLSF_MSVQ_Survivors = int(rng.choice([2, 3, 4, 6, 8, 16, 32]))
    return f


def same(a, b, order):
    return (list(a.NLSFIndices)[:order + 1] == list(b.NLSFIndices)[:order + 1] and list(a.NLSF_Q15)[:order] == list(b.NLSF_Q15)[:order] and
            list(a.PredCoef_Q12[0])[:order] == list(b.PredCoef_Q12[0])[:order] and list(a.PredCoef_Q12[1])[:order] == list(b.PredCoef_Q12[1])[:order])


@pytest.mark.parametrize('wb', [0, 1])
def test_stabilize_nlsf2a_inv_pred_gain(wb):
    R, O = util.load_ref('c'), util.load_oracle()
    order = 16 if wb else 10
    rng = np.random.default_rng(3 + wb)
    unstable = 0
    for k in range(600):
        x = random_nlsf(rng, order, ['spread', 'cluster', 'garbage'][k % 3])
        a, b = x.copy(), x.copy()
        R.refh_silk_NLSF_stabilize(a.ctypes.data_as(ctypes.c_void_p), wb)
        O.orc_silk_NLSF_stabilize(b.ctypes.data_as(ctypes.c_void_p), _delta(O, wb), order)
        assert np.array_equal(a, b), k
        pa, pb = np.zeros(order, np.int16), np.zeros(order, np.int16)
        R.refh_silk_NLSF2A(pa.ctypes.data_as(ctypes.c_void_p), a.ctypes.data_as(ctypes.c_void_p), order)
        O.orc_silk_NLSF2A(pb.ctypes.data_as(ctypes.c_void_p), a.ctypes.data_as(ctypes.c_void_p), order)
        assert np.array_equal(pa, pb), k
        q = (pa.astype(np.int32) * rng.uniform(0.9, 1.4)).clip(-32768, 32767).astype(np.int16)
        ga = R.refh_silk_LPC_inverse_pred_gain(q.ctypes.data_as(ctypes.c_void_p), order)
        gb = O.orc_silk_LPC_inverse_pred_gain(q.ctypes.data_as(ctypes.c_void_p), order)
        assert ga == gb, k
        unstable += int(ga == 0)
    assert 0 < unstable < 600


def _delta(O, wb):
    """the codebook's minimum spacings (the oracle's stabiliser takes them as an argument)"""
    nb = [250, 3, 6, 3, 3, 3, 4, 3, 3, 3, 461]
    wbv = [100, 3, 40, 3, 3, 3, 5, 14, 14, 10, 11, 3, 8, 9, 7, 3, 347]
    v = wbv if wb else nb
    return (ctypes.c_int16 * len(v))(*v)


@pytest.mark.parametrize('wb,nb_subfr', [(0, 4), (1, 4), (1, 2), (0, 2)])
def test_process_nlsfs(wb, nb_subfr):
    R, O = util.load_ref('c'), util.load_oracle()
    assert O.orc_silk_nlsf_size() == ctypes.sizeof(Nlsf)
    order = 16 if wb else 10
    rng = np.random.default_rng(10 * wb + nb_subfr)
    for stream in range(12):
        prev = None
        for frame in range(25):
            a = random_frame(rng, wb, prev)
            b = Nlsf.from_buffer_copy(bytes(a))
            R.refh_silk_process_NLSFs(ctypes.byref(a), wb, nb_subfr)
            O.orc_silk_process_NLSFs(ctypes.byref(b), wb, nb_subfr)
            assert same(a, b, order), (stream, frame, list(a.NLSFIndices), list(b.NLSFIndices))
            prev = list(a.NLSF_Q15)
