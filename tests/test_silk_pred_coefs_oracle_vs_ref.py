"""Pin the prediction analysis of a SILK frame (oracle/silk_nlsf_oracle.c: orc_silk_find_pred_coefs = silk_find_pred_coefs_FLP: LTP analysis, LTP gain
quantiser, LTP scaling, LTP residual, LPC estimation, NLSF quantiser, residual energies) to the reference's function of oracle/_ref: every index, every
coefficient, the carried sum_log_gain and previous NLSFs, bit for bit — voiced and unvoiced frames, WB / NB, 10 / 20 ms, frame after frame."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built (needs /root/reference)')


class PredCoefs(ctypes.Structure):   # OrcSilkPredCoefs == RefhSilkPredCoefs == OpusB200SilkPredCoefs
    _fields_ = [('Gains', ctypes.c_float * 4), ('pitchL', ctypes.c_int32 * 4), ('coding_quality', ctypes.c_float), ('signalType', ctypes.c_int32),
                ('condCoding', ctypes.c_int32), ('first_frame_after_reset', ctypes.c_int32), ('useInterpolatedNLSFs', ctypes.c_int32), ('PacketLoss_perc', ctypes.c_int32),
                ('nFramesPerPacket', ctypes.c_int32), ('LBRR_flag', ctypes.c_int32), ('SNR_dB_Q7', ctypes.c_int32), ('speech_activity_Q8', ctypes.c_int32),
                ('NLSF_MSVQ_Survivors', ctypes.c_int32), ('prev_NLSFq_Q15', ctypes.c_int16 * 16), ('sum_log_gain_Q7', ctypes.c_int32), ('LTPCoef', ctypes.c_float * 20),
                ('LTPredCodGain', ctypes.c_float), ('LTP_scale', ctypes.c_float), ('PredCoef', (ctypes.c_float * 16) * 2), ('ResNrg', ctypes.c_float * 4),
                ('LTPIndex', ctypes.c_int8 * 4), ('PERIndex', ctypes.c_int8), ('LTP_scaleIndex', ctypes.c_int8), ('NLSFInterpCoef_Q2', ctypes.c_int8), ('pad', ctypes.c_int8),
                ('NLSFIndices', ctypes.c_int8 * 17), ('pad2', ctypes.c_int8 * 3)]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


HIST = 400   # samples of history in front of the frame in both signals (pitch lag <= 288 + LTP_ORDER / 2 + order)


def frame_signals(rng, order, nb_subfr, subfr_length, fs_kHz, voiced):
    """(res_pitch, x): HIST samples of history + the frame + 8; a pitched signal for voiced frames (the LTP finds something), noise otherwise"""
    n = HIST + nb_subfr * subfr_length + 8
    t = np.arange(n)
    lag = int(rng.integers(2 * fs_kHz + 2, 18 * fs_kHz - 3))
    if voiced:
        base = rng.standard_normal(lag)
        x = np.tile(base, n // lag + 2)[:n] * (1 + 0.2 * np.sin(t / 97.0)) + 0.15 * rng.standard_normal(n)
    else:
        x = rng.standard_normal(n)
    x = (x * 10.0 ** rng.uniform(0.5, 3.0)).astype(np.float32)
    res = (x + 0.05 * np.abs(x).mean() * rng.standard_normal(n)).astype(np.float32)
    return res, x, lag


def random_case(rng, order, nb_subfr, subfr_length, fs_kHz, frame, prev, sum_log_gain):
    c = PredCoefs()
    voiced = rng.random() < 0.6
    res, x, lag = frame_signals(rng, order, nb_subfr, subfr_length, fs_kHz, voiced)
    c.signalType = 2 if voiced else int(rng.integers(0, 2))
    for k in range(4):
        c.Gains[k] = np.float32(10.0 ** rng.uniform(0, 3))
        c.pitchL[k] = int(np.clip(lag + rng.integers(-2, 3), 2 * fs_kHz, 18 * fs_kHz - 1)) if voiced else 0
    c.coding_quality = np.float32(rng.random())
    c.condCoding = int(rng.integers(0, 3)) if frame else 0
    c.first_frame_after_reset = int(frame == 0)
    c.useInterpolatedNLSFs = int(rng.integers(0, 2))
    c.PacketLoss_perc = int(rng.choice([0, 0, 5, 20]))
    c.nFramesPerPacket = int(rng.integers(1, 4))
    c.LBRR_flag = int(rng.integers(0, 2))
    c.SNR_dB_Q7 = int(rng.integers(10 * 128, 35 * 128))
    c.speech_activity_Q8 = int(rng.integers(0, 257))
    c.NLSF_MSVQ_Survivors = int(rng.choice([2, 4, 8, 16]))
    for i in range(16):
        c.prev_NLSFq_Q15[i] = prev[i]
    c.sum_log_gain_Q7 = sum_log_gain
    return c, res, x


def outputs(c, order, nb_subfr):
    f = lambda a, n: np.array(list(a)[:n], np.float32).view(np.uint32).tolist()
    return (f(c.LTPCoef, 5 * nb_subfr), np.float32(c.LTPredCodGain).view(np.uint32), np.float32(c.LTP_scale).view(np.uint32), f(c.PredCoef[0], order),
            f(c.PredCoef[1], order), f(c.ResNrg, nb_subfr), list(c.LTPIndex)[:nb_subfr] if c.signalType == 2 else None, c.PERIndex if c.signalType == 2 else None,
            c.LTP_scaleIndex if c.signalType == 2 else None, c.NLSFInterpCoef_Q2, list(c.NLSFIndices)[:order + 1], list(c.prev_NLSFq_Q15), c.sum_log_gain_Q7)


@pytest.mark.parametrize('fs_kHz,nb_subfr', [(16, 4), (8, 4), (12, 2), (16, 2)])
def test_find_pred_coefs(fs_kHz, nb_subfr):
    R, O = util.load_ref('c'), util.load_oracle()
    assert O.orc_silk_pred_coefs_size() == ctypes.sizeof(PredCoefs)
    order, subfr_length = (16 if fs_kHz == 16 else 10), 5 * fs_kHz
    rng = np.random.default_rng(fs_kHz * 10 + nb_subfr)
    voiced = 0
    for stream in range(10):
        prev, slg = [0] * 16, 0
        for frame in range(8):
            a, res, x = random_case(rng, order, nb_subfr, subfr_length, fs_kHz, frame, prev, slg)
            b = PredCoefs.from_buffer_copy(bytes(a))
            R.refh_silk_find_pred_coefs(ctypes.byref(a), _p(res[HIST:]), _p(x[HIST:]), order, nb_subfr, subfr_length)
            O.orc_silk_find_pred_coefs(ctypes.byref(b), _p(res[HIST:]), _p(x[HIST:]), order, nb_subfr, subfr_length)
            oa, ob = outputs(a, order, nb_subfr), outputs(b, order, nb_subfr)
            assert oa == ob, (stream, frame, [i for i in range(len(oa)) if oa[i] != ob[i]])
            prev, slg = list(a.prev_NLSFq_Q15), a.sum_log_gain_Q7
            voiced += int(a.signalType == 2 and a.LTPredCodGain > 3)
    assert voiced > 5
