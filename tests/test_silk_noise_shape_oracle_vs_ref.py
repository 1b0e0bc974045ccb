"""Pin the noise shaping analysis of a SILK frame (oracle/silk_oracle.c: orc_silk_noise_shape_analysis = silk_noise_shape_analysis_FLP: gain control,
sparseness measure, windowed (warped) autocorrelation, Schur / step-up, bandwidth expansion and coefficient limiting, gain tweaking, low-frequency /
tilt / harmonic shaping with their smoothers) to the reference's function of oracle/_ref, bit for bit."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built (needs /root/reference)')


class NoiseShape(ctypes.Structure):   # OrcSilkNoiseShape == RefhSilkNoiseShape == OpusB200SilkNoiseShape
    _fields_ = [('SNR_dB_Q7', ctypes.c_int32), ('input_quality_bands_Q15', ctypes.c_int32 * 2), ('useCBR', ctypes.c_int32), ('speech_activity_Q8', ctypes.c_int32),
                ('signalType', ctypes.c_int32), ('warping_Q16', ctypes.c_int32), ('LTPCorr', ctypes.c_float), ('predGain', ctypes.c_float), ('pitchL', ctypes.c_int32 * 4),
                ('HarmShapeGain_smth', ctypes.c_float), ('Tilt_smth', ctypes.c_float), ('input_quality', ctypes.c_float), ('coding_quality', ctypes.c_float),
                ('quantOffsetType', ctypes.c_int32), ('AR', ctypes.c_float * 96), ('Gains', ctypes.c_float * 4), ('LF_MA_shp', ctypes.c_float * 4),
                ('LF_AR_shp', ctypes.c_float * 4), ('HarmShapeGain', ctypes.c_float * 4), ('Tilt', ctypes.c_float * 4)]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def frame_input(rng, fs_kHz, nb_subfr, kind):
    """(pitch_res [frame], x with la_shape history + frame + la_shape look-ahead)"""
    fl, la = 5 * fs_kHz * nb_subfr, 5 * fs_kHz
    n = la + fl + la + 8
    t = np.arange(n)
    if kind == 0:
        x = sum(np.sin(2 * np.pi * f * t / (1000.0 * fs_kHz)) / (1 + h) for h, f in enumerate(rng.uniform(80, 3000, 6))) + 0.1 * rng.standard_normal(n)
    elif kind == 1:
        x = rng.standard_normal(n) * (1 + 4 * (t % 97 < 9))     # sparse bursts
    else:
        x = np.cumsum(rng.standard_normal(n)) * 0.05 + 0.01 * rng.standard_normal(n)   # strongly low-pass: large coefficients, the limiter works
    x = (x * 10.0 ** rng.uniform(0, 3.5)).astype(np.float32)
    res = (rng.standard_normal(fl) * np.abs(x).mean() * (1 + 5 * (np.arange(fl) % 61 < 5) * (kind == 1))).astype(np.float32)
    return res, x, la


def random_case(rng, fs_kHz, state):
    c = NoiseShape()
    c.SNR_dB_Q7 = int(rng.integers(5 * 128, 40 * 128))
    c.input_quality_bands_Q15[0], c.input_quality_bands_Q15[1] = int(rng.integers(0, 32768)), int(rng.integers(0, 32768))
    c.useCBR = int(rng.integers(0, 2))
    c.speech_activity_Q8 = int(rng.integers(0, 257))
    c.signalType = int(rng.integers(0, 3))
    c.warping_Q16 = int(rng.choice([0, fs_kHz * 983]))     # silk_control_encoder: fs_kHz * WARPING_MULTIPLIER in Q16, or 0 at low complexity
    c.LTPCorr = np.float32(rng.random())
    c.predGain = np.float32(10.0 ** rng.uniform(0, 3))
    for k in range(4):
        c.pitchL[k] = int(rng.integers(2 * fs_kHz, 18 * fs_kHz))
    c.HarmShapeGain_smth, c.Tilt_smth = state
    return c


def outputs(c, nb_subfr, order):
    f = lambda a: np.array(list(a), np.float32).view(np.uint32).tolist()
    ar = [f(c.AR[k * 24:k * 24 + order]) for k in range(nb_subfr)]
    return (ar, f(c.Gains[:nb_subfr]), f(c.LF_MA_shp[:nb_subfr]), f(c.LF_AR_shp[:nb_subfr]), f(c.HarmShapeGain[:nb_subfr]), f(c.Tilt[:nb_subfr]),
            f([c.input_quality, c.coding_quality, c.HarmShapeGain_smth, c.Tilt_smth]), c.quantOffsetType)


@pytest.mark.parametrize('fs_kHz,nb_subfr,order', [(16, 4, 24), (16, 4, 16), (8, 4, 12), (12, 2, 14), (16, 2, 24)])
def test_noise_shape_analysis(fs_kHz, nb_subfr, order):
    R, O = util.load_ref('c'), util.load_oracle()
    assert O.orc_silk_noise_shape_size() == ctypes.sizeof(NoiseShape)
    rng = np.random.default_rng(fs_kHz + nb_subfr + order)
    for stream in range(12):
        state = (np.float32(0), np.float32(0))
        for frame in range(8):
            res, x, la = frame_input(rng, fs_kHz, nb_subfr, (stream + frame) % 3)
            a = random_case(rng, fs_kHz, state)
            b = NoiseShape.from_buffer_copy(bytes(a))
            R.refh_silk_noise_shape_analysis(ctypes.byref(a), _p(res), _p(x[la:]), fs_kHz, nb_subfr, order)
            O.orc_silk_noise_shape_analysis(ctypes.byref(b), _p(res), _p(x[la:]), fs_kHz, nb_subfr, order)
            oa, ob = outputs(a, nb_subfr, order), outputs(b, nb_subfr, order)
            assert oa == ob, (stream, frame, [i for i in range(len(oa)) if oa[i] != ob[i]])
            state = (a.HarmShapeGain_smth, a.Tilt_smth)
