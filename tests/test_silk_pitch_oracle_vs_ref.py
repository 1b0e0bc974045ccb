"""Pin the pitch analysis of a SILK frame (oracle/silk_oracle.c: orc_silk_find_pitch_lags = silk_find_pitch_lags_FLP with silk_pitch_analysis_core_FLP:
whitening, decimation to 8 / 4 kHz, the three search stages) to the reference's function of oracle/_ref: the residual, the voicing decision, the four
lags, lag / contour index, LTPCorr and predGain, bit for bit; 8 / 12 / 16 kHz, 10 / 20 ms, the three complexities."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built (needs /root/reference)')


class Pitch(ctypes.Structure):   # OrcSilkPitch == RefhSilkPitch == OpusB200SilkPitch
    _fields_ = [('signalType', ctypes.c_int32), ('first_frame_after_reset', ctypes.c_int32), ('speech_activity_Q8', ctypes.c_int32), ('prevSignalType', ctypes.c_int32),
                ('input_tilt_Q15', ctypes.c_int32), ('prevLag', ctypes.c_int32), ('pitchEstimationThreshold_Q16', ctypes.c_int32), ('LTPCorr', ctypes.c_float),
                ('predGain', ctypes.c_float), ('pitchL', ctypes.c_int32 * 4), ('lagIndex', ctypes.c_int32), ('contourIndex', ctypes.c_int32)]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def pitch_input(rng, fs_kHz, nb_subfr, kind):
    """20 ms of history + the frame + 2 ms look-ahead, int16 range: a pitched pulse train through a resonance (with vibrato), noise, or a mixture"""
    n = (20 + 5 * nb_subfr + 2) * fs_kHz + 8
    t = np.arange(n)
    f0 = rng.uniform(60, 450)
    ph = np.cumsum(f0 * (1 + 0.05 * np.sin(2 * np.pi * t / (fs_kHz * 30.0))) / (1000.0 * fs_kHz))
    voiced = sum(np.sin(2 * np.pi * h * ph + rng.uniform(0, 6.28)) / h for h in range(1, 9))
    noise = rng.standard_normal(n)
    mix = [1.0, 0.0, 0.5][kind]
    x = mix * voiced + (1 - mix) * noise + 0.02 * noise
    return (x * 10.0 ** rng.uniform(1.5, 3.8)).astype(np.float32)


def random_case(rng, fs_kHz, frame, prev):
    c = Pitch()
    c.signalType = int(rng.integers(0, 3)) if frame else 1
    c.first_frame_after_reset = int(frame == 0 and rng.random() < 0.5)
    c.speech_activity_Q8 = int(rng.integers(0, 257))
    c.prevSignalType = prev[0]
    c.input_tilt_Q15 = int(rng.integers(-20000, 20000))
    c.prevLag = prev[1]
    c.pitchEstimationThreshold_Q16 = int(rng.uniform(0.6, 0.9) * 65536)
    c.LTPCorr = prev[2]
    return c


def outputs(c):
    return (c.signalType, list(c.pitchL), c.lagIndex, c.contourIndex, np.float32(c.LTPCorr).view(np.uint32), np.float32(c.predGain).view(np.uint32))


@pytest.mark.parametrize('fs_kHz,nb_subfr,order,complexity', [(16, 4, 16, 2), (16, 4, 12, 1), (8, 4, 10, 2), (8, 4, 8, 0), (12, 4, 12, 1), (16, 2, 16, 2), (12, 2, 12, 0)])
def test_find_pitch_lags(fs_kHz, nb_subfr, order, complexity):
    R, O = util.load_ref('c'), util.load_oracle()
    assert O.orc_silk_pitch_size() == ctypes.sizeof(Pitch)
    rng = np.random.default_rng(fs_kHz + nb_subfr + order)
    hist, buf_len = 20 * fs_kHz, (22 + 5 * nb_subfr) * fs_kHz
    voiced = 0
    for stream in range(14):
        prev = (0, 0, np.float32(0))
        for frame in range(6):
            x = pitch_input(rng, fs_kHz, nb_subfr, (stream + frame // 2) % 3)
            a = random_case(rng, fs_kHz, frame, prev)
            b = Pitch.from_buffer_copy(bytes(a))
            ra, rb = np.zeros(buf_len, np.float32), np.zeros(buf_len, np.float32)
            R.refh_silk_find_pitch_lags(ctypes.byref(a), _p(ra), _p(x[hist:]), fs_kHz, nb_subfr, order, complexity)
            O.orc_silk_find_pitch_lags(ctypes.byref(b), _p(rb), _p(x[hist:]), fs_kHz, nb_subfr, order, complexity)
            assert np.array_equal(ra.view(np.uint32), rb.view(np.uint32)), (stream, frame)
            assert outputs(a) == outputs(b), (stream, frame, outputs(a), outputs(b))
            voiced += int(a.signalType == 2)
            prev = (a.signalType, a.pitchL[nb_subfr - 1] if a.signalType == 2 else 0, a.LTPCorr)
    assert voiced > 8
