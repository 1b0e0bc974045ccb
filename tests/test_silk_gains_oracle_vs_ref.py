"""Pin the gain processing of a SILK frame (oracle/silk_oracle.c: orc_silk_process_gains, orc_silk_gains_quant, lin2log / log2lin) to the reference's
silk_process_gains_FLP / silk_gains_quant (silk/float/process_gains_FLP.c, silk/gain_quant.c) of oracle/_ref: indices, quantised gains, the carried
LastGainIndex, quantOffsetType and Lambda bit for bit, over voiced / unvoiced frames, both conditional-coding modes and 2 / 4 subframes."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built (needs /root/reference)')


class Gains(ctypes.Structure):   # OrcSilkGains == RefhSilkGains
    _fields_ = [('Gains', ctypes.c_float * 4), ('ResNrg', ctypes.c_float * 4), ('LTPredCodGain', ctypes.c_float), ('input_quality', ctypes.c_float),
                ('coding_quality', ctypes.c_float), ('Lambda', ctypes.c_float), ('GainsUnq_Q16', ctypes.c_int32 * 4), ('signalType', ctypes.c_int32),
                ('quantOffsetType', ctypes.c_int32), ('SNR_dB_Q7', ctypes.c_int32), ('input_tilt_Q15', ctypes.c_int32), ('speech_activity_Q8', ctypes.c_int32),
                ('condCoding', ctypes.c_int32), ('GainsIndices', ctypes.c_int8 * 4), ('LastGainIndex', ctypes.c_int8), ('lastGainIndexPrev', ctypes.c_int8),
                ('pad', ctypes.c_int8 * 2)]


def random_frame(rng, last_index, frame):
    g = Gains()
    scale = 10.0 ** rng.uniform(-1, 4.6)
    for k in range(4):
        g.Gains[k] = np.float32(scale * rng.uniform(0.3, 3.0))
        g.ResNrg[k] = np.float32(scale * scale * rng.uniform(0, 50.0) * 100)
    g.signalType = int(rng.integers(0, 3))
    g.quantOffsetType = int(rng.integers(0, 2))
    g.LTPredCodGain = np.float32(rng.uniform(-2, 30))
    g.input_quality, g.coding_quality = np.float32(rng.random()), np.float32(rng.random())
    g.SNR_dB_Q7 = int(rng.integers(0, 40 * 128))
    g.input_tilt_Q15 = int(rng.integers(-32768, 32768))
    g.speech_activity_Q8 = int(rng.integers(0, 257))
    g.condCoding = int(rng.integers(0, 3)) if frame else 0
    g.LastGainIndex = last_index
    return g


def fields(g):
    return (list(g.Gains), list(g.GainsUnq_Q16), list(g.GainsIndices), g.LastGainIndex, g.lastGainIndexPrev, g.quantOffsetType,
            np.float32(g.Lambda).view(np.uint32))


def test_lin2log_log2lin():
    R, O = util.load_ref('c'), util.load_oracle()
    rng = np.random.default_rng(1)
    xs = list(range(0, 300)) + [int(v) for v in rng.integers(0, 2 ** 31 - 1, 3000)] + [2 ** 31 - 1, 2 ** 30, 65536]
    for x in xs:
        assert R.refh_silk_lin2log(x) == O.orc_silk_lin2log(x), x
    for x in list(range(-5, 4200)):
        assert R.refh_silk_log2lin(x) == O.orc_silk_log2lin(x), x


@pytest.mark.parametrize('nb_subfr,subfr_length,nstates', [(4, 80, 4), (4, 40, 1), (2, 80, 2), (2, 60, 3)])
def test_process_gains(nb_subfr, subfr_length, nstates):
    R, O = util.load_ref('c'), util.load_oracle()
    assert O.orc_silk_gains_size() == ctypes.sizeof(Gains)
    rng = np.random.default_rng(nb_subfr * 100 + subfr_length)
    for stream in range(40):
        last_a = last_b = 10
        for frame in range(12):
            a = random_frame(rng, last_a, frame)
            b = Gains.from_buffer_copy(bytes(a))
            b.LastGainIndex = last_b
            R.refh_silk_process_gains(ctypes.byref(a), nb_subfr, subfr_length, nstates)
            O.orc_silk_process_gains(ctypes.byref(b), nb_subfr, subfr_length, nstates)
            fa, fb = fields(a), fields(b)
            assert np.array_equal(np.array(fa[0], np.float32).view(np.uint32), np.array(fb[0], np.float32).view(np.uint32)), (stream, frame)
            assert fa[1:] == fb[1:], (stream, frame, fa, fb)
            last_a, last_b = a.LastGainIndex, b.LastGainIndex
