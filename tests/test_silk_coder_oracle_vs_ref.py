"""Pins oracle/silk_coder_oracle.c (silk_encode_pulses + shell coder + sign coder on the range-encoder restatement) against the reference's
silk_encode_pulses (silk/encode_pulses.c:59) driving the reference's own ec_enc: the whole output buffer after ec_enc_done, ec_tell_frac before it,
and the final range, over signal types, quantisation offset types, frame lengths (incl. the zero-padded 10 ms @ 12 kHz frame) and pulse statistics
from sparse to the down-scaling (LSB) regime."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built')


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def make_pulses(rng, n, frame_length, kind):
    """kind 0: sparse +-1; 1: laplacian with a few larger values; 2: dense and large (forces the right shifts / LSB coding); 3: all zero blocks mixed in"""
    pad = ((frame_length + 15) // 16) * 16 + 16
    p = np.zeros((n, pad), np.int8)
    for s in range(n):
        if kind == 0:
            v = (rng.random(frame_length) < 0.15) * rng.choice([-1, 1], frame_length)
        elif kind == 1:
            v = np.round(rng.laplace(0, 0.9, frame_length))
        elif kind == 2:
            v = np.round(rng.laplace(0, 6.0, frame_length))
        else:
            v = np.round(rng.laplace(0, 1.5, frame_length)) * (np.repeat(rng.random((frame_length + 15) // 16) < 0.5, 16)[:frame_length])
        p[s, :frame_length] = np.clip(v, -80, 80).astype(np.int8)
        p[s, frame_length:] = 77   # garbage the coder must not read beyond its own zero padding
    return p


@pytest.mark.parametrize('frame_length', [320, 160, 240, 120, 80])
@pytest.mark.parametrize('kind', [0, 1, 2, 3])
def test_encode_pulses(frame_length, kind):
    R, O = util.load_ref('c'), util.load_oracle()
    R.refh_silk_coder_done.restype = ctypes.c_uint32
    O.orc_silk_coder_done.restype = ctypes.c_uint32
    R.refh_silk_coder_tell_frac.restype = ctypes.c_uint32
    O.orc_silk_coder_tell_frac.restype = ctypes.c_uint32
    rng = np.random.default_rng(frame_length + kind)
    n = 24
    P = make_pulses(rng, n, frame_length, kind)
    storage = 1275 if kind == 2 else 400
    for s in range(n):
        st, qo = int(rng.integers(0, 3)), int(rng.integers(0, 2))
        rc = ctypes.create_string_buffer(R.refh_silk_coder_size())
        oc = ctypes.create_string_buffer(O.orc_silk_coder_size())
        R.refh_silk_coder_init(rc, storage)
        O.orc_silk_coder_init(oc, storage)
        icdf = np.array([200, 120, 30, 0], np.uint8)   # symbols in front of the pulses, as the indices are in the real frame
        for sym in rng.integers(0, 4, 5):
            R.refh_silk_coder_icdf(rc, int(sym), _p(icdf), 8)
            O.orc_silk_coder_icdf(oc, int(sym), _p(icdf), 8)
        a, b = P[s].copy(), P[s].copy()
        R.refh_silk_encode_pulses(rc, st, qo, _p(a), frame_length)
        O.orc_silk_encode_pulses(oc, st, qo, _p(b), frame_length)
        assert R.refh_silk_coder_tell_frac(rc) == O.orc_silk_coder_tell_frac(oc), s
        ba, bb = np.zeros(storage, np.uint8), np.zeros(storage, np.uint8)
        ra, rb = R.refh_silk_coder_done(rc, _p(ba)), O.orc_silk_coder_done(oc, _p(bb))
        assert ra == rb and np.array_equal(ba, bb), s


class SilkIndices(ctypes.Structure):   # SideInfoIndices (silk/structs.h:129-141), same fields in the same order
    _fields_ = [('GainsIndices', ctypes.c_int8 * 4), ('LTPIndex', ctypes.c_int8 * 4), ('NLSFIndices', ctypes.c_int8 * 17), ('lagIndex', ctypes.c_int16),
                ('contourIndex', ctypes.c_int8), ('signalType', ctypes.c_int8), ('quantOffsetType', ctypes.c_int8), ('NLSFInterpCoef_Q2', ctypes.c_int8),
                ('PERIndex', ctypes.c_int8), ('LTP_scaleIndex', ctypes.c_int8), ('Seed', ctypes.c_int8)]


def random_indices(rng, fs_kHz, nb_subfr, cond):
    ix = SilkIndices()
    ix.signalType = int(rng.integers(0, 3))
    ix.quantOffsetType = int(rng.integers(0, 2))
    ix.GainsIndices[0] = int(rng.integers(0, 41 if cond == 2 else 64))
    for i in range(1, 4):
        ix.GainsIndices[i] = int(rng.integers(0, 41))
    order = 16 if fs_kHz == 16 else 10
    ix.NLSFIndices[0] = int(rng.integers(0, 32))
    for i in range(order):
        ix.NLSFIndices[i + 1] = int(np.clip(np.round(rng.laplace(0, 2.5)), -10, 10))
    ix.NLSFInterpCoef_Q2 = int(rng.integers(0, 5))
    ix.lagIndex = int(rng.integers(0, 16 * fs_kHz))
    ix.contourIndex = int(rng.integers(0, {(8, 4): 11, (8, 2): 3}.get((fs_kHz, nb_subfr), 34 if nb_subfr == 4 else 12)))
    ix.PERIndex = int(rng.integers(0, 3))
    for k in range(4):
        ix.LTPIndex[k] = int(rng.integers(0, 8 << ix.PERIndex))
    ix.LTP_scaleIndex = int(rng.integers(0, 3)) if cond == 0 else 0
    ix.Seed = int(rng.integers(0, 4))
    return ix


@pytest.mark.parametrize('fs_kHz,nb_subfr', [(16, 4), (16, 2), (12, 4), (12, 2), (8, 4), (8, 2)])
def test_encode_indices_then_pulses(fs_kHz, nb_subfr):
    """a stream of SILK frames: silk_encode_indices (independent, no-LTP-scaling and conditional coding, the pitch-lag delta state carried from frame
    to frame) followed by silk_encode_pulses in the same range coder, as silk_encode_frame_FLP writes them"""
    R, O = util.load_ref('c'), util.load_oracle()
    for L, d, t in ((R, 'refh_silk_coder_done', 'refh_silk_coder_tell_frac'), (O, 'orc_silk_coder_done', 'orc_silk_coder_tell_frac')):
        getattr(L, d).restype = ctypes.c_uint32
        getattr(L, t).restype = ctypes.c_uint32
    assert R.refh_silk_indices_size() == ctypes.sizeof(SilkIndices) == O.orc_silk_indices_size()
    rng = np.random.default_rng(fs_kHz * 10 + nb_subfr)
    frame_length = 5 * nb_subfr * fs_kHz
    for stream in range(12):
        prev_r, prev_o = (ctypes.c_int * 2)(0, 0), (ctypes.c_int * 2)(0, 0)
        for frame in range(6):
            cond = [0, 2, 2, 1, 2, 0][frame]
            ix = random_indices(rng, fs_kHz, nb_subfr, cond)
            if frame in (1, 2):
                ix.signalType = 2   # voiced after voiced: the delta lag path
            P = make_pulses(rng, 1, frame_length, int(rng.integers(0, 4)))[0]
            rc = ctypes.create_string_buffer(R.refh_silk_coder_size())
            oc = ctypes.create_string_buffer(O.orc_silk_coder_size())
            R.refh_silk_coder_init(rc, 1275)
            O.orc_silk_coder_init(oc, 1275)
            R.refh_silk_encode_indices(rc, fs_kHz, nb_subfr, ctypes.byref(ix), cond, prev_r)
            O.orc_silk_encode_indices(oc, fs_kHz, nb_subfr, ctypes.byref(ix), cond, prev_o)
            assert list(prev_r) == list(prev_o)
            a, b = P.copy(), P.copy()
            R.refh_silk_encode_pulses(rc, int(ix.signalType), int(ix.quantOffsetType), _p(a), frame_length)
            O.orc_silk_encode_pulses(oc, int(ix.signalType), int(ix.quantOffsetType), _p(b), frame_length)
            assert R.refh_silk_coder_tell_frac(rc) == O.orc_silk_coder_tell_frac(oc), (stream, frame)
            ba, bb = np.zeros(1275, np.uint8), np.zeros(1275, np.uint8)
            assert R.refh_silk_coder_done(rc, _p(ba)) == O.orc_silk_coder_done(oc, _p(bb)) and np.array_equal(ba, bb), (stream, frame)
