"""SILK integer kernels: oracle restatement vs the reference's functions, bit-exact."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built (reference absent)')


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize('d,length', [(16, 320), (10, 160), (16, 496), (6, 40), (24, 100)])
def test_lpc_analysis_filter(d, length):
    R, O = util.load_ref('c'), util.load_oracle()
    rng = np.random.default_rng(d * 1000 + length)
    for amp, camp in ((3000, 2000), (32767, 32767), (100, 30000)):
        x = rng.integers(-amp, amp + 1, size=length).astype(np.int16)
        B = rng.integers(-camp, camp + 1, size=d).astype(np.int16)
        a = np.zeros(length, np.int16)
        b = np.zeros(length, np.int16)
        R.silk_LPC_analysis_filter(_p(a), _p(x), _p(B), length, d, 0)
        O.orc_silk_lpc_analysis_filter(_p(b), _p(x), _p(B), length, d)
        assert np.array_equal(a, b)


@pytest.mark.parametrize('fs_in,fs_out', [(48000, 16000), (48000, 8000), (48000, 12000), (24000, 12000), (24000, 8000), (16000, 8000)])
def test_resampler_down(fs_in, fs_out):
    R, O = util.load_ref('c'), util.load_oracle()
    rs = ctypes.create_string_buffer(R.refh_resampler_size())
    os_ = ctypes.create_string_buffer(O.orc_resampler_size())
    assert R.silk_resampler_init(rs, fs_in, fs_out, 1) == 0
    assert O.orc_resampler_init(os_, fs_in, fs_out) == 0
    rng = np.random.default_rng(fs_in // 1000 + fs_out)
    n_in, n_out = fs_in // 50, fs_out // 50
    t = 0
    for frame in range(12):
        amp = [3000, 32767, 10][frame % 3]
        x = (amp * np.sin(2 * np.pi * 440 * (np.arange(n_in) + t) / fs_in) + rng.integers(-amp // 4 - 1, amp // 4 + 1, size=n_in)).clip(-32768, 32767).astype(np.int16)
        t += n_in
        a = np.zeros(n_out, np.int16)
        b = np.zeros(n_out, np.int16)
        R.silk_resampler(rs, _p(a), _p(x), n_in)
        O.orc_resampler_process(os_, _p(b), _p(x), n_in)
        assert np.array_equal(a, b), frame


ENC_PAIRS = [(i, o, 1) for i in (8000, 12000, 16000, 24000, 48000) for o in (8000, 12000, 16000)]
DEC_PAIRS = [(i, o, 0) for i in (8000, 12000, 16000) for o in (8000, 12000, 16000, 24000, 48000)]


@pytest.mark.parametrize('fs_in,fs_out,for_enc', ENC_PAIRS + DEC_PAIRS)
@pytest.mark.parametrize('ms', [20, 10])
def test_resampler_every_rate_pair(fs_in, fs_out, for_enc, ms):
    """every pair silk_resampler_init accepts, both directions: copy, 2x all-pass up-sampling, up-sampling with FIR interpolation, down-sampling with
    1 / 2 / 3 interpolation phases; 12 consecutive frames with the carried state"""
    R, O = util.load_ref('c'), util.load_oracle()
    rs = ctypes.create_string_buffer(R.refh_resampler_size())
    os_ = ctypes.create_string_buffer(O.orc_resampler_size())
    assert R.silk_resampler_init(rs, fs_in, fs_out, for_enc) == 0
    assert O.orc_resampler_init2(os_, fs_in, fs_out, for_enc) == 0
    rng = np.random.default_rng(fs_in // 1000 + fs_out + for_enc)
    n_in, n_out = fs_in * ms // 1000, fs_out * ms // 1000
    t = 0
    for frame in range(12):
        amp = [3000, 32767, 10][frame % 3]
        x = (amp * np.sin(2 * np.pi * 440 * (np.arange(n_in) + t) / fs_in) + rng.integers(-amp // 4 - 1, amp // 4 + 1, size=n_in)).clip(-32768, 32767).astype(np.int16)
        t += n_in
        a = np.zeros(n_out, np.int16)
        b = np.zeros(n_out, np.int16)
        R.silk_resampler(rs, _p(a), _p(x), n_in)
        O.orc_resampler_process(os_, _p(b), _p(x), n_in)
        assert np.array_equal(a, b), frame


@pytest.mark.parametrize('subfr,nb,D', [(96, 4, 16), (96, 2, 16), (50, 4, 10), (56, 4, 16)])
def test_burg_and_autocorr(subfr, nb, D):
    R, O = util.load_ref('c'), util.load_oracle()
    R.silk_burg_modified_FLP.restype = ctypes.c_float
    R.silk_burg_modified_FLP.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    O.orc_silk_burg_modified.restype = ctypes.c_float
    O.orc_silk_burg_modified.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    rng = np.random.default_rng(subfr + D)
    for trial in range(6):
        L = subfr * nb
        t = np.arange(L)
        x = (3000 * np.sin(0.05 * (trial + 1) * t) + 1000 * np.sin(0.31 * t + trial) + [300, 3, 3000][trial % 3] * rng.standard_normal(L)).astype(np.float32)
        for mig in (1e-4, 0.5, 1.0 / 1e4 * 100):
            a = np.zeros(D, np.float32)
            b = np.zeros(D, np.float32)
            ra = R.silk_burg_modified_FLP(_p(a), _p(x), mig, subfr, nb, D, 0)
            rb = O.orc_silk_burg_modified(_p(b), _p(x), mig, subfr, nb, D)
            assert ra == rb and np.array_equal(a, b)
        ca = np.zeros(D + 1, np.float32)
        cb = np.zeros(D + 1, np.float32)
        R.silk_autocorrelation_FLP(_p(ca), _p(x), L, D + 1, 0)
        O.orc_silk_autocorrelation(_p(cb), _p(x), L, D + 1)
        assert np.array_equal(ca, cb)
