"""opus_pcm_soft_clip (src/opus.c:39-166) and the int16 decode API opus_decode() = float decode + soft clip + FLOAT2INT16
(src/opus_decoder.c:875-878, 912-940): the oracle's restatement against the reference, bit-identical."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built')


def loud_signal(seed, n, ch, kind):
    """signals whose decoded version overshoots full scale: clipped sines, full-scale noise bursts, square-ish waves"""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 48000.0
    out = np.zeros((n, ch), np.float64)
    for c in range(ch):
        f0 = rng.uniform(80, 900)
        if kind == 0:
            x = 1.6 * np.sin(2 * np.pi * f0 * t + c)                  # clipped sine
        elif kind == 1:
            x = np.sign(np.sin(2 * np.pi * f0 * t + c)) * 0.98        # square wave at full scale
        else:
            x = 1.2 * np.sin(2 * np.pi * f0 * t) * (1 + 0.5 * np.sin(2 * np.pi * 3 * t)) + 0.4 * rng.standard_normal(n)
        out[:, c] = x
    return np.clip(np.rint(out * 32767), -32768, 32767).astype(np.int16)


@pytest.mark.parametrize('seed', range(10))
def test_soft_clip_function(seed):
    O, L = util.load_oracle(), util.load_ref('c')
    rng = np.random.default_rng(seed)
    C = 1 + seed % 2
    mem_a, mem_b = np.zeros(2, np.float32), np.zeros(2, np.float32)
    for call in range(6):   # the non-linearity continues across calls through declip_mem
        N = int(rng.choice([120, 240, 480, 960, 53]))
        t = np.arange(N)[:, None] / 48000.0
        x = (rng.uniform(0.5, 2.6) * np.sin(2 * np.pi * rng.uniform(100, 3000, C)[None, :] * t + rng.uniform(0, 6, C)[None, :])
             + 0.2 * rng.standard_normal((N, C))).astype(np.float32)
        if call == 3:
            x *= 0.3   # a frame that stays inside +-1
        a, b = np.ascontiguousarray(x.copy()), np.ascontiguousarray(x.copy())
        L.opus_pcm_soft_clip(util._ptr(a), N, C, util._ptr(mem_a))
        O.orc_pcm_soft_clip(util._ptr(b), N, C, util._ptr(mem_b))
        assert np.array_equal(a, b) and np.array_equal(mem_a[:C], mem_b[:C])
        assert np.abs(a).max() <= 1.0


@pytest.mark.parametrize('kind', [0, 1, 2])
@pytest.mark.parametrize('ch,fs,br', [(2, 960, 64000), (1, 480, 24000), (2, 240, 96000), (1, 120, 32000)])
def test_int16_decode_of_loud_streams(kind, ch, fs, br):
    O, L = util.load_oracle(), util.load_ref('c')
    nfr = 12
    x = loud_signal(kind * 10 + ch, nfr * fs, ch, kind)
    pk, sz, _ = util.ref_encode(x, ch, fs, br)
    sz = np.ascontiguousarray(sz, np.int32)
    if kind == 1:
        sz[5] = sz[6] = 0   # lost packets: concealed, and returned WITHOUT soft clipping (src/opus_decoder.c:764-778)
    T = len(sz)
    ref, rr = np.zeros((T * fs, ch), np.int16), np.zeros(T, np.uint32)
    assert L.refh_decode_stream(48000, ch, fs, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(ref), util._ptr(rr)) == 0
    got, gr = np.zeros((T * fs, ch), np.int16), np.zeros(T, np.uint32)
    assert O.orc_decode_stream_int16(48000, ch, fs, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(got), util._ptr(gr)) == 0
    assert (gr == rr).all() and np.array_equal(got, ref)
    # the streams really clip: the plain conversion of the float decode differs from opus_decode's int16 output
    flt, fr = np.zeros((T * fs, ch), np.float32), np.zeros(T, np.uint32)
    assert L.refh_decode_stream_float(48000, ch, fs, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(flt), util._ptr(fr)) == 0
    if kind != 2:
        assert np.abs(flt).max() > 1.0
        assert (util.to_int16(flt) != ref).any()
