"""The decoder oracle at output rates below 48 kHz (opus_decoder_create(Fs): decimating de-emphasis celt/celt_decoder.c:387-401, spectrum
bound celt/bands.c:199-200) against the reference decoder: final range and float PCM bit-identical."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built')


@pytest.mark.parametrize('Fs', [24000, 16000, 12000, 8000])
@pytest.mark.parametrize('ch,fs48,br', [(2, 960, 128000), (1, 960, 32000), (2, 480, 64000), (2, 240, 48000), (1, 120, 24000), (2, 960, 24000)])
def test_downsampled_decode(Fs, ch, fs48, br):
    O, L = util.load_oracle(), util.load_ref('c')
    x = util.sig_music(Fs // 1000 + fs48, 8 * fs48, ch).reshape(-1, ch)[:8 * fs48]
    pk, sz, _ = util.ref_encode(x, ch, fs48, br)
    T, fsd = len(sz), fs48 * Fs // 48000
    sz = np.ascontiguousarray(sz, np.int32)
    ref, rr = np.zeros((T * fsd, ch), np.float32), np.zeros(T, np.uint32)
    assert L.refh_decode_stream_float(Fs, ch, fsd, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(ref), util._ptr(rr)) == 0
    got, gr = np.zeros((T * fsd, ch), np.float32), np.zeros(T, np.uint32)
    assert O.orc_decode_stream_float_fs(Fs, ch, fsd, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(got), util._ptr(gr)) == 0
    assert (gr == rr).all() and np.array_equal(got, ref)
