"""Device decoder source at output rates below 48 kHz, run on the host SIMT emulator, against the oracle (which
test_oracle_downsample_vs_ref.py pins on the reference): final range bit-exact, PCM within +-1 LSB at 16 bits and < 2e-6 in float
(scan re-association only).
Reference: celt/celt_decoder.c:300-405 (decimating de-emphasis), celt/bands.c:199-200 (spectrum bound)."""
import numpy as np
import pytest

import emu_util
import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built')


@pytest.mark.parametrize('Fs', [24000, 16000, 12000, 8000])
@pytest.mark.parametrize('ch,fs48,br', [(2, 960, 128000), (1, 480, 32000), (2, 120, 48000)])
def test_emu_downsampled_decode(Fs, ch, fs48, br):
    O = util.load_oracle()
    x = util.sig_music(Fs // 1000 + fs48 + ch, 6 * fs48, ch).reshape(-1, ch)[:6 * fs48]
    pk, sz, _ = util.ref_encode(x, ch, fs48, br)
    T, fsd = len(sz), fs48 * Fs // 48000
    sz = np.ascontiguousarray(sz, np.int32)
    ref, rr = np.zeros((T * fsd, ch), np.float32), np.zeros(T, np.uint32)
    assert O.orc_decode_stream_float_fs(Fs, ch, fsd, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(ref), util._ptr(rr)) == 0
    got, gr = emu_util.emu_decode_stream(pk, sz, ch, fsd, Fs)
    assert (gr == rr).all()
    assert np.abs(util.to_int16(got).astype(np.int32) - util.to_int16(ref).astype(np.int32)).max() <= 1
    assert np.abs(got - ref).max() < 2e-6
