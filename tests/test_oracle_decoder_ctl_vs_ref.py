"""The decoder ctls that touch the decode path: OPUS_SET_GAIN (src/opus_decoder.c:686-700), OPUS_SET_PHASE_INVERSION_DISABLED
(celt/celt_decoder.c:1763-1771 -> disable_inv in the band decoder), and the getters OPUS_GET_PITCH / OPUS_GET_LAST_PACKET_DURATION:
the oracle against the reference decoder with the same ctls, bit-identical."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built')

OPUS_SET_GAIN, OPUS_SET_PHASE_INVERSION_DISABLED = 4034, 4046


def run_ctl(lib_kind, ch, fs, pk, sz, ctls):
    L = util.load_ref('c') if lib_kind == 'ref' else util.load_oracle()
    fn = L.refh_decode_stream_float_ctl if lib_kind == 'ref' else L.orc_decode_stream_float_ctl
    T = len(sz)
    pcm, rr = np.zeros((T * fs, ch), np.float32), np.zeros(T, np.uint32)
    pitch, dur = np.zeros(T, np.int32), np.zeros(T, np.int32)
    c = np.ascontiguousarray(np.array(ctls, np.int32).reshape(-1))
    assert fn(48000, ch, fs, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(pcm), util._ptr(rr), len(ctls), util._ptr(c),
              util._ptr(pitch), util._ptr(dur)) == 0
    return pcm, rr, pitch, dur


def ctl_stream(ch, fs, br, seed, nfr=10, harmonic=False):
    x = (util.sig_harmonic if harmonic else util.sig_music)(seed, nfr * fs, ch).reshape(-1, ch)[:nfr * fs]
    pk, sz, _ = util.ref_encode(x, ch, fs, br)
    sz = np.ascontiguousarray(sz, np.int32)
    sz[6] = 0   # one lost packet: the gain applies to concealed audio too, the duration getter reports the concealed length
    return pk, sz


@pytest.mark.parametrize('ctls', [[(OPUS_SET_GAIN, 1536)], [(OPUS_SET_GAIN, -3000)], [(OPUS_SET_PHASE_INVERSION_DISABLED, 1)],
                                  [(OPUS_SET_GAIN, 700), (OPUS_SET_PHASE_INVERSION_DISABLED, 1)], []])
@pytest.mark.parametrize('ch,fs,br,harm', [(2, 960, 24000, False), (2, 480, 16000, True), (1, 960, 32000, True), (2, 240, 48000, False)])
def test_decoder_ctls(ctls, ch, fs, br, harm):
    pk, sz = ctl_stream(ch, fs, br, 11 + ch + fs, harmonic=harm)
    a = run_ctl('ref', ch, fs, pk, sz, ctls)
    b = run_ctl('oracle', ch, fs, pk, sz, ctls)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    if ctls == [(OPUS_SET_PHASE_INVERSION_DISABLED, 1)] and ch == 2 and fs == 960:
        plain = run_ctl('ref', ch, fs, pk, sz, [])
        assert not np.array_equal(plain[0], a[0])   # low-rate stereo does use phase inversion: the ctl changes the output
