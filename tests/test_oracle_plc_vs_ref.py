"""Packet-loss concealment of the decoder oracle (oracle/celt_plc.inc: celt_decode_lost celt/celt_decoder.c:679-1090, prefilter_and_fold
:576-625, the post-loss energy safeguard :1362-1393, the Opus-layer PLC framing src/opus_decoder.c:316-365) against the reference decoder
fed the same loss pattern: final range and float PCM bit-identical."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built')

# loss patterns over 24 packets: 1 = lost
PATTERNS = {
    'single': [5, 13],
    'burst3': [6, 7, 8, 15],
    'burst8': list(range(4, 12)),            # > 40 short blocks at 20 ms after 5 frames: pitch PLC turns into noise PLC
    'alternating': [3, 5, 7, 9, 11],         # skip_plc: a loss right after a single good packet is concealed with noise
    'first': [0, 1, 6],                      # nothing decoded yet: zeros
}


def _run(Fs, ch, fs48, br, lost, sig, nfr=24, cx=10):
    O, L = util.load_oracle(), util.load_ref('c')
    x = util.SIGNALS[sig](fs48 + ch + br // 1000, nfr * fs48, ch).reshape(-1, ch)[:nfr * fs48]
    pk, sz, _ = util.ref_encode(x, ch, fs48, br, complexity=cx)
    T, fsd = len(sz), fs48 * Fs // 48000
    sz = np.ascontiguousarray(sz, np.int32)
    for t in lost:
        sz[t] = 0
    ref, rr = np.zeros((T * fsd, ch), np.float32), np.zeros(T, np.uint32)
    assert L.refh_decode_stream_float(Fs, ch, fsd, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(ref), util._ptr(rr)) == 0
    got, gr = np.zeros((T * fsd, ch), np.float32), np.zeros(T, np.uint32)
    assert O.orc_decode_stream_float_fs(Fs, ch, fsd, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(got), util._ptr(gr)) == 0
    assert (gr == rr).all()
    bad = np.argwhere(got != ref)
    assert bad.size == 0, (bad[0], float(np.abs(got - ref).max()))
    return ref


@pytest.mark.parametrize('pattern', list(PATTERNS))
@pytest.mark.parametrize('ch,fs48,br', [(2, 960, 128000), (1, 960, 32000), (2, 480, 64000), (2, 240, 48000), (1, 120, 24000)])
def test_plc_patterns(pattern, ch, fs48, br):
    sig = list(util.SIGNALS)[(ch + fs48 // 120) % 3]
    ref = _run(48000, ch, fs48, br, PATTERNS[pattern], sig)
    assert np.abs(ref).max() > 1e-3


@pytest.mark.parametrize('sig', ['harmonic', 'swell'])
@pytest.mark.parametrize('ch', [1, 2])
def test_plc_pitched_signals(sig, ch):
    """harmonic input: the post-filter is on around the loss (prefilter_and_fold with non-zero gains) and the pitch search has a clear peak"""
    O, L = util.load_oracle(), util.load_ref('c')
    fs48, nfr = 960, 20
    x = (util.sig_harmonic if sig == 'harmonic' else util.sig_swell)(7 + ch, nfr * fs48, ch).reshape(-1, ch)[:nfr * fs48]
    pk, sz, _ = util.ref_encode(x, ch, fs48, 64000, complexity=10)
    sz = np.ascontiguousarray(sz, np.int32)
    for t in (4, 9, 10, 15):
        sz[t] = 0
    T = len(sz)
    ref, rr = np.zeros((T * fs48, ch), np.float32), np.zeros(T, np.uint32)
    assert L.refh_decode_stream_float(48000, ch, fs48, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(ref), util._ptr(rr)) == 0
    got, gr = np.zeros((T * fs48, ch), np.float32), np.zeros(T, np.uint32)
    assert O.orc_decode_stream_float(ch, fs48, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(got), util._ptr(gr)) == 0
    assert (gr == rr).all() and np.array_equal(got, ref)


@pytest.mark.parametrize('Fs', [24000, 16000, 8000])
def test_plc_downsampled_output(Fs):
    _run(Fs, 2, 960, 64000, [4, 5, 12], 'music')


@pytest.mark.parametrize('seed', range(12))
def test_plc_randomised(seed):
    rng = np.random.default_rng(1000 + seed)
    ch = int(rng.integers(1, 3)); fs48 = int(rng.choice([120, 240, 480, 960])); br = int(rng.choice([16000, 32000, 64000, 128000, 256000]))
    nfr = 30
    lost = sorted(set(int(v) for v in rng.integers(1, nfr, size=int(rng.integers(2, 12)))))
    _run(48000, ch, fs48, br, lost, list(util.SIGNALS)[seed % 3], nfr=nfr, cx=int(rng.integers(0, 11)))


@pytest.mark.parametrize('fs_arg', [2880, 720, 360, 1920])
def test_plc_long_and_odd_frame_sizes(fs_arg):
    """opus_decode(NULL, frame_size): 60 ms = 3 x 20, 15 ms = 10 + 5, 7.5 ms = 5 + 2.5, 40 ms = 2 x 20 (src/opus_decoder.c:343-365);
    packets are 20 ms (fs_arg >= 960) or 2.5 ms frames otherwise, every slot of the output holds fs_arg samples"""
    O, L = util.load_oracle(), util.load_ref('c')
    ch, fs48 = 2, 960 if fs_arg >= 960 else 120
    nfr = 14
    x = util.sig_music(fs_arg, nfr * fs48, ch).reshape(-1, ch)[:nfr * fs48]
    pk, sz, _ = util.ref_encode(x, ch, fs48, 96000)
    sz = np.ascontiguousarray(sz, np.int32)
    for t in (0, 4, 5, 9):
        sz[t] = 0
    T = len(sz)
    ref, rr = np.zeros((T * fs_arg, ch), np.float32), np.zeros(T, np.uint32)
    assert L.refh_decode_stream_float(48000, ch, fs_arg, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(ref), util._ptr(rr)) == 0
    got, gr = np.zeros((T * fs_arg, ch), np.float32), np.zeros(T, np.uint32)
    assert O.orc_decode_stream_float(ch, fs_arg, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(got), util._ptr(gr)) == 0
    assert (gr == rr).all() and np.array_equal(got, ref)


def _dtx_packets(ch, fs48, nfr, seed):
    """code-2 packets of two frames: a real frame followed by a 1-byte frame, which the decoder conceals (len <= 1: celt_decoder.c:1281)"""
    x = util.sig_music(seed, 2 * nfr * fs48, ch).reshape(-1, ch)[:2 * nfr * fs48]
    pk, sz, _ = util.ref_encode(x, ch, fs48, 64000)
    out = np.zeros((nfr, 1500), np.uint8)
    osz = np.zeros(nfr, np.int32)
    for t in range(nfr):
        a = bytes(pk[2 * t, 1:sz[2 * t]])
        assert len(a) < 252
        dtx = b'' if t % 3 == 2 else b'\x00'     # a 0-byte and a 1-byte second frame are both "lost"
        p = bytes([int(pk[2 * t, 0]) | 2, len(a)]) + a + dtx
        out[t, :len(p)] = np.frombuffer(p, np.uint8)
        osz[t] = len(p)
    return out, osz


@pytest.mark.parametrize('ch,fs48', [(2, 480), (1, 240), (2, 120)])
def test_dtx_frames_inside_a_packet(ch, fs48):
    O, L = util.load_oracle(), util.load_ref('c')
    nfr = 10
    pk, sz = _dtx_packets(ch, fs48, nfr, 3 + ch)
    fs = 2 * fs48
    ref, rr = np.zeros((nfr * fs, ch), np.float32), np.zeros(nfr, np.uint32)
    assert L.refh_decode_stream_float(48000, ch, fs, nfr, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(ref), util._ptr(rr)) == 0
    got, gr = np.zeros((nfr * fs, ch), np.float32), np.zeros(nfr, np.uint32)
    assert O.orc_decode_stream_float(ch, fs, nfr, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(got), util._ptr(gr)) == 0
    assert (gr == rr).all() and np.array_equal(got, ref)
    assert (rr == 0).all()   # the last frame of every packet is concealed: OPUS_GET_FINAL_RANGE reports 0
