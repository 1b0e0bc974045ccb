"""Pins the oracle restatement (oracle/celt_dec_oracle.c) to the reference itself:
packets produced by the reference's pure-C build are decoded by both; the range decoder's
final state must be identical on every packet and the PCM identical sample for sample
(both are scalar C with the same operation order and -ffp-contract=off)."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built (reference absent)')

CASES = []
for sig in ('music', 'tonal', 'white'):
    for ch in (1, 2):
        for fs, br in ((960, 128000), (960, 64000), (960, 24000), (480, 96000), (240, 64000), (120, 96000), (960, 510000)):
            CASES.append((sig, ch, fs, br, 0, ()))
CASES += [('music', 2, 960, 64000, 1, ()), ('tonal', 2, 960, 32000, 1, ()), ('white', 1, 960, 12000, 0, ()),
          ('music', 2, 960, 48000, 0, ((util.OPUS_SET_BANDWIDTH_REQUEST, util.OPUS_BANDWIDTH['nb']),)),
          ('music', 2, 960, 48000, 0, ((util.OPUS_SET_BANDWIDTH_REQUEST, util.OPUS_BANDWIDTH['wb']),)),
          ('music', 1, 480, 48000, 0, ((util.OPUS_SET_BANDWIDTH_REQUEST, util.OPUS_BANDWIDTH['swb']),)),
          ('music', 2, 960, 6000, 0, ()), ('tonal', 2, 960, 8000, 1, ())]


@pytest.mark.parametrize('sig,ch,fs,br,vbr,ctls', CASES)
def test_oracle_matches_ref_c(sig, ch, fs, br, vbr, ctls):
    nfr = 25 if fs == 960 else 40
    pcm = util.SIGNALS[sig](hash((sig, ch, fs, br)) % 1000, nfr * fs, ch)
    pkts, sizes, enc_rng = util.ref_encode(pcm, ch, fs, br, vbr=vbr, ctls=ctls)
    ref_pcm, ref_rng = util.ref_decode_float(pkts, sizes, ch, fs)
    assert (ref_rng == enc_rng).all()
    orc_pcm, orc_rng = util.oracle_decode_float(pkts, sizes, ch, fs)
    assert (orc_rng == ref_rng).all(), np.nonzero(orc_rng != ref_rng)[0][:5]
    assert np.array_equal(orc_pcm, ref_pcm), float(np.abs(orc_pcm - ref_pcm).max())


def test_mono_stream_to_stereo_output_and_back():
    pcm = util.sig_music(5, 20 * 960, 1)
    pkts, sizes, _ = util.ref_encode(pcm, 1, 960, 48000)
    a, ar = util.ref_decode_float(pkts, sizes, 2, 960)
    b, br = util.oracle_decode_float(pkts, sizes, 2, 960)
    assert (ar == br).all() and np.array_equal(a, b)
    pcm = util.sig_music(6, 20 * 960, 2)
    pkts, sizes, _ = util.ref_encode(pcm, 2, 960, 96000)
    a, ar = util.ref_decode_float(pkts, sizes, 1, 960)
    b, br = util.oracle_decode_float(pkts, sizes, 1, 960)
    assert (ar == br).all() and np.array_equal(a, b)


def _random_packets(seed, n, stereo_flag=None, max_len=400):
    rng = np.random.default_rng(seed)
    pk = np.zeros((n, 1500), np.uint8)
    sizes = np.zeros(n, np.int32)
    for i in range(n):
        ln = int(rng.integers(2, max_len))
        body = rng.integers(0, 256, size=ln, dtype=np.uint8)
        if rng.random() < 0.3:  # long runs of identical bytes stress the range decoder differently
            body[:] = rng.integers(0, 256)
        cfg = int(rng.integers(16, 32))
        s = int(rng.integers(0, 2)) if stereo_flag is None else stereo_flag
        pk[i, 0] = (cfg << 3) | (s << 2) | 0
        pk[i, 1:1 + ln] = body
        sizes[i] = ln + 1
    return pk, sizes


@pytest.mark.parametrize('seed', range(6))
@pytest.mark.parametrize('ch', (1, 2))
def test_fuzzed_packets_same_as_ref(seed, ch):
    """Random payloads behind a valid CELT TOC (reference test strategy: tests/test_opus_decode.c).
    Frame size varies per packet, so decode packet by packet through both decoders."""
    import ctypes
    pk, sizes = _random_packets(seed, 60)
    L = util.load_ref('c')
    O = util.load_oracle()
    err = ctypes.c_int(0)
    L.opus_decoder_create.restype = ctypes.c_void_p
    dec = ctypes.c_void_p(L.opus_decoder_create(48000, ch, ctypes.byref(err)))
    st = ctypes.create_string_buffer(O.orc_decoder_size())
    O.orc_decoder_init(st, ch)
    a = np.zeros((5760, ch), np.float32)
    b = np.zeros((5760, ch), np.float32)
    for i in range(len(sizes)):
        na = L.opus_decode_float(dec, util._ptr(pk[i]), int(sizes[i]), util._ptr(a), 5760, 0)
        nb = O.orc_opus_decode_float(st, util._ptr(pk[i]), int(sizes[i]), util._ptr(b), 5760)
        assert na == nb, (i, na, nb)
        if na > 0:
            ra = ctypes.c_uint32(0)
            L.opus_decoder_ctl(dec, 4031, ctypes.byref(ra))  # OPUS_GET_FINAL_RANGE
            assert ra.value == O.orc_final_range(st), i
            assert np.array_equal(a[:na].view(np.uint32), b[:na].view(np.uint32)), i
    L.opus_decoder_destroy(dec)


def test_silence_and_tiny_packets():
    # the silence flag (payload 0xFF 0xFF.., as the reference itself uses for a CELT fade-out,
    # src/opus_decoder.c:629-637) between normal frames; very low rates -> the <15-bit
    # coarse-energy fallbacks.  (All-zero input makes the encoder emit 1-byte payloads = DTX/PLC,
    # which is out of scope this round.)
    x = util.sig_music(4, 12 * 960, 2)
    pkts, sizes, _ = util.ref_encode(x, 2, 960, 64000)
    for k, n in ((3, 2), (4, 9), (8, 2)):
        pkts[k, 1:1 + n] = 0xFF
        sizes[k] = 1 + n
    a, ar = util.ref_decode_float(pkts, sizes, 2, 960)
    b, br = util.oracle_decode_float(pkts, sizes, 2, 960)
    assert (ar == br).all() and np.array_equal(a, b)
    for fs, br in ((120, 16000), (240, 6000), (240, 10000), (480, 6000)):
        x = util.sig_music(3, 60 * fs, 2)
        pkts, sizes, _ = util.ref_encode(x, 2, fs, br)
        assert sizes.min() > 2
        a, ar = util.ref_decode_float(pkts, sizes, 2, fs)
        b, br = util.oracle_decode_float(pkts, sizes, 2, fs)
        assert (ar == br).all() and np.array_equal(a, b)


def test_multiframe_packets_code1_code2_code3():
    """Frame packing codes 1/2/3 built by hand from single-frame packets (RFC 6716 section 3.2)."""
    import ctypes
    x = util.sig_music(11, 12 * 480, 2)
    pkts, sizes, _ = util.ref_encode(x, 2, 480, 64000, vbr=1)
    toc = int(pkts[0, 0])
    packs = []
    f = [bytes(pkts[i, 1:sizes[i]]) for i in range(12)]
    # code 1: two equal-size frames (pad the shorter one is not allowed; pick CBR-equal pairs if any)
    if len(f[0]) == len(f[1]):
        packs.append((bytes([toc | 1]) + f[0] + f[1], 2))
    # code 2: two frames, first size explicit
    def enc_size(n):
        return bytes([n]) if n < 252 else bytes([252 + (n & 3), (n - (252 + (n & 3))) >> 2])
    packs.append((bytes([toc | 2]) + enc_size(len(f[2])) + f[2] + f[3], 2))
    # code 3 VBR, 3 frames, with 2 bytes of padding
    body = bytes([toc | 3, 0x80 | 0x40 | 3, 2]) + enc_size(len(f[4])) + enc_size(len(f[5])) + f[4] + f[5] + f[6] + b'\0\0'
    packs.append((body, 3))
    # code 3 CBR, 2 frames
    g = f[7][:100].ljust(100, b'\0'), f[8][:100].ljust(100, b'\0')
    packs.append((bytes([toc | 3, 2]) + g[0] + g[1], 2))
    L = util.load_ref('c')
    O = util.load_oracle()
    err = ctypes.c_int(0)
    L.opus_decoder_create.restype = ctypes.c_void_p
    dec = ctypes.c_void_p(L.opus_decoder_create(48000, 2, ctypes.byref(err)))
    st = ctypes.create_string_buffer(O.orc_decoder_size())
    O.orc_decoder_init(st, 2)
    a = np.zeros((5760, 2), np.float32)
    b = np.zeros((5760, 2), np.float32)
    for body, nf in packs:
        buf = np.frombuffer(body, np.uint8).copy()
        na = L.opus_decode_float(dec, util._ptr(buf), len(body), util._ptr(a), 5760, 0)
        nb = O.orc_opus_decode_float(st, util._ptr(buf), len(body), util._ptr(b), 5760)
        assert na == nb == nf * 480, (na, nb)
        assert np.array_equal(a[:na], b[:na])
    L.opus_decoder_destroy(dec)
