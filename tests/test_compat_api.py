"""libopus_b200_compat.so: the reference's public API names over the batched C ABI.  CPU part: the library loads, exports
every symbol include/opus_b200_compat.h declares, a libopus client program compiles and links against it, and the packet helpers agree
with the reference's.  GPU part: opus_encode / opus_decode / opus_multistream_decode through the shim against the reference."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import util

ROOT = util.ROOT


def compat_lib():
    from opus_b200 import capi
    capi.build_compat_library()
    return ctypes.CDLL(capi.compat_library_path())


def test_compat_symbols_and_header():
    from opus_b200 import capi
    L = compat_lib()
    missing = [s for s in capi.COMPAT_SYMBOLS if not hasattr(L, s)]
    assert not missing, missing
    hdr = open(os.path.join(ROOT, 'include', 'opus_b200_compat.h')).read()
    for s in capi.COMPAT_SYMBOLS:
        assert s + '(' in hdr, s


def build_client(tmp, extra=()):
    from opus_b200 import capi
    capi.build_compat_library()
    exe = os.path.join(tmp, 'compat_roundtrip')
    subprocess.check_call(['gcc', '-O1', '-Wall', '-Werror', '-std=c99', '-I' + os.path.join(ROOT, 'include'), *extra, '-o', exe,
                           os.path.join(ROOT, 'tests', 'compat', 'compat_roundtrip.c'), '-L' + os.path.dirname(capi.compat_library_path()), '-lopus_b200_compat',
                           '-lm', '-Wl,-rpath,' + os.path.dirname(capi.compat_library_path())])
    return exe


def test_client_program_compiles_and_links(tmp_path):
    """a plain C99 libopus client builds against the header and links against the shim (no compute without a GPU)"""
    build_client(str(tmp_path))
    if os.path.exists('/root/reference/include/opus.h'):   # and the same source against the reference's own header
        build_client(str(tmp_path), ['-DUSE_REFERENCE_HEADERS', '-I/root/reference/include'])


@pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built')
def test_packet_helpers_match_reference():
    L, R = compat_lib(), util.load_ref('c')
    rng = np.random.default_rng(5)
    for trial in range(4000):
        n = int(rng.integers(1, 60))
        p = rng.integers(0, 256, n, dtype=np.uint8)
        if trial % 3 == 0:   # plausible multi-frame headers
            p[0] = (p[0] & 0xFC) | int(rng.integers(0, 4))
            if n > 1:
                p[1] = int(rng.integers(0, 256)) if trial % 2 else int(rng.integers(1, 6)) | (0x80 if trial % 5 == 0 else 0)
        b = p.tobytes()
        for Fs in (48000, 16000, 8000):
            assert L.opus_packet_get_nb_samples(b, n, Fs) == R.opus_packet_get_nb_samples(b, n, Fs)
        assert L.opus_packet_get_nb_frames(b, n) == R.opus_packet_get_nb_frames(b, n)
        assert L.opus_packet_get_bandwidth(b) == R.opus_packet_get_bandwidth(b) and L.opus_packet_get_nb_channels(b) == R.opus_packet_get_nb_channels(b)
        res = []
        for lib in (L, R):
            toc = ctypes.c_ubyte(0)
            frames = (ctypes.c_void_p * 48)()
            size = (ctypes.c_int16 * 48)()
            off = ctypes.c_int(0)
            buf = ctypes.create_string_buffer(b, n)
            c = lib.opus_packet_parse(buf, n, ctypes.byref(toc), frames, size, ctypes.byref(off))
            base = ctypes.addressof(buf)
            res.append((c, toc.value if c > 0 else 0, [size[i] for i in range(max(c, 0))], [frames[i] - base for i in range(max(c, 0))], off.value if c > 0 else 0))
        assert res[0] == res[1], (p[:8], res)


# ---------------------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------------------
def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.gpu
def test_client_program_runs_and_matches_reference(tmp_path):
    exe = build_client(str(tmp_path))
    out = subprocess.run([exe, '12'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert 'compat roundtrip ok' in out.stdout
    ref = os.path.join(ROOT, 'oracle', '_ref', 'libopus_ref_c.so')
    if os.path.exists(ref) and os.path.exists('/root/reference/include/opus.h'):
        # the same program linked against the unmodified reference prints the same packets (sizes, TOCs, final ranges, digest)
        rexe = os.path.join(str(tmp_path), 'ref_roundtrip')
        subprocess.check_call(['gcc', '-O1', '-std=c99', '-DUSE_REFERENCE_HEADERS', '-I/root/reference/include', '-o', rexe,
                               os.path.join(ROOT, 'tests', 'compat', 'compat_roundtrip.c'), ref, '-lm', '-Wl,-rpath,' + os.path.dirname(ref)])
        want = subprocess.run([rexe, '12'], capture_output=True, text=True, timeout=600)
        assert want.returncode == 0, want.stdout
        assert [l for l in out.stdout.splitlines() if l.startswith(('frame', 'digest'))] == [l for l in want.stdout.splitlines() if l.startswith(('frame', 'digest'))]


@pytest.mark.gpu
def test_opus_decode_through_the_shim_matches_the_oracle():
    from test_oracle_golden import golden_cases
    L = compat_lib()
    L.opus_decoder_create.restype = ctypes.c_void_p
    for name, pkts, sizes, rng, pcm, ch, fs in list(golden_cases())[:4]:
        err = ctypes.c_int(0)
        dec = ctypes.c_void_p(L.opus_decoder_create(48000, ch, ctypes.byref(err)))
        assert dec and err.value == 0
        want, want_rng = util.oracle_decode_float(pkts, np.asarray(sizes, np.int32), ch, fs)
        for t in range(len(sizes)):
            out = np.zeros(fs * ch, np.int16)
            ns = L.opus_decode(dec, pkts[t].tobytes(), int(sizes[t]), _p(out), fs, 0)
            assert ns == fs, (name, t, ns)
            fr = ctypes.c_uint32(0)
            assert L.opus_decoder_ctl(dec, 4031, ctypes.byref(fr)) == 0 and fr.value == want_rng[t]
            d = np.abs(out.astype(np.int32) - util.to_int16(want[t * fs:(t + 1) * fs].reshape(-1)).astype(np.int32))
            assert d.max() <= 1, (name, t, d.max())
        L.opus_decoder_destroy(dec)


@pytest.mark.gpu
@pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built')
def test_multistream_decode_matches_reference():
    """5.1-style layout (2 coupled + 2 mono streams, 6 channels): packets from the reference's multistream encoder, decoded by the shim (two
    batched launches per packet) and by the reference's opus_multistream_decode"""
    L, R = compat_lib(), util.load_ref('c')
    Fs, fs, nframes = 48000, 960, 6
    streams, coupled, channels = 4, 2, 6
    mapping = bytes([0, 4, 1, 2, 3, 5])
    R.opus_multistream_encoder_create.restype = ctypes.c_void_p
    R.opus_multistream_decoder_create.restype = ctypes.c_void_p
    L.opus_multistream_decoder_create.restype = ctypes.c_void_p
    err = ctypes.c_int(0)
    enc = ctypes.c_void_p(R.opus_multistream_encoder_create(Fs, channels, streams, coupled, mapping, util.OPUS_APPLICATION_RESTRICTED_CELT, ctypes.byref(err)))
    assert enc and err.value == 0
    R.opus_multistream_encoder_ctl(enc, 4002, ctypes.c_int32(256000))
    rdec = ctypes.c_void_p(R.opus_multistream_decoder_create(Fs, channels, streams, coupled, mapping, ctypes.byref(err)))
    dec = ctypes.c_void_p(L.opus_multistream_decoder_create(Fs, channels, streams, coupled, mapping, ctypes.byref(err)))
    assert dec and err.value == 0
    x = np.concatenate([util.sig_music(3 + c, fs * nframes, 1) for c in range(channels)], axis=1)
    for t in range(nframes):
        frame = np.ascontiguousarray(x[t * fs:(t + 1) * fs])
        pkt = np.zeros(4000, np.uint8)
        nb = R.opus_multistream_encode(enc, _p(frame), fs, _p(pkt), 4000)
        assert nb > 0
        a, b = np.zeros((fs, channels), np.int16), np.zeros((fs, channels), np.int16)
        assert R.opus_multistream_decode(rdec, _p(pkt), nb, _p(a), fs, 0) == fs
        assert L.opus_multistream_decode(dec, _p(pkt), nb, _p(b), fs, 0) == fs
        assert np.abs(a.astype(np.int32) - b.astype(np.int32)).max() <= 1, t
        ra, rb = ctypes.c_uint32(0), ctypes.c_uint32(0)
        R.opus_multistream_decoder_ctl(rdec, 4031, ctypes.byref(ra))
        L.opus_multistream_decoder_ctl(dec, 4031, ctypes.byref(rb))
        assert ra.value == rb.value
    L.opus_multistream_decoder_destroy(dec)
