"""Pin the Opus-layer encoder restatement (oracle/opus_enc.inc: run_analysis + the CELT-only subset of opus_encode_native /
opus_encode_frame_native) to the reference's own opus_encode() / opus_encode_float() (oracle/_ref, pure-C build):
the AnalysisInfo of every frame, the packet bytes (TOC included) and OPUS_GET_FINAL_RANGE must be identical, frame after frame."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built (needs /root/reference)')

APP_CELT = util.OPUS_APPLICATION_RESTRICTED_CELT
APP_LD = util.OPUS_APPLICATION_RESTRICTED_LOWDELAY


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def resample_sig(x, Fs):
    """the 48 kHz test signals decimated to Fs (plain decimation: aliasing is welcome in a test signal)"""
    return np.ascontiguousarray(x[::48000 // Fs])


def analysis_both(pcm, Fs, ch, fs, is_int16, lsb):
    R, O = util.load_ref('c'), util.load_oracle()
    T = pcm.shape[0] // fs
    a, b = np.zeros((T, 30), np.float32), np.zeros((T, 30), np.float32)
    assert R.refh_run_analysis_stream(Fs, ch, fs, T, _p(pcm), is_int16, lsb, _p(a)) == 0
    assert O.orc_run_analysis_stream(Fs, ch, fs, T, _p(pcm), is_int16, lsb, _p(b)) == 0
    return a, b


@pytest.mark.parametrize('Fs', [48000, 24000, 16000])
@pytest.mark.parametrize('ch', [1, 2])
@pytest.mark.parametrize('fs_ms', [20, 10, 5, 2.5])
@pytest.mark.parametrize('sname', ['music', 'tonal', 'white'])
def test_run_analysis(Fs, ch, fs_ms, sname):
    fs = int(Fs * fs_ms / 1000)
    x = resample_sig(util.SIGNALS[sname](7 + ch, 48000 * 2, ch), Fs)
    x[Fs // 2:Fs // 2 + Fs // 10] = 0   # 100 ms of digital silence: the analysis copies the previous frame's result
    for is_int16 in (1, 0):
        pcm = x if is_int16 else (x.astype(np.float32) / 32768.0)
        a, b = analysis_both(np.ascontiguousarray(pcm), Fs, ch, fs, is_int16, 16 if is_int16 else 24)
        assert a[:, 0].sum() > 10   # the analysis produced valid frames
        bad = np.argwhere(a.view(np.uint32) != b.view(np.uint32))
        assert bad.size == 0, (bad[:5], a[bad[0][0]], b[bad[0][0]])


def encode_both(pcm, Fs, ch, fs, bitrate, vbr, cx, app=APP_CELT, ctls=(), is_int16=True, max_bytes=1276):
    R, O = util.load_ref('c'), util.load_oracle()
    T = pcm.shape[0] // fs
    c = np.array([v for p in ctls for v in p], dtype=np.int32)
    res = []
    for L, name in ((R, 'refh_encode_stream_ex' if is_int16 else 'refh_encode_stream_float_ex'), (O, 'orc_opus_encode_stream')):
        out = np.zeros((T, max_bytes), np.uint8)
        sizes, rngs = np.zeros(T, np.int32), np.zeros(T, np.uint32)
        if L is R:
            r = getattr(L, name)(Fs, ch, app, int(bitrate), int(vbr), int(cx), fs, T, _p(pcm), _p(out), max_bytes, _p(sizes), _p(rngs), _p(c), len(ctls))
        else:
            r = L.orc_opus_encode_stream(Fs, ch, app, int(bitrate), int(vbr), int(cx), fs, T, _p(pcm), int(is_int16), _p(out), max_bytes,
                                         _p(sizes), _p(rngs), _p(c), len(ctls))
        assert r == 0, (name, r)
        res.append((out, sizes, rngs))
    return res


def check_same(res, tag=''):
    (a, sa, ra), (b, sb, rb) = res
    assert np.array_equal(sa, sb), (tag, sa[:12], sb[:12])
    for f in range(len(sa)):
        assert np.array_equal(a[f, :sa[f]], b[f, :sb[f]]), (tag, f, int(np.argmax(a[f] != b[f])), sa[f])
    assert np.array_equal(ra, rb), tag


# BASELINE.json configs[2]: 48 kHz stereo 20 ms @64 kb/s VBR, complexity 10 (and the complexities around the analysis gate at 7)
@pytest.mark.parametrize('sname', ['music', 'tonal', 'white'])
@pytest.mark.parametrize('cx', [10, 8, 7, 6, 5, 0])
def test_config2_vbr64(sname, cx):
    x = util.SIGNALS[sname](11, 960 * 60, 2)
    check_same(encode_both(x, 48000, 2, 960, 64000, 1, cx), (sname, cx))


@pytest.mark.parametrize('ch,bitrate', [(2, 16000), (2, 24000), (2, 32000), (2, 48000), (2, 96000), (2, 128000), (2, 256000),
                                        (1, 12000), (1, 24000), (1, 48000), (1, 96000), (2, 8000), (1, 6000)])
@pytest.mark.parametrize('vbr,cvbr', [(1, 1), (1, 0), (0, 1)])
def test_rates_20ms(ch, bitrate, vbr, cvbr):
    """low rates switch the stream to mono / narrower bands / the stereo width fade; high ones stay full band"""
    x = util.sig_music(ch * 1000 + bitrate // 1000, 960 * 40, ch)
    check_same(encode_both(x, 48000, ch, 960, bitrate, vbr, 10, ctls=[(4020, cvbr)]), (ch, bitrate, vbr, cvbr))


@pytest.mark.parametrize('fs', [480, 240, 120])
@pytest.mark.parametrize('ch', [1, 2])
@pytest.mark.parametrize('vbr', [0, 1])
def test_short_frames(fs, ch, vbr):
    x = util.sig_tonal_noise(fs + ch, 48000, ch)
    for br in (32000, 96000):
        check_same(encode_both(x, 48000, ch, fs, br, vbr, 10), (fs, ch, vbr, br))


@pytest.mark.parametrize('Fs', [24000, 16000, 12000, 8000])
@pytest.mark.parametrize('ch', [1, 2])
def test_input_rates(Fs, ch):
    x = resample_sig(util.sig_music(Fs // 1000 + ch, 48000, ch), Fs)
    for br, vbr in ((24000, 1), (64000, 0)):
        check_same(encode_both(x, Fs, ch, Fs // 50, br, vbr, 10), (Fs, ch, br, vbr))


def test_float_api_and_lowdelay_application():
    x = util.sig_music(5, 960 * 30, 2)
    xf = np.ascontiguousarray(x.astype(np.float32) / 32768.0)
    check_same(encode_both(xf, 48000, 2, 960, 64000, 1, 10, is_int16=False), 'float')
    check_same(encode_both(x, 48000, 2, 960, 64000, 1, 10, app=APP_LD), 'lowdelay')
    # an absurd float input is cleared by the Opus layer's guard (src/opus_encoder.c:2011-2022)
    xf2 = xf.copy()
    xf2[960 * 5:960 * 6] *= 1e6
    xf2[960 * 9 + 17] = np.nan
    check_same(encode_both(xf2, 48000, 2, 960, 64000, 1, 10, is_int16=False), 'float-guard')


@pytest.mark.parametrize('ctl', [[(4008, 1101)], [(4008, 1103)], [(4008, 1104)], [(4004, 1103)], [(4022, 1)], [(4024, 3001)], [(4024, 3002)],
                                 [(4036, 12)], [(4042, 1)], [(4046, 1)], [(4014, 20)], [(4002, -1)], [(4002, -1000)]])
def test_ctls(ctl):
    """user bandwidth, max bandwidth, forced mono, signal type, lsb depth, prediction disabled, phase inversion disabled, loss %, bitrate max / auto"""
    x = util.sig_music(31, 960 * 25, 2)
    for vbr in (1, 0):
        check_same(encode_both(x, 48000, 2, 960, 40000, vbr, 10, ctls=ctl), (ctl, vbr))


def test_silence_and_level_changes():
    x = util.sig_music(77, 960 * 50, 2)
    x[960 * 10:960 * 20] = 0
    x[960 * 30:960 * 35] //= 64
    for vbr in (1, 0):
        check_same(encode_both(x, 48000, 2, 960, 64000, vbr, 10), vbr)


def test_small_output_buffers():
    """max_data_bytes below what the bitrate asks for bounds the packet (and the bitrate the frame is coded at)"""
    x = util.sig_music(3, 960 * 12, 2)
    for mb in (30, 61, 100, 200):
        check_same(encode_both(x, 48000, 2, 960, 64000, 1, 10, max_bytes=mb), mb)
        check_same(encode_both(x, 48000, 2, 960, 64000, 0, 10, max_bytes=mb), mb)


# 40 .. 120 ms: the 20 ms frames are coded one by one and repacketised (src/opus_encoder.c:1698-1838, src/repacketizer.c): code 1 / 2 / 3 packets,
# VBR and CBR (padded to the CBR size), run_analysis over the whole packet with tonality_get_info per 20 ms frame
@pytest.mark.parametrize('fs_ms', [40, 60, 80, 100, 120])
@pytest.mark.parametrize('ch,bitrate,vbr,cx', [(2, 64000, 1, 10), (2, 96000, 0, 10), (1, 32000, 1, 5), (2, 24000, 1, 8), (1, 20000, 0, 9)])
def test_multiframe_packets(fs_ms, ch, bitrate, vbr, cx):
    Fs = 48000
    fs = Fs * fs_ms // 1000
    x = util.SIGNALS['music'](20 + ch + fs_ms, fs * 8, ch)
    x[3 * fs:3 * fs + fs // 3] = 0
    res = encode_both(x, Fs, ch, fs, bitrate, vbr, cx, max_bytes=1276 * 6 if fs_ms <= 120 else 1276)
    check_same(res, (fs_ms, ch, bitrate, vbr, cx))
    sizes = res[0][1]
    assert sizes.min() > 10
    codes = {int(res[0][0][f, 0]) & 3 for f in range(len(sizes))}
    assert codes <= ({1, 2, 3} if fs_ms == 40 else {3}), codes


@pytest.mark.parametrize('Fs,fs_ms', [(24000, 60), (16000, 40)])
def test_multiframe_other_rates_float_api(Fs, fs_ms):
    fs = Fs * fs_ms // 1000
    x = resample_sig(util.SIGNALS['tonal'](31, 48000 * 2, 2), Fs)
    x = np.ascontiguousarray(x[:(x.shape[0] // fs) * fs].astype(np.float32) / 32768.0)
    res = encode_both(x, Fs, 2, fs, 40000, 1, 10, app=APP_LD, is_int16=False, max_bytes=1276 * 3)
    check_same(res, (Fs, fs_ms))
