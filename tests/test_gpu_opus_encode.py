"""Parity of the CUDA Opus encoder (opusb200_opus_encode*, through the C ABI) with the reference's own opus_encode() / opus_encode_float()
(oracle/_ref pure-C build when it travelled with the repo, else the oracle restatement that is pinned to it): packets with their TOC byte and
OPUS_GET_FINAL_RANGE identical for every stream and frame.  BASELINE.json configs[2] is the first case."""
import ctypes

import numpy as np
import pytest

import util
from test_emu_opus_enc import oracle_stream, three_signals

pytestmark = pytest.mark.gpu

APP_CELT = util.OPUS_APPLICATION_RESTRICTED_CELT


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def reference_stream(pcm, Fs, ch, fs, bitrate, vbr, cx, ctls=(), app=APP_CELT, max_bytes=1276):
    """the unmodified reference (opus_encode / opus_encode_float) when oracle/_ref is there, else the pinned oracle"""
    if not util.ref_available('c'):
        return oracle_stream(pcm, Fs, ch, fs, bitrate, vbr, cx, ctls, app, max_bytes)
    R = util.load_ref('c')
    T = pcm.shape[0] // fs
    c = np.array([v for p in ctls for v in p], dtype=np.int32)
    out = np.zeros((T, max_bytes), np.uint8)
    sizes, rngs = np.zeros(T, np.int32), np.zeros(T, np.uint32)
    f = R.refh_encode_stream_ex if pcm.dtype == np.int16 else R.refh_encode_stream_float_ex
    r = f(Fs, ch, app, int(bitrate), int(vbr), int(cx), fs, T, _p(pcm), _p(out), max_bytes, _p(sizes), _p(rngs), _p(c), len(ctls))
    assert r == 0, r
    return out, sizes, rngs


def run_case(base_sigs, Fs, ch, fs, bitrate, vbr, cx, ctls=(), nframes=8, n=200, max_bytes=1276, app=APP_CELT):
    import opus_b200
    base = len(base_sigs)
    want = [reference_stream(np.ascontiguousarray(x[:nframes * fs]), Fs, ch, fs, bitrate, vbr, cx, ctls, app, max_bytes) for x in base_sigs]
    enc = opus_b200.OpusEncoderBatch(n, channels=ch, fs=Fs, application=app, bitrate=bitrate, vbr=vbr, complexity=cx)
    for r, v in ctls:
        enc.ctl(r, v)
    for t in range(nframes):
        frame = np.stack([base_sigs[s % base][t * fs:(t + 1) * fs].reshape(-1) for s in range(n)])
        pk, nb = enc.encode(frame, max_bytes, frame_size=fs)
        rng = enc.final_range()
        for s in range(n):
            w, ws, wr = want[s % base]
            assert nb[s] == ws[t], (t, s, nb[s], ws[t])
            assert np.array_equal(pk[s, :nb[s]], w[t, :ws[t]]), (t, s, int(np.argmax(pk[s, :nb[s]] != w[t, :ws[t]])))
            assert rng[s] == wr[t], (t, s)
    enc.close()


def many_signals(seed, nsamp, ch, k=4):
    out = []
    for i in range(k):
        out += three_signals(seed + 10 * i, nsamp, ch)
    return out


def test_config2_vbr64_complexity10_vs_opus_encode():
    """BASELINE.json configs[2]: 48 kHz stereo 20 ms, 64 kb/s VBR, complexity 10, RESTRICTED_CELT: 30 warm frames of 12 signals"""
    run_case(many_signals(11, 960 * 30, 2), 48000, 2, 960, 64000, 1, 10, nframes=30, n=300)


@pytest.mark.parametrize('cx', [0, 3, 5, 6, 7, 8])
def test_complexities(cx):
    run_case(many_signals(50 + cx, 960 * 8, 2, 2), 48000, 2, 960, 64000, 1, cx)


@pytest.mark.parametrize('ch,bitrate,vbr', [(2, 16000, 1), (2, 24000, 0), (2, 32000, 1), (2, 48000, 0), (1, 12000, 1), (1, 24000, 1), (2, 128000, 0),
                                            (2, 256000, 1), (1, 96000, 0)])
def test_rates(ch, bitrate, vbr):
    run_case(many_signals(70 + ch, 960 * 10, ch, 2), 48000, ch, 960, bitrate, vbr, 10, nframes=10)


@pytest.mark.parametrize('fs', [480, 240, 120])
@pytest.mark.parametrize('ch', [1, 2])
def test_short_frames(fs, ch):
    run_case(many_signals(fs + ch, fs * 20, ch, 2), 48000, ch, fs, 64000, 1, 10, nframes=20)


@pytest.mark.parametrize('Fs', [24000, 16000, 12000, 8000])
def test_input_rates(Fs):
    sigs = [np.ascontiguousarray(x[::48000 // Fs]) for x in many_signals(Fs // 1000, 960 * 8, 2, 2)]
    run_case(sigs, Fs, 2, Fs // 50, 32000, 1, 10)


def test_float_api_silence_guard_and_lowdelay():
    x = util.sig_music(9, 960 * 10, 2).astype(np.float32) / 32768.0
    y = x.copy()
    y[960 * 3:960 * 5] = 0
    z = x.copy()
    z[960 * 4:960 * 5] *= 1e6
    z[960 * 7 + 5] = np.nan
    run_case([x, y, z], 48000, 2, 960, 64000, 1, 10, nframes=10, n=67)
    run_case(three_signals(3, 960 * 6, 2), 48000, 2, 960, 64000, 1, 10, nframes=6, n=40, app=util.OPUS_APPLICATION_RESTRICTED_LOWDELAY)


@pytest.mark.parametrize('ctl', [[(4008, 1101)], [(4008, 1103)], [(4004, 1104)], [(4022, 1)], [(4024, 3001)], [(4024, 3002)], [(4042, 1)], [(4046, 1)],
                                 [(4036, 12)], [(4020, 0)], [(4014, 20)], [(4002, -1)]])
def test_ctls(ctl):
    for vbr in (1, 0):
        run_case(three_signals(31, 960 * 6, 2), 48000, 2, 960, 40000, vbr, 10, ctls=ctl, nframes=6, n=50)


def test_unsupported_calls_fail_loudly():
    import opus_b200
    from opus_b200 import capi
    with pytest.raises(capi.OpusB200Error):
        opus_b200.OpusEncoderBatch(4, application=util.OPUS_APPLICATION_AUDIO)   # may select SILK: not built
    enc = opus_b200.OpusEncoderBatch(4)
    L = capi.lib()
    pcm = np.zeros((4, 2 * 960), np.int16)
    pk, nb = np.zeros((4, 1276), np.uint8), np.zeros(4, np.int32)
    r = L.opusb200_opus_encode(enc._h, _p(pcm), 960, _p(pk), 1276, 2, _p(nb))   # a 2-byte budget: the reference emits a 'PLC frame' packet, not built
    assert r == capi.OPUS_UNIMPLEMENTED, r
    r = L.opusb200_opus_encode(enc._h, _p(pcm), 1000, _p(pk), 1276, 1276, _p(nb))   # not an Opus frame size
    assert r == capi.OPUS_BAD_ARG, r
    enc.close()


def test_submit_wait_matches_the_synchronous_call():
    """opusb200_opus_encode_submit / _wait (double-buffered staging, transfers on their own streams) gives the packets of opusb200_opus_encode,
    call for call, with two calls in flight"""
    import opus_b200
    n, fs, nf = 96, 960, 6
    rng = np.random.default_rng(5)
    t = np.arange(nf * fs)[None, :, None] / 48000.0
    x = (8000 * np.sin(2 * np.pi * (200 + 30 * np.arange(n)[:, None, None]) * t * (1 + np.arange(2)[None, None, :])) + rng.normal(0, 900, (n, nf * fs, 2))).astype(np.int16)
    a = opus_b200.OpusEncoderBatch(n, channels=2, bitrate=64000, vbr=True, complexity=10)
    b = opus_b200.OpusEncoderBatch(n, channels=2, bitrate=64000, vbr=True, complexity=10)
    want = [a.encode(np.ascontiguousarray(x[:, k * fs:(k + 1) * fs].reshape(n, -1)), 1276) for k in range(nf)]
    pcm = [np.ascontiguousarray(x[:, k * fs:(k + 1) * fs].reshape(n, -1)) for k in range(nf)]
    pk = [np.zeros((n, 1276), np.uint8) for _ in range(nf)]
    nb = [np.zeros(n, np.int32) for _ in range(nf)]
    for k in range(nf):
        b.submit(pcm[k], pk[k], nb[k], 1276, fs)
        if k % 2 == 1:
            b.wait()
    b.wait()
    for k in range(nf):
        assert (nb[k] == want[k][1]).all()
        for s in range(n):
            assert (pk[k][s, :nb[k][s]] == want[k][0][s, :nb[k][s]]).all()
    assert (a.final_range() == b.final_range()).all()
    a.close(); b.close()


@pytest.mark.parametrize('fs_ms,ch,bitrate,vbr,cx,n', [(40, 2, 64000, 1, 10, 70), (60, 2, 24000, 1, 10, 40), (120, 1, 32000, 0, 9, 40), (80, 2, 96000, 0, 5, 40),
                                                         (100, 2, 20000, 1, 8, 40)])
def test_multiframe_packets(fs_ms, ch, bitrate, vbr, cx, n):
    """opus_encode() with 40 .. 120 ms frames: 20 ms frames coded one by one and repacketised (src/opus_encoder.c:1698-1838): packets (code 1 / 2 / 3,
    CBR padding) and final range identical to the reference's"""
    fs = 48 * fs_ms
    run_case(many_signals(70 + fs_ms, fs * 4, ch, k=2), 48000, ch, fs, bitrate, vbr, cx, nframes=4, n=n, max_bytes=1276 * 6)


def test_multiframe_float_api_other_rate_and_pipelined_call():
    import opus_b200
    Fs, fs, n = 24000, 24 * 60, 24
    base = [np.ascontiguousarray(s[::2]).astype(np.float32) / 32768.0 for s in three_signals(88, 2 * fs * 4, 2)]
    base[1][fs + 700:fs + 720] = 3e4
    run_case(base, Fs, 2, fs, 40000, 1, 10, nframes=4, n=n, max_bytes=1276 * 3, app=util.OPUS_APPLICATION_RESTRICTED_LOWDELAY)
    # the pipelined host-buffer call takes the same path
    x = np.stack([base[s % 3][:4 * fs].reshape(-1) for s in range(n)])
    a = opus_b200.OpusEncoderBatch(n, channels=2, fs=Fs, bitrate=40000, vbr=True, complexity=10)
    b = opus_b200.OpusEncoderBatch(n, channels=2, fs=Fs, bitrate=40000, vbr=True, complexity=10)
    for k in range(4):
        fr = np.ascontiguousarray(x[:, k * fs * 2:(k + 1) * fs * 2])
        want, wn = a.encode(fr, 1276 * 3, frame_size=fs)
        pk, nb = np.zeros((n, 1276 * 3), np.uint8), np.zeros(n, np.int32)
        b.submit(fr, pk, nb, 1276 * 3, fs)
        b.wait()
        assert (nb == wn).all()
        for s in range(n):
            assert (pk[s, :nb[s]] == want[s, :wn[s]]).all()
    a.close(); b.close()
