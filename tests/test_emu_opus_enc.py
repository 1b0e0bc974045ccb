"""The Opus-layer encoder's DEVICE SOURCE (opus_b200/csrc/opus_enc_layer.cuh + the CELT encoder kernels) on the host SIMT emulator
against the oracle's opus_encode restatement (itself pinned to the reference's opus_encode by test_opus_enc_oracle_vs_ref.py):
AnalysisInfo, packets (TOC included) and final range identical, frame after frame, several streams per batch."""
import ctypes

import numpy as np
import pytest

import emu_util
import util

APP_CELT = util.OPUS_APPLICATION_RESTRICTED_CELT


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class EmuOpusEncoder:
    def __init__(self, n, Fs, ch, app=APP_CELT):
        self.L = emu_util.load_emu_ops()
        self.L.emu_opus_encoder_create.restype = ctypes.c_void_p
        self.h = ctypes.c_void_p(self.L.emu_opus_encoder_create(n, Fs, ch, app))
        self.n, self.ch = n, ch

    def ctl(self, req, val):
        assert self.L.emu_opus_encoder_ctl(self.h, req, val) == 0, (req, val)

    def encode(self, pcm, frame_size, max_bytes=1276):
        pcm = np.ascontiguousarray(pcm)
        is16 = int(pcm.dtype == np.int16)
        out = np.zeros((self.n, max_bytes), np.uint8)
        nb = np.zeros(self.n, np.int32)
        rng = np.zeros(self.n, np.uint32)
        r = self.L.emu_opus_encode(self.h, _p(pcm), is16, frame_size, _p(out), max_bytes, max_bytes, _p(nb), _p(rng))
        assert r == 0, r
        return out, nb, rng

    def last_analysis(self):
        a = np.zeros((self.n, 30), np.float32)
        self.L.emu_opus_last_analysis(self.h, _p(a))
        return a

    def close(self):
        self.L.emu_opus_encoder_destroy(self.h)


def oracle_stream(pcm, Fs, ch, fs, bitrate, vbr, cx, ctls=(), app=APP_CELT, max_bytes=1276):
    O = util.load_oracle()
    T = pcm.shape[0] // fs
    c = np.array([v for p in ctls for v in p], dtype=np.int32)
    out = np.zeros((T, max_bytes), np.uint8)
    sizes, rngs = np.zeros(T, np.int32), np.zeros(T, np.uint32)
    r = O.orc_opus_encode_stream(Fs, ch, app, int(bitrate), int(vbr), int(cx), fs, T, _p(pcm), int(pcm.dtype == np.int16), _p(out), max_bytes,
                                 _p(sizes), _p(rngs), _p(c), len(ctls))
    assert r == 0, r
    return out, sizes, rngs


def run_case(sigs, Fs, ch, fs, bitrate, vbr, cx, ctls=(), nframes=10, max_bytes=1276, check_analysis=False):
    """sigs: list of [T*fs, ch] arrays (one per stream, same dtype)"""
    n = len(sigs)
    enc = EmuOpusEncoder(n, Fs, ch)
    enc.ctl(4002, bitrate); enc.ctl(4006, vbr); enc.ctl(4010, cx)
    for r, v in ctls:
        enc.ctl(r, v)
    want = [oracle_stream(np.ascontiguousarray(x[:nframes * fs]), Fs, ch, fs, bitrate, vbr, cx, ctls, max_bytes=max_bytes) for x in sigs]
    ana = None
    if check_analysis:
        O = util.load_oracle()
        ana = []
        for x in sigs:
            a = np.zeros((nframes, 30), np.float32)
            xx = np.ascontiguousarray(x[:nframes * fs])
            assert O.orc_run_analysis_stream(Fs, ch, fs, nframes, _p(xx), int(xx.dtype == np.int16), 16 if xx.dtype == np.int16 else 24, _p(a)) == 0
            ana.append(a)
    for t in range(nframes):
        frame = np.stack([x[t * fs:(t + 1) * fs].reshape(-1) for x in sigs])
        out, nb, rng = enc.encode(frame, fs, max_bytes)
        if ana is not None:
            got = enc.last_analysis()
            for s in range(n):
                assert np.array_equal(got[s].view(np.uint32), ana[s][t].view(np.uint32)), (t, s, got[s][:11], ana[s][t][:11])
        for s in range(n):
            w, ws, wr = want[s]
            assert nb[s] == ws[t], (t, s, nb[s], ws[t])
            assert np.array_equal(out[s, :nb[s]], w[t, :ws[t]]), (t, s, int(np.argmax(out[s, :nb[s]] != w[t, :ws[t]])), nb[s])
            assert rng[s] == wr[t], (t, s)
    enc.close()


def three_signals(seed, nsamp, ch):
    return [util.sig_music(seed, nsamp, ch), util.sig_tonal_noise(seed + 1, nsamp, ch), util.sig_white(seed + 2, nsamp, ch)]


def test_config2_analysis_and_packets():
    """BASELINE.json configs[2]: 48 kHz stereo 20 ms, 64 kb/s VBR, complexity 10 — analysis info and packets vs the oracle"""
    run_case(three_signals(11, 960 * 12, 2), 48000, 2, 960, 64000, 1, 10, nframes=12, check_analysis=True)


@pytest.mark.parametrize('cx', [7, 5, 0])
def test_complexities(cx):
    run_case(three_signals(5, 960 * 6, 2), 48000, 2, 960, 64000, 1, cx, nframes=6)


@pytest.mark.parametrize('ch,bitrate,vbr', [(2, 16000, 1), (2, 24000, 0), (2, 32000, 1), (1, 24000, 1), (2, 128000, 0), (1, 64000, 0)])
def test_rates(ch, bitrate, vbr):
    """low stereo rates: mono-coded streams, narrower bands, the stereo width fade"""
    run_case(three_signals(21 + ch, 960 * 8, ch), 48000, ch, 960, bitrate, vbr, 10, nframes=8)


@pytest.mark.parametrize('fs', [480, 240, 120])
def test_short_frames(fs):
    run_case(three_signals(fs, fs * 24, 2), 48000, 2, fs, 64000, 1, 10, nframes=24, check_analysis=True)


@pytest.mark.parametrize('Fs', [24000, 16000, 8000])
def test_input_rates(Fs):
    sigs = [np.ascontiguousarray(x[::48000 // Fs]) for x in three_signals(Fs // 1000, 960 * 8, 2)]
    run_case(sigs, Fs, 2, Fs // 50, 32000, 1, 10, nframes=8, check_analysis=Fs >= 16000)


def test_float_api_silence_and_guard():
    x = util.sig_music(9, 960 * 10, 2).astype(np.float32) / 32768.0
    y = x.copy()
    y[960 * 3:960 * 5] = 0          # digital silence: the analysis repeats its last result
    z = x.copy()
    z[960 * 4:960 * 5] *= 1e6       # absurd level: the frame is cleared by the float guard
    z[960 * 7 + 5] = np.nan
    run_case([x, y, z], 48000, 2, 960, 64000, 1, 10, nframes=10, check_analysis=True)


@pytest.mark.parametrize('ctl', [[(4008, 1103)], [(4022, 1)], [(4024, 3001)], [(4042, 1)], [(4036, 12)], [(4020, 0)]])
def test_ctls(ctl):
    run_case(three_signals(31, 960 * 6, 2), 48000, 2, 960, 40000, 1, 10, ctls=ctl, nframes=6)


def test_frame_coder_with_a_barrier_per_band():
    """OPUSB200_ENC_SYNC: blocks of several warps that meet before every band (shared instruction-cache lines); 40 streams = 2 warps in one
    block, streams of different end bands (a narrower bandwidth on some) and a mono-coded low rate in the same run"""
    L = emu_util.load_emu_ops()
    L.emu_celt_set_frame_sync(1)
    try:
        sigs = [s for k in range(14) for s in three_signals(50 + 3 * k, 960 * 3, 2)][:40]
        run_case(sigs, 48000, 2, 960, 64000, 1, 10, nframes=3)
        run_case(three_signals(47, 960 * 4, 2), 48000, 2, 960, 20000, 1, 10, ctls=[(4004, 1103)], nframes=4)
    finally:
        L.emu_celt_set_frame_sync(0)


@pytest.mark.parametrize('mode', [0, 2])
def test_other_front_end_forms(mode):
    """0 = OPUSB200_ENC_NO_CHAINS (every recurrence inside the CTA-per-stream kernel), 2 = OPUSB200_ENC_SERIAL_PF (run_prefilter too on the
    thread-per-stream kernel): same packets as the default form"""
    L = emu_util.load_emu_ops()
    L.emu_celt_set_chains(mode)
    try:
        run_case(three_signals(41, 960 * 5, 2), 48000, 2, 960, 64000, 1, 10, nframes=5)
        run_case(three_signals(43, 480 * 6, 1), 48000, 1, 480, 32000, 0, 5, nframes=6)
        run_case([np.ascontiguousarray(x[::2]) for x in three_signals(45, 960 * 6, 2)], 24000, 2, 480, 32000, 1, 8, nframes=6)
    finally:
        L.emu_celt_set_chains(1)


@pytest.mark.parametrize('fs_ms,ch,bitrate,vbr,cx', [(40, 2, 64000, 1, 10), (60, 2, 24000, 1, 10), (120, 1, 32000, 0, 9), (80, 2, 96000, 0, 5), (100, 2, 20000, 1, 8)])
def test_multiframe_packets(fs_ms, ch, bitrate, vbr, cx):
    """40 .. 120 ms: P1 / P2 per 20 ms frame, the packet's decisions, the CELT kernels frame by frame, the repacketizer (code 1 / 2 / 3, CBR padding;
    low stereo rates bring the stereo fade and mono-coded streams into it) against the oracle, which is pinned on opus_encode()"""
    fs = 48 * fs_ms
    nfr = 5
    run_case(three_signals(60 + fs_ms, fs * nfr, ch), 48000, ch, fs, bitrate, vbr, cx, nframes=nfr, max_bytes=1276 * 6)


def test_multiframe_float_api_with_guard_and_other_rate():
    Fs, fs = 24000, 24 * 60
    x = [np.ascontiguousarray(s[::2]).astype(np.float32) / 32768.0 for s in three_signals(77, 2 * fs * 4, 2)]
    x[1][fs + 700:fs + 720] = 3e4     # the float guard fires in one 20 ms frame of the second packet
    x[2][2 * fs:3 * fs] = 0           # a silent packet
    run_case(x, Fs, 2, fs, 40000, 1, 10, nframes=4, max_bytes=1276 * 3)
