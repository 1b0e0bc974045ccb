"""The CELT encoder kernels (analysis block + thread-per-stream frame coder) on the host SIMT emulator against the pinned
encoder oracle: packet bytes and final range identical, frame after frame."""
import ctypes

import numpy as np
import pytest

import emu_util
import util


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def oracle_encode(O, x, ch, fs, nbytes, bitrate=-1, lsb=24, complexity=0, vbr=0, cvbr=1, end=21, Fs=48000, coded=None, prediction=2):
    """x [nframes, fs*ch] float -> list of (packet bytes, range)"""
    O.orc_encoder_init.restype = None
    O.orc_encoder_ctl.restype = None
    O.orc_encoder_final_range.restype = ctypes.c_uint32
    st = np.zeros(O.orc_encoder_size(), np.uint8)
    O.orc_encoder_init(_p(st), ch)
    O.orc_encoder_ctl(_p(st), complexity, bitrate, vbr, lsb)
    O.orc_encoder_set_cvbr.restype = None
    O.orc_encoder_set_cvbr(_p(st), cvbr)
    assert O.orc_encoder_set_end_band(_p(st), end) == 0 and O.orc_encoder_set_fs(_p(st), Fs) == 0
    assert O.orc_encoder_set_channels(_p(st), coded or ch) == 0 and O.orc_encoder_set_prediction(_p(st), prediction) == 0
    out = []
    for f in range(len(x)):
        b = np.zeros(1275, np.uint8)
        nb = O.orc_celt_encode_frame(_p(st), _p(x[f]), fs, _p(b), nbytes)
        out.append((nb, b[:max(nb, 0)].copy(), O.orc_encoder_final_range(_p(st))))
    return out


def clicks(seed, n, ch):
    rng = np.random.default_rng(seed)
    x = util.sig_music(seed, n, ch).astype(np.float64).reshape(n, ch)
    w = min(300, n // 4)
    for pos in rng.integers(w, n - 2 * w, size=max(2, n // 2500)):
        x[pos:pos + w] += 12000 * np.exp(-np.arange(w) / 40.0)[:, None] * rng.standard_normal((w, ch))
    return np.clip(x, -32768, 32767).reshape(-1)


@pytest.mark.parametrize('ch,fs,nbytes,complexity,vbr,cvbr', [
    (2, 960, 160, 0, 0, 1), (1, 960, 80, 0, 0, 1), (2, 480, 60, 0, 0, 1), (2, 240, 40, 0, 0, 1), (1, 120, 20, 0, 0, 1), (2, 960, 40, 0, 0, 1),
    (2, 960, 160, 1, 0, 1), (2, 960, 160, 2, 0, 1), (2, 960, 200, 3, 0, 1), (2, 960, 160, 4, 0, 1), (1, 480, 80, 4, 0, 1), (2, 240, 50, 3, 0, 1),
    (2, 960, 1275, 4, 1, 1), (2, 960, 1275, 2, 1, 0), (1, 480, 1275, 4, 1, 1), (2, 120, 1275, 0, 1, 1),
    (2, 960, 160, 5, 0, 1), (1, 960, 80, 5, 0, 1), (2, 480, 100, 6, 0, 1), (2, 240, 60, 7, 0, 1), (1, 120, 30, 5, 0, 1), (2, 960, 1275, 5, 1, 1),
    (1, 480, 1275, 7, 1, 0),
    (2, 960, 160, 8, 0, 1), (2, 480, 100, 10, 0, 1), (2, 240, 60, 9, 0, 1), (2, 120, 40, 8, 0, 1), (1, 960, 80, 10, 0, 1), (2, 960, 1275, 10, 1, 1),
    (2, 960, 50, 8, 0, 1), (2, 480, 1275, 8, 1, 0)])
def test_emu_encoder_matches_oracle(ch, fs, nbytes, complexity, vbr, cvbr):
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    n, nf = (5, 4) if complexity < 5 else (6, 5)
    LM = {120: 0, 240: 1, 480: 2, 960: 3}[fs]
    sigs = [util.sig_music, util.sig_white, clicks] if complexity < 5 else [util.sig_harmonic, util.sig_swell, util.sig_music, clicks]
    bitrate = 64000 * ch // 2 * (2 if fs < 480 else 1) if vbr else -1
    x = np.stack([sigs[s % len(sigs)](77 + s + fs, fs * nf, ch).astype(np.float32).reshape(nf, fs * ch) / 32768.0 for s in range(n)])
    x[4, 2] = 0          # one silent frame in the middle of a stream
    want = [oracle_encode(O, x[s], ch, fs, nbytes, bitrate=bitrate, complexity=complexity, vbr=vbr, cvbr=cvbr) for s in range(n)]
    an = np.zeros(n * E.emu_enc_analysis_state_size(), np.uint8)
    es = np.zeros(n * E.emu_enc_state_size(), np.uint8)
    E.emu_enc_state_init(_p(es), n)
    pmem = np.zeros((n, ch, 1024), np.float32)
    cov = np.zeros(3, np.int32)
    E.emu_enc_coverage(_p(cov), 1)
    for f in range(nf):
        pcm = np.ascontiguousarray(x[:, f])
        pk = np.zeros((n, nbytes), np.uint8)
        nb = np.zeros(n, np.int32)
        rng = np.zeros(n, np.uint32)
        E.emu_celt_encode(_p(an), _p(es), _p(pcm), n, ch, LM, nbytes, bitrate, 24, _p(pk), _p(nb), _p(rng), complexity, vbr, cvbr, _p(pmem))
        for s in range(n):
            wn, wb, wr = want[s][f]
            assert nb[s] == wn, (f, s, nb[s], wn)
            assert np.array_equal(pk[s, :wn], wb), (f, s, int(np.argmax(pk[s, :wn] != wb)))
            assert rng[s] == wr
    E.emu_enc_coverage(_p(cov), 1)
    print('coverage pf_on/patched/transient', cov)
    if complexity >= 5 and fs >= 480:
        assert cov[0] > 0      # the pitch pre-filter switched on somewhere


@pytest.mark.parametrize('seed', range(24))
def test_emu_encoder_randomised(seed):
    """the device encoder's source, run lane by lane on the host emulator, against the pinned oracle over seeded random settings
    (channels, frame size, packet size / bit-rate, rate mode, complexity 0..10, signal type and level)"""
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    rng = np.random.default_rng(5000 + seed)
    ch = int(rng.integers(1, 3))
    fs = int(rng.choice([120, 240, 480, 960]))
    LM = {120: 0, 240: 1, 480: 2, 960: 3}[fs]
    complexity = int(rng.integers(0, 11))
    vbr, cvbr = int(rng.integers(0, 2)), int(rng.integers(0, 2))
    nbytes = 1275 if vbr else int(rng.integers(2, 400))
    bitrate = int(rng.integers(12000, 256000)) * ch if vbr else -1
    n, nf = 3, 6
    gens = [util.sig_music, util.sig_white, clicks, util.sig_harmonic, util.sig_swell]
    x = np.stack([gens[int(rng.integers(0, 5))](seed * 10 + s, fs * nf, ch).astype(np.float32).reshape(nf, fs * ch) / 32768.0 for s in range(n)])
    x[1] *= float(rng.choice([0.0005, 0.02, 1.0]))
    want = [oracle_encode(O, x[s], ch, fs, nbytes, bitrate=bitrate, complexity=complexity, vbr=vbr, cvbr=cvbr) for s in range(n)]
    an = np.zeros(n * E.emu_enc_analysis_state_size(), np.uint8)
    es = np.zeros(n * E.emu_enc_state_size(), np.uint8)
    E.emu_enc_state_init(_p(es), n)
    pmem = np.zeros((n, ch, 1024), np.float32)
    for f in range(nf):
        pcm = np.ascontiguousarray(x[:, f])
        pk = np.zeros((n, nbytes), np.uint8)
        nb = np.zeros(n, np.int32)
        rg = np.zeros(n, np.uint32)
        E.emu_celt_encode(_p(an), _p(es), _p(pcm), n, ch, LM, nbytes, bitrate, 24, _p(pk), _p(nb), _p(rg), complexity, vbr, cvbr, _p(pmem))
        for s in range(n):
            wn, wb, wr = want[s][f]
            assert nb[s] == wn, (seed, f, s, nb[s], wn, (ch, fs, nbytes, complexity, vbr, cvbr))
            assert np.array_equal(pk[s, :wn], wb), (seed, f, s, (ch, fs, nbytes, complexity, vbr, cvbr))
            assert rg[s] == wr


@pytest.mark.parametrize('end', [13, 17, 19])
@pytest.mark.parametrize('ch,fs,nbytes,complexity,vbr', [(2, 960, 120, 10, 0), (1, 480, 60, 5, 0), (2, 240, 1275, 8, 1), (2, 120, 30, 2, 0), (1, 960, 1275, 4, 1)])
def test_emu_encoder_end_band(end, ch, fs, nbytes, complexity, vbr):
    """CELT_SET_END_BAND (celt/celt_encoder.c:3046-3053): narrowband / wideband / super-wideband CELT streams (end = 13 / 17 / 19)"""
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    n, nf = 4, 5
    LM = {120: 0, 240: 1, 480: 2, 960: 3}[fs]
    sigs = [util.sig_harmonic, util.sig_swell, util.sig_music, clicks]
    bitrate = 48000 * ch // 2 * (2 if fs < 480 else 1) if vbr else -1
    x = np.stack([sigs[s % len(sigs)](end + s + fs, fs * nf, ch).astype(np.float32).reshape(nf, fs * ch) / 32768.0 for s in range(n)])
    want = [oracle_encode(O, x[s], ch, fs, nbytes, bitrate=bitrate, complexity=complexity, vbr=vbr, end=end) for s in range(n)]
    an = np.zeros(n * E.emu_enc_analysis_state_size(), np.uint8)
    es = np.zeros(n * E.emu_enc_state_size(), np.uint8)
    E.emu_enc_state_init(_p(es), n)
    pmem = np.zeros((n, ch, 1024), np.float32)
    E.emu_celt_set_end_band(end)
    try:
        for f in range(nf):
            pcm = np.ascontiguousarray(x[:, f])
            pk = np.zeros((n, nbytes), np.uint8)
            nb = np.zeros(n, np.int32)
            rng = np.zeros(n, np.uint32)
            E.emu_celt_encode(_p(an), _p(es), _p(pcm), n, ch, LM, nbytes, bitrate, 24, _p(pk), _p(nb), _p(rng), complexity, vbr, 1, _p(pmem))
            for s in range(n):
                wn, wb, wr = want[s][f]
                assert nb[s] == wn, (f, s, nb[s], wn)
                assert np.array_equal(pk[s, :wn], wb), (f, s, int(np.argmax(pk[s, :wn] != wb)))
                assert rng[s] == wr
    finally:
        E.emu_celt_set_end_band(21)


@pytest.mark.parametrize('Fs,end', [(24000, 19), (16000, 17), (12000, 17), (8000, 13)])
@pytest.mark.parametrize('ch,fs48,nbytes,complexity,vbr', [(2, 960, 100, 10, 0), (1, 480, 50, 5, 0), (2, 240, 1275, 8, 1), (2, 120, 25, 2, 0)])
def test_emu_encoder_input_rates(Fs, end, ch, fs48, nbytes, complexity, vbr):
    """celt_encoder_init(Fs < 48000): the input is zero-stuffed to 48 kHz in the pre-emphasis and the spectrum above Fs/2 is cleared
    (celt/celt_encoder.c:255, 583-589, 544-556); with the end band the Opus layer would pick for that rate"""
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    n, nf, up = 4, 5, 48000 // Fs
    fsd = fs48 // up
    LM = {120: 0, 240: 1, 480: 2, 960: 3}[fs48]
    sigs = [util.sig_harmonic, util.sig_swell, util.sig_music, clicks]
    bitrate = 40000 * ch // 2 * (2 if fs48 < 480 else 1) if vbr else -1
    x = np.stack([np.ascontiguousarray(sigs[s % len(sigs)](Fs // 1000 + s + fs48, fs48 * nf, ch).reshape(-1, ch)[::up][:fsd * nf]).astype(np.float32)
                  .reshape(nf, fsd * ch) / 32768.0 for s in range(n)])
    want = [oracle_encode(O, x[s], ch, fsd, nbytes, bitrate=bitrate, complexity=complexity, vbr=vbr, end=end, Fs=Fs) for s in range(n)]
    an = np.zeros(n * E.emu_enc_analysis_state_size(), np.uint8)
    es = np.zeros(n * E.emu_enc_state_size(), np.uint8)
    E.emu_enc_state_init(_p(es), n)
    pmem = np.zeros((n, ch, 1024), np.float32)
    E.emu_celt_set_end_band(end)
    E.emu_celt_set_upsample(up)
    try:
        for f in range(nf):
            pcm = np.ascontiguousarray(x[:, f])
            pk = np.zeros((n, nbytes), np.uint8)
            nb = np.zeros(n, np.int32)
            rng = np.zeros(n, np.uint32)
            E.emu_celt_encode(_p(an), _p(es), _p(pcm), n, ch, LM, nbytes, bitrate, 24, _p(pk), _p(nb), _p(rng), complexity, vbr, 1, _p(pmem))
            for s in range(n):
                wn, wb, wr = want[s][f]
                assert nb[s] == wn, (f, s, nb[s], wn)
                assert np.array_equal(pk[s, :wn], wb), (f, s, int(np.argmax(pk[s, :wn] != wb)))
                assert rng[s] == wr
    finally:
        E.emu_celt_set_end_band(21)
        E.emu_celt_set_upsample(1)


@pytest.mark.parametrize('fs,nbytes,complexity,vbr,end', [(960, 60, 10, 0, 21), (480, 40, 5, 0, 17), (240, 1275, 8, 1, 21), (120, 20, 2, 0, 21), (960, 1275, 4, 1, 19),
                                                          (960, 80, 0, 0, 21), (480, 50, 9, 0, 21)])
def test_emu_encoder_mono_coding_of_stereo_input(fs, nbytes, complexity, vbr, end):
    """CELT_SET_CHANNELS(1) on a stereo encoder (what the Opus layer does at low rates, src/opus_encoder.c:2285): analysis and pre-filter on
    both input channels, the two spectra averaged (celt/celt_encoder.c:539-543), everything after that mono"""
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    n, nf, ch = 4, 5, 2
    LM = {120: 0, 240: 1, 480: 2, 960: 3}[fs]
    sigs = [util.sig_harmonic, util.sig_swell, util.sig_music, clicks]
    bitrate = 32000 * (2 if fs < 480 else 1) if vbr else -1
    x = np.stack([sigs[s % len(sigs)](end + s + fs, fs * nf, ch).astype(np.float32).reshape(nf, fs * ch) / 32768.0 for s in range(n)])
    want = [oracle_encode(O, x[s], ch, fs, nbytes, bitrate=bitrate, complexity=complexity, vbr=vbr, end=end, coded=1) for s in range(n)]
    an = np.zeros(n * E.emu_enc_analysis_state_size(), np.uint8)
    es = np.zeros(n * E.emu_enc_state_size(), np.uint8)
    E.emu_enc_state_init(_p(es), n)
    pmem = np.zeros((n, ch, 1024), np.float32)
    E.emu_celt_set_end_band(end)
    E.emu_celt_set_channels(1)
    try:
        for f in range(nf):
            pcm = np.ascontiguousarray(x[:, f])
            pk = np.zeros((n, nbytes), np.uint8)
            nb = np.zeros(n, np.int32)
            rng = np.zeros(n, np.uint32)
            E.emu_celt_encode(_p(an), _p(es), _p(pcm), n, ch, LM, nbytes, bitrate, 24, _p(pk), _p(nb), _p(rng), complexity, vbr, 1, _p(pmem))
            for s in range(n):
                wn, wb, wr = want[s][f]
                assert nb[s] == wn, (f, s, nb[s], wn)
                assert np.array_equal(pk[s, :wn], wb), (f, s, int(np.argmax(pk[s, :wn] != wb)))
                assert rng[s] == wr
    finally:
        E.emu_celt_set_end_band(21)
        E.emu_celt_set_channels(0)


@pytest.mark.parametrize('prediction', [0, 1])
@pytest.mark.parametrize('ch,fs,nbytes,complexity,vbr', [(2, 960, 100, 10, 0), (1, 480, 50, 5, 0), (2, 240, 1275, 8, 1), (2, 960, 120, 4, 0)])
def test_emu_encoder_prediction_ctl(prediction, ch, fs, nbytes, complexity, vbr):
    """CELT_SET_PREDICTION (celt/celt_encoder.c:2974-2982): 0 = intra-coded energies only and no pitch pre-filter, 1 = no pre-filter"""
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    n, nf = 4, 5
    LM = {120: 0, 240: 1, 480: 2, 960: 3}[fs]
    sigs = [util.sig_harmonic, util.sig_swell, util.sig_music, clicks]
    bitrate = 48000 * ch // 2 * (2 if fs < 480 else 1) if vbr else -1
    x = np.stack([sigs[s % len(sigs)](prediction + s + fs, fs * nf, ch).astype(np.float32).reshape(nf, fs * ch) / 32768.0 for s in range(n)])
    want = [oracle_encode(O, x[s], ch, fs, nbytes, bitrate=bitrate, complexity=complexity, vbr=vbr, prediction=prediction) for s in range(n)]
    an = np.zeros(n * E.emu_enc_analysis_state_size(), np.uint8)
    es = np.zeros(n * E.emu_enc_state_size(), np.uint8)
    E.emu_enc_state_init(_p(es), n)
    pmem = np.zeros((n, ch, 1024), np.float32)
    E.emu_celt_set_prediction(prediction)
    try:
        for f in range(nf):
            pcm = np.ascontiguousarray(x[:, f])
            pk = np.zeros((n, nbytes), np.uint8)
            nb = np.zeros(n, np.int32)
            rng = np.zeros(n, np.uint32)
            E.emu_celt_encode(_p(an), _p(es), _p(pcm), n, ch, LM, nbytes, bitrate, 24, _p(pk), _p(nb), _p(rng), complexity, vbr, 1, _p(pmem))
            for s in range(n):
                wn, wb, wr = want[s][f]
                assert nb[s] == wn, (f, s, nb[s], wn)
                assert np.array_equal(pk[s, :wn], wb), (f, s, int(np.argmax(pk[s, :wn] != wb)))
                assert rng[s] == wr
    finally:
        E.emu_celt_set_prediction(2)
