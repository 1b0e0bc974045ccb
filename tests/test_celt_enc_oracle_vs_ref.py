"""Pin the CELT encoder restatement (oracle/celt_enc.inc) to the reference's celt_encode_with_ec driven directly
(oracle/_ref): packet bytes and final range identical, frame after frame (encoder state carried)."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built (needs /root/reference)')


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def encode_both(sig, ch, fs, nbytes, complexity, bitrate=-1000, vbr=0, lsb=24, kind='c', nframes=12, cvbr=1, end=21, coded=None, prediction=2):
    R, O = util.load_ref(kind), util.load_oracle()
    x = (sig.astype(np.float32) / 32768.0).reshape(-1, fs * ch)[:nframes]
    rs = np.zeros(R.refh_celt_encoder_size(ch) + 64, np.uint8)
    assert R.refh_celt_encoder_init(_p(rs), ch, complexity, bitrate, vbr, lsb) == 0
    R.refh_celt_encoder_set_cvbr(_p(rs), cvbr)
    assert R.refh_celt_encoder_set_end_band(_p(rs), end) == 0 and R.refh_celt_encoder_set_channels(_p(rs), coded or ch) == 0
    assert R.refh_celt_encoder_set_prediction(_p(rs), prediction) == 0
    os_ = np.zeros(O.orc_encoder_size(), np.uint8)
    O.orc_encoder_init.restype = None
    O.orc_encoder_ctl.restype = None
    O.orc_encoder_final_range.restype = ctypes.c_uint32
    O.orc_encoder_init(_p(os_), ch)
    O.orc_encoder_ctl(_p(os_), complexity, -1 if bitrate == -1000 else bitrate, vbr, lsb)
    O.orc_encoder_set_cvbr.restype = None
    O.orc_encoder_set_cvbr(_p(os_), cvbr)
    assert O.orc_encoder_set_end_band(_p(os_), end) == 0 and O.orc_encoder_set_channels(_p(os_), coded or ch) == 0
    assert O.orc_encoder_set_prediction(_p(os_), prediction) == 0
    out = []
    for f in range(len(x)):
        a, b = np.zeros(1275, np.uint8), np.zeros(1275, np.uint8)
        rng = ctypes.c_uint32(0)
        na = R.refh_celt_encode(_p(rs), _p(x[f]), fs, _p(a), nbytes, ctypes.byref(rng))
        nb = O.orc_celt_encode_frame(_p(os_), _p(x[f]), fs, _p(b), nbytes)
        out.append((f, na, nb, a[:max(na, 0)].copy(), b[:max(nb, 0)].copy(), rng.value, O.orc_encoder_final_range(_p(os_))))
    return out


CASES = [(ch, fs, nbytes, s) for ch in (1, 2) for fs in (960, 480, 240, 120) for nbytes in (40, 80, 160, 320)
         for s in ('music', 'noise')]


@pytest.mark.parametrize('ch,fs,nbytes,sname', CASES)
def test_complexity0_cbr(ch, fs, nbytes, sname):
    nb = max(8, nbytes * fs // 960)
    sig = (util.sig_music if sname == 'music' else util.sig_white)(ch * 100 + fs + nbytes, fs * 12, ch)
    for (f, na, nbo, a, b, ra, rb) in encode_both(sig, ch, fs, nb, 0):
        assert nbo != -5, 'frame needed a decision that is not restated'
        assert na == nbo, (f, na, nbo)
        assert np.array_equal(a, b), (f, int(np.argmax(a != b)))
        assert ra == rb


def _transient_signal(seed, n, ch):
    """music with castanet-like clicks so that transient_analysis fires on some frames"""
    rng = np.random.default_rng(seed)
    x = util.sig_music(seed, n, ch).astype(np.float64).reshape(n, ch)
    for pos in rng.integers(200, n - 400, size=max(2, n // 2500)):
        env = np.exp(-np.arange(300) / 40.0)[:, None]
        x[pos:pos + 300] += 12000 * env * rng.standard_normal((300, ch))
    return np.clip(x, -32768, 32767).reshape(-1)


@pytest.mark.parametrize('complexity', [1, 2, 3, 4])
@pytest.mark.parametrize('ch,fs,nbytes', [(2, 960, 160), (1, 960, 80), (2, 480, 100), (2, 240, 60), (1, 120, 30), (2, 960, 40), (2, 960, 320)])
def test_complexity_1_to_4_cbr(complexity, ch, fs, nbytes):
    for sname in ('music', 'clicks', 'noise'):
        n = fs * 12
        sig = util.sig_music(complexity + ch + fs, n, ch) if sname == 'music' else util.sig_white(complexity + fs, n, ch) if sname == 'noise' \
            else _transient_signal(complexity * 7 + fs + nbytes, n, ch)
        for (f, na, nbo, a, b, ra, rb) in encode_both(sig, ch, fs, nbytes, complexity):
            assert nbo != -5, 'frame needed a decision that is not restated'
            assert na == nbo, (sname, f, na, nbo)
            assert np.array_equal(a, b), (sname, f, int(np.argmax(a != b)))
            assert ra == rb


@pytest.mark.parametrize('complexity', [0, 2, 4])
@pytest.mark.parametrize('cvbr', [1, 0])
@pytest.mark.parametrize('ch,fs,bitrate', [(2, 960, 64000), (1, 960, 32000), (2, 480, 96000), (2, 240, 128000), (1, 120, 64000), (2, 960, 24000)])
def test_vbr(complexity, cvbr, ch, fs, bitrate):
    for sname in ('music', 'clicks'):
        n = fs * 16
        sig = util.sig_music(3 + ch + fs, n, ch) if sname == 'music' else _transient_signal(fs + bitrate // 1000, n, ch)
        sig = np.asarray(sig, np.float64)
        sig[fs * ch * 5:fs * ch * 7] = 0      # two silent frames: the VBR silence path
        sizes = []
        for (f, na, nbo, a, b, ra, rb) in encode_both(sig, ch, fs, 1275, complexity, bitrate=bitrate, vbr=1, cvbr=cvbr, nframes=16):
            assert nbo != -5
            assert na == nbo, (sname, f, na, nbo)
            assert np.array_equal(a, b), (sname, f, int(np.argmax(a != b)))
            assert ra == rb
            sizes.append(na)
        assert len(set(sizes)) > 2     # really variable


def _harmonic_signal(seed, n, ch):
    """voiced-speech-like: a drifting pitch with many harmonics plus a little noise, so that the pitch pre-filter switches on"""
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    f0 = 110 + 60 * np.sin(2 * np.pi * t / 48000 * 0.7 + rng.uniform(0, 6))
    ph = 2 * np.pi * np.cumsum(f0) / 48000
    x = np.zeros(n)
    for h in range(1, 25):
        x += np.sin(h * ph + rng.uniform(0, 6)) / h
    x = 6000 * x + 150 * rng.standard_normal(n)
    if ch == 2:
        x = np.stack([x, 0.7 * np.roll(x, 13) + 100 * rng.standard_normal(n)], axis=1)
    return np.clip(x, -32768, 32767).reshape(-1)


@pytest.mark.parametrize('complexity', [5, 6, 7])
@pytest.mark.parametrize('ch,fs,nbytes,vbr', [(2, 960, 160, 0), (1, 960, 80, 0), (2, 480, 100, 0), (2, 240, 60, 0), (1, 120, 30, 0), (2, 960, 60, 0),
                                              (2, 960, 1275, 1), (1, 480, 1275, 1)])
def test_complexity_5_to_7(complexity, ch, fs, nbytes, vbr):
    pf_frames = 0
    for sname in ('harmonic', 'music', 'clicks'):
        n = fs * 16
        sig = _harmonic_signal(complexity + fs, n, ch) if sname == 'harmonic' else util.sig_music(complexity + ch + fs, n, ch) if sname == 'music' \
            else _transient_signal(complexity * 7 + fs + nbytes, n, ch)
        for (f, na, nbo, a, b, ra, rb) in encode_both(sig, ch, fs, nbytes, complexity, bitrate=(48000 * ch if vbr else -1000), vbr=vbr, nframes=16):
            assert nbo != -5, 'frame needed a decision that is not restated'
            assert na == nbo, (sname, f, na, nbo)
            assert np.array_equal(a, b), (sname, f, int(np.argmax(a != b)))
            assert ra == rb
            if sname == 'harmonic' and len(a) > 1 and (a[0] & 0x40):
                pf_frames += 1      # second header bit set: post-filter on (after a 0 silence flag)
    if fs >= 480 and nbytes >= 80:
        assert pf_frames > 0


@pytest.mark.parametrize('complexity', [8, 10])
@pytest.mark.parametrize('ch,fs,nbytes,vbr', [(2, 960, 160, 0), (1, 960, 80, 0), (2, 480, 100, 0), (2, 240, 60, 0), (2, 120, 30, 0), (2, 960, 60, 0),
                                              (2, 960, 320, 0), (2, 960, 1275, 1), (2, 480, 1275, 1), (1, 480, 1275, 1)])
def test_complexity_8_to_10(complexity, ch, fs, nbytes, vbr):
    """second MDCT (transient frames) + theta RDO with resynthesis (stereo); CELT has no gate above 8"""
    for sname in ('harmonic', 'music', 'clicks', 'white'):
        n = fs * 12
        sig = {'harmonic': lambda: _harmonic_signal(complexity + fs, n, ch), 'music': lambda: util.sig_music(complexity + ch + fs, n, ch),
               'clicks': lambda: _transient_signal(complexity * 7 + fs + nbytes, n, ch), 'white': lambda: util.sig_white(3 + fs, n, ch)}[sname]()
        for (f, na, nbo, a, b, ra, rb) in encode_both(sig, ch, fs, nbytes, complexity, bitrate=(64000 * ch if vbr else -1000), vbr=vbr, nframes=12):
            assert nbo != -5, 'frame needed a decision that is not restated'
            assert na == nbo, (sname, f, na, nbo)
            assert np.array_equal(a, b), (sname, f, int(np.argmax(a != b)))
            assert ra == rb


@pytest.mark.parametrize('ch,fs,nbytes,complexity', [(2, 960, 160, 10), (1, 480, 9, 5), (2, 960, 1275, 10), (2, 120, 8, 8), (2, 960, 2, 5), (1, 240, 1275, 7),
                                                    (2, 960, 3, 10), (1, 120, 2, 0), (2, 480, 1275, 8), (2, 960, 10, 10)])
def test_edge_packet_sizes(ch, fs, nbytes, complexity):
    """the smallest (2 bytes) and the largest (1275 bytes) CBR packets"""
    sig = util.sig_harmonic(900 + nbytes, fs * 6, ch)
    for (f, na, nbo, a, b, ra, rb) in encode_both(sig, ch, fs, nbytes, complexity, bitrate=-1000, vbr=0, nframes=6):
        assert na == nbo and np.array_equal(a, b) and ra == rb, (f, na, nbo)


@pytest.mark.parametrize('seed', range(40))
def test_randomised_settings(seed):
    """seeded random draws over everything the encoder oracle takes: channels, frame size, packet size or bit-rate, rate mode,
    complexity and signal type"""
    rng = np.random.default_rng(1000 + seed)
    ch = int(rng.integers(1, 3))
    fs = int(rng.choice([120, 240, 480, 960]))
    complexity = int(rng.integers(0, 11))
    vbr = int(rng.integers(0, 2))
    cvbr = int(rng.integers(0, 2))
    nbytes = 1275 if vbr else int(rng.integers(2, 400))
    bitrate = int(rng.integers(12000, 256000)) * ch if vbr else -1000
    kind = int(rng.integers(0, 4))
    n = fs * 10
    sig = [lambda: util.sig_music(seed, n, ch), lambda: util.sig_white(seed, n, ch), lambda: _transient_signal(seed, n, ch),
           lambda: _harmonic_signal(seed, n, ch)][kind]()
    if rng.random() < 0.3:
        sig = sig * float(rng.choice([0.0005, 0.01, 0.2]))        # near-silent and quiet inputs
    for (f, na, nbo, a, b, ra, rb) in encode_both(sig, ch, fs, nbytes, complexity, bitrate=bitrate, vbr=vbr, nframes=10, cvbr=cvbr):
        assert na == nbo and np.array_equal(a, b) and ra == rb, (seed, ch, fs, nbytes, complexity, vbr, kind, f, na, nbo)


@pytest.mark.parametrize('end', [13, 17, 19])
@pytest.mark.parametrize('ch,fs,complexity,bitrate,vbr', [(2, 960, 10, 64000, 1), (1, 480, 5, -1000, 0), (2, 240, 8, 32000, 0), (2, 120, 0, -1000, 0),
                                                          (1, 960, 10, 24000, 1), (2, 480, 3, 96000, 1)])
def test_end_band(end, ch, fs, complexity, bitrate, vbr):
    """CELT_SET_END_BAND: NB / WB / SWB streams (end = 13 / 17 / 19), packets and final range bit-identical to the reference"""
    sig = util.sig_music(end + ch + fs, fs * 12, ch)
    nb = 1275 if bitrate != -1000 else max(8, 100 * fs // 960)
    for (f, na, nbo, a, b, ra, rb) in encode_both(sig, ch, fs, nb, complexity, bitrate=bitrate, vbr=vbr, end=end):
        assert na == nbo, (f, na, nbo)
        assert np.array_equal(a, b), (f, int(np.argmax(a != b)))
        assert ra == rb


@pytest.mark.parametrize('Fs,end', [(24000, 19), (16000, 17), (12000, 17), (8000, 13), (16000, 21)])
@pytest.mark.parametrize('ch,fs48,complexity,bitrate,vbr', [(2, 960, 10, 48000, 1), (1, 480, 5, -1000, 0), (2, 240, 8, 32000, 0), (2, 120, 0, -1000, 0), (1, 960, 10, 24000, 1)])
def test_input_rates(Fs, end, ch, fs48, complexity, bitrate, vbr):
    """celt_encoder_init(Fs < 48000) (st->upsample): zero-stuffed pre-emphasis, cleared upper spectrum; packets and final range bit-identical"""
    R, O = util.load_ref('c'), util.load_oracle()
    up, nframes = 48000 // Fs, 10
    fsd = fs48 // up
    sig = util.sig_music(end + ch + fs48, fs48 * nframes, ch).reshape(-1, ch)[::up][:fsd * nframes]
    x = (sig.astype(np.float32) / 32768.0).reshape(nframes, fsd * ch)
    nbytes = 1275 if bitrate != -1000 else max(8, 100 * fs48 // 960)
    rs = np.zeros(R.refh_celt_encoder_size(ch) + 64, np.uint8)
    assert R.refh_celt_encoder_init_fs(_p(rs), Fs, ch, complexity, bitrate, vbr, 24) == 0
    assert R.refh_celt_encoder_set_end_band(_p(rs), end) == 0
    os_ = np.zeros(O.orc_encoder_size(), np.uint8)
    O.orc_encoder_init.restype = None
    O.orc_encoder_ctl.restype = None
    O.orc_encoder_final_range.restype = ctypes.c_uint32
    O.orc_encoder_init(_p(os_), ch)
    O.orc_encoder_ctl(_p(os_), complexity, -1 if bitrate == -1000 else bitrate, vbr, 24)
    assert O.orc_encoder_set_end_band(_p(os_), end) == 0 and O.orc_encoder_set_fs(_p(os_), Fs) == 0
    for f in range(nframes):
        a, b = np.zeros(1275, np.uint8), np.zeros(1275, np.uint8)
        rng = ctypes.c_uint32(0)
        xf = np.ascontiguousarray(x[f])
        na = R.refh_celt_encode(_p(rs), _p(xf), fsd, _p(a), nbytes, ctypes.byref(rng))
        nb = O.orc_celt_encode_frame(_p(os_), _p(xf), fsd, _p(b), nbytes)
        assert na == nb, (f, na, nb)
        assert np.array_equal(a[:na], b[:na]) and rng.value == O.orc_encoder_final_range(_p(os_))


@pytest.mark.parametrize('fs', [960, 480, 240, 120])
@pytest.mark.parametrize('complexity,bitrate,vbr', [(0, -1000, 0), (5, -1000, 0), (10, 32000, 1), (8, 16000, 0), (3, 24000, 1)])
def test_mono_coding_of_stereo_input(fs, complexity, bitrate, vbr):
    """CELT_SET_CHANNELS(1) on a stereo encoder (celt/celt_encoder.c:539-543, 2095-2096, 2780-2782): bit-identical to the reference"""
    sig = util.sig_music(fs + complexity, fs * 12, 2)
    nb = 1275 if bitrate != -1000 else max(8, 60 * fs // 960)
    for (f, na, nbo, a, b, ra, rb) in encode_both(sig, 2, fs, nb, complexity, bitrate=bitrate, vbr=vbr, coded=1):
        assert na == nbo, (f, na, nbo)
        assert np.array_equal(a, b), (f, int(np.argmax(a != b)))
        assert ra == rb


@pytest.mark.parametrize('prediction', [0, 1])
@pytest.mark.parametrize('ch,fs,complexity,bitrate,vbr', [(2, 960, 10, 64000, 1), (1, 480, 5, -1000, 0), (2, 240, 8, 32000, 0), (2, 120, 4, -1000, 0), (1, 960, 10, 24000, 1)])
def test_prediction_ctl(prediction, ch, fs, complexity, bitrate, vbr):
    """CELT_SET_PREDICTION 0 / 1 (force_intra, disable_pf): bit-identical to the reference"""
    sig = util.sig_harmonic(prediction + ch + fs, fs * 12, ch)
    nb = 1275 if bitrate != -1000 else max(8, 100 * fs // 960)
    for (f, na, nbo, a, b, ra, rb) in encode_both(sig, ch, fs, nb, complexity, bitrate=bitrate, vbr=vbr, prediction=prediction):
        assert na == nbo, (f, na, nbo)
        assert np.array_equal(a, b), (f, int(np.argmax(a != b)))
        assert ra == rb
