"""Parity of the CUDA CELT encoder (through the C ABI) with the pinned encoder oracle, and encoder -> decoder round trips.
Bar: packet bytes and final range identical to the oracle (hence to the reference's celt_encode_with_ec at every complexity 0-10, CBR / VBR / CVBR)."""
import ctypes

import numpy as np
import pytest

import util
from test_emu_enc import oracle_encode, clicks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('ch,fs,nbytes,complexity,vbr,cvbr', [
    (2, 960, 160, 0, 0, 1), (2, 960, 320, 0, 0, 1), (1, 960, 80, 0, 0, 1), (2, 480, 80, 0, 0, 1), (2, 240, 40, 0, 0, 1), (1, 120, 20, 0, 0, 1),
    (2, 960, 30, 0, 0, 1), (2, 960, 160, 1, 0, 1), (2, 960, 160, 2, 0, 1), (2, 960, 200, 3, 0, 1), (2, 960, 160, 4, 0, 1), (1, 480, 80, 4, 0, 1),
    (2, 240, 50, 3, 0, 1), (2, 120, 30, 4, 0, 1), (2, 960, 1275, 4, 1, 1), (2, 960, 1275, 2, 1, 0), (1, 480, 1275, 4, 1, 1), (2, 120, 1275, 0, 1, 1),
    (2, 960, 160, 5, 0, 1), (1, 960, 80, 5, 0, 1), (2, 480, 100, 6, 0, 1), (2, 240, 60, 7, 0, 1), (1, 120, 30, 5, 0, 1), (2, 960, 60, 7, 0, 1),
    (2, 960, 1275, 5, 1, 1), (1, 480, 1275, 7, 1, 0), (2, 960, 1275, 6, 1, 0),
    (2, 960, 160, 8, 0, 1), (2, 960, 320, 10, 0, 1), (2, 480, 100, 10, 0, 1), (2, 240, 60, 9, 0, 1), (2, 120, 40, 8, 0, 1), (1, 960, 80, 10, 0, 1),
    (2, 960, 50, 8, 0, 1), (2, 960, 1275, 10, 1, 1), (2, 480, 1275, 8, 1, 0)])
def test_encoder_matches_oracle(ch, fs, nbytes, complexity, vbr, cvbr):
    import opus_b200
    O = util.load_oracle()
    n, nf, base = 257, 6, 12
    sigs = [util.sig_music, util.sig_white, clicks] if complexity < 5 else [util.sig_harmonic, util.sig_swell, util.sig_music, clicks]
    bitrate = 64000 * ch // 2 * (2 if fs < 480 else 1) if vbr else -1
    xb = np.stack([sigs[s % len(sigs)](500 + s + fs + nbytes, fs * nf, ch).astype(np.float32).reshape(nf, fs * ch) / 32768.0 for s in range(base)])
    xb[3, 2] = 0
    x = np.tile(xb, (n // base + 1, 1, 1))[:n]
    want = [oracle_encode(O, xb[s], ch, fs, nbytes, bitrate=bitrate, complexity=complexity, vbr=vbr, cvbr=cvbr) for s in range(base)]
    enc = opus_b200.EncoderBatch(n, channels=ch, complexity=complexity, bitrate=bitrate if vbr else None, vbr=bool(vbr), constrained_vbr=bool(cvbr))
    for f in range(nf):
        pk, nb = enc.encode(x[:, f], nbytes)
        rng = enc.final_range()
        for s in range(n):
            wn, wb, wr = want[s % base][f]
            assert nb[s] == wn, (f, s, nb[s], wn)
            if wn > 0:
                assert np.array_equal(pk[s, :wn], wb), (f, s)
                assert rng[s] == wr
    enc.close()


def test_encode_decode_round_trip():
    """packets from the device encoder decode on the device decoder: same final range on both sides, sane SNR"""
    import opus_b200
    ch, fs, nbytes, n, nf = 2, 960, 200, 300, 8
    x = np.stack([(util.sig_harmonic if s & 1 else util.sig_music)(40 + s, fs * nf, ch).astype(np.float32).reshape(nf, fs * ch) / 32768.0 for s in range(16)])
    x = np.tile(x, (n // 16 + 1, 1, 1))[:n]
    enc = opus_b200.EncoderBatch(n, channels=ch, complexity=10)
    dec = opus_b200.DecoderBatch(n, channels=ch)
    toc = 0xFC | 0x04   # CELT-only fullband 20 ms, stereo, one frame (RFC 6716 section 3.1: config 31, s=1, c=0)
    outs = []
    for f in range(nf):
        pk, nb = enc.encode(x[:, f], nbytes)
        assert (nb == nbytes).all()
        full = np.zeros((n, nbytes + 1), np.uint8)
        full[:, 0] = toc
        full[:, 1:] = pk
        pcm, ns = dec.decode(full, np.full(n, nbytes + 1, np.int32), frame_size=fs, dtype=np.float32)
        assert (ns == fs).all()
        assert (dec.final_range() == enc.final_range()).all()
        outs.append(pcm)
    # the codec delay is one overlap (120 samples): compare energy of the error on an aligned middle segment
    y = np.concatenate(outs, axis=1).reshape(n, nf * fs, ch)
    xr = x.reshape(n, nf * fs, ch)
    err = y[:, 960 + 120:5 * 960] - xr[:, 960:5 * 960 - 120]
    snr = 10 * np.log10((xr[:, 960:5 * 960 - 120] ** 2).sum() / (err ** 2).sum())
    assert snr > 8, snr
    enc.close(); dec.close()


def test_ctl_errors():
    import opus_b200
    from opus_b200 import encoder as E
    enc = opus_b200.EncoderBatch(4, channels=2)
    with pytest.raises(opus_b200.OpusB200Error):
        enc.ctl(E.OPUS_SET_COMPLEXITY_REQUEST, 11)    # OPUS_BAD_ARG like the reference (0..10)
    with pytest.raises(opus_b200.OpusB200Error):
        enc.ctl(E.OPUS_SET_BITRATE_REQUEST, 10)
    enc.close()


@pytest.mark.parametrize('n,ch,fs,nbytes,complexity', [(1, 2, 960, 160, 10), (33, 1, 480, 9, 5), (5, 2, 960, 1275, 10), (3, 2, 120, 8, 8), (2, 2, 960, 2, 5),
                                                     (7, 1, 240, 1275, 7)])
def test_encoder_edge_sizes(n, ch, fs, nbytes, complexity):
    """single-stream batches, the smallest and the largest CBR packets"""
    import opus_b200
    O = util.load_oracle()
    nf = 5
    x = np.stack([util.sig_harmonic(900 + s + nbytes, fs * nf, ch).astype(np.float32).reshape(nf, fs * ch) / 32768.0 for s in range(n)])
    want = [oracle_encode(O, x[s], ch, fs, nbytes, bitrate=-1, complexity=complexity, vbr=0, cvbr=1) for s in range(n)]
    enc = opus_b200.EncoderBatch(n, channels=ch, complexity=complexity)
    for f in range(nf):
        pk, nb = enc.encode(x[:, f], nbytes)
        rng = enc.final_range()
        for s in range(n):
            wn, wb, wr = want[s][f]
            assert nb[s] == wn, (f, s, nb[s], wn)
            if wn > 0:
                assert np.array_equal(pk[s, :wn], wb), (f, s)
                assert rng[s] == wr
    enc.close()


@pytest.mark.parametrize('complexity', [5, 10])
def test_encoder_large_batch(complexity):
    """a batch that spans many CTAs of both kernels (and the optional chunked two-stream mode, OPUSB200_ENC_CHUNKS): same packets"""
    import opus_b200
    O = util.load_oracle()
    ch, fs, nbytes, n, nf, base = 2, 960, 160, 9001, 4, 9
    sigs = [util.sig_harmonic, util.sig_swell, util.sig_music]
    xb = np.stack([sigs[s % 3](70 + s, fs * nf, ch).astype(np.float32).reshape(nf, fs * ch) / 32768.0 for s in range(base)])
    x = np.tile(xb, (n // base + 1, 1, 1))[:n]
    want = [oracle_encode(O, xb[s], ch, fs, nbytes, bitrate=-1, complexity=complexity, vbr=0, cvbr=1) for s in range(base)]
    enc = opus_b200.EncoderBatch(n, channels=ch, complexity=complexity)
    for f in range(nf):
        pk, nb = enc.encode(x[:, f], nbytes)
        rng = enc.final_range()
        assert (nb == nbytes).all()
        for s in range(base):
            wn, wb, wr = want[s][f]
            rows = np.arange(s, n, base)
            assert (pk[rows, :wn] == wb[None, :]).all(), (f, s)
            assert (rng[rows] == wr).all()
    enc.close()


@pytest.mark.parametrize('seed', range(16))
def test_encoder_randomised(seed):
    """seeded random settings (channels, frame size, packet size / bit-rate, CBR/VBR/CVBR, complexity 0..10, signal type and level):
    packets and final range equal to the oracle's, i.e. to the reference's"""
    import opus_b200
    O = util.load_oracle()
    rng = np.random.default_rng(7000 + seed)
    ch = int(rng.integers(1, 3))
    fs = int(rng.choice([120, 240, 480, 960]))
    complexity = int(rng.integers(0, 11))
    vbr, cvbr = int(rng.integers(0, 2)), int(rng.integers(0, 2))
    nbytes = 1275 if vbr else int(rng.integers(2, 400))
    bitrate = int(rng.integers(12000, 256000)) * ch if vbr else -1
    base, n, nf = 6, 70, 6
    gens = [util.sig_music, util.sig_white, clicks, util.sig_harmonic, util.sig_swell]
    xb = np.stack([gens[int(rng.integers(0, 5))](seed * 10 + s, fs * nf, ch).astype(np.float32).reshape(nf, fs * ch) / 32768.0 for s in range(base)])
    xb[1] *= float(rng.choice([0.0005, 0.02, 1.0]))
    x = np.tile(xb, (n // base + 1, 1, 1))[:n]
    want = [oracle_encode(O, xb[s], ch, fs, nbytes, bitrate=bitrate, complexity=complexity, vbr=vbr, cvbr=cvbr) for s in range(base)]
    enc = opus_b200.EncoderBatch(n, channels=ch, complexity=complexity, bitrate=bitrate if vbr else None, vbr=bool(vbr), constrained_vbr=bool(cvbr))
    for f in range(nf):
        pk, nb = enc.encode(x[:, f], nbytes)
        rg = enc.final_range()
        for s in range(n):
            wn, wb, wr = want[s % base][f]
            assert nb[s] == wn, (seed, f, s, nb[s], wn, (ch, fs, nbytes, complexity, vbr, cvbr))
            if wn > 0:
                assert np.array_equal(pk[s, :wn], wb), (seed, f, s)
                assert rg[s] == wr
    enc.close()


@pytest.mark.parametrize('end', [13, 17, 19])
@pytest.mark.parametrize('ch,fs,nbytes,complexity,vbr', [(2, 960, 120, 10, 0), (1, 480, 60, 5, 0), (2, 240, 1275, 8, 1)])
def test_encoder_end_band(end, ch, fs, nbytes, complexity, vbr):
    """CELT_SET_END_BAND (request 10012): narrowband / wideband / super-wideband CELT streams, bit-identical to the pinned oracle"""
    import opus_b200
    O = util.load_oracle()
    n, nf = 8, 5
    sigs = [util.sig_harmonic, util.sig_swell, util.sig_music, clicks]
    bitrate = 48000 * ch // 2 * (2 if fs < 480 else 1) if vbr else -1
    x = np.stack([sigs[s % len(sigs)](end + s + fs, fs * nf, ch).astype(np.float32).reshape(nf, fs * ch) / 32768.0 for s in range(n)])
    want = [oracle_encode(O, x[s], ch, fs, nbytes, bitrate=bitrate, complexity=complexity, vbr=vbr, end=end) for s in range(n)]
    enc = opus_b200.EncoderBatch(n, channels=ch, complexity=complexity, bitrate=bitrate if vbr else None, vbr=bool(vbr))
    enc.ctl(10012, end)
    for f in range(nf):
        pk, nb = enc.encode(np.ascontiguousarray(x[:, f]), nbytes)
        rng = enc.final_range()
        for s in range(n):
            wn, wb, wr = want[s][f]
            assert nb[s] == wn, (f, s, nb[s], wn)
            assert np.array_equal(pk[s, :wn], wb), (f, s)
            assert rng[s] == wr
    enc.close()
    with pytest.raises(opus_b200.OpusB200Error):
        opus_b200.EncoderBatch(1, channels=1).ctl(10012, 22)


@pytest.mark.parametrize('Fs,end', [(24000, 19), (16000, 17), (8000, 13)])
@pytest.mark.parametrize('ch,fs48,nbytes,complexity,vbr', [(2, 960, 100, 10, 0), (1, 480, 50, 5, 0), (2, 240, 1275, 8, 1)])
def test_encoder_input_rates(Fs, end, ch, fs48, nbytes, complexity, vbr):
    """opusb200_encoder_create(Fs < 48000) = celt_encoder_init(Fs): frame_size and the pcm are in Fs samples; with the end band the Opus
    layer would pick for that rate.  Bit-identical to the pinned oracle."""
    import opus_b200
    O = util.load_oracle()
    n, nf, up = 8, 5, 48000 // Fs
    fsd = fs48 // up
    sigs = [util.sig_harmonic, util.sig_swell, util.sig_music, clicks]
    bitrate = 40000 * ch // 2 * (2 if fs48 < 480 else 1) if vbr else -1
    x = np.stack([np.ascontiguousarray(sigs[s % len(sigs)](Fs // 1000 + s + fs48, fs48 * nf, ch).reshape(-1, ch)[::up][:fsd * nf]).astype(np.float32)
                  .reshape(nf, fsd * ch) / 32768.0 for s in range(n)])
    want = [oracle_encode(O, x[s], ch, fsd, nbytes, bitrate=bitrate, complexity=complexity, vbr=vbr, end=end, Fs=Fs) for s in range(n)]
    enc = opus_b200.EncoderBatch(n, channels=ch, fs=Fs, complexity=complexity, bitrate=bitrate if vbr else None, vbr=bool(vbr))
    enc.ctl(10012, end)
    for f in range(nf):
        pk, nb = enc.encode(np.ascontiguousarray(x[:, f]), nbytes)
        rng = enc.final_range()
        for s in range(n):
            wn, wb, wr = want[s][f]
            assert nb[s] == wn, (f, s, nb[s], wn)
            assert np.array_equal(pk[s, :wn], wb), (f, s)
            assert rng[s] == wr
    enc.close()


@pytest.mark.parametrize('fs,nbytes,complexity,vbr,end', [(960, 60, 10, 0, 21), (480, 40, 5, 0, 17), (240, 1275, 8, 1, 21), (960, 80, 0, 0, 21)])
def test_encoder_mono_coding_of_stereo_input(fs, nbytes, complexity, vbr, end):
    """CELT_SET_CHANNELS(1) (request 10008) on a stereo encoder: both input channels analysed, their spectra averaged, one channel coded"""
    import opus_b200
    O = util.load_oracle()
    n, nf, ch = 8, 5, 2
    sigs = [util.sig_harmonic, util.sig_swell, util.sig_music, clicks]
    bitrate = 32000 * (2 if fs < 480 else 1) if vbr else -1
    x = np.stack([sigs[s % len(sigs)](end + s + fs, fs * nf, ch).astype(np.float32).reshape(nf, fs * ch) / 32768.0 for s in range(n)])
    want = [oracle_encode(O, x[s], ch, fs, nbytes, bitrate=bitrate, complexity=complexity, vbr=vbr, end=end, coded=1) for s in range(n)]
    enc = opus_b200.EncoderBatch(n, channels=ch, complexity=complexity, bitrate=bitrate if vbr else None, vbr=bool(vbr))
    enc.ctl(10012, end)
    enc.ctl(10008, 1)
    for f in range(nf):
        pk, nb = enc.encode(np.ascontiguousarray(x[:, f]), nbytes)
        rng = enc.final_range()
        for s in range(n):
            wn, wb, wr = want[s][f]
            assert nb[s] == wn, (f, s, nb[s], wn)
            assert np.array_equal(pk[s, :wn], wb), (f, s)
            assert rng[s] == wr
    enc.close()


@pytest.mark.parametrize('prediction', [0, 1])
@pytest.mark.parametrize('ch,fs,nbytes,complexity,vbr', [(2, 960, 100, 10, 0), (1, 480, 50, 5, 0), (2, 240, 1275, 8, 1)])
def test_encoder_prediction_ctl(prediction, ch, fs, nbytes, complexity, vbr):
    """CELT_SET_PREDICTION (request 10002): 0 = intra-coded energies only and no pitch pre-filter, 1 = no pre-filter"""
    import opus_b200
    O = util.load_oracle()
    n, nf = 8, 5
    sigs = [util.sig_harmonic, util.sig_swell, util.sig_music, clicks]
    bitrate = 48000 * ch // 2 * (2 if fs < 480 else 1) if vbr else -1
    x = np.stack([sigs[s % len(sigs)](prediction + s + fs, fs * nf, ch).astype(np.float32).reshape(nf, fs * ch) / 32768.0 for s in range(n)])
    want = [oracle_encode(O, x[s], ch, fs, nbytes, bitrate=bitrate, complexity=complexity, vbr=vbr, prediction=prediction) for s in range(n)]
    enc = opus_b200.EncoderBatch(n, channels=ch, complexity=complexity, bitrate=bitrate if vbr else None, vbr=bool(vbr))
    enc.ctl(10002, prediction)
    for f in range(nf):
        pk, nb = enc.encode(np.ascontiguousarray(x[:, f]), nbytes)
        rng = enc.final_range()
        for s in range(n):
            wn, wb, wr = want[s][f]
            assert nb[s] == wn, (f, s, nb[s], wn)
            assert np.array_equal(pk[s, :wn], wb), (f, s)
            assert rng[s] == wr
    enc.close()
