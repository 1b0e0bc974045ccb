"""The decoder kernel SOURCES (opus_b200/csrc/*.cuh), executed lane-by-lane on the host SIMT
emulator, against the golden vectors: final range bit-exact, PCM within +-1 LSB (16 bit).
This is a CPU-side logic check of the device code; the real parity tests run on the GPU
(tests/test_gpu_decode.py)."""
import numpy as np
import pytest

import emu_util
import util
from test_oracle_golden import golden_cases

CASES = list(golden_cases())


@pytest.mark.parametrize('case', CASES, ids=lambda c: c[0])
def test_emulated_kernels_match_golden(case):
    name, pkts, sizes, rng, pcm, ch, fs = case
    nfr = min(len(sizes), 6)
    got, got_rng = emu_util.emu_decode_stream(pkts[:nfr], sizes[:nfr], ch, fs)
    assert (got_rng == rng[:nfr]).all(), (got_rng, rng[:nfr])
    ref = pcm[:nfr * fs]
    d = np.abs(util.to_int16(got).astype(np.int32) - util.to_int16(ref).astype(np.int32))
    assert d.max() <= 1, d.max()
    # and tight in float: the only differences are float re-association (scan / tree reductions)
    assert np.abs(got - ref).max() < 2e-5


@pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built')
@pytest.mark.parametrize('s', [6, 11, 18, 24])
def test_emulated_kernels_transient_dual_stereo_streams(s):
    """Streams whose first frames are transient with 16-block collapse masks, dual stereo and
    anti-collapse (these exposed an 8-bit collapse-mask truncation the golden set did not)."""
    from test_gpu_decode import _make_streams
    ch, fs, nfr = 2, 960, 6
    pk, sz = _make_streams(25, ch, fs, nfr, seed=ch * 10 + fs)
    ref, rr = util.ref_decode_float(pk[s], sz[s], ch, fs)
    got, gr = emu_util.emu_decode_stream(pk[s], sz[s], ch, fs)
    assert (gr == rr).all()
    assert np.abs(util.to_int16(got).astype(np.int32) - util.to_int16(ref).astype(np.int32)).max() <= 1
    assert np.abs(got - ref).max() < 2e-6


@pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built')
@pytest.mark.parametrize('ch,fs,br', [(2, 120, 96000), (2, 120, 24000), (2, 240, 32000), (1, 480, 20000)])
def test_emulated_kernels_short_frames(ch, fs, br):
    """2.5/5/10 ms frames incl. stereo N==1 bands (sign-only leaves) and low-rate folding."""
    x = util.sig_music(77 + fs, 10 * fs, ch)
    pk, sz, _ = util.ref_encode(x, ch, fs, br)
    ref, rr = util.ref_decode_float(pk, sz, ch, fs)
    got, gr = emu_util.emu_decode_stream(pk, sz, ch, fs)
    assert (gr == rr).all()
    assert np.abs(util.to_int16(got).astype(np.int32) - util.to_int16(ref).astype(np.int32)).max() <= 1
    assert np.abs(got - ref).max() < 2e-6


@pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built')
@pytest.mark.parametrize('seed', range(16))
def test_emulated_kernels_randomised(seed):
    """seeded random streams from the reference ENCODER (channels, frame size, bit-rate 8..400 kb/s, CBR/VBR, complexity, signal type):
    the three decode kernels' source, run on the host emulator, against the reference decoder — final range exact, PCM within 1 LSB.
    (Since the entropy kernel owns the energy state, this also walks the state hand-over between the kernels on every draw.)"""
    rng = np.random.default_rng(9000 + seed)
    ch = int(rng.integers(1, 3))
    fs = int(rng.choice([120, 240, 480, 960]))
    br = int(rng.choice([8000, 16000, 32000, 64000, 128000, 256000, 400000])) * ch // 2 + 6000
    vbr = int(rng.integers(0, 2))
    cx = int(rng.choice([0, 5, 10]))
    nfr = 8
    gens = [util.sig_music, util.sig_white, util.sig_harmonic, util.sig_swell]
    x = gens[int(rng.integers(0, 4))](seed, nfr * fs, ch)
    if rng.random() < 0.25:
        x = (x * 0.01).astype(np.int16)
    x = np.asarray(x).reshape(-1, ch)[:nfr * fs]       # (some generators return interleaved 1-D data)
    pk, sz, _ = util.ref_encode(x, ch, fs, br, vbr=vbr, complexity=cx)
    ref, rr = util.ref_decode_float(pk, sz, ch, fs)
    got, gr = emu_util.emu_decode_stream(pk, sz, ch, fs)
    assert (gr == rr).all(), (seed, ch, fs, br, vbr, cx)
    assert np.abs(util.to_int16(got).astype(np.int32) - util.to_int16(ref).astype(np.int32)).max() <= 1
    assert np.abs(got - ref).max() < 4e-6
