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
