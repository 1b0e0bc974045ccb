"""The oracle restatement against the committed golden vectors (generated from the reference's
pure-C build by tests/golden/make_golden.py): final range bit-exact, PCM bit-exact.
Runs anywhere (no /root/reference needed)."""
import numpy as np
import pytest

import util


def _golden():
    return np.load(util.GOLDEN + '/celt_decode_golden.npz')


def golden_cases():
    g = _golden()
    for name in g['names']:
        name = str(name)
        pk = g[name + '/pkts']
        pkts = np.zeros((pk.shape[0], 1500), np.uint8)
        pkts[:, :pk.shape[1]] = pk
        ch, fs = [int(v) for v in g[name + '/meta']]
        yield name, pkts, g[name + '/sizes'], g[name + '/rng'], g[name + '/pcm'], ch, fs


@pytest.mark.parametrize('case', list(golden_cases()), ids=lambda c: c[0])
def test_oracle_reproduces_golden(case):
    name, pkts, sizes, rng, pcm, ch, fs = case
    got, got_rng = util.oracle_decode_float(pkts, sizes, ch, fs)
    assert (got_rng == rng).all()
    assert np.array_equal(got, pcm)
