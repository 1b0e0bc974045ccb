#!/usr/bin/env python3
"""Generates tests/golden/bench_packets.npz from the reference's pure-C build: the synthetic
workload of BASELINE.json configs[1] ("batched CELT decode: 48 kHz stereo 20 ms frames @128 kbps").
256 distinct streams (music / tonal+noise / white-noise signals) x 8 consecutive
CBR packets each, taken after a 6-frame warm-up so the content is steady-state.  bench.py tiles
the 256 streams up to the requested stream count."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import util  # noqa: E402

S, T, WARM, FS, CH, BR = 256, 8, 6, 960, 2, 128000


def main():
    pk = np.zeros((S, T, 321), np.uint8)
    sz = np.zeros((S, T), np.int32)
    sigs = list(util.SIGNALS)
    for s in range(S):
        x = util.SIGNALS[sigs[s % 3]](7000 + s, (T + WARM) * FS, CH)
        p, z, _ = util.ref_encode(x, CH, FS, BR, vbr=0, complexity=10)
        assert z.max() <= 321
        pk[s] = p[WARM:, :321]
        sz[s] = z[WARM:]
    p = os.path.join(util.GOLDEN, 'bench_packets.npz')
    np.savez_compressed(p, pkts=pk, sizes=sz, meta=np.array([CH, FS, BR], np.int32))
    print('wrote', p, os.path.getsize(p), 'bytes; packet sizes', sz.min(), sz.max())


if __name__ == '__main__':
    main()
