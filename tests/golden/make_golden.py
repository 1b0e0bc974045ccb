#!/usr/bin/env python3
"""Generates tests/golden/celt_decode_golden.npz from the REFERENCE ITSELF (oracle/_ref/
libopus_ref_c.so = unmodified xiph/opus sources, pure-C float build):  synthetic PCM ->
reference encoder (CELT-only) -> packets; reference decoder -> float PCM + final range.
Run in the build container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
The fixture travels with the repo so the CPU suite and the GPU box need no reference tree."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import util  # noqa: E402

CASES = [  # name, signal, channels(enc), frame, bitrate, vbr, ctls, dec_channels
    ('st20_128k', 'music', 2, 960, 128000, 0, (), 2),
    ('st20_64k_vbr', 'tonal', 2, 960, 64000, 1, (), 2),
    ('st20_24k', 'music', 2, 960, 24000, 0, (), 2),
    ('mono20_48k', 'music', 1, 960, 48000, 0, (), 1),
    ('st10_96k', 'white', 2, 480, 96000, 0, (), 2),
    ('st5_64k', 'music', 2, 240, 64000, 0, (), 2),
    ('mono2p5_96k', 'tonal', 1, 120, 96000, 0, (), 1),
    ('st20_510k', 'white', 2, 960, 510000, 0, (), 2),
    ('st20_wb_48k', 'music', 2, 960, 48000, 0, ((util.OPUS_SET_BANDWIDTH_REQUEST, util.OPUS_BANDWIDTH['wb']),), 2),
    ('mono_to_stereo', 'music', 1, 960, 48000, 0, (), 2),
    ('stereo_to_mono', 'music', 2, 960, 96000, 0, (), 1),
]


def main():
    out = {}
    names = []
    for k, (name, sig, ch, fs, br, vbr, ctls, dch) in enumerate(CASES):
        nfr = 8 if fs == 960 else 16
        pcm = util.SIGNALS[sig](1000 + k, nfr * fs, ch)
        pkts, sizes, enc_rng = util.ref_encode(pcm, ch, fs, br, vbr=vbr, ctls=ctls)
        dec, dec_rng = util.ref_decode_float(pkts, sizes, dch, fs)
        if dch == ch:
            assert (dec_rng == enc_rng).all()
        mx = int(sizes.max())
        out[name + '/pkts'] = pkts[:, :mx].copy()
        out[name + '/sizes'] = sizes
        out[name + '/rng'] = dec_rng
        out[name + '/pcm'] = dec
        out[name + '/meta'] = np.array([dch, fs], np.int32)
        names.append(name)
    out['names'] = np.array(names)
    p = os.path.join(util.GOLDEN, 'celt_decode_golden.npz')
    np.savez_compressed(p, **out)
    print('wrote', p, os.path.getsize(p), 'bytes')


if __name__ == '__main__':
    main()
