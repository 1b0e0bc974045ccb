"""Pins oracle/fargan_oracle.c (our restatement of dnn/fargan.c) against the reference's own fargan.c compiled with seeded random
weights (oracle/gen_fargan_weights.py).  The reference library is built on dnn/vec.h's generic C branch (-U__SSE2__ ...):
that is the deterministic flavour every non-x86/NEON target runs; the x86 branch evaluates tanh/sigmoid through _mm_rcp_ps,
whose ~2e-4 error flips 8-bit activation roundings and makes two runs on different CPUs disagree at the 1e-2 level.
Bar: bit-identical PCM and state."""
import numpy as np
import pytest

import fargan_util as F
import util

R = F.load_ref()
pytestmark = pytest.mark.skipif(R is None, reason='oracle/_ref/libfargan_ref.so not built (needs /root/reference)')


@pytest.mark.parametrize('seed', [1, 2, 3, 4])
def test_fargan_oracle_matches_reference(seed):
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    ref = F.RefFargan(R)
    layers, shapes = ref.layers()
    assert shapes[7] == (272, 480) and shapes[17] == (688, 128) and shapes[18] == (128, 40)     # the sizes SURVEY 8(a) a19 quotes
    st = np.zeros(O.orc_fargan_size() + 64, np.uint8)
    O.orc_fargan_init(F.vp(st), layers)
    pcm0, feat0, feats = F.make_inputs(seed, 25)
    ref.cont(pcm0, feat0)
    O.orc_fargan_cont(F.vp(st), F.vp(pcm0), F.vp(feat0))
    energy = 0.0
    for k in range(len(feats)):
        want = ref.synth(feats[k])
        got = np.zeros(160, np.float32)
        O.orc_fargan_synthesize(F.vp(st), F.vp(got), F.vp(np.ascontiguousarray(feats[k])))
        assert np.array_equal(got, want), (k, float(np.abs(got - want).max()))
        energy += float((want ** 2).sum())
    assert energy > 1e-3          # the random model does produce a signal
    bufs = [np.zeros(n, np.float32) for n in (256, 128, 328, 160, 128, 128, 2)]
    O.orc_fargan_get_state(F.vp(st), *[F.vp(b) for b in bufs])
    for a, b in zip(bufs, ref.state()):
        assert np.array_equal(a, b)
