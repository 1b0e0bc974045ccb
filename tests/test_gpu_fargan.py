"""Parity of the batched CUDA FARGAN vocoder (opusb200_fargan_*) with the oracle restatement of dnn/fargan.c, which is pinned
bit-exact against the reference's own fargan.c (tests/test_fargan_oracle_vs_ref.py).  Weights: the seeded random model inside
oracle/_ref/libfargan_ref.so (it travels to the GPU box; /root/reference does not).
Bar: int8 layers are exact integer arithmetic and the float layers keep the reference's summation order, so the PCM is expected
bit-identical; the two libm calls of the graph (exp for the gain, pow for the period) are the device's, which allows a last-ulp
difference there: the test accepts 1e-6 absolute and reports how many frames were bit-identical."""
import numpy as np
import pytest

import fargan_util as F
import util

pytestmark = pytest.mark.gpu
R = F.load_ref()


@pytest.mark.skipif(R is None, reason='oracle/_ref/libfargan_ref.so missing')
@pytest.mark.parametrize('n', [3, 130])
def test_fargan_matches_oracle(n):
    import opus_b200
    from opus_b200 import nnet
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    ref = F.RefFargan(R)
    layers, shapes = ref.layers()
    nfr = 10
    ins = [F.make_inputs(100 + s, nfr) for s in range(min(n, 7))]
    ops = opus_b200.Ops()
    fg = nnet.Fargan(ops, layers, n)
    pcm0 = np.stack([ins[s % len(ins)][0] for s in range(n)])
    feat0 = np.stack([ins[s % len(ins)][1] for s in range(n)])
    fg.cont(pcm0, feat0)
    sts = []
    for s in range(len(ins)):
        st = np.zeros(O.orc_fargan_size() + 64, np.uint8)
        O.orc_fargan_init(F.vp(st), layers)
        O.orc_fargan_cont(F.vp(st), F.vp(ins[s][0]), F.vp(ins[s][1]))
        sts.append(st)
    exact = 0
    for k in range(nfr):
        feats = np.stack([ins[s % len(ins)][2][k] for s in range(n)])
        got = fg.synthesize(feats)
        for s in range(len(ins)):
            want = np.zeros(160, np.float32)
            O.orc_fargan_synthesize(F.vp(sts[s]), F.vp(want), F.vp(np.ascontiguousarray(ins[s][2][k])))
            for r in range(s, n, len(ins)):
                assert np.abs(got[r] - want).max() <= 1e-6, (k, r, float(np.abs(got[r] - want).max()))
            exact += int(np.array_equal(got[s], want))
    print('bit-identical frames: %d of %d' % (exact, nfr * len(ins)))
    assert exact >= nfr * len(ins) - 2
    fg.close()


def test_fargan_requires_cont_and_checks_sizes():
    import opus_b200
    from opus_b200 import nnet
    if R is None:
        pytest.skip('no weights')
    ref = F.RefFargan(R)
    layers, _ = ref.layers()
    ops = opus_b200.Ops()
    fg = nnet.Fargan(ops, layers, 4)
    with pytest.raises(opus_b200.OpusB200Error):
        fg.synthesize(np.zeros((4, 20), np.float32))       # celt_assert(st->cont_initialized)
    fg.close()
    layers[7].nb_inputs = 100
    with pytest.raises(opus_b200.OpusB200Error):
        nnet.Fargan(ops, layers, 4)
