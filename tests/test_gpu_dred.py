"""Parity of the batched CUDA DRED decoder network (opusb200_dred_decode_all) with the oracle restatement of dnn/dred_rdovae_dec.c,
itself bit-identical to the reference's own file (tests/test_dred_oracle_vs_ref.py).  Weights: the seeded random block-sparse
model inside oracle/_ref/libfargan_ref.so.  Bar: bit-identical feature frames."""
import numpy as np
import pytest

import fargan_util as F
import util

pytestmark = pytest.mark.gpu
R = F.load_ref()


@pytest.mark.skipif(R is None or not hasattr(R, 'refh_dred_decode_all'), reason='reference DRED build missing')
@pytest.mark.parametrize('n,nb', [(2, 3), (131, 12)])
def test_dred_decode_all_matches_oracle(n, nb):
    import opus_b200
    from opus_b200 import nnet
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    ref = F.RefDred(R)
    layers, shapes = ref.layers()
    ins = [F.make_dred_inputs(40 + s, nb, ref.latent, ref.state) for s in range(min(n, 5))]
    ops = opus_b200.Ops()
    dd = nnet.DredDecoder(ops, layers, n)
    state = np.stack([ins[s % len(ins)][0] for s in range(n)])
    lat = np.stack([ins[s % len(ins)][1] for s in range(n)])
    for rep in range(4):     # call 1 runs eagerly, call 2 is captured into a CUDA graph, later calls replay it
        got = dd.decode_all(state, lat)
    for s in range(len(ins)):   # (checked on the last, replayed call)
        want = np.zeros((nb, 80), np.float32)
        O.orc_dred_decode_all(layers, F.vp(want), F.vp(ins[s][0]), F.vp(ins[s][1]), nb, ref.latent)
        for r in range(s, n, len(ins)):
            assert np.array_equal(got[r], want), (r, float(np.abs(got[r] - want).max()))
    dd.close()


@pytest.mark.skipif(R is None or not hasattr(R, 'refh_dred_encode_dframe'), reason='oracle/_ref/libfargan_ref.so without the DRED encoder')
@pytest.mark.parametrize('n', [4, 150])
def test_dred_encoder_matches_oracle(n):
    """opusb200_dred_encode_dframe (the DRED encoder network, dnn/dred_rdovae_enc.c) for n streams x 12 consecutive double frames with carried
    state, against the oracle (pinned bit-exact on the reference's file): int8 layers are exact, the float layers keep the reference's order;
    tanh / sigmoid go through the shared approximations -> bit-identical expected, 1e-6 allowed"""
    import opus_b200
    from opus_b200 import nnet
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    ref = F.RefDredEnc(R)
    layers, _ = ref.layers()
    ops = opus_b200.Ops()
    enc = nnet.DredEncoder(ops, layers, n)
    kinds = min(n, 5)
    sts = []
    for s in range(kinds):
        st = np.zeros(O.orc_dred_enc_size() + 64, np.uint8)
        O.orc_dred_enc_init(F.vp(st), layers)
        sts.append(st)
    rng = np.random.default_rng(n)
    exact = 0
    for t in range(12):
        xk = (0.6 * rng.standard_normal((kinds, 40))).astype(np.float32)
        x = np.stack([xk[s % kinds] for s in range(n)])
        lat, stt = enc.encode_dframe(x)
        for s in range(kinds):
            wl, ws = np.zeros(25, np.float32), np.zeros(50, np.float32)
            O.orc_dred_encode_dframe(F.vp(sts[s]), F.vp(wl), F.vp(ws), F.vp(np.ascontiguousarray(xk[s])))
            for r in range(s, n, kinds):
                assert np.abs(lat[r] - wl).max() <= 1e-6 and np.abs(stt[r] - ws).max() <= 1e-6, (t, r)
            exact += int(np.array_equal(lat[s], wl) and np.array_equal(stt[s], ws))
    print('bit-identical frames: %d of %d' % (exact, 12 * kinds))
    assert exact >= 12 * kinds - 2
    enc.close()
