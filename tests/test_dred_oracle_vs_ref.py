"""Pins oracle/dred_oracle.c against the reference's own dnn/dred_rdovae_dec.c (DRED_rdovae_decode_all) compiled with a seeded random,
block-sparsified model (oracle/gen_dred_weights.py).  Same build flags as the FARGAN pin (generic C branch of dnn/vec.h).
Bar: bit-identical feature frames."""
import numpy as np
import pytest

import fargan_util as F
import util

R = F.load_ref()
pytestmark = pytest.mark.skipif(R is None or not hasattr(R, 'refh_dred_decode_all'), reason='oracle/_ref/libfargan_ref.so without the DRED decoder')


@pytest.mark.parametrize('seed,nb', [(1, 1), (2, 7), (3, 26), (4, 50)])
def test_dred_decode_all_matches_reference(seed, nb):
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    ref = F.RefDred(R)
    layers, shapes = ref.layers()
    assert shapes[0][:2] == (ref.latent + 1, 96) and shapes[7][:2] == (576, 80) and shapes[13][:2] == (128, 320)
    assert shapes[14][2] and shapes[8][2] and shapes[7][2]          # input-sparse GRU matrices, sparse conv_dense / output layers
    state, latents = F.make_dred_inputs(seed, nb, ref.latent, ref.state)
    want = ref.decode_all(state, latents)
    got = np.zeros_like(want)
    O.orc_dred_decode_all(layers, F.vp(got), F.vp(state), F.vp(latents), nb, ref.latent)
    assert np.array_equal(got, want), float(np.abs(got - want).max())
    assert float(np.abs(want).max()) > 1e-3
