"""dnn kernels (opus_b200/csrc/nnet.cuh) on the host SIMT emulator against the pinned oracle: bit-exact float32.
The tensor-core instruction is emulated from its documented fragment layout (simt/nnet.cuh: mma_m16n8k32_i8)."""
import ctypes

import numpy as np
import pytest

import emu_util
import util


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def oracle_linear(O, L, x, su):
    O.orc_nnet_set_su_bias(1 if su else 0)
    out = np.zeros((x.shape[0], L.nb_out), np.float32)
    for s in range(x.shape[0]):
        O.orc_nnet_linear(ctypes.byref(L.c), _p(out[s]), _p(x[s]))
    return out


@pytest.mark.parametrize('kind', ['f32', 'f32_sparse', 'i8', 'i8_sparse'])
@pytest.mark.parametrize('su', [1, 0])
def test_emu_linear(kind, su):
    if su == 0 and not kind.startswith('i8'):
        pytest.skip('convention only matters for int8 layers')
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    rng = np.random.default_rng(11)
    shapes = [(272, 480, False), (40, 24, False), (48, 144, True), (20, 13, False)] if kind == 'f32' else [(272, 480, False), (40, 24, False), (48, 144, True)]
    for si, (m, n_out, diag) in enumerate(shapes):
        if diag and kind == 'f32_sparse':
            continue
        L = util.Layer(7 * si + 1, m, n_out, kind, bias=True, diag=diag)
        n = 37 if m < 100 else 35
        x = np.clip(rng.standard_normal((n, m)) * 0.5, -1, 1).astype(np.float32)
        want = oracle_linear(O, L, x, su)
        got = np.zeros_like(want)
        E.emu_nnet_linear(ctypes.byref(L.c), su, _p(got), _p(x), n)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (kind, su, m, n_out)


def test_emu_gru_and_activation():
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    O.orc_nnet_set_su_bias(1)
    rng = np.random.default_rng(3)
    n, M, N = 9, 40, 24
    Wi = util.Layer(1, M, 3 * N, 'i8')
    Wr = util.Layer(2, N, 3 * N, 'i8', diag=True)
    x = np.clip(rng.standard_normal((n, M)) * 0.5, -1, 1).astype(np.float32)
    st = np.clip(rng.standard_normal((n, N)) * 0.5, -1, 1).astype(np.float32)
    want = st.copy()
    for s in range(n):
        O.orc_nnet_gru(ctypes.byref(Wi.c), ctypes.byref(Wr.c), _p(want[s]), _p(x[s]))
    zrh, rec = np.zeros((n, 3 * N), np.float32), np.zeros((n, 3 * N), np.float32)
    E.emu_nnet_linear(ctypes.byref(Wi.c), 1, _p(zrh), _p(x), n)
    E.emu_nnet_linear(ctypes.byref(Wr.c), 1, _p(rec), _p(st), n)
    got = st.copy()
    E.emu_nnet_gru_gates(_p(got), _p(zrh), _p(rec), n, N)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    for act in range(7):
        a = (rng.standard_normal((5, 33)) * 4).astype(np.float32)
        w = np.zeros_like(a)
        for s in range(5):
            O.orc_nnet_activation(_p(w[s]), _p(a[s]), 33, act)
        g = np.zeros_like(a)
        E.emu_nnet_activation(_p(g), _p(a), 5, 33, act)
        assert np.array_equal(g.view(np.uint32), w.view(np.uint32)), act
