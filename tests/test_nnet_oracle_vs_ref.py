"""Pin oracle/nnet_oracle.c to the reference's compute_linear_c / compute_activation_c / compute_conv2d_c
(dnn/nnet_default.c compiled into oracle/_ref/libopus_ref_c.so) on random layers.

On x86-64 that translation unit takes the vec_avx.h (SSE2) flavour: float GEMV in the same order as the generic code
(bit-exact), int8 GEMV with the unsigned-activation convention (orc_nnet_set_su_bias(1): bit-exact), and tanh / sigmoid
built on _mm_rcp_ps, whose ~12-bit result is CPU-specific - those are compared with the tolerance the reference itself
quotes (vec_avx.h:391-399, "max error is around 3e-4").  The oracle's activations follow the deterministic generic branch."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built (needs /root/reference)')


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


SHAPES = [(272, 480), (160, 480), (688, 128), (128, 40), (40, 24), (20, 13), (192, 192)]


@pytest.mark.parametrize('kind', ['f32', 'f32_sparse', 'i8', 'i8_sparse'])
def test_linear(kind):
    R, O = util.load_ref('c'), util.load_oracle()
    O.orc_nnet_set_su_bias(1)
    rng = np.random.default_rng(5)
    for si, (m, n) in enumerate(SHAPES):
        if kind != 'f32' and (m % 4 or n % 8):
            continue
        for diag in (False, True):
            if diag and kind == 'f32_sparse':
                continue
            mm, nn = (n // 3 // 8 * 8, n // 3 // 8 * 8 * 3) if diag else (m, n)
            if diag and (mm == 0 or mm % 4):
                continue
            L = util.Layer(100 * si + diag, mm, nn, kind, bias=(si % 2 == 0), diag=diag)
            for t in range(3):
                x = np.clip(rng.standard_normal(mm) * 0.5, -1, 1).astype(np.float32)
                a, b = np.zeros(nn, np.float32), np.zeros(nn, np.float32)
                R.refh_nnet_linear(_p(a), _p(x), _p(L.bias), _p(L.subias), _p(L.weights), _p(L.float_weights), _p(L.idx), _p(L.diag), _p(L.scale), mm, nn)
                O.orc_nnet_linear(ctypes.byref(L.c), _p(b), _p(x))
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (kind, m, n, diag)
                assert np.abs(a).max() > 0


@pytest.mark.parametrize('act', [0, 1, 2, 3, 4, 5, 6])
def test_activation(act):
    R, O = util.load_ref('c'), util.load_oracle()
    rng = np.random.default_rng(act)
    x = np.concatenate([rng.standard_normal(500) * 3, rng.standard_normal(200) * 30, [0, -0.0, 1e-8, -80, 80, 100, -100]]).astype(np.float32)
    a, b = np.zeros_like(x), np.zeros_like(x)
    R.refh_nnet_activation(_p(a), _p(x), len(x), act)
    O.orc_nnet_activation(_p(b), _p(x), len(x), act)
    if act in (0, 3):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    elif act in (1, 2):
        assert np.abs(a - b).max() < 6e-4
    elif act == 5:
        assert np.abs(a - b).max() < 6e-4 * np.abs(x).max()
    elif act == 4:   # softmax over the whole vector: the x86 build clamps exponents to +-50
        ok = np.abs(x) < 30
        assert np.allclose(a[ok] / a[ok].sum(), b[ok] / b[ok].sum(), rtol=1e-5)
    else:            # exp
        ok = np.abs(x) < 30
        assert np.allclose(a[ok], b[ok], rtol=2e-6, atol=0)


@pytest.mark.parametrize('k', [(3, 3), (3, 1), (2, 3)])
def test_conv2d(k):
    R, O = util.load_ref('c'), util.load_oracle()
    rng = np.random.default_rng(9)
    kt, kh = k
    inc, outc, height = 4, 6, 20
    hstride = height + 3
    w = (rng.standard_normal(outc * inc * kt * kh) * 0.3).astype(np.float32)
    bias = (rng.standard_normal(outc) * 0.1).astype(np.float32)
    ts = inc * (height + kh - 1)
    mem_a = (rng.standard_normal((kt - 1) * ts)).astype(np.float32)
    mem_b = mem_a.copy()
    c = util.CConv2d()
    c.bias, c.float_weights = bias.ctypes.data, w.ctypes.data
    c.in_channels, c.out_channels, c.ktime, c.kheight = inc, outc, kt, kh
    for t in range(3):
        x = rng.standard_normal(ts).astype(np.float32)
        a, b = np.zeros(outc * hstride, np.float32), np.zeros(outc * hstride, np.float32)
        R.refh_nnet_conv2d(_p(a), _p(mem_a), _p(x), _p(bias), _p(w), inc, outc, kt, kh, height, hstride, 2)
        O.orc_nnet_conv2d(ctypes.byref(c), _p(b), _p(mem_b), _p(x), height, hstride, 2)
        assert np.abs(a - b).max() < 6e-4      # tanh on top of a bit-exact convolution
        assert np.array_equal(mem_a, mem_b)
