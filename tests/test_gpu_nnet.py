"""dnn primitives on the GPU through the C ABI against the pinned oracle: float32 bit-exact."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


@pytest.fixture(scope='module')
def ctx():
    import opus_b200
    from opus_b200 import nnet
    ops = opus_b200.Ops(0)
    yield ops, nnet.Nnet(ops), nnet
    ops.close()


def upload(nnet, ops, L, su=True):
    return nnet.LinearLayer(ops, L.nb_in, L.nb_out, bias=L.bias, subias=L.subias, weights=L.weights, float_weights=L.float_weights,
                            weights_idx=L.idx, diag=L.diag, scale=L.scale, su_bias=su)


def oracle_rows(fn, n, width, *args):
    out = np.zeros((n, width), np.float32)
    for s in range(n):
        fn(s, out[s])
    return out


@pytest.mark.parametrize('kind', ['f32', 'f32_sparse', 'i8', 'i8_sparse'])
@pytest.mark.parametrize('su', [True, False])
def test_linear(ctx, kind, su):
    ops, nn, nnet = ctx
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(1 if su else 0)
    rng = np.random.default_rng(21)
    shapes = [(272, 480, False), (160, 480, True), (688, 128, False), (128, 40, False), (40, 24, False)]
    if kind == 'f32':
        shapes.append((20, 13, False))
    for si, (m, n_out, diag) in enumerate(shapes):
        if diag and kind == 'f32_sparse':
            continue
        L = util.Layer(31 * si + 5, m, n_out, kind, bias=(si != 2), diag=diag)
        n = 1000 + 7 * si
        x = np.clip(rng.standard_normal((n, m)) * 0.5, -1, 1).astype(np.float32)
        dl = upload(nnet, ops, L, su)
        got = nn.linear(dl, x)
        rows = list(range(0, n, 37)) + [n - 1]
        for s in rows:
            w = np.zeros(n_out, np.float32)
            O.orc_nnet_linear(ctypes.byref(L.c), _p(w), _p(x[s]))
            assert np.array_equal(got[s].view(np.uint32), w.view(np.uint32)), (kind, su, m, n_out, s)
        dl.close()


def test_gru_glu_dense_sequence(ctx):
    """a FARGAN-shaped step (dnn/fargan.c:69-140 uses exactly these primitives) iterated over 6 subframes"""
    ops, nn, nnet = ctx
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(1)
    rng = np.random.default_rng(8)
    n, M, N = 515, 272, 160
    Wi, Wr = util.Layer(1, M, 3 * N, 'i8'), util.Layer(2, N, 3 * N, 'i8', diag=True)
    G = util.Layer(3, N, N, 'i8')
    D = util.Layer(4, N, 40, 'f32')
    dWi, dWr, dG, dD = (upload(nnet, ops, l) for l in (Wi, Wr, G, D))
    st = np.zeros((n, N), np.float32)
    st_o = st.copy()
    for step in range(6):
        x = np.clip(rng.standard_normal((n, M)) * 0.4, -1, 1).astype(np.float32)
        st = nn.gru(dWi, dWr, st, x)
        g = nn.glu(dG, st)
        y = nn.dense(dD, g, 2)
        for s in range(0, n, 29):
            O.orc_nnet_gru(ctypes.byref(Wi.c), ctypes.byref(Wr.c), _p(st_o[s]), _p(x[s]))
            go = np.zeros(N, np.float32)
            O.orc_nnet_glu(ctypes.byref(G.c), _p(go), _p(st_o[s]))
            yo = np.zeros(40, np.float32)
            O.orc_nnet_dense(ctypes.byref(D.c), _p(yo), _p(go), 2)
            assert np.array_equal(st[s].view(np.uint32), st_o[s].view(np.uint32)), (step, s)
            assert np.array_equal(g[s].view(np.uint32), go.view(np.uint32))
            assert np.array_equal(y[s].view(np.uint32), yo.view(np.uint32))
        # rows the oracle did not follow: keep its copy in step with the device state
        mask = np.ones(n, bool)
        mask[::29] = False
        st_o[mask] = st[mask]


@pytest.mark.parametrize('dilation', [1, 2])
def test_conv1d(ctx, dilation):
    ops, nn, nnet = ctx
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(1)
    rng = np.random.default_rng(4)
    n, isz, ksize, nout = 130, 64, 3, 96
    L = util.Layer(9, isz * ksize, nout, 'i8')
    dl = upload(nnet, ops, L)
    mem_len = isz * (ksize - 1) if dilation == 1 else isz * dilation * (ksize - 1)
    mem = np.zeros((n, mem_len), np.float32)
    mem_o = mem.copy()
    for step in range(5):
        x = np.clip(rng.standard_normal((n, isz)) * 0.5, -1, 1).astype(np.float32)
        out, mem = nn.conv1d(dl, mem, x, 2, dilation)
        for s in range(n):
            w = np.zeros(nout, np.float32)
            if dilation == 1:
                O.orc_nnet_conv1d(ctypes.byref(L.c), _p(w), _p(mem_o[s]), _p(x[s]), isz, 2)
            else:
                O.orc_nnet_conv1d_dilation(ctypes.byref(L.c), _p(w), _p(mem_o[s]), _p(x[s]), isz, dilation, 2)
            assert np.array_equal(out[s].view(np.uint32), w.view(np.uint32)), (step, s)
        assert np.array_equal(mem, mem_o)


@pytest.mark.parametrize('k', [(3, 3), (2, 3), (1, 1)])
def test_conv2d(ctx, k):
    ops, nn, nnet = ctx
    O = util.load_oracle()
    rng = np.random.default_rng(6)
    kt, kh = k
    n, inc, outc, height = 70, 4, 6, 20
    hstride = height + 2
    w = (rng.standard_normal(outc * inc * kt * kh) * 0.3).astype(np.float32)
    bias = (rng.standard_normal(outc) * 0.1).astype(np.float32)
    ts = inc * (height + kh - 1)
    c = util.CConv2d()
    c.bias, c.float_weights = bias.ctypes.data, w.ctypes.data
    c.in_channels, c.out_channels, c.ktime, c.kheight = inc, outc, kt, kh
    dl = nnet.Conv2dLayer(ops, w, bias, inc, outc, kt, kh)
    mem = rng.standard_normal((n, (kt - 1) * ts)).astype(np.float32)
    mem_o = mem.copy()
    for step in range(3):
        x = rng.standard_normal((n, ts)).astype(np.float32)
        out, mem = nn.conv2d(dl, mem, x, height, hstride, 2)
        for s in range(n):
            wv = np.zeros(outc * hstride, np.float32)
            O.orc_nnet_conv2d(ctypes.byref(c), _p(wv), _p(mem_o[s]) if mem_o.size else None, _p(x[s]), height, hstride, 2)
            assert np.array_equal(out[s].view(np.uint32), wv.view(np.uint32)), (step, s)
        assert np.array_equal(mem, mem_o)
    dl.close()


def test_activation(ctx):
    ops, nn, nnet = ctx
    O = util.load_oracle()
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((300, 77)) * 5).astype(np.float32)
    for act in range(7):
        got = nn.activation(x, act)
        for s in range(0, 300, 11):
            w = np.zeros(77, np.float32)
            O.orc_nnet_activation(_p(w), _p(x[s]), 77, act)
            assert np.array_equal(got[s].view(np.uint32), w.view(np.uint32)), act
