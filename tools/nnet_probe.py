"""Times the dense int8 layer (272 -> 480, FARGAN's gru1 input matrix) and a GRU step; run under ncu for the per-kernel split."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import opus_b200, util
from opus_b200 import nnet as nnet_mod
n = int(os.environ.get('N', 65536))
ops = opus_b200.Ops(0)
L = ops._L
stream = torch.cuda.ExternalStream(L.opusb200_ctx_cuda_stream(ops._h))
dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
M, N = 272, 160
Wi, Wr = util.Layer(1, M, 3 * N, 'i8'), util.Layer(2, N, 3 * N, 'i8', diag=True)
upl = lambda Lh: nnet_mod.LinearLayer(ops, Lh.nb_in, Lh.nb_out, bias=Lh.bias, subias=Lh.subias, weights=Lh.weights, float_weights=Lh.float_weights,
                                      weights_idx=Lh.idx, diag=Lh.diag, scale=Lh.scale, su_bias=True)
dWi, dWr = upl(Wi), upl(Wr)
xs = torch.from_numpy(np.clip(rng.standard_normal((n, M)) * 0.4, -1, 1).astype(np.float32)).to(dev)
sts = torch.zeros((n, N), dtype=torch.float32, device=dev)
outl = torch.empty((n, 3 * N), dtype=torch.float32, device=dev)
dptr = lambda t: ctypes.c_void_p(t.data_ptr())
def timeit(fn, reps=10):
    with torch.cuda.stream(stream):
        for _ in range(3): fn()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps): fn()
        e1.record(stream); e1.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timeit(lambda: L.opusb200_nnet_compute_linear(ops._h, dWi._h, dptr(outl), dptr(xs), n, 1))
print('linear 272x480 n=%d: %.4f ms  %.1f TOPS' % (n, ms, 2.0 * n * M * 3 * N / (ms * 1e-3) / 1e12))
ms = timeit(lambda: L.opusb200_nnet_compute_gru(ops._h, dWi._h, dWr._h, dptr(sts), dptr(xs), n, 1))
print('gru 272/160 n=%d: %.4f ms' % (n, ms))
