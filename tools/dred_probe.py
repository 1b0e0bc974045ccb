"""Times opusb200_dred_decode_all (25 latents = 1 s of redundancy per stream) for a batch; run under ncu for the kernel split."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import opus_b200, fargan_util as FG
from opus_b200 import nnet as nn_mod
n, nbl = int(os.environ.get('N', 32768)), 25
ops = opus_b200.Ops(0)
L = ops._L
stream = torch.cuda.ExternalStream(L.opusb200_ctx_cuda_stream(ops._h))
dev = torch.device('cuda:0')
rd = FG.RefDred(FG.load_ref())
dl, _ = rd.layers()
dd = nn_mod.DredDecoder(ops, dl, n)
st0, lat0 = FG.make_dred_inputs(3, nbl, rd.latent, rd.state)
d_st = torch.from_numpy(np.tile(st0, (n, 1))).to(dev)
d_lat = torch.from_numpy(np.tile(lat0[None], (n, 1, 1))).to(dev)
d_ft = torch.empty((n, nbl, 80), dtype=torch.float32, device=dev)
dptr = lambda t: ctypes.c_void_p(t.data_ptr())
fn = lambda: L.opusb200_dred_decode_all(dd._h, dptr(d_ft), dptr(d_st), dptr(d_lat), nbl, 1)
with torch.cuda.stream(stream):
    for _ in range(3):
        fn()
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        fn()
    e1.record(stream); e1.synchronize()
ms = e0.elapsed_time(e1) / 3
print('dred_decode_all n=%d x %d latents: %.2f ms  (%.2f M stream-seconds of DRED per second)' % (n, nbl, ms, n / ms / 1e3))
