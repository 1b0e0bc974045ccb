"""One batched fargan_synthesize step, timed (run under ncu --metrics gpu__time_duration.sum for the per-kernel split)."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import opus_b200, fargan_util as FG
from opus_b200 import nnet as nn_mod
n = int(os.environ.get('N', 32768))
ops = opus_b200.Ops(0)
L = ops._L
stream = torch.cuda.ExternalStream(L.opusb200_ctx_cuda_stream(ops._h))
dev = torch.device('cuda:0')
ref = FG.RefFargan(FG.load_ref())
layers, _ = ref.layers()
fg = nn_mod.Fargan(ops, layers, n)
pcm0, feat0, feats = FG.make_inputs(5, 4)
fg.cont(np.tile(pcm0, (n, 1)), np.tile(feat0, (n, 1)))
d_feat = torch.from_numpy(np.tile(feats[0], (n, 1)).astype(np.float32)).to(dev)
d_out = torch.empty((n, 160), dtype=torch.float32, device=dev)
dptr = lambda t: ctypes.c_void_p(t.data_ptr())
reps = int(os.environ.get('REPS', 5))
with torch.cuda.stream(stream):
    for _ in range(2):
        L.opusb200_fargan_synthesize(fg._h, dptr(d_out), dptr(d_feat), 1)
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        L.opusb200_fargan_synthesize(fg._h, dptr(d_out), dptr(d_feat), 1)
    e1.record(stream); e1.synchronize()
ms = e0.elapsed_time(e1) / reps
print('fargan_synthesize n=%d: %.3f ms per 10 ms frame = %.2f M frames/s (%.2fx real time for the batch)' % (n, ms, n / ms / 1e3, 10.0 / ms))
