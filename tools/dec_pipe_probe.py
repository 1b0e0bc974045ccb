"""Device-resident decode throughput, stream-ordered vs pipelined calls (quick probe; bench.py is the reported number)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench, opus_b200
n = int(os.environ.get('N', 65536))
pk, sz = bench.load_fixture()
T = pk.shape[1] if pk.ndim == 3 else 8
pkt, szt = bench.tile_streams(pk, sz, n)
dev = torch.device('cuda:0')
d_pk = [torch.from_numpy(np.ascontiguousarray(pkt[:, t])).to(dev) for t in range(pkt.shape[1])]
d_len = [torch.from_numpy(np.ascontiguousarray(szt[:, t]).astype(np.int32)).to(dev) for t in range(pkt.shape[1])]
T = len(d_pk)
stride = d_pk[0].shape[1]
d_pcm = torch.empty((n, 1920), dtype=torch.int16, device=dev)
d_ns = torch.empty((n,), dtype=torch.int32, device=dev)
for pipelined in (False, True, False, True):
    dec = opus_b200.DecoderBatch(n, channels=2)
    dec.set_pipelined(pipelined)
    ext = torch.cuda.ExternalStream(dec.cuda_stream, device=dev)
    def step(i):
        t = i % T
        dec.decode_device(d_pk[t].data_ptr(), stride, d_len[t].data_ptr(), d_pcm.data_ptr(), 1920, int16=True, frame_size=960,
                          max_frames_per_packet=1, d_out_samples=d_ns.data_ptr())
    for i in range(3):
        step(i)
    dec.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    K = 16
    for i in range(K):
        step(3 + i)
    dec.join()
    e1.record(ext)
    dec.sync()
    ms = e0.elapsed_time(e1) / K
    print('pipelined=%d  %.3f ms/step  %.2f M frames/s' % (pipelined, ms, n / ms / 1e3))
    dec.close()
