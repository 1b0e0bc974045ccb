#!/usr/bin/env python3
"""bench.py — batched Opus (CELT) throughput on B200, two workloads in ONE JSON line:

  * top level = BASELINE.json configs[1], batched CELT DECODE: 65 536 independent 48 kHz stereo 20 ms streams @128 kb/s CBR per GPU,
    one packet per stream per step, synthetic packets from tests/golden/bench_packets.npz (256 distinct streams x 8 frames, produced by
    the reference encoder; tiled);
  * "encode" block = BASELINE.json configs[2], batched CELT ENCODE through the Opus API (opus_encode, OPUS_APPLICATION_RESTRICTED_CELT,
    complexity 10, 64 kb/s VBR): 65 536 independent 48 kHz stereo 20 ms streams per GPU, one frame per stream per step, >= 50 warm
    frames per stream before the timed region, synthetic int16 PCM (256 distinct streams x 32 consecutive frames, tiled), the
    north-star target workload.  It carries its own value / e2e / roofline / cpu_baseline.
Weak scaling: every GPU gets its own 65 536 streams.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--streams S] [--impl reference] [--workload both|decode|encode]

`value` = frames/s with inputs and outputs resident in HBM; `e2e` = the same through the host-buffer C-ABI call (H2D of the inputs and
D2H of the results inside the timed region, pinned host buffers).  `--impl reference` times the unmodified reference (oracle/_ref:
the AVX2-presumed build of BASELINE.md section 4 when the host has AVX2, else the RTCD build) on the host cores, same two workloads.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME = 960
CH = 2
PKT_STRIDE = 320
# algorithmic HBM bytes per stream-frame (DESIGN.md "Roofline bookkeeping")
SCRIPT_BYTES = 2600  # typical script per frame at 128 kb/s: ~110 leaves x 16 B + ~220 node bytes + 21 band records x 24 B
BYTES_KERNEL = {
    'ob_celt_dec_entropy': 320 + 340 + SCRIPT_BYTES + 168 + 64,            # packet in, side out, script out, energies in, state
    'ob_celt_dec_pvq': SCRIPT_BYTES + 340 + 2 * 960 * 4 + 2 * 672,         # script in, side in, X out, energy state in+out
    'ob_celt_dec_synth': 2 * 960 * 4 + 340 + 2 * 960 * 4 + 2 * 960 * 2 + 2 * 2 * 60 * 4,  # X in, side in, history out, PCM out, tails
}
BYTES_WHOLE_FRAME = 320 + 2 * 960 * 2                    # packet in + int16 PCM out


def load_fixture():
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'bench_packets.npz'))
    pk, sz = g['pkts'], g['sizes']  # [256][8][321], [256][8]
    assert int(sz.max()) <= PKT_STRIDE
    return np.ascontiguousarray(pk[:, :, :PKT_STRIDE]), np.ascontiguousarray(sz)


def tile_streams(pk, sz, n):
    reps = (n + pk.shape[0] - 1) // pk.shape[0]
    return np.tile(pk, (reps, 1, 1))[:n], np.tile(sz, (reps, 1))[:n]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.p = None
        self.th = None
        self.rows = []

    def _nvml_loop(self):
        import pynvml as nv
        h = nv.nvmlDeviceGetHandleByIndex(self._phys_index())
        bits = (('hw_slowdown', nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, 'nvmlClocksEventReasonHwSlowdown') else 0x8),
                ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((sm, mx, [n for n, b in bits if r & b]))
            except Exception:
                pass
            self._stop.wait(0.01)

    def _phys_index(self):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        if vis:
            try:
                return int(vis.split(',')[self.idx])
            except Exception:
                pass
        return self.idx

    def start(self):
        # NVML in a thread (10 ms period: even a 0.2 s timed region gets ~20 samples); nvidia-smi -lms as the fallback
        try:
            import pynvml as nv
            nv.nvmlInit()
            self._stop = threading.Event()
            self.th = threading.Thread(target=self._nvml_loop, daemon=True)
            self.th.start()
            return
        except Exception:
            self.th = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                       '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.th is not None:
            self._stop.set()
            self.th.join(timeout=2)
            sm = sorted(r[0] for r in self.rows)
            reasons = sorted({x for r in self.rows for x in r[2]})
            return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max((r[1] for r in self.rows), default=None), 'reasons': reasons,
                    'samples': len(sm), 'source': 'nvml'}
        if not self.p:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            out = ''
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm_sorted = sorted(sm)
        med = sm_sorted[len(sm_sorted) // 2] if sm_sorted else None
        return {'sm_mhz': med, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons), 'samples': len(sm), 'source': 'nvidia-smi'}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference decoder on the host cores
# ---------------------------------------------------------------------------------------------
def host_cores():
    """Usable host cores: the scheduler affinity mask, further limited by a cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        q, per = open('/sys/fs/cgroup/cpu.max').read().split()
        if q != 'max':
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return n


def load_cpu_decoder():
    """(lib, kind, build): the compiled reference if it travelled (oracle/_ref: the AVX2-presumed build of BASELINE.md section 4 when the
    host has AVX2, else the RTCD build), else the oracle port."""
    d = os.path.join(ROOT, 'oracle', '_ref')
    for name, ok in (('libopus_ref_avx2.so', host_has_avx2()), ('libopus_ref_simd.so', True)):
        ref = os.path.join(d, name)
        if ok and os.path.exists(ref):
            L = ctypes.CDLL(ref)
            L.refh_decode_bench.restype = ctypes.c_long
            return L, 'reference', name
    L = ctypes.CDLL(os.path.join(ROOT, 'oracle', 'liboracle.so'))
    return L, 'port', 'liboracle.so'


def cpu_decode_sample(L, kind, pk, sz, n_streams, reps, threads):
    """Decode n_streams x 8 x reps frames spread over `threads` host threads (ctypes releases the
    GIL; the reference has no internal threading, so one pinned-by-the-OS thread per core is its
    best case).  Returns (frames, seconds); per-thread input slices are prepared outside the timing."""
    T = pk.shape[1]
    per = max(1, n_streams // threads)
    chunks = [(i * per, min(n_streams, (i + 1) * per)) for i in range(threads) if i * per < n_streams]
    done = [0] * len(chunks)
    bufs = [(np.ascontiguousarray(pk[lo:hi]), np.ascontiguousarray(sz[lo:hi]), np.zeros(FRAME * CH, np.int16))
            for lo, hi in chunks]
    gate = threading.Barrier(len(chunks) + 1)

    def work(k):
        p, s, pcm = bufs[k]
        gate.wait()
        if kind == 'reference':
            r = L.refh_decode_bench(48000, CH, FRAME, p.shape[0], T, p.ctypes.data_as(ctypes.c_void_p), p.shape[2],
                                    s.ctypes.data_as(ctypes.c_void_p), reps, pcm.ctypes.data_as(ctypes.c_void_p))
            done[k] = int(r)
        else:
            rng = np.zeros(T, np.uint32)
            out = np.zeros((T * FRAME, CH), np.float32)
            for _ in range(reps):
                for i in range(p.shape[0]):
                    L.orc_decode_stream_float(CH, FRAME, T, p[i].ctypes.data_as(ctypes.c_void_p), p.shape[2],
                                              s[i].ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p),
                                              rng.ctypes.data_as(ctypes.c_void_p))
                    done[k] += T

    th = [threading.Thread(target=work, args=(k,)) for k in range(len(chunks))]
    for t in th:
        t.start()
    gate.wait()
    t0 = time.perf_counter()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    assert all(d > 0 for d in done), done
    return sum(done), dt


# ---------------------------------------------------------------------------------------------
# encode workload (BASELINE.json configs[2]): synthetic PCM and the CPU arm
# ---------------------------------------------------------------------------------------------
ENC_BASE, ENC_T, ENC_BITRATE, ENC_CX, ENC_MAXB = 256, 32, 64000, 10, 1276
APP_RESTRICTED_CELT = 2053
# algorithmic HBM bytes per stream-frame of the encoder kernels (DESIGN.md "Roofline bookkeeping")
ENC_BYTES_KERNEL = {
    'ob_opus_enc_pre': 2 * 960 * 2 + 2 * 960 * 4 + 480 * 4 + 64,                    # int16 PCM in, dc-rejected float PCM out, 24 kHz analysis samples out, state
    'ob_opus_enc_analysis': 480 * 4 + 2 * (3 * 240 * 4 + 720 * 4) + 2 * 64 + 96,      # new samples in, phase + buffer state in/out, info ring slot in/out, ctl out
    'ob_celt_enc_chains': 2 * 960 * 4 + 64 + 64,                                      # PCM in, recurrence state in/out, frame info out
    'ob_celt_enc_analysis': 2 * 960 * 4 + 2 * 2 * 1024 * 4 // 8 * 8 + 2 * 960 * 4 + 3 * 168 + 96,   # PCM in, pre-filter history in + frame out, X out, energies, ctl
    'ob_celt_enc_frame_rdo': 2 * 960 * 4 + 3 * 168 + 2 * 1100 + 96 + 170,            # X in, energies in, ObEncState in/out, ctl in, packet out
}
ENC_BYTES_WHOLE_FRAME = 2 * 960 * 2 + 160                                            # int16 PCM in + ~160 B packet out


def synth_pcm(n_base=ENC_BASE, T=ENC_T, seed=0x0B200):
    """[n_base][T][1920] int16: 48 kHz stereo, T consecutive 20 ms frames per stream.  Four families: drifting
    harmonic partials + transients, brown noise with a 440 Hz AM tone and a partly anti-phase 3 kHz tone, white noise at -20 dBFS,
    a pulse train through resonances."""
    rng = np.random.default_rng(seed)
    ns = 960 * T
    t = np.arange(ns) / 48000.0
    out = np.zeros((n_base, ns, 2))
    for s in range(n_base):
        k = s & 3
        if k == 0:
            x = 150 * rng.standard_normal((ns, 2))
            for _ in range(5):
                f0, a, vib = rng.uniform(80, 5000), rng.uniform(400, 4000), rng.uniform(0.3, 6)
                am = 0.6 + 0.4 * np.sin(2 * np.pi * rng.uniform(0.5, 6) * t + rng.uniform(0, 6))
                ph = 2 * np.pi * f0 * (t + 0.002 * np.sin(2 * np.pi * vib * t))
                x[:, 0] += a * am * np.sin(ph + rng.uniform(0, 6))
                x[:, 1] += a * am * np.sin(ph + rng.uniform(0, 6))
            for _ in range(T // 6):
                p0 = rng.integers(0, ns - 256)
                x[p0:p0 + 256] += rng.standard_normal((256, 2)) * rng.uniform(2000, 9000) * np.exp(-np.arange(256) / 48.0)[:, None]
        elif k == 1:
            b = np.cumsum(rng.standard_normal((ns, 2)), axis=0)
            b -= np.linspace(b[0], b[-1], ns)
            x = 6000 * b / (np.abs(b).max() + 1e-9)
            x += (5000 * (1 + 0.5 * np.sin(2 * np.pi * 3 * t)) * np.sin(2 * np.pi * 440 * t))[:, None]
            x += 2500 * np.sin(2 * np.pi * 3000 * t[:, None] + np.array([0, 2.5])[None, :])
        elif k == 2:
            c = rng.standard_normal((ns, 1))
            x = 3276.8 * (0.6 * c + 0.8 * rng.standard_normal((ns, 2)))
        else:
            f0 = rng.uniform(90, 240)
            pulses = (np.floor(t * f0) != np.floor((t - 1 / 48000.0) * f0)).astype(np.float64)
            y = pulses * 3000
            from scipy.signal import lfilter
            for fc in rng.uniform(300, 3500, 3):
                r, w = 0.985, 2 * np.pi * fc / 48000.0
                y = 0.5 * y + 0.12 * lfilter([1.0], [1.0, -2 * r * np.cos(w), r * r], y)
            x = np.stack([y, 0.9 * y], 1) + 60 * rng.standard_normal((ns, 2))
        out[s] = x
    return np.clip(out, -32768, 32767).astype(np.int16).reshape(n_base, T, 1920)


def mix_plan(gid):
    """Every stream of the batch is its own signal: stream gid (global id) = w * base[gid % 256] + (1 - w) * base[partner], the partner and the weight
    depend on the copy number gid // 256 (copy 0 is the base stream itself).  Without this the 65 536 streams would be 256 copies of 256 signals, and
    anything that groups streams by behaviour (the frame coder's frame-type binning) would meet identical streams in a warp.  Works on numpy arrays
    and torch tensors of integer ids; returns (a, b, w)."""
    a = gid % ENC_BASE
    c = gid // ENC_BASE
    b = (a + 1 + (c % (ENC_BASE - 1))) % ENC_BASE
    w = 0.60 + 0.35 * (((c * 89) % 256) / 256.0)
    w = w * (c > 0) + 1.0 * (c == 0)
    return a, b, w


def mix_streams_np(base, gids):
    """base [256][T][1920] int16 -> [len(gids)][T][1920] int16"""
    a, b, w = mix_plan(np.asarray(gids))
    x = base[a].astype(np.float32) * w[:, None, None].astype(np.float32) + base[b].astype(np.float32) * (1 - w)[:, None, None].astype(np.float32)
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)


def host_has_avx2():
    try:
        return 'avx2' in open('/proc/cpuinfo').read() and ' fma' in open('/proc/cpuinfo').read()
    except Exception:
        return False


def load_cpu_encoder():
    """(lib, kind, build): the compiled reference (oracle/_ref) — the AVX2-presumed build when the host can run it — else the oracle port"""
    d = os.path.join(ROOT, 'oracle', '_ref')
    for name, ok in (('libopus_ref_avx2.so', host_has_avx2()), ('libopus_ref_simd.so', True)):
        p = os.path.join(d, name)
        if ok and os.path.exists(p):
            L = ctypes.CDLL(p)
            L.refh_encode_bench.restype = ctypes.c_long
            return L, 'reference', name
    return ctypes.CDLL(os.path.join(ROOT, 'oracle', 'liboracle.so')), 'port', 'liboracle.so'


def cpu_encode_sample(L, kind, pcm, n_streams, reps, threads):
    """pcm [n_streams][T][1920] int16; one encoder per stream, T x reps consecutive frames each, streams spread over `threads` host
    threads (ctypes releases the GIL).  Returns (frames, seconds)."""
    T = pcm.shape[1]
    per = max(1, n_streams // threads)
    chunks = [(i * per, min(n_streams, (i + 1) * per)) for i in range(threads) if i * per < n_streams]
    bufs = [np.ascontiguousarray(pcm[lo:hi]) for lo, hi in chunks]
    done = [0] * len(chunks)
    gate = threading.Barrier(len(chunks) + 1)

    def work(k):
        p = bufs[k]
        gate.wait()
        if kind == 'reference':
            done[k] = int(L.refh_encode_bench(48000, CH, APP_RESTRICTED_CELT, ENC_BITRATE, 1, ENC_CX, FRAME, p.shape[0], T,
                                              p.ctypes.data_as(ctypes.c_void_p), reps, ENC_MAXB, None))
        else:
            out = np.zeros((T * reps, ENC_MAXB), np.uint8)
            sizes, rngs = np.zeros(T * reps, np.int32), np.zeros(T * reps, np.uint32)
            for i in range(p.shape[0]):
                x = np.ascontiguousarray(np.tile(p[i], (reps, 1)))
                r = L.orc_opus_encode_stream(48000, CH, APP_RESTRICTED_CELT, ENC_BITRATE, 1, ENC_CX, FRAME, T * reps, x.ctypes.data_as(ctypes.c_void_p), 1,
                                             out.ctypes.data_as(ctypes.c_void_p), ENC_MAXB, sizes.ctypes.data_as(ctypes.c_void_p),
                                             rngs.ctypes.data_as(ctypes.c_void_p), None, 0)
                assert r == 0, r
                done[k] += T * reps

    th = [threading.Thread(target=work, args=(k,)) for k in range(len(chunks))]
    for t in th:
        t.start()
    gate.wait()
    t0 = time.perf_counter()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    assert all(d > 0 for d in done), done
    return sum(done), dt


def cpu_encode_baseline(steps, warmup, streams_per_core=64, reps=4):
    """the reference's opus_encode on every host core: a bounded sample of the same workload (one step = streams_per_core streams per core
    x 32 frames x reps passes); returns the median step (the CPU arm is noisy on shared hosts) and its description"""
    L, kind, build = load_cpu_encoder()
    cores = host_cores()
    n = streams_per_core * cores
    pcm = mix_streams_np(synth_pcm(), np.arange(n))   # the first n streams of the GPU arm's batch
    for _ in range(warmup):
        cpu_encode_sample(L, kind, pcm, n, 1, cores)
    vals, tot_t = [], 0.0
    for _ in range(steps):
        f, dt = cpu_encode_sample(L, kind, pcm, n, reps, cores)
        vals.append(f / dt)
        tot_t += dt
    val = float(np.median(vals))
    sample = '%d streams x %d frames x %d passes per step on %d threads, median of %d steps (%s)' % (n, ENC_T, reps, cores, steps, build)
    return {'value': val, 'unit': 'frames/s', 'cores': cores, 'kind': kind, 'sample': sample, 'build': build,
            'min': float(min(vals)), 'max': float(max(vals))}, tot_t / max(1, steps)


_REAL_STDOUT_FD = None


def print_result(obj):
    """the one JSON line of the run, on the real stdout"""
    sys.stdout.flush()
    if _REAL_STDOUT_FD is not None:
        os.dup2(_REAL_STDOUT_FD, 1)
    sys.stdout.write(json.dumps(obj) + '\n')
    sys.stdout.flush()
    if _REAL_STDOUT_FD is not None:
        os.dup2(2, 1)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    pk, sz = load_fixture()
    L, kind, build = load_cpu_decoder()
    cores = host_cores()
    # one step = a bounded sample: (64 streams per core) x 8 frames x 4 passes
    n_streams = 64 * cores
    reps = 4
    pkt, szt = tile_streams(pk, sz, n_streams)
    for _ in range(args.warmup):
        cpu_decode_sample(L, kind, pkt, szt, n_streams, 1, cores)
    vals, t = [], 0.0
    for _ in range(args.steps):
        f, dt = cpu_decode_sample(L, kind, pkt, szt, n_streams, reps, cores)
        vals.append(f / dt)
        t += dt
    val = float(np.median(vals))   # the CPU arm swings on shared hosts: the median step is reported, min / max beside it
    sample = '%d streams x 8 frames x %d passes per step on %d threads, median of %d steps (%s)' % (n_streams, reps, cores, args.steps, build)
    enc_block = None
    if args.workload in ('both', 'encode'):
        cb, step_s = cpu_encode_baseline(max(3, min(args.steps, 7)), min(args.warmup, 2))
        enc_block = {'metric': 'celt_encode_frames_per_s', 'value': cb['value'], 'unit': 'frames/s', 'x_realtime': cb['value'] / 50.0,
                     'ms_per_step': 1000.0 * step_s, 'higher_is_better': True, 'dtype': 'f32', 'data': 'synthetic',
                     'config': {'workload': 'batched CELT encode via opus_encode, RESTRICTED_CELT, complexity 10, 48 kHz stereo 20 ms @64 kb/s VBR (BASELINE.json configs[2])',
                                'sample': cb['sample'], 'cpu': cb['kind']},
                     'cpu_baseline': cb, 'e2e': {'value': cb['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print_result({
        'impl': 'reference', 'metric': 'celt_decode_frames_per_s', 'value': val, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 * t / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'x_realtime': val / 50.0,
        'config': {'workload': 'batched CELT decode, 48 kHz stereo 20 ms @128 kb/s CBR (BASELINE.json configs[1])',
                   'sample': sample, 'cpu': kind},
        'cpu_baseline': {'value': val, 'unit': 'frames/s', 'cores': cores, 'kind': kind, 'sample': sample, 'build': build,
                         'min': float(min(vals)), 'max': float(max(vals))},
        'e2e': {'value': val, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        **({'encode': enc_block} if enc_block else {}),
    })


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu_decode(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import opus_b200

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    n = args.streams
    # the one init-time collective: rank 0 owns the workload fixture and broadcasts it
    from opus_b200 import parallel
    pk, sz = load_fixture() if rank == 0 else (None, None)
    t_pk, t_sz = parallel.broadcast_fixture(pk, sz, (256, 8, PKT_STRIDE), (256, 8), dev, rank, world)
    T = t_pk.shape[1]
    # [T][n][320] resident packets (168 MB at n = 65536), [T][n] lengths: this rank's streams replay fixture row (global id mod 256)
    rows = parallel.fixture_rows(rank, world, n, t_pk.shape[0]).to(dev)
    d_pk = t_pk.permute(1, 0, 2)[:, rows].contiguous()
    d_len = t_sz.permute(1, 0)[:, rows].contiguous()
    if args.loss > 0:   # non-default: a fraction of the packets is lost (len = 0) and concealed; never used for the headline number
        g = torch.Generator(device='cpu').manual_seed(1234 + rank)
        drop = (torch.rand(d_len.shape, generator=g) < args.loss).to(dev)
        drop[0] = False   # every stream decodes one packet first
        d_len = torch.where(drop, torch.zeros_like(d_len), d_len).contiguous()
    d_pcm = torch.empty((n, FRAME * CH), dtype=torch.int16, device=dev)
    d_ns = torch.empty((n,), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    dec = opus_b200.DecoderBatch(n, channels=CH, device=local_rank)
    ext = torch.cuda.ExternalStream(dec.cuda_stream, device=dev)

    def step_device(i):
        t = i % T
        dec.decode_device(d_pk[t].data_ptr(), PKT_STRIDE, d_len[t].data_ptr(), d_pcm.data_ptr(), FRAME * CH, int16=True,
                          frame_size=FRAME, max_frames_per_packet=1, d_out_samples=d_ns.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ----
    # pipelined calls: frame t's synthesis (own stream) overlaps frame t+1's range decoding; join() puts the last synthesis
    # back in front of the closing event, so the timed region holds every kernel of the K steps
    dec.set_pipelined(True)
    for i in range(args.warmup):
        step_device(i)
    dec.sync()
    assert int(d_ns.min()) == FRAME and int(d_ns.max()) == FRAME, (int(d_ns.min()), int(d_ns.max()))
    launches0 = dec.launch_count
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for i in range(args.steps):
        step_device(args.warmup + i)
    dec.join()
    e1.record(ext)
    dec.sync()
    barrier()
    ms = e0.elapsed_time(e1)
    dec.set_pipelined(False)
    clk = clocks.stop() if rank == 0 else None
    launches = dec.launch_count - launches0
    ms_max = parallel.max_over_ranks(ms, dev, world)
    value = parallel.aggregate_throughput(n, world, args.steps, ms_max)

    # ---- end to end through the host-buffer C-ABI call (pinned host buffers) ----
    L = opus_b200.lib()
    hp = L.opusb200_host_alloc(n * PKT_STRIDE * T)
    ho = [L.opusb200_host_alloc(n * FRAME * CH * 2) for _ in range(2)]
    h_pk = np.ctypeslib.as_array((ctypes.c_uint8 * (n * PKT_STRIDE * T)).from_address(hp)).reshape(T, n, PKT_STRIDE)
    h_pcm = [np.ctypeslib.as_array((ctypes.c_int16 * (n * FRAME * CH)).from_address(a)).reshape(n, FRAME * CH) for a in ho]
    h_ns = [np.zeros(n, np.int32) for _ in range(2)]
    h_pk[:] = d_pk.cpu().numpy()
    h_len = np.ascontiguousarray(d_len.cpu().numpy())
    e2e_steps = max(3, min(args.steps, 10))
    # The batch-server form of the host-buffer call: submit step k+1 while step k's PCM is still on its way back
    # (opusb200_decode_submit / _wait; every step's H2D and D2H are inside the timed region, which ends with the wait).
    for i in range(2):
        dec.submit(h_pk[i % T], h_len[i % T], h_pcm[i & 1], h_ns[i & 1], frame_size=FRAME)
    dec.wait()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(ext)
    for i in range(e2e_steps):
        dec.submit(h_pk[i % T], h_len[i % T], h_pcm[i & 1], h_ns[i & 1], frame_size=FRAME)
    dec.wait()
    f1.record(ext)
    dec.sync()
    barrier()
    assert (h_ns[0] == FRAME).all() and (h_ns[1] == FRAME).all()
    # the synchronous call (opusb200_decode: submit + wait per step), for comparison
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(ext)
    for i in range(e2e_steps):
        dec.decode(h_pk[i % T], h_len[i % T], frame_size=FRAME, dtype=np.int16, out=h_pcm[0])
    g1.record(ext)
    dec.sync()
    e2e_sync_val = n * e2e_steps / (g0.elapsed_time(g1) / 1000.0)
    e2e_val = parallel.aggregate_throughput(n, world, e2e_steps, parallel.max_over_ranks(f0.elapsed_time(f1), dev, world))
    h2d = n * PKT_STRIDE + n * 4
    d2h = n * FRAME * CH * 2 + n * 4

    # ---- per-kernel device time (CUDA events around each launch, on the launching stream) ----
    dec.set_timing(True)
    prof_steps = 5
    for i in range(prof_steps):
        step_device(i)
    kms, nfr = dec.get_timing()
    dec.set_timing(False)
    L.opusb200_host_free(hp)
    for a in ho:
        L.opusb200_host_free(a)

    dec.close()
    if rank != 0:
        return None
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    per = {}
    for short, name in (('entropy', 'ob_celt_dec_entropy'), ('pvq', 'ob_celt_dec_pvq'), ('synth', 'ob_celt_dec_synth')):
        t = kms[short] / nfr / 1000.0
        gbs = BYTES_KERNEL[name] * n / t / 1e9
        per[name] = {'ms_per_launch': t * 1000.0, 'achieved_GBs': gbs, 'frac': gbs / peak,
                     'algorithmic_bytes_per_launch': BYTES_KERNEL[name] * n}
    dom = max(per, key=lambda k: per[k]['ms_per_launch'])
    roof = {'kernel': dom, 'bound': 'hbm', 'achieved': per[dom]['achieved_GBs'], 'peak': peak, 'unit': 'GB/s',
            'frac': per[dom]['frac'], 'traffic': None, 'peak_source': peak_src,
            'algorithmic_bytes_per_launch': per[dom]['algorithmic_bytes_per_launch'],
            'ms_per_launch': per[dom]['ms_per_launch'],
            'note': 'entropy/PVQ kernels are latency- and issue-bound (serial range decoding), not HBM-bound; '
                    'the HBM fraction is reported as the contract requires, see DESIGN.md section 4',
            'all_kernels': per}
    prof = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(prof):
        try:
            roof['traffic'] = json.load(open(prof)).get(dom + '_dram_bytes_per_launch')
        except Exception:
            pass
    # ---- CPU baseline: bounded sample on the host cores ----
    Lc, kind, build = load_cpu_decoder()
    cores = host_cores()
    pkf, szf = load_fixture()
    ns_cpu = 64 * cores
    pkt, szt = tile_streams(pkf, szf, ns_cpu)
    cpu_decode_sample(Lc, kind, pkt, szt, ns_cpu, 1, cores)
    reps_cpu = 4
    cvals = []
    for _ in range(3):
        frames, dt = cpu_decode_sample(Lc, kind, pkt, szt, ns_cpu, reps_cpu, cores)
        cvals.append(frames / dt)
    cpu = {'value': float(np.median(cvals)), 'unit': 'frames/s', 'cores': cores, 'kind': kind, 'build': build,
           'sample': '%d streams x 8 frames x %d passes on %d threads, median of 3 (%s)' % (ns_cpu, reps_cpu, cores, build)}
    return ({
        'metric': 'celt_decode_frames_per_s', 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'x_realtime': value / 50.0,
        'config': {'workload': 'batched CELT decode, 48 kHz stereo 20 ms @128 kb/s CBR (BASELINE.json configs[1])',
                   'streams_per_gpu': n, 'packets': '256 distinct streams x 8 frames from the reference encoder, tiled',
                   'pipelining': 'consecutive steps are consecutive frames of the same streams; the synthesis kernel of step t runs on its own CUDA stream under the range decoding of step t+1 (joined before the closing event)',
                   'l2': 'per-step working set (spectrum 503 MB + history 503 MB at 65536 streams) exceeds the 126 MB L2',
                   'parallelism': 'streams sharded across GPUs, no steady-state collective',
                   **({'packet_loss': args.loss} if args.loss > 0 else {})},
        'e2e': {'value': e2e_val, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': e2e_steps,
                'call': 'opusb200_decode_submit/_wait (double-buffered)', 'sync_call_value_rank0': e2e_sync_val},
        'gpu_launches': launches, 'clocks': clk, 'roofline': roof, 'cpu_baseline': cpu,
    })


def run_gpu_encode(args, rank, world, local_rank):
    """BASELINE.json configs[2] through the Opus API of the library (opusb200_opus_encode*): returns the `encode` block (rank 0) or None"""
    import torch
    import torch.distributed as dist
    import opus_b200
    from opus_b200 import parallel

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    n, T = args.streams, ENC_T
    # the init-time collective: rank 0 synthesises the PCM fixture and broadcasts it
    if rank == 0:
        t_pcm = torch.from_numpy(synth_pcm()).to(dev)
    else:
        t_pcm = torch.empty((ENC_BASE, T, FRAME * CH), dtype=torch.int16, device=dev)
    if world > 1:
        dist.broadcast(t_pcm.view(torch.uint8), 0)   # (NCCL has no int16: the bytes travel)
    rows = parallel.fixture_rows(rank, world, n, ENC_BASE).to(dev)
    # experiments on SIMT divergence (never the reported configuration): 'family' puts streams of the same signal family in a warp,
    # 'same' gives all 32 lanes of a warp the same stream (the convergence ceiling)
    if args.enc_group == 'family':
        rows = rows[torch.argsort(rows % 4, stable=True)]
    elif args.enc_group == 'same':
        rows = (torch.arange(n, device=dev) // 32) % ENC_BASE
    if args.enc_group == 'none':   # every stream its own mixture of two base signals (mix_plan), built frame by frame on the device
        gid = torch.arange(n, device=dev, dtype=torch.int64) + rank * n
        ia, ib, w = mix_plan(gid)
        w = w.to(torch.float32)[:, None]
        d_pcm = torch.empty((T, n, FRAME * CH), dtype=torch.int16, device=dev)      # 8 GB at n = 65536
        for t in range(T):
            x = t_pcm[ia, t].to(torch.float32) * w + t_pcm[ib, t].to(torch.float32) * (1 - w)
            d_pcm[t] = torch.clamp(torch.round(x), -32768, 32767).to(torch.int16)
        del x
    else:
        d_pcm = t_pcm.permute(1, 0, 2)[:, rows].contiguous()          # [T][n][1920] int16
    d_pkt = torch.zeros((n, ENC_MAXB), dtype=torch.uint8, device=dev)
    d_nb = torch.zeros((n,), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    enc = opus_b200.OpusEncoderBatch(n, channels=CH, fs=48000, device=local_rank, bitrate=ENC_BITRATE, vbr=True, complexity=ENC_CX)
    ext = torch.cuda.ExternalStream(enc.cuda_stream, device=dev)

    def step_device(i):
        enc.encode_device(d_pcm[i % T].data_ptr(), True, FRAME, d_pkt.data_ptr(), ENC_MAXB, ENC_MAXB, d_nb.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(args.warmup, args.enc_warm_frames)   # >= 50 frames per stream: VBR reservoir, pre-filter and analysis state are warm
    for i in range(warm):
        step_device(i)
    enc.sync()
    nb = d_nb.cpu().numpy()
    assert nb.min() > 2 and nb.max() <= ENC_MAXB, (int(nb.min()), int(nb.max()))
    launches0 = enc.launch_count
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    bytes_sum = 0
    for i in range(args.steps):
        step_device(warm + i)
    e1.record(ext)
    enc.sync()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    launches = enc.launch_count - launches0
    ms_max = parallel.max_over_ranks(ms, dev, world)
    value = parallel.aggregate_throughput(n, world, args.steps, ms_max)
    mean_bytes = float(d_nb.float().mean().item())

    # ---- end to end: pinned HOST int16 PCM in and HOST packets out, every step (opusb200_opus_encode_submit / _wait: the upload of step
    # k+1 and the read-back of step k-1 run under the kernels of step k; every step's H2D and D2H are inside the timed region) ----
    L = opus_b200.lib()
    e2e_steps = max(3, min(args.steps, 8))
    hp = L.opusb200_host_alloc(n * FRAME * CH * 2 * e2e_steps)
    hk = [L.opusb200_host_alloc(n * ENC_MAXB) for _ in range(2)]
    h_pcm = np.ctypeslib.as_array((ctypes.c_int16 * (n * FRAME * CH * e2e_steps)).from_address(hp)).reshape(e2e_steps, n, FRAME * CH)
    h_pkt = [np.ctypeslib.as_array((ctypes.c_uint8 * (n * ENC_MAXB)).from_address(a)).reshape(n, ENC_MAXB) for a in hk]
    h_nb = [np.zeros(n, np.int32) for _ in range(2)]
    for i in range(e2e_steps):
        h_pcm[i] = d_pcm[(warm + args.steps + i) % T].cpu().numpy()
    vp = ctypes.c_void_p
    for i in range(2):
        enc.submit(h_pcm[i], h_pkt[i & 1], h_nb[i & 1], ENC_MAXB, FRAME)
    enc.wait()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(ext)
    for i in range(e2e_steps):
        enc.submit(h_pcm[i], h_pkt[i & 1], h_nb[i & 1], ENC_MAXB, FRAME)
    enc.wait()   # returns when the last read-back is on the host; the closing event is recorded after it
    f1.record(ext)
    enc.sync()
    e2e_ms = f0.elapsed_time(f1)
    barrier()
    assert h_nb[0].min() > 2 and h_nb[1].min() > 2
    # the synchronous call (opusb200_opus_encode: upload, kernels, read-back one after the other), for comparison
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(ext)
    for i in range(min(3, e2e_steps)):
        r = L.opusb200_opus_encode(enc._h, vp(h_pcm[i].ctypes.data), FRAME, vp(hk[0]), ENC_MAXB, ENC_MAXB, vp(h_nb[0].ctypes.data))
        assert r == 0, r
    g1.record(ext)
    enc.sync()
    e2e_sync_val = n * min(3, e2e_steps) / (g0.elapsed_time(g1) / 1000.0)
    e2e_val = parallel.aggregate_throughput(n, world, e2e_steps, parallel.max_over_ranks(e2e_ms, dev, world))
    h2d = n * FRAME * CH * 2
    d2h = n * ENC_MAXB + n * 4
    L.opusb200_host_free(hp)
    for a in hk:
        L.opusb200_host_free(a)

    # ---- per-kernel device time (CUDA events around each launch, on the launching stream) ----
    enc.set_timing(True)
    for i in range(4):
        step_device(i)
    kms, calls = enc.get_timing()
    enc.set_timing(False)
    enc.close()
    if rank != 0:
        return None
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    names = {'opus_pre': 'ob_opus_enc_pre', 'opus_analysis': 'ob_opus_enc_analysis', 'celt_chains': 'ob_celt_enc_chains', 'celt_analysis': 'ob_celt_enc_analysis', 'celt_frame': 'ob_celt_enc_frame_rdo'}
    per = {}
    for short, name in names.items():
        t = kms[short] / max(1, calls) / 1000.0
        gbs = ENC_BYTES_KERNEL[name] * n / max(t, 1e-9) / 1e9
        per[name] = {'ms_per_launch': t * 1000.0, 'achieved_GBs': gbs, 'frac': gbs / peak, 'algorithmic_bytes_per_launch': ENC_BYTES_KERNEL[name] * n}
    dom = max(per, key=lambda k: per[k]['ms_per_launch'])
    roof = {'kernel': dom, 'bound': 'hbm', 'achieved': per[dom]['achieved_GBs'], 'peak': peak, 'unit': 'GB/s', 'frac': per[dom]['frac'], 'traffic': None,
            'peak_source': peak_src, 'algorithmic_bytes_per_launch': per[dom]['algorithmic_bytes_per_launch'], 'ms_per_launch': per[dom]['ms_per_launch'],
            'note': 'the encoder kernels are instruction-issue / latency bound (serial range coding, greedy PVQ search), not HBM bound; the HBM '
                    'fraction is reported as the contract requires, see DESIGN.md section 4',
            'all_kernels': per}
    prof = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(prof):
        try:
            roof['traffic'] = json.load(open(prof)).get(dom + '_dram_bytes_per_launch')
        except Exception:
            pass
    cb, _ = cpu_encode_baseline(3, 1)
    return {
        'metric': 'celt_encode_frames_per_s', 'value': value, 'unit': 'frames/s', 'x_realtime': value / 50.0, 'n_gpus': world, 'steps': args.steps,
        'warmup': warm, 'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'target_frames_per_s_per_gpu': 6.25e6,
        'config': {'workload': 'batched CELT encode via opus_encode, RESTRICTED_CELT, complexity 10, 48 kHz stereo 20 ms @64 kb/s VBR (BASELINE.json configs[2])',
                   'streams_per_gpu': n, 'pcm': '65 536 distinct int16 streams per GPU x 32 consecutive frames: stream g = w(g) x base[g % 256] + (1 - w(g)) x base[partner(g)] over 256 synthetic base signals (4 families); the CPU arm encodes the first streams of the same batch',
                   'warm_frames_per_stream': warm, 'mean_packet_bytes': mean_bytes,
                   'l2': 'per-step working set (PCM 252 MB + spectra 503 MB + analysis state 0.9 GB at 65536 streams) exceeds the 126 MB L2',
                   'parallelism': 'streams sharded across GPUs, no steady-state collective'},
        'e2e': {'value': e2e_val, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': e2e_steps,
                'call': 'opusb200_opus_encode_submit/_wait (host int16 PCM in, host packets out, double-buffered)',
                'sync_call_value_rank0': e2e_sync_val},
        'gpu_launches': launches, 'clocks': clk, 'roofline': roof, 'cpu_baseline': cb,
    }


def run_mdct_kernel(local_rank, n_transforms=131072, reps=20):
    """The one kernel of the path that IS bound by HBM (the standalone transform): the long-block forward MDCT
    (ob_mdct_forward_tma), timed with CUDA events on the library stream; algorithmic bytes = 4320 in + 3840 out per transform, working set 1.07 GB."""
    import torch
    import opus_b200
    dev = torch.device('cuda', local_rank)
    ops = opus_b200.Ops(local_rank)
    L = ops._L
    stream = torch.cuda.ExternalStream(L.opusb200_ctx_cuda_stream(ops._h), device=dev)
    x = torch.randn((n_transforms, 1080), dtype=torch.float32, device=dev)
    o = torch.empty((n_transforms, 960), dtype=torch.float32, device=dev)
    fn = lambda: L.opusb200_mdct_forward(ops._h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(o.data_ptr()), n_transforms, 3, 0, 1)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for _ in range(3):
            assert fn() == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        stream.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ops.close()
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    peak = float(json.load(open(peaks_path))['hbm_gbs']) if os.path.exists(peaks_path) else 6650.0
    gbs = n_transforms * 8160 / (ms * 1e-3) / 1e9
    return {'kernel': 'ob_mdct_forward_tma', 'bound': 'hbm', 'what': 'clt_mdct_forward, 960-point long block, %d transforms per launch (65 536 stereo frames)' % n_transforms,
            'ms_per_launch': ms, 'algorithmic_bytes_per_launch': n_transforms * 8160, 'achieved': gbs, 'peak': peak, 'unit': 'GB/s', 'frac': gbs / peak,
            'launches_timed': reps, 'l2': 'working set 1.07 GB per launch exceeds the 126 MB L2'}


def run_gpu_arm(args, rank, world, local_rank):
    line = None
    if args.workload in ('both', 'decode'):
        line = run_gpu_decode(args, rank, world, local_rank)
    enc = run_gpu_encode(args, rank, world, local_rank) if args.workload in ('both', 'encode') else None
    if rank != 0:
        return
    if line is not None:
        try:
            line['mdct_kernel_roofline'] = run_mdct_kernel(local_rank)
        except Exception as ex:   # diagnostic block: never take the bench line down
            line['mdct_kernel_roofline'] = {'error': repr(ex)}
    if line is None:   # --workload encode: the encode block is the line
        line = enc
    elif enc is not None:
        line['encode'] = enc
        line['gpu_launches_decode'] = line['gpu_launches']
        line['gpu_launches'] = line['gpu_launches'] + enc['gpu_launches']
    print_result(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--streams', type=int, default=65536, help='streams per GPU')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--loss', type=float, default=0.0, help='fraction of packets lost and concealed (diagnostic runs only)')
    ap.add_argument('--workload', default='both', choices=['both', 'decode', 'encode'])
    ap.add_argument('--enc-group', default='none', choices=['none', 'family', 'same'], help='experiment: how fixture streams are laid out over warps')
    ap.add_argument('--enc-warm-frames', type=int, default=50, help='frames every stream encodes before the timed region of the encode workload')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    # stdout carries exactly ONE line, the JSON result: whatever libraries print while the job runs (NCCL's version banner comes from C code, past
    # sys.stdout) is sent to stderr by pointing file descriptor 1 there; print_result() points it back for the one line
    global _REAL_STDOUT_FD
    sys.stdout.flush()
    _REAL_STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == 'reference':
        run_reference_arm(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'   # the version banner goes to stdout, which must hold exactly one JSON line
        dist.init_process_group('nccl')
    try:
        run_gpu_arm(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
