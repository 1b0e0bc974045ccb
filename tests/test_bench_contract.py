"""bench.py's contract, the part that runs without a GPU: the reference arm prints exactly ONE line on stdout, a JSON object with the keys the driver
reads (both workloads), and the encode fixture gives every stream of the batch its own signal."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import util

ROOT = util.ROOT


@pytest.mark.skipif(not util.ref_available('simd') and not os.path.exists(os.path.join(ROOT, 'oracle', 'liboracle.so')), reason='no CPU arm built')
def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'], capture_output=True, text=True,
                       timeout=900, env=dict(os.environ, NCCL_DEBUG='VERSION'))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'celt_decode_frames_per_s' and d['unit'] == 'frames/s' and d['higher_is_better'] is True
    for k in ('value', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0 and d['e2e']['value'] == d['value']
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1
    e = d['encode']
    assert e['metric'] == 'celt_encode_frames_per_s' and e['value'] > 0 and e['cpu_baseline']['value'] == e['value']


def test_every_encode_stream_is_its_own_signal():
    sys.path.insert(0, ROOT)
    import bench
    a, b, w = bench.mix_plan(np.arange(65536 * 2))
    assert (a[:256] == np.arange(256)).all() and (w[:256] == 1.0).all()          # the first 256 streams are the base signals themselves
    assert (a != b).all()                                                        # a stream is never mixed with itself
    keys = set(zip(a.tolist(), b.tolist(), np.round(w, 6).tolist()))
    assert len(keys) == 65536 * 2                                                # no two streams of a 2-GPU job share (base, partner, weight)
    base = np.random.default_rng(0).integers(-9000, 9000, (256, 2, 1920)).astype(np.int16)
    x = bench.mix_streams_np(base, np.array([0, 256, 512, 70000]))
    assert (x[0] == base[0]).all() and not (x[1] == base[0]).all() and not (x[1] == x[2]).all()
