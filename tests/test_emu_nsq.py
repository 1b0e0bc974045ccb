"""The NSQ kernels (opus_b200/csrc/silk_nsq.cuh) run on the host SIMT emulator against the pinned oracle."""
import ctypes

import numpy as np
import pytest

import emu_util
import util

CONFIGS = [
    (16, 4, 16, 24, 15729, 4),
    (16, 4, 16, 16, 0, 2),
    (16, 4, 16, 20, 15729, 3),
    (12, 4, 10, 14, 11796, 2),
    (8, 2, 10, 16, 0, 4),
    (16, 4, 16, 16, 15729, 1),
    (16, 4, 16, 12, 0, 2),
    (16, 2, 16, 18, 15729, 4),
    (16, 4, 16, 16, 0, 1),
    (8, 4, 10, 12, 0, 1),
]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def oracle_run(O, cfg, x, frames):
    """x [n, nf, fl], frames [n, nf] -> pulses, seeds, state exports (after the last frame)"""
    n, nf, fl = x.shape
    pulses = np.zeros(x.shape, np.int8)
    seeds = np.zeros((n, nf), np.int32)
    exp = np.zeros((n, util.NSQ_STATE_EXPORT_WORDS), np.int32)
    for s in range(n):
        st = np.zeros(O.orc_nsq_state_size(), np.uint8)
        O.orc_nsq_state_init(_p(st))
        for f in range(nf):
            one = frames[s, f:f + 1].copy()
            O.orc_silk_nsq(_p(cfg), _p(st), _p(one), _p(x[s, f]), _p(pulses[s, f]))
            seeds[s, f] = one['Seed'][0]
        O.orc_nsq_state_export(_p(st), _p(exp[s]))
    return pulses, seeds, exp


def make_batch(seed, cfg, n, nf):
    xs, fs = [], []
    for s in range(n):
        x, fr = util.make_nsq_frames(seed * 7919 + s, cfg, nf)
        xs.append(x)
        fs.append(fr)
    return np.stack(xs), np.stack(fs)


@pytest.mark.parametrize('c', CONFIGS)
def test_emu_nsq(c):
    E, O = emu_util.load_emu_ops(), util.load_oracle()
    for L, names in ((E, ['emu_nsq_state_init', 'emu_silk_nsq', 'emu_nsq_state_export']),
                     (O, ['orc_nsq_state_init', 'orc_silk_nsq', 'orc_nsq_state_export'])):
        for nm in names:
            getattr(L, nm).restype = None
    cfg = util.nsq_cfg(*c)
    n, nf = 11, 3                      # 11 streams: one full warp of 8 + a ragged one
    x, frames = make_batch(c[0] + c[3] + c[5], cfg, n, nf)
    po, so, eo = oracle_run(O, cfg, x, frames)
    st = np.zeros(n * E.emu_nsq_state_size(), np.uint8)
    E.emu_nsq_state_init(_p(st), n)
    for f in range(nf):
        fr = np.ascontiguousarray(frames[:, f])
        xf = np.ascontiguousarray(x[:, f])
        pulses = np.zeros(xf.shape, np.int8)
        E.emu_silk_nsq(_p(cfg), _p(st), _p(fr), _p(xf), _p(pulses), n)
        assert np.array_equal(pulses, po[:, f]), (c, f, np.nonzero((pulses != po[:, f]).any(axis=1)))
        assert np.array_equal(fr['Seed'], so[:, f])
    ssz = E.emu_nsq_state_size()
    for s in range(n):
        e = np.zeros(util.NSQ_STATE_EXPORT_WORDS, np.int32)
        E.emu_nsq_state_export(_p(st[s * ssz:]), _p(e))
        assert np.array_equal(e, eo[s]), (c, s, np.nonzero(e != eo[s]))
