"""Pin the NSQ / NSQ_del_dec restatement (oracle/silk_nsq_oracle.c) to the reference's silk_NSQ / silk_NSQ_del_dec
(both the pure-C build and the RTCD/SIMD build) compiled under oracle/_ref."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built (needs /root/reference)')

CONFIGS = [
    # fs_khz nb_subfr predictLPCOrder shapingLPCOrder warping_Q16 nStates
    (16, 4, 16, 16, 0, 1),
    (16, 4, 16, 24, 15729, 4),
    (16, 4, 16, 20, 15729, 3),
    (16, 4, 16, 16, 0, 2),
    (16, 2, 16, 16, 15729, 4),
    (12, 4, 10, 14, 11796, 2),
    (8, 4, 10, 12, 0, 1),
    (8, 4, 10, 12, 7864, 1),
    (8, 2, 10, 16, 0, 4),
]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def run_ref(R, cfg, x, frames):
    st = np.zeros(R.refh_nsq_state_size(), np.uint8)
    R.refh_nsq_state_init(_p(st))
    fl = x.shape[1]
    pulses = np.zeros(x.shape, np.int8)
    seeds = np.zeros(len(frames), np.int32)
    exp = np.zeros((len(frames), util.NSQ_STATE_EXPORT_WORDS), np.int32)
    fr = frames.copy()
    for f in range(len(fr)):
        one = fr[f:f + 1]
        R.refh_silk_nsq(_p(cfg), _p(st), _p(one), _p(x[f]), _p(pulses[f]))
        seeds[f] = one['Seed'][0]
        R.refh_nsq_state_export(_p(st), _p(exp[f]))
    return pulses, seeds, exp


def run_oracle(O, cfg, x, frames):
    st = np.zeros(O.orc_nsq_state_size(), np.uint8)
    O.orc_nsq_state_init(_p(st))
    pulses = np.zeros(x.shape, np.int8)
    seeds = np.zeros(len(frames), np.int32)
    exp = np.zeros((len(frames), util.NSQ_STATE_EXPORT_WORDS), np.int32)
    fr = frames.copy()
    for f in range(len(fr)):
        one = fr[f:f + 1]
        O.orc_silk_nsq(_p(cfg), _p(st), _p(one), _p(x[f]), _p(pulses[f]))
        seeds[f] = one['Seed'][0]
        O.orc_nsq_state_export(_p(st), _p(exp[f]))
    return pulses, seeds, exp


def _setup(L, names):
    for n in names:
        getattr(L, n).restype = None
    return L


@pytest.mark.parametrize('c', CONFIGS)
@pytest.mark.parametrize('kind', ['c', 'simd'])
def test_nsq_oracle_matches_reference(c, kind):
    if not util.ref_available(kind):
        pytest.skip('no ' + kind)
    R = util.load_ref(kind)
    O = util.load_oracle()
    assert O.orc_nsq_frame_size() == util.NSQ_FRAME_DTYPE.itemsize
    _setup(R, ['refh_nsq_state_init', 'refh_silk_nsq', 'refh_nsq_state_export'])
    _setup(O, ['orc_nsq_state_init', 'orc_silk_nsq', 'orc_nsq_state_export'])
    cfg = util.nsq_cfg(*c)
    for seed in range(6):
        x, frames = util.make_nsq_frames(1000 * seed + c[0] + c[5], cfg, 6)
        pr, sr, er = run_ref(R, cfg, x, frames)
        po, so, eo = run_oracle(O, cfg, x, frames)
        assert np.array_equal(pr, po), (c, seed)
        assert np.array_equal(sr, so)
        # rand_seed is a scratch field in the delayed-decision path; everything else must agree
        assert np.array_equal(er, eo), (c, seed, np.nonzero(er != eo))
        assert np.abs(pr).max() > 0
