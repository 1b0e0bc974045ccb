"""Pins the neural packet-loss concealment restatement (oracle/fargan_oracle.c: orc_plc_*, after dnn/lpcnet_plc.c) against the reference's own
lpcnet_plc.c compiled with seeded random weights (oracle/gen_plc_weights.py + gen_fargan_weights.py, generic C branch of dnn/vec.h like the FARGAN
pin).  Covered: compute_plc_pred, get_fec_or_pred (FEC / DRED features, gaps, pure prediction), the attenuation table, queue_features, the PCM
history, fargan_cont + fargan_synthesize_int, lpcnet_plc_reset, fec_add / fec_clear, lpcnet_plc_update and the history analysis of a first concealed
frame that follows good audio (burg_cepstral_analysis + lpcnet_compute_single_frame_features_float + the predictor over the frames not analysed
yet, with lpcnet_enc_oracle.c as the feature extractor).  Bar: bit-identical int16 PCM and state."""
import ctypes

import numpy as np
import pytest

import fargan_util as F
import util

R = F.load_ref()
pytestmark = pytest.mark.skipif(R is None or not hasattr(R, 'refh_plc_state_size'), reason='oracle/_ref/libfargan_ref.so not built (needs /root/reference)')


def _oracle(ref):
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    pl, shapes = ref.layers()
    fl, _ = ref.fargan_layers()
    assert shapes == [(57, 128), (128, 20), (128, 384), (128, 384), (128, 384), (128, 384)]
    st = np.zeros(O.orc_plc_size() + 64, np.uint8)
    O.orc_plc_init(F.vp(st), pl, fl)
    return O, st, (pl, fl)


def _state(O, st):
    bufs = [np.zeros(k, np.float32) for k in (20, 100, 128, 128, 2400)] + [np.zeros(8, np.int32)]
    O.orc_plc_get_state(F.vp(st), *[F.vp(b) for b in bufs])
    return bufs


def test_predictor_alone():
    ref = F.RefPlc(R)
    O, st, keep = _oracle(ref)
    rng = np.random.default_rng(3)
    for k in range(12):
        x = (0.7 * rng.standard_normal(57)).astype(np.float32)
        a, b = np.zeros(20, np.float32), np.zeros(20, np.float32)
        R.refh_plc_pred(F.vp(ref.st), F.vp(a), F.vp(x))
        O.orc_plc_pred(F.vp(st), F.vp(b), F.vp(x))
        assert np.array_equal(a, b), k
    assert np.abs(a).max() > 1e-3


@pytest.mark.parametrize('seed', [1, 2, 3, 4, 5])
def test_conceal_sequences(seed):
    ref = F.RefPlc(R)
    O, st, keep = _oracle(ref)
    energy = 0
    for t, act in enumerate(F.plc_schedule(seed, 40)):
        if act[0] == 'feat':
            ref.fec_add(act[1])
            O.orc_plc_fec_add(F.vp(st), F.vp(act[1]))
        elif act[0] == 'skip':
            ref.fec_add(None)
            O.orc_plc_fec_add(F.vp(st), None)
        want = ref.conceal()
        got = np.zeros(160, np.int16)
        assert O.orc_plc_conceal(F.vp(st), F.vp(got)) == 0
        assert np.array_equal(got, want), (t, int(np.abs(got.astype(int) - want).max()))
        energy += int(np.abs(want.astype(np.int64)).sum())
        if t == 22 and seed % 2:   # a new burst: fec_clear + reset (what the decoder does between recoveries)
            ref.fec_clear(); ref.reset()
            O.orc_plc_fec_clear(F.vp(st)); O.orc_plc_reset(F.vp(st))
    assert energy > 1000
    for a, b in zip(_state(O, st), ref.state()):
        assert np.array_equal(a, b)


def test_history_analysis_needs_the_feature_extractor():
    """after lpcnet_plc_update (good audio) the first conceal runs the feature analysis: without an encoder object the oracle says so"""
    ref = F.RefPlc(R)
    O, st, keep = _oracle(ref)
    pcm = np.zeros(160, np.int16)
    O.orc_plc_update(F.vp(st), F.vp(pcm))
    assert O.orc_plc_conceal(F.vp(st), F.vp(pcm)) == -1


def loss_pattern(seed, nframes):
    """good / lost per frame: bursts of 1..6 losses between runs of good audio, the first frames good"""
    rng = np.random.default_rng(seed)
    lost = np.zeros(nframes, bool)
    t = int(rng.integers(3, 9))
    while t < nframes:
        k = int(rng.integers(1, 7))
        lost[t:t + k] = True
        t += k + int(rng.integers(1, 12))
    return lost


@pytest.mark.parametrize('seed', [1, 2, 3, 4])
def test_update_conceal_sequences(seed):
    """a decoder's life: good frames (lpcnet_plc_update), then a burst of losses, the first of which analyses the history"""
    ref = F.RefPlc(R)
    O, st, keep = _oracle(ref)
    enc = F.OrcLpcnetEnc(O, F.RefLpcnetEnc(R))
    O.orc_plc_set_encoder(F.vp(st), F.vp(enc.st))
    O.orc_plc_reset(F.vp(st))
    nfr = 70
    audio = np.round(F.speech_like(900 + seed, nfr, 9000.0)).astype(np.int16)
    lost = loss_pattern(seed, nfr)
    rng = np.random.default_rng(seed)
    analysed = 0
    for t in range(nfr):
        if not lost[t]:
            frame = np.ascontiguousarray(audio[t * 160:(t + 1) * 160])
            R.refh_plc_update_pcm(F.vp(ref.st), F.vp(frame))
            O.orc_plc_update(F.vp(st), F.vp(frame))
            continue
        if seed % 2 and rng.random() < 0.4:   # a DRED vector for this frame
            f = (0.5 * rng.standard_normal(20)).astype(np.float32)
            ref.fec_add(f)
            O.orc_plc_fec_add(F.vp(st), F.vp(f))
        before = _state(O, st)[5][5]
        want = ref.conceal()
        got = np.zeros(160, np.int16)
        assert O.orc_plc_conceal(F.vp(st), F.vp(got)) == 0
        analysed += int(before + 160 <= 2400 and not lost[t - 1])
        assert np.array_equal(got, want), (t, int(np.abs(got.astype(int) - want).max()))
        if not lost[min(t + 1, nfr - 1)] and seed % 2:
            ref.fec_clear()
            O.orc_plc_fec_clear(F.vp(st))
    assert analysed >= 3
    for a, b in zip(_state(O, st), ref.state()):
        assert np.array_equal(a, b)
