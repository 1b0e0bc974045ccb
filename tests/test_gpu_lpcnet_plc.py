"""Parity of the batched neural packet-loss concealment (opusb200_lpcnet_plc_*) with the oracle restatement of dnn/lpcnet_plc.c, which is pinned
bit-exact against the reference's own lpcnet_plc.c (tests/test_plc_oracle_vs_ref.py).  Weights: the seeded random predictor and vocoder inside
oracle/_ref/libfargan_ref.so (it travels to the GPU box).  Every stream of the batch follows its own FEC / gap schedule; the calls are in lockstep.
Bar: int16 PCM within 1 LSB of the oracle on every frame (the vocoder's two libm calls may differ in the last ulp, tests/test_gpu_fargan.py) and
bit-identical on almost all; the FEC counters and the loss count exact."""
import numpy as np
import pytest

import fargan_util as F
import util

pytestmark = pytest.mark.gpu
R = F.load_ref()


@pytest.mark.skipif(R is None or not hasattr(R, 'refh_plc_state_size'), reason='oracle/_ref/libfargan_ref.so missing')
@pytest.mark.parametrize('n', [5, 70])
def test_conceal_matches_oracle(n):
    import opus_b200
    from opus_b200 import nnet
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    ref = F.RefPlc(R)
    pl, shapes = ref.layers()
    fl, _ = ref.fargan_layers()
    ops = opus_b200.Ops()
    plc = nnet.LpcnetPlc(ops, pl, fl, n)
    nfr, kinds = 30, min(n, 6)
    sched = [F.plc_schedule(200 + s, nfr) for s in range(kinds)]
    sts = []
    for s in range(kinds):
        st = np.zeros(O.orc_plc_size() + 64, np.uint8)
        O.orc_plc_init(F.vp(st), pl, fl)
        sts.append(st)
    exact = 0
    for t in range(nfr):
        feats = np.zeros((n, 20), np.float32)
        have = np.full(n, -1, np.int32)
        for r in range(n):
            act = sched[r % kinds][t]
            if act[0] == 'feat':
                feats[r], have[r] = act[1], 1
            elif act[0] == 'skip':
                have[r] = 0
        plc.fec_add(feats, have)
        got = plc.conceal()
        for s in range(kinds):
            act = sched[s][t]
            if act[0] == 'feat':
                O.orc_plc_fec_add(F.vp(sts[s]), F.vp(act[1]))
            elif act[0] == 'skip':
                O.orc_plc_fec_add(F.vp(sts[s]), None)
            want = np.zeros(160, np.int16)
            assert O.orc_plc_conceal(F.vp(sts[s]), F.vp(want)) == 0
            for r in range(s, n, kinds):
                d = np.abs(got[r].astype(np.int32) - want.astype(np.int32)).max()
                assert d <= 1, (t, r, int(d))
            exact += int(np.array_equal(got[s], want))
        if t == 17:   # a new recovery for everybody
            plc.fec_clear(); plc.reset()
            for st in sts:
                O.orc_plc_fec_clear(F.vp(st)); O.orc_plc_reset(F.vp(st))
    print('bit-identical frames: %d of %d' % (exact, nfr * kinds))
    assert exact >= nfr * kinds - 3
    feat, cont, g1, g2, cnt = plc.state()
    for s in range(kinds):
        bufs = [np.zeros(k, np.float32) for k in (20, 100, 128, 128, 2400)] + [np.zeros(8, np.int32)]
        O.orc_plc_get_state(F.vp(sts[s]), *[F.vp(b) for b in bufs])
        assert (cnt[s] == [bufs[5][2], bufs[5][3], bufs[5][4], bufs[5][1]]).all()   # fec_read_pos, fec_fill_pos, fec_skip, loss_count
        assert np.abs(g1[s] - bufs[2]).max() < 1e-5 and np.abs(g2[s] - bufs[3]).max() < 1e-5
    plc.close()


@pytest.mark.skipif(R is None or not hasattr(R, 'refh_plc_state_size'), reason='oracle/_ref/libfargan_ref.so missing')
def test_history_analysis_needs_the_pitch_model():
    import opus_b200
    from opus_b200 import nnet
    ref = F.RefPlc(R)
    pl, _ = ref.layers()
    fl, _ = ref.fargan_layers()
    ops = opus_b200.Ops()
    plc = nnet.LpcnetPlc(ops, pl, fl, 4)
    plc.update(np.zeros((4, 160), np.int16))
    with pytest.raises(opus_b200.OpusB200Error) as e:
        plc.conceal()
    assert e.value.code == -6   # OPUS_INVALID_STATE: the pitch model of the history analysis was not loaded
    plc.close()


@pytest.mark.skipif(R is None or not hasattr(R, 'refh_lpcnet_enc_state_size'), reason='oracle/_ref/libfargan_ref.so missing')
@pytest.mark.parametrize('n', [6, 40])
def test_update_conceal_sequences_match_oracle(n):
    """a decoder's life in lockstep: good frames (lpcnet_plc_update), bursts of losses whose first frame analyses the decoded history
    (burg_cepstral_analysis + the feature extractor + the predictor); every stream has its own audio and its own DRED vectors.
    Bar: the libm note of tests/test_gpu_lpcnet_enc.py applies to the analysed features, and the vocoder turns a last-bit difference of a feature
    into a few LSB of PCM at most: within 1 LSB on almost every frame, never more than 8, most frames bit-identical."""
    import opus_b200
    from opus_b200 import nnet
    from test_plc_oracle_vs_ref import loss_pattern
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    ref = F.RefPlc(R)
    pl, _ = ref.layers()
    fl, _ = ref.fargan_layers()
    renc = F.RefLpcnetEnc(R)
    lin, _ = renc.linear()
    cv = renc.conv()
    ops = opus_b200.Ops()
    plc = nnet.LpcnetPlc(ops, pl, fl, n)
    plc.load_pitch_model(lin, cv)
    kinds, nfr = min(n, 6), 60
    sts, encs = [], []
    for s in range(kinds):
        st = np.zeros(O.orc_plc_size() + 64, np.uint8)
        O.orc_plc_init(F.vp(st), pl, fl)
        enc = F.OrcLpcnetEnc(O, renc)
        O.orc_plc_set_encoder(F.vp(st), F.vp(enc.st))
        O.orc_plc_reset(F.vp(st))
        sts.append(st); encs.append(enc)
    audio = np.stack([np.round(F.speech_like(700 + s, nfr, 9000.0)).astype(np.int16) for s in range(kinds)])
    lost = loss_pattern(5, nfr)
    rng = np.random.default_rng(9)
    exact = total = 0
    worst = 0
    for t in range(nfr):
        if not lost[t]:
            plc.update(np.ascontiguousarray(audio[np.arange(n) % kinds, t * 160:(t + 1) * 160]))
            for s in range(kinds):
                O.orc_plc_update(F.vp(sts[s]), F.vp(np.ascontiguousarray(audio[s, t * 160:(t + 1) * 160])))
            continue
        feats = np.zeros((n, 20), np.float32)
        have = np.full(n, -1, np.int32)
        for s in range(kinds):
            if rng.random() < 0.35:
                f = (0.5 * rng.standard_normal(20)).astype(np.float32)
                O.orc_plc_fec_add(F.vp(sts[s]), F.vp(f))
                feats[s::kinds], have[s::kinds] = f, 1
        plc.fec_add(feats, have)
        got = plc.conceal()
        for s in range(kinds):
            want = np.zeros(160, np.int16)
            assert O.orc_plc_conceal(F.vp(sts[s]), F.vp(want)) == 0
            for r in range(s, n, kinds):
                assert np.array_equal(got[r], got[s]), (t, r)
            d = int(np.abs(got[s].astype(np.int32) - want.astype(np.int32)).max())
            worst = max(worst, d)
            assert d <= 8, (t, s, d)
            total += 1
            exact += int(d == 0)
        if not lost[min(t + 1, nfr - 1)]:
            plc.fec_clear()
            for st in sts:
                O.orc_plc_fec_clear(F.vp(st))
    print('bit-identical frames: %d of %d, worst %d LSB' % (exact, total, worst))
    import conftest
    conftest.note_err('lpcnet_plc_history_pcm_lsb', worst)
    assert exact >= 0.8 * total
    plc.close()
