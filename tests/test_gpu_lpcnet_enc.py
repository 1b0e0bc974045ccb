"""Parity of the batched LPCNet feature extraction (opusb200_lpcnet_encoder_*, opusb200_burg_cepstral_analysis) with the oracle restatement of
dnn/lpcnet_enc.c / freq.c / pitchdnn.c / burg.c, which is pinned bit for bit on the reference's own files (tests/test_lpcnet_enc_oracle_vs_ref.py);
the same device source is bit-identical to the oracle on the host emulator (tests/test_emu_lpcnet_enc.py).  Weights: the seeded random pitch model
inside oracle/_ref/libfargan_ref.so (it travels to the GPU box).
Bar: the arithmetic is the reference's float order; what can differ on the GPU are the double-precision libm calls on the path (log, pow, exp, sqrt
of the device vs glibc: an ulp of a double now and then, which rarely survives the rounding to float) — and the int8 layers of the pitch network turn
such a last-bit difference of an activation into a step of 1/127 now and then.  So: almost every feature vector bit-identical; cepstrum and LPC
within 1e-4 on every frame; pitch within 0.02 (one step of the 1/60-octave grid) and the correlation feature, which is evaluated AT the integer pitch
lag, within 0.05 whenever the two sides picked the same lag."""
import ctypes

import numpy as np
import pytest

import conftest
import fargan_util as F
import util

pytestmark = pytest.mark.gpu
R = F.load_ref()
needs_ref = pytest.mark.skipif(R is None or not hasattr(R, 'refh_lpcnet_enc_state_size'), reason='oracle/_ref/libfargan_ref.so missing')


def _lag(dnn_pitch):
    return np.floor(.5 + 256. / np.power(2., dnn_pitch.astype(np.float64) + 1.5)).astype(np.int64)


@needs_ref
@pytest.mark.parametrize('n,is16', [(7, False), (130, True)])
def test_features_match_oracle(n, is16):
    import opus_b200
    from opus_b200 import nnet
    O = util.load_oracle()
    O.orc_nnet_set_su_bias(0)
    ref = F.RefLpcnetEnc(R)
    lin, _ = ref.linear()
    cv = ref.conv()
    ops = opus_b200.Ops()
    enc = nnet.LpcnetEncoder(ops, lin, cv, n)
    kinds, nfr = min(n, 10), 30
    orcs = [F.OrcLpcnetEnc(O, ref) for _ in range(kinds)]
    sig = np.stack([F.speech_like(300 + s, nfr, [8000.0, 30000.0, 60.0, 1500.0][s % 4]) for s in range(kinds)])
    if is16:
        sig = np.round(sig).astype(np.int16)
    exact = total = 0
    worst = [0.0, 0.0, 0.0]
    for t in range(nfr):
        if t == 21:   # lpcnet_encoder_init on a live object
            enc.reset()
            for o in orcs:
                o.reset()
        pcm = np.ascontiguousarray(sig[:, t * 160:(t + 1) * 160][np.arange(n) % kinds])
        got = enc.compute_single_frame_features(pcm)
        g_if, g_xc, g_lpc, g_misc = enc.get_state()
        for s in range(kinds):
            want = orcs[s].features(sig[s, t * 160:(t + 1) * 160].astype(np.float32))
            w_if, w_xc, w_lpc, w_misc = orcs[s].get()
            for r in range(s, n, kinds):
                assert np.array_equal(got[r], got[s]), (t, r)   # copies of a stream agree with each other exactly
            total += 1
            exact += int(np.array_equal(got[s], want))
            assert np.abs(g_if[s] - w_if).max() < 1e-4 and np.abs(g_xc[s] - w_xc).max() < 1e-5, (t, s)
            d = np.abs(got[s] - want)
            worst[0] = max(worst[0], float(np.delete(d, [18, 19]).max()))
            worst[1] = max(worst[1], float(d[18]))
            assert np.delete(d, [18, 19]).max() < 1e-4, (t, s, d)
            assert d[18] < 0.02, (t, s, got[s][18], want[18])
            if _lag(got[s][18:19])[0] == _lag(want[18:19])[0]:
                worst[2] = max(worst[2], float(d[19]))
                assert d[19] < 0.05, (t, s, got[s][19], want[19])
    print('bit-identical feature vectors: %d of %d; worst cepstrum/LPC %.3g, pitch %.3g, correlation %.3g' % (exact, total, *worst))
    conftest.note_err('lpcnet_enc_cepstrum_lpc', worst[0])
    conftest.note_err('lpcnet_enc_pitch', worst[1])
    conftest.note_err('lpcnet_enc_frame_corr', worst[2])
    assert exact >= 0.95 * total
    enc.close()


@needs_ref
def test_burg_cepstrum_matches_oracle():
    import opus_b200
    from opus_b200 import nnet
    O = util.load_oracle()
    ops = opus_b200.Ops()
    n = 150
    x = np.stack([F.speech_like(500 + s, 1, [8000.0, 30000.0, 3.0][s % 3]) for s in range(n)])
    x[7] = 0
    got = nnet.burg_cepstral_analysis(ops, x)
    exact = 0
    for s in range(n):
        want = np.zeros(36, np.float32)
        O.orc_burg_cepstral_analysis(F.vp(want), F.vp(x[s]))
        d = np.abs(got[s] - want).max()
        conftest.note_err('burg_cepstrum', d)
        assert d < 2e-5, (s, d)
        exact += int(np.array_equal(got[s], want))
    print('bit-identical: %d of %d' % (exact, n))
    assert exact >= 0.9 * n
