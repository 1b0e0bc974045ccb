"""Pin the CELT encoder restatement (oracle/celt_enc.inc) to the reference's celt_encode_with_ec driven directly
(oracle/_ref): packet bytes and final range identical, frame after frame (encoder state carried)."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built (needs /root/reference)')


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def encode_both(sig, ch, fs, nbytes, complexity, bitrate=-1000, vbr=0, lsb=24, kind='c', nframes=12):
    R, O = util.load_ref(kind), util.load_oracle()
    x = (sig.astype(np.float32) / 32768.0).reshape(-1, fs * ch)[:nframes]
    rs = np.zeros(R.refh_celt_encoder_size(ch) + 64, np.uint8)
    assert R.refh_celt_encoder_init(_p(rs), ch, complexity, bitrate, vbr, lsb) == 0
    os_ = np.zeros(O.orc_encoder_size(), np.uint8)
    O.orc_encoder_init.restype = None
    O.orc_encoder_ctl.restype = None
    O.orc_encoder_final_range.restype = ctypes.c_uint32
    O.orc_encoder_init(_p(os_), ch)
    O.orc_encoder_ctl(_p(os_), complexity, -1 if bitrate == -1000 else bitrate, vbr, lsb)
    out = []
    for f in range(len(x)):
        a, b = np.zeros(1275, np.uint8), np.zeros(1275, np.uint8)
        rng = ctypes.c_uint32(0)
        na = R.refh_celt_encode(_p(rs), _p(x[f]), fs, _p(a), nbytes, ctypes.byref(rng))
        nb = O.orc_celt_encode_frame(_p(os_), _p(x[f]), fs, _p(b), nbytes)
        out.append((f, na, nb, a[:max(na, 0)].copy(), b[:max(nb, 0)].copy(), rng.value, O.orc_encoder_final_range(_p(os_))))
    return out


CASES = [(ch, fs, nbytes, s) for ch in (1, 2) for fs in (960, 480, 240, 120) for nbytes in (40, 80, 160, 320)
         for s in ('music', 'noise')]


@pytest.mark.parametrize('ch,fs,nbytes,sname', CASES)
def test_complexity0_cbr(ch, fs, nbytes, sname):
    nb = max(8, nbytes * fs // 960)
    sig = (util.sig_music if sname == 'music' else util.sig_white)(ch * 100 + fs + nbytes, fs * 12, ch)
    for (f, na, nbo, a, b, ra, rb) in encode_both(sig, ch, fs, nb, 0):
        assert nbo != -5, 'frame needed a decision that is not restated'
        assert na == nbo, (f, na, nbo)
        assert np.array_equal(a, b), (f, int(np.argmax(a != b)))
        assert ra == rb
