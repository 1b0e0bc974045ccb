"""Decoder ctls of the decode path on the device source (host SIMT emulator): OPUS_SET_GAIN, OPUS_SET_PHASE_INVERSION_DISABLED,
OPUS_GET_PITCH, OPUS_GET_LAST_PACKET_DURATION against the reference decoder with the same ctls."""
import numpy as np
import pytest

import emu_util
import util
from test_oracle_decoder_ctl_vs_ref import OPUS_SET_GAIN, OPUS_SET_PHASE_INVERSION_DISABLED, ctl_stream, run_ctl

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built')


@pytest.mark.parametrize('ctls', [[(OPUS_SET_GAIN, 1536)], [(OPUS_SET_PHASE_INVERSION_DISABLED, 1)], [(OPUS_SET_GAIN, -2000), (OPUS_SET_PHASE_INVERSION_DISABLED, 1)]])
@pytest.mark.parametrize('ch,fs,br,harm', [(2, 960, 24000, False), (1, 480, 32000, True)])
def test_emu_decoder_ctls(ctls, ch, fs, br, harm):
    pk, sz = ctl_stream(ch, fs, br, 11 + ch + fs, harmonic=harm)
    ref, rr, pitch, dur = run_ctl('ref', ch, fs, pk, sz, ctls)
    d = emu_util.EmuDecoder(1, ch)
    for req, val in ctls:
        d.L.emu_decoder_ctl(d.h, req, val)
    lost_seen = False
    for t in range(len(sz)):
        pcm, ns, fr = d.decode(pk[t:t + 1], sz[t:t + 1], frame_size=fs)
        assert ns[0] == fs and fr[0] == rr[t]
        assert d.L.emu_decoder_get_ctl(d.h, 0, 4033) == pitch[t]
        assert d.L.emu_decoder_get_ctl(d.h, 0, 4039) == dur[t]
        lost_seen |= sz[t] == 0
        tol = 1e-3 * max(1.0, float(np.abs(ref).max())) if lost_seen else 4e-6 * max(1.0, float(np.abs(ref).max()))
        assert np.abs(pcm[0, :fs * ch].reshape(fs, ch) - ref[t * fs:(t + 1) * fs]).max() < tol
    d.close()
