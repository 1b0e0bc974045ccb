"""int16 output of the device decoder source (host SIMT emulator) on streams that overshoot full scale: opus_decode()'s soft clipping
(src/opus.c:39-166, src/opus_decoder.c:875-878) against the reference's int16 API.  +-1 LSB like every decode comparison (the float
signal ahead of the non-linearity carries re-association differences of ~1e-6); one-frame packets, where per-frame and per-packet
soft clipping are the same thing."""
import numpy as np
import pytest

import emu_util
import util
from test_oracle_softclip_vs_ref import loud_signal

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built')


@pytest.mark.parametrize('kind', [0, 1, 2])
@pytest.mark.parametrize('ch,fs,br', [(2, 960, 64000), (1, 480, 24000), (2, 120, 96000)])
def test_emu_int16_soft_clip(kind, ch, fs, br):
    L = util.load_ref('c')
    nfr = 10
    x = loud_signal(kind * 10 + ch, nfr * fs, ch, kind)
    pk, sz, _ = util.ref_encode(x, ch, fs, br)
    sz = np.ascontiguousarray(sz, np.int32)
    if kind == 1:
        sz[5] = 0    # a concealed frame goes through the same soft clipping
    T = len(sz)
    ref, rr = np.zeros((T * fs, ch), np.int16), np.zeros(T, np.uint32)
    assert L.refh_decode_stream(48000, ch, fs, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(ref), util._ptr(rr)) == 0
    d = emu_util.EmuDecoder(1, ch)
    got = np.zeros((T * fs, ch), np.int16)
    for t in range(T):
        pcm, ns, fr = d.decode(pk[t:t + 1], sz[t:t + 1], int16=True, frame_size=fs)
        assert ns[0] == fs and fr[0] == rr[t]
        got[t * fs:(t + 1) * fs] = pcm[0, :fs * ch].reshape(fs, ch)
    d.close()
    lossy = kind == 1
    diff = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= (40 if lossy else 1), int(diff.max())   # concealment: see test_emu_plc.py for the bound
    # and the plain conversion would have been wrong here
    flt, fr2 = np.zeros((T * fs, ch), np.float32), np.zeros(T, np.uint32)
    assert L.refh_decode_stream_float(48000, ch, fs, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(flt), util._ptr(fr2)) == 0
    if kind != 2:
        assert (np.abs(util.to_int16(flt).astype(np.int32) - ref.astype(np.int32)) > 1).any()


@pytest.mark.parametrize('kind', [0, 2])
@pytest.mark.parametrize('ch,fs,br', [(2, 1920, 64000), (1, 2880, 48000), (2, 960, 96000)])
def test_emu_int16_soft_clip_multiframe_packets(kind, ch, fs, br):
    """40 / 60 ms packets (two / three 20 ms frames, codes 1-3) and 20 ms packets of two 10 ms frames that overshoot full scale:
    opus_decode() soft-clips the whole PACKET (src/opus_decoder.c:875-878), so a clipped region may run across the frames of a packet"""
    L = util.load_ref('c')
    nfr = 6
    x = loud_signal(kind * 10 + ch + fs, nfr * fs, ch, kind)
    if fs == 960:   # two 10 ms frames per packet: encode at 10 ms and merge pairs into code-1 / code-2 packets by hand
        p10, s10, _ = util.ref_encode(x, ch, 480, br)
        pk = np.zeros((nfr, 1500), np.uint8)
        sz = np.zeros(nfr, np.int32)
        for t in range(nfr):
            a, b = bytes(p10[2 * t, 1:s10[2 * t]]), bytes(p10[2 * t + 1, 1:s10[2 * t + 1]])
            toc = int(p10[2 * t, 0]) & 0xFC
            if len(a) == len(b):
                pkt = bytes([toc | 1]) + a + b
            else:
                ln = bytes([len(a)]) if len(a) < 252 else bytes([252 + (len(a) & 3), (len(a) - (252 + (len(a) & 3))) >> 2])
                pkt = bytes([toc | 2]) + ln + a + b
            pk[t, :len(pkt)] = np.frombuffer(pkt, np.uint8)
            sz[t] = len(pkt)
    else:
        pk, sz, _ = util.ref_encode(x, ch, fs, br)
        sz = np.ascontiguousarray(sz, np.int32)
    T = len(sz)
    ref, rr = np.zeros((T * fs, ch), np.int16), np.zeros(T, np.uint32)
    assert L.refh_decode_stream(48000, ch, fs, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(ref), util._ptr(rr)) == 0
    d = emu_util.EmuDecoder(1, ch)
    got = np.zeros((T * fs, ch), np.int16)
    for t in range(T):
        pcm, ns, fr = d.decode(pk[t:t + 1], sz[t:t + 1], int16=True, frame_size=fs)
        assert ns[0] == fs and fr[0] == rr[t]
        got[t * fs:(t + 1) * fs] = pcm[0, :fs * ch].reshape(fs, ch)
    d.close()
    diff = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= 1, int(diff.max())
    flt, fr2 = np.zeros((T * fs, ch), np.float32), np.zeros(T, np.uint32)
    assert L.refh_decode_stream_float(48000, ch, fs, T, util._ptr(pk), pk.shape[1], util._ptr(sz), util._ptr(flt), util._ptr(fr2)) == 0
    assert (np.abs(util.to_int16(flt).astype(np.int32) - ref.astype(np.int32)) > 1).any()   # the clipping mattered
