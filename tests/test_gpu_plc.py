"""Packet-loss concealment of the CUDA decode path through the C ABI (len == 0 = lost packet, as opus_decode(data = NULL)), against the
reference decoder fed the same loss pattern.  Final range bit-exact (0 on a lost frame), sample counts equal.

PCM bound on streams with losses.  The concealment re-estimates a 24th-order LPC from the decoded history and extrapolates with it, which amplifies
last-bit differences of the frames before the loss by orders of magnitude: the reference's own C and SSE/AVX2 builds agree to 1e-7 without loss and
differ by up to 2.9e-2 of full scale (median 1.6e-4) on the concealed streams of this file.  Measured on a B200 (tools/plc_gap_probe.py,
profiles/r02_plc_gap_probe.txt, 214 lossy streams): the GPU is within 3.0e-7 of the reference's C build on every decoded frame, departs only at
lost frames, and on every stream stays CLOSER to the C build than 2.6x the C-vs-SIMD difference of that same stream (max 3.5e-3 where the
reference's two builds are 1.3e-2 apart; median 4.7e-5).  The bar is therefore per stream: max(5e-4, 4 x the reference's own C-vs-SIMD spread on
that stream); without the SIMD build at hand (it travels with oracle/_ref) the old absolute 5e-3 applies.  tests/test_emu_plc.py checks the
concealment arithmetic itself exactly, by giving the emulated device code the oracle's history before each lost frame.  Streams without loss in
the same batch keep the +-1 LSB bar (2e-6 in float)."""
import numpy as np
import pytest

import util
from conftest import note_err

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built')]

PLC_TOL = 5e-3      # absolute fallback (no SIMD build of the reference at hand) and the bound of the int16 / DTX tests
PLC_ABS, PLC_REL = 5e-4, 4.0   # per stream: max(PLC_ABS, PLC_REL x |ref_simd - ref_c|max of that stream)


def _streams(n, ch, fs, nfr, seed):
    sigs = [util.sig_music, util.sig_tonal_noise, util.sig_white, util.sig_harmonic, util.sig_swell]
    rates = [24000, 48000, 64000, 96000, 128000, 32000]
    pk = np.zeros((n, nfr, 1500), np.uint8)
    sz = np.zeros((n, nfr), np.int32)
    for s in range(n):
        x = sigs[s % 5](seed * 1000 + s, nfr * fs, ch).reshape(-1, ch)[:nfr * fs]
        p, z, _ = util.ref_encode(x, ch, fs, rates[s % len(rates)], vbr=(s // 3) % 2, complexity=5 + s % 6)
        pk[s], sz[s] = p, z
    return pk, sz


def _ref_decode(Fs, ch, fsd, pk, sz, kind='c'):
    L = util.load_ref(kind)
    T = len(sz)
    p, z = np.ascontiguousarray(pk), np.ascontiguousarray(sz, np.int32)
    ref, rr = np.zeros((T * fsd, ch), np.float32), np.zeros(T, np.uint32)
    assert L.refh_decode_stream_float(Fs, ch, fsd, T, util._ptr(p), p.shape[1], util._ptr(z), util._ptr(ref), util._ptr(rr)) == 0
    return ref, rr


@pytest.mark.parametrize('ch,fs48,Fs', [(2, 960, 48000), (1, 960, 48000), (2, 480, 48000), (2, 240, 48000), (1, 120, 48000), (2, 960, 16000)])
def test_random_loss_patterns_vs_reference(ch, fs48, Fs):
    import opus_b200
    n, nfr = 40, 30
    rng = np.random.default_rng(ch * 100 + fs48 + Fs // 1000)
    pk, sz = _streams(n, ch, fs48, nfr, seed=fs48 + ch)
    lost = rng.random((n, nfr)) < 0.18
    lost[:4] = False                    # clean streams in the same batch
    lost[4, 6:16] = True                # a long burst: pitch PLC fades into noise PLC
    lost[5, 0:2] = True                 # nothing decoded yet: zeros
    lost[6, 3::2] = True                # every other packet: skip_plc
    sz = np.where(lost, 0, sz).astype(np.int32)
    fsd = fs48 * Fs // 48000
    dec = opus_b200.DecoderBatch(n, channels=ch, Fs=Fs)
    outs = np.zeros((n, nfr * fsd, ch), np.float32)
    rngs = np.zeros((n, nfr), np.uint32)
    for t in range(nfr):
        got, ns = dec.decode(np.ascontiguousarray(pk[:, t]), np.ascontiguousarray(sz[:, t]), frame_size=fsd, dtype=np.float32)
        assert (ns == fsd).all(), (t, ns)
        outs[:, t * fsd:(t + 1) * fsd] = got.reshape(n, fsd, ch)
        rngs[:, t] = dec.final_range()
    dec.close()
    worst = 0.0
    for s in range(n):
        ref, rr = _ref_decode(Fs, ch, fsd, pk[s], sz[s])
        assert (rngs[s] == rr).all(), s
        err = float(np.abs(outs[s] - ref).max())
        if not lost[s].any():
            note_err('plc_clean_stream', err)
            assert err < 2e-6, (s, err)
            assert np.abs(util.to_int16(outs[s]).astype(np.int32) - util.to_int16(ref).astype(np.int32)).max() <= 1
        else:
            if util.ref_available('simd'):
                spread = float(np.abs(_ref_decode(Fs, ch, fsd, pk[s], sz[s], 'simd')[0] - ref).max())
                note_err('plc_concealed_over_reference_spread', err / max(spread, PLC_ABS / PLC_REL))
                assert err < max(PLC_ABS, PLC_REL * spread), (s, err, spread)
            else:
                assert err < PLC_TOL, (s, err)
            note_err('plc_concealed_stream', err)
            worst = max(worst, err)
            # frames before the first loss are ordinary decoded frames
            t0 = int(np.argmax(lost[s]))
            if t0:
                note_err('plc_before_first_loss', np.abs(outs[s][:t0 * fsd] - ref[:t0 * fsd]).max())
                assert np.abs(outs[s][:t0 * fsd] - ref[:t0 * fsd]).max() < 2e-6
    print('worst concealed-stream error', worst)


def test_int16_and_long_frame_sizes():
    """int16 output of concealed frames; opus_decode(NULL, 60 ms) = three 20 ms frames written one after the other"""
    import opus_b200
    ch, fs48, n, nfr, fs_arg = 2, 960, 8, 12, 2880
    pk, sz = _streams(n, ch, fs48, nfr, seed=77)
    lost = np.zeros((n, nfr), bool)
    lost[:, 5] = True
    lost[1::2, 6] = True
    lost[2, 0] = True
    sz = np.where(lost, 0, sz).astype(np.int32)
    dec = opus_b200.DecoderBatch(n, channels=ch)
    L = util.load_ref('c')
    refs = []
    for s in range(n):
        p, z = np.ascontiguousarray(pk[s]), np.ascontiguousarray(sz[s])
        ref, rr = np.zeros((nfr * fs_arg, ch), np.float32), np.zeros(nfr, np.uint32)
        assert L.refh_decode_stream_float(48000, ch, fs_arg, nfr, util._ptr(p), p.shape[1], util._ptr(z), util._ptr(ref), util._ptr(rr)) == 0
        refs.append((ref, rr))
    for t in range(nfr):
        got, ns = dec.decode(np.ascontiguousarray(pk[:, t]), np.ascontiguousarray(sz[:, t]), frame_size=fs_arg, dtype=np.int16)
        fr = dec.final_range()
        for s in range(n):
            want = fs_arg if lost[s, t] else fs48
            assert ns[s] == want, (s, t, ns[s])
            assert fr[s] == refs[s][1][t]
            r16 = util.to_int16(refs[s][0][t * fs_arg:t * fs_arg + want])
            assert np.abs(got[s, :want * ch].reshape(want, ch).astype(np.int32) - r16.astype(np.int32)).max() <= 164   # 5e-3 of full scale
    dec.close()


def test_negative_length_skips_the_stream():
    """this library's extension: len < 0 leaves the stream untouched in this call (0 samples), neither decoded nor concealed"""
    import opus_b200
    ch, fs, n, nfr = 2, 960, 4, 6
    pk, sz = _streams(n, ch, fs, nfr, seed=5)
    a, b = opus_b200.DecoderBatch(n, channels=ch), opus_b200.DecoderBatch(n, channels=ch)
    for t in range(nfr):
        ga, _ = a.decode(np.ascontiguousarray(pk[:, t]), np.ascontiguousarray(sz[:, t]), frame_size=fs, dtype=np.float32)
        if t == 2:   # stream 1 sits this call out in decoder b, then catches up with the same packet
            z = sz[:, t].copy(); z[1] = -1
            gb, ns = b.decode(np.ascontiguousarray(pk[:, t]), z, frame_size=fs, dtype=np.float32)
            assert ns[1] == 0 and (ns[[0, 2, 3]] == fs).all()
            z2 = np.full(n, -1, np.int32); z2[1] = sz[1, t]
            gb2, ns2 = b.decode(np.ascontiguousarray(pk[:, t]), z2, frame_size=fs, dtype=np.float32)
            assert ns2[1] == fs and (ns2[[0, 2, 3]] == 0).all()
            gb[1] = gb2[1]
        else:
            gb, _ = b.decode(np.ascontiguousarray(pk[:, t]), np.ascontiguousarray(sz[:, t]), frame_size=fs, dtype=np.float32)
        assert np.array_equal(ga, gb)
    a.close(); b.close()


@pytest.mark.parametrize('ch,fs', [(2, 960), (1, 480), (2, 120)])
def test_int16_soft_clipping_vs_reference(ch, fs):
    """opus_decode()'s int16 output on streams that overshoot full scale (opus_pcm_soft_clip, src/opus.c:39-166) against the reference's
    int16 API.  Bound: 3 LSB (the float signal ahead
    of the non-linearity is within 2e-5 of the reference; the emulated device code, 10x closer, stays within 1 LSB: test_emu_softclip.py)."""
    import opus_b200
    from test_oracle_softclip_vs_ref import loud_signal
    L = util.load_ref('c')
    n, nfr = 6, 10
    pk = np.zeros((n, nfr, 1500), np.uint8)
    sz = np.zeros((n, nfr), np.int32)
    for s in range(n):
        x = loud_signal(s * 7 + ch, nfr * fs, ch, s % 3)
        p, z, _ = util.ref_encode(x, ch, fs, [64000, 24000, 96000][s % 3])
        pk[s], sz[s] = p, z
    dec = opus_b200.DecoderBatch(n, channels=ch)
    out = np.zeros((n, nfr * fs, ch), np.int16)
    rngs = np.zeros((n, nfr), np.uint32)
    for t in range(nfr):
        got, ns = dec.decode(np.ascontiguousarray(pk[:, t]), np.ascontiguousarray(sz[:, t]), frame_size=fs, dtype=np.int16)
        assert (ns == fs).all()
        out[:, t * fs:(t + 1) * fs] = got.reshape(n, fs, ch)
        rngs[:, t] = dec.final_range()
    dec.close()
    clipped = 0
    for s in range(n):
        p, z = np.ascontiguousarray(pk[s]), np.ascontiguousarray(sz[s])
        ref, rr = np.zeros((nfr * fs, ch), np.int16), np.zeros(nfr, np.uint32)
        assert L.refh_decode_stream(48000, ch, fs, nfr, util._ptr(p), p.shape[1], util._ptr(z), util._ptr(ref), util._ptr(rr)) == 0
        assert (rngs[s] == rr).all()
        d = np.abs(out[s].astype(np.int32) - ref.astype(np.int32))
        assert d.max() <= 3, (s, int(d.max()))
        assert (d <= 1).mean() > 0.999
        flt, fr = np.zeros((nfr * fs, ch), np.float32), np.zeros(nfr, np.uint32)
        assert L.refh_decode_stream_float(48000, ch, fs, nfr, util._ptr(p), p.shape[1], util._ptr(z), util._ptr(flt), util._ptr(fr)) == 0
        clipped += int((np.abs(util.to_int16(flt).astype(np.int32) - ref.astype(np.int32)) > 1).any())
    assert clipped >= 2   # the test does exercise the non-linearity


@pytest.mark.parametrize('ch,fs48', [(2, 480), (1, 240)])
def test_dtx_frames_inside_a_packet(ch, fs48):
    """code-2 packets whose second frame is 0 or 1 byte long: a decoded frame and a concealed frame in one packet"""
    import opus_b200
    from test_oracle_plc_vs_ref import _dtx_packets
    n, nfr, fs = 5, 10, 2 * fs48
    pks, szs = zip(*[_dtx_packets(ch, fs48, nfr, 3 + ch + s) for s in range(n)])
    dec = opus_b200.DecoderBatch(n, channels=ch)
    outs = np.zeros((n, nfr * fs, ch), np.float32)
    rngs = np.zeros((n, nfr), np.uint32)
    for t in range(nfr):
        got, ns = dec.decode(np.ascontiguousarray(np.stack([p[t] for p in pks])), np.array([z[t] for z in szs], np.int32), frame_size=fs,
                             dtype=np.float32)
        assert (ns == fs).all(), ns
        outs[:, t * fs:(t + 1) * fs] = got.reshape(n, fs, ch)
        rngs[:, t] = dec.final_range()
    dec.close()
    for s in range(n):
        ref, rr = _ref_decode(48000, ch, fs, pks[s], szs[s])
        assert (rngs[s] == rr).all()
        note_err('plc_other', np.abs(outs[s] - ref).max())
        assert np.abs(outs[s] - ref).max() < PLC_TOL
