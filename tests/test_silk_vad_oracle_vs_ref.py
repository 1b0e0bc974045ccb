"""Pin the voice activity detector of a SILK frame (oracle/silk_oracle.c: orc_silk_vad_init, orc_silk_encode_do_vad, orc_silk_hp_variable_cutoff) to the
reference's silk_VAD_Init / silk_encode_do_VAD_FLP (silk_VAD_GetSA_Q8 inside) / silk_HP_variable_cutoff of oracle/_ref (silk/VAD.c,
silk/float/encode_frame_FLP.c:45-82, silk/HP_variable_cutoff.c): the whole silk_VAD_state, speech activity, tilt, the four band qualities, the VAD /
DTX bookkeeping and the cutoff smoother, bit for bit, frame after frame with the state carried."""
import ctypes

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available('c'), reason='oracle/_ref not built (needs /root/reference)')

I32 = ctypes.c_int32


class Vad(ctypes.Structure):   # OrcSilkVad == RefhSilkVad == OpusB200SilkVad
    _fields_ = [('AnaState', I32 * 2), ('AnaState1', I32 * 2), ('AnaState2', I32 * 2), ('XnrgSubfr', I32 * 4), ('NrgRatioSmth_Q8', I32 * 4), ('HPstate', I32),
                ('NL', I32 * 4), ('inv_NL', I32 * 4), ('NoiseLevelBias', I32 * 4), ('counter', I32), ('speech_activity_Q8', I32), ('input_tilt_Q15', I32),
                ('input_quality_bands_Q15', I32 * 4), ('signalType', I32), ('noSpeechCounter', I32), ('inDTX', I32), ('VAD_flag', I32),
                ('prevSignalType', I32), ('prevLag', I32), ('variable_HP_smth1_Q15', I32)]


def vad_signal(rng, fs_kHz, frame_length, nframes):
    """Speech-like test signal: silence, noise floors of several levels, voiced bursts and full-scale passages, frame by frame."""
    out = np.zeros((nframes, frame_length), np.int16)
    t = np.arange(frame_length)
    level = 10.0 ** rng.uniform(0, 2.5)
    for f in range(nframes):
        kind = rng.integers(0, 6)
        x = rng.standard_normal(frame_length) * level
        if kind == 0:
            x[:] = 0
        elif kind in (2, 3):
            f0 = rng.uniform(80, 350) / (fs_kHz * 1000.0)
            amp = 10.0 ** rng.uniform(2, 4.3)
            for h in range(1, 8):
                x += amp / h * np.sin(2 * np.pi * f0 * h * (t + f * frame_length))
        elif kind == 4:
            x = rng.standard_normal(frame_length) * 10.0 ** rng.uniform(2, 4.8)
        elif kind == 5 and rng.random() < 0.3:
            level = 10.0 ** rng.uniform(0, 3.0)
        out[f] = np.clip(np.rint(x), -32768, 32767).astype(np.int16)
    return out


@pytest.mark.parametrize('fs_kHz,frame_ms', [(16, 20), (8, 20), (12, 20), (16, 10), (12, 10), (8, 10)])
def test_encode_do_vad(fs_kHz, frame_ms):
    R, O = util.load_ref('c'), util.load_oracle()
    assert O.orc_silk_vad_size() == ctypes.sizeof(Vad)
    rng = np.random.default_rng(fs_kHz * 100 + frame_ms)
    frame_length = fs_kHz * frame_ms
    active = 0
    for stream in range(12):
        a, b = Vad(), Vad()
        R.refh_silk_vad_init(ctypes.byref(a))
        O.orc_silk_vad_init(ctypes.byref(b))
        assert bytes(a) == bytes(b)
        pcm = vad_signal(rng, fs_kHz, frame_length, 90)
        for f in range(pcm.shape[0]):
            activity = int(rng.integers(0, 2)) if stream % 3 == 0 else 1
            p = pcm[f].ctypes.data_as(ctypes.POINTER(ctypes.c_int16))
            R.refh_silk_encode_do_vad(ctypes.byref(a), p, frame_length, fs_kHz, activity)
            O.orc_silk_encode_do_vad(ctypes.byref(b), p, frame_length, fs_kHz, activity)
            assert bytes(a) == bytes(b), (stream, f, a.speech_activity_Q8, b.speech_activity_Q8)
            active += a.VAD_flag
            # the cutoff smoother runs on the previous frame's decision (silk_Encode calls it before the frame is coded)
            for v in (a, b):
                v.prevSignalType = int(2 if (v.VAD_flag and (f % 3)) else v.VAD_flag)
                v.prevLag = int(32 + (f * 37 + stream * 11) % (18 * fs_kHz - 32))
            R.refh_silk_hp_variable_cutoff(ctypes.byref(a), fs_kHz)
            O.orc_silk_hp_variable_cutoff(ctypes.byref(b), fs_kHz)
            assert bytes(a) == bytes(b), (stream, f, a.variable_HP_smth1_Q15, b.variable_HP_smth1_Q15)
    assert 100 < active < 12 * 90 - 100, active   # both decisions are exercised


def run_vad_streams(n, fs_kHz, frame_ms, nframes, seed):
    O = util.load_oracle()
    rng = np.random.default_rng(seed)
    frame_length = fs_kHz * frame_ms
    recs, want = (Vad * n)(), (Vad * n)()
    for s in range(n):
        O.orc_silk_vad_init(ctypes.byref(want[s]))
    sig = [vad_signal(rng, fs_kHz, frame_length, nframes) for _ in range(n)]
    return O, rng, frame_length, recs, want, sig


def drive_vad(init, do_vad, hp_cutoff, n, fs_kHz, frame_ms, nframes, seed, pad=0):
    """Shared by the emulator and GPU tests: n streams, nframes frames, the record carried; init(recs), do_vad(recs, pcm [n, frame_length + pad], activity,
    frame_length) and hp_cutoff(recs) run the path under test in place, the oracle runs beside it.  Returns (identical stream-frames, total, active)."""
    O, rng, frame_length, recs, want, sig = run_vad_streams(n, fs_kHz, frame_ms, nframes, seed)
    init(recs)
    assert bytes(recs) == bytes(want)
    same = active = 0
    for f in range(nframes):
        pcm = np.zeros((n, frame_length + pad), np.int16)
        pcm[:, :frame_length] = np.stack([sig[s][f] for s in range(n)])
        activity = rng.integers(0, 2, n).astype(np.int32) if f % 4 == 3 else np.ones(n, np.int32)
        for s in range(n):
            O.orc_silk_encode_do_vad(ctypes.byref(want[s]), pcm[s].ctypes.data_as(ctypes.POINTER(ctypes.c_int16)), frame_length, fs_kHz, int(activity[s]))
        do_vad(recs, pcm, activity, frame_length)
        for s in range(n):
            assert bytes(recs[s]) == bytes(want[s]), (f, s, recs[s].speech_activity_Q8, want[s].speech_activity_Q8)
            same += 1
            active += want[s].VAD_flag
            for v in (recs[s], want[s]):
                v.prevSignalType = int(2 if (v.VAD_flag and (f + s) % 3) else v.VAD_flag)
                v.prevLag = int(32 + (f * 37 + s * 11) % (18 * fs_kHz - 32))
            O.orc_silk_hp_variable_cutoff(ctypes.byref(want[s]), fs_kHz)
        hp_cutoff(recs)
        for s in range(n):
            assert bytes(recs[s]) == bytes(want[s]), (f, s, recs[s].variable_HP_smth1_Q15, want[s].variable_HP_smth1_Q15)
    return same, n * nframes, active
