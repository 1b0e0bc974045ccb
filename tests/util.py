"""Shared test plumbing: ctypes loaders for the oracle / the compiled reference, synthetic
signals, and fixture IO.  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, 'oracle')
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

OPUS_APPLICATION_VOIP = 2048
OPUS_APPLICATION_AUDIO = 2049
OPUS_APPLICATION_RESTRICTED_LOWDELAY = 2051
OPUS_APPLICATION_RESTRICTED_CELT = 2053
OPUS_SET_BANDWIDTH_REQUEST = 4008
OPUS_SET_FORCE_CHANNELS_REQUEST = 4022
OPUS_BANDWIDTH = {'nb': 1101, 'wb': 1103, 'swb': 1104, 'fb': 1105}


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def build_oracle():
    """(Re)build oracle/liboracle.so (and oracle/_ref when the reference is mounted)."""
    subprocess.check_call(['make', '-s', '-C', ORACLE_DIR, '-j8'], stdout=subprocess.DEVNULL)


_libs = {}


def load_oracle():
    if 'oracle' not in _libs:
        p = os.path.join(ORACLE_DIR, 'liboracle.so')
        if not os.path.exists(p):
            build_oracle()
        _libs['oracle'] = ctypes.CDLL(p)
    return _libs['oracle']


def ref_available(kind='c'):
    return os.path.exists(os.path.join(ORACLE_DIR, '_ref', 'libopus_ref_%s.so' % kind))


def load_ref(kind='c'):
    key = 'ref_' + kind
    if key not in _libs:
        p = os.path.join(ORACLE_DIR, '_ref', 'libopus_ref_%s.so' % kind)
        if not os.path.exists(p) and os.path.exists('/root/reference/celt/bands.c'):
            build_oracle()
        _libs[key] = ctypes.CDLL(p)
    return _libs[key]


# ---------------------------------------------------------------------------------------
# synthetic signals
# ---------------------------------------------------------------------------------------
def sig_music(seed, nsamp, channels):
    """Deterministic 'generate_music'-style test signal: a few drifting partials + noise bursts."""
    rng = np.random.default_rng(seed)
    t = np.arange(nsamp) / 48000.0
    x = np.zeros((nsamp, channels))
    for k in range(6):
        f0 = rng.uniform(80, 6000)
        am = 0.5 + 0.5 * np.sin(2 * np.pi * rng.uniform(0.5, 8) * t + rng.uniform(0, 6))
        ph = rng.uniform(0, 6, size=channels)
        for c in range(channels):
            x[:, c] += rng.uniform(500, 5000) * am * np.sin(2 * np.pi * f0 * t * (1 + 0.01 * np.sin(2 * np.pi * 0.7 * t)) + ph[c])
    # transients
    for k in range(max(1, nsamp // 9600)):
        p = rng.integers(0, nsamp - 200)
        x[p:p + 200] += rng.standard_normal((200, channels)) * rng.uniform(2000, 12000) * np.exp(-np.arange(200) / 40.0)[:, None]
    x += 200 * rng.standard_normal((nsamp, channels))
    return np.clip(x, -32768, 32767).astype(np.int16)


def sig_tonal_noise(seed, nsamp, channels):
    rng = np.random.default_rng(seed)
    t = np.arange(nsamp) / 48000.0
    brown = np.cumsum(rng.standard_normal((nsamp, channels)), axis=0)
    brown -= np.linspace(brown[0], brown[-1], nsamp)
    brown = brown / (np.abs(brown).max() + 1e-9)
    x = 0.3 * 20000 * brown
    x += (6000 * (1 + 0.5 * np.sin(2 * np.pi * 3 * t)) * np.sin(2 * np.pi * 440 * t))[:, None]
    ph = np.array([0, 2.5])[:channels]
    x += 3000 * np.sin(2 * np.pi * 3000 * t[:, None] + ph[None, :])
    return np.clip(x, -32768, 32767).astype(np.int16)


def sig_white(seed, nsamp, channels, dbfs=-20):
    rng = np.random.default_rng(seed)
    a = 32768 * 10 ** (dbfs / 20.0)
    return np.clip(a * rng.standard_normal((nsamp, channels)), -32768, 32767).astype(np.int16)


SIGNALS = {'music': sig_music, 'tonal': sig_tonal_noise, 'white': sig_white}


# ---------------------------------------------------------------------------------------
# reference drivers
# ---------------------------------------------------------------------------------------
def ref_encode(pcm, channels, frame_size, bitrate, vbr=0, complexity=10, application=OPUS_APPLICATION_RESTRICTED_CELT,
               ctls=(), kind='c', max_bytes=1500):
    L = load_ref(kind)
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    T = pcm.shape[0] // frame_size
    out = np.zeros((T, max_bytes), np.uint8)
    sizes = np.zeros(T, np.int32)
    rngs = np.zeros(T, np.uint32)
    c = np.array([v for p in ctls for v in p], dtype=np.int32)
    r = L.refh_encode_stream_ex(48000, channels, application, int(bitrate), int(vbr), int(complexity), frame_size, T,
                                _ptr(pcm), _ptr(out), max_bytes, _ptr(sizes), _ptr(rngs), _ptr(c), len(ctls))
    assert r == 0, r
    return out, sizes, rngs


def ref_decode_float(pkts, sizes, channels, frame_size, kind='c'):
    L = load_ref(kind)
    T = len(sizes)
    pcm = np.zeros((T * frame_size, channels), np.float32)
    rngs = np.zeros(T, np.uint32)
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    r = L.refh_decode_stream_float(48000, channels, frame_size, T, _ptr(pkts), pkts.shape[1], _ptr(sizes), _ptr(pcm), _ptr(rngs))
    assert r == 0, r
    return pcm, rngs


def oracle_decode_float(pkts, sizes, channels, frame_size):
    L = load_oracle()
    T = len(sizes)
    pcm = np.zeros((T * frame_size, channels), np.float32)
    rngs = np.zeros(T, np.uint32)
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    r = L.orc_decode_stream_float(channels, frame_size, T, _ptr(pkts), pkts.shape[1], _ptr(sizes), _ptr(pcm), _ptr(rngs))
    assert r == 0, r
    return pcm, rngs


def to_int16(pcm_float):
    x = np.clip(pcm_float.astype(np.float32) * np.float32(32768.0), -32768, 32767)
    return np.rint(x).astype(np.int16)
