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


# ---- SILK NSQ test-case generator (shared by the oracle-vs-reference, emulator and GPU tests) --------------------------------
NSQ_FRAME_DTYPE = np.dtype([('signalType', '<i4'), ('quantOffsetType', '<i4'), ('NLSFInterpCoef_Q2', '<i4'), ('Seed', '<i4'),
                            ('HarmShapeGain_Q14', '<i4', 4), ('Tilt_Q14', '<i4', 4), ('LF_shp_Q14', '<i4', 4), ('Gains_Q16', '<i4', 4),
                            ('pitchL', '<i4', 4), ('Lambda_Q10', '<i4'), ('LTP_scale_Q14', '<i4'),
                            ('PredCoef_Q12', '<i2', 32), ('LTPCoef_Q14', '<i2', 20), ('AR_Q13', '<i2', 96)])
NSQ_STATE_EXPORT_WORDS = 640 + 640 + 16 + 24 + 5


def nsq_cfg(fs_khz, nb_subfr, predict_order, shaping_order, warping_Q16, n_states):
    return np.array([fs_khz, nb_subfr, predict_order, shaping_order, warping_Q16, n_states], np.int32)


def _lpc_levinson(x, order):
    r = np.array([np.dot(x[:len(x) - k], x[k:]) for k in range(order + 1)], np.float64)
    r[0] = r[0] * 1.0001 + 1.0
    a = np.zeros(order + 1)
    a[0] = 1.0
    e = r[0]
    for i in range(1, order + 1):
        k = -(r[i] + np.dot(a[1:i], r[i - 1:0:-1])) / e
        a[1:i] = a[1:i] + k * a[i - 1:0:-1]
        a[i] = k
        e *= (1 - k * k)
    return -a[1:]          # prediction coefficients: x[n] ~ sum a[k] x[n-1-k]


def make_nsq_frames(seed, cfg, n_frames):
    """Plausible (stable, in-range) NSQ inputs for n_frames consecutive frames of one stream.
    Returns (x16 [n_frames, frame_length] int16, frames structured array [n_frames])."""
    rng = np.random.default_rng(seed)
    fs, nb, po, so = int(cfg[0]), int(cfg[1]), int(cfg[2]), int(cfg[3])
    sub = 5 * fs
    fl = sub * nb
    n = fl * n_frames
    t = np.arange(n)
    pitch = rng.integers(2 * fs + 4, 18 * fs - 4)
    sig = np.zeros(n)
    for h in range(1, 9):
        sig += rng.uniform(0.2, 1.0) / h * np.sin(2 * np.pi * h * t / pitch + rng.uniform(0, 6.28))
    sig = sig * rng.uniform(500, 6000) + rng.standard_normal(n) * rng.uniform(20, 1500)
    env = 0.3 + 0.7 * np.abs(np.sin(2 * np.pi * t / (fl * 2.7) + rng.uniform(0, 3)))
    x = np.clip(np.round(sig * env), -32768, 32767).astype(np.int16).reshape(n_frames, fl)
    frames = np.zeros(n_frames, NSQ_FRAME_DTYPE)
    for f in range(n_frames):
        fr = frames[f]
        st = int(rng.choice([0, 1, 2, 2]))
        fr['signalType'] = st
        fr['quantOffsetType'] = rng.integers(0, 2)
        fr['NLSFInterpCoef_Q2'] = rng.choice([4, 4, 0, 1, 2, 3])
        fr['Seed'] = rng.integers(0, 4)
        xf = x[f].astype(np.float64)
        for half in range(2):
            seg = xf[half * (fl // 2):(half + 1) * (fl // 2)] if nb == 4 else xf
            a = _lpc_levinson(seg, po) * 0.98 ** np.arange(1, po + 1)
            fr['PredCoef_Q12'][half * 16:half * 16 + po] = np.clip(np.round(a * 4096), -32768, 32767)
        for k in range(nb):
            seg = xf[k * sub:(k + 1) * sub]
            a = _lpc_levinson(seg, so)
            ar = a * 0.94 ** np.arange(1, so + 1) * rng.uniform(0.5, 1.0)
            fr['AR_Q13'][k * 24:k * 24 + so] = np.clip(np.round(ar * 8192), -32768, 32767)
            g = 2.0 ** rng.uniform(5, 12.5)
            fr['Gains_Q16'][k] = int(g * 65536) if rng.random() > 0.25 or k == 0 else fr['Gains_Q16'][k - 1]
            fr['Tilt_Q14'][k] = int(-rng.uniform(0.05, 0.45) * 16384)
            b = 0.2 / fs + 3.0 / max(pitch, 1)
            ma = int(round((-1.0 + b) * 16384))
            arq = int(round((1.0 - b - b * rng.uniform(0, 0.6)) * 16384))
            fr['LF_shp_Q14'][k] = np.array((arq << 16) | (ma & 0xFFFF), np.int64).astype(np.uint32).view(np.int32) if False else \
                np.int32(np.uint32(((arq & 0xFFFF) << 16) | (ma & 0xFFFF)))
            if st == 2:
                fr['pitchL'][k] = int(np.clip(pitch + rng.integers(-3, 4), 2 * fs, 18 * fs))
                taps = np.array([0.05, 0.2, 0.45, 0.2, 0.05]) * rng.uniform(0.2, 1.6) + rng.uniform(-0.05, 0.05, 5)
                fr['LTPCoef_Q14'][k * 5:k * 5 + 5] = np.round(taps * 16384)
                fr['HarmShapeGain_Q14'][k] = int(rng.uniform(0, 0.45) * 16384)
        fr['Lambda_Q10'] = rng.integers(900, 2900)
        fr['LTP_scale_Q14'] = rng.choice([15565, 12288, 8192])
    return x, frames


# ---- dnn layer generators (shared by oracle-vs-reference, emulator and GPU tests) ---------------------------------------------
class CLinear(ctypes.Structure):  # dnn/nnet.h:66-76 LinearLayer == OrcLinear == OpusB200LinearLayer
    _fields_ = [('bias', ctypes.c_void_p), ('subias', ctypes.c_void_p), ('weights', ctypes.c_void_p), ('float_weights', ctypes.c_void_p),
                ('weights_idx', ctypes.c_void_p), ('diag', ctypes.c_void_p), ('scale', ctypes.c_void_p), ('nb_inputs', ctypes.c_int),
                ('nb_outputs', ctypes.c_int)]


class CConv2d(ctypes.Structure):  # dnn/nnet.h:79-86
    _fields_ = [('bias', ctypes.c_void_p), ('float_weights', ctypes.c_void_p), ('in_channels', ctypes.c_int), ('out_channels', ctypes.c_int),
                ('ktime', ctypes.c_int), ('kheight', ctypes.c_int)]


class Layer:
    """A random LinearLayer with its arrays kept alive; .c is the C struct."""

    def __init__(self, seed, nb_in, nb_out, kind, bias=True, diag=False, density=0.3):
        rng = np.random.default_rng(seed)
        self.nb_in, self.nb_out, self.kind = nb_in, nb_out, kind
        self.bias = (rng.standard_normal(nb_out) * 0.1).astype(np.float32) if bias else None
        self.diag = (rng.standard_normal(nb_out) * 0.3).astype(np.float32) if diag else None
        self.weights = self.float_weights = self.idx = self.scale = None
        sparse = kind in ('f32_sparse', 'i8_sparse')
        if sparse:
            idx, nblk = [], 0
            for rb in range(nb_out // 8):
                cols = [4 * c for c in range(nb_in // 4) if rng.random() < density]
                idx.append(len(cols))
                idx.extend(cols)
                nblk += len(cols)
            self.idx = np.array(idx, np.int32)
        else:
            nblk = (nb_out // 8) * (nb_in // 4)
        if kind == 'f32':
            self.float_weights = (rng.standard_normal((nb_in, nb_out)) / np.sqrt(nb_in)).astype(np.float32)  # [j*N + i]
        elif kind == 'f32_sparse':
            self.float_weights = (rng.standard_normal(nblk * 32) / np.sqrt(nb_in * density)).astype(np.float32)
        else:
            self.weights = rng.integers(-127, 128, size=max(nblk, 1) * 32).astype(np.int8)
            self.scale = (rng.uniform(0.5, 1.5, nb_out) / (127.0 * 127.0 * np.sqrt(nb_in) * 0.6)).astype(np.float32)
        self.subias = None
        if self.weights is not None and self.bias is not None:
            # what the reference's weight exporter stores next to an int8 matrix: the bias compensated for the +127 offset
            if sparse:
                wsum, pos, blk = np.zeros(nb_out, np.float64), 0, 0
                for rb in range(nb_out // 8):
                    cnt = int(self.idx[pos]); pos += 1 + cnt
                    for _ in range(cnt):
                        wsum[rb * 8:rb * 8 + 8] += self.weights[blk * 32:blk * 32 + 32].reshape(8, 4).sum(axis=1)
                        blk += 1
            else:
                wsum = self.weights.reshape(nb_out // 8, nb_in // 4, 8, 4).astype(np.float64).sum(axis=(1, 3)).reshape(-1)
            self.subias = (self.bias - 127.0 * self.scale * wsum).astype(np.float32)
        c = CLinear()
        for name, arr in (('bias', self.bias), ('subias', self.subias), ('weights', self.weights), ('float_weights', self.float_weights), ('weights_idx', self.idx),
                          ('diag', self.diag), ('scale', self.scale)):
            setattr(c, name, arr.ctypes.data if arr is not None else None)
        c.nb_inputs, c.nb_outputs = nb_in, nb_out
        self.c = c

    def ptr(self, arr):
        return arr.ctypes.data_as(ctypes.c_void_p) if arr is not None else None


def sig_harmonic(seed, n, ch):
    """voiced-speech-like: a drifting pitch with many harmonics plus a little noise, so that the pitch pre-filter switches on"""
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    f0 = 110 + 60 * np.sin(2 * np.pi * t / 48000 * 0.7 + rng.uniform(0, 6))
    ph = 2 * np.pi * np.cumsum(f0) / 48000
    x = np.zeros(n)
    for h in range(1, 25):
        x += np.sin(h * ph + rng.uniform(0, 6)) / h
    x = 6000 * x + 150 * rng.standard_normal(n)
    if ch == 2:
        x = np.stack([x, 0.7 * np.roll(x, 13) + 100 * rng.standard_normal(n)], axis=1)
    return np.clip(x, -32768, 32767).reshape(-1)


def sig_swell(seed, n, ch, period=2900, rise=700):
    """noise bursts with a soft attack: too slow for transient_analysis, fast enough for patch_transient_decision now and then"""
    rng = np.random.default_rng(seed)
    ph = (np.arange(n) + rng.integers(0, period)) % period
    g = np.where(ph < rise, 0.02 + 0.98 * (ph / rise) ** 2, np.maximum(0.02, 1 - (ph - rise) / (period - rise) * 1.2))
    x = 5000 * g[:, None] * rng.standard_normal((n, ch))
    return np.clip(x, -32768, 32767).reshape(-1)
