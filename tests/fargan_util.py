"""Shared by the FARGAN tests: the reference's dnn/fargan.c built with seeded random weights (oracle/_ref/libfargan_ref.so, see
oracle/gen_fargan_weights.py) is both the pin and the weight source: its 20 LinearLayers are read back through the harness."""
import ctypes
import os

import numpy as np

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NL = 20


def load_ref():
    p = os.path.join(ROOT, 'oracle', '_ref', 'libfargan_ref.so')
    if not os.path.exists(p):
        return None
    L = ctypes.CDLL(p)
    L.refh_fargan_state_size.restype = ctypes.c_int
    return L


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class RefFargan:
    def __init__(self, L):
        self.L = L
        self.st = np.zeros(L.refh_fargan_state_size() + 64, np.uint8)
        L.refh_fargan_init(vp(self.st))

    def cont(self, pcm0, feat0):
        self.L.refh_fargan_cont(vp(self.st), vp(pcm0), vp(feat0))

    def synth(self, feat):
        out = np.zeros(160, np.float32)
        self.L.refh_fargan_synthesize(vp(self.st), vp(out), vp(feat))
        return out

    def layers(self):
        """(ctypes array of 20 CLinear pointing into the reference library's own weight arrays, list of (nb_in, nb_out))"""
        arr = (util.CLinear * NL)()
        ptrs, dims = (ctypes.c_void_p * 7)(), (ctypes.c_int * 2)()
        shapes = []
        for i in range(NL):
            self.L.refh_fargan_layer(vp(self.st), i, ptrs, dims)
            for k, name in enumerate(('bias', 'subias', 'weights', 'float_weights', 'weights_idx', 'diag', 'scale')):
                setattr(arr[i], name, ptrs[k])
            arr[i].nb_inputs, arr[i].nb_outputs = dims[0], dims[1]
            shapes.append((dims[0], dims[1]))
        return arr, shapes

    def state(self):
        bufs = [np.zeros(k, np.float32) for k in (256, 128, 328, 160, 128, 128, 2)]
        self.L.refh_fargan_get_state(vp(self.st), *[vp(b) for b in bufs])
        return bufs


def make_inputs(seed, nframes):
    """continuation PCM (320), 5 priming feature vectors, then nframes feature vectors; feature[18] steers the pitch period"""
    rng = np.random.default_rng(seed)
    pcm0 = (0.1 * rng.standard_normal(320)).astype(np.float32)
    feat0 = (0.5 * rng.standard_normal((5, 20))).astype(np.float32)
    feats = (0.5 * rng.standard_normal((nframes, 20))).astype(np.float32)
    feat0[:, 18] = rng.uniform(-1.4, 1.4, 5)
    feats[:, 18] = rng.uniform(-1.4, 1.4, nframes)     # periods from ~34 to ~240 samples
    return pcm0, feat0.reshape(-1), feats


class RefDred:
    """the reference's DRED decoder network (dnn/dred_rdovae_dec.c) with the seeded random model of oracle/gen_dred_weights.py"""
    NL = 29

    def __init__(self, L):
        self.L = L
        self.nf, self.latent, self.state = (L.refh_dred_const(i) for i in range(3))

    def layers(self):
        arr = (util.CLinear * self.NL)()
        ptrs, dims = (ctypes.c_void_p * 7)(), (ctypes.c_int * 2)()
        shapes = []
        for i in range(self.NL):
            self.L.refh_dred_layer(i, ptrs, dims)
            for k, name in enumerate(('bias', 'subias', 'weights', 'float_weights', 'weights_idx', 'diag', 'scale')):
                setattr(arr[i], name, ptrs[k])
            arr[i].nb_inputs, arr[i].nb_outputs = dims[0], dims[1]
            shapes.append((dims[0], dims[1], bool(ptrs[4])))
        return arr, shapes

    def decode_all(self, state, latents):
        nb = latents.shape[0]
        out = np.zeros(nb * 4 * self.nf, np.float32)
        self.L.refh_dred_decode_all(vp(out), vp(np.ascontiguousarray(state, np.float32)), vp(np.ascontiguousarray(latents, np.float32)), nb)
        return out.reshape(nb, 4 * self.nf)


def make_dred_inputs(seed, nb_latents, latent_dim, state_dim):
    rng = np.random.default_rng(seed)
    state = (0.5 * rng.standard_normal(state_dim)).astype(np.float32)
    latents = (0.5 * rng.standard_normal((nb_latents, latent_dim + 1))).astype(np.float32)
    latents[:, -1] = rng.uniform(0, 1, nb_latents)     # the quantiser-level input the encoder appends
    return state, latents


class RefPlc:
    """the reference's neural packet-loss concealment (dnn/lpcnet_plc.c) with the seeded random predictor of oracle/gen_plc_weights.py and the
    FARGAN model above; the history analysis is not part of the build (oracle/fargan_harness.c): only sequences without good audio before a loss"""
    NL = 6

    def __init__(self, L):
        self.L = L
        self.st = np.zeros(L.refh_plc_state_size() + 64, np.uint8)
        assert L.refh_plc_init(vp(self.st)) == 0

    def _layers(self, fn, count):
        arr = (util.CLinear * count)()
        ptrs, dims = (ctypes.c_void_p * 7)(), (ctypes.c_int * 2)()
        shapes = []
        for i in range(count):
            fn(vp(self.st), i, ptrs, dims)
            for k, name in enumerate(('bias', 'subias', 'weights', 'float_weights', 'weights_idx', 'diag', 'scale')):
                setattr(arr[i], name, ptrs[k])
            arr[i].nb_inputs, arr[i].nb_outputs = dims[0], dims[1]
            shapes.append((dims[0], dims[1]))
        return arr, shapes

    def layers(self):
        return self._layers(self.L.refh_plc_layer, self.NL)

    def fargan_layers(self):
        return self._layers(self.L.refh_plc_fargan_layer, NL)

    def fec_add(self, features):
        self.L.refh_plc_fec_add(vp(self.st), vp(np.ascontiguousarray(features, np.float32)) if features is not None else None)

    def fec_clear(self):
        self.L.refh_plc_fec_clear(vp(self.st))

    def reset(self):
        self.L.refh_plc_reset(vp(self.st))

    def conceal(self):
        pcm = np.zeros(160, np.int16)
        assert self.L.refh_plc_conceal(vp(self.st), vp(pcm)) == 0
        return pcm

    def state(self):
        bufs = [np.zeros(k, np.float32) for k in (20, 100, 128, 128, 2400)] + [np.zeros(8, np.int32)]
        self.L.refh_plc_get_state(vp(self.st), *[vp(b) for b in bufs])
        return bufs


def plc_schedule(seed, nframes):
    """a DRED-style recovery: which of the concealed frames get a feature vector (lpcnet_plc_fec_add) and which a gap (fec_add(NULL));
    returns a list of per-frame actions ('feat', vector) / ('skip',) / ('none',) applied BEFORE the conceal call of that frame"""
    rng = np.random.default_rng(seed)
    acts = []
    for t in range(nframes):
        u = rng.random()
        if u < 0.55:
            f = (0.5 * rng.standard_normal(20)).astype(np.float32)
            f[18] = rng.uniform(-1.4, 1.4)
            acts.append(('feat', f))
        elif u < 0.7:
            acts.append(('skip',))
        else:
            acts.append(('none',))
    return acts


class RefDredEnc:
    """the reference's DRED encoder network (dnn/dred_rdovae_enc.c) with the seeded random model of oracle/gen_dred_weights.py"""
    NL = 24

    def __init__(self, L):
        self.L = L
        self.st = np.zeros(L.refh_dred_enc_state_size() + 64, np.uint8)

    def layers(self):
        arr = (util.CLinear * self.NL)()
        ptrs, dims = (ctypes.c_void_p * 7)(), (ctypes.c_int * 2)()
        shapes = []
        for i in range(self.NL):
            self.L.refh_dred_enc_layer(i, ptrs, dims)
            for k, name in enumerate(('bias', 'subias', 'weights', 'float_weights', 'weights_idx', 'diag', 'scale')):
                setattr(arr[i], name, ptrs[k])
            arr[i].nb_inputs, arr[i].nb_outputs = dims[0], dims[1]
            shapes.append((dims[0], dims[1], bool(ptrs[4])))
        return arr, shapes

    def encode_dframe(self, x40):
        lat, st = np.zeros(25, np.float32), np.zeros(50, np.float32)
        self.L.refh_dred_encode_dframe(vp(self.st), vp(lat), vp(st), vp(np.ascontiguousarray(x40, np.float32)))
        return lat, st


class RefLpcnetEnc:
    """the reference's feature extraction (dnn/lpcnet_enc.c, freq.c, pitchdnn.c) with the seeded random pitch model of
    oracle/gen_pitchdnn_weights.py; also the weight source of the oracle and the device side"""

    def __init__(self, L):
        self.L = L
        L.refh_pitchdnn_compute.restype = ctypes.c_float
        self.st = np.zeros(L.refh_lpcnet_enc_state_size() + 64, np.uint8)
        L.refh_lpcnet_enc_init(vp(self.st))

    def linear(self):
        arr = (util.CLinear * 6)()
        ptrs, dims = (ctypes.c_void_p * 7)(), (ctypes.c_int * 2)()
        shapes = []
        for i in range(6):
            self.L.refh_pitchdnn_linear(vp(self.st), i, ptrs, dims)
            for k, name in enumerate(('bias', 'subias', 'weights', 'float_weights', 'weights_idx', 'diag', 'scale')):
                setattr(arr[i], name, ptrs[k])
            arr[i].nb_inputs, arr[i].nb_outputs = dims[0], dims[1]
            shapes.append((dims[0], dims[1]))
        return arr, shapes

    def conv(self):
        arr = (util.CConv2d * 2)()
        ptrs, dims = (ctypes.c_void_p * 2)(), (ctypes.c_int * 4)()
        for i in range(2):
            self.L.refh_pitchdnn_conv2d(vp(self.st), i, ptrs, dims)
            arr[i].bias, arr[i].float_weights = ptrs[0], ptrs[1]
            arr[i].in_channels, arr[i].out_channels, arr[i].ktime, arr[i].kheight = dims[0], dims[1], dims[2], dims[3]
        return arr

    def features(self, pcm):
        out = np.zeros(36, np.float32)
        self.L.refh_lpcnet_enc_features(vp(self.st), vp(np.ascontiguousarray(pcm, np.float32)), vp(out))
        return out

    def get(self):
        bufs = [np.zeros(k, np.float32) for k in (88, 224, 16, 3)]
        self.L.refh_lpcnet_enc_get(vp(self.st), *[vp(b) for b in bufs])
        return bufs


class OrcLpcnetEnc:
    """oracle/lpcnet_enc_oracle.c with the layers of a RefLpcnetEnc"""

    def __init__(self, O, ref):
        self.O, self.ref = O, ref
        O.orc_pitchdnn_compute.restype = ctypes.c_float
        self.lin, self.shapes = ref.linear()
        self.cv = ref.conv()
        self.st = np.zeros(O.orc_lpcnet_enc_size() + 64, np.uint8)
        O.orc_lpcnet_enc_init(vp(self.st), self.lin, self.cv)

    def reset(self):
        self.O.orc_lpcnet_enc_reset(vp(self.st))

    def features(self, pcm):
        out = np.zeros(36, np.float32)
        self.O.orc_lpcnet_enc_features(vp(self.st), vp(np.ascontiguousarray(pcm, np.float32)), vp(out))
        return out

    def get(self):
        bufs = [np.zeros(k, np.float32) for k in (88, 224, 16, 3)]
        self.O.orc_lpcnet_enc_get(vp(self.st), *[vp(b) for b in bufs])
        return bufs


def speech_like(seed, nframes, scale=8000.0):
    """16 kHz test signal in int16 range: a pitched pulse train through a resonance, noise bursts, level changes and silence"""
    rng = np.random.default_rng(seed)
    n = nframes * 160
    t = np.arange(n)
    f0 = rng.uniform(80, 300) * (1 + 0.2 * np.sin(2 * np.pi * t / rng.uniform(4000, 9000)))
    ph = np.cumsum(f0 / 16000.0)
    x = np.zeros(n)
    for h in range(1, 12):
        x += np.sin(2 * np.pi * h * ph + rng.uniform(0, 6.28)) / h
    x += 0.3 * rng.standard_normal(n)
    env = np.clip(np.sin(2 * np.pi * t / rng.uniform(3000, 8000) + rng.uniform(0, 6.28)) + 0.6, 0, 1)
    if rng.random() < 0.5:
        a = rng.integers(0, max(1, n - 400))
        env[a:a + 400] = 0
    return (scale * rng.uniform(0.05, 1.0) * env * x).astype(np.float32)


class RefDredFront:
    """the reference's DREDEnc object (dnn/dred_encoder.c) with the compiled-in seeded random models"""

    def __init__(self, L, Fs, channels):
        self.L = L
        self.st = np.zeros(L.refh_dredenc_size() + 64, np.uint8)
        L.refh_dredenc_init(vp(self.st), Fs, channels)
        assert L.refh_dredenc_loaded(vp(self.st)) == 1

    def compute_latents(self, pcm, frame_size, extra_delay=0):
        self.L.refh_dredenc_compute_latents(vp(self.st), vp(np.ascontiguousarray(pcm, np.float32)), frame_size, extra_delay)

    def get(self):
        bufs = [np.zeros(k, np.float32) for k in (104 * 25, 104 * 50, 640, 9)] + [np.zeros(5, np.int32)]
        self.L.refh_dredenc_get(vp(self.st), *[vp(b) for b in bufs])
        return bufs


class OrcDredFront:
    """oracle/dred_oracle.c's front end on the layers of the reference library"""

    def __init__(self, O, L, Fs, channels):
        self.O = O
        self.lp = OrcLpcnetEnc(O, RefLpcnetEnc(L))
        self.layers, _ = RefDredEnc(L).layers()
        self.net = np.zeros(O.orc_dred_enc_size() + 64, np.uint8)
        O.orc_dred_enc_init(vp(self.net), self.layers)
        self.st = np.zeros(O.orc_dredenc_size() + 64, np.uint8)
        O.orc_dredenc_init(vp(self.st), vp(self.net), vp(self.lp.st), Fs, channels)

    def compute_latents(self, pcm, frame_size, extra_delay=0):
        self.O.orc_dredenc_compute_latents(vp(self.st), vp(np.ascontiguousarray(pcm, np.float32)), frame_size, extra_delay)

    def reset(self):
        self.O.orc_dredenc_reset(vp(self.st))

    def get(self):
        bufs = [np.zeros(k, np.float32) for k in (104 * 25, 104 * 50, 640, 9)] + [np.zeros(5, np.int32)]
        self.O.orc_dredenc_get(vp(self.st), *[vp(b) for b in bufs])
        return bufs


def dred_input(seed, Fs, channels, nsamples):
    """float PCM in [-1, 1) at Fs: a 16 kHz speech-like signal held / interpolated to Fs, the second channel a delayed copy"""
    base = speech_like(seed, (nsamples * 16000 // Fs) // 160 + 2, 12000.0) / 32768.0
    t = np.arange(nsamples) * (16000.0 / Fs)
    x = np.interp(t, np.arange(base.size), base).astype(np.float32)
    if channels == 1:
        return x
    y = np.roll(x, 7) * np.float32(0.8)
    return np.stack([x, y], 1).reshape(-1).astype(np.float32)


def dred_stats(L):
    """the eight uint8 tables of the reference library's statistical model: (ctypes array of 8 pointers, list of numpy copies)"""
    L.refh_dred_stats.restype = ctypes.c_void_p
    ptrs = (ctypes.c_void_p * 8)()
    arrays = []
    for k in range(8):
        ln = ctypes.c_int()
        ptrs[k] = L.refh_dred_stats(k, ctypes.byref(ln))
        arrays.append(np.ctypeslib.as_array(ctypes.cast(ptrs[k], ctypes.POINTER(ctypes.c_uint8)), (ln.value,)).copy())
    return ptrs, arrays


def dred_activity(seed, kind):
    """activity_mem of dred_encode_silk_frame (DRED_MAX_FRAMES * 4 bytes, one per 2.5 ms, newest first)"""
    rng = np.random.default_rng(seed)
    a = np.zeros(416, np.uint8)
    if kind == 'all':
        a[:] = 1
    elif kind == 'bursts':
        t = 0
        while t < 416:
            k = int(rng.integers(4, 60))
            a[t:t + k] = rng.random() < 0.6
            t += k
    elif kind == 'late':      # silence now, speech further back
        a[int(rng.integers(20, 120)):] = 1
    return a
