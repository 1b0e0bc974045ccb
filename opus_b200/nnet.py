"""Host-side mirror of the reference's dnn/nnet.h layer API (LinearLayer, Conv2dLayer, compute_generic_*) with a leading
batch dimension: n streams share one set of weights.  numpy in / numpy out through the C ABI."""
import ctypes

import numpy as np

from . import capi

ACTIVATION_LINEAR, ACTIVATION_SIGMOID, ACTIVATION_TANH, ACTIVATION_RELU, ACTIVATION_SOFTMAX, ACTIVATION_SWISH, ACTIVATION_EXP = range(7)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class _CLinear(ctypes.Structure):  # OpusB200LinearLayer
    _fields_ = [('bias', ctypes.c_void_p), ('subias', ctypes.c_void_p), ('weights', ctypes.c_void_p), ('float_weights', ctypes.c_void_p),
                ('weights_idx', ctypes.c_void_p), ('diag', ctypes.c_void_p), ('scale', ctypes.c_void_p), ('nb_inputs', ctypes.c_int),
                ('nb_outputs', ctypes.c_int)]


class _CConv2d(ctypes.Structure):  # OpusB200Conv2dLayer
    _fields_ = [('bias', ctypes.c_void_p), ('float_weights', ctypes.c_void_p), ('in_channels', ctypes.c_int), ('out_channels', ctypes.c_int),
                ('ktime', ctypes.c_int), ('kheight', ctypes.c_int)]


class LinearLayer:
    """Device copy of a LinearLayer (dnn/nnet.h:66-76).  Arrays as in the reference's weight blobs."""

    def __init__(self, ops, nb_inputs, nb_outputs, bias=None, subias=None, weights=None, float_weights=None, weights_idx=None, diag=None,
                 scale=None, su_bias=True):
        self._ops, self._L = ops, ops._L
        self.nb_inputs, self.nb_outputs = nb_inputs, nb_outputs
        keep = []

        def arr(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dt)
            keep.append(a)
            return a.ctypes.data

        c = _CLinear(arr(bias, np.float32), arr(subias, np.float32), arr(weights, np.int8), arr(float_weights, np.float32),
                     arr(weights_idx, np.int32), arr(diag, np.float32), arr(scale, np.float32), nb_inputs, nb_outputs)
        h = ctypes.c_void_p()
        r = self._L.opusb200_nnet_linear_create(ops._h, ctypes.byref(c), 1 if su_bias else 0, ctypes.byref(h))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_nnet_linear_create')
        self._h = h

    def close(self):
        if getattr(self, '_h', None):
            self._L.opusb200_nnet_linear_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Conv2dLayer:
    def __init__(self, ops, float_weights, bias, in_channels, out_channels, ktime, kheight):
        self._ops, self._L = ops, ops._L
        self.in_channels, self.out_channels, self.ktime, self.kheight = in_channels, out_channels, ktime, kheight
        w = np.ascontiguousarray(float_weights, np.float32)
        b = np.ascontiguousarray(bias, np.float32) if bias is not None else None
        c = _CConv2d(b.ctypes.data if b is not None else None, w.ctypes.data, in_channels, out_channels, ktime, kheight)
        h = ctypes.c_void_p()
        r = self._L.opusb200_nnet_conv2d_create(ops._h, ctypes.byref(c), ctypes.byref(h))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_nnet_conv2d_create')
        self._h = h

    def close(self):
        if getattr(self, '_h', None):
            self._L.opusb200_nnet_conv2d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Nnet:
    """compute_* entry points of dnn/nnet.h:88-93 over a batch."""

    def __init__(self, ops):
        self._ops, self._L, self._c = ops, ops._L, ops._h

    def _chk(self, r, what):
        if r != 0:
            raise capi.OpusB200Error(r, what)

    def linear(self, layer, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty((x.shape[0], layer.nb_outputs), np.float32)
        self._chk(self._L.opusb200_nnet_compute_linear(self._c, layer._h, _p(out), _p(x), x.shape[0], 0), 'nnet_compute_linear')
        return out

    def activation(self, x, act):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(x)
        self._chk(self._L.opusb200_nnet_compute_activation(self._c, _p(out), _p(x), x.shape[0], x.shape[1], act, 0), 'nnet_compute_activation')
        return out

    def dense(self, layer, x, act):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty((x.shape[0], layer.nb_outputs), np.float32)
        self._chk(self._L.opusb200_nnet_compute_dense(self._c, layer._h, _p(out), _p(x), x.shape[0], act, 0), 'nnet_compute_dense')
        return out

    def gru(self, input_weights, recurrent_weights, state, x):
        """returns the new state [n][N]"""
        x = np.ascontiguousarray(x, np.float32)
        st = np.ascontiguousarray(state, np.float32).copy()
        self._chk(self._L.opusb200_nnet_compute_gru(self._c, input_weights._h, recurrent_weights._h, _p(st), _p(x), x.shape[0], 0), 'nnet_compute_gru')
        return st

    def glu(self, layer, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(x)
        self._chk(self._L.opusb200_nnet_compute_glu(self._c, layer._h, _p(out), _p(x), x.shape[0], 0), 'nnet_compute_glu')
        return out

    def conv1d(self, layer, mem, x, act, dilation=1):
        """returns (out, new mem)"""
        x = np.ascontiguousarray(x, np.float32)
        m = np.ascontiguousarray(mem, np.float32).copy()
        out = np.empty((x.shape[0], layer.nb_outputs), np.float32)
        self._chk(self._L.opusb200_nnet_compute_conv1d(self._c, layer._h, _p(out), _p(m) if m.size else None, _p(x), x.shape[0], x.shape[1],
                                                      dilation, act, 0), 'nnet_compute_conv1d')
        return out, m

    def conv2d(self, layer, mem, x, height, hstride, act):
        x = np.ascontiguousarray(x, np.float32)
        m = np.ascontiguousarray(mem, np.float32).copy()
        out = np.zeros((x.shape[0], layer.out_channels * hstride), np.float32)
        self._chk(self._L.opusb200_nnet_compute_conv2d(self._c, layer._h, _p(out), _p(m) if m.size else None, _p(x), x.shape[0], height, hstride,
                                                      act, 0), 'nnet_compute_conv2d')
        return out, m


class Fargan:
    """Batched FARGAN vocoder (dnn/fargan.h:59-65): n streams share the model; `layers` is a ctypes array of 20 OpusB200LinearLayer
    (= the reference's generated FARGAN struct, in declaration order) or a pointer to one."""
    NB_FEATURES, FRAME_SIZE, CONT_SAMPLES = 20, 160, 320

    def __init__(self, ops, layers, n_streams, su_bias=False):
        self._ops, self._L, self.n = ops, ops._L, n_streams
        h = ctypes.c_void_p()
        r = self._L.opusb200_fargan_create(ops._h, layers, 20, 1 if su_bias else 0, n_streams, ctypes.byref(h))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_fargan_create')
        self._h = h

    def cont(self, pcm0, features0):
        """fargan_cont: pcm0 [n][320], features0 [n][5*20]"""
        pcm0 = np.ascontiguousarray(pcm0, np.float32).reshape(self.n, self.CONT_SAMPLES)
        features0 = np.ascontiguousarray(features0, np.float32).reshape(self.n, 5 * self.NB_FEATURES)
        r = self._L.opusb200_fargan_cont(self._h, _p(pcm0), _p(features0), 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_fargan_cont')

    def synthesize(self, features):
        """fargan_synthesize: features [n][20] -> pcm [n][160] float"""
        features = np.ascontiguousarray(features, np.float32).reshape(self.n, self.NB_FEATURES)
        out = np.empty((self.n, self.FRAME_SIZE), np.float32)
        r = self._L.opusb200_fargan_synthesize(self._h, _p(out), _p(features), 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_fargan_synthesize')
        return out

    def close(self):
        if getattr(self, '_h', None):
            self._L.opusb200_fargan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DredEncoder:
    """Batched DRED encoder network (dnn/dred_rdovae_enc.h): `layers` = ctypes array of the 24 LinearLayers of RDOVAEEnc."""

    def __init__(self, ops, layers, n_streams, latent_dim=25, state_dim=50, su_bias=False):
        self._L, self.n, self.latent_dim, self.state_dim = ops._L, n_streams, latent_dim, state_dim
        h = ctypes.c_void_p()
        r = self._L.opusb200_dred_encoder_create(ops._h, layers, 24, 1 if su_bias else 0, latent_dim, state_dim, n_streams, ctypes.byref(h))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_dred_encoder_create')
        self._h = h

    def encode_dframe(self, x):
        """x [n][40] (two feature frames) -> (latents [n][latent_dim], initial_state [n][state_dim])"""
        x = np.ascontiguousarray(x, np.float32).reshape(self.n, 40)
        lat, st = np.empty((self.n, self.latent_dim), np.float32), np.empty((self.n, self.state_dim), np.float32)
        r = self._L.opusb200_dred_encode_dframe(self._h, _p(lat), _p(st), _p(x), 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_dred_encode_dframe')
        return lat, st

    def init_frontend(self, Fs, channels, linear, conv, su_bias=False):
        """dred_encoder_init + the pitch model (the PitchDNN layers, as LpcnetEncoder): enables compute_latents"""
        r = self._L.opusb200_dred_encoder_init_frontend(self._h, Fs, channels, linear, 6, conv, 2, 1 if su_bias else 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_dred_encoder_init_frontend')
        self.channels = channels

    def compute_latents(self, pcm, frame_size, extra_delay=0):
        """dred_compute_latents: pcm [n][frame_size * channels] float in [-1, 1)"""
        pcm = np.ascontiguousarray(pcm, np.float32).reshape(self.n, frame_size * self.channels)
        r = self._L.opusb200_dred_compute_latents(self._h, _p(pcm), frame_size, extra_delay, 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_dred_compute_latents')

    def load_stats(self, tables):
        """the statistical model: ctypes array of 8 uint8 pointers (latent quant_scales / dead_zone / r / p0, then the state's)"""
        r = self._L.opusb200_dred_encoder_load_stats(self._h, tables, 8)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_dred_encoder_load_stats')

    def encode_silk_frame(self, activity_mem, max_chunks, max_bytes, q0, dQ, qmax):
        """dred_encode_silk_frame: activity_mem [n][416] uint8 -> (payload [n][max_bytes] uint8, nbytes [n])"""
        act = np.ascontiguousarray(activity_mem, np.uint8).reshape(self.n, 416)
        buf, nb = np.zeros((self.n, max_bytes), np.uint8), np.zeros(self.n, np.int32)
        r = self._L.opusb200_dred_encode_silk_frame(self._h, _p(buf), max_bytes, max_chunks, max_bytes, q0, dQ, qmax, _p(act), _p(nb), 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_dred_encode_silk_frame')
        return buf, nb

    def get_latents(self):
        """(latents [n][104 * latent_dim], state [n][104 * state_dim], info [4] = input_buffer_fill, dred_offset, latent_offset, latents_buffer_fill)"""
        lat, st = np.empty((self.n, 104 * self.latent_dim), np.float32), np.empty((self.n, 104 * self.state_dim), np.float32)
        info = np.zeros(4, np.int32)
        r = self._L.opusb200_dred_encoder_get_latents(self._h, _p(lat), _p(st), _p(info))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_dred_encoder_get_latents')
        return lat, st, info

    def reset(self):
        self._L.opusb200_dred_encoder_reset(self._h)

    def close(self):
        if getattr(self, '_h', None):
            self._L.opusb200_dred_encoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LpcnetEncoder:
    """Batched LPCNet feature extraction (dnn/lpcnet_enc.c): `linear` = ctypes array of the 6 LinearLayers of the PitchDNN struct
    (dense_if_upsampler_1, dense_if_upsampler_2, dense_downsampler, dense_final_upsampler, gru_1_input, gru_1_recurrent), `conv` = its two Conv2dLayers."""
    FRAME_SIZE, NB_TOTAL_FEATURES = 160, 36

    def __init__(self, ops, linear, conv, n_streams, su_bias=False):
        self._L, self.n = ops._L, n_streams
        h = ctypes.c_void_p()
        r = self._L.opusb200_lpcnet_encoder_create(ops._h, linear, 6, conv, 2, 1 if su_bias else 0, n_streams, ctypes.byref(h))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_lpcnet_encoder_create')
        self._h = h

    def compute_single_frame_features(self, pcm):
        """pcm [n][160] int16 or float (int16 range) -> features [n][36]"""
        is16 = np.asarray(pcm).dtype == np.int16
        pcm = np.ascontiguousarray(pcm, np.int16 if is16 else np.float32).reshape(self.n, 160)
        out = np.empty((self.n, 36), np.float32)
        r = self._L.opusb200_lpcnet_compute_single_frame_features(self._h, _p(pcm), 1 if is16 else 0, _p(out), 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_lpcnet_compute_single_frame_features')
        return out

    def get_state(self):
        """(if_features [n][88], xcorr_features [n][224], lpc [n][16], misc [n][3] = dnn_pitch, pitch_filt, mem_preemph) of the last frame"""
        bufs = [np.empty((self.n, k), np.float32) for k in (88, 224, 16, 3)]
        r = self._L.opusb200_lpcnet_encoder_get_state(self._h, *[_p(b) for b in bufs])
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_lpcnet_encoder_get_state')
        return bufs

    def reset(self):
        self._L.opusb200_lpcnet_encoder_reset(self._h)

    def close(self):
        if getattr(self, '_h', None):
            self._L.opusb200_lpcnet_encoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def burg_cepstral_analysis(ops, x):
    """burg_cepstral_analysis (dnn/freq.c:179) of n frames: x [n][160] float (int16 range) -> ceps [n][36]"""
    x = np.ascontiguousarray(x, np.float32)
    n = x.shape[0]
    out = np.empty((n, 36), np.float32)
    r = ops._L.opusb200_burg_cepstral_analysis(ops._h, _p(out), _p(x.reshape(n, 160)), n, 0)
    if r != 0:
        raise capi.OpusB200Error(r, 'opusb200_burg_cepstral_analysis')
    return out


class LpcnetPlc:
    """Batched neural packet-loss concealment (dnn/lpcnet_plc.c): `plc_layers` = ctypes array of the 6 LinearLayers of PLCModel, `fargan_layers` = the
    20 of FARGAN.  The streams take update / conceal calls in lockstep; each has its own FEC / DRED feature queue."""
    NB_FEATURES, FRAME_SIZE = 20, 160

    def __init__(self, ops, plc_layers, fargan_layers, n_streams, su_bias=False):
        self._L, self.n = ops._L, n_streams
        h = ctypes.c_void_p()
        r = self._L.opusb200_lpcnet_plc_create(ops._h, plc_layers, 6, fargan_layers, 20, 1 if su_bias else 0, n_streams, ctypes.byref(h))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_lpcnet_plc_create')
        self._h = h

    def _chk(self, r, what):
        if r != 0:
            raise capi.OpusB200Error(r, what)

    def load_pitch_model(self, linear, conv, su_bias=False):
        """the PitchDNN layers (as LpcnetEncoder): needed by the history analysis of the first concealed frame after good audio"""
        self._chk(self._L.opusb200_lpcnet_plc_load_pitch_model(self._h, linear, 6, conv, 2, 1 if su_bias else 0), 'opusb200_lpcnet_plc_load_pitch_model')

    def reset(self):
        self._chk(self._L.opusb200_lpcnet_plc_reset(self._h), 'opusb200_lpcnet_plc_reset')

    def fec_add(self, features, have=None):
        """features [n][20]; have [n] int: 1 = this stream's vector, 0 = lpcnet_plc_fec_add(NULL), < 0 = no call (None: vectors for all)"""
        features = np.ascontiguousarray(features, np.float32).reshape(self.n, self.NB_FEATURES)
        hv = None if have is None else np.ascontiguousarray(have, np.int32)
        self._chk(self._L.opusb200_lpcnet_plc_fec_add(self._h, _p(features), _p(hv) if hv is not None else None, 0), 'opusb200_lpcnet_plc_fec_add')

    def fec_clear(self):
        self._chk(self._L.opusb200_lpcnet_plc_fec_clear(self._h), 'opusb200_lpcnet_plc_fec_clear')

    def update(self, pcm):
        pcm = np.ascontiguousarray(pcm, np.int16).reshape(self.n, self.FRAME_SIZE)
        self._chk(self._L.opusb200_lpcnet_plc_update(self._h, _p(pcm), 0), 'opusb200_lpcnet_plc_update')

    def conceal(self):
        out = np.empty((self.n, self.FRAME_SIZE), np.int16)
        self._chk(self._L.opusb200_lpcnet_plc_conceal(self._h, _p(out), 0), 'opusb200_lpcnet_plc_conceal')
        return out

    def state(self):
        bufs = [np.zeros((self.n, k), np.float32) for k in (20, 100, 128, 128)] + [np.zeros((self.n, 4), np.int32)]
        self._chk(self._L.opusb200_lpcnet_plc_get_state(self._h, *[_p(b) for b in bufs]), 'opusb200_lpcnet_plc_get_state')
        return bufs

    def close(self):
        if getattr(self, '_h', None):
            self._L.opusb200_lpcnet_plc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DredDecoder:
    """Batched DRED decoder network (dnn/dred_rdovae_dec.h): `layers` = ctypes array of the 29 LinearLayers of RDOVAEDec."""

    def __init__(self, ops, layers, n_streams, su_bias=False):
        self._L, self.n = ops._L, n_streams
        h = ctypes.c_void_p()
        r = self._L.opusb200_dred_decoder_create(ops._h, layers, 29, 1 if su_bias else 0, n_streams, ctypes.byref(h))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_dred_decoder_create')
        self._h = h

    def decode_all(self, state, latents):
        """state [n][state_dim], latents [n][nb_latents][latent_dim+1] -> features [n][nb_latents][80]"""
        state = np.ascontiguousarray(state, np.float32)
        latents = np.ascontiguousarray(latents, np.float32)
        nb = latents.shape[1]
        out = np.empty((self.n, nb, 80), np.float32)
        r = self._L.opusb200_dred_decode_all(self._h, _p(out), _p(state), _p(latents), nb, 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_dred_decode_all')
        return out

    def close(self):
        if getattr(self, '_h', None):
            self._L.opusb200_dred_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
