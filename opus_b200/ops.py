"""Batched kernel-level operations (host-side mirror of the reference's per-call SIMD plug-in
functions, with a leading batch dimension).  numpy in / numpy out through the C ABI."""
import ctypes

import numpy as np

from . import capi


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Ops:
    def __init__(self, device=0):
        self._L = capi.lib()
        h = ctypes.c_void_p()
        r = self._L.opusb200_ctx_create(device, ctypes.byref(h))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_ctx_create')
        self._h = h

    def _chk(self, r, what):
        if r != 0:
            raise capi.OpusB200Error(r, what)

    def mdct_forward(self, x, LM, transient):
        """x: [n, (120<<LM)+120] float32 -> [n, 120<<LM]  (clt_mdct_forward via compute_mdcts)"""
        x = np.ascontiguousarray(x, np.float32)
        n, N = x.shape[0], 120 << LM
        assert x.shape[1] == N + 120
        out = np.empty((n, N), np.float32)
        self._chk(self._L.opusb200_mdct_forward(self._h, _p(x), _p(out), n, LM, int(transient), 0), 'mdct_forward')
        return out

    def pitch_xcorr(self, x, y, max_pitch):
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(y, np.float32)
        n, length = x.shape
        assert y.shape == (n, length + max_pitch)
        out = np.empty((n, max_pitch), np.float32)
        self._chk(self._L.opusb200_pitch_xcorr(self._h, _p(x), _p(y), _p(out), n, length, max_pitch, 0), 'pitch_xcorr')
        return out

    def exp_rotation(self, X, K, direction, stride, spread):
        X = np.array(X, np.float32, order='C')
        K = np.ascontiguousarray(K, np.int32)
        n, N = X.shape
        self._chk(self._L.opusb200_exp_rotation(self._h, _p(X), _p(K), n, N, direction, stride, spread, 0), 'exp_rotation')
        return X

    def pvq_search(self, X, K):
        """Returns (|X|, iy, yy) like op_pvq_search_c (which overwrites X with |X|)."""
        X = np.array(X, np.float32, order='C')
        K = np.ascontiguousarray(K, np.int32)
        n, N = X.shape
        iy = np.empty((n, N), np.int32)
        yy = np.empty(n, np.float32)
        self._chk(self._L.opusb200_pvq_search(self._h, _p(X), _p(iy), _p(yy), _p(K), n, N, 0), 'pvq_search')
        return X, iy, yy

    def silk_lpc_analysis_filter(self, x, B):
        """x [n, len] int16, B [n, d] int16 Q12 -> [n, len] int16 (silk_LPC_analysis_filter)"""
        x = np.ascontiguousarray(x, np.int16)
        B = np.ascontiguousarray(B, np.int16)
        out = np.empty_like(x)
        self._chk(self._L.opusb200_silk_lpc_analysis_filter(self._h, _p(out), _p(x), _p(B), x.shape[0], x.shape[1], B.shape[1], 0),
                  'silk_lpc_analysis_filter')
        return out

    def silk_burg_modified(self, x, min_inv_gain, subfr_length, nb_subfr, D):
        """x [n, nb_subfr*subfr_length] float32 -> (A [n, D], residual energy [n])"""
        x = np.ascontiguousarray(x, np.float32)
        n = x.shape[0]
        A = np.empty((n, D), np.float32)
        nrg = np.empty(n, np.float32)
        self._chk(self._L.opusb200_silk_burg_modified(self._h, _p(A), _p(nrg), _p(x), float(min_inv_gain), n, subfr_length,
                                                     nb_subfr, D, 0), 'silk_burg_modified')
        return A, nrg

    def silk_autocorrelation(self, x, count):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty((x.shape[0], count), np.float32)
        self._chk(self._L.opusb200_silk_autocorrelation(self._h, _p(out), _p(x), x.shape[0], x.shape[1], count, 0), 'silk_autocorrelation')
        return out

    def silk_resampler(self, n_streams, fs_in, fs_out, for_enc=True):
        return SilkResampler(self, n_streams, fs_in, fs_out, for_enc)

    def silk_process_gains(self, frames, nb_subfr, subfr_length, n_states_delayed_decision):
        """silk_process_gains_FLP for n streams: frames = ctypes array of OpusB200SilkGains-shaped records (processed in place)"""
        n = len(frames)
        r = self._L.opusb200_silk_process_gains(self._h, ctypes.cast(frames, ctypes.c_void_p), n, nb_subfr, subfr_length, n_states_delayed_decision, 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_silk_process_gains')
        return frames

    def silk_process_nlsfs(self, frames, fs_kHz, nb_subfr):
        """silk_process_NLSFs for n streams: frames = ctypes array of OpusB200SilkNlsf-shaped records (processed in place)"""
        r = self._L.opusb200_silk_process_nlsfs(self._h, ctypes.cast(frames, ctypes.c_void_p), len(frames), fs_kHz, nb_subfr, 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_silk_process_nlsfs')
        return frames

    def silk_find_lpc(self, frames, x, order, nb_subfr, subfr_length):
        """silk_find_LPC_FLP for n streams: frames = ctypes array of OpusB200SilkFindLpc-shaped records (in place), x float32 [n, nb_subfr * (subfr_length + order)]"""
        x = np.ascontiguousarray(x, np.float32)
        r = self._L.opusb200_silk_find_lpc(self._h, ctypes.cast(frames, ctypes.c_void_p), x.ctypes.data_as(ctypes.c_void_p), len(frames), order, nb_subfr,
                                           subfr_length, 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_silk_find_lpc')
        return frames

    def silk_find_pred_coefs(self, frames, res_pitch, x, frame_off, fs_kHz, nb_subfr):
        """silk_find_pred_coefs_FLP for n streams: frames = ctypes array of OpusB200SilkPredCoefs-shaped records (in place); res_pitch / x float32
        [n, stride] with the current frame at frame_off"""
        res_pitch, x = np.ascontiguousarray(res_pitch, np.float32), np.ascontiguousarray(x, np.float32)
        r = self._L.opusb200_silk_find_pred_coefs(self._h, ctypes.cast(frames, ctypes.c_void_p), res_pitch.ctypes.data_as(ctypes.c_void_p),
                                                  x.ctypes.data_as(ctypes.c_void_p), len(frames), x.shape[1], frame_off, fs_kHz, nb_subfr, 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_silk_find_pred_coefs')
        return frames

    def silk_noise_shape_analysis(self, frames, pitch_res, x, x_off, fs_kHz, nb_subfr, shaping_order):
        """silk_noise_shape_analysis_FLP for n streams: frames = ctypes array of OpusB200SilkNoiseShape-shaped records (in place); pitch_res float32
        [n, >= frame], x float32 [n, stride] with the frame at x_off"""
        pitch_res, x = np.ascontiguousarray(pitch_res, np.float32), np.ascontiguousarray(x, np.float32)
        r = self._L.opusb200_silk_noise_shape_analysis(self._h, ctypes.cast(frames, ctypes.c_void_p), pitch_res.ctypes.data_as(ctypes.c_void_p), pitch_res.shape[1],
                                                       x.ctypes.data_as(ctypes.c_void_p), x.shape[1], x_off, len(frames), fs_kHz, nb_subfr, shaping_order, 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_silk_noise_shape_analysis')
        return frames

    def silk_find_pitch_lags(self, frames, x, x_off, fs_kHz, nb_subfr, lpc_order, complexity):
        """silk_find_pitch_lags_FLP for n streams: frames = ctypes array of OpusB200SilkPitch-shaped records (in place), x float32 [n, stride] with the
        frame at x_off; returns the residual float32 [n, (22 + 5 nb_subfr) fs_kHz]"""
        x = np.ascontiguousarray(x, np.float32)
        res = np.zeros((len(frames), (22 + 5 * nb_subfr) * fs_kHz), np.float32)
        r = self._L.opusb200_silk_find_pitch_lags(self._h, ctypes.cast(frames, ctypes.c_void_p), res.ctypes.data_as(ctypes.c_void_p), res.shape[1],
                                                  x.ctypes.data_as(ctypes.c_void_p), x.shape[1], x_off, len(frames), fs_kHz, nb_subfr, lpc_order, complexity, 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_silk_find_pitch_lags')
        return res

    def silk_vad_init(self, frames):
        """silk_VAD_Init for n streams: frames = ctypes array of OpusB200SilkVad-shaped records (filled on the host)"""
        r = self._L.opusb200_silk_vad_init(ctypes.cast(frames, ctypes.c_void_p), len(frames))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_silk_vad_init')
        return frames

    def silk_encode_do_vad(self, frames, pcm, fs_kHz, activity=None, frame_length=None):
        """silk_encode_do_VAD_FLP for n streams: frames = ctypes array of OpusB200SilkVad-shaped records (in place), pcm int16 [n, frame_length]
        (10 or 20 ms at fs_kHz; or [n, stride] with frame_length given), activity int32 [n] = the Opus-level decision per stream or None"""
        pcm = np.ascontiguousarray(pcm, np.int16)
        act = None if activity is None else np.ascontiguousarray(activity, np.int32)
        r = self._L.opusb200_silk_encode_do_vad(self._h, ctypes.cast(frames, ctypes.c_void_p), pcm.ctypes.data_as(ctypes.c_void_p), pcm.shape[1],
                                                None if act is None else act.ctypes.data_as(ctypes.c_void_p), len(frames), frame_length or pcm.shape[1], fs_kHz, 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_silk_encode_do_vad')
        return frames

    def silk_hp_variable_cutoff(self, frames, fs_kHz):
        """silk_HP_variable_cutoff for n streams: updates variable_HP_smth1_Q15 of the voiced streams' records in place"""
        r = self._L.opusb200_silk_hp_variable_cutoff(self._h, ctypes.cast(frames, ctypes.c_void_p), len(frames), fs_kHz, 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_silk_hp_variable_cutoff')
        return frames

    def range_encoder(self, n_streams, buf_bytes=1275):
        return RangeEncoder(self, n_streams, buf_bytes)

    def celt_analysis(self, n_streams, channels):
        return CeltAnalysis(self, n_streams, channels)

    def silk_nsq(self, n_streams, fs_khz, nb_subfr, predict_order, shaping_order, warping_Q16, n_states):
        return SilkNsq(self, n_streams, (fs_khz, nb_subfr, predict_order, shaping_order, warping_Q16, n_states))

    def close(self):
        if getattr(self, '_h', None):
            self._L.opusb200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RangeEncoder:
    """n_streams range encoders with device-resident state (ec_enc) and the SILK excitation coder on them (silk_encode_pulses)."""

    def __init__(self, ops, n_streams, buf_bytes=1275):
        self._L, self.n = ops._L, n_streams
        h = ctypes.c_void_p()
        r = self._L.opusb200_range_encoder_create(ops._h, n_streams, buf_bytes, ctypes.byref(h))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_range_encoder_create')
        self._h = h
        self.storage = buf_bytes

    def _chk(self, r, what):
        if r != 0:
            raise capi.OpusB200Error(r, what)

    def init(self, storage=None):
        self.storage = storage or self.storage
        self._chk(self._L.opusb200_range_encoder_init(self._h, self.storage), 'opusb200_range_encoder_init')

    def icdf(self, symbols, icdf, ftb=8):
        symbols = np.ascontiguousarray(symbols, np.int32)
        icdf = np.ascontiguousarray(icdf, np.uint8)
        self._chk(self._L.opusb200_range_encoder_icdf(self._h, _p(symbols), _p(icdf), len(icdf), ftb, 0), 'opusb200_range_encoder_icdf')

    def encode_pulses(self, signal_type, quant_offset_type, pulses, frame_length):
        st = np.ascontiguousarray(signal_type, np.int32)
        qo = np.ascontiguousarray(quant_offset_type, np.int32)
        pulses = np.ascontiguousarray(pulses, np.int8)
        assert pulses.shape[0] == self.n
        self._chk(self._L.opusb200_silk_encode_pulses(self._h, _p(st), _p(qo), _p(pulses), pulses.shape[1], frame_length, 0), 'opusb200_silk_encode_pulses')

    def encode_indices(self, fs_kHz, nb_subfr, indices, cond_coding):
        """indices: a ctypes array of n SideInfoIndices-layout structs (or a uint8 array [n][36]); cond_coding [n]"""
        cc = np.ascontiguousarray(cond_coding, np.int32)
        ptr = _p(indices) if isinstance(indices, np.ndarray) else ctypes.cast(indices, ctypes.c_void_p)
        self._chk(self._L.opusb200_silk_encode_indices(self._h, fs_kHz, nb_subfr, ptr, _p(cc), 0), 'opusb200_silk_encode_indices')

    def reset_silk_state(self):
        self._chk(self._L.opusb200_silk_coder_reset_state(self._h), 'opusb200_silk_coder_reset_state')

    def done(self):
        """-> (bytes [n][storage], tell_frac [n], final_range [n], error [n])"""
        b = np.zeros((self.n, self.storage), np.uint8)
        tf, fr, er = np.zeros(self.n, np.uint32), np.zeros(self.n, np.uint32), np.zeros(self.n, np.int32)
        self._chk(self._L.opusb200_range_encoder_done(self._h, _p(b), _p(tf), _p(fr), _p(er)), 'opusb200_range_encoder_done')
        return b, tf, fr, er

    def close(self):
        if getattr(self, '_h', None):
            self._L.opusb200_range_encoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SilkResampler:
    """n_streams SILK resamplers (silk_resampler_init(S, fs_in, fs_out, for_enc) / silk_resampler)."""

    def __init__(self, ops, n_streams, fs_in, fs_out, for_enc=True):
        self._ops, self._L, self.n, self.fs_in, self.fs_out = ops, ops._L, n_streams, fs_in, fs_out
        h = ctypes.c_void_p()
        r = self._L.opusb200_silk_resampler_create_ex(ops._h, n_streams, fs_in, fs_out, 1 if for_enc else 0, ctypes.byref(h))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_silk_resampler_create')
        self._h = h

    def process(self, x):
        x = np.ascontiguousarray(x, np.int16)
        assert x.shape[0] == self.n
        out = np.empty((self.n, x.shape[1] * self.fs_out // self.fs_in), np.int16)
        r = self._L.opusb200_silk_resampler_process(self._h, _p(x), _p(out), x.shape[1], 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_silk_resampler_process')
        return out

    def close(self):
        if getattr(self, '_h', None):
            self._L.opusb200_silk_resampler_destroy(self._h)
            self._h = None


NSQ_FRAME_DTYPE = np.dtype([('signalType', '<i4'), ('quantOffsetType', '<i4'), ('NLSFInterpCoef_Q2', '<i4'), ('Seed', '<i4'),
                            ('HarmShapeGain_Q14', '<i4', 4), ('Tilt_Q14', '<i4', 4), ('LF_shp_Q14', '<i4', 4), ('Gains_Q16', '<i4', 4),
                            ('pitchL', '<i4', 4), ('Lambda_Q10', '<i4'), ('LTP_scale_Q14', '<i4'),
                            ('PredCoef_Q12', '<i2', 32), ('LTPCoef_Q14', '<i2', 20), ('AR_Q13', '<i2', 96)])  # OpusB200NsqFrame


class SilkNsq:
    """n_streams SILK noise-shaping quantizers (silk_NSQ / silk_NSQ_del_dec, chosen as silk_NSQ_wrapper_FLP does)."""
    STATE_WORDS = 1325

    def __init__(self, ops, n_streams, cfg):
        self._ops, self._L, self.n = ops, ops._L, n_streams
        self.cfg = np.array(cfg, np.int32)
        self.frame_length = 5 * int(cfg[0]) * int(cfg[1])
        h = ctypes.c_void_p()
        r = self._L.opusb200_silk_nsq_create(ops._h, n_streams, _p(self.cfg), ctypes.byref(h))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_silk_nsq_create')
        self._h = h

    def process(self, frames, x16):
        """frames: structured array [n] of NSQ_FRAME_DTYPE (Seed updated in the returned copy); x16 [n, frame_length].
        Returns (pulses int8 [n, frame_length], frames_out)."""
        fr = np.ascontiguousarray(frames, NSQ_FRAME_DTYPE).copy()
        x = np.ascontiguousarray(x16, np.int16)
        assert fr.shape == (self.n,) and x.shape == (self.n, self.frame_length)
        pulses = np.empty(x.shape, np.int8)
        r = self._L.opusb200_silk_nsq_process(self._h, _p(fr), _p(x), _p(pulses), 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_silk_nsq_process')
        return pulses, fr

    def state(self):
        out = np.empty((self.n, self.STATE_WORDS), np.int32)
        r = self._L.opusb200_silk_nsq_get_state(self._h, _p(out))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_silk_nsq_get_state')
        return out

    def reset(self):
        r = self._L.opusb200_silk_nsq_reset(self._h)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_silk_nsq_reset')

    def close(self):
        if getattr(self, '_h', None):
            self._L.opusb200_silk_nsq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CeltAnalysis:
    """n_streams CELT encoder front ends (pre-emphasis -> MDCT -> band energies -> normalised bands), state carried."""

    def __init__(self, ops, n_streams, channels):
        self._ops, self._L, self.n, self.channels = ops, ops._L, n_streams, channels
        h = ctypes.c_void_p()
        r = self._L.opusb200_celt_analysis_create(ops._h, n_streams, channels, ctypes.byref(h))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_celt_analysis_create')
        self._h = h

    def process(self, pcm, short_blocks=None):
        """pcm float32 [n, frame_size*channels] interleaved -> (X [n, channels*frame_size], bandE [n, channels*21], bandLogE)"""
        pcm = np.ascontiguousarray(pcm, np.float32)
        fs = pcm.shape[1] // self.channels
        X = np.empty_like(pcm)
        E = np.empty((self.n, self.channels * 21), np.float32)
        Lg = np.empty_like(E)
        sb = np.ascontiguousarray(short_blocks, np.int32) if short_blocks is not None else None
        r = self._L.opusb200_celt_analysis_process(self._h, _p(pcm), _p(sb) if sb is not None else None, _p(X), _p(E), _p(Lg), fs, 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_celt_analysis_process')
        return X, E, Lg

    def close(self):
        if getattr(self, '_h', None):
            self._L.opusb200_celt_analysis_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
