"""Host-side mirror of the reference CELT encoder API (celt_encoder_init / celt_encoder_ctl / celt_encode_with_ec) for a batch
of independent streams.  numpy in / numpy out through the C ABI."""
import ctypes

import numpy as np

from . import capi

OPUS_SET_BITRATE_REQUEST, OPUS_SET_VBR_REQUEST, OPUS_SET_COMPLEXITY_REQUEST, OPUS_SET_LSB_DEPTH_REQUEST = 4002, 4006, 4010, 4036
# celt/celt.h:97-120: the CELT-level requests the Opus layer sets per frame
CELT_SET_PREDICTION_REQUEST, CELT_SET_CHANNELS_REQUEST, CELT_SET_END_BAND_REQUEST = 10002, 10008, 10012
OPUS_SET_PHASE_INVERSION_DISABLED_REQUEST = 4046
OPUS_SET_VBR_CONSTRAINT_REQUEST, OPUS_SET_PACKET_LOSS_PERC_REQUEST = 4020, 4014


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class EncoderBatch:
    def __init__(self, n_streams, channels=2, fs=48000, device=0, bitrate=None, complexity=5, lsb_depth=24, vbr=False, constrained_vbr=True):
        self._L = capi.lib()
        self.n_streams, self.channels = n_streams, channels
        h = ctypes.c_void_p()
        r = self._L.opusb200_encoder_create(device, fs, channels, n_streams, ctypes.byref(h))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_encoder_create')
        self._h = h
        self.ctl(OPUS_SET_COMPLEXITY_REQUEST, complexity)
        self.ctl(OPUS_SET_LSB_DEPTH_REQUEST, lsb_depth)
        if bitrate is not None:
            self.ctl(OPUS_SET_BITRATE_REQUEST, bitrate)
        self.ctl(OPUS_SET_VBR_REQUEST, 1 if vbr else 0)
        self.ctl(OPUS_SET_VBR_CONSTRAINT_REQUEST, 1 if constrained_vbr else 0)

    def ctl(self, request, value):
        r = self._L.opusb200_encoder_ctl_set(self._h, request, int(value))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_encoder_ctl_set(%d, %d)' % (request, value))

    def encode(self, pcm, max_data_bytes, frame_size=None):
        """pcm float32 [n, frame_size*channels] interleaved (+-1.0).  Returns (packets uint8 [n, max_data_bytes], nbytes int32 [n])."""
        pcm = np.ascontiguousarray(pcm, np.float32)
        assert pcm.shape[0] == self.n_streams
        fs = frame_size or pcm.shape[1] // self.channels
        pk = np.zeros((self.n_streams, max_data_bytes), np.uint8)
        nb = np.zeros(self.n_streams, np.int32)
        r = self._L.opusb200_encode_float(self._h, _p(pcm), fs, _p(pk), max_data_bytes, max_data_bytes, _p(nb))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_encode_float')
        return pk, nb

    def final_range(self):
        fr = np.empty(self.n_streams, np.uint32)
        r = self._L.opusb200_encoder_get_final_range(self._h, _p(fr))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_encoder_get_final_range')
        return fr

    def reset(self):
        r = self._L.opusb200_encoder_reset(self._h)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_encoder_reset')

    def sync(self):
        self._L.opusb200_encoder_sync(self._h)

    @property
    def cuda_stream(self):
        return self._L.opusb200_encoder_cuda_stream(self._h)

    @property
    def launch_count(self):
        return self._L.opusb200_encoder_launch_count(self._h)

    def close(self):
        if getattr(self, '_h', None):
            self._L.opusb200_encoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


OPUS_APPLICATION_RESTRICTED_LOWDELAY, OPUS_APPLICATION_RESTRICTED_CELT = 2051, 2053


class OpusEncoderBatch:
    """n independent Opus encoders (include/opus.h:170-355: opus_encoder_create / opus_encode / opus_encode_float / opus_encoder_ctl) on one GPU."""

    def __init__(self, n_streams, channels=2, fs=48000, application=OPUS_APPLICATION_RESTRICTED_CELT, device=0, bitrate=None, vbr=None, complexity=None):
        self._L = capi.lib()
        self.n_streams, self.channels = n_streams, channels
        h = ctypes.c_void_p()
        r = self._L.opusb200_opus_encoder_create(device, fs, channels, application, n_streams, ctypes.byref(h))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_opus_encoder_create')
        self._h = h
        if bitrate is not None:
            self.ctl(OPUS_SET_BITRATE_REQUEST, bitrate)
        if vbr is not None:
            self.ctl(OPUS_SET_VBR_REQUEST, 1 if vbr else 0)
        if complexity is not None:
            self.ctl(OPUS_SET_COMPLEXITY_REQUEST, complexity)

    def ctl(self, request, value):
        r = self._L.opusb200_opus_encoder_ctl(self._h, request, int(value))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_opus_encoder_ctl(%d, %d)' % (request, value))

    def encode(self, pcm, max_data_bytes=1276, frame_size=None):
        """pcm int16 or float32 [n, frame_size*channels] interleaved -> (packets uint8 [n, max_data_bytes], nbytes int32 [n]); the dtype selects
        opus_encode or opus_encode_float"""
        pcm = np.ascontiguousarray(pcm)
        assert pcm.shape[0] == self.n_streams and pcm.dtype in (np.int16, np.float32)
        fs = frame_size or pcm.shape[1] // self.channels
        pk = np.zeros((self.n_streams, max_data_bytes), np.uint8)
        nb = np.zeros(self.n_streams, np.int32)
        f = self._L.opusb200_opus_encode if pcm.dtype == np.int16 else self._L.opusb200_opus_encode_float
        r = f(self._h, _p(pcm), fs, _p(pk), max_data_bytes, max_data_bytes, _p(nb))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_opus_encode')
        return pk, nb

    def final_range(self):
        fr = np.empty(self.n_streams, np.uint32)
        r = self._L.opusb200_opus_encoder_get_final_range(self._h, _p(fr))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_opus_encoder_get_final_range')
        return fr

    def submit(self, pcm, packets, nbytes, max_data_bytes=None, frame_size=None):
        """opusb200_opus_encode_submit: enqueue one call on caller-owned host arrays (pcm int16 / float32 [n, frame_size*channels], packets uint8
        [n, stride], nbytes int32 [n]); they must stay alive until wait()"""
        assert pcm.flags.c_contiguous and packets.flags.c_contiguous and pcm.dtype in (np.int16, np.float32)
        fs = frame_size or pcm.shape[1] // self.channels
        stride = packets.shape[1]
        r = self._L.opusb200_opus_encode_submit(self._h, _p(pcm), 1 if pcm.dtype == np.int16 else 0, fs, _p(packets), stride,
                                                max_data_bytes or stride, _p(nbytes))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_opus_encode_submit')

    def wait(self):
        r = self._L.opusb200_opus_encode_wait(self._h)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_opus_encode_wait')

    def encode_device(self, d_pcm, pcm_is_int16, frame_size, d_packets, pkt_stride, max_data_bytes, d_nbytes):
        """device pointers (ints); enqueued on the encoder's stream"""
        r = self._L.opusb200_opus_encode_device(self._h, ctypes.c_void_p(d_pcm), int(pcm_is_int16), frame_size, ctypes.c_void_p(d_packets), pkt_stride,
                                                max_data_bytes, ctypes.c_void_p(d_nbytes))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_opus_encode_device')

    def set_timing(self, on):
        self._L.opusb200_opus_encoder_set_timing(self._h, 1 if on else 0)

    def get_timing(self):
        ms = (ctypes.c_double * 5)()
        calls = ctypes.c_int64(0)
        self._L.opusb200_opus_encoder_get_timing(self._h, ms, ctypes.byref(calls))
        return dict(zip(('opus_pre', 'opus_analysis', 'celt_chains', 'celt_analysis', 'celt_frame'), list(ms))), calls.value

    def sync(self):
        self._L.opusb200_opus_encoder_sync(self._h)

    @property
    def cuda_stream(self):
        return self._L.opusb200_opus_encoder_cuda_stream(self._h)

    @property
    def launch_count(self):
        return self._L.opusb200_opus_encoder_launch_count(self._h)

    def close(self):
        if getattr(self, '_h', None):
            self._L.opusb200_opus_encoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
