"""DecoderBatch: N independent Opus decoders on one GPU (host-side mirror of the reference's
OpusDecoder API, include/opus.h:458-591, one packet per stream per call)."""
import ctypes

import numpy as np

from . import capi

MAX_FRAME = 5760  # 120 ms at 48 kHz, the largest an Opus packet can carry


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class DecoderBatch:
    """`n_streams` decoders created like opus_decoder_create(Fs, channels) (src/opus_decoder.c:150)."""

    def __init__(self, n_streams, channels=2, Fs=48000, device=0):
        self._L = capi.lib()
        self.n_streams, self.channels, self.Fs, self.device = int(n_streams), int(channels), int(Fs), int(device)
        h = ctypes.c_void_p()
        r = self._L.opusb200_decoder_create(self.device, self.Fs, self.channels, self.n_streams, ctypes.byref(h))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_decoder_create')
        self._h = h

    # -- opus_decode() / opus_decode_float(), batched, host buffers ----------------------------
    def decode(self, packets, lens, frame_size=960, dtype=np.int16, out=None):
        """packets: uint8 [n_streams, stride]; lens: int32 [n_streams] (0 = packet lost: frame_size samples are concealed, like
        opus_decode(data=NULL); < 0 = the stream sits this call out).
        Returns (pcm [n_streams, frame_size*channels], samples_per_channel [n_streams])."""
        packets = np.ascontiguousarray(packets, np.uint8)
        lens = np.ascontiguousarray(lens, np.int32)
        assert packets.shape[0] == self.n_streams and lens.shape[0] == self.n_streams
        stride = frame_size * self.channels
        if out is None:
            out = np.empty((self.n_streams, stride), dtype)
        ns = np.empty(self.n_streams, np.int32)
        fn = self._L.opusb200_decode if np.dtype(dtype) == np.int16 else self._L.opusb200_decode_float
        r = fn(self._h, _ptr(packets), packets.shape[1], _ptr(lens), _ptr(out), stride, frame_size, _ptr(ns))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_decode')
        return out, ns

    # -- asynchronous host-buffer variant: submit k+1 before wait()ing for k to overlap read-back with compute ---------
    def submit(self, packets, lens, out, ns, frame_size=960):
        """packets uint8 [n, stride], lens int32 [n], out int16|float32 [n, frame_size*channels], ns int32 [n]: all C-contiguous
        and kept alive (ideally pinned) until wait() returns."""
        assert packets.flags.c_contiguous and out.flags.c_contiguous and packets.dtype == np.uint8 and lens.dtype == np.int32
        r = self._L.opusb200_decode_submit(self._h, _ptr(packets), packets.shape[1], _ptr(lens), _ptr(out), out.shape[1],
                                           1 if out.dtype == np.int16 else 0, frame_size, _ptr(ns))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_decode_submit')

    def wait(self):
        r = self._L.opusb200_decode_wait(self._h)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_decode_wait')

    # -- device-resident variant (torch CUDA tensors or raw pointers) --------------------------
    def decode_device(self, d_packets, pkt_stride, d_lens, d_pcm, pcm_stride, int16=True, frame_size=960,
                      max_frames_per_packet=1, d_out_samples=None):
        r = self._L.opusb200_decode_device(self._h, d_packets, pkt_stride, d_lens, d_pcm, pcm_stride, 1 if int16 else 0,
                                           frame_size, max_frames_per_packet, d_out_samples)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_decode_device')

    def sync(self):
        r = self._L.opusb200_decoder_sync(self._h)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_decoder_sync')

    def set_pipelined(self, enable):
        """device API: let a call's synthesis overlap the next call's parsing; join()/sync() before reading the PCM"""
        r = self._L.opusb200_decoder_set_pipelined(self._h, 1 if enable else 0)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_decoder_set_pipelined')

    def join(self):
        r = self._L.opusb200_decoder_join(self._h)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_decoder_join')

    def reset(self):  # OPUS_RESET_STATE
        r = self._L.opusb200_decoder_reset(self._h)
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_decoder_reset')

    def final_range(self):  # OPUS_GET_FINAL_RANGE
        fr = np.empty(self.n_streams, np.uint32)
        r = self._L.opusb200_decoder_get_final_range(self._h, _ptr(fr))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_decoder_get_final_range')
        return fr

    # -- opus_decoder_ctl(): integer requests of the decode path (include/opus_defines.h numbers) ----------------------------
    OPUS_GET_SAMPLE_RATE, OPUS_GET_PITCH, OPUS_SET_GAIN, OPUS_GET_LAST_PACKET_DURATION, OPUS_GET_GAIN = 4029, 4033, 4034, 4039, 4045
    OPUS_SET_PHASE_INVERSION_DISABLED, OPUS_GET_PHASE_INVERSION_DISABLED = 4046, 4047

    def ctl(self, request, value):
        """a setter, applied to every stream (OPUS_SET_GAIN, OPUS_SET_PHASE_INVERSION_DISABLED)"""
        r = self._L.opusb200_decoder_ctl(self._h, int(request), int(value))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_decoder_ctl(%d)' % request)

    def get_ctl(self, request):
        """a getter: int32 [n_streams] (OPUS_GET_PITCH, OPUS_GET_LAST_PACKET_DURATION, OPUS_GET_GAIN, ...)"""
        v = np.empty(self.n_streams, np.int32)
        r = self._L.opusb200_decoder_get_ctl(self._h, int(request), _ptr(v))
        if r != 0:
            raise capi.OpusB200Error(r, 'opusb200_decoder_get_ctl(%d)' % request)
        return v

    def set_timing(self, enable):
        self._L.opusb200_decoder_set_timing(self._h, 1 if enable else 0)

    def get_timing(self):
        """({'entropy': ms, 'pvq': ms, 'synth': ms}, frame steps) accumulated since set_timing(True)."""
        ms = (ctypes.c_double * 3)()
        n = ctypes.c_int64(0)
        self._L.opusb200_decoder_get_timing(self._h, ms, ctypes.byref(n))
        return {'entropy': ms[0], 'pvq': ms[1], 'synth': ms[2]}, n.value

    @property
    def cuda_stream(self):
        return self._L.opusb200_decoder_cuda_stream(self._h)

    @property
    def launch_count(self):
        return int(self._L.opusb200_decoder_launch_count(self._h))

    def close(self):
        if getattr(self, '_h', None):
            self._L.opusb200_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
