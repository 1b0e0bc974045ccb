"""ctypes binding of include/opus_b200.h (the same stub a maintainer of another host language
would write against the C ABI)."""
import ctypes
import os
import subprocess

OPUS_OK = 0
OPUS_BAD_ARG = -1
OPUS_BUFFER_TOO_SMALL = -2
OPUS_INTERNAL_ERROR = -3
OPUS_INVALID_PACKET = -4
OPUS_UNIMPLEMENTED = -5
OPUS_INVALID_STATE = -6
OPUS_ALLOC_FAIL = -7

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-fmad=false', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared', '-diag-suppress', '177']

# every symbol include/opus_b200.h declares
SYMBOLS = ['opusb200_version', 'opusb200_device_count', 'opusb200_decoder_create', 'opusb200_decoder_destroy',
           'opusb200_decoder_reset', 'opusb200_decoder_n_streams', 'opusb200_decode', 'opusb200_decode_float',
           'opusb200_decode_device', 'opusb200_decoder_sync', 'opusb200_decoder_set_pipelined', 'opusb200_decoder_join', 'opusb200_decoder_get_final_range', 'opusb200_decoder_ctl', 'opusb200_decoder_get_ctl',
           'opusb200_decoder_cuda_stream', 'opusb200_decoder_launch_count', 'opusb200_decoder_set_timing',
           'opusb200_decoder_get_timing', 'opusb200_decode_submit', 'opusb200_decode_wait', 'opusb200_celt_analysis_create', 'opusb200_encoder_create', 'opusb200_encoder_destroy',
           'opusb200_encoder_reset', 'opusb200_encoder_ctl_set', 'opusb200_encode_float', 'opusb200_encode_device',
           'opusb200_encoder_get_final_range', 'opusb200_encoder_sync', 'opusb200_encoder_cuda_stream', 'opusb200_encoder_launch_count',
           'opusb200_celt_analysis_destroy', 'opusb200_celt_analysis_reset', 'opusb200_celt_analysis_process', 'opusb200_ctx_create', 'opusb200_ctx_destroy', 'opusb200_ctx_sync',
           'opusb200_ctx_cuda_stream', 'opusb200_mdct_forward', 'opusb200_pitch_xcorr', 'opusb200_exp_rotation',
           'opusb200_pvq_search', 'opusb200_silk_lpc_analysis_filter', 'opusb200_silk_resampler_create', 'opusb200_silk_resampler_create_ex', 'opusb200_range_encoder_create', 'opusb200_range_encoder_destroy',
           'opusb200_range_encoder_init', 'opusb200_range_encoder_icdf', 'opusb200_silk_encode_pulses', 'opusb200_range_encoder_done', 'opusb200_silk_encode_indices', 'opusb200_silk_coder_reset_state',
           'opusb200_silk_resampler_destroy', 'opusb200_silk_resampler_reset', 'opusb200_silk_resampler_process',
           'opusb200_silk_burg_modified', 'opusb200_silk_autocorrelation',
           'opusb200_silk_nsq_create', 'opusb200_silk_nsq_destroy', 'opusb200_silk_nsq_reset', 'opusb200_silk_nsq_process',
           'opusb200_silk_nsq_get_state', 'opusb200_nnet_linear_create', 'opusb200_nnet_linear_destroy',
           'opusb200_nnet_compute_linear', 'opusb200_nnet_compute_activation', 'opusb200_nnet_compute_dense',
           'opusb200_nnet_compute_gru', 'opusb200_nnet_compute_glu', 'opusb200_nnet_compute_conv1d',
           'opusb200_nnet_conv2d_create', 'opusb200_nnet_conv2d_destroy', 'opusb200_nnet_compute_conv2d',
           'opusb200_fargan_create', 'opusb200_fargan_destroy', 'opusb200_fargan_cont', 'opusb200_fargan_synthesize',
           'opusb200_dred_encoder_create', 'opusb200_dred_encoder_destroy', 'opusb200_dred_encoder_reset', 'opusb200_dred_encode_dframe',
           'opusb200_dred_encoder_init_frontend', 'opusb200_dred_compute_latents', 'opusb200_dred_encoder_get_latents',
           'opusb200_dred_encoder_load_stats', 'opusb200_dred_encode_silk_frame',
           'opusb200_lpcnet_plc_create', 'opusb200_lpcnet_plc_destroy', 'opusb200_lpcnet_plc_reset', 'opusb200_lpcnet_plc_fec_add', 'opusb200_lpcnet_plc_fec_clear',
           'opusb200_lpcnet_plc_update', 'opusb200_lpcnet_plc_conceal', 'opusb200_lpcnet_plc_get_state', 'opusb200_lpcnet_plc_load_pitch_model',
           'opusb200_silk_process_gains', 'opusb200_silk_process_nlsfs', 'opusb200_silk_find_lpc', 'opusb200_silk_find_pred_coefs', 'opusb200_silk_noise_shape_analysis', 'opusb200_silk_vad_init', 'opusb200_silk_encode_do_vad', 'opusb200_silk_hp_variable_cutoff', 'opusb200_silk_find_pitch_lags', 'opusb200_lpcnet_encoder_create', 'opusb200_lpcnet_encoder_destroy', 'opusb200_lpcnet_encoder_reset',
           'opusb200_lpcnet_compute_single_frame_features', 'opusb200_lpcnet_encoder_get_state', 'opusb200_burg_cepstral_analysis',
           'opusb200_dred_decoder_create', 'opusb200_dred_decoder_destroy', 'opusb200_dred_decode_all', 'opusb200_host_alloc', 'opusb200_host_free',
           'opusb200_opus_encoder_create', 'opusb200_opus_encoder_destroy', 'opusb200_opus_encoder_reset', 'opusb200_opus_encoder_ctl',
           'opusb200_opus_encode', 'opusb200_opus_encode_float', 'opusb200_opus_encode_device', 'opusb200_opus_encode_submit', 'opusb200_opus_encode_wait', 'opusb200_opus_encoder_get_final_range',
           'opusb200_opus_encoder_sync', 'opusb200_opus_encoder_cuda_stream', 'opusb200_opus_encoder_launch_count',
           'opusb200_opus_encoder_set_timing', 'opusb200_opus_encoder_get_timing']


class OpusB200Error(RuntimeError):
    def __init__(self, code, what=''):
        super().__init__('opus_b200 error %d %s' % (code, what))
        self.code = code


def library_path():
    return os.path.join(_HERE, 'libopus_b200.so')


def build_library(force=False):
    """Compile opus_b200/csrc/opus_b200.cu for sm_100a into opus_b200/libopus_b200.so (in tree)."""
    out = library_path()
    src_dir = os.path.join(_HERE, 'csrc')
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir)] + [os.path.join(ROOT, 'include', 'opus_b200.h')]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    subprocess.check_call([nvcc] + NVCC_FLAGS + ['-o', out, os.path.join(src_dir, 'opus_b200.cu')])
    build_compat_library(force=True)
    return out


# the libopus-named entry points (include/opus_b200_compat.h) over the C ABI: plain C++ (no CUDA), linked against libopus_b200.so next to it
COMPAT_SYMBOLS = ['opus_strerror', 'opus_get_version_string', 'opus_decoder_get_size', 'opus_decoder_create', 'opus_decoder_init', 'opus_decode',
                  'opus_decode_float', 'opus_decoder_ctl', 'opus_decoder_destroy', 'opus_encoder_get_size', 'opus_encoder_create', 'opus_encoder_init',
                  'opus_encode', 'opus_encode_float', 'opus_encoder_ctl', 'opus_encoder_destroy', 'opus_packet_parse', 'opus_packet_get_bandwidth',
                  'opus_packet_get_samples_per_frame', 'opus_packet_get_nb_channels', 'opus_packet_get_nb_frames', 'opus_packet_get_nb_samples',
                  'opus_decoder_get_nb_samples', 'opus_multistream_decoder_get_size', 'opus_multistream_decoder_create', 'opus_multistream_decoder_init',
                  'opus_multistream_decode', 'opus_multistream_decode_float', 'opus_multistream_decoder_ctl', 'opus_multistream_decoder_destroy']


def compat_library_path():
    return os.path.join(_HERE, 'libopus_b200_compat.so')


def build_compat_library(force=False):
    out = compat_library_path()
    src = os.path.join(_HERE, 'csrc', 'opus_compat.cpp')
    deps = [src, os.path.join(ROOT, 'include', 'opus_b200_compat.h'), os.path.join(ROOT, 'include', 'opus_b200.h')]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-fvisibility=default', '-Wall', '-o', out, src, '-L' + _HERE, '-lopus_b200',
                           '-Wl,-rpath,$ORIGIN'])
    return out


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        p = library_path()
        if not os.path.exists(p):
            raise OpusB200Error(OPUS_INTERNAL_ERROR, 'libopus_b200.so is not built (run __graft_entry__.build()); '
                                'there is no CPU fallback')
        L = ctypes.CDLL(p)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        L.opusb200_version.restype = ctypes.c_char_p
        L.opusb200_decoder_create.argtypes = [ctypes.c_int, i32, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
        L.opusb200_decoder_destroy.argtypes = [vp]
        L.opusb200_decoder_destroy.restype = None
        L.opusb200_decoder_reset.argtypes = [vp]
        L.opusb200_decoder_n_streams.argtypes = [vp]
        for f in (L.opusb200_decode, L.opusb200_decode_float):
            f.argtypes = [vp, vp, ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_int, vp]
        L.opusb200_decode_device.argtypes = [vp, vp, ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, vp]
        L.opusb200_decoder_sync.argtypes = [vp]
        L.opusb200_decoder_get_final_range.argtypes = [vp, vp]
        L.opusb200_decoder_ctl.argtypes = [vp, ctypes.c_int, ctypes.c_int32]
        L.opusb200_decoder_get_ctl.argtypes = [vp, ctypes.c_int, vp]
        L.opusb200_decoder_cuda_stream.argtypes = [vp]
        L.opusb200_decoder_cuda_stream.restype = vp
        L.opusb200_decoder_launch_count.argtypes = [vp]
        L.opusb200_decoder_launch_count.restype = i64
        L.opusb200_decoder_set_timing.argtypes = [vp, ctypes.c_int]
        L.opusb200_decoder_get_timing.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64)]
        L.opusb200_ctx_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
        L.opusb200_ctx_destroy.argtypes = [vp]
        L.opusb200_ctx_destroy.restype = None
        L.opusb200_ctx_sync.argtypes = [vp]
        L.opusb200_ctx_cuda_stream.argtypes = [vp]
        L.opusb200_ctx_cuda_stream.restype = vp
        ci = ctypes.c_int
        L.opusb200_mdct_forward.argtypes = [vp, vp, vp, ci, ci, ci, ci]
        L.opusb200_pitch_xcorr.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci]
        L.opusb200_exp_rotation.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci]
        L.opusb200_pvq_search.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci]
        L.opusb200_silk_lpc_analysis_filter.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci]
        L.opusb200_silk_resampler_create.argtypes = [vp, ci, i32, i32, ctypes.POINTER(vp)]
        L.opusb200_silk_resampler_create_ex.argtypes = [vp, ci, i32, i32, ci, ctypes.POINTER(vp)]
        L.opusb200_range_encoder_create.argtypes = [vp, ci, ci, ctypes.POINTER(vp)]
        L.opusb200_range_encoder_destroy.argtypes = [vp]
        L.opusb200_range_encoder_destroy.restype = None
        L.opusb200_range_encoder_init.argtypes = [vp, ci]
        L.opusb200_range_encoder_icdf.argtypes = [vp, vp, vp, ci, ci, ci]
        L.opusb200_silk_encode_pulses.argtypes = [vp, vp, vp, vp, ci, ci, ci]
        L.opusb200_range_encoder_done.argtypes = [vp, vp, vp, vp, vp]
        L.opusb200_silk_encode_indices.argtypes = [vp, ci, ci, vp, vp, ci]
        L.opusb200_silk_coder_reset_state.argtypes = [vp]
        L.opusb200_silk_resampler_destroy.argtypes = [vp]
        L.opusb200_silk_resampler_destroy.restype = None
        L.opusb200_silk_resampler_reset.argtypes = [vp]
        L.opusb200_silk_resampler_process.argtypes = [vp, vp, vp, ci, ci]
        L.opusb200_silk_burg_modified.argtypes = [vp, vp, vp, vp, ctypes.c_float, ci, ci, ci, ci, ci]
        L.opusb200_silk_autocorrelation.argtypes = [vp, vp, vp, ci, ci, ci, ci]
        L.opusb200_silk_nsq_create.argtypes = [vp, ci, vp, ctypes.POINTER(vp)]
        L.opusb200_silk_nsq_destroy.argtypes = [vp]
        L.opusb200_silk_nsq_destroy.restype = None
        L.opusb200_silk_nsq_reset.argtypes = [vp]
        L.opusb200_silk_nsq_process.argtypes = [vp, vp, vp, vp, ci]
        L.opusb200_silk_nsq_get_state.argtypes = [vp, vp]
        L.opusb200_nnet_linear_create.argtypes = [vp, vp, ci, ctypes.POINTER(vp)]
        L.opusb200_nnet_linear_destroy.argtypes = [vp]
        L.opusb200_nnet_linear_destroy.restype = None
        L.opusb200_nnet_compute_linear.argtypes = [vp, vp, vp, vp, ci, ci]
        L.opusb200_nnet_compute_activation.argtypes = [vp, vp, vp, ci, ci, ci, ci]
        L.opusb200_nnet_compute_dense.argtypes = [vp, vp, vp, vp, ci, ci, ci]
        L.opusb200_nnet_compute_gru.argtypes = [vp, vp, vp, vp, vp, ci, ci]
        L.opusb200_nnet_compute_glu.argtypes = [vp, vp, vp, vp, ci, ci]
        L.opusb200_nnet_compute_conv1d.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci]
        L.opusb200_nnet_conv2d_create.argtypes = [vp, vp, ctypes.POINTER(vp)]
        L.opusb200_nnet_conv2d_destroy.argtypes = [vp]
        L.opusb200_nnet_conv2d_destroy.restype = None
        L.opusb200_nnet_compute_conv2d.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci]
        L.opusb200_fargan_create.argtypes = [vp, vp, ci, ci, ci, ctypes.POINTER(vp)]
        L.opusb200_dred_decoder_create.argtypes = [vp, vp, ci, ci, ci, ctypes.POINTER(vp)]
        L.opusb200_dred_decoder_destroy.argtypes = [vp]
        L.opusb200_dred_decoder_destroy.restype = None
        L.opusb200_dred_decode_all.argtypes = [vp, vp, vp, vp, ci, ci]
        L.opusb200_fargan_destroy.argtypes = [vp]
        L.opusb200_fargan_destroy.restype = None
        L.opusb200_fargan_cont.argtypes = [vp, vp, vp, ci]
        L.opusb200_fargan_synthesize.argtypes = [vp, vp, vp, ci]
        L.opusb200_dred_encoder_create.argtypes = [vp, vp, ci, ci, ci, ci, ci, ctypes.POINTER(vp)]
        L.opusb200_dred_encoder_destroy.argtypes = [vp]
        L.opusb200_dred_encoder_destroy.restype = None
        L.opusb200_dred_encoder_reset.argtypes = [vp]
        L.opusb200_dred_encode_dframe.argtypes = [vp, vp, vp, vp, ci]
        L.opusb200_dred_encoder_init_frontend.argtypes = [vp, ci, ci, vp, ci, vp, ci, ci]
        L.opusb200_dred_compute_latents.argtypes = [vp, vp, ci, ci, ci]
        L.opusb200_dred_encoder_get_latents.argtypes = [vp, vp, vp, vp]
        L.opusb200_dred_encoder_load_stats.argtypes = [vp, vp, ci]
        L.opusb200_dred_encode_silk_frame.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, vp, vp, ci]
        L.opusb200_lpcnet_plc_create.argtypes = [vp, vp, ci, vp, ci, ci, ci, ctypes.POINTER(vp)]
        L.opusb200_lpcnet_plc_destroy.argtypes = [vp]
        L.opusb200_lpcnet_plc_destroy.restype = None
        L.opusb200_lpcnet_plc_reset.argtypes = [vp]
        L.opusb200_lpcnet_plc_fec_add.argtypes = [vp, vp, vp, ci]
        L.opusb200_lpcnet_plc_fec_clear.argtypes = [vp]
        L.opusb200_lpcnet_plc_update.argtypes = [vp, vp, ci]
        L.opusb200_lpcnet_plc_conceal.argtypes = [vp, vp, ci]
        L.opusb200_lpcnet_plc_get_state.argtypes = [vp, vp, vp, vp, vp, vp]
        L.opusb200_lpcnet_plc_load_pitch_model.argtypes = [vp, vp, ci, vp, ci, ci]
        L.opusb200_silk_process_gains.argtypes = [vp, vp, ci, ci, ci, ci, ci]
        L.opusb200_silk_process_nlsfs.argtypes = [vp, vp, ci, ci, ci, ci]
        L.opusb200_silk_find_lpc.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci]
        L.opusb200_silk_find_pred_coefs.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, ci]
        L.opusb200_silk_find_pitch_lags.argtypes = [vp, vp, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, ci]
        L.opusb200_silk_vad_init.argtypes = [vp, ci]
        L.opusb200_silk_encode_do_vad.argtypes = [vp, vp, vp, ci, vp, ci, ci, ci, ci]
        L.opusb200_silk_hp_variable_cutoff.argtypes = [vp, vp, ci, ci, ci]
        L.opusb200_silk_noise_shape_analysis.argtypes = [vp, vp, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci]
        L.opusb200_lpcnet_encoder_create.argtypes = [vp, vp, ci, vp, ci, ci, ci, ctypes.POINTER(vp)]
        L.opusb200_lpcnet_encoder_destroy.argtypes = [vp]
        L.opusb200_lpcnet_encoder_destroy.restype = None
        L.opusb200_lpcnet_encoder_reset.argtypes = [vp]
        L.opusb200_lpcnet_compute_single_frame_features.argtypes = [vp, vp, ci, vp, ci]
        L.opusb200_lpcnet_encoder_get_state.argtypes = [vp, vp, vp, vp, vp]
        L.opusb200_burg_cepstral_analysis.argtypes = [vp, vp, vp, ci, ci]
        L.opusb200_decoder_set_pipelined.argtypes = [vp, ci]
        L.opusb200_decoder_join.argtypes = [vp]
        L.opusb200_decode_submit.argtypes = [vp, vp, ci, vp, vp, ci, ci, ci, vp]
        L.opusb200_decode_wait.argtypes = [vp]
        L.opusb200_celt_analysis_create.argtypes = [vp, ci, ci, ctypes.POINTER(vp)]
        L.opusb200_celt_analysis_destroy.argtypes = [vp]
        L.opusb200_celt_analysis_destroy.restype = None
        L.opusb200_celt_analysis_reset.argtypes = [vp]
        L.opusb200_celt_analysis_process.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci]
        L.opusb200_encoder_create.argtypes = [ci, i32, ci, ci, ctypes.POINTER(vp)]
        L.opusb200_encoder_destroy.argtypes = [vp]
        L.opusb200_encoder_destroy.restype = None
        L.opusb200_encoder_reset.argtypes = [vp]
        L.opusb200_encoder_ctl_set.argtypes = [vp, ci, i32]
        L.opusb200_encode_float.argtypes = [vp, vp, ci, vp, ci, ci, vp]
        L.opusb200_encode_device.argtypes = [vp, vp, ci, vp, ci, ci, vp]
        L.opusb200_encoder_get_final_range.argtypes = [vp, vp]
        L.opusb200_encoder_sync.argtypes = [vp]
        L.opusb200_encoder_cuda_stream.argtypes = [vp]
        L.opusb200_encoder_cuda_stream.restype = vp
        L.opusb200_encoder_launch_count.argtypes = [vp]
        L.opusb200_encoder_launch_count.restype = ctypes.c_int64
        L.opusb200_opus_encoder_create.argtypes = [ci, i32, ci, ci, ci, ctypes.POINTER(vp)]
        L.opusb200_opus_encoder_destroy.argtypes = [vp]
        L.opusb200_opus_encoder_destroy.restype = None
        L.opusb200_opus_encoder_reset.argtypes = [vp]
        L.opusb200_opus_encoder_ctl.argtypes = [vp, ci, i32]
        L.opusb200_opus_encode.argtypes = [vp, vp, ci, vp, ci, ci, vp]
        L.opusb200_opus_encode_float.argtypes = [vp, vp, ci, vp, ci, ci, vp]
        L.opusb200_opus_encode_device.argtypes = [vp, vp, ci, ci, vp, ci, ci, vp]
        L.opusb200_opus_encode_submit.argtypes = [vp, vp, ci, ci, vp, ci, ci, vp]
        L.opusb200_opus_encode_wait.argtypes = [vp]
        L.opusb200_opus_encoder_get_final_range.argtypes = [vp, vp]
        L.opusb200_opus_encoder_sync.argtypes = [vp]
        L.opusb200_opus_encoder_cuda_stream.argtypes = [vp]
        L.opusb200_opus_encoder_cuda_stream.restype = vp
        L.opusb200_opus_encoder_launch_count.argtypes = [vp]
        L.opusb200_opus_encoder_launch_count.restype = ctypes.c_int64
        L.opusb200_opus_encoder_set_timing.argtypes = [vp, ci]
        L.opusb200_opus_encoder_get_timing.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64)]
        L.opusb200_host_alloc.argtypes = [ctypes.c_size_t]
        L.opusb200_host_alloc.restype = vp
        L.opusb200_host_free.argtypes = [vp]
        L.opusb200_host_free.restype = None
        _lib = L
    return _lib
