/* opus_b200.h — C ABI of the B200-native batched Opus (CELT) hot path.
 *
 * Plain C, no CUDA/torch types in the signatures.  A "batch decoder" is N independent Opus
 * decoders (N streams) living on one GPU; one call decodes one packet per stream.  It is the
 * batched twin of the reference's per-stream API and keeps its conventions:
 *
 *   reference (xiph/opus include/opus.h)                     this library
 *   ------------------------------------------------------  -----------------------------------
 *   opus_decoder_create(Fs, channels, &err)        :458      opusb200_decoder_create(..., n_streams, ...)
 *   opus_decode(st, data, len, pcm, frame_size, fec) :516    opusb200_decode(dec, packets, stride, len[], pcm, ...)
 *   opus_decode_float(...)                         :566      opusb200_decode_float(...)
 *   opus_decoder_ctl(st, OPUS_RESET_STATE)         :586      opusb200_decoder_reset(dec)
 *   opus_decoder_ctl(st, OPUS_GET_FINAL_RANGE(&r)) :586      opusb200_decoder_get_final_range(dec, r[])
 *   opus_decoder_ctl(st, OPUS_SET_GAIN(g)) etc.    :586      opusb200_decoder_ctl(dec, request, value) / _get_ctl(dec, request, v[])
 *   opus_decoder_destroy(st)                       :591      opusb200_decoder_destroy(dec)
 *
 * Below the API the frame path is the reference's celt_decode_with_ec() (celt/celt.h:172,
 * celt/celt_decoder.c:1104) restated as two sm_100a kernels (opus_b200/csrc/).
 * Return values: >= 0 success / sample counts, < 0 the reference's error codes
 * (include/opus_defines.h:46-60): OPUS_BAD_ARG -1, OPUS_BUFFER_TOO_SMALL -2, OPUS_INTERNAL_ERROR -3,
 * OPUS_INVALID_PACKET -4, OPUS_UNIMPLEMENTED -5, OPUS_INVALID_STATE -6, OPUS_ALLOC_FAIL -7.
 *
 * Scope of this round: CELT-only packets (TOC config 16..31), output rates 48/24/16/12/8 kHz, 1 or 2 channels, frame
 * sizes 2.5/5/10/20 ms, all frame-packing codes, packet-loss concealment (a lost packet is len[s] == 0, like
 * opus_decode(data = NULL); DTX frames inside a packet are concealed too: celt_decode_lost, celt/celt_decoder.c:679-1090,
 * pitch-based and noise PLC; no DEEP_PLC).  int16 output is soft-clipped like opus_decode() does (opus_pcm_soft_clip, src/opus.c:39-166,
 * once over all frames of a packet, src/opus_decoder.c:875-878), float output is not, like opus_decode_float().
 * SILK/hybrid packets and FEC return OPUS_UNIMPLEMENTED per stream.
 * There is NO CPU fallback: every entry point fails with OPUS_INTERNAL_ERROR if no CUDA device
 * is usable. */
#ifndef OPUS_B200_H
#define OPUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPUSB200_OK 0
#define OPUSB200_BAD_ARG (-1)
#define OPUSB200_BUFFER_TOO_SMALL (-2)
#define OPUSB200_INTERNAL_ERROR (-3)
#define OPUSB200_INVALID_PACKET (-4)
#define OPUSB200_UNIMPLEMENTED (-5)
#define OPUSB200_INVALID_STATE (-6)
#define OPUSB200_ALLOC_FAIL (-7)

typedef struct OpusB200Decoder OpusB200Decoder;

/* Library / device info. */
const char *opusb200_version(void);
int opusb200_device_count(void);

/* N decoders in the state opus_decoder_create() leaves them (src/opus_decoder.c:150-188,
 * celt/celt_decoder.c:240-268).  device: CUDA ordinal.  Fs: 48000, 24000, 16000, 12000 or 8000 (the output rate: below
 * 48 kHz the decoder keeps only the spectrum under Fs/2 and decimates after the de-emphasis, celt/celt_decoder.c:219-231,
 * 300-405; frame_size / pcm_stride / out_samples of the decode calls are then in samples at Fs); channels 1 or 2. */
int opusb200_decoder_create(int device, int32_t Fs, int channels, int n_streams, OpusB200Decoder **out);
void opusb200_decoder_destroy(OpusB200Decoder *dec);
int opusb200_decoder_reset(OpusB200Decoder *dec); /* OPUS_RESET_STATE on every stream */
int opusb200_decoder_n_streams(const OpusB200Decoder *dec);

/* Decode one packet per stream; HOST buffers (copies to/from the device are part of the call).
 *   packets    : stream s's packet starts at packets + s*pkt_stride
 *   len[s]     : packet length in bytes.  0 = the packet is LOST: frame_size samples are concealed (frame_size must then be a
 *                multiple of 2.5 ms, src/opus_decoder.c:757-778), zeros if the stream has not decoded anything yet.
 *                < 0 (this library's extension) = the stream sits this call out: 0 samples, state untouched.
 *   pcm        : stream s's output starts at pcm + s*pcm_stride (interleaved if 2 channels)
 *   pcm_stride : elements (not bytes) between consecutive streams' outputs, >= frame_size*channels
 *   frame_size : capacity per channel of each stream's pcm area, like opus_decode()'s argument
 *   out_samples[s] : samples per channel decoded for stream s, or a negative error code
 * Returns 0, or a negative error code if the call as a whole failed. */
int opusb200_decode(OpusB200Decoder *dec, const unsigned char *packets, int pkt_stride, const int32_t *len,
                    int16_t *pcm, int pcm_stride, int frame_size, int32_t *out_samples);
int opusb200_decode_float(OpusB200Decoder *dec, const unsigned char *packets, int pkt_stride, const int32_t *len,
                          float *pcm, int pcm_stride, int frame_size, int32_t *out_samples);
/* Asynchronous form of the two calls above for batch servers: enqueue call k+1 before waiting for call k and the PCM
 * read-back of k overlaps the kernels of k+1 (device staging is double-buffered; results are identical).  The host buffers
 * of a submitted call must stay valid and untouched until opusb200_decode_wait returns; at most two calls may be in flight. */
int opusb200_decode_submit(OpusB200Decoder *dec, const unsigned char *packets, int pkt_stride, const int32_t *len, void *pcm,
                           int pcm_stride, int pcm_is_int16, int frame_size, int32_t *out_samples);
int opusb200_decode_wait(OpusB200Decoder *dec);

/* Same, with every buffer already resident in device memory (no copies, no synchronisation:
 * work is enqueued on the decoder's stream; call opusb200_decoder_sync() before reading results).
 * max_frames_per_packet: upper bound on frames per packet the caller guarantees (1 for the
 * usual one-frame packets, up to 48); a lost packet (d_len[s] == 0) needs frame_size / (the stream's packet frame size) of them,
 * i.e. 1 when frame_size is the stream's frame size; pcm_is_int16 selects int16 or float output. */
int opusb200_decode_device(OpusB200Decoder *dec, const unsigned char *d_packets, int pkt_stride, const int32_t *d_len,
                           void *d_pcm, int pcm_stride, int pcm_is_int16, int frame_size, int max_frames_per_packet,
                           int32_t *d_out_samples);
int opusb200_decoder_sync(OpusB200Decoder *dec);
/* The synthesis kernel of a frame runs on an internal stream so that it can overlap the next frame's bitstream parsing.
 * By default every opusb200_decode_device call ends by making the library stream wait for it (stream-ordered results).
 * set_pipelined(dec, 1) leaves that wait out: consecutive calls then overlap (frame t's synthesis under frame t+1's range
 * decoding); call opusb200_decoder_join (enqueue the wait) or opusb200_decoder_sync before consuming the PCM.
 * No reference counterpart: the reference decodes one stream synchronously (src/opus_decoder.c:opus_decode_native). */
int opusb200_decoder_set_pipelined(OpusB200Decoder *dec, int enable);
int opusb200_decoder_join(OpusB200Decoder *dec);

/* OPUS_GET_FINAL_RANGE for every stream (host array of n_streams). */
int opusb200_decoder_get_final_range(OpusB200Decoder *dec, uint32_t *final_range);
/* opus_decoder_ctl() (include/opus.h:586, src/opus_decoder.c:1057-1215) for the integer requests that touch the decode path; the
 * request numbers are include/opus_defines.h's.  Setters apply to every stream: 4034 OPUS_SET_GAIN (Q8 dB, -32768..32767),
 * 4046 OPUS_SET_PHASE_INVERSION_DISABLED (0/1).  Getters fill one value per stream: 4033 OPUS_GET_PITCH, 4039
 * OPUS_GET_LAST_PACKET_DURATION, 4045 OPUS_GET_GAIN, 4047 OPUS_GET_PHASE_INVERSION_DISABLED, 4029 OPUS_GET_SAMPLE_RATE.
 * Anything else: OPUSB200_UNIMPLEMENTED. */
int opusb200_decoder_ctl(OpusB200Decoder *dec, int request, int32_t value);
int opusb200_decoder_get_ctl(OpusB200Decoder *dec, int request, int32_t *values);

/* The CUDA stream the decoder enqueues on, as an opaque pointer (cudaStream_t). */
void *opusb200_decoder_cuda_stream(OpusB200Decoder *dec);
/* Number of kernels launched by this decoder since creation (for benchmarks' bookkeeping). */
int64_t opusb200_decoder_launch_count(const OpusB200Decoder *dec);

/* Profiling aid: when enabled, every kernel launch is bracketed by CUDA events on the decoder's
 * stream and the stream is drained after each frame (so do not enable it for throughput runs).
 * get_timing returns the accumulated device milliseconds of the three kernels (ms_kernels[3]:
 * entropy, pvq, synthesis) and the number of frame steps they cover. */
int opusb200_decoder_set_timing(OpusB200Decoder *dec, int enable);
int opusb200_decoder_get_timing(OpusB200Decoder *dec, double *ms_kernels, int64_t *frames);

/* ---- Kernel-level operations, batched (the reference's per-call SIMD plug-in surface with a leading
 * batch dimension).  Results are bit-identical to the reference's C functions.
 * Every buffer is a HOST pointer when on_device == 0 (copied in/out, call returns when done) or a
 * DEVICE pointer when on_device != 0 (work is enqueued on the context's stream; opusb200_ctx_sync). */
typedef struct OpusB200Ctx OpusB200Ctx;
int opusb200_ctx_create(int device, OpusB200Ctx **out);
void opusb200_ctx_destroy(OpusB200Ctx *ctx);
int opusb200_ctx_sync(OpusB200Ctx *ctx);
void *opusb200_ctx_cuda_stream(OpusB200Ctx *ctx);

/* Forward MDCT of n_items channel-frames, replaces clt_mdct_forward_c as driven by compute_mdcts
 * (celt/mdct.c:122, celt/celt_encoder.c:511-554; NE10 plug-in slot celt/mdct.h:80-98).
 * in: [n_items][(120<<LM)+120] time samples; out: [n_items][120<<LM] coefficients (short blocks
 * interleaved when transient != 0). */
int opusb200_mdct_forward(OpusB200Ctx *ctx, const float *in, float *out, int n_items, int LM, int transient, int on_device);

/* celt_pitch_xcorr (celt/pitch.c:230, RTCD slot celt/x86/x86_celt_map.c:95):
 * xcorr[i] = sum_j x[j]*y[i+j], x: [n_items][len], y: [n_items][len+max_pitch], xcorr: [n_items][max_pitch]. */
int opusb200_pitch_xcorr(OpusB200Ctx *ctx, const float *x, const float *y, float *xcorr, int n_items, int len, int max_pitch,
                         int on_device);

/* exp_rotation (celt/vq.c:104) on n_items vectors of N floats in place; K: [n_items] pulses per vector. */
int opusb200_exp_rotation(OpusB200Ctx *ctx, float *X, const int *K, int n_items, int N, int dir, int stride, int spread, int on_device);

/* op_pvq_search (celt/vq.c:205, RTCD slot celt/x86/x86_celt_map.c:175): X [n_items][N] is replaced
 * by |X| as in the reference, iy [n_items][N] receives the pulses, yy [n_items] the return value. */
int opusb200_pvq_search(OpusB200Ctx *ctx, float *X, int *iy, float *yy, const int *K, int n_items, int N, int on_device);

/* silk_LPC_analysis_filter (silk/LPC_analysis_filter.c:49; called from silk/NSQ_del_dec.c:252 and the
 * float wrappers): out/in [n_items][len] int16, B [n_items][d] Q12 coefficients, d even, 6..24. */
int opusb200_silk_lpc_analysis_filter(OpusB200Ctx *ctx, int16_t *out, const int16_t *in, const int16_t *B, int n_items, int len,
                                      int d, int on_device);

/* silk_resampler (silk/resampler.c:77 silk_resampler_init, :183 silk_resampler) with persistent per-stream state, for every rate pair
 * silk_resampler_init accepts: for_enc = 1 (the encoder's input resampler): 8/12/16/24/48 kHz -> 8/12/16 kHz (48->16 kHz is config 4's);
 * for_enc = 0 (the decoder's output resampler): 8/12/16 kHz -> 8/12/16/24/48 kHz.  The four resampler functions of the reference: copy,
 * silk_resampler_private_up2_HQ (2x), silk_resampler_private_IIR_FIR (2x + FIR interpolation: the other up-sampling ratios),
 * silk_resampler_private_down_FIR (AR2 + FIR: 1:2, 1:3, 1:4, 1:6, and 2:3 / 3:4 with interpolated phases).  Other pairs: OPUSB200_BAD_ARG.
 * in [n][in_len], out [n][in_len*Fs_out/Fs_in]; in_len a multiple of 1 ms, at most 40 ms.  _create = _create_ex(for_enc = 1). */
typedef struct OpusB200SilkResampler OpusB200SilkResampler;
int opusb200_silk_resampler_create(OpusB200Ctx *ctx, int n_streams, int32_t Fs_in, int32_t Fs_out, OpusB200SilkResampler **out);
int opusb200_silk_resampler_create_ex(OpusB200Ctx *ctx, int n_streams, int32_t Fs_in, int32_t Fs_out, int for_enc, OpusB200SilkResampler **out);
void opusb200_silk_resampler_destroy(OpusB200SilkResampler *rs);
int opusb200_silk_resampler_reset(OpusB200SilkResampler *rs);
int opusb200_silk_resampler_process(OpusB200SilkResampler *rs, const int16_t *in, int16_t *out, int in_len, int on_device);

/* ---- A batch of range encoders (ec_enc, celt/entcode.h:62-88, celt/entenc.c) whose state stays on the device between calls, and the SILK
 * excitation coder on it: silk_encode_pulses (silk/encode_pulses.c:59) with silk_shell_encoder (silk/shell_coder.c:78) and silk_encode_signs
 * (silk/code_signs.c:40) — the first slice of the SILK frame pipeline.  Other symbols can be coded in front of and behind the
 * pulses in the same coder (opusb200_range_encoder_icdf = ec_enc_icdf), as silk_encode_indices does ahead of them in
 * silk_encode_frame_FLP (silk/float/encode_frame_FLP.c:290-300).  Bit-exact: buffer bytes, ec_tell_frac and final range.
 *   ec_enc_init   celt/entenc.c:108   opusb200_range_encoder_init(re, storage)
 *   ec_enc_icdf   celt/entenc.c:195   opusb200_range_encoder_icdf(re, symbols[n] (< 0: skip the stream), icdf, icdf_len <= 64, ftb, on_device)
 *   silk_encode_pulses                opusb200_silk_encode_pulses(re, signal_type[n], quant_offset_type[n], pulses [n][pulses_stride] int8,
 *                                     pulses_stride, frame_length (80..320: a multiple of 16, or 120), on_device)
 *   ec_tell_frac + ec_enc_done :256   opusb200_range_encoder_done(re, bytes [n][storage], tell_frac[n], final_range[n], error[n]) ---- */
typedef struct OpusB200RangeEncoder OpusB200RangeEncoder;
int opusb200_range_encoder_create(OpusB200Ctx *ctx, int n_streams, int buf_bytes, OpusB200RangeEncoder **out);
void opusb200_range_encoder_destroy(OpusB200RangeEncoder *re);
int opusb200_range_encoder_init(OpusB200RangeEncoder *re, int storage);
int opusb200_range_encoder_icdf(OpusB200RangeEncoder *re, const int32_t *symbols, const unsigned char *icdf, int icdf_len, int ftb, int on_device);
int opusb200_silk_encode_pulses(OpusB200RangeEncoder *re, const int32_t *signal_type, const int32_t *quant_offset_type, const int8_t *pulses,
                                int pulses_stride, int frame_length, int on_device);
int opusb200_range_encoder_done(OpusB200RangeEncoder *re, unsigned char *bytes, uint32_t *tell_frac, uint32_t *final_range, int32_t *error);
/* silk_encode_indices(psEncC, &enc, FrameIndex, 0, condCoding) (silk/encode_indices.c:35, with silk_NLSF_unpack): the side information of one SILK frame
 * per stream.  OpusB200SilkIndices has SideInfoIndices' fields in SideInfoIndices' order (silk/structs.h:129-141), so the reference's struct casts to
 * it; cond_coding [n]: CODE_INDEPENDENTLY 0, CODE_INDEPENDENTLY_NO_LTP_SCALING 1, CODE_CONDITIONALLY 2; fs_kHz (8 / 12 / 16) and nb_subfr (2 / 4)
 * select the tables as silk_control_encoder does (silk/control_codec.c:212-296).  ec_prevSignalType / ec_prevLagIndex are kept per stream inside the
 * object (opusb200_silk_coder_reset_state zeroes them).  Indices followed by opusb200_silk_encode_pulses = the bitstream of a SILK frame
 * (silk/float/encode_frame_FLP.c:290-300). */
typedef struct {
  int8_t GainsIndices[4], LTPIndex[4], NLSFIndices[17];
  int16_t lagIndex;
  int8_t contourIndex, signalType, quantOffsetType, NLSFInterpCoef_Q2, PERIndex, LTP_scaleIndex, Seed;
} OpusB200SilkIndices;
int opusb200_silk_encode_indices(OpusB200RangeEncoder *re, int fs_kHz, int nb_subfr, const OpusB200SilkIndices *indices, const int32_t *cond_coding,
                                 int on_device);
int opusb200_silk_coder_reset_state(OpusB200RangeEncoder *re);

/* silk_burg_modified_FLP (silk/float/burg_modified_FLP.c:39; called from find_LPC_FLP.c:60,65):
 * x [n_items][nb_subfr*subfr_length] -> A [n_items][D] prediction coefficients, nrg [n_items] residual energy. */
int opusb200_silk_burg_modified(OpusB200Ctx *ctx, float *A, float *nrg, const float *x, float minInvGain, int n_items,
                                int subfr_length, int nb_subfr, int D, int on_device);
/* silk_autocorrelation_FLP (silk/float/autocorrelation_FLP.c:36): results [n_items][count] */
int opusb200_silk_autocorrelation(OpusB200Ctx *ctx, float *results, const float *x, int n_items, int size, int count, int on_device);

/* ---- Batched CELT encoder: n independent streams, one frame each per call.
 * Replaces celt_encoder_init / celt_encoder_ctl / celt_encode_with_ec (celt/celt.h:157-160, celt/celt_encoder.c:240,1726,2953):
 * the output is the CELT payload exactly as celt_encode_with_ec(st, pcm, frame_size, compressed, nbCompressedBytes, NULL) writes it
 * (the Opus layer's TOC byte, dc_reject and mode decisions stay in src/opus_encoder.c:1830-2200 on the host).
 * Every complexity 0..10 (transient_analysis, tf_analysis, spreading_decision, two-pass coarse energy, the pitch pre-filter and
 * patch_transient_decision, the second MDCT and the theta RDO), CBR, VBR and constrained VBR: packet bytes and OPUS_GET_FINAL_RANGE
 * are bit-identical to the reference's C build at the same settings.  Inputs the Opus layer would add (st->analysis from
 * run_analysis, silk_info, a shared range coder for hybrid frames) are absent, as in CELT-only custom-mode use.
 * nbytes[s] < 0 flags a per-stream failure (OPUS_INTERNAL_ERROR -3). ---- */
typedef struct OpusB200Encoder OpusB200Encoder;
/* Fs: 48000, 24000, 16000, 12000 or 8000, like celt_encoder_init (celt/celt_encoder.c:233-262): below 48 kHz the input is zero-stuffed to
 * 48 kHz ahead of the pre-emphasis and the spectrum above Fs/2 is cleared; frame_size and the pcm of the encode calls are in Fs samples. */
int opusb200_encoder_create(int device, int32_t Fs, int channels, int n_streams, OpusB200Encoder **out);
void opusb200_encoder_destroy(OpusB200Encoder *enc);
int opusb200_encoder_reset(OpusB200Encoder *enc);                                   /* OPUS_RESET_STATE */
/* request = the OPUS_SET_*_REQUEST number of include/opus_defines.h: 4002 bitrate, 4006 vbr, 4020 vbr constraint, 4010 complexity
 * (0..10, default 5 like celt_encoder_init), 4014 packet loss %, 4036 lsb depth, 4046 phase inversion disabled, 10002 CELT_SET_PREDICTION (0 intra only, 1 no pitch pre-filter, 2 default), 10008 CELT_SET_CHANNELS (celt/celt.h:111; 1 on a stereo encoder = mono coding of the stereo
 * input, as the Opus layer does at low rates), 10012 CELT_SET_END_BAND (celt/celt.h:119; 13 NB, 17 WB,
 * 19 SWB, 21 FB: what the Opus layer sets from the bandwidth); other requests
 * return OPUSB200_UNIMPLEMENTED */
int opusb200_encoder_ctl_set(OpusB200Encoder *enc, int request, int32_t value);
/* pcm [n][frame_size*channels] interleaved float (+-1.0); packets [n][pkt_stride]; nbytes [n] */
int opusb200_encode_float(OpusB200Encoder *enc, const float *pcm, int frame_size, unsigned char *packets, int pkt_stride,
                          int max_data_bytes, int32_t *nbytes);
int opusb200_encode_device(OpusB200Encoder *enc, const float *d_pcm, int frame_size, unsigned char *d_packets, int pkt_stride,
                           int max_data_bytes, int32_t *d_nbytes);
int opusb200_encoder_get_final_range(OpusB200Encoder *enc, uint32_t *final_range);   /* OPUS_GET_FINAL_RANGE */
int opusb200_encoder_sync(OpusB200Encoder *enc);
void *opusb200_encoder_cuda_stream(OpusB200Encoder *enc);
int64_t opusb200_encoder_launch_count(OpusB200Encoder *enc);

/* ---- Batched Opus encoder: the reference's public encoder API for n independent streams (include/opus.h:170-355, src/opus_encoder.c).
 *   opus_encoder_create(Fs, channels, application, &err)   include/opus.h:208   opusb200_opus_encoder_create(..., n_streams, ...)
 *   opus_encode(st, pcm, frame_size, data, max_data_bytes) include/opus.h:266   opusb200_opus_encode(enc, pcm, frame_size, packets, stride, max_data_bytes, nbytes[])
 *   opus_encode_float(...)                                 include/opus.h:343   opusb200_opus_encode_float(...)
 *   opus_encoder_ctl(st, request, value)                   include/opus.h:355   opusb200_opus_encoder_ctl(enc, request, value)
 *   opus_encoder_ctl(st, OPUS_GET_FINAL_RANGE(&r))                              opusb200_opus_encoder_get_final_range(enc, r[])
 *   opus_encoder_destroy(st)                               include/opus.h:350   opusb200_opus_encoder_destroy(enc)
 * application: OPUS_APPLICATION_RESTRICTED_CELT (2053) or OPUS_APPLICATION_RESTRICTED_LOWDELAY (2051), the two applications whose frames are
 * always CELT-only; VOIP / AUDIO / RESTRICTED_SILK can select SILK and return OPUSB200_UNIMPLEMENTED.  Everything opus_encode() does for such
 * an encoder runs on the device, per stream: run_analysis (src/analysis.c: tonality / bandwidth / music detector, from complexity 7 and Fs >= 16 kHz),
 * the mono/stereo and bandwidth decisions with their hysteresis, dc_reject, the stereo width fade, the float API's NaN / level guard, the CELT
 * encoder with st->analysis valid, the TOC byte.  Packets (TOC included) and OPUS_GET_FINAL_RANGE are bit-identical to the reference's C build
 * given the same calls.  Frame sizes 2.5 / 5 / 10 / 20 ms and 40 .. 120 ms (20 ms frames coded one by one and repacketised, code 1 / 2 / 3 packets with
 * the CBR padding: src/opus_encoder.c:1698-1838, src/repacketizer.c; pkt_stride must hold the packet: up to max_data_bytes); DTX and budgets below
 * 3 bytes per frame return OPUSB200_UNIMPLEMENTED for the call.  Setters (request numbers of include/opus_defines.h, value semantics of opus_encoder_ctl):
 * 4002 bitrate (-1000 OPUS_AUTO, -1 OPUS_BITRATE_MAX), 4004 max bandwidth, 4006 vbr, 4008 bandwidth, 4010 complexity, 4014 packet loss %,
 * 4020 vbr constraint, 4022 force channels, 4024 signal, 4028 OPUS_RESET_STATE, 4036 lsb depth, 4042 prediction disabled, 4046 phase inversion disabled.
 * nbytes[s] = packet length of stream s (TOC included) or a negative error code. ---- */
typedef struct OpusB200OpusEncoder OpusB200OpusEncoder;
int opusb200_opus_encoder_create(int device, int32_t Fs, int channels, int application, int n_streams, OpusB200OpusEncoder **out);
void opusb200_opus_encoder_destroy(OpusB200OpusEncoder *enc);
int opusb200_opus_encoder_reset(OpusB200OpusEncoder *enc);
int opusb200_opus_encoder_ctl(OpusB200OpusEncoder *enc, int request, int32_t value);
/* host buffers: pcm [n][frame_size*channels] interleaved, packets [n][pkt_stride] (pkt_stride >= max_data_bytes), nbytes [n] */
int opusb200_opus_encode(OpusB200OpusEncoder *enc, const int16_t *pcm, int frame_size, unsigned char *packets, int pkt_stride, int max_data_bytes,
                         int32_t *nbytes);
int opusb200_opus_encode_float(OpusB200OpusEncoder *enc, const float *pcm, int frame_size, unsigned char *packets, int pkt_stride, int max_data_bytes,
                               int32_t *nbytes);
/* The same call in two halves, for a batch server: submit enqueues the PCM upload, the kernels and the packet read-back of one call and returns;
 * wait returns when every submitted call's packets and nbytes[] are on the host.  Two calls may be in flight (double-buffered staging: the upload
 * of call k+1 and the read-back of call k-1 run under the kernels of call k); host buffers stay valid until the wait and should be pinned
 * (opusb200_host_alloc) for the transfers to overlap.  Per-stream results are those of opus_encode() called in submit order. */
int opusb200_opus_encode_submit(OpusB200OpusEncoder *enc, const void *pcm, int pcm_is_int16, int frame_size, unsigned char *packets, int pkt_stride,
                                int max_data_bytes, int32_t *nbytes);
int opusb200_opus_encode_wait(OpusB200OpusEncoder *enc);
/* device buffers (d_pcm int16 or float as pcm_is_int16 says); work is enqueued on the encoder's stream */
int opusb200_opus_encode_device(OpusB200OpusEncoder *enc, const void *d_pcm, int pcm_is_int16, int frame_size, unsigned char *d_packets,
                                int pkt_stride, int max_data_bytes, int32_t *d_nbytes);
int opusb200_opus_encoder_get_final_range(OpusB200OpusEncoder *enc, uint32_t *final_range);
int opusb200_opus_encoder_sync(OpusB200OpusEncoder *enc);
void *opusb200_opus_encoder_cuda_stream(OpusB200OpusEncoder *enc);
int64_t opusb200_opus_encoder_launch_count(OpusB200OpusEncoder *enc);
/* Profiling aid: when enabled every kernel of a call is bracketed by CUDA events on the encoder's stream and the call waits for them (so not
 * for throughput runs).  get_timing: accumulated device ms, ms_kernels[5] = [Opus-layer pre-pass, Opus-layer analysis,
 * CELT transient / tone recurrences, CELT analysis, CELT frame coder]. */
int opusb200_opus_encoder_set_timing(OpusB200OpusEncoder *enc, int enable);
int opusb200_opus_encoder_get_timing(OpusB200OpusEncoder *enc, double *ms_kernels, int64_t *calls);

/* ---- CELT encoder front end, fused: celt_preemphasis (celt/celt_encoder.c:557) -> compute_mdcts (:511) -> compute_band_energies
 * (celt/bands.c:151) -> amp2Log2 (celt/quant_bands.c:553) -> normalise_bands (celt/bands.c:169), i.e. lines 2007-2265 of
 * celt_encode_with_ec with the pitch pre-filter off.  The object keeps, per stream and channel, the pre-emphasis memory and
 * the `overlap` tail of the previous frame (CELTEncoder preemph_memE / prefilter_mem tail).
 * pcm [n][frame_size*channels] interleaved floats (+-1.0), short_blocks [n] (the transient decision per stream, NULL = none),
 * X [n][channels*frame_size] normalised MDCT, bandE / bandLogE [n][channels*21]. ---- */
typedef struct OpusB200CeltAnalysis OpusB200CeltAnalysis;
int opusb200_celt_analysis_create(OpusB200Ctx *ctx, int n_streams, int channels, OpusB200CeltAnalysis **out);
void opusb200_celt_analysis_destroy(OpusB200CeltAnalysis *a);
int opusb200_celt_analysis_reset(OpusB200CeltAnalysis *a);
int opusb200_celt_analysis_process(OpusB200CeltAnalysis *a, const float *pcm, const int32_t *short_blocks, float *X, float *bandE,
                                   float *bandLogE, int frame_size, int on_device);

/* ---- SILK noise-shaping quantizer, batched: silk_NSQ / silk_NSQ_del_dec (silk/main.h:250-300; silk/NSQ.c:76,
 * silk/NSQ_del_dec.c:113), called once per frame from silk_NSQ_wrapper_FLP (silk/float/wrappers_FLP.c:106-170).
 * One object holds the silk_nsq_state (silk/structs.h:56-69) of n_streams encoders in device memory. ---- */
typedef struct OpusB200SilkNsq OpusB200SilkNsq;
/* the fields of silk_encoder_state (silk/structs.h) the quantizers read; identical for every stream of the batch */
typedef struct {
  int32_t fs_khz;                 /* 8, 12, 16: subfr_length = 5*fs_khz, ltp_mem_length = 20*fs_khz */
  int32_t nb_subfr;               /* 2 or 4 */
  int32_t predictLPCOrder;        /* 10 or 16 */
  int32_t shapingLPCOrder;        /* even, <= 24 */
  int32_t warping_Q16;
  int32_t nStatesDelayedDecision; /* 1..4; >1 or warping_Q16>0 selects silk_NSQ_del_dec (wrappers_FLP.c:163) */
} OpusB200NsqConfig;
/* per stream and frame: SideInfoIndices fields + the arguments of silk_NSQ (silk/main.h:250-266) */
typedef struct {
  int32_t signalType, quantOffsetType, NLSFInterpCoef_Q2;
  int32_t Seed;                   /* in/out: silk_NSQ_del_dec rewrites psIndices->Seed */
  int32_t HarmShapeGain_Q14[4], Tilt_Q14[4], LF_shp_Q14[4], Gains_Q16[4], pitchL[4];
  int32_t Lambda_Q10, LTP_scale_Q14;
  int16_t PredCoef_Q12[2 * 16];   /* [2][MAX_LPC_ORDER] */
  int16_t LTPCoef_Q14[4 * 5];     /* [MAX_NB_SUBFR][LTP_ORDER] */
  int16_t AR_Q13[4 * 24];         /* [MAX_NB_SUBFR][MAX_SHAPE_LPC_ORDER] */
} OpusB200NsqFrame;
int opusb200_silk_nsq_create(OpusB200Ctx *ctx, int n_streams, const OpusB200NsqConfig *cfg, OpusB200SilkNsq **out);
void opusb200_silk_nsq_destroy(OpusB200SilkNsq *q);
int opusb200_silk_nsq_reset(OpusB200SilkNsq *q);
/* frames [n_streams], x16 [n_streams][frame_length], pulses [n_streams][frame_length] (out) */
int opusb200_silk_nsq_process(OpusB200SilkNsq *q, OpusB200NsqFrame *frames, const int16_t *x16, int8_t *pulses, int on_device);
/* debug / parity: the carried state as [n_streams][1325] int32 words (host):
 * xq[640] sLTP_shp_Q14[640] sLPC_Q14[16] sAR2_Q14[24] sLF_AR_shp_Q14 sDiff_shp_Q14 lagPrev rand_seed prev_gain_Q16 */
int opusb200_silk_nsq_get_state(OpusB200SilkNsq *q, int32_t *out);

/* ---- dnn/ inference primitives, batched over n streams that share one set of weights (dnn/nnet.h:88-93, dnn/nnet_arch.h:69-243).
 * Activations are row-major [n][dim] float.  int8 layers run on the tensor cores (exact integer dot products). ---- */
#define OPUSB200_ACTIVATION_LINEAR 0   /* dnn/nnet.h:34-40 */
#define OPUSB200_ACTIVATION_SIGMOID 1
#define OPUSB200_ACTIVATION_TANH 2
#define OPUSB200_ACTIVATION_RELU 3
#define OPUSB200_ACTIVATION_SOFTMAX 4
#define OPUSB200_ACTIVATION_SWISH 5
#define OPUSB200_ACTIVATION_EXP 6
/* = LinearLayer (dnn/nnet.h:66-76), host pointers */
typedef struct {
  const float *bias;
  const float *subias;
  const int8_t *weights;        /* 8x4-blocked int8 (cgemv8x4), with per-row `scale` */
  const float *float_weights;   /* [j*nb_outputs + i], or 8x4 blocks when weights_idx != NULL */
  const int *weights_idx;       /* block-sparse: per 8 rows a count followed by that many column positions */
  const float *diag;            /* GRU recurrent layers only (nb_outputs == 3*nb_inputs) */
  const float *scale;
  int nb_inputs;
  int nb_outputs;
} OpusB200LinearLayer;
/* = Conv2dLayer (dnn/nnet.h:79-86) */
typedef struct {
  const float *bias;
  const float *float_weights;   /* [out][in][ktime][kheight] */
  int in_channels, out_channels, ktime, kheight;
} OpusB200Conv2dLayer;
typedef struct OpusB200Linear OpusB200Linear;
typedef struct OpusB200Conv2d OpusB200Conv2d;
/* Uploads a layer.  su_bias != 0 selects the x86 reference convention for int8 layers (dnn/vec_avx.h:41 USE_SU_BIAS:
 * activations quantised to unsigned 127+q and `subias` as the bias, dnn/nnet_arch.h:136-139); 0 = the generic C convention. */
int opusb200_nnet_linear_create(OpusB200Ctx *ctx, const OpusB200LinearLayer *layer, int su_bias, OpusB200Linear **out);
void opusb200_nnet_linear_destroy(OpusB200Linear *l);
/* compute_linear (dnn/nnet_arch.h:122): out [n][nb_outputs], in [n][nb_inputs] */
int opusb200_nnet_compute_linear(OpusB200Ctx *ctx, const OpusB200Linear *l, float *out, const float *in, int n, int on_device);
/* compute_activation (dnn/nnet_arch.h:69): [n][N]; softmax normalises each row */
int opusb200_nnet_compute_activation(OpusB200Ctx *ctx, float *out, const float *in, int n, int N, int activation, int on_device);
/* compute_generic_dense (dnn/nnet.c:60) */
int opusb200_nnet_compute_dense(OpusB200Ctx *ctx, const OpusB200Linear *l, float *out, const float *in, int n, int activation, int on_device);
/* compute_generic_gru (dnn/nnet.c:76): state [n][N] in/out, in [n][input_weights.nb_inputs] */
int opusb200_nnet_compute_gru(OpusB200Ctx *ctx, const OpusB200Linear *input_weights, const OpusB200Linear *recurrent_weights, float *state,
                              const float *in, int n, int on_device);
/* compute_glu (dnn/nnet.c:107): out = in * sigmoid(W in + b); out may alias in */
int opusb200_nnet_compute_glu(OpusB200Ctx *ctx, const OpusB200Linear *l, float *out, const float *in, int n, int on_device);
/* compute_generic_conv1d (dilation 1; mem [n][nb_inputs-input_size]) / compute_generic_conv1d_dilation
 * (mem [n][input_size*dilation*(ksize-1)]) (dnn/nnet.c:124-153) */
int opusb200_nnet_compute_conv1d(OpusB200Ctx *ctx, const OpusB200Linear *l, float *out, float *mem, const float *in, int n, int input_size,
                                 int dilation, int activation, int on_device);
int opusb200_nnet_conv2d_create(OpusB200Ctx *ctx, const OpusB200Conv2dLayer *layer, OpusB200Conv2d **out);
void opusb200_nnet_conv2d_destroy(OpusB200Conv2d *q);
/* compute_conv2d (dnn/nnet_arch.h:217): in [n][in_ch*(height+kheight-1)], mem [n][(ktime-1)*in_ch*(height+kheight-1)],
 * out [n][out_ch*hstride] */
int opusb200_nnet_compute_conv2d(OpusB200Ctx *ctx, const OpusB200Conv2d *q, float *out, float *mem, const float *in, int n, int height,
                                 int hstride, int activation, int on_device);

/* ---- FARGAN vocoder (dnn/fargan.h:59-65, dnn/fargan.c): one synthesis state per stream, the model shared ----------------
 * layers: the 20 LinearLayers of the generated `FARGAN` model struct (fargan_data.h) in declaration order, i.e. what
 * init_fargan() fills from the weight blob (cond_net_pembed ... sig_net_gain_dense_out); sizes are checked like linear_init.
 * su_bias as for opusb200_nnet_linear_create.  All buffers are [n_streams][width] row-major floats.
 *   opusb200_fargan_create      = fargan_init + init_fargan / fargan_load_model   dnn/fargan.c:173-189
 *   opusb200_fargan_cont        = fargan_cont(st, pcm0[320], features0[5*20])     dnn/fargan.c:143
 *   opusb200_fargan_synthesize  = fargan_synthesize(st, pcm[160], features[20])   dnn/fargan.c:215  (BAD_ARG before _cont) */
typedef struct OpusB200Fargan OpusB200Fargan;
int opusb200_fargan_create(OpusB200Ctx *ctx, const OpusB200LinearLayer *layers, int n_layers, int su_bias, int n_streams, OpusB200Fargan **out);
void opusb200_fargan_destroy(OpusB200Fargan *f);
int opusb200_fargan_cont(OpusB200Fargan *f, const float *pcm0, const float *features0, int on_device);
int opusb200_fargan_synthesize(OpusB200Fargan *f, float *pcm, const float *features, int on_device);

/* ---- DRED encoder network (dnn/dred_rdovae_enc.h:48, dnn/dred_rdovae_enc.c:47-116; called per double frame from dred_process_frame,
 * dnn/dred_encoder.c): layers = the 24 LinearLayers of the generated RDOVAEEnc struct in declaration order (enc_dense1, enc_zdense, gdense1, gdense2,
 * enc_conv_dense1..5, enc_gru1..5 input / recurrent, enc_conv1..5), n streams in lockstep, one RDOVAEEncState per stream on the device.
 *   opusb200_dred_encode_dframe(enc, latents [n][latent_dim], initial_state [n][state_dim], input [n][2 * 20 features], on_device)
 * = dred_rdovae_encode_dframe for every stream; latents and state bit-identical to the reference given the same model.  The front end
 * (16 kHz conversion, feature extraction, buffers) and the entropy coding of the latents follow below. ---- */
typedef struct OpusB200DredEnc OpusB200DredEnc;
int opusb200_dred_encoder_create(OpusB200Ctx *ctx, const OpusB200LinearLayer *layers, int n_layers, int su_bias, int latent_dim, int state_dim,
                                 int n_streams, OpusB200DredEnc **out);
void opusb200_dred_encoder_destroy(OpusB200DredEnc *enc);
int opusb200_dred_encoder_reset(OpusB200DredEnc *enc);
int opusb200_dred_encode_dframe(OpusB200DredEnc *enc, float *latents, float *initial_state, const float *input, int on_device);
/* The encoder's front end (DREDEnc of dnn/dred_encoder.h:40-60; dred_encoder_init :80, dred_encoder_reset :70, dred_compute_latents :215-250 with
 * dred_convert_to_16k :133-213 and dred_process_frame :84-107): PCM at Fs in, per double frame of 20 ms two feature vectors
 * (opusb200_lpcnet_encoder_*) and one latent / state pair pushed at the front of the 104-deep buffers dred_encode_silk_frame reads.
 *   init_frontend: Fs in {8000, 12000, 16000, 24000, 48000}, channels 1 or 2, linear / conv = the PitchDNN layers (the other half of
 *                  dred_encoder_load_model); opusb200_dred_encoder_reset then is dred_encoder_reset.
 *   compute_latents: pcm [n][frame_size * channels] float in [-1, 1), frame_size * 16000 divisible by Fs; streams in lockstep.
 *   get_latents: the buffers in the reference's order (entry 0 = newest), info [4] = input_buffer_fill, dred_offset, latent_offset, latents_buffer_fill.
 *   load_stats: the statistical model of the latents (dred_rdovae_stats_data.h): 8 uint8 tables of 16 levels x dim in the order latent
 *               quant_scales / dead_zone / r / p0, then the state's four.
 *   encode_silk_frame: dred_encode_silk_frame (dnn/dred_encoder.c:281-383) for every stream — the DRED payload of this packet from the buffered
 *               latents: buf [n][buf_stride], activity_mem [n][416] (one byte per 2.5 ms, newest first), nbytes [n] = the reference's return
 *               value per stream (0 = no DRED this time); q0 / dQ / qmax / max_chunks / max_bytes are the batch's.  Bit-exact. */
int opusb200_dred_encoder_init_frontend(OpusB200DredEnc *enc, int Fs, int channels, const OpusB200LinearLayer *linear, int n_linear,
                                        const OpusB200Conv2dLayer *conv, int n_conv, int su_bias);
int opusb200_dred_compute_latents(OpusB200DredEnc *enc, const float *pcm, int frame_size, int extra_delay, int on_device);
int opusb200_dred_encoder_get_latents(OpusB200DredEnc *enc, float *latents, float *state, int32_t *info);
int opusb200_dred_encoder_load_stats(OpusB200DredEnc *enc, const unsigned char *const *tables, int n_tables);
int opusb200_dred_encode_silk_frame(OpusB200DredEnc *enc, unsigned char *buf, int buf_stride, int max_chunks, int max_bytes, int q0, int dQ, int qmax,
                                    const unsigned char *activity_mem, int32_t *nbytes, int on_device);

/* silk_process_gains_FLP (silk/float/process_gains_FLP.c:36-103; call site silk/float/encode_frame_FLP.c:166) with silk_gains_quant
 * (silk/gain_quant.c:40-95): the gain quantiser of a SILK frame — another slice of SURVEY 8(f).3.  One record per stream holds what the function
 * reads and writes in silk_encoder_state_FLP / silk_encoder_control_FLP: in: Gains, ResNrg, LTPredCodGain, input_quality, coding_quality, signalType,
 * quantOffsetType, SNR_dB_Q7, input_tilt_Q15, speech_activity_Q8, condCoding (0 / 1 / 2 as silk/define.h:97-99), LastGainIndex; out: Gains
 * (quantised), GainsUnq_Q16, GainsIndices, LastGainIndex, lastGainIndexPrev, quantOffsetType, Lambda.  nb_subfr 2 or 4.  Bit-exact. */
typedef struct {
  float Gains[4], ResNrg[4];
  float LTPredCodGain, input_quality, coding_quality, Lambda;
  int32_t GainsUnq_Q16[4];
  int32_t signalType, quantOffsetType, SNR_dB_Q7, input_tilt_Q15, speech_activity_Q8, condCoding;
  int8_t GainsIndices[4];
  int8_t LastGainIndex, lastGainIndexPrev, pad[2];
} OpusB200SilkGains;
int opusb200_silk_process_gains(OpusB200Ctx *ctx, OpusB200SilkGains *frames, int n, int nb_subfr, int subfr_length, int nStatesDelayedDecision,
                                int on_device);

/* silk_process_NLSFs (silk/process_NLSFs.c:35-107; call site silk/float/wrappers_FLP.c:84 from find_pred_coefs_FLP.c:108): the NLSF quantiser of a
 * SILK frame — laroia weights, interpolation, silk_NLSF_encode (first-stage VQ, survivors, delayed-decision trellis), silk_NLSF_decode, the stabiliser
 * and silk_NLSF2A for both half-frames.  One record per stream: in: NLSF_Q15 (unquantised), prev_NLSFq_Q15, NLSFInterpCoef_Q2, signalType,
 * useInterpolatedNLSFs, speech_activity_Q8, NLSF_MSVQ_Survivors (2..32); out: NLSF_Q15 (quantised), NLSFIndices [order + 1], PredCoef_Q12 [2][order].
 * fs_kHz 8 / 12 (order 10) or 16 (order 16).  Bit-exact. */
typedef struct {
  int16_t NLSF_Q15[16];
  int16_t prev_NLSFq_Q15[16];
  int16_t PredCoef_Q12[2][16];
  int8_t NLSFIndices[17];
  int8_t NLSFInterpCoef_Q2, signalType, useInterpolatedNLSFs;
  int32_t speech_activity_Q8, NLSF_MSVQ_Survivors;
} OpusB200SilkNlsf;
int opusb200_silk_process_nlsfs(OpusB200Ctx *ctx, OpusB200SilkNlsf *frames, int n, int fs_kHz, int nb_subfr, int on_device);

/* silk_find_LPC_FLP (silk/float/find_LPC_FLP.c:37-105; call site silk/float/find_pred_coefs_FLP.c:105): the LPC estimation of a SILK frame — Burg over
 * the frame and, with NLSF interpolation on, over its second half; silk_A2NLSF; the search over the interpolation factor (silk_interpolate, silk_NLSF2A,
 * the float LPC analysis filter, residual energies).  x [n][nb_subfr * (subfr_length + order)]: the pre-whitened frame with `order` samples of history
 * in front of every subframe; order 10 or 16.  in: prev_NLSFq_Q15, minInvGain, useInterpolatedNLSFs, first_frame_after_reset; out: NLSF_Q15 (the input
 * of opusb200_silk_process_nlsfs), NLSFInterpCoef_Q2.  Integer parts exact; the Burg recursion is double precision (a17). */
typedef struct {
  int16_t NLSF_Q15[16];
  int16_t prev_NLSFq_Q15[16];
  float minInvGain;
  int32_t useInterpolatedNLSFs, first_frame_after_reset;
  int32_t NLSFInterpCoef_Q2;
} OpusB200SilkFindLpc;
int opusb200_silk_find_lpc(OpusB200Ctx *ctx, OpusB200SilkFindLpc *frames, const float *x, int n, int order, int nb_subfr, int subfr_length, int on_device);

/* silk_find_pred_coefs_FLP (silk/float/find_pred_coefs_FLP.c:35-116; call site silk/float/encode_frame_FLP.c:161): the whole prediction analysis of a
 * SILK frame in one call per batch.  Voiced frames: silk_find_LTP_FLP, silk_quant_LTP_gains (silk_VQ_WMat_EC), silk_LTP_scale_ctrl_FLP,
 * silk_LTP_analysis_filter_FLP; all frames: silk_find_LPC_FLP, silk_process_NLSFs, silk_residual_energy_FLP.
 * res_pitch / x [n][sig_stride] float: the pitch-analysis residual and the input signal of every stream with the current frame at frame_off (history in
 * front: >= 18 ms + 2 + order samples; 5 samples behind the frame are read).  One record per stream — in: Gains, pitchL, coding_quality, signalType,
 * condCoding, first_frame_after_reset, useInterpolatedNLSFs, PacketLoss_perc, nFramesPerPacket, LBRR_flag, SNR_dB_Q7, speech_activity_Q8,
 * NLSF_MSVQ_Survivors; in / out: prev_NLSFq_Q15, sum_log_gain_Q7; out: LTPCoef, LTPIndex, PERIndex, LTP_scaleIndex, LTP_scale, LTPredCodGain,
 * NLSFInterpCoef_Q2, NLSFIndices, PredCoef [2][order], ResNrg.  The integer parts are exact; see tests/test_gpu_ops.py for the float bound. */
typedef struct {
  float Gains[4];
  int32_t pitchL[4];
  float coding_quality;
  int32_t signalType, condCoding, first_frame_after_reset, useInterpolatedNLSFs, PacketLoss_perc, nFramesPerPacket, LBRR_flag, SNR_dB_Q7, speech_activity_Q8,
      NLSF_MSVQ_Survivors;
  int16_t prev_NLSFq_Q15[16];
  int32_t sum_log_gain_Q7;
  float LTPCoef[20], LTPredCodGain, LTP_scale, PredCoef[2][16], ResNrg[4];
  int8_t LTPIndex[4], PERIndex, LTP_scaleIndex, NLSFInterpCoef_Q2, pad;
  int8_t NLSFIndices[17], pad2[3];
} OpusB200SilkPredCoefs;
int opusb200_silk_find_pred_coefs(OpusB200Ctx *ctx, OpusB200SilkPredCoefs *frames, const float *res_pitch, const float *x, int n, int sig_stride, int frame_off,
                                  int fs_kHz, int nb_subfr, int on_device);

/* silk_noise_shape_analysis_FLP (silk/float/noise_shape_analysis_FLP.c:147-350; call site silk/float/encode_frame_FLP.c:156): the noise shaping
 * analysis of a SILK frame — coding / input quality and the SNR adjustment, the sparseness measure (quantOffsetType of unvoiced frames), per subframe
 * the windowed (warped) autocorrelation, Schur recursion, step-up, bandwidth expansion and coefficient limiting (AR, 24 slots per subframe) and the
 * initial gains; gain tweaking; low-frequency, tilt and harmonic shaping with their smoothers (HarmShapeGain_smth / Tilt_smth carried in the record).
 * pitch_res [n][res_stride]: the frame's residual; x [n][x_stride] with the frame at x_off (5 ms in front, 5 ms look-ahead behind).  shapingLPCOrder
 * even, <= 24 (12 / 14 / 16 / 24 by complexity, silk/control_codec.c).  Float arithmetic in the reference's order; doubles where it uses them. */
typedef struct {
  int32_t SNR_dB_Q7, input_quality_bands_Q15[2], useCBR, speech_activity_Q8, signalType, warping_Q16;
  float LTPCorr, predGain;
  int32_t pitchL[4];
  float HarmShapeGain_smth, Tilt_smth;
  float input_quality, coding_quality;
  int32_t quantOffsetType;
  float AR[4 * 24], Gains[4], LF_MA_shp[4], LF_AR_shp[4], HarmShapeGain[4], Tilt[4];
} OpusB200SilkNoiseShape;
int opusb200_silk_noise_shape_analysis(OpusB200Ctx *ctx, OpusB200SilkNoiseShape *frames, const float *pitch_res, int res_stride, const float *x, int x_stride,
                                       int x_off, int n, int fs_kHz, int nb_subfr, int shapingLPCOrder, int on_device);

/* silk_find_pitch_lags_FLP (silk/float/find_pitch_lags_FLP.c:36-132; call site silk/float/encode_frame_FLP.c:151) with silk_pitch_analysis_core_FLP
 * (silk/float/pitch_analysis_core_FLP.c:66-630): the pitch analysis of a SILK frame — the whitened residual of history + frame + look-ahead (res out,
 * the input of opusb200_silk_noise_shape_analysis and opusb200_silk_find_pred_coefs), the voicing decision, the four pitch lags with their lag / contour
 * indices, LTPCorr and predGain.  x [n][x_stride] with the frame at x_off (20 ms in front, 2 ms behind); res [n][res_stride >= (22 + 5 nb_subfr) fs_kHz].
 * lpc_order in {6, 8, 10, 12, 16} (pitchEstimationLPCOrder), complexity 0..2 (pitchEstimationComplexity).  in: signalType (0 = no voice activity: no
 * search), first_frame_after_reset, speech_activity_Q8, prevSignalType, input_tilt_Q15, prevLag, pitchEstimationThreshold_Q16, LTPCorr. */
typedef struct {
  int32_t signalType;
  int32_t first_frame_after_reset, speech_activity_Q8, prevSignalType, input_tilt_Q15, prevLag, pitchEstimationThreshold_Q16;
  float LTPCorr;
  float predGain;
  int32_t pitchL[4];
  int32_t lagIndex, contourIndex;
} OpusB200SilkPitch;
int opusb200_silk_find_pitch_lags(OpusB200Ctx *ctx, OpusB200SilkPitch *frames, float *res, int res_stride, const float *x, int x_stride, int x_off, int n, int fs_kHz,
                                  int nb_subfr, int lpc_order, int complexity, int on_device);

/* silk_VAD_Init (silk/VAD.c:46-76), silk_encode_do_VAD_FLP (silk/float/encode_frame_FLP.c:45-82; call sites silk/enc_API.c:465,480) with
 * silk_VAD_GetSA_Q8_c (silk/VAD.c:82-297) inside, and silk_HP_variable_cutoff (silk/HP_variable_cutoff.c:39-77; call site silk/enc_API.c:409): the
 * voice activity detector of a SILK frame — another slice of SURVEY 8(f).3.  One record per stream = silk_VAD_state (silk/structs.h:74-85; HPstate
 * widened to 32 bits) followed by the silk_encoder_state fields the three functions read and write: speech_activity_Q8, input_tilt_Q15,
 * input_quality_bands_Q15, indices.signalType, noSpeechCounter, inDTX, VAD_flags[nFramesEncoded] (VAD_flag), prevSignalType, prevLag,
 * variable_HP_smth1_Q15.  The caller carries the record from frame to frame.
 *   opusb200_silk_vad_init          fills n records on the host (constants; no device work)
 *   opusb200_silk_encode_do_vad     pcm [n][pcm_stride] int16, the frame_length samples the reference reads at inputBuf + 1; frame_length = 10 or 20 ms
 *                                   at fs_kHz 8 / 12 / 16; activity [n] = the Opus-level decision (0 = VAD_NO_ACTIVITY) or NULL = active everywhere
 *   opusb200_silk_hp_variable_cutoff  updates variable_HP_smth1_Q15 where prevSignalType is voiced (prevLag in 1..32767 then)
 * on_device != 0: frames / pcm / activity are device pointers and the call only enqueues.  All integer, bit-exact. */
typedef struct {
  int32_t AnaState[2], AnaState1[2], AnaState2[2], XnrgSubfr[4], NrgRatioSmth_Q8[4], HPstate, NL[4], inv_NL[4], NoiseLevelBias[4], counter;
  int32_t speech_activity_Q8, input_tilt_Q15, input_quality_bands_Q15[4];
  int32_t signalType, noSpeechCounter, inDTX, VAD_flag;
  int32_t prevSignalType, prevLag, variable_HP_smth1_Q15;
} OpusB200SilkVad;
int opusb200_silk_vad_init(OpusB200SilkVad *frames, int n);
int opusb200_silk_encode_do_vad(OpusB200Ctx *ctx, OpusB200SilkVad *frames, const int16_t *pcm, int pcm_stride, const int32_t *activity, int n, int frame_length,
                                int fs_kHz, int on_device);
int opusb200_silk_hp_variable_cutoff(OpusB200Ctx *ctx, OpusB200SilkVad *frames, int n, int fs_kHz, int on_device);

/* ---- LPCNet feature extraction (dnn/lpcnet_enc.c, freq.c, pitchdnn.c, burg.c; struct LPCNetEncState of dnn/lpcnet_private.h:25-43): n streams
 * that take one 10 ms frame of 16 kHz audio per call.  It produces the 36 features the DRED encoder codes (dnn/dred_encoder.c:101-102) and is what
 * lpcnet_plc_conceal runs over the decoded history (dnn/lpcnet_plc.c:160-177).
 *   lpcnet_encoder_init + load_model (lpcnet_enc.c:53-62)          opusb200_lpcnet_encoder_create   linear = the 6 LinearLayers of the generated PitchDNN
 *                                                                   struct (pitchdnn_data.h: dense_if_upsampler_1, dense_if_upsampler_2, dense_downsampler,
 *                                                                   dense_final_upsampler, gru_1_input, gru_1_recurrent), conv = conv2d_1, conv2d_2
 *   lpcnet_encoder_init on a live object                            opusb200_lpcnet_encoder_reset
 *   lpcnet_compute_single_frame_features / _float (:217-230)        opusb200_lpcnet_compute_single_frame_features (pcm [n][160] int16 or float in
 *                                                                   int16 range, features [n][36] = 18 cepstral, pitch, correlation, 16 LPC)
 *   burg_cepstral_analysis (freq.c:179-190)                         opusb200_burg_cepstral_analysis (x [n][160], ceps [n][36])
 * Float arithmetic in the reference's order (generic C build); the double-precision libm calls on the path (log, pow, exp, sqrt) are the device's,
 * so results can differ from the reference in the last bit of a float (tests/test_gpu_lpcnet_enc.py states the bound). ---- */
typedef struct OpusB200LpcnetEnc OpusB200LpcnetEnc;
int opusb200_lpcnet_encoder_create(OpusB200Ctx *ctx, const OpusB200LinearLayer *linear, int n_linear, const OpusB200Conv2dLayer *conv, int n_conv,
                                   int su_bias, int n_streams, OpusB200LpcnetEnc **out);
void opusb200_lpcnet_encoder_destroy(OpusB200LpcnetEnc *enc);
int opusb200_lpcnet_encoder_reset(OpusB200LpcnetEnc *enc);
int opusb200_lpcnet_compute_single_frame_features(OpusB200LpcnetEnc *enc, const void *pcm, int pcm_is_int16, float *features, int on_device);
/* parity aid (host buffers): if_features [n][88], xcorr_features [n][224], lpc [n][16], misc [n][3] = dnn_pitch, pitch_filt, mem_preemph */
int opusb200_lpcnet_encoder_get_state(OpusB200LpcnetEnc *enc, float *if_features, float *xcorr_features, float *lpc, float *misc);
int opusb200_burg_cepstral_analysis(OpusB200Ctx *ctx, float *ceps, const float *x, int n, int on_device);

/* ---- Neural packet-loss concealment (dnn/lpcnet_plc.c, struct LPCNetPLCState of dnn/lpcnet_private.h:48-71): n streams that take their update /
 * conceal calls in lockstep (the recovery decode of a batch: every stream conceals frame k in the same call), each with its own FEC / DRED feature
 * queue, predictor state and PCM history.
 *   lpcnet_plc_init + load (dnn/lpcnet_plc.c:58-86)   opusb200_lpcnet_plc_create   plc_layers = the 6 LinearLayers of the generated PLCModel struct
 *                                                       (plc_data.h order: dense_in, dense_out, gru1 input / recurrent, gru2 input / recurrent),
 *                                                       fargan_layers = the 20 of FARGAN (as opusb200_fargan_create); sizes checked like init_plcmodel
 *   lpcnet_plc_reset (:44)                             opusb200_lpcnet_plc_reset
 *   lpcnet_plc_fec_add (:88) / _fec_clear (:98)        opusb200_lpcnet_plc_fec_add (features [n][20]; have [n]: 1 = vector, 0 = the NULL call = a gap,
 *                                                       < 0 = no call for this stream; have == NULL: vectors for all) / _fec_clear
 *   lpcnet_plc_update (:143)                           opusb200_lpcnet_plc_update   (pcm [n][160] int16: a good frame of every stream)
 *   lpcnet_plc_conceal (:156)                          opusb200_lpcnet_plc_conceal  (pcm [n][160] int16 out): compute_plc_pred / get_fec_or_pred /
 *                                                       the attenuation table / queue_features / fargan_cont / fargan_synthesize_int
 *   lpcnet_plc_load_model's pitch part (:75-86)        opusb200_lpcnet_plc_load_pitch_model (the PitchDNN layers, as opusb200_lpcnet_encoder_create)
 * PCM and state are bit-identical to the reference's lpcnet_plc.c (generic C branch of dnn/vec.h) given the same model and calls, up to the libm
 * note of the feature extraction above.  The first concealed frame after GOOD audio analyses the decoded history (:160-177: burg_cepstral_analysis,
 * lpcnet_compute_single_frame_features_float, the predictor over the frames not analysed yet); that needs the pitch model: without it such a call
 * returns OPUSB200_INVALID_STATE and changes nothing.  Concealment from a reset state (with or without queued features) never needs it. ---- */
typedef struct OpusB200LpcnetPlc OpusB200LpcnetPlc;
int opusb200_lpcnet_plc_create(OpusB200Ctx *ctx, const OpusB200LinearLayer *plc_layers, int n_plc_layers, const OpusB200LinearLayer *fargan_layers,
                               int n_fargan_layers, int su_bias, int n_streams, OpusB200LpcnetPlc **out);
void opusb200_lpcnet_plc_destroy(OpusB200LpcnetPlc *plc);
int opusb200_lpcnet_plc_reset(OpusB200LpcnetPlc *plc);
int opusb200_lpcnet_plc_load_pitch_model(OpusB200LpcnetPlc *plc, const OpusB200LinearLayer *linear, int n_linear, const OpusB200Conv2dLayer *conv,
                                         int n_conv, int su_bias);
int opusb200_lpcnet_plc_fec_add(OpusB200LpcnetPlc *plc, const float *features, const int32_t *have, int on_device);
int opusb200_lpcnet_plc_fec_clear(OpusB200LpcnetPlc *plc);
int opusb200_lpcnet_plc_update(OpusB200LpcnetPlc *plc, const int16_t *pcm, int on_device);
int opusb200_lpcnet_plc_conceal(OpusB200LpcnetPlc *plc, int16_t *pcm, int on_device);
/* parity aid: features [n][20], cont_features [n][100], gru1 / gru2 [n][128], counters [n][4] = fec_read_pos, fec_fill_pos, fec_skip, loss_count */
int opusb200_lpcnet_plc_get_state(OpusB200LpcnetPlc *plc, float *features, float *cont_features, float *gru1, float *gru2, int32_t *counters);

/* ---- DRED decoder network (dnn/dred_rdovae_dec.h:50-52, dnn/dred_rdovae_dec.c:44-145) -------------------------------------------
 * layers: the 29 LinearLayers of the generated `RDOVAEDec` struct (dred_rdovae_dec_data.h) in declaration order (init_rdovaedec).
 * opusb200_dred_decode_all = DRED_rdovae_decode_all(model, features, state, latents, nb_latents, arch) for n streams at once:
 * state [n][DRED_STATE_DIM], latents [n][nb_latents][DRED_LATENT_DIM+1] -> features [n][nb_latents][4*DRED_NUM_FEATURES]. */
typedef struct OpusB200DredDec OpusB200DredDec;
int opusb200_dred_decoder_create(OpusB200Ctx *ctx, const OpusB200LinearLayer *layers, int n_layers, int su_bias, int n_streams, OpusB200DredDec **out);
void opusb200_dred_decoder_destroy(OpusB200DredDec *d);
int opusb200_dred_decode_all(OpusB200DredDec *d, float *features, const float *state, const float *latents, int nb_latents, int on_device);

/* Pinned host memory helpers (so host-buffer calls can overlap and run at PCIe speed). */
void *opusb200_host_alloc(size_t bytes);
void opusb200_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
