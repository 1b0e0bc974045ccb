/* oracle/ref_harness.c — TEST INFRASTRUCTURE ONLY.
 * Linked into oracle/_ref/libopus_ref_{c,simd}.so next to the unmodified reference objects.
 * It holds OUR small helpers that drive the reference through its own public / internal
 * (non-static) entry points so tests and bench.py can call them in bulk through ctypes
 * without a Python loop per frame.  Nothing here re-implements codec logic. */
#include <stdlib.h>
#include <string.h>
#include "opus.h"

/* Which oracle flavour is this?  (0 = pure C, 1 = RTCD/SIMD) */
int refh_is_simd(void) {
#ifdef OPUS_HAVE_RTCD
  return 1;
#else
  return 0;
#endif
}

/* Encode `nframes` consecutive frames of one stream (interleaved int16 PCM) with a fresh
 * encoder.  Packets are written back-to-back into `out` (stride `max_bytes`), sizes into
 * `sizes`, the encoder's final range per packet into `ranges`.
 * application: OPUS_APPLICATION_* ; returns 0 or a negative opus error. */
int refh_encode_stream(int Fs, int channels, int application, int bitrate, int vbr, int complexity,
                       int frame_size, int nframes, const opus_int16 *pcm, unsigned char *out,
                       int max_bytes, int *sizes, unsigned int *ranges) {
  int err = 0, f;
  OpusEncoder *enc = opus_encoder_create(Fs, channels, application, &err);
  if (!enc) return err;
  opus_encoder_ctl(enc, OPUS_SET_BITRATE(bitrate));
  opus_encoder_ctl(enc, OPUS_SET_VBR(vbr));
  opus_encoder_ctl(enc, OPUS_SET_COMPLEXITY(complexity));
  for (f = 0; f < nframes; f++) {
    int n = opus_encode(enc, pcm + (size_t)f * frame_size * channels, frame_size,
                        out + (size_t)f * max_bytes, max_bytes);
    if (n < 0) { opus_encoder_destroy(enc); return n; }
    sizes[f] = n;
    opus_encoder_ctl(enc, OPUS_GET_FINAL_RANGE(&ranges[f]));
  }
  opus_encoder_destroy(enc);
  return 0;
}

/* As refh_encode_stream, plus a list of (request, value) integer "set" ctls applied after create
 * (e.g. OPUS_SET_BANDWIDTH_REQUEST=4008, OPUS_SET_FORCE_CHANNELS_REQUEST=4022,
 * OPUS_SET_PACKET_LOSS_PERC_REQUEST=4014 ...). */
int refh_encode_stream_ex(int Fs, int channels, int application, int bitrate, int vbr, int complexity,
                          int frame_size, int nframes, const opus_int16 *pcm, unsigned char *out,
                          int max_bytes, int *sizes, unsigned int *ranges, const int *ctls, int nctl) {
  int err = 0, f, k;
  OpusEncoder *enc = opus_encoder_create(Fs, channels, application, &err);
  if (!enc) return err;
  opus_encoder_ctl(enc, OPUS_SET_BITRATE(bitrate));
  opus_encoder_ctl(enc, OPUS_SET_VBR(vbr));
  opus_encoder_ctl(enc, OPUS_SET_COMPLEXITY(complexity));
  for (k = 0; k < nctl; k++) {
    err = opus_encoder_ctl(enc, ctls[2 * k], (opus_int32)ctls[2 * k + 1]);
    if (err != OPUS_OK) { opus_encoder_destroy(enc); return err; }
  }
  for (f = 0; f < nframes; f++) {
    int n = opus_encode(enc, pcm + (size_t)f * frame_size * channels, frame_size,
                        out + (size_t)f * max_bytes, max_bytes);
    if (n < 0) { opus_encoder_destroy(enc); return n; }
    sizes[f] = n;
    opus_encoder_ctl(enc, OPUS_GET_FINAL_RANGE(&ranges[f]));
  }
  opus_encoder_destroy(enc);
  return 0;
}

/* Decode `nframes` consecutive packets of one stream with a fresh decoder. */
int refh_decode_stream(int Fs, int channels, int frame_size, int nframes, const unsigned char *pkts,
                       int stride, const int *sizes, opus_int16 *pcm, unsigned int *ranges) {
  int err = 0, f;
  OpusDecoder *dec = opus_decoder_create(Fs, channels, &err);
  if (!dec) return err;
  for (f = 0; f < nframes; f++) {
    int n = opus_decode(dec, sizes[f] ? pkts + (size_t)f * stride : NULL, sizes[f],
                        pcm + (size_t)f * frame_size * channels, frame_size, 0);
    if (n < 0) { opus_decoder_destroy(dec); return n; }
    opus_decoder_ctl(dec, OPUS_GET_FINAL_RANGE(&ranges[f]));
  }
  opus_decoder_destroy(dec);
  return 0;
}

/* Same, float output (no int16 rounding / soft clip applied by opus_decode_float). */
int refh_decode_stream_float(int Fs, int channels, int frame_size, int nframes, const unsigned char *pkts,
                             int stride, const int *sizes, float *pcm, unsigned int *ranges) {
  int err = 0, f;
  OpusDecoder *dec = opus_decoder_create(Fs, channels, &err);
  if (!dec) return err;
  for (f = 0; f < nframes; f++) {
    int n = opus_decode_float(dec, sizes[f] ? pkts + (size_t)f * stride : NULL, sizes[f],
                              pcm + (size_t)f * frame_size * channels, frame_size, 0);
    if (n < 0) { opus_decoder_destroy(dec); return n; }
    opus_decoder_ctl(dec, OPUS_GET_FINAL_RANGE(&ranges[f]));
  }
  opus_decoder_destroy(dec);
  return 0;
}

/* Float decode with integer "set" ctls applied after create (request, value pairs), reporting OPUS_GET_PITCH and
 * OPUS_GET_LAST_PACKET_DURATION after every packet. */
int refh_decode_stream_float_ctl(int Fs, int channels, int frame_size, int nframes, const unsigned char *pkts, int stride, const int *sizes,
                                 float *pcm, unsigned int *ranges, int nctl, const int *ctls, int *pitch, int *duration) {
  int err = 0, f, i;
  OpusDecoder *dec = opus_decoder_create(Fs, channels, &err);
  if (!dec) return err;
  for (i = 0; i < nctl; i++)
    if (opus_decoder_ctl(dec, ctls[2 * i], ctls[2 * i + 1]) != OPUS_OK) { opus_decoder_destroy(dec); return OPUS_BAD_ARG; }
  for (f = 0; f < nframes; f++) {
    int n = opus_decode_float(dec, sizes[f] ? pkts + (size_t)f * stride : NULL, sizes[f], pcm + (size_t)f * frame_size * channels, frame_size, 0);
    if (n < 0) { opus_decoder_destroy(dec); return n; }
    opus_decoder_ctl(dec, OPUS_GET_FINAL_RANGE(&ranges[f]));
    opus_decoder_ctl(dec, OPUS_GET_PITCH(&pitch[f]));
    opus_decoder_ctl(dec, OPUS_GET_LAST_PACKET_DURATION(&duration[f]));
  }
  opus_decoder_destroy(dec);
  return 0;
}

/* Throughput harness for the CPU baseline (bench.py): `n` decoders are created once, then every
 * stream decodes its T packets `reps` times (state simply continues, as in the GPU bench).
 * pkts: [n][T][stride].  Returns the number of frames decoded, or a negative opus error. */
long refh_decode_bench(int Fs, int channels, int frame_size, int n, int T, const unsigned char *pkts, int stride,
                       const int *sizes, int reps, opus_int16 *scratch_pcm) {
  int err = 0, s, t, r;
  long frames = 0;
  OpusDecoder **dec = (OpusDecoder **)calloc(n, sizeof(*dec));
  if (!dec) return OPUS_ALLOC_FAIL;
  for (s = 0; s < n; s++) {
    dec[s] = opus_decoder_create(Fs, channels, &err);
    if (!dec[s]) { frames = err; goto done; }
  }
  for (r = 0; r < reps; r++)
    for (t = 0; t < T; t++)
      for (s = 0; s < n; s++) {
        int k = opus_decode(dec[s], pkts + ((size_t)s * T + t) * stride, sizes[s * T + t], scratch_pcm, frame_size, 0);
        if (k < 0) { frames = k; goto done; }
        frames++;
      }
done:
  for (s = 0; s < n; s++) if (dec[s]) opus_decoder_destroy(dec[s]);
  free(dec);
  return frames;
}

/* ---- kernel-level entry points of the reference, for the B3-level parity tests --------------
 * These wrappers only fetch the reference's static mode (mdct lookup, window) and call the
 * reference's own non-static functions. */
#include "celt/modes.h"
#include "celt/mdct.h"
#include "celt/pitch.h"
#include "celt/vq.h"
#include "celt/bands.h"

static const OpusCustomMode *refh_mode(void) { return opus_custom_mode_create(48000, 960, NULL); }

/* clt_mdct_forward on one block set: in has B*N+overlap samples (N = 120<<LM if !transient, else 120 with
 * B = 1<<LM blocks), out has B*N coefficients (interleaved by B).  celt/celt_encoder.c:511-554 */
void refh_mdct_forward(const float *in, float *out, int LM, int transient) {
  const OpusCustomMode *m = refh_mode();
  int B = transient ? 1 << LM : 1, N = transient ? 120 : 120 << LM, shift = transient ? 3 : 3 - LM, b;
  float tmp[960 + 120 + 8];
  memcpy(tmp, in, sizeof(float) * (B * N + 120));
  for (b = 0; b < B; b++) clt_mdct_forward_c(&m->mdct, tmp + b * N, &out[b], m->window, 120, shift, B, 0);
}
void refh_mdct_backward(const float *in, float *out, int LM, int transient) {
  const OpusCustomMode *m = refh_mode();
  int B = transient ? 1 << LM : 1, N = transient ? 120 : 120 << LM, shift = transient ? 3 : 3 - LM, b;
  float tmp[960 + 8];
  memcpy(tmp, in, sizeof(float) * B * N);
  for (b = 0; b < B; b++) clt_mdct_backward_c(&m->mdct, &tmp[b], out + N * b, m->window, 120, shift, B, 0);
}
void refh_pitch_xcorr(const float *x, const float *y, float *xcorr, int len, int max_pitch) {
  celt_pitch_xcorr_c(x, y, xcorr, len, max_pitch, 0);
}
void refh_exp_rotation(float *X, int len, int dir, int stride, int K, int spread) { exp_rotation(X, len, dir, stride, K, spread); }
float refh_pvq_search(float *X, int *iy, int K, int N) { return op_pvq_search_c(X, iy, K, N, 0); }

/* ---- SILK kernel-level entry points ---- */
#include "silk/SigProc_FIX.h"
#include "silk/resampler_structs.h"
int refh_resampler_size(void) { return (int)sizeof(silk_resampler_state_struct); }

/* ---- SILK noise-shaping quantizer (silk_NSQ / silk_NSQ_del_dec through the reference's own dispatch) ---- */
#include "silk/main.h"
#include "celt/cpu_support.h"
typedef struct { opus_int32 fs_khz, nb_subfr, predictLPCOrder, shapingLPCOrder, warping_Q16, nStates; } RefhNsqCfg;
typedef struct {
  opus_int32 signalType, quantOffsetType, NLSFInterpCoef_Q2, Seed;
  opus_int32 HarmShapeGain_Q14[4], Tilt_Q14[4], LF_shp_Q14[4], Gains_Q16[4], pitchL[4];
  opus_int32 Lambda_Q10, LTP_scale_Q14;
  opus_int16 PredCoef_Q12[32], LTPCoef_Q14[20], AR_Q13[96];
} RefhNsqFrame;
int refh_nsq_state_size(void) { return (int)sizeof(silk_nsq_state); }
void refh_nsq_state_init(silk_nsq_state *S) {
  memset(S, 0, sizeof(*S));
  S->prev_gain_Q16 = 65536;
  S->lagPrev = 100;
}
void refh_silk_nsq(const RefhNsqCfg *c, silk_nsq_state *S, RefhNsqFrame *f, const opus_int16 *x16, opus_int8 *pulses) {
  static silk_encoder_state enc; /* only the fields the quantizers read are set */
  SideInfoIndices idx;
  memset(&enc, 0, sizeof(enc));
  memset(&idx, 0, sizeof(idx));
  enc.fs_kHz = c->fs_khz;
  enc.nb_subfr = c->nb_subfr;
  enc.subfr_length = 5 * c->fs_khz;
  enc.frame_length = enc.subfr_length * c->nb_subfr;
  enc.ltp_mem_length = 20 * c->fs_khz;
  enc.predictLPCOrder = c->predictLPCOrder;
  enc.shapingLPCOrder = c->shapingLPCOrder;
  enc.warping_Q16 = c->warping_Q16;
  enc.nStatesDelayedDecision = c->nStates;
  enc.arch = opus_select_arch();
  idx.signalType = (opus_int8)f->signalType;
  idx.quantOffsetType = (opus_int8)f->quantOffsetType;
  idx.NLSFInterpCoef_Q2 = (opus_int8)f->NLSFInterpCoef_Q2;
  idx.Seed = (opus_int8)f->Seed;
  if (c->nStates > 1 || c->warping_Q16 > 0)
    silk_NSQ_del_dec(&enc, S, &idx, x16, pulses, f->PredCoef_Q12, f->LTPCoef_Q14, f->AR_Q13, f->HarmShapeGain_Q14, f->Tilt_Q14,
                     f->LF_shp_Q14, f->Gains_Q16, f->pitchL, f->Lambda_Q10, f->LTP_scale_Q14, enc.arch);
  else
    silk_NSQ(&enc, S, &idx, x16, pulses, f->PredCoef_Q12, f->LTPCoef_Q14, f->AR_Q13, f->HarmShapeGain_Q14, f->Tilt_Q14, f->LF_shp_Q14,
             f->Gains_Q16, f->pitchL, f->Lambda_Q10, f->LTP_scale_Q14, enc.arch);
  f->Seed = idx.Seed;
}
void refh_nsq_state_export(const silk_nsq_state *S, opus_int32 *o) {
  int i;
  for (i = 0; i < 640; i++) *o++ = S->xq[i];
  for (i = 0; i < 640; i++) *o++ = S->sLTP_shp_Q14[i];
  for (i = 0; i < 16; i++) *o++ = S->sLPC_Q14[i];
  for (i = 0; i < 24; i++) *o++ = S->sAR2_Q14[i];
  *o++ = S->sLF_AR_shp_Q14; *o++ = S->sDiff_shp_Q14; *o++ = S->lagPrev; *o++ = S->rand_seed; *o++ = S->prev_gain_Q16;
}

/* ---- dnn primitives (pure-C library only: the RTCD build would need the dnn/x86 dispatch tables) ---- */
#ifndef OPUS_HAVE_RTCD
#include "dnn/nnet.h"
void compute_linear_c(const LinearLayer *linear, float *out, const float *in);
void compute_activation_c(float *output, const float *input, int N, int activation);
void compute_conv2d_c(const Conv2dLayer *conv, float *out, float *mem, const float *in, int height, int hstride, int activation);
void refh_nnet_linear(float *out, const float *in, const float *bias, const float *subias, const opus_int8 *weights, const float *float_weights,
                      const int *idx, const float *diag, const float *scale, int nb_inputs, int nb_outputs) {
  LinearLayer l;
  l.bias = bias; l.subias = subias; l.weights = weights; l.float_weights = float_weights; l.weights_idx = idx; l.diag = diag;
  l.scale = scale; l.nb_inputs = nb_inputs; l.nb_outputs = nb_outputs;
  compute_linear_c(&l, out, in);
}
void refh_nnet_activation(float *out, const float *in, int N, int act) { compute_activation_c(out, in, N, act); }
void refh_nnet_conv2d(float *out, float *mem, const float *in, const float *bias, const float *w, int in_ch, int out_ch, int ktime,
                      int kheight, int height, int hstride, int act) {
  Conv2dLayer c;
  c.bias = bias; c.float_weights = w; c.in_channels = in_ch; c.out_channels = out_ch; c.ktime = ktime; c.kheight = kheight;
  compute_conv2d_c(&c, out, mem, in, height, hstride, act);
}
#endif

/* ---- encoder front end: the reference's celt_preemphasis / clt_mdct_forward (as compute_mdcts drives it) /
 * compute_band_energies / amp2Log2 / normalise_bands on one frame.  mem: [CC] pre-emphasis memories,
 * hist: [CC][120] previous frame tails (both in/out). ---- */
#include "celt/celt.h"
#include "celt/quant_bands.h"
void refh_celt_enc_analysis(float *mem, float *hist, const float *pcm, int CC, int LM, int shortBlocks, float *X, float *bandE, float *bandLogE) {
  const OpusCustomMode *m = refh_mode();
  const int N = 120 << LM, overlap = 120;
  float in[2 * (960 + 120)], freq[2 * 960];
  int c;
  for (c = 0; c < CC; c++) {
    memcpy(in + c * (N + overlap), hist + c * overlap, overlap * sizeof(float));
    celt_preemphasis(pcm + c, in + c * (N + overlap) + overlap, N, CC, 1, m->preemph, mem + c, 0);
    memcpy(hist + c * overlap, in + c * (N + overlap) + N, overlap * sizeof(float));
    refh_mdct_forward(in + c * (N + overlap), freq + c * N, LM, shortBlocks);
  }
  compute_band_energies(m, freq, bandE, 21, CC, LM, 0);
  amp2Log2(m, 21, 21, bandE, bandLogE, CC);
  normalise_bands(m, freq, X, bandE, 21, CC, 1 << LM);
}

/* ---- the reference CELT encoder driven directly (celt_encoder_init + celt_encode_with_ec), bypassing the Opus-layer glue ---- */
#include "celt/celt.h"
int refh_celt_encoder_size(int channels) { return celt_encoder_get_size(channels); }
int refh_celt_encoder_init(CELTEncoder *st, int channels, int complexity, int bitrate, int vbr, int lsb_depth) {
  int ret = celt_encoder_init(st, 48000, channels, opus_select_arch());
  if (ret != OPUS_OK) return ret;
  celt_encoder_ctl(st, CELT_SET_SIGNALLING(0));
  celt_encoder_ctl(st, OPUS_SET_COMPLEXITY(complexity));
  celt_encoder_ctl(st, OPUS_SET_BITRATE(bitrate));
  celt_encoder_ctl(st, OPUS_SET_VBR(vbr));
  celt_encoder_ctl(st, OPUS_SET_LSB_DEPTH(lsb_depth));
  return OPUS_OK;
}
/* pcm: interleaved floats +-1.0; returns bytes written (or < 0); *rng = final range */
int refh_celt_encode(CELTEncoder *st, const float *pcm, int frame_size, unsigned char *out, int max_bytes, unsigned int *rng) {
  int n = celt_encode_with_ec(st, pcm, frame_size, out, max_bytes, NULL);
  celt_encoder_ctl(st, OPUS_GET_FINAL_RANGE(rng));
  return n;
}
void refh_celt_encoder_set_cvbr(CELTEncoder *st, int v) { celt_encoder_ctl(st, OPUS_SET_VBR_CONSTRAINT(v)); }
/* as refh_celt_encoder_init for an input rate Fs in {48000, 24000, 16000, 12000, 8000} (st->upsample = 48000 / Fs) */
int refh_celt_encoder_init_fs(CELTEncoder *st, int Fs, int channels, int complexity, int bitrate, int vbr, int lsb_depth) {
  int ret = celt_encoder_init(st, Fs, channels, opus_select_arch());
  if (ret != OPUS_OK) return ret;
  celt_encoder_ctl(st, CELT_SET_SIGNALLING(0));
  celt_encoder_ctl(st, OPUS_SET_COMPLEXITY(complexity));
  celt_encoder_ctl(st, OPUS_SET_BITRATE(bitrate));
  celt_encoder_ctl(st, OPUS_SET_VBR(vbr));
  celt_encoder_ctl(st, OPUS_SET_LSB_DEPTH(lsb_depth));
  return OPUS_OK;
}
int refh_celt_encoder_set_prediction(CELTEncoder *st, int v) { return celt_encoder_ctl(st, CELT_SET_PREDICTION(v)); }
int refh_celt_encoder_set_channels(CELTEncoder *st, int c) { return celt_encoder_ctl(st, CELT_SET_CHANNELS(c)); }
int refh_celt_encoder_set_end_band(CELTEncoder *st, int end) { return celt_encoder_ctl(st, CELT_SET_END_BAND(end)); }

/* ---- the Opus layer's tonality / music analysis alone (src/analysis.c run_analysis), one stream, consecutive frames.
 * out: nframes records of 11 floats + 19 bytes = our OrcAnalysisInfo layout is NOT assumed: fields are written one by one:
 * [valid, tonality, tonality_slope, noisiness, activity, music_prob, music_prob_min, music_prob_max, bandwidth,
 *  activity_probability, max_pitch_ratio, leak_boost[0..18]] as 30 floats per frame. ---- */
#include "src/analysis.h"
#include "src/opus_private.h"
int refh_run_analysis_stream(int Fs, int channels, int frame_size, int nframes, const void *pcm, int is_int16, int lsb_depth, float *out) {
  TonalityAnalysisState *t = (TonalityAnalysisState *)calloc(1, sizeof(*t));
  const OpusCustomMode *m = refh_mode();
  int f, k;
  if (!t) return -7;
  tonality_analysis_init(t, Fs);
  for (f = 0; f < nframes; f++) {
    AnalysisInfo info;
    const void *p = is_int16 ? (const void *)((const opus_int16 *)pcm + (size_t)f * frame_size * channels)
                             : (const void *)((const float *)pcm + (size_t)f * frame_size * channels);
    memset(&info, 0, sizeof(info));
    run_analysis(t, m, p, frame_size, frame_size, 0, -2, channels, Fs, lsb_depth, is_int16 ? downmix_int : downmix_float, &info);
    float *o = out + 30 * f;
    o[0] = (float)info.valid; o[1] = info.tonality; o[2] = info.tonality_slope; o[3] = info.noisiness; o[4] = info.activity;
    o[5] = info.music_prob; o[6] = info.music_prob_min; o[7] = info.music_prob_max; o[8] = (float)info.bandwidth;
    o[9] = info.activity_probability; o[10] = info.max_pitch_ratio;
    for (k = 0; k < 19; k++) o[11 + k] = (float)info.leak_boost[k];
  }
  free(t);
  return 0;
}

/* As refh_encode_stream_ex through opus_encode_float (pcm: interleaved floats +-1.0) */
int refh_encode_stream_float_ex(int Fs, int channels, int application, int bitrate, int vbr, int complexity,
                                int frame_size, int nframes, const float *pcm, unsigned char *out,
                                int max_bytes, int *sizes, unsigned int *ranges, const int *ctls, int nctl) {
  int err = 0, f, k;
  OpusEncoder *enc = opus_encoder_create(Fs, channels, application, &err);
  if (!enc) return err;
  opus_encoder_ctl(enc, OPUS_SET_BITRATE(bitrate));
  opus_encoder_ctl(enc, OPUS_SET_VBR(vbr));
  opus_encoder_ctl(enc, OPUS_SET_COMPLEXITY(complexity));
  for (k = 0; k < nctl; k++) {
    err = opus_encoder_ctl(enc, ctls[2 * k], (opus_int32)ctls[2 * k + 1]);
    if (err != OPUS_OK) { opus_encoder_destroy(enc); return err; }
  }
  for (f = 0; f < nframes; f++) {
    int n = opus_encode_float(enc, pcm + (size_t)f * frame_size * channels, frame_size, out + (size_t)f * max_bytes, max_bytes);
    if (n < 0) { opus_encoder_destroy(enc); return n; }
    sizes[f] = n;
    opus_encoder_ctl(enc, OPUS_GET_FINAL_RANGE(&ranges[f]));
  }
  opus_encoder_destroy(enc);
  return 0;
}

/* Encode throughput sample for bench.py: n independent streams (one encoder each, created outside the timing by the caller's
 * warm-up call pattern: creation is cheap next to T frames), T consecutive frames per stream, `reps` passes over the same PCM with the
 * encoder state carried.  pcm: [n][T][frame_size*channels] int16.  Returns the number of frames encoded, or a negative error. */
long refh_encode_bench(int Fs, int channels, int application, int bitrate, int vbr, int complexity, int frame_size, int n, int T,
                       const opus_int16 *pcm, int reps, int max_bytes, long *bytes_out) {
  long frames = 0, bytes = 0;
  int s, f, r, err = 0;
  unsigned char *buf = (unsigned char *)malloc(max_bytes);
  OpusEncoder **enc = (OpusEncoder **)calloc(n, sizeof(*enc));
  if (!buf || !enc) return -7;
  for (s = 0; s < n; s++) {
    enc[s] = opus_encoder_create(Fs, channels, application, &err);
    if (!enc[s]) return err;
    opus_encoder_ctl(enc[s], OPUS_SET_BITRATE(bitrate));
    opus_encoder_ctl(enc[s], OPUS_SET_VBR(vbr));
    opus_encoder_ctl(enc[s], OPUS_SET_COMPLEXITY(complexity));
  }
  for (r = 0; r < reps; r++)
    for (f = 0; f < T; f++)   /* frame-major, like a batch server that advances every stream by one frame per step */
      for (s = 0; s < n; s++) {
        int nb = opus_encode(enc[s], pcm + ((size_t)s * T + f) * frame_size * channels, frame_size, buf, max_bytes);
        if (nb < 0) return nb;
        bytes += nb;
        frames++;
      }
  for (s = 0; s < n; s++) opus_encoder_destroy(enc[s]);
  free(enc);
  free(buf);
  if (bytes_out) *bytes_out = bytes;
  return frames;
}

/* ---- SILK excitation coder (silk/encode_pulses.c, shell_coder.c, code_signs.c) on a reference range encoder ---- */
#include "celt/entenc.h"
typedef struct { ec_enc enc; unsigned char buf[1280]; } RefhSilkCoder;
int refh_silk_coder_size(void) { return (int)sizeof(RefhSilkCoder); }
void refh_silk_coder_init(RefhSilkCoder *c, int storage) { ec_enc_init(&c->enc, c->buf, (opus_uint32)storage); }
void refh_silk_coder_icdf(RefhSilkCoder *c, int s, const unsigned char *icdf, int ftb) { ec_enc_icdf(&c->enc, s, icdf, (unsigned)ftb); }
opus_uint32 refh_silk_coder_tell_frac(RefhSilkCoder *c) { return ec_tell_frac(&c->enc); }
void refh_silk_encode_pulses(RefhSilkCoder *c, int signalType, int quantOffsetType, opus_int8 *pulses, int frame_length) {
  silk_encode_pulses(&c->enc, signalType, quantOffsetType, pulses, frame_length);
}
opus_uint32 refh_silk_coder_done(RefhSilkCoder *c, unsigned char *out) {
  ec_enc_done(&c->enc);
  memcpy(out, c->buf, c->enc.storage);
  return c->enc.rng;
}
/* silk_encode_indices on a scratch encoder state that holds only what the function reads (silk/encode_indices.c:35-185) */
#include "silk/tables.h"
void refh_silk_encode_indices(RefhSilkCoder *c, int fs_kHz, int nb_subfr, const SideInfoIndices *ix, int condCoding, int *prev) {
  static silk_encoder_state st;
  memset(&st, 0, sizeof(st));
  st.fs_kHz = fs_kHz; st.nb_subfr = nb_subfr;
  st.psNLSF_CB = fs_kHz == 16 ? &silk_NLSF_CB_WB : &silk_NLSF_CB_NB_MB;
  st.predictLPCOrder = fs_kHz == 16 ? 16 : 10;
  st.pitch_lag_low_bits_iCDF = fs_kHz == 16 ? silk_uniform8_iCDF : fs_kHz == 12 ? silk_uniform6_iCDF : silk_uniform4_iCDF;
  st.pitch_contour_iCDF = fs_kHz == 8 ? (nb_subfr == 4 ? silk_pitch_contour_NB_iCDF : silk_pitch_contour_10_ms_NB_iCDF)
                                      : (nb_subfr == 4 ? silk_pitch_contour_iCDF : silk_pitch_contour_10_ms_iCDF);
  st.indices = *ix;
  st.ec_prevSignalType = prev[0]; st.ec_prevLagIndex = (opus_int16)prev[1];
  silk_encode_indices(&st, &c->enc, 0, 0, condCoding);
  prev[0] = st.ec_prevSignalType; prev[1] = st.ec_prevLagIndex;
}
int refh_silk_indices_size(void) { return (int)sizeof(SideInfoIndices); }

/* ---- silk_process_gains_FLP on scratch encoder structs that hold what the function reads (silk/float/process_gains_FLP.c:35-103) ---- */
#include "silk/float/main_FLP.h"
typedef struct {
  float Gains[4], ResNrg[4];
  float LTPredCodGain, input_quality, coding_quality, Lambda;
  opus_int32 GainsUnq_Q16[4];
  opus_int32 signalType, quantOffsetType, SNR_dB_Q7, input_tilt_Q15, speech_activity_Q8, condCoding;
  opus_int8 GainsIndices[4];
  opus_int8 LastGainIndex, lastGainIndexPrev, pad[2];
} RefhSilkGains;   /* == OrcSilkGains */
void refh_silk_process_gains(RefhSilkGains *g, int nb_subfr, int subfr_length, int nStatesDelayedDecision) {
  static silk_encoder_state_FLP st;
  static silk_encoder_control_FLP ctl;
  int k;
  memset(&st, 0, sizeof(st));
  memset(&ctl, 0, sizeof(ctl));
  st.sCmn.nb_subfr = nb_subfr; st.sCmn.subfr_length = subfr_length; st.sCmn.nStatesDelayedDecision = nStatesDelayedDecision;
  st.sCmn.indices.signalType = (opus_int8)g->signalType; st.sCmn.indices.quantOffsetType = (opus_int8)g->quantOffsetType;
  st.sCmn.SNR_dB_Q7 = g->SNR_dB_Q7; st.sCmn.input_tilt_Q15 = g->input_tilt_Q15; st.sCmn.speech_activity_Q8 = g->speech_activity_Q8;
  st.sShape.LastGainIndex = g->LastGainIndex;
  for (k = 0; k < 4; k++) { ctl.Gains[k] = g->Gains[k]; ctl.ResNrg[k] = g->ResNrg[k]; }
  ctl.LTPredCodGain = g->LTPredCodGain; ctl.input_quality = g->input_quality; ctl.coding_quality = g->coding_quality;
  silk_process_gains_FLP(&st, &ctl, g->condCoding);
  for (k = 0; k < 4; k++) { g->Gains[k] = ctl.Gains[k]; g->GainsUnq_Q16[k] = ctl.GainsUnq_Q16[k]; g->GainsIndices[k] = st.sCmn.indices.GainsIndices[k]; }
  g->Lambda = ctl.Lambda; g->quantOffsetType = st.sCmn.indices.quantOffsetType; g->LastGainIndex = st.sShape.LastGainIndex;
  g->lastGainIndexPrev = ctl.lastGainIndexPrev;
}
opus_int32 refh_silk_lin2log(opus_int32 x) { return silk_lin2log(x); }
opus_int32 refh_silk_log2lin(opus_int32 x) { return silk_log2lin(x); }

/* ---- silk_process_NLSFs on a scratch encoder state that holds what the function reads (silk/process_NLSFs.c:35-107) ---- */
typedef struct {
  opus_int16 NLSF_Q15[16];
  opus_int16 prev_NLSFq_Q15[16];
  opus_int16 PredCoef_Q12[2][16];
  opus_int8 NLSFIndices[17];
  opus_int8 NLSFInterpCoef_Q2, signalType, useInterpolatedNLSFs;
  opus_int32 speech_activity_Q8, NLSF_MSVQ_Survivors;
} RefhSilkNlsf;   /* == OrcSilkNlsf */
void refh_silk_process_NLSFs(RefhSilkNlsf *f, int wb, int nb_subfr) {
  static silk_encoder_state st;
  memset(&st, 0, sizeof(st));
  st.psNLSF_CB = wb ? &silk_NLSF_CB_WB : &silk_NLSF_CB_NB_MB;
  st.predictLPCOrder = wb ? 16 : 10;
  st.nb_subfr = nb_subfr;
  st.speech_activity_Q8 = f->speech_activity_Q8;
  st.useInterpolatedNLSFs = f->useInterpolatedNLSFs;
  st.NLSF_MSVQ_Survivors = f->NLSF_MSVQ_Survivors;
  st.indices.NLSFInterpCoef_Q2 = f->NLSFInterpCoef_Q2;
  st.indices.signalType = f->signalType;
  silk_process_NLSFs(&st, f->PredCoef_Q12, f->NLSF_Q15, f->prev_NLSFq_Q15);
  memcpy(f->NLSFIndices, st.indices.NLSFIndices, 17);
}
void refh_silk_NLSF2A(opus_int16 *a_Q12, const opus_int16 *NLSF, int d) { silk_NLSF2A(a_Q12, NLSF, d, 0); }
void refh_silk_NLSF_stabilize(opus_int16 *NLSF_Q15, int wb) {
  const silk_NLSF_CB_struct *cb = wb ? &silk_NLSF_CB_WB : &silk_NLSF_CB_NB_MB;
  silk_NLSF_stabilize(NLSF_Q15, cb->deltaMin_Q15, cb->order);
}
opus_int32 refh_silk_LPC_inverse_pred_gain(const opus_int16 *A_Q12, int order) { return silk_LPC_inverse_pred_gain_c(A_Q12, order); }

/* ---- silk_find_LPC_FLP on a scratch encoder state (silk/float/find_LPC_FLP.c:36-105) ---- */
typedef struct {
  opus_int16 NLSF_Q15[16];
  opus_int16 prev_NLSFq_Q15[16];
  float minInvGain;
  opus_int32 useInterpolatedNLSFs, first_frame_after_reset;
  opus_int32 NLSFInterpCoef_Q2;
} RefhSilkFindLpc;   /* == OrcSilkFindLpc */
void refh_silk_find_LPC(RefhSilkFindLpc *f, const float *x, int order, int nb_subfr, int subfr_length) {
  static silk_encoder_state st;
  memset(&st, 0, sizeof(st));
  st.predictLPCOrder = order; st.nb_subfr = nb_subfr; st.subfr_length = subfr_length;
  st.useInterpolatedNLSFs = f->useInterpolatedNLSFs; st.first_frame_after_reset = f->first_frame_after_reset;
  memcpy(st.prev_NLSFq_Q15, f->prev_NLSFq_Q15, sizeof(st.prev_NLSFq_Q15));
  silk_find_LPC_FLP(&st, f->NLSF_Q15, x, f->minInvGain, 0);
  f->NLSFInterpCoef_Q2 = st.indices.NLSFInterpCoef_Q2;
}
void refh_silk_A2NLSF(opus_int16 *NLSF, opus_int32 *a_Q16, int d) { silk_A2NLSF(NLSF, a_Q16, d); }

/* ---- silk_find_pred_coefs_FLP on scratch encoder structs (silk/float/find_pred_coefs_FLP.c:35-120) ---- */
typedef struct {
  float Gains[4];
  opus_int32 pitchL[4];
  float coding_quality;
  opus_int32 signalType, condCoding, first_frame_after_reset, useInterpolatedNLSFs, PacketLoss_perc, nFramesPerPacket, LBRR_flag, SNR_dB_Q7, speech_activity_Q8,
      NLSF_MSVQ_Survivors;
  opus_int16 prev_NLSFq_Q15[16];
  opus_int32 sum_log_gain_Q7;
  float LTPCoef[20], LTPredCodGain, LTP_scale, PredCoef[2][16], ResNrg[4];
  opus_int8 LTPIndex[4], PERIndex, LTP_scaleIndex, NLSFInterpCoef_Q2, pad;
  opus_int8 NLSFIndices[17], pad2[3];
} RefhSilkPredCoefs;   /* == OrcSilkPredCoefs */
void refh_silk_find_pred_coefs(RefhSilkPredCoefs *c, const float *res_pitch, const float *x, int order, int nb_subfr, int subfr_length) {
  static silk_encoder_state_FLP st;
  static silk_encoder_control_FLP ctl;
  int k;
  memset(&st, 0, sizeof(st));
  memset(&ctl, 0, sizeof(ctl));
  st.sCmn.nb_subfr = nb_subfr; st.sCmn.subfr_length = subfr_length; st.sCmn.predictLPCOrder = order; st.sCmn.ltp_mem_length = 4 * subfr_length;
  st.sCmn.psNLSF_CB = order == 16 ? &silk_NLSF_CB_WB : &silk_NLSF_CB_NB_MB;
  st.sCmn.indices.signalType = (opus_int8)c->signalType;
  st.sCmn.first_frame_after_reset = c->first_frame_after_reset; st.sCmn.useInterpolatedNLSFs = c->useInterpolatedNLSFs;
  st.sCmn.PacketLoss_perc = c->PacketLoss_perc; st.sCmn.nFramesPerPacket = c->nFramesPerPacket; st.sCmn.LBRR_flag = c->LBRR_flag;
  st.sCmn.SNR_dB_Q7 = c->SNR_dB_Q7; st.sCmn.speech_activity_Q8 = c->speech_activity_Q8; st.sCmn.NLSF_MSVQ_Survivors = c->NLSF_MSVQ_Survivors;
  st.sCmn.sum_log_gain_Q7 = c->sum_log_gain_Q7;
  memcpy(st.sCmn.prev_NLSFq_Q15, c->prev_NLSFq_Q15, sizeof(st.sCmn.prev_NLSFq_Q15));
  for (k = 0; k < 4; k++) { ctl.Gains[k] = c->Gains[k]; ctl.pitchL[k] = c->pitchL[k]; }
  ctl.coding_quality = c->coding_quality;
  silk_find_pred_coefs_FLP(&st, &ctl, res_pitch, x, c->condCoding);
  memcpy(c->prev_NLSFq_Q15, st.sCmn.prev_NLSFq_Q15, sizeof(c->prev_NLSFq_Q15));
  c->sum_log_gain_Q7 = st.sCmn.sum_log_gain_Q7;
  memcpy(c->LTPCoef, ctl.LTPCoef, sizeof(c->LTPCoef));
  c->LTPredCodGain = ctl.LTPredCodGain; c->LTP_scale = ctl.LTP_scale;
  memcpy(c->PredCoef, ctl.PredCoef, sizeof(c->PredCoef));
  memcpy(c->ResNrg, ctl.ResNrg, sizeof(c->ResNrg));
  memcpy(c->LTPIndex, st.sCmn.indices.LTPIndex, 4);
  c->PERIndex = st.sCmn.indices.PERIndex; c->LTP_scaleIndex = st.sCmn.indices.LTP_scaleIndex; c->NLSFInterpCoef_Q2 = st.sCmn.indices.NLSFInterpCoef_Q2;
  memcpy(c->NLSFIndices, st.sCmn.indices.NLSFIndices, 17);
}

/* ---- silk_noise_shape_analysis_FLP on scratch encoder structs (silk/float/noise_shape_analysis_FLP.c:138-350) ---- */
typedef struct {
  opus_int32 SNR_dB_Q7, input_quality_bands_Q15[2], useCBR, speech_activity_Q8, signalType, warping_Q16;
  float LTPCorr, predGain;
  opus_int32 pitchL[4];
  float HarmShapeGain_smth, Tilt_smth;
  float input_quality, coding_quality;
  opus_int32 quantOffsetType;
  float AR[4 * 24], Gains[4], LF_MA_shp[4], LF_AR_shp[4], HarmShapeGain[4], Tilt[4];
} RefhSilkNoiseShape;   /* == OrcSilkNoiseShape */
void refh_silk_noise_shape_analysis(RefhSilkNoiseShape *c, const float *pitch_res, const float *x, int fs_kHz, int nb_subfr, int shapingLPCOrder) {
  static silk_encoder_state_FLP st;
  static silk_encoder_control_FLP ctl;
  int k;
  memset(&st, 0, sizeof(st));
  memset(&ctl, 0, sizeof(ctl));
  st.sCmn.fs_kHz = fs_kHz; st.sCmn.nb_subfr = nb_subfr; st.sCmn.subfr_length = 5 * fs_kHz; st.sCmn.la_shape = 5 * fs_kHz;
  st.sCmn.shapeWinLength = 5 * fs_kHz + 2 * st.sCmn.la_shape; st.sCmn.shapingLPCOrder = shapingLPCOrder;
  st.sCmn.SNR_dB_Q7 = c->SNR_dB_Q7; st.sCmn.input_quality_bands_Q15[0] = c->input_quality_bands_Q15[0]; st.sCmn.input_quality_bands_Q15[1] = c->input_quality_bands_Q15[1];
  st.sCmn.useCBR = c->useCBR; st.sCmn.speech_activity_Q8 = c->speech_activity_Q8; st.sCmn.indices.signalType = (opus_int8)c->signalType;
  st.sCmn.warping_Q16 = c->warping_Q16; st.LTPCorr = c->LTPCorr; ctl.predGain = c->predGain;
  st.sShape.HarmShapeGain_smth = c->HarmShapeGain_smth; st.sShape.Tilt_smth = c->Tilt_smth;
  for (k = 0; k < 4; k++) ctl.pitchL[k] = c->pitchL[k];
  silk_noise_shape_analysis_FLP(&st, &ctl, pitch_res, x);
  c->HarmShapeGain_smth = st.sShape.HarmShapeGain_smth; c->Tilt_smth = st.sShape.Tilt_smth;
  c->input_quality = ctl.input_quality; c->coding_quality = ctl.coding_quality; c->quantOffsetType = st.sCmn.indices.quantOffsetType;
  memcpy(c->AR, ctl.AR, sizeof(c->AR)); memcpy(c->Gains, ctl.Gains, sizeof(c->Gains)); memcpy(c->LF_MA_shp, ctl.LF_MA_shp, sizeof(c->LF_MA_shp));
  memcpy(c->LF_AR_shp, ctl.LF_AR_shp, sizeof(c->LF_AR_shp)); memcpy(c->HarmShapeGain, ctl.HarmShapeGain, sizeof(c->HarmShapeGain)); memcpy(c->Tilt, ctl.Tilt, sizeof(c->Tilt));
}

/* ---- silk_find_pitch_lags_FLP on scratch encoder structs (silk/float/find_pitch_lags_FLP.c:36-132) ---- */
typedef struct {
  opus_int32 signalType, first_frame_after_reset, speech_activity_Q8, prevSignalType, input_tilt_Q15, prevLag, pitchEstimationThreshold_Q16;
  float LTPCorr, predGain;
  opus_int32 pitchL[4];
  opus_int32 lagIndex, contourIndex;
} RefhSilkPitch;   /* == OrcSilkPitch */
void refh_silk_find_pitch_lags(RefhSilkPitch *c, float *res, const float *x, int fs_kHz, int nb_subfr, int lpc_order, int complexity) {
  static silk_encoder_state_FLP st;
  static silk_encoder_control_FLP ctl;
  int k;
  memset(&st, 0, sizeof(st));
  memset(&ctl, 0, sizeof(ctl));
  st.sCmn.fs_kHz = fs_kHz; st.sCmn.nb_subfr = nb_subfr; st.sCmn.la_pitch = 2 * fs_kHz; st.sCmn.frame_length = 5 * nb_subfr * fs_kHz; st.sCmn.ltp_mem_length = 20 * fs_kHz;
  st.sCmn.pitch_LPC_win_length = (5 * nb_subfr + 4) * fs_kHz; st.sCmn.pitchEstimationLPCOrder = lpc_order; st.sCmn.pitchEstimationComplexity = complexity;
  st.sCmn.indices.signalType = (opus_int8)c->signalType; st.sCmn.first_frame_after_reset = c->first_frame_after_reset; st.sCmn.speech_activity_Q8 = c->speech_activity_Q8;
  st.sCmn.prevSignalType = c->prevSignalType; st.sCmn.input_tilt_Q15 = c->input_tilt_Q15; st.sCmn.prevLag = c->prevLag;
  st.sCmn.pitchEstimationThreshold_Q16 = c->pitchEstimationThreshold_Q16; st.LTPCorr = c->LTPCorr;
  silk_find_pitch_lags_FLP(&st, &ctl, res, x, 0);
  c->signalType = st.sCmn.indices.signalType; c->LTPCorr = st.LTPCorr; c->predGain = ctl.predGain;
  for (k = 0; k < 4; k++) c->pitchL[k] = ctl.pitchL[k];
  c->lagIndex = st.sCmn.indices.lagIndex; c->contourIndex = st.sCmn.indices.contourIndex;
}

/* ---- silk_VAD_Init / silk_encode_do_VAD_FLP (silk_VAD_GetSA_Q8 inside) / silk_HP_variable_cutoff on a scratch encoder struct
 * (silk/VAD.c:46-360, silk/float/encode_frame_FLP.c:45-82, silk/HP_variable_cutoff.c:39-77) ---- */
typedef struct {
  opus_int32 AnaState[2], AnaState1[2], AnaState2[2], XnrgSubfr[4], NrgRatioSmth_Q8[4], HPstate, NL[4], inv_NL[4], NoiseLevelBias[4], counter;
  opus_int32 speech_activity_Q8, input_tilt_Q15, input_quality_bands_Q15[4];
  opus_int32 signalType, noSpeechCounter, inDTX, VAD_flag;
  opus_int32 prevSignalType, prevLag, variable_HP_smth1_Q15;
} RefhSilkVad;   /* == OrcSilkVad */
static void refh_vad_load(silk_encoder_state *s, const RefhSilkVad *c) {
  int b;
  silk_VAD_state *v = &s->sVAD;
  for (b = 0; b < 2; b++) { v->AnaState[b] = c->AnaState[b]; v->AnaState1[b] = c->AnaState1[b]; v->AnaState2[b] = c->AnaState2[b]; }
  for (b = 0; b < 4; b++) { v->XnrgSubfr[b] = c->XnrgSubfr[b]; v->NrgRatioSmth_Q8[b] = c->NrgRatioSmth_Q8[b]; v->NL[b] = c->NL[b]; v->inv_NL[b] = c->inv_NL[b];
    v->NoiseLevelBias[b] = c->NoiseLevelBias[b]; s->input_quality_bands_Q15[b] = c->input_quality_bands_Q15[b]; }
  v->HPstate = (opus_int16)c->HPstate; v->counter = c->counter;
  s->speech_activity_Q8 = c->speech_activity_Q8; s->input_tilt_Q15 = c->input_tilt_Q15; s->indices.signalType = (opus_int8)c->signalType;
  s->noSpeechCounter = c->noSpeechCounter; s->inDTX = c->inDTX; s->prevSignalType = c->prevSignalType; s->prevLag = c->prevLag;
  s->variable_HP_smth1_Q15 = c->variable_HP_smth1_Q15;
}
static void refh_vad_store(RefhSilkVad *c, const silk_encoder_state *s) {
  int b;
  const silk_VAD_state *v = &s->sVAD;
  for (b = 0; b < 2; b++) { c->AnaState[b] = v->AnaState[b]; c->AnaState1[b] = v->AnaState1[b]; c->AnaState2[b] = v->AnaState2[b]; }
  for (b = 0; b < 4; b++) { c->XnrgSubfr[b] = v->XnrgSubfr[b]; c->NrgRatioSmth_Q8[b] = v->NrgRatioSmth_Q8[b]; c->NL[b] = v->NL[b]; c->inv_NL[b] = v->inv_NL[b];
    c->NoiseLevelBias[b] = v->NoiseLevelBias[b]; c->input_quality_bands_Q15[b] = s->input_quality_bands_Q15[b]; }
  c->HPstate = v->HPstate; c->counter = v->counter;
  c->speech_activity_Q8 = s->speech_activity_Q8; c->input_tilt_Q15 = s->input_tilt_Q15; c->signalType = s->indices.signalType;
  c->noSpeechCounter = s->noSpeechCounter; c->inDTX = s->inDTX; c->prevSignalType = s->prevSignalType; c->prevLag = s->prevLag;
  c->variable_HP_smth1_Q15 = s->variable_HP_smth1_Q15;
}
void refh_silk_vad_init(RefhSilkVad *c) {
  static silk_encoder_state_FLP st;
  memset(&st, 0, sizeof(st));
  memset(c, 0, sizeof(*c));
  silk_VAD_Init(&st.sCmn.sVAD);
  refh_vad_store(c, &st.sCmn);
}
void refh_silk_encode_do_vad(RefhSilkVad *c, const opus_int16 *pIn, int frame_length, int fs_kHz, int activity) {
  static silk_encoder_state_FLP st;
  memset(&st, 0, sizeof(st));
  refh_vad_load(&st.sCmn, c);
  st.sCmn.frame_length = frame_length; st.sCmn.fs_kHz = fs_kHz; st.sCmn.nFramesEncoded = 0;
  memcpy(st.sCmn.inputBuf + 1, pIn, (size_t)frame_length * sizeof(opus_int16));
  silk_encode_do_VAD_FLP(&st, activity);
  refh_vad_store(c, &st.sCmn);
  c->VAD_flag = st.sCmn.VAD_flags[0];
}
void refh_silk_hp_variable_cutoff(RefhSilkVad *c, int fs_kHz) {
  static silk_encoder_state_FLP st[2];
  memset(st, 0, sizeof(st));
  refh_vad_load(&st[0].sCmn, c);
  st[0].sCmn.fs_kHz = fs_kHz;
  silk_HP_variable_cutoff(st);
  refh_vad_store(c, &st[0].sCmn);
}
