/* oracle/dred_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 * Plain-C restatement of the DRED decoder network, dnn/dred_rdovae_dec.c:32-145 (conv1_cond_init, DRED_rdovae_decode_all,
 * dred_rdovae_dec_init_states, dred_rdovae_decode_qframe) on the nnet primitives of nnet_oracle.c.  The model is the 29
 * LinearLayers of the generated RDOVAEDec struct (dred_rdovae_dec_data.h) in declaration order.  Pinned bit-exact against the
 * reference's own file compiled with seeded random weights (oracle/_ref/libfargan_ref.so, gen_dred_weights.py) by
 * tests/test_dred_oracle_vs_ref.py: the graph and its arithmetic are pinned, not the trained model. */
#include <stdint.h>
#include <string.h>

typedef struct {
  const float *bias, *subias;
  const int8_t *weights;
  const float *float_weights;
  const int *weights_idx;
  const float *diag, *scale;
  int nb_inputs, nb_outputs;
} OrcLinear; /* as in nnet_oracle.c */
void orc_nnet_dense(const OrcLinear *l, float *out, const float *in, int act);
void orc_nnet_gru(const OrcLinear *wi, const OrcLinear *wr, float *state, const float *in);
void orc_nnet_glu(const OrcLinear *l, float *out, const float *in);
void orc_nnet_conv1d(const OrcLinear *l, float *out, float *mem, const float *in, int input_size, int act);
enum { ACT_LINEAR = 0, ACT_TANH = 2 };

enum { D_DENSE1 = 0, D_GLU1 = 1, D_HIDDEN_INIT = 6, D_OUTPUT = 7, D_CONV_DENSE1 = 8, D_GRU_INIT = 13, D_GRU1_IN = 14, D_CONV1 = 24, D_COUNT = 29 };
#define NUM_FEATURES 20
#define GRU_N 64
#define CONV_N 32
#define DENSE1_N 96
#define BUF_N (DENSE1_N + 5 * (GRU_N + CONV_N))

typedef struct {
  OrcLinear l[D_COUNT];
  int initialized;
  float gru[5][GRU_N], conv[5][CONV_N];
} OrcDredDec;

int orc_dred_size(void) { return (int)sizeof(OrcDredDec); }
void orc_dred_init(OrcDredDec *st, const OrcLinear *layers) {
  memset(st, 0, sizeof(*st));
  memcpy(st->l, layers, sizeof(st->l));
}
/* dnn/dred_rdovae_dec.c:63-87 */
void orc_dred_init_states(OrcDredDec *st, const float *initial_state) {
  float hidden[128], state_init[5 * GRU_N];
  int k;
  orc_nnet_dense(&st->l[D_HIDDEN_INIT], hidden, initial_state, ACT_TANH);
  orc_nnet_dense(&st->l[D_GRU_INIT], state_init, hidden, ACT_TANH);
  for (k = 0; k < 5; k++) memcpy(st->gru[k], &state_init[k * GRU_N], GRU_N * sizeof(float));
  st->initialized = 0;
}
/* dnn/dred_rdovae_dec.c:90-145 */
void orc_dred_decode_qframe(OrcDredDec *st, float *qframe, const float *input) {
  float buffer[BUF_N], conv_tmp[64];
  int k, output_index = 0;
  orc_nnet_dense(&st->l[D_DENSE1], &buffer[output_index], input, ACT_TANH);
  output_index += DENSE1_N;
  for (k = 0; k < 5; k++) {
    orc_nnet_gru(&st->l[D_GRU1_IN + 2 * k], &st->l[D_GRU1_IN + 2 * k + 1], st->gru[k], buffer);
    orc_nnet_glu(&st->l[D_GLU1 + k], &buffer[output_index], st->gru[k]);
    output_index += GRU_N;
    if (!st->initialized) memset(st->conv[k], 0, CONV_N * sizeof(float)); /* conv1_cond_init, :32-41 */
    orc_nnet_dense(&st->l[D_CONV_DENSE1 + k], conv_tmp, buffer, ACT_TANH);
    orc_nnet_conv1d(&st->l[D_CONV1 + k], &buffer[output_index], st->conv[k], conv_tmp, CONV_N, ACT_TANH);
    output_index += CONV_N;
  }
  orc_nnet_dense(&st->l[D_OUTPUT], qframe, buffer, ACT_LINEAR);
  st->initialized = 1;
}
/* dnn/dred_rdovae_dec.c:44-61 */
void orc_dred_decode_all(const OrcLinear *layers, float *features, const float *state, const float *latents, int nb_latents, int latent_dim) {
  OrcDredDec dec;
  int i;
  orc_dred_init(&dec, layers);
  orc_dred_init_states(&dec, state);
  for (i = 0; i < 2 * nb_latents; i += 2) orc_dred_decode_qframe(&dec, &features[2 * i * NUM_FEATURES], &latents[(i / 2) * (latent_dim + 1)]);
}

/* ---- the encoder network: dred_rdovae_encode_dframe (dnn/dred_rdovae_enc.c:47-116).  24 LinearLayers in the order of the generated RDOVAEEnc struct
 * (dred_rdovae_enc_data.h as gen_dred_weights.py exports it): dense1, zdense, gdense1, gdense2, conv_dense1..5, gru1..5 (input, recurrent), conv1..5.
 * One call = one double feature frame (2 x 20 features) -> 25 latents + the 50-value initial state; conv1 has dilation 1, conv2..5 dilation 2. ---- */
void orc_nnet_conv1d_dilation(const OrcLinear *l, float *out, float *mem, const float *in, int input_size, int dilation, int act);
enum { E_DENSE1 = 0, E_ZDENSE = 1, E_GDENSE1 = 2, E_GDENSE2 = 3, E_CONV_DENSE1 = 4, E_GRU1_IN = 9, E_CONV1 = 19, E_COUNT = 24 };
#define E_GRU_N 32
#define E_CONV_N 64
#define E_DENSE1_N 64
#define E_BUF_N (E_DENSE1_N + 5 * (E_GRU_N + E_CONV_N))
typedef struct {
  OrcLinear l[E_COUNT];
  int initialized;
  float gru[5][E_GRU_N], conv[5][2 * E_CONV_N];
} OrcDredEnc;
int orc_dred_enc_size(void) { return (int)sizeof(OrcDredEnc); }
void orc_dred_enc_init(OrcDredEnc *st, const OrcLinear *layers) {
  memset(st, 0, sizeof(*st));
  memcpy(st->l, layers, sizeof(st->l));
}
void orc_dred_encode_dframe(OrcDredEnc *st, float *latents, float *initial_state, const float *input) {
  float buffer[E_BUF_N], conv_tmp[128], padded_latents[32], padded_state[56], state_hidden[128];
  int k, output_index = 0;
  orc_nnet_dense(&st->l[E_DENSE1], &buffer[output_index], input, ACT_TANH);
  output_index += E_DENSE1_N;
  for (k = 0; k < 5; k++) {
    const int dilation = k == 0 ? 1 : 2;
    orc_nnet_gru(&st->l[E_GRU1_IN + 2 * k], &st->l[E_GRU1_IN + 2 * k + 1], st->gru[k], buffer);
    memcpy(&buffer[output_index], st->gru[k], E_GRU_N * sizeof(float));
    output_index += E_GRU_N;
    if (!st->initialized) memset(st->conv[k], 0, dilation * E_CONV_N * sizeof(float)); /* conv1_cond_init, :39-45 */
    orc_nnet_dense(&st->l[E_CONV_DENSE1 + k], conv_tmp, buffer, ACT_TANH);
    if (k == 0) orc_nnet_conv1d(&st->l[E_CONV1 + k], &buffer[output_index], st->conv[k], conv_tmp, E_CONV_N, ACT_TANH);
    else orc_nnet_conv1d_dilation(&st->l[E_CONV1 + k], &buffer[output_index], st->conv[k], conv_tmp, E_CONV_N, 2, ACT_TANH);
    output_index += E_CONV_N;
  }
  orc_nnet_dense(&st->l[E_ZDENSE], padded_latents, buffer, ACT_LINEAR);
  memcpy(latents, padded_latents, 25 * sizeof(float));
  orc_nnet_dense(&st->l[E_GDENSE1], state_hidden, buffer, ACT_TANH);
  orc_nnet_dense(&st->l[E_GDENSE2], padded_state, state_hidden, ACT_LINEAR);
  memcpy(initial_state, padded_state, 50 * sizeof(float));
  st->initialized = 1;
}

/* ---- DRED encoder front end: dnn/dred_encoder.c:70-260 (dred_encoder_reset, dred_process_frame, filter_df2t, dred_convert_to_16k,
 * dred_compute_latents): the input is converted to 16 kHz mono in int16 range (zero-stuffing by `up`, a 7th-order elliptic low-pass in transposed
 * direct form II, decimation by 3), buffered to double frames of 320 samples, each of which gives two feature vectors (lpcnet_enc_oracle.c) and one
 * latent / state pair (orc_dred_encode_dframe) pushed at the front of the 104-deep ring buffers.  Pinned against the reference's own dred_encoder.c
 * (oracle/_ref/libfargan_ref.so) by tests/test_dred_oracle_vs_ref.py. ---- */
#include <math.h>
#include <stdlib.h>
#define DE_MAX_FRAMES 104     /* DRED_MAX_FRAMES, dnn/dred_config.h:52 */
#define DE_REDUNDANCY 52      /* DRED_NUM_REDUNDANCY_FRAMES */
#define DE_FRAME 160
#define DE_DELAY (79 + 12 - 80) /* DRED_SILK_ENCODER_DELAY */
#define DE_ORDER 8            /* RESAMPLING_ORDER, dnn/dred_encoder.h */
void orc_lpcnet_enc_reset(void *st);
void orc_lpcnet_enc_features(void *st, const float *pcm, float *features);
typedef struct {
  OrcDredEnc *net;
  void *lpcnet;               /* an OrcLpcnetEnc */
  int Fs, channels;
  float input_buffer[4 * DE_FRAME];
  int input_buffer_fill, dred_offset, latent_offset, last_extra_dred_offset;
  float latents_buffer[DE_MAX_FRAMES * 25];
  int latents_buffer_fill;
  float state_buffer[DE_MAX_FRAMES * 50];
  float resample_mem[DE_ORDER + 1];
} OrcDredFront;
int orc_dredenc_size(void) { return (int)sizeof(OrcDredFront); }
void orc_dredenc_reset(OrcDredFront *e) { /* dred_encoder_reset, :70-78 */
  memset(e->input_buffer, 0, sizeof(*e) - ((char *)e->input_buffer - (char *)e));
  e->input_buffer_fill = DE_DELAY;
  orc_lpcnet_enc_reset(e->lpcnet);
  e->net->initialized = 0;
  memset(e->net->gru, 0, sizeof(e->net->gru));
  memset(e->net->conv, 0, sizeof(e->net->conv));
}
void orc_dredenc_init(OrcDredFront *e, OrcDredEnc *net, void *lpcnet, int Fs, int channels) {
  memset(e, 0, sizeof(*e));
  e->net = net; e->lpcnet = lpcnet; e->Fs = Fs; e->channels = channels;
  orc_dredenc_reset(e);
}
static float de_to_int16(float x) { /* FLOAT2INT16, celt/float_cast.h:150-156 */
  x = x * 32768.f;
  x = x > -32768 ? x : -32768;
  x = x < 32767 ? x : 32767;
  return (float)(short)lrintf(x);
}
static void de_filter(const float *in, float *out, int len, float b0, const float *b, const float *a, float *mem) { /* filter_df2t, :109-126 */
  int i, j;
  for (i = 0; i < len; i++) {
    const float xi = in[i], yi = xi * b0 + mem[0], nyi = -yi;
    for (j = 0; j < DE_ORDER; j++) mem[j] = mem[j + 1] + b[j] * xi + a[j] * nyi;
    out[i] = yi;
  }
}
static void de_convert_to_16k(OrcDredFront *e, const float *in, int in_len, float *out, int out_len) { /* :133-213 */
  static const float b48[8] = {0.005873358047f, 0.012980854831f, 0.014531340042f, 0.014531340042f, 0.012980854831f, 0.005873358047f, 0.004523418224f, 0.f};
  static const float a48[8] = {-3.878718597768f, 7.748834257468f, -9.653651699533f, 8.007342726666f, -4.379450178552f, 1.463182111810f, -0.231720677804f, 0.f};
  static const float b12[8] = {-0.001017101081f, 0.003673127243f, 0.001009165267f, 0.001009165267f, 0.003673127243f, -0.001017101081f, 0.002033596776f, 0.f};
  static const float a12[8] = {-4.930414411612f, 11.291643096504f, -15.322037343815f, 13.216403930898f, -7.220409219553f, 2.310550142771f, -0.334338618782f, 0.f};
  static const float b8[8] = {0.081670120929f, 0.180401598565f, 0.259391051971f, 0.259391051971f, 0.180401598565f, 0.081670120929f, 0.020109185709f, 0.f};
  static const float a8[8] = {-1.393651933659f, 2.609789872676f, -2.403541968806f, 2.056814957331f, -1.148908574570f, 0.473001413788f, -0.110359852412f, 0.f};
  float downmix[960 * 2];
  const int up = e->Fs == 8000 ? 2 : e->Fs == 12000 ? 4 : e->Fs == 24000 ? 2 : 1;
  int i;
  memset(downmix, 0, up * in_len * sizeof(float));
  if (e->channels == 1) for (i = 0; i < in_len; i++) downmix[up * i] = up * de_to_int16(in[i]) + 1e-30f;
  else for (i = 0; i < in_len; i++) downmix[up * i] = up * de_to_int16(.5 * (in[2 * i] + in[2 * i + 1])) + 1e-30f;
  if (e->Fs == 16000) memcpy(out, downmix, out_len * sizeof(float));
  else if (e->Fs == 48000 || e->Fs == 24000) {
    de_filter(downmix, downmix, up * in_len, 0.004523418224f, b48, a48, e->resample_mem);
    for (i = 0; i < out_len; i++) out[i] = downmix[3 * i];
  } else if (e->Fs == 12000) {
    de_filter(downmix, downmix, up * in_len, 0.002033596776f, b12, a12, e->resample_mem);
    for (i = 0; i < out_len; i++) out[i] = downmix[3 * i];
  } else de_filter(downmix, out, out_len, 0.020109185709f, b8, a8, e->resample_mem);
}
static void de_process_frame(OrcDredFront *e) { /* dred_process_frame, :84-107 */
  float feature_buffer[2 * 36], input[2 * 20];
  memmove(e->latents_buffer + 25, e->latents_buffer, (DE_MAX_FRAMES - 1) * 25 * sizeof(float));
  memmove(e->state_buffer + 50, e->state_buffer, (DE_MAX_FRAMES - 1) * 50 * sizeof(float));
  orc_lpcnet_enc_features(e->lpcnet, e->input_buffer, feature_buffer);
  orc_lpcnet_enc_features(e->lpcnet, e->input_buffer + DE_FRAME, feature_buffer + 36);
  memcpy(input, feature_buffer, 20 * sizeof(float));
  memcpy(input + 20, feature_buffer + 36, 20 * sizeof(float));
  orc_dred_encode_dframe(e->net, e->latents_buffer, e->state_buffer, input);
  e->latents_buffer_fill = e->latents_buffer_fill + 1 < DE_REDUNDANCY ? e->latents_buffer_fill + 1 : DE_REDUNDANCY;
}
void orc_dredenc_compute_latents(OrcDredFront *e, const float *pcm, int frame_size, int extra_delay) { /* dred_compute_latents, :215-250 */
  int frame_size16k = frame_size * 16000 / e->Fs;
  int curr_offset16k = 40 + extra_delay * 16000 / e->Fs - e->input_buffer_fill;
  e->dred_offset = (int)floor((curr_offset16k + 20.f) / 40.f);
  e->latent_offset = 0;
  while (frame_size16k > 0) {
    const int process_size16k = 2 * DE_FRAME < frame_size16k ? 2 * DE_FRAME : frame_size16k;
    const int process_size = process_size16k * e->Fs / 16000;
    de_convert_to_16k(e, pcm, process_size, &e->input_buffer[e->input_buffer_fill], process_size16k);
    e->input_buffer_fill += process_size16k;
    if (e->input_buffer_fill >= 2 * DE_FRAME) {
      curr_offset16k += 320;
      de_process_frame(e);
      e->input_buffer_fill -= 2 * DE_FRAME;
      memmove(&e->input_buffer[0], &e->input_buffer[2 * DE_FRAME], e->input_buffer_fill * sizeof(float));
      if (e->dred_offset < 6) e->dred_offset += 8;
      else e->latent_offset++;
    }
    pcm += process_size * e->channels;
    frame_size16k -= process_size16k;
  }
}
void orc_dredenc_get(const OrcDredFront *e, float *latents, float *state, float *input_buffer, float *resample_mem, int *ints) {
  memcpy(latents, e->latents_buffer, sizeof(e->latents_buffer));
  memcpy(state, e->state_buffer, sizeof(e->state_buffer));
  memcpy(input_buffer, e->input_buffer, sizeof(e->input_buffer));
  memcpy(resample_mem, e->resample_mem, sizeof(e->resample_mem));
  ints[0] = e->input_buffer_fill; ints[1] = e->dred_offset; ints[2] = e->latent_offset; ints[3] = e->latents_buffer_fill; ints[4] = e->last_extra_dred_offset;
}

/* ---- entropy coding of the buffered latents: dred_encode_silk_frame (dnn/dred_encoder.c:281-383) with dred_encode_latents (:252-270),
 * dred_voice_active (:272-279), compute_quantizer (dnn/dred_coding.c:39-44), ec_laplace_encode_p0 (celt/laplace.c:136-162), ec_enc_icdf16
 * (celt/entenc.c:175-184) on the range-encoder restatement of celt_rangeenc.inc.  The statistical model (dred_rdovae_stats_data.c: eight uint8
 * tables) is passed in: stats[0..3] latent quant_scales / dead_zone / r / p0, stats[4..7] the same for the state. ---- */
static int ilog32(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }
#define BITRES 3
#define IMIN(a, b) ((a) < (b) ? (a) : (b))
#include "celt_rangeenc.inc"
void orc_nnet_activation(float *out, const float *in, int N, int act);
static void re_icdf16(RangeEnc *e, int s, const uint16_t *icdf, unsigned ftb) {
  const uint32_t r = e->rng >> ftb;
  if (s > 0) {
    e->val += e->rng - r * icdf[s - 1];
    e->rng = r * (icdf[s - 1] - icdf[s]);
  } else e->rng -= r * icdf[s];
  re_normalize(e);
}
static void laplace_p0(RangeEnc *e, int value, uint16_t p0, uint16_t decay) {
  uint16_t sign_icdf[3], icdf[8];
  int i;
  sign_icdf[0] = 32768 - p0;
  sign_icdf[1] = sign_icdf[0] / 2;
  sign_icdf[2] = 0;
  re_icdf16(e, value == 0 ? 0 : (value > 0 ? 1 : 2), sign_icdf, 15);
  value = abs(value);
  if (value) {
    icdf[0] = 7 > decay ? 7 : decay;
    for (i = 1; i < 7; i++) {
      const int v = (icdf[i - 1] * (int32_t)decay) >> 15;
      icdf[i] = 7 - i > v ? 7 - i : v;
    }
    icdf[7] = 0;
    value--;
    do {
      re_icdf16(e, value < 7 ? value : 7, icdf, 15);
      value -= 7;
    } while (value >= 0);
  }
}
static void encode_latents(RangeEnc *e, const float *x, const uint8_t *scale, const uint8_t *dzone, const uint8_t *r, const uint8_t *p0, int dim) {
  float xq[64], delta[64], deadzone[64];
  int q[64], i;
  const float eps = .1f;
  for (i = 0; i < dim; i++) {
    delta[i] = dzone[i] * (1.f / 256.f);
    xq[i] = x[i] * scale[i] * (1.f / 256.f);
    deadzone[i] = xq[i] / (delta[i] + eps);
  }
  orc_nnet_activation(deadzone, deadzone, dim, ACT_TANH);
  for (i = 0; i < dim; i++) {
    xq[i] = xq[i] - delta[i] * deadzone[i];
    q[i] = (int)floor(.5f + xq[i]);
  }
  for (i = 0; i < dim; i++) {
    if (r[i] == 0 || p0[i] == 255) q[i] = 0;
    else laplace_p0(e, q[i], p0[i] << 7, r[i] << 7);
  }
}
static int voice_active(const uint8_t *activity_mem, int offset) {
  int i;
  for (i = 0; i < 16; i++) {
    if (8 * offset + i >= DE_MAX_FRAMES * 4) break;
    if (activity_mem[8 * offset + i] == 1) return 1;
  }
  return 0;
}
static int dred_quantizer(int q0, int dQ, int qmax, int i) {
  static const int dQ_table[8] = {0, 2, 3, 4, 6, 8, 12, 16};
  const int quant = q0 + (dQ_table[dQ] * i + 8) / 16;
  return quant > qmax ? qmax : quant;
}
void orc_dredenc_set_last_extra(OrcDredFront *e, int v) { e->last_extra_dred_offset = v; }
int orc_dredenc_encode_silk_frame(OrcDredFront *e, const uint8_t *const *stats, uint8_t *buf, int max_chunks, int max_bytes, int q0, int dQ, int qmax,
                                  const uint8_t *activity_mem) {
  RangeEnc ec, bak;
  int i, latent_offset = e->latent_offset, extra_dred_offset = 0, dred_encoded = 0, delayed_dred = 0, prev_active = 0, total_offset, fill, lim;
  if (activity_mem[0] && e->last_extra_dred_offset > 0) {
    latent_offset = e->last_extra_dred_offset;
    delayed_dred = 1;
    e->last_extra_dred_offset = 0;
  }
  while (latent_offset < e->latents_buffer_fill - 1 && !voice_active(activity_mem, latent_offset)) {
    latent_offset++;
    extra_dred_offset++;
  }
  if (!delayed_dred) e->last_extra_dred_offset = extra_dred_offset;
  re_init(&ec, buf, (uint32_t)max_bytes);
  re_uint(&ec, q0, 16);
  re_uint(&ec, dQ, 8);
  total_offset = 16 - (e->dred_offset - extra_dred_offset * 8);
  if (total_offset > 31) {
    re_uint(&ec, 1, 2);
    re_uint(&ec, total_offset >> 5, 256);
    re_uint(&ec, total_offset & 31, 32);
  } else {
    re_uint(&ec, 0, 2);
    re_uint(&ec, total_offset, 32);
  }
  if (q0 < 14 && dQ > 0) {
    const int nvals = 15 - (q0 + 1);
    re_encode(&ec, qmax >= 15 ? 0 : nvals + qmax - (q0 + 1), qmax >= 15 ? nvals : nvals + qmax - q0, 2 * nvals);
  }
  encode_latents(&ec, &e->state_buffer[latent_offset * 50], stats[4] + q0 * 50, stats[5] + q0 * 50, stats[6] + q0 * 50, stats[7] + q0 * 50, 50);
  if (re_tell(&ec) > 8 * max_bytes) return 0;
  bak = ec;
  lim = 2 * max_chunks < e->latents_buffer_fill - latent_offset ? 2 * max_chunks : e->latents_buffer_fill - latent_offset;
  for (i = 0; i < lim; i += 2) {
    const int off = dred_quantizer(q0, dQ, qmax, i / 2) * 25;
    int active;
    encode_latents(&ec, e->latents_buffer + (i + latent_offset) * 25, stats[0] + off, stats[1] + off, stats[2] + off, stats[3] + off, 25);
    if (re_tell(&ec) > 8 * max_bytes) {
      if (i == 0) return 0;
      break;
    }
    active = voice_active(activity_mem, i + latent_offset);
    if (active || prev_active) {
      bak = ec;
      dred_encoded = i + 2;
    }
    prev_active = active;
  }
  if (dred_encoded == 0 || (dred_encoded <= 2 && extra_dred_offset)) return 0;
  ec = bak;
  fill = (re_tell(&ec) + 7) / 8;
  re_shrink(&ec, fill);
  re_done(&ec);
  return fill;
}
