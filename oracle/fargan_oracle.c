/* oracle/fargan_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 * Plain-C restatement of the reference's FARGAN vocoder graph, dnn/fargan.c:41-217 (compute_fargan_cond, run_fargan_subframe,
 * fargan_cont, fargan_synthesize), on top of the nnet primitives of nnet_oracle.c.  The model is 20 LinearLayers in the order
 * of the generated FARGAN struct (fargan_data.h).  Pinned against the reference's own fargan.c compiled with seeded random
 * weights (oracle/_ref/libfargan_ref.so, gen_fargan_weights.py) by tests/test_fargan_oracle_vs_ref.py; the real weights
 * are a separate download that is not in the tree: parity of the GRAPH is pinned, audio quality is not. */
#include <math.h>
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
enum { ACT_LINEAR = 0, ACT_SIGMOID = 1, ACT_TANH = 2 };

enum { L_PEMBED, L_FDENSE1, L_FCONV1, L_FDENSE2, L_COND_GAIN, L_FWC0_CONV, L_FWC0_GLU, L_GRU1_IN, L_GRU1_REC, L_GRU2_IN, L_GRU2_REC,
       L_GRU3_IN, L_GRU3_REC, L_GRU1_GLU, L_GRU2_GLU, L_GRU3_GLU, L_SKIP_GLU, L_SKIP_DENSE, L_SIG_OUT, L_GAIN_OUT, L_COUNT };
#define NB_FEATURES 20 /* dnn/lpcnet.h:32 */
#define NB_BANDS 18    /* dnn/freq.h:48 */
#define PITCH_MAX 256  /* dnn/pitchdnn.h:10 */
#define SUBFRAME 40
#define NB_SUBFRAMES 4
#define FRAME (SUBFRAME * NB_SUBFRAMES)
#define CONT_SAMPLES 320
#define COND_SIZE 80   /* COND_NET_FDENSE2_OUT_SIZE / 4 */
#define SIG_IN (COND_SIZE + 2 * SUBFRAME + 4)
#define PEMBED 12
#define FWC0_OUT 192
#define GRU1 160
#define GRU2 128
#define GRU3 128
#define SKIP_OUT 128
#define DEEMPH 0.85f
#define IMAX(a, b) ((a) > (b) ? (a) : (b))
#define IMIN(a, b) ((a) < (b) ? (a) : (b))

typedef struct {
  OrcLinear l[L_COUNT];
  int cont_initialized, last_period;
  float deemph_mem;
  float pitch_buf[PITCH_MAX], cond_conv1_state[128], fwc0_mem[2 * SIG_IN], gru1_state[GRU1], gru2_state[GRU2], gru3_state[GRU3];
} OrcFargan;

int orc_fargan_size(void) { return (int)sizeof(OrcFargan); }
void orc_fargan_init(OrcFargan *st, const OrcLinear *layers) {
  memset(st, 0, sizeof(*st));
  memcpy(st->l, layers, sizeof(st->l));
}
static int period_of(const float *features) { /* dnn/fargan.c:148,198 */
  return (int)floor(.5 + 256. / pow(2.f, ((1. / 60.) * ((features[NB_BANDS] + 1.5) * 60))));
}
/* dnn/fargan.c:41-59 */
static void fargan_cond(OrcFargan *st, float *cond, const float *features, int period) {
  float dense_in[NB_FEATURES + PEMBED], conv1_in[64], fdense2_in[128];
  memcpy(&dense_in[NB_FEATURES], &st->l[L_PEMBED].float_weights[IMAX(0, IMIN(period - 32, 223)) * PEMBED], PEMBED * sizeof(float));
  memcpy(dense_in, features, NB_FEATURES * sizeof(float));
  orc_nnet_dense(&st->l[L_FDENSE1], conv1_in, dense_in, ACT_TANH);
  orc_nnet_conv1d(&st->l[L_FCONV1], fdense2_in, st->cond_conv1_state, conv1_in, 64, ACT_TANH);
  orc_nnet_dense(&st->l[L_FDENSE2], cond, fdense2_in, ACT_TANH);
}
/* dnn/fargan.c:69-141 */
static void fargan_subframe(OrcFargan *st, float *pcm, const float *cond, int period) {
  int i, pos;
  float fwc0_in[SIG_IN], gru1_in[FWC0_OUT + 2 * SUBFRAME], gru2_in[GRU1 + 2 * SUBFRAME], gru3_in[GRU2 + 2 * SUBFRAME];
  float pred[SUBFRAME + 4], prev[SUBFRAME], pitch_gate[4], gain, gain_1;
  float skip_cat[GRU1 + GRU2 + GRU3 + FWC0_OUT + 2 * SUBFRAME], skip_out[SKIP_OUT];
  orc_nnet_dense(&st->l[L_COND_GAIN], &gain, cond, ACT_LINEAR);
  gain = exp(gain);
  gain_1 = 1.f / (1e-5f + gain);
  pos = PITCH_MAX - period - 2;
  for (i = 0; i < SUBFRAME + 4; i++) {
    float v = gain_1 * st->pitch_buf[IMAX(0, pos)];
    pred[i] = v > 1.f ? 1.f : (v < -1.f ? -1.f : v);   /* MIN32(1, MAX32(-1, v)) */
    pos++;
    if (pos == PITCH_MAX) pos -= period;
  }
  for (i = 0; i < SUBFRAME; i++) {
    float v = gain_1 * st->pitch_buf[PITCH_MAX - SUBFRAME + i];
    prev[i] = v < 1.f ? (v > -1.f ? v : -1.f) : 1.f;    /* MAX32(-1, MIN16(1, v)) */
  }
  memcpy(&fwc0_in[0], cond, COND_SIZE * sizeof(float));
  memcpy(&fwc0_in[COND_SIZE], pred, (SUBFRAME + 4) * sizeof(float));
  memcpy(&fwc0_in[COND_SIZE + SUBFRAME + 4], prev, SUBFRAME * sizeof(float));
  orc_nnet_conv1d(&st->l[L_FWC0_CONV], gru1_in, st->fwc0_mem, fwc0_in, SIG_IN, ACT_TANH);
  orc_nnet_glu(&st->l[L_FWC0_GLU], gru1_in, gru1_in);
  orc_nnet_dense(&st->l[L_GAIN_OUT], pitch_gate, gru1_in, ACT_SIGMOID);
  for (i = 0; i < SUBFRAME; i++) gru1_in[FWC0_OUT + i] = pitch_gate[0] * pred[i + 2];
  memcpy(&gru1_in[FWC0_OUT + SUBFRAME], prev, SUBFRAME * sizeof(float));
  orc_nnet_gru(&st->l[L_GRU1_IN], &st->l[L_GRU1_REC], st->gru1_state, gru1_in);
  orc_nnet_glu(&st->l[L_GRU1_GLU], gru2_in, st->gru1_state);
  for (i = 0; i < SUBFRAME; i++) gru2_in[GRU1 + i] = pitch_gate[1] * pred[i + 2];
  memcpy(&gru2_in[GRU1 + SUBFRAME], prev, SUBFRAME * sizeof(float));
  orc_nnet_gru(&st->l[L_GRU2_IN], &st->l[L_GRU2_REC], st->gru2_state, gru2_in);
  orc_nnet_glu(&st->l[L_GRU2_GLU], gru3_in, st->gru2_state);
  for (i = 0; i < SUBFRAME; i++) gru3_in[GRU2 + i] = pitch_gate[2] * pred[i + 2];
  memcpy(&gru3_in[GRU2 + SUBFRAME], prev, SUBFRAME * sizeof(float));
  orc_nnet_gru(&st->l[L_GRU3_IN], &st->l[L_GRU3_REC], st->gru3_state, gru3_in);
  orc_nnet_glu(&st->l[L_GRU3_GLU], &skip_cat[GRU1 + GRU2], st->gru3_state);
  memcpy(skip_cat, gru2_in, GRU1 * sizeof(float));
  memcpy(&skip_cat[GRU1], gru3_in, GRU2 * sizeof(float));
  memcpy(&skip_cat[GRU1 + GRU2 + GRU3], gru1_in, FWC0_OUT * sizeof(float));
  for (i = 0; i < SUBFRAME; i++) skip_cat[GRU1 + GRU2 + GRU3 + FWC0_OUT + i] = pitch_gate[3] * pred[i + 2];
  memcpy(&skip_cat[GRU1 + GRU2 + GRU3 + FWC0_OUT + SUBFRAME], prev, SUBFRAME * sizeof(float));
  orc_nnet_dense(&st->l[L_SKIP_DENSE], skip_out, skip_cat, ACT_TANH);
  orc_nnet_glu(&st->l[L_SKIP_GLU], skip_out, skip_out);
  orc_nnet_dense(&st->l[L_SIG_OUT], pcm, skip_out, ACT_TANH);
  for (i = 0; i < SUBFRAME; i++) pcm[i] *= gain;
  memmove(st->pitch_buf, &st->pitch_buf[SUBFRAME], (PITCH_MAX - SUBFRAME) * sizeof(float));
  memcpy(&st->pitch_buf[PITCH_MAX - SUBFRAME], pcm, SUBFRAME * sizeof(float));
  for (i = 0; i < SUBFRAME; i++) { /* fargan_deemphasis, :61-67 */
    pcm[i] += DEEMPH * st->deemph_mem;
    st->deemph_mem = pcm[i];
  }
}
/* dnn/fargan.c:143-170 */
void orc_fargan_cont(OrcFargan *st, const float *pcm0, const float *features0) {
  int i, period = 0;
  float cond[4 * COND_SIZE], x0[CONT_SAMPLES], dummy[SUBFRAME];
  for (i = 0; i < 5; i++) {
    const float *features = &features0[i * NB_FEATURES];
    st->last_period = period;
    period = period_of(features);
    fargan_cond(st, cond, features, period);
  }
  x0[0] = 0;
  for (i = 1; i < CONT_SAMPLES; i++) x0[i] = pcm0[i] - DEEMPH * pcm0[i - 1];
  memcpy(&st->pitch_buf[PITCH_MAX - FRAME], x0, FRAME * sizeof(float));
  st->cont_initialized = 1;
  for (i = 0; i < NB_SUBFRAMES; i++) {
    fargan_subframe(st, dummy, &cond[i * COND_SIZE], st->last_period);
    memcpy(&st->pitch_buf[PITCH_MAX - SUBFRAME], &x0[FRAME + i * SUBFRAME], SUBFRAME * sizeof(float));
  }
  st->deemph_mem = pcm0[CONT_SAMPLES - 1];
}
/* dnn/fargan.c:191-213 */
void orc_fargan_synthesize(OrcFargan *st, float *pcm, const float *features) {
  int subframe, period;
  float cond[4 * COND_SIZE];
  period = period_of(features);
  fargan_cond(st, cond, features, period);
  for (subframe = 0; subframe < NB_SUBFRAMES; subframe++) fargan_subframe(st, &pcm[subframe * SUBFRAME], &cond[subframe * COND_SIZE], st->last_period);
  st->last_period = period;
}
void orc_fargan_get_state(const OrcFargan *st, float *pitch_buf, float *cond_conv1, float *fwc0, float *g1, float *g2, float *g3, float *misc) {
  memcpy(pitch_buf, st->pitch_buf, sizeof(st->pitch_buf));
  memcpy(cond_conv1, st->cond_conv1_state, sizeof(st->cond_conv1_state));
  memcpy(fwc0, st->fwc0_mem, sizeof(st->fwc0_mem));
  memcpy(g1, st->gru1_state, sizeof(st->gru1_state));
  memcpy(g2, st->gru2_state, sizeof(st->gru2_state));
  memcpy(g3, st->gru3_state, sizeof(st->gru3_state));
  misc[0] = st->deemph_mem; misc[1] = (float)st->last_period;
}
