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

/* ---- neural packet-loss concealment: dnn/lpcnet_plc.c:44-207 (lpcnet_plc_reset, lpcnet_plc_fec_add / _clear, compute_plc_pred, get_fec_or_pred,
 * queue_features, lpcnet_plc_update, lpcnet_plc_conceal) on top of the FARGAN restatement above.  The predictor is the 6 LinearLayers of the
 * generated PLCModel struct (plc_data.h): dense_in, dense_out, gru1 input / recurrent, gru2 input / recurrent.
 * The history analysis a first concealed frame runs when good audio came before it (:160-177: burg_cepstral_analysis,
 * lpcnet_compute_single_frame_features_float) uses lpcnet_enc_oracle.c through orc_plc_set_encoder(); without an encoder orc_plc_conceal returns -1
 * and leaves the state alone when a call would need it.  Pinned against the reference's own lpcnet_plc.c compiled with seeded random weights
 * (oracle/_ref/libfargan_ref.so, gen_plc_weights.py) by tests/test_plc_oracle_vs_ref.py. ---- */
#define PLC_MAX_FEC 104            /* dnn/lpcnet_private.h:15 */
#define CONT_VECTORS 5
#define NB_TOTAL_FEATURES 36       /* dnn/lpcnet.h:36 */
#define PLC_BUF_SIZE ((CONT_VECTORS + 10) * FRAME)
#define PLC_IN (2 * NB_BANDS + NB_FEATURES + 1)
enum { P_DENSE_IN, P_DENSE_OUT, P_GRU1_IN, P_GRU1_REC, P_GRU2_IN, P_GRU2_REC, P_COUNT };
typedef struct { float gru1[128], gru2[128]; } OrcPlcNet;
void orc_lpcnet_enc_reset(void *st);                                          /* lpcnet_enc_oracle.c */
void orc_lpcnet_enc_features(void *st, const float *pcm, float *features);
void orc_burg_cepstral_analysis(float *ceps, const float *x);
typedef struct {
  OrcLinear l[P_COUNT];
  OrcFargan fargan;
  void *enc;                        /* LPCNetPLCState.enc: an OrcLpcnetEnc of lpcnet_enc_oracle.c, or NULL */
  float fec[PLC_MAX_FEC][NB_FEATURES];
  int analysis_gap, fec_read_pos, fec_fill_pos, fec_skip, analysis_pos, predict_pos;
  float pcm[PLC_BUF_SIZE];
  int blend;
  float features[NB_TOTAL_FEATURES];
  float cont_features[CONT_VECTORS * NB_FEATURES];
  int loss_count;
  OrcPlcNet net, bak[2];
} OrcPlc;
int orc_plc_size(void) { return (int)sizeof(OrcPlc); }
void orc_plc_reset(OrcPlc *st) { /* lpcnet_plc.c:44-56: everything from `fec` on, the model and the vocoder stay */
  memset(st->fec, 0, sizeof(*st) - ((char *)st->fec - (char *)st));
  if (st->enc) orc_lpcnet_enc_reset(st->enc); /* :48 */
  st->blend = 0;
  st->loss_count = 0;
  st->analysis_gap = 1;
  st->analysis_pos = PLC_BUF_SIZE;
  st->predict_pos = PLC_BUF_SIZE;
}
void orc_plc_init(OrcPlc *st, const OrcLinear *plc_layers, const OrcLinear *fargan_layers) {
  memset(st, 0, sizeof(*st));
  memcpy(st->l, plc_layers, sizeof(st->l));
  orc_fargan_init(&st->fargan, fargan_layers);
  orc_plc_reset(st);
}
void orc_plc_set_encoder(OrcPlc *st, void *enc) { st->enc = enc; }
void orc_plc_fec_add(OrcPlc *st, const float *features) { /* :88-96 */
  if (!features) { st->fec_skip++; return; }
  memcpy(st->fec[st->fec_fill_pos], features, NB_FEATURES * sizeof(float));
  st->fec_fill_pos++;
}
void orc_plc_fec_clear(OrcPlc *st) { st->fec_read_pos = st->fec_fill_pos = st->fec_skip = 0; }
/* compute_plc_pred, :103-113 */
void orc_plc_pred(OrcPlc *st, float *out, const float *in) {
  float tmp[128];
  orc_nnet_dense(&st->l[P_DENSE_IN], tmp, in, ACT_TANH);
  orc_nnet_gru(&st->l[P_GRU1_IN], &st->l[P_GRU1_REC], st->net.gru1, tmp);
  orc_nnet_gru(&st->l[P_GRU2_IN], &st->l[P_GRU2_REC], st->net.gru2, st->net.gru1);
  orc_nnet_dense(&st->l[P_DENSE_OUT], out, st->net.gru2, ACT_LINEAR);
}
static int plc_fec_or_pred(OrcPlc *st, float *out) { /* :115-133 */
  float in[PLC_IN];
  memset(in, 0, sizeof(in));
  if (st->fec_read_pos != st->fec_fill_pos && st->fec_skip == 0) {
    float discard[NB_FEATURES];
    memcpy(out, st->fec[st->fec_read_pos], NB_FEATURES * sizeof(float));
    st->fec_read_pos++;
    memcpy(&in[2 * NB_BANDS], out, NB_FEATURES * sizeof(float));
    in[2 * NB_BANDS + NB_FEATURES] = -1;
    orc_plc_pred(st, discard, in);
    return 1;
  }
  orc_plc_pred(st, out, in);
  if (st->fec_skip > 0) st->fec_skip--;
  return 0;
}
static void plc_queue(OrcPlc *st, const float *features) { /* :135-138 */
  memmove(&st->cont_features[0], &st->cont_features[NB_FEATURES], (CONT_VECTORS - 1) * NB_FEATURES * sizeof(float));
  memcpy(&st->cont_features[(CONT_VECTORS - 1) * NB_FEATURES], features, NB_FEATURES * sizeof(float));
}
int orc_plc_update(OrcPlc *st, const int16_t *pcm) { /* :143-153 */
  int i;
  if (st->analysis_pos - FRAME >= 0) st->analysis_pos -= FRAME;
  else st->analysis_gap = 1;
  if (st->predict_pos - FRAME >= 0) st->predict_pos -= FRAME;
  memmove(st->pcm, &st->pcm[FRAME], (PLC_BUF_SIZE - FRAME) * sizeof(float));
  for (i = 0; i < FRAME; i++) st->pcm[PLC_BUF_SIZE - FRAME + i] = (1.f / 32768.f) * pcm[i];
  st->loss_count = 0;
  st->blend = 0;
  return 0;
}
int orc_plc_conceal(OrcPlc *st, int16_t *pcm) { /* :156-207 */
  static const float att_table[10] = {0, 0, -.2, -.2, -.4, -.4, -.8, -.8, -1.6, -1.6};
  float fpcm[FRAME];
  int i;
  if (st->blend == 0) {
    int count = 0;
    if (st->analysis_pos + FRAME <= PLC_BUF_SIZE && !st->enc) return -1; /* the history analysis needs the feature extractor */
    st->net = st->bak[0];
    while (st->analysis_pos + FRAME <= PLC_BUF_SIZE) { /* :160-177 */
      float x[FRAME], plc_features[PLC_IN];
      for (i = 0; i < FRAME; i++) x[i] = 32768.f * st->pcm[st->analysis_pos + i];
      orc_burg_cepstral_analysis(plc_features, x);
      orc_lpcnet_enc_features(st->enc, x, st->features);
      if ((!st->analysis_gap || count > 0) && st->analysis_pos >= st->predict_pos) {
        plc_queue(st, st->features);
        memcpy(&plc_features[2 * NB_BANDS], st->features, NB_FEATURES * sizeof(float));
        plc_features[2 * NB_BANDS + NB_FEATURES] = 1;
        st->bak[0] = st->bak[1];
        st->bak[1] = st->net;
        orc_plc_pred(st, st->features, plc_features);
      }
      st->analysis_pos += FRAME;
      count++;
    }
    st->bak[0] = st->bak[1];
    st->bak[1] = st->net;
    plc_fec_or_pred(st, st->features);
    plc_queue(st, st->features);
    st->bak[0] = st->bak[1];
    st->bak[1] = st->net;
    plc_fec_or_pred(st, st->features);
    plc_queue(st, st->features);
    orc_fargan_cont(&st->fargan, &st->pcm[PLC_BUF_SIZE - CONT_SAMPLES], st->cont_features);
    st->analysis_gap = 0;
  }
  st->bak[0] = st->bak[1];
  st->bak[1] = st->net;
  if (plc_fec_or_pred(st, st->features)) st->loss_count = 0;
  else st->loss_count++;
  {
    const float a = st->loss_count >= 10 ? st->features[0] + att_table[9] - 2 * (st->loss_count - 9) : st->features[0] + att_table[st->loss_count];
    st->features[0] = a > -15 ? a : -15;
  }
  orc_fargan_synthesize(&st->fargan, fpcm, st->features);
  for (i = 0; i < FRAME; i++) { /* fargan_synthesize_int, dnn/fargan.c:219-225 */
    float v = 32768.f * fpcm[i];
    v = v > -32767 ? v : -32767;
    v = v < 32767 ? v : 32767;
    pcm[i] = (int16_t)(int)floor(.5 + v);
  }
  plc_queue(st, st->features);
  if (st->analysis_pos - FRAME >= 0) st->analysis_pos -= FRAME;
  else st->analysis_gap = 1;
  st->predict_pos = PLC_BUF_SIZE;
  memmove(st->pcm, &st->pcm[FRAME], (PLC_BUF_SIZE - FRAME) * sizeof(float));
  for (i = 0; i < FRAME; i++) st->pcm[PLC_BUF_SIZE - FRAME + i] = (1.f / 32768.f) * pcm[i];
  st->blend = 1;
  return 0;
}
void orc_plc_get_state(const OrcPlc *st, float *features, float *cont_features, float *gru1, float *gru2, float *pcm, int *ints) {
  memcpy(features, st->features, sizeof(float) * NB_FEATURES);
  memcpy(cont_features, st->cont_features, sizeof(st->cont_features));
  memcpy(gru1, st->net.gru1, sizeof(st->net.gru1));
  memcpy(gru2, st->net.gru2, sizeof(st->net.gru2));
  memcpy(pcm, st->pcm, sizeof(st->pcm));
  ints[0] = st->blend; ints[1] = st->loss_count; ints[2] = st->fec_read_pos; ints[3] = st->fec_fill_pos; ints[4] = st->fec_skip;
  ints[5] = st->analysis_pos; ints[6] = st->predict_pos; ints[7] = st->analysis_gap;
}
