/* oracle/fargan_harness.c — TEST INFRASTRUCTURE ONLY.  OURS: thin exported wrappers around the reference's FARGAN vocoder
 * (dnn/fargan.c, compiled unmodified from /root/reference with the generated random-weight fargan_data.c, see
 * gen_fargan_weights.py) so that tests can drive it through ctypes and read the model's layers back. */
#include <string.h>
#include "fargan.h"
#include "nnet.h"
#include "lpcnet.h"

int refh_fargan_state_size(void) { return (int)sizeof(FARGANState); }
void refh_fargan_init(FARGANState *st) { fargan_init(st); }
void refh_fargan_cont(FARGANState *st, const float *pcm0, const float *features0) { fargan_cont(st, pcm0, features0); }
void refh_fargan_synthesize(FARGANState *st, float *pcm, const float *features) { fargan_synthesize(st, pcm, features); }
int refh_fargan_nlayers(void) { return (int)(sizeof(FARGAN) / sizeof(LinearLayer)); }
/* the model struct is 20 LinearLayers in declaration order (fargan_data.h); dims[0..1] = nb_inputs, nb_outputs */
void refh_fargan_layer(FARGANState *st, int i, const void **ptrs /* bias subias weights float_weights weights_idx diag scale */, int *dims) {
  const LinearLayer *l = (const LinearLayer *)&st->model + i;
  ptrs[0] = l->bias; ptrs[1] = l->subias; ptrs[2] = l->weights; ptrs[3] = l->float_weights; ptrs[4] = l->weights_idx; ptrs[5] = l->diag;
  ptrs[6] = l->scale;
  dims[0] = l->nb_inputs; dims[1] = l->nb_outputs;
}
/* state access for parity checks */
void refh_fargan_get_state(const FARGANState *st, float *pitch_buf, float *cond_conv1, float *fwc0, float *g1, float *g2, float *g3, float *misc) {
  memcpy(pitch_buf, st->pitch_buf, sizeof(st->pitch_buf));
  memcpy(cond_conv1, st->cond_conv1_state, sizeof(st->cond_conv1_state));
  memcpy(fwc0, st->fwc0_mem, sizeof(st->fwc0_mem));
  memcpy(g1, st->gru1_state, sizeof(st->gru1_state));
  memcpy(g2, st->gru2_state, sizeof(st->gru2_state));
  memcpy(g3, st->gru3_state, sizeof(st->gru3_state));
  misc[0] = st->deemph_mem; misc[1] = (float)st->last_period;
}
int refh_fargan_const(int which) {
  switch (which) {
    case 0: return PITCH_MAX_PERIOD;
    case 1: return NB_FEATURES;
    case 2: return NB_BANDS;
    case 3: return COND_NET_FCONV1_STATE_SIZE;
    case 4: return SIG_NET_FWC0_STATE_SIZE;
    default: return -1;
  }
}

/* ---- DRED decoder (dnn/dred_rdovae_dec.c, compiled unmodified with the generated random-weight dred_rdovae_dec_data.c) ---- */
#include "dred_rdovae_dec.h"
#include "dred_rdovae_constants.h"
extern const WeightArray rdovaedec_arrays[];
static RDOVAEDec g_dred_model;
static int g_dred_ready = 0;
static void dred_ready(void) {
  if (!g_dred_ready) { init_rdovaedec(&g_dred_model, rdovaedec_arrays); g_dred_ready = 1; }
}
int refh_dred_nlayers(void) { return (int)(sizeof(RDOVAEDec) / sizeof(LinearLayer)); }
void refh_dred_layer(int i, const void **ptrs, int *dims) {
  const LinearLayer *l;
  dred_ready();
  l = (const LinearLayer *)&g_dred_model + i;
  ptrs[0] = l->bias; ptrs[1] = l->subias; ptrs[2] = l->weights; ptrs[3] = l->float_weights; ptrs[4] = l->weights_idx; ptrs[5] = l->diag;
  ptrs[6] = l->scale;
  dims[0] = l->nb_inputs; dims[1] = l->nb_outputs;
}
int refh_dred_const(int which) { return which == 0 ? DRED_NUM_FEATURES : which == 1 ? DRED_LATENT_DIM : which == 2 ? DRED_STATE_DIM : -1; }
int refh_dred_state_size(void) { return (int)sizeof(RDOVAEDecState); }
void refh_dred_init_states(RDOVAEDecState *h, const float *initial_state) { dred_ready(); dred_rdovae_dec_init_states(h, &g_dred_model, initial_state, 0); }
void refh_dred_decode_qframe(RDOVAEDecState *h, float *qframe, const float *z) { dred_ready(); dred_rdovae_decode_qframe(h, &g_dred_model, qframe, z, 0); }
void refh_dred_decode_all(float *features, const float *state, const float *latents, int nb_latents) {
  dred_ready();
  DRED_rdovae_decode_all(&g_dred_model, features, state, latents, nb_latents, 0);
}

/* ---- neural packet-loss concealment (dnn/lpcnet_plc.c, compiled unmodified with the generated random-weight plc_data.c) ----
 * The history analysis of the first concealed frame after good audio (burg_cepstral_analysis, lpcnet_compute_single_frame_features_float:
 * dnn/burg.c, lpcnet_enc.c, freq.c, pitchdnn.c) is compiled unmodified too, with the generated random-weight pitchdnn_data.c. */
#include <stdio.h>
#include <stdlib.h>
#include "lpcnet_private.h"

/* ---- feature extraction (dnn/lpcnet_enc.c, freq.c, pitchdnn.c, burg.c) ---- */
int refh_lpcnet_enc_state_size(void) { return (int)sizeof(LPCNetEncState); }
void refh_lpcnet_enc_init(LPCNetEncState *st) { lpcnet_encoder_init(st); }
void refh_lpcnet_enc_features(LPCNetEncState *st, const float *pcm, float *features) { lpcnet_compute_single_frame_features_float(st, pcm, features, 0); }
void refh_lpcnet_enc_get(const LPCNetEncState *st, float *if_features, float *xcorr_features, float *lpc, float *misc /* dnn_pitch, pitch_filt, mem_preemph */) {
  memcpy(if_features, st->if_features, sizeof(st->if_features));
  memcpy(xcorr_features, st->xcorr_features, sizeof(st->xcorr_features));
  memcpy(lpc, st->lpc, sizeof(st->lpc));
  misc[0] = st->dnn_pitch; misc[1] = st->pitch_filt; misc[2] = st->mem_preemph;
}
/* the six linear layers of the pitch model: 0..3 dense (if_upsampler_1/2, downsampler, final_upsampler), 4 / 5 the GRU's input / recurrent half */
void refh_pitchdnn_linear(LPCNetEncState *st, int i, const void **ptrs, int *dims) {
  const PitchDNN *m = &st->pitchdnn.model;
  const LinearLayer *l = i == 0 ? &m->dense_if_upsampler_1 : i == 1 ? &m->dense_if_upsampler_2 : i == 2 ? &m->dense_downsampler :
                         i == 3 ? &m->dense_final_upsampler : i == 4 ? &m->gru_1_input : &m->gru_1_recurrent;
  ptrs[0] = l->bias; ptrs[1] = l->subias; ptrs[2] = l->weights; ptrs[3] = l->float_weights; ptrs[4] = l->weights_idx; ptrs[5] = l->diag;
  ptrs[6] = l->scale;
  dims[0] = l->nb_inputs; dims[1] = l->nb_outputs;
}
void refh_pitchdnn_conv2d(LPCNetEncState *st, int i, const void **ptrs /* bias float_weights */, int *dims /* in_ch out_ch ktime kheight */) {
  const Conv2dLayer *c = i == 0 ? &st->pitchdnn.model.conv2d_1 : &st->pitchdnn.model.conv2d_2;
  ptrs[0] = c->bias; ptrs[1] = c->float_weights;
  dims[0] = c->in_channels; dims[1] = c->out_channels; dims[2] = c->ktime; dims[3] = c->kheight;
}
float refh_pitchdnn_compute(LPCNetEncState *st, const float *if_features, const float *xcorr_features) {
  return compute_pitchdnn(&st->pitchdnn, if_features, xcorr_features, 0);
}
void refh_burg_cepstral_analysis(float *ceps, const float *x) { burg_cepstral_analysis(ceps, x); }
float refh_lpc_from_cepstrum(float *lpc, const float *cepstrum) { return lpc_from_cepstrum(lpc, cepstrum); }

int refh_plc_state_size(void) { return (int)sizeof(LPCNetPLCState); }
int refh_plc_init(LPCNetPLCState *st) { return lpcnet_plc_init(st); }
void refh_plc_reset(LPCNetPLCState *st) { lpcnet_plc_reset(st); }
void refh_plc_fec_add(LPCNetPLCState *st, const float *features) { lpcnet_plc_fec_add(st, features); }
void refh_plc_fec_clear(LPCNetPLCState *st) { lpcnet_plc_fec_clear(st); }
int refh_plc_conceal(LPCNetPLCState *st, opus_int16 *pcm) { return lpcnet_plc_conceal(st, pcm); }
int refh_plc_update_pcm(LPCNetPLCState *st, opus_int16 *pcm) { return lpcnet_plc_update(st, pcm); }
int refh_plc_nlayers(void) { return (int)(sizeof(PLCModel) / sizeof(LinearLayer)); }
void refh_plc_layer(LPCNetPLCState *st, int i, const void **ptrs, int *dims) {
  const LinearLayer *l = (const LinearLayer *)&st->model + i;
  ptrs[0] = l->bias; ptrs[1] = l->subias; ptrs[2] = l->weights; ptrs[3] = l->float_weights; ptrs[4] = l->weights_idx; ptrs[5] = l->diag;
  ptrs[6] = l->scale;
  dims[0] = l->nb_inputs; dims[1] = l->nb_outputs;
}
/* the FARGAN model inside the PLC state (the same generated weights as refh_fargan_*) */
void refh_plc_fargan_layer(LPCNetPLCState *st, int i, const void **ptrs, int *dims) { refh_fargan_layer(&st->fargan, i, ptrs, dims); }
void refh_plc_get_state(const LPCNetPLCState *st, float *features, float *cont_features, float *gru1, float *gru2, float *pcm, int *ints) {
  memcpy(features, st->features, sizeof(float) * NB_FEATURES);
  memcpy(cont_features, st->cont_features, sizeof(st->cont_features));
  memcpy(gru1, st->plc_net.gru1_state, sizeof(st->plc_net.gru1_state));
  memcpy(gru2, st->plc_net.gru2_state, sizeof(st->plc_net.gru2_state));
  memcpy(pcm, st->pcm, sizeof(st->pcm));
  ints[0] = st->blend; ints[1] = st->loss_count; ints[2] = st->fec_read_pos; ints[3] = st->fec_fill_pos; ints[4] = st->fec_skip;
  ints[5] = st->analysis_pos; ints[6] = st->predict_pos; ints[7] = st->analysis_gap;
}
/* compute_plc_pred alone (static in lpcnet_plc.c): the same four calls on the model, for the layer-level pin */
void refh_plc_pred(LPCNetPLCState *st, float *out, const float *in) {
  float tmp[PLC_DENSE_IN_OUT_SIZE];
  compute_generic_dense(&st->model.plc_dense_in, tmp, in, ACTIVATION_TANH, 0);
  compute_generic_gru(&st->model.plc_gru1_input, &st->model.plc_gru1_recurrent, st->plc_net.gru1_state, tmp, 0);
  compute_generic_gru(&st->model.plc_gru2_input, &st->model.plc_gru2_recurrent, st->plc_net.gru2_state, st->plc_net.gru1_state, 0);
  compute_generic_dense(&st->model.plc_dense_out, out, st->plc_net.gru2_state, ACTIVATION_LINEAR, 0);
}

/* ---- DRED encoder network (dnn/dred_rdovae_enc.c, compiled unmodified with the generated random-weight dred_rdovae_enc_data.c) ---- */
#include "dred_rdovae_enc.h"
extern const WeightArray rdovaeenc_arrays[];
static RDOVAEEnc g_dred_enc_model;
static int g_dred_enc_ready = 0;
static void dred_enc_ready(void) {
  if (!g_dred_enc_ready) { init_rdovaeenc(&g_dred_enc_model, rdovaeenc_arrays); g_dred_enc_ready = 1; }
}
int refh_dred_enc_nlayers(void) { return (int)(sizeof(RDOVAEEnc) / sizeof(LinearLayer)); }
void refh_dred_enc_layer(int i, const void **ptrs, int *dims) {
  const LinearLayer *l;
  dred_enc_ready();
  l = (const LinearLayer *)&g_dred_enc_model + i;
  ptrs[0] = l->bias; ptrs[1] = l->subias; ptrs[2] = l->weights; ptrs[3] = l->float_weights; ptrs[4] = l->weights_idx; ptrs[5] = l->diag;
  ptrs[6] = l->scale;
  dims[0] = l->nb_inputs; dims[1] = l->nb_outputs;
}
int refh_dred_enc_state_size(void) { return (int)sizeof(RDOVAEEncState); }
/* input: 2 * DRED_NUM_FEATURES floats (a double feature frame); latents: DRED_LATENT_DIM, initial_state: DRED_STATE_DIM */
void refh_dred_encode_dframe(RDOVAEEncState *st, float *latents, float *initial_state, const float *input) {
  dred_enc_ready();
  dred_rdovae_encode_dframe(st, &g_dred_enc_model, latents, initial_state, input, 0);
}

/* ---- DRED encoder front end (dnn/dred_encoder.c, compiled unmodified): 16 kHz conversion, feature extraction, latents and state ring buffers ---- */
#include "dred_encoder.h"
int refh_dredenc_size(void) { return (int)sizeof(DREDEnc); }
void refh_dredenc_init(DREDEnc *e, int Fs, int channels) { dred_encoder_init(e, Fs, channels); }
void refh_dredenc_reset(DREDEnc *e) { dred_encoder_reset(e); }
int refh_dredenc_loaded(const DREDEnc *e) { return e->loaded; }
void refh_dredenc_compute_latents(DREDEnc *e, const float *pcm, int frame_size, int extra_delay) { dred_compute_latents(e, pcm, frame_size, extra_delay, 0); }
void refh_dredenc_get(const DREDEnc *e, float *latents, float *state, float *input_buffer, float *resample_mem, int *ints) {
  memcpy(latents, e->latents_buffer, sizeof(e->latents_buffer));
  memcpy(state, e->state_buffer, sizeof(e->state_buffer));
  memcpy(input_buffer, e->input_buffer, sizeof(e->input_buffer));
  memcpy(resample_mem, e->resample_mem, sizeof(e->resample_mem));
  ints[0] = e->input_buffer_fill; ints[1] = e->dred_offset; ints[2] = e->latent_offset; ints[3] = e->latents_buffer_fill; ints[4] = e->last_extra_dred_offset;
}
int refh_dredenc_encode_silk_frame(DREDEnc *e, unsigned char *buf, int max_chunks, int max_bytes, int q0, int dQ, int qmax, unsigned char *activity_mem) {
  return dred_encode_silk_frame(e, buf, max_chunks, max_bytes, q0, dQ, qmax, activity_mem, 0);
}
int refh_dredenc_const(int which) { return which == 0 ? DRED_MAX_FRAMES : which == 1 ? RESAMPLING_ORDER : which == 2 ? DRED_SILK_ENCODER_DELAY : -1; }
/* the statistical model of the latents (dred_rdovae_stats_data.c): which = 0..3 latent quant_scales / dead_zone / r / p0, 4..7 the same for the state */
#include "dred_rdovae_stats_data.h"
const unsigned char *refh_dred_stats(int which, int *len) {
  switch (which) {
    case 0: *len = sizeof(dred_latent_quant_scales_q8); return dred_latent_quant_scales_q8;
    case 1: *len = sizeof(dred_latent_dead_zone_q8); return dred_latent_dead_zone_q8;
    case 2: *len = sizeof(dred_latent_r_q8); return dred_latent_r_q8;
    case 3: *len = sizeof(dred_latent_p0_q8); return dred_latent_p0_q8;
    case 4: *len = sizeof(dred_state_quant_scales_q8); return dred_state_quant_scales_q8;
    case 5: *len = sizeof(dred_state_dead_zone_q8); return dred_state_dead_zone_q8;
    case 6: *len = sizeof(dred_state_r_q8); return dred_state_r_q8;
    case 7: *len = sizeof(dred_state_p0_q8); return dred_state_p0_q8;
  }
  *len = 0;
  return NULL;
}
void refh_dredenc_set_last_extra(DREDEnc *e, int v) { e->last_extra_dred_offset = v; }
