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
