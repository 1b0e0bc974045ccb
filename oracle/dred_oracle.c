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
