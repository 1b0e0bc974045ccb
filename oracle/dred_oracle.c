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
