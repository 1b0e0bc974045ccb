/* oracle/nnet_oracle.c — TEST INFRASTRUCTURE ONLY.
 * Plain-C restatement of the dnn/ inference primitives, generic (non-SIMD) flavour:
 *   compute_linear_c / compute_activation_c / compute_conv2d_c   dnn/nnet_arch.h:69-243
 *   sgemv, sparse_sgemv8x4, cgemv8x4, sparse_cgemv8x4,
 *   tanh_approx, sigmoid_approx, lpcnet_exp2                      dnn/vec.h:50-382 (the #else branch: no AVX / NEON)
 *   compute_generic_dense/gru/conv1d/conv1d_dilation, compute_glu dnn/nnet.c:60-153
 * The first group is pinned against the reference's own dnn/nnet_default.c compiled into oracle/_ref
 * (tests/test_nnet_oracle_vs_ref.py).  dnn/nnet.c itself cannot be compiled here (it includes generated model headers
 * that are not part of the source tree), so the second group — glue over the pinned primitives — is pinned by construction
 * only.  Model outputs (FARGAN/PLC/DRED weights) are unpinned: the weights are not in /root/reference. */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct {
  const float *bias;
  const float *subias;
  const int8_t *weights;
  const float *float_weights;
  const int *weights_idx;
  const float *diag;
  const float *scale;
  int nb_inputs;
  int nb_outputs;
} OrcLinear; /* dnn/nnet.h:66-76 */

typedef struct {
  const float *bias;
  const float *float_weights;
  int in_channels, out_channels, ktime, kheight;
} OrcConv2d; /* dnn/nnet.h:79-86 */

/* Which int8 convention: 0 = generic C (vec.h #else branch: signed 8-bit activations, layer->bias), 1 = the x86 build
 * (vec_avx.h:41,883 USE_SU_BIAS: activations offset to unsigned 127+q, layer->subias carries the compensated bias).
 * Both are exact integer dot products; they differ only in the float rounding of the bias compensation. */
static int g_su_bias = 0;
void orc_nnet_set_su_bias(int on) { g_su_bias = on; }

enum { ACT_LINEAR = 0, ACT_SIGMOID = 1, ACT_TANH = 2, ACT_RELU = 3, ACT_SOFTMAX = 4, ACT_SWISH = 5, ACT_EXP = 6 };

/* dnn/vec.h:317-336 */
static float exp2_approx(float x) {
  int integer;
  float frac;
  union { float f; uint32_t i; } res;
  integer = (int)floor(x);
  if (integer < -50) return 0;
  frac = x - integer;
  res.f = 0.99992522f + frac * (0.69583354f + frac * (0.22606716f + 0.078024523f * frac));
  res.i = (res.i + ((uint32_t)integer << 23)) & 0x7fffffff;
  return res.f;
}
static float exp_approx(float x) { return exp2_approx(x * 1.44269504f); }
/* dnn/vec.h:338-352 */
static float tanh_approx(float x) {
  const float N0 = 952.52801514f, N1 = 96.39235687f, N2 = 0.60863042f;
  const float D0 = 952.72399902f, D1 = 413.36801147f, D2 = 11.88600922f;
  float X2 = x * x, num, den;
  num = (N2 * X2 + N1) * X2 + N0;
  den = (D2 * X2 + D1) * X2 + D0;
  num = num * x / den;
  return num < -1.f ? -1.f : (num > 1.f ? 1.f : num);
}
static float sigmoid_approx(float x) { return .5f + .5f * tanh_approx(.5f * x); }

/* dnn/nnet_arch.h:69-119 (SOFTMAX_HACK is not defined in that translation unit) */
void orc_nnet_activation(float *out, const float *in, int N, int act) {
  int i;
  if (act == ACT_SIGMOID) for (i = 0; i < N; i++) out[i] = sigmoid_approx(in[i]);
  else if (act == ACT_TANH) for (i = 0; i < N; i++) out[i] = tanh_approx(in[i]);
  else if (act == ACT_SWISH) for (i = 0; i < N; i++) out[i] = in[i] * sigmoid_approx(in[i]);
  else if (act == ACT_RELU) for (i = 0; i < N; i++) out[i] = in[i] < 0 ? 0 : in[i];
  else if (act == ACT_SOFTMAX) {
    float sum = 0;
    for (i = 0; i < N; i++) out[i] = exp_approx(in[i]);
    for (i = 0; i < N; i++) sum += out[i];
    sum = 1.f / (sum + 1e-30);
    for (i = 0; i < N; i++) out[i] = sum * out[i];
  } else if (act == ACT_EXP) for (i = 0; i < N; i++) out[i] = exp_approx(in[i]);
  else if (in != out) for (i = 0; i < N; i++) out[i] = in[i];
}

/* dnn/nnet_arch.h:122-154 */
void orc_nnet_linear(const OrcLinear *l, float *out, const float *in) {
  const int M = l->nb_inputs, N = l->nb_outputs;
  const float *bias = l->bias;
  int i, j, r;
  if (l->float_weights) {
    const float *w = l->float_weights;
    if (l->weights_idx) { /* sparse_sgemv8x4, vec.h:124-176 */
      const int *idx = l->weights_idx;
      memset(out, 0, N * sizeof(float));
      for (i = 0; i < N; i += 8) {
        int cols = *idx++;
        for (j = 0; j < cols; j++) {
          int pos = *idx++, c;
          for (c = 0; c < 4; c++)
            for (r = 0; r < 8; r++) out[i + r] += w[c * 8 + r] * in[pos + c];
          w += 32;
        }
      }
    } else { /* sgemv, vec.h:50-122: every variant accumulates column by column */
      for (i = 0; i < N; i++) {
        float y = 0;
        for (j = 0; j < M; j++) y += w[j * N + i] * in[j];
        out[i] = y;
      }
    }
  } else if (l->weights && g_su_bias) { /* vec_avx.h:381-384,780-881 (SSE2 flavour): u8 x s8 integer dot, int32 accumulate */
    unsigned char x[2048];
    const int8_t *w = l->weights;
    const int *idx = l->weights_idx;
    for (i = 0; i < M; i++) x[i] = (unsigned char)(127 + (int)floor(.5 + 127 * in[i]));
    for (i = 0; i < N; i += 8) {
      int32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      int nblk = idx ? *idx++ : M / 4;
      for (j = 0; j < nblk; j++) {
        int pos = idx ? *idx++ : 4 * j;
        for (r = 0; r < 8; r++) acc[r] += w[r * 4 + 0] * x[pos + 0] + w[r * 4 + 1] * x[pos + 1] + w[r * 4 + 2] * x[pos + 2] + w[r * 4 + 3] * x[pos + 3];
        w += 32;
      }
      for (r = 0; r < 8; r++) out[i + r] = (float)acc[r] * l->scale[i + r];
    }
    bias = l->subias; /* nnet_arch.h:136-139 */
  } else if (l->weights) {
    int8_t x[2048];
    const int8_t *w = l->weights;
    for (i = 0; i < N; i++) out[i] = 0;
    for (i = 0; i < M; i++) x[i] = (int8_t)(int)floor(.5 + 127 * in[i]);
    if (l->weights_idx) { /* sparse_cgemv8x4, vec.h:249-282: integer products */
      const int *idx = l->weights_idx;
      for (i = 0; i < N; i += 8) {
        int colblocks = *idx++;
        for (j = 0; j < colblocks; j++) {
          int pos = *idx++;
          for (r = 0; r < 8; r++)
            out[i + r] += (w[r * 4 + 0] * x[pos + 0] + w[r * 4 + 1] * x[pos + 1] + w[r * 4 + 2] * x[pos + 2] + w[r * 4 + 3] * x[pos + 3]);
          w += 32;
        }
      }
    } else { /* cgemv8x4, vec.h:283-313: float products of small integers */
      for (i = 0; i < N; i += 8)
        for (j = 0; j < M; j += 4) {
          float x0 = x[j], x1 = x[j + 1], x2 = x[j + 2], x3 = x[j + 3];
          for (r = 0; r < 8; r++) out[i + r] += (w[r * 4 + 0] * x0 + w[r * 4 + 1] * x1 + w[r * 4 + 2] * x2 + w[r * 4 + 3] * x3);
          w += 32;
        }
    }
    for (i = 0; i < N; i++) out[i] *= l->scale[i];
  } else memset(out, 0, N * sizeof(float));
  if (bias) for (i = 0; i < N; i++) out[i] += bias[i];
  if (l->diag)
    for (i = 0; i < M; i++) {
      out[i] += l->diag[i] * in[i];
      out[i + M] += l->diag[i + M] * in[i];
      out[i + 2 * M] += l->diag[i + 2 * M] * in[i];
    }
}

/* dnn/nnet_arch.h:163-243 */
void orc_nnet_conv2d(const OrcConv2d *c, float *out, float *mem, const float *in, int height, int hstride, int act) {
  float in_buf[8192];
  const int time_stride = c->in_channels * (height + c->kheight - 1), in_stride = height + c->kheight - 1;
  const int kh = c->kheight, kt = c->ktime;
  int i, m, t, h, j;
  memcpy(in_buf, mem, (kt - 1) * time_stride * sizeof(float));
  memcpy(&in_buf[(kt - 1) * time_stride], in, time_stride * sizeof(float));
  memcpy(mem, &in_buf[time_stride], (kt - 1) * time_stride * sizeof(float));
  for (i = 0; i < c->out_channels; i++) {
    const float *w = &c->float_weights[i * c->in_channels * kt * kh];
    memset(&out[i * hstride], 0, height * sizeof(float));
    for (m = 0; m < c->in_channels; m++) {
      if (kh == 3 && kt == 3) { /* conv2d_3x3_float: the nine products are summed first, then added */
        for (j = 0; j < height; j++) {
          const float *x = &in_buf[m * in_stride + j];
          const float *wm = &w[m * 9];
          out[i * hstride + j] += wm[0] * x[0] + wm[1] * x[1] + wm[2] * x[2] + wm[3] * x[time_stride] + wm[4] * x[time_stride + 1] +
                                  wm[5] * x[time_stride + 2] + wm[6] * x[2 * time_stride] + wm[7] * x[2 * time_stride + 1] +
                                  wm[8] * x[2 * time_stride + 2];
        }
      } else {
        for (t = 0; t < kt; t++)
          for (h = 0; h < kh; h++)
            for (j = 0; j < height; j++)
              out[i * hstride + j] += w[m * kt * kh + t * kh + h] * in_buf[t * time_stride + m * in_stride + j + h];
      }
    }
  }
  if (c->bias)
    for (i = 0; i < c->out_channels; i++)
      for (j = 0; j < height; j++) out[i * hstride + j] += c->bias[i];
  for (i = 0; i < c->out_channels; i++) orc_nnet_activation(&out[i * hstride], &out[i * hstride], height, act);
}

/* ---- dnn/nnet.c:60-153 ---- */
void orc_nnet_dense(const OrcLinear *l, float *out, const float *in, int act) {
  orc_nnet_linear(l, out, in);
  orc_nnet_activation(out, out, l->nb_outputs, act);
}
void orc_nnet_gru(const OrcLinear *wi, const OrcLinear *wr, float *state, const float *in) {
  float zrh[3 * 1024], recur[3 * 1024];
  const int N = wr->nb_inputs;
  float *z = zrh, *r = &zrh[N], *h = &zrh[2 * N];
  int i;
  orc_nnet_linear(wi, zrh, in);
  orc_nnet_linear(wr, recur, state);
  for (i = 0; i < 2 * N; i++) zrh[i] += recur[i];
  orc_nnet_activation(zrh, zrh, 2 * N, ACT_SIGMOID);
  for (i = 0; i < N; i++) h[i] += recur[2 * N + i] * r[i];
  orc_nnet_activation(h, h, N, ACT_TANH);
  for (i = 0; i < N; i++) h[i] = z[i] * state[i] + (1 - z[i]) * h[i];
  for (i = 0; i < N; i++) state[i] = h[i];
}
void orc_nnet_glu(const OrcLinear *l, float *out, const float *in) {
  float act2[2048];
  int i;
  orc_nnet_linear(l, act2, in);
  orc_nnet_activation(act2, act2, l->nb_outputs, ACT_SIGMOID);
  for (i = 0; i < l->nb_outputs; i++) out[i] = in[i] * act2[i];
}
void orc_nnet_conv1d(const OrcLinear *l, float *out, float *mem, const float *in, int input_size, int act) {
  float tmp[2048];
  const int hist = l->nb_inputs - input_size;
  if (hist) memcpy(tmp, mem, hist * sizeof(float));
  memcpy(&tmp[hist], in, input_size * sizeof(float));
  orc_nnet_linear(l, out, tmp);
  orc_nnet_activation(out, out, l->nb_outputs, act);
  if (hist) memcpy(mem, &tmp[input_size], hist * sizeof(float));
}
void orc_nnet_conv1d_dilation(const OrcLinear *l, float *out, float *mem, const float *in, int input_size, int dilation, int act) {
  float tmp[2048];
  const int ksize = l->nb_inputs / input_size, hist = l->nb_inputs - input_size;
  int i;
  if (dilation == 1) memcpy(tmp, mem, hist * sizeof(float));
  else for (i = 0; i < ksize - 1; i++) memcpy(&tmp[i * input_size], &mem[i * input_size * dilation], input_size * sizeof(float));
  memcpy(&tmp[hist], in, input_size * sizeof(float));
  orc_nnet_linear(l, out, tmp);
  orc_nnet_activation(out, out, l->nb_outputs, act);
  if (dilation == 1) memcpy(mem, &tmp[input_size], hist * sizeof(float));
  else {
    memmove(mem, &mem[input_size], (input_size * dilation * (ksize - 1) - input_size) * sizeof(float));
    memcpy(&mem[input_size * dilation * (ksize - 1) - input_size], in, input_size * sizeof(float));
  }
}
