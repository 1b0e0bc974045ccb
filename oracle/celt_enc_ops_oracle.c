/* oracle/celt_enc_ops_oracle.c — TEST INFRASTRUCTURE ONLY.
 * Plain-C restatements of encoder-side CELT kernel operations, each
 * citing the reference lines it follows; float operation order is the reference's
 * (-ffp-contract=off).  Pinned bit-for-bit against the reference's own functions in
 * oracle/_ref by tests/test_ops_oracle_vs_ref.py. */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define OB_TABLE static const
#include "../opus_b200/csrc/celt_tables_data.h"

typedef struct { float r, i; } cpx;
void orc_fft_inplace(cpx *fout, int shift); /* celt_dec_oracle.c */

static const short *bitrev_table(int shift) {
  return shift == 0 ? ob_fft_bitrev480 : shift == 1 ? ob_fft_bitrev240 : shift == 2 ? ob_fft_bitrev120 : ob_fft_bitrev60;
}

/* One forward MDCT (celt/mdct.c:122-264): in[N/2 + overlap] -> out[N/2] with the given stride. */
static void mdct_forward_block(const float *in, float *out, int shift, int stride) {
  const int overlap = 120;
  int i, N = 1920 >> shift, N2 = N >> 1, N4 = N >> 2;
  const float *trig = ob_mdct_trig;
  const float scale = 1.f / (float)(480 >> shift);
  float f[960];
  cpx f2[480];
  for (i = 0; i < shift; i++) trig += (1920 >> i) >> 1;
  {
    const float *xp1 = in + (overlap >> 1), *xp2 = in + N2 - 1 + (overlap >> 1);
    float *yp = f;
    const float *wp1 = ob_window + (overlap >> 1), *wp2 = ob_window + (overlap >> 1) - 1;
    for (i = 0; i < ((overlap + 3) >> 2); i++) {
      *yp++ = xp1[N2] * (*wp2) + (*xp2) * (*wp1);
      *yp++ = (*xp1) * (*wp1) - xp2[-N2] * (*wp2);
      xp1 += 2; xp2 -= 2; wp1 += 2; wp2 -= 2;
    }
    wp1 = ob_window;
    wp2 = ob_window + overlap - 1;
    for (; i < N4 - ((overlap + 3) >> 2); i++) {
      *yp++ = *xp2;
      *yp++ = *xp1;
      xp1 += 2; xp2 -= 2;
    }
    for (; i < N4; i++) {
      *yp++ = -(xp1[-N2] * (*wp1)) + (*xp2) * (*wp2);
      *yp++ = (*xp1) * (*wp2) + xp2[N2] * (*wp1);
      xp1 += 2; xp2 -= 2; wp1 += 2; wp2 -= 2;
    }
  }
  {
    const float *yp = f;
    const short *bitrev = bitrev_table(shift);
    for (i = 0; i < N4; i++) {
      float t0 = trig[i], t1 = trig[N4 + i], re = *yp++, im = *yp++;
      float yr = re * t0 - im * t1;
      float yi = im * t0 + re * t1;
      f2[bitrev[i]].r = yr * scale;
      f2[bitrev[i]].i = yi * scale;
    }
  }
  orc_fft_inplace(f2, shift);
  {
    const cpx *fp = f2;
    float *yp1 = out, *yp2 = out + stride * (N2 - 1);
    for (i = 0; i < N4; i++) {
      float t0 = trig[i], t1 = trig[N4 + i];
      *yp1 = fp->i * t1 - fp->r * t0;
      *yp2 = fp->r * t1 + fp->i * t0;
      fp++;
      yp1 += 2 * stride;
      yp2 -= 2 * stride;
    }
  }
}

/* compute_mdcts for one channel (celt/celt_encoder.c:511-554): in[B*N+overlap] -> out[B*N] */
void orc_mdct_forward(const float *in, float *out, int LM, int transient) {
  int B = transient ? 1 << LM : 1, N = transient ? 120 : 120 << LM, shift = transient ? 3 : 3 - LM, b;
  for (b = 0; b < B; b++) mdct_forward_block(in + b * N, &out[b], shift, B);
}

/* celt_pitch_xcorr_c (celt/pitch.c:230-305): every lag is a sequential dot product */
void orc_pitch_xcorr(const float *x, const float *y, float *xcorr, int len, int max_pitch) {
  int i, j;
  for (i = 0; i < max_pitch; i++) {
    float sum = 0;
    for (j = 0; j < len; j++) sum = sum + x[j] * y[i + j];
    xcorr[i] = sum;
  }
}

/* exp_rotation (celt/vq.c:75-147), both directions */
static void rotate_pass(float *X, int len, int stride, float c, float s) {
  int i;
  float ms = -s, *p = X;
  for (i = 0; i < len - stride; i++) {
    float x1 = p[0], x2 = p[stride];
    p[stride] = c * x2 + s * x1;
    *p++ = c * x1 + ms * x2;
  }
  p = &X[len - 2 * stride - 1];
  for (i = len - 2 * stride - 1; i >= 0; i--) {
    float x1 = p[0], x2 = p[stride];
    p[stride] = c * x2 + s * x1;
    *p-- = c * x1 + ms * x2;
  }
}
void orc_exp_rotation(float *X, int len, int dir, int stride, int K, int spread) {
  static const int factor_tab[3] = {15, 10, 5};
  int i, stride2 = 0, factor;
  float c, s, gain, theta;
  if (2 * K >= len || spread == 0) return;
  factor = factor_tab[spread - 1];
  gain = ((float)(1.f * len)) / ((float)(len + factor * K));
  theta = .5f * (gain * gain);
  c = (float)cos((.5f * 3.1415926535897931) * (theta));
  s = (float)cos((.5f * 3.1415926535897931) * (1.0f - theta));
  if (len >= 8 * stride) {
    stride2 = 1;
    while ((stride2 * stride2 + stride2) * stride + (stride >> 2) < len) stride2++;
  }
  len = (unsigned)len / (unsigned)stride;
  for (i = 0; i < stride; i++) {
    if (dir < 0) {
      if (stride2) rotate_pass(X + i * len, len, stride2, s, c);
      rotate_pass(X + i * len, len, 1, c, s);
    } else {
      rotate_pass(X + i * len, len, 1, c, -s);
      if (stride2) rotate_pass(X + i * len, len, stride2, s, -c);
    }
  }
}

/* op_pvq_search_c (celt/vq.c:205-374): greedy pulse search; X is overwritten by |X|. Returns yy. */
float orc_pvq_search(float *X, int *iy, int K, int N) {
  float y[176 + 4];
  int signx[176 + 4];
  int i, j, pulsesLeft;
  float sum = 0, xy, yy;
  j = 0;
  do {
    signx[j] = X[j] < 0;
    X[j] = (float)fabs(X[j]);
    iy[j] = 0;
    y[j] = 0;
  } while (++j < N);
  xy = yy = 0;
  pulsesLeft = K;
  if (K > (N >> 1)) {
    float rcp;
    j = 0; do sum += X[j]; while (++j < N);
    if (!(sum > 1e-15f && sum < 64)) {
      X[0] = 1.f;
      j = 1; do X[j] = 0; while (++j < N);
      sum = 1.f;
    }
    rcp = (K + 0.8f) * (1.f / (sum));
    j = 0;
    do {
      iy[j] = (int)floor(rcp * X[j]);
      y[j] = (float)iy[j];
      yy = yy + y[j] * y[j];
      xy = xy + X[j] * y[j];
      y[j] *= 2;
      pulsesLeft -= iy[j];
    } while (++j < N);
  }
  if (pulsesLeft > N + 3) {
    float tmp = (float)pulsesLeft;
    yy = yy + tmp * tmp;
    yy = yy + tmp * y[0];
    iy[0] += pulsesLeft;
    pulsesLeft = 0;
  }
  for (i = 0; i < pulsesLeft; i++) {
    float Rxy, Ryy, best_num, best_den;
    int best_id = 0;
    yy = yy + 1;
    Rxy = xy + X[0];
    Ryy = yy + y[0];
    Rxy = Rxy * Rxy;
    best_den = Ryy;
    best_num = Rxy;
    j = 1;
    do {
      Rxy = xy + X[j];
      Ryy = yy + y[j];
      Rxy = Rxy * Rxy;
      if (best_den * Rxy > Ryy * best_num) { best_den = Ryy; best_num = Rxy; best_id = j; }
    } while (++j < N);
    xy = xy + X[best_id];
    yy = yy + y[best_id];
    y[best_id] += 2;
    iy[best_id]++;
  }
  j = 0;
  do iy[j] = (iy[j] ^ -signx[j]) + signx[j]; while (++j < N);
  return yy;
}

/* ---- encoder front end of celt_encode_with_ec without the pitch pre-filter (complexity < 5 or pf disabled):
 *   celt_preemphasis (celt/celt_encoder.c:557-646, 48 kHz fast path) -> [overlap history | frame]
 *   compute_mdcts (:511-554) -> compute_band_energies (celt/bands.c:151-165) -> amp2Log2 (celt/quant_bands.c:553-572)
 *   -> normalise_bands (celt/bands.c:169-183)
 * state per channel: the pre-emphasis memory and the last `overlap` pre-emphasised samples (what the reference keeps at
 * the tail of prefilter_mem, celt_encoder.c:2014).  pcm is interleaved, +-1.0 full scale. */
typedef struct { float preemph_mem[2]; float overlap_mem[2][120]; } OrcEncAnalysisState;
int orc_enc_analysis_state_size(void) { return (int)sizeof(OrcEncAnalysisState); }
void orc_celt_enc_analysis(OrcEncAnalysisState *st, const float *pcm, int CC, int LM, int shortBlocks, float *X, float *bandE, float *bandLogE) {
  const int N = 120 << LM, overlap = 120, M = 1 << LM;
  float in[960 + 120], freq[960];
  int c, i, j;
  for (c = 0; c < CC; c++) {
    float m = st->preemph_mem[c];
    memcpy(in, st->overlap_mem[c], overlap * sizeof(float));
    for (i = 0; i < N; i++) {
      float x = 32768.f * pcm[CC * i + c];
      in[overlap + i] = x - m;
      m = 0.85000610f * x;   /* mode->preemph[0] (celt/static_modes_float.h) */
    }
    st->preemph_mem[c] = m;
    memcpy(st->overlap_mem[c], in + N, overlap * sizeof(float));
    orc_mdct_forward(in, freq, LM, shortBlocks);
    for (i = 0; i < 21; i++) {
      float sum = 0;
      for (j = M * ob_ebands[i]; j < M * ob_ebands[i + 1]; j++) sum = sum + freq[j] * freq[j];   /* celt_inner_prod_c */
      sum = 1e-27f + sum;
      bandE[c * 21 + i] = (float)sqrt(sum);
      bandLogE[c * 21 + i] = (float)(1.442695040888963387 * log(bandE[c * 21 + i])) - ob_emeans[i];
      {
        float g = 1.f / (1e-27f + bandE[c * 21 + i]);
        for (j = M * ob_ebands[i]; j < M * ob_ebands[i + 1]; j++) X[c * N + j] = freq[j] * g;
      }
    }
  }
}
