/* oracle/celt_dec_oracle.c — TEST INFRASTRUCTURE ONLY (the checker; never the product path).
 *
 * Plain scalar C restatement of the reference's CELT decode hot path for the 48 kHz mode
 * (frames of 120/240/480/960 samples, overlap 120, 21 bands, float build):
 *   range decoder, Laplace/coarse+fine energy, TF/spread/dynalloc/trim, bit allocation,
 *   the band loop (theta splits, PVQ index -> pulses, spreading rotation, folding,
 *   anti-collapse), denormalisation, inverse MDCT, comb post-filter, de-emphasis,
 *   plus the CELT-only subset of the Opus packet layer (TOC + frame packing).
 * Every function cites the reference file:line it follows.  Float expressions keep the
 * reference's operation order (compile with -ffp-contract=off) so that the result can be
 * compared sample-for-sample with the reference's pure-C build (oracle/_ref/libopus_ref_c.so).
 *
 * Parity pin: tests/test_oracle_vs_ref.py (bit-exact final range, PCM identical to ref_c)
 * and the committed fixtures in tests/golden/ generated from ref_c.
 *
 * Not restated (out of scope this round): packet-loss concealment, SILK/hybrid,
 * sample rates other than 48 kHz, QEXT/DRED extensions, soft clipping.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OB_TABLE static const
#include "../opus_b200/csrc/celt_tables_data.h"

#define NBANDS 21
#define OVERLAP 120
#define SHORT_N 120
#define MAXN 960
#define DEC_BUF 2048
#define BITRES 3
#define MAX_FINE_BITS 8
#define FINE_OFFSET 21
#define QTHETA_OFFSET 4
#define QTHETA_OFFSET_TWOPHASE 16
#define ALLOC_STEPS 6
#define LOG_MAX_PSEUDO 6
#define SPREAD_NONE 0
#define SPREAD_NORMAL 2
#define SPREAD_AGGRESSIVE 3
#define COMB_MINPERIOD 15
#define PI_D 3.1415926535897931

#define IMIN(a, b) ((a) < (b) ? (a) : (b))
#define IMAX(a, b) ((a) > (b) ? (a) : (b))

/* ------------------------------------------------------------------------------------------
 * Range decoder  (celt/entdec.c:91-266, celt/entcode.c:69-88, celt/entcode.h:116-118)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const uint8_t *buf;
  uint32_t storage, end_offs, end_window;
  int nend_bits, nbits_total;
  uint32_t offs, rng, val, ext;
  int rem, error;
} RangeDec;

static int ilog32(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }

static int rd_next_byte(RangeDec *d) { return d->offs < d->storage ? d->buf[d->offs++] : 0; }
static int rd_next_byte_from_end(RangeDec *d) {
  return d->end_offs < d->storage ? d->buf[d->storage - ++(d->end_offs)] : 0;
}

/* entdec.c:102-117 */
static void rd_renorm(RangeDec *d) {
  while (d->rng <= (1u << 23)) {
    int sym;
    d->nbits_total += 8;
    d->rng <<= 8;
    sym = d->rem;
    d->rem = rd_next_byte(d);
    sym = (sym << 8 | d->rem) >> 1;
    d->val = ((d->val << 8) + (255 & ~sym)) & 0x7FFFFFFFu;
  }
}

/* entdec.c:119-137 */
static void rd_init(RangeDec *d, const uint8_t *buf, uint32_t len) {
  d->buf = buf; d->storage = len; d->end_offs = 0; d->end_window = 0; d->nend_bits = 0;
  d->nbits_total = 9; d->offs = 0; d->rng = 128; d->ext = 0;
  d->rem = rd_next_byte(d);
  d->val = d->rng - 1 - (d->rem >> 1);
  d->error = 0;
  rd_renorm(d);
}

static int rd_tell(const RangeDec *d) { return d->nbits_total - ilog32(d->rng); }

/* entcode.c:69-88 (the table-corrected variant) */
static uint32_t rd_tell_frac(const RangeDec *d) {
  static const unsigned corr[8] = {35733, 38967, 42495, 46340, 50535, 55109, 60097, 65535};
  uint32_t nbits = (uint32_t)d->nbits_total << BITRES;
  int l = ilog32(d->rng);
  uint32_t r = d->rng >> (l - 16);
  unsigned b = (r >> 12) - 8;
  b += r > corr[b];
  l = (l << 3) + b;
  return nbits - l;
}

/* entdec.c:139-144 */
static unsigned rd_decode(RangeDec *d, unsigned ft) {
  unsigned s;
  d->ext = d->rng / ft;
  s = (unsigned)(d->val / d->ext);
  return ft - IMIN(s + 1, ft);
}
/* entdec.c:146-151 */
static unsigned rd_decode_bin(RangeDec *d, unsigned bits) {
  unsigned s;
  d->ext = d->rng >> bits;
  s = (unsigned)(d->val / d->ext);
  return (1u << bits) - IMIN(s + 1u, 1u << bits);
}
/* entdec.c:153-160 */
static void rd_update(RangeDec *d, unsigned fl, unsigned fh, unsigned ft) {
  uint32_t s = d->ext * (ft - fh);
  d->val -= s;
  d->rng = fl > 0 ? d->ext * (fh - fl) : d->rng - s;
  rd_renorm(d);
}
/* entdec.c:162-177 */
static int rd_bit_logp(RangeDec *d, unsigned logp) {
  uint32_t r = d->rng, v = d->val, s = r >> logp;
  int ret = v < s;
  if (!ret) d->val = v - s;
  d->rng = ret ? s : r - s;
  rd_renorm(d);
  return ret;
}
/* entdec.c:179-198 */
static int rd_icdf(RangeDec *d, const unsigned char *icdf, unsigned ftb) {
  uint32_t s = d->rng, v = d->val, r = s >> ftb, t;
  int ret = -1;
  do { t = s; s = r * icdf[++ret]; } while (v < s);
  d->val = v - s;
  d->rng = t - s;
  rd_renorm(d);
  return ret;
}
/* entdec.c:246-266 */
static uint32_t rd_bits(RangeDec *d, unsigned bits) {
  uint32_t window = d->end_window, ret;
  int avail = d->nend_bits;
  if ((unsigned)avail < bits) {
    do { window |= (uint32_t)rd_next_byte_from_end(d) << avail; avail += 8; } while (avail <= 32 - 8);
  }
  ret = window & (((uint32_t)1 << bits) - 1u);
  window >>= bits; avail -= bits;
  d->end_window = window; d->nend_bits = avail; d->nbits_total += bits;
  return ret;
}
/* entdec.c:219-244 */
static uint32_t rd_uint(RangeDec *d, uint32_t ft) {
  unsigned s;
  int ftb;
  ft--;
  ftb = ilog32(ft);
  if (ftb > 8) {
    uint32_t t; unsigned f;
    ftb -= 8;
    f = (unsigned)(ft >> ftb) + 1;
    s = rd_decode(d, f);
    rd_update(d, s, s + 1, f);
    t = (uint32_t)s << ftb | rd_bits(d, ftb);
    if (t <= ft) return t;
    d->error = 1;
    return ft;
  } else {
    ft++;
    s = rd_decode(d, (unsigned)ft);
    rd_update(d, s, s + 1, (unsigned)ft);
    return s;
  }
}

/* celt/laplace.c:44-50, 94-134 */
static int rd_laplace(RangeDec *d, unsigned fs, int decay) {
  int val = 0;
  unsigned fl = 0, fm = rd_decode_bin(d, 15);
  if (fm >= fs) {
    unsigned ft = 32768 - (1 * (2 * 16)) - fs; /* LAPLACE_MINP=1, LAPLACE_NMIN=16 */
    val++;
    fl = fs;
    fs = ((ft * (int32_t)(16384 - decay)) >> 15) + 1;
    while (fs > 1 && fm >= fl + 2 * fs) {
      fs *= 2;
      fl += fs;
      fs = ((fs - 2) * (int32_t)decay) >> 15;
      fs += 1;
      val++;
    }
    if (fs <= 1) {
      int di = (fm - fl) >> 1;
      val += di;
      fl += 2 * di;
    }
    if (fm < fl + fs) val = -val; else fl += fs;
  }
  rd_update(d, fl, IMIN(fl + fs, 32768), 32768);
  return val;
}

#include "celt_rangeenc.inc"

/* ------------------------------------------------------------------------------------------
 * Small exact integer helpers (celt/bands.c:68-91, celt/mathops.c:45-68, celt/rate.h:48-87)
 * ---------------------------------------------------------------------------------------- */
static int frac_mul16(int a, int b) { return (16384 + ((int32_t)(int16_t)a * (int16_t)b)) >> 15; }

static int bitexact_cos(int x) {
  int32_t tmp = (4096 + ((int32_t)x * x)) >> 13;
  int x2 = tmp;
  x2 = (32767 - x2) + frac_mul16(x2, (-7651 + frac_mul16(x2, (8277 + frac_mul16(-626, x2)))));
  return 1 + x2;
}
static int bitexact_log2tan(int isin, int icos) {
  int lc = ilog32(icos), ls = ilog32(isin);
  icos <<= 15 - lc;
  isin <<= 15 - ls;
  return (ls - lc) * (1 << 11) + frac_mul16(isin, frac_mul16(isin, -2597) + 7932) -
         frac_mul16(icos, frac_mul16(icos, -2597) + 7932);
}
static unsigned isqrt_u32(uint32_t v) { /* floor(sqrt(v)) */
  unsigned g = 0;
  int bshift = (ilog32(v) - 1) >> 1;
  unsigned b = 1u << bshift;
  do {
    uint32_t t = (((uint32_t)g << 1) + b) << bshift;
    if (t <= v) { g += b; v -= t; }
    b >>= 1; bshift--;
  } while (bshift >= 0);
  return g;
}
static int pseudo_to_pulses(int i) { return i < 8 ? i : (8 + (i & 7)) << ((i >> 3) - 1); }

static const unsigned char *pulse_cache_row(int band, int LM) { /* LM in -1..3 */
  return ob_cache_bits + ob_cache_index[(LM + 1) * NBANDS + band];
}
static int bits_to_pseudo(int band, int LM, int bits) { /* rate.h:53-78 */
  const unsigned char *cache = pulse_cache_row(band, LM);
  int lo = 0, hi = cache[0], i;
  bits--;
  for (i = 0; i < LOG_MAX_PSEUDO; i++) {
    int mid = (lo + hi + 1) >> 1;
    if ((int)cache[mid] >= bits) hi = mid; else lo = mid;
  }
  if (bits - (lo == 0 ? -1 : (int)cache[lo]) <= (int)cache[hi] - bits) return lo;
  return hi;
}
static int pseudo_to_bits(int band, int LM, int q) { /* rate.h:80-87 */
  return q == 0 ? 0 : pulse_cache_row(band, LM)[q] + 1;
}

/* PVQ codebook sizes, U(n,k) symmetric, rows stored by min(n,k)  (celt/cwrs.c:197-211) */
static uint32_t pvq_u(int n, int k) {
  int a = IMIN(n, k), b = IMAX(n, k);
  return ob_pvq_u[ob_pvq_rowoff[a] + b];
}
static uint32_t pvq_v(int n, int k) { return pvq_u(n, k) + pvq_u(n, k + 1); }

/* index -> pulse vector; returns sum of squares  (celt/cwrs.c:467-541) */
static float pvq_index_to_pulses(int n, int k, uint32_t i, int *y) {
  uint32_t p;
  int s, k0;
  int16_t val;
  float yy = 0;
  while (n > 2) {
    uint32_t q;
    if (k >= n) { /* many pulses */
      p = pvq_u(n, k + 1);
      s = -(i >= p);
      i -= p & s;
      k0 = k;
      q = pvq_u(n, n);
      if (q > i) {
        k = n;
        do p = pvq_u(--k, n); while (p > i);
      } else {
        for (p = pvq_u(n, k); p > i; p = pvq_u(n, k)) k--;
      }
      i -= p;
      val = (k0 - k + s) ^ s;
      *y++ = val;
      yy = yy + (float)val * (float)val;
    } else { /* many dimensions */
      p = pvq_u(k, n);
      q = pvq_u(k + 1, n);
      if (p <= i && i < q) {
        i -= p;
        *y++ = 0;
      } else {
        s = -(i >= q);
        i -= q & s;
        k0 = k;
        do p = pvq_u(--k, n); while (p > i);
        i -= p;
        val = (k0 - k + s) ^ s;
        *y++ = val;
        yy = yy + (float)val * (float)val;
      }
    }
    n--;
  }
  /* n == 2 */
  p = 2 * k + 1;
  s = -(i >= p);
  i -= p & s;
  k0 = k;
  k = (i + 1) >> 1;
  if (k) i -= 2 * k - 1;
  val = (k0 - k + s) ^ s;
  *y++ = val;
  yy = yy + (float)val * (float)val;
  /* n == 1 */
  s = -(int)i;
  val = (k + s) ^ s;
  *y = val;
  yy = yy + (float)val * (float)val;
  return yy;
}

/* ------------------------------------------------------------------------------------------
 * Float helpers with the reference's rounding behaviour (celt/mathops.h:225-233,346-347)
 * ---------------------------------------------------------------------------------------- */
static float f_exp2(float x) { return (float)exp(0.6931471805599453094 * (x)); }
static float f_rsqrt(float x) { return 1.f / ((float)sqrt(x)); }
static float f_cos_norm(float x) { return (float)cos((.5f * PI_D) * (x)); }

/* ------------------------------------------------------------------------------------------
 * Vector primitives of the band decoder
 * ---------------------------------------------------------------------------------------- */
static float dot_seq(const float *x, const float *y, int n) { /* celt/pitch.h celt_inner_prod_c */
  float s = 0; int i;
  for (i = 0; i < n; i++) s = s + x[i] * y[i];
  return s;
}

/* celt/vq.c:75-102 */
static void rotate_pass(float *X, int len, int stride, float c, float s) {
  int i;
  float ms = -s;
  float *p = X;
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
/* celt/vq.c:104-147 */
static void spread_rotation(float *X, int len, int dir, int stride, int K, int spread) {
  static const int factor_tab[3] = {15, 10, 5};
  int i, stride2 = 0, factor;
  float c, s, gain, theta;
  if (2 * K >= len || spread == SPREAD_NONE) return;
  factor = factor_tab[spread - 1];
  gain = ((float)(1.f * len)) / ((float)(len + factor * K));
  theta = .5f * (gain * gain);
  c = f_cos_norm(theta);
  s = f_cos_norm(1.0f - theta);
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
/* celt/vq.c:150-181 */
static void pulses_to_unit(const int *iy, float *X, int N, float Ryy, float gain) {
  float g = f_rsqrt(Ryy) * gain;
  int i = 0;
  do X[i] = (float)iy[i] * g; while (++i < N);
}
/* celt/vq.c:183-203 */
static unsigned collapse_mask_of(const int *iy, int N, int B) {
  unsigned mask = 0;
  int N0, i;
  if (B <= 1) return 1;
  N0 = (unsigned)N / (unsigned)B;
  i = 0;
  do {
    int j = 0; unsigned tmp = 0;
    do tmp |= iy[i * N0 + j]; while (++j < N0);
    mask |= (tmp != 0) << i;
  } while (++i < B);
  return mask;
}
/* celt/vq.c:695-722 */
static void renormalise(float *X, int N, float gain) {
  float E = 1e-15f + dot_seq(X, X, N);
  float g = f_rsqrt(E) * gain;
  int i;
  for (i = 0; i < N; i++) X[i] = g * X[i];
}
/* celt/bands.c:623-636 */
static void haar_step(float *X, int N0, int stride) {
  int i, j;
  N0 >>= 1;
  for (i = 0; i < stride; i++)
    for (j = 0; j < N0; j++) {
      float t1 = .70710678f * X[stride * 2 * j + i];
      float t2 = .70710678f * X[stride * (2 * j + 1) + i];
      X[stride * 2 * j + i] = t1 + t2;
      X[stride * (2 * j + 1) + i] = t1 - t2;
    }
}
/* Hadamard ordering: natural order -> sequency order with DC last (celt/bands.c:564-572) */
static const int hadamard_order[30] = {1, 0, 3, 0, 2, 1, 7, 0, 4, 3, 6, 1, 5, 2,
                                       15, 0, 8, 7, 12, 3, 11, 4, 14, 1, 9, 6, 13, 2, 10, 5};
/* celt/bands.c:574-597 */
static void deinterleave(float *X, int N0, int stride, int hadamard) {
  float tmp[176];
  int i, j, N = N0 * stride;
  if (hadamard) {
    const int *ord = hadamard_order + stride - 2;
    for (i = 0; i < stride; i++)
      for (j = 0; j < N0; j++) tmp[ord[i] * N0 + j] = X[j * stride + i];
  } else {
    for (i = 0; i < stride; i++)
      for (j = 0; j < N0; j++) tmp[i * N0 + j] = X[j * stride + i];
  }
  memcpy(X, tmp, N * sizeof(float));
}
/* celt/bands.c:599-621 */
static void interleave(float *X, int N0, int stride, int hadamard) {
  float tmp[176];
  int i, j, N = N0 * stride;
  if (hadamard) {
    const int *ord = hadamard_order + stride - 2;
    for (i = 0; i < stride; i++)
      for (j = 0; j < N0; j++) tmp[j * stride + i] = X[ord[i] * N0 + j];
  } else {
    for (i = 0; i < stride; i++)
      for (j = 0; j < N0; j++) tmp[j * stride + i] = X[i * N0 + j];
  }
  memcpy(X, tmp, N * sizeof(float));
}
/* celt/bands.c:638-662 */
static int theta_levels(int N, int b, int offset, int pulse_cap, int stereo) {
  static const int16_t exp2_table8[8] = {16384, 17866, 19483, 21247, 23170, 25267, 27554, 30048};
  int qn, qb, N2 = 2 * N - 1;
  if (stereo && N == 2) N2--;
  qb = (b + N2 * offset) / N2; /* celt_sudiv: C division (both signs) */
  qb = IMIN(b - pulse_cap - (4 << BITRES), qb);
  qb = IMIN(8 << BITRES, qb);
  if (qb < (1 << BITRES >> 1)) qn = 1;
  else {
    qn = exp2_table8[qb & 0x7] >> (14 - (qb >> BITRES));
    qn = (qn + 1) >> 1 << 1;
  }
  return qn;
}

/* ------------------------------------------------------------------------------------------
 * Band decoder context (decode-only view of celt/bands.c:664-698)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  RangeDec *ec;
  int band, intensity, spread, tf_change, disable_inv;
  int32_t remaining_bits;
  uint32_t seed;
} BandCtx;

typedef struct { int inv, imid, iside, delta, itheta, qalloc; } SplitInfo;

/* decode side of celt/bands.c:700-933 */
static void decode_theta(BandCtx *ctx, SplitInfo *sp, int N, int *b, int B, int B0, int LM, int stereo, int *fill) {
  int qn, itheta = 0, delta, imid, iside, qalloc, pulse_cap, offset, inv = 0;
  int32_t tell;
  RangeDec *ec = ctx->ec;
  int i = ctx->band;
  pulse_cap = ob_logN[i] + LM * (1 << BITRES);
  offset = (pulse_cap >> 1) - (stereo && N == 2 ? QTHETA_OFFSET_TWOPHASE : QTHETA_OFFSET);
  qn = theta_levels(N, *b, offset, pulse_cap, stereo);
  if (stereo && i >= ctx->intensity) qn = 1;
  tell = rd_tell_frac(ec);
  if (qn != 1) {
    if (stereo && N > 2) { /* step pdf */
      int p0 = 3, x, x0 = qn / 2, ft = p0 * (x0 + 1) + x0, fs;
      fs = rd_decode(ec, ft);
      if (fs < (x0 + 1) * p0) x = fs / p0; else x = x0 + 1 + (fs - (x0 + 1) * p0);
      rd_update(ec, x <= x0 ? p0 * x : (x - 1 - x0) + (x0 + 1) * p0,
                x <= x0 ? p0 * (x + 1) : (x - x0) + (x0 + 1) * p0, ft);
      itheta = x;
    } else if (B0 > 1 || stereo) { /* uniform pdf */
      itheta = rd_uint(ec, qn + 1);
    } else { /* triangular pdf */
      int fs = 1, ft, fl = 0, fm;
      ft = ((qn >> 1) + 1) * ((qn >> 1) + 1);
      fm = rd_decode(ec, ft);
      if (fm < ((qn >> 1) * ((qn >> 1) + 1) >> 1)) {
        itheta = (isqrt_u32(8 * (uint32_t)fm + 1) - 1) >> 1;
        fs = itheta + 1;
        fl = itheta * (itheta + 1) >> 1;
      } else {
        itheta = (2 * (qn + 1) - isqrt_u32(8 * (uint32_t)(ft - fm - 1) + 1)) >> 1;
        fs = qn + 1 - itheta;
        fl = ft - ((qn + 1 - itheta) * (qn + 2 - itheta) >> 1);
      }
      rd_update(ec, fl, fl + fs, ft);
    }
    itheta = (uint32_t)((int32_t)itheta * 16384) / (uint32_t)qn;
  } else if (stereo) {
    if (*b > 2 << BITRES && ctx->remaining_bits > 2 << BITRES) inv = rd_bit_logp(ec, 2);
    else inv = 0;
    if (ctx->disable_inv) inv = 0;
    itheta = 0;
  }
  qalloc = rd_tell_frac(ec) - tell;
  *b -= qalloc;
  if (itheta == 0) {
    imid = 32767; iside = 0; *fill &= (1 << B) - 1; delta = -16384;
  } else if (itheta == 16384) {
    imid = 0; iside = 32767; *fill &= ((1 << B) - 1) << B; delta = 16384;
  } else {
    imid = bitexact_cos((int16_t)itheta);
    iside = bitexact_cos((int16_t)(16384 - itheta));
    delta = frac_mul16((N - 1) << 7, bitexact_log2tan(iside, imid));
  }
  sp->inv = inv; sp->imid = imid; sp->iside = iside; sp->delta = delta; sp->itheta = itheta; sp->qalloc = qalloc;
}

/* celt/bands.c:934-968 */
static unsigned decode_band_n1(BandCtx *ctx, float *X, float *Y, float *lowband_out) {
  int c, stereo = Y != NULL;
  float *x = X;
  c = 0;
  do {
    int sign = 0;
    if (ctx->remaining_bits >= 1 << BITRES) {
      sign = rd_bits(ctx->ec, 1);
      ctx->remaining_bits -= 1 << BITRES;
    }
    x[0] = sign ? -1.f : 1.f;
    x = Y;
  } while (++c < 1 + stereo);
  if (lowband_out) lowband_out[0] = X[0];
  return 1;
}

/* celt/vq.c:621-692 (alg_unquant) */
static unsigned pvq_decode(float *X, int N, int K, int spread, int B, RangeDec *ec, float gain) {
  int iy[176];
  float Ryy = pvq_index_to_pulses(N, K, rd_uint(ec, pvq_v(N, K)), iy);
  pulses_to_unit(iy, X, N, Ryy, gain);
  spread_rotation(X, N, -1, B, K, spread);
  return collapse_mask_of(iy, N, B);
}

/* decode side of celt/bands.c:973-1176 */
static unsigned decode_partition(BandCtx *ctx, float *X, int N, int b, int B, float *lowband, int LM, float gain, int fill) {
  const unsigned char *cache;
  int q, curr_bits, imid = 0, iside = 0, B0 = B;
  float mid = 0, side = 0;
  unsigned cm = 0;
  float *Y = NULL;
  int i = ctx->band;
  cache = pulse_cache_row(i, LM);
  if (LM != -1 && b > cache[cache[0]] + 12 && N > 2) {
    int mbits, sbits, delta, itheta, qalloc;
    SplitInfo sp;
    float *next_lowband2 = NULL;
    int32_t rebalance;
    N >>= 1;
    Y = X + N;
    LM -= 1;
    if (B == 1) fill = (fill & 1) | (fill << 1);
    B = (B + 1) >> 1;
    decode_theta(ctx, &sp, N, &b, B, B0, LM, 0, &fill);
    imid = sp.imid; iside = sp.iside; delta = sp.delta; itheta = sp.itheta; qalloc = sp.qalloc;
    mid = (1.f / 32768) * imid;
    side = (1.f / 32768) * iside;
    if (B0 > 1 && (itheta & 0x3fff)) {
      if (itheta > 8192) delta -= delta >> (4 - LM);
      else delta = IMIN(0, delta + (N << BITRES >> (5 - LM)));
    }
    mbits = IMAX(0, IMIN(b, (b - delta) / 2));
    sbits = b - mbits;
    ctx->remaining_bits -= qalloc;
    if (lowband) next_lowband2 = lowband + N;
    rebalance = ctx->remaining_bits;
    if (mbits >= sbits) {
      cm = decode_partition(ctx, X, N, mbits, B, lowband, LM, gain * mid, fill);
      rebalance = mbits - (rebalance - ctx->remaining_bits);
      if (rebalance > 3 << BITRES && itheta != 0) sbits += rebalance - (3 << BITRES);
      cm |= decode_partition(ctx, Y, N, sbits, B, next_lowband2, LM, gain * side, fill >> B) << (B0 >> 1);
    } else {
      cm = decode_partition(ctx, Y, N, sbits, B, next_lowband2, LM, gain * side, fill >> B) << (B0 >> 1);
      rebalance = sbits - (rebalance - ctx->remaining_bits);
      if (rebalance > 3 << BITRES && itheta != 16384) mbits += rebalance - (3 << BITRES);
      cm |= decode_partition(ctx, X, N, mbits, B, lowband, LM, gain * mid, fill);
    }
  } else {
    q = bits_to_pseudo(i, LM, b);
    curr_bits = pseudo_to_bits(i, LM, q);
    ctx->remaining_bits -= curr_bits;
    while (ctx->remaining_bits < 0 && q > 0) {
      ctx->remaining_bits += curr_bits;
      q--;
      curr_bits = pseudo_to_bits(i, LM, q);
      ctx->remaining_bits -= curr_bits;
    }
    if (q != 0) {
      cm = pvq_decode(X, N, pseudo_to_pulses(q), ctx->spread, B, ctx->ec, gain);
    } else {
      int j;
      unsigned cm_mask = (unsigned)(1UL << B) - 1;
      fill &= cm_mask;
      if (!fill) {
        memset(X, 0, N * sizeof(float));
      } else {
        if (lowband == NULL) { /* noise */
          for (j = 0; j < N; j++) {
            ctx->seed = 1664525u * ctx->seed + 1013904223u;
            X[j] = (float)((int32_t)ctx->seed >> 20);
          }
          cm = cm_mask;
        } else { /* folded spectrum */
          for (j = 0; j < N; j++) {
            float tmp;
            ctx->seed = 1664525u * ctx->seed + 1013904223u;
            tmp = 1.0f / 256;
            tmp = (ctx->seed) & 0x8000 ? tmp : -tmp;
            X[j] = lowband[j] + tmp;
          }
          cm = fill;
        }
        renormalise(X, N, gain);
      }
    }
  }
  return cm;
}

/* decode side of celt/bands.c:1248-1379 */
static unsigned decode_band_mono(BandCtx *ctx, float *X, int N, int b, int B, float *lowband, int LM,
                                 float *lowband_out, float gain, float *lowband_scratch, int fill) {
  static const unsigned char bit_interleave[16] = {0, 1, 1, 1, 2, 3, 3, 3, 2, 3, 3, 3, 2, 3, 3, 3};
  static const unsigned char bit_deinterleave[16] = {0x00, 0x03, 0x0C, 0x0F, 0x30, 0x33, 0x3C, 0x3F,
                                                     0xC0, 0xC3, 0xCC, 0xCF, 0xF0, 0xF3, 0xFC, 0xFF};
  int N0 = N, N_B = N, N_B0, B0 = B, time_divide = 0, recombine = 0, longBlocks, k;
  unsigned cm = 0;
  int tf_change = ctx->tf_change;
  longBlocks = B0 == 1;
  N_B = (unsigned)N_B / (unsigned)B;
  if (N == 1) return decode_band_n1(ctx, X, NULL, lowband_out);
  if (tf_change > 0) recombine = tf_change;
  if (lowband_scratch && lowband && (recombine || ((N_B & 1) == 0 && tf_change < 0) || B0 > 1)) {
    memcpy(lowband_scratch, lowband, N * sizeof(float));
    lowband = lowband_scratch;
  }
  for (k = 0; k < recombine; k++) {
    if (lowband) haar_step(lowband, N >> k, 1 << k);
    fill = bit_interleave[fill & 0xF] | bit_interleave[fill >> 4] << 2;
  }
  B >>= recombine;
  N_B <<= recombine;
  while ((N_B & 1) == 0 && tf_change < 0) {
    if (lowband) haar_step(lowband, N_B, B);
    fill |= fill << B;
    B <<= 1;
    N_B >>= 1;
    time_divide++;
    tf_change++;
  }
  B0 = B;
  N_B0 = N_B;
  if (B0 > 1 && lowband) deinterleave(lowband, N_B >> recombine, B0 << recombine, longBlocks);
  cm = decode_partition(ctx, X, N, b, B, lowband, LM, gain, fill);
  if (B0 > 1) interleave(X, N_B >> recombine, B0 << recombine, longBlocks);
  N_B = N_B0;
  B = B0;
  for (k = 0; k < time_divide; k++) {
    B >>= 1;
    N_B <<= 1;
    cm |= cm >> B;
    haar_step(X, N_B, B);
  }
  for (k = 0; k < recombine; k++) {
    cm = bit_deinterleave[cm];
    haar_step(X, N0 >> k, 1 << k);
  }
  B <<= recombine;
  if (lowband_out) {
    int j;
    float n = (float)sqrt((double)(float)N0);
    for (j = 0; j < N0; j++) lowband_out[j] = n * X[j];
  }
  cm &= (1 << B) - 1;
  return cm;
}

/* celt/bands.c:407-466 (stereo_merge) */
static void merge_mid_side(float *X, float *Y, float mid, int N) {
  int j;
  float xp, side, El, Er, lgain, rgain;
  xp = dot_seq(Y, X, N);
  side = dot_seq(Y, Y, N);
  xp = mid * xp;
  El = (mid * mid) + side - 2 * xp;
  Er = (mid * mid) + side + 2 * xp;
  if (Er < 6e-4f || El < 6e-4f) {
    memcpy(Y, X, N * sizeof(float));
    return;
  }
  lgain = f_rsqrt(El);
  rgain = f_rsqrt(Er);
  for (j = 0; j < N; j++) {
    float l = mid * X[j], r = Y[j];
    X[j] = lgain * (l - r);
    Y[j] = rgain * (l + r);
  }
}

/* decode side of celt/bands.c:1387-1570 */
static unsigned decode_band_stereo(BandCtx *ctx, float *X, float *Y, int N, int b, int B, float *lowband, int LM,
                                   float *lowband_out, float *lowband_scratch, int fill) {
  int imid, iside, inv, mbits, sbits, delta, itheta, qalloc, orig_fill;
  float mid, side;
  unsigned cm = 0;
  SplitInfo sp;
  if (N == 1) return decode_band_n1(ctx, X, Y, lowband_out);
  orig_fill = fill;
  decode_theta(ctx, &sp, N, &b, B, B, LM, 1, &fill);
  inv = sp.inv; imid = sp.imid; iside = sp.iside; delta = sp.delta; itheta = sp.itheta; qalloc = sp.qalloc;
  mid = (1.f / 32768) * imid;
  side = (1.f / 32768) * iside;
  if (N == 2) {
    int c, sign = 0;
    float *x2, *y2;
    mbits = b;
    sbits = 0;
    if (itheta != 0 && itheta != 16384) sbits = 1 << BITRES;
    mbits -= sbits;
    c = itheta > 8192;
    ctx->remaining_bits -= qalloc + sbits;
    x2 = c ? Y : X;
    y2 = c ? X : Y;
    if (sbits) sign = rd_bits(ctx->ec, 1);
    sign = 1 - 2 * sign;
    cm = decode_band_mono(ctx, x2, N, mbits, B, lowband, LM, lowband_out, 1.0f, lowband_scratch, orig_fill);
    y2[0] = -sign * x2[1];
    y2[1] = sign * x2[0];
    {
      float tmp;
      X[0] = mid * X[0];
      X[1] = mid * X[1];
      Y[0] = side * Y[0];
      Y[1] = side * Y[1];
      tmp = X[0];
      X[0] = tmp - Y[0];
      Y[0] = tmp + Y[0];
      tmp = X[1];
      X[1] = tmp - Y[1];
      Y[1] = tmp + Y[1];
    }
  } else {
    int32_t rebalance;
    mbits = IMAX(0, IMIN(b, (b - delta) / 2));
    sbits = b - mbits;
    ctx->remaining_bits -= qalloc;
    rebalance = ctx->remaining_bits;
    if (mbits >= sbits) {
      cm = decode_band_mono(ctx, X, N, mbits, B, lowband, LM, lowband_out, 1.0f, lowband_scratch, fill);
      rebalance = mbits - (rebalance - ctx->remaining_bits);
      if (rebalance > 3 << BITRES && itheta != 0) sbits += rebalance - (3 << BITRES);
      cm |= decode_band_mono(ctx, Y, N, sbits, B, NULL, LM, NULL, side, NULL, fill >> B);
    } else {
      cm = decode_band_mono(ctx, Y, N, sbits, B, NULL, LM, NULL, side, NULL, fill >> B);
      rebalance = sbits - (rebalance - ctx->remaining_bits);
      if (rebalance > 3 << BITRES && itheta != 16384) mbits += rebalance - (3 << BITRES);
      cm |= decode_band_mono(ctx, X, N, mbits, B, lowband, LM, lowband_out, 1.0f, lowband_scratch, fill);
    }
  }
  if (N != 2) merge_mid_side(X, Y, mid, N);
  if (inv) {
    int j;
    for (j = 0; j < N; j++) Y[j] = -Y[j];
  }
  return cm;
}

/* decode side of celt/bands.c:1589-1923 */
static void decode_all_bands(int start, int end, float *X_, float *Y_, unsigned char *collapse_masks,
                             const int *pulses, int shortBlocks, int spread, int dual_stereo, int intensity,
                             const int *tf_res, int32_t total_bits, int32_t balance, RangeDec *ec, int LM,
                             int codedBands, uint32_t *seed, int disable_inv) {
  int i, B, M = 1 << LM, lowband_offset = 0, update_lowband = 1, C = Y_ != NULL ? 2 : 1;
  int norm_offset = M * ob_ebands[start];
  int normlen = M * ob_ebands[NBANDS - 1] - norm_offset;
  float normbuf[2 * 8 * 78], scratch_buf[176];
  float *norm = normbuf, *norm2 = normbuf + normlen, *lowband_scratch;
  BandCtx ctx;
  B = shortBlocks ? M : 1;
  ctx.ec = ec; ctx.intensity = intensity; ctx.seed = *seed; ctx.spread = spread; ctx.disable_inv = disable_inv;
  memset(normbuf, 0, sizeof(normbuf));
  for (i = start; i < end; i++) {
    int32_t tell, curr_balance, remaining_bits;
    int b, N, effective_lowband = -1, tf_change, last;
    float *X, *Y;
    unsigned x_cm, y_cm;
    ctx.band = i;
    last = (i == end - 1);
    X = X_ + M * ob_ebands[i];
    Y = Y_ != NULL ? Y_ + M * ob_ebands[i] : NULL;
    N = M * ob_ebands[i + 1] - M * ob_ebands[i];
    tell = rd_tell_frac(ec);
    if (i != start) balance -= tell;
    remaining_bits = total_bits - tell - 1;
    ctx.remaining_bits = remaining_bits;
    if (i <= codedBands - 1) {
      curr_balance = balance / IMIN(3, codedBands - i);
      b = IMAX(0, IMIN(16383, IMIN(remaining_bits + 1, pulses[i] + curr_balance)));
    } else b = 0;
    if ((M * ob_ebands[i] - N >= M * ob_ebands[start] || i == start + 1) && (update_lowband || lowband_offset == 0))
      lowband_offset = i;
    if (i == start + 1) { /* special_hybrid_folding, bands.c:1572-1586 (a no-op copy for start==0) */
      int n1 = M * (ob_ebands[start + 1] - ob_ebands[start]);
      int n2 = M * (ob_ebands[start + 2] - ob_ebands[start + 1]);
      if (n2 > n1) {
        memmove(&norm[n1], &norm[2 * n1 - n2], (n2 - n1) * sizeof(float));
        if (dual_stereo) memmove(&norm2[n1], &norm2[2 * n1 - n2], (n2 - n1) * sizeof(float));
      }
    }
    tf_change = tf_res[i];
    ctx.tf_change = tf_change;
    lowband_scratch = last ? NULL : scratch_buf;
    if (lowband_offset != 0 && (spread != SPREAD_AGGRESSIVE || B > 1 || tf_change < 0)) {
      int fold_start, fold_end, fold_i;
      effective_lowband = IMAX(0, M * ob_ebands[lowband_offset] - norm_offset - N);
      fold_start = lowband_offset;
      while (M * ob_ebands[--fold_start] > effective_lowband + norm_offset);
      fold_end = lowband_offset - 1;
      while (++fold_end < i && M * ob_ebands[fold_end] < effective_lowband + norm_offset + N);
      x_cm = y_cm = 0;
      fold_i = fold_start;
      do {
        x_cm |= collapse_masks[fold_i * C + 0];
        y_cm |= collapse_masks[fold_i * C + C - 1];
      } while (++fold_i < fold_end);
    } else x_cm = y_cm = (1 << B) - 1;
    if (dual_stereo && i == intensity) {
      int j;
      dual_stereo = 0;
      for (j = 0; j < M * ob_ebands[i] - norm_offset; j++) norm[j] = .5f * (norm[j] + norm2[j]);
    }
    if (dual_stereo) {
      x_cm = decode_band_mono(&ctx, X, N, b / 2, B, effective_lowband != -1 ? norm + effective_lowband : NULL, LM,
                              last ? NULL : norm + M * ob_ebands[i] - norm_offset, 1.0f, lowband_scratch, x_cm);
      y_cm = decode_band_mono(&ctx, Y, N, b / 2, B, effective_lowband != -1 ? norm2 + effective_lowband : NULL, LM,
                              last ? NULL : norm2 + M * ob_ebands[i] - norm_offset, 1.0f, lowband_scratch, y_cm);
    } else {
      if (Y != NULL) {
        x_cm = decode_band_stereo(&ctx, X, Y, N, b, B, effective_lowband != -1 ? norm + effective_lowband : NULL, LM,
                                  last ? NULL : norm + M * ob_ebands[i] - norm_offset, lowband_scratch, x_cm | y_cm);
      } else {
        x_cm = decode_band_mono(&ctx, X, N, b, B, effective_lowband != -1 ? norm + effective_lowband : NULL, LM,
                                last ? NULL : norm + M * ob_ebands[i] - norm_offset, 1.0f, lowband_scratch, x_cm | y_cm);
      }
      y_cm = x_cm;
    }
    collapse_masks[i * C + 0] = (unsigned char)x_cm;
    collapse_masks[i * C + C - 1] = (unsigned char)y_cm;
    balance += pulses[i] + tell;
    update_lowband = b > (N << BITRES);
  }
  *seed = ctx.seed;
}

/* ------------------------------------------------------------------------------------------
 * Bit allocation (celt/rate.c:249-533, 535-646; celt/celt.c:329-338)
 * ---------------------------------------------------------------------------------------- */
static int interp_alloc(int start, int end, int skip_start, const int *bits1, const int *bits2, const int *thresh,
                        const int *cap, int32_t total, int32_t *balance_out, int skip_rsv, int *intensity,
                        int intensity_rsv, int *dual_stereo, int dual_stereo_rsv, int *bits, int *ebits,
                        int *fine_priority, int C, int LM, RangeDec *ec, AllocEnc *ae) {
  int32_t psum, left, percoeff, balance;
  int lo = 0, hi = 1 << ALLOC_STEPS, i, j, logM = LM << BITRES, stereo = C > 1, codedBands, alloc_floor = C << BITRES, done;
  for (i = 0; i < ALLOC_STEPS; i++) {
    int mid = (lo + hi) >> 1;
    psum = 0; done = 0;
    for (j = end; j-- > start;) {
      int tmp = bits1[j] + (mid * (int32_t)bits2[j] >> ALLOC_STEPS);
      if (tmp >= thresh[j] || done) { done = 1; psum += IMIN(tmp, cap[j]); }
      else if (tmp >= alloc_floor) psum += alloc_floor;
    }
    if (psum > total) hi = mid; else lo = mid;
  }
  psum = 0; done = 0;
  for (j = end; j-- > start;) {
    int tmp = bits1[j] + ((int32_t)lo * bits2[j] >> ALLOC_STEPS);
    if (tmp < thresh[j] && !done) tmp = tmp >= alloc_floor ? alloc_floor : 0;
    else done = 1;
    tmp = IMIN(tmp, cap[j]);
    bits[j] = tmp;
    psum += tmp;
  }
  for (codedBands = end;; codedBands--) {
    int band_width, band_bits, rem;
    j = codedBands - 1;
    if (j <= skip_start) { total += skip_rsv; break; }
    left = total - psum;
    percoeff = (uint32_t)left / (uint32_t)(ob_ebands[codedBands] - ob_ebands[start]);
    left -= (ob_ebands[codedBands] - ob_ebands[start]) * percoeff;
    rem = IMAX(left - (ob_ebands[j] - ob_ebands[start]), 0);
    band_width = ob_ebands[codedBands] - ob_ebands[j];
    band_bits = (int)(bits[j] + percoeff * band_width + rem);
    if (band_bits >= IMAX(thresh[j], alloc_floor + (1 << BITRES))) {
      if (ae) { /* encoder: the only free decision of the allocator (celt/rate.c:346-369) */
        int depth_threshold = codedBands > 17 ? (j < ae->prev ? 7 : 9) : 0;
        if (codedBands <= start + 2 || (band_bits > (depth_threshold * band_width << LM << BITRES) >> 4 && j <= ae->signalBandwidth)) {
          re_bit_logp(ae->enc, 1, 1);
          break;
        }
        re_bit_logp(ae->enc, 0, 1);
      } else if (rd_bit_logp(ec, 1)) break;
      psum += 1 << BITRES;
      band_bits -= 1 << BITRES;
    }
    psum -= bits[j] + intensity_rsv;
    if (intensity_rsv > 0) intensity_rsv = ob_log2_frac_tab[j - start];
    psum += intensity_rsv;
    if (band_bits >= alloc_floor) { psum += alloc_floor; bits[j] = alloc_floor; }
    else bits[j] = 0;
  }
  if (intensity_rsv > 0) {
    if (ae) {
      *intensity = IMIN(*intensity, codedBands);
      re_uint(ae->enc, *intensity - start, codedBands + 1 - start);
    } else *intensity = start + rd_uint(ec, codedBands + 1 - start);
  } else *intensity = 0;
  if (*intensity <= start) { total += dual_stereo_rsv; dual_stereo_rsv = 0; }
  if (dual_stereo_rsv > 0) {
    if (ae) re_bit_logp(ae->enc, *dual_stereo, 1);
    else *dual_stereo = rd_bit_logp(ec, 1);
  } else *dual_stereo = 0;
  left = total - psum;
  percoeff = (uint32_t)left / (uint32_t)(ob_ebands[codedBands] - ob_ebands[start]);
  left -= (ob_ebands[codedBands] - ob_ebands[start]) * percoeff;
  for (j = start; j < codedBands; j++) bits[j] += ((int)percoeff * (ob_ebands[j + 1] - ob_ebands[j]));
  for (j = start; j < codedBands; j++) {
    int tmp = (int)IMIN(left, ob_ebands[j + 1] - ob_ebands[j]);
    bits[j] += tmp;
    left -= tmp;
  }
  balance = 0;
  for (j = start; j < codedBands; j++) {
    int N0 = ob_ebands[j + 1] - ob_ebands[j], N = N0 << LM, den, offset, NClogN;
    int32_t excess, bit = (int32_t)bits[j] + balance;
    if (N > 1) {
      excess = IMAX(bit - cap[j], 0);
      bits[j] = bit - excess;
      den = (C * N + ((C == 2 && N > 2 && !*dual_stereo && j < *intensity) ? 1 : 0));
      NClogN = den * (ob_logN[j] + logM);
      offset = (NClogN >> 1) - den * FINE_OFFSET;
      if (N == 2) offset += den << BITRES >> 2;
      if (bits[j] + offset < den * 2 << BITRES) offset += NClogN >> 2;
      else if (bits[j] + offset < den * 3 << BITRES) offset += NClogN >> 3;
      ebits[j] = IMAX(0, (bits[j] + offset + (den << (BITRES - 1))));
      ebits[j] = ((uint32_t)ebits[j] / (uint32_t)den) >> BITRES;
      if (C * ebits[j] > (bits[j] >> BITRES)) ebits[j] = bits[j] >> stereo >> BITRES;
      ebits[j] = IMIN(ebits[j], MAX_FINE_BITS);
      fine_priority[j] = ebits[j] * (den << BITRES) >= bits[j] + offset;
      bits[j] -= C * ebits[j] << BITRES;
    } else {
      excess = IMAX(0, bit - (C << BITRES));
      bits[j] = bit - excess;
      ebits[j] = 0;
      fine_priority[j] = 1;
    }
    if (excess > 0) {
      int extra_fine = IMIN(excess >> (stereo + BITRES), MAX_FINE_BITS - ebits[j]);
      int extra_bits;
      ebits[j] += extra_fine;
      extra_bits = extra_fine * C << BITRES;
      fine_priority[j] = extra_bits >= excess - balance;
      excess -= extra_bits;
    }
    balance = excess;
  }
  *balance_out = balance;
  for (; j < end; j++) {
    ebits[j] = bits[j] >> stereo >> BITRES;
    bits[j] = 0;
    fine_priority[j] = ebits[j] < 1;
  }
  return codedBands;
}

static int compute_allocation(int start, int end, const int *offsets, const int *cap, int alloc_trim, int *intensity,
                              int *dual_stereo, int32_t total, int32_t *balance, int *pulses, int *ebits,
                              int *fine_priority, int C, int LM, RangeDec *ec, AllocEnc *ae) {
  int lo, hi, len = NBANDS, j, skip_start = start, skip_rsv, intensity_rsv = 0, dual_stereo_rsv = 0;
  int bits1[NBANDS], bits2[NBANDS], thresh[NBANDS], trim_offset[NBANDS];
  const int nbAllocVectors = 11;
  total = IMAX(total, 0);
  skip_rsv = total >= 1 << BITRES ? 1 << BITRES : 0;
  total -= skip_rsv;
  if (C == 2) {
    intensity_rsv = ob_log2_frac_tab[end - start];
    if (intensity_rsv > total) intensity_rsv = 0;
    else {
      total -= intensity_rsv;
      dual_stereo_rsv = total >= 1 << BITRES ? 1 << BITRES : 0;
      total -= dual_stereo_rsv;
    }
  }
  for (j = start; j < end; j++) {
    int w = ob_ebands[j + 1] - ob_ebands[j];
    thresh[j] = IMAX((C) << BITRES, (3 * w << LM << BITRES) >> 4);
    trim_offset[j] = C * w * (alloc_trim - 5 - LM) * (end - j - 1) * (1 << (LM + BITRES)) >> 6;
    if (w << LM == 1) trim_offset[j] -= C << BITRES;
  }
  lo = 1;
  hi = nbAllocVectors - 1;
  do {
    int done = 0, psum = 0, mid = (lo + hi) >> 1;
    for (j = end; j-- > start;) {
      int N = ob_ebands[j + 1] - ob_ebands[j];
      int bitsj = C * N * ob_band_alloc[mid * len + j] << LM >> 2;
      if (bitsj > 0) bitsj = IMAX(0, bitsj + trim_offset[j]);
      bitsj += offsets[j];
      if (bitsj >= thresh[j] || done) { done = 1; psum += IMIN(bitsj, cap[j]); }
      else if (bitsj >= C << BITRES) psum += C << BITRES;
    }
    if (psum > total) hi = mid - 1; else lo = mid + 1;
  } while (lo <= hi);
  hi = lo--;
  for (j = start; j < end; j++) {
    int N = ob_ebands[j + 1] - ob_ebands[j];
    int bits1j = C * N * ob_band_alloc[lo * len + j] << LM >> 2;
    int bits2j = hi >= nbAllocVectors ? cap[j] : C * N * ob_band_alloc[hi * len + j] << LM >> 2;
    if (bits1j > 0) bits1j = IMAX(0, bits1j + trim_offset[j]);
    if (bits2j > 0) bits2j = IMAX(0, bits2j + trim_offset[j]);
    if (lo > 0) bits1j += offsets[j];
    bits2j += offsets[j];
    if (offsets[j] > 0) skip_start = j;
    bits2j = IMAX(0, bits2j - bits1j);
    bits1[j] = bits1j;
    bits2[j] = bits2j;
  }
  return interp_alloc(start, end, skip_start, bits1, bits2, thresh, cap, total, balance, skip_rsv, intensity,
                      intensity_rsv, dual_stereo, dual_stereo_rsv, pulses, ebits, fine_priority, C, LM, ec, ae);
}

/* ------------------------------------------------------------------------------------------
 * Energy decoding (celt/quant_bands.c:431-547)
 * ---------------------------------------------------------------------------------------- */
static const float pred_coef[4] = {29440 / 32768., 26112 / 32768., 21248 / 32768., 16384 / 32768.};
static const float beta_coef[4] = {30147 / 32768., 22282 / 32768., 12124 / 32768., 6554 / 32768.};
static const float beta_intra = 4915 / 32768.;

static void decode_coarse_energy(int start, int end, float *oldE, int intra, RangeDec *dec, int C, int LM) {
  const unsigned char *prob = ob_e_prob + (LM * 2 + intra) * 42;
  int i, c;
  float prev[2] = {0, 0}, coef, beta;
  int32_t budget, tell;
  if (intra) { coef = 0; beta = beta_intra; } else { beta = beta_coef[LM]; coef = pred_coef[LM]; }
  budget = dec->storage * 8;
  for (i = start; i < end; i++) {
    c = 0;
    do {
      int qi;
      float q, tmp;
      tell = rd_tell(dec);
      if (budget - tell >= 15) {
        int pi = 2 * IMIN(i, 20);
        qi = rd_laplace(dec, prob[pi] << 7, prob[pi + 1] << 6);
      } else if (budget - tell >= 2) {
        qi = rd_icdf(dec, ob_small_energy_icdf, 2);
        qi = (qi >> 1) ^ -(qi & 1);
      } else if (budget - tell >= 1) {
        qi = -rd_bit_logp(dec, 1);
      } else qi = -1;
      q = (float)qi;
      oldE[i + c * NBANDS] = oldE[i + c * NBANDS] > -9.f ? oldE[i + c * NBANDS] : -9.f;
      tmp = coef * oldE[i + c * NBANDS] + prev[c] + q;
      oldE[i + c * NBANDS] = tmp;
      prev[c] = prev[c] + q - beta * q;
    } while (++c < C);
  }
}
static void decode_fine_energy(int start, int end, float *oldE, const int *fine_quant, RangeDec *dec, int C) {
  int i, c;
  for (i = start; i < end; i++) {
    int extra = fine_quant[i];
    if (fine_quant[i] <= 0) continue;
    if (rd_tell(dec) + C * fine_quant[i] > (int32_t)dec->storage * 8) continue;
    c = 0;
    do {
      int q2 = rd_bits(dec, extra);
      float offset = (q2 + .5f) * (1 << (14 - extra)) * (1.f / 16384) - .5f;
      offset *= (1 << (14 - 0)) * (1.f / 16384);
      oldE[i + c * NBANDS] += offset;
    } while (++c < C);
  }
}
static void decode_energy_finalise(int start, int end, float *oldE, const int *fine_quant, const int *fine_priority,
                                   int bits_left, RangeDec *dec, int C) {
  int i, prio, c;
  for (prio = 0; prio < 2; prio++) {
    for (i = start; i < end && bits_left >= C; i++) {
      if (fine_quant[i] >= MAX_FINE_BITS || fine_priority[i] != prio) continue;
      c = 0;
      do {
        int q2 = rd_bits(dec, 1);
        float offset = (q2 - .5f) * (1 << (14 - fine_quant[i] - 1)) * (1.f / 16384);
        oldE[i + c * NBANDS] += offset;
        bits_left--;
      } while (++c < C);
    }
  }
}

/* celt/celt_decoder.c:513-547 */
static void decode_tf(int start, int end, int isTransient, int *tf_res, int LM, RangeDec *dec) {
  int i, curr = 0, tf_select = 0, tf_select_rsv, tf_changed = 0, logp;
  uint32_t budget = dec->storage * 8, tell = rd_tell(dec);
  logp = isTransient ? 2 : 4;
  tf_select_rsv = LM > 0 && tell + logp + 1 <= budget;
  budget -= tf_select_rsv;
  for (i = start; i < end; i++) {
    if (tell + logp <= budget) {
      curr ^= rd_bit_logp(dec, logp);
      tell = rd_tell(dec);
      tf_changed |= curr;
    }
    tf_res[i] = curr;
    logp = isTransient ? 4 : 5;
  }
  if (tf_select_rsv && ob_tf_select[LM * 8 + 4 * isTransient + 0 + tf_changed] !=
                           ob_tf_select[LM * 8 + 4 * isTransient + 2 + tf_changed])
    tf_select = rd_bit_logp(dec, 1);
  for (i = start; i < end; i++) tf_res[i] = ob_tf_select[LM * 8 + 4 * isTransient + 2 * tf_select + tf_res[i]];
}

/* ------------------------------------------------------------------------------------------
 * Anti-collapse, denormalisation (celt/bands.c:188-353)
 * ---------------------------------------------------------------------------------------- */
static void anti_collapse(float *X_, const unsigned char *collapse_masks, int LM, int C, int size, int start, int end,
                          const float *logE, const float *prev1logE, const float *prev2logE, const int *pulses,
                          uint32_t seed) {
  int c, i, j, k;
  for (i = start; i < end; i++) {
    int N0 = ob_ebands[i + 1] - ob_ebands[i];
    int depth = ((uint32_t)(1 + pulses[i]) / (uint32_t)(ob_ebands[i + 1] - ob_ebands[i])) >> LM;
    float thresh = .5f * f_exp2(-.125f * depth);
    float sqrt_1 = f_rsqrt((float)(N0 << LM));
    c = 0;
    do {
      float *X, prev1, prev2, Ediff, r;
      int renorm = 0;
      prev1 = prev1logE[c * NBANDS + i];
      prev2 = prev2logE[c * NBANDS + i];
      if (C == 1) {
        prev1 = prev1 > prev1logE[NBANDS + i] ? prev1 : prev1logE[NBANDS + i];
        prev2 = prev2 > prev2logE[NBANDS + i] ? prev2 : prev2logE[NBANDS + i];
      }
      Ediff = logE[c * NBANDS + i] - (prev1 < prev2 ? prev1 : prev2);
      Ediff = Ediff > 0 ? Ediff : 0;
      r = 2.f * f_exp2(-Ediff);
      if (LM == 3) r *= 1.41421356f;
      r = thresh < r ? thresh : r;
      r = r * sqrt_1;
      X = X_ + c * size + (ob_ebands[i] << LM);
      for (k = 0; k < 1 << LM; k++) {
        if (!(collapse_masks[i * C + c] & 1 << k)) {
          for (j = 0; j < N0; j++) {
            seed = 1664525u * seed + 1013904223u;
            X[(j << LM) + k] = (seed & 0x8000 ? r : -r);
          }
          renorm = 1;
        }
      }
      if (renorm) renormalise(X, N0 << LM, 1.0f);
    } while (++c < C);
  }
}

static int g_denorm_downsample = 1; /* set per frame by the decoder (celt/bands.c:199-200) */
static void denormalise(const float *X, float *freq, const float *bandLogE, int start, int end, int M, int silence) {
  int i, N = M * SHORT_N, bound = IMIN(M * ob_ebands[end], N / g_denorm_downsample);
  float *f = freq;
  const float *x;
  if (silence) { bound = 0; start = end = 0; }
  x = X + M * ob_ebands[start];
  if (start != 0) { for (i = 0; i < M * ob_ebands[start]; i++) *f++ = 0; }
  else f += M * ob_ebands[start];
  for (i = start; i < end; i++) {
    int j = M * ob_ebands[i], band_end = M * ob_ebands[i + 1];
    float lg = bandLogE[i] + ob_emeans[i];
    float g = f_exp2(32.f < lg ? 32.f : lg);
    do { *f++ = (*x) * g; x++; } while (++j < band_end);
  }
  memset(&freq[bound], 0, (N - bound) * sizeof(float));
}

/* ------------------------------------------------------------------------------------------
 * Inverse MDCT with the reference's mixed-radix FFT (celt/kiss_fft.c:52-313,562-613;
 * celt/mdct.c:268-389)
 * ---------------------------------------------------------------------------------------- */
typedef struct { float r, i; } cpx;
#define TW(k) (*(const cpx *)&ob_fft_tw[2 * (k)])

static void bfly2(cpx *F, int m, int N) { /* m == 4 */
  const float tw = 0.7071067812f;
  int i;
  (void)m;
  for (i = 0; i < N; i++) {
    cpx *F2 = F + 4, t;
    t = F2[0];
    F2[0].r = F[0].r - t.r; F2[0].i = F[0].i - t.i; F[0].r += t.r; F[0].i += t.i;
    t.r = (F2[1].r + F2[1].i) * tw;
    t.i = (F2[1].i - F2[1].r) * tw;
    F2[1].r = F[1].r - t.r; F2[1].i = F[1].i - t.i; F[1].r += t.r; F[1].i += t.i;
    t.r = F2[2].i;
    t.i = -F2[2].r;
    F2[2].r = F[2].r - t.r; F2[2].i = F[2].i - t.i; F[2].r += t.r; F[2].i += t.i;
    t.r = (F2[3].i - F2[3].r) * tw;
    t.i = (-(F2[3].i + F2[3].r)) * tw;
    F2[3].r = F[3].r - t.r; F2[3].i = F[3].i - t.i; F[3].r += t.r; F[3].i += t.i;
    F += 8;
  }
}
#define CMUL(m, a, b) do { (m).r = (a).r * (b).r - (a).i * (b).i; (m).i = (a).r * (b).i + (a).i * (b).r; } while (0)
static void bfly4(cpx *F, int fstride, int m, int N, int mm) {
  int i, j;
  if (m == 1) {
    for (i = 0; i < N; i++) {
      cpx s0, s1;
      s0.r = F[0].r - F[2].r; s0.i = F[0].i - F[2].i;
      F[0].r += F[2].r; F[0].i += F[2].i;
      s1.r = F[1].r + F[3].r; s1.i = F[1].i + F[3].i;
      F[2].r = F[0].r - s1.r; F[2].i = F[0].i - s1.i;
      F[0].r += s1.r; F[0].i += s1.i;
      s1.r = F[1].r - F[3].r; s1.i = F[1].i - F[3].i;
      F[1].r = s0.r + s1.i; F[1].i = s0.i - s1.r;
      F[3].r = s0.r - s1.i; F[3].i = s0.i + s1.r;
      F += 4;
    }
  } else {
    cpx *Fb = F;
    const int m2 = 2 * m, m3 = 3 * m;
    for (i = 0; i < N; i++) {
      int t1 = 0, t2 = 0, t3 = 0;
      F = Fb + i * mm;
      for (j = 0; j < m; j++) {
        cpx s0, s1, s2, s3, s4, s5;
        CMUL(s0, F[m], TW(t1));
        CMUL(s1, F[m2], TW(t2));
        CMUL(s2, F[m3], TW(t3));
        s5.r = F->r - s1.r; s5.i = F->i - s1.i;
        F->r += s1.r; F->i += s1.i;
        s3.r = s0.r + s2.r; s3.i = s0.i + s2.i;
        s4.r = s0.r - s2.r; s4.i = s0.i - s2.i;
        F[m2].r = F->r - s3.r; F[m2].i = F->i - s3.i;
        t1 += fstride; t2 += fstride * 2; t3 += fstride * 3;
        F->r += s3.r; F->i += s3.i;
        F[m].r = s5.r + s4.i; F[m].i = s5.i - s4.r;
        F[m3].r = s5.r - s4.i; F[m3].i = s5.i + s4.r;
        ++F;
      }
    }
  }
}
static void bfly3(cpx *F, int fstride, int m, int N, int mm) {
  int i, k;
  const int m2 = 2 * m;
  cpx *Fb = F;
  cpx epi3 = TW(fstride * m);
  for (i = 0; i < N; i++) {
    int t1 = 0, t2 = 0;
    F = Fb + i * mm;
    k = m;
    do {
      cpx s0, s1, s2, s3;
      CMUL(s1, F[m], TW(t1));
      CMUL(s2, F[m2], TW(t2));
      s3.r = s1.r + s2.r; s3.i = s1.i + s2.i;
      s0.r = s1.r - s2.r; s0.i = s1.i - s2.i;
      t1 += fstride; t2 += fstride * 2;
      F[m].r = F->r - s3.r * .5f;
      F[m].i = F->i - s3.i * .5f;
      s0.r *= epi3.i; s0.i *= epi3.i;
      F->r += s3.r; F->i += s3.i;
      F[m2].r = F[m].r + s0.i;
      F[m2].i = F[m].i - s0.r;
      F[m].r = F[m].r - s0.i;
      F[m].i = F[m].i + s0.r;
      ++F;
    } while (--k);
  }
}
static void bfly5(cpx *F, int fstride, int m, int N, int mm) {
  int i, u;
  cpx *Fb = F;
  cpx ya = TW(fstride * m), yb = TW(fstride * 2 * m);
  for (i = 0; i < N; i++) {
    cpx *F0, *F1, *F2, *F3, *F4;
    F = Fb + i * mm;
    F0 = F; F1 = F0 + m; F2 = F0 + 2 * m; F3 = F0 + 3 * m; F4 = F0 + 4 * m;
    for (u = 0; u < m; ++u) {
      cpx s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12;
      s0 = *F0;
      CMUL(s1, *F1, TW(u * fstride));
      CMUL(s2, *F2, TW(2 * u * fstride));
      CMUL(s3, *F3, TW(3 * u * fstride));
      CMUL(s4, *F4, TW(4 * u * fstride));
      s7.r = s1.r + s4.r; s7.i = s1.i + s4.i;
      s10.r = s1.r - s4.r; s10.i = s1.i - s4.i;
      s8.r = s2.r + s3.r; s8.i = s2.i + s3.i;
      s9.r = s2.r - s3.r; s9.i = s2.i - s3.i;
      F0->r = F0->r + (s7.r + s8.r);
      F0->i = F0->i + (s7.i + s8.i);
      s5.r = s0.r + (s7.r * ya.r + s8.r * yb.r);
      s5.i = s0.i + (s7.i * ya.r + s8.i * yb.r);
      s6.r = s10.i * ya.i + s9.i * yb.i;
      s6.i = -(s10.r * ya.i + s9.r * yb.i);
      F1->r = s5.r - s6.r; F1->i = s5.i - s6.i;
      F4->r = s5.r + s6.r; F4->i = s5.i + s6.i;
      s11.r = s0.r + (s7.r * yb.r + s8.r * ya.r);
      s11.i = s0.i + (s7.i * yb.r + s8.i * ya.r);
      s12.r = s9.i * ya.i - s10.i * yb.i;
      s12.i = s10.r * yb.i - s9.r * ya.i;
      F2->r = s11.r + s12.r; F2->i = s11.i + s12.i;
      F3->r = s11.r - s12.r; F3->i = s11.i - s12.i;
      ++F0; ++F1; ++F2; ++F3; ++F4;
    }
  }
}
/* in-place FFT of length 480>>shift on digit-reversed data (celt/kiss_fft.c:562-613) */
void orc_fft_inplace(cpx *fout, int shift) {
  const short *fac = ob_fft_factors + 10 * shift;
  int fstride[6], L = 0, m, m2, i, p;
  fstride[0] = 1;
  do {
    p = fac[2 * L];
    m = fac[2 * L + 1];
    fstride[L + 1] = fstride[L] * p;
    L++;
  } while (m != 1);
  m = fac[2 * L - 1];
  for (i = L - 1; i >= 0; i--) {
    m2 = i != 0 ? fac[2 * i - 1] : 1;
    switch (fac[2 * i]) {
      case 2: bfly2(fout, m, fstride[i]); break;
      case 4: bfly4(fout, fstride[i] << shift, m, fstride[i], m2); break;
      case 3: bfly3(fout, fstride[i] << shift, m, fstride[i], m2); break;
      case 5: bfly5(fout, fstride[i] << shift, m, fstride[i], m2); break;
    }
    m = m2;
  }
}
static const short *bitrev_table(int shift) {
  return shift == 0 ? ob_fft_bitrev480 : shift == 1 ? ob_fft_bitrev240 : shift == 2 ? ob_fft_bitrev120 : ob_fft_bitrev60;
}
/* celt/mdct.c:268-389 ; `in` has N/2 coefficients with the given stride; out receives N/2+overlap/2.. */
void orc_imdct(const float *in, float *out, int shift, int stride) {
  int i, N = 1920 >> shift, N2 = N >> 1, N4 = N >> 2, overlap = OVERLAP;
  const float *trig = ob_mdct_trig;
  for (i = 0; i < shift; i++) trig += (1920 >> i) >> 1;
  {
    const float *xp1 = in, *xp2 = in + stride * (N2 - 1);
    float *yp = out + (overlap >> 1);
    const short *bitrev = bitrev_table(shift);
    for (i = 0; i < N4; i++) {
      int rev = *bitrev++;
      float x1 = *xp1, x2 = *xp2;
      float yr = x2 * trig[i] + x1 * trig[N4 + i];
      float yi = x1 * trig[i] - x2 * trig[N4 + i];
      yp[2 * rev + 1] = yr;
      yp[2 * rev] = yi;
      xp1 += 2 * stride;
      xp2 -= 2 * stride;
    }
  }
  orc_fft_inplace((cpx *)(out + (overlap >> 1)), shift);
  {
    float *yp0 = out + (overlap >> 1), *yp1 = out + (overlap >> 1) + N2 - 2;
    for (i = 0; i < (N4 + 1) >> 1; i++) {
      float re, im, yr, yi, t0, t1;
      re = yp0[1]; im = yp0[0];
      t0 = trig[i]; t1 = trig[N4 + i];
      yr = re * t0 + im * t1;
      yi = re * t1 - im * t0;
      re = yp1[1]; im = yp1[0];
      yp0[0] = yr;
      yp1[1] = yi;
      t0 = trig[N4 - i - 1]; t1 = trig[N2 - i - 1];
      yr = re * t0 + im * t1;
      yi = re * t1 - im * t0;
      yp1[0] = yr;
      yp0[1] = yi;
      yp0 += 2;
      yp1 -= 2;
    }
  }
  {
    float *xp1 = out + overlap - 1, *yp1 = out;
    const float *wp1 = ob_window, *wp2 = ob_window + overlap - 1;
    for (i = 0; i < overlap / 2; i++) {
      float x1 = *xp1, x2 = *yp1;
      *yp1++ = x2 * (*wp2) - x1 * (*wp1);
      *xp1-- = x2 * (*wp1) + x1 * (*wp2);
      wp1++;
      wp2--;
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Comb post-filter (celt/celt.c:166-313)
 * ---------------------------------------------------------------------------------------- */
static void comb_filter_inplace(float *y, int T0, int T1, int N, float g0, float g1, int tapset0, int tapset1) {
  static const float gains[3][3] = {{0.3066406250f, 0.2170410156f, 0.1296386719f},
                                    {0.4638671875f, 0.2680664062f, 0.f},
                                    {0.7998046875f, 0.1000976562f, 0.f}};
  float *x = y;
  int i, overlap = OVERLAP;
  float g00, g01, g02, g10, g11, g12, x0, x1, x2, x3, x4;
  if (g0 == 0 && g1 == 0) return;
  T0 = IMAX(T0, COMB_MINPERIOD);
  T1 = IMAX(T1, COMB_MINPERIOD);
  g00 = g0 * gains[tapset0][0]; g01 = g0 * gains[tapset0][1]; g02 = g0 * gains[tapset0][2];
  g10 = g1 * gains[tapset1][0]; g11 = g1 * gains[tapset1][1]; g12 = g1 * gains[tapset1][2];
  x1 = x[-T1 + 1]; x2 = x[-T1]; x3 = x[-T1 - 1]; x4 = x[-T1 - 2];
  if (g0 == g1 && T0 == T1 && tapset0 == tapset1) overlap = 0;
  for (i = 0; i < overlap; i++) {
    float f;
    x0 = x[i - T1 + 2];
    f = ob_window[i] * ob_window[i];
    y[i] = x[i] + ((1.0f - f) * g00) * x[i - T0] + ((1.0f - f) * g01) * (x[i - T0 + 1] + x[i - T0 - 1]) +
           ((1.0f - f) * g02) * (x[i - T0 + 2] + x[i - T0 - 2]) + (f * g10) * x2 + (f * g11) * (x1 + x3) +
           (f * g12) * (x0 + x4);
    x4 = x3; x3 = x2; x2 = x1; x1 = x0;
  }
  if (g1 == 0) return;
  /* constant part, celt.c:166-193 */
  y += i; x += i; N -= i;
  x4 = x[-T1 - 2]; x3 = x[-T1 - 1]; x2 = x[-T1]; x1 = x[-T1 + 1];
  for (i = 0; i < N; i++) {
    x0 = x[i - T1 + 2];
    y[i] = x[i] + g10 * x2 + g11 * (x1 + x3) + g12 * (x0 + x4);
    x4 = x3; x3 = x2; x2 = x1; x1 = x0;
  }
}

/* ------------------------------------------------------------------------------------------
 * Decoder state + frame decode (celt/celt_decoder.c:87-139, 1104-1617)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int channels;        /* output channels (CC) */
  int stream_channels; /* coded channels of the current packet (C) */
  int start, end;
  uint32_t rng;
  int error;
  int postfilter_period, postfilter_period_old, postfilter_tapset, postfilter_tapset_old;
  float postfilter_gain, postfilter_gain_old;
  float preemph_mem[2];
  float decode_mem[2][DEC_BUF + OVERLAP];
  float oldEBands[2 * NBANDS], oldLogE[2 * NBANDS], oldLogE2[2 * NBANDS], backgroundLogE[2 * NBANDS];
  /* Opus layer (src/opus_decoder.c) */
  int prev_mode_celt;
  uint32_t final_range;
  int downsample; /* 48000 / Fs of the output (celt/celt_decoder.c:222-240 resampling_factor), 1 after init */
  /* packet-loss concealment (celt/celt_decoder.c:110-124) */
  int loss_duration, plc_duration, skip_plc, last_frame_type, prefilter_and_fold, last_pitch_index;
  float lpc[2][24];
  int last_frame_size48; /* Opus layer st->frame_size (48 kHz samples) of the last packet, for len <= 1 frames */
  int disable_inv;       /* OPUS_SET_PHASE_INVERSION_DISABLED (celt_decoder.c:1763-1771) */
  int decode_gain;       /* OPUS_SET_GAIN, Q8 dB (src/opus_decoder.c:686-700) */
  int last_packet_duration;
} OrcDecoder;
static void decode_lost(OrcDecoder *st, int N, int LM);        /* celt_plc.inc */
static void prefilter_and_fold(OrcDecoder *st, int N);
static void deemphasis_out(OrcDecoder *st, float *out_syn[], float *pcm, int N);

int orc_decoder_size(void) { return (int)sizeof(OrcDecoder); }

/* celt_decoder.c:240-268 + OPUS_RESET_STATE 1794-1813 */
void orc_decoder_init(OrcDecoder *st, int channels) {
  int i;
  memset(st, 0, sizeof(*st));
  st->channels = st->stream_channels = channels;
  st->start = 0;
  st->end = NBANDS;
  for (i = 0; i < 2 * NBANDS; i++) st->oldLogE[i] = st->oldLogE2[i] = -28.f;
  st->downsample = 1;
  st->last_frame_size48 = 120; /* src/opus_decoder.c:178 */
}
/* opus_decoder_create(Fs): Fs in {48000, 24000, 16000, 12000, 8000} (celt/celt_decoder.c:222-240, resampling_factor celt/celt.c) */
int orc_decoder_set_fs(OrcDecoder *st, int Fs) {
  if (Fs != 48000 && Fs != 24000 && Fs != 16000 && Fs != 12000 && Fs != 8000) return -1;
  st->downsample = 48000 / Fs;
  return 0;
}

/* celt_decoder.c:413-511 */
static void synthesis(float *X, float *out_syn[], const float *oldBandE, int start, int effEnd, int C, int CC,
                      int isTransient, int LM, int silence) {
  int c, i, b, M = 1 << LM, B, N = SHORT_N << LM, NB, shift;
  float freq[MAXN];
  if (isTransient) { B = M; NB = SHORT_N; shift = 3; } else { B = 1; NB = SHORT_N << LM; shift = 3 - LM; }
  if (CC == 2 && C == 1) {
    float *freq2;
    denormalise(X, freq, oldBandE, start, effEnd, M, silence);
    freq2 = out_syn[1] + OVERLAP / 2;
    memcpy(freq2, freq, N * sizeof(float));
    for (b = 0; b < B; b++) orc_imdct(&freq2[b], out_syn[0] + NB * b, shift, B);
    for (b = 0; b < B; b++) orc_imdct(&freq[b], out_syn[1] + NB * b, shift, B);
  } else if (CC == 1 && C == 2) {
    float *freq2 = out_syn[0] + OVERLAP / 2;
    denormalise(X, freq, oldBandE, start, effEnd, M, silence);
    denormalise(X + N, freq2, oldBandE + NBANDS, start, effEnd, M, silence);
    for (i = 0; i < N; i++) freq[i] = (.5f * freq[i]) + (.5f * freq2[i]);
    for (b = 0; b < B; b++) orc_imdct(&freq[b], out_syn[0] + NB * b, shift, B);
  } else {
    c = 0;
    do {
      denormalise(X + c * N, freq, oldBandE + c * NBANDS, start, effEnd, M, silence);
      for (b = 0; b < B; b++) orc_imdct(&freq[b], out_syn[c] + NB * b, shift, B);
    } while (++c < CC);
  }
}

/* Debug tap for the tests: the last frame's normalised spectrum X[C][N] (after anti-collapse). */
float orc_dbg_X[2 * MAXN];
/* [0..15] scalars, [16] tf_res, [40] dynalloc offsets, [64] pulses, [88] fine_quant, [112] coarse energies*1024 */
int orc_dbg_side[160];

/* One CELT frame: `data` is the CELT payload (no TOC byte). Returns samples/channel or <0.
 * pcm: interleaved float, CC channels, +-1.0 full scale. */
int orc_celt_decode_frame(OrcDecoder *st, const uint8_t *data, int len, float *pcm, int frame_size) {
  int c, i, N, LM, M, C = st->stream_channels, CC = st->channels, start = st->start, end = st->end, effEnd;
  int spread, shortBlocks, isTransient, intra_ener, codedBands, alloc_trim, postfilter_pitch = 0, postfilter_tapset = 0;
  int intensity = 0, dual_stereo = 0, dynalloc_logp, anti_collapse_rsv, anti_collapse_on = 0, silence;
  float postfilter_gain = 0;
  int32_t bits, total_bits, balance, tell;
  RangeDec dec;
  float X[2 * MAXN];
  int fine_quant[NBANDS], pulses[NBANDS], cap[NBANDS], offsets[NBANDS], fine_priority[NBANDS], tf_res[NBANDS];
  unsigned char collapse_masks[2 * NBANDS];
  float *decode_mem[2], *out_syn[2];
  float *oldBandE = st->oldEBands, *oldLogE = st->oldLogE, *oldLogE2 = st->oldLogE2, *backgroundLogE = st->backgroundLogE;

  for (LM = 0; LM <= 3; LM++) if (SHORT_N << LM == frame_size) break;
  if (LM > 3) return -1;
  M = 1 << LM;
  if (len < 0 || len > 1275 || pcm == NULL) return -1;
  N = M * SHORT_N;
  g_denorm_downsample = st->downsample > 1 ? st->downsample : 1;
  c = 0;
  do { decode_mem[c] = st->decode_mem[c]; out_syn[c] = decode_mem[c] + DEC_BUF - N; } while (++c < CC);
  if (data == NULL || len <= 1) { /* celt_decoder.c:1281-1291 */
    decode_lost(st, N, LM);
    deemphasis_out(st, out_syn, pcm, N);
    return frame_size / (st->downsample > 1 ? st->downsample : 1);
  }
  if (st->loss_duration == 0) st->skip_plc = 0; /* :1299-1301 */
  effEnd = IMIN(end, NBANDS);
  rd_init(&dec, data, len);
  memset(X, 0, sizeof(X));
  memset(collapse_masks, 0, sizeof(collapse_masks));
  if (C == 1)
    for (i = 0; i < NBANDS; i++) oldBandE[i] = oldBandE[i] > oldBandE[NBANDS + i] ? oldBandE[i] : oldBandE[NBANDS + i];
  total_bits = len * 8;
  tell = rd_tell(&dec);
  if (tell >= total_bits) silence = 1;
  else if (tell == 1) silence = rd_bit_logp(&dec, 15);
  else silence = 0;
  if (silence) {
    tell = len * 8;
    dec.nbits_total += tell - rd_tell(&dec);
  }
  if (start == 0 && tell + 16 <= total_bits) {
    if (rd_bit_logp(&dec, 1)) {
      int qg, octave = rd_uint(&dec, 6);
      postfilter_pitch = (16 << octave) + rd_bits(&dec, 4 + octave) - 1;
      qg = rd_bits(&dec, 3);
      if (rd_tell(&dec) + 2 <= total_bits) postfilter_tapset = rd_icdf(&dec, ob_tapset_icdf, 2);
      postfilter_gain = .09375f * (qg + 1);
    }
    tell = rd_tell(&dec);
  }
  if (LM > 0 && tell + 3 <= total_bits) {
    isTransient = rd_bit_logp(&dec, 3);
    tell = rd_tell(&dec);
  } else isTransient = 0;
  shortBlocks = isTransient ? M : 0;
  intra_ener = tell + 3 <= total_bits ? rd_bit_logp(&dec, 3) : 0;
  if (!intra_ener && st->loss_duration != 0) { /* safe energy prediction after a loss, celt_decoder.c:1362-1393 */
    c = 0;
    do {
      float safety = 0;
      const int missing = IMIN(10, st->loss_duration >> LM);
      if (LM == 0) safety = 1.5f;
      else if (LM == 1) safety = .5f;
      for (i = start; i < end; i++) {
        float *E = &oldBandE[c * NBANDS + i];
        const float l1 = oldLogE[c * NBANDS + i], l2 = oldLogE2[c * NBANDS + i];
        if (*E < (l1 > l2 ? l1 : l2)) {
          float E0 = *E, slope, a = l1 - E0, b = .5f * (l2 - E0);
          slope = a > b ? a : b;
          slope = slope < 2.f ? slope : 2.f;
          { float d = (1 + missing) * slope; E0 -= 0 > d ? 0 : d; }
          *E = -20.f > E0 ? -20.f : E0;
        } else {
          float m = *E < l1 ? *E : l1;
          *E = m < l2 ? m : l2;
        }
        *E -= safety;
      }
    } while (++c < 2);
  }
  decode_coarse_energy(start, end, oldBandE, intra_ener, &dec, C, LM);
  decode_tf(start, end, isTransient, tf_res, LM, &dec);
  tell = rd_tell(&dec);
  spread = 2; /* SPREAD_NORMAL */
  if (tell + 4 <= total_bits) spread = rd_icdf(&dec, ob_spread_icdf, 5);
  for (i = 0; i < NBANDS; i++) { /* init_caps, celt.c:329-338 */
    int n = (ob_ebands[i + 1] - ob_ebands[i]) << LM;
    cap[i] = (ob_cache_caps[NBANDS * (2 * LM + C - 1) + i] + 64) * C * n >> 2;
  }
  dynalloc_logp = 6;
  total_bits <<= BITRES;
  tell = rd_tell_frac(&dec);
  for (i = start; i < end; i++) {
    int width = C * (ob_ebands[i + 1] - ob_ebands[i]) << LM;
    int quanta = IMIN(width << BITRES, IMAX(6 << BITRES, width));
    int loop_logp = dynalloc_logp, boost = 0;
    while (tell + (loop_logp << BITRES) < total_bits && boost < cap[i]) {
      int flag = rd_bit_logp(&dec, loop_logp);
      tell = rd_tell_frac(&dec);
      if (!flag) break;
      boost += quanta;
      total_bits -= quanta;
      loop_logp = 1;
    }
    offsets[i] = boost;
    if (boost > 0) dynalloc_logp = IMAX(2, dynalloc_logp - 1);
  }
  for (i = 0; i < start; i++) offsets[i] = 0;
  alloc_trim = tell + (6 << BITRES) <= total_bits ? rd_icdf(&dec, ob_trim_icdf, 7) : 5;
  bits = (((int32_t)len * 8) << BITRES) - (int32_t)rd_tell_frac(&dec) - 1;
  anti_collapse_rsv = isTransient && LM >= 2 && bits >= ((LM + 2) << BITRES) ? (1 << BITRES) : 0;
  bits -= anti_collapse_rsv;
  codedBands = compute_allocation(start, end, offsets, cap, alloc_trim, &intensity, &dual_stereo, bits, &balance,
                                  pulses, fine_quant, fine_priority, C, LM, &dec, NULL);
  {  /* debug tap: the side information of this frame (tests diff it between two packets) */
    int *t = orc_dbg_side;
    t[0] = silence; t[1] = isTransient; t[2] = intra_ener; t[3] = spread; t[4] = alloc_trim; t[5] = intensity; t[6] = dual_stereo;
    t[7] = codedBands; t[8] = postfilter_pitch; t[9] = (int)(postfilter_gain * 1000); t[10] = postfilter_tapset; t[11] = bits;
    for (i = 0; i < NBANDS; i++) { t[16 + i] = tf_res[i]; t[40 + i] = offsets[i]; t[64 + i] = pulses[i]; t[88 + i] = fine_quant[i]; }
    for (i = 0; i < 2 * NBANDS; i++) t[112 + i] = (int)floor(oldBandE[i] * 1024 + .5);
  }
  decode_fine_energy(start, end, oldBandE, fine_quant, &dec, C);
  c = 0;
  do memmove(decode_mem[c], decode_mem[c] + N, (DEC_BUF - N + OVERLAP) * sizeof(float)); while (++c < CC);
  decode_all_bands(start, end, X, C == 2 ? X + N : NULL, collapse_masks, pulses, shortBlocks, spread, dual_stereo,
                   intensity, tf_res, len * (8 << BITRES) - anti_collapse_rsv, balance, &dec, LM, codedBands, &st->rng,
                   CC == 1 || st->disable_inv /* disable_inv = (channels==1) at init, celt_decoder.c:253; OPUS_SET_PHASE_INVERSION_DISABLED */);
  if (anti_collapse_rsv > 0) anti_collapse_on = rd_bits(&dec, 1);
  decode_energy_finalise(start, end, oldBandE, fine_quant, fine_priority, len * 8 - rd_tell(&dec), &dec, C);
  if (anti_collapse_on)
    anti_collapse(X, collapse_masks, LM, C, N, start, end, oldBandE, oldLogE, oldLogE2, pulses, st->rng);
  memcpy(orc_dbg_X, X, sizeof(float) * C * N);
  if (silence) for (i = 0; i < C * NBANDS; i++) oldBandE[i] = -28.f;
  if (st->prefilter_and_fold) prefilter_and_fold(st, N);
  synthesis(X, out_syn, oldBandE, start, effEnd, C, CC, isTransient, LM, silence);
  c = 0;
  do {
    st->postfilter_period = IMAX(st->postfilter_period, COMB_MINPERIOD);
    st->postfilter_period_old = IMAX(st->postfilter_period_old, COMB_MINPERIOD);
    comb_filter_inplace(out_syn[c], st->postfilter_period_old, st->postfilter_period, SHORT_N, st->postfilter_gain_old,
                        st->postfilter_gain, st->postfilter_tapset_old, st->postfilter_tapset);
    if (LM != 0)
      comb_filter_inplace(out_syn[c] + SHORT_N, st->postfilter_period, postfilter_pitch, N - SHORT_N, st->postfilter_gain,
                          postfilter_gain, st->postfilter_tapset, postfilter_tapset);
  } while (++c < CC);
  st->postfilter_period_old = st->postfilter_period;
  st->postfilter_gain_old = st->postfilter_gain;
  st->postfilter_tapset_old = st->postfilter_tapset;
  st->postfilter_period = postfilter_pitch;
  st->postfilter_gain = postfilter_gain;
  st->postfilter_tapset = postfilter_tapset;
  if (LM != 0) {
    st->postfilter_period_old = st->postfilter_period;
    st->postfilter_gain_old = st->postfilter_gain;
    st->postfilter_tapset_old = st->postfilter_tapset;
  }
  if (C == 1) memcpy(&oldBandE[NBANDS], oldBandE, NBANDS * sizeof(float));
  if (!isTransient) {
    memcpy(oldLogE2, oldLogE, 2 * NBANDS * sizeof(float));
    memcpy(oldLogE, oldBandE, 2 * NBANDS * sizeof(float));
  } else {
    for (i = 0; i < 2 * NBANDS; i++) oldLogE[i] = oldLogE[i] < oldBandE[i] ? oldLogE[i] : oldBandE[i];
  }
  {
    float inc = IMIN(160, st->loss_duration + M) * 0.001f;
    for (i = 0; i < 2 * NBANDS; i++) {
      float a = backgroundLogE[i] + inc;
      backgroundLogE[i] = a < oldBandE[i] ? a : oldBandE[i];
    }
  }
  c = 0;
  do {
    for (i = 0; i < start; i++) { oldBandE[c * NBANDS + i] = 0; oldLogE[c * NBANDS + i] = oldLogE2[c * NBANDS + i] = -28.f; }
    for (i = end; i < NBANDS; i++) { oldBandE[c * NBANDS + i] = 0; oldLogE[c * NBANDS + i] = oldLogE2[c * NBANDS + i] = -28.f; }
  } while (++c < 2);
  st->rng = dec.rng;
  deemphasis_out(st, out_syn, pcm, N);
  st->loss_duration = 0;
  st->plc_duration = 0;
  st->last_frame_type = 1; /* FRAME_NORMAL */
  st->prefilter_and_fold = 0;
  if (rd_tell(&dec) > 8 * len) return -3;
  if (dec.error) st->error = 1;
  return frame_size / (st->downsample > 1 ? st->downsample : 1);
}

/* de-emphasis, celt_decoder.c:288-408 (coef0 = 0.85000610f at 48 kHz; the output rate is 48000/downsample) */
static void deemphasis_out(OrcDecoder *st, float *out_syn[], float *pcm, int N) {
  const float coef0 = 0.85000610f;
  const int ds = st->downsample > 1 ? st->downsample : 1, CC = st->channels;
  int c = 0;
  do {
    float m = st->preemph_mem[c];
    const float *x = out_syn[c];
    int j;
    for (j = 0; j < N; j++) {
      float tmp = x[j] + 1e-30f + m;
      m = coef0 * tmp;
      if (ds == 1) pcm[j * CC + c] = (1 / 32768.f) * tmp;
      else if (j % ds == 0) pcm[(j / ds) * CC + c] = (1 / 32768.f) * tmp;   /* :387-401: keep every downsample-th sample */
    }
    st->preemph_mem[c] = m;
  } while (++c < CC);
}

/* ------------------------------------------------------------------------------------------
 * Opus packet layer, CELT-only subset (src/opus.c:203-420 packet parse; src/opus_decoder.c:271-714)
 * ---------------------------------------------------------------------------------------- */
static int parse_size(const uint8_t *data, int len, int *size) {
  if (len < 1) { *size = -1; return -1; }
  if (data[0] < 252) { *size = data[0]; return 1; }
  if (len < 2) { *size = -1; return -1; }
  *size = 4 * data[1] + data[0];
  return 2;
}

/* Parses an Opus packet (RFC 6716 section 3) into frame offsets/sizes. Returns count or <0. */
int orc_packet_parse(const uint8_t *data, int len, int *toc_out, int offsets[48], int sizes[48]) {
  const uint8_t *data0 = data;
  int i, bytes, count, cbr, framesize, last_size, pad = 0;
  uint8_t ch, toc;
  if (len <= 0) return -4;
  {
    int t = data[0];
    if (t & 0x80) framesize = (48000 << ((t >> 3) & 3)) / 400;
    else if ((t & 0x60) == 0x60) framesize = (t & 0x08) ? 960 : 480;
    else { int a = (t >> 3) & 3; framesize = a == 3 ? 2880 : (48000 << a) / 100; }
  }
  cbr = 0;
  toc = *data++;
  len--;
  last_size = len;
  switch (toc & 0x3) {
    case 0: count = 1; break;
    case 1:
      count = 2; cbr = 1;
      if (len & 0x1) return -4;
      last_size = len / 2;
      sizes[0] = last_size;
      break;
    case 2:
      count = 2;
      bytes = parse_size(data, len, sizes);
      len -= bytes;
      if (sizes[0] < 0 || sizes[0] > len) return -4;
      data += bytes;
      last_size = len - sizes[0];
      break;
    default:
      if (len < 1) return -4;
      ch = *data++;
      count = ch & 0x3F;
      if (count <= 0 || framesize * count > 5760) return -4;
      len--;
      if (ch & 0x40) {
        int p;
        do {
          int tmp;
          if (len <= 0) return -4;
          p = *data++;
          len--;
          tmp = p == 255 ? 254 : p;
          len -= tmp;
          pad += tmp;
        } while (p == 255);
      }
      if (len < 0) return -4;
      cbr = !(ch & 0x80);
      if (!cbr) {
        last_size = len;
        for (i = 0; i < count - 1; i++) {
          bytes = parse_size(data, len, sizes + i);
          len -= bytes;
          if (sizes[i] < 0 || sizes[i] > len) return -4;
          data += bytes;
          last_size -= bytes + sizes[i];
        }
        if (last_size < 0) return -4;
      } else {
        last_size = len / count;
        if (last_size * count != len) return -4;
        for (i = 0; i < count - 1; i++) sizes[i] = last_size;
      }
      break;
  }
  if (last_size > 1275) return -4;
  sizes[count - 1] = last_size;
  for (i = 0; i < count; i++) { offsets[i] = (int)(data - data0); data += sizes[i]; }
  *toc_out = toc;
  return count;
}

/* OPUS_SET_GAIN: pcm *= 2^(gain/(256*6.0206)) (src/opus_decoder.c:686-700, float build) */
static void apply_decode_gain(const OrcDecoder *st, float *pcm, int n) {
  if (st->decode_gain) {
    const float gain = (float)exp(0.6931471805599453094 * (6.48814081e-4f * st->decode_gain));
    int i;
    for (i = 0; i < n * st->channels; i++) pcm[i] = pcm[i] * gain;
  }
}
/* the integer ctls of the decode path: 4034 OPUS_SET_GAIN, 4046 OPUS_SET_PHASE_INVERSION_DISABLED */
int orc_decoder_ctl(OrcDecoder *st, int request, int value) {
  if (request == 4034) { if (value < -32768 || value > 32767) return -1; st->decode_gain = value; return 0; }
  if (request == 4046) { if (value < 0 || value > 1) return -1; st->disable_inv = value; return 0; }
  return -5;
}
/* OPUS_GET_PITCH for a CELT-only stream (src/opus_decoder.c:1117-1133): the post-filter period, 0 before the first packet */
int orc_decoder_get_pitch(const OrcDecoder *st) { return st->prev_mode_celt ? st->postfilter_period : 0; }
int orc_decoder_get_last_packet_duration(const OrcDecoder *st) { return st->last_packet_duration; }

/* opus_decode_frame() with data == NULL (src/opus_decoder.c:316-365, 598-613, 701-712): conceal `frame_size` output-rate samples */
static int decode_frame_lost(OrcDecoder *st, float *pcm, int frame_size) {
  const int ds = st->downsample > 1 ? st->downsample : 1, F20 = 960 / ds, F10 = F20 >> 1, F5 = F10 >> 1, F2_5 = F5 >> 1;
  int audiosize, i, ret;
  if (frame_size < F2_5) return -2;
  frame_size = IMIN(frame_size, 48000 / ds / 25 * 3);
  frame_size = IMIN(frame_size, st->last_frame_size48 / ds); /* "don't conceal more than what the ToC says" (:316-321) */
  audiosize = frame_size;
  if (!st->prev_mode_celt) { /* no packet yet: zeros */
    for (i = 0; i < audiosize * st->channels; i++) pcm[i] = 0;
    return audiosize;
  }
  if (audiosize > F20) {
    do {
      ret = decode_frame_lost(st, pcm, IMIN(audiosize, F20));
      if (ret < 0) return ret;
      pcm += ret * st->channels;
      audiosize -= ret;
    } while (audiosize > 0);
    return frame_size;
  } else if (audiosize < F20) {
    if (audiosize > F10) audiosize = F10;
    else if (audiosize > F5 && audiosize < F10) audiosize = F5;
  }
  ret = orc_celt_decode_frame(st, NULL, 0, pcm, audiosize * ds);
  st->final_range = 0;
  if (ret > 0) apply_decode_gain(st, pcm, ret);
  return ret < 0 ? ret : audiosize;
}

/* Decode one Opus packet (CELT-only TOC configurations 16..31). pcm: interleaved float.
 * data == NULL or len == 0: the packet is lost, conceal max_frame_size samples (src/opus_decoder.c:757-778).
 * Returns samples per channel or a negative opus error code. */
int orc_opus_decode_float(OrcDecoder *st, const uint8_t *data, int len, float *pcm, int max_frame_size) {
  int toc, count, offsets[48], sizes[48], i, framesize, nb = 0, bw, endband;
  const int ds = st->downsample > 1 ? st->downsample : 1;
  if (len < 0) return -1;
  if (data == NULL || len == 0) {
    int pcm_count = 0;
    if (max_frame_size % (120 / ds) != 0) return -1;
    do {
      int ret = decode_frame_lost(st, pcm + pcm_count * st->channels, max_frame_size - pcm_count);
      if (ret < 0) return ret;
      pcm_count += ret;
    } while (pcm_count < max_frame_size);
    st->last_packet_duration = pcm_count;
    return pcm_count;
  }
  count = orc_packet_parse(data, len, &toc, offsets, sizes);
  if (count < 0) return count;
  if (!(toc & 0x80)) return -5; /* SILK / hybrid: not restated */
  framesize = (48000 << ((toc >> 3) & 3)) / 400;   /* at 48 kHz; max_frame_size and the return value are in output-rate samples */
  if (count * framesize / ds > max_frame_size) return -2;
  bw = (toc >> 5) & 3; /* CELT-only: 0 NB, 1 WB, 2 SWB, 3 FB  (src/opus.c opus_packet_get_bandwidth) */
  endband = bw == 0 ? 13 : bw == 1 ? 17 : bw == 2 ? 19 : 21;
  st->stream_channels = (toc & 0x4) ? 2 : 1;
  st->start = 0;
  st->last_frame_size48 = framesize; /* st->frame_size, src/opus_decoder.c:842 */
  for (i = 0; i < count; i++) {
    int ret;
    if (sizes[i] <= 1) { /* DTX / lost frame inside the packet: concealed at the packet's frame size; the end band stays */
      ret = decode_frame_lost(st, pcm + nb * st->channels, framesize / ds);
    } else {
      st->end = endband;
      ret = orc_celt_decode_frame(st, data + offsets[i], sizes[i], pcm + nb * st->channels, framesize);
      st->final_range = st->rng;
      st->prev_mode_celt = 1;
      if (ret > 0) apply_decode_gain(st, pcm + nb * st->channels, ret);
    }
    if (ret < 0) return ret;
    nb += ret;
  }
  st->last_packet_duration = nb;
  return nb;
}

uint32_t orc_final_range(const OrcDecoder *st) { return st->final_range; }
/* debug taps for the tests: the synthesis history (decode_mem[c][0 .. DEC_BUF + OVERLAP)) and the de-emphasis memory */
const float *orc_dbg_decode_mem(const OrcDecoder *st, int c) { return st->decode_mem[c]; }
float orc_dbg_preemph_mem(const OrcDecoder *st, int c) { return st->preemph_mem[c]; }

/* float -> int16 as the reference's int16 API does when nothing clips (celt/float_cast.h:150-158) */
void orc_float_to_int16(const float *in, int16_t *out, int n) {
  int i;
  for (i = 0; i < n; i++) {
    float x = in[i] * 32768.f;
    x = x > -32768.f ? x : -32768.f;
    x = x < 32767.f ? x : 32767.f;
    out[i] = (int16_t)lrintf(x);
  }
}

/* opus_pcm_soft_clip (src/opus.c:39-166): what opus_decode() applies to a packet's float PCM before the int16 conversion
 * (src/opus_decoder.c:875-878).  x: interleaved, N samples per channel; declip_mem[C] carries the non-linearity into the next call. */
void orc_pcm_soft_clip(float *x_, int N, int C, float *declip_mem) {
  int c, i;
  if (C < 1 || N < 1 || !x_ || !declip_mem) return;
  for (i = 0; i < N * C; i++) x_[i] = x_[i] > 2.f ? 2.f : x_[i] < -2.f ? -2.f : x_[i]; /* opus_limit2_checkwithin1_c */
  for (c = 0; c < C; c++) {
    float *x = x_ + c, a = declip_mem[c], x0;
    int curr = 0;
    for (i = 0; i < N; i++) { /* continue the previous frame's curve up to the first zero crossing */
      if (x[i * C] * a >= 0) break;
      x[i * C] = x[i * C] + a * x[i * C] * x[i * C];
    }
    x0 = x[0];
    while (1) {
      int start, end, special, peak_pos;
      float maxval;
      for (i = curr; i < N; i++)
        if (x[i * C] > 1 || x[i * C] < -1) break;
      if (i == N) { a = 0; break; }
      peak_pos = i;
      start = end = i;
      maxval = (float)fabs(x[i * C]);
      while (start > 0 && x[i * C] * x[(start - 1) * C] >= 0) start--;   /* zero crossing before the clipping */
      while (end < N && x[i * C] * x[end * C] >= 0) {                    /* ... and after, tracking the peak */
        if ((float)fabs(x[end * C]) > maxval) { maxval = (float)fabs(x[end * C]); peak_pos = end; }
        end++;
      }
      special = (start == 0 && x[i * C] * x[0] >= 0);   /* clipping before the first zero crossing */
      a = (maxval - 1) / (maxval * maxval);             /* maxval + a*maxval^2 = 1 */
      a += a * 2.4e-7f;
      if (x[i * C] > 0) a = -a;
      for (i = start; i < end; i++) x[i * C] = x[i * C] + a * x[i * C] * x[i * C];
      if (special && peak_pos >= 2) {   /* linear ramp from the first sample to the peak: no step at the frame start */
        float offset = x0 - x[0], delta = offset / peak_pos;
        for (i = curr; i < peak_pos; i++) {
          offset -= delta;
          x[i * C] += offset;
          x[i * C] = x[i * C] > 1.f ? 1.f : x[i * C] < -1.f ? -1.f : x[i * C];
        }
      }
      curr = end;
      if (curr == N) break;
    }
    declip_mem[c] = a;
  }
}

/* opus_decode() (int16): float decode, soft clip over the packet, FLOAT2INT16.  One stream, fresh decoder, like orc_decode_stream_float. */
int orc_decode_stream_int16(int Fs, int channels, int frame_size, int nframes, const uint8_t *pkts, int stride, const int *sizes, int16_t *pcm,
                            uint32_t *ranges) {
  OrcDecoder *st = (OrcDecoder *)malloc(sizeof(OrcDecoder));
  float *tmp = (float *)malloc(sizeof(float) * 5760 * 2), declip[2] = {0, 0};
  int f;
  orc_decoder_init(st, channels);
  if (orc_decoder_set_fs(st, Fs)) { free(st); free(tmp); return -1; }
  for (f = 0; f < nframes; f++) {
    int n = orc_opus_decode_float(st, sizes[f] ? pkts + (size_t)f * stride : NULL, sizes[f], tmp, frame_size);
    if (n < 0) { free(st); free(tmp); return n; }
    if (sizes[f]) orc_pcm_soft_clip(tmp, n, channels, declip); /* a lost packet returns before the soft clipping (src/opus_decoder.c:764-778) */
    orc_float_to_int16(tmp, pcm + (size_t)f * frame_size * channels, n * channels);
    ranges[f] = st->final_range;
  }
  free(st); free(tmp);
  return 0;
}

/* as orc_decode_stream_float_fs, with set-ctls (request, value pairs) applied first and the pitch / last-packet-duration getters read back */
int orc_decode_stream_float_ctl(int Fs, int channels, int frame_size, int nframes, const uint8_t *pkts, int stride, const int *sizes, float *pcm,
                                uint32_t *ranges, int nctl, const int *ctls, int *pitch, int *duration) {
  OrcDecoder *st = (OrcDecoder *)malloc(sizeof(OrcDecoder));
  int f, i;
  orc_decoder_init(st, channels);
  if (orc_decoder_set_fs(st, Fs)) { free(st); return -1; }
  for (i = 0; i < nctl; i++) if (orc_decoder_ctl(st, ctls[2 * i], ctls[2 * i + 1])) { free(st); return -1; }
  for (f = 0; f < nframes; f++) {
    int n = orc_opus_decode_float(st, sizes[f] ? pkts + (size_t)f * stride : NULL, sizes[f], pcm + (size_t)f * frame_size * channels, frame_size);
    if (n < 0) { free(st); return n; }
    ranges[f] = st->final_range;
    pitch[f] = orc_decoder_get_pitch(st);
    duration[f] = orc_decoder_get_last_packet_duration(st);
  }
  free(st);
  return 0;
}

/* Convenience for the tests: decode `nframes` packets of one stream with a fresh decoder. */
int orc_decode_stream_float(int channels, int frame_size, int nframes, const uint8_t *pkts, int stride,
                            const int *sizes, float *pcm, uint32_t *ranges) {
  OrcDecoder *st = (OrcDecoder *)malloc(sizeof(OrcDecoder));
  int f;
  orc_decoder_init(st, channels);
  for (f = 0; f < nframes; f++) {
    int n = orc_opus_decode_float(st, pkts + (size_t)f * stride, sizes[f], pcm + (size_t)f * frame_size * channels, frame_size);
    if (n < 0) { free(st); return n; }
    ranges[f] = st->final_range;
  }
  free(st);
  return 0;
}
/* the same with an output rate Fs <= 48000: frame_size and the PCM are in Fs samples */
int orc_decode_stream_float_fs(int Fs, int channels, int frame_size, int nframes, const uint8_t *pkts, int stride, const int *sizes, float *pcm,
                               uint32_t *ranges) {
  OrcDecoder *st = (OrcDecoder *)malloc(sizeof(OrcDecoder));
  int f;
  orc_decoder_init(st, channels);
  if (orc_decoder_set_fs(st, Fs)) { free(st); return -1; }
  for (f = 0; f < nframes; f++) {
    int n = orc_opus_decode_float(st, pkts + (size_t)f * stride, sizes[f], pcm + (size_t)f * frame_size * channels, frame_size);
    if (n < 0) { free(st); return n; }
    ranges[f] = st->final_range;
  }
  free(st);
  return 0;
}

#include "celt_enc.inc"
#include "opus_enc.inc"
#include "celt_plc.inc"
