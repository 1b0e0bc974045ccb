// celt_dec_core.cuh — one warp decodes one CELT frame: header symbols, energies, allocation and
// the band loop (theta splits, PVQ index -> pulses, spreading, folding, anti-collapse).
// Output: the normalised spectrum X (HBM) and the frame's side info for the synthesis kernel.
//
// Execution model: all scalar / control code (range decoder, allocation, recursion over band
// partitions) is warp-uniform — every lane runs it redundantly on identical registers, so there
// is no divergence and no broadcast traffic; vector work on a band (N <= 176) is lane-strided
// over shared memory (celt_vec.cuh).
//
// Reference behaviour restated here (decode side only):
//   celt/celt_decoder.c:1104-1617 (frame driver), celt/quant_bands.c:431-547 (energies),
//   celt/rate.c:249-646 (allocation), celt/bands.c:664-1923 (band loop), celt/vq.c:621-692,
//   celt/cwrs.c:467-545 (PVQ), src/opus.c:224-420 + src/opus_decoder.c:271-714 (Opus packet
//   layer, CELT-only subset).
#pragma once
#include "celt_types.h"
#include "celt_vec.cuh"
#include "rangedec.cuh"
#include "celt_tables_data.h"

namespace ob {

#define OB_BITRES 3
#define OB_IMIN(a, b) ((a) < (b) ? (a) : (b))
#define OB_IMAX(a, b) ((a) > (b) ? (a) : (b))

// Shared-memory working set of one warp (band decoder).  ~10 KB + packet.
struct BandWS {
  float norm[2 * 624];  // folding source: per-band normalised mid (or L/R in dual stereo)
  float xb[176];        // current band, channel 0 / mid
  float yb[176];        // current band, channel 1 / side
  float scr[176];       // lowband scratch
  float tmp[176];       // permutation / rotation scratch
  int iy[176];          // decoded pulses
  float oldE[2 * OB_NBANDS];
  int pulses[OB_NBANDS], fine_quant[OB_NBANDS], fine_prio[OB_NBANDS], tf_res[OB_NBANDS], cap[OB_NBANDS],
      offsets[OB_NBANDS];
  int bits1[OB_NBANDS], bits2[OB_NBANDS], thresh[OB_NBANDS], trim_off[OB_NBANDS], bits_tmp[OB_NBANDS];
  unsigned char collapse[2 * OB_NBANDS];
  unsigned char pad[2];
  // warp-uniform context of the band being decoded: written by lane 0 between bands (behind a
  // __syncwarp), read-only while the band's partitions are processed
  int c_band, c_intensity, c_spread, c_tf_change, c_disable_inv;
};

// The mutable part of the band decoder's state.  It travels BY VALUE through the (non-inlined,
// recursive) partition functions, i.e. in registers, never through local memory.
struct BandMut {
  RangeDec ec;
  int32_t remaining_bits;
  uint32_t seed;
  unsigned cm;  // collapse mask returned by the last band/partition call
};

// ---- exact integer helpers (celt/bands.c:68-91, celt/mathops.c:45-68, celt/rate.h:48-87) ----
OB_DEV int frac_mul16(int a, int b) { return (16384 + ((int32_t)(int16_t)a * (int16_t)b)) >> 15; }
OB_DEV int bitexact_cos(int x) {
  int32_t tmp = (4096 + ((int32_t)x * x)) >> 13;
  int x2 = tmp;
  x2 = (32767 - x2) + frac_mul16(x2, (-7651 + frac_mul16(x2, (8277 + frac_mul16(-626, x2)))));
  return 1 + x2;
}
OB_DEV int bitexact_log2tan(int isin, int icos) {
  int lc = ilog32(icos), ls = ilog32(isin);
  icos <<= 15 - lc;
  isin <<= 15 - ls;
  return (ls - lc) * (1 << 11) + frac_mul16(isin, frac_mul16(isin, -2597) + 7932) -
         frac_mul16(icos, frac_mul16(icos, -2597) + 7932);
}
OB_DEV unsigned isqrt_u32(uint32_t v) {
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
OB_DEV int pseudo_to_pulses(int i) { return i < 8 ? i : (8 + (i & 7)) << ((i >> 3) - 1); }
OB_DEV const unsigned char *pulse_cache_row(int band, int LM) {
  return ob_cache_bits + ob_cache_index[(LM + 1) * OB_NBANDS + band];
}
OB_DEV int bits_to_pseudo(int band, int LM, int bits) {
  const unsigned char *cache = pulse_cache_row(band, LM);
  int lo = 0, hi = cache[0];
  bits--;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    int mid = (lo + hi + 1) >> 1;
    if ((int)cache[mid] >= bits) hi = mid; else lo = mid;
  }
  if (bits - (lo == 0 ? -1 : (int)cache[lo]) <= (int)cache[hi] - bits) return lo;
  return hi;
}
OB_DEV int pseudo_to_bits(int band, int LM, int q) { return q == 0 ? 0 : pulse_cache_row(band, LM)[q] + 1; }

OB_DEV uint32_t pvq_u(int n, int k) {
  int a = OB_IMIN(n, k), b = OB_IMAX(n, k);
  return ob_pvq_u[ob_pvq_rowoff[a] + b];
}

// PVQ index -> pulse vector (celt/cwrs.c:467-541).  Warp-uniform; lane 0 stores.  Returns |y|^2.
OB_DEV float pvq_index_to_pulses(int n, int k, uint32_t i, int *y, int lane) {
  uint32_t p;
  int s, k0, val, yy = 0, pos = 0;
  while (n > 2) {
    uint32_t q;
    if (k >= n) {
      p = pvq_u(n, k + 1);
      s = -(int)(i >= p);
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
      if (lane == 0) y[pos] = val;
      pos++;
      yy += val * val;
    } else {
      p = pvq_u(k, n);
      q = pvq_u(k + 1, n);
      if (p <= i && i < q) {
        i -= p;
        if (lane == 0) y[pos] = 0;
        pos++;
      } else {
        s = -(int)(i >= q);
        i -= q & s;
        k0 = k;
        do p = pvq_u(--k, n); while (p > i);
        i -= p;
        val = (k0 - k + s) ^ s;
        if (lane == 0) y[pos] = val;
        pos++;
        yy += val * val;
      }
    }
    n--;
  }
  p = 2 * k + 1;
  s = -(int)(i >= p);
  i -= p & s;
  k0 = k;
  k = (i + 1) >> 1;
  if (k) i -= 2 * k - 1;
  val = (k0 - k + s) ^ s;
  if (lane == 0) y[pos] = val;
  pos++;
  yy += val * val;
  s = -(int)i;
  val = (k + s) ^ s;
  if (lane == 0) y[pos] = val;
  yy += val * val;
  return (float)yy;  // exact: integers < 2^24
}

OB_DEV float exp2_ref(float x) { return (float)exp(0.6931471805599453094 * (double)x); }  // celt/mathops.h:347

// celt/bands.c:638-662
OB_DEV int theta_levels(int N, int b, int offset, int pulse_cap, int stereo) {
  int N2 = 2 * N - 1;
  if (stereo && N == 2) N2--;
  int qb = (b + N2 * offset) / N2;
  qb = OB_IMIN(b - pulse_cap - (4 << OB_BITRES), qb);
  qb = OB_IMIN(8 << OB_BITRES, qb);
  if (qb < (1 << OB_BITRES >> 1)) return 1;
  int e = qb & 7;
  int t = e == 0 ? 16384 : e == 1 ? 17866 : e == 2 ? 19483 : e == 3 ? 21247 : e == 4 ? 23170 : e == 5 ? 25267
        : e == 6 ? 27554 : 30048;
  int qn = t >> (14 - (qb >> OB_BITRES));
  return (qn + 1) >> 1 << 1;
}

struct SplitInfo { int inv, imid, iside, delta, itheta, qalloc, b, fill; };
struct ThetaRet { RangeDec ec; SplitInfo sp; };

// decode side of celt/bands.c:700-933.  b and fill go in by value and come back in sp.b / sp.fill.
OB_DEVNI ThetaRet decode_theta(RangeDec ec, const BandWS *ws, int32_t remaining_bits, int N, int b, int B, int B0, int LM,
                              int stereo, int fill) {
  int i = ws->c_band, itheta = 0, inv = 0;
  int pulse_cap = ob_logN[i] + LM * (1 << OB_BITRES);
  int offset = (pulse_cap >> 1) - (stereo && N == 2 ? 16 : 4);
  int qn = theta_levels(N, b, offset, pulse_cap, stereo);
  if (stereo && i >= ws->c_intensity) qn = 1;
  int32_t tell = ec.tell_frac();
  if (qn != 1) {
    if (stereo && N > 2) {
      int p0 = 3, x, x0 = qn / 2, ft = p0 * (x0 + 1) + x0;
      int fs = ec.decode(ft);
      if (fs < (x0 + 1) * p0) x = fs / p0; else x = x0 + 1 + (fs - (x0 + 1) * p0);
      ec.update(x <= x0 ? p0 * x : (x - 1 - x0) + (x0 + 1) * p0, x <= x0 ? p0 * (x + 1) : (x - x0) + (x0 + 1) * p0, ft);
      itheta = x;
    } else if (B0 > 1 || stereo) {
      itheta = ec.uint(qn + 1);
    } else {
      int fs = 1, fl = 0;
      int ft = ((qn >> 1) + 1) * ((qn >> 1) + 1);
      int fm = ec.decode(ft);
      if (fm < ((qn >> 1) * ((qn >> 1) + 1) >> 1)) {
        itheta = (isqrt_u32(8 * (uint32_t)fm + 1) - 1) >> 1;
        fs = itheta + 1;
        fl = itheta * (itheta + 1) >> 1;
      } else {
        itheta = (2 * (qn + 1) - isqrt_u32(8 * (uint32_t)(ft - fm - 1) + 1)) >> 1;
        fs = qn + 1 - itheta;
        fl = ft - ((qn + 1 - itheta) * (qn + 2 - itheta) >> 1);
      }
      ec.update(fl, fl + fs, ft);
    }
    itheta = (uint32_t)((int32_t)itheta * 16384) / (uint32_t)qn;
  } else if (stereo) {
    if (b > 2 << OB_BITRES && remaining_bits > 2 << OB_BITRES) inv = ec.bit_logp(2);
    else inv = 0;
    if (ws->c_disable_inv) inv = 0;
    itheta = 0;
  }
  int qalloc = ec.tell_frac() - tell;
  b -= qalloc;
  int imid, iside, delta;
  if (itheta == 0) {
    imid = 32767; iside = 0; fill &= (1 << B) - 1; delta = -16384;
  } else if (itheta == 16384) {
    imid = 0; iside = 32767; fill &= ((1 << B) - 1) << B; delta = 16384;
  } else {
    imid = bitexact_cos((int16_t)itheta);
    iside = bitexact_cos((int16_t)(16384 - itheta));
    delta = frac_mul16((N - 1) << 7, bitexact_log2tan(iside, imid));
  }
  ThetaRet r;
  r.ec = ec;
  r.sp.inv = inv; r.sp.imid = imid; r.sp.iside = iside; r.sp.delta = delta; r.sp.itheta = itheta; r.sp.qalloc = qalloc;
  r.sp.b = b; r.sp.fill = fill;
  return r;
}

// celt/bands.c:934-968.  X, Y: shared memory.
OB_DEV BandMut decode_band_n1(BandMut m, float *X, float *Y, float *lowband_out, int lane) {
  int stereo = Y != nullptr;
  float *x = X;
  for (int c = 0; c < 1 + stereo; c++) {
    int sign = 0;
    if (m.remaining_bits >= 1 << OB_BITRES) {
      sign = m.ec.bits(1);
      m.remaining_bits -= 1 << OB_BITRES;
    }
    if (lane == 0) x[0] = sign ? -1.f : 1.f;
    x = Y;
  }
  wsync();
  if (lowband_out && lane == 0) lowband_out[0] = X[0];
  wsync();
  m.cm = 1;
  return m;
}

// celt/vq.c:621-692 (alg_unquant).  Returns the collapse mask.
OB_DEV unsigned pvq_decode(RangeDec &ec, BandWS *ws, float *X, int N, int K, int B, float gain, int lane) {
  int *iy = ws->iy;
  uint32_t idx = ec.uint(pvq_u(N, K) + pvq_u(N, K + 1));
  wsync();  // previous readers of iy are done
  float Ryy = pvq_index_to_pulses(N, K, idx, iy, lane);
  wsync();
  float g = rsqrt_ref(Ryy) * gain;
  unsigned mask = 0;
  const int N0 = B > 1 ? (unsigned)N / (unsigned)B : N;
#pragma unroll 1
  for (int j = lane; j < N; j += 32) {
    int v = iy[j];
    X[j] = (float)v * g;
    if (B > 1 && v != 0) mask |= 1u << (j / N0);
  }
  mask = B > 1 ? warp_or(mask) : 1u;
  spread_unrotate(X, ws->tmp, N, B, K, ws->c_spread, lane);
  wsync();
  return mask;
}

// decode side of celt/bands.c:973-1176 (recursive; depth <= 5)
OB_DEVNI BandMut decode_partition(BandMut m, BandWS *ws, float *X, int N, int b, int B, float *lowband, int LM, float gain,
                                 int fill) {
  const int lane = lane_id();
  const int i = ws->c_band;
  int B0 = B;
  unsigned cm = 0;
  const unsigned char *cache = pulse_cache_row(i, LM);
  if (LM != -1 && b > cache[cache[0]] + 12 && N > 2) {
    N >>= 1;
    float *Y = X + N;
    LM -= 1;
    if (B == 1) fill = (fill & 1) | (fill << 1);
    B = (B + 1) >> 1;
    ThetaRet tr = decode_theta(m.ec, ws, m.remaining_bits, N, b, B, B0, LM, 0, fill);
    m.ec = tr.ec;
    b = tr.sp.b; fill = tr.sp.fill;
    int delta = tr.sp.delta, itheta = tr.sp.itheta;
    float mid = (1.f / 32768) * tr.sp.imid;
    float side = (1.f / 32768) * tr.sp.iside;
    if (B0 > 1 && (itheta & 0x3fff)) {
      if (itheta > 8192) delta -= delta >> (4 - LM);
      else delta = OB_IMIN(0, delta + (N << OB_BITRES >> (5 - LM)));
    }
    int mbits = OB_IMAX(0, OB_IMIN(b, (b - delta) / 2));
    int sbits = b - mbits;
    m.remaining_bits -= tr.sp.qalloc;
    float *next_lowband2 = lowband ? lowband + N : nullptr;
    int32_t rebalance = m.remaining_bits;
    if (mbits >= sbits) {
      m = decode_partition(m, ws, X, N, mbits, B, lowband, LM, gain * mid, fill);
      cm = m.cm;
      rebalance = mbits - (rebalance - m.remaining_bits);
      if (rebalance > 3 << OB_BITRES && itheta != 0) sbits += rebalance - (3 << OB_BITRES);
      m = decode_partition(m, ws, Y, N, sbits, B, next_lowband2, LM, gain * side, fill >> B);
      cm |= m.cm << (B0 >> 1);
    } else {
      m = decode_partition(m, ws, Y, N, sbits, B, next_lowband2, LM, gain * side, fill >> B);
      cm = m.cm << (B0 >> 1);
      rebalance = sbits - (rebalance - m.remaining_bits);
      if (rebalance > 3 << OB_BITRES && itheta != 16384) mbits += rebalance - (3 << OB_BITRES);
      m = decode_partition(m, ws, X, N, mbits, B, lowband, LM, gain * mid, fill);
      cm |= m.cm;
    }
  } else {
    int q = bits_to_pseudo(i, LM, b);
    int curr_bits = pseudo_to_bits(i, LM, q);
    m.remaining_bits -= curr_bits;
#pragma unroll 1
    while (m.remaining_bits < 0 && q > 0) {
      m.remaining_bits += curr_bits;
      q--;
      curr_bits = pseudo_to_bits(i, LM, q);
      m.remaining_bits -= curr_bits;
    }
    if (q != 0) {
      cm = pvq_decode(m.ec, ws, X, N, pseudo_to_pulses(q), B, gain, lane);
    } else {
      unsigned cm_mask = (unsigned)(1UL << B) - 1;
      fill &= cm_mask;
      wsync();
      if (!fill) {
#pragma unroll 1
        for (int j = lane; j < N; j += 32) X[j] = 0.f;
        wsync();
      } else {
        // every lane steps the LCG through all N draws and keeps its own elements
        uint32_t seed = m.seed;
        if (lowband == nullptr) {
#pragma unroll 1
          for (int j = 0; j < N; j++) {
            seed = 1664525u * seed + 1013904223u;
            if ((j & 31) == lane) X[j] = (float)((int32_t)seed >> 20);
          }
          cm = cm_mask;
        } else {
#pragma unroll 1
          for (int j = 0; j < N; j++) {
            seed = 1664525u * seed + 1013904223u;
            if ((j & 31) == lane) {
              float t = 1.0f / 256;
              t = (seed & 0x8000) ? t : -t;
              X[j] = lowband[j] + t;
            }
          }
          cm = fill;
        }
        m.seed = seed;
        renormalise(X, N, gain, lane);
      }
    }
  }
  m.cm = cm;
  return m;
}

// decode side of celt/bands.c:1248-1379
OB_DEVNI BandMut decode_band_mono(BandMut m, BandWS *ws, float *X, int N, int b, int B, float *lowband, int LM,
                                 float *lowband_out, float gain, float *lowband_scratch, int fill) {
  const int lane = lane_id();
  float *tmp = ws->tmp;
  int N0 = N, N_B = N, N_B0, B0 = B, time_divide = 0, recombine = 0;
  int tf_change = ws->c_tf_change;
  const int longBlocks = B0 == 1;
  N_B = (unsigned)N_B / (unsigned)B;
  if (N == 1) return decode_band_n1(m, X, nullptr, lowband_out, lane);
  if (tf_change > 0) recombine = tf_change;
  if (lowband_scratch && lowband && (recombine || ((N_B & 1) == 0 && tf_change < 0) || B0 > 1)) {
    wsync();
#pragma unroll 1
    for (int j = lane; j < N; j += 32) lowband_scratch[j] = lowband[j];
    wsync();
    lowband = lowband_scratch;
  }
#pragma unroll 1
  for (int k = 0; k < recombine; k++) {
    if (lowband) haar_step(lowband, N >> k, 1 << k, lane);
    int lo = fill & 0xF, hi = fill >> 4;  // bit_interleave_table, celt/bands.c:1288
    int tl = lo == 0 ? 0 : (lo & 3 ? 1 : 0) | (lo & 12 ? 2 : 0);
    int th = hi == 0 ? 0 : (hi & 3 ? 1 : 0) | (hi & 12 ? 2 : 0);
    fill = tl | th << 2;
  }
  B >>= recombine;
  N_B <<= recombine;
#pragma unroll 1
  while ((N_B & 1) == 0 && tf_change < 0) {
    if (lowband) haar_step(lowband, N_B, B, lane);
    fill |= fill << B;
    B <<= 1;
    N_B >>= 1;
    time_divide++;
    tf_change++;
  }
  B0 = B;
  N_B0 = N_B;
  if (B0 > 1 && lowband) deinterleave_hadamard(lowband, tmp, N_B >> recombine, B0 << recombine, longBlocks, lane);
  m = decode_partition(m, ws, X, N, b, B, lowband, LM, gain, fill);
  unsigned cm = m.cm;
  if (B0 > 1) interleave_hadamard(X, tmp, N_B >> recombine, B0 << recombine, longBlocks, lane);
  N_B = N_B0;
  B = B0;
#pragma unroll 1
  for (int k = 0; k < time_divide; k++) {
    B >>= 1;
    N_B <<= 1;
    cm |= cm >> B;
    haar_step(X, N_B, B, lane);
  }
#pragma unroll 1
  for (int k = 0; k < recombine; k++) {
    // bit_deinterleave_table, celt/bands.c:1352: each of the low 4 bits is doubled
    unsigned c4 = cm & 0xF;
    cm = (c4 & 1 ? 0x03u : 0) | (c4 & 2 ? 0x0Cu : 0) | (c4 & 4 ? 0x30u : 0) | (c4 & 8 ? 0xC0u : 0);
    haar_step(X, N0 >> k, 1 << k, lane);
  }
  B <<= recombine;
  if (lowband_out) {
    float n = f_sqrt((float)N0);
    wsync();
#pragma unroll 1
    for (int j = lane; j < N0; j += 32) lowband_out[j] = n * X[j];
    wsync();
  }
  cm &= (1u << B) - 1;
  m.cm = cm;
  return m;
}

// celt/bands.c:407-466 (stereo_merge)
OB_DEVNI void merge_mid_side(float *X, float *Y, float mid, int N, int lane) {
  wsync();
  float xp = warp_dot(Y, X, N, lane);
  float side = warp_dot(Y, Y, N, lane);
  xp = mid * xp;
  float El = (mid * mid) + side - 2 * xp;
  float Er = (mid * mid) + side + 2 * xp;
  if (Er < 6e-4f || El < 6e-4f) {
#pragma unroll 1
    for (int j = lane; j < N; j += 32) Y[j] = X[j];
    wsync();
    return;
  }
  float lgain = rsqrt_ref(El), rgain = rsqrt_ref(Er);
#pragma unroll 1
  for (int j = lane; j < N; j += 32) {
    float l = mid * X[j], r = Y[j];
    X[j] = lgain * (l - r);
    Y[j] = rgain * (l + r);
  }
  wsync();
}

// decode side of celt/bands.c:1387-1570
OB_DEVNI BandMut decode_band_stereo(BandMut m, BandWS *ws, float *X, float *Y, int N, int b, int B, float *lowband, int LM,
                                   float *lowband_out, float *lowband_scratch, int fill) {
  const int lane = lane_id();
  unsigned cm = 0;
  if (N == 1) return decode_band_n1(m, X, Y, lowband_out, lane);
  int orig_fill = fill;
  ThetaRet tr = decode_theta(m.ec, ws, m.remaining_bits, N, b, B, B, LM, 1, fill);
  m.ec = tr.ec;
  b = tr.sp.b; fill = tr.sp.fill;
  int inv = tr.sp.inv, delta = tr.sp.delta, itheta = tr.sp.itheta, qalloc = tr.sp.qalloc;
  float mid = (1.f / 32768) * tr.sp.imid;
  float side = (1.f / 32768) * tr.sp.iside;
  if (N == 2) {
    int mbits = b, sbits = 0, sign = 0;
    if (itheta != 0 && itheta != 16384) sbits = 1 << OB_BITRES;
    mbits -= sbits;
    int c = itheta > 8192;
    m.remaining_bits -= qalloc + sbits;
    float *x2 = c ? Y : X;
    float *y2 = c ? X : Y;
    if (sbits) sign = m.ec.bits(1);
    sign = 1 - 2 * sign;
    m = decode_band_mono(m, ws, x2, N, mbits, B, lowband, LM, lowband_out, 1.0f, lowband_scratch, orig_fill);
    cm = m.cm;
    wsync();
    if (lane == 0) {
      y2[0] = -sign * x2[1];
      y2[1] = sign * x2[0];
      X[0] = mid * X[0];
      X[1] = mid * X[1];
      Y[0] = side * Y[0];
      Y[1] = side * Y[1];
      float t = X[0];
      X[0] = t - Y[0];
      Y[0] = t + Y[0];
      t = X[1];
      X[1] = t - Y[1];
      Y[1] = t + Y[1];
    }
    wsync();
  } else {
    int mbits = OB_IMAX(0, OB_IMIN(b, (b - delta) / 2));
    int sbits = b - mbits;
    m.remaining_bits -= qalloc;
    int32_t rebalance = m.remaining_bits;
    // one call site per role (mid / side) keeps the code small: order is decided by `first`
    const int mid_first = mbits >= sbits;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
      const int do_mid = (pass == 0) == (mid_first != 0);
      if (do_mid) {
        m = decode_band_mono(m, ws, X, N, mbits, B, lowband, LM, lowband_out, 1.0f, lowband_scratch, fill);
        if (pass == 0) {
          cm = m.cm;
          rebalance = mbits - (rebalance - m.remaining_bits);
          if (rebalance > 3 << OB_BITRES && itheta != 0) sbits += rebalance - (3 << OB_BITRES);
        } else cm |= m.cm;
      } else {
        m = decode_band_mono(m, ws, Y, N, sbits, B, nullptr, LM, nullptr, side, nullptr, fill >> B);
        if (pass == 0) {
          cm = m.cm;
          rebalance = sbits - (rebalance - m.remaining_bits);
          if (rebalance > 3 << OB_BITRES && itheta != 16384) mbits += rebalance - (3 << OB_BITRES);
        } else cm |= m.cm;
      }
    }
  }
  if (N != 2) merge_mid_side(X, Y, mid, N, lane);
  if (inv) {
    wsync();
#pragma unroll 1
    for (int j = lane; j < N; j += 32) Y[j] = -Y[j];
    wsync();
  }
  m.cm = cm;
  return m;
}

// decode side of celt/bands.c:1589-1923.  Xg/Yg: this frame's normalised spectrum in HBM
// (each band is produced in shared memory and stored once it is final).
OB_DEV void decode_all_bands(BandMut &m, BandWS *wsp, int start, int end, float *Xg, float *Yg, int shortBlocks,
                             int dual_stereo, int32_t total_bits, int32_t balance, int LM, int codedBands, int lane) {
  BandWS &ws = *wsp;
  const int M = 1 << LM, C = Yg != nullptr ? 2 : 1;
  const int B = shortBlocks ? M : 1;
  const int norm_offset = M * ob_ebands[start];
  float *norm = ws.norm, *norm2 = ws.norm + (M * ob_ebands[OB_NBANDS - 1] - norm_offset);
  int lowband_offset = 0, update_lowband = 1;
  const int spread = ws.c_spread, intensity = ws.c_intensity;
#pragma unroll 1
  for (int i = start; i < end; i++) {
    const int last = (i == end - 1);
    const int N = M * ob_ebands[i + 1] - M * ob_ebands[i];
    int32_t tell = m.ec.tell_frac();
    if (i != start) balance -= tell;
    int32_t remaining_bits = total_bits - tell - 1;
    m.remaining_bits = remaining_bits;
    int b;
    if (i <= codedBands - 1) {
      int32_t curr_balance = balance / OB_IMIN(3, codedBands - i);
      b = OB_IMAX(0, OB_IMIN(16383, OB_IMIN(remaining_bits + 1, ws.pulses[i] + curr_balance)));
    } else b = 0;
    if ((M * ob_ebands[i] - N >= M * ob_ebands[start] || i == start + 1) && (update_lowband || lowband_offset == 0))
      lowband_offset = i;
    // (special_hybrid_folding, bands.c:1572-1586, copies nothing when start == 0)
    const int tf_change = ws.tf_res[i];
    if (lane == 0) { ws.c_band = i; ws.c_tf_change = tf_change; }
    float *lowband_scratch = last ? nullptr : ws.scr;
    int effective_lowband = -1;
    unsigned x_cm, y_cm;
    if (lowband_offset != 0 && (spread != 3 || B > 1 || tf_change < 0)) {
      effective_lowband = OB_IMAX(0, M * ob_ebands[lowband_offset] - norm_offset - N);
      int fold_start = lowband_offset;
      while (M * ob_ebands[--fold_start] > effective_lowband + norm_offset) {}
      int fold_end = lowband_offset - 1;
      while (++fold_end < i && M * ob_ebands[fold_end] < effective_lowband + norm_offset + N) {}
      x_cm = y_cm = 0;
      int fold_i = fold_start;
      do {
        x_cm |= ws.collapse[fold_i * C + 0];
        y_cm |= ws.collapse[fold_i * C + C - 1];
      } while (++fold_i < fold_end);
    } else x_cm = y_cm = (1u << B) - 1;
    wsync();  // c_band / c_tf_change visible
    if (dual_stereo && i == intensity) {
      dual_stereo = 0;
#pragma unroll 1
      for (int j = lane; j < M * ob_ebands[i] - norm_offset; j += 32) norm[j] = .5f * (norm[j] + norm2[j]);
      wsync();
    }
    float *X = ws.xb, *Y = ws.yb;
    float *lb1 = effective_lowband != -1 ? norm + effective_lowband : nullptr;
    float *lb2 = effective_lowband != -1 ? norm2 + effective_lowband : nullptr;
    float *lo1 = last ? nullptr : norm + M * ob_ebands[i] - norm_offset;
    float *lo2 = last ? nullptr : norm2 + M * ob_ebands[i] - norm_offset;
    if (dual_stereo) {
      m = decode_band_mono(m, wsp, X, N, b / 2, B, lb1, LM, lo1, 1.0f, lowband_scratch, x_cm);
      x_cm = m.cm;
      m = decode_band_mono(m, wsp, Y, N, b / 2, B, lb2, LM, lo2, 1.0f, lowband_scratch, y_cm);
      y_cm = m.cm;
    } else {
      if (C == 2) m = decode_band_stereo(m, wsp, X, Y, N, b, B, lb1, LM, lo1, lowband_scratch, x_cm | y_cm);
      else m = decode_band_mono(m, wsp, X, N, b, B, lb1, LM, lo1, 1.0f, lowband_scratch, x_cm | y_cm);
      x_cm = y_cm = m.cm;
    }
    wsync();
    // the band is final: store it (coalesced) to the frame's spectrum in HBM
    {
      float *gx = Xg + M * ob_ebands[i];
#pragma unroll 1
      for (int j = lane; j < N; j += 32) gx[j] = X[j];
      if (C == 2) {
        float *gy = Yg + M * ob_ebands[i];
#pragma unroll 1
        for (int j = lane; j < N; j += 32) gy[j] = Y[j];
      }
    }
    if (lane == 0) {
      ws.collapse[i * C + 0] = (unsigned char)x_cm;
      ws.collapse[i * C + C - 1] = (unsigned char)y_cm;
    }
    wsync();
    balance += ws.pulses[i] + tell;
    update_lowband = b > (N << OB_BITRES);
  }
}

// ---- bit allocation: celt/rate.c:249-646 (decode side).  Arrays in shared memory; warp-uniform,
// lane 0 stores. -------------------------------------------------------------------------------
OB_DEV int compute_allocation(RangeDec &ec, BandWS *wsp, int lane, int start, int end, int alloc_trim, int &intensity,
                              int &dual_stereo, int32_t total, int32_t &balance_out, int C, int LM) {
  BandWS &ws = *wsp;
  const int len = OB_NBANDS, nbAllocVectors = 11;
  int skip_start = start, intensity_rsv = 0, dual_stereo_rsv = 0;
  total = OB_IMAX(total, 0);
  int skip_rsv = total >= 1 << OB_BITRES ? 1 << OB_BITRES : 0;
  total -= skip_rsv;
  if (C == 2) {
    intensity_rsv = ob_log2_frac_tab[end - start];
    if (intensity_rsv > total) intensity_rsv = 0;
    else {
      total -= intensity_rsv;
      dual_stereo_rsv = total >= 1 << OB_BITRES ? 1 << OB_BITRES : 0;
      total -= dual_stereo_rsv;
    }
  }
  const int *offsets = ws.offsets, *cap = ws.cap;
  int *bits1 = ws.bits1, *bits2 = ws.bits2, *thresh = ws.thresh, *trim_offset = ws.trim_off;
  int *bits = ws.bits_tmp, *pulses_out = ws.pulses, *ebits = ws.fine_quant, *fine_priority = ws.fine_prio;
  wsync();
  // per-band constants: one band per lane
  for (int j = start + lane; j < end; j += 32) {
    int w = ob_ebands[j + 1] - ob_ebands[j];
    thresh[j] = OB_IMAX((C) << OB_BITRES, (3 * w << LM << OB_BITRES) >> 4);
    int t = C * w * (alloc_trim - 5 - LM) * (end - j - 1) * (1 << (LM + OB_BITRES)) >> 6;
    if (w << LM == 1) t -= C << OB_BITRES;
    trim_offset[j] = t;
  }
  wsync();
  int lo = 1, hi = nbAllocVectors - 1;
  do {
    int done = 0, psum = 0, mid = (lo + hi) >> 1;
    for (int j = end; j-- > start;) {
      int N = ob_ebands[j + 1] - ob_ebands[j];
      int bitsj = C * N * ob_band_alloc[mid * len + j] << LM >> 2;
      if (bitsj > 0) bitsj = OB_IMAX(0, bitsj + trim_offset[j]);
      bitsj += offsets[j];
      if (bitsj >= thresh[j] || done) { done = 1; psum += OB_IMIN(bitsj, cap[j]); }
      else if (bitsj >= C << OB_BITRES) psum += C << OB_BITRES;
    }
    if (psum > total) hi = mid - 1; else lo = mid + 1;
  } while (lo <= hi);
  hi = lo--;
  for (int j = start; j < end; j++) {
    if (offsets[j] > 0) skip_start = j;
  }
  for (int j = start + lane; j < end; j += 32) {
    int N = ob_ebands[j + 1] - ob_ebands[j];
    int bits1j = C * N * ob_band_alloc[lo * len + j] << LM >> 2;
    int bits2j = hi >= nbAllocVectors ? cap[j] : C * N * ob_band_alloc[hi * len + j] << LM >> 2;
    if (bits1j > 0) bits1j = OB_IMAX(0, bits1j + trim_offset[j]);
    if (bits2j > 0) bits2j = OB_IMAX(0, bits2j + trim_offset[j]);
    if (lo > 0) bits1j += offsets[j];
    bits2j += offsets[j];
    bits2j = OB_IMAX(0, bits2j - bits1j);
    bits1[j] = bits1j;
    bits2[j] = bits2j;
  }
  wsync();
  // ---- interp_bits2pulses (rate.c:249-533) ----
  int32_t psum;
  const int logM = LM << OB_BITRES, stereo = C > 1, alloc_floor = C << OB_BITRES;
  lo = 0;
  hi = 1 << 6;
  for (int it = 0; it < 6; it++) {
    int mid = (lo + hi) >> 1, done = 0;
    psum = 0;
    for (int j = end; j-- > start;) {
      int tmp = bits1[j] + (mid * (int32_t)bits2[j] >> 6);
      if (tmp >= thresh[j] || done) { done = 1; psum += OB_IMIN(tmp, cap[j]); }
      else if (tmp >= alloc_floor) psum += alloc_floor;
    }
    if (psum > total) hi = mid; else lo = mid;
  }
  psum = 0;
  {
    int done = 0;
    for (int j = end; j-- > start;) {
      int tmp = bits1[j] + ((int32_t)lo * bits2[j] >> 6);
      if (tmp < thresh[j] && !done) tmp = tmp >= alloc_floor ? alloc_floor : 0;
      else done = 1;
      tmp = OB_IMIN(tmp, cap[j]);
      if (lane == 0) bits[j] = tmp;
      psum += tmp;
    }
  }
  wsync();
  int codedBands;
  for (codedBands = end;; codedBands--) {
    int j = codedBands - 1;
    if (j <= skip_start) { total += skip_rsv; break; }
    int32_t left = total - psum;
    int32_t percoeff = (uint32_t)left / (uint32_t)(ob_ebands[codedBands] - ob_ebands[start]);
    left -= (ob_ebands[codedBands] - ob_ebands[start]) * percoeff;
    int rem = OB_IMAX(left - (ob_ebands[j] - ob_ebands[start]), 0);
    int band_width = ob_ebands[codedBands] - ob_ebands[j];
    int bj = bits[j];
    int band_bits = (int)(bj + percoeff * band_width + rem);
    if (band_bits >= OB_IMAX(thresh[j], alloc_floor + (1 << OB_BITRES))) {
      if (ec.bit_logp(1)) break;
      psum += 1 << OB_BITRES;
      band_bits -= 1 << OB_BITRES;
    }
    psum -= bj + intensity_rsv;
    if (intensity_rsv > 0) intensity_rsv = ob_log2_frac_tab[j - start];
    psum += intensity_rsv;
    int nb;
    if (band_bits >= alloc_floor) { psum += alloc_floor; nb = alloc_floor; }
    else nb = 0;
    wsync();
    if (lane == 0) bits[j] = nb;
    wsync();
  }
  if (intensity_rsv > 0) intensity = start + ec.uint(codedBands + 1 - start);
  else intensity = 0;
  if (intensity <= start) { total += dual_stereo_rsv; dual_stereo_rsv = 0; }
  if (dual_stereo_rsv > 0) dual_stereo = ec.bit_logp(1);
  else dual_stereo = 0;
  int32_t left = total - psum;
  int32_t percoeff = (uint32_t)left / (uint32_t)(ob_ebands[codedBands] - ob_ebands[start]);
  left -= (ob_ebands[codedBands] - ob_ebands[start]) * percoeff;
  int32_t balance = 0;
  // the remaining per-band arithmetic is a serial chain through `left` and `balance`
  for (int j = start; j < codedBands; j++) {
    int w = ob_ebands[j + 1] - ob_ebands[j];
    int bj = bits[j] + (int)percoeff * w;
    int tmp = (int)OB_IMIN(left, w);
    bj += tmp;
    left -= tmp;
    int N0 = w, N = N0 << LM, ej, fp;
    int32_t excess, bit = (int32_t)bj + balance;
    if (N > 1) {
      excess = OB_IMAX(bit - cap[j], 0);
      bj = bit - excess;
      int den = (C * N + ((C == 2 && N > 2 && !dual_stereo && j < intensity) ? 1 : 0));
      int NClogN = den * (ob_logN[j] + logM);
      int offset = (NClogN >> 1) - den * 21;
      if (N == 2) offset += den << OB_BITRES >> 2;
      if (bj + offset < den * 2 << OB_BITRES) offset += NClogN >> 2;
      else if (bj + offset < den * 3 << OB_BITRES) offset += NClogN >> 3;
      ej = OB_IMAX(0, (bj + offset + (den << (OB_BITRES - 1))));
      ej = ((uint32_t)ej / (uint32_t)den) >> OB_BITRES;
      if (C * ej > (bj >> OB_BITRES)) ej = bj >> stereo >> OB_BITRES;
      ej = OB_IMIN(ej, 8);
      fp = ej * (den << OB_BITRES) >= bj + offset;
      bj -= C * ej << OB_BITRES;
    } else {
      excess = OB_IMAX(0, bit - (C << OB_BITRES));
      bj = bit - excess;
      ej = 0;
      fp = 1;
    }
    if (excess > 0) {
      int extra_fine = OB_IMIN(excess >> (stereo + OB_BITRES), 8 - ej);
      ej += extra_fine;
      int extra_bits = extra_fine * C << OB_BITRES;
      fp = extra_bits >= excess - balance;
      excess -= extra_bits;
    }
    balance = excess;
    if (lane == 0) { pulses_out[j] = bj; ebits[j] = ej; fine_priority[j] = fp; }
  }
  balance_out = balance;
  for (int j = codedBands; j < end; j++) {
    int bj = bits[j];
    int ej = bj >> stereo >> OB_BITRES;
    if (lane == 0) { ebits[j] = ej; pulses_out[j] = 0; fine_priority[j] = ej < 1; }
  }
  wsync();
  return codedBands;
}

// celt/bands.c:259-353; X in HBM (this warp wrote it).  Rare path (transient frames that collapsed).
OB_DEV void anti_collapse(BandWS *wsp, int lane, float *Xg, int LM, int C, int size, int start, int end, const float *logE,
                          const float *prev1logE, const float *prev2logE, uint32_t seed) {
  BandWS &ws = *wsp;
  wsync();
  for (int i = start; i < end; i++) {
    int N0 = ob_ebands[i + 1] - ob_ebands[i];
    int depth = ((uint32_t)(1 + ws.pulses[i]) / (uint32_t)N0) >> LM;
    float thresh = .5f * exp2_ref(-.125f * depth);
    float sqrt_1 = rsqrt_ref((float)(N0 << LM));
    for (int c = 0; c < C; c++) {
      float prev1 = prev1logE[c * OB_NBANDS + i], prev2 = prev2logE[c * OB_NBANDS + i];
      if (C == 1) {
        float o1 = prev1logE[OB_NBANDS + i], o2 = prev2logE[OB_NBANDS + i];
        prev1 = prev1 > o1 ? prev1 : o1;
        prev2 = prev2 > o2 ? prev2 : o2;
      }
      float Ediff = logE[c * OB_NBANDS + i] - (prev1 < prev2 ? prev1 : prev2);
      Ediff = Ediff > 0 ? Ediff : 0;
      float r = 2.f * exp2_ref(-Ediff);
      if (LM == 3) r *= 1.41421356f;
      r = thresh < r ? thresh : r;
      r = r * sqrt_1;
      float *X = Xg + c * size + (ob_ebands[i] << LM);
      int renorm = 0;
      unsigned mask = ws.collapse[i * C + c];
      for (int k = 0; k < 1 << LM; k++) {
        if (!(mask & 1u << k)) {
          for (int j = 0; j < N0; j++) {
            seed = 1664525u * seed + 1013904223u;
            if (lane == 0) X[(j << LM) + k] = (seed & 0x8000 ? r : -r);
          }
          renorm = 1;
        }
      }
      if (renorm) renormalise(X, N0 << LM, 1.0f, lane);
    }
  }
}

}  // namespace ob
