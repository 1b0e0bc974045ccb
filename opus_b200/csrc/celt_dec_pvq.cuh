// celt_dec_pvq.cuh — kernel A2: ONE WARP PER STREAM turns the entropy stage's script into the
// frame's normalised spectrum X[C][N].
//
// Phase 1 (lane per leaf): every PVQ leaf of the frame is independent given its codeword index, so
//   lane l handles leaves l, l+32, ...: index -> pulses (celt/cwrs.c:467-541), unit-norm scaling
//   (celt/vq.c:150-181), inverse spreading rotation (celt/vq.c:75-147) and the collapse mask
//   (celt/vq.c:183-203), each run SEQUENTIALLY by its lane in the reference's exact order — the
//   serial Givens chains that cannot be split across lanes without changing the rounding become
//   32 chains in flight, and the results are bit-identical to the scalar reference.
// Phase 2 (band order, warp cooperative): fill / collapse-mask propagation through the partition
//   tree, zero / noise / folded leaves (the LCG and the folding source are band-order dependent),
//   Haar and Hadamard (de)interleaving, mid/side merge, the folding buffer, anti-collapse; one
//   coalesced store of the whole frame at the end; the per-stream energy state update.
//
// Reference (resynthesis side only): celt/bands.c:259-353, 407-466, 574-636, 934-1923,
// celt/vq.c:75-203,621-722, celt/cwrs.c:467-541, celt/celt_decoder.c:1527-1600.
#pragma once
#include "celt_types.h"
#include "celt_vec.cuh"
#include "celt_tables_data.h"

namespace ob {

#ifndef OB_BITRES
#define OB_BITRES 3
#define OB_IMIN(a, b) ((a) < (b) ? (a) : (b))
#define OB_IMAX(a, b) ((a) > (b) ? (a) : (b))
#endif

struct DecPvqParams {
  const ObFrameSide *side;  // [n_streams]
  const ObScript *script;   // [n_streams]
  ObDecBandState *state;    // [n_streams]
  float *X;                 // [n_streams][2][960]
  int n_streams;
};

// Shared by the warps of a CTA: the PVQ U(n,k) table (celt/cwrs.c:197-211), ~5 KB
struct PvqShared {
  uint32_t u[1272];
  short rowoff[16];
  unsigned long long bar;   // the mbarrier the bulk copy of `u` completes on (it lives in the dynamic allocation: the kernel opts in to ALL of the
                            // SM's shared memory, a static variable would not fit beside it)
};
// Per warp
struct PvqWS {
  float X[2 * OB_MAXN];  // the frame's normalised spectrum, [C][N]
  float norm[2 * 624];   // folding source
  float scr[176];        // lowband scratch
  float tmp[176];        // permutation scratch
  unsigned short leaf_cm[OB_MAX_LEAVES];  // up to 16 blocks after a time-divide (bands.c:1303)
  ObBandRec bands[OB_NBANDS];             // staged copy of the script's band records
  unsigned char collapse[2 * OB_NBANDS];
  unsigned char pad[6];
};

OB_DEV float exp2_ref(float x) { return (float)exp(0.6931471805599453094 * (double)x); }  // celt/mathops.h:347

OB_DEV uint32_t pvq_u_s(const PvqShared *t, int n, int k) {
  int a = OB_IMIN(n, k), b = OB_IMAX(n, k);
  return t->u[t->rowoff[a] + b];
}

// index -> pulses written as floats; returns |y|^2 (exact integer).  celt/cwrs.c:467-541
OB_DEV int pvq_index_to_pulses_f(const PvqShared *t, int n, int k, uint32_t i, float *y) {
  uint32_t p;
  int s, k0, val, yy = 0;
#pragma unroll 1
  while (n > 2) {
    uint32_t q;
    if (k >= n) {
      p = pvq_u_s(t, n, k + 1);
      s = -(int)(i >= p);
      i -= p & s;
      k0 = k;
      q = pvq_u_s(t, n, n);
      if (q > i) {
        k = n;
#pragma unroll 1
        do p = pvq_u_s(t, --k, n); while (p > i);
      } else {
#pragma unroll 1
        for (p = pvq_u_s(t, n, k); p > i; p = pvq_u_s(t, n, k)) k--;
      }
      i -= p;
      val = (k0 - k + s) ^ s;
      *y++ = (float)val;
      yy += val * val;
    } else {
      p = pvq_u_s(t, k, n);
      q = pvq_u_s(t, k + 1, n);
      if (p <= i && i < q) {
        i -= p;
        *y++ = 0.f;
      } else {
        s = -(int)(i >= q);
        i -= q & s;
        k0 = k;
#pragma unroll 1
        do p = pvq_u_s(t, --k, n); while (p > i);
        i -= p;
        val = (k0 - k + s) ^ s;
        *y++ = (float)val;
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
  *y++ = (float)val;
  yy += val * val;
  s = -(int)i;
  val = (k + s) ^ s;
  *y = (float)val;
  yy += val * val;
  return yy;
}

// celt/vq.c:75-102, one lane, reference order
OB_DEV void rotate_seq(float *X, int len, int stride, float c, float s) {
  const float ms = -s;
  float *p = X;
#pragma unroll 1
  for (int i = 0; i < len - stride; i++) {
    float x1 = p[0], x2 = p[stride];
    p[stride] = c * x2 + s * x1;
    *p++ = c * x1 + ms * x2;
  }
  p = &X[len - 2 * stride - 1];
#pragma unroll 1
  for (int i = len - 2 * stride - 1; i >= 0; i--) {
    float x1 = p[0], x2 = p[stride];
    p[stride] = c * x2 + s * x1;
    *p-- = c * x1 + ms * x2;
  }
}
// celt/vq.c:104-147 with dir = -1
OB_DEV void unrotate_seq(float *X, int len, int stride, int K, int spread) {
  if (2 * K >= len || spread == 0) return;
  const int factor = spread == 1 ? 15 : spread == 2 ? 10 : 5;
  float gain = f_div((float)len, (float)(len + factor * K));
  float theta = .5f * (gain * gain);
  float c = cos_norm_ref(theta);
  float s = cos_norm_ref(1.0f - theta);
  int stride2 = 0;
  if (len >= 8 * stride) {
    stride2 = 1;
    while ((stride2 * stride2 + stride2) * stride + (stride >> 2) < len) stride2++;
  }
  len = (unsigned)len / (unsigned)stride;
#pragma unroll 1
  for (int i = 0; i < stride; i++) {
    if (stride2) rotate_seq(X + i * len, len, stride2, s, c);
    rotate_seq(X + i * len, len, 1, c, s);
  }
}

// ---- phase 1 -------------------------------------------------------------------------------
OB_DEV void pvq_phase1(const PvqShared *tab, PvqWS *ws, const ObScript *sc, int n_leaves, int spread, int Nframe, int lane) {
#pragma unroll 1
  for (int l = lane; l < n_leaves; l += 32) {
    const ObLeaf lf = sc->leaves[l];
    float *X = ws->X + lf.off;
    const int kind = lf.B_kind >> 5;
    if (kind == 0) {
      const int n = lf.n, K = lf.K, B = lf.B_kind & 31;
      int yy = pvq_index_to_pulses_f(tab, n, K, lf.index, X);
      float g = rsqrt_ref((float)yy) * lf.gain;
      unsigned mask = 1;
      if (B > 1) {  // extract_collapse_mask
        mask = 0;
        const int N0 = (unsigned)n / (unsigned)B;
        int j = 0;
#pragma unroll 1
        for (int blk = 0; blk < B; blk++) {
          int nz = 0;
#pragma unroll 1
          for (int e = 0; e < N0; e++, j++) nz |= X[j] != 0.f;
          mask |= (unsigned)nz << blk;
        }
      }
#pragma unroll 1
      for (int j = 0; j < n; j++) X[j] = X[j] * g;
      unrotate_seq(X, n, B, K, spread);
      ws->leaf_cm[l] = (unsigned short)mask;
    } else if (kind == 2) {  // N == 1 bands: a sign per channel (celt/bands.c:934-968)
      X[0] = (lf.index & 1) ? -1.f : 1.f;
      if (lf.K == 2) X[Nframe] = (lf.index & 2) ? -1.f : 1.f;
      ws->leaf_cm[l] = 1;
    }
  }
  wsync();
}

// ---- phase 2 -------------------------------------------------------------------------------
struct PartRet { unsigned cm; uint32_t seed; };

// A leaf without pulses (celt/bands.c:1131-1172): zero, LCG noise, or the folded lower band, then
// renormalised.  Warp cooperative; all arguments warp-uniform.  Returns collapse mask + new seed.
OB_DEVNI PartRet fold_leaf(PvqWS *ws, int off, int n, float gain, int has_lb, const float *lbv, int B, int fill, uint32_t seed,
                          int lane) {
  float *X = ws->X + off;
  const unsigned cm_mask = (unsigned)(1UL << B) - 1;
  PartRet r;
  r.cm = 0; r.seed = seed;
  fill &= cm_mask;
  wsync();
  if (!fill) {
#pragma unroll 1
    for (int j = lane; j < n; j += 32) X[j] = 0.f;
    wsync();
    return r;
  }
  if (!has_lb) {
#pragma unroll 1
    for (int j = 0; j < n; j++) {
      seed = 1664525u * seed + 1013904223u;
      if ((j & 31) == lane) X[j] = (float)((int32_t)seed >> 20);
    }
    r.cm = cm_mask;
  } else {
#pragma unroll 1
    for (int j = 0; j < n; j++) {
      seed = 1664525u * seed + 1013904223u;
      if ((j & 31) == lane) {
        float t = 1.0f / 256;
        t = (seed & 0x8000) ? t : -t;
        X[j] = lbv[j] + t;
      }
    }
    r.cm = fill;
  }
  r.seed = seed;
  renormalise(X, n, gain, lane);
  return r;
}

// One band part (resynthesis side of celt/bands.c:973-1176) WITHOUT walking its partition tree:
// lane l takes leaf l0+l, derives the leaf's `fill` from the part's root fill and the leaf's split
// path (bands.c:1010-1012: B==1 duplicates the bit, theta 0 / 16384 clear one half, the second
// child sees fill>>B), pulse-less leaves are produced one after the other in decode order (their
// LCG draws are order dependent), and the part's collapse mask is the OR of the leaves' masks
// shifted into place (bands.c:1121).
OB_DEV PartRet part_leaves(PvqWS *ws, const ObScript *sc, int l0, int l1, int part_off, int B_root, const float *lowband,
                           int fill_root, uint32_t seed, int lane) {
  unsigned cm_part = 0;
#pragma unroll 1
  for (int base = l0; base < l1; base += 32) {
    const int l = base + lane;
    const bool valid = l < l1;
    ObLeaf lf;
    lf.index = 0; lf.gain = 0.f; lf.off = 0; lf.n = 0; lf.K = 0; lf.path = 0; lf.B_kind = 0; lf.cmshift = 0;
    if (valid) lf = sc->leaves[l];
    const int kind = lf.B_kind >> 5, Bl = lf.B_kind & 31;
    // leaf fill from the split path
    int fill = fill_root, B0 = B_root;
    const int depth = (lf.path >> 12) & 7;
    unsigned path = lf.path & 0xFFFu;
#pragma unroll 1
    for (int d = 0; d < depth; d++, path >>= 3) {
      if (B0 == 1) fill = (fill & 1) | (fill << 1);
      const int B = (B0 + 1) >> 1;
      const int cls = path & 3;
      if (cls == 0) fill &= (1 << B) - 1;
      else if (cls == 1) fill &= ((1 << B) - 1) << B;
      if (path & 4) fill >>= B;
      B0 = B;
    }
    unsigned cm = (valid && kind != 1) ? ws->leaf_cm[l] : 0u;
    unsigned pending = ballot(valid && kind == 1);
#pragma unroll 1
    while (pending) {
      const int src = ffs32(pending) - 1;
      pending &= pending - 1;
      const int off = shfl((int)lf.off, src), n = shfl((int)lf.n, src), Bs = shfl(Bl, src), fs = shfl(fill, src);
      const float gain = shfl(lf.gain, src);
      const int has_lb = shfl((int)(lf.path >> 15), src) && lowband != nullptr;
      PartRet fr = fold_leaf(ws, off, n, gain, has_lb, has_lb ? lowband + (off - part_off) : nullptr, Bs, fs, seed, lane);
      seed = fr.seed;
      if (lane == src) cm = fr.cm;
    }
    cm_part |= warp_or(cm << lf.cmshift);
  }
  PartRet r;
  r.cm = cm_part; r.seed = seed;
  return r;
}

// resynthesis side of celt/bands.c:1248-1379
OB_DEVNI PartRet band_mono_vec(PvqWS *ws, const ObScript *sc, int l0, int l1, uint32_t seed, int part_off, int N, int B,
                              float *lowband, float *lowband_out, float *lowband_scratch, int fill, int tf_change, int lane) {
  float *X = ws->X + part_off;
  float *tmp = ws->tmp;
  int N0 = N, N_B = N, N_B0, B0 = B, time_divide = 0, recombine = 0;
  const int longBlocks = B0 == 1;
  PartRet ret;
  ret.cm = 1; ret.seed = seed;
  N_B = (unsigned)N_B / (unsigned)B;
  if (N == 1) {  // sign(s) already placed by phase 1
    wsync();
    if (lowband_out && lane == 0) lowband_out[0] = X[0];
    wsync();
    return ret;
  }
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
  ret = part_leaves(ws, sc, l0, l1, part_off, B, lowband, fill, seed, lane);
  unsigned cm = ret.cm;
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
    unsigned c4 = cm & 0xF;  // bit_deinterleave_table, celt/bands.c:1352
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
  ret.cm = cm;
  return ret;
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

// celt/bands.c:259-353, X in shared memory
OB_DEV void anti_collapse_vec(PvqWS *ws, const ObFrameSide *side, int lane, int LM, int C, int size, int start, int end,
                              const float *logE, const float *prev1logE, const float *prev2logE, uint32_t seed) {
  wsync();
#pragma unroll 1
  for (int i = start; i < end; i++) {
    int N0 = ob_ebands[i + 1] - ob_ebands[i];
    int depth = ((uint32_t)(1 + side->pulses[i]) / (uint32_t)N0) >> LM;
    float thresh = .5f * exp2_ref(-.125f * depth);
    float sqrt_1 = rsqrt_ref((float)(N0 << LM));
#pragma unroll 1
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
      float *X = ws->X + c * size + (ob_ebands[i] << LM);
      int renorm = 0;
      unsigned mask = ws->collapse[i * C + c];
#pragma unroll 1
      for (int k = 0; k < 1 << LM; k++) {
        if (!(mask & 1u << k)) {
#pragma unroll 1
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

// ------------------------------------------------------------------------------------------
// Kernel A2 body: one warp, one stream.
// ------------------------------------------------------------------------------------------
OB_DEV void dec_pvq_warp(const DecPvqParams &p, int stream, const PvqShared *tab, unsigned char *smem_warp) {
  const int lane = lane_id();
  PvqWS *ws = (PvqWS *)smem_warp;
  const ObFrameSide *side = p.side + stream;
  if (side->status <= 0) return;
  if (OB_PLC_TYPE(side->plc) != 0) {  // a lost frame
    // Noise PLC (celt_decoder.c:773-790): every band below `end` is LCG noise of unit norm, channel after channel; the synthesis
    // kernel then runs as for a decoded frame.  (Pitch PLC and the all-zero case have no spectrum: ob_celt_dec_plc.)
    if (OB_PLC_TYPE(side->plc) != OB_PLC_NOISE) return;
    const int LM = side->LM, N = OB_SHORT_N << LM, end = side->end_band;
    float *Xg = p.X + (size_t)stream * 2 * OB_MAXN;
    uint32_t seed = side->seed;
#pragma unroll 1
    for (int c = 0; c < side->C; c++) {
#pragma unroll 1
      for (int i = 0; i < end; i++) {
        float *X = Xg + c * N + (ob_ebands[i] << LM);
        const int n = (ob_ebands[i + 1] - ob_ebands[i]) << LM;
#pragma unroll 1
        for (int j = 0; j < n; j++) {  // every lane runs the generator, lane j mod 32 keeps the sample
          seed = 1664525u * seed + 1013904223u;
          if ((j & 31) == lane) X[j] = (float)((int32_t)seed >> 20);
        }
        renormalise(X, n, 1.0f, lane);
      }
    }
    return;
  }
  const ObScript *sc = p.script + stream;
  const int LM = side->LM, C = side->C, isTransient = side->transient, end = side->end_band, start = 0;
  const int N = OB_SHORT_N << LM, M = 1 << LM;
  const int spread = side->spread, intensity = side->intensity;
  const int B = isTransient ? M : 1;

  // stage the script's band records in shared memory (coalesced 4-byte copies)
  {
    const uint32_t *bs = (const uint32_t *)sc->bands;
    uint32_t *bd = (uint32_t *)ws->bands;
#pragma unroll 1
    for (int i = lane; i < (int)(OB_NBANDS * sizeof(ObBandRec) / 4); i += 32) bd[i] = bs[i];
  }
  pvq_phase1(tab, ws, sc, side->n_leaves, spread, N, lane);

  // ---- phase 2: bands in order (resynthesis side of celt/bands.c:1589-1923) ----
  uint32_t seed = side->seed;
  const int norm_offset = M * ob_ebands[start];
  float *norm = ws->norm, *norm2 = ws->norm + (M * ob_ebands[OB_NBANDS - 1] - norm_offset);
  for (int i = lane; i < 2 * OB_NBANDS; i += 32) ws->collapse[i] = 0;
  wsync();
#pragma unroll 1
  for (int i = start; i < end; i++) {
    const ObBandRec br = ws->bands[i];
    const int last = (i == end - 1);
    const int Nb = M * ob_ebands[i + 1] - M * ob_ebands[i];
    const int xoff = M * ob_ebands[i];
    const int tf_change = side->tf_res[i];
    const int effective_lowband = br.eff_lowband;
    const int flags = br.flags;
    float *lowband_scratch = last ? nullptr : ws->scr;
    unsigned x_cm, y_cm;
    if (effective_lowband != -1) {
      const int lowband_offset = br.lowband_offset;
      int fold_start = lowband_offset;
      while (M * ob_ebands[--fold_start] > effective_lowband + norm_offset) {}
      int fold_end = lowband_offset - 1;
      while (++fold_end < i && M * ob_ebands[fold_end] < effective_lowband + norm_offset + Nb) {}
      x_cm = y_cm = 0;
      int fold_i = fold_start;
      do {
        x_cm |= ws->collapse[fold_i * C + 0];
        y_cm |= ws->collapse[fold_i * C + C - 1];
      } while (++fold_i < fold_end);
    } else x_cm = y_cm = (1u << B) - 1;
    if (flags & 2) {  // dual stereo switches to intensity here: norm = (norm + norm2)/2
      wsync();
#pragma unroll 1
      for (int j = lane; j < M * ob_ebands[i] - norm_offset; j += 32) norm[j] = .5f * (norm[j] + norm2[j]);
      wsync();
    }
    float *lb1 = effective_lowband != -1 ? norm + effective_lowband : nullptr;
    float *lb2 = effective_lowband != -1 ? norm2 + effective_lowband : nullptr;
    float *lo1 = last ? nullptr : norm + M * ob_ebands[i] - norm_offset;
    float *lo2 = last ? nullptr : norm2 + M * ob_ebands[i] - norm_offset;
    const int l0 = br.leaf_begin, lm = br.leaf_mid, l1 = br.leaf_end;
    PartRet pr;
    if (flags & 1) {
      pr = band_mono_vec(ws, sc, l0, lm, seed, xoff, Nb, B, lb1, lo1, lowband_scratch, x_cm, tf_change, lane);
      x_cm = pr.cm; seed = pr.seed;
      pr = band_mono_vec(ws, sc, lm, l1, seed, N + xoff, Nb, B, lb2, lo2, lowband_scratch, y_cm, tf_change, lane);
      y_cm = pr.cm; seed = pr.seed;
    } else if (flags & 4) {
      // resynthesis side of quant_band_stereo, celt/bands.c:1387-1570
      float *X = ws->X + xoff, *Y = ws->X + N + xoff;
      int fill = x_cm | y_cm;
      unsigned cm;
      if (Nb == 1) {
        wsync();
        if (lo1 && lane == 0) lo1[0] = X[0];
        wsync();
        cm = 1;
      } else {
        const int orig_fill = fill;
        if (br.theta_class == 0) fill &= (1 << B) - 1;
        else if (br.theta_class == 1) fill &= ((1 << B) - 1) << B;
        const float mid = br.mid, sidef = br.side;
        if (flags & 16) {  // N == 2
          const int c = (flags & 32) != 0;
          const int sign = (flags & 64) ? -1 : 1;
          pr = band_mono_vec(ws, sc, l0, l1, seed, (c ? N : 0) + xoff, Nb, B, lb1, lo1, lowband_scratch, orig_fill, tf_change, lane);
          cm = pr.cm; seed = pr.seed;
          float *x2 = c ? Y : X, *y2 = c ? X : Y;
          wsync();
          if (lane == 0) {
            y2[0] = -sign * x2[1];
            y2[1] = sign * x2[0];
            X[0] = mid * X[0];
            X[1] = mid * X[1];
            Y[0] = sidef * Y[0];
            Y[1] = sidef * Y[1];
            float t = X[0];
            X[0] = t - Y[0];
            Y[0] = t + Y[0];
            t = X[1];
            X[1] = t - Y[1];
            Y[1] = t + Y[1];
          }
          wsync();
        } else {
          const int side_first = (flags & 128) != 0;
          cm = 0;
#pragma unroll 1
          for (int pass = 0; pass < 2; pass++) {
            const int do_mid = (pass == 0) != side_first;
            const int pl0 = pass == 0 ? l0 : lm, pl1 = pass == 0 ? lm : l1;
            if (do_mid) pr = band_mono_vec(ws, sc, pl0, pl1, seed, xoff, Nb, B, lb1, lo1, lowband_scratch, fill, tf_change, lane);
            else pr = band_mono_vec(ws, sc, pl0, pl1, seed, N + xoff, Nb, B, nullptr, nullptr, nullptr, fill >> B, tf_change, lane);
            cm |= pr.cm; seed = pr.seed;
          }
          merge_mid_side(X, Y, mid, Nb, lane);
        }
        if (flags & 8) {  // inv
          wsync();
#pragma unroll 1
          for (int j = lane; j < Nb; j += 32) Y[j] = -Y[j];
          wsync();
        }
      }
      x_cm = y_cm = cm;
    } else {
      pr = band_mono_vec(ws, sc, l0, l1, seed, xoff, Nb, B, lb1, lo1, lowband_scratch, x_cm | y_cm, tf_change, lane);
      x_cm = y_cm = pr.cm; seed = pr.seed;
    }
    wsync();
    if (lane == 0) {
      ws->collapse[i * C + 0] = (unsigned char)x_cm;
      ws->collapse[i * C + C - 1] = (unsigned char)y_cm;
    }
    wsync();
  }
  if (side->anti_collapse_on)
    anti_collapse_vec(ws, side, lane, LM, C, N, start, end, side->bandE, side->prevLogE[0], side->prevLogE[1], seed);
  wsync();
  // ---- one coalesced store of the frame's spectrum ----
  {
    float *Xg = p.X + (size_t)stream * 2 * OB_MAXN;
    const int bound = M * ob_ebands[end];
#pragma unroll 1
    for (int c = 0; c < C; c++)
#pragma unroll 1
      for (int j = lane; j < N; j += 32) Xg[c * N + j] = j < bound ? ws->X[c * N + j] : 0.f;
  }
  // (the energy-state update of celt/celt_decoder.c:1575-1600 is done by the entropy kernel)
}

}  // namespace ob
