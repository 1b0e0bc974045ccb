// celt_enc_rdo.cuh — the resynthesis-enabled band coder of the frame-encoder thread: quant_all_bands with encode = 1 and
// theta_rdo = 1 (stereo, not dual stereo, complexity >= 8; celt/bands.c:1589-1923).  Every partition is also reconstructed
// the way the decoder does it (alg_quant's resynth branch celt/vq.c:552-617, the noise / folding fill of quant_partition
// :1113-1160, quant_band's inverse Haar / Hadamard steps :1330-1370, stereo_merge :407-466), each stereo band below the
// intensity threshold is coded twice (theta rounded down / up) and the candidate with the larger weighted correlation to the
// input is kept.  Thread-private scalar code like the rest of celt_enc_frame.cuh (included from there).
#pragma once

namespace ob {

OB_TABLE unsigned char ob_bit_interleave[16] = {0, 1, 1, 1, 2, 3, 3, 3, 2, 3, 3, 3, 2, 3, 3, 3};   // celt/bands.c:1288
OB_TABLE unsigned char ob_bit_deinterleave[16] = {0x00, 0x03, 0x0C, 0x0F, 0x30, 0x33, 0x3C, 0x3F, 0xC0, 0xC3, 0xCC, 0xCF, 0xF0, 0xF3, 0xFC, 0xFF};   // :1352

OB_DEVNI void rdo_exp_rotation_inv(float *X, int len, int stride, int K, int spread) {  // celt/vq.c:104-147, dir = -1
  if (2 * K >= len || spread == 0) return;
  const int factor = spread == 1 ? 15 : spread == 2 ? 10 : 5;
  const float gain = ((float)(1.f * len)) / ((float)(len + factor * K));
  const float theta = .5f * (gain * gain);
  const float c = cos_norm_ref(theta), s = cos_norm_ref(1.0f - theta);
  int stride2 = 0;
  if (len >= 8 * stride) {
    stride2 = 1;
    while ((stride2 * stride2 + stride2) * stride + (stride >> 2) < len) stride2++;
  }
  len = (unsigned)len / (unsigned)stride;
#pragma unroll 1
  for (int i = 0; i < stride; i++) {
    if (stride2) enc_rotate_pass(X + i * len, len, stride2, s, c);
    enc_rotate_pass(X + i * len, len, 1, c, s);
  }
}
OB_DEVNI void rdo_interleave(float *__restrict__ X, float *__restrict__ tmp, int N0, int stride, int hadamard) {  // celt/bands.c:599-621
  const int N = N0 * stride;
#pragma unroll 1
  for (int i = 0; i < stride; i++) {
    const int row = hadamard ? ob_enc_hadamard_order[stride - 2 + i] : i;
    const float *__restrict__ q = X + row * N0;
    float *__restrict__ d = tmp + i;
#pragma unroll 4
    for (int j = 0; j < N0; j++) d[j * stride] = q[j];
  }
#pragma unroll 4
  for (int j = 0; j < N; j++) X[j] = tmp[j];
}
OB_DEVNI void rdo_renormalise(float *X, int N, float gain) {  // celt/vq.c:695-722
  const float E = 1e-15f + enc_dot(X, X, N);
  const float g = rsqrt_ref(E) * gain;
#pragma unroll 4
  for (int i = 0; i < N; i++) X[i] = g * X[i];
}
OB_DEVNI void rdo_stereo_merge(float *X, float *Y, float mid, int N) {  // celt/bands.c:407-466
  float xp = enc_dot(Y, X, N);
  const float side = enc_dot(Y, Y, N);
  xp = mid * xp;
  const float El = (mid * mid) + side - 2 * xp, Er = (mid * mid) + side + 2 * xp;
  if (Er < 6e-4f || El < 6e-4f) {
#pragma unroll 1
    for (int j = 0; j < N; j++) Y[j] = X[j];
    return;
  }
  const float lgain = rsqrt_ref(El), rgain = rsqrt_ref(Er);
#pragma unroll 2
  for (int j = 0; j < N; j++) {
    const float l = mid * X[j], r = Y[j];
    X[j] = lgain * (l - r);
    Y[j] = rgain * (l + r);
  }
}
// alg_quant with resynth = 1 (celt/vq.c:552-617): X <- the decoded partition; returns the collapse mask
OB_DEVNI unsigned rdo_pvq(float *X, float *y2, int N, int K, int spread, int B, RangeEnc &ec, float gain) {
  uint32_t signs[6];
  enc_exp_rotation(X, N, B, K, spread);
  const float yy = enc_pvq_search(X, y2, N, K, signs);
  unsigned cm = 1;
  if (B > 1) {   // extract_collapse_mask (celt/vq.c:183-203)
    const int N0 = (unsigned)N / (unsigned)B;
    cm = 0;
#pragma unroll 1
    for (int i = 0; i < B; i++) {
      int any = 0;
#pragma unroll 4
      for (int j = 0; j < N0; j++) any |= y2[i * N0 + j] != 0.f;
      cm |= (unsigned)any << i;
    }
  }
  re_uint(ec, enc_icwrs(y2, signs, N), pvq_u(N, K) + pvq_u(N, K + 1));
  const float g = rsqrt_ref(yy) * gain;   // normalise_residual (celt/vq.c:150-181)
#pragma unroll 4
  for (int j = 0; j < N; j++) {
    const int k = (int)(y2[j] * .5f);
    const int iy = ((signs[j >> 5] >> (j & 31)) & 1u) ? -k : k;
    X[j] = (float)iy * g;
  }
  rdo_exp_rotation_inv(X, N, B, K, spread);
  return cm;
}
// quant_partition with encode = resynth = 1 (celt/bands.c:973-1176); the recursion runs on an explicit stack (depth <= LM + 2)
struct RdoPart {
  short off, N, lb, itheta;   // lb: offset of this partition's folding source in `lowband`, -1 = none
  int b, second_bits, rebalance, fill;
  signed char B, Bc, LM, phase, second_x;   // B: this partition's block count; Bc: its children's
  float gain, mid, side;
  unsigned cm;
};
OB_DEVNI unsigned rdo_partition(EncBand &ctx, float *X, float *y2, const float *lowband, int N0, int b0, int B0_, int LM0, float gain0, int fill0) {
  RdoPart st[6];
  int sp = 0;
  unsigned ret = 0;
  st[0].off = 0; st[0].N = (short)N0; st[0].lb = lowband ? 0 : -1; st[0].b = b0; st[0].B = (signed char)B0_; st[0].LM = (signed char)LM0; st[0].phase = 0;
  st[0].gain = gain0; st[0].fill = fill0;
  // The walk alternates two steps for all lanes together: (A) advance the tree — splits, theta, pulse-less leaves — until a leaf with pulses is
  // reached or the tree is done; (B) every lane that has such a leaf quantises it (rotation, pulse search, index, resynthesis: most of the
  // band coder's instructions).  The vote between the steps keeps the lanes of a warp (= different streams) in the same pass of (B) instead
  // of each running it alone whenever its own walk happens to get there.
  const unsigned lanes = converged_lanes();
  int pend_K = 0;   // > 0: the leaf at st[sp] waits for step (B)
  for (;;) {
  while (sp >= 0 && !pend_K) {
    RdoPart &f = st[sp];
    const int i = ctx.band;
    if (f.phase == 0) {
      const unsigned char *cache = pulse_cache_row(i, f.LM);
      if (f.LM != -1 && f.b > cache[cache[0]] + 12 && f.N > 2) {
        const int N = f.N >> 1, LM = f.LM - 1, Bp = f.B, B = (f.B + 1) >> 1;
        int b = f.b, fill = f.fill;
        if (Bp == 1) fill = (fill & 1) | (fill << 1);
        const EncTheta t = enc_theta(ctx, X + f.off, X + f.off + N, N, b, B, Bp, LM, 0);
        if (t.itheta == 0) fill &= (1 << B) - 1;
        else if (t.itheta == 16384) fill &= ((1 << B) - 1) << B;
        f.mid = (1.f / 32768) * t.imid;
        f.side = (1.f / 32768) * t.iside;
        int delta = t.delta;
        if (Bp > 1 && (t.itheta & 0x3fff)) {
          if (t.itheta > 8192) delta -= delta >> (4 - LM);
          else delta = OB_IMIN(0, delta + (N << OB_BITRES >> (5 - LM)));
        }
        const int mbits = OB_IMAX(0, OB_IMIN(b, (b - delta) / 2)), sbits = b - mbits;
        ctx.remaining_bits -= t.qalloc;
        f.rebalance = ctx.remaining_bits;
        f.itheta = (short)t.itheta;
        f.fill = fill;
        f.Bc = (signed char)B;
        f.phase = 1;
        RdoPart &ch = st[sp + 1];
        ch.N = (short)N; ch.B = (signed char)B; ch.LM = (signed char)LM; ch.phase = 0;
        if (mbits >= sbits) {   // mid first
          ch.off = f.off; ch.b = mbits; ch.gain = f.gain * f.mid; ch.fill = fill; ch.lb = f.lb;
          f.second_x = 0; f.second_bits = sbits; f.b = mbits;
        } else {                // side first
          ch.off = (short)(f.off + N); ch.b = sbits; ch.gain = f.gain * f.side; ch.fill = fill >> B; ch.lb = (short)(f.lb >= 0 ? f.lb + N : -1);
          f.second_x = 1; f.second_bits = mbits; f.b = sbits;
        }
        sp++;
      } else {
        int q = bits_to_pseudo(cache, f.b);
        int curr_bits = q == 0 ? 0 : cache[q] + 1;
        ctx.remaining_bits -= curr_bits;
        while (ctx.remaining_bits < 0 && q > 0) {
          ctx.remaining_bits += curr_bits;
          q--;
          curr_bits = q == 0 ? 0 : cache[q] + 1;
          ctx.remaining_bits -= curr_bits;
        }
        float *Xp = X + f.off;
        unsigned cm = 0;
        if (q != 0) { pend_K = pseudo_to_pulses(q); continue; }   // step (B) below
        else {   // no pulse: fill the band anyway (celt/bands.c:1113-1160)
          const unsigned cm_mask = (1u << f.B) - 1;
          const int fill = f.fill & (int)cm_mask;
          if (!fill) {
#pragma unroll 1
            for (int j = 0; j < f.N; j++) Xp[j] = 0.f;
          } else {
            if (f.lb < 0) {
#pragma unroll 1
              for (int j = 0; j < f.N; j++) {
                ctx.seed = 1664525u * ctx.seed + 1013904223u;
                Xp[j] = (float)((int32_t)ctx.seed >> 20);
              }
              cm = cm_mask;
            } else {
#pragma unroll 1
              for (int j = 0; j < f.N; j++) {
                ctx.seed = 1664525u * ctx.seed + 1013904223u;
                const float tmp = (ctx.seed & 0x8000) ? 1.0f / 256 : -(1.0f / 256);
                Xp[j] = lowband[f.lb + j] + tmp;
              }
              cm = (unsigned)fill;
            }
            rdo_renormalise(Xp, f.N, f.gain);
          }
        }
        ret = cm;
        sp--;
      }
    } else if (f.phase == 1) {   // first half done: rebalance, then the other half
      const int N = f.N >> 1;
      f.cm = f.second_x ? ret << (f.B >> 1) : ret;   // (second_x: the first half was the side)
      int second = f.second_bits;
      const int rebalance = f.b - (f.rebalance - ctx.remaining_bits);
      if (rebalance > 3 << OB_BITRES && f.itheta != (f.second_x ? 16384 : 0)) second += rebalance - (3 << OB_BITRES);
      f.phase = 2;
      RdoPart &ch = st[sp + 1];
      ch.N = (short)N; ch.B = f.Bc; ch.LM = (signed char)(f.LM - 1); ch.phase = 0; ch.b = second;
      if (f.second_x) { ch.off = f.off; ch.gain = f.gain * f.mid; ch.fill = f.fill; ch.lb = f.lb; }
      else { ch.off = (short)(f.off + N); ch.gain = f.gain * f.side; ch.fill = f.fill >> f.Bc; ch.lb = (short)(f.lb >= 0 ? f.lb + N : -1); }
      sp++;
    } else {
      f.cm |= f.second_x ? ret : ret << (f.B >> 1);
      ret = f.cm;
      sp--;
    }
  }
    if (!ballot_among(lanes, pend_K != 0)) break;   // every lane's tree is done
    if (pend_K) {
      RdoPart &f = st[sp];
      ret = rdo_pvq(X + f.off, y2, f.N, pend_K, ctx.spread, f.B, *ctx.ec, f.gain);
      pend_K = 0;
      sp--;
    }
  }
  return ret;
}
OB_DEV unsigned rdo_band_n1(EncBand &ctx, float *X, float *Y, float *lowband_out) {  // celt/bands.c:934-968, resynth = 1
  float *x = X;
#pragma unroll 1
  for (int c = 0; c < (Y ? 2 : 1); c++) {
    int sign = 0;
    if (ctx.remaining_bits >= 1 << OB_BITRES) {
      sign = x[0] < 0;
      re_bits(*ctx.ec, sign, 1);
      ctx.remaining_bits -= 1 << OB_BITRES;
    }
    x[0] = sign ? -1.f : 1.f;
    x = Y;
  }
  if (lowband_out) lowband_out[0] = X[0];
  return 1;
}
// quant_band with encode = resynth = 1 (celt/bands.c:1248-1379).  tmp: 176 floats (Hadamard scratch / pulse scratch)
OB_DEVNI unsigned rdo_band_mono(EncBand &ctx, float *X, float *tmp, int N, int b, int B, float *lowband, int LM, float *lowband_out, float gain,
                                float *lowband_scratch, int fill) {
  const int N0 = N, longBlocks = B == 1;
  int N_B = (unsigned)N / (unsigned)B, recombine = 0, time_divide = 0, tf_change = ctx.tf_change;
  if (N == 1) return rdo_band_n1(ctx, X, nullptr, lowband_out);
  if (tf_change > 0) recombine = tf_change;
  if (lowband_scratch && lowband && (recombine || ((N_B & 1) == 0 && tf_change < 0) || B > 1)) {
#pragma unroll 1
    for (int j = 0; j < N; j++) lowband_scratch[j] = lowband[j];
    lowband = lowband_scratch;
  }
#pragma unroll 1
  for (int k = 0; k < recombine; k++) {
    enc_haar(X, N >> k, 1 << k);
    if (lowband) enc_haar(lowband, N >> k, 1 << k);
    fill = ob_bit_interleave[fill & 0xF] | ob_bit_interleave[fill >> 4] << 2;
  }
  B >>= recombine;
  N_B <<= recombine;
  while ((N_B & 1) == 0 && tf_change < 0) {
    enc_haar(X, N_B, B);
    if (lowband) enc_haar(lowband, N_B, B);
    fill |= fill << B;
    B <<= 1;
    N_B >>= 1;
    time_divide++;
    tf_change++;
  }
  const int B0 = B, N_B0 = N_B;
  if (B0 > 1) {
    enc_deinterleave(X, tmp, N_B >> recombine, B0 << recombine, longBlocks);
    if (lowband) enc_deinterleave(lowband, tmp, N_B >> recombine, B0 << recombine, longBlocks);
  }
  unsigned cm = rdo_partition(ctx, X, tmp, lowband, N, b, B, LM, gain, fill);
  if (B0 > 1) rdo_interleave(X, tmp, N_B >> recombine, B0 << recombine, longBlocks);
  N_B = N_B0;
  B = B0;
#pragma unroll 1
  for (int k = 0; k < time_divide; k++) {
    B >>= 1;
    N_B <<= 1;
    cm |= cm >> B;
    enc_haar(X, N_B, B);
  }
#pragma unroll 1
  for (int k = 0; k < recombine; k++) {
    cm = ob_bit_deinterleave[cm];
    enc_haar(X, N0 >> k, 1 << k);
  }
  B <<= recombine;
  if (lowband_out) {
    const float n = (float)sqrt((double)(float)N0);
#pragma unroll 1
    for (int j = 0; j < N0; j++) lowband_out[j] = n * X[j];
  }
  cm &= (1u << B) - 1;
  return cm;
}
// quant_band_stereo with encode = resynth = 1 (celt/bands.c:1387-1573)
OB_DEVNI unsigned rdo_band_stereo(EncBand &ctx, float *X, float *Y, float *tmp, int N, int b, int B, float *lowband, int LM, float *lowband_out,
                                  float *lowband_scratch, int fill) {
  if (N == 1) return rdo_band_n1(ctx, X, Y, lowband_out);
  const int orig_fill = fill;
  const float eL = ctx.bandE[ctx.band], eR = ctx.bandE[OB_NBANDS + ctx.band];
  if (eL < 1e-10f || eR < 1e-10f) {
    if (eL > eR) { for (int j = 0; j < N; j++) Y[j] = X[j]; }
    else { for (int j = 0; j < N; j++) X[j] = Y[j]; }
  }
  const EncTheta t = enc_theta(ctx, X, Y, N, b, B, B, LM, 1);
  if (t.itheta == 0) fill &= (1 << B) - 1;
  else if (t.itheta == 16384) fill &= ((1 << B) - 1) << B;
  const float mid = (1.f / 32768) * t.imid, side = (1.f / 32768) * t.iside;
  unsigned cm;
  // The three shapes of the band (N == 2; mid first; side first) go through the SAME two calls of rdo_band_mono: the lanes of a warp
  // code different streams, and threads that reach a function from different call sites do not reconverge inside it.
  int mbits, sbits, sign = 0, first_is_mid = 1, do_second = N != 2;
  float *x2 = X, *y2 = Y;
  if (N == 2) {
    mbits = b; sbits = 0;
    if (t.itheta != 0 && t.itheta != 16384) sbits = 1 << OB_BITRES;
    mbits -= sbits;
    const int c = t.itheta > 8192;
    ctx.remaining_bits -= t.qalloc + sbits;
    x2 = c ? Y : X; y2 = c ? X : Y;
    if (sbits) {
      sign = x2[0] * y2[1] - x2[1] * y2[0] < 0;
      re_bits(*ctx.ec, sign, 1);
    }
    sign = 1 - 2 * sign;
  } else {
    mbits = OB_IMAX(0, OB_IMIN(b, (b - t.delta) / 2)); sbits = b - mbits;
    ctx.remaining_bits -= t.qalloc;
    first_is_mid = mbits >= sbits;
  }
  int32_t rebalance = ctx.remaining_bits;
  {   // first part: the mid (with the folding source and the lowband output) or the side
    float *xf = N == 2 ? x2 : first_is_mid ? X : Y;
    const int bf = (N == 2 || first_is_mid) ? mbits : sbits, midlike = N == 2 || first_is_mid;
    cm = rdo_band_mono(ctx, xf, tmp, N, bf, B, midlike ? lowband : nullptr, LM, midlike ? lowband_out : nullptr, midlike ? 1.0f : side,
                       midlike ? lowband_scratch : nullptr, N == 2 ? orig_fill : midlike ? fill : fill >> B);
  }
  if (do_second) {
    if (first_is_mid) {
      rebalance = mbits - (rebalance - ctx.remaining_bits);
      if (rebalance > 3 << OB_BITRES && t.itheta != 0) sbits += rebalance - (3 << OB_BITRES);
    } else {
      rebalance = sbits - (rebalance - ctx.remaining_bits);
      if (rebalance > 3 << OB_BITRES && t.itheta != 16384) mbits += rebalance - (3 << OB_BITRES);
    }
    float *xs = first_is_mid ? Y : X;
    const int midlike = !first_is_mid;
    cm |= rdo_band_mono(ctx, xs, tmp, N, midlike ? mbits : sbits, B, midlike ? lowband : nullptr, LM, midlike ? lowband_out : nullptr, midlike ? 1.0f : side,
                        midlike ? lowband_scratch : nullptr, midlike ? fill : fill >> B);
  }
  if (N == 2) {
    y2[0] = -sign * x2[1];
    y2[1] = sign * x2[0];
    X[0] = mid * X[0];
    X[1] = mid * X[1];
    Y[0] = side * Y[0];
    Y[1] = side * Y[1];
    float t0 = X[0];
    X[0] = t0 - Y[0];
    Y[0] = t0 + Y[0];
    t0 = X[1];
    X[1] = t0 - Y[1];
    Y[1] = t0 + Y[1];
  }
  if (N != 2) rdo_stereo_merge(X, Y, mid, N);
  if (t.inv)
#pragma unroll 1
    for (int j = 0; j < N; j++) Y[j] = -Y[j];
  return cm;
}

// The band loop of quant_all_bands for the theta-RDO configuration (stereo, no dual stereo, start = 0, end = 21).
// Xg: this stream's normalised bands [2][N] in HBM (read only: they double as X_save / Y_save);  bx/by/tmp: 176-float work
// arrays.  Returns 0 or a negative status.
struct RdoScratch {
  float norm[8 * 78];        // folding history of the mid channel (celt/bands.c:1638)
  float lowband_scratch[176];
  float norm_save2[176];
  unsigned char bytes_save[320];
  unsigned char collapse_masks[2 * OB_NBANDS];
};
// SYNC: the CTA's warps meet at a barrier before every band, so that the warps of an SM walk the (large) band-coder code together and share
// its instruction-cache lines; every thread then runs all OB_NBANDS iterations (the barrier count must match), idle past its end band
template <bool SYNC>
OB_DEVNI int rdo_all_bands(RangeEnc &enc, const float *Xg, const float *bandE, const int *pulses, const int *tf_res, int shortBlocks, int spread,
                           int intensity, int32_t total_bits, int32_t balance, int LM, int codedBands, uint32_t seed, int disable_inv, float *bx,
                           float *by, float *tmp, RdoScratch &w, int end) {
  const int M = 1 << LM, N_ = M * OB_SHORT_N, B = shortBlocks ? M : 1, start = 0, C = 2;
  int lowband_offset = 0, update_lowband = 1, status = 0;
  EncBand ctx;
  ctx.ec = &enc; ctx.bandE = bandE; ctx.intensity = intensity; ctx.spread = spread; ctx.disable_inv = disable_inv; ctx.seed = seed;
  ctx.theta_round = 0;
  ctx.avoid_split_noise = B > 1;
#pragma unroll 1
  for (int j = 0; j < 8 * 78; j++) w.norm[j] = 0.f;
#pragma unroll 1
  for (int i = start; i < (SYNC ? OB_NBANDS : end); i++) {
    if (SYNC) {
      bsync();
      if (i >= end) continue;
    }
    const int off = M * ob_ebands[i], N = M * ob_ebands[i + 1] - off, last = i == end - 1;
    ctx.band = i;
    const int32_t tell = re_tell_frac(enc);
    if (i != start) balance -= tell;
    const int32_t remaining_bits = total_bits - tell - 1;
    ctx.remaining_bits = remaining_bits;
    int b = 0;
    if (i <= codedBands - 1) {
      const int32_t curr_balance = balance / OB_IMIN(3, codedBands - i);
      b = OB_IMAX(0, OB_IMIN(16383, OB_IMIN(remaining_bits + 1, pulses[i] + curr_balance)));
    }
    if ((off - N >= M * ob_ebands[start] || i == start + 1) && (update_lowband || lowband_offset == 0)) lowband_offset = i;
    const int tf_change = tf_res[i];
    ctx.tf_change = tf_change;
    int effective_lowband = -1;
    unsigned x_cm, y_cm;
    if (lowband_offset != 0 && (spread != 3 || B > 1 || tf_change < 0)) {   // 3 = SPREAD_AGGRESSIVE
      effective_lowband = OB_IMAX(0, M * ob_ebands[lowband_offset] - N);
      int fold_start = lowband_offset;
      while (M * ob_ebands[--fold_start] > effective_lowband);
      int fold_end = lowband_offset - 1;
      while (++fold_end < i && M * ob_ebands[fold_end] < effective_lowband + N);
      x_cm = y_cm = 0;
      int fold_i = fold_start;
      do {
        x_cm |= w.collapse_masks[fold_i * C + 0];
        y_cm |= w.collapse_masks[fold_i * C + C - 1];
      } while (++fold_i < fold_end);
    } else x_cm = y_cm = (1u << B) - 1;
    float *lowband = effective_lowband != -1 ? w.norm + effective_lowband : nullptr;
    float *lowband_out = last ? nullptr : w.norm + off;
#pragma unroll 1
    for (int j = 0; j < N; j++) { bx[j] = Xg[off + j]; by[j] = Xg[N_ + off + j]; }
    {
      // Below the intensity band the band is coded twice (theta rounded down, then up) and the better trial kept; above it once.  One
      // call site in a trial loop, so that the streams of a warp stay together whichever case each of them is in.
      const int rdo = i < intensity;
      const unsigned cm = x_cm | y_cm;
      const float minE = enc_fmin(bandE[i], bandE[i + OB_NBANDS]);   // compute_channel_weights (celt/bands.c:362-377)
      const float w0 = bandE[i] + minE / 3, w1 = bandE[i + OB_NBANDS] + minE / 3;
      const RangeEnc ec_save = enc;
      const EncBand ctx_save = ctx;
      RangeEnc ec_save2 = enc;
      EncBand ctx_save2 = ctx;
      float dist0 = 0.f;
      unsigned cm2 = 0;
      uint32_t nfront = 0, nend = 0;
#pragma unroll 1
      for (int trial = 0; trial < (rdo ? 2 : 1); trial++) {
        if (trial) {
          cm2 = x_cm;
          ec_save2 = enc;
          ctx_save2 = ctx;
          if (!last)
#pragma unroll 1
            for (int j = 0; j < N; j++) w.norm_save2[j] = lowband_out[j];
          // only the bytes the first trial wrote can differ between the trials: save those instead of the whole tail of the packet
          nfront = ec_save2.offs - ec_save.offs; nend = ec_save2.end_offs - ec_save.end_offs;
          if (nfront + nend > sizeof(w.bytes_save)) status = -3;
#pragma unroll 1
          for (uint32_t k = 0; k < nfront && k < sizeof(w.bytes_save); k++) w.bytes_save[k] = enc.buf[ec_save.offs + k];
#pragma unroll 1
          for (uint32_t k = 0; k < nend && nfront + k < sizeof(w.bytes_save); k++) w.bytes_save[nfront + k] = enc.buf[enc.storage - ec_save2.end_offs + k];
          enc = ec_save;
          ctx = ctx_save;
#pragma unroll 1
          for (int j = 0; j < N; j++) { bx[j] = Xg[off + j]; by[j] = Xg[N_ + off + j]; }
        }
        ctx.theta_round = rdo ? (trial ? 1 : -1) : 0;
        x_cm = rdo_band_stereo(ctx, bx, by, tmp, N, b, B, lowband, LM, lowband_out, w.lowband_scratch, (int)cm);
        if (rdo) {
          const float dist = w0 * enc_dot(Xg + off, bx, N) + w1 * enc_dot(Xg + N_ + off, by, N);
          if (!trial) dist0 = dist;
          else if (dist0 >= dist) {
            x_cm = cm2;
            enc = ec_save2;
            ctx = ctx_save2;
            if (!last)
#pragma unroll 1
              for (int j = 0; j < N; j++) lowband_out[j] = w.norm_save2[j];
#pragma unroll 1
            for (uint32_t k = 0; k < nfront && k < sizeof(w.bytes_save); k++) enc.buf[ec_save.offs + k] = w.bytes_save[k];
#pragma unroll 1
            for (uint32_t k = 0; k < nend && nfront + k < sizeof(w.bytes_save); k++) enc.buf[enc.storage - ec_save2.end_offs + k] = w.bytes_save[nfront + k];
          }
        }
      }
    }
    y_cm = x_cm;
    w.collapse_masks[i * C + 0] = (unsigned char)x_cm;
    w.collapse_masks[i * C + C - 1] = (unsigned char)y_cm;
    balance += pulses[i] + tell;
    update_lowband = b > (N << OB_BITRES);
    ctx.avoid_split_noise = 0;
  }
  return status;
}

}  // namespace ob
