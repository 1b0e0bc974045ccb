// celt_enc_frame.cuh — the coding half of celt_encode_with_ec (celt/celt_encoder.c:1726-2834), one THREAD per stream.
//
// The front half (pre-emphasis, MDCT, band energies, normalisation, tone detector) is ob_celt_enc_analysis; this kernel
// takes its outputs and produces the packet: silence / post-filter / transient flags, dynalloc_analysis, coarse energy
// (quant_coarse_energy), tf_encode, spread, dynamic allocation, stereo_analysis + intensity hysteresis,
// alloc_trim_analysis, clt_compute_allocation (encode side), fine energy, quant_all_bands (encode side: theta from the
// signal, exp_rotation + op_pvq_search + icwrs per leaf, no resynthesis), anti-collapse flag, quant_energy_finalise,
// ec_enc_done.  Everything after the MDCT is a chain of data-dependent scalar decisions on one range-coder state, so the
// parallel axis is the batch (like the decoder's entropy kernel); the vector work inside a band (rotation, PVQ search,
// Hadamard reordering) runs on thread-private copies of the band in local memory, which the hardware interleaves per lane.
// Round-1 scope = what the reference does at complexity 0 (see DESIGN.md §7b for the decision code still to come); the
// library is compiled with -fmad=false, so the float expressions below evaluate exactly as written.
#pragma once
#include "celt_dec_entropy.cuh"   // integer helpers shared with the decoder: bitexact_cos, theta_levels, pulse cache, pvq_u
#include "celt_ops.cuh"

namespace ob {

struct EncFrameParams {
  ObEncState *state;            // [n]
  const ObEncFrameInfo *info;   // [n]
  float *X;                     // [n][C*N] normalised MDCT (scratch: the band transforms work in place)
  const float *bandE;           // [n][C*21]
  const float *bandLogE;        // [n][C*21]
  const float *bandLogE2;       // [n][C*21] valid where info.patched (else the same as bandLogE)
  unsigned char *packets;       // [n][pkt_stride]
  int32_t *nbytes;              // [n] bytes produced, or a negative error
  const int32_t *perm = nullptr;   // frame-type binning: thread t codes stream perm[t] (null: stream t)
  int first_stream = 0;   // this launch covers streams [first_stream, n_streams)
  int end = OB_NBANDS;    // CELT_SET_END_BAND (celt/celt_encoder.c:3046-3053): 13 NB, 17 WB, 19 SWB, 21 FB
  int CC = 0;             // input channels when they differ from the coded channels C (CELT_SET_CHANNELS); 0 = C
  int n_streams, C, LM, pkt_stride, max_bytes, complexity, bitrate, lsb_depth, force_intra, disable_inv, disable_pf, vbr, constrained_vbr,
      loss_rate;
  // Opus layer: per-stream coded channels / end band / AnalysisInfo (see EncAnalysisParams::ctl); toc_bytes is added to nbytes (the
  // packet pointer then already skips the TOC byte)
  const ObOpusFrameCtl *ctl = nullptr;
  int toc_bytes = 0;
};

// ---- range encoder (celt/entenc.c:60-305), state in registers, bytes straight to the packet in HBM ----
struct RangeEnc {
  unsigned char *buf;
  uint32_t storage, end_offs, end_window, offs, rng, val, ext;
  int nend_bits, nbits_total, rem, error;
};
OB_DEV int re_tell(const RangeEnc &e) { return e.nbits_total - ilog32(e.rng); }
OB_DEV uint32_t re_tell_frac(const RangeEnc &e) { return rd_tell_frac(e.rng, e.nbits_total); }
OB_DEV void re_write_byte(RangeEnc &e, unsigned v) {
  if (e.offs + e.end_offs >= e.storage) { e.error = -1; return; }
  e.buf[e.offs++] = (unsigned char)v;
}
OB_DEV void re_write_byte_at_end(RangeEnc &e, unsigned v) {
  if (e.offs + e.end_offs >= e.storage) { e.error = -1; return; }
  e.buf[e.storage - ++e.end_offs] = (unsigned char)v;
}
OB_DEVNI void re_carry_out(RangeEnc &e, int c) {
  if (c != 0xFF) {
    const int carry = c >> 8;
    if (e.rem >= 0) re_write_byte(e, e.rem + carry);
    if (e.ext > 0) {
      const unsigned sym = (0xFF + carry) & 0xFF;
      do re_write_byte(e, sym); while (--e.ext > 0);
    }
    e.rem = c & 0xFF;
  } else e.ext++;
}
OB_DEVNI void re_normalize(RangeEnc &e) {
  while (e.rng <= (1u << 23)) {
    re_carry_out(e, (int)(e.val >> 23));
    e.val = (e.val << 8) & ((1u << 31) - 1);
    e.rng <<= 8;
    e.nbits_total += 8;
  }
}
OB_DEV void re_init(RangeEnc &e, unsigned char *buf, uint32_t size) {
  e.buf = buf; e.end_offs = 0; e.end_window = 0; e.nend_bits = 0; e.nbits_total = 33; e.offs = 0;
  e.rng = 1u << 31; e.rem = -1; e.val = 0; e.ext = 0; e.storage = size; e.error = 0;
}
OB_DEVNI void re_encode(RangeEnc &e, unsigned fl, unsigned fh, unsigned ft) {
  const uint32_t r = e.rng / ft;
  if (fl > 0) { e.val += e.rng - r * (ft - fl); e.rng = r * (fh - fl); }
  else e.rng -= r * (ft - fh);
  re_normalize(e);
}
OB_DEVNI void re_encode_bin(RangeEnc &e, unsigned fl, unsigned fh, unsigned bits) {
  const uint32_t r = e.rng >> bits;
  if (fl > 0) { e.val += e.rng - r * ((1u << bits) - fl); e.rng = r * (fh - fl); }
  else e.rng -= r * ((1u << bits) - fh);
  re_normalize(e);
}
OB_DEVNI void re_bit_logp(RangeEnc &e, int val, unsigned logp) {
  uint32_t r = e.rng;
  const uint32_t l = e.val, s = r >> logp;
  r -= s;
  if (val) e.val = l + r;
  e.rng = val ? s : r;
  re_normalize(e);
}
OB_DEVNI void re_icdf(RangeEnc &e, int s, const unsigned char *icdf, unsigned ftb) {
  const uint32_t r = e.rng >> ftb;
  if (s > 0) { e.val += e.rng - r * icdf[s - 1]; e.rng = r * (icdf[s - 1] - icdf[s]); }
  else e.rng -= r * icdf[s];
  re_normalize(e);
}
OB_DEVNI void re_bits(RangeEnc &e, uint32_t fl, unsigned bits) {
  uint32_t window = e.end_window;
  int used = e.nend_bits;
  if (used + bits > 32) {
    do {
      re_write_byte_at_end(e, window & 0xFF);
      window >>= 8;
      used -= 8;
    } while (used >= 8);
  }
  window |= fl << used;
  used += bits;
  e.end_window = window; e.nend_bits = used; e.nbits_total += bits;
}
OB_DEVNI void re_uint(RangeEnc &e, uint32_t fl, uint32_t ft) {
  ft--;
  int ftb = ilog32(ft);
  if (ftb > 8) {
    ftb -= 8;
    const unsigned ft2 = (ft >> ftb) + 1, fl2 = (unsigned)(fl >> ftb);
    re_encode(e, fl2, fl2 + 1, ft2);
    re_bits(e, fl & (((uint32_t)1 << ftb) - 1U), ftb);
  } else re_encode(e, fl, fl + 1, ft + 1);
}
OB_DEVNI void re_done(RangeEnc &e) {
  int l = 32 - ilog32(e.rng);
  uint32_t msk = ((1u << 31) - 1) >> l;
  uint32_t end = (e.val + msk) & ~msk;
  if ((end | msk) >= e.val + e.rng) {
    l++;
    msk >>= 1;
    end = (e.val + msk) & ~msk;
  }
  while (l > 0) {
    re_carry_out(e, (int)(end >> 23));
    end = (end << 8) & ((1u << 31) - 1);
    l -= 8;
  }
  if (e.rem >= 0 || e.ext > 0) re_carry_out(e, 0);
  uint32_t window = e.end_window;
  int used = e.nend_bits;
  while (used >= 8) {
    re_write_byte_at_end(e, window & 0xFF);
    window >>= 8;
    used -= 8;
  }
  if (!e.error) {
#pragma unroll 1
    for (uint32_t i = e.offs; i < e.storage - e.end_offs; i++) e.buf[i] = 0;
    if (used > 0) {
      if (e.end_offs >= e.storage) e.error = -1;
      else {
        l = -l;
        if (e.offs + e.end_offs >= e.storage && l < used) {
          window &= (1 << l) - 1;
          e.error = -1;
        }
        e.buf[e.storage - e.end_offs - 1] |= (unsigned char)window;
      }
    }
  }
}
OB_DEVNI void re_laplace(RangeEnc &e, int &value, unsigned fs, int decay) {  // celt/laplace.c:44-93
  unsigned fl = 0;
  int val = value;
  if (val) {
    const int s = -(val < 0);
    int i;
    val = (val + s) ^ s;
    fl = fs;
    fs = ((32768 - 32 - fs) * (int32_t)(16384 - decay)) >> 15;
    for (i = 1; fs > 0 && i < val; i++) {
      fs *= 2;
      fl += fs + 2;
      fs = (fs * (int32_t)decay) >> 15;
    }
    if (!fs) {
      int ndi_max = (int)(32768 - fl);
      ndi_max = (ndi_max - s) >> 1;
      const int di = OB_IMIN(val - i, ndi_max - 1);
      fl += (2 * di + 1 + s);
      fs = OB_IMIN(1, (int)(32768 - fl));
      value = (i + di + s) ^ s;
    } else {
      fs += 1;
      fl += fs & ~s;
    }
  }
  re_encode_bin(e, fl, fl + fs, 15);
}

// ---- float helpers in the reference's rounding (celt/mathops.h:226,346-347) ----
OB_DEV float enc_exp2(float x) { return (float)exp(0.6931471805599453094 * (double)x); }
OB_DEV float enc_log2(float x) { return (float)(1.442695040888963387 * log((double)x)); }
OB_DEV float enc_fmax(float a, float b) { return a > b ? a : b; }
OB_DEV float enc_fmin(float a, float b) { return a < b ? a : b; }
OB_DEVNI float enc_dot(const float *x, const float *y, int n) {  // celt_inner_prod_c
  float s = 0.f;
#pragma unroll 4
  for (int i = 0; i < n; i++) s = s + x[i] * y[i];
  return s;
}
OB_DEV float enc_median5(const float *x) {  // celt/celt_encoder.c:990-1029
  float t0, t1, t2 = x[2], t3, t4, tmp;
  if (x[0] > x[1]) { t0 = x[1]; t1 = x[0]; } else { t0 = x[0]; t1 = x[1]; }
  if (x[3] > x[4]) { t3 = x[4]; t4 = x[3]; } else { t3 = x[3]; t4 = x[4]; }
  if (t0 > t3) { tmp = t0; t0 = t3; t3 = tmp; tmp = t1; t1 = t4; t4 = tmp; }
  if (t2 > t1) return t1 < t3 ? enc_fmin(t2, t3) : enc_fmin(t4, t1);
  return t2 < t3 ? enc_fmin(t1, t3) : enc_fmin(t2, t4);
}
OB_DEV float enc_median3(const float *x) {
  float t0, t1;
  const float t2 = x[2];
  if (x[0] > x[1]) { t0 = x[1]; t1 = x[0]; } else { t0 = x[0]; t1 = x[1]; }
  if (t1 < t2) return t1;
  if (t0 < t2) return t2;
  return t0;
}

// ---- dynalloc_analysis (celt/celt_encoder.c:1049-1274; no lfe / surround, analysis invalid, CBR) ----
struct EncDyn { float maxDepth; int32_t tot_boost; };
OB_DEVNI EncDyn enc_dynalloc(const float *bandLogE, const float *bandLogE2, const float *oldBandE, int C, int *offsets, int lsb_depth, int isTransient, int LM,
                             int effectiveBytes, float tone_freq, float toneishness, int vbr, int constrained_vbr, int *importance,
                             int *spread_weight, int end, const ObAnalysisInfo *an) {
  float follower[2 * OB_NBANDS], noise_floor[OB_NBANDS], bandLogE3[OB_NBANDS];
  int32_t tot_boost = 0;
#pragma unroll 1
  for (int i = 0; i < end; i++) {
    offsets[i] = 0;
    noise_floor[i] = 0.0625f * ob_logN[i] + .5f + (9 - lsb_depth) - ob_emeans[i] + .0062f * (i + 5) * (i + 5);
  }
  EncDyn ret;
  float maxDepth = -31.9f;
#pragma unroll 1
  for (int c = 0; c < C; c++)
#pragma unroll 1
    for (int i = 0; i < end; i++) maxDepth = enc_fmax(maxDepth, bandLogE[c * OB_NBANDS + i] - noise_floor[i]);
  ret.maxDepth = maxDepth;
  {  // simple masking model -> spread_weight (celt/celt_encoder.c:1082-1112); follower[] doubles as mask[] / sig[] scratch
    float *mask = follower, *sig = follower + OB_NBANDS;
#pragma unroll 1
    for (int i = 0; i < end; i++) {
      mask[i] = bandLogE[i] - noise_floor[i];
      if (C == 2) mask[i] = enc_fmax(mask[i], bandLogE[OB_NBANDS + i] - noise_floor[i]);
      sig[i] = mask[i];
    }
#pragma unroll 1
    for (int i = 1; i < end; i++) mask[i] = enc_fmax(mask[i], mask[i - 1] - 2.f);
#pragma unroll 1
    for (int i = end - 2; i >= 0; i--) mask[i] = enc_fmax(mask[i], mask[i + 1] - 3.f);
#pragma unroll 1
    for (int i = 0; i < end; i++) {
      const float smr = sig[i] - enc_fmax(enc_fmax(0.f, maxDepth - 12.f), mask[i]);
      const int shift = OB_IMIN(5, OB_IMAX(0, -(int)floor((double)(.5f + smr))));
      spread_weight[i] = 32 >> shift;
    }
  }
  ret.tot_boost = 0;
  if (effectiveBytes < 30 + 5 * LM) {
#pragma unroll 1
    for (int i = 0; i < end; i++) importance[i] = 13;
    return ret;
  }
  int last = 0;
#pragma unroll 1
  for (int c = 0; c < C; c++) {
    float *f = &follower[c * OB_NBANDS];
#pragma unroll 1
    for (int i = 0; i < end; i++) bandLogE3[i] = bandLogE2[c * OB_NBANDS + i];
    if (LM == 0)
      for (int i = 0; i < 8; i++) bandLogE3[i] = enc_fmax(bandLogE2[c * OB_NBANDS + i], oldBandE[c * OB_NBANDS + i]);
    f[0] = bandLogE3[0];
#pragma unroll 1
    for (int i = 1; i < end; i++) {
      if (bandLogE3[i] > bandLogE3[i - 1] + .5f) last = i;
      f[i] = enc_fmin(f[i - 1] + 1.5f, bandLogE3[i]);
    }
#pragma unroll 1
    for (int i = last - 1; i >= 0; i--) f[i] = enc_fmin(f[i], enc_fmin(f[i + 1] + 2.f, bandLogE3[i]));
#pragma unroll 1
    for (int i = 2; i < end - 2; i++) f[i] = enc_fmax(f[i], enc_median5(&bandLogE3[i - 2]) - 1.f);
    float tmp = enc_median3(&bandLogE3[0]) - 1.f;
    f[0] = enc_fmax(f[0], tmp);
    f[1] = enc_fmax(f[1], tmp);
    tmp = enc_median3(&bandLogE3[end - 3]) - 1.f;
    f[end - 2] = enc_fmax(f[end - 2], tmp);
    f[end - 1] = enc_fmax(f[end - 1], tmp);
#pragma unroll 1
    for (int i = 0; i < end; i++) f[i] = enc_fmax(f[i], noise_floor[i]);
  }
  if (C == 2) {
#pragma unroll 1
    for (int i = 0; i < end; i++) {
      follower[OB_NBANDS + i] = enc_fmax(follower[OB_NBANDS + i], follower[i] - 4.f);
      follower[i] = enc_fmax(follower[i], follower[OB_NBANDS + i] - 4.f);
      follower[i] = .5f * (enc_fmax(0.f, bandLogE[i] - follower[i]) + enc_fmax(0.f, bandLogE[OB_NBANDS + i] - follower[OB_NBANDS + i]));
    }
  } else {
#pragma unroll 1
    for (int i = 0; i < end; i++) follower[i] = enc_fmax(0.f, bandLogE[i] - follower[i]);
  }
#pragma unroll 1
  for (int i = 0; i < end; i++) importance[i] = (int)floor((double)(.5f + 13 * enc_exp2(enc_fmin(follower[i], 4.f))));
  if ((!vbr || constrained_vbr) && !isTransient)
#pragma unroll 1
    for (int i = 0; i < end; i++) follower[i] = .5f * follower[i];
#pragma unroll 1
  for (int i = 0; i < end; i++) {
    if (i < 8) follower[i] *= 2;
    if (i >= 12) follower[i] = .5f * follower[i];
  }
  if (toneishness > .98f) {
    const int freq_bin = (int)floor(.5 + tone_freq * 120 / 3.14159265358979323846);
#pragma unroll 1
    for (int i = 0; i < end; i++) {
      if (freq_bin >= ob_ebands[i] && freq_bin <= ob_ebands[i + 1]) follower[i] += 2.f;
      if (freq_bin >= ob_ebands[i] - 1 && freq_bin <= ob_ebands[i + 1] + 1) follower[i] += 1.f;
      if (freq_bin >= ob_ebands[i] - 2 && freq_bin <= ob_ebands[i + 1] + 2) follower[i] += 1.f;
      if (freq_bin >= ob_ebands[i] - 3 && freq_bin <= ob_ebands[i + 1] + 3) follower[i] += .5f;
    }
    if (freq_bin >= ob_ebands[end]) {
      follower[end - 1] += 2.f;
      follower[end - 2] += 1.f;
    }
  }
  if (an && an->valid)   // celt/celt_encoder.c:1226-1230
#pragma unroll 1
    for (int i = 0; i < OB_IMIN(19, end); i++) follower[i] = follower[i] + (1.f / 64.f) * an->leak_boost[i];
  if (effectiveBytes > 320) follower[0] += enc_fmin(1.5f, 1e-3f * (effectiveBytes - 320));
#pragma unroll 1
  for (int i = 0; i < end; i++) {
    int boost, boost_bits;
    follower[i] = enc_fmin(follower[i], 4.f);
    const int width = C * (ob_ebands[i + 1] - ob_ebands[i]) << LM;
    if (width < 6) {
      boost = (int)follower[i];
      boost_bits = boost * width << OB_BITRES;
    } else if (width > 48) {
      boost = (int)(follower[i] * 8);
      boost_bits = (boost * width << OB_BITRES) / 8;
    } else {
      boost = (int)(follower[i] * width / 6);
      boost_bits = boost * 6 << OB_BITRES;
    }
    if ((!vbr || (constrained_vbr && !isTransient)) && (tot_boost + boost_bits) >> OB_BITRES >> 3 > 2 * effectiveBytes / 3) {
      const int32_t cap = ((2 * effectiveBytes / 3) << OB_BITRES << 3);
      offsets[i] = cap - tot_boost;
      tot_boost = cap;
      break;
    } else {
      offsets[i] = boost;
      tot_boost += boost_bits;
    }
  }
  ret.tot_boost = tot_boost;
  return ret;
}

// ---- quant_coarse_energy_impl (celt/quant_bands.c:156-254) ----
OB_DEVNI int enc_coarse_impl(RangeEnc &enc, const float *eBands, float *oldEBands, int32_t budget, int32_t tell, const unsigned char *prob_model,
                             float *error, int C, int LM, int intra, float max_decay, int end) {
  const int start = 0;
  int badness = 0;
  float prev[2] = {0.f, 0.f}, coef, beta;
  if (tell + 3 <= budget) re_bit_logp(enc, intra, 3);
  if (intra) { coef = 0.f; beta = 4915 / 32768.f; }
  else {
    coef = LM == 0 ? 29440 / 32768.f : LM == 1 ? 26112 / 32768.f : LM == 2 ? 21248 / 32768.f : 16384 / 32768.f;
    beta = LM == 0 ? 30147 / 32768.f : LM == 1 ? 22282 / 32768.f : LM == 2 ? 12124 / 32768.f : 6554 / 32768.f;
  }
#pragma unroll 1
  for (int i = start; i < end; i++)
#pragma unroll 1
    for (int c = 0; c < C; c++) {
      const float x = eBands[i + c * OB_NBANDS];
      const float oldE = enc_fmax(-9.f, oldEBands[i + c * OB_NBANDS]);
      const float f = x - coef * oldE - prev[c];
      int qi = (int)floor((double)(.5f + f));
      const float decay_bound = enc_fmax(-28.f, oldEBands[i + c * OB_NBANDS]) - max_decay;
      if (qi < 0 && x < decay_bound) {
        qi += (int)(decay_bound - x);
        if (qi > 0) qi = 0;
      }
      const int qi0 = qi;
      tell = re_tell(enc);
      const int bits_left = budget - tell - 3 * C * (end - i);
      if (i != start && bits_left < 30) {
        if (bits_left < 24) qi = OB_IMIN(1, qi);
        if (bits_left < 16) qi = OB_IMAX(-1, qi);
      }
      if (budget - tell >= 15) {
        const int pi = 2 * OB_IMIN(i, 20);
        re_laplace(enc, qi, prob_model[pi] << 7, prob_model[pi + 1] << 6);
      } else if (budget - tell >= 2) {
        qi = OB_IMAX(-1, OB_IMIN(qi, 1));
        re_icdf(enc, 2 * qi ^ -(qi < 0), ob_small_energy_icdf, 2);
      } else if (budget - tell >= 1) {
        qi = OB_IMIN(0, qi);
        re_bit_logp(enc, -qi, 1);
      } else qi = -1;
      error[i + c * OB_NBANDS] = f - qi;
      badness += qi0 > qi ? qi0 - qi : qi - qi0;
      const float q = (float)qi;
      oldEBands[i + c * OB_NBANDS] = coef * oldE + prev[c] + q;
      prev[c] = prev[c] + q - beta * q;
    }
  return badness;
}

OB_DEVNI void enc_haar(float *X, int N0, int stride);
// ---- tf_analysis (celt/celt_encoder.c:648-822) ----
OB_DEVNI float enc_l1(const float *tmp, int N, int LM, float bias) {
  float L1 = 0.f;
#pragma unroll 1
  for (int i = 0; i < N; i++) L1 += fabsf(tmp[i]);
  return L1 + (LM * bias) * L1;
}
OB_DEVNI int enc_tf_analysis(int isTransient, int *tf_res, int lambda, const float *X, int N0, int LM, float tf_estimate, int tf_chan,
                             const int *importance, float *tmp, float *tmp_1, int len) {
  int metric[OB_NBANDS];
  unsigned path0 = 0, path1 = 0;   // one bit per band
  const float bias = .04f * enc_fmax(-.25f, .5f - tf_estimate);
#pragma unroll 1
  for (int i = 0; i < len; i++) {
    const int N = (ob_ebands[i + 1] - ob_ebands[i]) << LM, narrow = (ob_ebands[i + 1] - ob_ebands[i]) == 1;
    int best_level = 0;
#pragma unroll 1
    for (int j = 0; j < N; j++) tmp[j] = X[tf_chan * N0 + (ob_ebands[i] << LM) + j];
    float L1 = enc_l1(tmp, N, isTransient ? LM : 0, bias), best_L1 = L1;
    if (isTransient && !narrow) {
#pragma unroll 1
      for (int j = 0; j < N; j++) tmp_1[j] = tmp[j];
      enc_haar(tmp_1, N >> LM, 1 << LM);
      L1 = enc_l1(tmp_1, N, LM + 1, bias);
      if (L1 < best_L1) { best_L1 = L1; best_level = -1; }
    }
#pragma unroll 1
    for (int k = 0; k < LM + !(isTransient || narrow); k++) {
      const int B = isTransient ? (LM - k - 1) : k + 1;
      enc_haar(tmp, N >> k, 1 << k);
      L1 = enc_l1(tmp, N, B, bias);
      if (L1 < best_L1) { best_L1 = L1; best_level = k + 1; }
    }
    metric[i] = isTransient ? 2 * best_level : -2 * best_level;
    if (narrow && (metric[i] == 0 || metric[i] == -2 * LM)) metric[i] -= 1;
  }
  int selcost[2], tf_select = 0, cost0, cost1;
#define OB_TFSEL(sel, w) (2 * ob_tf_select[LM * 8 + 4 * isTransient + 2 * (sel) + (w)])
#define OB_ABSD(a, b) ((a) > (b) ? (a) - (b) : (b) - (a))
#pragma unroll 1
  for (int sel = 0; sel < 2; sel++) {
    cost0 = importance[0] * OB_ABSD(metric[0], OB_TFSEL(sel, 0));
    cost1 = importance[0] * OB_ABSD(metric[0], OB_TFSEL(sel, 1)) + (isTransient ? 0 : lambda);
#pragma unroll 1
    for (int i = 1; i < len; i++) {
      const int curr0 = OB_IMIN(cost0, cost1 + lambda), curr1 = OB_IMIN(cost0 + lambda, cost1);
      cost0 = curr0 + importance[i] * OB_ABSD(metric[i], OB_TFSEL(sel, 0));
      cost1 = curr1 + importance[i] * OB_ABSD(metric[i], OB_TFSEL(sel, 1));
    }
    selcost[sel] = OB_IMIN(cost0, cost1);
  }
  if (selcost[1] < selcost[0] && isTransient) tf_select = 1;
  cost0 = importance[0] * OB_ABSD(metric[0], OB_TFSEL(tf_select, 0));
  cost1 = importance[0] * OB_ABSD(metric[0], OB_TFSEL(tf_select, 1)) + (isTransient ? 0 : lambda);
#pragma unroll 1
  for (int i = 1; i < len; i++) {
    int curr0, curr1, from0 = cost0, from1 = cost1 + lambda;
    if (from0 < from1) curr0 = from0; else { curr0 = from1; path0 |= 1u << i; }
    from0 = cost0 + lambda;
    from1 = cost1;
    if (from0 < from1) curr1 = from0; else { curr1 = from1; path1 |= 1u << i; }
    cost0 = curr0 + importance[i] * OB_ABSD(metric[i], OB_TFSEL(tf_select, 0));
    cost1 = curr1 + importance[i] * OB_ABSD(metric[i], OB_TFSEL(tf_select, 1));
  }
#undef OB_TFSEL
#undef OB_ABSD
  tf_res[len - 1] = cost0 < cost1 ? 0 : 1;
#pragma unroll 1
  for (int i = len - 2; i >= 0; i--) tf_res[i] = (int)(((tf_res[i + 1] == 1 ? path1 : path0) >> (i + 1)) & 1u);
  return tf_select;
}

// ---- spreading_decision (celt/bands.c:470-562) ----
OB_DEVNI int enc_spreading_decision(const float *X, ObEncState &st, int update_hf, int C, int M, const int *spread_weight, int end) {
  const int N0 = M * OB_SHORT_N;
  int sum = 0, nbBands = 0, hf_sum = 0;
  if (M * (ob_ebands[end] - ob_ebands[end - 1]) <= 8) return 0;
#pragma unroll 1
  for (int c = 0; c < C; c++)
#pragma unroll 1
    for (int i = 0; i < end; i++) {
      const int N = M * (ob_ebands[i + 1] - ob_ebands[i]);
      if (N <= 8) continue;
      const float *x = X + M * ob_ebands[i] + c * N0;
      int t0 = 0, t1 = 0, t2 = 0;
#pragma unroll 1
      for (int j = 0; j < N; j++) {
        const float x2N = (x[j] * x[j]) * N;
        if (x2N < 0.25f) t0++;
        if (x2N < 0.0625f) t1++;
        if (x2N < 0.015625f) t2++;
      }
      if (i > OB_NBANDS - 4) hf_sum += (uint32_t)(32 * (t1 + t0)) / (uint32_t)N;
      const int tmp = (2 * t2 >= N) + (2 * t1 >= N) + (2 * t0 >= N);
      sum += tmp * spread_weight[i];
      nbBands += spread_weight[i];
    }
  if (update_hf) {
    if (hf_sum) hf_sum = (uint32_t)hf_sum / (uint32_t)(C * (4 - OB_NBANDS + end));
    st.hf_average = (st.hf_average + hf_sum) >> 1;
    hf_sum = st.hf_average;
    if (st.tapset_decision == 2) hf_sum += 4;
    else if (st.tapset_decision == 0) hf_sum -= 4;
    if (hf_sum > 22) st.tapset_decision = 2;
    else if (hf_sum > 18) st.tapset_decision = 1;
    else st.tapset_decision = 0;
  }
  sum = (uint32_t)((int32_t)sum << 8) / (uint32_t)nbBands;
  sum = (sum + st.tonal_average) >> 1;
  st.tonal_average = sum;
  sum = (3 * sum + (((3 - st.spread_decision) << 7) + 64) + 2) >> 2;
  return sum < 80 ? 3 : sum < 256 ? 2 : sum < 384 ? 1 : 0;
}

// ---- compute_vbr (celt/celt_encoder.c:1604-1717; analysis invalid, no lfe / surround mask) ----
OB_DEVNI int32_t enc_compute_vbr(int32_t base_target, int LM, int32_t bitrate, int lastCodedBands, int C, int intensity, int constrained_vbr,
                                 float stereo_saving, int tot_boost, float tf_estimate, float maxDepth, float temporal_vbr,
                                 const ObAnalysisInfo *an, int pitch_change) {
  const int coded_bands = lastCodedBands ? lastCodedBands : OB_NBANDS;
  int coded_bins = ob_ebands[coded_bands] << LM;
  if (C == 2) coded_bins += ob_ebands[OB_IMIN(intensity, coded_bands)] << LM;
  int32_t target = base_target;
  if (an && an->valid && (double)an->activity < .4) target -= (int32_t)((coded_bins << OB_BITRES) * (.4f - an->activity));   // :1632-1633
  if (C == 2) {
    const int coded_stereo_bands = OB_IMIN(intensity, coded_bands);
    const int coded_stereo_dof = (ob_ebands[coded_stereo_bands] << LM) - coded_stereo_bands;
    const float max_frac = (0.8f * coded_stereo_dof) / coded_bins;
    stereo_saving = enc_fmin(stereo_saving, 1.f);
    const float a = max_frac * target, b = (stereo_saving - 0.1f) * (coded_stereo_dof << OB_BITRES);
    target -= (int32_t)enc_fmin(a, b);
  }
  target += tot_boost - (19 << LM);
  target += (int32_t)((tf_estimate - 0.044f) * target);
  if (an && an->valid) {   // tonality boost (:1658-1670)
    const float tonal = enc_fmax(0.f, an->tonality - .15f) - 0.12f;
    int32_t tonal_target = target + (int32_t)((coded_bins << OB_BITRES) * 1.2f * tonal);
    if (pitch_change) tonal_target += (int32_t)((coded_bins << OB_BITRES) * .8f);
    target = tonal_target;
  }
  const int bins = ob_ebands[OB_NBANDS - 2] << LM;
  int32_t floor_depth = (int32_t)((C * bins << OB_BITRES) * maxDepth);
  floor_depth = OB_IMAX(floor_depth, target >> 2);
  target = OB_IMIN(target, floor_depth);
  if (constrained_vbr) target = base_target + (int32_t)(0.67f * (target - base_target));
  if (tf_estimate < .2f) {
    const float amount = .0000031f * OB_IMAX(0, OB_IMIN(32000, 96000 - bitrate));
    const float tvbr_factor = temporal_vbr * amount;
    target += (int32_t)(tvbr_factor * target);
  }
  return OB_IMIN(2 * base_target, target);
}

// ---- clt_compute_allocation, encode side (celt/rate.c:248-637) ----
struct EncAlloc { int codedBands, intensity, dual_stereo; int32_t balance; };
OB_DEVNI EncAlloc enc_compute_allocation(RangeEnc &enc, const int *offsets, const int *cap, int alloc_trim, int intensity, int dual_stereo,
                                          int32_t total, int *bits, int *ebits, int *fine_priority, int C, int LM, int prev,
                                          int signalBandwidth, int end) {
  const int start = 0, len = OB_NBANDS, nbAllocVectors = 11, alloc_floor = C << OB_BITRES, logM = LM << OB_BITRES,
            stereo = C > 1;
  int bits1[OB_NBANDS], bits2[OB_NBANDS], thresh[OB_NBANDS], trim_offset[OB_NBANDS];
  int skip_start = start, skip_rsv, intensity_rsv = 0, dual_stereo_rsv = 0, lo, hi, j, codedBands;
  int32_t psum, left, percoeff, balance;
  total = OB_IMAX(total, 0);
  skip_rsv = total >= 1 << OB_BITRES ? 1 << OB_BITRES : 0;
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
#pragma unroll 1
  for (j = start; j < end; j++) {
    const int w = ob_ebands[j + 1] - ob_ebands[j];
    thresh[j] = OB_IMAX((C) << OB_BITRES, (3 * w << LM << OB_BITRES) >> 4);
    trim_offset[j] = C * w * (alloc_trim - 5 - LM) * (end - j - 1) * (1 << (LM + OB_BITRES)) >> 6;
    if (w << LM == 1) trim_offset[j] -= C << OB_BITRES;
  }
  lo = 1;
  hi = nbAllocVectors - 1;
  do {
    int done = 0;
    const int mid = (lo + hi) >> 1;
    psum = 0;
#pragma unroll 1
    for (j = end; j-- > start;) {
      const int N = ob_ebands[j + 1] - ob_ebands[j];
      int bitsj = C * N * ob_band_alloc[mid * len + j] << LM >> 2;
      if (bitsj > 0) bitsj = OB_IMAX(0, bitsj + trim_offset[j]);
      bitsj += offsets[j];
      if (bitsj >= thresh[j] || done) { done = 1; psum += OB_IMIN(bitsj, cap[j]); }
      else if (bitsj >= C << OB_BITRES) psum += C << OB_BITRES;
    }
    if (psum > total) hi = mid - 1; else lo = mid + 1;
  } while (lo <= hi);
  hi = lo--;
#pragma unroll 1
  for (j = start; j < end; j++) {
    const int N = ob_ebands[j + 1] - ob_ebands[j];
    int bits1j = C * N * ob_band_alloc[lo * len + j] << LM >> 2;
    int bits2j = hi >= nbAllocVectors ? cap[j] : C * N * ob_band_alloc[hi * len + j] << LM >> 2;
    if (bits1j > 0) bits1j = OB_IMAX(0, bits1j + trim_offset[j]);
    if (bits2j > 0) bits2j = OB_IMAX(0, bits2j + trim_offset[j]);
    if (lo > 0) bits1j += offsets[j];
    bits2j += offsets[j];
    if (offsets[j] > 0) skip_start = j;
    bits2j = OB_IMAX(0, bits2j - bits1j);
    bits1[j] = bits1j;
    bits2[j] = bits2j;
  }
  // interp_bits2pulses
  lo = 0;
  hi = 1 << 6;
#pragma unroll 1
  for (int it = 0; it < 6; it++) {
    const int mid = (lo + hi) >> 1;
    int done = 0;
    psum = 0;
#pragma unroll 1
    for (j = end; j-- > start;) {
      const int tmp = bits1[j] + (mid * (int32_t)bits2[j] >> 6);
      if (tmp >= thresh[j] || done) { done = 1; psum += OB_IMIN(tmp, cap[j]); }
      else if (tmp >= alloc_floor) psum += alloc_floor;
    }
    if (psum > total) hi = mid; else lo = mid;
  }
  psum = 0;
  {
    int done = 0;
#pragma unroll 1
    for (j = end; j-- > start;) {
      int tmp = bits1[j] + ((int32_t)lo * bits2[j] >> 6);
      if (tmp < thresh[j] && !done) tmp = tmp >= alloc_floor ? alloc_floor : 0;
      else done = 1;
      tmp = OB_IMIN(tmp, cap[j]);
      bits[j] = tmp;
      psum += tmp;
    }
  }
#pragma unroll 1
  for (codedBands = end;; codedBands--) {
    j = codedBands - 1;
    if (j <= skip_start) { total += skip_rsv; break; }
    left = total - psum;
    percoeff = (uint32_t)left / (uint32_t)(ob_ebands[codedBands] - ob_ebands[start]);
    left -= (ob_ebands[codedBands] - ob_ebands[start]) * percoeff;
    const int rem = OB_IMAX(left - (ob_ebands[j] - ob_ebands[start]), 0);
    const int band_width = ob_ebands[codedBands] - ob_ebands[j];
    int band_bits = (int)(bits[j] + percoeff * band_width + rem);
    if (band_bits >= OB_IMAX(thresh[j], alloc_floor + (1 << OB_BITRES))) {
      const int depth_threshold = codedBands > 17 ? (j < prev ? 7 : 9) : 0;   // the encoder's only free choice (rate.c:346-369)
      if (codedBands <= start + 2 || (band_bits > (depth_threshold * band_width << LM << OB_BITRES) >> 4 && j <= signalBandwidth)) {
        re_bit_logp(enc, 1, 1);
        break;
      }
      re_bit_logp(enc, 0, 1);
      psum += 1 << OB_BITRES;
      band_bits -= 1 << OB_BITRES;
    }
    psum -= bits[j] + intensity_rsv;
    if (intensity_rsv > 0) intensity_rsv = ob_log2_frac_tab[j - start];
    psum += intensity_rsv;
    if (band_bits >= alloc_floor) { psum += alloc_floor; bits[j] = alloc_floor; }
    else bits[j] = 0;
  }
  if (intensity_rsv > 0) {
    intensity = OB_IMIN(intensity, codedBands);
    re_uint(enc, intensity - start, codedBands + 1 - start);
  } else intensity = 0;
  if (intensity <= start) { total += dual_stereo_rsv; dual_stereo_rsv = 0; }
  if (dual_stereo_rsv > 0) re_bit_logp(enc, dual_stereo, 1);
  else dual_stereo = 0;
  left = total - psum;
  percoeff = (uint32_t)left / (uint32_t)(ob_ebands[codedBands] - ob_ebands[start]);
  left -= (ob_ebands[codedBands] - ob_ebands[start]) * percoeff;
#pragma unroll 1
  for (j = start; j < codedBands; j++) bits[j] += ((int)percoeff * (ob_ebands[j + 1] - ob_ebands[j]));
#pragma unroll 1
  for (j = start; j < codedBands; j++) {
    const int tmp = (int)OB_IMIN(left, ob_ebands[j + 1] - ob_ebands[j]);
    bits[j] += tmp;
    left -= tmp;
  }
  balance = 0;
#pragma unroll 1
  for (j = start; j < codedBands; j++) {
    const int N0 = ob_ebands[j + 1] - ob_ebands[j], N = N0 << LM;
    int32_t excess;
    const int32_t bit = (int32_t)bits[j] + balance;
    if (N > 1) {
      excess = OB_IMAX(bit - cap[j], 0);
      bits[j] = bit - excess;
      const int den = (C * N + ((C == 2 && N > 2 && !dual_stereo && j < intensity) ? 1 : 0));
      const int NClogN = den * (ob_logN[j] + logM);
      int offset = (NClogN >> 1) - den * 21;   // FINE_OFFSET
      if (N == 2) offset += den << OB_BITRES >> 2;
      if (bits[j] + offset < den * 2 << OB_BITRES) offset += NClogN >> 2;
      else if (bits[j] + offset < den * 3 << OB_BITRES) offset += NClogN >> 3;
      ebits[j] = OB_IMAX(0, (bits[j] + offset + (den << (OB_BITRES - 1))));
      ebits[j] = ((uint32_t)ebits[j] / (uint32_t)den) >> OB_BITRES;
      if (C * ebits[j] > (bits[j] >> OB_BITRES)) ebits[j] = bits[j] >> stereo >> OB_BITRES;
      ebits[j] = OB_IMIN(ebits[j], 8);   // MAX_FINE_BITS
      fine_priority[j] = ebits[j] * (den << OB_BITRES) >= bits[j] + offset;
      bits[j] -= C * ebits[j] << OB_BITRES;
    } else {
      excess = OB_IMAX(0, bit - (C << OB_BITRES));
      bits[j] = bit - excess;
      ebits[j] = 0;
      fine_priority[j] = 1;
    }
    if (excess > 0) {
      const int extra_fine = OB_IMIN(excess >> (stereo + OB_BITRES), 8 - ebits[j]);
      ebits[j] += extra_fine;
      const int extra_bits = extra_fine * C << OB_BITRES;
      fine_priority[j] = extra_bits >= excess - balance;
      excess -= extra_bits;
    }
    balance = excess;
  }
#pragma unroll 1
  for (; j < end; j++) {
    ebits[j] = bits[j] >> stereo >> OB_BITRES;
    bits[j] = 0;
    fine_priority[j] = ebits[j] < 1;
  }
  EncAlloc r;
  r.codedBands = codedBands; r.intensity = intensity; r.dual_stereo = dual_stereo; r.balance = balance;
  return r;
}

// ---- band-level vector helpers on thread-private arrays ----
// (the band vectors live in thread-private local memory: every loop below issues its loads ahead of the arithmetic and the stores that
// depend on them, four elements at a time where the elements are independent, so one memory latency is paid per group, not per element)
OB_DEVNI void enc_haar(float *X, int N0, int stride) {  // celt/bands.c:623-636
  N0 >>= 1;
#pragma unroll 1
  for (int i = 0; i < stride; i++) {
    float *P = X + i;
    const int s2 = 2 * stride;
    int j = 0;
#pragma unroll 1
    for (; j + 4 <= N0; j += 4) {   // the butterflies are independent
      float *Q = P + s2 * j;
      const float a0 = Q[0], b0 = Q[stride], a1 = Q[s2], b1 = Q[s2 + stride], a2 = Q[2 * s2], b2 = Q[2 * s2 + stride], a3 = Q[3 * s2], b3 = Q[3 * s2 + stride];
      const float t10 = .70710678f * a0, t20 = .70710678f * b0, t11 = .70710678f * a1, t21 = .70710678f * b1;
      const float t12 = .70710678f * a2, t22 = .70710678f * b2, t13 = .70710678f * a3, t23 = .70710678f * b3;
      Q[0] = t10 + t20; Q[stride] = t10 - t20;
      Q[s2] = t11 + t21; Q[s2 + stride] = t11 - t21;
      Q[2 * s2] = t12 + t22; Q[2 * s2 + stride] = t12 - t22;
      Q[3 * s2] = t13 + t23; Q[3 * s2 + stride] = t13 - t23;
    }
#pragma unroll 1
    for (; j < N0; j++) {
      const float t1 = .70710678f * P[s2 * j], t2 = .70710678f * P[s2 * j + stride];
      P[s2 * j] = t1 + t2;
      P[s2 * j + stride] = t1 - t2;
    }
  }
}
OB_TABLE unsigned char ob_enc_hadamard_order[30] = {1, 0, 3, 0, 2, 1, 7, 0, 4, 3, 6, 1, 5, 2, 15, 0, 8, 7, 12, 3, 11, 4, 14, 1, 9, 6, 13, 2, 10, 5};
OB_DEVNI void enc_deinterleave(float *__restrict__ X, float *__restrict__ tmp, int N0, int stride, int hadamard) {  // celt/bands.c:574-597
  const int N = N0 * stride;
#pragma unroll 1
  for (int i = 0; i < stride; i++) {
    const int row = hadamard ? ob_enc_hadamard_order[stride - 2 + i] : i;
    float *__restrict__ d = tmp + row * N0;
    const float *__restrict__ q = X + i;
#pragma unroll 4
    for (int j = 0; j < N0; j++) d[j] = q[j * stride];
  }
#pragma unroll 4
  for (int j = 0; j < N; j++) X[j] = tmp[j];
}
OB_DEVNI void enc_rotate_pass(float *X, int len, int stride, float c, float s) {  // celt/vq.c:75-102
  const float ms = -s;
  if (stride == 1) {   // the element a step has just written is the next step's x1 (forward) / x2 (backward): carried in a register
    if (len < 2) return;
    float cur = X[0];
#pragma unroll 4
    for (int i = 0; i < len - 1; i++) {
      const float x2 = X[i + 1];
      X[i] = c * cur + ms * x2;
      cur = c * x2 + s * cur;
    }
    X[len - 1] = cur;
    if (len < 3) return;
    cur = X[len - 2];
#pragma unroll 4
    for (int i = len - 3; i >= 0; i--) {
      const float x1 = X[i];
      X[i + 1] = c * cur + s * x1;
      cur = c * x1 + ms * cur;
    }
    X[0] = cur;
    return;
  }
  // stride >= 2: steps i and i + 1 touch four different elements
  int i = 0;
#pragma unroll 1
  for (; i + 2 <= len - stride; i += 2) {
    const float x1 = X[i], x2 = X[i + stride], z1 = X[i + 1], z2 = X[i + 1 + stride];
    X[i + stride] = c * x2 + s * x1;
    X[i] = c * x1 + ms * x2;
    X[i + 1 + stride] = c * z2 + s * z1;
    X[i + 1] = c * z1 + ms * z2;
  }
  for (; i < len - stride; i++) {
    const float x1 = X[i], x2 = X[i + stride];
    X[i + stride] = c * x2 + s * x1;
    X[i] = c * x1 + ms * x2;
  }
  i = len - 2 * stride - 1;
#pragma unroll 1
  for (; i >= 1; i -= 2) {
    const float x1 = X[i], x2 = X[i + stride], z1 = X[i - 1], z2 = X[i - 1 + stride];
    X[i + stride] = c * x2 + s * x1;
    X[i] = c * x1 + ms * x2;
    X[i - 1 + stride] = c * z2 + s * z1;
    X[i - 1] = c * z1 + ms * z2;
  }
  for (; i >= 0; i--) {
    const float x1 = X[i], x2 = X[i + stride];
    X[i + stride] = c * x2 + s * x1;
    X[i] = c * x1 + ms * x2;
  }
}
OB_DEVNI void enc_exp_rotation(float *X, int len, int stride, int K, int spread) {  // celt/vq.c:104-147, dir = +1
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
    enc_rotate_pass(X + i * len, len, 1, c, -s);
    if (stride2) enc_rotate_pass(X + i * len, len, stride2, s, -c);
  }
}
// op_pvq_search_c (celt/vq.c:205-374) followed by icwrs (celt/cwrs.c:444-460): returns the codeword index.
// X: the (rotated) partition, overwritten with |X|; y2: scratch holding 2*pulses
// (the search returns yy for the resynthesis-enabled encoder; signs: bit set of the negative inputs)
OB_DEVNI float enc_pvq_search(float *__restrict__ X, float *__restrict__ y2, int N, int K, uint32_t *signs) {
  float sum = 0.f, xy = 0.f, yy = 0.f;   // sign bits live in a small bit set (N <= 176)
#pragma unroll
  for (int j = 0; j < 6; j++) signs[j] = 0;
#pragma unroll 2
  for (int j = 0; j < N; j++) {
    const float x = X[j];
    if (x < 0) signs[j >> 5] |= 1u << (j & 31);
    X[j] = fabsf(x);
    y2[j] = 0.f;
  }
  int pulsesLeft = K;
  if (K > (N >> 1)) {
#pragma unroll 4
    for (int j = 0; j < N; j++) sum += X[j];
    if (!(sum > 1e-15f && sum < 64.f)) {
      X[0] = 1.f;
#pragma unroll 1
      for (int j = 1; j < N; j++) X[j] = 0.f;
      sum = 1.f;
    }
    const float rcp = (K + 0.8f) * (1.f / sum);
#pragma unroll 4
    for (int j = 0; j < N; j++) {
      const float x = X[j];
      const int iyj = (int)floorf(rcp * x);
      const float yj = (float)iyj;
      yy = yy + yj * yj;
      xy = xy + x * yj;
      y2[j] = yj * 2;
      pulsesLeft -= iyj;
    }
  }
  if (pulsesLeft > N + 3) {
    const float tmp = (float)pulsesLeft;
    yy = yy + tmp * tmp;
    yy = yy + tmp * y2[0];
    y2[0] += 2.f * tmp;
    pulsesLeft = 0;
  }
#pragma unroll 1
  for (int i = 0; i < pulsesLeft; i++) {
    int best_id = 0;
    yy = yy + 1;
    float Rxy = xy + X[0], Ryy = yy + y2[0];
    Rxy = Rxy * Rxy;
    float best_den = Ryy, best_num = Rxy;
#pragma unroll 4
    for (int j = 1; j < N; j++) {
      Rxy = xy + X[j];
      Ryy = yy + y2[j];
      Rxy = Rxy * Rxy;
      if (best_den * Rxy > Ryy * best_num) { best_den = Ryy; best_num = Rxy; best_id = j; }
    }
    xy = xy + X[best_id];
    yy = yy + y2[best_id];
    y2[best_id] += 2.f;
  }
  return yy;
}
// icwrs over the signed pulse vector
OB_DEVNI uint32_t enc_icwrs(const float *y2, const uint32_t *signs, int N) {
  int j = N - 1;
  int yj = (int)(y2[j] * .5f);
  uint32_t idx = ((signs[j >> 5] >> (j & 31)) & 1u) && yj ? 1u : 0u;
  int k = yj;
  do {
    j--;
    idx += pvq_u(N - j, k);
    yj = (int)(y2[j] * .5f);
    k += yj;
    if (((signs[j >> 5] >> (j & 31)) & 1u) && yj) idx += pvq_u(N - j, k + 1);
  } while (j > 0);
  return idx;
}
OB_DEV uint32_t enc_pvq_index(float *X, float *y2, int N, int K) {
  uint32_t signs[6];
  enc_pvq_search(X, y2, N, K, signs);
  return enc_icwrs(y2, signs, N);
}

// ---- stereo_itheta (celt/vq.c:724-757) and the float arctangent (celt/mathops.h:142-186) ----
OB_DEV float enc_atan_norm(float x) {
  const float x_sq = x * x;
  return 0.636619772367581f *
         (x + x * x_sq * (-3.3331659436225891113281250000e-01f +
                          x_sq * (1.99627041816711425781250000000e-01f +
                                  x_sq * (-1.3976582884788513183593750000e-01f +
                                          x_sq * (9.79423448443412780761718750000e-02f +
                                                  x_sq * (-5.7773590087890625000000000000e-02f +
                                                          x_sq * (2.30401363223791122436523437500e-02f + x_sq * (-4.3554059229791164398193359375e-03f))))))));
}
OB_DEV int enc_itheta(const float *X, const float *Y, int stereo, int N) {
  float Emid = 0.f, Eside = 0.f;
  if (stereo) {
#pragma unroll 1
    for (int i = 0; i < N; i++) {
      const float m = X[i] + Y[i], s = X[i] - Y[i];
      Emid = Emid + m * m;
      Eside = Eside + s * s;
    }
  } else {
    Emid += enc_dot(X, X, N);
    Eside += enc_dot(Y, Y, N);
  }
  const float mid = f_sqrt(Emid), side = f_sqrt(Eside);
  float a;
  if ((mid * mid + side * side) < 1e-18f) a = 0.f;
  else if (side < mid) a = enc_atan_norm(side / mid);
  else a = 1.f - enc_atan_norm(mid / side);
  return (int)((int32_t)floor((double)(.5f + 65536.f * 16384 * a)) >> 16);
}

// ---- the band coder ----
struct EncBand {
  RangeEnc *ec;
  const float *bandE;   // this stream's [C][21]
  int band, intensity, spread, tf_change, disable_inv, avoid_split_noise;
  int32_t remaining_bits;
  int theta_round;   // -1 / +1: the two candidates of the theta RDO (complexity >= 8); 0: plain rounding
  uint32_t seed;     // noise-fill LCG of the resynthesis path
};
struct EncTheta { int itheta, delta, qalloc, imid, iside, inv; };
// encode side of compute_theta (celt/bands.c:700-933); X/Y are transformed for the stereo case
OB_DEVNI EncTheta enc_theta(EncBand &ctx, float *X, float *Y, int N, int &b, int B, int B0, int LM, int stereo) {
  RangeEnc &ec = *ctx.ec;
  const int i = ctx.band;
  const int pulse_cap = ob_logN[i] + LM * (1 << OB_BITRES);
  const int offset = (pulse_cap >> 1) - (stereo && N == 2 ? 16 : 4);   // QTHETA_OFFSET_TWOPHASE : QTHETA_OFFSET
  int qn = theta_levels(N, b, offset, pulse_cap, stereo);
  if (stereo && i >= ctx.intensity) qn = 1;
  int itheta = enc_itheta(X, Y, stereo, N), inv = 0;
  const int32_t tell = re_tell_frac(ec);
  (void)B;
  if (qn != 1) {
    if (!stereo || ctx.theta_round == 0) {
      itheta = (itheta * (int32_t)qn + 8192) >> 14;
      if (!stereo && ctx.avoid_split_noise && itheta > 0 && itheta < qn) {
        const int unquantized = (uint32_t)((int32_t)itheta * 16384) / (uint32_t)qn;
        const int imid = bitexact_cos((int16_t)unquantized), iside = bitexact_cos((int16_t)(16384 - unquantized));
        const int delta = frac_mul16((N - 1) << 7, bitexact_log2tan(iside, imid));
        if (delta > b) itheta = qn;
        else if (delta < -b) itheta = 0;
      }
    } else {   // celt/bands.c:769-778: the RDO candidates, biased towards 0 and 16384
      const int bias = itheta > 8192 ? 32767 / qn : -32767 / qn;
      const int down = OB_IMIN(qn - 1, OB_IMAX(0, (itheta * (int32_t)qn + bias) >> 14));
      itheta = ctx.theta_round < 0 ? down : down + 1;
    }
    if (stereo && N > 2) {
      const int p0 = 3, x = itheta, x0 = qn / 2, ft = p0 * (x0 + 1) + x0;
      re_encode(ec, x <= x0 ? p0 * x : (x - 1 - x0) + (x0 + 1) * p0, x <= x0 ? p0 * (x + 1) : (x - x0) + (x0 + 1) * p0, ft);
    } else if (B0 > 1 || stereo) {
      re_uint(ec, itheta, qn + 1);
    } else {
      const int ft = ((qn >> 1) + 1) * ((qn >> 1) + 1);
      const int fs = itheta <= (qn >> 1) ? itheta + 1 : qn + 1 - itheta;
      const int fl = itheta <= (qn >> 1) ? itheta * (itheta + 1) >> 1 : ft - ((qn + 1 - itheta) * (qn + 2 - itheta) >> 1);
      re_encode(ec, fl, fl + fs, ft);
    }
    itheta = (uint32_t)((int32_t)itheta * 16384) / (uint32_t)qn;
    if (stereo) {
      if (itheta == 0) {   // intensity_stereo (celt/bands.c:379-403)
        const float left = ctx.bandE[i], right = ctx.bandE[i + OB_NBANDS];
        const float norm = 1e-15f + f_sqrt(1e-15f + left * left + right * right);
        const float a1 = left / norm, a2 = right / norm;
#pragma unroll 1
        for (int j = 0; j < N; j++) X[j] = a1 * X[j] + a2 * Y[j];
      } else {             // stereo_split
#pragma unroll 1
        for (int j = 0; j < N; j++) {
          const float l = .70710678f * X[j], r = .70710678f * Y[j];
          X[j] = l + r;
          Y[j] = r - l;
        }
      }
    }
  } else if (stereo) {
    inv = itheta > 8192 && !ctx.disable_inv;
    if (inv)
#pragma unroll 1
      for (int j = 0; j < N; j++) Y[j] = -Y[j];
    {
      const float left = ctx.bandE[i], right = ctx.bandE[i + OB_NBANDS];
      const float norm = 1e-15f + f_sqrt(1e-15f + left * left + right * right);
      const float a1 = left / norm, a2 = right / norm;
#pragma unroll 1
      for (int j = 0; j < N; j++) X[j] = a1 * X[j] + a2 * Y[j];
    }
    if (b > 2 << OB_BITRES && ctx.remaining_bits > 2 << OB_BITRES) re_bit_logp(ec, inv, 2);
    else inv = 0;
    if (ctx.disable_inv) inv = 0;
    itheta = 0;
  }
  EncTheta r;
  r.qalloc = (int)(re_tell_frac(ec) - tell);
  b -= r.qalloc;
  r.itheta = itheta;
  r.inv = inv;
  if (itheta == 0) { r.delta = -16384; r.imid = 32767; r.iside = 0; }
  else if (itheta == 16384) { r.delta = 16384; r.imid = 0; r.iside = 32767; }
  else {
    r.imid = bitexact_cos((int16_t)itheta);
    r.iside = bitexact_cos((int16_t)(16384 - itheta));
    r.delta = frac_mul16((N - 1) << 7, bitexact_log2tan(r.iside, r.imid));
  }
  return r;
}
// quant_partition, encode side (celt/bands.c:973-1176), the recursion unrolled onto an explicit stack (depth <= LM+2)
struct EncPart { short off, N; int b; signed char B, LM, phase, second_x; short itheta; int second_bits, rebalance; };
OB_DEVNI void enc_partition(EncBand &ctx, float *X, float *y2, int N0, int b0, int B0_, int LM0) {
  EncPart st[6];
  int sp = 0;
  st[0].off = 0; st[0].N = (short)N0; st[0].b = b0; st[0].B = (signed char)B0_; st[0].LM = (signed char)LM0; st[0].phase = 0;
  // two alternating steps for all lanes together, see rdo_partition: (A) walk the tree to the next leaf with pulses, (B) quantise it
  const unsigned lanes = converged_lanes();
  int pend_K = 0;
  for (;;) {
  while (sp >= 0 && !pend_K) {
    EncPart &f = st[sp];
    const int i = ctx.band;
    if (f.phase == 0) {
      const unsigned char *cache = pulse_cache_row(i, f.LM);
      if (f.LM != -1 && f.b > cache[cache[0]] + 12 && f.N > 2) {
        const int N = f.N >> 1, LM = f.LM - 1, Bp = f.B, B = (f.B + 1) >> 1;
        int b = f.b;
        const EncTheta t = enc_theta(ctx, X + f.off, X + f.off + N, N, b, B, Bp, LM, 0);
        int delta = t.delta;
        if (Bp > 1 && (t.itheta & 0x3fff)) {
          if (t.itheta > 8192) delta -= delta >> (4 - LM);
          else delta = OB_IMIN(0, delta + (N << OB_BITRES >> (5 - LM)));
        }
        const int mbits = OB_IMAX(0, OB_IMIN(b, (b - delta) / 2)), sbits = b - mbits;
        ctx.remaining_bits -= t.qalloc;
        f.rebalance = ctx.remaining_bits;
        f.itheta = (short)t.itheta;
        f.phase = 1;
        EncPart &ch = st[sp + 1];
        ch.N = (short)N; ch.B = (signed char)B; ch.LM = (signed char)LM; ch.phase = 0;
        if (mbits >= sbits) { ch.off = f.off; ch.b = mbits; f.second_x = 0; f.second_bits = sbits; f.b = mbits; }
        else { ch.off = (short)(f.off + N); ch.b = sbits; f.second_x = 1; f.second_bits = mbits; f.b = sbits; }
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
        if (q != 0) { pend_K = pseudo_to_pulses(q); continue; }   // step (B) below
        sp--;
      }
    } else if (f.phase == 1) {   // first half done: rebalance, then the other half
      const int N = f.N >> 1;
      int second = f.second_bits;
      const int rebalance = f.b - (f.rebalance - ctx.remaining_bits);
      if (rebalance > 3 << OB_BITRES && f.itheta != (f.second_x ? 16384 : 0)) second += rebalance - (3 << OB_BITRES);
      f.phase = 2;
      EncPart &ch = st[sp + 1];
      ch.N = (short)N; ch.B = (signed char)((f.B + 1) >> 1); ch.LM = (signed char)(f.LM - 1); ch.phase = 0;
      ch.off = (short)(f.second_x ? f.off : f.off + N);
      ch.b = second;
      sp++;
    } else sp--;
  }
    if (!ballot_among(lanes, pend_K != 0)) break;
    if (pend_K) {   // alg_quant (celt/vq.c:552-617) without resynthesis
      EncPart &f = st[sp];
      const int K = pend_K;
      enc_exp_rotation(X + f.off, f.N, f.B, K, ctx.spread);
      const uint32_t idx = enc_pvq_index(X + f.off, y2, f.N, K);
      re_uint(*ctx.ec, idx, pvq_u(f.N, K) + pvq_u(f.N, K + 1));
      pend_K = 0;
      sp--;
    }
  }
}
OB_DEV void enc_band_n1(EncBand &ctx, const float *X, const float *Y) {  // celt/bands.c:934-968
  if (ctx.remaining_bits >= 1 << OB_BITRES) {
    re_bits(*ctx.ec, X[0] < 0, 1);
    ctx.remaining_bits -= 1 << OB_BITRES;
  }
  if (Y && ctx.remaining_bits >= 1 << OB_BITRES) {
    re_bits(*ctx.ec, Y[0] < 0, 1);
    ctx.remaining_bits -= 1 << OB_BITRES;
  }
}
// quant_band, encode side (celt/bands.c:1248-1379)
OB_DEVNI void enc_band_mono(EncBand &ctx, float *X, float *tmp, int N, int b, int B, int LM) {
  int N_B = N, recombine = 0, tf_change = ctx.tf_change;
  const int longBlocks = B == 1;
  N_B = (unsigned)N_B / (unsigned)B;
  if (N == 1) { enc_band_n1(ctx, X, nullptr); return; }
  if (tf_change > 0) recombine = tf_change;
#pragma unroll 1
  for (int k = 0; k < recombine; k++) enc_haar(X, N >> k, 1 << k);
  B >>= recombine;
  N_B <<= recombine;
  while ((N_B & 1) == 0 && tf_change < 0) {
    enc_haar(X, N_B, B);
    B <<= 1;
    N_B >>= 1;
    tf_change++;
  }
  if (B > 1) enc_deinterleave(X, tmp, N_B >> recombine, B << recombine, longBlocks);
  enc_partition(ctx, X, tmp, N, b, B, LM);
}
// quant_band_stereo, encode side (celt/bands.c:1387-1573)
OB_DEVNI void enc_band_stereo(EncBand &ctx, float *X, float *Y, float *tmp, int N, int b, int B, int LM) {
  if (N == 1) { enc_band_n1(ctx, X, Y); return; }
  const float eL = ctx.bandE[ctx.band], eR = ctx.bandE[OB_NBANDS + ctx.band];
  if (eL < 1e-10f || eR < 1e-10f) {
    if (eL > eR) { for (int j = 0; j < N; j++) Y[j] = X[j]; }
    else { for (int j = 0; j < N; j++) X[j] = Y[j]; }
  }
  const EncTheta t = enc_theta(ctx, X, Y, N, b, B, B, LM, 1);
  // one pair of enc_band_mono calls for the three shapes of the band (see rdo_band_stereo): threads that enter a function from
  // different call sites do not reconverge inside it
  int mbits, sbits, first_is_mid = 1;
  float *xf = X;
  if (N == 2) {
    mbits = b; sbits = 0;
    if (t.itheta != 0 && t.itheta != 16384) sbits = 1 << OB_BITRES;
    mbits -= sbits;
    const int c = t.itheta > 8192;
    ctx.remaining_bits -= t.qalloc + sbits;
    float *x2 = c ? Y : X, *y2 = c ? X : Y;
    if (sbits) re_bits(*ctx.ec, x2[0] * y2[1] - x2[1] * y2[0] < 0, 1);
    xf = x2;
  } else {
    mbits = OB_IMAX(0, OB_IMIN(b, (b - t.delta) / 2)); sbits = b - mbits;
    ctx.remaining_bits -= t.qalloc;
    first_is_mid = mbits >= sbits;
    xf = first_is_mid ? X : Y;
  }
  int32_t rebalance = ctx.remaining_bits;
  enc_band_mono(ctx, xf, tmp, N, (N == 2 || first_is_mid) ? mbits : sbits, B, LM);
  if (N != 2) {
    if (first_is_mid) {
      rebalance = mbits - (rebalance - ctx.remaining_bits);
      if (rebalance > 3 << OB_BITRES && t.itheta != 0) sbits += rebalance - (3 << OB_BITRES);
    } else {
      rebalance = sbits - (rebalance - ctx.remaining_bits);
      if (rebalance > 3 << OB_BITRES && t.itheta != 16384) mbits += rebalance - (3 << OB_BITRES);
    }
    enc_band_mono(ctx, first_is_mid ? Y : X, tmp, N, first_is_mid ? sbits : mbits, B, LM);
  }
}

}  // namespace ob
#include "celt_enc_rdo.cuh"
namespace ob {

// ---- frame driver ----  RDO: compiled with the theta-RDO band coder (complexity >= 8; a second kernel, because its
// folding / save buffers triple the thread's local memory)
template <bool RDO, bool SYNC = false>
OB_DEV void enc_frame_thread(const EncFrameParams &p, int stream) {
  const int LM = p.LM, M = 1 << LM, N = M * OB_SHORT_N, start = 0;
  const int C = p.ctl ? p.ctl[stream].C : p.C, end = p.ctl ? p.ctl[stream].end : p.end, SC = p.ctl ? (p.CC ? p.CC : p.C) : p.C;
  const ObAnalysisInfo *an = p.ctl ? &p.ctl[stream].analysis : nullptr;
  ObEncState &st = p.state[stream];
  const ObEncFrameInfo fi = p.info[stream];
  float *Xg = p.X + (size_t)stream * SC * N;
  const float *bandE = p.bandE + (size_t)stream * SC * OB_NBANDS;
  float bandLogE[2 * OB_NBANDS], error[2 * OB_NBANDS];
  int fine_quant[OB_NBANDS], pulses[OB_NBANDS], cap[OB_NBANDS], offsets[OB_NBANDS], fine_priority[OB_NBANDS], tf_res[OB_NBANDS];
  float bx[176], by[176], tmp[176];
#pragma unroll 1
  for (int i = 0; i < C * OB_NBANDS; i++) bandLogE[i] = p.bandLogE[(size_t)stream * SC * OB_NBANDS + i];
  float bl2[2 * OB_NBANDS];
  // bandLogE2 differs from bandLogE for a patched frame and, from complexity 8, for every transient frame (second MDCT)
  const bool have_bl2 = p.bandLogE2 && (fi.patched || (p.complexity >= 8 && fi.short_blocks));
  if (have_bl2)
#pragma unroll 1
    for (int i = 0; i < C * OB_NBANDS; i++) bl2[i] = p.bandLogE2[(size_t)stream * SC * OB_NBANDS + i];

  int importance[OB_NBANDS], spread_weight[OB_NBANDS];
  const EncBudget bud = enc_budget(p.max_bytes, p.bitrate, p.vbr, p.constrained_vbr, N, st.vbr_reservoir);
  int nbCompressedBytes = bud.nbCompressedBytes, nbAvailableBytes = bud.nbAvailableBytes, effectiveBytes = bud.effectiveBytes;
  const int32_t vbr_rate = bud.vbr_rate;
  int32_t equiv_rate = ((int32_t)OB_IMIN(p.max_bytes, 1275) * 8 * 50 << (3 - LM)) - (40 * C + 20) * ((400 >> LM) - 50);
  if (!(p.vbr && p.bitrate != -1) && p.bitrate != -1)
    equiv_rate = ((int32_t)OB_IMAX(2, OB_IMIN(OB_IMIN(p.max_bytes, 1275), (p.bitrate * N + 4 * 48000) / (8 * 48000))) * 8 * 50 << (3 - LM)) - (40 * C + 20) * ((400 >> LM) - 50);
  if (p.bitrate != -1) equiv_rate = OB_IMIN(equiv_rate, p.bitrate - (40 * C + 20) * ((400 >> LM) - 50));
  RangeEnc enc;
  re_init(enc, p.packets + (size_t)stream * p.pkt_stride, nbCompressedBytes);
  int32_t tell = 1;
  int32_t total_bits = nbCompressedBytes * 8;
  const int silence = fi.sample_max <= 1.f / (float)(1 << p.lsb_depth);
  re_bit_logp(enc, silence, 15);
  if (silence) {
    if (vbr_rate > 0) {
      effectiveBytes = nbCompressedBytes = OB_IMIN(nbCompressedBytes, 2);
      total_bits = nbCompressedBytes * 8;
      nbAvailableBytes = 2;
      enc.storage = nbCompressedBytes;
    }
    tell = nbCompressedBytes * 8;
    enc.nbits_total += tell - re_tell(enc);
  }
  int isTransient = fi.is_transient && !fi.patched;   // the time-domain decision; a patched frame turns transient after the energies
  const float tf_estimate = fi.tf_estimate;
  const int tf_chan = fi.tf_chan;
  const float toneishness = fi.toneishness;   // already limited to 1 - tf_estimate by the analysis kernel
  // the post-filter parameters chosen by run_prefilter in the analysis kernel (celt/celt_encoder.c:2040-2057)
  int status = 0;
  const int pf_on = fi.pf_on;
  int pitch_index = fi.pitch_index;
  // celt/celt_encoder.c:2043-2045 (read by compute_vbr's tonality boost only)
  const int pitch_change = (fi.gain1 > .4f || st.prefilter_gain > .4f) && (!(an && an->valid) || (double)an->tonality > .3) &&
                           ((double)pitch_index > 1.26 * st.prefilter_period || (double)pitch_index < .79 * st.prefilter_period);
  if (pf_on == 0) {
    if (tell + 16 <= total_bits) re_bit_logp(enc, 0, 1);
  } else {
    re_bit_logp(enc, 1, 1);
    pitch_index += 1;
    const int octave = ilog32(pitch_index) - 5;
    re_uint(enc, octave, 6);
    re_bits(enc, pitch_index - (16 << octave), 4 + octave);
    pitch_index -= 1;
    re_bits(enc, fi.qg, 3);
    re_icdf(enc, fi.pf_tapset, ob_tapset_icdf, 2);
  }
  int shortBlocks = 0, transient_got_disabled = 0;
  if (LM > 0 && re_tell(enc) + 3 <= total_bits) {
    if (isTransient) shortBlocks = M;
  } else {
    isTransient = 0;
    transient_got_disabled = 1;
  }
  if (fi.patched) {   // patch_transient_decision fired in the analysis kernel (needs the room for the transient flag, like the reference)
    if (!transient_got_disabled && !isTransient) { isTransient = 1; shortBlocks = M; }
    else status = -3;
  }
  if ((shortBlocks != 0) != (fi.short_blocks != 0)) status = -3;   // the analysis kernel transformed with the other block size
  float temporal_vbr;
  { // temporal VBR statistics (:2213-2228) keep spec_avg current even in CBR
    float follow = -10.0f, frame_avg = 0.f;
    // (computed on the energies of the first transform: for a patched frame those are the long-block ones, offset 0)
    const float offset = (shortBlocks && !fi.patched) ? .5f * LM : 0.f;
    const float *le = fi.patched ? bl2 : bandLogE;
#pragma unroll 1
    for (int i = start; i < end; i++) {
      follow = enc_fmax(follow - 1.0f, le[i] - offset);
      if (C == 2) follow = enc_fmax(follow, le[i + OB_NBANDS] - offset);
      frame_avg += follow;
    }
    frame_avg /= (end - start);
    temporal_vbr = frame_avg - st.spec_avg;
    temporal_vbr = enc_fmin(3.f, enc_fmax(-1.5f, temporal_vbr));
    st.spec_avg += .02f * temporal_vbr;
  }
  if (have_bl2)   // celt/celt_encoder.c:2087, :2243-2245: compensate the long-block energies for the block-size change
#pragma unroll 1
    for (int i = 0; i < C * OB_NBANDS; i++) bl2[i] += .5f * LM;
  if (LM > 0 && re_tell(enc) + 3 <= total_bits) re_bit_logp(enc, isTransient, 3);
  const EncDyn dyn = enc_dynalloc(bandLogE, have_bl2 ? bl2 : bandLogE, st.oldBandE, C, offsets, p.lsb_depth, isTransient, LM, effectiveBytes, fi.tone_freq, toneishness, p.vbr,
                                  p.constrained_vbr, importance, spread_weight, end, an);
  int tf_select = 0;
  if (effectiveBytes >= 15 * C && p.complexity >= 2 && toneishness < .98f) {   // enable_tf_analysis
    const int lambda = OB_IMAX(80, 20480 / effectiveBytes + 2);
    tf_select = enc_tf_analysis(isTransient, tf_res, lambda, Xg, N, LM, tf_estimate, tf_chan, importance, bx, by, end);
  } else {
#pragma unroll 1
    for (int i = 0; i < end; i++) tf_res[i] = isTransient;
  }
#pragma unroll 1
  for (int c = 0; c < C; c++)
#pragma unroll 1
    for (int i = start; i < end; i++)
      if (fabsf(bandLogE[i + c * OB_NBANDS] - st.oldBandE[i + c * OB_NBANDS]) < 2.f) bandLogE[i + c * OB_NBANDS] -= 0.25f * st.energyError[i + c * OB_NBANDS];
  { // quant_coarse_energy (celt/quant_bands.c:256-365)
    int two_pass = p.complexity >= 4;
    int intra = p.force_intra || (!two_pass && st.delayedIntra > 2 * C * (end - start) && nbAvailableBytes > (end - start) * C);
    const int32_t intra_bias = (int32_t)((total_bits * st.delayedIntra * p.loss_rate) / (C * 512));
    float new_distortion = 0.f;
#pragma unroll 1
    for (int c = 0; c < C; c++)
#pragma unroll 1
      for (int i = start; i < end; i++) {
        const float d = bandLogE[i + c * OB_NBANDS] - st.oldBandE[i + c * OB_NBANDS];
        new_distortion = new_distortion + d * d;
      }
    new_distortion = enc_fmin(200.f, new_distortion);
    const int32_t t0 = re_tell(enc);
    if (t0 + 3 > total_bits) two_pass = intra = 0;
    const float max_decay = enc_fmin(16.f, .125f * nbAvailableBytes);   // end - start > 10
    const float pc = LM == 0 ? 29440 / 32768.f : LM == 1 ? 26112 / 32768.f : LM == 2 ? 21248 / 32768.f : 16384 / 32768.f;
    if (two_pass || intra) {
      // by/bx double as oldEBands_intra / error_intra (42 floats each); a handful of packet bytes is set aside for the retry
      float *old_intra = by, *err_intra = by + 2 * OB_NBANDS;
      const RangeEnc start_state = enc;
#pragma unroll 1
      for (int i = 0; i < C * OB_NBANDS; i++) old_intra[i] = st.oldBandE[i];
      const int badness1 = enc_coarse_impl(enc, bandLogE, old_intra, total_bits, t0, ob_e_prob + (LM * 2 + 1) * 42, err_intra, C, LM, 1, max_decay, end);
      if (!intra) {
        unsigned char saved[160];
        const int32_t tell_intra = re_tell_frac(enc);
        const RangeEnc intra_state = enc;
        const uint32_t nb0 = start_state.offs, nb1 = intra_state.offs;
        if (nb1 - nb0 > 160) status = -3;
#pragma unroll 1
        for (uint32_t i = 0; i < nb1 - nb0 && i < 160; i++) saved[i] = enc.buf[nb0 + i];
        enc = start_state;
        const int badness2 = enc_coarse_impl(enc, bandLogE, st.oldBandE, total_bits, t0, ob_e_prob + (LM * 2 + 0) * 42, error, C, LM, 0, max_decay, end);
        if (two_pass && (badness1 < badness2 || (badness1 == badness2 && ((int32_t)re_tell_frac(enc)) + intra_bias > tell_intra))) {
          enc = intra_state;
#pragma unroll 1
          for (uint32_t i = 0; i < nb1 - nb0 && i < 160; i++) enc.buf[nb0 + i] = saved[i];
#pragma unroll 1
          for (int i = 0; i < C * OB_NBANDS; i++) { st.oldBandE[i] = old_intra[i]; error[i] = err_intra[i]; }
          intra = 1;
        }
      } else {
#pragma unroll 1
        for (int i = 0; i < C * OB_NBANDS; i++) { st.oldBandE[i] = old_intra[i]; error[i] = err_intra[i]; }
      }
    } else {
      enc_coarse_impl(enc, bandLogE, st.oldBandE, total_bits, t0, ob_e_prob + (LM * 2 + 0) * 42, error, C, LM, 0, max_decay, end);
    }
    if (intra) st.delayedIntra = new_distortion;
    else st.delayedIntra = (pc * pc) * st.delayedIntra + new_distortion;
  }
  { // tf_encode (celt/celt_encoder.c:824-862)
    int curr = 0, tf_changed = 0, logp = isTransient ? 2 : 4;
    uint32_t budget = enc.storage * 8, t = re_tell(enc);
    const int tf_select_rsv = LM > 0 && t + logp + 1 <= budget;
    budget -= tf_select_rsv;
#pragma unroll 1
    for (int i = start; i < end; i++) {
      if (t + logp <= budget) {
        re_bit_logp(enc, tf_res[i] ^ curr, logp);
        t = re_tell(enc);
        curr = tf_res[i];
        tf_changed |= curr;
      } else tf_res[i] = curr;
      logp = isTransient ? 4 : 5;
    }
    if (tf_select_rsv && ob_tf_select[LM * 8 + 4 * isTransient + 0 + tf_changed] != ob_tf_select[LM * 8 + 4 * isTransient + 2 + tf_changed])
      re_bit_logp(enc, tf_select, 1);
    else tf_select = 0;
#pragma unroll 1
    for (int i = start; i < end; i++) tf_res[i] = ob_tf_select[LM * 8 + 4 * isTransient + 2 * tf_select + tf_res[i]];
  }
  if (re_tell(enc) + 4 <= total_bits) {
    if (shortBlocks || p.complexity < 3 || nbAvailableBytes < 10 * C) st.spread_decision = p.complexity == 0 ? 0 : 2;   // SPREAD_NONE : SPREAD_NORMAL
    else st.spread_decision = enc_spreading_decision(Xg, st, pf_on && !shortBlocks, C, M, spread_weight, end);
    re_icdf(enc, st.spread_decision, ob_spread_icdf, 5);
  } else st.spread_decision = 2;
#pragma unroll 1
  for (int i = 0; i < OB_NBANDS; i++) cap[i] = (ob_cache_caps[OB_NBANDS * (2 * LM + C - 1) + i] + 64) * C * ((ob_ebands[i + 1] - ob_ebands[i]) << LM) >> 2;
  int dynalloc_logp = 6;
  total_bits <<= OB_BITRES;
  int32_t total_boost = 0;
  tell = re_tell_frac(enc);
#pragma unroll 1
  for (int i = start; i < end; i++) {
    const int width = C * (ob_ebands[i + 1] - ob_ebands[i]) << LM;
    const int quanta = OB_IMIN(width << OB_BITRES, OB_IMAX(6 << OB_BITRES, width));
    int dynalloc_loop_logp = dynalloc_logp, boost = 0, j;
    for (j = 0; tell + (dynalloc_loop_logp << OB_BITRES) < total_bits - total_boost && boost < cap[i]; j++) {
      const int flag = j < offsets[i];
      re_bit_logp(enc, flag, dynalloc_loop_logp);
      tell = re_tell_frac(enc);
      if (!flag) break;
      boost += quanta;
      total_boost += quanta;
      dynalloc_loop_logp = 1;
    }
    if (j) dynalloc_logp = OB_IMAX(2, dynalloc_logp - 1);
    offsets[i] = boost;
  }
  int dual_stereo = 0;
  if (C == 2) {
    if (LM != 0) {   // stereo_analysis (celt/celt_encoder.c:954-987)
      float sumLR = 1e-15f, sumMS = 1e-15f;
#pragma unroll 1
      for (int j = 0; j < ob_ebands[13] << LM; j++) {
        const float L = Xg[j], R = Xg[N + j], Mm = L + R, S = L - R;
        sumLR = sumLR + (fabsf(L) + fabsf(R));
        sumMS = sumMS + (fabsf(Mm) + fabsf(S));
      }
      sumMS = 0.707107f * sumMS;
      const int thetas = LM <= 1 ? 5 : 13;
      dual_stereo = ((ob_ebands[13] << (LM + 1)) + thetas) * sumMS > (ob_ebands[13] << (LM + 1)) * sumLR;
    }
    { // hysteresis_decision on the intensity threshold table (celt/celt_encoder.c:2394-2408, celt/bands.c:46-59)
      const float th[21] = {1, 2, 3, 4, 5, 6, 7, 8, 16, 24, 36, 44, 50, 56, 62, 67, 72, 79, 88, 106, 134};
      const float hy[21] = {1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 3, 3, 4, 5, 6, 8, 8};
      const float val = (float)(equiv_rate / 1000);
      const int prev = st.intensity;
      int i;
      for (i = 0; i < 21; i++) if (val < th[i]) break;
      if (i > prev && val < th[prev] + hy[prev]) i = prev;
      if (i < prev && val > th[prev - 1] - hy[prev - 1]) i = prev;
      st.intensity = OB_IMIN(end, OB_IMAX(start, i));
    }
  }
  int alloc_trim = 5;
  if (tell + (6 << OB_BITRES) <= total_bits - total_boost) {   // alloc_trim_analysis (celt/celt_encoder.c:865-952)
    float trim = 5.f, diff = 0.f;
    if (equiv_rate < 64000) trim = 4.f;
    else if (equiv_rate < 80000) trim = 4.f + (1.f / 16.f) * ((equiv_rate - 64000) >> 10);
    if (C == 2) {
      float sum = 0.f;
#pragma unroll 1
      for (int i = 0; i < 8; i++)
        sum = sum + enc_dot(&Xg[ob_ebands[i] << LM], &Xg[N + (ob_ebands[i] << LM)], (ob_ebands[i + 1] - ob_ebands[i]) << LM);
      sum = (1.f / 8) * sum;
      sum = enc_fmin(1.f, fabsf(sum));
      float minXC = sum;
#pragma unroll 1
      for (int i = 8; i < st.intensity; i++)
        minXC = enc_fmin(minXC, fabsf(enc_dot(&Xg[ob_ebands[i] << LM], &Xg[N + (ob_ebands[i] << LM)], (ob_ebands[i + 1] - ob_ebands[i]) << LM)));
      minXC = enc_fmin(1.f, fabsf(minXC));
      const float logXC = enc_log2(1.001f - sum * sum);
      const float logXC2 = enc_fmax(.5f * logXC, enc_log2(1.001f - minXC * minXC));
      trim += enc_fmax(-4.f, .75f * logXC);
      st.stereo_saving = enc_fmin(st.stereo_saving + 0.25f, -(.5f * logXC2));
    }
#pragma unroll 1
    for (int c = 0; c < C; c++)
#pragma unroll 1
      for (int i = 0; i < end - 1; i++) diff += bandLogE[i + c * OB_NBANDS] * (int32_t)(2 + 2 * i - end);
    diff /= C * (end - 1);
    trim -= enc_fmax(-2.f, enc_fmin(2.f, (diff + 1.f) / 6));
    trim -= 2 * tf_estimate;
    if (an && an->valid) trim -= enc_fmax(-2.f, enc_fmin(2.f, (float)(2.f * (an->tonality_slope + .05f))));   // celt/celt_encoder.c:935-939
    alloc_trim = OB_IMAX(0, OB_IMIN(10, (int)floor((double)(.5f + trim))));
    re_icdf(enc, alloc_trim, ob_trim_icdf, 7);
    tell = re_tell_frac(enc);
  }
  if (vbr_rate > 0) {   // :2432-2522
    const int32_t min_allowed = ((tell + total_boost + (1 << (OB_BITRES + 3)) - 1) >> (OB_BITRES + 3)) + 2;
    const int lm_diff = 3 - LM;
    nbCompressedBytes = OB_IMIN(nbCompressedBytes, 1275 >> (3 - LM));
    int32_t base_target = vbr_rate - ((40 * C + 20) << OB_BITRES);
    if (p.constrained_vbr) base_target += (st.vbr_offset >> lm_diff);
    int32_t target = enc_compute_vbr(base_target, LM, equiv_rate, st.lastCodedBands, C, st.intensity, p.constrained_vbr, st.stereo_saving, dyn.tot_boost,
                                     tf_estimate, dyn.maxDepth, temporal_vbr, an, pitch_change);
    target = target + tell;
    nbAvailableBytes = (target + (1 << (OB_BITRES + 2))) >> (OB_BITRES + 3);
    nbAvailableBytes = OB_IMAX(min_allowed, nbAvailableBytes);
    nbAvailableBytes = OB_IMIN(nbCompressedBytes, nbAvailableBytes);
    int32_t delta = target - vbr_rate;
    target = nbAvailableBytes << (OB_BITRES + 3);
    if (silence) {
      nbAvailableBytes = 2;
      target = 2 * 8 << OB_BITRES;
      delta = 0;
    }
    float alpha;
    if (st.vbr_count < 970) {
      st.vbr_count++;
      alpha = 1.0f / (st.vbr_count + 20);
    } else alpha = .001f;
    if (p.constrained_vbr) {
      st.vbr_reservoir += target - vbr_rate;
      st.vbr_drift += (int32_t)(alpha * ((delta * (1 << lm_diff)) - st.vbr_offset - st.vbr_drift));
      st.vbr_offset = -st.vbr_drift;
      if (st.vbr_reservoir < 0) {
        const int adjust = (-st.vbr_reservoir) / (8 << OB_BITRES);
        nbAvailableBytes += silence ? 0 : adjust;
        st.vbr_reservoir = 0;
      }
    }
    nbCompressedBytes = OB_IMIN(nbCompressedBytes, nbAvailableBytes);
    enc.storage = nbCompressedBytes;   // ec_enc_shrink: no raw bits have been written yet
  }
  int32_t bits = (((int32_t)nbCompressedBytes * 8) << OB_BITRES) - (int32_t)re_tell_frac(enc) - 1;
  const int anti_collapse_rsv = isTransient && LM >= 2 && bits >= ((LM + 2) << OB_BITRES) ? (1 << OB_BITRES) : 0;
  bits -= anti_collapse_rsv;
  int signalBandwidth = end - 1;
  if (an && an->valid) {   // celt/celt_encoder.c:2610-2624
    const int min_bandwidth = equiv_rate < (int32_t)32000 * C ? 13 : equiv_rate < (int32_t)48000 * C ? 16 : equiv_rate < (int32_t)60000 * C ? 18
                              : equiv_rate < (int32_t)80000 * C ? 19 : 20;
    signalBandwidth = OB_IMAX(an->bandwidth, min_bandwidth);
  }
  const EncAlloc al = enc_compute_allocation(enc, offsets, cap, alloc_trim, st.intensity, dual_stereo, bits, pulses, fine_quant, fine_priority, C, LM,
                                             st.lastCodedBands, signalBandwidth, end);
  st.intensity = al.intensity;
  dual_stereo = al.dual_stereo;
  const int codedBands = al.codedBands;
  if (st.lastCodedBands) st.lastCodedBands = OB_IMIN(st.lastCodedBands + 1, OB_IMAX(st.lastCodedBands - 1, codedBands));
  else st.lastCodedBands = codedBands;
  // quant_fine_energy (celt/quant_bands.c:367-410)
#pragma unroll 1
  for (int i = start; i < end; i++) {
    const int extra = 1 << fine_quant[i];
    if (fine_quant[i] <= 0) continue;
    if (re_tell(enc) + C * fine_quant[i] > (int32_t)enc.storage * 8) continue;
#pragma unroll 1
    for (int c = 0; c < C; c++) {
      int q2 = (int)floor((double)((error[i + c * OB_NBANDS] * 1 + .5f) * extra));
      if (q2 > extra - 1) q2 = extra - 1;
      if (q2 < 0) q2 = 0;
      re_bits(enc, q2, fine_quant[i]);
      float offset = (q2 + .5f) * (1 << (14 - fine_quant[i])) * (1.f / 16384) - .5f;
      offset *= (1 << 14) * (1.f / 16384);
      st.oldBandE[i + c * OB_NBANDS] += offset;
      error[i + c * OB_NBANDS] -= offset;
    }
  }
  // quant_all_bands, encode side (celt/bands.c:1589-1923): no resynthesis, so no folding state is carried
  if (RDO && C == 2 && !dual_stereo && p.complexity >= 8) {   // theta_rdo (celt/bands.c:1618)
    if constexpr (RDO) {
      RdoScratch w;
      const int r = rdo_all_bands<SYNC>(enc, Xg, bandE, pulses, tf_res, shortBlocks, st.spread_decision, st.intensity,
                                  nbCompressedBytes * (8 << OB_BITRES) - anti_collapse_rsv, al.balance, LM, codedBands, st.rng, p.disable_inv, bx, by, tmp, w, end);
      if (r) status = r;
    }
  } else {
    EncBand ctx;
    ctx.ec = &enc; ctx.bandE = bandE; ctx.intensity = st.intensity; ctx.spread = st.spread_decision; ctx.disable_inv = p.disable_inv;
    ctx.theta_round = 0; ctx.seed = 0;
    const int B = shortBlocks ? M : 1;
    ctx.avoid_split_noise = B > 1;
    int32_t balance = al.balance;
    const int32_t tb = nbCompressedBytes * (8 << OB_BITRES) - anti_collapse_rsv;
#pragma unroll 1
    for (int i = start; i < (SYNC ? OB_NBANDS : end); i++) {
      if (SYNC) {   // (see rdo_all_bands)
        bsync();
        if (i >= end) continue;
      }
      const int off = M * ob_ebands[i], Nb = M * ob_ebands[i + 1] - off;
      ctx.band = i;
      const int32_t t = re_tell_frac(enc);
      if (i != start) balance -= t;
      const int32_t remaining_bits = tb - t - 1;
      ctx.remaining_bits = remaining_bits;
      int b = 0;
      if (i <= codedBands - 1) {
        const int32_t curr_balance = balance / OB_IMIN(3, codedBands - i);
        b = OB_IMAX(0, OB_IMIN(16383, OB_IMIN(remaining_bits + 1, pulses[i] + curr_balance)));
      }
      ctx.tf_change = tf_res[i];
      if (dual_stereo && i == st.intensity) dual_stereo = 0;
#pragma unroll 1
      for (int j = 0; j < Nb; j++) bx[j] = Xg[off + j];
      if (C == 2)
#pragma unroll 1
        for (int j = 0; j < Nb; j++) by[j] = Xg[N + off + j];
      if (C == 2 && !dual_stereo) enc_band_stereo(ctx, bx, by, tmp, Nb, b, B, LM);
      else {   // mono, or the two halves of a dual-stereo band: one call site
#pragma unroll 1
        for (int h = 0; h < (dual_stereo ? 2 : 1); h++) enc_band_mono(ctx, h ? by : bx, tmp, Nb, dual_stereo ? b / 2 : b, B, LM);
      }
      balance += pulses[i] + t;
      ctx.avoid_split_noise = 0;
    }
  }
  if (anti_collapse_rsv > 0) re_bits(enc, st.consec_transient < 2, 1);
  { // quant_energy_finalise (celt/quant_bands.c:412-443)
    int bits_left = nbCompressedBytes * 8 - re_tell(enc);
#pragma unroll 1
    for (int prio = 0; prio < 2; prio++)
#pragma unroll 1
      for (int i = start; i < end && bits_left >= C; i++) {
        if (fine_quant[i] >= 8 || fine_priority[i] != prio) continue;
#pragma unroll 1
        for (int c = 0; c < C; c++) {
          const int q2 = error[i + c * OB_NBANDS] < 0 ? 0 : 1;
          re_bits(enc, q2, 1);
          const float offset = (q2 - .5f) * (1 << (14 - fine_quant[i] - 1)) * (1.f / 16384);
          st.oldBandE[i + c * OB_NBANDS] += offset;
          error[i + c * OB_NBANDS] -= offset;
          bits_left--;
        }
      }
  }
#pragma unroll 1
  for (int i = 0; i < 2 * OB_NBANDS; i++) st.energyError[i] = 0.f;
#pragma unroll 1
  for (int c = 0; c < C; c++)
#pragma unroll 1
    for (int i = start; i < end; i++) st.energyError[i + c * OB_NBANDS] = enc_fmax(-0.5f, enc_fmin(0.5f, error[i + c * OB_NBANDS]));
  if (silence)
#pragma unroll 1
    for (int i = 0; i < C * OB_NBANDS; i++) st.oldBandE[i] = -28.f;
  st.prefilter_period = pitch_index;
  st.prefilter_gain = fi.gain1;
  st.prefilter_tapset = fi.pf_tapset;
  const int CC = p.CC ? p.CC : C;
  if (CC == 2 && C == 1)   // celt/celt_encoder.c:2780-2782
#pragma unroll 1
    for (int i = 0; i < OB_NBANDS; i++) st.oldBandE[OB_NBANDS + i] = st.oldBandE[i];
  if (!isTransient) {
#pragma unroll 1
    for (int i = 0; i < CC * OB_NBANDS; i++) { st.oldLogE2[i] = st.oldLogE[i]; st.oldLogE[i] = st.oldBandE[i]; }
  } else {
#pragma unroll 1
    for (int i = 0; i < CC * OB_NBANDS; i++) st.oldLogE[i] = enc_fmin(st.oldLogE[i], st.oldBandE[i]);
  }
  if (end < OB_NBANDS) {  // in case the end band were to change (celt/celt_encoder.c:2795-2807)
#pragma unroll 1
    for (int c = 0; c < CC; c++)
#pragma unroll 1
      for (int i = end; i < OB_NBANDS; i++) { st.oldBandE[c * OB_NBANDS + i] = 0.f; st.oldLogE[c * OB_NBANDS + i] = st.oldLogE2[c * OB_NBANDS + i] = -28.f; }
  }
  if (isTransient || transient_got_disabled) st.consec_transient++;
  else st.consec_transient = 0;
  st.rng = enc.rng;
  re_done(enc);
  p.nbytes[stream] = status ? status : (enc.error ? -3 : nbCompressedBytes + p.toc_bytes);
}

}  // namespace ob
