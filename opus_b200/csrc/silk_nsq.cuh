// silk_nsq.cuh — SILK noise-shaping quantizers, batched over streams.
//   silk_NSQ_c          silk/NSQ.c:76-437         -> nsq_plain_thread : one THREAD per stream
//   silk_NSQ_del_dec_c  silk/NSQ_del_dec.c:113-730 -> nsq_del_dec_warp : FOUR LANES per stream, lane k = delayed-decision state k
//
// Both are sample-serial feedback loops, so the parallel axes are the batch and (for the delayed-decision trellis) the
// up-to-4 surviving states.  What differs from the scalar layout of the reference:
//   * a state's short-term history (16 taps) and its warped noise-shaping filter (24 taps) are REGISTERS of its lane: the
//     history is a shift register, all tap loops are fully unrolled, and every silk_SMULWB/SMLAWB is ONE IMAD.HI against a
//     coefficient pre-shifted by 16 bits ((a*b)>>16 == mulhi(a, b<<16) exactly for 16-bit b);
//   * the five 40-deep decision rings (RandState, Q, Xq, Pred, Shape) stay put in shared memory, [ring][slot][lane]; when the
//     trellis replaces a state by a copy of another the reference memcpy's ~1.3 KB — here the lane only takes over a 2-bit-per-
//     slot MAP (3 registers) saying which physical column holds its entry for each time slot, plus 40+ register shuffles;
//   * winner / worst / best searches are two-step butterflies over the 4 lanes with the reference's tie rules (first index wins).
// Results are bit-identical to the reference (tests/test_emu_nsq.py on the host emulator, tests/test_gpu_ops.py on the GPU).
#pragma once
#include "silk_ops.cuh"

namespace ob {

struct ObNsqCfg { int32_t fs_khz, nb_subfr, predictLPCOrder, shapingLPCOrder, warping_Q16, nStates; };
// = OpusB200NsqFrame (include/opus_b200.h)
struct ObNsqFrame {
  int32_t signalType, quantOffsetType, NLSFInterpCoef_Q2, Seed;
  int32_t HarmShapeGain_Q14[4], Tilt_Q14[4], LF_shp_Q14[4], Gains_Q16[4], pitchL[4];
  int32_t Lambda_Q10, LTP_scale_Q14;
  int16_t PredCoef_Q12[32], LTPCoef_Q14[20], AR_Q13[96];
};
// per-stream carried state in HBM (silk_nsq_state, silk/structs.h:56-69) + the per-call LTP scratch
struct ObNsqState {
  int32_t sLTP_shp_Q14[640];
  int32_t sLTP_Q15[640];  // scratch, rebuilt every frame
  int16_t xq[640];
  int32_t sLPC_Q14[16];
  int32_t sAR2_Q14[24];
  int32_t sLF_AR_shp_Q14, sDiff_shp_Q14, lagPrev, rand_seed, prev_gain_Q16, pad[3];
};
struct SilkNsqParams {
  ObNsqState *state;     // [n_streams]
  ObNsqFrame *frames;    // [n_streams]  (Seed is updated by the delayed-decision variant)
  const int16_t *x16;    // [n_streams][frame_length]
  int8_t *pulses;        // [n_streams][frame_length]
  int n_streams;
  ObNsqCfg cfg;
};

constexpr int kNsqDelay = 40;
constexpr int kNsqRingWords = 5 * kNsqDelay * 32;            // per warp
constexpr int kNsqWarpSmemWords = kNsqRingWords + kNsqDelay * 8;  // 26.9 KB: 8 warps per SM
enum { RING_RAND = 0, RING_Q = 1, RING_XQ = 2, RING_PRED = 3, RING_SHAPE = 4 };

OB_DEV int32_t nsq_addw(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
OB_DEV int32_t nsq_subw(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
OB_DEV int32_t nsq_shl(int32_t a, int s) { return (int32_t)((uint32_t)a << s); }
OB_DEV int32_t nsq_hi16(int32_t c) { return (int32_t)((uint32_t)(int32_t)(int16_t)c << 16); }  // low half-word, moved up
OB_DEV int32_t nsq_smulww(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * b) >> 16); }
OB_DEV int32_t nsq_smulbb(int32_t a, int32_t b) { return (int32_t)(int16_t)a * (int32_t)(int16_t)b; }
OB_DEV int32_t nsq_add_sat(int32_t a, int32_t b) {
  const int64_t s = (int64_t)a + b;
  return s > 2147483647LL ? 2147483647 : (s < -2147483648LL ? (int32_t)(-2147483647 - 1) : (int32_t)s);
}
OB_DEV int32_t nsq_sub_sat(int32_t a, int32_t b) {
  const int64_t s = (int64_t)a - b;
  return s > 2147483647LL ? 2147483647 : (s < -2147483648LL ? (int32_t)(-2147483647 - 1) : (int32_t)s);
}
OB_DEV int32_t nsq_lcg(int32_t seed) { return (int32_t)(907633515u + (uint32_t)seed * 196314165u); }
OB_DEV int nsq_clz(int32_t x) { return x ? clz32((uint32_t)x) : 32; }
OB_DEV int32_t nsq_min(int32_t a, int32_t b) { return a < b ? a : b; }
OB_DEV int32_t nsq_max(int32_t a, int32_t b) { return a > b ? a : b; }
OB_DEV int32_t nsq_abs(int32_t a) { return a > 0 ? a : -a; }
OB_DEV int32_t nsq_lshift_sat(int32_t a, int s) {
  const int32_t lo = (int32_t)(-2147483647 - 1) >> s, hi = 2147483647 >> s;
  return nsq_shl(a < lo ? lo : (a > hi ? hi : a), s);
}
// silk/Inlines.h:93-136
OB_DEV int32_t nsq_div32_varq(int32_t a32, int32_t b32, int Qres) {
  const int a_headrm = nsq_clz(nsq_abs(a32)) - 1, b_headrm = nsq_clz(nsq_abs(b32)) - 1;
  int32_t a_nrm = nsq_shl(a32, a_headrm);
  const int32_t b_nrm = nsq_shl(b32, b_headrm);
  const int32_t b_inv = (2147483647 >> 2) / (b_nrm >> 16);
  int32_t result = silk_smulwb(a_nrm, b_inv);
  a_nrm = nsq_subw(a_nrm, nsq_shl(mulhi32(b_nrm, result), 3));
  result = nsq_addw(result, silk_smulwb(a_nrm, b_inv));
  const int lshift = 29 + a_headrm - b_headrm - Qres;
  if (lshift < 0) return nsq_lshift_sat(result, -lshift);
  return lshift < 32 ? result >> lshift : 0;
}
// silk/Inlines.h:139-178
OB_DEV int32_t nsq_inverse32_varq(int32_t b32, int Qres) {
  const int b_headrm = nsq_clz(nsq_abs(b32)) - 1;
  const int32_t b_nrm = nsq_shl(b32, b_headrm);
  const int32_t b_inv = (2147483647 >> 2) / (b_nrm >> 16);
  int32_t result = nsq_shl(b_inv, 16);
  const int32_t err_Q32 = nsq_shl(((int32_t)1 << 29) - silk_smulwb(b_nrm, b_inv), 3);
  result = nsq_addw(result, nsq_smulww(err_Q32, b_inv));
  const int lshift = 61 - b_headrm - Qres;
  if (lshift <= 0) return nsq_lshift_sat(result, -lshift);
  return lshift < 32 ? result >> lshift : 0;
}

// The two quantization candidates around r (silk/NSQ.c:277-331 == NSQ_del_dec.c:456-505): q1 < q2 one level apart,
// with their rate terms.  Branch-free selects on the sign class of q1_Q0.
OB_DEV void nsq_candidates(int32_t r_Q10, int32_t offset_Q10, int32_t Lambda_Q10, int32_t &q1, int32_t &q2, int32_t &rd1, int32_t &rd2) {
  int32_t q1_Q10 = r_Q10 - offset_Q10;
  int32_t q1_Q0 = q1_Q10 >> 10;
  if (Lambda_Q10 > 2048) {
    const int32_t rdo = Lambda_Q10 / 2 - 512;
    q1_Q0 = q1_Q10 > rdo ? (q1_Q10 - rdo) >> 10 : (q1_Q10 < -rdo ? (q1_Q10 + rdo) >> 10 : (q1_Q10 < 0 ? -1 : 0));
  }
  if (q1_Q0 > 0) {
    q1 = (q1_Q0 << 10) - 80 + offset_Q10;
    q2 = q1 + 1024;
    rd1 = nsq_smulbb(q1, Lambda_Q10);
    rd2 = nsq_smulbb(q2, Lambda_Q10);
  } else if (q1_Q0 == 0) {
    q1 = offset_Q10;
    q2 = q1 + (1024 - 80);
    rd1 = nsq_smulbb(q1, Lambda_Q10);
    rd2 = nsq_smulbb(q2, Lambda_Q10);
  } else if (q1_Q0 == -1) {
    q2 = offset_Q10;
    q1 = q2 - (1024 - 80);
    rd1 = nsq_smulbb(-q1, Lambda_Q10);
    rd2 = nsq_smulbb(q2, Lambda_Q10);
  } else {
    q1 = nsq_shl(q1_Q0, 10) + 80 + offset_Q10;
    q2 = q1 + 1024;
    rd1 = nsq_smulbb(-q1, Lambda_Q10);
    rd2 = nsq_smulbb(-q2, Lambda_Q10);
  }
}
OB_DEV int32_t nsq_quant_offset(int signalType, int quantOffsetType) {  // silk/tables_other.c:77
  return (signalType >> 1) ? (quantOffsetType ? 100 : 32) : (quantOffsetType ? 240 : 100);
}

// ================================================================================================================
// delayed-decision quantizer: 8 streams per warp, lane = 4*g + k
// ================================================================================================================
struct NsqMap {  // which physical column (lane of the group) holds this state's ring entry for each of the 40 time slots
  uint32_t w0, w1, w2;
  OB_MDEV int get(int t) const { const uint32_t w = t < 16 ? w0 : (t < 32 ? w1 : w2); return (int)((w >> ((t & 15) * 2)) & 3u); }
  OB_MDEV void set(int t, int v) {
    const int sh = (t & 15) * 2;
    const uint32_t m = ~(3u << sh), b = (uint32_t)v << sh;
    if (t < 16) w0 = (w0 & m) | b;
    else if (t < 32) w1 = (w1 & m) | b;
    else w2 = (w2 & m) | b;
  }
};

// argmin / argmax over the 4 lanes of a group; ties go to the lower index (reference scans with strict compares)
OB_DEV void nsq_grp_argmin(int32_t &v, int &idx) {
#pragma unroll
  for (int m = 1; m <= 2; m <<= 1) {
    const int32_t ov = shfl_xor(v, m);
    const int oi = shfl_xor(idx, m);
    if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
}
OB_DEV void nsq_grp_argmax(int32_t &v, int &idx) {
#pragma unroll
  for (int m = 1; m <= 2; m <<= 1) {
    const int32_t ov = shfl_xor(v, m);
    const int oi = shfl_xor(idx, m);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
}

// ORDER = shapingLPCOrder when it is one of the reference's settings (12, 14, 16, 20, 24: silk/control_codec.c:316-386),
// 0 = read it from the configuration (sections past the order are then skipped with warp-uniform selects).
template <int ORDER>
OB_DEV void nsq_del_dec_warp(const SilkNsqParams &p, int stream0, int32_t *smem) {
  const int lane = lane_id(), g = lane >> 2, k = lane & 3, gbase = lane & ~3;
  const bool live = stream0 + g < p.n_streams;
  const int stream = live ? stream0 + g : p.n_streams - 1;  // dead groups shadow the last stream read-only
  const ObNsqCfg c = p.cfg;
  const int subfr_length = 5 * c.fs_khz, ltp_mem = 20 * c.fs_khz, frame_length = subfr_length * c.nb_subfr;
  const bool act = k < c.nStates;
  ObNsqState &S = p.state[stream];
  ObNsqFrame &F = p.frames[stream];
  const int16_t *x16 = p.x16 + (size_t)stream * frame_length;
  int8_t *pulses = p.pulses + (size_t)stream * frame_length;
  int32_t *ring = smem;                         // [5][40][32]
  int32_t *dgain = smem + kNsqRingWords;        // [40][8]
#define RING(a, t, col) ring[((a) * kNsqDelay + (t)) * 32 + (col)]

  // ---- per-state registers ----
  int32_t sLPC[16], sAR2[24];  // sLPC[0] = newest
  int32_t LF_AR = S.sLF_AR_shp_Q14, Diff = S.sDiff_shp_Q14;
  int32_t Seed = (k + F.Seed) & 3, SeedInit = Seed, RD = 0;
  NsqMap map;
  map.w0 = map.w1 = map.w2 = 0x55555555u * (uint32_t)k;  // identity: every slot -> own column
#pragma unroll
  for (int j = 0; j < 16; j++) sLPC[j] = S.sLPC_Q14[15 - j];
#pragma unroll
  for (int j = 0; j < 24; j++) sAR2[j] = S.sAR2_Q14[j];
#pragma unroll 1
  for (int a = 0; a < 5; a++)
#pragma unroll 1
    for (int t = 0; t < kNsqDelay; t++) RING(a, t, lane) = 0;
  RING(RING_SHAPE, 0, lane) = S.sLTP_shp_Q14[ltp_mem - 1];

  const int signalType = F.signalType;
  const bool voiced = signalType == 2;
  const int32_t offset_Q10 = nsq_quant_offset(signalType, F.quantOffsetType);
  const int32_t Lambda_Q10 = F.Lambda_Q10;
  const int interp = F.NLSFInterpCoef_Q2 == 4 ? 0 : 1;
  int lag = S.lagPrev;
  int decisionDelay = kNsqDelay < subfr_length ? kNsqDelay : subfr_length;
  if (voiced) {
#pragma unroll 1
    for (int s = 0; s < c.nb_subfr; s++) decisionDelay = nsq_min(decisionDelay, F.pitchL[s] - 3);
  } else if (lag > 0) decisionDelay = nsq_min(decisionDelay, lag - 3);
  int smpl = 0, subfr = 0;
  int shp_idx = ltp_mem, ltp_idx = ltp_mem;  // NSQ->sLTP_shp_buf_idx, NSQ->sLTP_buf_idx
  int32_t prev_gain_Q16 = S.prev_gain_Q16;
  wsync();

#pragma unroll 1
  for (int sf = 0; sf < c.nb_subfr; sf++) {
    const int16_t *A_Q12 = &F.PredCoef_Q12[((sf >> 1) | (1 - interp)) * 16];
    const int32_t Gain_Q16 = F.Gains_Q16[sf];
    const bool rewhite = voiced && (sf & (3 - (interp << 1))) == 0;
    if (voiced) lag = F.pitchL[sf];
    if (sf == 2) {
      // ---- reset delayed decisions: flush the winner (NSQ_del_dec.c:211-241).  The searches and broadcasts are
      //      warp-uniform; only their effects are predicated on this stream re-whitening here. ----
      int32_t v = act ? RD : 2147483647;
      int w = act ? k : 4 + k;
      nsq_grp_argmin(v, w);
      NsqMap wm;
      wm.w0 = shfl(map.w0, gbase + w); wm.w1 = shfl(map.w1, gbase + w); wm.w2 = shfl(map.w2, gbase + w);
      if (rewhite) {
        if (k != w) RD += 2147483647 >> 4;
        const int32_t g1 = F.Gains_Q16[1];
#pragma unroll 1
        for (int i = k; i < decisionDelay; i += 4) {  // each lane flushes every 4th sample of the window
          const int t = (smpl + decisionDelay - 1 - i) % kNsqDelay;
          const int col = gbase + wm.get(t);
          const int o = sf * subfr_length + i - decisionDelay;
          if (live) {
            pulses[o] = (int8_t)silk_rshift_round(RING(RING_Q, t, col), 10);
            S.xq[ltp_mem + o] = silk_sat16(silk_rshift_round(nsq_smulww(RING(RING_XQ, t, col), g1), 14));
            S.sLTP_shp_Q14[shp_idx - decisionDelay + i] = RING(RING_SHAPE, t, col);
          }
        }
        subfr = 0;
      }
      wsync();
    }
    if (rewhite) ltp_idx = ltp_mem;
    // ---- scale states (NSQ_del_dec.c:655-730), the 4 lanes of a group share the per-stream loops ----
    int32_t inv_gain_Q31 = nsq_inverse32_varq(Gain_Q16 > 1 ? Gain_Q16 : 1, 47);
    const int32_t inv_gain_Q26 = silk_rshift_round(inv_gain_Q31, 5);
    if (rewhite) {
      // re-whitening (silk_LPC_analysis_filter over xq, NSQ_del_dec.c:243-248) fused with its rescale into sLTP_Q15
      if (sf == 0) inv_gain_Q31 = nsq_shl(silk_smulwb(inv_gain_Q31, F.LTP_scale_Q14), 2);
      // Each lane filters one contiguous quarter of the lag+2 samples with the 16 previous inputs in a register window.
      const int16_t *xin = &S.xq[sf * subfr_length];  // index n here = xq[n + sf*subfr_length]
      int32_t A[16], w[16];
#pragma unroll
      for (int j = 0; j < 16; j++) A[j] = j < c.predictLPCOrder ? (int32_t)A_Q12[j] : 0;
      const int chunk = (lag + 2 + 3) >> 2;
      const int nb = ltp_mem - lag - 2 + k * chunk, ne = nsq_min(nb + chunk, ltp_mem);
#pragma unroll
      for (int j = 0; j < 16; j++) w[j] = xin[nb - 1 - j];
      int32_t cur = xin[nb];
#pragma unroll 2
      for (int n = nb; n < ne; n++) {
        const int32_t nxt = xin[n + 1];
        uint32_t acc = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) acc += (uint32_t)(w[j] * A[j]);
        acc = ((uint32_t)cur << 12) - acc;
        const int32_t res = silk_sat16(silk_rshift_round((int32_t)acc, 12));
        if (live) S.sLTP_Q15[n] = silk_smulwb(inv_gain_Q31, res);
#pragma unroll
        for (int j = 15; j > 0; j--) w[j] = w[j - 1];
        w[0] = cur;
        cur = nxt;
      }
    }
    if (Gain_Q16 != prev_gain_Q16) {
      const int32_t adj = nsq_div32_varq(prev_gain_Q16, Gain_Q16, 16);
      if (live) {
#pragma unroll 8
        for (int i = shp_idx - ltp_mem + k; i < shp_idx; i += 4) S.sLTP_shp_Q14[i] = nsq_smulww(adj, S.sLTP_shp_Q14[i]);
        if (voiced && !rewhite) {
#pragma unroll 4
          for (int i = ltp_idx - lag - 2 + k; i < ltp_idx - decisionDelay; i += 4) S.sLTP_Q15[i] = nsq_smulww(adj, S.sLTP_Q15[i]);
        }
      }
      LF_AR = nsq_smulww(adj, LF_AR);
      Diff = nsq_smulww(adj, Diff);
#pragma unroll
      for (int j = 0; j < 16; j++) sLPC[j] = nsq_smulww(adj, sLPC[j]);
#pragma unroll
      for (int j = 0; j < 24; j++) sAR2[j] = nsq_smulww(adj, sAR2[j]);
#pragma unroll 1
      for (int t = 0; t < kNsqDelay; t++) {
        RING(RING_PRED, t, lane) = nsq_smulww(adj, RING(RING_PRED, t, lane));
        RING(RING_SHAPE, t, lane) = nsq_smulww(adj, RING(RING_SHAPE, t, lane));
      }
      prev_gain_Q16 = Gain_Q16;
    }
    wsync();

    // ---- the subframe's coefficients, pre-shifted so that every tap is one IMAD.HI ----
    int32_t a_sh[16], ar_sh[24], b_sh[5];
#pragma unroll
    for (int j = 0; j < 16; j++) a_sh[j] = j < c.predictLPCOrder ? nsq_hi16(A_Q12[j]) : 0;
#pragma unroll
    for (int j = 0; j < 24; j++) ar_sh[j] = j < c.shapingLPCOrder ? nsq_hi16(F.AR_Q13[sf * 24 + j]) : 0;
#pragma unroll
    for (int j = 0; j < 5; j++) b_sh[j] = nsq_hi16(F.LTPCoef_Q14[sf * 5 + j]);
    const int32_t hsg = F.HarmShapeGain_Q14[sf];
    const int32_t HarmPacked = (hsg >> 2) | nsq_shl(hsg >> 1, 16);
    const int32_t harm_lo = nsq_hi16(HarmPacked), harm_hi = (int32_t)((uint32_t)HarmPacked & 0xffff0000u);
    const int32_t tilt_sh = nsq_hi16(F.Tilt_Q14[sf]);
    const int32_t LF = F.LF_shp_Q14[sf];
    const int32_t lf_lo = nsq_hi16(LF), lf_hi = (int32_t)((uint32_t)LF & 0xffff0000u);
    const int32_t warp_sh = nsq_hi16(c.warping_Q16);
    const int32_t Gain_Q10 = Gain_Q16 >> 6;
    const int order = ORDER ? ORDER : c.shapingLPCOrder;
    const int32_t *shp_ptr = &S.sLTP_shp_Q14[shp_idx - lag + 1];
    const int32_t *pred_ptr = &S.sLTP_Q15[ltp_idx - lag + 2];
    const int out0 = sf * subfr_length - decisionDelay;  // frame-relative index of the sample decided at i == 0
    // Register windows over the LTP and harmonic-shaping taps.  The element entering at sample i+1 was written (by this
    // group's winner) at sample i+1-lag+2+delay <= i at the latest, so it is fetched right after sample i's closing wsync.
    int32_t pw[5] = {0, 0, 0, 0, 0}, sw[3] = {0, 0, 0};
    if (voiced) {
#pragma unroll
      for (int j = 0; j < 5; j++) pw[j] = pred_ptr[-j];
    }
    if (lag > 0) {
#pragma unroll
      for (int j = 0; j < 3; j++) sw[j] = shp_ptr[-j];
    }

#pragma unroll 1
    for (int i = 0; i < subfr_length; i++) {
      // ---- common to all states of the stream (tap windows pw/sw were loaded one sample ahead) ----
      int32_t LTP_pred_Q14 = 0, n_LTP_Q14 = 0;
      if (voiced) {
        int32_t acc = 2;
#pragma unroll
        for (int j = 0; j < 5; j++) acc = nsq_addw(acc, mulhi32(pw[j], b_sh[j]));
        LTP_pred_Q14 = nsq_shl(acc, 1);
      }
      if (lag > 0) {
        int32_t n = mulhi32(nsq_add_sat(sw[0], sw[2]), harm_lo);
        n = nsq_addw(n, mulhi32(sw[1], harm_hi));
        n_LTP_Q14 = LTP_pred_Q14 - nsq_shl(n, 2);
      }
      const int32_t x_Q10 = nsq_smulww(x16[sf * subfr_length + i], inv_gain_Q26);

      // ---- this lane's state ----
      Seed = nsq_lcg(Seed);
      int32_t LPC_pred_Q14 = c.predictLPCOrder >> 1;
#pragma unroll
      for (int j = 0; j < 16; j++) LPC_pred_Q14 = nsq_addw(LPC_pred_Q14, mulhi32(sLPC[j], a_sh[j]));
      LPC_pred_Q14 = nsq_shl(LPC_pred_Q14, 4);
      // warped noise-shaping filter (NSQ_del_dec.c:420-443)
      int32_t t2 = nsq_addw(Diff, mulhi32(sAR2[0], warp_sh));
      int32_t t1 = nsq_addw(sAR2[0], mulhi32(nsq_subw(sAR2[1], t2), warp_sh));
      sAR2[0] = t2;
      int32_t n_AR_Q14 = nsq_addw(order >> 1, mulhi32(t2, ar_sh[0]));
      // With a run-time order, sections beyond it are computed and discarded with warp-uniform selects, which keeps the
      // whole delay line in registers.  ar_sh[] is zero from `order` on, so the first skipped section contributes
      // exactly the reference's closing tap t1*AR[order-1] and the later ones nothing.
#pragma unroll
      for (int j = 2; j < (ORDER ? ORDER : 24); j += 2) {
        const bool on = ORDER ? true : j < order;
        const int32_t u2 = nsq_addw(sAR2[j - 1], mulhi32(nsq_subw(sAR2[j], t1), warp_sh));
        const int32_t u1 = nsq_addw(sAR2[j], mulhi32(nsq_subw(sAR2[j + 1], u2), warp_sh));
        n_AR_Q14 = nsq_addw(n_AR_Q14, mulhi32(t1, ar_sh[j - 1]));
        n_AR_Q14 = nsq_addw(n_AR_Q14, mulhi32(u2, on ? ar_sh[j] : 0));
        sAR2[j - 1] = (ORDER || j <= order) ? t1 : sAR2[j - 1];
        sAR2[j] = on ? u2 : sAR2[j];
        t1 = on ? u1 : t1;
      }
      if (ORDER) {
        n_AR_Q14 = nsq_addw(n_AR_Q14, mulhi32(t1, ar_sh[ORDER ? ORDER - 1 : 23]));
        sAR2[ORDER ? ORDER - 1 : 23] = t1;
      } else {
        n_AR_Q14 = nsq_addw(n_AR_Q14, mulhi32(t1, ar_sh[23]));
        sAR2[23] = order == 24 ? t1 : sAR2[23];
      }
      n_AR_Q14 = nsq_shl(n_AR_Q14, 1);
      n_AR_Q14 = nsq_addw(n_AR_Q14, mulhi32(LF_AR, tilt_sh));
      n_AR_Q14 = nsq_shl(n_AR_Q14, 2);
      int32_t n_LF_Q14 = mulhi32(RING(RING_SHAPE, smpl, gbase + map.get(smpl)), lf_lo);
      n_LF_Q14 = nsq_addw(n_LF_Q14, mulhi32(LF_AR, lf_hi));
      n_LF_Q14 = nsq_shl(n_LF_Q14, 2);
      int32_t tmp1 = nsq_add_sat(n_AR_Q14, n_LF_Q14);
      const int32_t tmp2 = nsq_addw(n_LTP_Q14, LPC_pred_Q14);
      tmp1 = nsq_sub_sat(tmp2, tmp1);
      tmp1 = silk_rshift_round(tmp1, 4);
      int32_t r_Q10 = x_Q10 - tmp1;
      if (Seed < 0) r_Q10 = -r_Q10;
      r_Q10 = nsq_min(nsq_max(r_Q10, -(31 << 10)), 30 << 10);
      int32_t q1, q2, rd1, rd2;
      nsq_candidates(r_Q10, offset_Q10, Lambda_Q10, q1, q2, rd1, rd2);
      int32_t rr = r_Q10 - q1;
      rd1 = (rd1 + nsq_smulbb(rr, rr)) >> 10;
      rr = r_Q10 - q2;
      rd2 = (rd2 + nsq_smulbb(rr, rr)) >> 10;
      // best (c0) and second best (c1) continuation of this state
      const bool first = rd1 < rd2;
      int32_t c0_Q = first ? q1 : q2, c1_Q = first ? q2 : q1;
      int32_t c0_RD = RD + (first ? rd1 : rd2), c1_RD = RD + (first ? rd2 : rd1);
      int32_t e0 = nsq_shl(c0_Q, 4), e1 = nsq_shl(c1_Q, 4);
      if (Seed < 0) { e0 = -e0; e1 = -e1; }
      const int32_t x_Q14 = nsq_shl(x_Q10, 4);
      int32_t c0_exc = e0 + LTP_pred_Q14, c1_exc = e1 + LTP_pred_Q14;
      int32_t c0_xq = nsq_addw(c0_exc, LPC_pred_Q14), c1_xq = nsq_addw(c1_exc, LPC_pred_Q14);
      int32_t c0_Diff = nsq_subw(c0_xq, x_Q14), c1_Diff = nsq_subw(c1_xq, x_Q14);
      int32_t c0_LF = nsq_subw(c0_Diff, n_AR_Q14), c1_LF = nsq_subw(c1_Diff, n_AR_Q14);
      int32_t c0_shp = nsq_sub_sat(c0_LF, n_LF_Q14), c1_shp = nsq_sub_sat(c1_LF, n_LF_Q14);

      // ---- trellis bookkeeping across the 4 lanes ----
      smpl = smpl == 0 ? kNsqDelay - 1 : smpl - 1;
      int last = smpl + decisionDelay;
      if (last >= kNsqDelay) last -= kNsqDelay;
      int32_t v = act ? c0_RD : 2147483647;
      int winner = act ? k : 4 + k;
      nsq_grp_argmin(v, winner);
      const int32_t my_rand = RING(RING_RAND, last, gbase + map.get(last));
      const int32_t win_rand = shfl(my_rand, gbase + winner);
      if (my_rand != win_rand) {
        c0_RD += 2147483647 >> 4;
        c1_RD += 2147483647 >> 4;
      }
      int32_t vmax = act ? c0_RD : (int32_t)(-2147483647 - 1), vmin = act ? c1_RD : 2147483647;
      int imax = act ? k : 4 + k, imin = imax;
      nsq_grp_argmax(vmax, imax);
      nsq_grp_argmin(vmin, imin);
      const bool take = vmin < vmax && k == imax;   // this lane's state is replaced by a copy of state imin
      const int src = take ? gbase + imin : lane;
#pragma unroll
      for (int j = 0; j < 16; j++) sLPC[j] = shfl(sLPC[j], src);
      // taps past a compile-time order never diverge between the states (same start, same rescaling): nothing to copy
#pragma unroll
      for (int j = 0; j < (ORDER ? ORDER : 24); j++) sAR2[j] = shfl(sAR2[j], src);
      Seed = shfl(Seed, src);
      SeedInit = shfl(SeedInit, src);
      map.w0 = shfl(map.w0, src); map.w1 = shfl(map.w1, src); map.w2 = shfl(map.w2, src);
      {
        const int32_t tQ = shfl(c1_Q, src), tRD = shfl(c1_RD, src), txq = shfl(c1_xq, src), tLF = shfl(c1_LF, src),
                      tDiff = shfl(c1_Diff, src), tshp = shfl(c1_shp, src), texc = shfl(c1_exc, src);
        if (take) { c0_Q = tQ; c0_RD = tRD; c0_xq = txq; c0_LF = tLF; c0_Diff = tDiff; c0_shp = tshp; c0_exc = texc; }
      }
      // ---- the winner writes the sample that falls out of the decision window ----
      if (k == winner && live && (subfr > 0 || i >= decisionDelay)) {
        const int col = gbase + map.get(last);
        pulses[out0 + i] = (int8_t)silk_rshift_round(RING(RING_Q, last, col), 10);
        S.xq[ltp_mem + out0 + i] = silk_sat16(silk_rshift_round(nsq_smulww(RING(RING_XQ, last, col), dgain[last * 8 + g]), 8));
        S.sLTP_shp_Q14[shp_idx + i - decisionDelay] = RING(RING_SHAPE, last, col);
        S.sLTP_Q15[ltp_idx + i - decisionDelay] = RING(RING_PRED, last, col);
      }
      wsync();  // ring reads of this sample are done before slot `smpl` (== `last` when the delay is 40) is rewritten
      // ---- update states ----
      LF_AR = c0_LF;
      Diff = c0_Diff;
#pragma unroll
      for (int j = 15; j > 0; j--) sLPC[j] = sLPC[j - 1];
      sLPC[0] = c0_xq;
      RING(RING_XQ, smpl, lane) = c0_xq;
      RING(RING_Q, smpl, lane) = c0_Q;
      RING(RING_PRED, smpl, lane) = nsq_shl(c0_exc, 1);
      RING(RING_SHAPE, smpl, lane) = c0_shp;
      Seed = nsq_addw(Seed, silk_rshift_round(c0_Q, 10));
      RING(RING_RAND, smpl, lane) = Seed;
      RD = c0_RD;
      map.set(smpl, k);
      if (k == 0) dgain[smpl * 8 + g] = Gain_Q10;
      wsync();  // publishes the rings and the sLTP / shaping samples to the group's other lanes
      if (voiced) {
#pragma unroll
        for (int j = 4; j > 0; j--) pw[j] = pw[j - 1];
        pw[0] = pred_ptr[i + 1];
      }
      if (lag > 0) {
        sw[2] = sw[1]; sw[1] = sw[0];
        sw[0] = shp_ptr[i + 1];
      }
    }
    shp_idx += subfr_length;
    ltp_idx += subfr_length;
    subfr++;
  }

  // ---- final winner: flush the decision window, hand the state back (NSQ_del_dec.c:287-327) ----
  {
    int32_t v = act ? RD : 2147483647;
    int w = act ? k : 4 + k;
    nsq_grp_argmin(v, w);
    NsqMap wm;
    wm.w0 = shfl(map.w0, gbase + w); wm.w1 = shfl(map.w1, gbase + w); wm.w2 = shfl(map.w2, gbase + w);
    const int32_t Gain_Q10 = F.Gains_Q16[c.nb_subfr - 1] >> 6;
#pragma unroll 1
    for (int i = k; i < decisionDelay; i += 4) {
      const int t = (smpl + decisionDelay - 1 - i) % kNsqDelay;
      const int col = gbase + wm.get(t);
      const int o = frame_length + i - decisionDelay;
      if (live) {
        pulses[o] = (int8_t)silk_rshift_round(RING(RING_Q, t, col), 10);
        S.xq[ltp_mem + o] = silk_sat16(silk_rshift_round(nsq_smulww(RING(RING_XQ, t, col), Gain_Q10), 8));
        S.sLTP_shp_Q14[shp_idx - decisionDelay + i] = RING(RING_SHAPE, t, col);
      }
    }
    if (k == w && live) {
      F.Seed = SeedInit;
#pragma unroll
      for (int j = 0; j < 16; j++) S.sLPC_Q14[15 - j] = sLPC[j];
#pragma unroll
      for (int j = 0; j < 24; j++) S.sAR2_Q14[j] = sAR2[j];
      S.sLF_AR_shp_Q14 = LF_AR;
      S.sDiff_shp_Q14 = Diff;
      S.lagPrev = F.pitchL[c.nb_subfr - 1];
      S.prev_gain_Q16 = prev_gain_Q16;
    }
  }
  wsync();
  // ---- slide the output / shaping histories by one frame (memmove, NSQ_del_dec.c:325-326): blocks of frame_length
  //      elements never overlap their own source, so only block boundaries need ordering ----
#pragma unroll 1
  for (int base = 0; base < ltp_mem; base += frame_length) {
    const int end = nsq_min(base + frame_length, ltp_mem);
    if (live) {
#pragma unroll 4
      for (int i = base + k; i < end; i += 4) {
        S.xq[i] = S.xq[i + frame_length];
        S.sLTP_shp_Q14[i] = S.sLTP_shp_Q14[i + frame_length];
      }
    }
    wsync();
  }
#undef RING
}

// ================================================================================================================
// plain quantizer (complexity 0-1: one state, no warping): one thread per stream
// ================================================================================================================
OB_DEV void nsq_plain_thread(const SilkNsqParams &p, int stream) {
  const ObNsqCfg c = p.cfg;
  const int subfr_length = 5 * c.fs_khz, ltp_mem = 20 * c.fs_khz, frame_length = subfr_length * c.nb_subfr;
  ObNsqState &S = p.state[stream];
  const ObNsqFrame &F = p.frames[stream];
  const int16_t *x16 = p.x16 + (size_t)stream * frame_length;
  int8_t *pulses = p.pulses + (size_t)stream * frame_length;
  int32_t sLPC[16], sAR2[24];
#pragma unroll
  for (int j = 0; j < 16; j++) sLPC[j] = S.sLPC_Q14[15 - j];
#pragma unroll
  for (int j = 0; j < 24; j++) sAR2[j] = S.sAR2_Q14[j];
  int32_t sLF_AR = S.sLF_AR_shp_Q14, sDiff = S.sDiff_shp_Q14, prev_gain_Q16 = S.prev_gain_Q16;
  int32_t rand_seed = F.Seed;
  const int signalType = F.signalType;
  const bool voiced = signalType == 2;
  const int32_t offset_Q10 = nsq_quant_offset(signalType, F.quantOffsetType);
  const int32_t Lambda_Q10 = F.Lambda_Q10;
  const int interp = F.NLSFInterpCoef_Q2 == 4 ? 0 : 1;
  const int order = c.shapingLPCOrder;
  int lag = S.lagPrev;
  int shp_idx = ltp_mem, ltp_idx = ltp_mem;
#pragma unroll 1
  for (int sf = 0; sf < c.nb_subfr; sf++) {
    const int16_t *A_Q12 = &F.PredCoef_Q12[((sf >> 1) | (1 - interp)) * 16];
    const int32_t Gain_Q16 = F.Gains_Q16[sf];
    bool rewhite = false;
    if (voiced) {
      lag = F.pitchL[sf];
      if ((sf & (3 - (interp << 1))) == 0) {
        rewhite = true;
        ltp_idx = ltp_mem;
      }
    }
    // scale states (NSQ.c:390-437)
    int32_t inv_gain_Q31 = nsq_inverse32_varq(Gain_Q16 > 1 ? Gain_Q16 : 1, 47);
    const int32_t inv_gain_Q26 = silk_rshift_round(inv_gain_Q31, 5);
    if (rewhite) {
      if (sf == 0) inv_gain_Q31 = nsq_shl(silk_smulwb(inv_gain_Q31, F.LTP_scale_Q14), 2);
      const int d = c.predictLPCOrder;
      const int16_t *xin = &S.xq[sf * subfr_length];
#pragma unroll 1
      for (int n = ltp_mem - lag - 2; n < ltp_mem; n++) {
        uint32_t acc = 0;
#pragma unroll 1
        for (int j = 0; j < d; j++) acc += (uint32_t)((int32_t)xin[n - 1 - j] * A_Q12[j]);
        acc = ((uint32_t)((int32_t)xin[n]) << 12) - acc;
        S.sLTP_Q15[n] = silk_smulwb(inv_gain_Q31, silk_sat16(silk_rshift_round((int32_t)acc, 12)));
      }
    }
    if (Gain_Q16 != prev_gain_Q16) {
      const int32_t adj = nsq_div32_varq(prev_gain_Q16, Gain_Q16, 16);
#pragma unroll 1
      for (int i = shp_idx - ltp_mem; i < shp_idx; i++) S.sLTP_shp_Q14[i] = nsq_smulww(adj, S.sLTP_shp_Q14[i]);
      if (voiced && !rewhite) {
#pragma unroll 1
        for (int i = ltp_idx - lag - 2; i < ltp_idx; i++) S.sLTP_Q15[i] = nsq_smulww(adj, S.sLTP_Q15[i]);
      }
      sLF_AR = nsq_smulww(adj, sLF_AR);
      sDiff = nsq_smulww(adj, sDiff);
#pragma unroll
      for (int j = 0; j < 16; j++) sLPC[j] = nsq_smulww(adj, sLPC[j]);
#pragma unroll
      for (int j = 0; j < 24; j++) sAR2[j] = nsq_smulww(adj, sAR2[j]);
      prev_gain_Q16 = Gain_Q16;
    }
    int32_t a_sh[16], ar_sh[24], b_sh[5];
#pragma unroll
    for (int j = 0; j < 16; j++) a_sh[j] = j < c.predictLPCOrder ? nsq_hi16(A_Q12[j]) : 0;
#pragma unroll
    for (int j = 0; j < 24; j++) ar_sh[j] = j < order ? nsq_hi16(F.AR_Q13[sf * 24 + j]) : 0;
#pragma unroll
    for (int j = 0; j < 5; j++) b_sh[j] = nsq_hi16(F.LTPCoef_Q14[sf * 5 + j]);
    const int32_t hsg = F.HarmShapeGain_Q14[sf];
    const int32_t HarmPacked = (hsg >> 2) | nsq_shl(hsg >> 1, 16);
    const int32_t harm_lo = nsq_hi16(HarmPacked), harm_hi = (int32_t)((uint32_t)HarmPacked & 0xffff0000u);
    const int32_t tilt_sh = nsq_hi16(F.Tilt_Q14[sf]);
    const int32_t LF = F.LF_shp_Q14[sf];
    const int32_t lf_lo = nsq_hi16(LF), lf_hi = (int32_t)((uint32_t)LF & 0xffff0000u);
    const int32_t Gain_Q10 = Gain_Q16 >> 6;
    const int32_t *shp_ptr = &S.sLTP_shp_Q14[shp_idx - lag + 1];
    const int32_t *pred_ptr = &S.sLTP_Q15[ltp_idx - lag + 2];
    int32_t shp_prev = S.sLTP_shp_Q14[shp_idx - 1];
#pragma unroll 1
    for (int i = 0; i < subfr_length; i++) {
      const int32_t x_Q10 = nsq_smulww(x16[sf * subfr_length + i], inv_gain_Q26);
      rand_seed = nsq_lcg(rand_seed);
      int32_t LPC_pred_Q10 = c.predictLPCOrder >> 1;
#pragma unroll
      for (int j = 0; j < 16; j++) LPC_pred_Q10 = nsq_addw(LPC_pred_Q10, mulhi32(sLPC[j], a_sh[j]));
      int32_t LTP_pred_Q13 = 0;
      if (voiced) {
        LTP_pred_Q13 = 2;
#pragma unroll
        for (int j = 0; j < 5; j++) LTP_pred_Q13 = nsq_addw(LTP_pred_Q13, mulhi32(pred_ptr[i - j], b_sh[j]));
      }
      // noise shape feedback (NSQ.h:71-97) is a delay line fed by sDiff: shift, then one dot product
#pragma unroll
      for (int j = 23; j > 0; j--) sAR2[j] = j < order ? sAR2[j - 1] : sAR2[j];
      sAR2[0] = sDiff;
      int32_t n_AR_Q12 = order >> 1;
#pragma unroll
      for (int j = 0; j < 24; j++) n_AR_Q12 = nsq_addw(n_AR_Q12, mulhi32(sAR2[j], ar_sh[j]));
      n_AR_Q12 = nsq_shl(n_AR_Q12, 1);
      n_AR_Q12 = nsq_addw(n_AR_Q12, mulhi32(sLF_AR, tilt_sh));
      int32_t n_LF_Q12 = mulhi32(shp_prev, lf_lo);
      n_LF_Q12 = nsq_addw(n_LF_Q12, mulhi32(sLF_AR, lf_hi));
      int32_t tmp1 = nsq_subw(nsq_shl(LPC_pred_Q10, 2), n_AR_Q12);
      tmp1 = nsq_subw(tmp1, n_LF_Q12);
      if (lag > 0) {
        int32_t n = mulhi32(nsq_add_sat(shp_ptr[i], shp_ptr[i - 2]), harm_lo);
        n = nsq_addw(n, mulhi32(shp_ptr[i - 1], harm_hi));
        n = nsq_shl(n, 1);
        const int32_t tmp2 = LTP_pred_Q13 - n;
        tmp1 = nsq_addw(tmp2, nsq_shl(tmp1, 1));
        tmp1 = silk_rshift_round(tmp1, 3);
      } else tmp1 = silk_rshift_round(tmp1, 2);
      int32_t r_Q10 = x_Q10 - tmp1;
      if (rand_seed < 0) r_Q10 = -r_Q10;
      r_Q10 = nsq_min(nsq_max(r_Q10, -(31 << 10)), 30 << 10);
      int32_t q1, q2, rd1, rd2;
      nsq_candidates(r_Q10, offset_Q10, Lambda_Q10, q1, q2, rd1, rd2);
      int32_t rr = r_Q10 - q1;
      rd1 = rd1 + nsq_smulbb(rr, rr);
      rr = r_Q10 - q2;
      rd2 = rd2 + nsq_smulbb(rr, rr);
      if (rd2 < rd1) q1 = q2;
      const int32_t pulse = (int8_t)silk_rshift_round(q1, 10);
      pulses[sf * subfr_length + i] = (int8_t)pulse;
      int32_t exc_Q14 = nsq_shl(q1, 4);
      if (rand_seed < 0) exc_Q14 = -exc_Q14;
      const int32_t LPC_exc_Q14 = exc_Q14 + nsq_shl(LTP_pred_Q13, 1);
      const int32_t xq_Q14 = nsq_addw(LPC_exc_Q14, nsq_shl(LPC_pred_Q10, 4));
      S.xq[ltp_mem + sf * subfr_length + i] = silk_sat16(silk_rshift_round(nsq_smulww(xq_Q14, Gain_Q10), 8));
#pragma unroll
      for (int j = 15; j > 0; j--) sLPC[j] = sLPC[j - 1];
      sLPC[0] = xq_Q14;
      sDiff = nsq_subw(xq_Q14, nsq_shl(x_Q10, 4));
      sLF_AR = nsq_subw(sDiff, nsq_shl(n_AR_Q12, 2));
      shp_prev = nsq_subw(sLF_AR, nsq_shl(n_LF_Q12, 2));
      S.sLTP_shp_Q14[shp_idx + i] = shp_prev;
      S.sLTP_Q15[ltp_idx + i] = nsq_shl(LPC_exc_Q14, 1);
      rand_seed = nsq_addw(rand_seed, pulse);
    }
    shp_idx += subfr_length;
    ltp_idx += subfr_length;
  }
#pragma unroll
  for (int j = 0; j < 16; j++) S.sLPC_Q14[15 - j] = sLPC[j];
#pragma unroll
  for (int j = 0; j < 24; j++) S.sAR2_Q14[j] = sAR2[j];
  S.sLF_AR_shp_Q14 = sLF_AR;
  S.sDiff_shp_Q14 = sDiff;
  S.rand_seed = rand_seed;
  S.prev_gain_Q16 = prev_gain_Q16;
  S.lagPrev = F.pitchL[c.nb_subfr - 1];
#pragma unroll 1
  for (int i = 0; i < ltp_mem; i++) {
    S.xq[i] = S.xq[i + frame_length];
    S.sLTP_shp_Q14[i] = S.sLTP_shp_Q14[i + frame_length];
  }
}

}  // namespace ob
