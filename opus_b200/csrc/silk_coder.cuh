// silk_coder.cuh — the SILK excitation coder on the device (first slice): silk_encode_pulses (silk/encode_pulses.c:59-206),
// silk_shell_encoder (silk/shell_coder.c:78-112), silk_encode_signs (silk/code_signs.c:40-72), one THREAD per stream on a range-encoder state
// that lives in HBM between calls, so that other symbols can be coded in front of and behind the pulses in the same range coder — the order of
// silk_encode_frame_FLP (silk/float/encode_frame_FLP.c:290-300: indices, then pulses) and, with CELT behind it, of a hybrid frame
// (celt/celt_encoder.c:1861-1869: the seam B2 of SURVEY 8(b)).  Integer arithmetic only: bit-exact.
#pragma once
#include "celt_enc_frame.cuh"   // RangeEnc and its primitives (celt/entenc.c)
#include "silk_tables_data.h"

namespace ob {

struct ObRangeEncState { uint32_t storage, end_offs, end_window, offs, rng, val, ext; int32_t nend_bits, nbits_total, rem, error; };   // ec_enc minus the buffer pointer
struct SilkCoderParams {
  ObRangeEncState *state;     // [n]
  unsigned char *buf;         // [n][buf_stride]
  int n, buf_stride, storage;
  // encode_pulses
  const int32_t *signal_type, *quant_offset_type;   // [n]
  const int8_t *pulses;                             // [n][pulses_stride]
  int pulses_stride, frame_length;
  // one symbol per stream with a caller-supplied table (ec_enc_icdf)
  const int32_t *symbols;     // [n], < 0 = nothing for this stream
  unsigned char icdf[64];
  int ftb;
  // done
  uint32_t *final_range, *tell_frac;   // [n]
  // encode_indices
  const struct ObSilkIndices *indices;  // [n]
  const int32_t *cond_coding;           // [n]: CODE_INDEPENDENTLY 0, CODE_INDEPENDENTLY_NO_LTP_SCALING 1, CODE_CONDITIONALLY 2 (silk/define.h:75-77)
  int32_t *silk_prev;                   // [n][2]: ec_prevSignalType, ec_prevLagIndex of the encoder state (silk/structs.h)
  int fs_kHz, nb_subfr;
};
struct ObSilkIndices {   // SideInfoIndices (silk/structs.h:129-141): the same fields in the same order
  int8_t GainsIndices[4], LTPIndex[4], NLSFIndices[17];
  int16_t lagIndex;
  int8_t contourIndex, signalType, quantOffsetType, NLSFInterpCoef_Q2, PERIndex, LTP_scaleIndex, Seed;
};
OB_DEV void coder_load(RangeEnc &e, const SilkCoderParams &p, int s) {
  const ObRangeEncState &st = p.state[s];
  e.buf = p.buf + (size_t)s * p.buf_stride;
  e.storage = st.storage; e.end_offs = st.end_offs; e.end_window = st.end_window; e.offs = st.offs; e.rng = st.rng; e.val = st.val; e.ext = st.ext;
  e.nend_bits = st.nend_bits; e.nbits_total = st.nbits_total; e.rem = st.rem; e.error = st.error;
}
OB_DEV void coder_store(const RangeEnc &e, const SilkCoderParams &p, int s) {
  ObRangeEncState &st = p.state[s];
  st.storage = e.storage; st.end_offs = e.end_offs; st.end_window = e.end_window; st.offs = e.offs; st.rng = e.rng; st.val = e.val; st.ext = e.ext;
  st.nend_bits = e.nend_bits; st.nbits_total = e.nbits_total; st.rem = e.rem; st.error = e.error;
}
OB_DEV void silk_coder_init_thread(const SilkCoderParams &p, int s) {   // ec_enc_init (celt/entenc.c:108-122)
  RangeEnc e;
  re_init(e, p.buf + (size_t)s * p.buf_stride, (uint32_t)p.storage);
  coder_store(e, p, s);
}
OB_DEV void silk_coder_icdf_thread(const SilkCoderParams &p, int s) {
  const int sym = p.symbols[s];
  if (sym < 0) return;
  RangeEnc e;
  coder_load(e, p, s);
  re_icdf(e, sym, p.icdf, (unsigned)p.ftb);
  coder_store(e, p, s);
}
OB_DEV void silk_coder_done_thread(const SilkCoderParams &p, int s) {   // ec_tell_frac, then ec_enc_done (celt/entenc.c:256-305); final range = rng
  RangeEnc e;
  coder_load(e, p, s);
  if (p.tell_frac) p.tell_frac[s] = re_tell_frac(e);
  re_done(e);
  coder_store(e, p, s);
  if (p.final_range) p.final_range[s] = e.rng;
}

OB_DEV int silk_pulse_at(const int8_t *q, int i, int frame_length) { return i < frame_length ? (int)q[i] : 0; }   // the reference zero-pads 10 ms @ 12 kHz in place
OB_DEV void silk_encode_split(RangeEnc &e, int p_child1, int pp, int level) {   // shell_coder.c:49-59
  if (pp > 0) re_icdf(e, p_child1, &ob_silk_shell_code_tables[152 * level + ob_silk_shell_code_table_offsets[pp]], 8);
}
// one 16-pulse block: |pulses| >> shift in registers
OB_DEV void silk_block_abs(int *ap, const int8_t *q, int base, int frame_length, int shift) {
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const int v = silk_pulse_at(q, base + k, frame_length);
    ap[k] = (v < 0 ? -v : v) >> shift;
  }
}
// combine_and_check cascade (encode_pulses.c:38-56, 111-131): 0 and the block sum, or 1 when a partial sum exceeds its limit
OB_DEV int silk_block_sum(const int *ap, int &sum) {
  int c8[8], c4[4], c2[2], bad = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) { c8[k] = ap[2 * k] + ap[2 * k + 1]; bad |= c8[k] > ob_silk_max_pulses_table[0]; }
#pragma unroll
  for (int k = 0; k < 4; k++) { c4[k] = c8[2 * k] + c8[2 * k + 1]; bad |= c4[k] > ob_silk_max_pulses_table[1]; }
#pragma unroll
  for (int k = 0; k < 2; k++) { c2[k] = c4[2 * k] + c4[2 * k + 1]; bad |= c2[k] > ob_silk_max_pulses_table[2]; }
  sum = c2[0] + c2[1];
  bad |= sum > ob_silk_max_pulses_table[3];
  return bad;
}
OB_DEVNI void silk_shell_encoder_dev(RangeEnc &e, const int *p0) {   // shell_coder.c:78-112
  int p1[8], p2[4], p3[2];
#pragma unroll
  for (int k = 0; k < 8; k++) p1[k] = p0[2 * k] + p0[2 * k + 1];
#pragma unroll
  for (int k = 0; k < 4; k++) p2[k] = p1[2 * k] + p1[2 * k + 1];
#pragma unroll
  for (int k = 0; k < 2; k++) p3[k] = p2[2 * k] + p2[2 * k + 1];
  silk_encode_split(e, p3[0], p3[0] + p3[1], 3);
#pragma unroll
  for (int k = 0; k < 2; k++) {
    silk_encode_split(e, p2[2 * k], p3[k], 2);
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int i1 = 2 * k + q;
      silk_encode_split(e, p1[2 * i1], p2[i1], 1);
      silk_encode_split(e, p0[4 * i1], p1[2 * i1], 0);
      silk_encode_split(e, p0[4 * i1 + 2], p1[2 * i1 + 1], 0);
    }
  }
}
// silk_encode_pulses for one stream
OB_DEV void silk_encode_pulses_thread(const SilkCoderParams &p, int s) {
  RangeEnc e;
  coder_load(e, p, s);
  const int8_t *q = p.pulses + (size_t)s * p.pulses_stride;
  const int frame_length = p.frame_length, signalType = p.signal_type[s], quantOffsetType = p.quant_offset_type[s];
  int iter = frame_length >> 4;
  if (iter * 16 < frame_length) iter++;
  unsigned char sum_pulses[21], nRshifts[21];
  int ap[16];
#pragma unroll 1
  for (int i = 0; i < iter; i++) {
    int shift = 0, sum = 0;
#pragma unroll 1
    for (;;) {
      silk_block_abs(ap, q, 16 * i, frame_length, shift);
      if (!silk_block_sum(ap, sum)) break;
      shift++;
    }
    sum_pulses[i] = (unsigned char)sum;
    nRshifts[i] = (unsigned char)shift;
  }
  // the rate level with the fewest bits for the pulses-per-block symbols (:137-153)
  int RateLevelIndex = 0;
  int32_t minSumBits_Q5 = 0x7fffffff;
#pragma unroll 1
  for (int k = 0; k < 9; k++) {
    const unsigned char *nBits = &ob_silk_pulses_per_block_BITS_Q5[18 * k];
    int32_t sumBits_Q5 = ob_silk_rate_levels_BITS_Q5[9 * (signalType >> 1) + k];
#pragma unroll 1
    for (int i = 0; i < iter; i++) sumBits_Q5 += nRshifts[i] > 0 ? nBits[17] : nBits[sum_pulses[i]];
    if (sumBits_Q5 < minSumBits_Q5) { minSumBits_Q5 = sumBits_Q5; RateLevelIndex = k; }
  }
  re_icdf(e, RateLevelIndex, &ob_silk_rate_levels_iCDF[9 * (signalType >> 1)], 8);
  {
    const unsigned char *cdf = &ob_silk_pulses_per_block_iCDF[18 * RateLevelIndex], *last = &ob_silk_pulses_per_block_iCDF[18 * 9];
#pragma unroll 1
    for (int i = 0; i < iter; i++) {
      if (nRshifts[i] == 0) re_icdf(e, sum_pulses[i], cdf, 8);
      else {
        re_icdf(e, 17, cdf, 8);
#pragma unroll 1
        for (int k = 0; k < nRshifts[i] - 1; k++) re_icdf(e, 17, last, 8);
        re_icdf(e, sum_pulses[i], last, 8);
      }
    }
  }
#pragma unroll 1
  for (int i = 0; i < iter; i++)
    if (sum_pulses[i] > 0) {
      silk_block_abs(ap, q, 16 * i, frame_length, nRshifts[i]);
      silk_shell_encoder_dev(e, ap);
    }
#pragma unroll 1
  for (int i = 0; i < iter; i++)
    if (nRshifts[i] > 0) {
      const int nLS = nRshifts[i] - 1;
#pragma unroll 1
      for (int k = 0; k < 16; k++) {
        const int v = silk_pulse_at(q, 16 * i + k, frame_length);
        const int abs_q = (int)(int8_t)(v < 0 ? -v : v);
#pragma unroll 1
        for (int j = nLS; j > 0; j--) re_icdf(e, (abs_q >> j) & 1, ob_silk_lsb_iCDF, 8);
        re_icdf(e, abs_q & 1, ob_silk_lsb_iCDF, 8);
      }
    }
  {   // silk_encode_signs (code_signs.c:40-72)
    unsigned char icdf[2];
    const unsigned char *icdf_ptr = &ob_silk_sign_iCDF[7 * (quantOffsetType + (signalType << 1))];
    const int blocks = (frame_length + 8) >> 4;
    icdf[1] = 0;
#pragma unroll 1
    for (int i = 0; i < blocks; i++) {
      const int pp = sum_pulses[i];
      if (pp > 0) {
        icdf[0] = icdf_ptr[(pp & 0x1F) < 6 ? (pp & 0x1F) : 6];
#pragma unroll 1
        for (int j = 0; j < 16; j++) {
          const int v = silk_pulse_at(q, 16 * i + j, frame_length);
          if (v != 0) re_icdf(e, (v >> 15) + 1, icdf, 8);
        }
      }
    }
  }
  coder_store(e, p, s);
}

// silk_encode_indices (silk/encode_indices.c:35-185) with silk_NLSF_unpack (silk/NLSF_unpack.c:35-55) for one stream; the tables by fs_kHz /
// nb_subfr as silk_control_encoder selects them (silk/control_codec.c:212-296)
OB_DEV void silk_encode_indices_thread(const SilkCoderParams &p, int s) {
  RangeEnc e;
  coder_load(e, p, s);
  const ObSilkIndices ix = p.indices[s];
  const int condCoding = p.cond_coding[s], fs_kHz = p.fs_kHz, nb_subfr = p.nb_subfr;
  int32_t *prev = p.silk_prev + 2 * (size_t)s;
  const int wb = fs_kHz == 16, order = wb ? 16 : 10, typeOffset = 2 * ix.signalType + ix.quantOffsetType;
  const unsigned char *cb1 = &ob_silk_NLSF_CB1_iCDF[wb ? 64 : 0], *sel = &ob_silk_NLSF_CB2_SELECT[wb ? 160 : 0], *cb2 = &ob_silk_NLSF_CB2_iCDF[wb ? 72 : 0];
  if (typeOffset >= 2) re_icdf(e, typeOffset - 2, ob_silk_type_offset_VAD_iCDF, 8);
  else re_icdf(e, typeOffset, ob_silk_type_offset_no_VAD_iCDF, 8);
  if (condCoding == 2) re_icdf(e, ix.GainsIndices[0], ob_silk_delta_gain_iCDF, 8);
  else {
    re_icdf(e, ix.GainsIndices[0] >> 3, &ob_silk_gain_iCDF[8 * ix.signalType], 8);
    re_icdf(e, ix.GainsIndices[0] & 7, ob_silk_uniform8_iCDF, 8);
  }
#pragma unroll 1
  for (int i = 1; i < nb_subfr; i++) re_icdf(e, ix.GainsIndices[i], ob_silk_delta_gain_iCDF, 8);
  re_icdf(e, ix.NLSFIndices[0], &cb1[(ix.signalType >> 1) * 32], 8);
#pragma unroll 1
  for (int i = 0; i < order; i++) {
    const int entry = sel[ix.NLSFIndices[0] * order / 2 + (i >> 1)];
    const int ec_ix = (((i & 1) ? entry >> 5 : entry >> 1) & 7) * 9;
    const int v = ix.NLSFIndices[i + 1];
    if (v >= 4) {
      re_icdf(e, 8, &cb2[ec_ix], 8);
      re_icdf(e, v - 4, ob_silk_NLSF_EXT_iCDF, 8);
    } else if (v <= -4) {
      re_icdf(e, 0, &cb2[ec_ix], 8);
      re_icdf(e, -v - 4, ob_silk_NLSF_EXT_iCDF, 8);
    } else re_icdf(e, v + 4, &cb2[ec_ix], 8);
  }
  if (nb_subfr == 4) re_icdf(e, ix.NLSFInterpCoef_Q2, ob_silk_NLSF_interpolation_factor_iCDF, 8);
  if (ix.signalType == 2) {
    int encode_absolute = 1;
    if (condCoding == 2 && prev[0] == 2) {
      int delta = ix.lagIndex - prev[1];
      if (delta < -8 || delta > 11) delta = 0;
      else { delta += 9; encode_absolute = 0; }
      re_icdf(e, delta, ob_silk_pitch_delta_iCDF, 8);
    }
    if (encode_absolute) {
      const int half = fs_kHz >> 1, hi = ix.lagIndex / half, lo = ix.lagIndex - hi * half;
      re_icdf(e, hi, ob_silk_pitch_lag_iCDF, 8);
      re_icdf(e, lo, fs_kHz == 16 ? ob_silk_uniform8_iCDF : fs_kHz == 12 ? ob_silk_uniform6_iCDF : ob_silk_uniform4_iCDF, 8);
    }
    prev[1] = ix.lagIndex;
    re_icdf(e, ix.contourIndex,
            fs_kHz == 8 ? (nb_subfr == 4 ? ob_silk_pitch_contour_NB_iCDF : ob_silk_pitch_contour_10_ms_NB_iCDF)
                        : (nb_subfr == 4 ? ob_silk_pitch_contour_iCDF : ob_silk_pitch_contour_10_ms_iCDF), 8);
    re_icdf(e, ix.PERIndex, ob_silk_LTP_per_index_iCDF, 8);
#pragma unroll 1
    for (int k = 0; k < nb_subfr; k++) re_icdf(e, ix.LTPIndex[k], &ob_silk_LTP_gain_iCDF[ix.PERIndex == 0 ? 0 : ix.PERIndex == 1 ? 8 : 24], 8);
    if (condCoding == 0) re_icdf(e, ix.LTP_scaleIndex, ob_silk_LTPscale_iCDF, 8);
  }
  prev[0] = ix.signalType;
  re_icdf(e, ix.Seed, ob_silk_uniform4_iCDF, 8);
  coder_store(e, p, s);
}

}  // namespace ob
