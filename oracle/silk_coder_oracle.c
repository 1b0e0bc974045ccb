/* oracle/silk_coder_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 * Plain-C restatement of the SILK excitation coder: silk_encode_pulses (silk/encode_pulses.c:59-206), silk_shell_encoder
 * (silk/shell_coder.c:36-112), silk_encode_signs (silk/code_signs.c:40-72), on the range-encoder restatement of celt_rangeenc.inc
 * (celt/entenc.c).  The coder state is an object of its own so that other SILK symbols can precede / follow the pulses in the same
 * range coder, as in silk_encode_frame_FLP (silk/float/encode_frame_FLP.c:290-300).  Pinned against the reference's functions by
 * tests/test_silk_coder_oracle_vs_ref.py (buffer bytes, ec_tell_frac, final range). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OB_TABLE static const
#include "../opus_b200/csrc/silk_tables_data.h"

static int ilog32(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }
#define BITRES 3
#define IMIN(a, b) ((a) < (b) ? (a) : (b))
#include "celt_rangeenc.inc"

typedef struct { RangeEnc e; uint8_t buf[1280]; } OrcSilkCoder;
int orc_silk_coder_size(void) { return (int)sizeof(OrcSilkCoder); }
void orc_silk_coder_init(OrcSilkCoder *c, int storage) { re_init(&c->e, c->buf, (uint32_t)storage); }
void orc_silk_coder_icdf(OrcSilkCoder *c, int s, const unsigned char *icdf, int ftb) { re_icdf(&c->e, s, icdf, (unsigned)ftb); }
uint32_t orc_silk_coder_tell_frac(OrcSilkCoder *c) { return re_tell_frac(&c->e); }
/* ec_enc_done: buf [storage] bytes out, returns the final range */
uint32_t orc_silk_coder_done(OrcSilkCoder *c, uint8_t *out) {
  re_done(&c->e);
  memcpy(out, c->buf, c->e.storage);
  return c->e.rng;
}

static int combine_and_check(int *out, const int *in, int max_pulses, int len) { /* encode_pulses.c:38-56 */
  int k;
  for (k = 0; k < len; k++) {
    const int sum = in[2 * k] + in[2 * k + 1];
    if (sum > max_pulses) return 1;
    out[k] = sum;
  }
  return 0;
}
static void encode_split(RangeEnc *e, int p_child1, int p, int level) { /* shell_coder.c:49-59 */
  if (p > 0) re_icdf(e, p_child1, &ob_silk_shell_code_tables[152 * level + ob_silk_shell_code_table_offsets[p]], 8);
}
static void shell_encoder(RangeEnc *e, const int *p0) { /* shell_coder.c:78-112: the 15 splits of a 16-pulse block, parent first */
  int p1[8], p2[4], p3[2], p4, k;
  for (k = 0; k < 8; k++) p1[k] = p0[2 * k] + p0[2 * k + 1];
  for (k = 0; k < 4; k++) p2[k] = p1[2 * k] + p1[2 * k + 1];
  for (k = 0; k < 2; k++) p3[k] = p2[2 * k] + p2[2 * k + 1];
  p4 = p3[0] + p3[1];
  encode_split(e, p3[0], p4, 3);
  for (k = 0; k < 2; k++) {
    int q;
    encode_split(e, p2[2 * k], p3[k], 2);
    for (q = 0; q < 2; q++) {
      const int i1 = 2 * k + q;
      encode_split(e, p1[2 * i1], p2[i1], 1);
      encode_split(e, p0[4 * i1], p1[2 * i1], 0);
      encode_split(e, p0[4 * i1 + 2], p1[2 * i1 + 1], 0);
    }
  }
}
/* silk_encode_pulses; pulses must have room for iter * 16 values (the reference zero-pads the 10 ms @ 12 kHz frame in place) */
void orc_silk_encode_pulses(OrcSilkCoder *c, int signalType, int quantOffsetType, int8_t *pulses, int frame_length) {
  RangeEnc *e = &c->e;
  int abs_pulses[336], sum_pulses[21], nRshifts[21], comb[8];
  int i, k, j, iter, RateLevelIndex = 0;
  int32_t minSumBits_Q5, sumBits_Q5;
  memset(comb, 0, sizeof(comb));
  iter = frame_length >> 4;
  if (iter * 16 < frame_length) {
    iter++;
    memset(&pulses[frame_length], 0, 16);
  }
  for (i = 0; i < iter * 16; i++) abs_pulses[i] = abs((int)pulses[i]);
  for (i = 0; i < iter; i++) {
    int *ap = &abs_pulses[16 * i];
    nRshifts[i] = 0;
    for (;;) {
      int scale_down = combine_and_check(comb, ap, ob_silk_max_pulses_table[0], 8);
      scale_down += combine_and_check(comb, comb, ob_silk_max_pulses_table[1], 4);
      scale_down += combine_and_check(comb, comb, ob_silk_max_pulses_table[2], 2);
      scale_down += combine_and_check(&sum_pulses[i], comb, ob_silk_max_pulses_table[3], 1);
      if (!scale_down) break;
      nRshifts[i]++;
      for (k = 0; k < 16; k++) ap[k] >>= 1;
    }
  }
  minSumBits_Q5 = INT32_MAX;
  for (k = 0; k < 9; k++) {
    const unsigned char *nBits = &ob_silk_pulses_per_block_BITS_Q5[18 * k];
    sumBits_Q5 = ob_silk_rate_levels_BITS_Q5[9 * (signalType >> 1) + k];
    for (i = 0; i < iter; i++) sumBits_Q5 += nRshifts[i] > 0 ? nBits[17] : nBits[sum_pulses[i]];
    if (sumBits_Q5 < minSumBits_Q5) { minSumBits_Q5 = sumBits_Q5; RateLevelIndex = k; }
  }
  re_icdf(e, RateLevelIndex, &ob_silk_rate_levels_iCDF[9 * (signalType >> 1)], 8);
  for (i = 0; i < iter; i++) {
    const unsigned char *cdf = &ob_silk_pulses_per_block_iCDF[18 * RateLevelIndex], *last = &ob_silk_pulses_per_block_iCDF[18 * 9];
    if (nRshifts[i] == 0) re_icdf(e, sum_pulses[i], cdf, 8);
    else {
      re_icdf(e, 17, cdf, 8);
      for (k = 0; k < nRshifts[i] - 1; k++) re_icdf(e, 17, last, 8);
      re_icdf(e, sum_pulses[i], last, 8);
    }
  }
  for (i = 0; i < iter; i++)
    if (sum_pulses[i] > 0) shell_encoder(e, &abs_pulses[16 * i]);
  for (i = 0; i < iter; i++)
    if (nRshifts[i] > 0) {
      const int8_t *pp = &pulses[16 * i];
      const int nLS = nRshifts[i] - 1;
      for (k = 0; k < 16; k++) {
        const int abs_q = (int8_t)abs((int)pp[k]);
        for (j = nLS; j > 0; j--) re_icdf(e, (abs_q >> j) & 1, ob_silk_lsb_iCDF, 8);
        re_icdf(e, abs_q & 1, ob_silk_lsb_iCDF, 8);
      }
    }
  { /* silk_encode_signs, code_signs.c:40-72 */
    unsigned char icdf[2];
    const unsigned char *icdf_ptr = &ob_silk_sign_iCDF[7 * (quantOffsetType + (signalType << 1))];
    const int blocks = (frame_length + 8) >> 4;
    icdf[1] = 0;
    for (i = 0; i < blocks; i++) {
      const int p = sum_pulses[i];
      if (p > 0) {
        const int8_t *q = &pulses[16 * i];
        icdf[0] = icdf_ptr[(p & 0x1F) < 6 ? (p & 0x1F) : 6];
        for (j = 0; j < 16; j++)
          if (q[j] != 0) re_icdf(e, (q[j] >> 15) + 1, icdf, 8);   /* silk_enc_map: 0 for negative, 1 for positive */
      }
    }
  }
}

/* ---- silk_encode_indices (silk/encode_indices.c:35-185) with silk_NLSF_unpack (silk/NLSF_unpack.c:35-55): the side information of one SILK frame.
 * OrcSilkIndices has SideInfoIndices' fields in SideInfoIndices' order (silk/structs.h:129-141); prev = {ec_prevSignalType, ec_prevLagIndex} of the
 * encoder state, read and updated.  Table selection by fs_kHz / nb_subfr as silk_control_encoder does (silk/control_codec.c:212-296). ---- */
typedef struct {
  int8_t GainsIndices[4], LTPIndex[4], NLSFIndices[17];
  int16_t lagIndex;
  int8_t contourIndex, signalType, quantOffsetType, NLSFInterpCoef_Q2, PERIndex, LTP_scaleIndex, Seed;
} OrcSilkIndices;
int orc_silk_indices_size(void) { return (int)sizeof(OrcSilkIndices); }
void orc_silk_encode_indices(OrcSilkCoder *c, int fs_kHz, int nb_subfr, const OrcSilkIndices *ix, int condCoding, int *prev) {
  RangeEnc *e = &c->e;
  const int wb = fs_kHz == 16, order = wb ? 16 : 10, typeOffset = 2 * ix->signalType + ix->quantOffsetType;
  const unsigned char *cb1 = &ob_silk_NLSF_CB1_iCDF[wb ? 64 : 0], *sel = &ob_silk_NLSF_CB2_SELECT[wb ? 160 : 0], *cb2 = &ob_silk_NLSF_CB2_iCDF[wb ? 72 : 0];
  int i, k;
  if (typeOffset >= 2) re_icdf(e, typeOffset - 2, ob_silk_type_offset_VAD_iCDF, 8);
  else re_icdf(e, typeOffset, ob_silk_type_offset_no_VAD_iCDF, 8);
  if (condCoding == 2) re_icdf(e, ix->GainsIndices[0], ob_silk_delta_gain_iCDF, 8);
  else {
    re_icdf(e, ix->GainsIndices[0] >> 3, &ob_silk_gain_iCDF[8 * ix->signalType], 8);
    re_icdf(e, ix->GainsIndices[0] & 7, ob_silk_uniform8_iCDF, 8);
  }
  for (i = 1; i < nb_subfr; i++) re_icdf(e, ix->GainsIndices[i], ob_silk_delta_gain_iCDF, 8);
  re_icdf(e, ix->NLSFIndices[0], &cb1[(ix->signalType >> 1) * 32], 8);
  for (i = 0; i < order; i++) { /* ec_ix of silk_NLSF_unpack: the entropy table of coefficient i, three bits per coefficient in the select table */
    const int entry = sel[ix->NLSFIndices[0] * order / 2 + (i >> 1)];
    const int ec_ix = (((i & 1) ? entry >> 5 : entry >> 1) & 7) * 9;
    const int v = ix->NLSFIndices[i + 1];
    if (v >= 4) {
      re_icdf(e, 8, &cb2[ec_ix], 8);
      re_icdf(e, v - 4, ob_silk_NLSF_EXT_iCDF, 8);
    } else if (v <= -4) {
      re_icdf(e, 0, &cb2[ec_ix], 8);
      re_icdf(e, -v - 4, ob_silk_NLSF_EXT_iCDF, 8);
    } else re_icdf(e, v + 4, &cb2[ec_ix], 8);
  }
  if (nb_subfr == 4) re_icdf(e, ix->NLSFInterpCoef_Q2, ob_silk_NLSF_interpolation_factor_iCDF, 8);
  if (ix->signalType == 2) {
    int encode_absolute = 1;
    if (condCoding == 2 && prev[0] == 2) {
      int delta = ix->lagIndex - prev[1];
      if (delta < -8 || delta > 11) delta = 0;
      else { delta += 9; encode_absolute = 0; }
      re_icdf(e, delta, ob_silk_pitch_delta_iCDF, 8);
    }
    if (encode_absolute) {
      const int half = fs_kHz >> 1, hi = ix->lagIndex / half, lo = ix->lagIndex - hi * half;
      re_icdf(e, hi, ob_silk_pitch_lag_iCDF, 8);
      re_icdf(e, lo, fs_kHz == 16 ? ob_silk_uniform8_iCDF : fs_kHz == 12 ? ob_silk_uniform6_iCDF : ob_silk_uniform4_iCDF, 8);
    }
    prev[1] = ix->lagIndex;
    re_icdf(e, ix->contourIndex,
            fs_kHz == 8 ? (nb_subfr == 4 ? ob_silk_pitch_contour_NB_iCDF : ob_silk_pitch_contour_10_ms_NB_iCDF)
                        : (nb_subfr == 4 ? ob_silk_pitch_contour_iCDF : ob_silk_pitch_contour_10_ms_iCDF), 8);
    re_icdf(e, ix->PERIndex, ob_silk_LTP_per_index_iCDF, 8);
    for (k = 0; k < nb_subfr; k++) re_icdf(e, ix->LTPIndex[k], &ob_silk_LTP_gain_iCDF[ix->PERIndex == 0 ? 0 : ix->PERIndex == 1 ? 8 : 24], 8);
    if (condCoding == 0) re_icdf(e, ix->LTP_scaleIndex, ob_silk_LTPscale_iCDF, 8);
  }
  prev[0] = ix->signalType;
  re_icdf(e, ix->Seed, ob_silk_uniform4_iCDF, 8);
}
