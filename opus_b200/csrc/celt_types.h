// celt_types.h — plain-C layouts shared by the host side and the kernels (HBM resident).
#pragma once
#include <stdint.h>

#define OB_NBANDS 21
#define OB_OVERLAP 120
#define OB_SHORT_N 120
#define OB_MAXN 960
#define OB_MAX_PACKET 1275
#define OB_HIST 2048 /* comb-filter history ring per channel (>= 1024+2+960, power of two) */

/* Per-stream entropy/band-decoder state carried frame to frame (one record per stream).
 * Mirrors the persistent part of the reference's decoder state, celt/celt_decoder.c:87-139. */
typedef struct {
  float oldEBands[2 * OB_NBANDS];
  float oldLogE[2 * OB_NBANDS];
  float oldLogE2[2 * OB_NBANDS];
  float backgroundLogE[2 * OB_NBANDS];
  uint32_t rng;         /* final range of the last frame; also seeds the noise LCG */
  uint32_t final_range; /* what OPUS_GET_FINAL_RANGE reports */
  int32_t error;
  int32_t pf_period, pf_period_old, pf_tapset, pf_tapset_old;
  float pf_gain, pf_gain_old;
  int32_t last_status;
  int32_t pad[2];
} ObDecBandState; /* 720 bytes */

/* Per-stream synthesis state (one record per stream). */
typedef struct {
  float hist[2][OB_HIST];           /* post-filtered output ring, indexed by absolute sample position */
  float tail[2][OB_OVERLAP / 2];    /* last overlap/2 raw IMDCT outputs of the previous frame */
  float deemph_mem[2];
  uint32_t pos;                     /* absolute position of the next output sample (mod 2^32) */
  int32_t pad[1];
} ObDecSynthState;

/* What the band decoder hands to the synthesis kernel for one stream-frame. */
typedef struct {
  int32_t status;       /* >=0: samples per channel in this frame; <0: opus error code; 0: no frame */
  int32_t LM, transient, C, silence, end_band;
  int32_t pf_T[3], pf_tap[3]; /* post-filter period/tapset: [0]=old, [1]=current, [2]=new (this frame) */
  float pf_g[3];
  float bandE[2 * OB_NBANDS]; /* final log2 band energies (without the means) */
  int32_t pad[5];
} ObFrameSide; /* 240 bytes */
