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
  /* packet-loss concealment bookkeeping (celt/celt_decoder.c:110-124), decided in the entropy kernel; packed in two words so that a
     loss-free stream costs the hot path two loads and no store per frame */
  uint32_t plc_dur;   /* bits0-15 loss_duration, bits16-31 plc_duration (2.5 ms units, saturating at 10000) */
  uint32_t plc_flags; /* OB_PF_* */
} ObDecBandState;
#define OB_PF_SKIP_PLC 1u        /* noise PLC only, until two consecutive packets arrive */
#define OB_PF_FOLD 2u            /* prefilter_and_fold pending (the last frame was a pitch-based PLC frame) */
#define OB_PF_HAS_PACKET 4u      /* a packet has been decoded (Opus layer prev_mode != 0) */
#define OB_PF_TYPE_SHIFT 3       /* bits3-4: last_frame_type (OB_FRAME_*) */
#define OB_PF_END_SHIFT 5        /* bits5-9: end band of the last decoded packet (a lost frame keeps it) */
#define OB_PF_LM_SHIFT 10        /* bits10-11: LM of the last accepted packet (Opus layer st->frame_size) */
#define OB_FRAME_NONE 0
#define OB_FRAME_NORMAL 1
#define OB_FRAME_PLC_NOISE 2
#define OB_FRAME_PLC_PERIODIC 3
/* ObFrameSide.plc: bits0-1 what the frame is (0 decoded, 1 noise PLC, 2 pitch PLC, 3 zeros: nothing decoded yet),
   bit2 prefilter_and_fold before the synthesis, bit3 pitch PLC analyses (first lost frame of a burst), bit4 the whole packet is lost */
#define OB_PLC_TYPE(x) ((x) & 3)
#define OB_PLC_NOISE 1
#define OB_PLC_PITCH 2
#define OB_PLC_ZEROS 3
#define OB_PLC_FOLD 4
#define OB_PLC_FIRST 8
#define OB_PLC_PACKET 16 /* the whole packet is lost (len == 0): opus_decode() returns such frames without soft clipping */
#define OB_PLC_LPC_ORDER 24

/* Per-stream synthesis state (one record per stream). */
typedef struct {
  float hist[2][OB_HIST];           /* post-filtered output ring, indexed by absolute sample position */
  float tail[2][OB_OVERLAP];        /* [0,60): last overlap/2 raw IMDCT outputs of the previous frame; after a pitch-based PLC frame
                                       all 120 samples that it extrapolated past the frame (celt_decoder.c:948-952) */
  float deemph_mem[2];
  float declip[2];                  /* opus_pcm_soft_clip's declip_mem (int16 output only; Opus layer st->softclip_mem) */
  uint32_t pos;                     /* absolute position of the next output sample (mod 2^32) */
  int32_t last_pitch_index;         /* pitch-based PLC: period found at the first lost frame of the burst */
  float lpc[2][OB_PLC_LPC_ORDER];   /* pitch-based PLC: LPC of the last good signal */
} ObDecSynthState;

/* ---- entropy kernel -> PVQ kernel hand-off ("frame script"), one per stream ------------------ */
#define OB_MAX_LEAVES 448  /* worst case 416: every band of a stereo 20 ms frame split down to LM=-1 */

/* One leaf partition of the band tree (celt/bands.c quant_partition no-split case). 16 bytes.
 * `path` describes the splits between the band part's root and this leaf, root first, 3 bits per
 * split: bits0-1 theta class (0: itheta==0, 1: itheta==16384, 2: other), bit2 set if the leaf lies in
 * the second (Y) half.  It lets any lane derive the leaf's `fill` mask and the position of its
 * collapse mask in the part's mask without walking the tree (celt/bands.c:1010-1012, 1121). */
typedef struct {
  uint32_t index;  /* kind 0: PVQ codeword index (ec_dec_uint result); kind 2: sign bits (bit c = channel c) */
  float gain;      /* product of the mid/side gains down the tree (celt/bands.c:1078-1089) */
  uint16_t off;    /* offset in the frame's X[C][N] array */
  uint8_t n;       /* dimensions (2..176; 1 for kind 2) */
  uint8_t K;       /* pulses (kind 0); channels (kind 2) */
  uint16_t path;   /* bits0-11 split path, bits12-14 depth (number of splits), bit15 lowband present */
  uint8_t B_kind;  /* bits0-4 blocks at this leaf, bits5-6 kind: 0 PVQ pulses, 1 no pulses, 2 N==1 signs */
  uint8_t cmshift; /* left shift of this leaf's collapse mask inside the part's mask */
} ObLeaf;

/* Per band: what the vector stage cannot derive without the range decoder. 24 bytes. */
typedef struct {
  int16_t eff_lowband;   /* effective_lowband (offset into norm) or -1 */
  uint8_t lowband_offset;
  uint8_t flags;         /* bit0 dual stereo band, bit1 dual->intensity switch happens at this band,
                            bit2 stereo split band, bit3 inv, bit4 N==2 special, bit5 n2 c, bit6 n2 sign(neg),
                            bit7 side decoded first */
  uint8_t theta_class;   /* stereo theta: 0 -> itheta==0, 1 -> itheta==16384, 2 -> other */
  uint8_t pad;
  uint16_t leaf_begin;   /* first leaf of this band */
  uint16_t leaf_mid;     /* end of the first decoded part's leaves (= leaf_end when the band has one part) */
  uint16_t leaf_end;
  float mid, side;       /* stereo mid/side gains (imid/32768, iside/32768) */
  int32_t b;             /* bits given to the band (debug / update_lowband) */
} ObBandRec;

/* What the band decoder hands to the synthesis kernel for one stream-frame. */
typedef struct {
  int32_t status;       /* >=0: samples per channel in this frame; <0: opus error code; 0: no frame */
  int32_t LM, transient, C, silence, end_band;
  int32_t pf_T[3], pf_tap[3]; /* post-filter period/tapset: [0]=old, [1]=current, [2]=new (this frame) */
  float pf_g[3];
  float bandE[2 * OB_NBANDS]; /* final log2 band energies (without the means) */
  /* entropy -> PVQ stage */
  int32_t spread, intensity, dual_stereo, anti_collapse_on, n_leaves, plc; /* plc: OB_PLC_* */
  uint32_t seed;              /* LCG seed at the start of the band loop (previous frame's final range) */
  int8_t tf_res[OB_NBANDS];
  uint8_t pad8[3];
  int16_t pulses[OB_NBANDS];  /* PVQ bit allocation per band (anti-collapse depth) */
  int16_t pad16;
  int32_t pcm_off;            /* per-channel offset of this frame in the stream's pcm area, output-rate samples */
  /* oldLogE / oldLogE2 as they were BEFORE this frame's state update (written only when anti_collapse_on): the entropy kernel
     owns the energy state, so that the next frame's parsing does not wait for this frame's vector stage */
  float prevLogE[2][2 * OB_NBANDS];
} ObFrameSide; /* 664 bytes */

/* Per-stream script buffer written by the entropy kernel, read by the PVQ kernel. */
typedef struct {
  ObLeaf leaves[OB_MAX_LEAVES];
  ObBandRec bands[OB_NBANDS];
  uint8_t pad[8]; /* keeps sizeof(ObScript) a multiple of 16 */
} ObScript;
