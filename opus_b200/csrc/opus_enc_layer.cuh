// opus_enc_layer.cuh — the Opus layer of the CELT-only encode path on the device: what opus_encode() does around
// celt_encode_with_ec() for an encoder created with OPUS_APPLICATION_RESTRICTED_CELT / RESTRICTED_LOWDELAY
// (src/opus_encoder.c:1182-1853 opus_encode_native, :1855-2657 opus_encode_frame_native, CELT-only subset):
//
//   kernel P1  ob_opus_enc_pre       one THREAD per stream: every per-sample recurrence of the layer in one pass over the frame —
//                                    dc_reject (:510-550), the float API's NaN / level guard (:2011-2022), is_digital_silence
//                                    (:1060-1077), and the analysis input: downmix (:748-802) + 48 -> 24 kHz all-pass
//                                    resampler with its high-band energy (src/analysis.c:115-214).  Streams of a warp are staged
//                                    through transposed shared-memory tiles, so HBM sees coalesced rows and each thread a private column.
//   kernel P2  ob_opus_enc_analysis  one WARP per stream: tonality_analysis (src/analysis.c:445-952: 480-point FFT of the windowed
//                                    24 kHz signal, per-bin phase tracking, band energies / tonality / stationarity, bandwidth
//                                    detection, the 25 -> 32 -> GRU 24 -> 2 network of src/mlp.c), tonality_get_info (:232-408),
//                                    then the layer's per-stream decisions on one lane: voice estimate, mono/stereo switch with
//                                    hysteresis, automatic + detected bandwidth, stereo width fade gains, TOC byte.
//   The CELT kernels read the result (ObOpusFrameCtl: coded channels, end band, AnalysisInfo, fade gains) per stream.
// Float expressions keep the reference's operand order and double-precision detours (compiled with -fmad=false).
#pragma once
#include "celt_fft.cuh"
#include "analysis_tables_data.h"

#ifndef OB_IMIN
#define OB_IMIN(a, b) ((a) < (b) ? (a) : (b))
#define OB_IMAX(a, b) ((a) > (b) ? (a) : (b))
#endif

namespace ob {

#define OB_AN_NB_FRAMES 8
#define OB_AN_NB_TBANDS 18
#define OB_AN_BUF 720
#define OB_AN_DETECT 100
#define OB_AN_SKIP 9

// AnalysisInfo (celt/celt.h:65-79)
struct ObAnalysisInfo {
  int32_t valid;
  float tonality, tonality_slope, noisiness, activity, music_prob, music_prob_min, music_prob_max;
  int32_t bandwidth;
  float activity_probability, max_pitch_ratio;
  unsigned char leak_boost[20];   // LEAK_BANDS 19 (+1 pad)
};
// TonalityAnalysisState (src/analysis.h:47-80), one per stream in HBM
struct ObTonality {
  float angle[240], d_angle[240], d2_angle[240];
  float inmem[OB_AN_BUF];
  float prev_band_tonality[OB_AN_NB_TBANDS];
  float E[OB_AN_NB_FRAMES][OB_AN_NB_TBANDS];
  float logE[OB_AN_NB_FRAMES][OB_AN_NB_TBANDS];
  float lowE[OB_AN_NB_TBANDS], highE[OB_AN_NB_TBANDS], meanE[OB_AN_NB_TBANDS + 1];
  float mem[32], cmean[8], std[9];
  float rnn_state[32];
  float downmix_state[3];
  float prev_tonality, Etracker, lowECount, hp_ener_accum;
  int32_t mem_fill, prev_bandwidth, E_count, count, write_pos, read_pos, read_subframe, initialized;
  ObAnalysisInfo info[OB_AN_DETECT];
};
// the Opus layer's own per-stream state (OpusEncoder, src/opus_encoder.c:74-145, what the CELT-only path reads and writes)
struct ObOpusEncState {
  int32_t stream_channels, bandwidth, auto_bandwidth, voice_ratio, first, prev_mode, hybrid_stereo_width_Q14, pad;
  float hp_mem[4];
};
// P1 -> P2 hand-off
struct ObOpusPre {
  float newsamp[480];   // this frame's 24 kHz analysis samples (at most 20 ms)
  float hp1, hp2;       // high-band energy of the samples that complete the 30 ms buffer / of the rest (downmix_and_resample's two calls)
  int32_t silence;      // is_digital_silence of the frame
  int32_t clear;        // the float guard fired: the frame is coded as zeros
};
// per stream and frame: what the layer tells the CELT kernels
struct ObOpusFrameCtl {
  int32_t C, end, toc, fade;   // coded channels (CELT_SET_CHANNELS), end band (CELT_SET_END_BAND), TOC byte, stereo_fade on
  float fade_g1, fade_g2;      // 1 - previous / current stereo width
  int32_t clear, pad;
  ObAnalysisInfo analysis;     // CELT_SET_ANALYSIS
};
// encoder-wide settings (opus_encoder_ctl) + the per-call quantities that are the same for every stream
struct OpusLayerParams {
  int Fs, channels, application, force_channels, signal_type, user_bandwidth, max_bandwidth, use_vbr, complexity, packet_loss, lsb_depth;
  int frame_size;          // samples per channel at Fs
  int32_t bitrate_bps;     // after user_bitrate_to_bitrate and the CBR rounding (:1325-1334)
  int run_analysis;        // complexity >= 7 && Fs >= 16000 (:1252)
  int n_streams;
  ObOpusEncState *st;
  ObTonality *tonal;
  ObOpusPre *pre;
  ObOpusFrameCtl *ctl;
  const void *pcm;         // [n][frame_size*channels] float (+-1.0) or int16
  int pcm_is_int16;
  float *pcm_buf;          // [n][frame_size*channels] out: dc-rejected input of the CELT encoder
  unsigned char *packets;  // [n][pkt_stride]: byte 0 receives the TOC
  int pkt_stride;
  // 40..120 ms packets (src/opus_encoder.c:1698-1838): the 20 ms frames go through P1 / P2 one by one with `multi` set — P1 reads its frame at
  // pcm_off inside rows of pcm_row samples, P2 runs the tonality analysis of the chunk and leaves get_info / the decisions to opus_multiframe_decide_thread
  int multi = 0, pcm_row = 0, pcm_off = 0;
  int nb_frames = 1, packet_frame_size = 0;   // decide kernel: the packet's frame count and whole size (p.frame_size is the 20 ms frame then)
  ObOpusPre *pre_all = nullptr;               // [nb_frames][n]
  ObOpusFrameCtl *ctl_all = nullptr;          // [nb_frames][n]
  unsigned char *tmp_packets = nullptr;       // [nb_frames][n][tmp_stride]: the one-frame packets (TOC + payload)
  int tmp_stride = 0;
  const int32_t *tmp_nbytes = nullptr;        // [nb_frames][n]: bytes of every one-frame packet (TOC included), negative = error
  int32_t *out_nbytes = nullptr;              // [n]: the repacketised size
  int repacketize_len = 0, pad = 0;
};

// ---------------------------------------------------------------------------------------------------------------------------
// Host side of a call: the part of opus_encode_native that is the same for every stream of the batch (settings and the byte budget)
// ---------------------------------------------------------------------------------------------------------------------------
struct OpusEncSettings {   // opus_encoder_ctl state (OpusEncoder, src/opus_encoder.c:74-110)
  int Fs, channels, application, force_channels, signal_type, user_bandwidth, max_bandwidth, use_vbr, vbr_constraint, user_bitrate_bps, lsb_depth,
      complexity, packet_loss, prediction_disabled;
};
struct OpusCallPlan {
  int nb_frames, enc_frame_size;         // 40..120 ms: nb_frames 20 ms frames, repacketised (:1698-1838); else 1 / frame_size
  int32_t repacketize_len, curr_max;     // multi-frame: the packet's size bound, and every frame's (TOC included; the third bound of :1795 never binds)
  int32_t bitrate_bps, max_data_bytes;   // st->bitrate_bps and the packet size bound after the CBR rounding (:1325-1334), TOC included
  int run_analysis, lsb;
  int celt_vbr, celt_bitrate, celt_cvbr, celt_disable_pf, celt_force_intra;   // what opus_encode_frame_native sets on the CELT encoder (:2284-2294, :2445-2463)
};
inline void opus_enc_settings_init(OpusEncSettings *e, int Fs, int channels, int application) {   // opus_encoder_init defaults (:257-320)
  e->Fs = Fs; e->channels = channels; e->application = application;
  e->force_channels = -1000; e->signal_type = -1000; e->user_bandwidth = -1000; e->max_bandwidth = 1105; e->use_vbr = 1; e->vbr_constraint = 1;
  e->user_bitrate_bps = -1000; e->lsb_depth = 24; e->complexity = 9; e->packet_loss = 0; e->prediction_disabled = 0;
}
// opus_encoder_ctl: the integer setters of the CELT-only path (src/opus_encoder.c:2783-3290; request numbers of include/opus_defines.h).
// Returns 0, -1 (OPUS_BAD_ARG) or -5 (not a request handled here)
inline int opus_enc_settings_ctl(OpusEncSettings *e, int request, int32_t value) {
  switch (request) {
    case 4002:   // OPUS_SET_BITRATE (:2811-2825)
      if (value != -1000 && value != -1) {
        if (value <= 0) return -1;
        else if (value <= 500) value = 500;
        else if (value > (int32_t)750000 * e->channels) value = (int32_t)750000 * e->channels;
      }
      e->user_bitrate_bps = value; return 0;
    case 4022: if ((value < 1 || value > e->channels) && value != -1000) return -1; e->force_channels = value; return 0;   // OPUS_SET_FORCE_CHANNELS
    case 4004: if (value < 1101 || value > 1105) return -1; e->max_bandwidth = value; return 0;                               // OPUS_SET_MAX_BANDWIDTH
    case 4008: if ((value < 1101 || value > 1105) && value != -1000) return -1; e->user_bandwidth = value; return 0;            // OPUS_SET_BANDWIDTH
    case 4010: if (value < 0 || value > 10) return -1; e->complexity = value; return 0;                                          // OPUS_SET_COMPLEXITY
    case 4014: if (value < 0 || value > 100) return -1; e->packet_loss = value; return 0;                                        // OPUS_SET_PACKET_LOSS_PERC
    case 4006: if (value < 0 || value > 1) return -1; e->use_vbr = value; return 0;                                              // OPUS_SET_VBR
    case 4020: if (value < 0 || value > 1) return -1; e->vbr_constraint = value; return 0;                                       // OPUS_SET_VBR_CONSTRAINT
    case 4024: if (value != -1000 && value != 3001 && value != 3002) return -1; e->signal_type = value; return 0;                // OPUS_SET_SIGNAL
    case 4036: if (value < 8 || value > 24) return -1; e->lsb_depth = value; return 0;                                           // OPUS_SET_LSB_DEPTH
    case 4042: if (value < 0 || value > 1) return -1; e->prediction_disabled = value; return 0;                                  // OPUS_SET_PREDICTION_DISABLED
    default: return -5;
  }
}
inline int32_t opus_bits_to_bitrate(int32_t bits, int32_t Fs, int32_t frame_size) { return bits * (6 * Fs / frame_size) / 6; }   // celt/celt.h:146
inline int32_t opus_bitrate_to_bits(int32_t bitrate, int32_t Fs, int32_t frame_size) { return bitrate * 6 / (6 * Fs / frame_size); }
// 0, or the negative Opus code the call returns for every stream
inline int opus_enc_plan(const OpusEncSettings &e, int frame_size, int max_data_bytes_in, int pcm_is_int16, OpusCallPlan *o) {
  const int Fs = e.Fs;
  int32_t max_data_bytes = max_data_bytes_in < 1276 * 6 ? max_data_bytes_in : 1276 * 6;
  if (frame_size <= 0 || max_data_bytes <= 0) return -1;
  // frame_size_select with OPUS_FRAMESIZE_ARG (src/opus_encoder.c:827-851)
  if (frame_size < Fs / 400) return -1;
  if (400 * frame_size != Fs && 200 * frame_size != Fs && 100 * frame_size != Fs && 50 * frame_size != Fs && 25 * frame_size != Fs &&
      50 * frame_size != 3 * Fs && 50 * frame_size != 4 * Fs && 50 * frame_size != 5 * Fs && 50 * frame_size != 6 * Fs)
    return -1;
  const int frame_rate = Fs / frame_size;
  int32_t bitrate_bps, cbr_bytes = -1;
  {   // user_bitrate_to_bitrate (:733-745)
    const int32_t max_bitrate = opus_bits_to_bitrate(max_data_bytes * 8, Fs, frame_size);
    const int32_t user = e.user_bitrate_bps == -1000 ? 60 * Fs / frame_size + Fs * e.channels : e.user_bitrate_bps == -1 ? 1500000 : e.user_bitrate_bps;
    bitrate_bps = user < max_bitrate ? user : max_bitrate;
  }
  if (!e.use_vbr) {
    cbr_bytes = (opus_bitrate_to_bits(bitrate_bps, Fs, frame_size) + 4) / 8;
    if (cbr_bytes > max_data_bytes) cbr_bytes = max_data_bytes;
    bitrate_bps = opus_bits_to_bitrate(cbr_bytes * 8, Fs, frame_size);
    max_data_bytes = cbr_bytes > 1 ? cbr_bytes : 1;
  }
  if (max_data_bytes < 3 || bitrate_bps < 3 * frame_rate * 8 || (frame_rate < 50 && (max_data_bytes * (int32_t)frame_rate < 300 || bitrate_bps < 2400)))
    return -5;   // the 'PLC frame' packets of a budget too small to code anything (:1340-1406) are not built
  o->max_data_bytes = max_data_bytes < 1276 ? max_data_bytes : 1276;
  o->nb_frames = 1; o->enc_frame_size = frame_size; o->repacketize_len = o->curr_max = 0;
  if (frame_size > Fs / 50) {   // :1698-1795
    o->enc_frame_size = Fs / 50;
    o->nb_frames = frame_size / o->enc_frame_size;
    const int max_header_bytes = o->nb_frames == 2 ? 3 : (2 + (o->nb_frames - 1) * 2);
    o->repacketize_len = (e.use_vbr || e.user_bitrate_bps == -1) ? max_data_bytes_in : (cbr_bytes < max_data_bytes_in ? cbr_bytes : max_data_bytes_in);
    const int32_t max_len_sum = o->nb_frames + o->repacketize_len - max_header_bytes, a = opus_bitrate_to_bits(bitrate_bps, Fs, o->enc_frame_size) / 8,
                  b = max_len_sum / o->nb_frames;
    o->curr_max = a < b ? a : b;
    if (o->curr_max < 2) return -5;
    o->max_data_bytes = o->curr_max < 1276 ? o->curr_max : 1276;   // what opus_encode_frame_native gets for every 20 ms frame
  }
  o->bitrate_bps = bitrate_bps;
  o->run_analysis = e.complexity >= 7 && Fs >= 16000;   // :1252
  o->lsb = (pcm_is_int16 ? 16 : 24) < e.lsb_depth ? (pcm_is_int16 ? 16 : 24) : e.lsb_depth;
  o->celt_disable_pf = o->celt_force_intra = e.prediction_disabled ? 1 : 0;   // CELT_SET_PREDICTION(0 / 2)
  o->celt_vbr = e.use_vbr; o->celt_bitrate = -1; o->celt_cvbr = 1;
  if (e.use_vbr) {
    o->celt_cvbr = e.vbr_constraint;
    if (bitrate_bps > 500) o->celt_bitrate = bitrate_bps < 750000 * e.channels ? bitrate_bps : 750000 * e.channels;   // celt_encoder_ctl (celt/celt_encoder.c:3020-3028)
  }
  return 0;
}
inline void opus_layer_params_init(OpusLayerParams *p, const OpusEncSettings &e, const OpusCallPlan &pl, int frame_size, int pcm_is_int16, int n_streams) {
  p->Fs = e.Fs; p->channels = e.channels; p->application = e.application; p->force_channels = e.force_channels; p->signal_type = e.signal_type;
  p->user_bandwidth = e.user_bandwidth; p->max_bandwidth = e.max_bandwidth; p->use_vbr = e.use_vbr; p->complexity = e.complexity; p->packet_loss = e.packet_loss;
  p->lsb_depth = e.lsb_depth; p->frame_size = frame_size; p->bitrate_bps = pl.bitrate_bps; p->run_analysis = pl.run_analysis; p->n_streams = n_streams;
  p->pcm_is_int16 = pcm_is_int16;
}
// opus_encoder_init's part of the per-stream state (src/opus_encoder.c:272-320); everything else is zero
OB_HD void opus_enc_state_init(ObOpusEncState *st, int channels) {
  st->stream_channels = channels; st->bandwidth = 1105; st->auto_bandwidth = 0; st->voice_ratio = -1; st->first = 1; st->prev_mode = 0;
  st->hybrid_stereo_width_Q14 = 1 << 14; st->pad = 0;
  st->hp_mem[0] = st->hp_mem[1] = st->hp_mem[2] = st->hp_mem[3] = 0.f;
}

// ---------------------------------------------------------------------------------------------------------------------------
// P1: one thread per stream.  tile: [32][TS] floats per warp (TS odd: a thread reads its own row without bank conflicts)
// ---------------------------------------------------------------------------------------------------------------------------
#define OB_PRE_TS 65   // 64 samples (32 stereo pairs) per stream and step, +1 padding
OB_DEV float pre_load_sample(const void *pcm, int is16, size_t idx) {
  return is16 ? (float)((const int16_t *)pcm)[idx] : ((const float *)pcm)[idx];
}
OB_DEV void opus_enc_pre_warp(const OpusLayerParams &p, int first_stream, float *tile) {
  const int lane = lane_id();
  const int s = first_stream + lane;
  const bool live = s < p.n_streams;
  const int CC = p.channels, fs = p.frame_size, total = fs * CC;   // floats per stream
  ObOpusEncState *st = live ? p.st + s : nullptr;
  ObTonality *tn = (live && p.run_analysis) ? p.tonal + s : nullptr;
  ObOpusPre *pre = live ? p.pre + s : nullptr;
  const float coef = 6.3f * 3 / p.Fs, coef2 = 1 - coef;
  float m0 = live ? st->hp_mem[0] : 0.f, m2 = live ? st->hp_mem[2] : 0.f;
  float S0 = 0.f, S1 = 0.f, S2 = 0.f, hp1 = 0.f, hp2 = 0.f;
  int n1 = 0;   // 24 kHz samples that go to the first call of downmix_and_resample (they complete the 30 ms buffer)
  if (tn) {
    S0 = tn->downmix_state[0]; S1 = tn->downmix_state[1]; S2 = tn->downmix_state[2];
    const int mem_fill = tn->initialized ? tn->mem_fill : 240;
    const int len = p.Fs == 48000 ? fs / 2 : p.Fs == 16000 ? 3 * fs / 2 : fs;
    n1 = OB_IMIN(len, OB_AN_BUF - mem_fill);
  }
  float mx = 0.f, mn = 0.f, guard = 0.f;
  // analysis resampler bookkeeping: `pend` holds the first sample of a pair (48 kHz: consecutive inputs; 16 kHz: the input tripled)
  float pend = 0.f;
  int have_pend = 0, nout = 0;
  const int is16 = p.pcm_is_int16;
  for (int base = 0; base < total; base += 64) {
    const int cnt = OB_IMIN(64, total - base);
    // coalesced load: row r = stream first_stream + r
    for (int r = 0; r < 32; r++) {
      const int sr = first_stream + r;
      if (sr >= p.n_streams) break;
      for (int j = lane; j < cnt; j += 32) tile[r * OB_PRE_TS + j] = pre_load_sample(p.pcm, is16, (size_t)sr * (p.pcm_row ? p.pcm_row : total) + p.pcm_off + base + j);
    }
    wsync();
    if (live) {
      float *row = tile + lane * OB_PRE_TS;
      for (int j = 0; j < cnt; j += CC) {
        // the sample as opus_encode() hands it on (int16 API: x / 32768, src/opus_encoder.c:2687-2688)
        const float r0 = row[j], r1 = CC == 2 ? row[j + 1] : 0.f;
        const float x0 = is16 ? r0 * (1 / 32768.f) : r0, x1 = is16 ? r1 * (1 / 32768.f) : r1;
        mx = fmaxf(mx, x0); mn = fminf(mn, x0);
        if (CC == 2) { mx = fmaxf(mx, x1); mn = fminf(mn, x1); }
        // dc_reject (src/opus_encoder.c:510-550)
        const float o0 = x0 - m0;
        m0 = coef * x0 + 1e-30f + coef2 * m0;
        guard = guard + o0 * o0;
        row[j] = o0;
        if (CC == 2) {
          const float o1 = x1 - m2;
          m2 = coef * x1 + 1e-30f + coef2 * m2;
          guard = guard + o1 * o1;
          row[j + 1] = o1;
        }
        if (tn) {
          // downmix_int / downmix_float with c1 = 0, c2 = -2 (:748-802), then the halving of downmix_and_resample (src/analysis.c:185-189)
          float y;
          if (is16) { y = r0; if (CC == 2) y += r1; }
          else {
            y = r0 * 32768.f;
            if (CC == 2) y += r1 * 32768.f;
            if (y < -65536.f) y = -65536.f;
            if (y > 65536.f) y = 65536.f;
            if (y != y) y = 0;
          }
          if (CC == 2) y = .5f * y;
          if (p.Fs == 24000) {
            pre->newsamp[nout++] = y;
          } else {
            // 48 kHz: every input sample feeds the 2:1 all-pass pair once; 16 kHz: three times (the 3x zero-order hold)
            const int reps = p.Fs == 16000 ? 3 : 1;
            for (int k = 0; k < reps; k++) {
              if (!have_pend) { pend = y; have_pend = 1; continue; }
              have_pend = 0;
              // silk_resampler_down2_hp (src/analysis.c:115-163)
              float in32 = pend, Y = in32 - S0, X = 0.6074371f * Y, out32 = S0 + X;
              S0 = in32 + X;
              float out32_hp = out32;
              in32 = y;
              Y = in32 - S1;
              X = 0.15063f * Y;
              out32 = out32 + S1;
              out32 = out32 + X;
              S1 = in32 + X;
              Y = -in32 - S2;
              X = 0.15063f * Y;
              out32_hp = out32_hp + S2;
              out32_hp = out32_hp + X;
              S2 = -in32 + X;
              if (p.Fs == 48000) { if (nout < n1) hp1 += out32_hp * out32_hp; else hp2 += out32_hp * out32_hp; }
              pre->newsamp[nout++] = .5f * out32;
            }
          }
        }
      }
    }
    wsync();
    for (int r = 0; r < 32; r++) {
      const int sr = first_stream + r;
      if (sr >= p.n_streams) break;
      for (int j = lane; j < cnt; j += 32) p.pcm_buf[(size_t)sr * total + base + j] = tile[r * OB_PRE_TS + j];
    }
    wsync();
  }
  if (!live) return;
  const int lsb = p.pcm_is_int16 ? OB_IMIN(16, p.lsb_depth) : OB_IMIN(24, p.lsb_depth);
  pre->silence = fmaxf(mx, -mn) <= 1.f / (float)(1 << lsb);
  const int clear = !(guard < 1e9f) || guard != guard;   // :2017-2021
  pre->clear = clear;
  if (clear) { m0 = 0.f; m2 = 0.f; st->hp_mem[1] = 0.f; st->hp_mem[3] = 0.f; }
  st->hp_mem[0] = m0;
  st->hp_mem[2] = m2;
  if (tn) {
    tn->downmix_state[0] = S0; tn->downmix_state[1] = S1; tn->downmix_state[2] = S2;
    pre->hp1 = hp1 * (1.f / 32768 / 32768);
    pre->hp2 = hp2 * (1.f / 32768 / 32768);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// src/mlp.c on a warp: one lane per neuron, the inputs in shared memory; every accumulation in the reference's order
// ---------------------------------------------------------------------------------------------------------------------------
OB_DEV float an_tansig(float x) {   // src/mlp.c:39-54
  const float N0 = 952.52801514f, N1 = 96.39235687f, N2 = 0.60863042f, D0 = 952.72399902f, D1 = 413.36801147f, D2 = 11.88600922f;
  const float X2 = x * x;
  float num = (N2 * X2 + N1) * X2 + N0;
  const float den = (D2 * X2 + D1) * X2 + D0;
  num = num * x / den;
  return fmaxf(-1.f, fminf(1.f, num));
}
OB_DEV float an_sigmoid(float x) { return .5f + .5f * an_tansig(.5f * x); }
OB_DEV float an_gemm1(float acc, const signed char *w, int i, int cols, int col_stride, const float *x) {
  for (int j = 0; j < cols; j++) acc += w[j * col_stride + i] * x[j];
  return acc;
}
OB_DEV float an_fast_atan2f(float y, float x) {   // celt/mathops.h:56-77
  const float cA = 0.43157974f, cB = 0.67848403f, cC = 0.08595542f, cE = (float)3.14159265358979323846 / 2;
  const float x2 = x * x, y2 = y * y;
  if (x2 + y2 < 1e-18f) return 0;
  if (x2 < y2) {
    const float den = (y2 + cB * x2) * (y2 + cC * x2);
    return -x * y * (y2 + cA * x2) / den + (y < 0 ? -cE : cE);
  } else {
    const float den = (x2 + cB * y2) * (x2 + cC * y2);
    return x * y * (x2 + cA * y2) / den + (y < 0 ? -cE : cE) - (x * y < 0 ? -cE : cE);
  }
}
OB_DEV float an_lrintf(float x) {
#if defined(__CUDACC__)
  return (float)__float2int_rn(x);
#else
  return (float)lrintf(x);
#endif
}

// shared-memory workspace of one warp
struct AnWS {
  cpx out[480];                 // FFT in / out
  float tonality[240], noisiness[240], tonality2[240];
  float bE[OB_AN_NB_TBANDS], btE[OB_AN_NB_TBANDS], bnE[OB_AN_NB_TBANDS], bE2[OB_AN_NB_TBANDS], blogE[OB_AN_NB_TBANDS], bstat[OB_AN_NB_TBANDS],
      btone[OB_AN_NB_TBANDS], brel[OB_AN_NB_TBANDS];
  float mind[OB_AN_NB_FRAMES][OB_AN_NB_FRAMES];
  float feat[32], l0[32], gz[32], gr[32], gt[32];
  float E0;                     // the DC band's energy (band_log2[0])
  int32_t flag;
};

// tonality_get_info (src/analysis.c:232-408), one lane
OB_DEVNI void an_get_info(ObTonality *tonal, ObAnalysisInfo *info_out, int len, int Fs) {
  int pos = tonal->read_pos;
  int curr_lookahead = tonal->write_pos - tonal->read_pos;
  if (curr_lookahead < 0) curr_lookahead += OB_AN_DETECT;
  tonal->read_subframe += len / (Fs / 400);
  while (tonal->read_subframe >= 8) { tonal->read_subframe -= 8; tonal->read_pos++; }
  if (tonal->read_pos >= OB_AN_DETECT) tonal->read_pos -= OB_AN_DETECT;
  if (len > Fs / 50 && pos != tonal->write_pos) { pos++; if (pos == OB_AN_DETECT) pos = 0; }
  if (pos == tonal->write_pos) pos--;
  if (pos < 0) pos = OB_AN_DETECT - 1;
  const int pos0 = pos;
  *info_out = tonal->info[pos];
  if (!info_out->valid) return;
  float tonality_max = info_out->tonality, tonality_avg = info_out->tonality;
  int tonality_count = 1, bandwidth_span = 6;
  for (int i = 0; i < 3; i++) {
    pos++;
    if (pos == OB_AN_DETECT) pos = 0;
    if (pos == tonal->write_pos) break;
    tonality_max = fmaxf(tonality_max, tonal->info[pos].tonality);
    tonality_avg += tonal->info[pos].tonality;
    tonality_count++;
    info_out->bandwidth = OB_IMAX(info_out->bandwidth, tonal->info[pos].bandwidth);
    bandwidth_span--;
  }
  pos = pos0;
  for (int i = 0; i < bandwidth_span; i++) {
    pos--;
    if (pos < 0) pos = OB_AN_DETECT - 1;
    if (pos == tonal->write_pos) break;
    info_out->bandwidth = OB_IMAX(info_out->bandwidth, tonal->info[pos].bandwidth);
  }
  info_out->tonality = fmaxf(tonality_avg / tonality_count, tonality_max - .2f);
  int mpos = pos0, vpos = pos0;
  if (curr_lookahead > 15) {
    mpos += 5;
    if (mpos >= OB_AN_DETECT) mpos -= OB_AN_DETECT;
    vpos += 1;
    if (vpos >= OB_AN_DETECT) vpos -= OB_AN_DETECT;
  }
  float prob_min = 1.f, prob_max = 0.f;
  const float vad_prob = tonal->info[vpos].activity_probability;
  float prob_count = fmaxf(.1f, vad_prob);
  float prob_avg = fmaxf(.1f, vad_prob) * tonal->info[mpos].music_prob;
  while (1) {
    mpos++;
    if (mpos == OB_AN_DETECT) mpos = 0;
    if (mpos == tonal->write_pos) break;
    vpos++;
    if (vpos == OB_AN_DETECT) vpos = 0;
    if (vpos == tonal->write_pos) break;
    const float pos_vad = tonal->info[vpos].activity_probability;
    prob_min = fminf((prob_avg - 10 * (vad_prob - pos_vad)) / prob_count, prob_min);
    prob_max = fmaxf((prob_avg + 10 * (vad_prob - pos_vad)) / prob_count, prob_max);
    prob_count += fmaxf(.1f, pos_vad);
    prob_avg += fmaxf(.1f, pos_vad) * tonal->info[mpos].music_prob;
  }
  info_out->music_prob = prob_avg / prob_count;
  prob_min = fminf(prob_avg / prob_count, prob_min);
  prob_max = fmaxf(prob_avg / prob_count, prob_max);
  prob_min = fmaxf(prob_min, 0.f);
  prob_max = fminf(prob_max, 1.f);
  if (curr_lookahead < 10) {
    float pmin = prob_min, pmax = prob_max;
    pos = pos0;
    for (int i = 0; i < OB_IMIN(tonal->count - 1, 15); i++) {
      pos--;
      if (pos < 0) pos = OB_AN_DETECT - 1;
      pmin = fminf(pmin, tonal->info[pos].music_prob);
      pmax = fmaxf(pmax, tonal->info[pos].music_prob);
    }
    pmin = fmaxf(0.f, pmin - .1f * vad_prob);
    pmax = fminf(1.f, pmax + .1f * vad_prob);
    prob_min += (1.f - .1f * curr_lookahead) * (pmin - prob_min);
    prob_max += (1.f - .1f * curr_lookahead) * (pmax - prob_max);
  }
  info_out->music_prob_min = prob_min;
  info_out->music_prob_max = prob_max;
}

// the per-band / per-frame scalar tail of tonality_analysis (src/analysis.c:614-950), one lane; band sums come from AnWS
OB_DEVNI void an_frame_tail(ObTonality *tonal, ObAnalysisInfo *info, AnWS *w, float hp_ener, int lsb_depth, int Fs, float alpha, float alphaE,
                            float alphaE2) {
  const short *tbands = ob_an_tbands;
  float band_log2[OB_AN_NB_TBANDS + 1], leakage_from[OB_AN_NB_TBANDS + 1], leakage_to[OB_AN_NB_TBANDS + 1], logE[OB_AN_NB_TBANDS], BFCC[8], midE[8],
      features[25];
  int is_masked[OB_AN_NB_TBANDS + 1];
  float frame_tonality = 0, max_frame_tonality = 0, frame_noisiness = 0, frame_stationarity = 0, relativeE = 0, frame_loudness = 0, slope = 0;
  int b;
  info->activity = 0;
  band_log2[0] = .5f * 1.442695f * (float)log((double)(w->E0 + 1e-10f));
  for (b = 0; b < OB_AN_NB_TBANDS; b++) {
    const float E = w->bE[b];
    frame_noisiness += w->bnE[b] / (1e-15f + E);
    frame_loudness += (float)sqrt((double)(E + 1e-10f));
    logE[b] = w->blogE[b];
    band_log2[b + 1] = .5f * 1.442695f * logE[b];
    relativeE += w->brel[b];
    frame_stationarity += w->bstat[b];
    const float bt = w->btone[b];
    frame_tonality += bt;
    if (b >= OB_AN_NB_TBANDS - OB_AN_SKIP) frame_tonality -= w->btone[b - OB_AN_NB_TBANDS + OB_AN_SKIP];
    max_frame_tonality = fmaxf(max_frame_tonality, (1.f + .03f * (b - OB_AN_NB_TBANDS)) * frame_tonality);
    slope += bt * (b - 8);
  }
  leakage_from[0] = band_log2[0];
  leakage_to[0] = band_log2[0] - 2.5f;
  for (b = 1; b < OB_AN_NB_TBANDS + 1; b++) {
    const float leak_slope = 2.f * (tbands[b] - tbands[b - 1]) / 4;
    leakage_from[b] = fminf(leakage_from[b - 1] + leak_slope, band_log2[b]);
    leakage_to[b] = fmaxf(leakage_to[b - 1] - leak_slope, band_log2[b] - 2.5f);
  }
  for (b = OB_AN_NB_TBANDS - 2; b >= 0; b--) {
    const float leak_slope = 2.f * (tbands[b + 1] - tbands[b]) / 4;
    leakage_from[b] = fminf(leakage_from[b + 1] + leak_slope, leakage_from[b]);
    leakage_to[b] = fmaxf(leakage_to[b + 1] - leak_slope, leakage_to[b]);
  }
  for (b = 0; b < OB_AN_NB_TBANDS + 1; b++) {
    const float boost = fmaxf(0, leakage_to[b] - band_log2[b]) + fmaxf(0, band_log2[b] - (leakage_from[b] + 2.5f));
    info->leak_boost[b] = (unsigned char)OB_IMIN(255, (int)floor(.5 + (double)(64.f * boost)));
  }
  info->leak_boost[19] = 0;
  float spec_variability = 0;
  for (int i = 0; i < OB_AN_NB_FRAMES; i++) {
    float mindist = 1e15f;
    for (int j = 0; j < OB_AN_NB_FRAMES; j++)
      if (j != i) mindist = fminf(mindist, w->mind[i][j]);
    spec_variability += mindist;
  }
  spec_variability = (float)sqrt((double)(spec_variability / OB_AN_NB_FRAMES / OB_AN_NB_TBANDS));
  float bandwidth_mask = 0, maxE = 0, below_max_pitch = 0, above_max_pitch = 0;
  int bandwidth = 0;
  float noise_floor = 5.7e-4f / (1 << (OB_IMAX(0, lsb_depth - 8)));
  noise_floor *= noise_floor;
  for (b = 0; b < OB_AN_NB_TBANDS; b++) {
    const int band_start = tbands[b], band_end = tbands[b + 1];
    const float E = w->bE2[b];
    maxE = fmaxf(maxE, E);
    if (band_start < 64) below_max_pitch += E;
    else above_max_pitch += E;
    tonal->meanE[b] = fmaxf((1 - alphaE2) * tonal->meanE[b], E);
    const float Em = fmaxf(E, tonal->meanE[b]);
    if (E * 1e9f > maxE && (Em > 3 * noise_floor * (band_end - band_start) || E > noise_floor * (band_end - band_start))) bandwidth = b + 1;
    is_masked[b] = E < (tonal->prev_bandwidth >= b + 1 ? .01f : .05f) * bandwidth_mask;
    bandwidth_mask = fmaxf(.05f * bandwidth_mask, E);
  }
  if (Fs == 48000) {
    const float E = hp_ener * (1.f / (60 * 60));
    const float noise_ratio = tonal->prev_bandwidth == 20 ? 10.f : 30.f;
    above_max_pitch += E;
    tonal->meanE[b] = fmaxf((1 - alphaE2) * tonal->meanE[b], E);
    const float Em = fmaxf(E, tonal->meanE[b]);
    if (Em > 3 * noise_ratio * noise_floor * 160 || E > noise_ratio * noise_floor * 160) bandwidth = 20;
    is_masked[b] = E < (tonal->prev_bandwidth == 20 ? .01f : .05f) * bandwidth_mask;
  } else is_masked[b] = 0;
  if (above_max_pitch > below_max_pitch) info->max_pitch_ratio = below_max_pitch / above_max_pitch;
  else info->max_pitch_ratio = 1;
  if (bandwidth == 20 && is_masked[OB_AN_NB_TBANDS]) bandwidth -= 2;
  else if (bandwidth > 0 && bandwidth <= OB_AN_NB_TBANDS && is_masked[bandwidth - 1]) bandwidth--;
  if (tonal->count <= 2) bandwidth = 20;
  frame_loudness = 20 * (float)log10((double)frame_loudness);
  tonal->Etracker = fmaxf(tonal->Etracker - .003f, frame_loudness);
  tonal->lowECount *= (1 - alphaE);
  if (frame_loudness < tonal->Etracker - 30) tonal->lowECount += alphaE;
  for (int i = 0; i < 8; i++) {
    float sum = 0;
    for (b = 0; b < 16; b++) sum += ob_an_dct[i * 16 + b] * logE[b];
    BFCC[i] = sum;
  }
  for (int i = 0; i < 8; i++) {
    float sum = 0;
    for (b = 0; b < 16; b++) sum += ob_an_dct[i * 16 + b] * .5f * (tonal->highE[b] + tonal->lowE[b]);
    midE[i] = sum;
  }
  frame_stationarity /= OB_AN_NB_TBANDS;
  relativeE /= OB_AN_NB_TBANDS;
  if (tonal->count < 10) relativeE = .5f;
  frame_noisiness /= OB_AN_NB_TBANDS;
  info->activity = frame_noisiness + (1 - frame_noisiness) * relativeE;
  frame_tonality = (max_frame_tonality / (OB_AN_NB_TBANDS - OB_AN_SKIP));
  frame_tonality = fmaxf(frame_tonality, tonal->prev_tonality * .8f);
  tonal->prev_tonality = frame_tonality;
  slope /= 8 * 8;
  info->tonality_slope = slope;
  tonal->E_count = (tonal->E_count + 1) % OB_AN_NB_FRAMES;
  tonal->count = OB_IMIN(tonal->count + 1, 10000);
  info->tonality = frame_tonality;
  for (int i = 0; i < 4; i++)
    features[i] = -0.12299f * (BFCC[i] + tonal->mem[i + 24]) + 0.49195f * (tonal->mem[i] + tonal->mem[i + 16]) + 0.69693f * tonal->mem[i + 8] -
                  1.4349f * tonal->cmean[i];
  for (int i = 0; i < 4; i++) tonal->cmean[i] = (1 - alpha) * tonal->cmean[i] + alpha * BFCC[i];
  for (int i = 0; i < 4; i++) features[4 + i] = 0.63246f * (BFCC[i] - tonal->mem[i + 24]) + 0.31623f * (tonal->mem[i] - tonal->mem[i + 16]);
  for (int i = 0; i < 3; i++)
    features[8 + i] = 0.53452f * (BFCC[i] + tonal->mem[i + 24]) - 0.26726f * (tonal->mem[i] + tonal->mem[i + 16]) - 0.53452f * tonal->mem[i + 8];
  if (tonal->count > 5)
    for (int i = 0; i < 9; i++) tonal->std[i] = (1 - alpha) * tonal->std[i] + alpha * features[i] * features[i];
  for (int i = 0; i < 4; i++) features[i] = BFCC[i] - midE[i];
  for (int i = 0; i < 8; i++) {
    tonal->mem[i + 24] = tonal->mem[i + 16];
    tonal->mem[i + 16] = tonal->mem[i + 8];
    tonal->mem[i + 8] = tonal->mem[i];
    tonal->mem[i] = BFCC[i];
  }
  for (int i = 0; i < 9; i++) features[11 + i] = (float)sqrt((double)tonal->std[i]) - ob_an_std_bias[i];
  features[18] = spec_variability - 0.78f;
  features[20] = info->tonality - 0.154723f;
  features[21] = info->activity - 0.724643f;
  features[22] = frame_stationarity - 0.743717f;
  features[23] = info->tonality_slope + 0.069216f;
  features[24] = tonal->lowECount - 0.067930f;
  for (int i = 0; i < 25; i++) w->feat[i] = features[i];
  info->bandwidth = bandwidth;
  tonal->prev_bandwidth = bandwidth;
  info->noisiness = frame_noisiness;
}

// the Opus layer's per-stream decisions for a CELT-only frame (src/opus_encoder.c:1273-1305, 1407-1456, 1583-1686, 2266-2349, 2551), one lane
OB_DEV int32_t opus_equiv_rate(int32_t bitrate, int channels, int frame_rate, int vbr, int celt_mode, int complexity, int loss) {   // :1027-1058
  int32_t equiv = bitrate;
  if (frame_rate > 50) equiv -= (40 * channels + 20) * (frame_rate - 50);
  if (!vbr) equiv -= equiv / 12;
  equiv = equiv * (90 + complexity) / 100;
  if (celt_mode) { if (complexity < 5) equiv = equiv * 9 / 10; }
  else equiv -= equiv * loss / (12 * loss + 20);
  return equiv;
}
OB_DEV int opus_toc_celt(int framerate, int bandwidth, int channels) {   // gen_toc :330-360, MODE_CELT_ONLY
  int period = 0;
  while (framerate < 400) { framerate <<= 1; period++; }
  int tmp = bandwidth - 1102;
  if (tmp < 0) tmp = 0;
  return 0x80 | (tmp << 5) | (period << 3) | ((channels == 2) << 2);
}
// frame_size: the packet's (the rate decisions see the packet's frame rate); toc_frame_size: the coded frame's (20 ms inside a multi-frame packet)
OB_DEVNI void opus_layer_decide(const OpusLayerParams &p, ObOpusEncState *st, int silence, int clear, const ObAnalysisInfo &ai, ObOpusFrameCtl *ctl,
                                int frame_size, int toc_frame_size) {
  const int BW_NB = 1101, BW_MB = 1102, BW_WB = 1103, BW_SWB = 1104, BW_FB = 1105, AUTO = -1000;
  const int frame_rate = p.Fs / frame_size;
  if (!silence) st->voice_ratio = -1;
  int detected_bandwidth = 0;
  if (ai.valid) {
    if (p.signal_type == AUTO) {
      const float prob = st->prev_mode == 0 ? ai.music_prob : ai.music_prob_max;
      st->voice_ratio = (int)floor(.5 + (double)(100 * (1 - prob)));
    }
    const int ab = ai.bandwidth;
    detected_bandwidth = ab <= 12 ? BW_NB : ab <= 14 ? BW_MB : ab <= 16 ? BW_WB : ab <= 18 ? BW_SWB : BW_FB;
  }
  int32_t equiv_rate = opus_equiv_rate(p.bitrate_bps, p.channels, frame_rate, p.use_vbr, 0, p.complexity, p.packet_loss);
  int voice_est;
  if (p.signal_type == 3001) voice_est = 127;
  else if (p.signal_type == 3002) voice_est = 0;
  else if (st->voice_ratio >= 0) {
    voice_est = st->voice_ratio * 327 >> 8;
    if (p.application == 2049) voice_est = OB_IMIN(voice_est, 115);
  } else if (p.application == 2048) voice_est = 115;
  else voice_est = 48;
  if (p.force_channels != AUTO && p.channels == 2) st->stream_channels = p.force_channels;
  else if (p.channels == 2) {
    int32_t stereo_threshold = 17000 + ((voice_est * voice_est * (19000 - 17000)) >> 14);
    if (st->stream_channels == 2) stereo_threshold -= 1000;
    else stereo_threshold += 1000;
    st->stream_channels = (equiv_rate > stereo_threshold) ? 2 : 1;
  } else st->stream_channels = p.channels;
  equiv_rate = opus_equiv_rate(p.bitrate_bps, st->stream_channels, frame_rate, p.use_vbr, 1, p.complexity, p.packet_loss);
  {
    const int32_t voice_thr[8] = {9000, 700, 9000, 700, 13500, 1000, 14000, 2000}, music_thr[8] = {9000, 700, 9000, 700, 11000, 1000, 12000, 2000};
    int bandwidth = BW_FB;
    do {
      const int k = 2 * (bandwidth - BW_MB);
      int threshold = music_thr[k] + ((voice_est * voice_est * (voice_thr[k] - music_thr[k])) >> 14);
      const int hysteresis = music_thr[k + 1] + ((voice_est * voice_est * (voice_thr[k + 1] - music_thr[k + 1])) >> 14);
      if (!st->first) {
        if (st->auto_bandwidth >= bandwidth) threshold -= hysteresis;
        else threshold += hysteresis;
      }
      if (equiv_rate >= threshold) break;
    } while (--bandwidth > BW_NB);
    if (bandwidth == BW_MB) bandwidth = BW_WB;
    st->bandwidth = st->auto_bandwidth = bandwidth;
  }
  if (st->bandwidth > p.max_bandwidth) st->bandwidth = p.max_bandwidth;
  if (p.user_bandwidth != AUTO) st->bandwidth = p.user_bandwidth;
  if (p.Fs <= 24000 && st->bandwidth > BW_SWB) st->bandwidth = BW_SWB;
  if (p.Fs <= 16000 && st->bandwidth > BW_WB) st->bandwidth = BW_WB;
  if (p.Fs <= 12000 && st->bandwidth > BW_MB) st->bandwidth = BW_MB;
  if (p.Fs <= 8000 && st->bandwidth > BW_NB) st->bandwidth = BW_NB;
  if (detected_bandwidth && p.user_bandwidth == AUTO) {
    int min_detected;
    if (equiv_rate <= 18000 * st->stream_channels) min_detected = BW_NB;
    else if (equiv_rate <= 24000 * st->stream_channels) min_detected = BW_MB;
    else if (equiv_rate <= 30000 * st->stream_channels) min_detected = BW_WB;
    else if (equiv_rate <= 44000 * st->stream_channels) min_detected = BW_SWB;
    else min_detected = BW_FB;
    detected_bandwidth = OB_IMAX(detected_bandwidth, min_detected);
    st->bandwidth = OB_IMIN(st->bandwidth, detected_bandwidth);
  }
  if (st->bandwidth == BW_MB) st->bandwidth = BW_WB;
  const int bw = st->bandwidth;
  ctl->end = bw == BW_NB ? 13 : (bw == BW_MB || bw == BW_WB) ? 17 : bw == BW_SWB ? 19 : 21;
  ctl->C = st->stream_channels;
  ctl->toc = opus_toc_celt(p.Fs / toc_frame_size, bw, st->stream_channels);
  ctl->clear = clear;
  ctl->fade = 0;
  ctl->fade_g1 = ctl->fade_g2 = 0.f;
  {
    int w14;
    if (equiv_rate > 32000) w14 = 16384;
    else if (equiv_rate < 16000) w14 = 0;
    else w14 = 16384 - 2048 * (int32_t)(32000 - equiv_rate) / (equiv_rate - 14000);
    if (p.channels == 2 && (st->hybrid_stereo_width_Q14 < (1 << 14) || w14 < (1 << 14))) {
      float g1 = (float)st->hybrid_stereo_width_Q14, g2 = (float)w14;
      g1 *= (1.f / 16384);
      g2 *= (1.f / 16384);
      ctl->fade = 1;
      ctl->fade_g1 = 1.f - g1;
      ctl->fade_g2 = 1.f - g2;
      st->hybrid_stereo_width_Q14 = w14;
    }
  }
  ctl->analysis = ai;
  st->prev_mode = 1002;
  st->first = 0;
}

// P2: one warp per stream
OB_DEV void opus_enc_analysis_warp(const OpusLayerParams &p, int s, AnWS *w) {
  const int lane = lane_id();
  ObOpusEncState *st = p.st + s;
  const ObOpusPre *pre = p.pre + s;
  ObOpusFrameCtl *ctl = p.ctl + s;
  ObAnalysisInfo ai;
  ai.valid = 0;
  if (!p.run_analysis) {
    // tonality_analysis_reset when the analysis is off (:1261-1263) happens on the host side (memset at the complexity change)
    if (lane == 0) {
      ai.tonality = ai.tonality_slope = ai.noisiness = ai.activity = ai.music_prob = ai.music_prob_min = ai.music_prob_max = 0.f;
      ai.bandwidth = 0; ai.activity_probability = ai.max_pitch_ratio = 0.f;
      for (int i = 0; i < 20; i++) ai.leak_boost[i] = 0;
      if (!p.multi) {
        opus_layer_decide(p, st, pre->silence, pre->clear, ai, ctl, p.frame_size, p.frame_size);
        p.packets[(size_t)s * p.pkt_stride] = (unsigned char)ctl->toc;
      }
    }
    return;
  }
  ObTonality *tonal = p.tonal + s;
  const int Fs = p.Fs, N = 480, N2 = 240;
  const int len = Fs == 48000 ? p.frame_size / 2 : Fs == 16000 ? 3 * p.frame_size / 2 : p.frame_size;
  const int count = tonal->count;
  int mem_fill = tonal->initialized ? tonal->mem_fill : 240;
  const float alpha = 1.f / OB_IMIN(10, 1 + count), alphaE = 1.f / OB_IMIN(25, 1 + count);
  float alphaE2 = 1.f / OB_IMIN(100, 1 + count);
  if (count <= 1) alphaE2 = 1;
  const int n1 = OB_IMIN(len, OB_AN_BUF - mem_fill);
  wsync();
  for (int i = lane; i < n1; i += 32) tonal->inmem[mem_fill + i] = pre->newsamp[i];
  wsync();
  if (mem_fill + len < OB_AN_BUF) {
    if (lane == 0) { tonal->hp_ener_accum += pre->hp1; tonal->mem_fill = mem_fill + len; tonal->initialized = 1; }
  } else {
    const float hp_ener = tonal->hp_ener_accum + pre->hp1;
    const int wp = tonal->write_pos;
    ObAnalysisInfo *info = &tonal->info[wp];
    // is_digital_silence over the 30 ms buffer, and the windowed FFT input (src/analysis.c:529-542)
    float mx = 0.f, mn = 0.f;
    for (int i = lane; i < OB_AN_BUF; i += 32) { const float v = tonal->inmem[i]; mx = fmaxf(mx, v); mn = fminf(mn, v); }
    mx = warp_max(mx);
    mn = -warp_max(-mn);
    const int is_silence = fmaxf(mx, -mn) <= 1.f / (float)(1 << (p.pcm_is_int16 ? OB_IMIN(16, p.lsb_depth) : OB_IMIN(24, p.lsb_depth)));
    const float scale = 0.0020833334f;   // kiss_fft_state.scale of the 480-point transform (celt/static_modes_float.h)
    for (int i = lane; i < N2; i += 32) {
      const float wv = ob_an_window[i];
      cpx a, b;
      a.r = wv * tonal->inmem[i];
      a.i = wv * tonal->inmem[N2 + i];
      b.r = wv * tonal->inmem[N - i - 1];
      b.i = wv * tonal->inmem[N + N2 - i - 1];
      // opus_fft_c: scaled, digit-reversed copy (celt/kiss_fft.c:596-607)
      cpx sa, sb;
      sa.r = scale * a.r; sa.i = scale * a.i; sb.r = scale * b.r; sb.i = scale * b.i;
      w->out[ob_fft_bitrev480[i]] = sa;
      w->out[ob_fft_bitrev480[N - i - 1]] = sb;
    }
    wsync();
    // OPUS_MOVE(inmem, inmem + 480, 240) and the rest of this frame's samples (:543-548)
    float keep[8];
    for (int k = 0; k < 8; k++) { const int i = lane + 32 * k; keep[k] = i < 240 ? tonal->inmem[OB_AN_BUF - 240 + i] : 0.f; }
    wsync();
    for (int k = 0; k < 8; k++) { const int i = lane + 32 * k; if (i < 240) tonal->inmem[i] = keep[k]; }
    const int remaining = len - n1;
    for (int i = lane; i < remaining; i += 32) tonal->inmem[240 + i] = pre->newsamp[n1 + i];
    wsync();
    if (lane == 0) {
      tonal->write_pos = wp + 1 >= OB_AN_DETECT ? wp + 1 - OB_AN_DETECT : wp + 1;
      tonal->hp_ener_accum = pre->hp2;
      tonal->mem_fill = 240 + remaining;
      tonal->initialized = 1;
    }
    if (is_silence) {
      if (lane == 0) {
        int prev_pos = wp + 1 - 2;
        if (prev_pos < 0) prev_pos += OB_AN_DETECT;
        *info = tonal->info[prev_pos];
      }
    } else {
      fft_stages(w->out, 0, 1, lane);
      const cpx *out = w->out;
      const bool isnan0 = out[0].r != out[0].r;
      if (isnan0) { if (lane == 0) info->valid = 0; }
      else {
        const float pi4 = (float)(3.14159265358979323846 * 3.14159265358979323846 * 3.14159265358979323846 * 3.14159265358979323846);
        const float inv2pi = (float)(.5f / 3.14159265358979323846);
        float *A = tonal->angle, *dA = tonal->d_angle, *d2A = tonal->d2_angle;
        for (int i = 1 + lane; i < N2; i += 32) {
          const float X1r = (float)out[i].r + out[N - i].r, X1i = (float)out[i].i - out[N - i].i;
          const float X2r = (float)out[i].i + out[N - i].i, X2i = (float)out[N - i].r - out[i].r;
          const float angle = inv2pi * an_fast_atan2f(X1i, X1r);
          const float d_angle = angle - A[i];
          const float d2_angle = d_angle - dA[i];
          const float angle2 = inv2pi * an_fast_atan2f(X2i, X2r);
          const float d_angle2 = angle2 - angle;
          const float d2_angle2 = d_angle2 - d_angle;
          float mod1 = d2_angle - an_lrintf(d2_angle);
          float nz = fabsf(mod1);
          mod1 *= mod1;
          mod1 *= mod1;
          float mod2 = d2_angle2 - an_lrintf(d2_angle2);
          nz += fabsf(mod2);
          mod2 *= mod2;
          mod2 *= mod2;
          const float avg_mod = .25f * (d2A[i] + mod1 + 2 * mod2);
          w->noisiness[i] = nz;
          w->tonality[i] = 1.f / (1.f + 40.f * 16.f * pi4 * avg_mod) - .015f;
          w->tonality2[i] = 1.f / (1.f + 40.f * 16.f * pi4 * mod2) - .015f;
          A[i] = angle2;
          dA[i] = d_angle2;
          d2A[i] = mod2;
        }
        wsync();
        float sm[8];
        for (int k = 0; k < 8; k++) {
          const int i = 2 + lane + 32 * k;
          sm[k] = 0.f;
          if (i < N2 - 1) {
            const float tt = fminf(w->tonality2[i], fmaxf(w->tonality2[i - 1], w->tonality2[i + 1]));
            sm[k] = .9f * fmaxf(w->tonality[i], tt - .1f);
          }
        }
        wsync();
        for (int k = 0; k < 8; k++) { const int i = 2 + lane + 32 * k; if (i < N2 - 1) w->tonality[i] = sm[k]; }
        wsync();
        // band sums: one lane per band, bins in order (:646-659, :783-797); lane 18 takes the DC band (:631-645)
        const short *tbands = ob_an_tbands;
        const int E_count = tonal->E_count;
        int bad = 0;
        if (lane < OB_AN_NB_TBANDS) {
          const int b = lane;
          float E = 0, tE = 0, nE = 0, E2 = 0;
          for (int i = tbands[b]; i < tbands[b + 1]; i++) {
            float binE = out[i].r * (float)out[i].r + out[N - i].r * (float)out[N - i].r + out[i].i * (float)out[i].i + out[N - i].i * (float)out[N - i].i;
            E2 += binE;
            binE = (1.f / 32768 / 32768) * binE;
            E += binE;
            tE += binE * fmaxf(0, w->tonality[i]);
            nE += binE * 2.f * (.5f - w->noisiness[i]);
          }
          E2 = (1.f / 32768 / 32768) * E2;
          bad = !(E < 1e9f) || E != E;
          w->bE[b] = E; w->btE[b] = tE; w->bnE[b] = nE; w->bE2[b] = E2;
        } else if (lane == OB_AN_NB_TBANDS) {
          const float X1r = 2 * (float)out[0].r, X2r = 2 * (float)out[0].i;
          float E = X1r * X1r + X2r * X2r;
          for (int i = 1; i < 4; i++) {
            const float binE = out[i].r * (float)out[i].r + out[N - i].r * (float)out[N - i].r + out[i].i * (float)out[i].i + out[N - i].i * (float)out[N - i].i;
            E += binE;
          }
          w->E0 = (1.f / 32768 / 32768) * E;
        }
        // the reference returns at the first band whose energy is out of range, after having updated the bands below it
        const unsigned badmask = ballot(bad);
        const int nb_ok = badmask ? ffs32(badmask) - 1 : OB_AN_NB_TBANDS;
        if (lane == 0) info->activity = 0;
        if (!count && lane < OB_AN_NB_TBANDS) { tonal->lowE[lane] = 1e10; tonal->highE[lane] = -1e10; }
        if (lane < nb_ok) {
          const int b = lane;
          const float E = w->bE[b];
          tonal->E[E_count][b] = E;
          const float lE = (float)log((double)(E + 1e-10f));
          w->blogE[b] = lE;
          tonal->logE[E_count][b] = lE;
          float hi = tonal->highE[b], lo = tonal->lowE[b];
          if (count == 0) hi = lo = lE;
          if (hi > lo + 7.5) {
            if (hi - lE > lE - lo) hi -= .01f;
            else lo += .01f;
          }
          if (lE > hi) { hi = lE; lo = fmaxf(hi - 15, lo); }
          else if (lE < lo) { lo = lE; hi = fminf(lo + 15, hi); }
          tonal->highE[b] = hi; tonal->lowE[b] = lo;
          w->brel[b] = (lE - lo) / (1e-5f + (hi - lo));
          float L1 = 0, L2 = 0;
          for (int i = 0; i < OB_AN_NB_FRAMES; i++) {
            const float e = i == E_count ? E : tonal->E[i][b];
            L1 += (float)sqrt((double)e);
            L2 += e;
          }
          float stationarity = fminf(0.99f, L1 / (float)sqrt(1e-15 + (double)(OB_AN_NB_FRAMES * L2)));
          stationarity *= stationarity;
          stationarity *= stationarity;
          w->bstat[b] = stationarity;
          const float bt = fmaxf(w->btE[b] / (1e-15f + E), stationarity * tonal->prev_band_tonality[b]);
          w->btone[b] = bt;
          tonal->prev_band_tonality[b] = bt;
        }
        wsync();
        if (badmask) { if (lane == 0) info->valid = 0; }
        else {
          // spectral variability: the 64 frame pairs over the lanes (:756-775)
          for (int q = lane; q < OB_AN_NB_FRAMES * OB_AN_NB_FRAMES; q += 32) {
            const int i = q >> 3, j = q & 7;
            float dist = 0;
            for (int k = 0; k < OB_AN_NB_TBANDS; k++) {
              const float tmp = tonal->logE[i][k] - tonal->logE[j][k];
              dist += tmp * tmp;
            }
            w->mind[i][j] = dist;
          }
          wsync();
          if (lane == 0) an_frame_tail(tonal, info, w, hp_ener, p.pcm_is_int16 ? OB_IMIN(16, p.lsb_depth) : OB_IMIN(24, p.lsb_depth), Fs, alpha, alphaE, alphaE2);
          wsync();
          // the network (src/mlp.c, src/analysis.c:931-937): a lane per neuron
          {
            float o = (float)ob_an_layer0_bias[lane];
            o = an_gemm1(o, ob_an_layer0_weights, lane, 25, 32, w->feat);
            o *= (1.f / 128);
            w->l0[lane] = an_tansig(o);
          }
          wsync();
          const int NN = 24;
          float z = 0, r = 0, h = 0, sprev = 0;
          if (lane < NN) {
            sprev = tonal->rnn_state[lane];
            w->gt[lane] = sprev;
          }
          wsync();
          if (lane < NN) {
            z = (float)ob_an_layer1_bias[lane];
            z = an_gemm1(z, ob_an_layer1_weights, lane, 32, 3 * NN, w->l0);
            z = an_gemm1(z, ob_an_layer1_recur_weights, lane, NN, 3 * NN, w->gt);
            z = an_sigmoid((1.f / 128) * z);
            r = (float)ob_an_layer1_bias[NN + lane];
            r = an_gemm1(r, ob_an_layer1_weights + NN, lane, 32, 3 * NN, w->l0);
            r = an_gemm1(r, ob_an_layer1_recur_weights + NN, lane, NN, 3 * NN, w->gt);
            r = an_sigmoid((1.f / 128) * r);
            w->gr[lane] = sprev * r;
          }
          wsync();
          if (lane < NN) {
            h = (float)ob_an_layer1_bias[2 * NN + lane];
            h = an_gemm1(h, ob_an_layer1_weights + 2 * NN, lane, 32, 3 * NN, w->l0);
            h = an_gemm1(h, ob_an_layer1_recur_weights + 2 * NN, lane, NN, 3 * NN, w->gr);
            h = z * sprev + (1 - z) * an_tansig((1.f / 128) * h);
            w->gz[lane] = h;
            tonal->rnn_state[lane] = h;
          }
          wsync();
          if (lane < 2) {
            float o = (float)ob_an_layer2_bias[lane];
            o = an_gemm1(o, ob_an_layer2_weights, lane, NN, 2, w->gz);
            o *= (1.f / 128);
            o = an_sigmoid(o);
            if (lane == 1) info->activity_probability = o;
            else info->music_prob = o;
          }
          wsync();
          if (lane == 0) info->valid = 1;
        }
      }
    }
  }
  wsync();
  if (lane == 0 && !p.multi) {
    an_get_info(tonal, &ai, p.frame_size, Fs);
    opus_layer_decide(p, st, pre->silence, pre->clear, ai, ctl, p.frame_size, p.frame_size);
    p.packets[(size_t)s * p.pkt_stride] = (unsigned char)ctl->toc;
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// 40..120 ms packets: one thread per stream after the P1 / P2 passes of all the packet's 20 ms frames.  The packet's decisions from
// tonality_get_info over the whole packet (run_analysis, src/analysis.c:968-971), then — read position reset as in :1727-1733 — the analysis info of
// every 20 ms frame for its own CELT call (:1799-1802), and the frame controls: TOC of a 20 ms frame, float guard of that frame, stereo fade from the
// previous width in the first frame and at the new width in the others (hybrid_stereo_width_Q14 is updated by the first).
// ---------------------------------------------------------------------------------------------------------------------------
OB_DEV void opus_multiframe_decide_thread(const OpusLayerParams &p, int s) {
  ObOpusEncState *st = p.st + s;
  ObAnalysisInfo ai;
  ai.valid = 0;
  ai.tonality = ai.tonality_slope = ai.noisiness = ai.activity = ai.music_prob = ai.music_prob_min = ai.music_prob_max = 0.f;
  ai.bandwidth = 0; ai.activity_probability = ai.max_pitch_ratio = 0.f;
  for (int i = 0; i < 20; i++) ai.leak_boost[i] = 0;
  ObTonality *tonal = p.run_analysis ? p.tonal + s : nullptr;
  int pos_bak = 0, sub_bak = 0, silence = 1;
  if (tonal) {
    pos_bak = tonal->read_pos; sub_bak = tonal->read_subframe;
    an_get_info(tonal, &ai, p.packet_frame_size, p.Fs);
  }
  for (int i = 0; i < p.nb_frames; i++) silence &= p.pre_all[(size_t)i * p.n_streams + s].silence;
  ObOpusFrameCtl c0;
  opus_layer_decide(p, st, silence, 0, ai, &c0, p.packet_frame_size, p.frame_size);
  if (tonal) { tonal->read_pos = pos_bak; tonal->read_subframe = sub_bak; }
  for (int i = 0; i < p.nb_frames; i++) {
    ObOpusFrameCtl *c = p.ctl_all + (size_t)i * p.n_streams + s;
    *c = c0;
    if (tonal) an_get_info(tonal, &c->analysis, p.frame_size, p.Fs);
    c->clear = p.pre_all[(size_t)i * p.n_streams + s].clear;
    if (i > 0) {
      c->fade = c0.fade && c0.fade_g2 > 0.f;
      c->fade_g1 = c->fade_g2 = c->fade ? c0.fade_g2 : 0.f;
    }
    p.tmp_packets[((size_t)i * p.n_streams + s) * p.tmp_stride] = (unsigned char)c0.toc;
  }
}
// opus_repacketizer_cat of the packet's one-frame packets + opus_repacketizer_out_range_impl(rp, 0, nb_frames, data, repacketize_len, 0, pad, NULL, 0)
// (src/repacketizer.c:88-152, 160-365; no extensions): code 1 / 2 for two frames, code 3 otherwise or when the CBR padding is needed
OB_DEV int opus_repack_size(int size, unsigned char *data) {   // encode_size
  if (size < 252) { data[0] = (unsigned char)size; return 1; }
  data[0] = (unsigned char)(252 + (size & 0x3));
  data[1] = (unsigned char)((size - (int)data[0]) >> 2);
  return 2;
}
OB_DEV void opus_repacketize_thread(const OpusLayerParams &p, int s) {
  const int nb = p.nb_frames, maxlen = p.repacketize_len, pad = p.pad;
  unsigned char *data = p.packets + (size_t)s * p.pkt_stride, *ptr = data;
  int len[6], tot_size = 0, bad = 0;
  for (int i = 0; i < nb; i++) { len[i] = p.tmp_nbytes[(size_t)i * p.n_streams + s] - 1; bad |= len[i] < 0; }   // the frame kernel counts the TOC
  if (bad) { p.out_nbytes[s] = -3; return; }
  const unsigned char toc = p.tmp_packets[(size_t)s * p.tmp_stride] & 0xFC;
  if (nb == 2) {
    if (len[1] == len[0]) {
      tot_size = 2 * len[0] + 1;
      if (tot_size > maxlen) { p.out_nbytes[s] = -3; return; }
      *ptr++ = toc | 0x1;
    } else {
      tot_size = len[0] + len[1] + 2 + (len[0] >= 252);
      if (tot_size > maxlen) { p.out_nbytes[s] = -3; return; }
      *ptr++ = toc | 0x2;
      ptr += opus_repack_size(len[0], ptr);
    }
  }
  if (nb > 2 || (pad && tot_size < maxlen)) {
    int vbr = 0;
    ptr = data;
    tot_size = 0;
    for (int i = 1; i < nb; i++) if (len[i] != len[0]) { vbr = 1; break; }
    if (vbr) {
      tot_size += 2;
      for (int i = 0; i < nb - 1; i++) tot_size += 1 + (len[i] >= 252) + len[i];
      tot_size += len[nb - 1];
    } else tot_size += nb * len[0] + 2;
    if (tot_size > maxlen) { p.out_nbytes[s] = -3; return; }
    *ptr++ = toc | 0x3;
    *ptr++ = (unsigned char)(vbr ? (nb | 0x80) : nb);
    const int pad_amount = pad ? (maxlen - tot_size) : 0;
    if (pad_amount != 0) {
      const int nb_255s = (pad_amount - 1) / 255;
      data[1] |= 0x40;
      if (tot_size + nb_255s + 1 > maxlen) { p.out_nbytes[s] = -3; return; }
      for (int i = 0; i < nb_255s; i++) *ptr++ = 255;
      *ptr++ = (unsigned char)(pad_amount - 255 * nb_255s - 1);
      tot_size += pad_amount;
    }
    if (vbr) for (int i = 0; i < nb - 1; i++) ptr += opus_repack_size(len[i], ptr);
  }
  for (int i = 0; i < nb; i++) {
    const unsigned char *f = p.tmp_packets + ((size_t)i * p.n_streams + s) * p.tmp_stride + 1;
    for (int k = 0; k < len[i]; k++) ptr[k] = f[k];
    ptr += len[i];
  }
  if (pad) while (ptr < data + maxlen) *ptr++ = 0;
  p.out_nbytes[s] = tot_size;
}

}  // namespace ob
