// celt_dec_entropy.cuh — kernel A1: ONE THREAD PER STREAM runs everything that touches the
// bitstream: Opus packet framing, CELT frame header, Laplace/fine energies, TF, spread, dynalloc,
// trim, bit allocation, and the band loop's *control* — theta decoding, the partition recursion with
// its bit re-balancing, and the PVQ codeword indices (ec_dec_uint) — and writes a compact script
// (ObScript) for the vector stage.  Nothing here depends on a float signal value: in the decoder no
// decoded sample ever feeds back into parsing (SURVEY §8(a) notes), so this stage is pure integer
// plus the band-energy floats, and the 32 streams of a warp never communicate.
//
// Why thread-per-stream: the first design ran this code once per WARP (uniform, 32x redundant);
// ncu showed the SM starved on instruction fetch (gcc instruction requests 97 % of peak, 53-72 %
// of stall samples "no instruction") at 0.6 IPC.  Per-thread execution amortises every fetched
// instruction over up to 32 streams; divergence is bounded by writing the partition recursion as
// an explicit-stack loop (threads reconverge at the loop head once per tree node).
//
// Reference: celt/celt_decoder.c:1104-1617, celt/quant_bands.c:431-547, celt/rate.c:249-646,
// celt/bands.c:700-1176,1387-1923 (control flow only), src/opus.c:224-420.
#pragma once
#include "celt_types.h"
#include "rangedec.cuh"
#include "celt_tables_data.h"

namespace ob {

#define OB_BITRES 3
#define OB_IMIN(a, b) ((a) < (b) ? (a) : (b))
#define OB_IMAX(a, b) ((a) > (b) ? (a) : (b))

struct DecEntropyParams {
  const unsigned char *packets;  // [n_streams][pkt_stride], Opus packets (with TOC)
  const int32_t *pkt_len;        // [n_streams], <=0 : nothing to decode for this stream
  int pkt_stride;
  int n_streams;
  int channels;     // output channels (decoder was created with), 1 or 2
  int frame_index;  // which frame of a multi-frame packet to decode in this launch
  ObDecBandState *state;  // [n_streams]
  ObFrameSide *side;      // [n_streams]
  ObScript *script;       // [n_streams]
  int32_t *out_samples;   // [n_streams] or null: running per-stream result (samples/channel or error)
  int max_samples;        // per-channel capacity of each stream's pcm area, in OUTPUT-rate samples
  int downsample;         // 48000 / output rate (1, 2, 3, 4 or 6): celt_decoder.c:219-231
  int disable_inv;        // OPUS_SET_PHASE_INVERSION_DISABLED (celt_decoder.c:1763-1771); mono output disables it by itself
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
#pragma unroll 1
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
OB_DEV int bits_to_pseudo(const unsigned char *cache, int bits) {
  int lo = 0, hi = cache[0];
  bits--;
#pragma unroll 1
  for (int i = 0; i < 6; i++) {
    int mid = (lo + hi + 1) >> 1;
    if ((int)cache[mid] >= bits) hi = mid; else lo = mid;
  }
  if (bits - (lo == 0 ? -1 : (int)cache[lo]) <= (int)cache[hi] - bits) return lo;
  return hi;
}
OB_DEV uint32_t pvq_u(int n, int k) {
  int a = OB_IMIN(n, k), b = OB_IMAX(n, k);
  return ob_pvq_u[ob_pvq_rowoff[a] + b];
}
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

struct ThetaInfo { int inv, imid, iside, delta, itheta, qalloc, b; };
struct ThetaRet { RangeDec ec; ThetaInfo t; };

// decode side of celt/bands.c:700-933 (without the `fill` bookkeeping, which belongs to the vector stage)
OB_DEVNI ThetaRet ent_theta(RangeDec ec, int band, int intensity, int disable_inv, int32_t remaining_bits, int N, int b,
                           int B0, int LM, int stereo) {
  int itheta = 0, inv = 0;
  int pulse_cap = ob_logN[band] + LM * (1 << OB_BITRES);
  int offset = (pulse_cap >> 1) - (stereo && N == 2 ? 16 : 4);
  int qn = theta_levels(N, b, offset, pulse_cap, stereo);
  if (stereo && band >= intensity) qn = 1;
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
    if (disable_inv) inv = 0;
    itheta = 0;
  }
  int qalloc = ec.tell_frac() - tell;
  b -= qalloc;
  int imid, iside, delta;
  if (itheta == 0) { imid = 32767; iside = 0; delta = -16384; }
  else if (itheta == 16384) { imid = 0; iside = 32767; delta = 16384; }
  else {
    imid = bitexact_cos((int16_t)itheta);
    iside = bitexact_cos((int16_t)(16384 - itheta));
    delta = frac_mul16((N - 1) << 7, bitexact_log2tan(iside, imid));
  }
  ThetaRet r;
  r.ec = ec;
  r.t.inv = inv; r.t.imid = imid; r.t.iside = iside; r.t.delta = delta; r.t.itheta = itheta; r.t.qalloc = qalloc; r.t.b = b;
  return r;
}
OB_DEV int theta_class(int itheta) { return itheta == 0 ? 0 : itheta == 16384 ? 1 : 2; }

// Per-thread emission state
struct Emit {
  RangeDec ec;
  int32_t remaining_bits;
  ObLeaf *leaves;
  int n_leaves;
  int band, intensity, disable_inv;
};

OB_DEV void emit_leaf(Emit &e, uint32_t index, float gain, int off, int n, int K, int B, int kind, int lb, unsigned path,
                      int depth, int cmshift) {
  if (e.n_leaves < OB_MAX_LEAVES) {
    ObLeaf l;
    l.index = index; l.gain = gain; l.off = (uint16_t)off; l.n = (uint8_t)n; l.K = (uint8_t)K;
    l.path = (uint16_t)((path & 0xFFFu) | (unsigned)depth << 12 | (lb ? 0x8000u : 0u));
    l.B_kind = (uint8_t)(B | kind << 5);
    l.cmshift = (uint8_t)cmshift;
    e.leaves[e.n_leaves] = l;
  }
  e.n_leaves++;
}

// One frame of the explicit-stack depth-first walk over a band part's partition tree (entropy
// side of celt/bands.c:973-1176); each leaf records its split path (see ObLeaf).
struct PFrame {
  int b, mbits, sbits, rebalance;
  float gain, mid, side;
  short off, N, itheta;
  unsigned short path;
  signed char LM;
  unsigned char B, B0, stage, lb, mid_first, depth, cmshift;
};

// ---- Opus packet framing (RFC 6716 section 3.2; src/opus.c:183-420) -------------------------------
OB_HD int parse_size(const unsigned char *d, int len, int &size) {
  if (len < 1) { size = -1; return -1; }
  if (d[0] < 252) { size = d[0]; return 1; }
  if (len < 2) { size = -1; return -1; }
  size = 4 * d[1] + d[0];
  return 2;
}
// Returns number of frames in the packet (<0 on a malformed packet); off/flen describe frame k.
// (also compiled for the host: the host-buffer API uses it to know how many frame launches a call needs)
OB_HDNI int locate_frame(const unsigned char *data, int len, int k, int *off_out, int *flen_out) {
  int pos = 0, count, cbr = 0, last_size, sz0 = 0, off = 0, flen = 0;
  *off_out = 0; *flen_out = 0;
  if (len <= 0) return -4;
  int toc = data[0];
  int framesize = (48000 << ((toc >> 3) & 3)) / 400;  // CELT-only TOC (checked by the caller)
  pos = 1; len--;
  last_size = len;
  switch (toc & 3) {
    case 0: count = 1; break;
    case 1:
      count = 2; cbr = 1;
      if (len & 1) return -4;
      last_size = len / 2;
      sz0 = last_size;
      break;
    case 2: {
      count = 2;
      int bytes = parse_size(data + pos, len, sz0);
      len -= bytes;
      if (sz0 < 0 || sz0 > len) return -4;
      pos += bytes;
      last_size = len - sz0;
      break;
    }
    default: {
      if (len < 1) return -4;
      int ch = data[pos++];
      count = ch & 0x3F;
      if (count <= 0 || framesize * count > 5760) return -4;
      len--;
      if (ch & 0x40) {
        int p;
        do {
          if (len <= 0) return -4;
          p = data[pos++];
          len--;
          int t = p == 255 ? 254 : p;
          len -= t;
        } while (p == 255);
      }
      if (len < 0) return -4;
      cbr = !(ch & 0x80);
      if (!cbr) {
        last_size = len;
        int accum = 0, mysize = -1;
        for (int i = 0; i < count - 1; i++) {
          int sz;
          int bytes = parse_size(data + pos, len, sz);
          len -= bytes;
          if (sz < 0 || sz > len) return -4;
          pos += bytes;
          last_size -= bytes + sz;
          if (i < k) accum += sz;
          if (i == k) mysize = sz;
        }
        if (last_size < 0) return -4;
        if (last_size > 1275) return -4;
        if (k < count) { *off_out = pos + accum; *flen_out = k == count - 1 ? last_size : mysize; }
        return count;
      } else {
        last_size = len / count;
        if (last_size * count != len) return -4;
        sz0 = last_size;
      }
      break;
    }
  }
  if (last_size > 1275) return -4;
  if (k < count) {
    if (cbr || count == 1) { off = pos + k * last_size; flen = last_size; }
    else { off = pos + (k == 0 ? 0 : sz0); flen = k == 0 ? sz0 : last_size; }  // code 2
  }
  *off_out = off; *flen_out = flen;
  return count;
}

OB_DEV void ent_finish(const DecEntropyParams &p, int stream, ObFrameSide *side, ObDecBandState *st, int status) {
  side->status = status;  // the later kernels see the 48 kHz frame length
  const int rep = status > 0 ? status / p.downsample : status;  // the caller sees output-rate samples (opus_decoder.c:735)
  if (status < 0 || p.frame_index == 0) {
    // OPUS_GET_LAST_PACKET_DURATION is set by successful calls only (src/opus_decoder.c:880-883): an error or a stream that sits the
    // call out (len < 0) leaves it alone; the caller sees the code in out_samples
    if (status > 0) st->last_status = rep;
    if (p.out_samples) p.out_samples[stream] = rep;
  } else if (status > 0) {
    if (st->last_status >= 0) st->last_status += rep;   // OPUS_GET_LAST_PACKET_DURATION
    if (p.out_samples && p.out_samples[stream] >= 0) p.out_samples[stream] += rep;
  }
}

// ------------------------------------------------------------------------------------------
// A lost frame (celt_decode_lost, celt/celt_decoder.c:679-1090): everything that is a DECISION or band-energy / range-coder state
// happens here, in stream order with the parsing of the neighbouring frames; the signal work is done by the later kernels:
//   noise PLC  -> the vector kernel draws the noise spectrum, the synthesis kernel runs unchanged on the decayed energies;
//   pitch PLC  -> ob_celt_dec_plc (celt_dec_plc.cuh) extrapolates the history ring.
// ------------------------------------------------------------------------------------------
OB_DEVNI void ent_lost_frame(const DecEntropyParams &p, int stream, ObFrameSide *side, ObDecBandState *st, int LM, int pcm_off, uint32_t fl,
                             int packet_lost) {
  const int CC = p.channels, N = OB_SHORT_N << LM, M = 1 << LM;
  side->pcm_off = pcm_off;
  side->LM = LM;
  st->final_range = 0;  // src/opus_decoder.c:701-704
  if (!(fl & OB_PF_HAS_PACKET)) {  // nothing decoded yet: zeros, the CELT state is not touched (src/opus_decoder.c:333-341)
    side->plc = OB_PLC_ZEROS | (packet_lost ? OB_PLC_PACKET : 0);
    st->plc_flags = fl;
    ent_finish(p, stream, side, st, N);
    return;
  }
  const uint32_t dur = st->plc_dur;
  const int loss_duration = dur & 0xffff, plc_duration = dur >> 16;
  const int noise = plc_duration >= 40 || (fl & OB_PF_SKIP_PLC);  // start == 0 always here
  if (noise) {
    const int end = (fl >> OB_PF_END_SHIFT) & 31;  // effEnd == end: the mode has 21 bands
    const float decay = loss_duration == 0 ? 1.5f : .5f;
#pragma unroll 1
    for (int c = 0; c < CC; c++) {
#pragma unroll 1
      for (int i = 0; i < end; i++) {
        const float a = st->backgroundLogE[c * OB_NBANDS + i], b = st->oldEBands[c * OB_NBANDS + i] - decay;
        st->oldEBands[c * OB_NBANDS + i] = a > b ? a : b;
      }
    }
#pragma unroll 1
    for (int i = 0; i < 2 * OB_NBANDS; i++) side->bandE[i] = st->oldEBands[i];
    // the LCG runs once per coefficient of every band below `end`, channel after channel (celt_decoder.c:773-790)
    uint32_t seed = st->rng;
    side->seed = seed;
    const int draws = CC * M * ob_ebands[end];
#pragma unroll 4
    for (int j = 0; j < draws; j++) seed = 1664525u * seed + 1013904223u;
    st->rng = seed;
    side->transient = 0; side->C = CC; side->silence = 0; side->end_band = end;
    side->anti_collapse_on = 0; side->n_leaves = 0;
    {  // post-filter with the last parameters (celt_decoder.c:794-808)
      const int T_old = OB_IMAX(st->pf_period_old, 15), T_cur = OB_IMAX(st->pf_period, 15);
      side->pf_T[0] = T_old; side->pf_T[1] = T_cur; side->pf_T[2] = T_cur;
      side->pf_g[0] = st->pf_gain_old; side->pf_g[1] = st->pf_gain; side->pf_g[2] = st->pf_gain;
      side->pf_tap[0] = st->pf_tapset_old; side->pf_tap[1] = st->pf_tapset; side->pf_tap[2] = st->pf_tapset;
      st->pf_period = T_cur;
      st->pf_period_old = T_cur; st->pf_gain_old = st->pf_gain; st->pf_tapset_old = st->pf_tapset;
    }
    side->plc = OB_PLC_NOISE | ((fl & OB_PF_FOLD) ? OB_PLC_FOLD : 0) | (packet_lost ? OB_PLC_PACKET : 0);
    fl = (fl & ~OB_PF_FOLD) | OB_PF_SKIP_PLC;
  } else {
    const int last_type = (fl >> OB_PF_TYPE_SHIFT) & 3;
    side->plc = OB_PLC_PITCH | (last_type != OB_FRAME_PLC_PERIODIC ? OB_PLC_FIRST : 0) | (packet_lost ? OB_PLC_PACKET : 0);
    fl |= OB_PF_FOLD;
  }
  fl = (fl & ~(3u << OB_PF_TYPE_SHIFT)) | ((uint32_t)(noise ? OB_FRAME_PLC_NOISE : OB_FRAME_PLC_PERIODIC) << OB_PF_TYPE_SHIFT);
  st->plc_flags = fl;
  st->plc_dur = (uint32_t)OB_IMIN(10000, loss_duration + M) | ((uint32_t)OB_IMIN(10000, plc_duration + M) << 16);
  ent_finish(p, stream, side, st, N);
}

// A lost PACKET (len == 0): opus_decode(NULL) conceals frame_size samples as a run of 20 / 10 / 5 / 2.5 ms frames
// (src/opus_decoder.c:757-778 and 316-365); launch k of the call produces the k-th of them.
OB_DEVNI void ent_lost_packet(const DecEntropyParams &p, int stream, ObFrameSide *side, ObDecBandState *st) {
  const int ds = p.downsample, F20 = 960 / ds, F10 = F20 >> 1, F5 = F10 >> 1, F2_5 = F5 >> 1, fsz = p.max_samples;
  const int k = p.frame_index;
  if (fsz % F2_5 != 0) { ent_finish(p, stream, side, st, k == 0 ? -1 : 0); return; }  // OPUS_BAD_ARG
  const uint32_t fl = st->plc_flags;
  if (!(fl & OB_PF_HAS_PACKET)) {
    if (k != 0) { ent_finish(p, stream, side, st, 0); return; }
    side->pcm_off = 0; side->LM = 0; side->plc = OB_PLC_ZEROS;
    st->final_range = 0;
    ent_finish(p, stream, side, st, fsz * ds);
    return;
  }
  int rem = fsz, a = 0;
#pragma unroll 1
  for (int idx = 0;; idx++) {
    if (rem <= 0) { ent_finish(p, stream, side, st, 0); return; }  // no k-th frame
    a = OB_IMIN(rem, (OB_SHORT_N << ((fl >> OB_PF_LM_SHIFT) & 3)) / ds);  // "don't conceal more than what the ToC says" (src/opus_decoder.c:316-321)
    if (a > F20) a = F20;
    else if (a < F20) {
      if (a > F10) a = F10;
      else if (a > F5 && a < F10) a = F5;
    }
    if (idx == k) break;
    rem -= a;
  }
  const int N48 = a * ds;
  const int LM = N48 == 120 ? 0 : N48 == 240 ? 1 : N48 == 480 ? 2 : 3;
  ent_lost_frame(p, stream, side, st, LM, fsz - rem, fl, 1);
}

// ------------------------------------------------------------------------------------------
// Kernel A1 body: one thread, one stream.
// ------------------------------------------------------------------------------------------
OB_DEV void dec_entropy_thread(const DecEntropyParams &p, int stream) {
  ObFrameSide *side = p.side + stream;
  ObDecBandState *st = p.state + stream;
  ObScript *sc = p.script + stream;
  const int CC = p.channels;
  const int plen = p.pkt_len[stream];
  if (plen < 0) { ent_finish(p, stream, side, st, 0); return; }   // this library's extension: the stream sits this call out
  if (plen == 0) { ent_lost_packet(p, stream, side, st); return; }  // opus_decode(data = NULL / len = 0): conceal frame_size samples
  if (plen > p.pkt_stride) { ent_finish(p, stream, side, st, -1); return; }   // the packet would run into the next stream's row: OPUS_BAD_ARG
  const unsigned char *pkt = p.packets + (size_t)stream * p.pkt_stride;
  const int toc = pkt[0];
  int foff, flen;
  const int count = locate_frame(pkt, plen, p.frame_index, &foff, &flen);
  const int LM = (toc >> 3) & 3;
  const int N = OB_SHORT_N << LM, M = 1 << LM;
  int status = 1;
  if (count < 0) status = count;                      // OPUS_INVALID_PACKET
  else if (!(toc & 0x80)) status = -5;                // SILK / hybrid: OPUS_UNIMPLEMENTED here
  else if (p.frame_index >= count) status = 0;        // no such frame in this packet
  if (status > 0 && count * (N / p.downsample) > p.max_samples) status = p.frame_index == 0 ? -2 : 0;  // BUFFER_TOO_SMALL: the whole packet (src/opus_decoder.c:806-807)
  if (status <= 0) { ent_finish(p, stream, side, st, status); return; }
  const uint32_t fl0 = st->plc_flags, dur0 = st->plc_dur;   // two loads; a loss-free stream stores neither back
  uint32_t fl = (fl0 & ~(3u << OB_PF_LM_SHIFT)) | ((uint32_t)LM << OB_PF_LM_SHIFT);  // Opus layer st->frame_size, per accepted packet (src/opus_decoder.c:842)
  if (flen <= 1) { ent_lost_frame(p, stream, side, st, LM, p.frame_index * (N / p.downsample), fl, 0); return; }  // DTX / lost frame inside the packet
  side->pcm_off = p.frame_index * (N / p.downsample);
  const int loss_dur = dur0 & 0xffff;
  if (loss_dur == 0) fl &= ~OB_PF_SKIP_PLC;  // celt_decoder.c:1299-1301

  const int C = (toc & 4) ? 2 : 1;
  const int bw = (toc >> 5) & 3;
  const int start = 0, end = bw == 0 ? 13 : bw == 1 ? 17 : bw == 2 ? 19 : 21;
  const int len = flen;

  Emit e;
  e.leaves = sc->leaves; e.n_leaves = 0;
  e.disable_inv = CC == 1 || p.disable_inv;
  e.remaining_bits = 0;
  RangeDec &dec = e.ec;
  dec.init(pkt + foff, len);

  int32_t total_bits = len * 8;
  int32_t tell = dec.tell();
  int silence = 0;
  if (tell >= total_bits) silence = 1;
  else if (tell == 1) silence = dec.bit_logp(15);
  if (silence) {
    tell = len * 8;
    dec.nbits_total += tell - dec.tell();
  }
  int pf_pitch = 0, pf_tapset = 0;
  float pf_gain = 0.f;
  if (start == 0 && tell + 16 <= total_bits) {
    if (dec.bit_logp(1)) {
      int octave = dec.uint(6);
      pf_pitch = (16 << octave) + dec.bits(4 + octave) - 1;
      int qg = dec.bits(3);
      if (dec.tell() + 2 <= total_bits) pf_tapset = dec.icdf(ob_tapset_icdf, 2);
      pf_gain = .09375f * (qg + 1);
    }
    tell = dec.tell();
  }
  int isTransient = 0;
  if (LM > 0 && tell + 3 <= total_bits) {
    isTransient = dec.bit_logp(3);
    tell = dec.tell();
  }
  const int shortBlocks = isTransient ? M : 0;
  const int intra = tell + 3 <= total_bits ? dec.bit_logp(3) : 0;

  // thread-private working arrays (local memory, L1 resident)
  float E[2 * OB_NBANDS];
  int pulses[OB_NBANDS];
  short offsets[OB_NBANDS];
  unsigned char ebits[OB_NBANDS];
  uint32_t fine_prio = 0;  // bit j = fine_priority[j]
  signed char tf_res[OB_NBANDS];

  // ---- coarse energy (celt/quant_bands.c:431-494) ----
  {
    const unsigned char *prob = ob_e_prob + (LM * 2 + intra) * 42;
    float prev0 = 0.f, prev1 = 0.f, coef, beta;
    if (intra) { coef = 0.f; beta = 4915 / 32768.f; }
    else {
      beta = LM == 0 ? 30147 / 32768.f : LM == 1 ? 22282 / 32768.f : LM == 2 ? 12124 / 32768.f : 6554 / 32768.f;
      coef = LM == 0 ? 29440 / 32768.f : LM == 1 ? 26112 / 32768.f : LM == 2 ? 21248 / 32768.f : 16384 / 32768.f;
    }
    const int32_t budget = dec.storage * 8;
    const float *oldg = st->oldEBands;
#pragma unroll 1
    for (int i = 0; i < 2 * OB_NBANDS; i++) E[i] = oldg[i];
    if (C == 1)  // celt_decoder.c:1306-1310
#pragma unroll 1
      for (int i = 0; i < OB_NBANDS; i++) E[i] = E[i] > E[OB_NBANDS + i] ? E[i] : E[OB_NBANDS + i];
    if (!intra && loss_dur != 0) {  // safe energy prediction after a loss (celt_decoder.c:1362-1393), rare
      const int missing = OB_IMIN(10, loss_dur >> LM);
      const float safety = LM == 0 ? 1.5f : LM == 1 ? .5f : 0.f;
#pragma unroll 1
      for (int ci = 0; ci < 2 * OB_NBANDS; ci++) {
        const int band = ci >= OB_NBANDS ? ci - OB_NBANDS : ci;
        if (band < start || band >= end) continue;
        const float l1 = st->oldLogE[ci], l2 = st->oldLogE2[ci];
        float E0 = E[ci];
        if (E0 < (l1 > l2 ? l1 : l2)) {  // going down already: continue the trend
          const float sa = l1 - E0, sb = .5f * (l2 - E0);
          float slope = sa > sb ? sa : sb;
          slope = slope < 2.f ? slope : 2.f;
          const float d = (1 + missing) * slope;
          E0 -= 0.f > d ? 0.f : d;
          E0 = -20.f > E0 ? -20.f : E0;
        } else {
          E0 = E0 < l1 ? E0 : l1;
          E0 = E0 < l2 ? E0 : l2;
        }
        E[ci] = E0 - safety;
      }
    }
#pragma unroll 1
    for (int i = start; i < end; i++) {
#pragma unroll 1
      for (int c = 0; c < C; c++) {
        int qi;
        int32_t t = dec.tell();
        if (budget - t >= 15) {
          int pi = 2 * OB_IMIN(i, 20);
          qi = dec.laplace(prob[pi] << 7, prob[pi + 1] << 6);
        } else if (budget - t >= 2) {
          qi = dec.icdf(ob_small_energy_icdf, 2);
          qi = (qi >> 1) ^ -(qi & 1);
        } else if (budget - t >= 1) {
          qi = -dec.bit_logp(1);
        } else qi = -1;
        float q = (float)qi;
        float o = E[i + c * OB_NBANDS];
        o = o > -9.f ? o : -9.f;
        float prev = c ? prev1 : prev0;
        float tmp = coef * o + prev + q;
        prev = prev + q - beta * q;
        if (c) prev1 = prev; else prev0 = prev;
        E[i + c * OB_NBANDS] = tmp;
      }
    }
  }
  // ---- tf_decode (celt/celt_decoder.c:513-547) ----
  {
    int curr = 0, tf_select = 0, tf_changed = 0;
    uint32_t budget = dec.storage * 8, t = dec.tell();
    int logp = isTransient ? 2 : 4;
    int tf_select_rsv = LM > 0 && t + logp + 1 <= budget;
    budget -= tf_select_rsv;
    uint32_t raw = 0;
#pragma unroll 1
    for (int i = start; i < end; i++) {
      if (t + logp <= budget) {
        curr ^= dec.bit_logp(logp);
        t = dec.tell();
        tf_changed |= curr;
      }
      raw |= (uint32_t)curr << i;
      logp = isTransient ? 4 : 5;
    }
    if (tf_select_rsv && ob_tf_select[LM * 8 + 4 * isTransient + 0 + tf_changed] !=
                             ob_tf_select[LM * 8 + 4 * isTransient + 2 + tf_changed])
      tf_select = dec.bit_logp(1);
#pragma unroll 1
    for (int i = 0; i < OB_NBANDS; i++)
      tf_res[i] = (i >= start && i < end) ? ob_tf_select[LM * 8 + 4 * isTransient + 2 * tf_select + ((raw >> i) & 1)] : 0;
  }
  tell = dec.tell();
  int spread = 2;
  if (tell + 4 <= total_bits) spread = dec.icdf(ob_spread_icdf, 5);
#define OB_CAP(i) ((ob_cache_caps[OB_NBANDS * (2 * LM + C - 1) + (i)] + 64) * C * ((ob_ebands[(i) + 1] - ob_ebands[(i)]) << LM) >> 2)
  // ---- dynalloc (celt/celt_decoder.c:1412-1442) ----
  {
    int dynalloc_logp = 6;
    total_bits <<= OB_BITRES;
    int32_t tf = dec.tell_frac();
#pragma unroll 1
    for (int i = 0; i < OB_NBANDS; i++) offsets[i] = 0;
#pragma unroll 1
    for (int i = start; i < end; i++) {
      int width = C * (ob_ebands[i + 1] - ob_ebands[i]) << LM;
      int quanta = OB_IMIN(width << OB_BITRES, OB_IMAX(6 << OB_BITRES, width));
      int loop_logp = dynalloc_logp, boost = 0;
      const int capi = OB_CAP(i);
#pragma unroll 1
      while (tf + (loop_logp << OB_BITRES) < total_bits && boost < capi) {
        int flag = dec.bit_logp(loop_logp);
        tf = dec.tell_frac();
        if (!flag) break;
        boost += quanta;
        total_bits -= quanta;
        loop_logp = 1;
      }
      offsets[i] = (short)boost;
      if (boost > 0) dynalloc_logp = OB_IMAX(2, dynalloc_logp - 1);
    }
    tell = tf;
  }
  const int alloc_trim = tell + (6 << OB_BITRES) <= total_bits ? dec.icdf(ob_trim_icdf, 7) : 5;
  int32_t bits = (((int32_t)len * 8) << OB_BITRES) - (int32_t)dec.tell_frac() - 1;
  const int anti_collapse_rsv = isTransient && LM >= 2 && bits >= ((LM + 2) << OB_BITRES) ? (1 << OB_BITRES) : 0;
  bits -= anti_collapse_rsv;

  // ---- bit allocation (celt/rate.c:535-646 and 249-533).  Per-band terms that the reference keeps
  // in temporary arrays (thresh, trim_offset, bits1, bits2) are recomputed where they are used. ----
  int intensity = 0, dual_stereo = 0, codedBands;
  int32_t balance = 0;
  {
    const int len_b = OB_NBANDS, nbAllocVectors = 11;
    int32_t total = OB_IMAX(bits, 0);
    int skip_start = start, intensity_rsv = 0, dual_stereo_rsv = 0;
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
#define OB_W(j) (ob_ebands[(j) + 1] - ob_ebands[(j)])
#define OB_THRESH(j) OB_IMAX((C) << OB_BITRES, (3 * OB_W(j) << LM << OB_BITRES) >> 4)
#define OB_TRIM(j) ((C * OB_W(j) * (alloc_trim - 5 - LM) * (end - (j) - 1) * (1 << (LM + OB_BITRES)) >> 6) - \
                    ((OB_W(j) << LM) == 1 ? (C << OB_BITRES) : 0))
    int lo = 1, hi = nbAllocVectors - 1;
#pragma unroll 1
    do {
      int done = 0, psum = 0, mid = (lo + hi) >> 1;
#pragma unroll 1
      for (int j = end; j-- > start;) {
        int bitsj = C * OB_W(j) * ob_band_alloc[mid * len_b + j] << LM >> 2;
        if (bitsj > 0) bitsj = OB_IMAX(0, bitsj + OB_TRIM(j));
        bitsj += offsets[j];
        if (bitsj >= OB_THRESH(j) || done) { done = 1; psum += OB_IMIN(bitsj, OB_CAP(j)); }
        else if (bitsj >= C << OB_BITRES) psum += C << OB_BITRES;
      }
      if (psum > total) hi = mid - 1; else lo = mid + 1;
    } while (lo <= hi);
    hi = lo--;
#pragma unroll 1
    for (int j = start; j < end; j++) if (offsets[j] > 0) skip_start = j;
    // bits1[j], bits2[j] as functions (rate.c:606-628)
#define OB_BITS1(j, out1, out2) do { \
      int b1_ = C * OB_W(j) * ob_band_alloc[lo * len_b + (j)] << LM >> 2; \
      int b2_ = hi >= nbAllocVectors ? OB_CAP(j) : C * OB_W(j) * ob_band_alloc[hi * len_b + (j)] << LM >> 2; \
      int tr_ = OB_TRIM(j); \
      if (b1_ > 0) b1_ = OB_IMAX(0, b1_ + tr_); \
      if (b2_ > 0) b2_ = OB_IMAX(0, b2_ + tr_); \
      if (lo > 0) b1_ += offsets[j]; \
      b2_ += offsets[j]; \
      b2_ = OB_IMAX(0, b2_ - b1_); \
      out1 = b1_; out2 = b2_; } while (0)
    // ---- interp_bits2pulses ----
    int32_t psum;
    const int logM = LM << OB_BITRES, stereo = C > 1, alloc_floor = C << OB_BITRES;
    int ilo = 0, ihi = 1 << 6;
#pragma unroll 1
    for (int it = 0; it < 6; it++) {
      int mid = (ilo + ihi) >> 1, done = 0;
      psum = 0;
#pragma unroll 1
      for (int j = end; j-- > start;) {
        int b1, b2;
        OB_BITS1(j, b1, b2);
        int tmp = b1 + (mid * (int32_t)b2 >> 6);
        if (tmp >= OB_THRESH(j) || done) { done = 1; psum += OB_IMIN(tmp, OB_CAP(j)); }
        else if (tmp >= alloc_floor) psum += alloc_floor;
      }
      if (psum > total) ihi = mid; else ilo = mid;
    }
    psum = 0;
    {
      int done = 0;
#pragma unroll 1
      for (int j = end; j-- > start;) {
        int b1, b2;
        OB_BITS1(j, b1, b2);
        int tmp = b1 + ((int32_t)ilo * b2 >> 6);
        if (tmp < OB_THRESH(j) && !done) tmp = tmp >= alloc_floor ? alloc_floor : 0;
        else done = 1;
        tmp = OB_IMIN(tmp, OB_CAP(j));
        pulses[j] = tmp;
        psum += tmp;
      }
    }
#pragma unroll 1
    for (codedBands = end;; codedBands--) {
      int j = codedBands - 1;
      if (j <= skip_start) { total += skip_rsv; break; }
      int32_t left = total - psum;
      int32_t percoeff = (uint32_t)left / (uint32_t)(ob_ebands[codedBands] - ob_ebands[start]);
      left -= (ob_ebands[codedBands] - ob_ebands[start]) * percoeff;
      int rem = OB_IMAX(left - (ob_ebands[j] - ob_ebands[start]), 0);
      int band_width = ob_ebands[codedBands] - ob_ebands[j];
      int bj = pulses[j];
      int band_bits = (int)(bj + percoeff * band_width + rem);
      if (band_bits >= OB_IMAX(OB_THRESH(j), alloc_floor + (1 << OB_BITRES))) {
        if (dec.bit_logp(1)) break;
        psum += 1 << OB_BITRES;
        band_bits -= 1 << OB_BITRES;
      }
      psum -= bj + intensity_rsv;
      if (intensity_rsv > 0) intensity_rsv = ob_log2_frac_tab[j - start];
      psum += intensity_rsv;
      if (band_bits >= alloc_floor) { psum += alloc_floor; pulses[j] = alloc_floor; }
      else pulses[j] = 0;
    }
    if (intensity_rsv > 0) intensity = start + dec.uint(codedBands + 1 - start);
    else intensity = 0;
    if (intensity <= start) { total += dual_stereo_rsv; dual_stereo_rsv = 0; }
    if (dual_stereo_rsv > 0) dual_stereo = dec.bit_logp(1);
    else dual_stereo = 0;
    int32_t left = total - psum;
    int32_t percoeff = (uint32_t)left / (uint32_t)(ob_ebands[codedBands] - ob_ebands[start]);
    left -= (ob_ebands[codedBands] - ob_ebands[start]) * percoeff;
    // two reference loops fused: the serial dependence is only through `left` and `balance`, in band order
    int32_t bal = 0;
    // the reference first adds percoeff*w to all bands, then distributes `left` from the lowest band up
#pragma unroll 1
    for (int j = start; j < codedBands; j++) {
      int w = OB_W(j);
      int bj = pulses[j] + (int)percoeff * w;
      int tmp = (int)OB_IMIN(left, w);
      bj += tmp;
      left -= tmp;
      int N0 = w, Nn = N0 << LM, ej, fp;
      int32_t excess, bit = (int32_t)bj + bal;
      if (Nn > 1) {
        excess = OB_IMAX(bit - OB_CAP(j), 0);
        bj = bit - excess;
        int den = (C * Nn + ((C == 2 && Nn > 2 && !dual_stereo && j < intensity) ? 1 : 0));
        int NClogN = den * (ob_logN[j] + logM);
        int offset = (NClogN >> 1) - den * 21;
        if (Nn == 2) offset += den << OB_BITRES >> 2;
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
        fp = extra_bits >= excess - bal;
        excess -= extra_bits;
      }
      bal = excess;
      pulses[j] = bj; ebits[j] = (unsigned char)ej;
      fine_prio |= (uint32_t)(fp & 1) << j;
    }
    balance = bal;
#pragma unroll 1
    for (int j = codedBands; j < end; j++) {
      int ej = pulses[j] >> stereo >> OB_BITRES;
      ebits[j] = (unsigned char)ej;
      pulses[j] = 0;
      fine_prio |= (uint32_t)(ej < 1) << j;
    }
  }
  e.intensity = intensity;
  // ---- fine energy (celt/quant_bands.c:496-523) ----
#pragma unroll 1
  for (int i = start; i < end; i++) {
    int extra = ebits[i];
    if (extra <= 0) continue;
    if (dec.tell() + C * extra > (int32_t)dec.storage * 8) continue;
#pragma unroll 1
    for (int c = 0; c < C; c++) {
      int q2 = dec.bits(extra);
      float offset = (q2 + .5f) * (1 << (14 - extra)) * (1.f / 16384) - .5f;
      E[i + c * OB_NBANDS] += offset;
    }
  }
  // ---- band loop: control side of celt/bands.c:1589-1923, 1387-1570, 1248-1379, 973-1176 ----
  // Written as ONE flat loop whose iteration is either "start the next band" or "process one node of
  // the current part's partition tree".  The threads of a warp are different streams with different
  // trees; a nested band/part/node loop would make them wait for each other at every band and part
  // boundary, whereas here they reconverge once per iteration no matter which band they are in.
  {
    const int B_frame = shortBlocks ? M : 1;
    const int norm_offset = M * ob_ebands[start];
    const int32_t total_b = len * (8 << OB_BITRES) - anti_collapse_rsv;
    int lowband_offset = 0, update_lowband = 1;
    int ds = dual_stereo;
    int i = start;
    int in_band = 0;           // 0: start band i next; 1: walking a part
    PFrame st[6];
    int sp = 0;
    // per-band context
    ObBandRec br;
    int32_t tl = 0, rebal0 = 0;
    int b = 0, Nb = 0, xoff = 0, lb = 0, tf_change = 0, leaf_mid = -1;
    int nparts = 0, part = 0;  // parts of the current band and the one being walked
    int p2_off = 0, p2_bits = 0, p1_bits = 0, p2_lb = 0, p2_itheta = 0, p2_is_mid = 0;
    float p2_gain = 1.f;
    unsigned flags = 0;
    br.mid = 1.f; br.side = 0.f;
    // start part (off, bits, gain, lb): N == 1 is a leaf by itself (quant_band_n1), otherwise the TF
    // changes fix the block count and the tree walk begins (celt/bands.c:1262-1313)
#define OB_START_PART(OFF, BITS, GAIN, LBF) do { \
      if (Nb == 1) { \
        int sign_ = 0; \
        if (e.remaining_bits >= 1 << OB_BITRES) { sign_ = dec.bits(1); e.remaining_bits -= 1 << OB_BITRES; } \
        emit_leaf(e, (uint32_t)sign_, (GAIN), (OFF), 1, 1, 1, 2, 0, 0, 0, 0); \
        sp = 0; \
      } else { \
        int Bp_ = B_frame, tfc_ = tf_change; \
        int N_B_ = (unsigned)Nb / (unsigned)Bp_; \
        int rec_ = tfc_ > 0 ? tfc_ : 0; \
        Bp_ >>= rec_; N_B_ <<= rec_; \
        while ((N_B_ & 1) == 0 && tfc_ < 0) { Bp_ <<= 1; N_B_ >>= 1; tfc_++; } \
        st[0].off = (short)(OFF); st[0].N = (short)Nb; st[0].b = (BITS); st[0].B = (unsigned char)Bp_; \
        st[0].LM = (signed char)LM; st[0].gain = (GAIN); st[0].lb = (unsigned char)(LBF); st[0].stage = 0; \
        st[0].path = 0; st[0].depth = 0; st[0].cmshift = 0; \
        sp = 1; \
      } } while (0)
#pragma unroll 1
    while (i < end) {
      if (!in_band) {
        // ------------------------------ start band i ------------------------------
        e.band = i;
        Nb = M * ob_ebands[i + 1] - M * ob_ebands[i];
        xoff = M * ob_ebands[i];
        tl = dec.tell_frac();
        if (i != start) balance -= tl;
        const int32_t remaining_bits = total_b - tl - 1;
        e.remaining_bits = remaining_bits;
        if (i <= codedBands - 1) {
          int32_t curr_balance = balance / OB_IMIN(3, codedBands - i);
          b = OB_IMAX(0, OB_IMIN(16383, OB_IMIN(remaining_bits + 1, pulses[i] + curr_balance)));
        } else b = 0;
        if ((M * ob_ebands[i] - Nb >= M * ob_ebands[start] || i == start + 1) && (update_lowband || lowband_offset == 0))
          lowband_offset = i;
        tf_change = tf_res[i];
        int effective_lowband = -1;
        if (lowband_offset != 0 && (spread != 3 || B_frame > 1 || tf_change < 0))
          effective_lowband = OB_IMAX(0, M * ob_ebands[lowband_offset] - norm_offset - Nb);
        flags = 0;
        if (ds && i == intensity) { ds = 0; flags |= 2; }
        br.eff_lowband = (int16_t)effective_lowband;
        br.lowband_offset = (uint8_t)lowband_offset;
        br.leaf_begin = (uint16_t)OB_IMIN(e.n_leaves, 65535);
        br.theta_class = 0; br.pad = 0; br.mid = 1.f; br.side = 0.f; br.b = b;
        lb = effective_lowband != -1;
        leaf_mid = -1;
        part = 0;
        in_band = 1;
        if (ds) {
          flags |= 1;
          nparts = 2;
          p1_bits = b / 2; p2_bits = b / 2; p2_off = N + xoff; p2_gain = 1.0f; p2_lb = lb; p2_itheta = -1;
          OB_START_PART(xoff, b / 2, 1.0f, lb);
        } else if (C == 2) {
          // control side of quant_band_stereo, celt/bands.c:1387-1570
          flags |= 4;
          if (Nb == 1) {
            int s0 = 0, s1 = 0;
            if (e.remaining_bits >= 1 << OB_BITRES) { s0 = dec.bits(1); e.remaining_bits -= 1 << OB_BITRES; }
            if (e.remaining_bits >= 1 << OB_BITRES) { s1 = dec.bits(1); e.remaining_bits -= 1 << OB_BITRES; }
            emit_leaf(e, (uint32_t)(s0 | s1 << 1), 1.0f, xoff, 1, 2, 1, 2, 0, 0, 0, 0);
            nparts = 1;
            sp = 0;
          } else {
            ThetaRet tr = ent_theta(dec, i, intensity, e.disable_inv, e.remaining_bits, Nb, b, B_frame, LM, 1);
            dec = tr.ec;
            const int bb = tr.t.b;
            const int itheta = tr.t.itheta, delta = tr.t.delta, qalloc = tr.t.qalloc;
            br.theta_class = (uint8_t)theta_class(itheta);
            br.mid = (1.f / 32768) * tr.t.imid;
            br.side = (1.f / 32768) * tr.t.iside;
            if (tr.t.inv) flags |= 8;
            if (Nb == 2) {
              int mbits = bb, sbits = 0, sign = 0;
              if (itheta != 0 && itheta != 16384) sbits = 1 << OB_BITRES;
              mbits -= sbits;
              const int c = itheta > 8192;
              e.remaining_bits -= qalloc + sbits;
              if (sbits) sign = dec.bits(1);
              flags |= 16 | (c ? 32 : 0) | (sign ? 64 : 0);
              nparts = 1;
              OB_START_PART((c ? N : 0) + xoff, mbits, 1.0f, lb);  // the mid goes to x2 = c ? Y : X
            } else {
              int mbits = OB_IMAX(0, OB_IMIN(bb, (bb - delta) / 2));
              int sbits = bb - mbits;
              e.remaining_bits -= qalloc;
              rebal0 = e.remaining_bits;
              nparts = 2;
              p2_itheta = itheta;
              if (mbits >= sbits) {
                p1_bits = mbits; p2_bits = sbits; p2_off = N + xoff; p2_gain = br.side; p2_lb = 0; p2_is_mid = 0;
                OB_START_PART(xoff, mbits, 1.0f, lb);
              } else {
                flags |= 128;
                p1_bits = sbits; p2_bits = mbits; p2_off = xoff; p2_gain = 1.0f; p2_lb = lb; p2_is_mid = 1;
                OB_START_PART(N + xoff, sbits, br.side, 0);
              }
            }
          }
        } else {
          nparts = 1;
          OB_START_PART(xoff, b, 1.0f, lb);
        }
      } else if (sp > 0) {
        // ------------------------------ one tree node ------------------------------
        PFrame &f = st[sp - 1];
        if (f.stage == 0) {
          int Np = f.N, LMp = f.LM, bp = f.b, Bp = f.B;
          const unsigned char *cache = pulse_cache_row(i, LMp);
          if (LMp != -1 && bp > cache[cache[0]] + 12 && Np > 2) {
            const int B0 = Bp;
            Np >>= 1;
            LMp -= 1;
            Bp = (Bp + 1) >> 1;
            ThetaRet tr = ent_theta(dec, i, intensity, e.disable_inv, e.remaining_bits, Np, bp, B0, LMp, 0);
            dec = tr.ec;
            bp = tr.t.b;
            int delta = tr.t.delta;
            const int itheta = tr.t.itheta;
            if (B0 > 1 && (itheta & 0x3fff)) {
              if (itheta > 8192) delta -= delta >> (4 - LMp);
              else delta = OB_IMIN(0, delta + (Np << OB_BITRES >> (5 - LMp)));
            }
            int mbits = OB_IMAX(0, OB_IMIN(bp, (bp - delta) / 2));
            int sbits = bp - mbits;
            e.remaining_bits -= tr.t.qalloc;
            f.mid = (1.f / 32768) * tr.t.imid;
            f.side = (1.f / 32768) * tr.t.iside;
            f.mbits = mbits; f.sbits = sbits; f.rebalance = e.remaining_bits; f.itheta = (short)itheta;
            f.mid_first = mbits >= sbits;
            f.N = (short)Np; f.LM = (signed char)LMp; f.B = (unsigned char)Bp; f.B0 = (unsigned char)B0;
            f.stage = 1;
            PFrame &c = st[sp++];
            c.stage = 0; c.N = (short)Np; c.LM = (signed char)LMp; c.B = (unsigned char)Bp; c.lb = f.lb;
            c.depth = (unsigned char)(f.depth + 1);
            const unsigned cls = (unsigned)theta_class(itheta) << (3 * f.depth);
            if (f.mid_first) {
              c.off = f.off; c.b = mbits; c.gain = f.gain * f.mid;
              c.path = (unsigned short)(f.path | cls); c.cmshift = f.cmshift;
            } else {
              c.off = (short)(f.off + Np); c.b = sbits; c.gain = f.gain * f.side;
              c.path = (unsigned short)(f.path | cls | 4u << (3 * f.depth)); c.cmshift = (unsigned char)(f.cmshift + (B0 >> 1));
            }
          } else {
            int q = bits_to_pseudo(cache, bp);
            int curr_bits = q == 0 ? 0 : cache[q] + 1;
            e.remaining_bits -= curr_bits;
#pragma unroll 1
            while (e.remaining_bits < 0 && q > 0) {
              e.remaining_bits += curr_bits;
              q--;
              curr_bits = q == 0 ? 0 : cache[q] + 1;
              e.remaining_bits -= curr_bits;
            }
            if (q != 0) {
              const int K = pseudo_to_pulses(q);
              uint32_t idx = dec.uint(pvq_u(Np, K) + pvq_u(Np, K + 1));
              emit_leaf(e, idx, f.gain, f.off, Np, K, Bp, 0, f.lb, f.path, f.depth, f.cmshift);
            } else {
              emit_leaf(e, 0, f.gain, f.off, Np, 0, Bp, 1, f.lb, f.path, f.depth, f.cmshift);
            }
            sp--;
          }
        } else if (f.stage == 1) {
          // first child done: re-balance and descend into the other half (bands.c:1104-1122)
          int rebalance;
          PFrame &c = st[sp];
          c.stage = 0; c.N = f.N; c.LM = f.LM; c.B = f.B; c.lb = f.lb;
          c.depth = (unsigned char)(f.depth + 1);
          const unsigned cls = (unsigned)theta_class(f.itheta) << (3 * f.depth);
          if (f.mid_first) {
            rebalance = f.mbits - (f.rebalance - e.remaining_bits);
            if (rebalance > 3 << OB_BITRES && f.itheta != 0) f.sbits += rebalance - (3 << OB_BITRES);
            c.off = (short)(f.off + f.N); c.b = f.sbits; c.gain = f.gain * f.side;
            c.path = (unsigned short)(f.path | cls | 4u << (3 * f.depth)); c.cmshift = (unsigned char)(f.cmshift + (f.B0 >> 1));
          } else {
            rebalance = f.sbits - (f.rebalance - e.remaining_bits);
            if (rebalance > 3 << OB_BITRES && f.itheta != 16384) f.mbits += rebalance - (3 << OB_BITRES);
            c.off = f.off; c.b = f.mbits; c.gain = f.gain * f.mid;
            c.path = (unsigned short)(f.path | cls); c.cmshift = f.cmshift;
          }
          f.stage = 2;
          sp++;
        } else {
          sp--;
        }
      }
      if (in_band && sp == 0) {
        // ------------------------------ a part is complete ------------------------------
        part++;
        if (part < nparts) {
          leaf_mid = e.n_leaves;
          int bits2 = p2_bits;
          if (p2_itheta >= 0) {  // stereo split: give unused bits of the first part to the second (bands.c:1523-1547)
            int32_t rebalance = p1_bits - (rebal0 - e.remaining_bits);
            if (rebalance > 3 << OB_BITRES && p2_itheta != (p2_is_mid ? 16384 : 0)) bits2 += rebalance - (3 << OB_BITRES);
          }
          OB_START_PART(p2_off, bits2, p2_gain, p2_lb);
        }
        if (part >= nparts && sp == 0) {
          br.flags = (uint8_t)flags;
          br.leaf_end = (uint16_t)OB_IMIN(e.n_leaves, 65535);
          br.leaf_mid = (uint16_t)(leaf_mid < 0 ? br.leaf_end : OB_IMIN(leaf_mid, 65535));
          sc->bands[i] = br;
          balance += pulses[i] + tl;
          update_lowband = b > (Nb << OB_BITRES);
          in_band = 0;
          i++;
        }
      }
    }
#undef OB_START_PART
  }
  int anti_collapse_on = 0;
  if (anti_collapse_rsv > 0) anti_collapse_on = dec.bits(1);
  // ---- unquant_energy_finalise (celt/quant_bands.c:525-547) ----
  {
    int bits_left = len * 8 - dec.tell();
#pragma unroll 1
    for (int prio = 0; prio < 2; prio++) {
#pragma unroll 1
      for (int i = start; i < end && bits_left >= C; i++) {
        int fq = ebits[i];
        if (fq >= 8 || (int)((fine_prio >> i) & 1) != prio) continue;
#pragma unroll 1
        for (int c = 0; c < C; c++) {
          int q2 = dec.bits(1);
          float offset = (q2 - .5f) * (1 << (14 - fq - 1)) * (1.f / 16384);
          E[i + c * OB_NBANDS] += offset;
          bits_left--;
        }
      }
    }
  }
  // ---- side info for the vector / synthesis stages ----
#pragma unroll 1
  for (int i = 0; i < 2 * OB_NBANDS; i++) {
    const int band = i >= OB_NBANDS ? i - OB_NBANDS : i, c = i >= OB_NBANDS;
    float v = E[i];
    if (C == 1 && c == 1) v = E[band];
    if (silence && (i < C * OB_NBANDS || C == 1)) v = -28.f;
    side->bandE[i] = v;
  }
#pragma unroll 1
  for (int i = 0; i < OB_NBANDS; i++) { side->tf_res[i] = tf_res[i]; side->pulses[i] = (int16_t)pulses[i]; }
  const int overflow = e.n_leaves > OB_MAX_LEAVES;
  const int ok = (dec.tell() > 8 * len || overflow) ? -3 : N;  // OPUS_INTERNAL_ERROR
  side->LM = LM; side->transient = isTransient; side->C = C; side->silence = silence; side->end_band = end;
  side->spread = spread; side->intensity = intensity; side->dual_stereo = dual_stereo;
  side->anti_collapse_on = anti_collapse_on; side->n_leaves = e.n_leaves;
  side->plc = (fl & OB_PF_FOLD) ? OB_PLC_FOLD : 0;
  side->seed = st->rng;
  {  // post-filter parameter pipeline (celt_decoder.c:1540-1563)
    int T_old = OB_IMAX(st->pf_period_old, 15), T_cur = OB_IMAX(st->pf_period, 15);
    side->pf_T[0] = T_old; side->pf_T[1] = T_cur; side->pf_T[2] = pf_pitch;
    side->pf_g[0] = st->pf_gain_old; side->pf_g[1] = st->pf_gain; side->pf_g[2] = pf_gain;
    side->pf_tap[0] = st->pf_tapset_old; side->pf_tap[1] = st->pf_tapset; side->pf_tap[2] = pf_tapset;
    float g_cur = st->pf_gain;
    int tap_cur = st->pf_tapset;
    st->pf_period_old = T_cur; st->pf_gain_old = g_cur; st->pf_tapset_old = tap_cur;
    st->pf_period = pf_pitch; st->pf_gain = pf_gain; st->pf_tapset = pf_tapset;
    if (LM != 0) { st->pf_period_old = pf_pitch; st->pf_gain_old = pf_gain; st->pf_tapset_old = pf_tapset; }
  }
  st->rng = dec.rng;
  st->final_range = dec.rng;
  if (dec.error) st->error = 1;
  if (ok > 0) {
    // energy-state update (celt/celt_decoder.c:1575-1600).  It only needs the decoded energies, so it lives here and not behind the
    // vector stage: frame t+1 can be parsed while frame t's PVQ / synthesis kernels still run.  Anti-collapse (vector stage) wants
    // the log energies of the two previous frames as they were before the update: handed over in the side info.
    const int M = 1 << LM;
#pragma unroll 1
    for (int i = 0; i < 2 * OB_NBANDS; i++) {
      const int band = i >= OB_NBANDS ? i - OB_NBANDS : i;
      float eb = side->bandE[i];
      float l1 = st->oldLogE[i], l2 = st->oldLogE2[i], bg = st->backgroundLogE[i];
      if (anti_collapse_on) { side->prevLogE[0][i] = l1; side->prevLogE[1][i] = l2; }
      if (!isTransient) { l2 = l1; l1 = eb; }
      else l1 = l1 < eb ? l1 : eb;
      const float inc = OB_IMIN(160, loss_dur + M) * 0.001f;
      const float a = bg + inc;
      bg = a < eb ? a : eb;
      if (band >= end) { eb = 0.f; l1 = l2 = -28.f; }
      st->oldEBands[i] = eb;
      st->oldLogE[i] = l1;
      st->oldLogE2[i] = l2;
      st->backgroundLogE[i] = bg;
    }
  }
  // celt_decoder.c:1604-1607: loss counters cleared, FRAME_NORMAL, no fold pending; plus the Opus layer's prev_mode / end band
  fl = (fl & ~(OB_PF_FOLD | (3u << OB_PF_TYPE_SHIFT) | (31u << OB_PF_END_SHIFT))) | OB_PF_HAS_PACKET |
       ((uint32_t)OB_FRAME_NORMAL << OB_PF_TYPE_SHIFT) | ((uint32_t)end << OB_PF_END_SHIFT);
  if (fl != fl0) st->plc_flags = fl;
  if (dur0 != 0) st->plc_dur = 0;
  ent_finish(p, stream, side, st, ok);
}

}  // namespace ob
