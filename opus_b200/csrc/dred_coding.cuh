// dred_coding.cuh — device code of the DRED encoder's entropy coding (see dred_coding_host.cuh): one thread codes one stream's DRED payload.
#pragma once
#include "celt_enc_frame.cuh"
#include "nnet.cuh"

namespace ob {
struct DredCodeParams {
  const float *lat_ring, *state_ring;   // [n][104 * dim], entry (head + k) % 104 is the reference's index k
  const uint8_t *stats[8];              // dred_rdovae_stats_data.c: latent quant_scales / dead_zone / r / p0, then the state's
  const uint8_t *activity;              // [n][416]: activity_mem, one byte per 2.5 ms, newest first
  int32_t *last_extra;                  // [n]: last_extra_dred_offset
  unsigned char *buf;                   // [n][buf_stride]
  int32_t *nbytes;                      // [n]: the return value (0 = no DRED for this packet)
  int n, buf_stride, head, latent_dim, state_dim, latents_buffer_fill, latent_offset, dred_offset, max_chunks, max_bytes, q0, dQ, qmax;
};
constexpr int kDredFrames = 104;
OB_DEV void re_icdf16(RangeEnc &e, int s, const uint16_t *icdf, unsigned ftb) {
  const uint32_t r = e.rng >> ftb;
  if (s > 0) {
    e.val += e.rng - r * icdf[s - 1];
    e.rng = r * (uint32_t)(icdf[s - 1] - icdf[s]);
  } else e.rng -= r * icdf[s];
  re_normalize(e);
}
OB_DEV void re_shrink(RangeEnc &e, uint32_t size) {   // ec_enc_shrink (celt/entenc.c:263-269); no raw bits are in flight on this path
  for (uint32_t i = 0; i < e.end_offs; i++) e.buf[size - e.end_offs + i] = e.buf[e.storage - e.end_offs + i];
  e.storage = size;
}
OB_DEVNI void dred_laplace_p0(RangeEnc &e, int value, uint16_t p0, uint16_t decay) {
  uint16_t sign_icdf[3], icdf[8];
  sign_icdf[0] = (uint16_t)(32768 - p0);
  sign_icdf[1] = sign_icdf[0] / 2;
  sign_icdf[2] = 0;
  re_icdf16(e, value == 0 ? 0 : (value > 0 ? 1 : 2), sign_icdf, 15);
  value = value < 0 ? -value : value;
  if (value) {
    icdf[0] = 7 > decay ? 7 : decay;
    for (int i = 1; i < 7; i++) {
      const int v = (icdf[i - 1] * (int32_t)decay) >> 15;
      icdf[i] = (uint16_t)(7 - i > v ? 7 - i : v);
    }
    icdf[7] = 0;
    value--;
    do {
      re_icdf16(e, value < 7 ? value : 7, icdf, 15);
      value -= 7;
    } while (value >= 0);
  }
}
OB_DEVNI void dred_encode_latents(RangeEnc &e, const float *x, const uint8_t *scale, const uint8_t *dzone, const uint8_t *r, const uint8_t *p0, int dim) {
#pragma unroll 1
  for (int i = 0; i < dim; i++) {
    const float delta = dzone[i] * (1.f / 256.f);
    float xq = x[i] * scale[i] * (1.f / 256.f);
    const float dz = nn_tanh(xq / (delta + .1f));
    xq = xq - delta * dz;
    const int q = (int)floorf(.5f + xq);
    if (!(r[i] == 0 || p0[i] == 255)) dred_laplace_p0(e, q, (uint16_t)(p0[i] << 7), (uint16_t)(r[i] << 7));
  }
}
OB_DEV int dred_voice_active(const uint8_t *activity_mem, int offset) {
  for (int i = 0; i < 16; i++) {
    if (8 * offset + i >= kDredFrames * 4) break;
    if (activity_mem[8 * offset + i] == 1) return 1;
  }
  return 0;
}
OB_DEV int dred_quantizer(int q0, int dQ, int qmax, int i) {
  const int tab[8] = {0, 2, 3, 4, 6, 8, 12, 16};
  const int quant = q0 + (tab[dQ] * i + 8) / 16;
  return quant > qmax ? qmax : quant;
}
OB_DEV int dred_encode_silk_frame_thread(const DredCodeParams &p, int s) {
  const uint8_t *act = p.activity + (size_t)s * kDredFrames * 4;
  const float *lat = p.lat_ring + (size_t)s * kDredFrames * p.latent_dim, *st = p.state_ring + (size_t)s * kDredFrames * p.state_dim;
  int latent_offset = p.latent_offset, extra = 0, encoded = 0, delayed = 0, prev_active = 0;
  int last_extra = p.last_extra[s];
  if (act[0] && last_extra > 0) {
    latent_offset = last_extra;
    delayed = 1;
    last_extra = 0;
  }
  while (latent_offset < p.latents_buffer_fill - 1 && !dred_voice_active(act, latent_offset)) {
    latent_offset++;
    extra++;
  }
  if (!delayed) last_extra = extra;
  p.last_extra[s] = last_extra;
  RangeEnc ec, bak;
  re_init(ec, p.buf + (size_t)s * p.buf_stride, (uint32_t)p.max_bytes);
  re_uint(ec, p.q0, 16);
  re_uint(ec, p.dQ, 8);
  const int total_offset = 16 - (p.dred_offset - extra * 8);
  if (total_offset > 31) {
    re_uint(ec, 1, 2);
    re_uint(ec, total_offset >> 5, 256);
    re_uint(ec, total_offset & 31, 32);
  } else {
    re_uint(ec, 0, 2);
    re_uint(ec, total_offset, 32);
  }
  if (p.q0 < 14 && p.dQ > 0) {
    const int nvals = 15 - (p.q0 + 1);
    re_encode(ec, p.qmax >= 15 ? 0 : nvals + p.qmax - (p.q0 + 1), p.qmax >= 15 ? nvals : nvals + p.qmax - p.q0, 2 * nvals);
  }
  const int so = p.q0 * p.state_dim;
  dred_encode_latents(ec, st + (size_t)((p.head + latent_offset) % kDredFrames) * p.state_dim, p.stats[4] + so, p.stats[5] + so, p.stats[6] + so, p.stats[7] + so,
                      p.state_dim);
  if (re_tell(ec) > 8 * p.max_bytes) return 0;
  bak = ec;
  const int lim = 2 * p.max_chunks < p.latents_buffer_fill - latent_offset ? 2 * p.max_chunks : p.latents_buffer_fill - latent_offset;
  for (int i = 0; i < lim; i += 2) {
    const int off = dred_quantizer(p.q0, p.dQ, p.qmax, i / 2) * p.latent_dim;
    dred_encode_latents(ec, lat + (size_t)((p.head + i + latent_offset) % kDredFrames) * p.latent_dim, p.stats[0] + off, p.stats[1] + off, p.stats[2] + off,
                        p.stats[3] + off, p.latent_dim);
    if (re_tell(ec) > 8 * p.max_bytes) {
      if (i == 0) return 0;
      break;
    }
    const int active = dred_voice_active(act, i + latent_offset);
    if (active || prev_active) {
      bak = ec;
      encoded = i + 2;
    }
    prev_active = active;
  }
  if (encoded == 0 || (encoded <= 2 && extra)) return 0;
  ec = bak;
  const int fill = (re_tell(ec) + 7) / 8;
  re_shrink(ec, (uint32_t)fill);
  re_done(ec);
  return fill;
}
}  // namespace ob
