// celt_mdct_fast.cuh — the long-block forward MDCT (N = 960: clt_mdct_forward_c with shift 0, celt/mdct.c:122-264, over the
// 480-point opus_fft_impl, celt/kiss_fft.c:52-313) with the FFT in REGISTERS.
//
// The shared-memory FFT of celt_fft.cuh walks all 480 complex values through shared memory once per radix stage (5 stages, lane-strided
// butterflies, an index split per butterfly).  Here the five stages (4.2.4.3.5, innermost first, as the reference's factor table runs
// them) are regrouped by where their operands live:
//   * stages 4(m=1), 2(m=4), 4(m=8) only combine values inside an aligned group of 32 consecutive points: ONE LANE takes a whole group
//     (32 complex = 64 registers), runs the three stages fully unrolled with compile-time twiddle indices, and writes the group back.
//     15 groups per transform: a warp works on TWO transforms at once (lanes 0..14 and 16..30);
//   * stages 3(m=32), 5(m=96) only combine points whose index is equal modulo 32: LANE j takes the 15 points j + 32k of one transform
//     and runs both stages in registers; its twiddles (tw[5j], tw[10j], tw[j + 32a], ...) depend on the lane only;
//   * the post-rotation pairs point i with point 479 - i, which lane 31 - j holds in register 14 - k: one shuffle per point, then a
//     coalesced 8-byte store per lane (outputs 2i, 2i + 1).
// So a transform crosses shared memory twice (the digit-reversed scatter of the pre-rotation, and the regrouping between the two register
// phases) instead of ten times, and no butterfly index is computed at run time.  Every butterfly is evaluated with the reference's
// expression order (compiled with -fmad=false), so the result is bit-identical to clt_mdct_forward_c — and to mdct_forward_from_ws.
//
// Shared-memory layout of one transform's 480 points: group g (32 points) starts at complex index 34 g (15 x 34 = 510 complex, 4080 B):
// the 16-byte loads of the group phase (lane = group) and the 8-byte loads of the column phase (lane = point mod 32) are both free of
// bank conflicts.
#pragma once
#include "celt_fft.cuh"

namespace ob {

constexpr int kF480Pad = 34, kF480Cpx = 15 * kF480Pad;
OB_DEV int f480_pos(int i) { return (i >> 5) * kF480Pad + (i & 31); }

// the constant tables of the transform, staged in shared memory once per CTA (ob_mdct_fast_kernel: by cp.async.bulk)
struct alignas(16) Mdct480Tabs {
  float trig[960];     // ob_mdct_trig[0..960): t0 = trig[i], t1 = trig[480 + i]
  float tw[960];       // ob_fft_tw: 480 complex twiddles
  float window[120];
  short bitrev[480];
};

OB_DEV cpx f480_tw(const Mdct480Tabs *t, int k) { return ((const cpx *)t->tw)[k]; }

// radix-4 butterfly of the first stage (m = 1, no twiddles), celt/kiss_fft.c:98-128
OB_DEV void f480_bfly4_m1(cpx &a0, cpx &a1, cpx &a2, cpx &a3) {
  cpx s0, s1;
  s0.r = a0.r - a2.r; s0.i = a0.i - a2.i;
  a0.r += a2.r; a0.i += a2.i;
  s1.r = a1.r + a3.r; s1.i = a1.i + a3.i;
  a2.r = a0.r - s1.r; a2.i = a0.i - s1.i;
  a0.r += s1.r; a0.i += s1.i;
  s1.r = a1.r - a3.r; s1.i = a1.i - a3.i;
  a1.r = s0.r + s1.i; a1.i = s0.i - s1.r;
  a3.r = s0.r - s1.i; a3.i = s0.i + s1.r;
}
// radix-2 butterfly of the m = 4 stage: fixed twiddles per j (celt/kiss_fft.c:52-96)
template <int J> OB_DEV void f480_bfly2_m4(cpx &a, cpx &b) {
  const float tw = 0.7071067812f;
  cpx t;
  if (J == 0) t = b;
  else if (J == 1) { t.r = (b.r + b.i) * tw; t.i = (b.i - b.r) * tw; }
  else if (J == 2) { t.r = b.i; t.i = -b.r; }
  else { t.r = (b.i - b.r) * tw; t.i = (-(b.i + b.r)) * tw; }
  b.r = a.r - t.r; b.i = a.i - t.i;
  a.r = a.r + t.r; a.i = a.i + t.i;
}
// radix-4 butterfly with twiddles (celt/kiss_fft.c:129-170)
OB_DEV void f480_bfly4(cpx &f0, cpx &f1, cpx &f2, cpx &f3, cpx t1, cpx t2, cpx t3) {
  const cpx s0 = cmul(f1, t1), s1 = cmul(f2, t2), s2 = cmul(f3, t3);
  cpx s5, s3, s4;
  s5.r = f0.r - s1.r; s5.i = f0.i - s1.i;
  f0.r += s1.r; f0.i += s1.i;
  s3.r = s0.r + s2.r; s3.i = s0.i + s2.i;
  s4.r = s0.r - s2.r; s4.i = s0.i - s2.i;
  f2.r = f0.r - s3.r; f2.i = f0.i - s3.i;
  f0.r += s3.r; f0.i += s3.i;
  f1.r = s5.r + s4.i; f1.i = s5.i - s4.r;
  f3.r = s5.r - s4.i; f3.i = s5.i + s4.r;
}
// radix-3 butterfly (celt/kiss_fft.c:174-222)
OB_DEV void f480_bfly3(cpx &f0, cpx &f1, cpx &f2, cpx t1, cpx t2, float epi3_i) {
  const cpx s1 = cmul(f1, t1), s2 = cmul(f2, t2);
  cpx s3, s0, fm;
  s3.r = s1.r + s2.r; s3.i = s1.i + s2.i;
  s0.r = s1.r - s2.r; s0.i = s1.i - s2.i;
  fm.r = f0.r - s3.r * .5f;
  fm.i = f0.i - s3.i * .5f;
  s0.r *= epi3_i; s0.i *= epi3_i;
  f0.r += s3.r; f0.i += s3.i;
  f2.r = fm.r + s0.i; f2.i = fm.i - s0.r;
  f1.r = fm.r - s0.i; f1.i = fm.i + s0.r;
}
// radix-5 butterfly (celt/kiss_fft.c:226-313)
OB_DEV void f480_bfly5(cpx &f0, cpx &f1, cpx &f2, cpx &f3, cpx &f4, cpx t1, cpx t2, cpx t3, cpx t4, cpx ya, cpx yb) {
  const cpx s0 = f0;
  const cpx s1 = cmul(f1, t1), s2 = cmul(f2, t2), s3 = cmul(f3, t3), s4 = cmul(f4, t4);
  cpx s7, s10, s8, s9, s5, s6, s11, s12;
  s7.r = s1.r + s4.r; s7.i = s1.i + s4.i;
  s10.r = s1.r - s4.r; s10.i = s1.i - s4.i;
  s8.r = s2.r + s3.r; s8.i = s2.i + s3.i;
  s9.r = s2.r - s3.r; s9.i = s2.i - s3.i;
  f0.r = s0.r + (s7.r + s8.r); f0.i = s0.i + (s7.i + s8.i);
  s5.r = s0.r + (s7.r * ya.r + s8.r * yb.r);
  s5.i = s0.i + (s7.i * ya.r + s8.i * yb.r);
  s6.r = s10.i * ya.i + s9.i * yb.i;
  s6.i = -(s10.r * ya.i + s9.r * yb.i);
  f1.r = s5.r - s6.r; f1.i = s5.i - s6.i;
  f4.r = s5.r + s6.r; f4.i = s5.i + s6.i;
  s11.r = s0.r + (s7.r * yb.r + s8.r * ya.r);
  s11.i = s0.i + (s7.i * yb.r + s8.i * ya.r);
  s12.r = s9.i * ya.i - s10.i * yb.i;
  s12.i = s10.r * yb.i - s9.r * ya.i;
  f2.r = s11.r + s12.r; f2.i = s11.i + s12.i;
  f3.r = s11.r - s12.r; f3.i = s11.i - s12.i;
}

// window + fold + pre-rotation of one transform (celt/mdct.c:157-226, N2 = 960, N4 = 480, overlap 120): lane j produces points
// j + 32k and scatters them to their digit-reversed place in the padded layout.  in: 1080 samples.
OB_DEV void mdct480_prerotate(const float *__restrict__ in, cpx *__restrict__ F, const Mdct480Tabs *__restrict__ t, int lane) {
  constexpr int N2 = 960, N4 = 480, overlap = 120, q1 = (overlap + 3) >> 2;
  const float scale = f_div(1.f, (float)N4);
#pragma unroll 3
  for (int k = 0; k < 15; k++) {
    const int i = lane + 32 * k;
    const float *xp1 = in + (overlap >> 1) + 2 * i, *xp2 = in + N2 - 1 + (overlap >> 1) - 2 * i;
    float re, im;
    if (i < q1) {
      const float w1 = t->window[(overlap >> 1) + 2 * i], w2 = t->window[(overlap >> 1) - 1 - 2 * i];
      re = xp1[N2] * w2 + (*xp2) * w1;
      im = (*xp1) * w1 - xp2[-N2] * w2;
    } else if (i < N4 - q1) {
      re = *xp2;
      im = *xp1;
    } else {
      const int kk = i - (N4 - q1);
      const float w1 = t->window[2 * kk], w2 = t->window[overlap - 1 - 2 * kk];
      re = -(xp1[-N2] * w1) + (*xp2) * w2;
      im = (*xp1) * w2 + xp2[N2] * w1;
    }
    const float t0 = t->trig[i], t1 = t->trig[N4 + i];
    const float yr = re * t0 - im * t1;
    const float yi = im * t0 + re * t1;
    cpx v;
    v.r = yr * scale;
    v.i = yi * scale;
    F[f480_pos(t->bitrev[i])] = v;
  }
}

// stages 4(m=1), 2(m=4), 4(m=8) of one aligned group of 32 points, in registers.  G: the group's 32 complex values (16-byte aligned).
OB_DEV void f480_group_stages(cpx *__restrict__ G, const Mdct480Tabs *__restrict__ t) {
  cpx v[32];
#pragma unroll
  for (int e = 0; e < 32; e += 2) {
    const float4 q = *(const float4 *)(G + e);
    v[e].r = q.x; v[e].i = q.y; v[e + 1].r = q.z; v[e + 1].i = q.w;
  }
#pragma unroll
  for (int q = 0; q < 8; q++) f480_bfly4_m1(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
#pragma unroll
  for (int h = 0; h < 4; h++) {
    f480_bfly2_m4<0>(v[8 * h + 0], v[8 * h + 4]);
    f480_bfly2_m4<1>(v[8 * h + 1], v[8 * h + 5]);
    f480_bfly2_m4<2>(v[8 * h + 2], v[8 * h + 6]);
    f480_bfly2_m4<3>(v[8 * h + 3], v[8 * h + 7]);
  }
#pragma unroll
  for (int j = 0; j < 8; j++)   // twiddle stride 15 (fstride of this stage): indices are compile-time constants, the loads broadcast
    f480_bfly4(v[j], v[j + 8], v[j + 16], v[j + 24], f480_tw(t, 15 * j), f480_tw(t, 30 * j), f480_tw(t, 45 * j));
#pragma unroll
  for (int e = 0; e < 32; e += 2) {
    float4 q;
    q.x = v[e].r; q.y = v[e].i; q.z = v[e + 1].r; q.w = v[e + 1].i;
    *(float4 *)(G + e) = q;
  }
}

// stages 3(m=32), 5(m=96) and the post-rotation (celt/mdct.c:236-262) of one transform: lane j owns points j + 32k.  out: 960 coefficients
// (8-byte aligned).
OB_DEV void f480_column_stages_postrotate(const cpx *__restrict__ F, const Mdct480Tabs *__restrict__ t, float *__restrict__ out, int lane) {
  cpx u[15];
#pragma unroll
  for (int k = 0; k < 15; k++) u[k] = F[k * kF480Pad + lane];
  {
    const cpx t1 = f480_tw(t, 5 * lane), t2 = f480_tw(t, 10 * lane);
    const float epi3_i = f480_tw(t, 160).i;
#pragma unroll
    for (int g = 0; g < 5; g++) f480_bfly3(u[3 * g], u[3 * g + 1], u[3 * g + 2], t1, t2, epi3_i);
  }
  {
    const cpx ya = f480_tw(t, 96), yb = f480_tw(t, 192);
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const int jp = lane + 32 * a;
      f480_bfly5(u[a], u[3 + a], u[6 + a], u[9 + a], u[12 + a], f480_tw(t, jp), f480_tw(t, 2 * jp), f480_tw(t, 3 * jp), f480_tw(t, 4 * jp), ya, yb);
    }
  }
  // out[2i] = fp.i * t1 - fp.r * t0 from point i;  out[2i + 1] = out[N2 - 1 - 2i'] = fp.r * t1 + fp.i * t0 from point i' = 479 - i,
  // which is register 14 - k of lane 31 - j
  float A[15], B[15];
#pragma unroll
  for (int k = 0; k < 15; k++) {
    const int i = lane + 32 * k;
    const float t0 = t->trig[i], t1 = t->trig[480 + i];
    A[k] = u[k].i * t1 - u[k].r * t0;
    B[k] = u[k].r * t1 + u[k].i * t0;
  }
#pragma unroll
  for (int k = 0; k < 15; k++) {
    const float b = shfl(B[14 - k], 31 - lane);
    float2 o;
    o.x = A[k]; o.y = b;
    *(float2 *)(out + 2 * (lane + 32 * k)) = o;
  }
}

// Two transforms by one warp.  in0 / in1: 1080 samples each (shared memory); F0 / F1: kF480Cpx complex each; out0 / out1: 960 coefficients
// (in1 == nullptr: one transform only).
OB_DEV void mdct480_forward_pair(const float *in0, const float *in1, cpx *F0, cpx *F1, const Mdct480Tabs *t, float *out0, float *out1, int lane) {
  mdct480_prerotate(in0, F0, t, lane);
  if (in1) mdct480_prerotate(in1, F1, t, lane);
  wsync();
  {
    const int g = lane & 15, which = lane >> 4;
    if (g < 15 && (which == 0 || in1)) f480_group_stages((which ? F1 : F0) + g * kF480Pad, t);
  }
  wsync();
  f480_column_stages_postrotate(F0, t, out0, lane);
  if (in1) f480_column_stages_postrotate(F1, t, out1, lane);
}

}  // namespace ob

#if defined(__CUDACC__)
namespace ob {
// ---- ob_mdct_forward_tma: persistent kernel, tables and inputs staged with cp.async.bulk (the bulk-copy engine of the TMA unit) ----
// One CTA of kMdctFastWarps warps per SM slot; the four constant tables (9 KB) are copied into shared memory once per CTA, every warp
// then loops over PAIRS of transforms: its two 4320-byte inputs arrive by bulk copy on the warp's own mbarrier, and the copy of the next
// pair is issued as soon as the pre-rotation has consumed the current one, so it lands under the FFT.  Outputs go straight from
// registers to HBM as coalesced 8-byte stores.  HBM traffic = the algorithmic 8160 B per transform (4320 in + 3840 out).
namespace tma {
__device__ __forceinline__ uint32_t saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {   // bounded: a protocol error traps instead of hanging the GPU
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; spin++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
}  // namespace tma

constexpr int kMdctFastWarps = 6;
struct alignas(16) MdctFastWarpWS { float in[2][1080]; cpx F[2][kF480Cpx]; };
struct alignas(16) MdctFastSmem { Mdct480Tabs tabs; MdctFastWarpWS w[kMdctFastWarps]; unsigned long long bar[kMdctFastWarps + 1]; };

OB_DEV void mdct_forward_tma_block(const MdctFwdParams &p, unsigned char *smem) {
  MdctFastSmem *S = (MdctFastSmem *)smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tab_bar = tma::saddr(&S->bar[kMdctFastWarps]), my_bar = tma::saddr(&S->bar[warp]);
  if (threadIdx.x == 0) {
    for (int i = 0; i <= kMdctFastWarps; i++) tma::mbar_init(tma::saddr(&S->bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tma::mbar_expect_tx(tab_bar, sizeof(Mdct480Tabs));
    tma::bulk_g2s(tma::saddr(S->tabs.trig), ob_mdct_trig, sizeof(S->tabs.trig), tab_bar);
    tma::bulk_g2s(tma::saddr(S->tabs.tw), ob_fft_tw, sizeof(S->tabs.tw), tab_bar);
    tma::bulk_g2s(tma::saddr(S->tabs.window), ob_window, sizeof(S->tabs.window), tab_bar);
    tma::bulk_g2s(tma::saddr(S->tabs.bitrev), ob_fft_bitrev480, sizeof(S->tabs.bitrev), tab_bar);
  }
  MdctFastWarpWS *w = &S->w[warp];
  const int npairs = (p.n_items + 1) >> 1, stride = gridDim.x * kMdctFastWarps;
  int pair = blockIdx.x * kMdctFastWarps + warp;
  auto issue = [&](int pr) {   // lane 0: the two inputs of pair pr
    const int two = 2 * pr + 1 < p.n_items;
    tma::mbar_expect_tx(my_bar, (two ? 2u : 1u) * 4320u);
    tma::bulk_g2s(tma::saddr(w->in[0]), p.in + (size_t)(2 * pr) * p.in_stride, 4320u, my_bar);
    if (two) tma::bulk_g2s(tma::saddr(w->in[1]), p.in + (size_t)(2 * pr + 1) * p.in_stride, 4320u, my_bar);
  };
  if (lane == 0 && pair < npairs) issue(pair);
  tma::mbar_wait(tab_bar, 0);
  uint32_t phase = 0;
  for (; pair < npairs; pair += stride) {
    const bool two = 2 * pair + 1 < p.n_items;
    tma::mbar_wait(my_bar, phase);
    phase ^= 1;
    mdct480_prerotate(w->in[0], w->F[0], &S->tabs, lane);
    if (two) mdct480_prerotate(w->in[1], w->F[1], &S->tabs, lane);
    __syncwarp();
    if (lane == 0 && pair + stride < npairs) {   // the inputs are consumed: the next pair's copy runs under the FFT
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(pair + stride);
    }
    {
      const int g = lane & 15, which = lane >> 4;
      if (g < 15 && (which == 0 || two)) f480_group_stages(w->F[which] + g * kF480Pad, &S->tabs);
    }
    __syncwarp();
    f480_column_stages_postrotate(w->F[0], &S->tabs, p.out + (size_t)(2 * pair) * p.out_stride, lane);
    if (two) f480_column_stages_postrotate(w->F[1], &S->tabs, p.out + (size_t)(2 * pair + 1) * p.out_stride, lane);
    __syncwarp();
  }
}
}  // namespace ob
#endif
