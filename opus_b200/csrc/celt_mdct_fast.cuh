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
template <int PAD = kF480Pad> OB_DEV int f480_pos(int i) { return (i >> 5) * PAD + (i & 31); }
// PAD = 34 (16-byte group loads) everywhere except the encoder's analysis kernel, which has 3960 bytes to spare per warp, not 4080: PAD = 33 there
// (8-byte group loads; the lane-per-group and lane-per-column access patterns stay conflict-free for any odd or even stride that is not a multiple of 16 words)
constexpr int kF480PadSmall = 33;

// the constant tables of the transform, staged in shared memory once per CTA (ob_mdct_fast_kernel: by cp.async.bulk)
struct alignas(16) Mdct480Tabs {
  float trig[960];     // ob_mdct_trig[0..960): t0 = trig[i], t1 = trig[480 + i]
  float tw[960];       // ob_fft_tw: 480 complex twiddles
  float window[120];
  short bitrev[480];
  OB_MDEV cpx get_tw(int k) const { return ((const cpx *)tw)[k]; }
  OB_MDEV float get_trig(int i) const { return trig[i]; }
  OB_MDEV float get_window(int i) const { return window[i]; }
  OB_MDEV int get_bitrev(int i) const { return bitrev[i]; }
};
// the same tables where the library keeps them (global memory, served by L1 / L2): for the kernels that run one transform per warp inside a
// larger job and have no table stage of their own (the decoder's synthesis kernel)
struct Mdct480Global {
  OB_MDEV cpx get_tw(int k) const { return ldtw(k); }
  OB_MDEV float get_trig(int i) const { return ob_mdct_trig[i]; }
  OB_MDEV float get_window(int i) const { return ob_window[i]; }
  OB_MDEV int get_bitrev(int i) const { return ob_fft_bitrev480[i]; }
};
template <class T> OB_DEV cpx f480_tw(const T *t, int k) { return t->get_tw(k); }

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
template <int PAD = kF480Pad, class T> OB_DEV void mdct480_prerotate(const float *__restrict__ in, cpx *__restrict__ F, const T *__restrict__ t, int lane) {
  constexpr int N2 = 960, N4 = 480, overlap = 120, q1 = (overlap + 3) >> 2;
  const float scale = f_div(1.f, (float)N4);
#pragma unroll 3
  for (int k = 0; k < 15; k++) {
    const int i = lane + 32 * k;
    const float *xp1 = in + (overlap >> 1) + 2 * i, *xp2 = in + N2 - 1 + (overlap >> 1) - 2 * i;
    float re, im;
    if (i < q1) {
      const float w1 = t->get_window((overlap >> 1) + 2 * i), w2 = t->get_window((overlap >> 1) - 1 - 2 * i);
      re = xp1[N2] * w2 + (*xp2) * w1;
      im = (*xp1) * w1 - xp2[-N2] * w2;
    } else if (i < N4 - q1) {
      re = *xp2;
      im = *xp1;
    } else {
      const int kk = i - (N4 - q1);
      const float w1 = t->get_window(2 * kk), w2 = t->get_window(overlap - 1 - 2 * kk);
      re = -(xp1[-N2] * w1) + (*xp2) * w2;
      im = (*xp1) * w2 + xp2[N2] * w1;
    }
    const float t0 = t->get_trig(i), t1 = t->get_trig(N4 + i);
    const float yr = re * t0 - im * t1;
    const float yi = im * t0 + re * t1;
    cpx v;
    v.r = yr * scale;
    v.i = yi * scale;
    F[f480_pos<PAD>(t->get_bitrev(i))] = v;
  }
}

// stages 4(m=1), 2(m=4), 4(m=8) of one aligned group of 32 points, in registers.  G: the group's 32 complex values (16-byte aligned).
template <bool WIDE = true, class T> OB_DEV void f480_group_stages(cpx *__restrict__ G, const T *__restrict__ t) {
  cpx v[32];
  if (WIDE) {
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
      const float4 q = *(const float4 *)(G + e);
      v[e].r = q.x; v[e].i = q.y; v[e + 1].r = q.z; v[e + 1].i = q.w;
    }
  } else {
#pragma unroll
    for (int e = 0; e < 32; e++) v[e] = G[e];
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
  if (WIDE) {
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
      float4 q;
      q.x = v[e].r; q.y = v[e].i; q.z = v[e + 1].r; q.w = v[e + 1].i;
      *(float4 *)(G + e) = q;
    }
  } else {
#pragma unroll
    for (int e = 0; e < 32; e++) G[e] = v[e];
  }
}

// stages 3(m=32), 5(m=96) and the post-rotation (celt/mdct.c:236-262) of one transform: lane j owns points j + 32k.  out: 960 coefficients
// (8-byte aligned).
template <class T> OB_DEV void f480_column_stages(cpx (&u)[15], const T *__restrict__ t, int lane) {
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
}
template <int PAD = kF480Pad, class T> OB_DEV void f480_column_stages_postrotate(const cpx *__restrict__ F, const T *__restrict__ t, float *__restrict__ out, int lane) {
  cpx u[15];
#pragma unroll
  for (int k = 0; k < 15; k++) u[k] = F[k * PAD + lane];
  f480_column_stages(u, t, lane);
  // out[2i] = fp.i * t1 - fp.r * t0 from point i;  out[2i + 1] = out[N2 - 1 - 2i'] = fp.r * t1 + fp.i * t0 from point i' = 479 - i,
  // which is register 14 - k of lane 31 - j
  float A[15], B[15];
#pragma unroll
  for (int k = 0; k < 15; k++) {
    const int i = lane + 32 * k;
    const float t0 = t->get_trig(i), t1 = t->get_trig(480 + i);
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
template <class T> OB_DEV void mdct480_forward_pair(const float *in0, const float *in1, cpx *F0, cpx *F1, const T *t, float *out0, float *out1, int lane) {
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

// ---- the long-block INVERSE transform (clt_mdct_backward_c with shift 0, celt/mdct.c:268-369: pre-rotation into digit-reversed order, the same
// 480-point FFT, post-rotation from both ends) for one warp.  freq: the 960 coefficients; buf: 1020 floats (16-byte aligned), used first as the
// padded point array and overwritten in place with the 960 time-domain values in natural order (the caller's overlap-add area: the TDAC window
// mirror of mdct.c:372-388 stays with the caller).  Per point p the post-rotation writes out[2p] = re*t0 + im*t1 of point p and
// out[2p + 1] = re*t1 - im*t0 of point 479 - p (re / im = the FFT output's imaginary / real part), the pairing of the forward transform. ----
template <class T> OB_DEV void imdct480_warp(const float *__restrict__ freq, float *__restrict__ buf, const T *__restrict__ t, int lane) {
  constexpr int N2 = 960, N4 = 480;
  cpx *F = (cpx *)buf;
#pragma unroll 3
  for (int k = 0; k < 15; k++) {
    const int i = lane + 32 * k;
    const float x1 = freq[2 * i], x2 = freq[N2 - 1 - 2 * i];
    const float t0 = t->get_trig(i), t1 = t->get_trig(N4 + i);
    const float yr = x2 * t0 + x1 * t1;
    const float yi = x1 * t0 - x2 * t1;
    cpx v;
    v.r = yi;
    v.i = yr;
    F[f480_pos(t->get_bitrev(i))] = v;
  }
  wsync();
  if (lane < 15) f480_group_stages(F + lane * kF480Pad, t);
  wsync();
  cpx u[15];
#pragma unroll
  for (int k = 0; k < 15; k++) u[k] = F[k * kF480Pad + lane];
  wsync();   // every point is in a register: the buffer can take the natural-order output
  f480_column_stages(u, t, lane);
  float A[15], B[15];
#pragma unroll
  for (int k = 0; k < 15; k++) {
    const int i = lane + 32 * k;
    const float re = u[k].i, im = u[k].r;
    const float t0 = t->get_trig(i), t1 = t->get_trig(N4 + i);
    A[k] = re * t0 + im * t1;
    B[k] = re * t1 - im * t0;
  }
#pragma unroll
  for (int k = 0; k < 15; k++) {
    const float b = shfl(B[14 - k], 31 - lane);
    float2 o;
    o.x = A[k]; o.y = b;
    *(float2 *)(buf + 2 * (lane + 32 * k)) = o;
  }
  wsync();
}

}  // namespace ob

