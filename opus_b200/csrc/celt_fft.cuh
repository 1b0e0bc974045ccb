// celt_fft.cuh — the reference's mixed-radix FFT (celt/kiss_fft.c:52-313,562-613) as a
// warp-cooperative routine over shared memory: stage order, twiddle indexing and per-butterfly
// arithmetic are the reference's (so with -fmad=false the results are bit-identical to
// opus_fft_impl), the independent butterflies of a stage are lane-strided.
#pragma once
#include "simt.cuh"
#include "celt_tables_data.h"

namespace ob {

OB_DEV const short *bitrev_table(int shift) {
  return shift == 0 ? ob_fft_bitrev480 : shift == 1 ? ob_fft_bitrev240 : shift == 2 ? ob_fft_bitrev120 : ob_fft_bitrev60;
}

struct alignas(8) cpx { float r, i; };   // 8-byte aligned: one LDS.64 / STS.64 per complex value
OB_DEV cpx cmul(cpx a, cpx b) { cpx m; m.r = a.r * b.r - a.i * b.i; m.i = a.r * b.i + a.i * b.r; return m; }
OB_DEV cpx ldtw(int k) { cpx t; t.r = ob_fft_tw[2 * k]; t.i = ob_fft_tw[2 * k + 1]; return t; }

// One stage of the transform for all blocks: radix P butterflies of span M (M2 = P * M), twiddle stride FS = fstride << shift.  Every constant
// is a template parameter, so the butterfly index split (q -> block, i, j) costs no integer division.  Per-butterfly arithmetic as in
// celt/kiss_fft.c:52-313.
template <int P, int M, int M2, int FS, int NFFT> OB_DEV void fft_stage(cpx *F, int nblk, int lane) {
  constexpr int p = P, m = M, m2 = M2, fs = FS, nfft = NFFT, per_blk = NFFT / P;
  wsync();
  for (int q = lane; q < per_blk * nblk; q += 32) {
      const int blk = q / per_blk, w = q - blk * per_blk;
      const int i = w / m, j = w - i * m;
      cpx *Fo = F + blk * nfft + i * m2 + j;
      if constexpr (p == 4) {
        if constexpr (m == 1) {
          cpx a0 = Fo[0], a1 = Fo[1], a2 = Fo[2], a3 = Fo[3], s0, s1;
          s0.r = a0.r - a2.r; s0.i = a0.i - a2.i;
          a0.r += a2.r; a0.i += a2.i;
          s1.r = a1.r + a3.r; s1.i = a1.i + a3.i;
          a2.r = a0.r - s1.r; a2.i = a0.i - s1.i;
          a0.r += s1.r; a0.i += s1.i;
          s1.r = a1.r - a3.r; s1.i = a1.i - a3.i;
          a1.r = s0.r + s1.i; a1.i = s0.i - s1.r;
          a3.r = s0.r - s1.i; a3.i = s0.i + s1.r;
          Fo[0] = a0; Fo[1] = a1; Fo[2] = a2; Fo[3] = a3;
        } else {
          cpx f0 = Fo[0];
          cpx s0 = cmul(Fo[m], ldtw(j * fs)), s1 = cmul(Fo[2 * m], ldtw(2 * j * fs)), s2 = cmul(Fo[3 * m], ldtw(3 * j * fs));
          cpx s5, s3, s4;
          s5.r = f0.r - s1.r; s5.i = f0.i - s1.i;
          f0.r += s1.r; f0.i += s1.i;
          s3.r = s0.r + s2.r; s3.i = s0.i + s2.i;
          s4.r = s0.r - s2.r; s4.i = s0.i - s2.i;
          cpx o2; o2.r = f0.r - s3.r; o2.i = f0.i - s3.i;
          f0.r += s3.r; f0.i += s3.i;
          cpx o1, o3;
          o1.r = s5.r + s4.i; o1.i = s5.i - s4.r;
          o3.r = s5.r - s4.i; o3.i = s5.i + s4.r;
          Fo[0] = f0; Fo[m] = o1; Fo[2 * m] = o2; Fo[3 * m] = o3;
        }
      } else if constexpr (p == 2) {  // m == 4 : Fo[0] pairs with Fo[4], fixed twiddles per j
        const float tw = 0.7071067812f;
        cpx *F2 = Fo + 4;
        cpx t, a = Fo[0], b = F2[0];
        if (j == 0) t = b;
        else if (j == 1) { t.r = (b.r + b.i) * tw; t.i = (b.i - b.r) * tw; }
        else if (j == 2) { t.r = b.i; t.i = -b.r; }
        else { t.r = (b.i - b.r) * tw; t.i = (-(b.i + b.r)) * tw; }
        F2[0].r = a.r - t.r; F2[0].i = a.i - t.i;
        Fo[0].r = a.r + t.r; Fo[0].i = a.i + t.i;
      } else if constexpr (p == 3) {
        cpx epi3 = ldtw(fs * m);
        cpx f0 = Fo[0];
        cpx s1 = cmul(Fo[m], ldtw(j * fs)), s2 = cmul(Fo[2 * m], ldtw(2 * j * fs)), s3, s0;
        s3.r = s1.r + s2.r; s3.i = s1.i + s2.i;
        s0.r = s1.r - s2.r; s0.i = s1.i - s2.i;
        cpx fm;
        fm.r = f0.r - s3.r * .5f;
        fm.i = f0.i - s3.i * .5f;
        s0.r *= epi3.i; s0.i *= epi3.i;
        f0.r += s3.r; f0.i += s3.i;
        cpx o2, o1;
        o2.r = fm.r + s0.i; o2.i = fm.i - s0.r;
        o1.r = fm.r - s0.i; o1.i = fm.i + s0.r;
        Fo[0] = f0; Fo[m] = o1; Fo[2 * m] = o2;
      } else {  // p == 5
        cpx ya = ldtw(fs * m), yb = ldtw(fs * 2 * m);
        cpx s0 = Fo[0];
        cpx s1 = cmul(Fo[m], ldtw(j * fs)), s2 = cmul(Fo[2 * m], ldtw(2 * j * fs));
        cpx s3 = cmul(Fo[3 * m], ldtw(3 * j * fs)), s4 = cmul(Fo[4 * m], ldtw(4 * j * fs));
        cpx s7, s10, s8, s9, s5, s6, s11, s12, o;
        s7.r = s1.r + s4.r; s7.i = s1.i + s4.i;
        s10.r = s1.r - s4.r; s10.i = s1.i - s4.i;
        s8.r = s2.r + s3.r; s8.i = s2.i + s3.i;
        s9.r = s2.r - s3.r; s9.i = s2.i - s3.i;
        o.r = s0.r + (s7.r + s8.r); o.i = s0.i + (s7.i + s8.i);
        Fo[0] = o;
        s5.r = s0.r + (s7.r * ya.r + s8.r * yb.r);
        s5.i = s0.i + (s7.i * ya.r + s8.i * yb.r);
        s6.r = s10.i * ya.i + s9.i * yb.i;
        s6.i = -(s10.r * ya.i + s9.r * yb.i);
        o.r = s5.r - s6.r; o.i = s5.i - s6.i; Fo[m] = o;
        o.r = s5.r + s6.r; o.i = s5.i + s6.i; Fo[4 * m] = o;
        s11.r = s0.r + (s7.r * yb.r + s8.r * ya.r);
        s11.i = s0.i + (s7.i * yb.r + s8.i * ya.r);
        s12.r = s9.i * ya.i - s10.i * yb.i;
        s12.i = s10.r * yb.i - s9.r * ya.i;
        o.r = s11.r + s12.r; o.i = s11.i + s12.i; Fo[2 * m] = o;
        o.r = s11.r - s12.r; o.i = s11.i - s12.i; Fo[3 * m] = o;
      }
    }
}

// All B blocks' FFTs of size nfft = 480>>shift, in place on digit-reversed data, stage order and per-butterfly arithmetic of
// celt/kiss_fft.c:52-313,562-613 (factors of celt/static_modes_float.h: 480 = 4.2.4.3.5, 240 = 4.4.3.5, 120 = 4.2.3.5, 60 = 4.3.5, executed
// from the innermost factor out).  F: B consecutive arrays of nfft cpx.
OB_DEV void fft_stages(cpx *F, int shift, int nblk, int lane) {
  if (shift == 0) {
    fft_stage<4, 1, 4, 120, 480>(F, nblk, lane);
    fft_stage<2, 4, 8, 60, 480>(F, nblk, lane);
    fft_stage<4, 8, 32, 15, 480>(F, nblk, lane);
    fft_stage<3, 32, 96, 5, 480>(F, nblk, lane);
    fft_stage<5, 96, 1, 1, 480>(F, nblk, lane);
  } else if (shift == 1) {
    fft_stage<4, 1, 4, 120, 240>(F, nblk, lane);
    fft_stage<4, 4, 16, 30, 240>(F, nblk, lane);
    fft_stage<3, 16, 48, 10, 240>(F, nblk, lane);
    fft_stage<5, 48, 1, 2, 240>(F, nblk, lane);
  } else if (shift == 2) {
    fft_stage<4, 1, 4, 120, 120>(F, nblk, lane);
    fft_stage<2, 4, 8, 60, 120>(F, nblk, lane);
    fft_stage<3, 8, 24, 20, 120>(F, nblk, lane);
    fft_stage<5, 24, 1, 4, 120>(F, nblk, lane);
  } else {
    fft_stage<4, 1, 4, 120, 60>(F, nblk, lane);
    fft_stage<3, 4, 12, 40, 60>(F, nblk, lane);
    fft_stage<5, 12, 1, 8, 60>(F, nblk, lane);
  }
  wsync();
}

}  // namespace ob
