// celt_vec.cuh — warp-cooperative vector primitives of the CELT band decoder.
// One warp owns one stream; vectors (N <= 176) live in that warp's slice of shared memory and
// are processed lane-strided.  Every routine starts and ends with the warp converged and the
// vector visible to all lanes (they __syncwarp() on entry where they read what another phase
// wrote, and on exit after writing).
//
// Float results follow the reference's formulas; where the reference sums serially (dot products)
// the warp uses a tree instead, so values can differ from the scalar reference in the last bits
// (decode tolerance: +-1 LSB of 16-bit PCM).
#pragma once
#include "simt.cuh"

namespace ob {

OB_DEV float rsqrt_ref(float x) { return f_div(1.f, f_sqrt(x)); }  // 1.f/(float)sqrt(x), celt/mathops.h:227-228

OB_DEV float warp_dot(const float *x, const float *y, int n, int lane) {
  float s = 0.f;
#pragma unroll 1
  for (int i = lane; i < n; i += 32) s += x[i] * y[i];
  return warp_sum(s);
}

// X *= rsqrt(eps + |X|^2) * gain   (celt/vq.c:695-722)
OB_DEVNI void renormalise(float *X, int N, float gain, int lane) {
  wsync();
  float E = 1e-15f + warp_dot(X, X, N, lane);
  float g = rsqrt_ref(E) * gain;
#pragma unroll 1
  for (int i = lane; i < N; i += 32) X[i] = g * X[i];
  wsync();
}

// One Haar butterfly level (celt/bands.c:623-636): pairs (stride*2j+i, stride*(2j+1)+i)
OB_DEVNI void haar_step(float *X, int N0, int stride, int lane) {
  wsync();
  int half = N0 >> 1;
  int npairs = half * stride;
#pragma unroll 1
  for (int p = lane; p < npairs; p += 32) {
    int i = p % stride, j = p / stride;
    int a = stride * 2 * j + i, b = stride * (2 * j + 1) + i;
    float t1 = .70710678f * X[a], t2 = .70710678f * X[b];
    X[a] = t1 + t2;
    X[b] = t1 - t2;
  }
  wsync();
}

// natural -> sequency order with DC last, rows for strides 2,4,8,16 (celt/bands.c:564-572)
OB_TABLE unsigned char ob_hadamard_order[30] = {1, 0, 3, 0, 2, 1, 7, 0, 4, 3, 6, 1, 5, 2,
                                               15, 0, 8, 7, 12, 3, 11, 4, 14, 1, 9, 6, 13, 2, 10, 5};

// celt/bands.c:574-597.  tmp: scratch of N floats.
OB_DEVNI void deinterleave_hadamard(float *X, float *tmp, int N0, int stride, int hadamard, int lane) {
  wsync();
  int N = N0 * stride;
#pragma unroll 1
  for (int p = lane; p < N; p += 32) {
    int j = p / stride, i = p - j * stride;  // source index p = j*stride+i
    int row = hadamard ? ob_hadamard_order[stride - 2 + i] : i;
    tmp[row * N0 + j] = X[p];
  }
  wsync();
#pragma unroll 1
  for (int p = lane; p < N; p += 32) X[p] = tmp[p];
  wsync();
}
// celt/bands.c:599-621
OB_DEVNI void interleave_hadamard(float *X, float *tmp, int N0, int stride, int hadamard, int lane) {
  wsync();
  int N = N0 * stride;
#pragma unroll 1
  for (int p = lane; p < N; p += 32) {
    int j = p / stride, i = p - j * stride;  // destination index p = j*stride+i
    int row = hadamard ? ob_hadamard_order[stride - 2 + i] : i;
    tmp[p] = X[row * N0 + j];
  }
  wsync();
#pragma unroll 1
  for (int p = lane; p < N; p += 32) X[p] = tmp[p];
  wsync();
}

// First-order linear recurrences y' = a*y + b are evaluated across lanes by composing affine maps
// with a (segmented) warp scan; used by the synthesis kernel's de-emphasis.
struct Affine { float a, b; };  // y -> a*y + b

// inclusive scan of affine maps inside aligned groups of P lanes (P power of two)
OB_DEV Affine group_scan(Affine m, int P, int lane) {
  int gl = lane & (P - 1);
#pragma unroll 1
  for (int d = 1; d < P; d <<= 1) {
    float pa = shfl_up(m.a, d), pb = shfl_up(m.b, d);
    if (gl >= d) {  // compose: first the left map (pa,pb), then mine
      m.b = m.a * pb + m.b;
      m.a = m.a * pa;
    }
  }
  return m;
}

OB_DEVNI float cos_norm_ref(float x) {  // (float)cos((.5f*PI)*(x)) with PI a double, celt/mathops.h:229
  return (float)cos((.5f * 3.1415926535897931) * (double)x);
}

}  // namespace ob
