// celt_vec.cuh — warp-cooperative vector primitives of the CELT band decoder.
// One warp owns one stream; vectors (N <= 176) live in that warp's slice of shared memory and
// are processed lane-strided.  Every routine starts and ends with the warp converged and the
// vector visible to all lanes (they __syncwarp() on entry where they read what another phase
// wrote, and on exit after writing).
//
// Float results follow the reference's formulas; where the reference runs a serial chain
// (spreading rotation, dot products) the warp uses a scan / tree instead, so values can differ
// from the scalar reference in the last bits (decode tolerance: +-1 LSB of 16-bit PCM).
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

// -----------------------------------------------------------------------------------------
// Spreading rotation (celt/vq.c:75-147).  A pass with stride s over a vector of length len is s
// independent chains; along a chain (elements e_0, e_1, ... at distance s) the forward sweep is
//    a_0 = e_0 ;  a_{t+1} = c*e_{t+1} + sn*a_t ;  out_t = c*a_t - sn*e_{t+1} ;  out_last = a_last
// and the backward sweep (t descending from L-2) is
//    d_{L-2} = e_{L-1} ; final_{t+1} = c*d_t + sn*e_t ; d_{t-1} = c*e_t - sn*d_t ; final_0 = d_{-1}
// Both are first-order linear recurrences y' = alpha*y + beta_t, evaluated here by giving each
// lane a contiguous segment of a chain and composing the segments' affine maps with a
// segmented warp scan (P lanes per chain).
// -----------------------------------------------------------------------------------------
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

// One sweep pair (forward then backward) with stride s over `nvec` vectors of length len stored
// back to back at X (vector v at X + v*len).  Chains = nvec*s.  T is scratch of nvec*len floats:
// the forward sweep goes X -> T, the backward sweep T -> X (out of place: no read/write hazards
// between lanes working on neighbouring segments).
OB_DEVNI void rotate_pass(float *X, float *T, int len, int nvec, int s, float c, float sn, int lane) {
  const int nch = nvec * s;
  int lgP = 0;  // lanes per chain P = 2^lgP: largest power of two with P*nch <= 32
  while ((2 << lgP) * nch <= 32) lgP++;
  const int P = 1 << lgP;
  const int G = 32 >> lgP;  // chains in flight
  const int gl = lane & (P - 1), g = lane >> lgP;
  // x / s for 0 <= x < 5041 as a multiply-shift (s <= 13): exact, and keeps integer division
  // (a ~30-instruction sequence each) out of the loops
  const unsigned inv_s = 65535u / (unsigned)s + 1u;
  wsync();
  // ---------------- forward sweep X -> T : steps t = 0 .. L-2 ----------------
#pragma unroll 1
  for (int q0 = 0; q0 < nch; q0 += G) {
    const int q = q0 + g;
    const bool live = q < nch;
    const int v = live ? (int)(((unsigned)q * inv_s) >> 16) : 0, r = live ? q - v * s : 0;
    const float *src = X + v * len + r;
    float *dst = T + v * len + r;
    const int L = live ? (int)(((unsigned)(len - r + s - 1) * inv_s) >> 16) : 0;
    const int steps = L > 0 ? L - 1 : 0;
    const int seg = (steps + P - 1) >> lgP;
    const int t0 = gl * seg < steps ? gl * seg : steps;
    const int t1 = t0 + seg < steps ? t0 + seg : steps;
    Affine m = {1.f, 0.f};
#pragma unroll 1
    for (int t = t0; t < t1; t++) {  // a_{t+1} = sn*a_t + c*e_{t+1}
      float e1 = src[(t + 1) * s];
      m.b = sn * m.b + c * e1;
      m.a = sn * m.a;
    }
    Affine inc = group_scan(m, P, lane);
    float pa = shfl_up(inc.a, 1), pb = shfl_up(inc.b, 1);
    float e0 = L > 0 ? src[0] : 0.f;
    float a = gl == 0 ? e0 : pa * e0 + pb;  // a_{t0}
#pragma unroll 1
    for (int t = t0; t < t1; t++) {
      float e1 = src[(t + 1) * s];
      dst[t * s] = c * a - sn * e1;
      a = c * e1 + sn * a;
    }
    // a_last (or the untouched single element) is stored by the lane whose segment ends the chain
    if (L > 0 && t1 == steps && (t0 < t1 || (steps == 0 && gl == 0))) dst[steps * s] = a;
  }
  wsync();
  // ---------------- backward sweep T -> X : steps t = L-3 .. 0 ----------------
#pragma unroll 1
  for (int q0 = 0; q0 < nch; q0 += G) {
    const int q = q0 + g;
    const bool live = q < nch;
    const int v = live ? (int)(((unsigned)q * inv_s) >> 16) : 0, r = live ? q - v * s : 0;
    const float *src = T + v * len + r;
    float *dst = X + v * len + r;
    const int L = live ? (int)(((unsigned)(len - r + s - 1) * inv_s) >> 16) : 0;
    const int steps = L > 2 ? L - 2 : 0;  // step u handles t = L-3-u
    const int seg = (steps + P - 1) >> lgP;
    const int t0 = gl * seg < steps ? gl * seg : steps;
    const int t1 = t0 + seg < steps ? t0 + seg : steps;
    Affine m = {1.f, 0.f};
#pragma unroll 1
    for (int u = t0; u < t1; u++) {  // d <- c*f_t - sn*d
      float ft = src[(L - 3 - u) * s];
      m.b = c * ft - sn * m.b;
      m.a = -sn * m.a;
    }
    Affine inc = group_scan(m, P, lane);
    float pa = shfl_up(inc.a, 1), pb = shfl_up(inc.b, 1);
    float d0 = L >= 2 ? src[(L - 2) * s] : 0.f;  // x2 of the first backward step
    float d = gl == 0 ? d0 : pa * d0 + pb;
#pragma unroll 1
    for (int u = t0; u < t1; u++) {
      int t = L - 3 - u;
      float ft = src[t * s];
      dst[(t + 1) * s] = c * d + sn * ft;
      d = c * ft - sn * d;
    }
    if (steps > 0) {
      if (t1 == steps && t0 < t1) dst[0] = d;           // final value of the chain head
      if (gl == 0) dst[(L - 1) * s] = src[(L - 1) * s];  // tail element is not touched by this sweep
    } else if (gl == 0) {
      for (int t = 0; t < L; t++) dst[t * s] = src[t * s];  // chains of length <= 2: copy through
    }
  }
  wsync();
}

OB_DEVNI float cos_norm_ref(float x) {  // (float)cos((.5f*PI)*(x)) with PI a double, celt/mathops.h:229
  return (float)cos((.5f * 3.1415926535897931) * (double)x);
}

// exp_rotation(X, len, dir=-1, stride, K, spread): the decoder only ever undoes the rotation.
OB_DEVNI void spread_unrotate(float *X, float *T, int len, int stride, int K, int spread, int lane) {
  if (2 * K >= len || spread == 0) return;
  int factor = spread == 1 ? 15 : spread == 2 ? 10 : 5;
  float gain = f_div((float)len, (float)(len + factor * K));
  float theta = .5f * (gain * gain);
  float c = cos_norm_ref(theta);
  float s = cos_norm_ref(1.0f - theta);
  int stride2 = 0;
  if (len >= 8 * stride) {
    stride2 = 1;
    while ((stride2 * stride2 + stride2) * stride + (stride >> 2) < len) stride2++;
  }
  int sub = (unsigned)len / (unsigned)stride;
  // reference: per sub-vector { if (stride2) pass(stride2, s, c); pass(1, c, s); } -- the
  // sub-vectors are independent, so all stride2 passes can run before all stride-1 passes.
  if (stride2) rotate_pass(X, T, sub, stride, stride2, s, c, lane);
  rotate_pass(X, T, sub, stride, 1, c, s, lane);
}

}  // namespace ob
