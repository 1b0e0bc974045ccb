// simt.cuh — the small set of warp-level primitives every kernel in this library is written
// against.  Under nvcc they are the sm_100a intrinsics.  Under a plain host compiler with
// -DOB_HOST_EMU they map onto tests/emu/emu_runtime.h, a fiber-based 32-lane warp emulator
// that exists ONLY so the kernel sources can be debugged and unit-tested in a container
// without a GPU.  The emulator is never compiled into libopus_b200.so and is not a fallback:
// the product library has no host execution path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)

#define OB_DEV __device__ __forceinline__
#define OB_DEVNI __device__ __noinline__
#define OB_HDNI __host__ __device__ __noinline__
#define OB_MDEV __device__ __forceinline__
#define OB_HD __host__ __device__ __forceinline__
#define OB_TABLE static __device__ const __align__(16)   // 16-byte aligned: the tables are cp.async.bulk sources (celt_mdct_fast.cuh)
#define OB_RESTRICT __restrict__

namespace ob {
OB_DEV int lane_id() { return threadIdx.x & 31; }
OB_DEV int warp_in_block() { return threadIdx.x >> 5; }
OB_DEV int warps_per_block() { return blockDim.x >> 5; }
OB_DEV int block_id() { return blockIdx.x; }
OB_DEV int num_blocks() { return gridDim.x; }
OB_DEV void wsync() { __syncwarp(); }
// Block barrier.  (Round 2 tried a per-team NAMED barrier here — bar.sync with a run-time id, several streams per CTA in the encoder's analysis
// kernel — and measured what it costs: a run-time barrier id makes ptxas reserve all 16 hardware barriers for EVERY kernel that inlines this helper,
// and CTAs per SM drop accordingly: ob_celt_dec_synth went from 2.0 to 4.2 ms, the one-stream analysis launch from 11.5 to 18-21 ms.  The 4 %
// the grouping bought the analysis kernel was not worth a second barrier helper threaded through the shared pitch code.)
OB_DEV void bsync() { __syncthreads(); }
template <class T> OB_DEV T shfl(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <class T> OB_DEV T shfl_up(T v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
template <class T> OB_DEV T shfl_down(T v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
template <class T> OB_DEV T shfl_xor(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
OB_DEV unsigned ballot(int p) { return __ballot_sync(0xffffffffu, p); }
// For thread-per-item code whose lanes run independent items: the lanes that are together right now, and a vote among exactly those.
// Used to line up an expensive step of a data-dependent loop (all lanes that need it execute it in the same pass).
OB_DEV unsigned converged_lanes() { return __activemask(); }
OB_DEV unsigned ballot_among(unsigned lanes, int p) { return __ballot_sync(lanes, p); }
OB_DEV int clz32(uint32_t v) { return __clz((int)v); }
OB_DEV int ffs32(uint32_t v) { return __ffs((int)v); }
OB_DEV float f_mul(float a, float b) { return __fmul_rn(a, b); }
OB_DEV float f_add(float a, float b) { return __fadd_rn(a, b); }
OB_DEV float f_sqrt(float a) { return __fsqrt_rn(a); }
OB_DEV float f_div(float a, float b) { return __fdiv_rn(a, b); }
OB_DEV unsigned char *dyn_smem() {
  extern __shared__ __align__(16) unsigned char ob_dyn_smem[];
  return ob_dyn_smem;
}
}  // namespace ob
#define OB_KERNEL(name, ...) extern "C" __global__ void name(__VA_ARGS__)

#elif defined(OB_HOST_EMU)

#include "emu_runtime.h"

#else
#error "simt.cuh: compile with nvcc, or with -DOB_HOST_EMU for the test-only host emulator"
#endif

namespace ob {
// high word of the signed 32x32 product (one IMAD.HI on the device)
OB_DEV int32_t mulhi32(int32_t a, int32_t b) {
#if defined(__CUDACC__)
  return __mulhi(a, b);
#else
  return (int32_t)(((int64_t)a * b) >> 32);
#endif
}
// 32 - clz, 0 for 0 (the range coder's EC_ILOG)
OB_DEV int ilog32(uint32_t v) { return 32 - clz32(v); }

template <class T> OB_DEV T warp_sum(T v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += shfl_xor(v, m);
  return v;
}
template <class T> OB_DEV T warp_or(T v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v |= shfl_xor(v, m);
  return v;
}
template <class T> OB_DEV T warp_max(T v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    T o = shfl_xor(v, m);
    v = o > v ? o : v;
  }
  return v;
}
}  // namespace ob
