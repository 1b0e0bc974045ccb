// nnet_umma.cuh — the dense int8 layer of dnn/ (cgemv8x4, dnn/vec.h:249-312 / vec_avx.h:780-881) on the 5th-generation tensor
// cores: tcgen05.mma kind::i8 with the accumulators in TMEM.
//
//   out[s][i] = (float)( sum_k x8[s][k] * w8[i][k] ) * scale[i] (+ bias[i]) (+ diag)         s < n streams, i < N outputs
//
// One persistent CTA per SM (320 threads, warp-specialised):
//   warp 0    producer: 1-D bulk copies (cp.async.bulk, completion on an mbarrier) — the whole weight matrix once, then one
//             128-stream activation tile per step.  Both operands are stored in HBM already in the tensor core's canonical
//             K-major no-swizzle layout ([k/16][row][16 bytes]: 8x16-byte core matrices), so a linear copy lands them right.
//   warp 1    allocates TMEM, then one elected lane issues tcgen05.mma (M=128, N=tile, K=32 per instruction) and commits to
//             mbarriers; the two N-halves of a layer use two accumulators, so the epilogue of one overlaps the MMAs of the other.
//   warps 2-9 epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> transpose in shared memory -> scale/bias in the
//             reference's float order -> coalesced stores.
// Integer accumulation is exact, so the result is bit-identical to the mma.sync kernel and to the oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ob {

struct UmmaLinearParams {
  const int8_t *wt;     // [Kp/16][Npad][16]   weights, zero padded
  const int8_t *xq;     // [ceil(n/128)][Kp/16][128][16]   quantised activations, zero padded
  const float *scale, *bias, *diag;
  const float *in;      // [n][M] float input (only the diag term reads it)
  float *out;           // [n][N]
  int n, M, N, Kp, Npad, n_tiles_n, tileN, a_stages, su, tmem_cols;
  int act;              // elementwise activation fused into the epilogue (compute_generic_dense / conv1d), 0 = none
  const float *glu_in;  // non-null: out = glu_in * sigmoid(y)  (compute_glu, dnn/nnet.c:107-120; may alias out)
};
struct UmmaQuantParams { int8_t *xq; const float *in; int n, M, Kp, su; int in_stride; };   // in_stride: floats between rows (>= M)
// The operand of a layer whose input is a concatenation: up to 6 pieces, each [n][width] with its own row stride and an
// optional per-row multiplier (FARGAN's pitch-gated prediction); quantised in place of materialising the float concatenation.
struct UmmaSeg { const float *ptr; const float *gate; int width, stride, gate_stride; };
struct UmmaQuantSegParams { int8_t *xq; UmmaSeg seg[6]; int nseg, n, M, Kp, su; };

#if defined(__CUDACC__)
namespace umma {
__device__ __forceinline__ uint32_t saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
// bounded wait: a protocol error traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; spin++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
// K-major, no swizzle (cute/atom/mma_traits_sm100.hpp: ((8,n),2):((1,SBO),LBO) in 16-byte units), descriptor version 1
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
        "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
        "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
}  // namespace umma

// float -> (u8|s8) activations (dnn/vec.h:253 / vec_avx.h:383) straight into the tensor core's operand layout;
// one thread per (16-byte k chunk, stream), streams fastest: a warp writes 512 contiguous bytes and every lane reads the 64
// contiguous bytes (two full sectors) of its own row with four 128-bit loads
extern "C" __global__ void __launch_bounds__(256) ob_nnet_quant_umma(UmmaQuantParams p) {
  const int KC = p.Kp >> 4;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long rows = (long)((p.n + 127) / 128) * 128;
  if (t >= rows * KC) return;
  const int kc = (int)(t / rows), s = (int)(t - (long)kc * rows);
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  if (s < p.n) {
    const float *src = p.in + (size_t)s * p.in_stride + kc * 16;
    float x[16];
    if (kc * 16 + 16 <= p.M && (p.in_stride & 3) == 0 && ((size_t)p.in & 15) == 0) {
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const float4 f = *(const float4 *)(src + 4 * b);
        x[4 * b] = f.x; x[4 * b + 1] = f.y; x[4 * b + 2] = f.z; x[4 * b + 3] = f.w;
      }
    } else {
#pragma unroll
      for (int b = 0; b < 16; b++) x[b] = kc * 16 + b < p.M ? src[b] : 0.f;
    }
#pragma unroll
    for (int b = 0; b < 16; b++) {
      int q = 0;
      if (kc * 16 + b < p.M) {
        q = (int)floor(0.5 + (double)__fmul_rn(127.f, x[b]));
        if (p.su) q += 127;
      }
      w[b >> 2] |= ((uint32_t)q & 0xffu) << (8 * (b & 3));
    }
  }
  uint4 *dst = (uint4 *)(p.xq + (((size_t)(s >> 7) * KC + kc) * 128 + (s & 127)) * 16);
  *dst = make_uint4(w[0], w[1], w[2], w[3]);
}

// (every piece's width and row stride must be a multiple of 4 floats: a group of 4 columns then lies in one piece, is read with
// one 128-bit load when aligned, and the segment walk runs per group; lanes of a warp share kc, so the walk is warp-uniform)
extern "C" __global__ void __launch_bounds__(256) ob_nnet_quant_umma_seg(UmmaQuantSegParams p) {
  const int KC = p.Kp >> 4;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long rows = (long)((p.n + 127) / 128) * 128;
  if (t >= rows * KC) return;
  const int kc = (int)(t / rows), s = (int)(t - (long)kc * rows);
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  if (s < p.n) {
    int g = 0, base = 0;
#pragma unroll
    for (int gq = 0; gq < 4; gq++) {
      const int j = kc * 16 + gq * 4;
      if (j >= p.M) break;
      while (g < p.nseg - 1 && j >= base + p.seg[g].width) { base += p.seg[g].width; g++; }
      const UmmaSeg sg = p.seg[g];
      const float *src = sg.ptr + (size_t)s * sg.stride + (j - base);
      float x[4];
      if ((((size_t)src) & 15) == 0) {
        const float4 f = *(const float4 *)src;
        x[0] = f.x; x[1] = f.y; x[2] = f.z; x[3] = f.w;
      } else {
        x[0] = src[0]; x[1] = src[1]; x[2] = src[2]; x[3] = src[3];
      }
      if (sg.gate) {
        const float gt = sg.gate[(size_t)s * sg.gate_stride];
#pragma unroll
        for (int b = 0; b < 4; b++) x[b] = __fmul_rn(gt, x[b]);
      }
      uint32_t word = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        int q = (int)floor(0.5 + (double)__fmul_rn(127.f, x[b]));
        if (p.su) q += 127;
        word |= ((uint32_t)q & 0xffu) << (8 * b);
      }
      w[gq] = word;
    }
  }
  uint4 *dst = (uint4 *)(p.xq + (((size_t)(s >> 7) * KC + kc) * 128 + (s & 127)) * 16);
  *dst = make_uint4(w[0], w[1], w[2], w[3]);
}

extern "C" __global__ void __launch_bounds__(320, 1) ob_nnet_linear_i8_umma(UmmaLinearParams p) {
  using namespace umma;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w_bytes = (uint32_t)p.Npad * p.Kp, a_bytes = 128u * p.Kp;
  unsigned char *sW = smem_raw, *sA = smem_raw + w_bytes;
  uint64_t *bars = (uint64_t *)(sA + (size_t)p.a_stages * a_bytes);   // [0] w_full, [1..2] a_full, [3..4] a_empty, [5..6] acc_full, [7..8] acc_empty
  uint32_t *tmem_slot = (uint32_t *)(bars + 9);
  float *stage = (float *)(bars + 12) + (warp >= 2 ? (warp - 2) * 32 * 33 : 0);   // per epilogue warp: 32 x 33 floats
  const uint32_t b_w = saddr(&bars[0]);
  auto b_afull = [&](int s) { return saddr(&bars[1 + s]); };
  auto b_aempty = [&](int s) { return saddr(&bars[3 + s]); };
  auto b_accfull = [&](int s) { return saddr(&bars[5 + s]); };
  auto b_accempty = [&](int s) { return saddr(&bars[7 + s]); };
  if (threadIdx.x == 0) {
    mbar_init(b_w, 1);
    for (int s = 0; s < 2; s++) { mbar_init(b_afull(s), 1); mbar_init(b_aempty(s), 1); mbar_init(b_accfull(s), 1); mbar_init(b_accempty(s), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_mtiles = (p.n + 127) / 128, KC = p.Kp >> 4, KS = p.Kp >> 5;
  constexpr uint32_t kChunk = 16384;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(b_w, w_bytes);
      for (uint32_t o = 0; o < w_bytes; o += kChunk) bulk_g2s(saddr(sW + o), p.wt + o, (w_bytes - o < kChunk ? w_bytes - o : kChunk), b_w);
      int it = 0;
      for (int t = blockIdx.x; t < n_mtiles; t += gridDim.x, it++) {
        const int s = it % p.a_stages;
        const uint32_t ph = (uint32_t)(it / p.a_stages) & 1u;
        mbar_wait(b_aempty(s), ph ^ 1u);
        mbar_expect_tx(b_afull(s), a_bytes);
        const int8_t *src = p.xq + (size_t)t * a_bytes;
        for (uint32_t o = 0; o < a_bytes; o += kChunk) bulk_g2s(saddr(sA + (size_t)s * a_bytes + o), src + o, (a_bytes - o < kChunk ? a_bytes - o : kChunk), b_afull(s));
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D = S32, A = u8|s8, B = s8, both K-major, N>>3, M>>4
      const uint32_t idesc = (2u << 4) | ((p.su ? 0u : 1u) << 7) | (1u << 10) | ((uint32_t)(p.tileN >> 3) << 17) | ((128u >> 4) << 24);
      mbar_wait(b_w, 0);
      int it = 0;
      for (int t = blockIdx.x; t < n_mtiles; t += gridDim.x, it++) {
        const int s = it % p.a_stages;
        const uint32_t ph = (uint32_t)(it / p.a_stages) & 1u;
        mbar_wait(b_afull(s), ph);
        fence_after();
        const uint32_t a0 = saddr(sA + (size_t)s * a_bytes);
        for (int nt = 0; nt < p.n_tiles_n; nt++) {
          mbar_wait(b_accempty(nt), ((uint32_t)it & 1u) ^ 1u);
          fence_after();
          const uint32_t b0 = saddr(sW) + (uint32_t)nt * p.tileN * 16u;
          for (int ks = 0; ks < KS; ks++) {
            const uint64_t ad = smem_desc(a0 + (uint32_t)ks * 2u * 2048u, 2048u, 128u);
            const uint64_t bd = smem_desc(b0 + (uint32_t)ks * 2u * (uint32_t)p.Npad * 16u, (uint32_t)p.Npad * 16u, 128u);
            mma_i8(tmem_base + (uint32_t)nt * p.tileN, ad, bd, idesc, ks > 0 ? 1u : 0u);
          }
          commit(b_accfull(nt));   // arrives when the MMAs above have written the accumulator
        }
        commit(b_aempty(s));       // ... and have finished reading this activation tile
      }
    }
    __syncwarp();
  } else {
    // 8 epilogue warps: warp & 3 = the TMEM lane quadrant it may read; the two warps of a quadrant split the 32-column chunks
    const int q = warp & 3, g = (warp - 2) >> 2;
    int it = 0;
    for (int t = blockIdx.x; t < n_mtiles; t += gridDim.x, it++) {
      const int row0 = t * 128 + q * 32;
      for (int nt = 0; nt < p.n_tiles_n; nt++) {
        mbar_wait(b_accfull(nt), (uint32_t)it & 1u);
        fence_after();
        for (int c0 = g * 32; c0 < p.tileN; c0 += 64) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(nt * p.tileN + c0), v);
          // lane = row here; transpose through shared memory so that a store instruction covers 32 consecutive outputs of a row
#pragma unroll
          for (int j = 0; j < 32; j++) stage[lane * 33 + j] = __int_as_float((int)v[j]);
          __syncwarp();
          const int i = nt * p.tileN + c0 + lane;
          const bool col_ok = c0 + lane < p.tileN && i < p.N;
          const float sc = col_ok ? p.scale[i] : 0.f, bi = (col_ok && p.bias) ? p.bias[i] : 0.f;
          float y[32];
#pragma unroll
          for (int r = 0; r < 32; r++) {
            y[r] = __fmul_rn((float)__float_as_int(stage[r * 33 + lane]), sc);
            if (p.bias) y[r] = __fadd_rn(y[r], bi);
          }
          if (p.diag && col_ok) {   // GRU recurrent matrices: + diag[i] * in[i mod M] (dnn/nnet_arch.h:145-152); loads issued together
            const int m = i >= 2 * p.M ? i - 2 * p.M : (i >= p.M ? i - p.M : i);
            const float dg = p.diag[i];
            float xin[32];
#pragma unroll
            for (int r = 0; r < 32; r++) xin[r] = row0 + r < p.n ? p.in[(size_t)(row0 + r) * p.M + m] : 0.f;
#pragma unroll
            for (int r = 0; r < 32; r++) y[r] = __fadd_rn(y[r], __fmul_rn(dg, xin[r]));
          }
          if (p.glu_in && col_ok) {
            float gin[32];
#pragma unroll
            for (int r = 0; r < 32; r++) gin[r] = row0 + r < p.n ? p.glu_in[(size_t)(row0 + r) * p.N + i] : 0.f;
#pragma unroll
            for (int r = 0; r < 32; r++) y[r] = __fmul_rn(gin[r], nn_sigmoid(y[r]));
          } else if (p.act) {
#pragma unroll
            for (int r = 0; r < 32; r++) y[r] = nn_act(y[r], p.act);
          }
          if (col_ok) {
#pragma unroll
            for (int r = 0; r < 32; r++)
              if (row0 + r < p.n) p.out[(size_t)(row0 + r) * p.N + i] = y[r];
          }
          __syncwarp();
        }
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_accempty(nt));
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
}
#endif  // __CUDACC__

}  // namespace ob
