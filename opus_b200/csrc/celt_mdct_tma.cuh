// celt_mdct_tma.cuh — the persistent long-block forward MDCT kernel: the register FFT of celt_mdct_fast.cuh with tables and inputs staged by
// cp.async.bulk.  Included by opus_b200.cu after celt_ops.cuh (MdctFwdParams).
#pragma once
#include "celt_mdct_fast.cuh"
#include "celt_ops.cuh"

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
