// tests/emu/emu_runtime.h — TEST INFRASTRUCTURE ONLY.
//
// A tiny SIMT emulator: each CUDA "thread" of a warp is a user-level fiber (ucontext); every
// warp-synchronous primitive (shuffle, ballot, __syncwarp, __syncthreads) is a scheduling point
// at which the running lane yields to the next one, round robin, so a primitive completes only
// once all 32 lanes have reached it.  Lanes run strictly one after another between
// synchronisation points, so a kernel that forgets a __syncwarp() between a shared-memory write
// phase and a read phase produces WRONG RESULTS here (lane 0 reads before lane 5 has written)
// instead of working by accident — the emulator is deliberately stricter than the hardware.
//
// Blocks are one warp wide (blockDim.x == 32) and are executed sequentially.  This header is
// included by opus_b200/csrc/simt.cuh only under -DOB_HOST_EMU, i.e. only by tests/emu/*.cpp.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <vector>

#define OB_DEV static inline
#define OB_DEVNI static
#define OB_HDNI static
#define OB_MDEV inline
#define OB_HD static inline
#define OB_TABLE static const
#define OB_RESTRICT __restrict__
#define OB_KERNEL(name, ...) void name(__VA_ARGS__)

// the CUDA vector types the kernel sources use for 8- and 16-byte accesses
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };

namespace ob {
struct EmuState {
  ucontext_t sched;
  ucontext_t lanes[256];
  bool done[256];
  int cur, nthreads;
  int block, nblocks;
  int bar_arrived;
  unsigned bar_gen;
  uint64_t xchg[256];
  unsigned char *smem;
  size_t smem_bytes;
  long yields;
};
inline EmuState &emu() {
  static EmuState s;
  return s;
}

OB_DEV int lane_id() { return emu().cur & 31; }
OB_DEV int warp_in_block() { return emu().cur >> 5; }
OB_DEV int warps_per_block() { return emu().nthreads >> 5; }
OB_DEV int block_id() { return emu().block; }
OB_DEV int num_blocks() { return emu().nblocks; }
OB_DEV unsigned char *dyn_smem() { return emu().smem; }

// yield to the scheduler; returns when every other live lane has also reached a yield point
OB_DEV void emu_yield() {
  EmuState &e = emu();
  int me = e.cur;
  e.yields++;
  swapcontext(&e.lanes[me], &e.sched);
}
OB_DEV void wsync() { emu_yield(); }
// block barrier: a real arrival counter (warps may run a different number of yields between two barriers)
OB_DEV void bsync() {
  EmuState &e = emu();
  const unsigned gen = e.bar_gen;
  if (++e.bar_arrived == e.nthreads) {
    e.bar_arrived = 0;
    e.bar_gen++;
  } else {
    while (e.bar_gen == gen) emu_yield();
  }
  emu_yield();
}

template <class T> OB_DEV T shfl(T v, int src) {
  EmuState &e = emu();
  uint64_t b = 0;
  memcpy(&b, &v, sizeof(T));
  e.xchg[e.cur] = b;
  emu_yield();
  uint64_t r = e.xchg[(e.cur & ~31) | (src & 31)];
  emu_yield();
  T out;
  memcpy(&out, &r, sizeof(T));
  return out;
}
template <class T> OB_DEV T shfl_up(T v, int d) {
  int l = lane_id();
  T r = shfl(v, l - d < 0 ? l : l - d);
  return l - d < 0 ? v : r;
}
template <class T> OB_DEV T shfl_down(T v, int d) {
  int l = lane_id();
  T r = shfl(v, l + d > 31 ? l : l + d);
  return l + d > 31 ? v : r;
}
template <class T> OB_DEV T shfl_xor(T v, int m) { return shfl(v, lane_id() ^ m); }
OB_DEV unsigned ballot(int p) {
  EmuState &e = emu();
  e.xchg[e.cur] = p ? 1 : 0;
  emu_yield();
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (unsigned)(e.xchg[(e.cur & ~31) | i] & 1) << i;
  emu_yield();
  return r;
}
// thread-per-item code: on the emulator every lane runs its item alone, so "the lanes that are together" is the lane itself
OB_DEV unsigned converged_lanes() { return 1u; }
OB_DEV unsigned ballot_among(unsigned, int p) { return p ? 1u : 0u; }
OB_DEV int clz32(uint32_t v) { return v ? __builtin_clz(v) : 32; }
OB_DEV int ffs32(uint32_t v) { return __builtin_ffs((int)v); }
OB_DEV float f_mul(float a, float b) { return a * b; }
OB_DEV float f_add(float a, float b) { return a + b; }
OB_DEV float f_sqrt(float a) { return sqrtf(a); }
OB_DEV float f_div(float a, float b) { return a / b; }

// ---- launcher ---------------------------------------------------------------------------
struct EmuLaunch {
  void (*fn)(void *);
  void *arg;
};
inline void emu_trampoline(unsigned lo, unsigned hi) {
  EmuLaunch *l = (EmuLaunch *)(((uintptr_t)hi << 32) | (uintptr_t)lo);
  l->fn(l->arg);
  EmuState &e = emu();
  e.done[e.cur] = true;
  swapcontext(&e.lanes[e.cur], &e.sched);
}
// Run `fn(arg)` as a grid of `nblocks` blocks of `nthreads` (a multiple of 32, at most 256) threads with `smem_bytes` of
// dynamic shared memory.  With two warps per block every scheduling point synchronises the whole
// block, which is only valid for kernels whose warps execute identical sync sequences.
inline void emu_launch(void (*fn)(void *), void *arg, int nblocks, int nthreads, size_t smem_bytes) {
  EmuState &e = emu();
  const size_t STK = 1 << 20;
  std::vector<unsigned char> stacks((size_t)nthreads * STK);
  e.nthreads = nthreads;
  e.bar_arrived = 0;
  e.bar_gen = 0;
  std::vector<unsigned char> smem(smem_bytes + 64);
  EmuLaunch l = {fn, arg};
  e.nblocks = nblocks;
  e.smem = smem.data();
  e.smem_bytes = smem_bytes;
  for (int b = 0; b < nblocks; b++) {
    e.block = b;
    memset(smem.data(), 0xCD, smem.size());  // poison: uninitialised smem reads show up
    for (int i = 0; i < nthreads; i++) {
      getcontext(&e.lanes[i]);
      e.lanes[i].uc_stack.ss_sp = stacks.data() + i * STK;
      e.lanes[i].uc_stack.ss_size = STK;
      e.lanes[i].uc_link = &e.sched;
      e.done[i] = false;
      uintptr_t p = (uintptr_t)&l;
      makecontext(&e.lanes[i], (void (*)())emu_trampoline, 2, (unsigned)(p & 0xffffffffu), (unsigned)(p >> 32));
    }
    int live = nthreads;
    while (live) {
      live = 0;
      for (int i = 0; i < nthreads; i++) {
        if (e.done[i]) continue;
        e.cur = i;
        swapcontext(&e.sched, &e.lanes[i]);
        if (!e.done[i]) live++;
      }
    }
  }
}
}  // namespace ob
