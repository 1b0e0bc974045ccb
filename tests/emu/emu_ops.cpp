// tests/emu/emu_ops.cpp — TEST INFRASTRUCTURE ONLY: the encoder-side op KERNEL SOURCES
// (opus_b200/csrc/celt_ops.cuh) on the host SIMT emulator.
#define OB_HOST_EMU 1
#include "../../opus_b200/csrc/celt_types.h"
#include "../../opus_b200/csrc/celt_ops.cuh"

using namespace ob;

static void mdct_entry(void *a) { const MdctFwdParams &p = *(const MdctFwdParams *)a; mdct_forward_warp(p, block_id(), dyn_smem()); }
static void xcorr_entry(void *a) { const XcorrParams &p = *(const XcorrParams *)a; pitch_xcorr_warp(p, block_id(), dyn_smem()); }
static void rot_entry(void *a) { const VqParams &p = *(const VqParams *)a; exp_rotation_warp(p, block_id(), dyn_smem()); }
static void search_entry(void *a) { const VqParams &p = *(const VqParams *)a; pvq_search_warp(p, block_id(), dyn_smem()); }

extern "C" {
void emu_mdct_forward(const float *in, float *out, int n, int LM, int transient) {
  MdctFwdParams p;
  p.in = in; p.out = out; p.n_items = n; p.in_stride = (120 << LM) + 120; p.out_stride = 120 << LM; p.LM = LM; p.transient = transient;
  emu_launch(mdct_entry, &p, n, 32, sizeof(MdctWS));
}
void emu_pitch_xcorr(const float *x, const float *y, float *xc, int n, int len, int max_pitch) {
  XcorrParams p;
  p.x = x; p.y = y; p.xcorr = xc; p.n_items = n; p.len = len; p.max_pitch = max_pitch;
  emu_launch(xcorr_entry, &p, n, 32, (((len + 3) & ~3) + len + max_pitch + 4) * sizeof(float));
}
void emu_exp_rotation(float *X, const int *K, int n, int N, int dir, int stride, int spread) {
  VqParams p;
  p.X = X; p.iy = nullptr; p.yy = nullptr; p.K = K; p.n_items = n; p.N = N; p.dir = dir; p.stride = stride; p.spread = spread;
  emu_launch(rot_entry, &p, (n + 31) / 32, 32, (size_t)N * 32 * sizeof(float));
}
void emu_pvq_search(float *X, int *iy, float *yy, const int *K, int n, int N) {
  VqParams p;
  p.X = X; p.iy = iy; p.yy = yy; p.K = K; p.n_items = n; p.N = N; p.dir = 1; p.stride = 1; p.spread = 0;
  emu_launch(search_entry, &p, (n + 31) / 32, 32, (size_t)N * 32 * (2 * sizeof(float) + 1));
}
}
