// nnet_host.cuh — kernels' entry points and the C ABI of the dnn/ primitives (included at the end of opus_b200.cu).
#pragma once
#include "nnet.cuh"
#include "nnet_umma.cuh"

extern "C" __global__ void __launch_bounds__(256) ob_nnet_activation(NnetActParams p) {
  nnet_activation_thread(p, (long)blockIdx.x * blockDim.x + threadIdx.x);
}
extern "C" __global__ void __launch_bounds__(256) ob_nnet_quant(NnetQuantParams p) {
  nnet_quant_thread(p, (long)blockIdx.x * blockDim.x + threadIdx.x);
}
extern "C" __global__ void __launch_bounds__(128) ob_nnet_linear_i8_mma_u8(NnetLinearParams p) {
  nnet_linear_i8_mma_warp<true>(p, blockIdx.x * 32, blockIdx.y * 256 + (threadIdx.x >> 5) * 64);
}
extern "C" __global__ void __launch_bounds__(128) ob_nnet_linear_i8_mma_s8(NnetLinearParams p) {
  nnet_linear_i8_mma_warp<false>(p, blockIdx.x * 32, blockIdx.y * 256 + (threadIdx.x >> 5) * 64);
}
extern "C" __global__ void __launch_bounds__(128) ob_nnet_linear_i8_sparse(NnetLinearParams p) {
  nnet_linear_i8_sparse_thread(p, (long)blockIdx.x * blockDim.x + threadIdx.x);
}
extern "C" __global__ void __launch_bounds__(128) ob_nnet_linear_f32_sparse(NnetLinearParams p) {
  nnet_linear_f32_sparse_thread(p, (long)blockIdx.x * blockDim.x + threadIdx.x);
}
extern "C" __global__ void __launch_bounds__(256) ob_nnet_linear_f32(NnetLinearParams p) {
  __shared__ NnetF32Smem sm;
  nnet_linear_f32_block(p, blockIdx.x * kNnF32Tile, blockIdx.y * kNnF32Tile, sm, threadIdx.x);
}
extern "C" __global__ void __launch_bounds__(256) ob_nnet_linear_noweights(NnetLinearParams p) {
  nnet_linear_noweights_thread(p, (long)blockIdx.x * blockDim.x + threadIdx.x);
}
extern "C" __global__ void __launch_bounds__(256) ob_nnet_gru(NnetGruParams p) { nnet_gru_thread(p, (long)blockIdx.x * blockDim.x + threadIdx.x); }
extern "C" __global__ void __launch_bounds__(256) ob_nnet_glu(NnetGluParams p) { nnet_glu_thread(p, (long)blockIdx.x * blockDim.x + threadIdx.x); }
extern "C" __global__ void __launch_bounds__(256) ob_nnet_conv1d(NnetConvParams p) {
  nnet_conv1d_thread(p, (long)blockIdx.x * blockDim.x + threadIdx.x);
}
extern "C" __global__ void __launch_bounds__(256) ob_nnet_conv2d(NnetConv2dParams p) {
  nnet_conv2d_thread(p, (long)blockIdx.x * blockDim.x + threadIdx.x);
}
extern "C" __global__ void __launch_bounds__(256) ob_nnet_conv2d_shift(NnetConv2dParams p) {
  nnet_conv2d_shift_thread(p, (long)blockIdx.x * blockDim.x + threadIdx.x);
}

struct OpusB200Linear {
  OpusB200Ctx *ctx;
  NnetLinear L;            // device pointers
  void *blob;              // one allocation behind all of them
  // dense int8 layers also keep the weights in the tcgen05 operand layout (nnet_umma.cuh); null when the layer does not qualify
  int8_t *wt_umma;
  int Kp, Npad, n_tiles_n, tileN, a_stages, tmem_cols;
  size_t umma_smem;
};
struct OpusB200Conv2d {
  OpusB200Ctx *ctx;
  float *w, *bias;
  int in_ch, out_ch, ktime, kheight;
};

// growable per-context scratch (quantised activations, GRU / GLU / conv temporaries); work is stream-ordered
static int nnet_scratch(OpusB200Ctx *c, int slot, size_t bytes, void **out) {
  if (c->scratch_bytes[slot] < bytes) {
    if (c->scratch[slot]) { cudaStreamSynchronize(c->stream); cudaFree(c->scratch[slot]); }
    c->scratch[slot] = nullptr; c->scratch_bytes[slot] = 0;
    c->scratch_gen++;   // graphs captured with the old pointer are stale (ob_graph_run re-captures)
    const size_t want = bytes + bytes / 4;
    if (cudaMalloc(&c->scratch[slot], want) != cudaSuccess) return -1;
    c->scratch_bytes[slot] = want;
  }
  *out = c->scratch[slot];
  return 0;
}
static inline unsigned nblk(long total, int per) { return (unsigned)((total + per - 1) / per); }

extern "C" int opusb200_nnet_linear_create(OpusB200Ctx *c, const OpusB200LinearLayer *h, int su_bias, OpusB200Linear **out) {
  if (!c || !h || !out || h->nb_inputs <= 0 || h->nb_outputs <= 0 || h->nb_inputs > 2048) return OPUSB200_BAD_ARG;
  *out = nullptr;
  const int M = h->nb_inputs, N = h->nb_outputs;
  const bool i8 = !h->float_weights && h->weights, f32 = h->float_weights != nullptr, sparse = h->weights_idx != nullptr;
  if ((i8 || sparse) && ((M & 3) || (N & 7))) return OPUSB200_BAD_ARG;   // 8x4 blocking
  if (i8 && !h->scale) return OPUSB200_BAD_ARG;
  if (h->diag && N != 3 * M) return OPUSB200_BAD_ARG;                     // dnn/nnet_arch.h:146
  OB_CUDA(cudaSetDevice(c->device));
  // sparse index list -> per-row-block offsets + block column positions
  int *rb_off = nullptr, *pos = nullptr;
  long nb = 0;
  if (sparse) {
    rb_off = (int *)malloc(sizeof(int) * (N / 8 + 1));
    const int *idx = h->weights_idx;
    for (int rb = 0; rb < N / 8; rb++) { rb_off[rb] = (int)nb; const int cnt = *idx++; nb += cnt; idx += cnt; }
    rb_off[N / 8] = (int)nb;
    pos = (int *)malloc(sizeof(int) * (nb ? nb : 1));
    idx = h->weights_idx;
    long k = 0;
    for (int rb = 0; rb < N / 8; rb++) { const int cnt = *idx++; for (int j = 0; j < cnt; j++) pos[k++] = *idx++; }
    for (long j = 0; j < nb; j++) if (pos[j] < 0 || pos[j] + 4 > M || (pos[j] & 3)) { free(rb_off); free(pos); return OPUSB200_BAD_ARG; }
  } else nb = (long)(N / 8) * (M / 4);
  const size_t w_bytes = f32 ? (sparse ? (size_t)nb * 32 * 4 : (size_t)M * N * 4) : (i8 ? (size_t)nb * 32 : 0);
  const float *eff_bias = (i8 && su_bias) ? h->subias : h->bias;           // dnn/nnet_arch.h:136-139
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t o_w = 0, o_bias = o_w + al(w_bytes), o_diag = o_bias + al(eff_bias ? N * 4 : 0), o_scale = o_diag + al(h->diag ? N * 4 : 0),
               o_rb = o_scale + al(i8 ? N * 4 : 0), o_pos = o_rb + al(sparse ? (N / 8 + 1) * 4 : 0), total = o_pos + al(sparse ? nb * 4 : 0);
  OpusB200Linear *l = new (std::nothrow) OpusB200Linear;
  if (!l) { free(rb_off); free(pos); return OPUSB200_ALLOC_FAIL; }
  l->ctx = c; l->blob = nullptr;
  if (cudaMalloc(&l->blob, total ? total : 256) != cudaSuccess) { delete l; free(rb_off); free(pos); return OPUSB200_ALLOC_FAIL; }
  char *b = (char *)l->blob;
  NnetLinear &L = l->L;
  memset(&L, 0, sizeof(L));
  L.M = M; L.N = N; L.su = (i8 && su_bias) ? 1 : 0;
  cudaError_t e = cudaSuccess;
  auto up = [&](size_t off, const void *src, size_t bytes) { if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(b + off, src, bytes, cudaMemcpyHostToDevice, c->stream); };
  if (f32) { up(o_w, h->float_weights, w_bytes); L.wf = (const float *)(b + o_w); }
  else if (i8) { up(o_w, h->weights, w_bytes); L.w8 = (const int8_t *)(b + o_w); up(o_scale, h->scale, (size_t)N * 4); L.scale = (const float *)(b + o_scale); }
  if (eff_bias) { up(o_bias, eff_bias, (size_t)N * 4); L.bias = (const float *)(b + o_bias); }
  if (h->diag) { up(o_diag, h->diag, (size_t)N * 4); L.diag = (const float *)(b + o_diag); }
  if (sparse) {
    up(o_rb, rb_off, (size_t)(N / 8 + 1) * 4); L.rb_off = (const int *)(b + o_rb);
    up(o_pos, pos, (size_t)nb * 4); L.pos = (const int *)(b + o_pos);
  }
  l->wt_umma = nullptr;
  int8_t *wt_host = nullptr;
  if (i8 && !sparse && N <= 512 && getenv("OPUSB200_NNET_NO_UMMA") == nullptr) {
    // tcgen05 tiling: one or two N tiles (multiples of 16, <= 256) per 128-stream tile; K padded to the MMA's 32
    const int nt = N <= 256 ? 1 : 2, tileN = (((N + nt - 1) / nt) + 15) & ~15, Npad = nt * tileN, Kp = (M + 31) & ~31;
    int cols = 32;
    while (cols < ((Npad + 31) & ~31)) cols <<= 1;
    const size_t wb = (size_t)Npad * Kp, ab = (size_t)128 * Kp, extra = 12 * 8 + 8 * 32 * 33 * 4 + 128;   // barriers + TMEM slot, the epilogue's transpose tiles
    int max_optin = 0;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
    const int stages = wb + 2 * ab + extra <= (size_t)max_optin ? 2 : (wb + ab + extra <= (size_t)max_optin ? 1 : 0);
    if (stages && cudaMalloc(&l->wt_umma, wb) == cudaSuccess) {
      wt_host = (int8_t *)calloc(wb, 1);
      const int CB = M / 4;
      for (int i = 0; i < N; i++)
        for (int k = 0; k < M; k++)   // reference blocking: 8 rows x 4 columns per 32-byte block (dnn/vec.h:249)
          wt_host[((size_t)(k >> 4) * Npad + i) * 16 + (k & 15)] = h->weights[((size_t)(i >> 3) * CB + (k >> 2)) * 32 + (i & 7) * 4 + (k & 3)];
      if (e == cudaSuccess) e = cudaMemcpyAsync(l->wt_umma, wt_host, wb, cudaMemcpyHostToDevice, c->stream);
      l->Kp = Kp; l->Npad = Npad; l->n_tiles_n = nt; l->tileN = tileN; l->a_stages = stages; l->tmem_cols = cols;
      l->umma_smem = wb + stages * ab + extra;
    }
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  free(rb_off); free(pos); free(wt_host);
  if (e != cudaSuccess) { cudaFree(l->blob); cudaFree(l->wt_umma); delete l; return OPUSB200_INTERNAL_ERROR; }
  *out = l;
  return OPUSB200_OK;
}
extern "C" void opusb200_nnet_linear_destroy(OpusB200Linear *l) {
  if (!l) return;
  cudaSetDevice(l->ctx->device);
  cudaStreamSynchronize(l->ctx->stream);
  cudaFree(l->blob);
  cudaFree(l->wt_umma);
  delete l;
}

// device-pointer core of compute_linear: out [n][N], in [n][M]
// ---- CUDA-graph cache for the launch-bound model graphs (FARGAN frame: ~200 launches, DRED decode: ~60 per latent) -----------------
// The first call with a given key (pointers + shape) runs eagerly (it sizes the scratch buffers and sets kernel attributes), the
// second one is captured from the very same enqueue code and every later one replays the instantiated graph.  A failed capture
// (e.g. a scratch buffer that had to grow) falls back to eager enqueueing.  OPUSB200_NO_GRAPHS=1 disables it.
// The key is the caller's pointers + shape AND the context's scratch generation: the captured kernels hold the shared scratch pointers,
// and another object on the same context may have grown (= reallocated) them since.
struct ObGraphCache { cudaGraphExec_t exec; const void *a, *b, *c3; long k; uint64_t gen; int state; };   // state: 0 new, 1 warmed, 2 captured, -1 disabled
template <class F> static int ob_graph_run(OpusB200Ctx *c, ObGraphCache &g, const void *a, const void *b, long k, F &&enqueue, const void *c3 = nullptr) {
  static const bool off = getenv("OPUSB200_NO_GRAPHS") != nullptr;
  if (off || g.state < 0) return enqueue();
  if (g.a != a || g.b != b || g.c3 != c3 || g.k != k || g.gen != c->scratch_gen) {
    if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    g.a = a; g.b = b; g.c3 = c3; g.k = k; g.gen = c->scratch_gen; g.state = 0;
  }
  if (g.state == 2) { OB_CUDA(cudaGraphLaunch(g.exec, c->stream)); return OPUSB200_OK; }
  if (g.state == 0) {   // the eager warm-up may grow the scratch: key the next (captured) run on the generation it leaves behind
    g.state = 1;
    const int r0 = enqueue();
    g.gen = c->scratch_gen;
    return r0;
  }
  if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); g.state = -1; return enqueue(); }
  const int r = enqueue();
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
  if (r != OPUSB200_OK || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    g.state = -1;
    return r != OPUSB200_OK ? r : enqueue();
  }
  const cudaError_t ei = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ei != cudaSuccess) { cudaGetLastError(); g.exec = nullptr; g.state = -1; return enqueue(); }
  g.state = 2;
  OB_CUDA(cudaGraphLaunch(g.exec, c->stream));
  return OPUSB200_OK;
}

// the tcgen05 GEMM on an operand that is already quantised in the tensor core's layout (xq: [ceil(n/128)][Kp/16][128][16])
static int nnet_umma_gemm(OpusB200Ctx *c, const OpusB200Linear *l, float *out, const int8_t *xq, const float *in_for_diag, int n, int act = NN_LINEAR,
                          const float *glu_in = nullptr, bool *fused = nullptr) {
  const NnetLinear &L = l->L;
  const int mt = (n + 127) / 128;
  UmmaLinearParams u;
  u.wt = l->wt_umma; u.xq = xq; u.scale = L.scale; u.bias = L.bias; u.diag = L.diag; u.in = in_for_diag; u.out = out; u.n = n; u.M = L.M; u.N = L.N;
  u.Kp = l->Kp; u.Npad = l->Npad; u.n_tiles_n = l->n_tiles_n; u.tileN = l->tileN; u.a_stages = l->a_stages; u.su = L.su; u.tmem_cols = l->tmem_cols;
  // Fusing tanh/sigmoid (an IEEE division each) into the 8-warp epilogue lengthens the kernel's critical path more than the
  // separate full-occupancy activation kernel costs (measured: FARGAN 4.61 vs 4.49 ms), so it is off unless asked for.
  static const bool fuse = getenv("OPUSB200_NNET_FUSE") != nullptr;
  u.act = (fuse && act != NN_SOFTMAX) ? act : NN_LINEAR; u.glu_in = fuse ? glu_in : nullptr;
  if (fused) *fused = fuse && act != NN_SOFTMAX;
  if (l->umma_smem > c->umma_smem_set) {   // per device (= per context), not per process
    OB_CUDA(cudaFuncSetAttribute(ob_nnet_linear_i8_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l->umma_smem));
    c->umma_smem_set = l->umma_smem;
  }
  if (!c->sms) cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, c->device);
  const int sms = c->sms;
  ob_nnet_linear_i8_umma<<<mt < sms ? mt : sms, 320, l->umma_smem, c->stream>>>(u);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  return OPUSB200_OK;
}
// quantise a concatenation of up to 6 pieces straight into the operand of layer l (which must take the tcgen05 path)
static int nnet_umma_quant_seg(OpusB200Ctx *c, const OpusB200Linear *l, const UmmaSeg *segs, int nseg, int n, const int8_t **xq_out) {
  if (!l->wt_umma || nseg < 1 || nseg > 6) return OPUSB200_BAD_ARG;
  int tot = 0;
  for (int i = 0; i < nseg; i++) {
    if ((segs[i].width & 3) || (segs[i].stride & 3)) return OPUSB200_BAD_ARG;   // groups of 4 columns stay inside one piece
    tot += segs[i].width;
  }
  if (tot != l->L.M) return OPUSB200_BAD_ARG;
  const int mt = (n + 127) / 128;
  void *xq;
  if (nnet_scratch(c, 0, (size_t)mt * 128 * l->Kp, &xq)) return OPUSB200_ALLOC_FAIL;
  UmmaQuantSegParams q;
  q.xq = (int8_t *)xq; q.nseg = nseg; q.n = n; q.M = l->L.M; q.Kp = l->Kp; q.su = l->L.su;
  for (int i = 0; i < nseg; i++) q.seg[i] = segs[i];
  ob_nnet_quant_umma_seg<<<nblk((long)mt * 128 * (l->Kp / 16), 256), 256, 0, c->stream>>>(q);
  c->launches++;
  *xq_out = (const int8_t *)xq;
  return OPUSB200_OK;
}
// act / glu_in: fused into the tcgen05 kernel's epilogue when the layer takes that path (*fused says so); other paths ignore them
// in_stride: the rows of `in` are in_stride floats apart (a layer reading the leading columns of a wider activation matrix, as
// the DRED decoder's layers do); int8 layers only -- their quantiser absorbs the stride.  0 = contiguous.
static int nnet_linear_dev(OpusB200Ctx *c, const OpusB200Linear *l, float *out, const float *in, int n, int act = NN_LINEAR,
                           const float *glu_in = nullptr, bool *fused = nullptr, int in_stride = 0) {
  if (fused) *fused = false;
  if (in_stride && in_stride != l->L.M && (!l->L.w8 || l->L.diag)) return OPUSB200_BAD_ARG;
  if (!in_stride) in_stride = l->L.M;
  NnetLinearParams p;
  p.L = l->L; p.out = out; p.in = in; p.n = n; p.xq = nullptr; p.Mp = (l->L.M + 31) & ~31;
  const NnetLinear &L = l->L;
  if (L.w8 && l->wt_umma) {   // dense int8 layer: tcgen05 / TMEM path (nnet_umma.cuh)
    const int mt = (n + 127) / 128;
    void *xq;
    if (nnet_scratch(c, 0, (size_t)mt * 128 * l->Kp, &xq)) return OPUSB200_ALLOC_FAIL;
    UmmaQuantParams q;
    q.xq = (int8_t *)xq; q.in = in; q.n = n; q.M = L.M; q.Kp = l->Kp; q.su = L.su; q.in_stride = in_stride;
    ob_nnet_quant_umma<<<nblk((long)mt * 128 * (l->Kp / 16), 256), 256, 0, c->stream>>>(q);
    c->launches++;
    return nnet_umma_gemm(c, l, out, (const int8_t *)xq, in, n, act, glu_in, fused);
  }
  if (L.w8) {
    void *xq;
    if (nnet_scratch(c, 0, (size_t)n * p.Mp, &xq)) return OPUSB200_ALLOC_FAIL;
    NnetQuantParams q;
    q.xq = (int8_t *)xq; q.in = in; q.n = n; q.M = L.M; q.Mp = p.Mp; q.su = L.su; q.in_stride = in_stride;
    ob_nnet_quant<<<nblk((long)n * (p.Mp / 4), 256), 256, 0, c->stream>>>(q);
    p.xq = (const int8_t *)xq;
    c->launches++;
    if (L.rb_off) ob_nnet_linear_i8_sparse<<<nblk((long)n * (L.N / 8), 128), 128, 0, c->stream>>>(p);
    else {
      dim3 grid((n + 31) / 32, (L.N + 255) / 256);
      if (L.su) ob_nnet_linear_i8_mma_u8<<<grid, 128, 0, c->stream>>>(p);
      else ob_nnet_linear_i8_mma_s8<<<grid, 128, 0, c->stream>>>(p);
    }
  } else if (L.wf) {
    if (L.rb_off) ob_nnet_linear_f32_sparse<<<nblk((long)n * (L.N / 8), 128), 128, 0, c->stream>>>(p);
    else {
      dim3 grid((n + kNnF32Tile - 1) / kNnF32Tile, (L.N + kNnF32Tile - 1) / kNnF32Tile);
      ob_nnet_linear_f32<<<grid, 256, 0, c->stream>>>(p);
    }
  } else ob_nnet_linear_noweights<<<nblk((long)n * L.N, 256), 256, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  return OPUSB200_OK;
}
static int nnet_act_dev(OpusB200Ctx *c, float *out, const float *in, int n, int N, int act) {
  if (act == NN_LINEAR && in == out) return OPUSB200_OK;
  NnetActParams p;
  p.out = out; p.in = in; p.n = n; p.N = N; p.act = act;
  ob_nnet_activation<<<nblk(act == NN_SOFTMAX ? n : (long)n * N, 256), 256, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  return OPUSB200_OK;
}

// host <-> device staging shared by the entry points below
struct NnetStage {
  OpusB200Ctx *c;
  DevBuf bufs[4];
  int used = 0;
  explicit NnetStage(OpusB200Ctx *ctx) : c(ctx) {}
  // returns the device view of `host` (uploads when copy_in)
  void *in(const void *host, size_t bytes, int on_device, bool copy_in, int *err) {
    if (on_device) return (void *)host;
    DevBuf &d = bufs[used++];
    if (d.alloc(bytes)) { *err = OPUSB200_ALLOC_FAIL; return nullptr; }
    if (copy_in && cudaMemcpyAsync(d.p, host, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) *err = OPUSB200_INTERNAL_ERROR;
    return d.p;
  }
  int out(void *host, const void *dev, size_t bytes, int on_device) {
    if (on_device) return OPUSB200_OK;
    return cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess ? OPUSB200_OK : OPUSB200_INTERNAL_ERROR;
  }
  int finish(int on_device) {
    if (on_device) return OPUSB200_OK;
    return cudaStreamSynchronize(c->stream) == cudaSuccess ? OPUSB200_OK : OPUSB200_INTERNAL_ERROR;
  }
};

extern "C" int opusb200_nnet_compute_linear(OpusB200Ctx *c, const OpusB200Linear *l, float *out, const float *in, int n, int on_device) {
  if (!c || !l || !out || !in || n <= 0 || in == out) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  NnetStage st(c);
  int err = OPUSB200_OK;
  const size_t ib = sizeof(float) * (size_t)n * l->L.M, ob_ = sizeof(float) * (size_t)n * l->L.N;
  const float *din = (const float *)st.in(in, ib, on_device, true, &err);
  float *dout = (float *)st.in(out, ob_, on_device, false, &err);
  if (err) return err;
  if ((err = nnet_linear_dev(c, l, dout, din, n))) return err;
  if ((err = st.out(out, dout, ob_, on_device))) return err;
  return st.finish(on_device);
}
extern "C" int opusb200_nnet_compute_activation(OpusB200Ctx *c, float *out, const float *in, int n, int N, int activation, int on_device) {
  if (!c || !out || !in || n <= 0 || N <= 0 || activation < 0 || activation > 6) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  NnetStage st(c);
  int err = OPUSB200_OK;
  const size_t b = sizeof(float) * (size_t)n * N;
  const float *din = (const float *)st.in(in, b, on_device, true, &err);
  float *dout = (on_device || in != out) ? (float *)st.in(out, b, on_device, false, &err) : (float *)din;
  if (err) return err;
  if ((err = nnet_act_dev(c, dout, din, n, N, activation))) return err;
  if ((err = st.out(out, dout, b, on_device))) return err;
  return st.finish(on_device);
}
extern "C" int opusb200_nnet_compute_dense(OpusB200Ctx *c, const OpusB200Linear *l, float *out, const float *in, int n, int activation, int on_device) {
  if (!c || !l || !out || !in || n <= 0 || in == out || activation < 0 || activation > 6) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  NnetStage st(c);
  int err = OPUSB200_OK;
  const size_t ib = sizeof(float) * (size_t)n * l->L.M, ob_ = sizeof(float) * (size_t)n * l->L.N;
  const float *din = (const float *)st.in(in, ib, on_device, true, &err);
  float *dout = (float *)st.in(out, ob_, on_device, false, &err);
  if (err) return err;
  bool fused = false;
  if ((err = nnet_linear_dev(c, l, dout, din, n, activation, nullptr, &fused))) return err;
  if (!fused && (err = nnet_act_dev(c, dout, dout, n, l->L.N, activation))) return err;
  if ((err = st.out(out, dout, ob_, on_device))) return err;
  return st.finish(on_device);
}
extern "C" int opusb200_nnet_compute_gru(OpusB200Ctx *c, const OpusB200Linear *wi, const OpusB200Linear *wr, float *state, const float *in, int n,
                                         int on_device) {
  if (!c || !wi || !wr || !state || !in || n <= 0 || in == state) return OPUSB200_BAD_ARG;
  const int N = wr->L.M;
  if (wr->L.N != 3 * N || wi->L.N != 3 * N) return OPUSB200_BAD_ARG;   // dnn/nnet.c:86-87
  OB_CUDA(cudaSetDevice(c->device));
  NnetStage st(c);
  int err = OPUSB200_OK;
  const size_t ib = sizeof(float) * (size_t)n * wi->L.M, sb = sizeof(float) * (size_t)n * N;
  const float *din = (const float *)st.in(in, ib, on_device, true, &err);
  float *dstate = (float *)st.in(state, sb, on_device, true, &err);
  if (err) return err;
  void *tmp;
  if (nnet_scratch(c, 1, sizeof(float) * (size_t)n * 6 * N, &tmp)) return OPUSB200_ALLOC_FAIL;
  float *zrh = (float *)tmp, *recur = zrh + (size_t)n * 3 * N;
  if ((err = nnet_linear_dev(c, wi, zrh, din, n))) return err;
  if ((err = nnet_linear_dev(c, wr, recur, dstate, n))) return err;
  NnetGruParams g;
  g.state = dstate; g.zrh = zrh; g.recur = recur; g.n = n; g.N = N;
  ob_nnet_gru<<<nblk((long)n * N, 256), 256, 0, c->stream>>>(g);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if ((err = st.out(state, dstate, sb, on_device))) return err;
  return st.finish(on_device);
}
extern "C" int opusb200_nnet_compute_glu(OpusB200Ctx *c, const OpusB200Linear *l, float *out, const float *in, int n, int on_device) {
  if (!c || !l || !out || !in || n <= 0 || l->L.M != l->L.N) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  NnetStage st(c);
  int err = OPUSB200_OK;
  const size_t b = sizeof(float) * (size_t)n * l->L.N;
  const float *din = (const float *)st.in(in, b, on_device, true, &err);
  float *dout = (on_device || in != out) ? (float *)st.in(out, b, on_device, false, &err) : (float *)din;
  if (err) return err;
  if (l->wt_umma && getenv("OPUSB200_NNET_FUSE") != nullptr) {   // gate fused into the tensor-core kernel's epilogue (see nnet_linear_dev)
    bool fused = false;
    if ((err = nnet_linear_dev(c, l, dout, din, n, NN_LINEAR, din, &fused))) return err;
    if ((err = st.out(out, dout, b, on_device))) return err;
    return st.finish(on_device);
  }
  void *tmp;
  if (nnet_scratch(c, 1, b, &tmp)) return OPUSB200_ALLOC_FAIL;
  if ((err = nnet_linear_dev(c, l, (float *)tmp, din, n))) return err;
  NnetGluParams g;
  g.out = dout; g.in = din; g.act2 = (const float *)tmp; g.total = (long)n * l->L.N;
  ob_nnet_glu<<<nblk(g.total, 256), 256, 0, c->stream>>>(g);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if ((err = st.out(out, dout, b, on_device))) return err;
  return st.finish(on_device);
}
// compute_generic_conv1d (dilation == 1) and compute_generic_conv1d_dilation (dnn/nnet.c:124-153).
// mem: [n][nb_inputs - input_size] for dilation 1, [n][input_size*dilation*(ksize-1)] otherwise.
extern "C" int opusb200_nnet_compute_conv1d(OpusB200Ctx *c, const OpusB200Linear *l, float *out, float *mem, const float *in, int n, int input_size,
                                            int dilation, int activation, int on_device) {
  if (!c || !l || !out || !in || n <= 0 || input_size <= 0 || dilation < 1 || l->L.M % input_size || in == out || activation < 0 || activation > 6)
    return OPUSB200_BAD_ARG;
  const int M = l->L.M, ksize = M / input_size, hist = M - input_size;
  const int mem_len = dilation == 1 ? hist : input_size * dilation * (ksize - 1);
  if (mem_len && !mem) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  NnetStage st(c);
  int err = OPUSB200_OK;
  const size_t ib = sizeof(float) * (size_t)n * input_size, ob_ = sizeof(float) * (size_t)n * l->L.N, mb = sizeof(float) * (size_t)n * mem_len;
  const float *din = (const float *)st.in(in, ib, on_device, true, &err);
  float *dout = (float *)st.in(out, ob_, on_device, false, &err);
  float *dmem = mem_len ? (float *)st.in(mem, mb, on_device, true, &err) : nullptr;
  if (err) return err;
  void *tmp;
  if (nnet_scratch(c, 1, sizeof(float) * (size_t)n * M, &tmp)) return OPUSB200_ALLOC_FAIL;
  NnetConvParams p;
  p.tmp = (float *)tmp; p.mem = dmem; p.in = din; p.n = n; p.nb_inputs = M; p.input_size = input_size; p.dilation = dilation; p.phase = 0;
  ob_nnet_conv1d<<<nblk((long)n * M, 256), 256, 0, c->stream>>>(p);
  c->launches++;
  bool fused = false;
  if ((err = nnet_linear_dev(c, l, dout, (const float *)tmp, n, activation, nullptr, &fused))) return err;
  if (!fused && (err = nnet_act_dev(c, dout, dout, n, l->L.N, activation))) return err;
  if (mem_len) {
    p.phase = 1;
    ob_nnet_conv1d<<<nblk(dilation == 1 ? (long)n * hist : (long)n * input_size, 256), 256, 0, c->stream>>>(p);
    c->launches++;
  }
  OB_CUDA(cudaGetLastError());
  if ((err = st.out(out, dout, ob_, on_device))) return err;
  if (mem_len && (err = st.out(mem, dmem, mb, on_device))) return err;
  return st.finish(on_device);
}

extern "C" int opusb200_nnet_conv2d_create(OpusB200Ctx *c, const OpusB200Conv2dLayer *h, OpusB200Conv2d **out) {
  if (!c || !h || !out || !h->float_weights || h->in_channels <= 0 || h->out_channels <= 0 || h->ktime <= 0 || h->kheight <= 0) return OPUSB200_BAD_ARG;
  *out = nullptr;
  OB_CUDA(cudaSetDevice(c->device));
  OpusB200Conv2d *q = new (std::nothrow) OpusB200Conv2d;
  if (!q) return OPUSB200_ALLOC_FAIL;
  q->ctx = c; q->in_ch = h->in_channels; q->out_ch = h->out_channels; q->ktime = h->ktime; q->kheight = h->kheight; q->w = q->bias = nullptr;
  const size_t wb = sizeof(float) * (size_t)q->in_ch * q->out_ch * q->ktime * q->kheight;
  if (cudaMalloc(&q->w, wb) != cudaSuccess || (h->bias && cudaMalloc(&q->bias, sizeof(float) * q->out_ch) != cudaSuccess)) {
    cudaFree(q->w); delete q; return OPUSB200_ALLOC_FAIL;
  }
  cudaMemcpyAsync(q->w, h->float_weights, wb, cudaMemcpyHostToDevice, c->stream);
  if (h->bias) cudaMemcpyAsync(q->bias, h->bias, sizeof(float) * q->out_ch, cudaMemcpyHostToDevice, c->stream);
  OB_CUDA(cudaStreamSynchronize(c->stream));
  *out = q;
  return OPUSB200_OK;
}
extern "C" void opusb200_nnet_conv2d_destroy(OpusB200Conv2d *q) {
  if (!q) return;
  cudaSetDevice(q->ctx->device);
  cudaStreamSynchronize(q->ctx->stream);
  cudaFree(q->w); cudaFree(q->bias);
  delete q;
}
// compute_conv2d (dnn/nnet_arch.h:217-243): in [n][in_ch*(height+kheight-1)], mem [n][(ktime-1)*that], out [n][out_ch*hstride]
extern "C" int opusb200_nnet_compute_conv2d(OpusB200Ctx *c, const OpusB200Conv2d *q, float *out, float *mem, const float *in, int n, int height,
                                            int hstride, int activation, int on_device) {
  if (!c || !q || !out || !in || n <= 0 || height <= 0 || hstride < height || in == out || activation < 0 || activation > 6 ||
      activation == NN_SOFTMAX)
    return OPUSB200_BAD_ARG;
  const int ts = q->in_ch * (height + q->kheight - 1);
  if (q->ktime * ts > 8192 || (q->ktime > 1 && !mem)) return OPUSB200_BAD_ARG;   // MAX_CONV2D_INPUTS
  OB_CUDA(cudaSetDevice(c->device));
  NnetStage st(c);
  int err = OPUSB200_OK;
  const size_t ib = sizeof(float) * (size_t)n * ts, ob_ = sizeof(float) * (size_t)n * q->out_ch * hstride, mb = sizeof(float) * (size_t)n * (q->ktime - 1) * ts;
  const float *din = (const float *)st.in(in, ib, on_device, true, &err);
  float *dout = (float *)st.in(out, ob_, on_device, true, &err);   // gaps between rows (hstride > height) keep the caller's values
  float *dmem = q->ktime > 1 ? (float *)st.in(mem, mb, on_device, true, &err) : nullptr;
  if (err) return err;
  NnetConv2dParams p;
  p.w = q->w; p.bias = q->bias; p.out = dout; p.mem = dmem; p.in = din; p.n = n; p.in_ch = q->in_ch; p.out_ch = q->out_ch; p.ktime = q->ktime;
  p.kheight = q->kheight; p.height = height; p.hstride = hstride; p.act = activation;
  ob_nnet_conv2d<<<nblk((long)n * q->out_ch * height, 256), 256, 0, c->stream>>>(p);
  c->launches++;
  if (q->ktime > 1) {
    ob_nnet_conv2d_shift<<<nblk((long)n * ts, 256), 256, 0, c->stream>>>(p);
    c->launches++;
  }
  OB_CUDA(cudaGetLastError());
  if ((err = st.out(out, dout, ob_, on_device))) return err;
  if (q->ktime > 1 && (err = st.out(mem, dmem, mb, on_device))) return err;
  return st.finish(on_device);
}
