// nnet.cuh — dnn/ inference primitives with a leading batch dimension: the reference runs one
// GEMV per stream and layer (dnn/nnet_arch.h:122 compute_linear); with n streams sharing the weights it is a GEMM
//   out[n][N] = in[n][M] · W[M][N].
//   * int8 layers (cgemv8x4, dnn/vec_avx.h:831 / vec.h:283): exact integer arithmetic, so they go to the TENSOR CORES —
//     mma.sync m16n8k32 (u8|s8 x s8 -> s32).  The reference's own 8-row x 4-column weight blocking is, word for word,
//     the B-fragment layout of that instruction (one 32-bit word = 4 consecutive k of one output row), so the weight
//     blob is used as stored.  Results are bit-identical to the reference (int32 dot, then float scale/bias).
//     At FARGAN sizes (M,N of a few hundred) with fp32 activations in HBM the layer is bandwidth-bound, which is why this
//     is mma.sync straight from L2/global fragments and not a tcgen05/TMEM pipeline.
//   * float layers (sgemv, vec.h:110): fp32 FMUL+FADD in the reference's column order (bit-exact), shared-memory tiled.
//   * block-sparse layers (sparse_sgemv8x4 / sparse_cgemv8x4): one thread per (stream, 8-row block), dp4a for int8.
//   * activations follow the deterministic generic branch of dnn/vec.h:317-382 (the x86 build's _mm_rcp_ps variants
//     are CPU-specific; they agree within the reference's stated 3e-4).
#pragma once
#include "simt.cuh"

namespace ob {

enum { NN_LINEAR = 0, NN_SIGMOID = 1, NN_TANH = 2, NN_RELU = 3, NN_SOFTMAX = 4, NN_SWISH = 5, NN_EXP = 6 };

struct NnetLinear {     // device-side LinearLayer (dnn/nnet.h:66-76)
  const float *bias;    // already the effective one (subias for int8 layers under the x86 convention)
  const int8_t *w8;     // 8x4-blocked int8 weights, or null
  const float *wf;      // float weights (dense: [j*N+i]; sparse: 8x4 blocks, column-major inside), or null
  const int *rb_off;    // sparse: [N/8+1] first block of each 8-row block
  const int *pos;       // sparse: [nblk] first input column of each block
  const float *diag;
  const float *scale;
  int M, N, su;         // su: unsigned activations 127+q (dnn/vec_avx.h:41 USE_SU_BIAS)
};

// ---- activations (dnn/vec.h:317-382, generic branch) ----------------------------------------------------------------
OB_DEV float nn_tanh(float x) {
  const float N0 = 952.52801514f, N1 = 96.39235687f, N2 = 0.60863042f;
  const float D0 = 952.72399902f, D1 = 413.36801147f, D2 = 11.88600922f;
  const float X2 = f_mul(x, x);
  float num = f_add(f_mul(f_add(f_mul(N2, X2), N1), X2), N0);
  const float den = f_add(f_mul(f_add(f_mul(D2, X2), D1), X2), D0);
  num = f_div(f_mul(num, x), den);
  return num < -1.f ? -1.f : (num > 1.f ? 1.f : num);
}
OB_DEV float nn_sigmoid(float x) { return f_add(.5f, f_mul(.5f, nn_tanh(f_mul(.5f, x)))); }
OB_DEV float nn_exp(float x0) {
  const float x = f_mul(x0, 1.44269504f);
  const int integer = (int)floorf(x);
  if (integer < -50) return 0.f;
  const float frac = f_add(x, -(float)integer);
  float p = f_add(0.22606716f, f_mul(0.078024523f, frac));
  p = f_add(0.69583354f, f_mul(frac, p));
  p = f_add(0.99992522f, f_mul(frac, p));
  union { float f; uint32_t i; } res;
  res.f = p;
  res.i = (res.i + ((uint32_t)integer << 23)) & 0x7fffffffu;
  return res.f;
}
OB_DEV float nn_act(float x, int act) {
  switch (act) {
    case NN_SIGMOID: return nn_sigmoid(x);
    case NN_TANH: return nn_tanh(x);
    case NN_RELU: return x < 0 ? 0.f : x;
    case NN_SWISH: return f_mul(x, nn_sigmoid(x));
    case NN_EXP: return nn_exp(x);
    default: return x;
  }
}
struct NnetActParams { float *out; const float *in; int n, N, act; };
// elementwise activations: one thread per element; softmax (compute_activation's normalised form, nnet_arch.h:99-108):
// one thread per stream, sums in index order
OB_DEV void nnet_activation_thread(const NnetActParams &p, long t) {
  if (p.act == NN_SOFTMAX) {
    if (t >= p.n) return;
    const float *in = p.in + (size_t)t * p.N;
    float *out = p.out + (size_t)t * p.N;
    float sum = 0;
#pragma unroll 1
    for (int i = 0; i < p.N; i++) {
      const float e = nn_exp(in[i]);
      out[i] = e;
      sum = f_add(sum, e);
    }
    sum = (float)(1.0 / ((double)sum + 1e-30));
#pragma unroll 1
    for (int i = 0; i < p.N; i++) out[i] = f_mul(sum, out[i]);
  } else if (t < (long)p.n * p.N) p.out[t] = nn_act(p.in[t], p.act);
}

// ---- shared epilogue of compute_linear (nnet_arch.h:141-153) --------------------------------------------------------
OB_DEV float nnet_epilogue(const NnetLinear &L, float y, const float *in_row, int i) {
  if (L.bias) y = f_add(y, L.bias[i]);
  if (L.diag) {  // GRU recurrent weights only: N == 3*M
    const int m = i >= 2 * L.M ? i - 2 * L.M : (i >= L.M ? i - L.M : i);
    y = f_add(y, f_mul(L.diag[i], in_row[m]));
  }
  return y;
}

// ---- activation quantisation (vec.h:253/288 and vec_avx.h:383) -------------------------------------------------------
struct NnetQuantParams { int8_t *xq; const float *in; int n, M, Mp, su; int in_stride; };   // in_stride: floats between rows (>= M)
OB_DEV void nnet_quant_thread(const NnetQuantParams &p, long t) {  // one thread per 4 outputs (one word)
  const int wpr = p.Mp >> 2;
  if (t >= (long)p.n * wpr) return;
  const int s = (int)(t / wpr), j0 = (int)(t - (long)s * wpr) * 4;
  uint32_t word = 0;
#pragma unroll
  for (int b = 0; b < 4; b++) {
    const int j = j0 + b;
    int q = 0;
    if (j < p.M) {
      q = (int)floor(0.5 + (double)f_mul(127.f, p.in[(size_t)s * p.in_stride + j]));
      if (p.su) q += 127;
    }
    word |= ((uint32_t)q & 0xffu) << (8 * b);
  }
  ((uint32_t *)p.xq)[t] = word;
}

// ---- tensor-core int8 GEMM ----------------------------------------------------------------------------------------
// D(16x8,s32) += A(16x32, row-major u8|s8) * B(32x8, "col": each output's k contiguous, s8)
template <bool UNSIGNED_A>
OB_DEV void mma_m16n8k32_i8(int32_t (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
#if defined(__CUDACC__)
  if (UNSIGNED_A)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  else
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
#else
  // host emulator: rebuild the fragments' rows / columns with shuffles (PTX ISA "Matrix Fragments for mma.m16n8k32")
  const int lane = lane_id(), gid = lane >> 2, tig = lane & 3;
  int32_t acc[4] = {0, 0, 0, 0};
  for (int q = 0; q < 4; q++) {      // k-words q (k = 4q..4q+3) and q+4 (k = 16+4q..)
    uint32_t arow[2][2], bcol[2][2];  // [row gid / gid+8][k half], [col 2*tig / 2*tig+1][k half]
    arow[0][0] = shfl(a[0], gid * 4 + q); arow[1][0] = shfl(a[1], gid * 4 + q);
    arow[0][1] = shfl(a[2], gid * 4 + q); arow[1][1] = shfl(a[3], gid * 4 + q);
    for (int cc = 0; cc < 2; cc++) {
      bcol[cc][0] = shfl(b[0], (2 * tig + cc) * 4 + q);
      bcol[cc][1] = shfl(b[1], (2 * tig + cc) * 4 + q);
    }
    for (int rr = 0; rr < 2; rr++)
      for (int cc = 0; cc < 2; cc++)
        for (int h = 0; h < 2; h++)
          for (int by = 0; by < 4; by++) {
            const int av = UNSIGNED_A ? (int)((arow[rr][h] >> (8 * by)) & 0xff) : (int)(int8_t)((arow[rr][h] >> (8 * by)) & 0xff);
            const int bv = (int)(int8_t)((bcol[cc][h] >> (8 * by)) & 0xff);
            acc[rr * 2 + cc] += av * bv;
          }
  }
  for (int i = 0; i < 4; i++) d[i] += acc[i];
#endif
}

struct NnetLinearParams {
  NnetLinear L;
  float *out;         // [n][N]
  const float *in;    // [n][M]
  const int8_t *xq;   // [n][Mp] quantised activations (int8 layers)
  int n, Mp;
};

// Block = 4 warps; block tile = 32 streams x 256 outputs, warp w owns outputs [64w, 64w+64).
// A fragments come straight from the quantised activations (32 contiguous bytes per row and k-step), B fragments straight
// from the reference's blocked weights (256 contiguous bytes per 8 rows and k-step); both live in L2 after first touch.
template <bool UNSIGNED_A>
OB_DEV void nnet_linear_i8_mma_warp(const NnetLinearParams &p, int stream0, int out0) {
  const NnetLinear &L = p.L;
  const int lane = lane_id(), gid = lane >> 2, tig = lane & 3;
  const int CB = L.M >> 2, NB = L.N >> 3;      // column blocks of 4, row blocks of 8
  const int rb0 = out0 >> 3;
  const uint32_t *W = (const uint32_t *)L.w8;  // word (rb*CB + cb)*8 + r  = weights of row 8rb+r, columns 4cb..4cb+3
  const uint32_t *X = (const uint32_t *)p.xq;
  const int wpr = p.Mp >> 2;
  int32_t acc[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int nt = 0; nt < 8; nt++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[mt][nt][e] = 0;
  int rowA[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int s = stream0 + mt * 16 + h * 8 + gid;
      rowA[mt][h] = (s < p.n ? s : p.n - 1) * wpr;
    }
  // software pipeline: the fragments of k-step s+1 are in flight while the 16 MMAs of step s issue
  uint32_t a[2][2][4], b[2][8][2];
  auto load_frags = [&](int k0, uint32_t (&af)[2][4], uint32_t (&bf)[8][2]) {
    const int cb = (k0 >> 2) + tig;
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
      af[mt][0] = X[rowA[mt][0] + cb];
      af[mt][1] = X[rowA[mt][1] + cb];
      af[mt][2] = X[rowA[mt][0] + cb + 4];
      af[mt][3] = X[rowA[mt][1] + cb + 4];
    }
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      const int rb = rb0 + nt;
      bf[nt][0] = bf[nt][1] = 0u;
      if (rb < NB) {
        if (cb < CB) bf[nt][0] = W[((size_t)rb * CB + cb) * 8 + gid];
        if (cb + 4 < CB) bf[nt][1] = W[((size_t)rb * CB + cb + 4) * 8 + gid];
      }
    }
  };
  load_frags(0, a[0], b[0]);
#pragma unroll 1
  for (int k0 = 0; k0 < p.Mp; k0 += 64) {
#pragma unroll
    for (int st = 0; st < 2; st++) {
      const int k = k0 + st * 32;
      if (k < p.Mp) {
        if (k + 32 < p.Mp) load_frags(k + 32, a[st ^ 1], b[st ^ 1]);
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
          mma_m16n8k32_i8<UNSIGNED_A>(acc[0][nt], a[st][0], b[st][nt]);
          mma_m16n8k32_i8<UNSIGNED_A>(acc[1][nt], a[st][1], b[st][nt]);
        }
      }
    }
  }
  // epilogue: (float)acc * scale (+bias) (+diag.in), vec.h:312 / nnet_arch.h:141-153
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int s = stream0 + mt * 16 + h * 8 + gid;
      if (s >= p.n) continue;
      const float *in_row = p.in + (size_t)s * L.M;
      float *out_row = p.out + (size_t)s * L.N;
#pragma unroll
      for (int nt = 0; nt < 8; nt++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int i = out0 + nt * 8 + tig * 2 + e;
          if (i < L.N) out_row[i] = nnet_epilogue(L, f_mul((float)acc[mt][nt][h * 2 + e], L.scale[i]), in_row, i);
        }
    }
}

// ---- int8 x (u8|s8) dot of 4, for the block-sparse layers ----
OB_DEV int32_t nn_dp4a(uint32_t w, uint32_t x, int32_t acc, bool unsigned_x) {
#if defined(__CUDACC__)
  int32_t d;
  if (unsigned_x) asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(x), "r"(acc));
  else asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(x), "r"(acc));
  return d;
#else
  for (int b = 0; b < 4; b++) {
    const int wv = (int)(int8_t)((w >> (8 * b)) & 0xff);
    const int xv = unsigned_x ? (int)((x >> (8 * b)) & 0xff) : (int)(int8_t)((x >> (8 * b)) & 0xff);
    acc += wv * xv;
  }
  return acc;
#endif
}
// one thread per (stream, 8-row block); sparse_cgemv8x4 (vec.h:249 / vec_avx.h:780)
OB_DEV void nnet_linear_i8_sparse_thread(const NnetLinearParams &p, long t) {
  const NnetLinear &L = p.L;
  const int NB = L.N >> 3;
  if (t >= (long)p.n * NB) return;
  const int s = (int)(t / NB), rb = (int)(t - (long)s * NB);
  const uint32_t *W = (const uint32_t *)L.w8;
  const uint32_t *X = (const uint32_t *)(p.xq + (size_t)s * p.Mp);
  int32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
  for (int blk = L.rb_off[rb]; blk < L.rb_off[rb + 1]; blk++) {
    const uint32_t x = X[L.pos[blk] >> 2];
#pragma unroll
    for (int r = 0; r < 8; r++) acc[r] = nn_dp4a(W[(size_t)blk * 8 + r], x, acc[r], L.su != 0);
  }
  const float *in_row = p.in + (size_t)s * L.M;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int i = rb * 8 + r;
    p.out[(size_t)s * L.N + i] = nnet_epilogue(L, f_mul((float)acc[r], L.scale[i]), in_row, i);
  }
}
// sparse_sgemv8x4 (vec.h:124): same mapping, fp32 in the reference's order (column by column inside a block)
OB_DEV void nnet_linear_f32_sparse_thread(const NnetLinearParams &p, long t) {
  const NnetLinear &L = p.L;
  const int NB = L.N >> 3;
  if (t >= (long)p.n * NB) return;
  const int s = (int)(t / NB), rb = (int)(t - (long)s * NB);
  const float *in_row = p.in + (size_t)s * L.M;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
  for (int blk = L.rb_off[rb]; blk < L.rb_off[rb + 1]; blk++) {
    const float *w = L.wf + (size_t)blk * 32;
    const int pos = L.pos[blk];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const float x = in_row[pos + c];
#pragma unroll
      for (int r = 0; r < 8; r++) acc[r] = f_add(acc[r], f_mul(w[c * 8 + r], x));
    }
  }
#pragma unroll
  for (int r = 0; r < 8; r++) p.out[(size_t)s * L.N + rb * 8 + r] = nnet_epilogue(L, acc[r], in_row, rb * 8 + r);
}

// ---- dense fp32 (sgemv, vec.h:50-122): 64 streams x 64 outputs per 256-thread block, k in the reference's order ----
constexpr int kNnF32Tile = 64, kNnF32K = 16;
struct NnetF32Smem { float xs[kNnF32K][kNnF32Tile + 1]; float ws[kNnF32K][kNnF32Tile]; };
OB_DEV void nnet_linear_f32_block(const NnetLinearParams &p, int stream0, int out0, NnetF32Smem &sm, int tid) {
  const NnetLinear &L = p.L;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int b = 0; b < 4; b++)
#pragma unroll
    for (int a = 0; a < 4; a++) acc[b][a] = 0.f;
#pragma unroll 1
  for (int j0 = 0; j0 < L.M; j0 += kNnF32K) {
    // stage x[64 streams][16 k] (transposed) and w[16 k][64 outputs]
#pragma unroll
    for (int e = tid; e < kNnF32Tile * kNnF32K; e += 256) {
      const int sx = e >> 4, kx = e & 15;
      const int s = stream0 + sx, j = j0 + kx;
      sm.xs[kx][sx] = (s < p.n && j < L.M) ? p.in[(size_t)s * L.M + j] : 0.f;
      const int kw = e >> 6, iw = e & 63;
      const int jw = j0 + kw, i = out0 + iw;
      sm.ws[kw][iw] = (jw < L.M && i < L.N) ? L.wf[(size_t)jw * L.N + i] : 0.f;
    }
    bsync();
    const int kmax = L.M - j0 < kNnF32K ? L.M - j0 : kNnF32K;
#pragma unroll 4
    for (int k = 0; k < kmax; k++) {
      float w[4], x[4];
#pragma unroll
      for (int a = 0; a < 4; a++) w[a] = sm.ws[k][tx + 16 * a];
#pragma unroll
      for (int b = 0; b < 4; b++) x[b] = sm.xs[k][ty + 16 * b];
#pragma unroll
      for (int b = 0; b < 4; b++)
#pragma unroll
        for (int a = 0; a < 4; a++) acc[b][a] = f_add(acc[b][a], f_mul(w[a], x[b]));
    }
    bsync();
  }
#pragma unroll
  for (int b = 0; b < 4; b++) {
    const int s = stream0 + ty + 16 * b;
    if (s >= p.n) continue;
    const float *in_row = p.in + (size_t)s * L.M;
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const int i = out0 + tx + 16 * a;
      if (i < L.N) p.out[(size_t)s * L.N + i] = nnet_epilogue(L, acc[b][a], in_row, i);
    }
  }
}
// a layer without weights: OPUS_CLEAR + bias (+diag) (nnet_arch.h:140)
OB_DEV void nnet_linear_noweights_thread(const NnetLinearParams &p, long t) {
  if (t >= (long)p.n * p.L.N) return;
  const int s = (int)(t / p.L.N), i = (int)(t - (long)s * p.L.N);
  p.out[t] = nnet_epilogue(p.L, 0.f, p.in + (size_t)s * p.L.M, i);
}

// ---- GRU gate update (compute_generic_gru, dnn/nnet.c:76-105) after the two linears ----
struct NnetGruParams { float *state; const float *zrh; const float *recur; int n, N; };
OB_DEV void nnet_gru_thread(const NnetGruParams &p, long t) {
  if (t >= (long)p.n * p.N) return;
  const int s = (int)(t / p.N), i = (int)(t - (long)s * p.N), N = p.N;
  const float *zrh = p.zrh + (size_t)s * 3 * N, *rec = p.recur + (size_t)s * 3 * N;
  const float z = nn_sigmoid(f_add(zrh[i], rec[i]));
  const float r = nn_sigmoid(f_add(zrh[N + i], rec[N + i]));
  const float h = nn_tanh(f_add(zrh[2 * N + i], f_mul(rec[2 * N + i], r)));
  const float st = p.state[t];
  p.state[t] = f_add(f_mul(z, st), f_mul(f_add(1.f, -z), h));
}
// ---- GLU (dnn/nnet.c:107-120): out = in * sigmoid(linear(in)); act2 holds the linear output ----
struct NnetGluParams { float *out; const float *in; const float *act2; long total; };
OB_DEV void nnet_glu_thread(const NnetGluParams &p, long t) {
  if (t < p.total) p.out[t] = f_mul(p.in[t], nn_sigmoid(p.act2[t]));
}
// ---- conv1d window assembly (dnn/nnet.c:124-153): tmp = [history | input]; afterwards the history slides ----
struct NnetConvParams { float *tmp; float *mem; const float *in; int n, nb_inputs, input_size, dilation, phase; };
OB_DEV void nnet_conv1d_thread(const NnetConvParams &p, long t) {
  const int ksize = p.nb_inputs / p.input_size, hist = p.nb_inputs - p.input_size;
  const int mem_len = p.dilation == 1 ? hist : p.input_size * p.dilation * (ksize - 1);
  if (p.phase == 0) {  // gather
    if (t >= (long)p.n * p.nb_inputs) return;
    const int s = (int)(t / p.nb_inputs), j = (int)(t - (long)s * p.nb_inputs);
    float v;
    if (j >= hist) v = p.in[(size_t)s * p.input_size + (j - hist)];
    else if (p.dilation == 1) v = p.mem[(size_t)s * mem_len + j];
    else v = p.mem[(size_t)s * mem_len + (j / p.input_size) * p.input_size * p.dilation + j % p.input_size];
    p.tmp[t] = v;
  } else if (p.dilation == 1) {  // history <- tmp[input_size ..] (tmp is complete, mem is no longer read)
    if (t >= (long)p.n * hist) return;
    const int s = (int)(t / hist), j = (int)(t - (long)s * hist);
    p.mem[(size_t)s * mem_len + j] = p.tmp[(size_t)s * p.nb_inputs + p.input_size + j];
  } else {                       // dilated history slides by one frame: one thread per (stream, channel) walks its column
    if (t >= (long)p.n * p.input_size) return;
    const int s = (int)(t / p.input_size), c = (int)(t - (long)s * p.input_size);
    float *mem = p.mem + (size_t)s * mem_len;
#pragma unroll 1
    for (int q = c; q < mem_len - p.input_size; q += p.input_size) mem[q] = mem[q + p.input_size];
    mem[mem_len - p.input_size + c] = p.in[(size_t)s * p.input_size + c];
  }
}

// ---- conv2d (compute_conv2d_c, nnet_arch.h:163-243): one thread per (stream, out channel, j) ----
struct NnetConv2dParams {
  const float *w, *bias;
  float *out, *mem;      // out [n][out_ch*hstride], mem [n][(ktime-1)*time_stride]
  const float *in;       // [n][time_stride]
  int n, in_ch, out_ch, ktime, kheight, height, hstride, act;
};
OB_DEV float nnet_conv2d_in(const NnetConv2dParams &p, int s, int t, int off) {  // in_buf[t*time_stride + off]
  const int ts = p.in_ch * (p.height + p.kheight - 1);
  return t < p.ktime - 1 ? p.mem[(size_t)s * (p.ktime - 1) * ts + t * ts + off] : p.in[(size_t)s * ts + off];
}
OB_DEV void nnet_conv2d_thread(const NnetConv2dParams &p, long tt) {
  const int per = p.out_ch * p.height;
  if (tt >= (long)p.n * per) return;
  const int s = (int)(tt / per), rem = (int)(tt - (long)s * per), i = rem / p.height, j = rem - i * p.height;
  const int in_stride = p.height + p.kheight - 1, kt = p.ktime, kh = p.kheight;
  const float *w = p.w + (size_t)i * p.in_ch * kt * kh;
  float y = 0.f;
#pragma unroll 1
  for (int m = 0; m < p.in_ch; m++) {
    if (kt == 3 && kh == 3) {  // conv2d_3x3_float: the nine products are summed left to right, then added
      float sum = f_mul(w[m * 9 + 0], nnet_conv2d_in(p, s, 0, m * in_stride + j));
#pragma unroll
      for (int q = 1; q < 9; q++) sum = f_add(sum, f_mul(w[m * 9 + q], nnet_conv2d_in(p, s, q / 3, m * in_stride + j + q % 3)));
      y = f_add(y, sum);
    } else {
#pragma unroll 1
      for (int t = 0; t < kt; t++)
#pragma unroll 1
        for (int h = 0; h < kh; h++) y = f_add(y, f_mul(w[m * kt * kh + t * kh + h], nnet_conv2d_in(p, s, t, m * in_stride + j + h)));
    }
  }
  if (p.bias) y = f_add(y, p.bias[i]);
  p.out[(size_t)s * p.out_ch * p.hstride + i * p.hstride + j] = nn_act(y, p.act);
}

// history update of compute_conv2d_c (nnet_arch.h:229-231): one thread per (stream, offset) walks its own column
OB_DEV void nnet_conv2d_shift_thread(const NnetConv2dParams &p, long tt) {
  const int ts = p.in_ch * (p.height + p.kheight - 1);
  if (tt >= (long)p.n * ts || p.ktime < 2) return;
  const int s = (int)(tt / ts), off = (int)(tt - (long)s * ts);
  float *mem = p.mem + (size_t)s * (p.ktime - 1) * ts;
#pragma unroll 1
  for (int t = 0; t < p.ktime - 2; t++) mem[t * ts + off] = mem[(t + 1) * ts + off];
  mem[(p.ktime - 2) * ts + off] = p.in[(size_t)s * ts + off];
}

}  // namespace ob
