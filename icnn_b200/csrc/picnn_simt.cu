// K1 (baseline): fully-connected PICNN energy f and df/dy for a whole minibatch, FP32 FFMA.
//
// Restates on the GPU what the reference evaluates through TensorFlow:
//   forward   multi-label-cls/icnn_ebundle.py:349-387, RL/src/icnn.py:356-404
//   gradient  tf.gradients(E_, y_)  (multi-label-cls/icnn_ebundle.py:146)
//
//   a_i = (z_{i-1} o cz_i) Wz_i + (y o cy_i) Wy_i + d_i ,  z_i = act(a_i) (i < L),  f = a_L
//   delta_L = 1 ;  g += cy_i o (delta_i Wy_i^T) ;  delta_{i-1} = act'(a_{i-1}) o cz_i o (delta_i Wz_i^T)
//
// Each hidden layer is one gated GEMM against Wcat_i = [Wz_i ; Wy_i] (forward: K-concatenated,
// backward: the same buffer read transposed), with the per-sample gates applied while the A tile
// is staged and bias/activation/act'-mask/g-accumulate fused into the epilogue.  The width-1
// output layer and its backward seed are a warp-per-sample kernel (shuffle reduction).
// This FP32 kernel is the accuracy anchor for the tcgen05 path (picnn_tc.cu).
#include "common.cuh"

#include <cooperative_groups.h>

#include <cstdlib>

namespace cg = cooperative_groups;

namespace icnn {

struct GemmArgs {
  int M, N, K0, K1;
  const float* A0; const float* G0; int lda0;
  const float* A1; const float* G1; int lda1;
  float a1_scale, a1_shift;
  const float* W; int ldw;
  // forward epilogue
  const float* D; float* Z; float alpha;
  // backward epilogue
  int N0; const float* Zprev; const float* Cz; float* dprev;
  const float* Cy; float* g; long long g_row_stride; const int* perm; const int* count; int KS; int n;
  float g_scale;
  const int* skip_if_zero;
};

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

__device__ __forceinline__ float* g_row_ptr(const GemmArgs& a, int m) {
  if (a.perm == nullptr) return a.g + (long long)m * a.g_row_stride;
  int slot = a.perm[(long long)m * a.KS + a.count[m]];
  return a.g + ((long long)m * a.KS + slot) * a.n;
}

// One output element of the fused epilogue (shared by the register and the split-K paths).
template <int MODE>
__device__ __forceinline__ void epilogue_elem(const GemmArgs& a, int m, int nn, float acc, float* grow) {
  if (MODE == 0) {
    const float v = acc + a.D[(long long)m * a.N + nn];
    a.Z[(long long)m * a.N + nn] = v > 0.f ? v : a.alpha * v;
  } else {
    if (nn < a.N0) {
      const long long idx = (long long)m * a.N0 + nn;
      const float da = a.Zprev[idx] > 0.f ? 1.f : a.alpha;
      a.dprev[idx] = da * a.Cz[idx] * acc;
    } else {
      const int e = nn - a.N0;
      grow[e] = fmaf(a.g_scale * a.Cy[(long long)m * a.n + e], acc, grow[e]);
    }
  }
}

// MODE 0: forward (W is [K, N]);  MODE 1: backward (W is [N, K], K = K0, no second segment)
// Split-K: gridDim.z = S CTAs of one thread-block cluster share an output tile; each reduces a
// K-slice, the partial tiles are summed through distributed shared memory (rank r owns BM/S rows
// of the tile for the reduction + epilogue).  S = 1 is the plain kernel.
template <int MODE>
__global__ void __launch_bounds__(256) gated_gemm_kernel(GemmArgs a) {
  if (a.skip_if_zero != nullptr && *a.skip_if_zero == 0) return;
  __shared__ __align__(16) float smem_f[2 * BK * (BM + PAD) + 2 * BK * (BN + PAD)];
  float (*As)[BK][BM + PAD] = reinterpret_cast<float (*)[BK][BM + PAD]>(smem_f);
  float (*Bs)[BK][BN + PAD] = reinterpret_cast<float (*)[BK][BN + PAD]>(smem_f + 2 * BK * (BM + PAD));
  const int t = threadIdx.x;
  const int S = gridDim.z;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int K = a.K0 + a.K1;
  const int ty = t / 16, tx = t % 16;

  // loader coordinates
  const int a_row = t / 4, a_k = (t % 4) * 4;  // A tile: 64 rows x 16 k
  const int b_k = t / 16, b_n = (t % 16) * 4;  // fwd W tile: 16 k x 64 n
  const int bt_n = t / 4, bt_k = (t % 4) * 4;  // bwd W tile: 64 n x 16 k

  float ra[4], rb[4];
  auto load_tiles = [&](int k0) {
    const int m = m0 + a_row;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = k0 + a_k + i;
      float v = 0.f;
      if (m < a.M && kk < K) {
        if (kk < a.K0) {
          v = a.A0[(long long)m * a.lda0 + kk];
          if (a.G0) v *= a.G0[(long long)m * a.lda0 + kk];
        } else {
          const int k1 = kk - a.K0;
          v = fmaf(a.a1_scale, a.A1[(long long)m * a.lda1 + k1], a.a1_shift);
          if (a.G1) v *= a.G1[(long long)m * a.lda1 + k1];
        }
      }
      ra[i] = v;
    }
    if (MODE == 0) {
      const int kk = k0 + b_k;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int nn = n0 + b_n + i;
        rb[i] = (kk < K && nn < a.N) ? a.W[(long long)kk * a.ldw + nn] : 0.f;
      }
    } else {
      const int nn = n0 + bt_n;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = k0 + bt_k + i;
        rb[i] = (kk < K && nn < a.N) ? a.W[(long long)nn * a.ldw + kk] : 0.f;
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) As[buf][a_k + i][a_row] = ra[i];
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[buf][b_k][b_n + i] = rb[i];
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[buf][bt_k + i][bt_n] = rb[i];
    }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk_all = (K + BK - 1) / BK;
  const int kt0 = (int)(((long long)nk_all * blockIdx.z) / S);
  const int nk = (int)(((long long)nk_all * (blockIdx.z + 1)) / S) - kt0;
  if (nk > 0) { load_tiles(kt0 * BK); store_tiles(0); }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt0 + kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  if (S == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m >= a.M) continue;
      float* grow = (MODE == 1) ? g_row_ptr(a, m) : nullptr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nn = n0 + tx * 4 + j;
        if (nn < a.N) epilogue_elem<MODE>(a, m, nn, acc[i][j], grow);
      }
    }
    return;
  }
  // ---- split-K: partial tile -> own shared memory -> DSMEM reduction ----
  cg::cluster_group cluster = cg::this_cluster();
  float (*Ps)[BN + 1] = reinterpret_cast<float (*)[BN + 1]>(smem_f);   // [BM][BN+1] floats fit
  static_assert(BM * (BN + 1) <= 2 * BK * (BM + PAD) + 2 * BK * (BN + PAD), "partial tile must fit");
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Ps[ty * 4 + i][tx * 4 + j] = acc[i][j];
  cluster.sync();
  const int rank = (int)cluster.block_rank();
  const int r_lo = (BM * rank) / S, r_hi = (BM * (rank + 1)) / S;
  for (int idx = t; idx < (r_hi - r_lo) * BN; idx += 256) {
    const int rr = r_lo + idx / BN, cc = idx % BN;
    float v = 0.f;
    for (int q = 0; q < S; ++q) {
      const float* rp = cluster.map_shared_rank(&Ps[rr][cc], q);
      v += *rp;
    }
    const int m = m0 + rr, nn = n0 + cc;
    if (m < a.M && nn < a.N) {
      float* grow = (MODE == 1 && nn >= a.N0) ? g_row_ptr(a, m) : nullptr;
      epilogue_elem<MODE>(a, m, nn, v, grow);
    }
  }
  cluster.sync();   // keep every CTA's partial tile alive until all ranks have read it
}

struct OutArgs {
  int M, S, n;                  // S = s_{L-1}
  const float* Z; const float* Cz;      // [M, S]
  const float* y; const float* Cy;      // [M, n]
  const float* D;                        // [M]
  const float* w;                        // [S + n]  = [wz_L ; wy_L]
  float in_scale, in_shift, g_scale, alpha;
  float* f; float* delta;                // [M], [M, S]
  float* delta_hi; float* delta_lo;      // optional TF32 hi/lo split of delta (tensor-core path)
  float* g; long long g_row_stride; const int* perm; const int* count; int KS;
  const int* skip_if_zero;
};

// TPS threads per sample (32 = a warp, 256 = a CTA for small batches where a warp per sample is
// latency-bound): f = d_L + (z o cz_L).wz_L + (y o cy_L).wy_L ; seeds delta_{L-1} and g.
template <int TPS>
__global__ void __launch_bounds__(256) out_layer_kernel(OutArgs a) {
  if (a.skip_if_zero != nullptr && *a.skip_if_zero == 0) return;
  __shared__ float red[8];
  const int m = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) / TPS);
  const int tl = threadIdx.x % TPS;
  const int lane = threadIdx.x & 31;
  const bool mv = m < a.M;
  float acc = 0.f;
  if (mv) {
    for (int j = tl; j < a.S; j += TPS) {
      const long long idx = (long long)m * a.S + j;
      const float z = a.Z[idx], c = a.Cz[idx] * a.w[j];
      acc = fmaf(z, c, acc);
      const float dl = (z > 0.f ? 1.f : a.alpha) * c;
      if (a.delta) a.delta[idx] = dl;
      if (a.delta_hi) {
        // TF32 hi/lo with round-to-nearest on both parts (see tf32_rn in picnn_tc.cu)
        const float h = __uint_as_float((__float_as_uint(dl) + 0x00001000u) & 0xFFFFE000u);
        a.delta_hi[idx] = h;
        a.delta_lo[idx] = __uint_as_float((__float_as_uint(dl - h) + 0x00001000u) & 0xFFFFE000u);
      }
    }
    float* grow;
    if (a.perm == nullptr) grow = a.g + (long long)m * a.g_row_stride;
    else grow = a.g + ((long long)m * a.KS + a.perm[(long long)m * a.KS + a.count[m]]) * a.n;
    for (int e = tl; e < a.n; e += TPS) {
      const long long idx = (long long)m * a.n + e;
      const float c = a.Cy[idx] * a.w[a.S + e];
      acc = fmaf(fmaf(a.in_scale, a.y[idx], a.in_shift), c, acc);
      grow[e] = a.g_scale * c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (TPS == 32) {
    if (mv && lane == 0) a.f[m] = acc + a.D[m];
  } else {   // TPS == 256: one sample per CTA, fixed-order sum of the 8 warp partials
    if (lane == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0 && mv) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w];
      a.f[m] = s + a.D[m];
    }
  }
}

__global__ void concat_rows_kernel(float* dst, const float* top, long long ntop, const float* bot,
                                   long long nbot) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < ntop) dst[i] = top[i];
  else if (i < ntop + nbot) dst[i] = bot[i - ntop];
}

// momentum GD update, multi-label-cls/icnn-back.py:122-128
__global__ void gd_update_kernel(float* y, float* v, const float* g, long long N, float lr, float mom) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float vp = v[i];
  const float vn = mom * vp - lr * g[i];
  y[i] = y[i] - mom * vp + (1.f + mom) * vn;
  v[i] = vn;
}

// Launch with a (1,1,S) thread-block cluster; S chosen so that the grid covers the chip at least
// ~2x while every CTA keeps >= 4 k-tiles.
template <int MODE>
static cudaError_t launch_gemm(const GemmArgs& a, cudaStream_t st) {
  const int gx = cdiv(a.N, BN), gy = cdiv(a.M, BM);
  const int nk = cdiv(a.K0 + a.K1, BK);
  int S = 1;
  while (S < 8 && gx * gy * S < 296 && nk / (S * 2) >= 4) S *= 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, gy, S);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = S;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, gated_gemm_kernel<MODE>, a);
}

// workspace layout: Z_0..Z_{L-1} [B, s_i], then two delta buffers [B, smax]
size_t picnn_simt_ws_floats(const icnn_picnn* h, int B, size_t* zoff, size_t* doff) {
  size_t off = 0;
  int smax = 0;
  for (int i = 0; i < h->L; ++i) {
    if (zoff) zoff[i] = off;
    off += (size_t)B * h->hidden[i];
    off = (off + 63) & ~(size_t)63;
    smax = h->hidden[i] > smax ? h->hidden[i] : smax;
  }
  if (doff) { doff[0] = off; doff[1] = off + (((size_t)B * smax + 63) & ~(size_t)63); }
  off += 2 * (((size_t)B * smax + 63) & ~(size_t)63);
  return off;
}

void out_layer_launch(const icnn_picnn* h, const icnn_gates* gt, const float* Zlast, const float* y32, float* f,
                      float* delta, float* delta_hi, float* delta_lo, float* g, long long g_row_stride,
                      const int* perm, const int* count, int KS, const int* skip, cudaStream_t st) {
  const int B = gt->B, L = h->L;
  OutArgs o{};
  o.M = B; o.S = h->hidden[L - 1]; o.n = h->n; o.Z = Zlast; o.Cz = gt->cz[L]; o.y = y32; o.Cy = gt->cy[L];
  o.D = gt->d[L]; o.w = h->Wcat[L]; o.in_scale = gt->in_scale; o.in_shift = gt->in_shift;
  o.g_scale = gt->g_scale; o.alpha = h->alpha; o.f = f; o.delta = delta; o.delta_hi = delta_hi; o.delta_lo = delta_lo;
  o.g = g; o.g_row_stride = g_row_stride; o.perm = perm; o.count = count; o.KS = KS; o.skip_if_zero = skip;
  if (B <= 2048) out_layer_kernel<256><<<B, 256, 0, st>>>(o);          // few samples: a CTA each
  else out_layer_kernel<32><<<cdiv(B * 32, 256), 256, 0, st>>>(o);
}

bool picnn_tc_supported(const icnn_picnn* h);
int picnn_tc_prepare_weights(icnn_picnn* h, cudaStream_t st);
void picnn_tc_free_weights(icnn_picnn* h);
void picnn_xpath_free(icnn_picnn* h);
size_t picnn_tc_ws_floats(const icnn_picnn* h, int B, size_t* aoff, size_t* doff);
int picnn_fg_tc(const icnn_picnn* h, const icnn_gates* gt, const float* y32, float* f, float* g,
                long long g_row_stride, const int* perm, const int* count, int KS, void* workspace,
                const int* skip, cudaStream_t st);

int picnn_fg_simt(const icnn_picnn* h, const icnn_gates* gt, const float* y32, float* f, float* g,
                  long long g_row_stride, const int* perm, const int* count, int KS, void* workspace,
                  const int* skip, cudaStream_t st) {
  const int B = gt->B, n = h->n, L = h->L;
  size_t zoff[ICNN_MAX_LAYERS], doff[2];
  picnn_simt_ws_floats(h, B, zoff, doff);
  float* ws = static_cast<float*>(workspace);
  float* Z[ICNN_MAX_LAYERS];
  for (int i = 0; i < L; ++i) Z[i] = ws + zoff[i];
  float* dl[2] = {ws + doff[0], ws + doff[1]};

  for (int i = 0; i < L; ++i) {  // forward hidden layers
    GemmArgs a{};
    a.M = B; a.N = h->hidden[i]; a.K0 = h->prev(i); a.K1 = n;
    a.A0 = i ? Z[i - 1] : nullptr; a.G0 = i ? gt->cz[i] : nullptr; a.lda0 = a.K0;
    a.A1 = y32; a.G1 = gt->cy[i]; a.lda1 = n; a.a1_scale = gt->in_scale; a.a1_shift = gt->in_shift;
    a.W = h->Wcat[i]; a.ldw = a.N;
    a.D = gt->d[i]; a.Z = Z[i]; a.alpha = h->alpha; a.skip_if_zero = skip;
    cudaError_t le = launch_gemm<0>(a, st);
    if (le != cudaSuccess) { set_error("gated_gemm<0> launch: %s", cudaGetErrorString(le)); return ICNN_E_CUDA; }
  }
  out_layer_launch(h, gt, Z[L - 1], y32, f, dl[0], nullptr, nullptr, g, g_row_stride, perm, count, KS, skip, st);
  int cur = 0;
  for (int i = L - 1; i >= 0; --i) {  // backward hidden layers
    GemmArgs a{};
    a.M = B; a.N0 = h->prev(i); a.N = a.N0 + n; a.K0 = h->hidden[i]; a.K1 = 0;
    a.A0 = dl[cur]; a.G0 = nullptr; a.lda0 = a.K0;
    a.W = h->Wcat[i]; a.ldw = a.K0; a.alpha = h->alpha;
    a.Zprev = i ? Z[i - 1] : nullptr; a.Cz = i ? gt->cz[i] : nullptr; a.dprev = dl[cur ^ 1];
    a.Cy = gt->cy[i]; a.g = g; a.g_row_stride = g_row_stride; a.perm = perm; a.count = count; a.KS = KS;
    a.n = n; a.g_scale = gt->g_scale; a.skip_if_zero = skip;
    cudaError_t le = launch_gemm<1>(a, st);
    if (le != cudaSuccess) { set_error("gated_gemm<1> launch: %s", cudaGetErrorString(le)); return ICNN_E_CUDA; }
    cur ^= 1;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("picnn_fg launch: %s", cudaGetErrorString(e)); return ICNN_E_CUDA; }
  return ICNN_OK;
}

// K1 dispatch: tcgen05 path for TMA-compatible shapes (every width % 4 == 0) with enough rows to
// fill a 128-row tile, FP32 FFMA path otherwise.  ICNN_K1=simt|tc forces one.
int picnn_fg_dispatch(const icnn_picnn* h, const icnn_gates* gt, const float* y32, float* f, float* g,
                      long long g_row_stride, const int* perm, const int* count, int KS, void* workspace,
                      const int* skip, cudaStream_t st) {
  if (h->use_tc && gt->B >= 64)
    return picnn_fg_tc(h, gt, y32, f, g, g_row_stride, perm, count, KS, workspace, skip, st);
  return picnn_fg_simt(h, gt, y32, f, g, g_row_stride, perm, count, KS, workspace, skip, st);
}

}  // namespace icnn

using namespace icnn;

extern "C" int icnn_picnn_create(const icnn_picnn_desc* d, icnn_picnn_t** out, void* stream) {
  ICNN_REQUIRE(d && out, "null descriptor");
  ICNN_REQUIRE(d->L >= 1 && d->L <= ICNN_MAX_LAYERS, "L must be in [1, 8]");
  ICNN_REQUIRE(d->n >= 1, "n must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  icnn_picnn* h = new icnn_picnn();
  h->n = d->n; h->L = d->L; h->alpha = d->alpha;
  for (int i = 0; i < d->L; ++i) {
    if (d->hidden[i] < 1) { delete h; set_error("hidden width must be positive"); return ICNN_E_INVALID; }
    h->hidden[i] = d->hidden[i];
  }
  for (int i = 0; i <= d->L; ++i) h->Wcat[i] = nullptr;
  for (int i = 0; i < ICNN_MAX_LAYERS; ++i) h->Wb_hi[i] = h->Wb_lo[i] = h->Wf_hi[i] = h->Wf_lo[i] = nullptr;
  h->use_tc = false;
  h->m = 0; h->has_xpath = false;
  for (int i = 0; i <= ICNN_MAX_LAYERS; ++i) h->Xw_hi[i] = h->Xw_lo[i] = h->Xbias[i] = nullptr;
  for (int i = 0; i <= d->L; ++i) {
    const long long si = h->width(i), sp = h->prev(i);
    const long long ntop = sp * si, nbot = (long long)h->n * si;
    if (!d->Wy[i] || (i > 0 && !d->Wz[i])) { icnn_picnn_destroy(h); set_error("null weight pointer, layer %d", i); return ICNN_E_INVALID; }
    cudaError_t e = cudaMalloc(&h->Wcat[i], sizeof(float) * (ntop + nbot));
    if (e != cudaSuccess) { icnn_picnn_destroy(h); set_error("cudaMalloc weights: %s", cudaGetErrorString(e)); return ICNN_E_CUDA; }
    const long long tot = ntop + nbot;
    concat_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(h->Wcat[i], d->Wz[i], ntop, d->Wy[i], nbot);
  }
  {
    const char* k1 = getenv("ICNN_K1");
    const bool want = !(k1 && k1[0] == 's');
    if (want && picnn_tc_supported(h)) {
      int rc = picnn_tc_prepare_weights(h, st);
      if (rc) { icnn_picnn_destroy(h); return rc; }
      h->use_tc = true;
    } else if (k1 && k1[0] == 't') {
      icnn_picnn_destroy(h); set_error("ICNN_K1=tc but the shape is not TMA-compatible (widths %% 4)"); return ICNN_E_UNSUPPORTED;
    }
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { icnn_picnn_destroy(h); set_error("picnn_create: %s", cudaGetErrorString(e)); return ICNN_E_CUDA; }
  *out = h;
  return ICNN_OK;
}

extern "C" int icnn_picnn_destroy(icnn_picnn_t* h) {
  if (!h) return ICNN_OK;
  for (int i = 0; i <= h->L && i <= ICNN_MAX_LAYERS; ++i)
    if (h->Wcat[i]) cudaFree(h->Wcat[i]);
  picnn_tc_free_weights(h);
  picnn_xpath_free(h);
  delete h;
  return ICNN_OK;
}

extern "C" size_t icnn_picnn_workspace_bytes(const icnn_picnn_t* h, int32_t B) {
  if (!h || B <= 0) return 0;
  return sizeof(float) * (picnn_simt_ws_floats(h, B, nullptr, nullptr) +
                          (h->use_tc ? picnn_tc_ws_floats(h, B, nullptr, nullptr) : 0));
}

extern "C" int icnn_picnn_fg(const icnn_picnn_t* h, const icnn_gates* gates, const float* y32, float* f,
                             float* g, int64_t g_row_stride, const int32_t* perm, const int32_t* count,
                             int32_t KS, void* workspace, const int32_t* skip_if_zero, void* stream) {
  ICNN_REQUIRE(h && gates && y32 && f && g && workspace, "null pointer");
  ICNN_REQUIRE(gates->B > 0, "empty batch");
  ICNN_REQUIRE((perm == nullptr) == (count == nullptr), "perm and count go together");
  return picnn_fg_dispatch(h, gates, y32, f, g, g_row_stride, perm, count, KS, workspace, skip_if_zero,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int icnn_gd_solve(const icnn_picnn_t* h, const icnn_gates* gates, float* y32, float* v, float* g,
                             float* f_out, int32_t nIter, float lr, float momentum, void* workspace,
                             void* stream) {
  ICNN_REQUIRE(h && gates && y32 && v && g && f_out && workspace, "null pointer");
  ICNN_REQUIRE(nIter >= 0, "nIter < 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long N = (long long)gates->B * h->n;
  ICNN_CUDA_CHECK(cudaMemsetAsync(v, 0, sizeof(float) * N, st));
  for (int it = 0; it < nIter; ++it) {
    int rc = picnn_fg_dispatch(h, gates, y32, f_out, g, h->n, nullptr, nullptr, 0, workspace, nullptr, st);
    if (rc) return rc;
    gd_update_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(y32, v, g, N, lr, momentum);
  }
  int rc = picnn_fg_dispatch(h, gates, y32, f_out, g, h->n, nullptr, nullptr, 0, workspace, nullptr, st);
  if (rc) return rc;
  ICNN_CUDA_CHECK(cudaGetLastError());
  return ICNN_OK;
}
