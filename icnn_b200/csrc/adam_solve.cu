// RL Adam argmin on the device.
//
// Reproduces the RL agent's default inner optimiser, RL/src/icnn.py:160-215 (`Agent.adam`), applied to
// func = [negQ - entropy(act), d/dact] (RL/src/icnn.py:60-63,127-131; entropy :455-458):
//   Adam (b1 .9, b2 .999, alpha .01, eps 1e-8; the step divides by sqrt(v), as the reference does) on the
//   whole minibatch of actions, clip to (-1+1e-8, 1-1e-8), best-so-far tracking per sample, stop when
//   the rolling average (lam .5) of mean_u ||act_best - prev_act_best|| drops below 1e-3 after i > 5.
// K1 (picnn f/grad) evaluates negQ; one warp-per-sample kernel adds the entropy term, tracks the best
// action, accumulates the stop statistic and applies the Adam update; a one-block kernel folds the
// statistic deterministically and raises the device-side `active` flag that turns the remaining
// launches of a chunk into no-ops.
#include "common.cuh"

namespace icnn {

int picnn_fg_dispatch(const icnn_picnn* h, const icnn_gates* gt, const float* y32, float* f, float* g,
                      long long g_row_stride, const int* perm, const int* count, int KS, void* workspace,
                      const int* skip, cudaStream_t st);

struct AdamArgs {
  int B, n, it;
  double* act; float* act32; double* m; double* v; double* act_best; double* f_best;
  const float* f; const float* g;
  double* nrm;      // [B] per-sample ||act_best - prev_act_best||
  double b1t;       // b1^(it+1)
  int* active;      // [1] 1 = still iterating
  double* stat;     // [2] a_diff, have_a_diff
  int* iters;       // [1]
};

__global__ void __launch_bounds__(256) adam_step_kernel(AdamArgs a) {
  if (*a.active == 0) return;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.B) return;
  const int u = warp, n = a.n;
  const double b1 = 0.9, b2 = 0.999, eps = 1e-8, alpha = 0.01;
  // f_entr = negQ + sum pen(act),  pen = xr log xr + (1-xr) log(1-xr),  xr = clip((act+1)/2, 1e-4, 1-1e-4)
  double fs = 0.0;
  for (int e = lane; e < n; e += 32) {
    const double xr = fmin(fmax((a.act[(size_t)u * n + e] + 1.0) * 0.5, 0.0001), 0.9999);
    fs += xr * log(xr) + (1.0 - xr) * log(1.0 - xr);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fs += __shfl_xor_sync(0xffffffffu, fs, o);
  const double fe = (double)a.f[u] + fs;
  bool better;
  if (a.it == 0) better = true;
  else better = fe < a.f_best[u];
  double nr = 0.0;
  for (int e = lane; e < n; e += 32) {
    const size_t i = (size_t)u * n + e;
    const double act = a.act[i];
    if (better) {
      if (a.it > 0) { const double d = act - a.act_best[i]; nr = fma(d, d, nr); }
      a.act_best[i] = act;
    }
    // Adam update on g_entr = g + dpen/dact (gradient passes inside the clip range only)
    const double xraw = (act + 1.0) * 0.5;
    const double xr = fmin(fmax(xraw, 0.0001), 0.9999);
    const double gp = (xraw >= 0.0001 && xraw <= 0.9999) ? 0.5 * (log(xr) - log(1.0 - xr)) : 0.0;
    const double ge = (double)a.g[i] + gp;
    const double mm = b1 * a.m[i] + (1.0 - b1) * ge;
    const double vv = b2 * a.v[i] + (1.0 - b2) * (ge * ge);
    a.m[i] = mm; a.v[i] = vv;
    const double mhat = mm / (1.0 - a.b1t);
    double an = act - alpha * mhat / (sqrt(vv) + eps);
    an = fmin(fmax(an, -1.0 + 1e-8), 1.0 - 1e-8);
    a.act[i] = an;
    a.act32[i] = (float)an;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nr += __shfl_xor_sync(0xffffffffu, nr, o);
  if (lane == 0) {
    if (better) a.f_best[u] = fe;
    a.nrm[u] = sqrt(nr);
  }
}

// one block: a_diff_i = mean_u nrm[u] (fixed summation order), rolling average, stop test
__global__ void __launch_bounds__(256) adam_finalize_kernel(AdamArgs a) {
  if (*a.active == 0) return;
  __shared__ double part[256];
  double s = 0.0;
  for (int u = threadIdx.x; u < a.B; u += 256) s += a.nrm[u];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) {
    *a.iters = a.it;
    if (a.it >= 1) {
      const double di = part[0] / a.B;
      const double ad = (a.stat[1] == 0.0) ? di : 0.5 * a.stat[0] + 0.5 * di;
      a.stat[0] = ad; a.stat[1] = 1.0;
      if (ad < 1e-3 && a.it > 5) *a.active = 0;
    }
  }
}

__global__ void adam_init_kernel(AdamArgs a) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long N = (long long)a.B * a.n;
  if (i < N) { a.act[i] = 0.0; a.act32[i] = 0.f; a.m[i] = 0.0; a.v[i] = 0.0; a.act_best[i] = 0.0; }
  if (i < a.B) a.f_best[i] = 0.0;
  if (i == 0) { *a.active = 1; a.stat[0] = 0.0; a.stat[1] = 0.0; *a.iters = 0; }
}

}  // namespace icnn

using namespace icnn;

// scratch (doubles): act, m, v [B*n] each, nrm [B], stat [2]; (floats) act32, g [B*n], f [B]; (ints) active, iters
extern "C" size_t icnn_adam_workspace_bytes(int32_t B, int32_t n) {
  if (B <= 0 || n <= 0) return 0;
  const size_t N = (size_t)B * n;
  return sizeof(double) * (3 * N + B + 2) + sizeof(float) * (2 * N + B) + sizeof(int) * 4 + 64;
}

extern "C" int icnn_adam_solve(const icnn_picnn_t* h, const icnn_gates* gates, double* act_best, double* f_best,
                               int32_t max_iter, int32_t* iters_out, void* scratch, void* workspace, void* stream) {
  ICNN_REQUIRE(h && gates && act_best && f_best && iters_out && scratch && workspace, "null pointer");
  ICNN_REQUIRE(max_iter >= 1, "max_iter < 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int B = gates->B, n = h->n;
  const size_t N = (size_t)B * n;
  double* dp = static_cast<double*>(scratch);
  AdamArgs a{};
  a.B = B; a.n = n;
  a.act = dp; a.m = dp + N; a.v = dp + 2 * N; a.nrm = dp + 3 * N; a.stat = a.nrm + B;
  float* fp = reinterpret_cast<float*>(a.stat + 2);
  a.act32 = fp; float* g = fp + N; float* f = g + N;
  a.f = f; a.g = g;
  a.active = reinterpret_cast<int*>(f + B + ((B & 1) ? 1 : 0));
  a.iters = a.active + 1;
  a.act_best = act_best; a.f_best = f_best;
  adam_init_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(a);
  double b1t = 1.0;
  int host_active = 1, it = 0;
  const int CHUNK = 16;   // iterations enqueued between two reads of the device-side stop flag
  while (it < max_iter && host_active) {
    const int end = (it + CHUNK < max_iter) ? it + CHUNK : max_iter;
    for (; it < end; ++it) {
      int rc = picnn_fg_dispatch(h, gates, a.act32, f, g, n, nullptr, nullptr, 0, workspace, a.active, st);
      if (rc) return rc;
      b1t *= 0.9;
      a.it = it; a.b1t = b1t;
      adam_step_kernel<<<cdiv(B * 32, 256), 256, 0, st>>>(a);
      adam_finalize_kernel<<<1, 256, 0, st>>>(a);
    }
    ICNN_CUDA_CHECK(cudaMemcpyAsync(&host_active, a.active, sizeof(int), cudaMemcpyDeviceToHost, st));
    ICNN_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  int its = 0;
  ICNN_CUDA_CHECK(cudaMemcpyAsync(&its, a.iters, sizeof(int), cudaMemcpyDeviceToHost, st));
  ICNN_CUDA_CHECK(cudaStreamSynchronize(st));
  *iters_out = host_active ? max_iter : its;
  return ICNN_OK;
}
