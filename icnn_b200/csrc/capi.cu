// C-ABI glue: error string, device query, the fused (K1, K2) x nIter loop.
#include "common.cuh"

#include <cstring>

namespace icnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int picnn_fg_dispatch(const icnn_picnn* h, const icnn_gates* gt, const float* y32, float* f, float* g,
                  long long g_row_stride, const int* perm, const int* count, int KS, void* workspace,
                  const int* skip, cudaStream_t st);
int bundle_step_launch(const icnn_bundle_cfg* cfg, const icnn_bundle_bufs* b, int t, cudaStream_t st);

}  // namespace icnn

using namespace icnn;

extern "C" const char* icnn_last_error(void) { return g_err; }
extern "C" int icnn_abi_version(void) { return ICNN_ABI_VERSION; }
extern "C" int icnn_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); return ICNN_E_CUDA; }
  return n;
}

namespace icnn {
__global__ void __launch_bounds__(256) fp64_mma_probe_kernel(int iters, double* sink) {
  double acc[8][2];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[t][0] = acc[t][1] = 0.0;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < 8; ++t)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(acc[t][0]), "+d"(acc[t][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int t = 0; t < 8; ++t) s += acc[t][0] + acc[t][1];
  if (s == 123.456) *sink = s;   // never true: keeps the loop alive
}
}  // namespace icnn

extern "C" int icnn_fp64_mma_probe(int32_t iters, double* sink, double* flops_out, void* stream) {
  ICNN_REQUIRE(iters > 0 && sink && flops_out, "bad arguments");
  int dev = 0, sms = 0;
  ICNN_CUDA_CHECK(cudaGetDevice(&dev));
  ICNN_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8;
  fp64_mma_probe_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(iters, sink);
  ICNN_CUDA_CHECK(cudaGetLastError());
  *flops_out = (double)blocks * 8.0 /* warps */ * (double)iters * 8.0 /* mma */ * 2.0 * 8 * 8 * 4;
  return ICNN_OK;
}

extern "C" int icnn_solve_batch_fused(const icnn_picnn_t* h, const icnn_gates* gates, const icnn_bundle_cfg* cfg,
                                      const icnn_bundle_bufs* b, void* workspace, void* stream) {
  ICNN_REQUIRE(h && gates && cfg && b && workspace, "null pointer");
  ICNN_REQUIRE(gates->B == b->B, "gates.B != bufs.B");
  ICNN_REQUIRE(h->n == b->n, "picnn.n != bufs.n");
  ICNN_REQUIRE(cfg->nIter >= 1, "nIter < 1");
  ICNN_REQUIRE(b->KS >= 2, "KS < 2");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = icnn_bundle_init(b, cfg->nIter, stream);
  if (rc) return rc;
  for (int t = 0; t < cfg->nIter; ++t) {
    rc = picnn_fg_dispatch(h, gates, b->y32, b->f, b->G, 0, b->perm, b->count, b->KS, workspace,
                       b->nactive + t, st);
    if (rc) return rc;
    rc = icnn_bundle_step(cfg, b, t, stream);
    if (rc) return rc;
  }
  return ICNN_OK;
}
