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
