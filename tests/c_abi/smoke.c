/* Plain-C consumer of the C ABI (include/icnn_b200.h): no Python, no torch.
 * Builds a tiny PICNN (n=4, hidden [8], ReLU), evaluates f / df/dy through icnn_picnn_fg and checks
 * them against a straightforward host loop, then runs one bundle-entropy step and checks the
 * multipliers sum to one, then one step of the GD training backward against its closed form.
 * Compiled and run by tests/test_gpu_c_abi.py. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "icnn_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s\n", cudaGetErrorString(e_)); return 2; } } while (0)
#define RC(x) do { int r_ = (x); if (r_ != 0) { printf("icnn error %d: %s\n", r_, icnn_last_error()); return 3; } } while (0)

static unsigned s_ = 12345u;
static float rnd(void) { s_ = s_ * 1664525u + 1013904223u; return ((s_ >> 8) / 16777216.0f) * 2.f - 1.f; }
static float* dev(const float* h, size_t n) { float* d; cudaMalloc((void**)&d, n * 4); cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice); return d; }

int main(void) {
  enum { B = 5, N = 4, S = 8, KS = 3 };
  if (icnn_abi_version() != ICNN_ABI_VERSION) { printf("ABI mismatch\n"); return 1; }
  float Wy0[N * S], Wy1[N], Wz1[S], cy0[B * N], cy1[B * N], cz1[B * S], d0[B * S], d1[B], y[B * N];
  for (int i = 0; i < N * S; ++i) Wy0[i] = rnd();
  for (int i = 0; i < N; ++i) Wy1[i] = rnd();
  for (int i = 0; i < S; ++i) Wz1[i] = fabsf(rnd());
  for (int i = 0; i < B * N; ++i) { cy0[i] = rnd(); cy1[i] = rnd(); y[i] = 0.25f + 0.5f * fabsf(rnd()); }
  for (int i = 0; i < B * S; ++i) { cz1[i] = fabsf(rnd()); d0[i] = rnd(); }
  for (int i = 0; i < B; ++i) d1[i] = rnd();

  const float* Wy[2] = {dev(Wy0, N * S), dev(Wy1, N)};
  const float* Wz[2] = {NULL, dev(Wz1, S)};
  int32_t hidden[1] = {S};
  icnn_picnn_desc desc = {N, 1, hidden, 0.0f, Wy, Wz};
  icnn_picnn_t* h = NULL;
  RC(icnn_picnn_create(&desc, &h, NULL));
  const float* cy[2] = {dev(cy0, B * N), dev(cy1, B * N)};
  const float* cz[2] = {NULL, dev(cz1, B * S)};
  const float* dd[2] = {dev(d0, B * S), dev(d1, B)};
  icnn_gates gates = {B, cy, cz, dd, 1.f, 0.f, 1.f};
  void* ws; CK(cudaMalloc(&ws, icnn_picnn_workspace_bytes(h, B) + 256));
  float *yd = dev(y, B * N), *fd, *gd;
  CK(cudaMalloc((void**)&fd, B * 4)); CK(cudaMalloc((void**)&gd, B * N * 4));
  RC(icnn_picnn_fg(h, &gates, yd, fd, gd, N, NULL, NULL, 0, ws, NULL, NULL));
  float f[B], g[B * N];
  CK(cudaMemcpy(f, fd, sizeof f, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(g, gd, sizeof g, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int b = 0; b < B; ++b) {   /* host reference of the same PICNN */
    double z[S], fr = d1[b], gr[N];
    for (int j = 0; j < S; ++j) { double a = d0[b * S + j]; for (int e = 0; e < N; ++e) a += (double)y[b * N + e] * cy0[b * N + e] * Wy0[e * S + j]; z[j] = a > 0 ? a : 0; }
    for (int j = 0; j < S; ++j) fr += z[j] * cz1[b * S + j] * Wz1[j];
    for (int e = 0; e < N; ++e) { fr += (double)y[b * N + e] * cy1[b * N + e] * Wy1[e]; gr[e] = (double)cy1[b * N + e] * Wy1[e]; }
    for (int j = 0; j < S; ++j) if (z[j] > 0) for (int e = 0; e < N; ++e) gr[e] += (double)cz1[b * S + j] * Wz1[j] * cy0[b * N + e] * Wy0[e * S + j];
    maxerr = fmax(maxerr, fabs(fr - f[b]));
    for (int e = 0; e < N; ++e) maxerr = fmax(maxerr, fabs(gr[e] - g[b * N + e]));
  }
  printf("fg max err %.3e\n", maxerr);
  if (!(maxerr < 1e-4)) return 4;

  /* one bundle step through the ABI */
  icnn_bundle_bufs bb; memset(&bb, 0, sizeof bb);   /* optional members (f64, iter_stats) = NULL */
  bb.B = B; bb.n = N; bb.KS = KS;
  double yh[B * N]; for (int i = 0; i < B * N; ++i) yh[i] = y[i];
  CK(cudaMalloc((void**)&bb.y, B * N * 8)); CK(cudaMemcpy(bb.y, yh, sizeof yh, cudaMemcpyHostToDevice));
  CK(cudaMalloc((void**)&bb.y32, B * N * 4)); CK(cudaMalloc((void**)&bb.f, B * 4)); CK(cudaMalloc((void**)&bb.G, B * KS * N * 4));
  CK(cudaMalloc((void**)&bb.ys, B * KS * N * 8)); CK(cudaMalloc((void**)&bb.h, B * KS * 8)); CK(cudaMalloc((void**)&bb.lam, B * KS * 8));
  CK(cudaMalloc((void**)&bb.rsum, B * KS * 8)); CK(cudaMalloc((void**)&bb.gram, B * KS * KS * 8)); CK(cudaMalloc((void**)&bb.perm, B * KS * 4));
  CK(cudaMalloc((void**)&bb.count, B * 4)); CK(cudaMalloc((void**)&bb.status, B * 4)); CK(cudaMalloc((void**)&bb.finished, B * 4));
  CK(cudaMalloc((void**)&bb.nIters, B * 4)); CK(cudaMalloc((void**)&bb.nactive, 8 * 4)); CK(cudaMalloc((void**)&bb.newton_its, B * 4));
  CK(cudaMalloc((void**)&bb.ksum, B * 4));
  CK(cudaMemset(bb.gram, 0, B * KS * KS * 8)); CK(cudaMemset(bb.lam, 0, B * KS * 8));
  icnn_bundle_cfg cfg = {ICNN_VARIANT_LIB, ICNN_SOLVER_PC, 1, 0, 1e-8, 1e-12, 2, 0};
  RC(icnn_bundle_init(&bb, 2, NULL));
  RC(icnn_bundle_put_fg(&bb, fd, gd, NULL));
  RC(icnn_bundle_step(&cfg, &bb, 0, NULL));
  CK(cudaDeviceSynchronize());
  double lam[B * KS], ynew[B * N]; int cnt[B];
  CK(cudaMemcpy(lam, bb.lam, sizeof lam, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(cnt, bb.count, sizeof cnt, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ynew, bb.y, sizeof ynew, cudaMemcpyDeviceToHost));
  for (int b = 0; b < B; ++b) {
    double s = 0; for (int j = 0; j < KS; ++j) s += lam[b * KS + j];
    if (cnt[b] != 1 || fabs(s - 1.0) > 1e-6) { printf("sample %d: count %d sum lam %.9f\n", b, cnt[b], s); return 5; }
    for (int e = 0; e < N; ++e) {   /* one cut: y = sigma(-g) */
      const double want = 1.0 / (1.0 + exp((double)g[b * N + e]));
      if (fabs(ynew[b * N + e] - want) > 1e-6) { printf("y mismatch %g %g\n", ynew[b * N + e], want); return 6; }
    }
  }
  /* training backward of ONE momentum-GD step through the ABI; the output layer has a closed form:
   * y1 = y - lr (1+m) g(y), a = scale (y1 - trueY), dWy_L[e] = -lr (1+m) sum_b a[b,e] cy_L[b,e] */
  {
    const float lr = 0.1f, mom = 0.3f, scale = 2.0f / (B * N);
    float zero[B * N] = {0};
    float *tYd = dev(zero, B * N), *yNd, *dWy0, *dWy1, *dWz1, *dcy0, *dcy1, *dcz1;
    CK(cudaMalloc((void**)&yNd, B * N * 4)); CK(cudaMalloc((void**)&dWy0, N * S * 4)); CK(cudaMalloc((void**)&dWy1, N * 4));
    CK(cudaMalloc((void**)&dWz1, S * 4)); CK(cudaMalloc((void**)&dcy0, B * N * 4)); CK(cudaMalloc((void**)&dcy1, B * N * 4));
    CK(cudaMalloc((void**)&dcz1, B * S * 4));
    float* dWy[2] = {dWy0, dWy1}; float* dWz[2] = {NULL, dWz1}; float* dcy[2] = {dcy0, dcy1}; float* dcz[2] = {NULL, dcz1};
    icnn_gd_grads gr = {dWy, dWz, dcy, dcz};
    void* ws2; CK(cudaMalloc(&ws2, icnn_gd_backward_workspace_bytes(h, B, 1) + 256));
    RC(icnn_gd_backward(h, &gates, yd, tYd, scale, 1, lr, mom, yNd, &gr, ws2, NULL));
    CK(cudaDeviceSynchronize());
    float y1[B * N], gw[N];
    CK(cudaMemcpy(y1, yNd, sizeof y1, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(gw, dWy1, sizeof gw, cudaMemcpyDeviceToHost));
    double e1 = 0, e2 = 0;
    for (int e = 0; e < N; ++e) {
      double want = 0;
      for (int b = 0; b < B; ++b) {
        const double yy = (double)y[b * N + e] - (double)lr * (1.0 + mom) * g[b * N + e];
        e1 = fmax(e1, fabs(yy - y1[b * N + e]));
        want += -(double)lr * (1.0 + mom) * scale * yy * cy1[b * N + e];
      }
      e2 = fmax(e2, fabs(want - gw[e]));
    }
    printf("gd_backward: y1 err %.3e, dWy_L err %.3e\n", e1, e2);
    if (!(e1 < 1e-5 && e2 < 1e-6)) return 7;
  }
  icnn_picnn_destroy(h);
  printf("C ABI OK\n");
  return 0;
}
