// Stand-alone timing harness for the QR panel kernel (tuning aid; compile variants with -DQRV_xxx):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dqmc_b200/csrc -o tools/bin/qr_bench tools/qr_bench.cu
#include "../dqmc_b200/csrc/qr.cu"
#include <vector>
#include <random>
char g_errbuf[512];
long long g_launches = 0;
int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1024;
  std::vector<cplx> h((size_t)n * n);
  std::mt19937_64 rng(1);
  std::normal_distribution<double> nd;
  for (auto& v : h) v = cmake(nd(rng), nd(rng));
  cplx *A, *A0, *tau, *tf; double* dabs;
  cudaMalloc(&A, sizeof(cplx) * n * n); cudaMalloc(&A0, sizeof(cplx) * n * n); cudaMalloc(&tau, sizeof(cplx) * n);
  cudaMalloc(&tf, sizeof(cplx) * 1024 * (n / 32 + 1)); cudaMalloc(&dabs, 8 * n);
  cudaMemcpy(A0, h.data(), sizeof(cplx) * n * n, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaMemcpy(A, A0, sizeof(cplx) * n * n, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e0);
    if (qr_panels_only(0, A, n, n, tau, dabs, tf)) { printf("error %s\n", g_errbuf); return 1; }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  extern int g_larfb_cluster_max_cols;
  QrAsync as; cudaStreamCreate(&as.st2); cudaEventCreateWithFlags(&as.eA, cudaEventDisableTiming); cudaEventCreateWithFlags(&as.eB, cudaEventDisableTiming);
  cudaStream_t st; cudaStreamCreate(&st);
  cplx* Q; cudaMalloc(&Q, sizeof(cplx) * n * n);
  for (int thr : {0, 32, 128, 256, 512, 1024, 4096}) {
    g_larfb_cluster_max_cols = thr;
    float bq = 1e30f, bf = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaMemcpyAsync(A, A0, sizeof(cplx) * n * n, cudaMemcpyDeviceToDevice, st);
      cudaEventRecord(e0, st);
      if (qr_factor(st, A, n, n, tau, dabs, tf, nullptr, 0, 0, 148, &as)) { printf("error %s\n", g_errbuf); return 1; }
      cudaEventRecord(e1, st); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < bq) bq = ms;
      cudaEventRecord(e0, st);
      if (qr_form_q(st, A, n, n, tf, Q, n, 148)) { printf("error %s\n", g_errbuf); return 1; }
      cudaEventRecord(e1, st); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1); if (ms < bf) bf = ms;
    }
    printf("  cluster-larfb up to %4d columns: qr_factor %.3f ms, form_q %.3f ms  %s\n", thr, bq, bf, cudaGetErrorString(cudaDeviceSynchronize()));
  }
  std::vector<double> d(n);
  cudaMemcpy(d.data(), dabs, 8 * n, cudaMemcpyDeviceToHost);
  printf("n=%d panel chain %.3f ms = %.0f ns/column   (|R00|=%.6f |R11|=%.6f) %s\n", n, best, best * 1e6 / n, d[0], d[1],
         cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
