// FP64 pipe microbenchmark for sm_100a: measures the roofline denominator used by bench.py.
//   dfma   : register-resident DFMA chains (vector FP64 pipe)
//   dmma884: mma.sync.m8n8k4.f64 chains (SASS DMMA.8x8x4)
//   dmma16816: mma.sync.m16n8k16.f64 chains (how ptxas lowers the bigger shape)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
// Prints one JSON object; times with CUDA events, best of 5.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_dmma884(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-3; c1[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
               "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void k_dmma16816(double* out, int iters, double av, double bv) {
  double c[ILP][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = av + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = bv + i;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 0; c[i][3] = 1; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// smem-fed DMMA: warp tile (MT m-tiles x NT n-tiles), A (64 x K) and B (K x 128) in shared memory,
// same access pattern as the propagation kernel's H * [Mqq|Mqp] product.
template <int MT, int NT>
__global__ void k_dmma_smem(double* out, int iters, int K) {
  extern __shared__ double sm[];
  const int lda = 60, ldb = 124;
  double* A = sm;               // 64 x lda
  double* B = sm + 64 * lda;    // 64 x ldb
  for (int i = threadIdx.x; i < 64 * lda + 64 * ldb; i += blockDim.x) sm[i] = 1e-3 * (i % 17);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m0 = ((warp % (8 / MT)) * MT) * 8, n0 = ((warp / (8 / MT)) * NT) * 8;
  double c0[MT][NT], c1[MT][NT];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) { c0[i][j] = 0; c1[i][j] = 0; }
  for (int it = 0; it < iters; ++it) {
    for (int k0 = 0; k0 < K; k0 += 4) {
      double a[MT], b[NT];
#pragma unroll
      for (int i = 0; i < MT; ++i) a[i] = A[(m0 + 8 * i + (lane >> 2)) * lda + k0 + (lane & 3)];
#pragma unroll
      for (int j = 0; j < NT; ++j) b[j] = B[(k0 + (lane & 3)) * ldb + (n0 + 8 * j) % 120 + (lane >> 2)];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma884(c0[i][j], c1[i][j], a[i], b[j]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) s += c0[i][j] + c1[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    int blocks = sms * (2048 / threads > 2 ? 2 : 2048 / threads);
    double ms = time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double tf = 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    printf(", \"dfma_tflops_t%d\": %.2f", threads, tf);
  }
  for (int threads : {128, 256, 320, 512, 1024}) {
    int blocks = sms * (threads <= 512 ? 2 : 1);
    double ms = time_ms([&] { k_dmma884<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double tf = 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
    printf(", \"dmma884_tflops_t%d\": %.2f", threads, tf);
  }
  for (int threads : {128, 256, 512}) {
    int blocks = sms * 2;
    double ms = time_ms([&] { k_dmma16816<4><<<blocks, threads>>>(out, iters / 4, 1.0000001, 1e-9); });
    double tf = 2.0 * 16 * 8 * 16 * 4 * (iters / 4) * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
    printf(", \"dmma16816_tflops_t%d\": %.2f", threads, tf);
  }
  {
    size_t smem = sizeof(double) * (64 * 60 + 64 * 124);
    CK(cudaFuncSetAttribute(k_dmma_smem<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_dmma_smem<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_dmma_smem<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_dmma_smem<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int K = 60, it2 = 2000;
    double ms;
    ms = time_ms([&] { k_dmma_smem<4, 3><<<sms, 320, smem>>>(out, it2, K); });
    printf(", \"dmma_smem_4x3_w10_tflops\": %.2f", 2.0 * 256 * 12 * (K / 4) * it2 * (double)sms * 10 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_dmma_smem<8, 1><<<sms, 480, smem>>>(out, it2, K); });
    printf(", \"dmma_smem_8x1_w15_tflops\": %.2f", 2.0 * 256 * 8 * (K / 4) * it2 * (double)sms * 15 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_dmma_smem<2, 3><<<sms, 640, smem>>>(out, it2, K); });
    printf(", \"dmma_smem_2x3_w20_tflops\": %.2f", 2.0 * 256 * 6 * (K / 4) * it2 * (double)sms * 20 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_dmma_smem<4, 4><<<sms, 256, smem>>>(out, it2, K); });
    printf(", \"dmma_smem_4x4_w8_tflops\": %.2f", 2.0 * 256 * 16 * (K / 4) * it2 * (double)sms * 8 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_dmma_smem<4, 3><<<2 * sms, 320, smem>>>(out, it2, K); });
    printf(", \"dmma_smem_4x3_w10_2cta_tflops\": %.2f", 2.0 * 256 * 12 * (K / 4) * it2 * 2.0 * sms * 10 / (ms * 1e-3) / 1e12);
  }
  int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  printf(", \"sm_clock_khz_max\": %d}\n", clk);
  return 0;
}
