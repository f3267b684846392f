// FP64 pipe microbenchmarks for B200 (sm_100a).
//
// MEASURED_PEAKS.json holds only the HBM and bf16 peaks; the DP-GP-LVM psi statistics are bound by
// the FP64 pipes, so the roofline denominators for this repository are measured here:
//   * DFMA issue peak (vector FP64 FMA pipe),
//   * DMMA peak (mma.sync .f64 shapes) and whether DMMA overlaps DFMA (separate pipe or not),
//   * cost and accuracy of the FP64 exp variants used by the psi kernels.
// Output: one JSON object on stdout (bench.py and profiles/ consume it).
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo fp64_peaks.cu -o fp64_peaks
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <string>
#include <cuda_runtime.h>

#include "../fast_exp.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)

static __device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}

// ---------------------------------------------------------------- DFMA
template <int CH>
__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters, double a, double b) {
  double x[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) x[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < CH; ++i) x[i] = fma(x[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
}

// ---------------------------------------------------------------- DMMA
static __device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
static __device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
static __device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int CH>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
  double d0[CH], d1[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) { d0[i] = threadIdx.x; d1[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < CH; ++i) dmma884(d0[i], d1[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += d0[i] + d1[i];
  if (s == 123.456) out[0] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double av, double bv) {
  double d[CH][4];
  double a[4] = {av, av * 0.5, av * 0.25, av * 0.125};
  double b[2] = {bv, bv * 0.5};
#pragma unroll
  for (int i = 0; i < CH; ++i) { d[i][0] = threadIdx.x; d[i][1] = i; d[i][2] = 1; d[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < CH; ++i) dmma1688(d[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 123.456) out[0] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double av, double bv) {
  double d[CH][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = av * (1.0 / (1 + i));
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = bv * (1.0 / (1 + i));
#pragma unroll
  for (int i = 0; i < CH; ++i) { d[i][0] = threadIdx.x; d[i][1] = i; d[i][2] = 1; d[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int i = 0; i < CH; ++i) dmma16816(d[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 123.456) out[0] = s;
}

// DMMA and DFMA interleaved in the same warp: NF DFMAs per DMMA (m8n8k4).  If the two share a pipe the
// time is the sum of the parts, if they are separate pipes it is the max.
template <int NF>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double a, double b) {
  constexpr int CH = 4;
  double d0[CH], d1[CH], x[CH][NF > 0 ? NF : 1];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    d0[i] = threadIdx.x; d1[i] = i;
#pragma unroll
    for (int j = 0; j < NF; ++j) x[i][j] = 1.0 + 1e-9 * (i + j);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        dmma884(d0[i], d1[i], a, b);
#pragma unroll
        for (int j = 0; j < NF; ++j) x[i][j] = fma(x[i][j], a, b);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    s += d0[i] + d1[i];
#pragma unroll
    for (int j = 0; j < NF; ++j) s += x[i][j];
  }
  if (s == 123.456) out[0] = s;
}

// ---------------------------------------------------------------- exp variants
// mode 0: libdevice exp, 1: polynomial fast_exp (deg 11), 2: shuffle-table exp (deg 6)
template <int MODE>
__global__ void __launch_bounds__(256) k_exp(double* out, int iters, double x0, double dx) {
  constexpr int CH = 4;
  double acc[CH], x[CH];
  dpgp::ExpTable tab; tab.init();
#pragma unroll
  for (int i = 0; i < CH; ++i) { acc[i] = 0; x[i] = x0 - 0.37 * i - 1e-3 * threadIdx.x; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if (MODE == 0) acc[i] += exp(x[i]);
        else if (MODE == 1) acc[i] = dpgp::exp_acc(x[i], 1.0, acc[i]);
        else acc[i] = tab.exp_acc(x[i], acc[i]);
        x[i] += dx;   // keeps the argument moving (1 extra DADD per exp; subtracted in the report)
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

// accuracy: max relative error of the fast variants against libdevice exp (itself <= 1 ulp)
__global__ void k_exp_acc(double* maxerr, double lo, double hi, int n) {
  dpgp::ExpTable tab; tab.init();
  double e1 = 0, e2 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double x = lo + (hi - lo) * ((double)i + 0.31) / n;
    double ref = exp(x);
    double a = dpgp::exp_acc(x, 1.0, 0.0);
    double b = tab.exp_acc(x, 0.0);
    if (ref > 1e-290) {
      e1 = fmax(e1, fabs(a - ref) / ref);
      e2 = fmax(e2, fabs(b - ref) / ref);
    } else {   // deep underflow: only require "no garbage"
      e1 = fmax(e1, fabs(a - ref));
      e2 = fmax(e2, fabs(b - ref));
    }
  }
  for (int o = 16; o; o >>= 1) {
    e1 = fmax(e1, __shfl_xor_sync(0xffffffffu, e1, o));
    e2 = fmax(e2, __shfl_xor_sync(0xffffffffu, e2, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax((unsigned long long*)&maxerr[0], (unsigned long long)__double_as_longlong(e1));
    atomicMax((unsigned long long*)&maxerr[1], (unsigned long long)__double_as_longlong(e2));
  }
}

__global__ void k_clock(unsigned long long* out, int iters) {
  unsigned long long t0 = gtime(); long long c0 = clock64();
  double x = threadIdx.x;
  for (int i = 0; i < iters; ++i) x = fma(x, 1.0000001, 1e-9);
  long long c1 = clock64(); unsigned long long t1 = gtime();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = (unsigned long long)(c1 - c0); out[1] = t1 - t0; out[2] = (x == 1.5); }
}

template <typename F>
static double time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 64));
  CK(cudaMemset(out, 0, 64));
  const int blocks = sms * 8, threads = 256;
  const int iters = (argc > 1) ? atoi(argv[1]) : 20000;
  const double nthreads = (double)blocks * threads;

  printf("{\"gpu\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_khz_max\": %d", prop.name, sms, prop.major, prop.minor, prop.clockRate);

  // DFMA
  {
    double ms = time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double fma = nthreads * 8.0 * 8.0 * iters;
    printf(", \"dfma_tflops\": %.3f, \"dfma_ms\": %.3f", 2.0 * fma / ms * 1e-9, ms);
    double ms4 = time_ms([&] { k_dfma<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf(", \"dfma_ch4_tflops\": %.3f", 2.0 * nthreads * 4.0 * 8.0 * iters / ms4 * 1e-9);
    double ms2 = time_ms([&] { k_dfma<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf(", \"dfma_ch2_tflops\": %.3f", 2.0 * nthreads * 2.0 * 8.0 * iters / ms2 * 1e-9);
    // one CTA of 128 threads per SM (1 warp per SMSP): latency-bound view
    double ms1w = time_ms([&] { k_dfma<1><<<sms, 128>>>(out, iters, 1.0000001, 1e-9); });
    printf(", \"dfma_dep_chain_ns_per_op\": %.4f", ms1w * 1e6 / (8.0 * iters));
  }
  // DFMA throughput vs resident warps per SMSP (one CTA per SM): how much TLP the FP64 pipe needs
  {
    printf(", \"dfma_tflops_vs_warps_per_smsp\": {");
    const int wps[6] = {1, 2, 3, 4, 6, 8};
    for (int i = 0; i < 6; ++i) {
      const int th = wps[i] * 128;
      double m8 = time_ms([&] { k_dfma<8><<<sms, th>>>(out, iters, 1.0000001, 1e-9); });
      double m4 = time_ms([&] { k_dfma<4><<<sms, th>>>(out, iters, 1.0000001, 1e-9); });
      printf("%s\"%d\": {\"ch8\": %.2f, \"ch4\": %.2f}", i ? ", " : "", wps[i],
             2.0 * sms * th * 8.0 * 8.0 * iters / m8 * 1e-9, 2.0 * sms * th * 4.0 * 8.0 * iters / m4 * 1e-9);
    }
    printf("}");
  }
  // SM clock during an FP64 loop
  {
    unsigned long long* c; CK(cudaMalloc(&c, 32));
    k_clock<<<sms * 8, 256>>>(c, 2000000); CK(cudaDeviceSynchronize());
    unsigned long long h[3]; CK(cudaMemcpy(h, c, 24, cudaMemcpyDeviceToHost));
    printf(", \"sm_mhz_under_fp64\": %.1f", (double)h[0] / (double)h[1] * 1e3);
  }
  // DMMA shapes
  {
    const int it = iters / 4;
    double ms = time_ms([&] { k_dmma884<8><<<blocks, threads>>>(out, it, 1.0000001, 1e-9); });
    double fma = (nthreads / 32.0) * 8.0 * 8.0 * it * 256.0;
    printf(", \"dmma_m8n8k4_tflops\": %.3f", 2.0 * fma / ms * 1e-9);
    ms = time_ms([&] { k_dmma884<2><<<blocks, threads>>>(out, it, 1.0000001, 1e-9); });
    printf(", \"dmma_m8n8k4_ch2_tflops\": %.3f", 2.0 * (nthreads / 32.0) * 2.0 * 8.0 * it * 256.0 / ms * 1e-9);
    ms = time_ms([&] { k_dmma1688<4><<<blocks, threads>>>(out, it, 1.0000001, 1e-9); });
    printf(", \"dmma_m16n8k8_tflops\": %.3f", 2.0 * (nthreads / 32.0) * 4.0 * 4.0 * it * 1024.0 / ms * 1e-9);
    ms = time_ms([&] { k_dmma16816<4><<<blocks, threads>>>(out, it, 1.0000001, 1e-9); });
    printf(", \"dmma_m16n8k16_tflops\": %.3f", 2.0 * (nthreads / 32.0) * 4.0 * 2.0 * it * 2048.0 / ms * 1e-9);
  }
  // mix: per DMMA (256 FMA/warp = 8 DFMA-equivalents) add NF DFMAs
  {
    const int it = iters / 4;
    double t0 = time_ms([&] { k_mix<0><<<blocks, threads>>>(out, it, 1.0000001, 1e-9); });
    double t4 = time_ms([&] { k_mix<4><<<blocks, threads>>>(out, it, 1.0000001, 1e-9); });
    double t8 = time_ms([&] { k_mix<8><<<blocks, threads>>>(out, it, 1.0000001, 1e-9); });
    double t16 = time_ms([&] { k_mix<16><<<blocks, threads>>>(out, it, 1.0000001, 1e-9); });
    // pure-DFMA time for the same number of DFMAs (NF per slot)
    double slots = nthreads * 4.0 * 4.0 * it;   // (thread, chain, unroll, iter)
    printf(", \"mix_ms\": {\"dmma_only\": %.3f, \"dmma+4dfma\": %.3f, \"dmma+8dfma\": %.3f, \"dmma+16dfma\": %.3f, \"slots\": %.4g}", t0, t4, t8, t16, slots);
  }
  // exp variants
  {
    const int it = iters / 8;
    double n = nthreads * 4.0 * 4.0 * it;
    double m0 = time_ms([&] { k_exp<0><<<blocks, threads>>>(out, it, -3.0, -1e-7); });
    double m1 = time_ms([&] { k_exp<1><<<blocks, threads>>>(out, it, -3.0, -1e-7); });
    double m2 = time_ms([&] { k_exp<2><<<blocks, threads>>>(out, it, -3.0, -1e-7); });
    printf(", \"exp_gexp_per_s\": {\"libdevice\": %.2f, \"poly11\": %.2f, \"shfl_table\": %.2f}", n / m0 * 1e-6, n / m1 * 1e-6, n / m2 * 1e-6);
    double* me; CK(cudaMalloc(&me, 16));
    double h[2];
    const double ranges[3][2] = {{-40.0, 5.0}, {-700.0, 700.0}, {-2000.0, -700.0}};
    const char* names[3] = {"[-40,5]", "[-700,700]", "[-2000,-700]"};
    printf(", \"exp_max_rel_err\": {");
    for (int r = 0; r < 3; ++r) {
      CK(cudaMemset(me, 0, 16));
      k_exp_acc<<<sms * 4, 256>>>(me, ranges[r][0], ranges[r][1], 1 << 24); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, me, 16, cudaMemcpyDeviceToHost));
      printf("%s\"%s\": {\"poly11\": %.3e, \"shfl_table\": %.3e}", r ? ", " : "", names[r], h[0], h[1]);
    }
    printf("}");
  }
  printf("}\n");
  return 0;
}
