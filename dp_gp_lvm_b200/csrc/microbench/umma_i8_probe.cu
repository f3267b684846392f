// Probe of the tcgen05 int8 path used by the sliced ("Ozaki") reductions of the fused psi2 backward (csrc/psi2_bwd_umma.cuh):
// pins the shared-memory descriptor semantics (no swizzle; K-major and MN-major views of ONE byte tile), the N-stacking of
// operand slices with shifted accumulator columns, mixed u8 x s8 operands and the TMEM read-back, against a host integer
// reference.  One CTA, one launch:
//
//   G tile   bytes g[pair][row], blocks of [8 pairs][16 rows] (128 B): pair-group stride SP, row-group stride SR
//            * as A of  dD[pair, n] = sum_row  g[pair][row] * V[n][row]    (M = pairs, K = rows : K-major,  SBO = SP, LBO = SR)
//            * as A of  dv[row,  n] = sum_pair g[pair][row] * D[n][pair]   (M = rows,  K = pairs: MN-major, SBO = SR, LBO = SP)
//   V        K-major B  [n][row],  blocks [8 n][16 rows];   D   MN-major B [pair][n], blocks [8 pairs][16 n]
//   slices   the G tile exists NSA times (planes), V / D hold NSB slices stacked along n (n = 16 j + q); the product of A
//            plane i with B slices 0 .. L - i lands in accumulator columns 16 i .. : one MMA per (i, k-step).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo umma_i8_probe.cu -o umma_i8_probe && ./umma_i8_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../umma.cuh"

using namespace dpgp;

constexpr int NSA = 6, NSB = 6, LV = 6;            // A planes, B slices, accumulator levels (i + j < LV)
constexpr int PAIRS = 64, ROWS = 64;               // valid extent of the G tile (the MMAs run M = 128 and over-read)
constexpr int SP = 128, SR = PAIRS / 8 * 128 + 16; // strides of the [8 pairs][16 rows] blocks
constexpr int PLANE = ROWS / 16 * SR;              // bytes per G plane
constexpr int NB = 16 * NSB;                       // stacked n extent
constexpr int V_LBO = NB / 8 * 128;                // K-major V: n-groups of 8 adjacent (SBO 128), 16-row chunks V_LBO apart
constexpr int D_SBO = PAIRS * 16;                  // MN-major D: one slice (16 n) per plane; pair-groups of 8 adjacent (LBO 128)
constexpr int G_BYTES = NSA * PLANE + 4096;        // + slack: the M = 128 MMAs read past the 64 valid pairs / rows
constexpr int V_BYTES = ROWS / 16 * V_LBO, D_BYTES = NSB * D_SBO;
constexpr int COLS = 16 * LV;

__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* g, const uint8_t* v, const uint8_t* d, int* out_dd, int* out_dv) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* sg = sm; uint8_t* sv = sg + G_BYTES; uint8_t* sd = sv + V_BYTES;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < G_BYTES; i += 128) sg[i] = g[i];
  for (int i = tid; i < V_BYTES; i += 128) sv[i] = v[i];
  for (int i = tid; i < D_BYTES; i += 128) sd[i] = d[i];
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 256);
  fence_proxy_async_smem();                         // generic-proxy writes above -> visible to the tensor core (async proxy)
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_base;
  if (warp == 0 && elect_one()) {
    // dD: columns [0, COLS); dv: columns [128, 128 + COLS)
    for (int ks = 0; ks < ROWS / 32; ++ks)
      for (int i = 0; i < NSA && i < LV; ++i) {
        const int nj = min(NSB, LV - i);
        const uint64_t da = umma_smem_desc(sg + i * PLANE + ks * 2 * SR, /*lbo*/ SR, /*sbo*/ SP);
        const uint64_t db = umma_smem_desc(sv + ks * 2 * V_LBO, /*lbo*/ V_LBO, /*sbo*/ 128);
        umma_i8(tm + 16 * i, da, db, umma_idesc_i8(128, 16 * nj, /*a signed*/ false, /*b signed*/ true, /*a MN*/ false, /*b MN*/ false), !(i == 0 && ks == 0));
      }
    for (int ks = 0; ks < PAIRS / 32; ++ks)
      for (int i = 0; i < NSA && i < LV; ++i) {
        const int nj = min(NSB, LV - i);
        const uint64_t da = umma_smem_desc(sg + i * PLANE + ks * 4 * SP, /*lbo*/ SP, /*sbo*/ SR);
        const uint64_t db = umma_smem_desc(sd + ks * 4 * 128, /*lbo*/ 128, /*sbo*/ D_SBO);
        umma_i8(tm + 128 + 16 * i, da, db, umma_idesc_i8(128, 16 * nj, false, true, /*a MN*/ true, /*b MN*/ true), !(i == 0 && ks == 0));
      }
    umma_commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0);
  tc_fence_after();
  // every warp reads its 32 lanes
  for (int c0 = 0; c0 < COLS; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tm + ((uint32_t)(32 * warp) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out_dd[(32 * warp + (tid & 31)) * COLS + c0 + j] = (int)r[j];
    tmem_ld16(tm + ((uint32_t)(32 * warp) << 16) + 128 + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out_dv[(32 * warp + (tid & 31)) * COLS + c0 + j] = (int)r[j];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_free(tm, 256);
}


// ---- where do the 64 rows of an M = 64 accumulator live in TMEM?  One MMA (K-major A = G plane 0 rows 0..31, B = V slice 0),
// all 128 lanes x 16 columns read back.
__global__ void __launch_bounds__(128, 1) m64_kernel(const uint8_t* g, const uint8_t* v, int* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* sg = sm; uint8_t* sv = sg + G_BYTES;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < G_BYTES; i += 128) sg[i] = g[i];
  for (int i = tid; i < V_BYTES; i += 128) sv[i] = v[i];
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 32);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_base;
  {   // poison the accumulator so untouched lanes are recognisable
    uint32_t r[16];
    for (int j = 0; j < 16; ++j) r[j] = 0x7fffffffu;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(tm + ((uint32_t)(32 * warp) << 16)), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (tid == 0) {
    umma_i8(tm, umma_smem_desc(sg, SR, SP), umma_smem_desc(sv, V_LBO, 128), umma_idesc_i8(64, 16, false, true, false, false), false);
    umma_commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t r[16];
  tmem_ld16(tm + ((uint32_t)(32 * warp) << 16), r);
  tmem_ld_wait();
  for (int j = 0; j < 16; ++j) out[tid * 16 + j] = (int)r[j];
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_free(tm, 32);
}

// ---- issue-rate measurement: `iters` back-to-back MMAs of one shape from one thread; cycles per MMA (clock64 around the
// issue loop and the commit wait).  kind: 0 = i8 (u8 x s8), 1 = f16 (fp16 operands, fp32 accumulate), 2 = f8f6f4 (e4m3)
__device__ __forceinline__ void umma_generic(int kind, uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, bool acc) {
  if (kind == 0) umma_i8(tmem_d, da, db, idesc, acc);
  else if (kind == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)acc) : "memory");
}
__global__ void __launch_bounds__(128, 1) rate_kernel(int kind, int n, int mn_major, int iters, long long* out, int nacc = 1, int m = 128) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    // K-major: 8-row groups 128 B apart, K chunks 2048 + 16 B apart; MN-major: the same tile viewed the other way
    const uint64_t da = mn_major ? umma_smem_desc(sm, 128, 2064) : umma_smem_desc(sm, 2064, 128);
    const uint64_t db = mn_major ? umma_smem_desc(sm + 32768, 128, 1024) : umma_smem_desc(sm + 32768, 2064, 128);
    uint32_t idesc;
    if (kind == 0) idesc = umma_idesc_i8(m, n, false, true, mn_major, mn_major);
    else if (kind == 1) idesc = (1u << 4) | ((mn_major ? 1u : 0u) << 15) | ((mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32
    else idesc = (1u << 4) | ((mn_major ? 1u : 0u) << 15) | ((mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);                // e4m3 x e4m3 -> f32
    const long long t0 = clock64();
    if (kind == 0) {
      // straight-line issue, 8 MMAs per trip, accumulators round-robin: measures the tensor pipe, not the issuing thread
      if (n > 128) nacc = 1;
      const uint32_t t1 = nacc > 1 ? tm + 128 : tm, t2 = nacc > 2 ? tm + 256 : tm, t3 = nacc > 2 ? tm + 384 : t1;
      umma_i8(tm, da, db, idesc, false); umma_i8(t1, da, db, idesc, false); umma_i8(t2, da, db, idesc, false); umma_i8(t3, da, db, idesc, false);
#pragma unroll 1
      for (int i = 0; i < iters; i += 8) {
        umma_i8(tm, da, db, idesc, true); umma_i8(t1, da, db, idesc, true); umma_i8(t2, da, db, idesc, true); umma_i8(t3, da, db, idesc, true);
        umma_i8(tm, da, db, idesc, true); umma_i8(t1, da, db, idesc, true); umma_i8(t2, da, db, idesc, true); umma_i8(t3, da, db, idesc, true);
      }
    } else {
      for (int i = 0; i < iters; ++i) umma_generic(kind, tm, da, db, idesc, i > 0);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_free(tm, 512);
}

int rate_main() {
  long long* out; cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const char* names[3] = {"i8  (u8 x s8 -> s32, K = 32)", "f16 (fp16 -> fp32,  K = 16)", "f8  (e4m3 -> fp32,  K = 32)"};
  for (int kind = 0; kind < 3; ++kind)
    for (int mn = 0; mn < 2; ++mn)
      for (int n : {16, 32, 64, 112, 128, 256}) {
        const int iters = 2000;
        rate_kernel<<<148, 128, 64 * 1024>>>(kind, n, mn, iters, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s N %d mn %d: CUDA error %s\n", names[kind], n, mn, cudaGetErrorString(e)); return 1; }
        long long h[148]; cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i] / 148;
        const double cyc = avg / iters, kk = kind == 1 ? 16 : 32;
        printf("%s  M 128 N %3d %s: %7.1f cycles / MMA = %7.0f MAC / clk / SM (all 148 SMs busy)\n", names[kind], n, mn ? "MN-major" : "K-major ", cyc, 128.0 * n * kk / cyc);
      }
  printf("-- independent accumulators (i8, K-major): the same MMA stream round-robin over nacc TMEM accumulators; M = 64\n");
  for (int m : {128, 64})
    for (int nacc : {1, 2, 4})
      for (int n : {16, 32, 64, 112, 128, 256}) {
        if (n == 256 && nacc > 1) continue;
        const int iters = 4000;
        rate_kernel<<<148, 128, 64 * 1024>>>(0, n, 0, iters, out, nacc, m);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("M %d nacc %d N %d: CUDA error %s\n", m, nacc, n, cudaGetErrorString(e)); return 1; }
        long long h[148]; cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i] / 148;
        printf("i8 M %3d N %3d, %d accumulators: %7.1f cycles / MMA = %6.0f MAC / clk / SM\n", m, n, nacc, avg / iters, (double)m * n * 32 / (avg / iters));
      }
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1) return rate_main();
  std::vector<uint8_t> g(G_BYTES), v(V_BYTES), d(D_BYTES);
  srand(1);
  for (auto& x : g) x = rand() & 255;             // includes the over-read slack: garbage must not matter for valid outputs
  std::vector<int> ga(NSA * PAIRS * ROWS), va(NB * ROWS), da(NB * PAIRS);
  for (int i = 0; i < NSA; ++i)
    for (int p = 0; p < PAIRS; ++p)
      for (int r = 0; r < ROWS; ++r) {
        const int val = rand() & 255;
        ga[(i * PAIRS + p) * ROWS + r] = val;
        g[i * PLANE + (p / 8) * SP + (r / 16) * SR + (p % 8) * 16 + r % 16] = (uint8_t)val;
      }
  for (int n = 0; n < NB; ++n)
    for (int r = 0; r < ROWS; ++r) {
      const int val = (rand() & 255) - 128;
      va[n * ROWS + r] = val;
      v[(n / 8) * 128 + (r / 16) * V_LBO + (n % 8) * 16 + r % 16] = (uint8_t)(int8_t)val;
    }
  for (int n = 0; n < NB; ++n)
    for (int p = 0; p < PAIRS; ++p) {
      const int val = (rand() & 255) - 128;
      da[n * PAIRS + p] = val;
      d[(n / 16) * D_SBO + (p / 8) * 128 + (p % 8) * 16 + n % 16] = (uint8_t)(int8_t)val;
    }
  uint8_t *dg, *dvv, *dd; int *odd, *odv;
  cudaMalloc(&dg, G_BYTES); cudaMalloc(&dvv, V_BYTES); cudaMalloc(&dd, D_BYTES);
  cudaMalloc(&odd, 128 * COLS * 4); cudaMalloc(&odv, 128 * COLS * 4);
  cudaMemcpy(dg, g.data(), G_BYTES, cudaMemcpyHostToDevice);
  cudaMemcpy(dvv, v.data(), V_BYTES, cudaMemcpyHostToDevice);
  cudaMemcpy(dd, d.data(), D_BYTES, cudaMemcpyHostToDevice);
  const int smem = G_BYTES + V_BYTES + D_BYTES + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(dg, dvv, dd, odd, odv);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<int> hdd(128 * COLS), hdv(128 * COLS);
  cudaMemcpy(hdd.data(), odd, hdd.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hdv.data(), odv, hdv.size() * 4, cudaMemcpyDeviceToHost);
  long bad_dd = 0, bad_dv = 0;
  for (int lv = 0; lv < LV; ++lv)
    for (int q = 0; q < 16; ++q) {
      for (int p = 0; p < PAIRS; ++p) {          // dD[pair][level, q] = sum_{i + j = lv} sum_row g_i[pair][row] v_j[q][row]
        long s = 0;
        for (int i = 0; i <= lv && i < NSA; ++i) { const int j = lv - i; if (j >= NSB) continue;
          for (int r = 0; r < ROWS; ++r) s += (long)ga[(i * PAIRS + p) * ROWS + r] * va[(16 * j + q) * ROWS + r]; }
        if ((int)s != hdd[p * COLS + 16 * lv + q]) { if (bad_dd < 5) printf("dD mismatch pair %d level %d q %d: %d vs %ld\n", p, lv, q, hdd[p * COLS + 16 * lv + q], s); ++bad_dd; }
      }
      for (int r = 0; r < ROWS; ++r) {           // dv[row][level, q] = sum_{i + j = lv} sum_pair g_i[pair][row] d_j[q][pair]
        long s = 0;
        for (int i = 0; i <= lv && i < NSA; ++i) { const int j = lv - i; if (j >= NSB) continue;
          for (int p = 0; p < PAIRS; ++p) s += (long)ga[(i * PAIRS + p) * ROWS + r] * da[(16 * j + q) * PAIRS + p]; }
        if ((int)s != hdv[r * COLS + 16 * lv + q]) { if (bad_dv < 5) printf("dv mismatch row %d level %d q %d: %d vs %ld\n", r, lv, q, hdv[r * COLS + 16 * lv + q], s); ++bad_dv; }
      }
    }
  printf("umma_i8_probe: dD (K-major A, K-major B) mismatches %ld / %d, dv (MN-major A, MN-major B) mismatches %ld / %d\n",
         bad_dd, PAIRS * COLS, bad_dv, ROWS * COLS);
  {
    int* o64; cudaMalloc(&o64, 128 * 16 * 4);
    cudaFuncSetAttribute(m64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    m64_kernel<<<1, 128, smem>>>(dg, dvv, o64);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e2 != cudaSuccess) { printf("m64 kernel: CUDA error %s\n", cudaGetErrorString(e2)); return 1; }
    std::vector<int> h64(128 * 16);
    cudaMemcpy(h64.data(), o64, h64.size() * 4, cudaMemcpyDeviceToHost);
    printf("M = 64 accumulator layout (TMEM lane -> matrix row; '.' = untouched):\n");
    for (int lane = 0; lane < 128; ++lane) {
      int found = -1;
      if (h64[lane * 16] != 0x7fffffff)
        for (int p = 0; p < PAIRS && found < 0; ++p) {
          bool ok = true;
          for (int q = 0; q < 16 && ok; ++q) {
            long sacc = 0;
            for (int r = 0; r < 32; ++r) sacc += (long)ga[(0 * PAIRS + p) * ROWS + r] * va[q * ROWS + r];
            ok = (int)sacc == h64[lane * 16 + q];
          }
          if (ok) found = p;
        }
      if (h64[lane * 16] == 0x7fffffff) printf(" .");
      else printf(" %d", found);
      if (lane % 32 == 31) printf("\n");
    }
  }
  return (bad_dd || bad_dv) ? 2 : 0;
}
