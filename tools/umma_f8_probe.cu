// Bring-up probe for the planned fp8 correction passes: tcgen05.mma kind::f8f6f4 with an e5m2 A operand
// and an e4m3 B operand, both in shared memory in the conv kernel's no-swizzle K-major layout (a core
// matrix row is 16 bytes = 16 fp8 values; one MMA is K=32 = two core matrices along K).
//   1. correctness: D[128 x N] = A[128 x 32] * B[N x 32]^T with exactly representable values
//   2. sustained rate by N for cta_group::1 and ::2 (same loop as tools/umma_rate_probe.cu)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_f8_probe tools/umma_f8_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  for (int it = 0; it < (1 << 24); it++) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return true;
  }
  return false;
}
// instruction descriptor, kind::f8f6f4: D=f32 (bit 4), A format bits 7-9 (E4M3=0, E5M2=1), B format bits 10-12,
// K-major both, N>>3 at 17, M>>4 at 24
__host__ __device__ constexpr uint32_t make_idesc_f8(int M, int N, int afmt, int bfmt) {
  return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
template <int CG>
__device__ __forceinline__ void umma_f8(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t acc) {
  if constexpr (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n\t}" ::"r"(d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], da, db, %5, p;\n\t}" ::"r"(d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr int kABytes = 16 * 1024, kBBytes = 64 * 1024, kSmem = kABytes + kBBytes + 1024;

// mode 0: one MMA (K=32), write D; mode 1: rate loop
template <int CG, int N>
__global__ void __launch_bounds__(128) f8_kernel(const uint8_t* a_img, const uint8_t* b_img, int mode, int reps,
                                                 float* d_out, long long* cycles, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t rank = 0;
  if constexpr (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (mode == 0) {
    for (int i = tid; i < kABytes / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    for (int i = tid; i < kBBytes / 16; i += 128)
      reinterpret_cast<uint4*>(smem + kABytes)[i] = reinterpret_cast<const uint4*>(b_img)[i];
  } else {
    for (uint32_t i = tid; i < (kABytes + kBBytes) / 4; i += 128)  // finite fp8 values, random mantissas
      reinterpret_cast<uint32_t*>(smem)[i] = 0x38383838u | ((i * 2654435761u) & 0x03030303u);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  constexpr uint32_t idesc = make_idesc_f8(128 * CG, N, /*A e5m2*/ 1, /*B e4m3*/ 0);
  constexpr int NB = N / CG;
  // A: [k16 0|1][128 rows][16 B]: LBO = 128*16, SBO = 128 (8-row groups contiguous)
  const uint32_t a_lo32 = (smem_u32(smem) >> 4) | ((uint32_t)((128 * 16) >> 4) << 16);
  const uint32_t b_lo32 = (smem_u32(smem + kABytes) >> 4) | ((uint32_t)((NB * 16) >> 4) << 16);
  constexpr uint32_t hi32 = (128u >> 4) | (1u << 14);
  bool ok = true;
  if (rank == 0 && warp == 1) {
    long long t0 = clock64();
    if (mode == 0) {
      if ((tid & 31) == 0) {
        umma_f8<CG>(tmem_base, a_lo32, hi32, b_lo32, hi32, idesc, 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
    } else {
      for (int r = 0; r < reps; r++) {
        uint32_t elected;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
        if (elected) {
#pragma unroll
          for (int j = 0; j < 40; j++) {
            constexpr int kAccs = 512 / N >= 4 ? 4 : 512 / N;
            umma_f8<CG>(tmem_base + (uint32_t)((j % kAccs) * N), a_lo32 + (uint32_t)(j % 5), hi32,
                        b_lo32 + (uint32_t)((j % 4) * ((NB * 32) >> 4)), hi32, idesc, 1u);
          }
          if (r == reps - 1) {
            if constexpr (CG == 1)
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            else
              asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
          }
        }
        __syncwarp();
      }
    }
    ok = mbar_wait_bounded(&bar, 0);
    long long t1 = clock64();
    if ((tid & 31) == 0) {
      cycles[blockIdx.x / CG] = t1 - t0;
      if (!ok) status[0] = 1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (mode == 0) {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int j = 0; j < 16; j++) d_out[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 0) {
    if constexpr (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// exactly representable values and their encodings
static const float kVals[8] = {0.f, 1.f, -1.f, 2.f, 0.5f, -2.f, 1.5f, -0.5f};
static const uint8_t kE5M2[8] = {0x00, 0x3C, 0xBC, 0x40, 0x38, 0xC0, 0x3E, 0xB8};  // bias 15, 2 mantissa bits
static const uint8_t kE4M3[8] = {0x00, 0x38, 0xB8, 0x40, 0x30, 0xC0, 0x3C, 0xB0};  // bias 7, 3 mantissa bits

template <int CG, int N>
static void launch(const uint8_t* a, const uint8_t* b, int mode, int reps, int grid, float* d_out, long long* cyc, int* st) {
  auto kern = f8_kernel<CG, N>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = kSmem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, kern, a, b, mode, reps, d_out, cyc, st));
  CK(cudaDeviceSynchronize());
}

template <int CG, int N>
static void rate(int grid, int reps, long long* d_cyc, int* d_st) {
  CK(cudaMemset(d_st, 0, sizeof(int)));
  for (int it = 0; it < 2; it++) launch<CG, N>(nullptr, nullptr, 1, reps, grid, nullptr, d_cyc, d_st);
  long long c[256];
  int st;
  CK(cudaMemcpy(c, d_cyc, (grid / CG) * sizeof(long long), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&st, d_st, sizeof(int), cudaMemcpyDeviceToHost));
  double sum = 0;
  for (int i = 0; i < grid / CG; i++) sum += (double)c[i];
  const double per = sum / (grid / CG) / (40.0 * reps);
  printf("f8f6f4 cta_group::%d M=%d N=%3d K=32 grid=%3d  %.1f cycles/MMA  %.0f MAC/clk/SM  %.0f B/clk/SM operand reads%s\n", CG,
         128 * CG, N, grid, per, 128.0 * N * 32 / per, (128 + N / CG) * 32.0 / per, st ? "  TIMEOUT" : "");
  fflush(stdout);
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 2000;
  constexpr int N = 64;
  // ---- correctness: A[128][32] e5m2, B[64][32] e4m3, layout [k16][rows][16 B]
  std::vector<uint8_t> a(kABytes, 0), b(kBBytes, 0);
  std::vector<float> af(128 * 32), bf(N * 32);
  for (int r = 0; r < 128; r++)
    for (int k = 0; k < 32; k++) {
      int v = (r * 7 + k * 3 + (r >> 3)) & 7;
      af[r * 32 + k] = kVals[v];
      a[(size_t)(k / 16) * 128 * 16 + r * 16 + (k % 16)] = kE5M2[v];
    }
  for (int n = 0; n < N; n++)
    for (int k = 0; k < 32; k++) {
      int v = (n * 5 + k * 11 + 1) & 7;
      bf[n * 32 + k] = kVals[v];
      b[(size_t)(k / 16) * N * 16 + n * 16 + (k % 16)] = kE4M3[v];
    }
  uint8_t *d_a, *d_b;
  float* d_out;
  long long* d_cyc;
  int* d_st;
  CK(cudaMalloc(&d_a, kABytes));
  CK(cudaMalloc(&d_b, kBBytes));
  CK(cudaMalloc(&d_out, 128 * N * sizeof(float)));
  CK(cudaMalloc(&d_cyc, 256 * sizeof(long long)));
  CK(cudaMalloc(&d_st, sizeof(int)));
  CK(cudaMemcpy(d_a, a.data(), kABytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, b.data(), kBBytes, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_st, 0, sizeof(int)));
  launch<1, N>(d_a, d_b, 0, 1, 1, d_out, d_cyc, d_st);
  std::vector<float> out(128 * N);
  CK(cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost));
  int bad = 0;
  double maxerr = 0;
  for (int r = 0; r < 128; r++)
    for (int n = 0; n < N; n++) {
      double ref = 0;
      for (int k = 0; k < 32; k++) ref += (double)af[r * 32 + k] * bf[n * 32 + k];
      double e = fabs(ref - out[r * N + n]);
      if (e > maxerr) maxerr = e;
      if (e > 1e-6) bad++;
    }
  printf("f8f6f4 e5m2 x e4m3, M=128 N=%d K=32, no-swizzle K-major (16 fp8 per 16-byte row): %s  bad=%d maxerr=%.3g\n", N,
         bad ? "FAIL" : "PASS", bad, maxerr);
  fflush(stdout);
  // ---- rates
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = sms & ~1;
  rate<1, 64>(grid, reps, d_cyc, d_st);
  rate<1, 128>(grid, reps, d_cyc, d_st);
  rate<1, 256>(grid, reps, d_cyc, d_st);
  rate<2, 64>(grid, reps, d_cyc, d_st);
  rate<2, 128>(grid, reps, d_cyc, d_st);
  rate<2, 256>(grid, reps, d_cyc, d_st);
  return 0;
}
