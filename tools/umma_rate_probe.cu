// Bring-up probe: sustained tcgen05.mma rate by shape and CTA-group size, with both operands in
// shared memory in the conv kernel's no-swizzle K-major layout.  Answers: how much of a layer's time
// is shared-memory operand traffic, and what cta_group::2 (each CTA supplies half of B) buys.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/umma_rate_probe tools/umma_rate_probe.cu
//   ./umma_rate_probe            (prints one line per shape; run under `timeout`)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

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
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int CG>
__device__ __forceinline__ void umma(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                     uint32_t idesc) {
  if constexpr (CG == 1)
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, 1;\n\t}" ::"r"(d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, 1;\n\t}" ::"r"(d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
}

constexpr int kHaloW = 36;                      // 5x5 layer, two 8x16 sub-tiles side by side
constexpr int kPlane = kHaloW * 12 * 16;        // one 8-channel plane of the halo tile
constexpr int kABytes = 4 * kPlane;             // hi k0, hi k1, lo k0, lo k1
constexpr int kBStage = 5 * 256 * 64;           // room for 5 taps of N <= 256
constexpr int kSmem = kABytes + 2 * kBStage + 1024;

// CG CTAs per cluster; the leader of each cluster issues `reps` x 40 MMAs of shape (128*CG) x N x 16:
// 5 taps x 2 sub-tiles x 4 accumulators-worth of rotation, operands addressed like the conv kernel does.
template <int CG, int N>
__global__ void __launch_bounds__(128) rate_kernel(int reps, long long* cycles, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t rank = 0;
  if constexpr (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  // operands: finite bf16 values with random mantissas
  for (uint32_t i = tid; i < (kABytes + 2 * kBStage) / 4; i += 128)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u | ((i * 2654435761u) & 0x007f007fu);
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
  constexpr uint32_t idesc = make_idesc(128 * CG, N);
  constexpr uint32_t a_hi32 = ((uint32_t)(kHaloW * 16) >> 4) | (1u << 14);  // SBO = one halo row
  constexpr uint32_t b_hi32 = (128u >> 4) | (1u << 14);
  constexpr int NB = N / CG;  // B rows held by this CTA
  const uint32_t a_lo32 = (smem_u32(smem) >> 4) | ((uint32_t)(kPlane >> 4) << 16);
  const uint32_t b_lo32 = (smem_u32(smem + kABytes) >> 4) | ((uint32_t)((NB * 16) >> 4) << 16);
  bool ok = true;
  if (rank == 0 && warp == 1) {
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
      const uint32_t bst = b_lo32 + (uint32_t)((r & 1) * (kBStage >> 4));
      uint32_t elected;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
      if (elected) {
#pragma unroll
        for (int t = 0; t < 5; t++) {
          const uint32_t bt = bst + (uint32_t)(t * ((NB * 64) >> 4));
#pragma unroll
          for (int pass = 0; pass < 4; pass++) {  // hi*hi, lo*hi, hi*lo (x) two sub-tiles -> 4 accumulators
#pragma unroll
            for (int s = 0; s < 2; s++) {
              constexpr int kAccs = 512 / N >= 4 ? 4 : 512 / N;
              const uint32_t d = tmem_base + (uint32_t)(((pass * 2 + s) % kAccs) * N);
              umma<CG>(d, a_lo32 + (uint32_t)(t + s * 8) + (pass == 1 ? (uint32_t)(2 * kPlane >> 4) : 0u), a_hi32,
                       bt + (pass == 2 ? (uint32_t)((NB * 32) >> 4) : 0u), b_hi32, idesc);
            }
          }
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
    ok = mbar_wait_bounded(&bar, 0);
    long long t1 = clock64();
    if ((tid & 31) == 0) {
      cycles[blockIdx.x / CG] = t1 - t0;
      if (!ok) status[0] = 1;
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

template <int CG, int N>
static void run(int grid, int reps) {
  long long* d_cycles;
  int* d_status;
  CK(cudaMalloc(&d_cycles, 256 * sizeof(long long)));
  CK(cudaMalloc(&d_status, sizeof(int)));
  CK(cudaMemset(d_status, 0, sizeof(int)));
  auto kern = rate_kernel<CG, N>;
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
  for (int it = 0; it < 2; it++) {  // second launch is the measurement
    CK(cudaLaunchKernelEx(&cfg, kern, reps, d_cycles, d_status));
    CK(cudaDeviceSynchronize());
  }
  long long c[256];
  int st;
  CK(cudaMemcpy(c, d_cycles, (grid / CG) * sizeof(long long), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&st, d_status, sizeof(int), cudaMemcpyDeviceToHost));
  double mx = 0, sum = 0;
  for (int i = 0; i < grid / CG; i++) { sum += (double)c[i]; if ((double)c[i] > mx) mx = (double)c[i]; }
  const double mmas = 40.0 * reps;
  const double per = sum / (grid / CG) / mmas;
  // operand bytes one CTA reads per MMA: A 128x16 + its share of B (N/CG)x16, bf16
  const double bytes = (128 + N / CG) * 32.0;
  printf("cta_group::%d M=%d N=%3d grid=%3d  %.1f cycles/MMA (max %.1f)  floor %.0f  %.0f B/clk/SM operand reads  %.0f MAC/clk/SM%s\n",
         CG, 128 * CG, N, grid, per, mx / mmas, 128.0 * N / 256.0, bytes / per, 128.0 * N * 16 / per,
         st ? "  TIMEOUT" : "");
  fflush(stdout);
  cudaFree(d_cycles);
  cudaFree(d_status);
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 2000;
  int dev_sms = 0;
  CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = dev_sms & ~1;
  for (int g : {2, grid}) {
    run<1, 32>(g, reps);
    run<1, 64>(g, reps);
    run<1, 128>(g, reps);
    run<1, 256>(g, reps);
    run<2, 64>(g, reps);
    run<2, 128>(g, reps);
    run<2, 256>(g, reps);
  }
  return 0;
}
