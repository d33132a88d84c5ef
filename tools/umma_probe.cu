// Bring-up probe for the tcgen05 / TMA building blocks the conv kernel relies on.
// Not part of the product: it answers, on real hardware, the questions the public
// headers leave open (descriptor field semantics for the no-swizzle K-major layout,
// 16-byte-granular start addresses and strides, 5-D TMA boxes with out-of-bounds
// zero fill, TMEM lane mapping, MMA issue rates by shape).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/umma_probe tools/umma_probe.cu
//   ./umma_probe <test> [args]   (each test in its own process; run under `timeout`)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  for (int it = 0; it < (1 << 22); it++) {
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
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= 1ull << 46;  // descriptor version for sm_100
  return d;          // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}
__host__ __device__ inline uint32_t make_idesc(int M, int N, int mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24) |
         (mn_major ? (1u << 15) | (1u << 16) : 0u);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct MmaArgs {
  int N, ksteps;                 // UMMA N; number of K=16 steps
  uint32_t a_bytes, b_bytes;     // smem image sizes
  uint32_t a_off, a_lbo, a_sbo, a_kstep;
  uint32_t b_off, b_lbo, b_sbo, b_kstep;
  int reps;                      // >1: timing mode, repeat the K loop
  int b_region_off;              // where the B image starts in smem (>= a_bytes, 128B aligned)
  int nacc;                      // timing mode: rotate over this many accumulators
  int mn_major;                  // operands are MN-major (K = slow index inside a core matrix)
};

// One CTA, 128 threads.  smem images are copied verbatim from global.
__global__ void __launch_bounds__(128) mma_probe_kernel(const uint8_t* a_img, const uint8_t* b_img,
                                                        MmaArgs g, float* d_out, int* status,
                                                        long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid; i < g.a_bytes / 16; i += 128)
    reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(a_img)[i];
  for (uint32_t i = tid; i < g.b_bytes / 16; i += 128)
    reinterpret_cast<uint4*>(smem + g.b_region_off)[i] = reinterpret_cast<const uint4*>(b_img)[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_s)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = make_idesc(128, g.N, g.mn_major);
  const uint32_t a_base = smem_u32(smem) + g.a_off;
  const uint32_t b_base = smem_u32(smem + g.b_region_off) + g.b_off;
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int r = 0; r < g.reps; r++)
      for (int k = 0; k < g.ksteps; k++)
        umma_bf16(tmem_base + (uint32_t)(((r * g.ksteps + k) % g.nacc) * g.N),
                  make_desc(a_base + k * g.a_kstep, g.a_lbo, g.a_sbo),
                  make_desc(b_base + k * g.b_kstep, g.b_lbo, g.b_sbo), idesc, (r * g.ksteps + k) >= g.nacc);
    umma_commit(&bar);
  }
  __syncwarp();
  bool ok = mbar_wait_bounded(&bar, 0);
  if (tid == 0) {
    t1 = clock64();
    cycles[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok) {
    if (tid == 0) status[0] = 1;
  } else {
    for (int c0 = 0; c0 < g.N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int j = 0; j < 32; j++)
        if (c0 + j < g.N) d_out[(size_t)tid * g.N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u)
                 : "memory");
}

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  uint32_t r = u + 0x7fff + ((u >> 16) & 1);
  return (uint16_t)(r >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// K-major no-swizzle image: element (row, k) at (row/8)*sbo + (k/8)*lbo + (row%8)*16 + (k%8)*2
static void put(std::vector<uint8_t>& img, uint32_t off, uint32_t lbo, uint32_t sbo, int row, int k,
                uint16_t v) {
  size_t b = (size_t)off + (size_t)(row / 8) * sbo + (size_t)(k / 8) * lbo + (row % 8) * 16 + (k % 8) * 2;
  if (b + 2 > img.size()) {
    printf("image overflow\n");
    exit(3);
  }
  memcpy(&img[b], &v, 2);
}

// MN-major no-swizzle image: element (mn, k) at (mn/8)*sbo + (k/8)*lbo + (k%8)*16 + (mn%8)*2
static void put_mn(std::vector<uint8_t>& img, uint32_t off, uint32_t lbo, uint32_t sbo, int mn, int k, uint16_t v) {
  size_t b = (size_t)off + (size_t)(mn / 8) * sbo + (size_t)(k / 8) * lbo + (k % 8) * 16 + (mn % 8) * 2;
  if (b + 2 > img.size()) {
    printf("image overflow\n");
    exit(3);
  }
  memcpy(&img[b], &v, 2);
}

static int run_mma(const char* name, int N, int K, uint32_t a_off, uint32_t a_lbo, uint32_t a_sbo,
                   uint32_t b_lbo, uint32_t b_sbo, bool swap_fields, int reps, int nacc = 1, int mn_major = 0) {
  const int M = 128;
  std::vector<float> A((size_t)M * K), B((size_t)N * K);
  srand(1234);
  for (auto& v : A) v = bf2f(f2bf((rand() % 2001 - 1000) / 1000.0f));
  for (auto& v : B) v = bf2f(f2bf((rand() % 2001 - 1000) / 1000.0f));
  uint32_t a_bytes = a_off + (M / 8) * a_sbo + (K / 8) * a_lbo + 256;
  uint32_t b_bytes = a_off + (N / 8) * b_sbo + (K / 8) * b_lbo + 256;
  a_bytes = (a_bytes + 1023) / 1024 * 1024;
  b_bytes = (b_bytes + 1023) / 1024 * 1024;
  std::vector<uint8_t> ai(a_bytes, 0), bi(b_bytes, 0);
  // poison so that wrong addressing shows up as large errors rather than zeros
  for (size_t i = 0; i + 1 < ai.size(); i += 2) { uint16_t p = f2bf(77.0f); memcpy(&ai[i], &p, 2); }
  for (size_t i = 0; i + 1 < bi.size(); i += 2) { uint16_t p = f2bf(55.0f); memcpy(&bi[i], &p, 2); }
  for (int r = 0; r < M; r++)
    for (int k = 0; k < K; k++)
      (mn_major ? put_mn : put)(ai, a_off, a_lbo, a_sbo, r, k, f2bf(A[(size_t)r * K + k]));
  for (int r = 0; r < N; r++)
    for (int k = 0; k < K; k++)
      (mn_major ? put_mn : put)(bi, mn_major ? a_off : 0, b_lbo, b_sbo, r, k, f2bf(B[(size_t)r * K + k]));
  uint8_t *da, *db;
  float* dd;
  int* ds;
  long long* dc;
  CK(cudaMalloc(&da, a_bytes));
  CK(cudaMalloc(&db, b_bytes));
  CK(cudaMalloc(&dd, (size_t)M * N * 4));
  CK(cudaMalloc(&ds, 4));
  CK(cudaMalloc(&dc, 8));
  CK(cudaMemcpy(da, ai.data(), a_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, bi.data(), b_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0, (size_t)M * N * 4));
  CK(cudaMemset(ds, 0, 4));
  MmaArgs g;
  g.N = N;
  g.ksteps = K / 16;
  g.a_bytes = a_bytes;
  g.b_bytes = b_bytes;
  g.a_off = a_off;
  g.b_off = mn_major ? a_off : 0;
  g.b_region_off = a_bytes;
  g.reps = reps;
  g.nacc = nacc;
  g.mn_major = mn_major;
  if (!swap_fields) {
    g.a_lbo = a_lbo; g.a_sbo = a_sbo; g.b_lbo = b_lbo; g.b_sbo = b_sbo;
  } else {
    g.a_lbo = a_sbo; g.a_sbo = a_lbo; g.b_lbo = b_sbo; g.b_sbo = b_lbo;
  }
  g.a_kstep = 2 * a_lbo;
  g.b_kstep = 2 * b_lbo;
  size_t smem = (size_t)a_bytes + b_bytes;
  if (smem > 227 * 1024) {
    printf("%s: smem %zu too large\n", name, smem);
    return 1;
  }
  CK(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mma_probe_kernel<<<1, 128, smem>>>(da, db, g, dd, ds, dc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: KERNEL ERROR %s\n", name, cudaGetErrorString(e));
    return 1;
  }
  int st;
  long long cyc;
  std::vector<float> D((size_t)M * N);
  CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost));
  if (st) {
    printf("%s: TIMEOUT waiting for the MMA commit\n", name);
    return 1;
  }
  double maxerr = 0;
  int bad = 0;
  for (int r = 0; r < M; r++)
    for (int c = 0; c < N; c++) {
      double ref = 0;
      for (int k = 0; k < K; k++) ref += (double)A[(size_t)r * K + k] * B[(size_t)c * K + k];
      ref *= reps;
      if (nacc > 1) continue;  // rotating accumulators: timing only
      double err = fabs(ref - D[(size_t)r * N + c]);
      if (err > maxerr) maxerr = err;
      if (err > 1e-2 * reps) bad++;
    }
  if (nacc > 1) printf("[nacc=%d] ", nacc);
  printf("%s: N=%d K=%d a_off=%u a_lbo=%u a_sbo=%u b_lbo=%u b_sbo=%u swap=%d reps=%d -> %s maxerr=%.3g bad=%d "
         "cycles=%lld (%.1f per MMA)\n",
         name, N, K, a_off, a_lbo, a_sbo, b_lbo, b_sbo, (int)swap_fields, reps, bad ? "FAIL" : "PASS",
         maxerr, bad, cyc, (double)cyc / (reps * (K / 16)));
  return bad ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// TMA 5-D box with out-of-bounds zero fill into the [c8][y][x][8ch] layout
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

__global__ void tma_probe_kernel(const __grid_constant__ CUtensorMap tmap, int x0, int y0, int c0, int n,
                                 uint32_t bytes, uint8_t* out, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  for (uint32_t i = tid; i < bytes; i += blockDim.x) smem[i] = 0xAB;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem)), "l"(&tmap), "r"(smem_u32(&bar)), "r"(0), "r"(x0), "r"(y0), "r"(c0), "r"(n)
        : "memory");
  }
  bool ok = mbar_wait_bounded(&bar, 0);
  if (!ok && tid == 0) status[0] = 1;
  __syncthreads();
  for (uint32_t i = tid; i < bytes; i += blockDim.x) out[i] = smem[i];
}

static int run_tma() {
  // global activation tensor [N][C8][H][W][8] bf16
  const int N = 2, C8 = 4, H = 20, W = 24;
  const int BW = 12, BH = 10, BC = 2;  // box: 8 ch x BW x BH x BC planes
  std::vector<uint16_t> g((size_t)N * C8 * H * W * 8);
  for (size_t i = 0; i < g.size(); i++) g[i] = (uint16_t)(i * 2654435761u >> 16);
  uint16_t* dg;
  CK(cudaMalloc(&dg, g.size() * 2));
  CK(cudaMemcpy(dg, g.data(), g.size() * 2, cudaMemcpyHostToDevice));
  EncodeTiledFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) {
    printf("tma: no cuTensorMapEncodeTiled entry point\n");
    return 1;
  }
  CUtensorMap tmap;
  cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C8, (cuuint64_t)N};
  cuuint64_t strides[4] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)C8 * H * W * 16};
  cuuint32_t box[5] = {8, BW, BH, BC, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dg, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("tma: cuTensorMapEncodeTiled failed %d\n", (int)r);
    return 1;
  }
  const uint32_t bytes = 16 * BW * BH * BC;
  uint8_t* dout;
  int* ds;
  CK(cudaMalloc(&dout, bytes));
  CK(cudaMalloc(&ds, 4));
  int fails = 0;
  const int cases[4][4] = {{4, 3, 1, 0}, {-3, -2, 0, 1}, {18, 15, 2, 1}, {-5, 16, 1, 0}};
  for (int t = 0; t < 4; t++) {
    int x0 = cases[t][0], y0 = cases[t][1], c0 = cases[t][2], n = cases[t][3];
    CK(cudaMemset(ds, 0, 4));
    tma_probe_kernel<<<1, 128, bytes>>>(tmap, x0, y0, c0, n, bytes, dout, ds);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("tma case %d: KERNEL ERROR %s\n", t, cudaGetErrorString(e));
      return 1;
    }
    int st;
    CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
    std::vector<uint16_t> o(bytes / 2);
    CK(cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int c = 0; c < BC; c++)
      for (int y = 0; y < BH; y++)
        for (int x = 0; x < BW; x++)
          for (int e8 = 0; e8 < 8; e8++) {
            int gx = x0 + x, gy = y0 + y, gc = c0 + c;
            uint16_t want = 0;
            if (gx >= 0 && gx < W && gy >= 0 && gy < H && gc >= 0 && gc < C8)
              want = g[((((size_t)n * C8 + gc) * H + gy) * W + gx) * 8 + e8];
            uint16_t got = o[(((size_t)c * BH + y) * BW + x) * 8 + e8];
            if (want != got) bad++;
          }
    printf("tma case %d (x0=%d y0=%d c0=%d n=%d): %s bad=%d timeout=%d\n", t, x0, y0, c0, n,
           (bad || st) ? "FAIL" : "PASS", bad, st);
    fails += (bad || st) ? 1 : 0;
  }
  return fails;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    printf("usage: umma_probe <basic|swapped|conv|conv_swapped|n224|n64|n16|n256|tma|rate N reps>\n");
    return 64;
  }
  const char* t = argv[1];
  // canonical dense layout: core matrices 128 B apart along rows (SBO=128), K planes M*16 apart (LBO)
  if (!strcmp(t, "basic")) return run_mma(t, 128, 64, 0, 128 * 16, 128, 128 * 16, 128, false, 1);
  if (!strcmp(t, "swapped")) return run_mma(t, 128, 64, 0, 128 * 16, 128, 128 * 16, 128, true, 1);
  // conv-like A: halo tile 20 px wide, 28 rows -> plane stride 560*16, row-group stride 20*16, tap (ky=3,kx=2)
  if (!strcmp(t, "conv")) return run_mma(t, 128, 64, (3 * 20 + 2) * 16, 560 * 16, 20 * 16, 128 * 16, 128, false, 1);
  if (!strcmp(t, "conv_swapped")) return run_mma(t, 128, 64, (3 * 20 + 2) * 16, 560 * 16, 20 * 16, 128 * 16, 128, true, 1);
  if (!strcmp(t, "n224")) return run_mma(t, 224, 32, 0, 128 * 16, 128, 224 * 16, 128, false, 1);
  if (!strcmp(t, "n64")) return run_mma(t, 64, 64, 0, 128 * 16, 128, 64 * 16, 128, false, 1);
  if (!strcmp(t, "n16")) return run_mma(t, 16, 64, 0, 128 * 16, 128, 16 * 16, 128, false, 1);
  if (!strcmp(t, "n256")) return run_mma(t, 256, 32, 0, 128 * 16, 128, 256 * 16, 128, false, 1);
  if (!strcmp(t, "tma")) return run_tma();
  // wgrad-like: K = 16 consecutive pixels (16 B apart), MN = channels (8 per 16 B, planes SBO apart);
  // LBO = 128 B between the two 8-pixel K groups; start address offset by a pixel shift (tap)
  if (!strcmp(t, "mnmajor")) return run_mma(t, 128, 16, 0, 128, 4096, 128, 4096, false, 1, 1, 1);
  if (!strcmp(t, "mnmajor_k64")) return run_mma(t, 128, 64, 0, 128, 4096, 128, 4096, false, 1, 1, 1);
  if (!strcmp(t, "mnmajor_shift")) return run_mma(t, 64, 64, 7 * 16, 128, 4096 + 320, 128, 4096 + 320, false, 1, 1, 1);
  if (!strcmp(t, "rate2") && argc >= 5) {
    int N = atoi(argv[2]), reps = atoi(argv[3]), nacc = atoi(argv[4]);
    return run_mma("rate2", N, 64, (3 * 20 + 2) * 16, 560 * 16, 20 * 16, N * 16, 128, false, reps, nacc);
  }
  if (!strcmp(t, "rate") && argc >= 4) {
    int N = atoi(argv[2]), reps = atoi(argv[3]);
    int conv = argc >= 5 ? atoi(argv[4]) : 0;
    if (conv) return run_mma("rate_conv", N, 64, (3 * 20 + 2) * 16, 560 * 16, 20 * 16, N * 16, 128, false, reps);
    return run_mma("rate", N, 64, 0, 128 * 16, 128, N * 16, 128, false, reps);
  }
  printf("unknown test %s\n", t);
  return 64;
}
