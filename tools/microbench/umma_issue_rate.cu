// What slows tcgen05.mma kind::i8 (M = 128, N = 8) below its 47 cycles per instruction in the codec kernel?  One issuer thread
// issues NMMA MMAs over a shared-memory-resident A image (canonical no-swizzle layout) and measures cycles per MMA until the final
// commit completes, under: commits every C MMAs; accumulator column offsets; runtime (non-unrolled) descriptors; concurrent
// LDS.128 traffic from other warps; concurrent cp.async.bulk streaming into another region of shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_issue_rate umma_issue_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { while (!mbar_try(bar, parity)) {} }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t sbo_bytes) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (uint64_t)(128 >> 4) << 16 | (uint64_t)(sbo_bytes >> 4) << 32 | (uint64_t)1 << 46;
}
__device__ __forceinline__ void umma_i8_n8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((8u >> 3) << 17) | ((128u >> 4) << 24);
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory"); }

struct Cfg { int nmma, commit_every, col_a, col_b, lds_warps, tma, wait_each_commit, lean, tiles, sbo; };

__global__ void __launch_bounds__(384, 1) k(Cfg c, const unsigned char *src, long long *cycles, float *sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *sA = smem;                    // 32 KB: 256 rows x 128 B of K (4 k-blocks), SBO = 1024
  unsigned char *sB = smem + 32768;            // 8 streams x 1024 B of K
  unsigned char *sT = smem + 65536;            // 2 x 32 KB TMA landing zone
  float *sL = reinterpret_cast<float *>(smem + 131072);   // 32 KB read by the LDS warps
  __shared__ uint64_t done_bar, ring_bar[8], tma_bar[2];
  __shared__ uint32_t tmem_s;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u * (i & 3);
  for (int i = tid; i < 32768 / 4; i += blockDim.x) sL[i] = 1.0f;
  if (tid == 0) { mbar_init(&done_bar, 1); for (int i = 0; i < 8; i++) mbar_init(&ring_bar[i], 1); mbar_init(&tma_bar[0], 1); mbar_init(&tma_bar[1], 1); stop = 0;
                  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_s)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  if (warp == 11 && c.lean) {                    // lean issuer: warp-uniform loop, one elected lane issues
    const uint32_t a_hi = (uint32_t)(c.sbo >> 4) | (1u << 14), b_hi = (8192u >> 4) | (1u << 14);
    const uint32_t a_base = ((smem_u32(sA) & 0x3FFFF) >> 4) | (8u << 16), b_base = ((smem_u32(sB) & 0x3FFFF) >> 4) | (8u << 16);
    const uint32_t tile_off = (16u * c.sbo) >> 4;
    int nc = 0;
    long long t0 = clock64();
    for (int i = 0; i < c.nmma; i += c.tiles) {
      const uint32_t a_lo = a_base + ((i / c.tiles) & 3) * 16, b_lo = b_base + ((i / c.tiles) & 31) * 16;
      if (lane == 0) {
        umma_i8_n8(tmem + c.col_a, (uint64_t)a_hi << 32 | a_lo, (uint64_t)b_hi << 32 | b_lo, i >= 8);
        if (c.tiles > 1) umma_i8_n8(tmem + c.col_b, (uint64_t)a_hi << 32 | (a_lo + tile_off), (uint64_t)b_hi << 32 | b_lo, i >= 8);
      }
      if ((i + c.tiles) % c.commit_every == 0) { if (lane == 0) umma_commit(&ring_bar[nc & 7]); nc++; }
      __syncwarp();
    }
    if (lane == 0) umma_commit(&done_bar);
    __syncwarp();
    mbar_wait(&done_bar, 0);
    long long t1 = clock64();
    if (lane == 0) { cycles[blockIdx.x] = t1 - t0; stop = 1; }
  } else if (warp == 11) {                       // naive issuer: everything inside `if (lane == 0)`
    if (lane == 0) {
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      uint32_t phase[8] = {0, 0, 0, 0, 0, 0, 0, 0}; int nc = 0;
      long long t0 = clock64();
      for (int i = 0; i < c.nmma; i++) {
        const int k = i & 3, g = (i >> 2) & 1;   // 4 k-blocks, two tiles (rows 0 and 128) -> accumulators col_a / col_b
        umma_i8_n8(tmem + (g ? c.col_b : c.col_a), umma_desc(a0 + g * 16 * 1024 + k * 256, 1024), umma_desc(b0 + (i & 31) * 256, 8192), i >= 8);
        if ((i + 1) % c.commit_every == 0) {
          const int s = nc & 7;
          umma_commit(&ring_bar[s]);
          if (c.wait_each_commit) { mbar_wait(&ring_bar[s], phase[s]); phase[s] ^= 1; }
          nc++;
        }
      }
      umma_commit(&done_bar);
      mbar_wait(&done_bar, 0);
      long long t1 = clock64();
      cycles[blockIdx.x] = t1 - t0;
      stop = 1;
    }
  } else if (warp == 10 && c.tma) {              // TMA streaming into the landing zone, two copies of 32 KB in flight
    if (lane == 0) {
      uint32_t ph[2] = {0, 0}; size_t off = 0;
      for (int s = 0; s < 2; s++) { mbar_expect_tx(&tma_bar[s], 32768); bulk_g2s(sT + s * 32768, src + off, 32768, &tma_bar[s]); off = (off + 32768) & ((1 << 20) - 1); }
      int s = 0;
      while (!stop) {
        mbar_wait(&tma_bar[s], ph[s]); ph[s] ^= 1;
        mbar_expect_tx(&tma_bar[s], 32768); bulk_g2s(sT + s * 32768, src + off, 32768, &tma_bar[s]); off = (off + 32768) & ((1 << 20) - 1);
        s ^= 1;
      }
      mbar_wait(&tma_bar[0], ph[0]); mbar_wait(&tma_bar[1], ph[1]);
    }
  } else if (warp < c.lds_warps) {               // LDS.128 + FMUL/FADD traffic like the codec's float warps
    float acc[4] = {0, 0, 0, 0};
    const float4 *W4 = reinterpret_cast<const float4 *>(sL) + (lane >> 3);
    int j = 0;
    while (!stop) {
#pragma unroll 8
      for (int u = 0; u < 8; u++) {
        const float4 w = W4[((j + u) * 20) & 2047];
        acc[0] += w.x * 1.0001f; acc[1] += w.y * 1.0001f; acc[2] += w.z * 1.0001f; acc[3] += w.w * 1.0001f;
      }
      j += 8;
    }
    if (acc[0] == 12345.f) sink[tid] = acc[0] + acc[1] + acc[2] + acc[3];
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128) : "memory");
}
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int nsm = prop.multiProcessorCount;
  unsigned char *src; cudaMalloc(&src, 2 << 20); cudaMemset(src, 1, 2 << 20);
  long long *cyc; cudaMalloc(&cyc, sizeof(long long) * nsm); float *sink; cudaMalloc(&sink, 4096);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
  long long h[256];
  const Cfg cfgs[] = {
    {512, 8, 0, 8, 0, 0, 0, 0, 2, 1024},   // naive issuer (descriptors computed inside `if (lane == 0)`)
    {512, 512, 0, 8, 0, 0, 0, 1, 1, 2048}, // lean issuer, one tile, one accumulator, SBO 2048, one commit at the end
    {512, 512, 0, 8, 0, 0, 0, 1, 1, 1024},
    {512, 512, 0, 8, 0, 0, 0, 1, 1, 512},
    {512, 512, 0, 8, 0, 0, 0, 1, 2, 1024}, // two tiles per k-block -> two accumulators
    {512, 8, 0, 8, 0, 0, 0, 1, 2, 1024},   // + a commit every 8 MMAs
    {512, 8, 0, 8, 0, 0, 0, 1, 1, 2048},
    {512, 8, 0, 8, 5, 0, 0, 1, 2, 1024},   // + 5 warps of LDS.128 traffic
    {512, 8, 0, 8, 10, 0, 0, 1, 2, 1024},
    {512, 8, 0, 8, 0, 1, 0, 1, 2, 1024},   // + TMA streaming
    {512, 8, 0, 8, 5, 1, 0, 1, 2, 1024},   // + both
  };
  printf("%-6s %-7s %-5s %-5s %-5s %-5s %-5s %-5s %12s\n", "nmma", "commit", "lean", "tiles", "sbo", "lds", "tma", "grid", "cycles/MMA");
  for (const Cfg &c : cfgs)
    for (int grid : {nsm}) {
      k<<<grid, 384, 164 * 1024>>>(c, src, cyc, sink);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
      printf("%-6d %-7d %-5d %-5d %-5d %-5d %-5d %-5d %12.1f\n", c.nmma, c.commit_every, c.lean, c.tiles, c.sbo, c.lds_warps, c.tma, grid, (double)mx / c.nmma);
    }
  return 0;
}
