// What slows tcgen05.mma kind::i8 M = 128 N = 8 from 47 cycles (alone) to ~84 inside the codec kernel?  An elected thread issues
// MMAs back to back (elect.sync, unrolled by 8, commit every 8) while other warps generate one kind of traffic:
//   LDS.128 readers (the float warps), cp.async.bulk streaming into shared memory (the weight rings), tcgen05.ld readers (the
//   epilogue warps), FP32 ALU work without memory traffic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_contention umma_contention.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
template <int N> __device__ __forceinline__ void umma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t accumulate) {
  constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const uint64_t da = (uint64_t)a_hi << 32 | a_lo, db = (uint64_t)b_hi << 32 | b_lo;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory"); }

struct Cfg { int lds_warps, tma, ldtm_warps, alu_warps, n16; };
__global__ void __launch_bounds__(512, 1) k(Cfg c, const unsigned char *src, int nchunks, long long *cycles, float *sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *sT = smem + 65536;            // 2 x 40 KB TMA landing zone
  float *sL = reinterpret_cast<float *>(smem + 65536 + 81920);   // 32 KB read by the LDS warps
  __shared__ uint64_t done_bar, chunk_bar[4], tma_bar[2];
  __shared__ uint32_t tmem_s;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u * (i & 3);
  for (int i = tid; i < 32768 / 4; i += blockDim.x) sL[i] = 1.0f;
  if (tid == 0) { mbar_init(&done_bar, 1); for (int i = 0; i < 4; i++) mbar_init(&chunk_bar[i], 1); mbar_init(&tma_bar[0], 1); mbar_init(&tma_bar[1], 1); stop = 0;
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
  if (warp == 15) {                              // issuer
    const uint32_t a_hi = (1024u >> 4) | (1u << 14), b_hi = (8192u >> 4) | (1u << 14);
    const uint32_t a_base = ((smem_u32(smem) & 0x3FFFF) >> 4) | (8u << 16), b_base = ((smem_u32(smem + 32768) & 0x3FFFF) >> 4) | (8u << 16);
    const bool leader = elect_one();
    long long t0 = clock64();
    for (int ch = 0; ch < nchunks; ch++) {
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
          if (c.n16) umma<16>(tmem + (kk & 1) * 16, a_base + (kk & 3) * 16 + (kk >> 2) * 1024, a_hi, b_base + kk * 16, b_hi, ch != 0);
          else umma<8>(tmem + (kk & 1) * 8, a_base + (kk & 3) * 16 + (kk >> 2) * 1024, a_hi, b_base + kk * 16, b_hi, ch != 0);
        }
        umma_commit(&chunk_bar[ch & 3]);
      }
      __syncwarp();
    }
    if (leader) umma_commit(&done_bar);
    __syncwarp();
    mbar_wait(&done_bar, 0);
    long long t1 = clock64();
    if (lane == 0) { cycles[blockIdx.x] = t1 - t0; stop = 1; }
  } else if (warp == 14 && c.tma) {              // TMA streaming, two 40 KB copies in flight
    if (lane == 0) {
      uint32_t ph[2] = {0, 0}; size_t off = 0;
      for (int s = 0; s < 2; s++) { mbar_expect_tx(&tma_bar[s], 40960); bulk_g2s(sT + s * 40960, src + off, 40960, &tma_bar[s]); off = (off + 40960) % (1 << 20); }
      int s = 0;
      while (!stop) {
        mbar_wait(&tma_bar[s], ph[s]); ph[s] ^= 1;
        mbar_expect_tx(&tma_bar[s], 40960); bulk_g2s(sT + s * 40960, src + off, 40960, &tma_bar[s]); off = (off + 40960) % (1 << 20);
        s ^= 1;
      }
      mbar_wait(&tma_bar[0], ph[0]); mbar_wait(&tma_bar[1], ph[1]);
    }
  } else if (warp < c.lds_warps) {               // the float warps' access pattern: LDS.128, 4 addresses per warp, + FMUL/FADD
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
  } else if (warp >= 4 && warp < 4 + c.ldtm_warps) {   // the epilogue warps' TMEM reads (other columns than the accumulators)
    int s = 0;
    while (!stop) {
      int v0, v1, v2, v3;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(tmem + 64 + ((uint32_t)(32 * (warp & 3)) << 16)) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      s += v0 + v3;
    }
    if (s == 0x7fffffff) sink[tid] = (float)s;
  } else if (warp >= 4 && warp < 4 + c.alu_warps) {    // pure FP32 ALU pressure
    float a = lane, b2 = 1.0001f;
    while (!stop) {
#pragma unroll 16
      for (int u = 0; u < 16; u++) { a = a * b2 + 1.f; b2 = b2 * 0.9999f + a; }
    }
    if (a == 12345.f) sink[tid] = a + b2;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128) : "memory");
}
int main() {
  unsigned char *src; cudaMalloc(&src, 2 << 20); cudaMemset(src, 1, 2 << 20);
  long long *cyc; cudaMalloc(&cyc, sizeof(long long) * 256); float *sink; cudaMalloc(&sink, 8192);
  const int smem = 65536 + 81920 + 32768;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const Cfg cfgs[] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 1}, {3, 0, 0, 0, 0}, {4, 0, 0, 0, 0}, {0, 1, 0, 0, 0}, {0, 0, 4, 0, 0}, {0, 0, 8, 0, 0}, {0, 0, 0, 8, 0}, {4, 1, 8, 0, 0}, {4, 1, 8, 0, 1}};
  printf("%-10s %-5s %-11s %-10s %-4s %14s\n", "lds warps", "tma", "ldtm warps", "alu warps", "N", "cycles per MMA");
  for (const Cfg &c : cfgs) {
    const int nchunks = 128;
    k<<<148, 512, smem>>>(c, src, nchunks, cyc, sink);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; i++) mx = h[i] > mx ? h[i] : mx;
    printf("%-10d %-5d %-11d %-10d %-4d %14.1f\n", c.lds_warps, c.tma, c.ldtm_warps, c.alu_warps, c.n16 ? 16 : 8, (double)mx / (nchunks * 8));
  }
  return 0;
}
