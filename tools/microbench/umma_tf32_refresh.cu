// Prototype of rx_track's 48-row |Dt| refresh (acquisition.check_pilots, /root/reference/radae/dsp.py:291-300) as a
// 3xTF32 tensor-core GEMM (round-2 plan), self-checking and timed.
//
//   Dt[t, f] = sum_n conj(rx[t + n]) p_w[n, f],  n < 160, f < 40, for 48 timing rows t at two pilot positions (t, t + 960);
//   the kernel keeps only rowsum[t] = sum_f |Dt[t, f]| (what sigma_r is built from).
// As a real GEMM, D[128][96] = A[128][320] x B[96][320]^T with
//   B row j  = [ Re rx[t_j + n] | Im rx[t_j + n] ]                     (96 = 48 rows x 2 pilot positions)
//   A row f      = [ Re p_w[n, f] |  Im p_w[n, f] ]  -> Re Dt          (conj(x) p = (xr pr + xi pi) + j (xr pi - xi pr))
//   A row 40 + f = [ Im p_w[n, f] | -Re p_w[n, f] ]  -> Im Dt          rows 80..127 are zero
// tcgen05.mma kind::tf32 reads fp32 words and keeps 10 mantissa bits, so every operand is split hi + lo (hi = the word with
// the low 13 bits cleared, lo = the exact remainder) and D = Ah Bh + Ah Bl + Al Bh accumulates in fp32 in TMEM: row sums
// within ~4e-7 of double precision on real data (tools/tf32_refresh_study.py), the same class as today's fp32 FFMA2 path.
// K is processed in 5 chunks of 64 floats; A chunks come pre-baked (canonical K-major layout) from global memory, B chunks
// are built from the stream's samples by the CTA.  One CTA = one stream; nothing is pipelined yet — this is the numerics /
// layout / cost check, not the final kernel.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o umma_tf32_refresh umma_tf32_refresh.cu
//   timeout 60 ./umma_tf32_refresh
// Written at the end of round 1 without GPU time left: compiled (SASS shows UTCHMMA), not yet run.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int NTAP = 160, NF = 40, NROW = 48, NPOS = 2, NMF = 960;
constexpr int M = 128, N = NROW * NPOS, K = 2 * NTAP;          // 128 x 96 x 320
constexpr int KC = 64, NCHUNK = K / KC, KCB = KC * 4;          // 64 floats = 256 bytes of K per chunk, 8 MMAs of K = 8
constexpr int TMEM_COLS = 128;
constexpr int RXBUF = NMF + 20 * NROW + NTAP + NMF;            // enough samples for every row at both positions

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ constexpr int canon(int r, int b, int kbytes) { return (r / 8) * (kbytes * 8) + (b / 16) * 128 + (r % 8) * 16 + b % 16; }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (uint64_t)(lbo_bytes >> 4) << 16 | (uint64_t)(sbo_bytes >> 4) << 32 | (uint64_t)1 << 46;
}
__host__ __device__ constexpr uint32_t instr_desc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  for (long long spin = 0; !ok; spin++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (spin > (1ll << 24)) __trap();
  }
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__host__ __device__ inline float tf32_hi(float v) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
#else
  uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; memcpy(&v, &u, 4); return v;
#endif
}

// A_hi / A_lo: [NCHUNK][M x KCB bytes] pre-baked canonical chunks; rx: [RXBUF] complex samples of one stream; t0: first row
// (rows t0 + 20 i); rowsum: [N] outputs (position-major: j = pos * 48 + i)
__global__ void __launch_bounds__(128, 1)
refresh_tf32(const uint8_t *__restrict__ A_hi, const uint8_t *__restrict__ A_lo, const float2 *__restrict__ rx_all, int t0,
             float *__restrict__ rowsum_all, long long *__restrict__ cyc) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *sAh = smem, *sAl = sAh + M * KCB, *sBh = sAl + M * KCB, *sBl = sBh + N * KCB;    // 32 + 32 + 24 + 24 KB
  float *sD = reinterpret_cast<float *>(sBl + N * KCB);                                  // [80][N + 1] for the |.| stage
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;
  const float2 *rx = rx_all + (size_t)blockIdx.x * RXBUF;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  constexpr uint32_t idesc = instr_desc_tf32(M, N);
  uint32_t phase = 0;
  long long t_start = clock64();

  for (int c = 0; c < NCHUNK; c++) {
    // stage this K chunk: A linear copies (16 B per thread per step), B built from the samples (k < 160: real parts, else imag)
    const uint4 *gh = reinterpret_cast<const uint4 *>(A_hi + (size_t)c * M * KCB), *gl = reinterpret_cast<const uint4 *>(A_lo + (size_t)c * M * KCB);
    for (int i = tid; i < M * KCB / 16; i += 128) { reinterpret_cast<uint4 *>(sAh)[i] = gh[i]; reinterpret_cast<uint4 *>(sAl)[i] = gl[i]; }
    for (int i = tid; i < N * KC; i += 128) {
      const int j = i / KC, kk = i % KC, k = c * KC + kk;                 // j = pos * 48 + row index
      const int t = t0 + 20 * (j % NROW) + NMF * (j / NROW);
      const float2 s = rx[t + (k % NTAP)];
      const float v = k < NTAP ? s.x : s.y;
      const float h = tf32_hi(v);
      *reinterpret_cast<float *>(sBh + canon(j, 4 * kk, KCB)) = h;
      *reinterpret_cast<float *>(sBl + canon(j, 4 * kk, KCB)) = v - h;    // exact; the tensor core keeps its top 10 mantissa bits
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < KC / 8; ks++) {
        const uint32_t off = ks * 256;                                    // 8 floats = 32 bytes = two 16-byte core-matrix columns
        const uint64_t ah = smem_desc(smem_u32(sAh) + off, 128, KCB * 8), al = smem_desc(smem_u32(sAl) + off, 128, KCB * 8);
        const uint64_t bh = smem_desc(smem_u32(sBh) + off, 128, KCB * 8), bl = smem_desc(smem_u32(sBl) + off, 128, KCB * 8);
        umma_tf32(tmem, ah, bh, idesc, (c | ks) != 0);
        umma_tf32(tmem, ah, bl, idesc, 1);
        umma_tf32(tmem, al, bh, idesc, 1);
      }
      umma_commit(&bar);                                                  // the staging buffers are free again when this fires
    }
    mbar_wait(&bar, phase); phase ^= 1;                                   // (un-pipelined prototype: everybody waits per chunk)
    __syncthreads();
  }
  // epilogue: lanes 0..39 hold Re Dt, 40..79 Im Dt; meet in shared memory, then one thread per column sums |Dt| over f
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 3) {
    for (int col = 0; col < N; col += 8) {
      float v[8];
      tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + col, v);
      if (tid < 2 * NF)
#pragma unroll
        for (int q = 0; q < 8; q++) sD[tid * (N + 1) + col + q] = v[q];
    }
  }
  __syncthreads();
  if (tid < N) {
    float acc = 0.f;
    for (int f = 0; f < NF; f++) { const float re = sD[f * (N + 1) + tid], im = sD[(NF + f) * (N + 1) + tid]; acc += sqrtf(re * re + im * im); }
    rowsum_all[(size_t)blockIdx.x * N + tid] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0 && blockIdx.x == 0) *cyc = clock64() - t_start;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TMEM_COLS) : "memory");
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  if (prop.major != 10) { printf("needs sm_100 (found sm_%d%d)\n", prop.major, prop.minor); return 1; }
  const int S = prop.multiProcessorCount * 7;                     // ~1024 streams: 7 per SM, one CTA each
  srand(3);
  auto frand = [] { return (float)rand() / (float)RAND_MAX * 2.f - 1.f; };
  // p_w[n][f] = p[n] exp(-j 2 pi f_k n / Fs), f_k = -50 + 2.5 k Hz, unit-magnitude stand-in pilot sequence
  std::vector<float> pr(NTAP * NF), pi(NTAP * NF);
  for (int n = 0; n < NTAP; n++) {
    const double ph = 6.283185307179586 * rand() / RAND_MAX;
    for (int f = 0; f < NF; f++) {
      const double a = ph - 6.283185307179586 * (-50.0 + 2.5 * f) * n / 8000.0;
      pr[n * NF + f] = (float)cos(a); pi[n * NF + f] = (float)sin(a);
    }
  }
  // A [M][K] then hi / lo canonical chunks
  std::vector<float> A((size_t)M * K, 0.f);
  for (int f = 0; f < NF; f++) for (int n = 0; n < NTAP; n++) {
    A[(size_t)f * K + n] = pr[n * NF + f];        A[(size_t)f * K + NTAP + n] = pi[n * NF + f];
    A[(size_t)(NF + f) * K + n] = pi[n * NF + f]; A[(size_t)(NF + f) * K + NTAP + n] = -pr[n * NF + f];
  }
  std::vector<uint8_t> Ah((size_t)NCHUNK * M * KCB), Al(Ah.size());
  for (int c = 0; c < NCHUNK; c++) for (int r = 0; r < M; r++) for (int kk = 0; kk < KC; kk++) {
    const float v = A[(size_t)r * K + c * KC + kk], h = tf32_hi(v), l = v - h;
    memcpy(&Ah[(size_t)c * M * KCB + canon(r, 4 * kk, KCB)], &h, 4);
    memcpy(&Al[(size_t)c * M * KCB + canon(r, 4 * kk, KCB)], &l, 4);
  }
  std::vector<float2> rx((size_t)S * RXBUF);
  for (auto &s : rx) { s.x = frand(); s.y = frand(); }
  const int t0 = 7;
  // CPU reference (double) for stream 0 and the last stream
  auto reference = [&](int s, std::vector<double> &out) {
    out.assign(N, 0.0);
    for (int j = 0; j < N; j++) {
      const int t = t0 + 20 * (j % NROW) + NMF * (j / NROW);
      for (int f = 0; f < NF; f++) {
        double re = 0, im = 0;
        for (int n = 0; n < NTAP; n++) {
          const double xr = rx[(size_t)s * RXBUF + t + n].x, xi = rx[(size_t)s * RXBUF + t + n].y;
          re += xr * pr[n * NF + f] + xi * pi[n * NF + f]; im += xr * pi[n * NF + f] - xi * pr[n * NF + f];
        }
        out[j] += sqrt(re * re + im * im);
      }
    }
  };
  uint8_t *dAh, *dAl; float2 *drx; float *drs; long long *dc;
  CK(cudaMalloc(&dAh, Ah.size())); CK(cudaMalloc(&dAl, Al.size())); CK(cudaMalloc(&drx, rx.size() * 8)); CK(cudaMalloc(&drs, (size_t)S * N * 4)); CK(cudaMalloc(&dc, 8));
  CK(cudaMemcpy(dAh, Ah.data(), Ah.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dAl, Al.data(), Al.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(drx, rx.data(), rx.size() * 8, cudaMemcpyHostToDevice));
  const int smem_bytes = 2 * M * KCB + 2 * N * KCB + 2 * NF * (N + 1) * 4;
  CK(cudaFuncSetAttribute(refresh_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  refresh_tf32<<<S, 128, smem_bytes>>>(dAh, dAl, drx, t0, drs, dc);      // warm-up
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  refresh_tf32<<<S, 128, smem_bytes>>>(dAh, dAl, drx, t0, drs, dc);
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<float> rs((size_t)S * N); long long cyc;
  CK(cudaMemcpy(rs.data(), drs, rs.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
  double worst = 0;
  for (int s : {0, S - 1}) {
    std::vector<double> ref; reference(s, ref);
    for (int j = 0; j < N; j++) worst = fmax(worst, fabs(rs[(size_t)s * N + j] - ref[j]) / ref[j]);
  }
  printf("3xTF32 refresh: %d streams, worst row-sum error %.2e relative (fp32 FFMA2 path: ~7e-8, bar for the plan: < 1e-6) -> %s\n", S, worst,
         worst < 1e-6 ? "OK" : "TOO COARSE");
  printf("  %.1f us for %d streams (%d CTAs of one stream, un-pipelined), %lld cycles per stream in CTA 0; today's FFMA2 refresh: ~100 us\n",
         ms * 1000.f, S, S, cyc);
  return worst < 1e-6 ? 0 : 1;
}
