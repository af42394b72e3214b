// Bring-up probe: tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (written from registers with tcgen05.st).
// Hypothesis: row m of A lives in TMEM lane m, element k in column (base + k), one 32-bit column per tf32 element, and an MMA
// of K = 8 reads 8 consecutive columns.   nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_a_probe tmem_a_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

constexpr int M = 128, N = 64, K = 32;

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int variant) {
  __shared__ __align__(1024) uint8_t btile[N * 128];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B tile: K-major SWIZZLE_128B, row n = 32 floats
  for (int i = tid; i < N * 8; i += 128) {
    const int n = i >> 3, c = i & 7;
    const float4 v = *reinterpret_cast<const float4*>(B + n * K + c * 4);
    *reinterpret_cast<float4*>(btile + swz(n, c)) = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t a_col = 64;
  // A: thread (warp, lane) owns row m = 32*warp + lane; 32 k-values -> TMEM columns a_col .. a_col+31 of its lane
  {
    const int m = 32 * warp + lane;
    uint32_t r[32];
    for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(A[m * K + k]);
    for (int h = 0; h < 2; ++h) {
      const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + a_col + 16 * h;
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
          "r"(r[16 * h + 0]), "r"(r[16 * h + 1]), "r"(r[16 * h + 2]), "r"(r[16 * h + 3]), "r"(r[16 * h + 4]), "r"(r[16 * h + 5]),
          "r"(r[16 * h + 6]), "r"(r[16 * h + 7]), "r"(r[16 * h + 8]), "r"(r[16 * h + 9]), "r"(r[16 * h + 10]), "r"(r[16 * h + 11]),
          "r"(r[16 * h + 12]), "r"(r[16 * h + 13]), "r"(r[16 * h + 14]), "r"(r[16 * h + 15])
          : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    // idesc: D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t bdesc = make_smem_desc(smem_u32(btile));
    for (int j = 0; j < K / 8; ++j) {
      const uint32_t a_t = tmem + a_col + (variant == 0 ? 8 * j : 4 * j);     // variant 1: K step of 4 columns (64-bit packing?)
      const uint32_t acc = j > 0 ? 1u : 0u;
      asm volatile(
          "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem),
          "r"(a_t), "l"(bdesc + 2ull * j), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait for the MMAs
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DONE_%=;\nbra W_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(&bar))
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int m = 32 * warp + lane;
    for (int c = 0; c < N; c += 16) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) D[m * N + c + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

int main() {
  float *hA = (float*)malloc(M * K * 4), *hB = (float*)malloc(N * K * 4), *hD = (float*)malloc(M * N * 4);
  srand(1);
  for (int i = 0; i < M * K; ++i) hA[i] = (float)(rand() % 17 - 8);      // small integers: exact in tf32
  for (int i = 0; i < N * K; ++i) hB[i] = (float)(rand() % 13 - 6);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dD, 0, M * N * 4);
    probe<<<1, 128>>>(dA, dB, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
    int bad = 0; double maxerr = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)hA[m * K + k] * hB[n * K + k];
        const double err = fabs(ref - hD[m * N + n]);
        if (err > maxerr) maxerr = err;
        if (err > 1e-3) { if (bad < 4) printf("  variant %d mismatch (%d,%d): got %g want %g\n", variant, m, n, hD[m * N + n], ref); ++bad; }
      }
    printf("variant %d (K step = %d TMEM columns): %d / %d mismatches, max abs err %g\n", variant, variant == 0 ? 8 : 4, bad, M * N, maxerr);
  }
  return 0;
}
