// fp32 CUDA-core gather-GEMM (OFFK_PREC_FP32): the exact-arithmetic mode used for the
// "fp32 max-abs per level" parity figures.  Classic 64x64x16 shared-memory tiling, 4x4 register
// micro-tiles, FFMA with fp32 accumulation; operands come through the same separable index tables
// as the tensor-core kernel (offk.h).
#include "offk_gemm.cuh"

namespace offk {

constexpr int SM_BM = 64, SM_BN = 64, SM_BK = 16, SM_THREADS = 256;

__global__ void __launch_bounds__(SM_THREADS) gather_gemm_simt_kernel(const offk_gemm_t g, int k_per_split) {
  __shared__ float As[SM_BK][SM_BM + 4];
  __shared__ float Bs[SM_BK][SM_BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * SM_BM, n0 = blockIdx.y * SM_BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(g.K, k_begin + k_per_split);
  const int tx = tid & 15, ty = tid >> 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += SM_BK) {
    // ---- A tile: 64 x 16
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int row, kk;
      if (g.a_mode == OFFK_LOAD_SCALAR_K || g.a_mode == OFFK_LOAD_VEC_K) { kk = tid & 15; row = (tid >> 4) + 16 * i; }
      else           { row = tid & 63; kk = (tid >> 6) + 4 * i; }
      const int m = m0 + row, k = k0 + kk;
      float v = 0.f;
      if (m < g.M && k < k_end) {
        const bool ones = (m == g.a_ones_row);
        offk_idx_t r = ones ? offk_idx_t{0, 0, 0} : g.a_row[m];
        v = gemm_load_a(g, r, g.a_col[k], ones);
      }
      As[kk][row] = v;
    }
    // ---- B tile: 64 x 16
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int row, kk;
      if (g.b_mode == OFFK_LOAD_SCALAR_K || g.b_mode == OFFK_LOAD_VEC_K) { kk = tid & 15; row = (tid >> 4) + 16 * i; }
      else           { row = tid & 63; kk = (tid >> 6) + 4 * i; }
      const int n = n0 + row, k = k0 + kk;
      float v = 0.f;
      if (n < g.N && k < k_end) v = __ldg(g.b_src + (g.b_row[n] + g.b_col[k]));
      Bs[kk][row] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SM_BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool atomic = (g.split_k > 1) || g.atomic_out;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
    const EpiRow r = epi_row(g, m);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < g.N) epi_store(g, r, n, acc[i][j], atomic);
    }
  }
}

int launch_gemm_simt(const offk_gemm_t& g, cudaStream_t st) {
  const int split = g.split_k > 1 ? g.split_k : 1;
  int k_per = (g.K + split - 1) / split;
  k_per = (k_per + SM_BK - 1) / SM_BK * SM_BK;
  dim3 grid((g.M + SM_BM - 1) / SM_BM, (g.N + SM_BN - 1) / SM_BN, (g.K + k_per - 1) / k_per);
  if (grid.y > 65535 || grid.z > 65535) return fail(OFFK_E_LIMIT, "gather_gemm: grid too large");
  gather_gemm_simt_kernel<<<grid, SM_THREADS, 0, st>>>(g, k_per);
  return OFFK_LAUNCH_CHECK("gather_gemm_simt");
}

}  // namespace offk
