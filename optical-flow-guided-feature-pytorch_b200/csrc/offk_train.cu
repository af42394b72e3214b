// The training step around the OFF path, on the flat parameter / gradient buffers:
//   offk_ce_loss_fwd_bwd   nn.CrossEntropyLoss over [P, C] logits with the clip label repeated per frame pair
//                          (train_off.py:72,133-146): loss and dL/dlogits in one pass, one warp per row
//   offk_grad_sumsq        sum of squares of the gradients (the global norm of clip_grad_norm, train_off.py:149)
//   offk_clip_adam_step    clip coefficient from that norm + optim.Adam(lr, betas, weight_decay) (train_off.py:72,151) in
//                          one pass over p, g, m, v -- no host round trip: the norm is read from device memory
// All three are plain streaming kernels (HBM-bound, 16-byte accesses); nothing here allocates or synchronises.
#include <math.h>
#include "offk_common.cuh"

namespace offk {

// ---------------------------------------------------------------------------- cross entropy
// row p of logits [P, C]; label = target[p / repeat] (the clip's label for each of its `repeat` frame pairs)
__global__ void __launch_bounds__(256) ce_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ target, int P,
                                                       int C, int repeat, float grad_scale, float inv_rows,
                                                       float* __restrict__ loss_accum, float* __restrict__ dlogits) {
  pdl_sync();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P) return;
  const float* row = logits + (size_t)warp * C;
  const int label = (int)target[warp / repeat];
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, row[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += expf(row[c] - mx);
  sum = warp_sum(sum);
  const float lse = mx + logf(sum);
  if (dlogits) {
    float* drow = dlogits + (size_t)warp * C;
    const float s = grad_scale * inv_rows;
    for (int c = lane; c < C; c += 32) drow[c] = s * (expf(row[c] - lse) - (c == label ? 1.f : 0.f));
  }
  if (lane == 0 && loss_accum) atomicAdd(loss_accum, (lse - row[label]) * inv_rows);
}

// ---------------------------------------------------------------------------- global gradient norm
__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, long long n4, long long n,
                                                          double* __restrict__ out) {
  pdl_sync();
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - 4 * n4)) {   // tail (n % 4 elements)
    const float v = g[4 * n4 + threadIdx.x];
    acc += v * v;
  }
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += (double)part[w];
    atomicAdd(out, s);
  }
}

// ---------------------------------------------------------------------------- clip + Adam
struct AdamRanges {
  long long lo[OFFK_ADAM_MAX_RANGES], hi[OFFK_ADAM_MAX_RANGES];   // element ranges [lo, hi), multiples of 4
  int n;
};

__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, const AdamRanges r, float lr, float b1, float b2,
                                                         float eps, float wd, float bc1, float bc2_sqrt, float max_norm,
                                                         const double* __restrict__ sumsq) {
  pdl_sync();
  // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), applied only when < 1
  float coef = 1.f;
  if (sumsq && max_norm > 0.f) {
    const float total = (float)sqrt(*sumsq);
    coef = fminf(max_norm / (total + 1e-6f), 1.f);
  }
  const float step_size = lr / bc1;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (int k = 0; k < r.n; ++k) {
    for (long long i = r.lo[k] + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < r.hi[k]; i += stride) {
      const float4 g4 = *reinterpret_cast<const float4*>(g + i);
      float4 p4 = *reinterpret_cast<float4*>(p + i), m4 = *reinterpret_cast<float4*>(m + i), v4 = *reinterpret_cast<float4*>(v + i);
      float gg[4] = {g4.x, g4.y, g4.z, g4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w},
            vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gr = gg[j] * coef + wd * pp[j];             // optim.Adam weight_decay: L2 term added to the gradient
        mm[j] = b1 * mm[j] + (1.f - b1) * gr;
        vv[j] = b2 * vv[j] + (1.f - b2) * gr * gr;
        pp[j] -= step_size * mm[j] / (sqrtf(vv[j]) / bc2_sqrt + eps);
      }
      *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
      *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
      *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
  }
}

}  // namespace offk

using namespace offk;

extern "C" int offk_ce_loss_fwd_bwd(const float* logits, const long long* target, int P, int C, int repeat, float grad_scale,
                                    float* loss_accum, float* dlogits, void* stream) {
  OFFK_REQUIRE(logits && target && P > 0 && C > 0 && repeat > 0, "ce_loss: bad arguments");
  const int rows_per_block = 256 / 32;
  cudaError_t e = launch_pdl(ce_loss_kernel, dim3((P + rows_per_block - 1) / rows_per_block), dim3(256), 0, as_stream(stream), logits,
                             target, P, C, repeat, grad_scale, 1.0f / (float)P, loss_accum, dlogits);
  if (e != cudaSuccess) return cuda_check(e, "ce_loss launch");
  return OFFK_LAUNCH_CHECK("ce_loss");
}

extern "C" int offk_grad_sumsq(const float* g, long long n, double* sumsq_accum, void* stream) {
  OFFK_REQUIRE(g && sumsq_accum && n > 0 && (reinterpret_cast<uintptr_t>(g) & 15u) == 0, "grad_sumsq: bad arguments");
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cudaError_t e = launch_pdl(grad_sumsq_kernel, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream), g, n4, n, sumsq_accum);
  if (e != cudaSuccess) return cuda_check(e, "grad_sumsq launch");
  return OFFK_LAUNCH_CHECK("grad_sumsq");
}

extern "C" int offk_clip_adam_step(float* p, const float* g, float* m, float* v, const long long* range_lo, const long long* range_hi,
                                   int n_ranges, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                                   float max_norm, const double* sumsq, void* stream) {
  OFFK_REQUIRE(p && g && m && v && range_lo && range_hi && n_ranges >= 1 && n_ranges <= OFFK_ADAM_MAX_RANGES && step >= 1,
               "clip_adam_step: bad arguments");
  OFFK_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v)) & 15u) == 0, "clip_adam_step: buffers must be 16-byte aligned");
  AdamRanges r;
  r.n = n_ranges;
  long long total = 0;
  for (int k = 0; k < n_ranges; ++k) {
    OFFK_REQUIRE(range_lo[k] % 4 == 0 && range_hi[k] % 4 == 0 && range_lo[k] <= range_hi[k], "clip_adam_step: ranges must be multiples of 4 elements");
    r.lo[k] = range_lo[k]; r.hi[k] = range_hi[k];
    total += range_hi[k] - range_lo[k];
  }
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  long long blocks = (total / 4 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cudaError_t e = launch_pdl(clip_adam_kernel, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream), p, g, m, v, r, lr, beta1, beta2, eps,
                             weight_decay, bc1, sqrtf(bc2), max_norm, sumsq);
  if (e != cudaSuccess) return cuda_check(e, "clip_adam_step launch");
  return OFFK_LAUNCH_CHECK("clip_adam_step");
}
