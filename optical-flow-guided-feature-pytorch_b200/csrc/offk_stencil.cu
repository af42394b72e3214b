// Fused OFF stencil kernels (forward and backward), HBM-bound by construction.
//
// Forward (one launch per OFF unit):
//   temporal CTAs : thread owns up to 4 float4 positions of the contiguous [Cg*H*W] span of a clip and walks
//                   t = 0..L-1 keeping the previous frame in registers, so every G frame is read exactly once and
//                   every difference G(t+1)-G(t) is written once, straight into the stage-fusion buffer.
//   spatial CTAs  : (pair, 8-channel group): the D planes are staged in shared memory with a zero halo (that IS the
//                   zero padding), each thread then produces 4 consecutive outputs per step with the 3x3 taps held
//                   in registers, applies bias and dropout and stores float4 at the unit's channel offset.
// Both halves write straight into the concatenated [P, Ctot, H, W] stage buffer: no torch.cat, no sub, no conv2d.
//
// Backward mirrors it: dG = (dT(t-1) - dT(t)) * [G > 0], dD = transposed stencil of the (dropped) spatial gradient,
// and the learned-tap / bias gradients are reduced with warp shuffles and one atomicAdd per warp.
#include "offk_common.cuh"

namespace offk {

constexpr int ST_THREADS = 256;
constexpr int ST_TPOS = 4;    // float4 positions per thread (temporal half)
constexpr int ST_CG = 8;      // channels per spatial CTA (forward)
constexpr int ST_CGB = 4;     // channels per spatial CTA (backward: two staged tiles)
constexpr int ST_MAXW = 32;   // spatial planes up to 32x32 (28/14/7 on the path)

__device__ __forceinline__ float4 f4sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

__device__ __forceinline__ int spatial_frame(const offk_stencil_t& s, int p) {
  if (s.index_mode == OFFK_INDEX_REFERENCE_FLAT) return p;
  const int b = p / (s.L - 1), t = p - b * (s.L - 1);
  return b * s.L + t;
}
// pair fed by frame f through the spatial branch, or -1
__device__ __forceinline__ int spatial_pair_of_frame(const offk_stencil_t& s, int f) {
  const int P = s.B * (s.L - 1);
  if (s.index_mode == OFFK_INDEX_REFERENCE_FLAT) return f < P ? f : -1;
  const int b = f / s.L, t = f - b * s.L;
  return t < s.L - 1 ? b * (s.L - 1) + t : -1;
}
__device__ __forceinline__ float drop_factor(const offk_stencil_t& s, uint32_t thr, size_t idx) {
  if (s.drop_mode == OFFK_DROP_NONE) return 1.f;
  const bool keep = s.drop_mode == OFFK_DROP_MASK ? (s.keep_mask[idx] != 0) : drop_keep(s.seed, idx, thr);
  return keep ? s.keep_scale : 0.f;
}

// ---------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(ST_THREADS)
stencil_diff_fwd_kernel(const offk_stencil_t s, const float* __restrict__ g, const float* __restrict__ d,
                        const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
                        int t_blocks_per_clip, int n_tblocks) {
  const int HW = s.H * s.W;
  const int tid = threadIdx.x;
  if ((int)blockIdx.x < n_tblocks) {
    // ------------------------------ temporal difference (RGB_OFF.py:599-604)
    const int b = blockIdx.x / t_blocks_per_clip;
    const int blk = blockIdx.x - b * t_blocks_per_clip;
    const int span4 = (s.Cg * HW) >> 2;
    const int base4 = blk * (ST_THREADS * ST_TPOS) + tid;
    const float* gb = g + (size_t)b * s.L * s.g_fs;
    float* ob = out + ((size_t)b * (s.L - 1) * s.out_ctot + s.out_coff + s.K * s.Cs) * HW;
    const size_t o_ps = (size_t)s.out_ctot * HW;
    float4 prev[ST_TPOS];
#pragma unroll
    for (int u = 0; u < ST_TPOS; ++u) {
      const int e = base4 + u * ST_THREADS;
      if (e < span4) prev[u] = ldg_stream4(gb + 4 * (size_t)e);
    }
    for (int t = 1; t < s.L; ++t) {
      float4 cur[ST_TPOS];
#pragma unroll
      for (int u = 0; u < ST_TPOS; ++u) {
        const int e = base4 + u * ST_THREADS;
        if (e < span4) cur[u] = ldg_stream4(gb + (size_t)t * s.g_fs + 4 * (size_t)e);
      }
#pragma unroll
      for (int u = 0; u < ST_TPOS; ++u) {
        const int e = base4 + u * ST_THREADS;
        if (e < span4) {
          stg_stream4(ob + (size_t)(t - 1) * o_ps + 4 * (size_t)e, f4sub(cur[u], prev[u]));
          prev[u] = cur[u];
        }
      }
    }
    return;
  }
  // ------------------------------ spatial gradient (RGB_OFF.py:611 / Flow_OFF.py:622 / util.py:46-50) + dropout
  __shared__ float tile[ST_CG][ST_MAXW + 2][ST_MAXW + 3];
  const int sblk = blockIdx.x - n_tblocks;
  const int groups = (s.Cs + ST_CG - 1) / ST_CG;
  const int p = sblk / groups, c0 = (sblk - p * groups) * ST_CG;
  const int nc = min(ST_CG, s.Cs - c0);
  const int f = spatial_frame(s, p);
  const float* dp = d + (size_t)f * s.d_fs + (size_t)c0 * HW;
  const int PH = s.H + 2, PW = s.W + 2;
  // zero halo + interior in one pass over the padded tile
  for (int i = tid; i < nc * PH * PW; i += ST_THREADS) {
    const int c = i / (PH * PW), r = i - c * (PH * PW);
    const int yy = r / PW, xx = r - yy * PW;
    float v = 0.f;
    if (yy >= 1 && yy <= s.H && xx >= 1 && xx <= s.W) v = __ldg(dp + (size_t)c * HW + (yy - 1) * s.W + (xx - 1));
    tile[c][yy][xx] = v;
  }
  __syncthreads();
  const uint32_t thr = drop_threshold24(s.drop_p);
  const int KC = s.K * s.Cs;
  const int hw4 = HW >> 2;  // HW % 4 == 0 checked by the host for the float4 path; else scalar path below
  const bool vec = (HW & 3) == 0;
  for (int kk = 0; kk < s.K; ++kk) {
    if (vec) {
      for (int i = tid; i < nc * hw4; i += ST_THREADS) {
        const int c = i / hw4, q = i - c * hw4;
        const float* wk = w + ((size_t)(c0 + c) * s.K + kk) * 9;
        float wr[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) wr[j] = __ldg(wk + j);
        const float bv = bias ? __ldg(bias + kk * s.Cs + c0 + c) : 0.f;
        const int och = kk * s.Cs + c0 + c;
        float r4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int pix = q * 4 + e;
          const int y = pix / s.W, x = pix - y * s.W;
          float acc = bv;
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) acc = fmaf(wr[a * 3 + bb], tile[c][y + a][x + bb], acc);
          const size_t midx = ((size_t)p * KC + och) * HW + pix;
          r4[e] = acc * drop_factor(s, thr, midx);
        }
        float* op = out + ((size_t)p * s.out_ctot + s.out_coff + och) * HW + q * 4;
        stg_stream4(op, make_float4(r4[0], r4[1], r4[2], r4[3]));
      }
    } else {
      for (int i = tid; i < nc * HW; i += ST_THREADS) {
        const int c = i / HW, pix = i - c * HW;
        const float* wk = w + ((size_t)(c0 + c) * s.K + kk) * 9;
        const int y = pix / s.W, x = pix - y * s.W;
        float acc = bias ? __ldg(bias + kk * s.Cs + c0 + c) : 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int bb = 0; bb < 3; ++bb) acc = fmaf(__ldg(wk + a * 3 + bb), tile[c][y + a][x + bb], acc);
        const int och = kk * s.Cs + c0 + c;
        const size_t midx = ((size_t)p * KC + och) * HW + pix;
        out[((size_t)p * s.out_ctot + s.out_coff + och) * HW + pix] = acc * drop_factor(s, thr, midx);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- backward
__global__ void __launch_bounds__(ST_THREADS)
stencil_diff_bwd_kernel(const offk_stencil_t s, const float* __restrict__ dout, const float* __restrict__ g,
                        const float* __restrict__ d, const float* __restrict__ w, float* __restrict__ dg,
                        long long dg_fs, float* __restrict__ dd, long long dd_fs, float* __restrict__ dw,
                        float* __restrict__ dbias, int t_blocks_per_clip, int n_tblocks) {
  const int HW = s.H * s.W;
  const int tid = threadIdx.x;
  if ((int)blockIdx.x < n_tblocks) {
    // ------------------------------ dG[b,t] = (dT[b,t-1] - dT[b,t]) * (G[b,t] > 0)
    const int b = blockIdx.x / t_blocks_per_clip;
    const int blk = blockIdx.x - b * t_blocks_per_clip;
    const int span4 = (s.Cg * HW) >> 2;
    const int base4 = blk * (ST_THREADS * ST_TPOS) + tid;
    const float* gb = g + (size_t)b * s.L * s.g_fs;
    float* dgb = dg + (size_t)b * s.L * dg_fs;
    const float* ob = dout + ((size_t)b * (s.L - 1) * s.out_ctot + s.out_coff + s.K * s.Cs) * HW;
    const size_t o_ps = (size_t)s.out_ctot * HW;
    float4 prev[ST_TPOS];
#pragma unroll
    for (int u = 0; u < ST_TPOS; ++u) prev[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < s.L; ++t) {
      float4 cur[ST_TPOS], gv[ST_TPOS];
#pragma unroll
      for (int u = 0; u < ST_TPOS; ++u) {
        const int e = base4 + u * ST_THREADS;
        cur[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < span4) {
          if (t < s.L - 1) cur[u] = ldg_stream4(ob + (size_t)t * o_ps + 4 * (size_t)e);
          gv[u] = ldg_stream4(gb + (size_t)t * s.g_fs + 4 * (size_t)e);
        }
      }
#pragma unroll
      for (int u = 0; u < ST_TPOS; ++u) {
        const int e = base4 + u * ST_THREADS;
        if (e < span4) {
          float4 r = f4sub(prev[u], cur[u]);
          r.x = gv[u].x > 0.f ? r.x : 0.f;
          r.y = gv[u].y > 0.f ? r.y : 0.f;
          r.z = gv[u].z > 0.f ? r.z : 0.f;
          r.w = gv[u].w > 0.f ? r.w : 0.f;
          stg_stream4(dgb + (size_t)t * dg_fs + 4 * (size_t)e, r);
          prev[u] = cur[u];
        }
      }
    }
    return;
  }
  // ------------------------------ spatial: one CTA per (frame, 4-channel group)
  __shared__ float gt[ST_CGB][ST_MAXW + 2][ST_MAXW + 3];   // dropped upstream gradient, zero halo
  __shared__ float dt[ST_CGB][ST_MAXW + 2][ST_MAXW + 3];   // D planes, zero halo (only for the tap gradient)
  const int sblk = blockIdx.x - n_tblocks;
  const int groups = (s.Cs + ST_CGB - 1) / ST_CGB;
  const int f = sblk / groups, c0 = (sblk - f * groups) * ST_CGB;
  const int nc = min(ST_CGB, s.Cs - c0);
  const int p = spatial_pair_of_frame(s, f);
  float* ddp = dd + (size_t)f * dd_fs + (size_t)c0 * HW;
  if (p < 0) {  // frame feeds no pair: its spatial-branch gradient is zero
    for (int i = tid; i < nc * HW; i += ST_THREADS) ddp[i] = 0.f;
    return;
  }
  const int PH = s.H + 2, PW = s.W + 2;
  const uint32_t thr = drop_threshold24(s.drop_p);
  const int KC = s.K * s.Cs;
  const bool need_dw = (dw != nullptr);
  if (need_dw) {
    for (int i = tid; i < nc * PH * PW; i += ST_THREADS) {
      const int c = i / (PH * PW), r = i - c * (PH * PW);
      const int yy = r / PW, xx = r - yy * PW;
      float v = 0.f;
      if (yy >= 1 && yy <= s.H && xx >= 1 && xx <= s.W)
        v = __ldg(d + (size_t)f * s.d_fs + (size_t)(c0 + c) * HW + (yy - 1) * s.W + (xx - 1));
      dt[c][yy][xx] = v;
    }
  }
  for (int kk = 0; kk < s.K; ++kk) {
    __syncthreads();  // previous kk's readers of gt are done (and dt is complete on the first pass)
    for (int i = tid; i < nc * PH * PW; i += ST_THREADS) {
      const int c = i / (PH * PW), r = i - c * (PH * PW);
      const int yy = r / PW, xx = r - yy * PW;
      float v = 0.f;
      if (yy >= 1 && yy <= s.H && xx >= 1 && xx <= s.W) {
        const int och = kk * s.Cs + c0 + c;
        const int pix = (yy - 1) * s.W + (xx - 1);
        v = __ldg(dout + ((size_t)p * s.out_ctot + s.out_coff + och) * HW + pix);
        v *= drop_factor(s, thr, ((size_t)p * KC + och) * HW + pix);
      }
      gt[c][yy][xx] = v;
    }
    __syncthreads();
    // dD(y,x) (+)= sum_ab w[a][b] * dS(y-a+1, x-b+1)
    for (int i = tid; i < nc * HW; i += ST_THREADS) {
      const int c = i / HW, pix = i - c * HW;
      const int y = pix / s.W, x = pix - y * s.W;
      const float* wk = w + ((size_t)(c0 + c) * s.K + kk) * 9;
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) acc = fmaf(__ldg(wk + a * 3 + bb), gt[c][y + 2 - a][x + 2 - bb], acc);
      if (kk == 0) ddp[i] = acc;
      else ddp[i] += acc;
    }
    if (need_dw || dbias) {
      // tap / bias gradients: per channel 9 (+1) sums over the plane; warp w handles channel w (nc <= 8 warps)
      const int warp = tid >> 5, lane = tid & 31;
      if (warp < nc) {
        const int c = warp;
        float sums[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) sums[j] = 0.f;
        for (int pix = lane; pix < HW; pix += 32) {
          const int y = pix / s.W, x = pix - y * s.W;
          const float gvv = gt[c][y + 1][x + 1];
          sums[9] += gvv;
          if (need_dw) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
              for (int bb = 0; bb < 3; ++bb) sums[a * 3 + bb] = fmaf(gvv, dt[c][y + a][x + bb], sums[a * 3 + bb]);
          }
        }
#pragma unroll
        for (int j = 0; j < 10; ++j) sums[j] = warp_sum(sums[j]);
        if (lane == 0) {
          if (need_dw)
            for (int j = 0; j < 9; ++j) atomicAdd(dw + ((size_t)(c0 + c) * s.K + kk) * 9 + j, sums[j]);
          if (dbias) atomicAdd(dbias + kk * s.Cs + c0 + c, sums[9]);
        }
      }
    }
  }
}

static int check_stencil(const offk_stencil_t* s) {
  OFFK_REQUIRE(s != nullptr, "stencil: null descriptor");
  OFFK_REQUIRE(s->B >= 1 && s->L >= 2, "stencil: need B >= 1 and L >= 2 (got B=%d L=%d)", s->B, s->L);
  OFFK_REQUIRE(s->Cg >= 0 && s->Cs >= 0 && s->Cg + s->Cs > 0, "stencil: bad channel counts");
  OFFK_REQUIRE(s->H >= 1 && s->W >= 1 && s->H <= ST_MAXW && s->W <= ST_MAXW, "stencil: plane %dx%d beyond %d", s->H,
               s->W, ST_MAXW);
  OFFK_REQUIRE(s->K >= 1 && s->K <= 2, "stencil: K must be 1 or 2");
  OFFK_REQUIRE(s->out_coff >= 0 && s->out_coff + s->K * s->Cs + s->Cg <= s->out_ctot, "stencil: channel slice");
  OFFK_REQUIRE(s->drop_mode >= 0 && s->drop_mode <= 2, "stencil: drop_mode");
  OFFK_REQUIRE(s->drop_mode != OFFK_DROP_MASK || s->keep_mask, "stencil: OFFK_DROP_MASK needs keep_mask");
  const long long HW = (long long)s->H * s->W;
  if (s->Cg > 0) {
    OFFK_REQUIRE((s->Cg * HW) % 4 == 0 && s->g_fs % 4 == 0 && (s->out_ctot * HW) % 4 == 0 &&
                     ((s->out_coff + s->K * s->Cs) * HW) % 4 == 0,
                 "stencil: temporal spans must be float4-aligned");
  }
  return 0;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace offk

using namespace offk;

extern "C" int offk_stencil_diff_fwd(const offk_stencil_t* s, const float* g, const float* d, const float* w,
                                     const float* bias, float* out, void* stream) {
  if (int e = check_stencil(s)) return e;
  OFFK_REQUIRE(out != nullptr && aligned16(out), "stencil_fwd: out must be non-null and 16-byte aligned");
  OFFK_REQUIRE(s->Cg == 0 || (g && aligned16(g)), "stencil_fwd: g must be 16-byte aligned");
  OFFK_REQUIRE(s->Cs == 0 || (d && w), "stencil_fwd: d / w missing");
  const int HW = s->H * s->W;
  const int span4 = (s->Cg * HW) / 4;
  const int per_blk = ST_THREADS * ST_TPOS;
  const int tb = s->Cg > 0 ? (span4 + per_blk - 1) / per_blk : 0;
  const int n_t = tb * s->B;
  const int P = s->B * (s->L - 1);
  const int n_s = s->Cs > 0 ? P * ((s->Cs + ST_CG - 1) / ST_CG) : 0;
  if (n_t + n_s == 0) return 0;
  stencil_diff_fwd_kernel<<<n_t + n_s, ST_THREADS, 0, as_stream(stream)>>>(*s, g, d, w, bias, out, tb > 0 ? tb : 1,
                                                                           n_t);
  return OFFK_LAUNCH_CHECK("stencil_diff_fwd");
}

extern "C" int offk_stencil_diff_bwd(const offk_stencil_t* s, const float* dout, const float* g, const float* d,
                                     const float* w, float* dg, int64_t dg_fs, float* dd, int64_t dd_fs, float* dw,
                                     float* dbias, void* stream) {
  if (int e = check_stencil(s)) return e;
  OFFK_REQUIRE(dout != nullptr && aligned16(dout), "stencil_bwd: dout must be 16-byte aligned");
  OFFK_REQUIRE(s->Cg == 0 || (g && dg && aligned16(g) && aligned16(dg) && dg_fs % 4 == 0), "stencil_bwd: g/dg");
  OFFK_REQUIRE(s->Cs == 0 || (w && dd), "stencil_bwd: w/dd missing");
  OFFK_REQUIRE(dw == nullptr || d != nullptr, "stencil_bwd: tap gradient needs d");
  const int HW = s->H * s->W;
  const int span4 = (s->Cg * HW) / 4;
  const int per_blk = ST_THREADS * ST_TPOS;
  const int tb = s->Cg > 0 ? (span4 + per_blk - 1) / per_blk : 0;
  const int n_t = tb * s->B;
  const int n_s = s->Cs > 0 ? s->B * s->L * ((s->Cs + ST_CGB - 1) / ST_CGB) : 0;
  if (n_t + n_s == 0) return 0;
  stencil_diff_bwd_kernel<<<n_t + n_s, ST_THREADS, 0, as_stream(stream)>>>(*s, dout, g, d, w, dg, (long long)dg_fs, dd,
                                                                           (long long)dd_fs, dw, dbias,
                                                                           tb > 0 ? tb : 1, n_t);
  return OFFK_LAUNCH_CHECK("stencil_diff_bwd");
}
