// Fused OFF stencil kernels (forward and backward) on channels-last tensors; HBM-bound by construction.
//
// Forward (one launch per OFF unit) reads the unit's reduced features once and writes the unit's 160 channels
// straight into the concatenated stage-fusion buffer [P, H, W, Ctot] at its channel offset:
//   temporal blocks : a thread owns up to 4 (pixel, 4-channel) positions of a clip and walks t = 0..L-1 keeping the
//                     previous frame in registers, so every G frame is read exactly once (128-bit, coalesced: one
//                     warp = the 128 channels of one pixel) and every difference G(t+1)-G(t) is written once.
//   spatial blocks  : a thread owns (pair, pixel, 4 channels): nine predicated 128-bit neighbour loads (zero padding by
//                     predication, neighbours come from L1/L2), per-channel 3x3 taps cached in shared memory, bias,
//                     dropout, one 128-bit store.  K = 2 emits two maps per channel (Sobel x and y).
// No torch.cat, no sub, no conv2d, no separate dropout launch.
//
// Backward mirrors it: dG = (dT(t-1) - dT(t)) * [G > 0], dD = transposed stencil of the dropped spatial gradient, and
// the learned-tap / bias gradients are accumulated in registers by a few persistent blocks, reduced with warp
// shuffles + shared-memory atomics, then one global atomicAdd per tap and block.
#include "offk_common.cuh"

namespace offk {

constexpr int ST_THREADS = 256;
constexpr int ST_TPOS = 4;      // positions per thread (temporal half)
constexpr int ST_MAX_CS = 64;   // spatial channels cached in smem (32 on the path)

__device__ __forceinline__ float4 f4sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

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
// keep-factor of spatial output element (p, och, y, x); the mask / hash index is the NCHW flat index of the
// reference's dropout input [P, K*Cs, H, W] (RGB_OFF.py:612), so injected masks keep the reference's layout.
__device__ __forceinline__ float drop_factor(const offk_stencil_t& s, uint32_t thr, int p, int och, int pix) {
  if (s.drop_mode == OFFK_DROP_NONE) return 1.f;
  const size_t idx = ((size_t)p * (s.K * s.Cs) + och) * (size_t)(s.H * s.W) + pix;
  const bool keep = s.drop_mode == OFFK_DROP_MASK ? (s.keep_mask[idx] != 0) : drop_keep(s.seed, idx, thr);
  return keep ? s.keep_scale : 0.f;
}
__device__ __forceinline__ float4 drop_factor4(const offk_stencil_t& s, uint32_t thr, int p, int och, int pix) {
  return make_float4(drop_factor(s, thr, p, och, pix), drop_factor(s, thr, p, och + 1, pix),
                     drop_factor(s, thr, p, och + 2, pix), drop_factor(s, thr, p, och + 3, pix));
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
    const int cq = s.Cg >> 2;                       // channel quads per pixel
    const int npos = HW * cq;
    const int base = blk * (ST_THREADS * ST_TPOS) + tid;
    const float* gb = g + (size_t)b * s.L * s.g_fs;
    float* ob = out + (size_t)b * (s.L - 1) * HW * s.out_ctot + s.out_coff + s.K * s.Cs;
    const size_t o_ps = (size_t)HW * s.out_ctot;
    size_t goff[ST_TPOS], ooff[ST_TPOS];
    float4 prev[ST_TPOS];
#pragma unroll
    for (int u = 0; u < ST_TPOS; ++u) {
      const int e = base + u * ST_THREADS;
      const int pix = e / cq, c = (e - pix * cq) << 2;
      goff[u] = (size_t)pix * s.g_ps + c;
      ooff[u] = (size_t)pix * s.out_ctot + c;
      if (e < npos) prev[u] = ldg_stream4(gb + goff[u]);
    }
    for (int t = 1; t < s.L; ++t) {
      float4 cur[ST_TPOS];
#pragma unroll
      for (int u = 0; u < ST_TPOS; ++u)
        if (base + u * ST_THREADS < npos) cur[u] = ldg_stream4(gb + (size_t)t * s.g_fs + goff[u]);
#pragma unroll
      for (int u = 0; u < ST_TPOS; ++u)
        if (base + u * ST_THREADS < npos) {
          stg_stream4(ob + (size_t)(t - 1) * o_ps + ooff[u], f4sub(cur[u], prev[u]));
          prev[u] = cur[u];
        }
    }
    return;
  }
  // ------------------------------ spatial gradient (RGB_OFF.py:611 / Flow_OFF.py:622 / util.py:46-50) + dropout
  __shared__ float ws[ST_MAX_CS * 2 * 9];
  __shared__ float bs[ST_MAX_CS * 2];
  for (int i = tid; i < s.Cs * s.K * 9; i += ST_THREADS) ws[i] = __ldg(w + i);
  for (int i = tid; i < s.Cs * s.K; i += ST_THREADS) bs[i] = bias ? __ldg(bias + i) : 0.f;
  __syncthreads();
  const int cq = s.Cs >> 2;
  const int P = s.B * (s.L - 1);
  const size_t total = (size_t)P * HW * cq;
  const size_t e = (size_t)(blockIdx.x - n_tblocks) * ST_THREADS + tid;
  if (e >= total) return;
  const int c = (int)(e % cq) << 2;
  const int pix = (int)((e / cq) % HW);
  const int p = (int)(e / ((size_t)cq * HW));
  const int y = pix / s.W, x = pix - y * s.W;
  const float* dp = d + (size_t)spatial_frame(s, p) * s.d_fs + c;
  float4 nb[9];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int bb = 0; bb < 3; ++bb) {
      const int yy = y + a - 1, xx = x + bb - 1;
      nb[a * 3 + bb] = (yy >= 0 && yy < s.H && xx >= 0 && xx < s.W) ? ldg4(dp + (size_t)(yy * s.W + xx) * s.d_ps)
                                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  const uint32_t thr = drop_threshold24(s.drop_p);
  float* op = out + ((size_t)p * HW + pix) * s.out_ctot + s.out_coff;
  for (int kk = 0; kk < s.K; ++kk) {
    const int och = kk * s.Cs + c;
    float4 acc = make_float4(bs[och], bs[och + 1], bs[och + 2], bs[och + 3]);
    const float* w0 = ws + ((c + 0) * s.K + kk) * 9;
    const float* w1 = ws + ((c + 1) * s.K + kk) * 9;
    const float* w2 = ws + ((c + 2) * s.K + kk) * 9;
    const float* w3 = ws + ((c + 3) * s.K + kk) * 9;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      acc.x = fmaf(w0[j], nb[j].x, acc.x);
      acc.y = fmaf(w1[j], nb[j].y, acc.y);
      acc.z = fmaf(w2[j], nb[j].z, acc.z);
      acc.w = fmaf(w3[j], nb[j].w, acc.w);
    }
    const float4 k4 = drop_factor4(s, thr, p, och, pix);
    stg_stream4(op + och, make_float4(acc.x * k4.x, acc.y * k4.y, acc.z * k4.z, acc.w * k4.w));
  }
}

// ---------------------------------------------------------------------------------------------- backward
__global__ void __launch_bounds__(ST_THREADS)
stencil_diff_bwd_kernel(const offk_stencil_t s, const float* __restrict__ dout, const float* __restrict__ g,
                        const float* __restrict__ d, const float* __restrict__ w, float* __restrict__ dg,
                        long long dg_fs, float* __restrict__ dd, long long dd_fs,
                        int t_blocks_per_clip, int n_tblocks) {
  const int HW = s.H * s.W;
  const int tid = threadIdx.x;
  if ((int)blockIdx.x < n_tblocks) {
    // ------------------------------ dG[b,t] = (dT[b,t-1] - dT[b,t]) * (G[b,t] > 0)
    const int b = blockIdx.x / t_blocks_per_clip;
    const int blk = blockIdx.x - b * t_blocks_per_clip;
    const int cq = s.Cg >> 2;
    const int npos = HW * cq;
    const int base = blk * (ST_THREADS * ST_TPOS) + tid;
    const float* gb = g + (size_t)b * s.L * s.g_fs;
    float* dgb = dg + (size_t)b * s.L * dg_fs;
    const float* ob = dout + (size_t)b * (s.L - 1) * HW * s.out_ctot + s.out_coff + s.K * s.Cs;
    const size_t o_ps = (size_t)HW * s.out_ctot;
    size_t goff[ST_TPOS], ooff[ST_TPOS];
    float4 prev[ST_TPOS];
#pragma unroll
    for (int u = 0; u < ST_TPOS; ++u) {
      const int e = base + u * ST_THREADS;
      const int pix = e / cq, c = (e - pix * cq) << 2;
      goff[u] = (size_t)pix * s.g_ps + c;     // dg uses the same pixel stride as g
      ooff[u] = (size_t)pix * s.out_ctot + c;
      prev[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int t = 0; t < s.L; ++t) {
      float4 cur[ST_TPOS], gv[ST_TPOS];
#pragma unroll
      for (int u = 0; u < ST_TPOS; ++u) {
        cur[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (base + u * ST_THREADS < npos) {
          if (t < s.L - 1) cur[u] = ldg_stream4(ob + (size_t)t * o_ps + ooff[u]);
          gv[u] = ldg_stream4(gb + (size_t)t * s.g_fs + goff[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < ST_TPOS; ++u)
        if (base + u * ST_THREADS < npos) {
          float4 r = f4sub(prev[u], cur[u]);
          r.x = gv[u].x > 0.f ? r.x : 0.f;
          r.y = gv[u].y > 0.f ? r.y : 0.f;
          r.z = gv[u].z > 0.f ? r.z : 0.f;
          r.w = gv[u].w > 0.f ? r.w : 0.f;
          stg_stream4(dgb + (size_t)t * dg_fs + goff[u], r);
          prev[u] = cur[u];
        }
    }
    return;
  }
  __shared__ float ws[ST_MAX_CS * 2 * 9];
  const uint32_t thr = drop_threshold24(s.drop_p);
  const int cq = s.Cs >> 2;
  {
    // ------------------------------ dD(y,x) = sum_kk sum_ab w[kk][a][b] * dS[kk](y-a+1, x-b+1)
    for (int i = tid; i < s.Cs * s.K * 9; i += ST_THREADS) ws[i] = __ldg(w + i);
    __syncthreads();
    const size_t total = (size_t)s.B * s.L * HW * cq;
    const size_t e = (size_t)(blockIdx.x - n_tblocks) * ST_THREADS + tid;
    if (e >= total) return;
    const int c = (int)(e % cq) << 2;
    const int pix = (int)((e / cq) % HW);
    const int f = (int)(e / ((size_t)cq * HW));
    const int y = pix / s.W, x = pix - y * s.W;
    float* ddp = dd + (size_t)f * dd_fs + (size_t)pix * s.d_ps + c;   // dd uses the same pixel stride as d
    const int p = spatial_pair_of_frame(s, f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p >= 0) {
      for (int kk = 0; kk < s.K; ++kk) {
        const int och = kk * s.Cs + c;
        const float* w0 = ws + ((c + 0) * s.K + kk) * 9;
        const float* w1 = ws + ((c + 1) * s.K + kk) * 9;
        const float* w2 = ws + ((c + 2) * s.K + kk) * 9;
        const float* w3 = ws + ((c + 3) * s.K + kk) * 9;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int bb = 0; bb < 3; ++bb) {
            const int yy = y - a + 1, xx = x - bb + 1;
            if (yy >= 0 && yy < s.H && xx >= 0 && xx < s.W) {
              const int np_ = yy * s.W + xx;
              float4 v = ldg4(dout + ((size_t)p * HW + np_) * s.out_ctot + s.out_coff + och);
              const float4 k4 = drop_factor4(s, thr, p, och, np_);
              acc.x = fmaf(w0[a * 3 + bb], v.x * k4.x, acc.x);
              acc.y = fmaf(w1[a * 3 + bb], v.y * k4.y, acc.y);
              acc.z = fmaf(w2[a * 3 + bb], v.z * k4.z, acc.z);
              acc.w = fmaf(w3[a * 3 + bb], v.w * k4.w, acc.w);
            }
          }
      }
    }
    *reinterpret_cast<float4*>(ddp) = acc;   // zero for frames that feed no pair
    return;
  }
}

// tap / bias gradients: persistent blocks, register accumulation (separate kernel: it needs ~120 registers, the
// streaming halves above should keep their occupancy)
//   dw[c,kk,a,b] = sum_{p,y,x} dS[p,kk*Cs+c,y,x] * D[fs(p),c,y+a-1,x+b-1],  dbias[kk*Cs+c] = sum dS
__global__ void __launch_bounds__(ST_THREADS)
stencil_tapgrad_kernel(const offk_stencil_t s, const float* __restrict__ dout, const float* __restrict__ d,
                       float* __restrict__ dw, float* __restrict__ dbias) {
  __shared__ float acc_s[ST_MAX_CS * 2 * 10];
  const int HW = s.H * s.W;
  const int tid = threadIdx.x;
  const uint32_t thr = drop_threshold24(s.drop_p);
  const int cq = s.Cs >> 2;
  const int n_wblocks = gridDim.x;
  const int wb = blockIdx.x;
  for (int i = tid; i < s.Cs * s.K * 10; i += ST_THREADS) acc_s[i] = 0.f;
  __syncthreads();
  const int P = s.B * (s.L - 1);
  const int groups = ST_THREADS / cq;                 // (pair,pixel) positions handled per block iteration
  const int c = (tid % cq) << 2;
  const int slot = tid / cq;
  const bool need_dw = dw != nullptr;
  for (int kk = 0; kk < s.K; ++kk) {
    float sums[4][10];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 10; ++j) sums[i][j] = 0.f;
    const int och = kk * s.Cs + c;
    if (slot < groups) {
      for (size_t pos = (size_t)wb * groups + slot; pos < (size_t)P * HW; pos += (size_t)n_wblocks * groups) {
        const int p = (int)(pos / HW), pix = (int)(pos - (size_t)p * HW);
        const int y = pix / s.W, x = pix - y * s.W;
        float4 v = ldg4(dout + ((size_t)p * HW + pix) * s.out_ctot + s.out_coff + och);
        const float4 k4 = drop_factor4(s, thr, p, och, pix);
        v.x *= k4.x; v.y *= k4.y; v.z *= k4.z; v.w *= k4.w;
        sums[0][9] += v.x; sums[1][9] += v.y; sums[2][9] += v.z; sums[3][9] += v.w;
        if (need_dw) {
          const float* dp = d + (size_t)spatial_frame(s, p) * s.d_fs + c;
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) {
              const int yy = y + a - 1, xx = x + bb - 1;
              if (yy >= 0 && yy < s.H && xx >= 0 && xx < s.W) {
                const float4 nv = ldg4(dp + (size_t)(yy * s.W + xx) * s.d_ps);
                sums[0][a * 3 + bb] = fmaf(v.x, nv.x, sums[0][a * 3 + bb]);
                sums[1][a * 3 + bb] = fmaf(v.y, nv.y, sums[1][a * 3 + bb]);
                sums[2][a * 3 + bb] = fmaf(v.z, nv.z, sums[2][a * 3 + bb]);
                sums[3][a * 3 + bb] = fmaf(v.w, nv.w, sums[3][a * 3 + bb]);
              }
            }
        }
      }
    }
    // lanes l and l^cq, l^2cq, ... hold the same channels (cq is a power of two <= 32 on this path)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        float v = sums[i][j];
        for (int o = 16; o >= cq; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) < cq && slot < groups) atomicAdd(&acc_s[((c + i) * s.K + kk) * 10 + j], v);
      }
  }
  __syncthreads();
  for (int i = tid; i < s.Cs * s.K * 10; i += ST_THREADS) {
    const int ck = i / 10, j = i - ck * 10;   // ck = c*K + kk
    if (j < 9) {
      if (need_dw) atomicAdd(dw + (size_t)ck * 9 + j, acc_s[i]);
    } else if (dbias) {
      const int cc = ck / s.K, kk = ck - cc * s.K;
      atomicAdd(dbias + kk * s.Cs + cc, acc_s[i]);
    }
  }
}

static int check_stencil(const offk_stencil_t* s) {
  OFFK_REQUIRE(s != nullptr, "stencil: null descriptor");
  OFFK_REQUIRE(s->B >= 1 && s->L >= 2, "stencil: need B >= 1 and L >= 2 (got B=%d L=%d)", s->B, s->L);
  OFFK_REQUIRE(s->Cg >= 0 && s->Cs >= 0 && s->Cg + s->Cs > 0, "stencil: bad channel counts");
  OFFK_REQUIRE(s->Cg % 4 == 0 && s->Cs % 4 == 0 && s->Cs <= ST_MAX_CS, "stencil: channels must be multiples of 4, Cs <= %d",
               ST_MAX_CS);
  OFFK_REQUIRE(s->Cs == 0 || ((s->Cs >> 2) <= 32 && (((s->Cs >> 2) & ((s->Cs >> 2) - 1)) == 0)),
               "stencil: Cs/4 must be a power of two <= 32");
  OFFK_REQUIRE(s->H >= 1 && s->W >= 1 && s->H < 32768 && s->W < 32768, "stencil: plane size");
  OFFK_REQUIRE(s->K >= 1 && s->K <= 2, "stencil: K must be 1 or 2");
  OFFK_REQUIRE(s->out_coff >= 0 && s->out_coff + s->K * s->Cs + s->Cg <= s->out_ctot, "stencil: channel slice");
  OFFK_REQUIRE(s->out_ctot % 4 == 0 && s->out_coff % 4 == 0, "stencil: output slice must be float4-aligned");
  OFFK_REQUIRE(s->g_fs % 4 == 0 && s->d_fs % 4 == 0 && s->g_ps % 4 == 0 && s->d_ps % 4 == 0, "stencil: strides");
  OFFK_REQUIRE(s->drop_mode >= 0 && s->drop_mode <= 2, "stencil: drop_mode");
  OFFK_REQUIRE(s->drop_mode != OFFK_DROP_MASK || s->keep_mask, "stencil: OFFK_DROP_MASK needs keep_mask");
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
  OFFK_REQUIRE(s->Cs == 0 || (d && w && aligned16(d)), "stencil_fwd: d / w missing or unaligned");
  const long long HW = (long long)s->H * s->W;
  const long long npos = HW * (s->Cg / 4);
  const int per_blk = ST_THREADS * ST_TPOS;
  const int tb = s->Cg > 0 ? (int)((npos + per_blk - 1) / per_blk) : 0;
  const long long n_t = (long long)tb * s->B;
  const long long P = (long long)s->B * (s->L - 1);
  const long long n_s = s->Cs > 0 ? (P * HW * (s->Cs / 4) + ST_THREADS - 1) / ST_THREADS : 0;
  if (n_t + n_s == 0) return 0;
  OFFK_REQUIRE(n_t + n_s < 2147483647LL, "stencil_fwd: grid too large");
  stencil_diff_fwd_kernel<<<(unsigned)(n_t + n_s), ST_THREADS, 0, as_stream(stream)>>>(*s, g, d, w, bias, out,
                                                                                      tb > 0 ? tb : 1, (int)n_t);
  return OFFK_LAUNCH_CHECK("stencil_diff_fwd");
}

extern "C" int offk_stencil_diff_bwd(const offk_stencil_t* s, const float* dout, const float* g, const float* d,
                                     const float* w, float* dg, int64_t dg_fs, float* dd, int64_t dd_fs, float* dw,
                                     float* dbias, void* stream) {
  if (int e = check_stencil(s)) return e;
  OFFK_REQUIRE(dout != nullptr && aligned16(dout), "stencil_bwd: dout must be 16-byte aligned");
  OFFK_REQUIRE(s->Cg == 0 || (g && dg && aligned16(g) && aligned16(dg) && dg_fs % 4 == 0), "stencil_bwd: g/dg");
  OFFK_REQUIRE(s->Cs == 0 || (w && dd && aligned16(dd) && dd_fs % 4 == 0), "stencil_bwd: w/dd missing");
  OFFK_REQUIRE(dw == nullptr || d != nullptr, "stencil_bwd: tap gradient needs d");
  const long long HW = (long long)s->H * s->W;
  const long long npos = HW * (s->Cg / 4);
  const int per_blk = ST_THREADS * ST_TPOS;
  const int tb = s->Cg > 0 ? (int)((npos + per_blk - 1) / per_blk) : 0;
  const long long n_t = (long long)tb * s->B;
  const long long n_s = s->Cs > 0 ? ((long long)s->B * s->L * HW * (s->Cs / 4) + ST_THREADS - 1) / ST_THREADS : 0;
  long long n_w = 0;
  if (s->Cs > 0 && (dw || dbias)) {
    const long long groups = ST_THREADS / (s->Cs / 4);
    const long long need = ((long long)s->B * (s->L - 1) * HW + groups - 1) / groups;
    n_w = need < 2 * sm_count() ? need : 2 * sm_count();
  }
  if (n_t + n_s == 0) return 0;
  OFFK_REQUIRE(n_t + n_s < 2147483647LL, "stencil_bwd: grid too large");
  stencil_diff_bwd_kernel<<<(unsigned)(n_t + n_s), ST_THREADS, 0, as_stream(stream)>>>(
      *s, dout, g, d, w, dg, (long long)dg_fs, dd, (long long)dd_fs, tb > 0 ? tb : 1, (int)n_t);
  if (int e = OFFK_LAUNCH_CHECK("stencil_diff_bwd")) return e;
  if (n_w > 0) {
    stencil_tapgrad_kernel<<<(unsigned)n_w, ST_THREADS, 0, as_stream(stream)>>>(*s, dout, d, dw, dbias);
    return OFFK_LAUNCH_CHECK("stencil_tapgrad");
  }
  return 0;
}
