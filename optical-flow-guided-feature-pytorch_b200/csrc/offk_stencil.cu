// Fused OFF stencil kernels (forward and backward) on channels-last tensors; HBM-bound by construction.
//
// One launch serves a BATCH of OFF units (the units feeding one stage-fusion buffer, or all nine in the backward
// pass), so the 14x14 and 7x7 levels -- a few MB each -- do not pay one launch latency apiece.  A block is either
//   a temporal block : (clip, 32 pixels).  One warp = the 128 reduced channels of a pixel (lane = channel quad), so
//                      there is no index arithmetic beyond pointer increments.  The warp walks t = 0..L-1 with the
//                      loads of frame t+1 issued before the difference of frame t is stored: every G frame is read
//                      exactly once (128-bit, 512 contiguous bytes per warp), every difference written once, and
//                      8 independent 16-byte loads per thread are in flight.
//   a spatial block  : (pair, band of rows).  The D band plus a one-pixel halo is staged in shared memory with
//                      16-byte cp.async (zero-fill gives the conv's zero padding), then each thread produces
//                      (pixel, 4 channels) outputs from nine LDS.128 neighbours and per-channel taps kept in
//                      shared memory; bias, dropout and the store at the unit's channel offset of the stage buffer
//                      follow.  K = 2 emits two maps per channel (Sobel x and y).
// No torch.cat, no sub, no conv2d, no separate dropout launch.
//
// Backward mirrors it: dG = (dT(t-1) - dT(t)) * [G > 0]; the spatial blocks stage the DROPPED spatial gradient dS
// (band + halo) once and use the same nine neighbours twice: dD = transposed stencil of dS, and
// dw[a,b] += D(y,x) * dS(y-a+1, x-b+1) (the tap gradient re-indexed so that it needs D only at the centre pixel).
// Tap / bias gradients are accumulated in registers by persistent spatial blocks, reduced with warp shuffles and
// shared-memory atomics, then one global atomicAdd per tap and block.
#include "offk_common.cuh"

namespace offk {

constexpr int ST_THREADS = 256;
constexpr int ST_WARPS = ST_THREADS / 32;
constexpr int ST_MAX_LEVELS = 12;
constexpr int ST_MAX_CS = 64;                 // spatial channels per level (32 on the path)
constexpr int ST_TJ = 2;                      // pixels per warp of a temporal block (lane = channel quad)
constexpr int ST_FC = 3;                      // frames loaded per chunk, forward
constexpr int ST_BC = 2;                      // frames loaded per chunk, backward (two tensors per frame)
constexpr int ST_TILE_BUDGET = 24 * 1024;     // target size of a halo tile (bytes)
constexpr int ST_SMEM_MAX = 64 * 1024;

struct StLevel {
  offk_stencil_t s;
  const float* g;
  const float* d;
  const float* w;
  const float* bias;
  float* out;          // forward: stage buffer
  const float* dout;   // backward: gradient of the stage buffer
  float* dg;
  float* dd;
  float* dw;
  float* dbias;
  long long dg_fs, dd_fs;
  int blk0;            // first block of this level
  int n_sblocks;       // spatial blocks (forward: one per item; backward: persistent over items)
  int n_sitems;        // spatial items = (pair | frame) x band
  int bands, band_rows;
  int t_chunks;        // temporal blocks per clip
  int lq, ltcp, lxw;   // log2: channel quads per pixel, tile column pitch, x positions per pass
};
struct StBatch {
  int n;
  StLevel lv[ST_MAX_LEVELS];
};

__device__ __forceinline__ float4 f4sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_cp_async16(uint32_t dst, const float* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void st_cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

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
// keep-factors of the 4 spatial outputs (p, och..och+3, pix).  OFFK_DROP_MASK: the mask is indexed like the
// reference's dropout input [P, K*Cs, H, W] (RGB_OFF.py:612), so injected masks keep the reference's layout.
// OFFK_DROP_SEED: one 64-bit hash per channel quad of the channels-last element index ((p*HW + pix)*K*Cs + och).
__device__ __forceinline__ float4 drop_factor4(const offk_stencil_t& s, uint32_t thr, int p, int och, int pix, int HW) {
  if (s.drop_mode == OFFK_DROP_NONE) return make_float4(1.f, 1.f, 1.f, 1.f);
  const int KC = s.K * s.Cs;
  uint32_t keep;
  if (s.drop_mode == OFFK_DROP_MASK) {
    const uint8_t* m = s.keep_mask + ((size_t)p * KC + och) * (size_t)HW + pix;
    keep = (m[0] != 0 ? 1u : 0u) | (m[HW] != 0 ? 2u : 0u) | (m[2 * (size_t)HW] != 0 ? 4u : 0u) | (m[3 * (size_t)HW] != 0 ? 8u : 0u);
  } else {
    keep = drop_keep4(s.seed, (((uint64_t)p * HW + pix) * KC + och) >> 2, thr);
  }
  const float k = s.keep_scale;
  return make_float4(keep & 1u ? k : 0.f, keep & 2u ? k : 0.f, keep & 4u ? k : 0.f, keep & 8u ? k : 0.f);
}

__device__ __forceinline__ const StLevel& find_level(const StBatch& bt, int blk) {
  int li = 0;
#pragma unroll 1
  while (li + 1 < bt.n && blk >= bt.lv[li + 1].blk0) ++li;
  return bt.lv[li];
}

// per-channel taps and bias as float4 over 4 consecutive channels: ws4[(kk*9 + j)*CQ + c4], bs4[kk*CQ + c4]
__device__ __forceinline__ void load_taps(const StLevel& lv, float4* ws4, float4* bs4, int tid) {
  const offk_stencil_t& s = lv.s;
  const int CQ = 1 << lv.lq, K9 = s.K * 9;
  for (int i = tid; i < K9 * CQ; i += ST_THREADS) {
    const int c4 = i & (CQ - 1), kj = i >> lv.lq;
    const float* b = lv.w + (size_t)(c4 * 4) * K9 + kj;          // w[c][kk][3][3]: channel stride K*9
    ws4[i] = make_float4(__ldg(b), __ldg(b + K9), __ldg(b + 2 * K9), __ldg(b + 3 * K9));
  }
  if (bs4) {
    for (int i = tid; i < s.K * CQ; i += ST_THREADS) {
      const int c4 = i & (CQ - 1), kk = i >> lv.lq;
      bs4[i] = lv.bias ? ldg4(lv.bias + kk * s.Cs + c4 * 4) : f4zero();
    }
  }
}

// ---------------------------------------------------------------------------------------------- forward
// temporal difference (RGB_OFF.py:599-604): thread = (ST_TJ pixels, one channel quad); frames are walked in chunks of
// ST_FC so that ST_TJ*ST_FC independent 16-byte loads are in flight per thread at ~50 % occupancy.
__device__ __forceinline__ void temporal_fwd(const StLevel& lv, int r2, int tid) {
  const int L = lv.s.L, Cg = lv.s.Cg, HW = lv.s.H * lv.s.W;
  const int g_ps = lv.s.g_ps, ctot = lv.s.out_ctot;
  const size_t g_fs = (size_t)lv.s.g_fs, o_fs = (size_t)HW * ctot;
  const int b = r2 / lv.t_chunks, chunk = r2 - b * lv.t_chunks;
  const int warp = tid >> 5, lane = tid & 31;
  const int pix0 = chunk * (ST_WARPS * ST_TJ) + warp * ST_TJ;
  bool ok[ST_TJ];
#pragma unroll
  for (int j = 0; j < ST_TJ; ++j) ok[j] = pix0 + j < HW;
  for (int c = lane * 4; c < Cg; c += 128) {
    const float* gp = lv.g + (size_t)b * L * g_fs + (size_t)pix0 * g_ps + c;
    float* op = lv.out + ((size_t)b * (L - 1) * HW + pix0) * ctot + lv.s.out_coff + lv.s.K * lv.s.Cs + c;
    float4 prev[ST_TJ];
#pragma unroll
    for (int j = 0; j < ST_TJ; ++j) prev[j] = ok[j] ? ldg_stream4(gp + j * g_ps) : f4zero();
    for (int t0 = 1; t0 < L; t0 += ST_FC) {
      float4 cur[ST_FC][ST_TJ];
#pragma unroll
      for (int i = 0; i < ST_FC; ++i) {
        gp += g_fs;
#pragma unroll
        for (int j = 0; j < ST_TJ; ++j) cur[i][j] = (ok[j] && t0 + i < L) ? ldg_stream4(gp + j * g_ps) : f4zero();
      }
#pragma unroll
      for (int i = 0; i < ST_FC; ++i) {
        if (t0 + i < L) {
#pragma unroll
          for (int j = 0; j < ST_TJ; ++j) {
            if (ok[j]) stg_stream4(op + j * ctot, f4sub(cur[i][j], prev[j]));
            prev[j] = cur[i][j];
          }
        }
        op += o_fs;
      }
    }
  }
}

__global__ void __launch_bounds__(ST_THREADS, 4)
stencil_diff_fwd_kernel(const __grid_constant__ StBatch bt) {
  extern __shared__ __align__(16) float4 st_smem[];
  const StLevel& lv = find_level(bt, (int)blockIdx.x);
  const int rel = (int)blockIdx.x - lv.blk0;
  const int tid = threadIdx.x;
  if (rel >= lv.n_sblocks) {
    temporal_fwd(lv, rel - lv.n_sblocks, tid);
    return;
  }
  // ------------------------------ spatial gradient (RGB_OFF.py:611 / Flow_OFF.py:622 / util.py:46-50) + dropout
  const offk_stencil_t& s = lv.s;
  const int H = s.H, W = s.W, HW = H * W, K = s.K, Cs = s.Cs, ctot = s.out_ctot, d_ps = s.d_ps;
  const int p = rel / lv.bands, band = rel - p * lv.bands;
  const int y0 = band * lv.band_rows;
  const int rows = min(lv.band_rows, H - y0);
  const int lq = lv.lq, ltcp = lv.ltcp, CQ = 1 << lq, TCP = 1 << ltcp;
  float4* tile = st_smem;                                        // [rows+2][TCP][CQ]
  float4* ws4 = st_smem + (((lv.band_rows + 2) << ltcp) << lq);  // [K*9][CQ]
  float4* bs4 = ws4 + K * 9 * CQ;                                // [K][CQ]
  {
    const float* dp = lv.d + (size_t)spatial_frame(s, p) * s.d_fs;
    const int n_cells = ((rows + 2) << ltcp) << lq;
    const uint32_t tile_s = st_smem_u32(tile);
    for (int i = tid; i < n_cells; i += ST_THREADS) {
      const int c4 = i & (CQ - 1), col = (i >> lq) & (TCP - 1), r = i >> (lq + ltcp);
      if (col < W + 2) {
        const int yy = y0 - 1 + r, xx = col - 1;
        const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
        st_cp_async16(tile_s + (uint32_t)i * 16u, in ? dp + (size_t)(yy * W + xx) * d_ps + c4 * 4 : dp, in ? 16u : 0u);
      }
    }
  }
  load_taps(lv, ws4, bs4, tid);
  st_cp_async_wait_all();
  __syncthreads();

  const int c4 = tid & (CQ - 1), slot = tid >> lq;
  const int XW = 1 << lv.lxw;
  const int rpp = (ST_THREADS >> lq) >> lv.lxw;                  // rows per pass
  const int sx = slot & (XW - 1), sy = slot >> lv.lxw;
  const uint32_t thr = drop_threshold16(s.drop_p);
  const float4* wk = ws4 + c4;
  const float4 b0 = bs4[c4], b1 = K > 1 ? bs4[CQ + c4] : f4zero();
  float* obase = lv.out + (size_t)p * HW * ctot + s.out_coff + c4 * 4;
  for (int yb = 0; yb < rows; yb += rpp) {
    const int y = yb + sy;
    for (int x0 = 0; x0 < W; x0 += XW) {
      const int x = x0 + sx;
      if (y < rows && x < W) {
        const float4* t0 = tile + (((y << ltcp) + x) << lq) + c4;     // halo coordinates: (y, x) = top-left neighbour
        float4 acc0 = b0, acc1 = b1;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int bb = 0; bb < 3; ++bb) {
            const float4 nb = t0[((a << ltcp) + bb) << lq];
            const float4 w0 = wk[(a * 3 + bb) << lq];
            acc0.x = fmaf(w0.x, nb.x, acc0.x);
            acc0.y = fmaf(w0.y, nb.y, acc0.y);
            acc0.z = fmaf(w0.z, nb.z, acc0.z);
            acc0.w = fmaf(w0.w, nb.w, acc0.w);
            if (K > 1) {
              const float4 w1 = wk[(9 + a * 3 + bb) << lq];
              acc1.x = fmaf(w1.x, nb.x, acc1.x);
              acc1.y = fmaf(w1.y, nb.y, acc1.y);
              acc1.z = fmaf(w1.z, nb.z, acc1.z);
              acc1.w = fmaf(w1.w, nb.w, acc1.w);
            }
          }
        const int pix = (y0 + y) * W + x;
        float* op = obase + (size_t)pix * ctot;
        const float4 k0 = drop_factor4(s, thr, p, c4 * 4, pix, HW);
        stg_stream4(op, make_float4(acc0.x * k0.x, acc0.y * k0.y, acc0.z * k0.z, acc0.w * k0.w));
        if (K > 1) {
          const float4 k1 = drop_factor4(s, thr, p, Cs + c4 * 4, pix, HW);
          stg_stream4(op + Cs, make_float4(acc1.x * k1.x, acc1.y * k1.y, acc1.z * k1.z, acc1.w * k1.w));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- backward
// dG[b,t] = (dT[b,t-1] - dT[b,t]) * (G[b,t] > 0)      (ReLU' of RGB_OFF.py:598); a streaming kernel of its own so that
// it keeps a high occupancy (the spatial kernel below needs ~40 accumulator registers)
__global__ void __launch_bounds__(ST_THREADS, 4)
stencil_diff_bwd_temporal_kernel(const __grid_constant__ StBatch bt) {
  const StLevel& lv = find_level(bt, (int)blockIdx.x);
  const int r2 = (int)blockIdx.x - lv.blk0;
  const int tid = threadIdx.x;
  const int L = lv.s.L, Cg = lv.s.Cg, HW = lv.s.H * lv.s.W;
  const int g_ps = lv.s.g_ps, ctot = lv.s.out_ctot;
  const size_t g_fs = (size_t)lv.s.g_fs, dg_fs = (size_t)lv.dg_fs, o_fs = (size_t)HW * ctot;
  const int b = r2 / lv.t_chunks, chunk = r2 - b * lv.t_chunks;
  const int warp = tid >> 5, lane = tid & 31;
  const int pix0 = chunk * (ST_WARPS * ST_TJ) + warp * ST_TJ;
  bool ok[ST_TJ];
#pragma unroll
  for (int j = 0; j < ST_TJ; ++j) ok[j] = pix0 + j < HW;
  for (int c = lane * 4; c < Cg; c += 128) {
    const float* gp = lv.g + (size_t)b * L * g_fs + (size_t)pix0 * g_ps + c;
    float* dgp = lv.dg + (size_t)b * L * dg_fs + (size_t)pix0 * g_ps + c;        // dg uses the pixel stride of g
    const float* op = lv.dout + ((size_t)b * (L - 1) * HW + pix0) * ctot + lv.s.out_coff + lv.s.K * lv.s.Cs + c;
    float4 prev[ST_TJ];
#pragma unroll
    for (int j = 0; j < ST_TJ; ++j) prev[j] = f4zero();
    for (int t0 = 0; t0 < L; t0 += ST_BC) {
      float4 gv[ST_BC][ST_TJ], dt[ST_BC][ST_TJ];
#pragma unroll
      for (int i = 0; i < ST_BC; ++i) {
#pragma unroll
        for (int j = 0; j < ST_TJ; ++j) {
          gv[i][j] = (ok[j] && t0 + i < L) ? ldg_stream4(gp + j * g_ps) : f4zero();
          dt[i][j] = (ok[j] && t0 + i < L - 1) ? ldg_stream4(op + j * ctot) : f4zero();
        }
        gp += g_fs;
        op += o_fs;
      }
#pragma unroll
      for (int i = 0; i < ST_BC; ++i) {
        if (t0 + i < L) {
#pragma unroll
          for (int j = 0; j < ST_TJ; ++j) {
            float4 r = f4sub(prev[j], dt[i][j]);
            r.x = gv[i][j].x > 0.f ? r.x : 0.f;
            r.y = gv[i][j].y > 0.f ? r.y : 0.f;
            r.z = gv[i][j].z > 0.f ? r.z : 0.f;
            r.w = gv[i][j].w > 0.f ? r.w : 0.f;
            if (ok[j]) stg_stream4(dgp + j * g_ps, r);
            prev[j] = dt[i][j];
          }
        }
        dgp += dg_fs;
      }
    }
  }
}

// spatial: dD = transposed stencil of the dropped dS; tap / bias gradients.  Persistent blocks over (frame, band) items.
__global__ void __launch_bounds__(ST_THREADS, 2)
stencil_diff_bwd_spatial_kernel(const __grid_constant__ StBatch bt) {
  extern __shared__ __align__(16) float4 st_smem[];
  const StLevel& lv = find_level(bt, (int)blockIdx.x);
  const offk_stencil_t& s = lv.s;
  const int rel = (int)blockIdx.x - lv.blk0;
  const int tid = threadIdx.x;
  const int H = s.H, W = s.W, HW = H * W, K = s.K, Cs = s.Cs, ctot = s.out_ctot, d_ps = s.d_ps;
  const int lq = lv.lq, ltcp = lv.ltcp, CQ = 1 << lq, TCP = 1 << ltcp;
  float4* tile = st_smem;                                        // dS band + halo [rows+2][TCP][CQ]
  float4* ws4 = st_smem + (((lv.band_rows + 2) << ltcp) << lq);  // [K*9][CQ]
  float* red = reinterpret_cast<float*>(ws4 + K * 9 * CQ);       // [10][Cs] block reduction of the tap gradients
  load_taps(lv, ws4, nullptr, tid);
  const bool need_dw = lv.dw != nullptr, need_db = lv.dbias != nullptr;
  const bool need_acc = need_dw || need_db;
  for (int i = tid; i < 10 * Cs; i += ST_THREADS) red[i] = 0.f;
  const int c4 = tid & (CQ - 1), slot = tid >> lq;
  const int XW = 1 << lv.lxw;
  const int rpp = (ST_THREADS >> lq) >> lv.lxw;
  const int sx = slot & (XW - 1), sy = slot >> lv.lxw;
  const uint32_t thr = drop_threshold16(s.drop_p);

  for (int kk = 0; kk < K; ++kk) {
    float4 wacc[9], bacc = f4zero();
#pragma unroll
    for (int j = 0; j < 9; ++j) wacc[j] = f4zero();
    const int och = kk * Cs + c4 * 4;
    const float4* wk = ws4 + kk * 9 * CQ + c4;
    for (int item = rel; item < lv.n_sitems; item += lv.n_sblocks) {
      const int f = item / lv.bands, band = item - f * lv.bands;
      const int y0 = band * lv.band_rows;
      const int rows = min(lv.band_rows, H - y0);
      const int p = spatial_pair_of_frame(s, f);
      __syncthreads();                                           // previous item's tile fully consumed (and ws4 / red ready)
      if (p >= 0) {
        const float* sp = lv.dout + (size_t)p * HW * ctot + s.out_coff + och;
        const int n_cells = ((rows + 2) << ltcp) << lq;
        for (int i = tid; i < n_cells; i += ST_THREADS) {        // (the cell's c4 equals this thread's c4: 256 % CQ == 0)
          const int col = (i >> lq) & (TCP - 1), r = i >> (lq + ltcp);
          if (col < W + 2) {
            const int yy = y0 - 1 + r, xx = col - 1;
            float4 v = f4zero();
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
              const int np_ = yy * W + xx;
              v = ldg4(sp + (size_t)np_ * ctot);
              const float4 k4 = drop_factor4(s, thr, p, och, np_, HW);
              v.x *= k4.x; v.y *= k4.y; v.z *= k4.z; v.w *= k4.w;
            }
            tile[i] = v;
          }
        }
      }
      __syncthreads();
      float* ddf = lv.dd + (size_t)f * lv.dd_fs + c4 * 4;          // dd uses the pixel stride of d
      const float* dfp = (need_dw && p >= 0) ? lv.d + (size_t)spatial_frame(s, p) * s.d_fs + c4 * 4 : nullptr;
      for (int yb = 0; yb < rows; yb += rpp) {
        const int y = yb + sy;
        for (int x0 = 0; x0 < W; x0 += XW) {
          const int x = x0 + sx;
          if (y < rows && x < W) {
            const int pix = (y0 + y) * W + x;
            float4 acc = f4zero();
            if (p >= 0) {
              float4 dv = f4zero();
              if (dfp) dv = ldg4(dfp + (size_t)pix * d_ps);
              // dS(y-a+1, x-b+1) sits at halo coordinates (y+2-a, x+2-b)
              const float4* t2 = tile + ((((y + 2) << ltcp) + x + 2) << lq) + c4;
#pragma unroll
              for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int bb = 0; bb < 3; ++bb) {
                  const float4 nv = *(t2 - (((a << ltcp) + bb) << lq));
                  const float4 wv = wk[(a * 3 + bb) << lq];
                  acc.x = fmaf(wv.x, nv.x, acc.x);
                  acc.y = fmaf(wv.y, nv.y, acc.y);
                  acc.z = fmaf(wv.z, nv.z, acc.z);
                  acc.w = fmaf(wv.w, nv.w, acc.w);
                  wacc[a * 3 + bb].x = fmaf(dv.x, nv.x, wacc[a * 3 + bb].x);
                  wacc[a * 3 + bb].y = fmaf(dv.y, nv.y, wacc[a * 3 + bb].y);
                  wacc[a * 3 + bb].z = fmaf(dv.z, nv.z, wacc[a * 3 + bb].z);
                  wacc[a * 3 + bb].w = fmaf(dv.w, nv.w, wacc[a * 3 + bb].w);
                  if (a == 1 && bb == 1) { bacc.x += nv.x; bacc.y += nv.y; bacc.z += nv.z; bacc.w += nv.w; }
                }
            }
            float4* o = reinterpret_cast<float4*>(ddf + (size_t)pix * d_ps);
            if (kk > 0) {
              const float4 old = *o;
              acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w;
            }
            *o = acc;                                              // zero for frames that feed no pair
          }
        }
      }
    }
    if (need_acc) {
      // lanes l, l^CQ, l^2CQ, ... hold the same channels
      float vals[40];
#pragma unroll
      for (int j = 0; j < 9; ++j) { vals[4 * j] = wacc[j].x; vals[4 * j + 1] = wacc[j].y; vals[4 * j + 2] = wacc[j].z; vals[4 * j + 3] = wacc[j].w; }
      vals[36] = bacc.x; vals[37] = bacc.y; vals[38] = bacc.z; vals[39] = bacc.w;
#pragma unroll
      for (int q = 0; q < 40; ++q) {
        float v = vals[q];
        for (int o = 16; o >= CQ; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) < CQ) atomicAdd(&red[(q >> 2) * Cs + c4 * 4 + (q & 3)], v);
      }
      __syncthreads();
      for (int i = tid; i < 10 * Cs; i += ST_THREADS) {
        const int j = i / Cs, c = i - j * Cs;
        const float v = red[i];
        red[i] = 0.f;
        if (j < 9) {
          if (need_dw) atomicAdd(lv.dw + ((size_t)c * K + kk) * 9 + j, v);
        } else if (need_db) {
          atomicAdd(lv.dbias + kk * Cs + c, v);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- host
static int check_stencil(const offk_stencil_t* s) {
  OFFK_REQUIRE(s != nullptr, "stencil: null descriptor");
  OFFK_REQUIRE(s->B >= 1 && s->L >= 2, "stencil: need B >= 1 and L >= 2 (got B=%d L=%d)", s->B, s->L);
  OFFK_REQUIRE(s->Cg >= 0 && s->Cs >= 0 && s->Cg + s->Cs > 0, "stencil: bad channel counts");
  OFFK_REQUIRE(s->Cg % 4 == 0 && s->Cs % 4 == 0 && s->Cs <= ST_MAX_CS, "stencil: channels must be multiples of 4, Cs <= %d",
               ST_MAX_CS);
  OFFK_REQUIRE(s->Cs == 0 || (((s->Cs >> 2) & ((s->Cs >> 2) - 1)) == 0), "stencil: Cs/4 must be a power of two");
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
static int ilog2_ceil(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// geometry of the spatial blocks of one level; returns the dynamic shared memory the level needs (0 if Cs == 0)
static int plan_spatial(StLevel& lv, bool backward) {
  const offk_stencil_t& s = lv.s;
  lv.lq = lv.ltcp = lv.lxw = 0;
  lv.bands = lv.band_rows = 1;
  lv.n_sitems = lv.n_sblocks = 0;
  if (s.Cs == 0) return 0;
  const int CQ = s.Cs / 4;
  lv.lq = ilog2_ceil(CQ);
  lv.ltcp = ilog2_ceil(s.W + 2);
  const int row_bytes = (1 << lv.ltcp) * CQ * 16;
  int max_rows = ST_TILE_BUDGET / row_bytes - 2;
  if (max_rows < 1) max_rows = 1;
  lv.bands = (s.H + max_rows - 1) / max_rows;
  lv.band_rows = (s.H + lv.bands - 1) / lv.bands;
  lv.bands = (s.H + lv.band_rows - 1) / lv.band_rows;
  const int slots = ST_THREADS / CQ;
  int lxw = ilog2_ceil(s.W);
  while ((1 << lxw) > slots) --lxw;
  lv.lxw = lxw;
  const long long units = backward ? (long long)s.B * s.L : (long long)s.B * (s.L - 1);
  const long long items = units * lv.bands;
  lv.n_sitems = (int)items;
  lv.n_sblocks = (int)items;
  if (backward && (lv.dw || lv.dbias)) {
    const long long cap = 2LL * sm_count();      // persistent: register-resident tap-gradient partial sums
    if (items > cap) lv.n_sblocks = (int)cap;
  }
  const int tile_bytes = (lv.band_rows + 2) * row_bytes;
  const int tap_bytes = s.K * 9 * CQ * 16;
  const int extra = backward ? 10 * s.Cs * 4 : s.K * CQ * 16;
  return tile_bytes + tap_bytes + extra;
}

static int launch_batch(int n, const offk_stencil_t* s, const offk_stencil_io_t* io, bool backward, void* stream) {
  OFFK_REQUIRE(n >= 1 && n <= ST_MAX_LEVELS, "stencil batch: 1 <= n <= %d (got %d)", ST_MAX_LEVELS, n);
  OFFK_REQUIRE(s != nullptr && io != nullptr, "stencil batch: null arrays");
  StBatch bt;
  bt.n = n;
  long long n_t[ST_MAX_LEVELS];
  int smem = 0;
  for (int i = 0; i < n; ++i) {
    if (int e = check_stencil(&s[i])) return e;
    StLevel& lv = bt.lv[i];
    lv.s = s[i];
    lv.g = io[i].g; lv.d = io[i].d; lv.w = io[i].w; lv.bias = io[i].bias; lv.out = io[i].out;
    lv.dout = io[i].dout; lv.dg = io[i].dg; lv.dd = io[i].dd; lv.dw = io[i].dw; lv.dbias = io[i].dbias;
    lv.dg_fs = io[i].dg_fs; lv.dd_fs = io[i].dd_fs;
    if (!backward) {
      OFFK_REQUIRE(lv.out != nullptr && aligned16(lv.out), "stencil_fwd: out must be non-null and 16-byte aligned");
      OFFK_REQUIRE(s[i].Cg == 0 || (lv.g && aligned16(lv.g)), "stencil_fwd: g must be 16-byte aligned");
      OFFK_REQUIRE(s[i].Cs == 0 || (lv.d && lv.w && aligned16(lv.d)), "stencil_fwd: d / w missing or unaligned");
      OFFK_REQUIRE(lv.bias == nullptr || aligned16(lv.bias), "stencil_fwd: bias must be 16-byte aligned");
    } else {
      OFFK_REQUIRE(lv.dout != nullptr && aligned16(lv.dout), "stencil_bwd: dout must be 16-byte aligned");
      OFFK_REQUIRE(s[i].Cg == 0 || (lv.g && lv.dg && aligned16(lv.g) && aligned16(lv.dg) && lv.dg_fs % 4 == 0), "stencil_bwd: g/dg");
      OFFK_REQUIRE(s[i].Cs == 0 || (lv.w && lv.dd && aligned16(lv.dd) && lv.dd_fs % 4 == 0), "stencil_bwd: w/dd missing");
      OFFK_REQUIRE(lv.dw == nullptr || (lv.d != nullptr && aligned16(lv.d)), "stencil_bwd: tap gradient needs d");
    }
    const int need = plan_spatial(lv, backward);
    OFFK_REQUIRE(need <= ST_SMEM_MAX, "stencil: halo tile of %d bytes exceeds %d (W=%d, Cs=%d)", need, ST_SMEM_MAX, s[i].W, s[i].Cs);
    if (need > smem) smem = need;
    const int HW = s[i].H * s[i].W;
    const int per_blk = ST_WARPS * ST_TJ;
    lv.t_chunks = s[i].Cg > 0 ? (HW + per_blk - 1) / per_blk : 1;
    n_t[i] = s[i].Cg > 0 ? (long long)lv.t_chunks * s[i].B : 0;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(stencil_diff_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM_MAX);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(stencil_diff_bwd_spatial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM_MAX);
    if (e != cudaSuccess) return cuda_check(e, "cudaFuncSetAttribute(stencil)");
    attr_set = true;
  }
  if (!backward) {
    // one grid: per level, spatial blocks first (heavier per byte), then temporal blocks
    long long blk = 0;
    for (int i = 0; i < n; ++i) {
      bt.lv[i].blk0 = (int)blk;
      blk += bt.lv[i].n_sblocks + n_t[i];
      OFFK_REQUIRE(blk < 2147483647LL, "stencil: grid too large");
    }
    if (blk == 0) return 0;
    stencil_diff_fwd_kernel<<<(unsigned)blk, ST_THREADS, smem, as_stream(stream)>>>(bt);
    return OFFK_LAUNCH_CHECK("stencil_diff_fwd");
  }
  // backward: a streaming kernel for dG and a persistent kernel for dD + tap gradients
  long long blk = 0;
  for (int i = 0; i < n; ++i) {
    bt.lv[i].blk0 = (int)blk;
    blk += n_t[i];
    OFFK_REQUIRE(blk < 2147483647LL, "stencil: grid too large");
  }
  if (blk > 0) {
    stencil_diff_bwd_temporal_kernel<<<(unsigned)blk, ST_THREADS, 0, as_stream(stream)>>>(bt);
    if (int e = OFFK_LAUNCH_CHECK("stencil_diff_bwd_temporal")) return e;
  }
  blk = 0;
  for (int i = 0; i < n; ++i) {
    bt.lv[i].blk0 = (int)blk;
    blk += bt.lv[i].n_sblocks;
  }
  if (blk > 0) {
    stencil_diff_bwd_spatial_kernel<<<(unsigned)blk, ST_THREADS, smem, as_stream(stream)>>>(bt);
    return OFFK_LAUNCH_CHECK("stencil_diff_bwd_spatial");
  }
  return 0;
}

}  // namespace offk

using namespace offk;

extern "C" int offk_stencil_diff_fwd_batch(int n, const offk_stencil_t* s, const offk_stencil_io_t* io, void* stream) {
  return launch_batch(n, s, io, false, stream);
}

extern "C" int offk_stencil_diff_bwd_batch(int n, const offk_stencil_t* s, const offk_stencil_io_t* io, void* stream) {
  return launch_batch(n, s, io, true, stream);
}

extern "C" int offk_stencil_diff_fwd(const offk_stencil_t* s, const float* g, const float* d, const float* w,
                                     const float* bias, float* out, void* stream) {
  offk_stencil_io_t io = {};
  io.g = g; io.d = d; io.w = w; io.bias = bias; io.out = out;
  return launch_batch(1, s, &io, false, stream);
}

extern "C" int offk_stencil_diff_bwd(const offk_stencil_t* s, const float* dout, const float* g, const float* d,
                                     const float* w, float* dg, int64_t dg_fs, float* dd, int64_t dd_fs, float* dw,
                                     float* dbias, void* stream) {
  offk_stencil_io_t io = {};
  io.g = g; io.d = d; io.w = w; io.dout = dout; io.dg = dg; io.dg_fs = dg_fs; io.dd = dd; io.dd_fs = dd_fs;
  io.dw = dw; io.dbias = dbias;
  return launch_batch(1, s, &io, true, stream);
}
