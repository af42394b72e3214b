// tcgen05 gather-GEMM (OFFK_PREC_TF32) for sm_100a.
//
//   D[128 x BN] (fp32, TMEM) += A[128 x 32] (tf32, smem) * B[BN x 32]^T (tf32, smem)   per K-block
//
// Warp roles (288 threads):
//   warps 0-7  producers: gather A and B elements through the index tables (any layout: NCHW 1x1,
//              KxK im2col, col2im for data gradients, pixel-major for weight gradients), write them
//              into the canonical K-major SWIZZLE_128B shared-memory layout, fence to the async
//              proxy and arrive on the stage's "full" mbarrier.  After the main loop the same
//              warps run the epilogue: tcgen05.ld the accumulator rows out of TMEM, apply
//              bias / ReLU / gate / residual, store coalesced along pixels (or atomically for
//              split-K and weight gradients).
//   warp 8     allocates TMEM; one elected lane waits on "full", issues 4 x tcgen05.mma
//              (kind::tf32, M=128, N=BN, K=8) per stage and tcgen05.commit's to the stage's "empty"
//              mbarrier (and to "accum_full" after the last K-block).
// One output tile per CTA; two CTAs co-reside per SM (<=110 KB smem, <=256 TMEM columns each) so one
// CTA's epilogue overlaps the other's main loop.
#include "offk_gemm.cuh"

namespace offk {

constexpr int TC_BM = 128;            // UMMA M
constexpr int TC_BK = 32;             // K-block: 32 tf32 = one 128-byte swizzle row
constexpr int TC_PRODUCERS = 256;     // producer / epilogue threads (8 warps)
constexpr int TC_THREADS = TC_PRODUCERS + 32;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;  // 16 KB
constexpr int TC_MAX_STAGES = 8;

// ---------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "OFFK_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra OFFK_DONE_%=;\n"
      "bra OFFK_WAIT_%=;\n"
      "OFFK_DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued tcgen05.mma of this thread arrive on `bar` when complete
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns of this warp's TMEM quadrant
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 format, cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1)
//   [32,46) stride byte offset >> 4 (1024 B between 8-row groups) | [46,48) version = 1 | [61,64) layout = 2
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2,
// A,B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
// byte offset of (row r, 16-byte chunk c) inside a K-major SWIZZLE_128B tile (8-row atoms of 1024 B)
__device__ __forceinline__ uint32_t swz(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float a) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory");
}

// ---------------------------------------------------------------------------- producers
struct TcShared {
  uint64_t full[TC_MAX_STAGES];
  uint64_t empty[TC_MAX_STAGES];
  uint64_t accum_full;
  uint32_t tmem_base;
};

// A tile, lanes along rows (sources contiguous along m: NCHW activations with m = pixel).
// Thread owns row (tid & 127) and the four 16-byte chunks [4*half, 4*half+4) of the 128-byte K row.
struct ARowLane {
  offk_idx_t r;
  bool rvalid, ones;
  int half;
  __device__ __forceinline__ void init(const offk_gemm_t& g, int m0, int tid) {
    const int m = m0 + (tid & 127);
    half = tid >> 7;
    rvalid = m < g.M;
    ones = rvalid && (m == g.a_ones_row);
    r = (rvalid && !ones) ? g.a_row[m] : offk_idx_t{0, 0, 0};
  }
  __device__ __forceinline__ void load(const offk_gemm_t& g, int k0, int k_end, float (&v)[16]) const {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int k = k0 + half * 16 + i;
      float x = 0.f;
      if (rvalid && k < k_end) x = gemm_load_a(g, r, g.a_col[k], ones);
      v[i] = x;
    }
  }
  __device__ __forceinline__ void store(uint32_t a_base, int tid, const float (&v)[16]) const {
    const int row = tid & 127;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      sts128(a_base + swz(row, half * 4 + c), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
  }
};

// A tile, lanes along k (sources contiguous along k: weight-gradient GEMMs where k = pixel).
// Lane owns column k0+lane; warp w owns rows w, w+8, ... (16 rows).
struct AKLane {
  __device__ __forceinline__ void load(const offk_gemm_t& g, int m0, int k0, int k_end, int warp, int lane,
                                       float (&v)[16]) const {
    const int k = k0 + lane;
    const bool kvalid = k < k_end;
    const offk_idx_t c = kvalid ? g.a_col[k] : offk_idx_t{0, 0, 0};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int m = m0 + warp + 8 * i;
      float x = 0.f;
      if (kvalid && m < g.M) {
        const bool ones = (m == g.a_ones_row);
        const offk_idx_t r = ones ? offk_idx_t{0, 0, 0} : g.a_row[m];
        x = gemm_load_a(g, r, c, ones);
      }
      v[i] = x;
    }
  }
  __device__ __forceinline__ void store(uint32_t a_base, int warp, int lane, const float (&v)[16]) const {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int row = warp + 8 * i;
      sts32(a_base + swz(row, lane >> 2) + (lane & 3) * 4, v[i]);
    }
  }
};

// B tile (bn rows x 32 k).  KLANE: lanes along k, warp w owns rows w, w+8, ...   (up to 32 rows/warp)
// ROWLANE: lanes along rows: thread owns row (tid % bn_pad) ... implemented as a strided loop.
template <bool KLANE>
__device__ __forceinline__ void load_store_b(const offk_gemm_t& g, uint32_t b_base, int n0, int bn, int k0,
                                             int k_end, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  if (KLANE) {
    const int k = k0 + lane;
    const bool kvalid = k < k_end;
    const int c = kvalid ? g.b_col[k] : 0;
    for (int row = warp; row < bn; row += 8 * 4) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = row + 8 * i;
        const int n = n0 + rr;
        v[i] = (kvalid && rr < bn && n < g.N) ? __ldg(g.b_src + (g.b_row[n] + c)) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = row + 8 * i;
        if (rr < bn) sts32(b_base + swz(rr, lane >> 2) + (lane & 3) * 4, v[i]);
      }
    }
  } else {
    // lanes along rows; each thread gathers one 16-byte chunk (4 consecutive k) of one row
    for (int item = tid; item < bn * 8; item += TC_PRODUCERS) {
      const int rr = item % bn, chunk = item / bn;
      const int n = n0 + rr;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (n < g.N) {
        const int rb = g.b_row[n];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = k0 + chunk * 4 + e;
          if (k < k_end) v[e] = __ldg(g.b_src + (rb + g.b_col[k]));
        }
      }
      sts128(b_base + swz(rr, chunk), v[0], v[1], v[2], v[3]);
    }
  }
}

// B tile when b_col is the identity and rows are 16-byte aligned (dense [N,K] weights): one float4
// per (row, chunk); 8 lanes cover a 128-byte row -> conflict-free swizzled STS.128.
__device__ __forceinline__ void load_store_b_vec4(const offk_gemm_t& g, uint32_t b_base, int n0, int bn, int k0,
                                                  int k_end, int tid) {
  for (int item = tid; item < bn * 8; item += TC_PRODUCERS) {
    const int rr = item >> 3, chunk = item & 7;
    const int n = n0 + rr, k = k0 + chunk * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < g.N && k < k_end) v = __ldg(reinterpret_cast<const float4*>(g.b_src + (g.b_row[n] + k)));
    sts128(b_base + swz(rr, chunk), v.x, v.y, v.z, v.w);
  }
}

// ---------------------------------------------------------------------------- kernel
template <bool A_KLANE, int B_MODE /*0 rowlane, 1 klane, 2 dense vec4*/>
__global__ void __launch_bounds__(TC_THREADS, 2)
gather_gemm_tc_kernel(const offk_gemm_t g, int bn, int stages, int kb_per_split, int tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A 16 KB | B bn*128)] then the barrier block
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = (uint32_t)bn * 128u;
  const uint32_t stage_bytes = TC_A_BYTES + ((b_bytes + 1023u) & ~1023u);
  TcShared* sh = reinterpret_cast<TcShared*>(smem_raw + (smem_base - smem_u32(smem_raw)) + stages * stage_bytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * bn;
  const int num_kb_total = (g.K + TC_BK - 1) / TC_BK;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int kb_end = min(num_kb_total, kb_begin + kb_per_split);
  const int nkb = kb_end - kb_begin;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&sh->full[s]), TC_PRODUCERS);
      mbar_init(smem_u32(&sh->empty[s]), 1);
    }
    mbar_init(smem_u32(&sh->accum_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) tmem_alloc(smem_u32(&sh->tmem_base), (uint32_t)tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = sh->tmem_base;

  if (warp < 8) {
    // ================= producers =================
    ARowLane arow;
    AKLane akl;
    if (!A_KLANE) arow.init(g, m0, tid);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % stages;
      const uint32_t round = (uint32_t)(i / stages);
      const int k0 = (kb_begin + i) * TC_BK;
      const uint32_t a_base = smem_base + s * stage_bytes;
      const uint32_t b_base = a_base + TC_A_BYTES;
      float va[16];
      if (A_KLANE) akl.load(g, m0, k0, g.K, warp, lane, va);
      else         arow.load(g, k0, g.K, va);
      mbar_wait(smem_u32(&sh->empty[s]), (round & 1u) ^ 1u);   // slot free (first round passes at once)
      if (A_KLANE) akl.store(a_base, warp, lane, va);
      else         arow.store(a_base, tid, va);
      if (B_MODE == 2)      load_store_b_vec4(g, b_base, n0, bn, k0, g.K, tid);
      else if (B_MODE == 1) load_store_b<true>(g, b_base, n0, bn, k0, g.K, tid);
      else                  load_store_b<false>(g, b_base, n0, bn, k0, g.K, tid);
      fence_proxy_async_smem();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(smem_u32(&sh->full[s]));
    }
  } else {
   if (lane == 0) {
    // ================= MMA issuer (one thread) =================
    const uint32_t idesc = make_idesc_tf32(bn);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % stages;
      const uint32_t round = (uint32_t)(i / stages);
      mbar_wait(smem_u32(&sh->full[s]), round & 1u);
      tc_fence_after();
      const uint32_t a_base = smem_base + s * stage_bytes;
      const uint32_t b_base = a_base + TC_A_BYTES;
      const uint64_t adesc = make_smem_desc(a_base), bdesc = make_smem_desc(b_base);
#pragma unroll
      for (int j = 0; j < TC_BK / 8; ++j) {
        // advance 8 tf32 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
        umma_tf32(tmem_d, adesc + (uint64_t)(2 * j), bdesc + (uint64_t)(2 * j), idesc, (i > 0 || j > 0) ? 1u : 0u);
      }
      umma_commit(smem_u32(&sh->empty[s]));          // frees the smem slot when these MMAs retire
    }
    umma_commit(smem_u32(&sh->accum_full));          // accumulator complete
   }
   __syncwarp();
  }

  // ================= epilogue (warps 0-7) =================
  if (warp < 8 && nkb > 0) {
    mbar_wait(smem_u32(&sh->accum_full), 0u);
    tc_fence_after();
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may read
    const int m = m0 + quad * 32 + lane;
    const bool mvalid = m < g.M;
    EpiRow er = {0, 0, 0, false};
    if (mvalid) er = epi_row(g, m);
    const bool atomic = (g.split_k > 1) || g.atomic_out;
    const int nchunks = bn >> 4;
    for (int c = (warp >> 2); c < nchunks; c += 2) {  // warps 0-3 even 16-column chunks, 4-7 odd ones
      float v[16];
      tmem_ld16(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 16), v);
      if (mvalid) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = n0 + c * 16 + j;
          if (n < g.N) epi_store(g, er, n, v[j], atomic);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc(tmem_d, (uint32_t)tmem_cols);
  }
}

// ---------------------------------------------------------------------------- host
static int pick_bn(int N) {
  // one N tile when it fits a single UMMA (N <= 256, multiple of 16); otherwise 256/128-wide tiles
  const int n16 = (N + 15) / 16 * 16;
  if (n16 <= 256) return n16;
  if (N % 256 == 0) return 256;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  // minimise padding
  int best = 256, waste = 1 << 30;
  for (int bn = 256; bn >= 128; bn -= 16) {
    const int w = (N + bn - 1) / bn * bn - N;
    if (w < waste) { waste = w; best = bn; }
  }
  return best;
}

template <bool A_KLANE, int B_MODE>
static int launch_tc_t(const offk_gemm_t& g, int bn, int stages, int kb_per, int tmem_cols, dim3 grid, size_t smem,
                       cudaStream_t st) {
  auto kern = gather_gemm_tc_kernel<A_KLANE, B_MODE>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_check(e, "cudaFuncSetAttribute(gather_gemm_tc)");
    attr_set = true;
  }
  kern<<<grid, TC_THREADS, smem, st>>>(g, bn, stages, kb_per, tmem_cols);
  return OFFK_LAUNCH_CHECK("gather_gemm_tc");
}

int launch_gemm_tc(const offk_gemm_t& g, cudaStream_t st) {
  int bn = g.tile_n > 0 ? g.tile_n : pick_bn(g.N);
  if (bn % 16 != 0 || bn < 16 || bn > 256) return fail(OFFK_E_BADARG, "gather_gemm: bad N tile %d", bn);
  const int num_kb = (g.K + TC_BK - 1) / TC_BK;
  const int split = g.split_k > 1 ? g.split_k : 1;
  const int kb_per = (num_kb + split - 1) / split;
  const uint32_t stage_bytes = TC_A_BYTES + ((bn * 128 + 1023) & ~1023);
  int stages = (108 * 1024) / (int)stage_bytes;
  if (stages < 2) stages = 2;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages > kb_per) stages = kb_per < 2 ? 2 : kb_per;
  const size_t smem = (size_t)stages * stage_bytes + sizeof(TcShared) + 1024;
  int tmem_cols = 32;
  while (tmem_cols < bn) tmem_cols <<= 1;
  dim3 grid((g.M + TC_BM - 1) / TC_BM, (g.N + bn - 1) / bn, (num_kb + kb_per - 1) / kb_per);
  if (grid.y > 65535 || grid.z > 65535) return fail(OFFK_E_LIMIT, "gather_gemm: grid too large");
  const int bmode = g.b_dense ? 2 : (g.b_klane ? 1 : 0);
  if (g.a_klane) {
    if (bmode == 2) return launch_tc_t<true, 2>(g, bn, stages, kb_per, tmem_cols, grid, smem, st);
    if (bmode == 1) return launch_tc_t<true, 1>(g, bn, stages, kb_per, tmem_cols, grid, smem, st);
    return launch_tc_t<true, 0>(g, bn, stages, kb_per, tmem_cols, grid, smem, st);
  }
  if (bmode == 2) return launch_tc_t<false, 2>(g, bn, stages, kb_per, tmem_cols, grid, smem, st);
  if (bmode == 1) return launch_tc_t<false, 1>(g, bn, stages, kb_per, tmem_cols, grid, smem, st);
  return launch_tc_t<false, 0>(g, bn, stages, kb_per, tmem_cols, grid, smem, st);
}

}  // namespace offk
