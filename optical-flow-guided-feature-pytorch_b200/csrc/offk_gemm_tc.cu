// tcgen05 gather-GEMM (OFFK_PREC_TF32) for sm_100a.
//
//   D[128 x BN] (fp32, TMEM) += A[128 x 32] (tf32, smem) * B[BN x 32]^T (tf32, smem)   per K-block
//
// Warp roles (288 threads):
//   warps 0-7  producers: gather A and B through the index tables (NHWC implicit-GEMM im2col / col2im,
//              NCHW taps, pixel-major weight-gradient operands), with 16-byte loads wherever the
//              caller promises contiguity (OFFK_LOAD_VEC_K: straight LDG.128 -> STS.128;
//              OFFK_LOAD_VEC_ROW: LDG.128 along rows + 4x4 register transpose), write them into the
//              canonical K-major SWIZZLE_128B shared-memory layout, fence to the async proxy and
//              arrive on the stage's "full" mbarrier.  After the main loop the same
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

// 16-byte asynchronous global->shared copy; src_bytes = 0 zero-fills the destination (padding / out-of-box taps)
__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const float* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const float* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// this thread's arrival on `bar` fires once all its prior cp.async have landed (counts as one expected arrival)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// ---------------------------------------------------------------------------- producers
struct TcShared {
  uint64_t full[TC_MAX_STAGES];
  uint64_t empty[TC_MAX_STAGES];
  uint64_t accum_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ float4 ldg128(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4relu(float4 v) {
  return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
}

// Operand views: A has the validity box / ReLU-on-load / ones row, B is a plain gather.
struct OpA {
  const offk_gemm_t& g;
  __device__ __forceinline__ OpA(const offk_gemm_t& g_) : g(g_) {}
  __device__ __forceinline__ int rows() const { return g.M; }
  __device__ __forceinline__ const float* src() const { return g.a_src; }
  __device__ __forceinline__ offk_idx_t row(int m) const { return g.a_row[m]; }
  __device__ __forceinline__ offk_idx_t col(int k) const { return g.a_col[k]; }
  __device__ __forceinline__ bool ok(offk_idx_t r, offk_idx_t c) const {
    return (g.a_h == 0) || ((unsigned)((int)r.y + (int)c.y) < (unsigned)g.a_h &&
                            (unsigned)((int)r.x + (int)c.x) < (unsigned)g.a_w);
  }
  __device__ __forceinline__ bool relu() const { return g.a_relu != 0; }
  __device__ __forceinline__ int ones_row() const { return g.a_ones_row; }
};
struct OpB {
  const offk_gemm_t& g;
  __device__ __forceinline__ OpB(const offk_gemm_t& g_) : g(g_) {}
  __device__ __forceinline__ int rows() const { return g.N; }
  __device__ __forceinline__ const float* src() const { return g.b_src; }
  __device__ __forceinline__ offk_idx_t row(int n) const { return offk_idx_t{g.b_row[n], 0, 0}; }
  __device__ __forceinline__ offk_idx_t col(int k) const { return offk_idx_t{g.b_col[k], 0, 0}; }
  __device__ __forceinline__ bool ok(offk_idx_t, offk_idx_t) const { return true; }
  __device__ __forceinline__ bool relu() const { return false; }
  __device__ __forceinline__ int ones_row() const { return -1; }
};

// One pass = a [128 rows x 32 k] sub-tile (rows row0..row0+127 of the operand tile, limited to `nrows`).
// Each of the 256 producer threads moves 16 floats of it per K-block.  The column-table entries a thread needs
// for a K-block ("Cols") are fetched one or two K-blocks ahead so that no table load sits on the critical path.
struct Cols {
  offk_idx_t c[4];
  uint32_t valid;   // bit e: entry e is a real column (k < K)
};

template <int MODE, class OP>
struct TileLoader {
  // ---- per-pass row state (constant over the K loop)
  offk_idx_t r[4];
  uint32_t dst[4];   // swizzled byte offset of this thread's first chunk of row (group) i inside the tile
  uint32_t flags;    // bit i: row (group) i valid, bit 8+i: it is the ones row
  __device__ __forceinline__ void init(const OP& op, int row_base /*global row of tile row 0*/, int row0, int nrows,
                                       int tid) {
    flags = 0;
    const int warp = tid >> 5, lane = tid & 31;
    if (MODE == OFFK_LOAD_VEC_K) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int tr = row0 + (tid >> 3) + 32 * i;
        const int m = row_base + tr;
        const bool intile = tr < nrows;
        const bool v = intile && m < op.rows();
        const bool one = v && m == op.ones_row();
        r[i] = (v && !one) ? op.row(m) : offk_idx_t{0, 0, 0};
        dst[i] = swz(tr, tid & 7);
        flags |= (v ? 1u : 0u) << i;
        flags |= (one ? 1u : 0u) << (8 + i);
        flags |= (intile ? 1u : 0u) << (16 + i);
      }
    } else if (MODE == OFFK_LOAD_VEC_ROW) {
      const int rq = (warp & 3) * 8 + (lane >> 2), kq = (warp >> 2) * 4 + (lane & 3);
      const int tr = row0 + 4 * rq;
      const int m = row_base + tr;
      const bool intile = tr < nrows;
      const bool v = intile && m < op.rows();
      const bool one = v && m == op.ones_row();
      r[0] = (v && !one) ? op.row(m) : offk_idx_t{0, 0, 0};
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[j] = swz(tr + j, kq);
      flags = (v ? 1u : 0u) | ((one ? 1u : 0u) << 8) | ((intile ? 1u : 0u) << 16);
    } else if (MODE == OFFK_LOAD_SCALAR_ROW) {
      const int tr = row0 + (tid & 127);
      const int m = row_base + tr;
      const bool intile = tr < nrows;
      const bool v = intile && m < op.rows();
      const bool one = v && m == op.ones_row();
      r[0] = (v && !one) ? op.row(m) : offk_idx_t{0, 0, 0};
      flags = (v ? 1u : 0u) | ((one ? 1u : 0u) << 8) | ((intile ? 1u : 0u) << 16);
    }
  }
  // column entries this thread needs for the K-block starting at k0 (vector modes only; scalar modes look them
  // up inside load())
  __device__ __forceinline__ void fetch_cols(const OP& op, int k0, int K, int tid, Cols& cc) const {
    if (MODE == OFFK_LOAD_VEC_K) {
      const int k = k0 + 4 * (tid & 7);
      cc.c[0] = k < K ? op.col(k) : offk_idx_t{0, 0, 0};
      cc.valid = k < K ? 1u : 0u;
    } else if (MODE == OFFK_LOAD_VEC_ROW) {
      const int kq = ((tid >> 5) >> 2) * 4 + (tid & 3);
      cc.valid = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + 4 * kq + e;
        cc.c[e] = k < K ? op.col(k) : offk_idx_t{0, 0, 0};
        cc.valid |= (k < K ? 1u : 0u) << e;
      }
    }
  }
  __device__ __forceinline__ float scalar(const OP& op, offk_idx_t rr, offk_idx_t cc, bool one) const {
    if (one) return 1.f;
    float x = 0.f;
    if (op.ok(rr, cc)) {
      x = __ldg(op.src() + (rr.off + cc.off));
      if (op.relu()) x = fmaxf(x, 0.f);
    }
    return x;
  }
  __device__ __forceinline__ void load(const OP& op, const Cols& cc, int row_base, int row0, int nrows, int k0, int K,
                                       int tid, float4 (&v)[4]) const {
    const int warp = tid >> 5, lane = tid & 31;
    if (MODE == OFFK_LOAD_VEC_K) {
      const bool kv = cc.valid != 0;
      const offk_idx_t c = cc.c[0];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 x = f4zero();
        if (kv && ((flags >> i) & 1u)) {
          if ((flags >> (8 + i)) & 1u) x = make_float4(1.f, 1.f, 1.f, 1.f);
          else if (op.ok(r[i], c)) {
            x = ldg128(op.src() + (r[i].off + c.off));
            if (op.relu()) x = f4relu(x);
          }
        }
        v[i] = x;
      }
    } else if (MODE == OFFK_LOAD_VEC_ROW) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float4 x = f4zero();
        const offk_idx_t c = cc.c[e];
        if (((cc.valid >> e) & 1u) && (flags & 1u)) {
          if ((flags >> 8) & 1u) x = make_float4(1.f, 0.f, 0.f, 0.f);  // ones row heads its own row-quad
          else if (op.ok(r[0], c)) {
            x = ldg128(op.src() + (r[0].off + c.off));
            if (op.relu()) x = f4relu(x);
          }
        }
        v[e] = x;
      }
    } else if (MODE == OFFK_LOAD_SCALAR_ROW) {
      const int half = tid >> 7;
      float t[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = k0 + half * 16 + i;
        t[i] = ((flags & 1u) && k < K) ? scalar(op, r[0], op.col(k), (flags >> 8) & 1u) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = make_float4(t[4 * i], t[4 * i + 1], t[4 * i + 2], t[4 * i + 3]);
    } else {  // OFFK_LOAD_SCALAR_K: lane owns column k0+lane, warp w owns tile rows row0 + w + 8*i
      const int k = k0 + lane;
      const bool kv = k < K;
      const offk_idx_t c = kv ? op.col(k) : offk_idx_t{0, 0, 0};
      float t[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int tr = row0 + warp + 8 * i;
        const int m = row_base + tr;
        float x = 0.f;
        if (kv && tr < nrows && m < op.rows()) {
          const bool one = (m == op.ones_row());
          x = scalar(op, one ? offk_idx_t{0, 0, 0} : op.row(m), c, one);
        }
        t[i] = x;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = make_float4(t[4 * i], t[4 * i + 1], t[4 * i + 2], t[4 * i + 3]);
    }
  }
  // OFFK_LOAD_VEC_K without ReLU-on-load: no register staging at all.  One cp.async per (row, 16-byte chunk);
  // invalid rows / out-of-box taps / k >= K are zero-filled by the copy engine.  Returns true when a plain
  // st.shared was used (the all-ones row), which then needs the generic->async proxy fence.
  template <bool L1_ALLOCATE>
  __device__ __forceinline__ bool issue_async(const OP& op, const Cols& cc, uint32_t tile_base) const {
    const bool kv = cc.valid != 0;
    const offk_idx_t c = cc.c[0];
    bool stored = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (!((flags >> (16 + i)) & 1u)) continue;
      const uint32_t d = tile_base + dst[i];
      if ((flags >> (8 + i)) & 1u) {
        const float one = kv ? 1.f : 0.f;
        sts128(d, one, one, one, one);
        stored = true;
        continue;
      }
      const bool ok = kv && ((flags >> i) & 1u) && op.ok(r[i], c);
      const float* src = ok ? op.src() + (r[i].off + c.off) : op.src();
      if (L1_ALLOCATE) cp_async16_ca(d, src, ok ? 16u : 0u);
      else             cp_async16_cg(d, src, ok ? 16u : 0u);
    }
    return stored;
  }
  // tile_base: smem address of tile row 0
  __device__ __forceinline__ void store(uint32_t tile_base, int row0, int nrows, int tid, const float4 (&v)[4]) const {
    const int warp = tid >> 5, lane = tid & 31;
    if (MODE == OFFK_LOAD_VEC_K) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if ((flags >> (16 + i)) & 1u) sts128(tile_base + dst[i], v[i].x, v[i].y, v[i].z, v[i].w);
    } else if (MODE == OFFK_LOAD_VEC_ROW) {
      if ((flags >> 16) & 1u) {
        sts128(tile_base + dst[0], v[0].x, v[1].x, v[2].x, v[3].x);
        sts128(tile_base + dst[1], v[0].y, v[1].y, v[2].y, v[3].y);
        sts128(tile_base + dst[2], v[0].z, v[1].z, v[2].z, v[3].z);
        sts128(tile_base + dst[3], v[0].w, v[1].w, v[2].w, v[3].w);
      }
    } else if (MODE == OFFK_LOAD_SCALAR_ROW) {
      const int tr = row0 + (tid & 127), half = tid >> 7;
      if (tr < nrows) {
#pragma unroll
        for (int c = 0; c < 4; ++c) sts128(tile_base + swz(tr, half * 4 + c), v[c].x, v[c].y, v[c].z, v[c].w);
      }
    } else {
      const float t[16] = {v[0].x, v[0].y, v[0].z, v[0].w, v[1].x, v[1].y, v[1].z, v[1].w,
                           v[2].x, v[2].y, v[2].z, v[2].w, v[3].x, v[3].y, v[3].z, v[3].w};
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int tr = row0 + warp + 8 * i;
        if (tr < nrows) sts32(tile_base + swz(tr, lane >> 2) + (lane & 3) * 4, t[i]);
      }
    }
  }
};

// An operand's whole producer pipeline (one or two 128-row passes).  Async (cp.async) operands prefetch their
// column entries one K-block ahead; register-staged operands prefetch column entries two and data one ahead.
template <int MODE, class OP, bool L1_ALLOCATE>
struct OperandPipe {
  TileLoader<MODE, OP> p0, p1;
  Cols c_cur, c_nxt;        // async: cols of block i / i+1;  register: cols of block i+1 / i+2
  float4 d_cur[4], d_nxt[4];
  int row_base, nrows;
  bool two;
  static constexpr bool async = (MODE == OFFK_LOAD_VEC_K);   // ReLU-on-load operands are never launched as VEC_K
  __device__ __forceinline__ void init(const OP& op, int row_base_, int nrows_, int tid) {
    row_base = row_base_;
    nrows = nrows_;
    two = nrows_ > 128;
    p0.init(op, row_base, 0, nrows, tid);
    if (two) p1.init(op, row_base, 128, nrows, tid);
  }
  __device__ __forceinline__ void prologue(const OP& op, int k_first, int K, int tid) {
    p0.fetch_cols(op, k_first, K, tid, c_cur);
    if (!async) {
      p0.load(op, c_cur, row_base, 0, nrows, k_first, K, tid, d_cur);
      p0.fetch_cols(op, k_first + TC_BK, K, tid, c_cur);
    }
  }
  // before blocking on the smem slot of K-block k0: issue the long-latency prefetches
  __device__ __forceinline__ void pre_wait(const OP& op, int k0, int K, int tid, bool has_next) {
    if (async) {
      p0.fetch_cols(op, k0 + TC_BK, K, tid, c_nxt);
    } else {
      if (has_next) p0.load(op, c_cur, row_base, 0, nrows, k0 + TC_BK, K, tid, d_nxt);
      p0.fetch_cols(op, k0 + 2 * TC_BK, K, tid, c_nxt);
    }
  }
  // after the slot is free: move K-block k0 into shared memory.  Returns true if st.shared was used.
  __device__ __forceinline__ bool post_wait(const OP& op, uint32_t tile_base, int k0, int K, int tid) {
    bool stored = false;
    if (async) {
      stored |= p0.template issue_async<L1_ALLOCATE>(op, c_cur, tile_base);
      if (two) stored |= p1.template issue_async<L1_ALLOCATE>(op, c_cur, tile_base);
    } else {
      p0.store(tile_base, 0, nrows, tid, d_cur);
      if (two) {   // second pass of a wide B tile: not prefetched (only dgrad / wgrad tiles wider than 128)
        Cols ct;
        float4 vt[4];
        p1.fetch_cols(op, k0, K, tid, ct);
        p1.load(op, ct, row_base, 128, nrows, k0, K, tid, vt);
        p1.store(tile_base, 128, nrows, tid, vt);
      }
      stored = true;
#pragma unroll
      for (int q = 0; q < 4; ++q) d_cur[q] = d_nxt[q];
    }
    c_cur = c_nxt;
    return stored;
  }
};

// ---------------------------------------------------------------------------- epilogue
// 4 consecutive output columns of one accumulator row, contiguous in memory (NHWC): float4 everywhere.
__device__ __forceinline__ void epi_store4(const offk_gemm_t& g, const EpiRow& r, int n, float4 v, bool atomic) {
  const int oc = g.out_col[n];
  float* o = g.out + (r.out + oc);
  if (atomic) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
    return;
  }
  if (g.bias) {
    const float4 b = ldg128(g.bias + n);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  if (n < g.relu_pre_cols) v = f4relu(v);   // relu_pre_cols is a multiple of 4 on this path
  const bool gated = g.gate && n >= g.gate_col0;
  float4 gt = make_float4(1.f, 1.f, 1.f, 1.f);
  if (gated) gt = ldg128(g.gate + (r.gate + (g.gate_col ? g.gate_col[n] : oc)));
  if (gated && g.gate_first) {
    v.x = gt.x > 0.f ? v.x : 0.f; v.y = gt.y > 0.f ? v.y : 0.f; v.z = gt.z > 0.f ? v.z : 0.f; v.w = gt.w > 0.f ? v.w : 0.f;
  }
  if (g.addend) {
    const float4 a = ldg128(g.addend + (r.add + (g.add_col ? g.add_col[n] : oc)));
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
  }
  if (gated && !g.gate_first) {
    v.x = gt.x > 0.f ? v.x : 0.f; v.y = gt.y > 0.f ? v.y : 0.f; v.z = gt.z > 0.f ? v.z : 0.f; v.w = gt.w > 0.f ? v.w : 0.f;
  }
  if (g.relu_post) v = f4relu(v);
  *reinterpret_cast<float4*>(o) = v;
}

// ---------------------------------------------------------------------------- kernel
template <int A_MODE, int B_MODE>
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
    // VEC_K operands go global->smem with cp.async (up to `stages` K-blocks in flight, no registers); the other
    // modes are staged through registers with the NEXT K-block's loads issued before this K-block is stored.
    const OpA opa(g);
    const OpB opb(g);
    OperandPipe<A_MODE, OpA, true> pa;
    OperandPipe<B_MODE, OpB, false> pb;
    pa.init(opa, m0, TC_BM, tid);
    pb.init(opb, n0, bn, tid);
    if (nkb > 0) {
      pa.prologue(opa, kb_begin * TC_BK, g.K, tid);
      pb.prologue(opb, kb_begin * TC_BK, g.K, tid);
    }
    int s = 0;
    uint32_t parity = 1;                                       // empty-barrier parity of the current round
    for (int i = 0; i < nkb; ++i) {
      const int k0 = (kb_begin + i) * TC_BK;
      const uint32_t a_base = smem_base + s * stage_bytes;
      const uint32_t b_base = a_base + TC_A_BYTES;
      pa.pre_wait(opa, k0, g.K, tid, i + 1 < nkb);
      pb.pre_wait(opb, k0, g.K, tid, i + 1 < nkb);
      mbar_wait(smem_u32(&sh->empty[s]), parity);              // slot free (first round passes at once)
      bool stored = pa.post_wait(opa, a_base, k0, g.K, tid);
      stored |= pb.post_wait(opb, b_base, k0, g.K, tid);
      if (stored) fence_proxy_async_smem();   // generic-proxy st.shared -> visible to the tensor core (async proxy)
      cp_async_mbar_arrive_noinc(smem_u32(&sh->full[s]));
      if (++s == stages) { s = 0; parity ^= 1u; }
    }
  } else {
    if (lane == 0) {
      // ================= MMA issuer (one thread) =================
      const uint32_t idesc = make_idesc_tf32(bn);
      int s = 0;
      uint32_t parity = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(smem_u32(&sh->full[s]), parity);
        fence_proxy_async_smem();
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
        if (++s == stages) { s = 0; parity ^= 1u; }
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
        if (g.out_vec && !er.ones) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const int n = n0 + c * 16 + j;
            if (n < g.N) epi_store4(g, er, n, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]), atomic);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + c * 16 + j;
            if (n < g.N) epi_store(g, er, n, v[j], atomic);
          }
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
  // one N tile when it fits a single UMMA (N <= 256, multiple of 16); otherwise the tile with least padding
  const int n16 = (N + 15) / 16 * 16;
  if (n16 <= 256) return n16;
  int best = 256, waste = 1 << 30;
  for (int bn = 256; bn >= 128; bn -= 16) {
    const int w = (N + bn - 1) / bn * bn - N;
    if (w < waste) { waste = w; best = bn; }
  }
  return best;
}

template <int A_MODE, int B_MODE>
static int launch_tc_t(const offk_gemm_t& g, int bn, int stages, int kb_per, int tmem_cols, dim3 grid, size_t smem,
                       cudaStream_t st) {
  auto kern = gather_gemm_tc_kernel<A_MODE, B_MODE>;
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
  const long long ctas = (long long)((g.M + TC_BM - 1) / TC_BM) * ((g.N + bn - 1) / bn) * ((num_kb + kb_per - 1) / kb_per);
  const int budget = ctas <= sm_count() ? 200 * 1024 : 108 * 1024;   // one resident CTA per SM -> deeper pipeline
  int stages = budget / (int)stage_bytes;
  if (stages < 2) stages = 2;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages > kb_per) stages = kb_per < 2 ? 2 : kb_per;
  const size_t smem = (size_t)stages * stage_bytes + sizeof(TcShared) + 1024;
  int tmem_cols = 32;
  while (tmem_cols < bn) tmem_cols <<= 1;
  dim3 grid((g.M + TC_BM - 1) / TC_BM, (g.N + bn - 1) / bn, (num_kb + kb_per - 1) / kb_per);
  if (grid.y > 65535 || grid.z > 65535) return fail(OFFK_E_LIMIT, "gather_gemm: grid too large");
  // cp.async cannot apply ReLU-on-load: such operands (one small 1x1 conv) take the scalar register path
  const int a_mode = (g.a_relu && g.a_mode == OFFK_LOAD_VEC_K) ? OFFK_LOAD_SCALAR_ROW : g.a_mode;
#define OFFK_TC_CASE(AM, BM) \
  if (a_mode == AM && g.b_mode == BM) return launch_tc_t<AM, BM>(g, bn, stages, kb_per, tmem_cols, grid, smem, st);
  OFFK_TC_CASE(0, 0) OFFK_TC_CASE(0, 1) OFFK_TC_CASE(0, 2) OFFK_TC_CASE(0, 3)
  OFFK_TC_CASE(1, 0) OFFK_TC_CASE(1, 1) OFFK_TC_CASE(1, 2) OFFK_TC_CASE(1, 3)
  OFFK_TC_CASE(2, 0) OFFK_TC_CASE(2, 1) OFFK_TC_CASE(2, 2) OFFK_TC_CASE(2, 3)
  OFFK_TC_CASE(3, 0) OFFK_TC_CASE(3, 1) OFFK_TC_CASE(3, 2) OFFK_TC_CASE(3, 3)
#undef OFFK_TC_CASE
  return fail(OFFK_E_BADARG, "gather_gemm: unsupported load modes %d/%d", a_mode, g.b_mode);
}

}  // namespace offk
