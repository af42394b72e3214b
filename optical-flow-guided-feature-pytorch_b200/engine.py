"""OFFEngine: the OFF sub-network (RGB_OFF.py:596-860 / Flow_OFF.py:606-884) as a static plan of
liboffk kernel launches over preallocated fp32 buffers.

Layout: the taps arrive NCHW (the reference's layout); the unit's fused 1x1 GEMM reads them in place through a TMA
tensor map (MN-major operand), converts to channels-last in its epilogue, and everything downstream (reduced
features, stage-fusion buffers, residual blocks, gradients) is NHWC, so that every operand tile is a TMA box or a
16-byte cp.async gather.

Every step is one C-ABI call with a descriptor bound at plan-build time; a forward or backward pass is a flat list of
launches spread over up to three streams by a hazard-derived schedule (class Schedule).  PyTorch supplies device
memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _lib as L
from . import spec as S
from . import tables as T

_SM_TARGET = 148
_PICK_A = float(os.environ.get("OFFK_PICK_A", "128"))      # _pick_tile_n: per-CTA cost ~ _PICK_A + bn (A rows + B rows of a K-block)
_X3_MAX_BN = int(os.environ.get("OFFK_X3_MAX_BN", "256"))     # widest N tile of a TMA-fed GEMM in the 3xTF32 mode
_TC_PRECS = (L.PREC_TF32, L.PREC_TF32X3)      # tcgen05 modes: 1 MMA per product / error-compensated 3xTF32


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _FreeGeom:
    """im2col geometry with a caller-given output grid and separate top / left padding (OFFK_TGEMM_FREE_GEOM)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def _pick_tile_n(M: int, N: int, slots: int = 2 * _SM_TARGET) -> int:
    """N tile (multiple of 16, <= 256; the last tile may be partial) of a GEMM without split-K: the one that minimises
    rounds x per-CTA cost, rounds = ceil(CTAs / resident CTA slots) and cost ~ 128 + bn (rows of A and B a K-block moves
    through shared memory).  Wave quantisation matters: 37 x 11 tiles of 96 columns are 1.4 rounds, 37 x 8 of 144 are one."""
    mt = math.ceil(M / 128)
    best, best_cost = 0, None
    for bn in range(256, 15, -16):
        if bn > (N + 15) // 16 * 16:
            continue
        nt = math.ceil(N / bn)
        cost = math.ceil(mt * nt / slots) * (_PICK_A + bn) * (1.0 if mt * nt >= _SM_TARGET else _SM_TARGET / (mt * nt)) ** 0.5
        if best_cost is None or cost < best_cost - 1e-9:
            best, best_cost = bn, cost
    return best


def _wgrad_split(tiles: int, kb: int, slots: int, bk: int = 32) -> int:
    """Split-K factor of a weight-gradient GEMM with `tiles` output tiles and `kb` K-blocks of `bk` pixels: the one that
    minimises rounds x (K-blocks per CTA + fixed cost), rounds = ceil(CTAs / resident CTA slots); an under-filled round is
    charged as idle SMs.  The fixed cost (prologue, first TMA round trip, adding the tile into dW) is worth ~10 K-blocks of
    32 pixels (profiles/timeline_*_r02z.txt); the result is insensitive to it between 5 and 20 (profiles/env_r03j.log)."""
    fixed = float(os.environ.get("OFFK_WGRAD_FIXED_KB", "10")) * 32 / bk
    cost = lambda s_: math.ceil(tiles * s_ / slots) * (math.ceil(kb / s_) + fixed) * max(1.0, _SM_TARGET / (tiles * s_))
    return min(range(1, max(1, kb // 4) + 1), key=lambda s_: (cost(s_), s_))


class Gemm:
    """One bound gather-GEMM launch (descriptor + device tables kept alive)."""

    def __init__(self, eng: "OFFEngine", spc: T.GemmSpec, key, *, a_src, b_src, out, bias=None, relu_pre_cols=0,
                 a_relu=False, gate=None, gate_tabs=None, gate_col0=0, gate_first=False, addend=None,
                 add_tabs=None, relu_post=False, atomic=False, ones_out=None, split_k=1, tile_n=0, name="",
                 finish=False, aux=None, out_vec=None):
        self.eng, self.name, self.spec = eng, name, spc
        tabs = eng._tables(key, spc)
        d = L.OffkGemm()
        d.M, d.N, d.K = spc.M, spc.N, spc.K
        d.a_src, d.a_row, d.a_col = a_src.data_ptr(), tabs["a_row"].data_ptr(), tabs["a_col"].data_ptr()
        d.a_h, d.a_w = (spc.a_h, spc.a_w) if spc.a_h else (T.NO_BOX, T.NO_BOX)
        d.a_relu, d.a_ones_row, d.a_mode = int(a_relu), spc.a_ones_row, spc.a_mode
        d.b_src, d.b_row, d.b_col, d.b_mode = b_src.data_ptr(), tabs["b_row"].data_ptr(), tabs["b_col"].data_ptr(), spc.b_mode
        d.out, d.out_row, d.out_col = out.data_ptr(), tabs["out_row"].data_ptr(), tabs["out_col"].data_ptr()
        d.bias = bias.data_ptr() if bias is not None else None
        d.relu_pre_cols = relu_pre_cols
        d.gate = gate.data_ptr() if gate is not None else None
        d.gate_row = gate_tabs[0].data_ptr() if gate_tabs else None
        d.gate_col = gate_tabs[1].data_ptr() if gate_tabs else None
        d.gate_col0, d.gate_first = gate_col0, int(gate_first)
        d.addend = addend.data_ptr() if addend is not None else None
        d.add_row = add_tabs[0].data_ptr() if add_tabs else None
        d.add_col = add_tabs[1].data_ptr() if add_tabs else None
        d.relu_post, d.atomic_out = int(relu_post), int(atomic)
        d.ones_row_out = ones_out.data_ptr() if ones_out is not None else None
        if tile_n == 0 and split_k == 1:
            tile_n = _auto_tile_n(spc.M, spc.N, isinstance(self, TGemm), eng.prec == L.PREC_TF32X3)
        if isinstance(self, TGemm) and eng.prec == L.PREC_TF32X3 and _X3_MAX_BN < 256:
            # 3xTF32 keeps the A operand in tensor memory when the accumulators leave room (N tile <= 192, offk_gemm_tma.cu);
            # a wide tile would fall back to both operands in shared memory and two pipeline stages
            bn = tile_n if tile_n else (256 if spc.N > 256 else (spc.N + 15) // 16 * 16)
            if bn > 192:
                tile_n = min(_X3_MAX_BN, (math.ceil(spc.N / math.ceil(spc.N / _X3_MAX_BN)) + 15) // 16 * 16)
        d.split_k, d.tile_n = split_k, tile_n
        d.out_vec = spc.out_vec if out_vec is None else out_vec
        counter = None
        if finish:                         # split-K finished in-kernel by the last CTA of each output tile (offk.h)
            bn = tile_n if tile_n else (256 if spc.N > 256 else (spc.N + 15) // 16 * 16)
            counter = torch.zeros(math.ceil(spc.M / 128) * math.ceil(spc.N / bn), dtype=torch.int32, device=eng.device)
            d.finish_counter = counter.data_ptr()
        if aux is not None:                # (aux_out, aux_row table or None, aux_col0, aux_addend or None)
            d.aux_out = aux[0].data_ptr()
            d.aux_row = aux[1].data_ptr() if aux[1] is not None else None
            d.aux_col0 = aux[2]
            d.aux_addend = aux[3].data_ptr() if aux[3] is not None else None
        T.check_modes(spc)
        self.desc = d
        self._keep = (tabs, a_src, b_src, out, bias, gate, gate_tabs, addend, add_tabs, ones_out, counter, aux)
        self.flops = 2.0 * spc.M * spc.N * spc.K
        self.launches = [name]
        self.reads = [a_src, b_src, addend, gate, bias, aux[3] if aux else None]
        self.writes = [out, ones_out, aux[0] if aux else None]
        self.lane = 0

    def set_a_src(self, ptr):
        self.desc.a_src = ptr

    def __call__(self, stream):
        L.check(self.eng.lib.offk_gather_gemm(C.byref(self.desc), self.eng.prec, stream), self.name)


class TGemm(Gemm):
    """A forward conv / FC on a channels-last input as ONE TMA-fed launch (offk_tma_gemm): A through a dense 2-D
    tensor map (1x1, stride 1) or TMA im2col mode (KxK), B = the [cout, K] weight matrix.  Same epilogue fields and
    output tables as the gather-GEMM it stands in for."""

    @staticmethod
    def eligible(geom: T.ConvGeom, prec, a_relu=False, x_layout="nhwc") -> bool:
        if x_layout == "nchw":      # the taps, read in place as an MN-major operand: 1x1 conv over whole frames
            return (prec in _TC_PRECS and not a_relu and T._is1x1(geom) and geom.cin % 32 == 0 and geom.x_coff == 0
                    and geom.x_ctot == geom.cin and (geom.hin * geom.win) % 4 == 0)
        return (prec in _TC_PRECS and not a_relu and geom.cin % 32 == 0 and geom.x_ctot % 4 == 0 and geom.x_coff % 4 == 0
                and geom.kdim % 4 == 0)

    @staticmethod
    def wgrad_eligible(geom: T.ConvGeom, prec, a_relu=False, x_layout="nhwc") -> bool:
        """Weight gradient with both operands TMA-fed: A = im2col(x)^T (channels-last) or the NCHW tap, B = dY."""
        y_ok = geom.y_ctot % 4 == 0 and geom.y_coff % 4 == 0 and geom.cout % 4 == 0
        if x_layout == "nchw":
            return (prec in _TC_PRECS and not a_relu and y_ok and T._is1x1(geom) and geom.x_coff == 0
                    and geom.x_ctot == geom.cin and (geom.hin * geom.win) % 4 == 0)
        return (prec in _TC_PRECS and not a_relu and y_ok and geom.cin % 32 == 0 and geom.x_ctot % 4 == 0
                and geom.x_coff % 4 == 0)

    def __init__(self, eng, spc, key, geom: T.ConvGeom, x_layout="nhwc", wgrad=False, bk=0, **kw):
        super().__init__(eng, spc, key, **kw)
        t = L.OffkTGemm()
        C.memmove(C.byref(t.g), C.byref(self.desc), C.sizeof(L.OffkGemm))
        one = geom.kh == 1 and geom.kw == 1 and geom.stride == 1 and geom.pad == 0
        if wgrad:
            t.a_kind = L.TMA_A_NCHW_T if x_layout == "nchw" else L.TMA_A_IM2COL_T
        elif x_layout == "nchw":
            t.a_kind = L.TMA_A_NCHW
        elif one:
            t.a_kind, t.lda = L.TMA_A_DENSE, geom.x_ctot
            t.g.a_src = kw["a_src"].data_ptr() + 4 * geom.x_coff
        else:
            t.a_kind = L.TMA_A_IM2COL
        t.a_coff = geom.x_coff
        if not wgrad and spc.out_vec == 1 and spc.M > 1:
            # linear output rows: a plain epilogue may leave through TMA tile stores (offk.h: out_ld)
            r = np.asarray(spc.out_row, dtype=np.int64)
            ld = int(r[1] - r[0])
            if ld > 0 and r[0] == 0 and np.array_equal(r, np.arange(spc.M, dtype=np.int64) * ld):
                t.out_ld, t.out_c0 = ld, int(spc.out_col[0])
        t.precision = eng.prec
        t.bk = bk if (wgrad and x_layout != "nchw") else 0
        if eng.presplit and not wgrad:
            # B is a weight matrix inside the parameter buffer or the per-step weight copies: its residuals sit at the same
            # offset of the twin buffer
            bp = kw["b_src"].data_ptr()
            for hi, lo in ((eng.params_flat, eng.params_lo), (eng.wc_flat, eng.wc_lo)):
                if hi.data_ptr() <= bp < hi.data_ptr() + hi.numel() * 4:
                    t.b_lo_delta = (lo.data_ptr() - hi.data_ptr()) // 4
                    self.reads.append(lo)
        t.n_img, t.hin, t.win, t.ctot, t.cin = geom.n_img, geom.hin, geom.win, geom.x_ctot, geom.cin
        t.kh, t.kw, t.stride, t.pad, t.hout, t.wout = geom.kh, geom.kw, geom.stride, geom.pad, geom.hout, geom.wout
        if isinstance(geom, _FreeGeom):
            t.geom_flags, t.pad_w = L.TGEMM_FREE_GEOM, geom.pad_w
        if wgrad:                         # B = dY[pixel, cout slice], row-major
            t.b_kind, t.ldb = L.TMA_B_DENSE_T, geom.y_ctot
            t.g.b_src = kw["b_src"].data_ptr() + 4 * geom.y_coff
        else:
            t.b_kind, t.ldb = L.TMA_B_DENSE, geom.kdim
        self.tdesc = t
        self._by_src = {}                 # a_src pointer -> prepared descriptor (the two tap sets alternate)

    def set_a_src(self, ptr):
        if self.tdesc.g.a_src == ptr:
            return
        self._by_src[self.tdesc.g.a_src] = self.tdesc
        t = self._by_src.get(ptr)
        if t is None:
            t = L.OffkTGemm()
            C.memmove(C.byref(t), C.byref(self.tdesc), C.sizeof(L.OffkTGemm))
            t.g.a_src, t.prepared = ptr, 0
        self.tdesc = t
        self.desc.a_src = ptr

    def __call__(self, stream):
        lib = self.eng.lib
        if not self.tdesc.prepared:                                  # tensor maps are encoded on first use (needs a driver)
            L.check(lib.offk_tma_gemm_prepare(C.byref(self.tdesc)), self.name + " (prepare)")
        L.check(lib.offk_tma_gemm(C.byref(self.tdesc), stream), self.name)


class OFFEngine:
    """Plan + buffers for one (batch, length) shape on one GPU.

    variant: 'rgb'  learned depth-wise 3x3 spatial gradient + bias, per-pair logits (RGB_OFF.py)
             'flow' fixed diagonal Sobel, segment consensus over the L-1 pairs (Flow_OFF.py / RGB_OFF_v2.py)
    precision: 'fp32' = fp32-parity mode ON THE TENSOR CORES (OFFK_PREC_TF32X3: error-compensated 3xTF32 tcgen05 MMAs,
               fp32 accumulate; the reference's arithmetic is fp32, RGB_OFF.py:597), 'tf32' = one kind::tf32 MMA per
               product (the fast mode, tolerance stated separately), 'fp32_simt' = CUDA-core FFMA cross-check (tests only)
    """

    def __init__(self, batch: int, length: int, variant: str = "rgb", device="cuda", precision: str = "tf32",
                 index_mode: str = "reference_flat", consensus=None, tap_grads: bool = False):
        assert variant in ("rgb", "flow") and length >= 2 and batch >= 1
        self.lib = L.lib()
        self.B, self.Lseg, self.variant = batch, length, variant
        self.N, self.P = batch * length, batch * (length - 1)
        self.device = torch.device(device)
        self.prec = {"fp32": L.PREC_TF32X3, "tf32": L.PREC_TF32, "fp32_simt": L.PREC_FP32}[precision]
        self.tc = self.prec in _TC_PRECS
        self.precision = precision
        self.index_mode = {"reference_flat": L.INDEX_REFERENCE_FLAT, "aligned": L.INDEX_ALIGNED}[index_mode]
        self.consensus = (variant != "rgb") if consensus is None else bool(consensus)
        self.tap_grads = tap_grads
        self.single_stream = False       # True: issue every lane on the caller's stream (profiling / debugging)
        self.use_tma = os.environ.get("OFFK_NO_TMA", "0") != "1"     # bring-up switch: gather-fed GEMMs everywhere
        # weight gradients with TMA-fed operands: on for the NCHW taps (measured 81 -> 54 us at level 3a); off by default
        # for channels-last convs, where 32-pixel im2col boxes lose to the cp.async gather on the big KxK layers
        # (motion_conv_trans_28: 388 vs 211 us) and tie elsewhere -- OFFK_TMA_WGRAD=all forces them on, =none disables both
        self.tma_wgrad = os.environ.get("OFFK_TMA_WGRAD", "taps")
        self.tma_strided_dgrad = os.environ.get("OFFK_NO_TMA_SDGRAD", "0") != "1"
        # channels-last weight gradients through TMA im2col with DEEP K-blocks (bk output pixels per stage: 64 / 128 instead
        # of 32 -> fewer, larger TMA boxes).  OFFK_WGRAD_TMA: which layers ("big" = the three stage-entry KxK convs, "kxk" =
        # every KxK conv, "all"); OFFK_WGRAD_BK: the depth
        # Measured at config 2 (profiles/wgrad_sweep_r02x.txt).  tf32 mode: the cp.async gather kernel is faster everywhere
        # except on motion_conv_trans_28, where 128-pixel boxes win (197 us gather, 389 / 204 / 169 us TMA at bk = 32 / 64 / 128;
        # motion_conv_trans_14: 97 vs 136, motion_conv_trans: 60 vs 77).  fp32 (3xTF32) mode, where the gather kernel's producers
        # also write the residual tiles: the TMA-fed kernel wins on every KxK layer (784 -> 501, 344 -> 296, 194 -> 177 us at the
        # deepest K-block that still leaves two pipeline stages).  Defaults follow the measurement.
        x3 = self.prec == L.PREC_TF32X3
        self.wgrad_tma = os.environ.get("OFFK_WGRAD_TMA", "all" if x3 else "t28")
        self.wgrad_bk = int(os.environ.get("OFFK_WGRAD_BK", "64" if x3 else "128"))
        self._tab_cache = {}
        self._keep = []
        self.generation = 0

        # ---- parameters: one flat buffer, reference-named views
        self.layout, self.n_flat = S.flat_layout(variant)
        # fp32 (3xTF32) mode: the tf32 residuals of the weights are produced once per step into a twin buffer directly behind
        # the parameters, so that the GEMMs receive both weight tiles by TMA (offk.h: b_lo_delta) instead of splitting them in
        # shared memory in every CTA and K-block
        self.presplit = self.prec == L.PREC_TF32X3 and os.environ.get("OFFK_NO_PRESPLIT", "0") != "1"
        self._params_all = torch.zeros(self.n_flat * (2 if self.presplit else 1), device=self.device)
        self.params_flat = self._params_all[:self.n_flat]
        self.params_lo = self._params_all[self.n_flat:] if self.presplit else None
        self.grads_flat = torch.zeros(self.n_flat, device=self.device)
        self.params = OrderedDict((n, self._view(self.params_flat, n)) for n in S.param_shapes(variant))
        self.grads = OrderedDict((n, self._view(self.grads_flat, n)) for n in S.param_shapes(variant))
        if variant == "flow":
            # frozen util.SobelFilter_Diagonal taps (util.py:61), [32,1,3,3]
            k = torch.tensor([[0., 1., 0.], [-1., 0., 1.], [0., -1., 0.]], device=self.device)
            self.sobel_w = k.expand(S.DOWN_C, 1, 3, 3).contiguous()

        self._alloc()
        self._build()

    # ------------------------------------------------------------------ helpers
    def _view(self, flat, name):
        off, shape = self.layout[name]
        return flat[off:off + int(np.prod(shape))].view(shape)

    def _unit_w(self, flat, tag):
        off, _ = self.layout[f"motion_conv_gen_{tag}.weight"]
        cin = S.LEVELS[tag][0]
        return flat[off:off + S.UNIT_C * cin].view(S.UNIT_C, cin)

    def _unit_b(self, flat, tag):
        off, _ = self.layout[f"motion_conv_gen_{tag}.bias"]
        return flat[off:off + S.UNIT_C]

    def _tables(self, key, spc: T.GemmSpec):
        if key not in self._tab_cache:
            dev = self.device
            self._tab_cache[key] = {k: torch.from_numpy(v).to(dev) for k, v in T.padded_tables(spc).items()}
        return self._tab_cache[key]

    def _buf(self, name, *shape):
        t = torch.zeros(*shape, device=self.device)
        self.buf[name] = t
        return t

    def load_params(self, prm: dict):
        """Copy reference-named tensors (a state_dict subset) into the flat buffer."""
        with torch.no_grad():
            for n, v in self.params.items():
                v.copy_(prm[n].to(self.device, torch.float32))

    # ------------------------------------------------------------------ buffers
    def _alloc(self):
        P, N = self.P, self.N
        self.buf = {}
        for tag, (cin, s) in S.LEVELS.items():
            self._buf("gd_" + tag, N, s, s, S.UNIT_C)
            self._buf("dgd_" + tag, N, s, s, S.UNIT_C)
        for st, (ctot, s, _) in S.STAGES.items():
            self._buf("F" + st, P, s, s, ctot)
            self._buf("dF" + st, P, s, s, ctot)
        a = lambda n, c, s: (self._buf(n, P, s, s, c), self._buf("d_" + n, P, s, s, c))
        a("t28", 64, 14)
        self._buf("t28r", P, 14, 14, 64)          # relu(t28): conv1_trans_28a consumes the activated copy (RGB_OFF.py:658-659)
        a("tmp28", 64, 14)
        for blk in "abc":
            a("h1_28" + blk, 64, 14)
            a("h2_28" + blk, 64, 14)
        a("br28", 256, 14)
        a("s28a", 256, 14)
        a("s28b", 256, 14)
        a("t14", 128, 7)
        a("tmp14", 128, 7)
        for blk in "ab":
            a("h1_14" + blk, 128, 7)
            a("h2_14" + blk, 128, 7)
        a("ex14", 512, 7)
        a("s14a", 512, 7)
        a("h3_14b", 512, 7)
        a("t7", 256, 7)
        a("tmp7", 256, 7)
        a("h1_7", 256, 7)
        a("h2_7", 256, 7)
        a("br7", 1024, 7)
        a("s7", 1024, 7)
        self._buf("p28", P, 7, 7, 256)
        for k, c in (("7", 1024), ("14", 512), ("28", 256)):
            self._buf("pool" + k, P, c)
            self._buf("fc" + k, P, S.NUM_CLASSES)
            if self.consensus:
                self._buf("cfc" + k, self.B, S.NUM_CLASSES)
        # two input sets: while one is being consumed (forward AND backward read the taps) the other can be filled by
        # an asynchronous host->device copy (stage_taps)
        self.tap_sets = [OrderedDict((tag, torch.zeros(N, cin, s, s, device=self.device)) for tag, (cin, s) in S.LEVELS.items())
                         for _ in range(2)]
        self.taps = self.tap_sets[0]
        self._tap_set = 0
        self._tap_users = {tag: [] for tag in S.LEVELS}      # bound GEMMs whose A operand is the tap
        self._copy_stream = None
        if self.tap_grads:
            self.tap_grad = OrderedDict((tag, torch.zeros_like(t)) for tag, t in self.taps.items())
        # Per-step weight re-layouts, produced by ONE gather-copy launch (wc_flat[i] = params_flat[wc_idx[i]]):
        #   wp[name] : OHWI [cout, kh, kw, cin] copies of the KxK conv weights (K order of the channels-last implicit GEMM)
        #   wd[name] : [cin, kh, kw, cout] copies with flipped taps of every stride-1 conv = the B operand of its
        #              data-gradient GEMM (dX = dY * W^T as a forward conv over dY)
        # dwp[name]: OHWI accumulators of the KxK weight gradients (un-permuted into grads_flat after the backward)
        self.wp, self.dwp, self.wd = {}, {}, {}
        kxk = [(name, cout, cin, k) for name, cout, cin, k, _, _ in S.STAGE_CONVS if k > 1]
        total = sum(cout * cin * k * k for _, cout, cin, k in kxk)
        self.dwp_flat = torch.zeros(total, device=self.device)
        idx, off, views = [], 0, []
        for name, cout, cin, k, stride, pad in S.STAGE_CONVS:
            p_off = self.layout[name + ".weight"][0]
            oihw = p_off + np.arange(cout * cin * k * k, dtype=np.int64).reshape(cout, cin, k, k)
            if k > 1:
                idx.append(oihw.transpose(0, 2, 3, 1).reshape(-1))
                views.append(("wp", name, off, (cout, k, k, cin)))
                off += oihw.size
            if stride == 1:
                idx.append(p_off + T.dgrad_class_weight_index(cout, cin, k, 1, pad, 0, 0))
                views.append(("wd", name, off, (cin, k * k * cout)))
                off += oihw.size
            else:
                # one [cin, R, Q, cout] block per stride-parity class (a, b) of the input pixel
                for a in range(stride):
                    for b in range(stride):
                        sub = p_off + T.dgrad_class_weight_index(cout, cin, k, stride, pad, a, b)
                        idx.append(sub)
                        views.append(("wd", (name, a, b), off, (cin, sub.size // cin)))
                        off += sub.size
        self.wc_idx = torch.from_numpy(np.concatenate(idx).astype(np.int32)).to(self.device)
        off = (off + 3) // 4 * 4
        self._wc_all = torch.zeros(off * (2 if self.presplit else 1), device=self.device)
        self.wc_flat = self._wc_all[:off]
        self.wc_lo = self._wc_all[off:] if self.presplit else None
        for kind, name, o, shape in views:
            getattr(self, kind)[name] = self.wc_flat[o:o + int(np.prod(shape))].view(shape)
        off = 0
        for name, cout, cin, k in kxk:
            n = cout * cin * k * k
            self.dwp[name] = self.dwp_flat[off:off + n].view(cout, k, k, cin)
            off += n
        # dropout state (filled per forward call)
        self.drop_mode = L.DROP_NONE
        self.drop_seed = 0
        self.masks = None
        self.seed_state = torch.zeros(1, dtype=torch.int64, device=self.device)   # per-step dropout seed (device word)
        self._graphs = {}                # (pass, tap set, dropout mode) -> torch.cuda.CUDAGraph
        self._cap_stream = None

    # ------------------------------------------------------------------ plan
    def _conv_fwd(self, name, x, y, geom, w, b, *, relu=False, relu_cols=None, a_relu=False, addend=None,
                  add_tabs=None, relu_post=False, x_layout="nhwc", no_split=False, aux=None):
        """aux = (aux_out, aux_row table | None, aux_col0, aux_addend | None): second output max(v + aux_addend, 0) written by
        the GEMM epilogue (offk.h); only honoured on the TMA-fed path -- the returned step has .aux_fused = True then."""
        spc = T.conv_fwd_spec(geom, x_layout, "nhwc")
        m_tiles, kb = math.ceil(spc.M / 128), math.ceil(spc.K / 32)
        n_tiles = max(1, math.ceil(spc.N / 256))
        if self.prec == L.PREC_TF32X3 and self.use_tma and spc.N > 192:
            n_tiles = math.ceil(spc.N / _X3_MAX_BN)      # Gemm.__init__ re-tiles wide 3xTF32 GEMMs
        split = 1
        if self.tc and addend is None and m_tiles * n_tiles < 200 and kb >= 32:
            # split-K factor: minimise rounds x K-blocks per CTA, rounds = ceil(CTAs / resident CTA slots) (two CTAs per SM in
            # the tf32 mode, one in the 3xTF32 mode whose stages are twice as large); ties go to the smaller factor.
            # 147 tiles: 2 splits = 294 CTAs = one round (3 splits were 1.49 rounds at 31 % tensor-pipe activity).
            slots = _SM_TARGET * (1 if self.prec == L.PREC_TF32X3 else 2)
            tiles = m_tiles * n_tiles
            split = min(range(1, max(1, kb // 8) + 1),
                        key=lambda s_: (math.ceil(tiles * s_ / slots) * math.ceil(kb / s_) * (1 if tiles * s_ >= 0.9 * _SM_TARGET else 4), s_))
            forced = dict(kv.split("=") for kv in os.environ.get("OFFK_FWD_SPLIT", "").split(",") if "=" in kv)
            split = int(forced.get(name, split))          # tuning hook: OFFK_FWD_SPLIT=motion_conv_trans_28=3,...
        cols = (geom.cout if relu else 0) if relu_cols is None else relu_cols
        tma = self.use_tma and TGemm.eligible(geom, self.prec, a_relu, x_layout)
        mk = (lambda *a, **k: TGemm(*a, geom=geom, x_layout=x_layout, **k)) if tma else Gemm
        if (tma and x_layout == "nchw") or no_split:
            split = 1                     # per-frame M tiles already fill the machine (N * ceil(hw/128) CTAs)
        fuse = tma and spc.out_vec and os.environ.get("OFFK_NO_FINISHER", "0") != "1"
        # in-kernel finishing pays when there are enough output tiles for the finishing CTAs to keep the machine busy
        # (motion_conv_trans_28: 147 tiles, same time as the bare split-K GEMM, bias_act and ReLU copy for free); with the
        # 37-tile 7x7 layers the fences + a finishing pass by 37 CTAs cost 9-27 us more than a separate full-grid bias_act
        # pass (measured, profiles/launches_tf32_r02g*.txt), so those keep the two-kernel form
        fin_ok = fuse and m_tiles * n_tiles >= int(os.environ.get("OFFK_FINISH_MIN_TILES", "100"))
        if split > 1 and fin_ok:
            # split-K partial tiles are added into the zeroed output; the last CTA of each tile applies bias / ReLU (and
            # writes the second output) in place: no separate bias_act pass
            g = mk(self, spc, ("fwd", x_layout, _gkey(geom)), a_src=x, b_src=w, out=y, bias=b, relu_pre_cols=cols, a_relu=a_relu,
                   split_k=split, name=name, finish=True, aux=aux)

            def run(stream, g=g, y=y):
                L.check(self.lib.offk_fill_zero(_ptr(y), y.numel(), stream), name + ".zero")
                g(stream)
            assert geom.y_coff == 0 and geom.y_ctot == geom.cout
            self.flops_fwd += g.flops
            run.name, run.flops, run.gemm, run.aux_fused = name, g.flops, g, aux is not None
            run.launches = [name + ".zero", name + ".splitk"]
            run.reads, run.writes, run.lane = g.reads, g.writes, 0
            return run
        if split > 1:
            g = mk(self, spc, ("fwd", x_layout, _gkey(geom)), a_src=x, b_src=w, out=y, a_relu=a_relu, split_k=split, name=name)
            hw = geom.hout * geom.wout

            def run(stream, g=g, y=y, b=b, geom=geom, hw=hw, cols=cols):
                L.check(self.lib.offk_fill_zero(_ptr(y), y.numel(), stream), name + ".zero")
                g(stream)
                L.check(self.lib.offk_bias_act(_ptr(y), _ptr(b), geom.n_img, geom.cout, hw, geom.y_ctot, geom.y_coff,
                                               cols, stream), name + ".bias_act")
            assert geom.y_coff == 0 and geom.y_ctot == geom.cout
            self.flops_fwd += g.flops
            run._split = True
            run.name = name
            run.flops = g.flops
            run.launches = [name + ".zero", name + ".splitk", name + ".bias_act"]
            run.gemm = g
            run.reads, run.writes, run.lane = g.reads, g.writes, 0
            return run
        # nchw TMA: one N tile (re-reading the tap per N tile would multiply the HBM traffic of an HBM-bound GEMM)
        tn = (spc.N + 15) // 16 * 16 if (tma and x_layout == "nchw" and spc.N <= 256) else 0
        g = mk(self, spc, ("fwd", x_layout, _gkey(geom)), a_src=x, b_src=w, out=y, bias=b, relu_pre_cols=cols, a_relu=a_relu,
               addend=addend, add_tabs=add_tabs, relu_post=relu_post, tile_n=tn, name=name, aux=aux if fuse else None)
        g.aux_fused = bool(fuse and aux is not None)
        self.flops_fwd += g.flops
        return g

    def _conv_wgrad(self, name, x, dy, geom, dw, db, *, a_relu=False, x_layout="nhwc", force_tma=False):
        # dW is accumulated in the k order of the implicit GEMM (OHWI: consecutive accumulator rows = consecutive
        # addresses, so the split-K reductions coalesce) and un-permuted once at the end; writing OIHW directly
        # (dw_layout="oihw") was measured slower: the strided atomics cost more than the permute pass
        spc = T.conv_wgrad_spec(geom, x_layout, "nhwc")
        tma = (self.use_tma and (self.tma_wgrad == "all" or (self.tma_wgrad == "taps" and (x_layout == "nchw" or force_tma)))
               and TGemm.wgrad_eligible(geom, self.prec, a_relu, x_layout))
        bk = 0
        base = name.split(".")[0]
        deep = {"t28": base == "motion_conv_trans_28",
                "big": base in ("motion_conv_trans_28", "motion_conv_trans_14", "motion_conv_trans"),
                "kxk": geom.kh > 1, "all": True}.get(self.wgrad_tma, False)
        if (deep and self.use_tma and x_layout == "nhwc" and not force_tma and self.wgrad_bk in (32, 64, 128)
                and TGemm.wgrad_eligible(geom, self.prec, a_relu, x_layout)):
            # deepest K-block <= the requested one whose pipeline stage (A: 4 stacks of {32 rows x bk pixels}, B: one per 32
            # columns of the N tile; twice that with the 3xTF32 residual tiles) still leaves room for two stages per SM
            bn = 256 if spc.N > 256 else (spc.N + 15) // 16 * 16
            mul = 2 if self.prec == L.PREC_TF32X3 else 1
            fits = [b for b in (128, 64, 32) if b <= self.wgrad_bk and mul * (b * 512 + math.ceil(bn / 32) * b * 128) <= 108 * 1024]
            if fits:
                tma, bk = True, fits[0]
        m_tiles, kb = math.ceil(spc.M / 128), math.ceil(spc.K / (bk or 32))
        n_tiles = max(1, math.ceil(spc.N / 256))
        # split-K factor: minimise rounds x (K-blocks per CTA + fixed cost), rounds = ceil(CTAs / resident CTA slots) -- two CTAs
        # per SM in the tf32 mode, one in the 3xTF32 mode.  The fixed cost (prologue, first TMA round trip, the epilogue that
        # adds the CTA's whole tile into dW) is worth ~10 K-blocks (profiles/timeline_*_r02z.txt); an under-filled round is
        # charged as idle SMs.  OFFK_WGRAD_CTAS=n restores the plain "n CTAs" rule.
        tiles = m_tiles * n_tiles
        fixed_ctas = int(os.environ.get("OFFK_WGRAD_CTAS", "0"))
        if fixed_ctas:
            split = max(1, min(math.ceil(fixed_ctas / tiles), max(1, kb // 4)))
        else:
            split = _wgrad_split(tiles, kb, _SM_TARGET * (1 if self.prec == L.PREC_TF32X3 else 2), bk or 32)
        mk = (lambda *a, **k: TGemm(*a, geom=geom, x_layout=x_layout, wgrad=True, bk=bk, **k)) if tma else Gemm
        # dW[n][m]: consecutive accumulator rows are consecutive addresses -> transposed float4 adds (offk.h: out_vec = 2)
        rows_vec = (self.tc and dw.data_ptr() % 16 == 0 and geom.kdim % 4 == 0
                    and os.environ.get("OFFK_WGRAD_VEC", "1") != "0")
        g = mk(self, spc, ("wgrad", x_layout, _gkey(geom)), a_src=x, b_src=dy, out=dw, ones_out=db, a_relu=a_relu,
               atomic=True, split_k=split, name=name + ".wgrad", out_vec=2 if rows_vec else None)
        self.flops_bwd += g.flops
        return g

    def _conv_dgrad(self, name, dy, w, dx, geom, *, gate=None, gate_col0=0, gate_first=False, addend=None,
                    add_geom=None, x_layout="nhwc"):
        out = []
        for i, spc in enumerate(T.conv_dgrad_specs(geom, x_layout, "nhwc", "nhwc")):
            add_tabs = None
            if addend is not None and add_geom is not None:
                # addend lives in a differently-shaped buffer: tables of the same logical (img, c, y, x) element
                aspec = T.conv_dgrad_specs(add_geom, "nhwc", "nhwc", "nhwc")[i]
                t = self._tables(("dgrad", "nhwc", _gkey(add_geom), i), aspec)
                add_tabs = (t["out_row"], t["out_col"])
            kw_ = dict(a_src=dy, b_src=w, out=dx, gate=gate, gate_col0=gate_col0, gate_first=gate_first, addend=addend,
                       add_tabs=add_tabs, name=f"{name}.dgrad{i}")
            # dX of a stride-1 conv = a forward conv over dY with flipped taps and pad' = k-1-pad: TMA-fed
            gd = T.ConvGeom(geom.n_img, geom.cout, geom.hout, geom.wout, geom.cin, geom.kh, geom.kw, 1,
                            geom.kh - 1 - geom.pad, geom.y_ctot, geom.y_coff, geom.x_ctot, geom.x_coff) if geom.stride == 1 else None
            ex = spc.extra
            if (self.use_tma and x_layout == "nhwc" and gd is not None and name in self.wd and geom.kh == geom.kw
                    and TGemm.eligible(gd, self.prec) and (gd.hout, gd.wout) == (geom.hin, geom.win)):
                kw_["b_src"] = self.wd[name]
                g = TGemm(self, spc, ("dgrad", x_layout, _gkey(geom), i), geom=gd, **kw_)
            elif (self.use_tma and self.tma_strided_dgrad and x_layout == "nhwc" and geom.stride > 1
                  and (name, ex["a"], ex["b"]) in self.wd and self.tc and geom.cout % 32 == 0
                  and geom.y_ctot % 4 == 0 and geom.y_coff % 4 == 0 and ex["pad_h"] >= 0 and ex["pad_w"] >= 0):
                # one stride-parity class = a stride-1 correlation over dY with the class's R x Q taps (free geometry)
                R, Q = len(ex["rs"]), len(ex["qs"])
                fg = _FreeGeom(n_img=geom.n_img, cin=geom.cout, hin=geom.hout, win=geom.wout, cout=geom.cin, kh=R, kw=Q,
                               stride=1, pad=ex["pad_h"], pad_w=ex["pad_w"], x_ctot=geom.y_ctot, x_coff=geom.y_coff,
                               hout=ex["hc"], wout=ex["wc"], kdim=R * Q * geom.cout)
                kw_["b_src"] = self.wd[(name, ex["a"], ex["b"])]
                kw_["tile_n"] = _pick_tile_n(spc.M, spc.N, _SM_TARGET * (1 if self.prec == L.PREC_TF32X3 else 2))
                g = TGemm(self, spc, ("dgrad", x_layout, _gkey(geom), i), geom=fg, **kw_)
            else:
                g = Gemm(self, spc, ("dgrad", x_layout, _gkey(geom), i), **kw_)
            self.flops_bwd += g.flops
            out.append(g)
        return out

    def _build(self):
        P, N, B, Lg = self.P, self.N, self.B, self.Lseg
        bf, pr, gr = self.buf, self.params, self.grads
        self.flops_fwd = self.flops_bwd = 0.0
        fwd, bwd_units, bwd_stage = [], [], []
        lib = self.lib

        # ============ OFF units (RGB_OFF.py:596-616 and the eight copies)
        self._stencils = {}
        fwd_lane = {"28": 0, "14": 1, "7": 2}
        self.stencil_fwd_steps = OrderedDict()        # stage -> batched stencil launch (bench.py times these)
        stage_levels = OrderedDict((st, [t for t in S.LEVELS if S.LEVEL_STAGE[t] == st]) for st in S.STAGES)
        n_lv = len(S.LEVELS)
        self._st_desc = (L.OffkStencil * n_lv)()      # one contiguous array: stage batches are slices of it
        self._st_io = (L.OffkStencilIO * n_lv)()
        k4_steps = []
        k4_by_stage = {st: [] for st in S.STAGES}
        for li, (tag, (cin, s)) in enumerate(S.LEVELS.items()):
            st = S.LEVEL_STAGE[tag]
            fl = fwd_lane[st]
            ctot, _, members = S.STAGES[st]
            coff = dict(members)[tag]
            geom = T.ConvGeom(N, cin, s, s, S.UNIT_C)
            gd, dgd = bf["gd_" + tag], bf["dgd_" + tag]
            # K1: gen (ReLU) and down (linear) 1x1 convs as ONE GEMM with 160 output channels
            # 7x7 taps: a 196-byte channel stride is not a legal TMA stride -> one channels-last copy of the tap per
            # step feeds both the forward GEMM (dense tensor map) and the weight gradient (im2col_t), all TMA-fed
            via_copy = (self.tc and self.use_tma and (s * s) % 4 != 0 and cin % 32 == 0
                        and os.environ.get("OFFK_NO_TAP_COPY", "0") != "1")
            if via_copy:
                tapT = self._buf("tapT_" + tag, N, s, s, cin)

                def kt(stream, tag=tag, tapT=tapT, cin=cin, hw=s * s):
                    L.check(lib.offk_nchw_to_nhwc(_ptr(self.taps[tag]), _ptr(tapT), N, cin, hw, stream), "tapT_" + tag)
                fwd.append(_nm(kt, "tapT_" + tag, writes=[tapT], lane=fl))
                k1 = self._conv_fwd("unit_" + tag, tapT, gd, geom, self._unit_w(self.params_flat, tag),
                                    self._unit_b(self.params_flat, tag), relu_cols=S.GEN_C, x_layout="nhwc", no_split=True)
            else:
                k1 = self._conv_fwd("unit_" + tag, self.taps[tag], gd, geom, self._unit_w(self.params_flat, tag),
                                    self._unit_b(self.params_flat, tag), relu_cols=S.GEN_C, x_layout="nchw")
                self._tap_users[tag].append(getattr(k1, "gemm", k1))
            fwd.append(_on(k1, fl))
            # K2 / K3 descriptors: fused spatial stencil + temporal difference + dropout + cat, into the stage buffer
            sd = self._st_desc[li]
            sd.B, sd.L, sd.Cg, sd.Cs, sd.K, sd.H, sd.W = B, Lg, S.GEN_C, S.DOWN_C, 1, s, s
            sd.g_fs = sd.d_fs = S.UNIT_C * s * s
            sd.g_ps = sd.d_ps = S.UNIT_C
            sd.out_ctot, sd.out_coff, sd.index_mode = ctot, coff, self.index_mode
            sd.drop_mode, sd.keep_scale, sd.drop_p = L.DROP_NONE, 1.0 / (1.0 - S.DROP_P), S.DROP_P
            self._stencils[tag] = sd
            if self.variant == "rgb":
                w3, b3 = pr[f"motion_spatial_grad_{tag}.weight"], pr[f"motion_spatial_grad_{tag}.bias"]
                dw3, db3 = gr[f"motion_spatial_grad_{tag}.weight"], gr[f"motion_spatial_grad_{tag}.bias"]
            else:
                w3, b3, dw3, db3 = self.sobel_w, None, None, None
            io = self._st_io[li]
            io.g, io.d = gd.data_ptr(), gd.data_ptr() + 4 * S.GEN_C            # D = channels [128,160) of every pixel
            io.w, io.bias = w3.data_ptr(), (b3.data_ptr() if b3 is not None else None)
            io.out, io.dout = bf["F" + st].data_ptr(), bf["dF" + st].data_ptr()
            io.dg, io.dd = dgd.data_ptr(), dgd.data_ptr() + 4 * S.GEN_C
            io.dg_fs = io.dd_fs = S.UNIT_C * s * s
            io.dw = dw3.data_ptr() if dw3 is not None else None
            io.dbias = db3.data_ptr() if db3 is not None else None
            # K4: weight / bias gradient of the fused 1x1 (no dX for the frozen taps unless asked, train_off.py:39-46)
            if via_copy:
                k4 = self._conv_wgrad("unit_" + tag, tapT, dgd, geom, self._unit_w(self.grads_flat, tag),
                                      self._unit_b(self.grads_flat, tag), x_layout="nhwc", force_tma=True)
            else:
                k4 = self._conv_wgrad("unit_" + tag, self.taps[tag], dgd, geom,
                                      self._unit_w(self.grads_flat, tag), self._unit_b(self.grads_flat, tag), x_layout="nchw")
                self._tap_users[tag].append(k4)
            k4.unit_tag = tag                       # data parallelism: this unit's gradients are final once k4 is done
            grp = [_on(k4, li % 3)]
            if self.tap_grads:
                grp += [_on(g, li % 3) for g in self._conv_dgrad("unit_" + tag, dgd, self._unit_w(self.params_flat, tag),
                                                                  self.tap_grad[tag], geom, x_layout="nchw")]
            k4_steps += grp
            k4_by_stage[st] += grp
        # K2: ONE stencil launch per stage-fusion buffer (its units' GEMMs precede it on the same lane)
        tags = list(S.LEVELS)
        # forward stencil launches: the 28-stage units on their own (they open the critical path); the 14- and 7-stage
        # units together (the 7x7 levels are 7 MB each: on their own they run at a fifth of the bandwidth)
        merge = os.environ.get("OFFK_STENCIL_FWD_MERGE", "1") == "1"
        groups = [("28", ["28"]), ("14+7", ["14", "7"])] if merge else [(st, [st]) for st in S.STAGES]
        self.stencil_fwd_tags = OrderedDict()         # launch name -> level tags (bench.py: algorithmic bytes)
        for gname, sts in groups:
            lv_tags = [t for st in sts for t in stage_levels[st]]
            i0, n = tags.index(lv_tags[0]), len(lv_tags)
            assert tags[i0:i0 + n] == lv_tags

            def k2(stream, i0=i0, n=n, gname=gname):
                L.check(lib.offk_stencil_diff_fwd_batch(n, C.byref(self._st_desc[i0]), C.byref(self._st_io[i0]), stream),
                        "stencil_fwd_" + gname)
            step = _nm(k2, "stencil_fwd_" + gname, reads=[bf["gd_" + t] for t in lv_tags], writes=[bf["F" + st] for st in sts],
                       lane=fwd_lane[sts[0]])
            fwd.append(step)
            self.stencil_fwd_steps[gname] = step
            self.stencil_fwd_tags[gname] = lv_tags
        # K3: the backward of all nine units in ONE launch (every stage gradient dF* is complete by then)
        grad3 = [gr[f"motion_spatial_grad_{t}.{k}"] for t in tags for k in ("weight", "bias")] if self.variant == "rgb" else []

        # two launches on two lanes: the temporal half streams (dG), the spatial half (dD, tap gradients) is
        # instruction-heavy; they write disjoint channels of dgd_X, so the pair is exempt from the hazard analysis
        def k3t(stream):
            L.check(lib.offk_stencil_diff_bwd_batch_part(n_lv, self._st_desc, self._st_io, 1, stream), "stencil_bwd_temporal")

        def k3s(stream):
            L.check(lib.offk_stencil_diff_bwd_batch_part(n_lv, self._st_desc, self._st_io, 2, stream), "stencil_bwd_spatial")
        rd3 = [bf["dF" + st] for st in S.STAGES] + [bf["gd_" + t] for t in tags]
        k3t = _nm(k3t, "stencil_bwd_temporal", reads=rd3, writes=[bf["dgd_" + t] for t in tags], lane=0)
        k3s = _nm(k3s, "stencil_bwd_spatial", reads=rd3, writes=[bf["dgd_" + t] for t in tags] + grad3, lane=1)
        k3s.independent_of = (k3t,)

        def k3(stream):
            L.check(lib.offk_stencil_diff_bwd_batch(n_lv, self._st_desc, self._st_io, stream), "stencil_bwd")
        k3 = _nm(k3, "stencil_bwd", reads=rd3, writes=[bf["dgd_" + t] for t in tags] + grad3, lane=0)
        # Early unit backward (optional): dF7[:, :320] is final right after motion_conv_trans' data gradient and
        # dF14[:, :800] after motion_conv_trans_14's, so the 7- and 14-stage units' stencil backward + weight gradients
        # can run on side lanes under the rest of the stage chain.  Measured at config 2: 2.90 ms/step vs 2.84 ms with
        # all nine units at the tail -- the HBM-bound unit kernels slow the chain more than the overlap saves -- so off.
        self.unit_bwd_early = os.environ.get("OFFK_UNIT_BWD_EARLY", "0") == "1"
        unit_bwd = {}
        for st, lv_tags in stage_levels.items():
            i0, n = tags.index(lv_tags[0]), len(lv_tags)
            g3 = [gr[f"motion_spatial_grad_{t}.{k}"] for t in lv_tags for k in ("weight", "bias")] if self.variant == "rgb" else []

            def k3st(stream, i0=i0, n=n, st=st):
                L.check(lib.offk_stencil_diff_bwd_batch(n, C.byref(self._st_desc[i0]), C.byref(self._st_io[i0]), stream),
                        "stencil_bwd_" + st)
            lane3 = 0 if st == "28" else 2
            stp = _nm(k3st, "stencil_bwd_" + st, reads=[bf["dF" + st]] + [bf["gd_" + t] for t in lv_tags],
                      writes=[bf["dgd_" + t] for t in lv_tags] + g3, lane=lane3)
            grp = list(k4_by_stage[st])
            if self.unit_bwd_early:
                for j, g_ in enumerate(grp):                       # early groups stay off lane 0 (the dgrad chain)
                    g_.lane = (2, 1)[j % 2] if st != "28" else (0, 2, 1)[j % 3]
            unit_bwd[st] = [stp] + grp
        if self.unit_bwd_early:
            bwd_units += unit_bwd["28"]
        else:
            bwd_units += [k3t, k3s] if os.environ.get("OFFK_STENCIL_BWD_SPLIT", "0") == "1" else [k3]
            bwd_units += k4_steps

        # ============ stage convs
        def geom_of(name, n_img, s_in, x_ctot=0, x_coff=0, y_ctot=0, y_coff=0):
            _, cout, cin, k, stride, pad = S.CONV_BY_NAME[name]
            return T.ConvGeom(n_img, cin, s_in, s_in, cout, k, k, stride, pad, x_ctot, x_coff, y_ctot, y_coff)

        def W(n): return self.wp[n] if n in self.wp else pr[n + ".weight"]     # OHWI copy for KxK convs
        def Bv(n): return pr[n + ".bias"]
        def dW(n): return self.dwp[n] if n in self.dwp else gr[n + ".weight"]
        def dB(n): return gr[n + ".bias"]

        def layer(name, x, y, s_in, *, relu=False, a_relu=False, addend=None, relu_post=False,
                  x_ctot=0, x_coff=0, y_ctot=0, y_coff=0, lane=0, aux=None):
            """forward conv; returns geom for the gradient wiring."""
            geom = geom_of(name, P, s_in, x_ctot, x_coff, y_ctot, y_coff)
            fwd.append(_on(self._conv_fwd(name, x, y, geom, W(name), Bv(name), relu=relu, a_relu=a_relu, addend=addend,
                                          relu_post=relu_post, aux=aux), lane))
            return geom

        # ---- resolution 28 (RGB_OFF.py:655-685)
        # pre-ReLU output kept for the branch (:665); its ReLU'd copy (:658) is the epilogue's second output
        g_t28 = layer("motion_conv_trans_28", bf["F28"], bf["t28"], 28, aux=(bf["t28r"], None, 0, None))
        if not getattr(fwd[-1], "aux_fused", False):
            n28 = bf["t28"].numel()
            fwd.append(_nm(lambda stream: L.check(lib.offk_relu_gate(_ptr(bf["t28"]), _ptr(bf["t28"]), n28, _ptr(bf["t28r"]),
                                                                      stream), "relu_t28"), "relu_t28",
                           reads=[bf["t28"]], writes=[bf["t28r"]]))
        g_c1a = layer("motion_conv1_trans_28a", bf["t28r"], bf["h1_28a"], 14, relu=True)
        g_c2a = layer("motion_conv2_trans_28a", bf["h1_28a"], bf["h2_28a"], 14, relu=True)
        g_bra = layer("motion_conv_branch_28a", bf["t28"], bf["br28"], 14, lane=2)
        g_c3a = layer("motion_conv3_trans_28a", bf["h2_28a"], bf["s28a"], 14, addend=bf["br28"], relu_post=True)
        g_c1b = layer("motion_conv1_trans_28b", bf["s28a"], bf["h1_28b"], 14, relu=True)
        g_c2b = layer("motion_conv2_trans_28b", bf["h1_28b"], bf["h2_28b"], 14, relu=True)
        g_c3b = layer("motion_conv3_trans_28b", bf["h2_28b"], bf["s28b"], 14, addend=bf["s28a"], relu_post=True)
        g_c1c = layer("motion_conv1_trans_28c", bf["s28b"], bf["h1_28c"], 14, relu=True)
        g_c2c = layer("motion_conv2_trans_28c", bf["h1_28c"], bf["h2_28c"], 14, relu=True)
        # sum_28c goes straight into the 14-stage fusion buffer at channel 800 (cat, :760)
        g_c3c = geom_of("motion_conv3_trans_28c", P, 14, y_ctot=1056, y_coff=800)
        g_s28b_as_f14 = T.ConvGeom(P, 256, 14, 14, 256, y_ctot=256)  # tables for the s28b addend (plain [P,14,14,256])
        spc_add = T.conv_fwd_spec(g_s28b_as_f14, "nhwc", "nhwc")
        tabs_add = self._tables(("fwd", "nhwc", _gkey(g_s28b_as_f14)), spc_add)
        fwd.append(self._conv_fwd("motion_conv3_trans_28c", bf["h2_28c"], bf["F14"], g_c3c,
                                  W("motion_conv3_trans_28c"), Bv("motion_conv3_trans_28c"), addend=bf["s28b"],
                                  add_tabs=(tabs_add["out_row"], tabs_add["out_col"]), relu_post=True))

        # ---- resolution 14 (RGB_OFF.py:759-780)
        g_t14 = layer("motion_conv_trans_14", bf["F14"], bf["t14"], 14, relu=True)
        g_14a1 = layer("motion_conv1_trans_14a", bf["t14"], bf["h1_14a"], 7, relu=True)
        g_14a2 = layer("motion_conv2_trans_14a", bf["h1_14a"], bf["h2_14a"], 7, relu=True)
        g_14ex = layer("motion_conv_expand_trans_14a", bf["t14"], bf["ex14"], 7, lane=2)
        g_14a3 = layer("motion_conv3_trans_14a", bf["h2_14a"], bf["s14a"], 7, addend=bf["ex14"], relu_post=True)
        g_14b1 = layer("motion_conv1_trans_14b", bf["s14a"], bf["h1_14b"], 7, relu=True)
        g_14b2 = layer("motion_conv2_trans_14b", bf["h1_14b"], bf["h2_14b"], 7, relu=True)
        # 3x3 + ReLU (:316,:778), kept for the backward gate; sum_14b = relu(s14a + h3_14b) -> 7-stage fusion buffer channels
        # [320,832) (:779-780, cat :832) is the same epilogue's second output
        g_f7slice = T.ConvGeom(P, 512, 7, 7, 512, y_ctot=832, y_coff=320)
        tabs_f7 = self._tables(("fwd", "nhwc", _gkey(g_f7slice)), T.conv_fwd_spec(g_f7slice, "nhwc", "nhwc"))
        g_14b3 = layer("motion_conv3_trans_14b", bf["h2_14b"], bf["h3_14b"], 7, relu=True,
                       aux=(bf["F7"], tabs_f7["out_row"], 0, bf["s14a"]))      # the row table carries the channel offset 320
        if not getattr(fwd[-1], "aux_fused", False):
            fwd.append(self._add_into_slice(bf["s14a"], bf["h3_14b"], bf["F7"], 832, 320, 512, 49))

        # ---- heads 28 / 14 (RGB_OFF.py:783-793)
        fwd.append(_nm(lambda stream: L.check(lib.offk_maxpool3s2_fwd(_ptr(bf["F14"]), P, 256, 14, 14, 1056, 800,
                                                                       _ptr(bf["p28"]), stream), "maxpool28"), "maxpool28",
                       reads=[bf["F14"]], writes=[bf["p28"]], lane=1))
        # each head = ONE launch: avgpool 7x7 -> dropout -> Linear (-> segment consensus)   
        fwd.append(_on(self._head_fwd("28", bf["p28"], 256, 256, 0, "fc_action_motion_28"), 1))
        fwd.append(_on(self._head_fwd("14", bf["F7"], 512, 832, 320, "fc_action_motion_14"), 1))

        # ---- resolution 7 (RGB_OFF.py:831-847)
        g_t7 = layer("motion_conv_trans", bf["F7"], bf["t7"], 7, relu=True)
        g_71 = layer("motion_conv1_trans", bf["t7"], bf["h1_7"], 7, relu=True)
        g_72 = layer("motion_conv2_trans", bf["h1_7"], bf["h2_7"], 7, relu=True)
        g_7br = layer("motion_conv_branch_trans", bf["t7"], bf["br7"], 7, lane=2)
        g_73 = layer("motion_conv3_trans", bf["h2_7"], bf["s7"], 7, addend=bf["br7"])        # no final ReLU (:841)
        fwd.append(self._head_fwd("7", bf["s7"], 1024, 1024, 0, "fc_action_motion"))

        # ============ backward of the stages (reverse order); every d_* buffer holds dL/d(pre-activation)
        bs = bwd_stage
        d = lambda n: bf["d_" + n]
        # upstream gradients: [B,101] with the consensus (its backward, g.expand / T, is folded into the head kernels),
        # [P,101] without
        n_up = B if self.consensus else P
        self.d_out7 = torch.zeros(n_up, S.NUM_CLASSES, device=self.device)
        self.d_out14 = torch.zeros(n_up, S.NUM_CLASSES, device=self.device)
        # heads (fc28 receives no gradient: never returned, RGB_OFF.py:860): Linear weight / bias gradient, and
        # ds7 = avgpool'(dropout'(Linear'(d_out7))) in one pass
        bs.extend(self._head_bwd("7", self.d_out7, d("s7"), 1024, 1024, 0, "fc_action_motion", act=None, accumulate=False))

        def back(name, x, dy, geom, *, a_relu=False):
            bs.append(_on(self._conv_wgrad(name, x, dy, geom, dW(name), dB(name), a_relu=a_relu), 1))   # wgrad lane

        def dgrad(name, dy, dx, geom, lane=0, **kw):
            bs.extend(_on(g, lane) for g in self._conv_dgrad(name, dy, W(name), dx, geom, **kw))

        # ---- 7
        back("motion_conv3_trans", bf["h2_7"], d("s7"), g_73)
        dgrad("motion_conv3_trans", d("s7"), d("h2_7"), g_73, gate=bf["h2_7"])
        back("motion_conv_branch_trans", bf["t7"], d("s7"), g_7br)
        dgrad("motion_conv_branch_trans", d("s7"), d("tmp7"), g_7br, lane=2)
        back("motion_conv2_trans", bf["h1_7"], d("h2_7"), g_72)
        dgrad("motion_conv2_trans", d("h2_7"), d("h1_7"), g_72, gate=bf["h1_7"])
        back("motion_conv1_trans", bf["t7"], d("h1_7"), g_71)
        dgrad("motion_conv1_trans", d("h1_7"), d("t7"), g_71, gate=bf["t7"], addend=d("tmp7"))
        back("motion_conv_trans", bf["F7"], d("t7"), g_t7)
        dgrad("motion_conv_trans", d("t7"), bf["dF7"], g_t7)
        # d sum_14b = (dF7[:,320:] + head-14 pool gradient) * [sum_14b > 0]   (in place in the dF7 slice)
        bs.extend(self._head_bwd("14", self.d_out14, bf["dF7"], 512, 832, 320, "fc_action_motion_14", act=bf["F7"], accumulate=True))
        if self.unit_bwd_early:          # after the in-place update of dF7[:, 320:], so the main lane never waits on these
            bs.extend(unit_bwd["7"])
        # ---- 14b:  sum_14b = relu(s14a + relu(conv3_14b(h2b)))
        bs.append(_nm(lambda stream: L.check(lib.offk_gate_copy(_ptr(bf["dF7"]), 832, 320, _ptr(bf["h3_14b"]), 512, 0,
                                                                _ptr(d("h3_14b")), 512, 0, P, 512, 49, stream), "gate_h3b"),
                      "gate_h3b", reads=[bf["dF7"], bf["h3_14b"]], writes=[d("h3_14b")]))
        back("motion_conv3_trans_14b", bf["h2_14b"], d("h3_14b"), g_14b3)
        dgrad("motion_conv3_trans_14b", d("h3_14b"), d("h2_14b"), g_14b3, gate=bf["h2_14b"])
        back("motion_conv2_trans_14b", bf["h1_14b"], d("h2_14b"), g_14b2)
        dgrad("motion_conv2_trans_14b", d("h2_14b"), d("h1_14b"), g_14b2, gate=bf["h1_14b"])
        back("motion_conv1_trans_14b", bf["s14a"], d("h1_14b"), g_14b1)
        g_slice7 = T.ConvGeom(P, 512, 7, 7, 128, x_ctot=832, x_coff=320)   # addresses dF7[:,320:832] like a conv input
        dgrad("motion_conv1_trans_14b", d("h1_14b"), d("s14a"), g_14b1, gate=bf["s14a"], addend=bf["dF7"],
              add_geom=g_slice7)
        # ---- 14a
        back("motion_conv3_trans_14a", bf["h2_14a"], d("s14a"), g_14a3)
        dgrad("motion_conv3_trans_14a", d("s14a"), d("h2_14a"), g_14a3, gate=bf["h2_14a"])
        back("motion_conv_expand_trans_14a", bf["t14"], d("s14a"), g_14ex)
        dgrad("motion_conv_expand_trans_14a", d("s14a"), d("tmp14"), g_14ex, lane=2)
        back("motion_conv2_trans_14a", bf["h1_14a"], d("h2_14a"), g_14a2)
        dgrad("motion_conv2_trans_14a", d("h2_14a"), d("h1_14a"), g_14a2, gate=bf["h1_14a"])
        back("motion_conv1_trans_14a", bf["t14"], d("h1_14a"), g_14a1)
        dgrad("motion_conv1_trans_14a", d("h1_14a"), d("t14"), g_14a1, gate=bf["t14"], addend=d("tmp14"))
        back("motion_conv_trans_14", bf["F14"], d("t14"), g_t14)
        # dF14; channels >= 800 are sum_28c = relu(.) -> gate them here (fc28 contributes nothing)
        dgrad("motion_conv_trans_14", d("t14"), bf["dF14"], g_t14, gate=bf["F14"], gate_col0=800)
        if self.unit_bwd_early:
            bs.extend(unit_bwd["14"])
        # ---- 28c / 28b (identity residuals): the incoming gradient of 28c is the dF14 slice [800,1056)
        g_dy_c3c = g_c3c
        back("motion_conv3_trans_28c", bf["h2_28c"], bf["dF14"], g_dy_c3c)
        dgrad("motion_conv3_trans_28c", bf["dF14"], d("h2_28c"), g_dy_c3c, gate=bf["h2_28c"])
        back("motion_conv2_trans_28c", bf["h1_28c"], d("h2_28c"), g_c2c)
        dgrad("motion_conv2_trans_28c", d("h2_28c"), d("h1_28c"), g_c2c, gate=bf["h1_28c"])
        back("motion_conv1_trans_28c", bf["s28b"], d("h1_28c"), g_c1c)
        g_slice14 = T.ConvGeom(P, 256, 14, 14, 64, x_ctot=1056, x_coff=800)
        dgrad("motion_conv1_trans_28c", d("h1_28c"), d("s28b"), g_c1c, gate=bf["s28b"], addend=bf["dF14"],
              add_geom=g_slice14)
        back("motion_conv3_trans_28b", bf["h2_28b"], d("s28b"), g_c3b)
        dgrad("motion_conv3_trans_28b", d("s28b"), d("h2_28b"), g_c3b, gate=bf["h2_28b"])
        back("motion_conv2_trans_28b", bf["h1_28b"], d("h2_28b"), g_c2b)
        dgrad("motion_conv2_trans_28b", d("h2_28b"), d("h1_28b"), g_c2b, gate=bf["h1_28b"])
        back("motion_conv1_trans_28b", bf["s28a"], d("h1_28b"), g_c1b)
        dgrad("motion_conv1_trans_28b", d("h1_28b"), d("s28a"), g_c1b, gate=bf["s28a"], addend=d("s28b"))
        # ---- 28a: the branch consumes the PRE-ReLU conv_trans_28 output (:665), conv1 the ReLU'd one (:659)
        back("motion_conv3_trans_28a", bf["h2_28a"], d("s28a"), g_c3a)
        dgrad("motion_conv3_trans_28a", d("s28a"), d("h2_28a"), g_c3a, gate=bf["h2_28a"])
        back("motion_conv_branch_28a", bf["t28"], d("s28a"), g_bra)
        dgrad("motion_conv_branch_28a", d("s28a"), d("tmp28"), g_bra, lane=2)
        back("motion_conv2_trans_28a", bf["h1_28a"], d("h2_28a"), g_c2a)
        dgrad("motion_conv2_trans_28a", d("h2_28a"), d("h1_28a"), g_c2a, gate=bf["h1_28a"])
        back("motion_conv1_trans_28a", bf["t28r"], d("h1_28a"), g_c1a)
        dgrad("motion_conv1_trans_28a", d("h1_28a"), d("t28"), g_c1a, gate=bf["t28"], gate_first=True,
              addend=d("tmp28"))
        back("motion_conv_trans_28", bf["F28"], d("t28"), g_t28)
        dgrad("motion_conv_trans_28", d("t28"), bf["dF28"], g_t28)

        # weight re-layouts before the forward (one launch), OHWI -> OIHW weight gradients after the backward
        wc, wci, pflat = self.wc_flat, self.wc_idx, self.params_flat
        wc_lane = int(os.environ.get("OFFK_WC_LANE", "2"))
        pre = [_nm(lambda stream: L.check(lib.offk_gather_copy(_ptr(pflat), _ptr(wci), _ptr(wc), wc.numel(), stream),
                                          "weight copies"), "weight_copies", writes=[wc], lane=wc_lane)]
        if self.presplit:
            # 3xTF32 residuals of every weight, once per step (the GEMMs then get both weight tiles by TMA)
            plo, wlo = self.params_lo, self.wc_lo
            pre.append(_nm(lambda stream: L.check(lib.offk_tf32_residual(_ptr(pflat), _ptr(plo), pflat.numel(), stream),
                                                  "weight residuals"), "weight_residuals_params", writes=[plo], lane=0))
            pre.append(_nm(lambda stream: L.check(lib.offk_tf32_residual(_ptr(wc), _ptr(wlo), wc.numel(), stream),
                                                  "weight residuals"), "weight_residuals_copies", reads=[wc], writes=[wlo],
                           lane=wc_lane))
        # OHWI -> OIHW accumulation of every KxK weight gradient: one launch
        kxk = [(name, cout, cin, k) for name, cout, cin, k, _, _ in S.STAGE_CONVS if k > 1]
        self._unperm = (L.OffkPermute * len(kxk))()
        for i, (name, cout, cin, k) in enumerate(kxk):
            it = self._unperm[i]
            it.src, it.dst = self.dwp[name].data_ptr(), gr[name + ".weight"].data_ptr()
            it.cout, it.cin, it.kh, it.kw = cout, cin, k, k
        post = [_nm(lambda stream: L.check(lib.offk_permute_weight_batch(len(kxk), self._unperm, 2, stream), "unpermute"),
                    "unpermute_kxk_weight_grads", reads=[self.dwp_flat], writes=[gr[n + ".weight"] for n, _, _, _ in kxk], lane=1)]
        self.fwd_steps = pre + fwd
        # kernels of liboffk launched per pass (cudaMemsetAsync zero-fills are not kernels)
        count = lambda steps: sum(len([n for n in _names(st) if not _is_memset(n)]) for st in steps)
        # gradient accumulators start from zero (split-K / atomic accumulation); grads_flat only when asked
        self._zero_grads = True
        gflat, dwflat = self.grads_flat, self.dwp_flat
        zero = [_nm(lambda stream: self._zero_grads and L.check(lib.offk_fill_zero(_ptr(gflat), gflat.numel(), stream),
                                                                 "zero grads"), "zero grads_flat", writes=[gflat]),
                _nm(lambda stream: L.check(lib.offk_fill_zero(_ptr(dwflat), dwflat.numel(), stream), "zero dwp"),
                    "zero dwp_flat", writes=[dwflat], lane=1)]
        # the stage / head gradients are complete before the units run: their bucket can be all-reduced meanwhile
        self.bwd_stage_steps = zero + bwd_stage + post
        self.bwd_unit_steps = bwd_units
        self.bwd_steps = self.bwd_stage_steps + self.bwd_unit_steps
        self.launches_fwd = count(self.fwd_steps)
        self.launches_bwd = count(self.bwd_steps)
        # parameters and taps are read-only inside a pass: never a hazard
        ro = [self.params_flat] + [t for ts in self.tap_sets for t in ts.values()]
        if self.variant == "flow":
            ro.append(self.sobel_w)
        self.fwd_sched = Schedule(self.fwd_steps, ignore=ro)
        self.bwd_sched = Schedule(self.bwd_steps, ignore=ro)
        self._lanes = max(self.fwd_sched.n_lanes, self.bwd_sched.n_lanes)
        self._side = None
        unit_end = max(self.layout[f"motion_{k}_{t}.bias"][0] + (S.GEN_C if k == "conv_gen" else S.DOWN_C)
                       for t in S.LEVELS for k in (("conv_gen", "spatial_down") + (("spatial_grad",) if self.variant == "rgb" else ())))
        self.unit_range = (0, (unit_end + 3) // 4 * 4)          # flat offsets of the nine units' parameters
        # per unit: [first parameter of the level, first parameter of the next level) -- contiguous, 16-byte aligned
        starts = [self.layout[f"motion_conv_gen_{t}.weight"][0] for t in S.LEVELS] + [self.unit_range[1]]
        self.unit_ranges = OrderedDict((t, (starts[i], starts[i + 1])) for i, t in enumerate(S.LEVELS))
        self.stage_range = (self.unit_range[1], self.n_flat)    # stage convs + FC heads

    # ------------------------------------------------------------------ small step factories
    def _add_into_slice(self, a, b, dst, ctot, coff, c, hw):
        P, lib = self.P, self.lib
        return _nm(lambda stream: L.check(lib.offk_add_relu_slice(_ptr(a), _ptr(b), _ptr(dst), ctot, coff, P, c, hw, 1,
                                                                   stream), "sum_14b"), "sum_14b", reads=[a, b], writes=[dst])

    @staticmethod
    def _site_salt(site):
        """Constant 64-bit salt of dropout call site ``site`` (12 per forward, RGB_OFF.py:612..:845): splitmix64(site).
        The kernels add the step seed, which lives in device memory (``seed_state``; offk.h: seed_dev), and hash the sum
        with the element index, so every bit of both reaches every keep decision -- and a captured CUDA graph, whose
        kernel arguments are frozen, still draws new masks whenever the device word changes."""
        m = 0xFFFFFFFFFFFFFFFF
        z = ((site + 1) * 0xD1B54A32D192ED03) & m
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
        return z ^ (z >> 31)

    def _seed_dev(self):
        return _ptr(self.seed_state) if self.drop_mode == L.DROP_SEED else None

    def _head_fwd(self, k, x, c, ctot, coff, fcname):
        """global_pool -> dropout -> fc_action_motion* (-> consensus) as ONE launch (RGB_OFF.py:784-787,790-793,844-847;
        Flow_OFF.py:867-876).  Linear ops only: exact fp32 in every precision mode."""
        lib, P, bf = self.lib, self.P, self.buf
        site = {"28": 9, "14": 10, "7": 11}[k]
        w, b = self.params[fcname + ".weight"], self.params[fcname + ".bias"]
        T_ = self.Lseg - 1 if self.consensus else 1
        cons = bf["cfc" + k] if self.consensus else None

        def run(stream):
            m = self.masks["fc" + k] if self.drop_mode == L.DROP_MASK else None
            L.check(lib.offk_head_fwd(_ptr(x), P, c, 49, ctot, coff, self.drop_mode, _ptr(m), self._site_salt(site),
                                      self._seed_dev(), S.DROP_P, 1.0 / (1.0 - S.DROP_P), _ptr(w), _ptr(b), S.NUM_CLASSES, T_,
                                      _ptr(bf["pool" + k]), _ptr(bf["fc" + k]), _ptr(cons), stream), "head_fwd" + k)
        self.flops_fwd += 2.0 * P * c * S.NUM_CLASSES
        return _nm(run, "head_fwd" + k, reads=[x], writes=[bf["pool" + k], bf["fc" + k], cons])

    def _head_bwd(self, k, dout, dx, c, ctot, coff, fcname, act, accumulate):
        """Backward of _head_fwd in two launches: Linear weight / bias gradient; Linear data gradient + dropout' + average
        pool backward (+ the producer's ReLU', + accumulation into a slice that already holds a gradient)."""
        lib, P, bf = self.lib, self.P, self.buf
        site = {"28": 9, "14": 10, "7": 11}[k]
        w = self.params[fcname + ".weight"]
        dw, db = self.grads[fcname + ".weight"], self.grads[fcname + ".bias"]
        T_ = self.Lseg - 1 if self.consensus else 1

        def call(stream, dx_, dw_, db_):
            m = self.masks["fc" + k] if self.drop_mode == L.DROP_MASK else None
            L.check(lib.offk_head_bwd(_ptr(dout), P, c, 49, ctot, coff, self.drop_mode, _ptr(m), self._site_salt(site),
                                      self._seed_dev(), S.DROP_P, 1.0 / (1.0 - S.DROP_P), _ptr(w), S.NUM_CLASSES, T_,
                                      _ptr(bf["pool" + k]), _ptr(act), int(accumulate), _ptr(dx_), _ptr(dw_), _ptr(db_), stream),
                    "head_bwd" + k)
        self.flops_bwd += 4.0 * P * c * S.NUM_CLASSES
        # the data gradient opens the dgrad chain (lane 0); the weight gradient runs beside it on the weight-gradient lane
        dgrad = _nm(lambda stream: call(stream, dx, None, None), "head_bwd%s.dgrad" % k,
                    reads=[dout, act, dx if accumulate else None], writes=[dx], lane=0)
        wgrad = _nm(lambda stream: call(stream, None, dw, db), "head_bwd%s.wgrad" % k, reads=[dout, bf["pool" + k]],
                    writes=[dw, db], lane=1)
        return [dgrad, wgrad]

    # ------------------------------------------------------------------ run
    def _set_dropout(self, train, masks, seed):
        if not train:
            self.drop_mode, self.masks = L.DROP_NONE, None
        elif masks is not None:
            self.drop_mode = L.DROP_MASK
            self.masks = {k: v.to(self.device, torch.uint8).contiguous() for k, v in masks.items()}
        else:
            self.drop_mode, self.masks = L.DROP_SEED, None
            self.drop_seed = int(seed)
        for i, (tag, sd) in enumerate(self._stencils.items()):
            sd.drop_mode = self.drop_mode
            sd.seed = self._site_salt(i)
            sd.seed_dev = self.seed_state.data_ptr() if self.drop_mode == L.DROP_SEED else None
            sd.keep_mask = self.masks[tag].data_ptr() if self.drop_mode == L.DROP_MASK else None

    def set_taps(self, taps: dict):
        for tag, t in self.taps.items():
            t.copy_(taps[tag], non_blocking=True)

    def select_taps(self, idx: int):
        """Make input set ``idx`` the one forward() and backward() read (re-binds the A operand of the unit GEMMs)."""
        self._tap_set = idx
        self.taps = self.tap_sets[idx]
        for tag, users in self._tap_users.items():
            ptr = self.taps[tag].data_ptr()
            for g in users:
                g.set_a_src(ptr)

    def stage_taps(self, taps: dict):
        """Asynchronously copy ``taps`` (host -- ideally pinned -- or device tensors) into the IDLE input set on a
        side stream, so the transfer overlaps the forward/backward still running on the other set.  Returns
        (set index, event); pass the index to select_taps() after making the compute stream wait on the event."""
        idx = 1 - self._tap_set
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        # everything enqueued so far (incl. the last backward that read set idx) must finish before it is overwritten
        self._copy_stream.wait_stream(main)
        with torch.cuda.stream(self._copy_stream):
            for tag, t in self.tap_sets[idx].items():
                t.copy_(taps[tag], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        return idx, ev

    def forward(self, taps: dict = None, train: bool = False, masks: dict = None, seed: int = 0, graph: bool = False):
        """Run the forward plan.  Returns (fc7, fc28, fc14): [P,101] each, or [B,101] with consensus.
        graph=True replays the plan as ONE captured CUDA graph (captured on first use per input set and dropout mode):
        the same kernels, bit for bit, without the ~60 host-side launches."""
        if taps is not None:
            self.set_taps(taps)
        self._set_dropout(train, masks, seed)
        self.generation += 1            # backward() consumes the activations of THIS forward (one set of static buffers)
        if self.drop_mode == L.DROP_SEED:
            main = torch.cuda.current_stream(self.device)
            L.check(self.lib.offk_seed_set(_ptr(self.seed_state), self.drop_seed & 0xFFFFFFFFFFFFFFFF,
                                           C.c_void_p(main.cuda_stream)), "seed_set")
        if graph:
            self._replay("fwd")
        else:
            self._issue("fwd")
        pre = "cfc" if self.consensus else "fc"
        return self.buf[pre + "7"], self.buf[pre + "28"], self.buf[pre + "14"]

    # ------------------------------------------------------------------ CUDA graphs
    def _issue(self, which):
        streams = self._fork()
        (self.fwd_sched if which == "fwd" else self.bwd_sched).run(streams)
        self._join(streams)

    def _replay(self, which):
        if self.drop_mode == L.DROP_MASK:
            raise RuntimeError("graph replay supports eval mode and seeded dropout; injected masks are an eager-only test facility")
        key = (which, self._tap_set, self.drop_mode)
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = self._capture(which)
        g.replay()

    def _capture(self, which):
        """Capture one pass (all lanes: the side streams fork from and join the capturing stream through the schedule's
        events, so the graph keeps the multi-stream concurrency of the eager plan)."""
        main = torch.cuda.current_stream(self.device)
        if self._cap_stream is None:
            self._cap_stream = torch.cuda.Stream(device=self.device)
        cap = self._cap_stream
        cap.wait_stream(main)
        with torch.cuda.stream(cap):
            self._issue(which)          # eager once on this stream: tensor maps encoded, kernel attributes set
        cap.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cap):
            self._issue(which)
        main.wait_stream(cap)
        return g

    def launch_names(self):
        """One name per device launch of forward() + backward(), in issue order (for annotating ncu launch lists)."""
        out = [n for st in self.fwd_steps for n in _names(st)]
        out += [n for st in self.bwd_steps for n in _names(st)]    # (the d_out copies are cudaMemcpyAsync, not kernels)
        return [n for n in out if not _is_memset(n)]

    def backward(self, g7: torch.Tensor, g14: torch.Tensor, zero_grads: bool = True, after_stage=None, after_unit=None,
                 graph: bool = False):
        """Run the backward plan for dL/dfc7 and dL/dfc14; fills grads_flat (views in self.grads).
        ``after_stage()`` is called once the stage/head gradients (flat range ``stage_range``) are final and
        before the unit gradients are computed (used to overlap their all-reduce); ``after_unit(tag, stream)`` right
        after the last kernel that writes unit ``tag``'s gradients (flat range ``unit_ranges[tag]``) was issued on
        ``stream``."""
        self.d_out7.copy_(g7.reshape(self.d_out7.shape))
        self.d_out14.copy_(g14.reshape(self.d_out14.shape))
        if graph:
            if not zero_grads or after_stage is not None or after_unit is not None:
                raise RuntimeError("graph replay of the backward pass zeroes the gradient buffers and takes no hooks")
            self._zero_grads = True
            self._replay("bwd")
            return self.grads
        self._zero_grads = bool(zero_grads)
        streams = self._fork()
        n_stage = len(self.bwd_stage_steps)
        self.bwd_sched.run(streams, 0, n_stage)
        if after_stage is not None:
            self._join(streams)                      # the caller's stream now trails every stage-gradient kernel
            after_stage()
        on_step = None
        if after_unit is not None:
            def on_step(step, stream):
                tag = getattr(step, "unit_tag", None)
                if tag is not None:
                    after_unit(tag, stream)
        self.bwd_sched.run(streams, n_stage, on_step=on_step)
        self._join(streams)
        return self.grads

    def _fork(self):
        """[caller's stream, side streams...]; the side lanes start behind everything already enqueued by the caller."""
        main = torch.cuda.current_stream(self.device)
        if self.single_stream or self._lanes == 1:
            return [main] * self._lanes
        if self._side is None:
            self._side = [torch.cuda.Stream(device=self.device) for _ in range(self._lanes - 1)]
        for s_ in self._side:
            s_.wait_stream(main)
        return [main] + self._side

    def _join(self, streams):
        for s_ in streams[1:]:
            if s_ is not streams[0]:
                streams[0].wait_stream(s_)


def _auto_tile_n(M: int, N: int, tma: bool = False, x3: bool = False) -> int:
    """N tile of a GEMM launched without split-K; 0 = the kernel's default (widest legal tile).
    TMA-fed: two CTAs fit an SM and one CTA's epilogue should overlap another's main loop, so aim for >= 2 x 148 CTAs
    with the widest tile that gets there (re-reading A per N tile is one cheap TMA instruction).
    Gather-fed: the A gather is the expensive part and is repeated per N tile, so only narrow the tile when the grid
    would otherwise leave most SMs idle (the 7x7-resolution layers: 37 M tiles)."""
    mt = math.ceil(M / 128)
    bn0 = (N + 15) // 16 * 16 if N <= 256 else 256
    want, enough = (120, 140)
    if tma:         # (2 x 148 before the MMA issue path was fixed: narrow tiles were issue-bound then; profiles/env_r03t.log)
        want = enough = int(os.environ.get("OFFK_TF32_WANT_CTAS", str(_SM_TARGET)))
    if tma and x3:
        # 3xTF32: one CTA per SM, and every extra N tile repeats the A tile AND its residual pass (motion_conv_trans_28,
        # 147 x 1 tiles of 64: two tiles of 32 took 586 us against 371 us; unit_5a, 56 M tiles: five tiles of 32 columns
        # 38 us against ~20 for one of 160).  Only the 37-tile 7x7 layers are narrowed (sweep: profiles/env_r03s.log)
        want = enough = int(os.environ.get("OFFK_X3_WANT_CTAS", "40"))
    if mt * math.ceil(N / bn0) >= want:
        return 0
    best = 0
    for bn in (128, 64, 32):
        if bn < bn0 and N % bn == 0:
            best = bn
            if mt * (N // bn) >= enough:
                break
    return best


def _is_memset(name):
    return name.endswith(".zero") or name.startswith("zero ")


def _nm(fn, name, reads=(), writes=(), lane=0):
    """Annotate a plan step: launch name, the buffers it reads / writes (for cross-stream hazard analysis), its lane."""
    fn.launches = [name]
    fn.reads, fn.writes, fn.lane = list(reads), list(writes), lane
    return fn


def _on(step, lane):
    step.lane = lane
    return step


def _span(t):
    return (t.data_ptr(), t.data_ptr() + t.numel() * t.element_size())


class Schedule:
    """A fixed list of steps spread over a few CUDA streams ("lanes").  Steps on one lane run in issue order; a step
    that touches a buffer last written (or still being read) by a step on another lane waits on an event recorded
    behind that step.  The hazards (RAW / WAW / WAR on byte ranges) are derived from the steps' declared reads and
    writes, so lane assignment is purely a performance choice: any assignment is correct."""

    def __init__(self, steps, ignore=()):
        self.steps = list(steps)
        n = len(self.steps)
        self.n_lanes = 1 + max((getattr(st, "lane", 0) for st in self.steps), default=0)
        ign = [_span(t) for t in ignore]

        def spans(ts):
            out = []
            for t in ts:
                if t is None:
                    continue
                sp = _span(t)
                if any(lo <= sp[0] and sp[1] <= hi for lo, hi in ign):
                    continue                      # read-only inputs (parameters, taps): never a hazard
                out.append(sp)
            return out

        rd = [spans(getattr(st, "reads", ())) for st in self.steps]
        wr = [spans(getattr(st, "writes", ())) for st in self.steps]
        hit = lambda A, B: any(a[0] < b[1] and b[0] < a[1] for a in A for b in B)
        lanes = [getattr(st, "lane", 0) for st in self.steps]
        self.lanes = lanes
        covered = [[-1] * self.n_lanes for _ in range(self.n_lanes)]   # covered[a][b]: latest step of lane b that a awaited
        self.waits = [[] for _ in range(n)]
        record = set()
        for i in range(n):
            need = {}
            for j in range(i):
                if lanes[j] == lanes[i] or j <= covered[lanes[i]][lanes[j]]:
                    continue
                if self.steps[j] in getattr(self.steps[i], "independent_of", ()):
                    continue                      # declared to touch disjoint elements of the same buffers
                if hit(wr[j], rd[i]) or hit(wr[j], wr[i]) or hit(rd[j], wr[i]):
                    need[lanes[j]] = j
            for lane_j, j in need.items():
                self.waits[i].append(j)
                covered[lanes[i]][lane_j] = j
                record.add(j)
        self.record = record
        self.events = {j: torch.cuda.Event() for j in record}

    def run(self, streams, lo=0, hi=None, on_step=None):
        """Issue steps [lo, hi) on ``streams`` (torch.cuda.Stream per lane; lane 0 = the caller's stream).
        ``on_step(step, stream)`` is called after each step has been issued."""
        hi = len(self.steps) if hi is None else hi
        handles = [C.c_void_p(st.cuda_stream) for st in streams]
        for i in range(lo, hi):
            lane = self.lanes[i]
            for j in self.waits[i]:
                streams[lane].wait_event(self.events[j])
            self.steps[i](handles[lane])
            if i in self.record:
                self.events[i].record(streams[lane])
            if on_step is not None:
                on_step(self.steps[i], streams[lane])


def _names(step):
    ls = getattr(step, "launches", None)
    if ls is not None:
        return list(ls)
    return [getattr(step, "name", None) or getattr(step, "__name__", "step")]


def _gkey(g: T.ConvGeom):
    return (g.n_img, g.cin, g.hin, g.win, g.cout, g.kh, g.kw, g.stride, g.pad, g.x_ctot, g.x_coff, g.y_ctot, g.y_coff)
