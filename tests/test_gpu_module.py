"""GPU: nn.Module surface, state_dict contract, autograd semantics, mirrors of basic_ops / util, full-size properties."""
import os

import numpy as np
import pytest
import torch

import off_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def dev():
    import off_b200  # noqa: F401
    return torch.device("cuda")


@pytest.mark.parametrize("variant", ["rgb", "flow"])
def test_state_dict_round_trip(dev, variant):
    from off_b200.modules import OFFSubNetwork
    net = OFFSubNetwork(1, 3, variant, device=dev)
    want = {}
    for line in open(os.path.join(GOLD, f"state_dict_keys_{variant}.txt")):
        f = line.split()
        want[f[0]] = tuple(int(x) for x in f[1:])
    sd = net.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == want          # keys AND shapes of the reference
    assert all("motion" in n for n, p in net.named_parameters() if p.requires_grad)   # train_off.py:40 filter works
    other = OFFSubNetwork(1, 3, variant, device=dev)
    other.load_state_dict(sd)
    assert torch.equal(other.engine.params_flat, net.engine.params_flat)
    # DataParallel-style checkpoints carry a "module." prefix (test_flow_off.py:52-58)
    other.load_state_dict({k[len("module."):]: v for k, v in {"module." + k: v for k, v in sd.items()}.items()})


def test_autograd_accumulates_like_torch(dev):
    from off_b200.modules import OFFSubNetwork
    B, Lg = 1, 3
    net = OFFSubNetwork(B, Lg, "rgb", precision="fp32", device=dev).eval()
    taps = {k: v.to(dev) for k, v in O.make_taps(3, B, Lg).items()}
    fc7, _, fc14 = net(taps)
    (fc7.sum() + fc14.sum()).backward()
    g1 = net.motion_conv_trans.weight.grad.clone()
    fc7, _, fc14 = net(taps)
    (fc7.sum() + fc14.sum()).backward()                     # second backward without zero_grad: gradients add up
    assert torch.allclose(net.motion_conv_trans.weight.grad, 2 * g1, rtol=1e-4, atol=1e-6)
    net.zero_grad(set_to_none=True)
    fc7, _, fc14 = net(taps)
    (fc7.sum() + fc14.sum()).backward()
    assert torch.allclose(net.motion_conv_trans.weight.grad, g1, rtol=1e-4, atol=1e-6)
    # fc_action_motion_28 is not part of any returned output (RGB_OFF.py:787,860): no gradient at all, as in the reference
    assert net.fc_action_motion_28.weight.grad is None and net.fc_action_motion_28.bias.grad is None


def test_backward_must_follow_its_own_forward(dev):
    """the engine keeps one set of activation buffers, so a backward whose forward has been
    overwritten by a later forward raises instead of silently using the wrong activations."""
    from off_b200.modules import OFFSubNetwork
    net = OFFSubNetwork(1, 3, "rgb", precision="tf32", device=dev).eval()
    taps = {k: v.to(dev) for k, v in O.make_taps(3, 1, 3).items()}
    fc7_a, _, _ = net(taps)
    fc7_b, _, _ = net(taps)
    with pytest.raises(RuntimeError, match="another forward"):
        fc7_a.sum().backward()
    fc7_b.sum().backward()                                   # the latest forward is fine
    assert net.motion_conv_trans.weight.grad is not None


def test_tap_gradients_when_requested(dev):
    from off_b200.modules import OFFSubNetwork
    B, Lg = 1, 3
    taps = O.make_taps(9, B, Lg)
    prm = O.make_params(9, "rgb")
    net = OFFSubNetwork(B, Lg, "rgb", precision="fp32", tap_grads=True, device=dev).eval()
    net.load_state_dict(prm)
    tg = {k: v.to(dev).requires_grad_(True) for k, v in taps.items()}
    fc7, _, fc14 = net(tg)
    (fc7.sum() + fc14.sum()).backward()
    # the oracle follows the engine's ReLU gates (helpers.engine_gates): arithmetic only, no gate flips
    from helpers import engine_gates
    _, _, tref = O.off_forward_backward(taps, prm, B, Lg, "rgb", None, torch.float64, tap_grads=True,
                                        gates=engine_gates(net.engine))
    for k in taps:
        err = (tg[k].grad.double().cpu() - tref[k]).norm() / tref[k].norm()
        assert err < 2e-5, (k, float(err))


def test_consensus_module(dev):
    """basic_ops.ConsensusModule: mean over segments and its expand/T backward."""
    from off_b200.basic_ops import ConsensusModule, Identity
    x = torch.randn(5, 6, 101, device=dev, requires_grad=True)
    y = ConsensusModule("avg", dim=1)(x)
    assert y.shape == (5, 1, 101) and torch.allclose(y, x.mean(1, keepdim=True), atol=1e-6)
    g = torch.randn_like(y)
    y.backward(g)
    assert torch.allclose(x.grad, g.expand(5, 6, 101) / 6.0, atol=1e-7)
    assert ConsensusModule("rnn")(x) is x and ConsensusModule("identity").consensus_type == "identity"
    assert Identity()(x) is x
    with pytest.raises(RuntimeError):
        ConsensusModule("avg")(x.cpu())                     # no CPU fallback


def test_sobel_modules_known_answers(dev):
    """util.SobelFilter / SobelFilter_Diagonal against the reference's own outputs (tests/golden/sobel_kat.npz)."""
    from off_b200.util import SobelFilter, SobelFilter_Diagonal
    fix = np.load(os.path.join(GOLD, "sobel_kat.npz"))
    xr = O.hash_normal(int(fix["rand_seed"]), (2, 4, 9, 7)).to(dev)
    gx, gy = SobelFilter(4, 4)(xr)
    assert torch.allclose(gx.cpu(), torch.from_numpy(fix["rand_gx"]), atol=2e-6)
    assert torch.allclose(gy.cpu(), torch.from_numpy(fix["rand_gy"]), atol=2e-6)
    assert torch.allclose(SobelFilter_Diagonal(4, 4)(xr).cpu(), torch.from_numpy(fix["rand_diag"]), atol=2e-6)
    # the reference's own smoke shape (util.py:101-108): SobelFilter(192,192) on [64,192,56,56] -> reduced batch here
    x = torch.rand(2, 192, 56, 56, device=dev, requires_grad=True)
    m = SobelFilter(192, 192).to(dev)
    ox, oy = m(x)
    rx, ry = O.sobel_xy(x.detach().cpu())
    assert torch.allclose(ox.cpu(), rx, atol=1e-5) and torch.allclose(oy.cpu(), ry, atol=1e-5)
    (ox.sum() + 2 * oy.sum()).backward()
    xc = x.detach().cpu().requires_grad_(True)
    a, b = O.sobel_xy(xc)
    (a.sum() + 2 * b.sum()).backward()
    assert torch.allclose(x.grad.cpu(), xc.grad, atol=1e-4)
    assert set(m.state_dict()) == {"conv1.weight", "conv2.weight"}


def test_full_size_properties(dev):
    """BASELINE config 2 size (48 clips x 3 segments): size-independent properties instead of an oracle run."""
    from off_b200.engine import OFFEngine
    B, Lg = 48, 3
    eng = OFFEngine(B, Lg, "rgb", dev, "tf32", index_mode="aligned")
    torch.manual_seed(0)
    with torch.no_grad():
        for n, v in eng.params.items():
            v.uniform_(-0.05, 0.05)
    for t in eng.taps.values():
        t.copy_(torch.relu(torch.randn_like(t)))
    a = [x.clone() for x in eng.forward(train=False)]
    b = [x.clone() for x in eng.forward(train=False)]
    # split-K layers accumulate with fp32 atomics: two runs agree to round-off of the tf32 operands downstream of
    # them (an fp32 ulp can move a tf32 rounding boundary), not bit-for-bit
    rep = [(x - y).abs().max().item() / x.abs().max().item() for x, y in zip(a, b)]
    assert max(rep) <= 2e-3, rep
    # clips are independent (aligned mode): reversing the clip order reverses the per-pair logits
    for t in eng.taps.values():
        t.copy_(t.view(B, Lg, *t.shape[1:]).flip(0).reshape(t.shape))
    c = [x.clone() for x in eng.forward(train=False)]
    for x, y in zip(a, c):
        xf = x.view(B, Lg - 1, -1).flip(0).reshape(x.shape)
        assert (xf - y).abs().max().item() <= 2e-3 * x.abs().max().item()
    # telescoping of the temporal channels in the 28x28 fusion buffer: sum_t T[b,t] = G[b,L-1] - G[b,0]
    F28 = eng.buf["F28"].view(B, Lg - 1, 28, 28, 320)
    G = eng.buf["gd_3a"].view(B, Lg, 28, 28, 160)[..., :128]
    assert (F28[..., 32:160].sum(1) - (G[:, -1] - G[:, 0])).abs().max().item() < 1e-4
    # seeded dropout: ~20 % kept, scaled by 5, identical mask regenerated for the same seed
    eng.forward(train=True, seed=7)
    s1 = eng.buf["F7"][..., :32].clone()
    eng.forward(train=True, seed=7)
    assert torch.equal(s1 != 0, eng.buf["F7"][..., :32] != 0)
    eng.forward(train=False)
    s0 = eng.buf["F7"][..., :32]
    kept = (s1 != 0).float().mean().item() / max((s0 != 0).float().mean().item(), 1e-9)
    assert 0.17 < kept < 0.23
    nz = s1 != 0
    assert torch.allclose(s1[nz], 5.0 * s0[nz], rtol=1e-5, atol=1e-6)
    # backward linearity in the upstream gradient
    g7, g14 = torch.randn(eng.P, 101, device=dev), torch.randn(eng.P, 101, device=dev)
    g1 = eng.backward(g7, g14)["motion_conv_trans.weight"].clone()
    g2 = eng.backward(2 * g7, 2 * g14)["motion_conv_trans.weight"].clone()
    assert (g2 - 2 * g1).norm().item() <= 1e-3 * g2.norm().item()
    assert eng.grads["fc_action_motion_28.weight"].abs().max().item() == 0


def test_bad_arguments_raise(dev):
    from off_b200 import _lib as L
    import ctypes as C
    lib = L.lib()
    s = L.OffkStencil()
    s.B, s.L, s.Cg, s.Cs, s.K, s.H, s.W = 1, 1, 128, 32, 1, 7, 7          # L = 1: no pairs
    with pytest.raises(RuntimeError, match="L >= 2"):
        L.check(lib.offk_stencil_diff_fwd(C.byref(s), None, None, None, None, None, None), "stencil")
    x = torch.zeros(4, device=dev)
    with pytest.raises(RuntimeError):
        L.check(lib.offk_avgpool_drop_fwd(x.data_ptr(), 1, 8, 49, 4, 0, 0, None, 0, None, 0.0, 1.0, x.data_ptr(), None), "pool")


def test_reference_model_classes(dev):
    """BNInception_OFF / bninception_off under the reference's module names: RGB_OFF_forward returns the per-pair tuple of
    RGB_OFF.py:860, Flow_OFF.forward the consensus tuple of Flow_OFF.py:884 or the modality_fuse sum (:881); the OFF
    outputs equal OFFSubNetwork's (same engine), the backbone score goes through the segment consensus."""
    from off_b200 import RGB_OFF, Flow_OFF
    from off_b200.modules import OFFSubNetwork
    B, Lg, NC = 2, 3, 101
    taps = {k: v.to(dev) for k, v in O.make_taps(4, B, Lg).items()}
    score = torch.randn(B * Lg, NC, device=dev)

    class Backbone(torch.nn.Module):
        def forward(self, x):
            return taps, score

    for mod, variant in ((RGB_OFF, "rgb"), (Flow_OFF, "flow")):
        prm = O.make_params(4, variant)
        ref = OFFSubNetwork(B, Lg, variant, device=dev).eval()
        ref.load_state_dict(prm, strict=False)                    # (the frozen Sobel taps are not in make_params)
        want7, _, want14 = [t.clone() for t in ref(taps)]
        m = mod.bninception_off(NC, B, Lg, backbone=Backbone(), device=dev).eval()
        m.load_state_dict(prm, strict=False)                      # strict=False: the stub backbone has no entries
        x = torch.zeros(B * Lg, 3, 8, 8, device=dev)
        if variant == "rgb":
            fc7, sc, fc14 = m.RGB_OFF_forward(x)
            assert fc7.shape == (B * (Lg - 1), NC) and sc is score
        else:
            fc7, sc, fc14 = m(x)
            assert fc7.shape == (B, NC)
            assert torch.allclose(sc, score.view(B, Lg, NC).mean(1), atol=1e-6)
        # same plan, same weights; split-K partial sums land in atomic order, hence not bit-for-bit
        assert torch.allclose(fc7, want7, rtol=1e-3, atol=1e-4) and torch.allclose(fc14, want14, rtol=1e-3, atol=1e-4)
        if variant == "flow":
            m.modality_fuse = True
            fused = m(x)
            assert torch.allclose(fused, want7 + score.view(B, Lg, NC).mean(1) + want14, rtol=1e-3, atol=1e-3)
    # without a backbone the taps are the input
    m = RGB_OFF.bninception_off(NC, B, Lg, device=dev).eval()
    fc7, sc, fc14 = m.RGB_OFF_forward(taps)
    assert sc is None and fc7.shape == (B * (Lg - 1), NC)


@pytest.mark.parametrize("variant,max_norm", [("rgb", 20.0), ("flow", 1e9)])
def test_fused_training_step_matches_torch(dev, variant, max_norm):
    """fused CE (labels repeated per pair, train_off.py:133-146) + clip_grad_norm (:149) + Adam (:72,151) on the
    flat buffers against F.cross_entropy + torch.nn.utils.clip_grad_norm_ + torch.optim.Adam on per-parameter tensors."""
    import torch.nn.functional as F
    from off_b200.modules import OFFSubNetwork
    from off_b200.train import FusedOFFTrainer
    B, Lg = 2, 3
    net = OFFSubNetwork(B, Lg, variant, precision="tf32", device=dev).eval()
    net.load_state_dict(O.make_params(4, variant), strict=(variant == "rgb"))
    eng = net.engine
    tr = FusedOFFTrainer(net, lr=1e-3, betas=(0.9, 0.99), weight_decay=5e-4, max_norm=max_norm)
    names = [n for n in eng.params if not n.startswith("fc_action_motion_28")]
    ref_p = [eng.params[n].detach().clone().requires_grad_(True) for n in names]
    fc28_before = eng.params["fc_action_motion_28.weight"].clone()
    opt = torch.optim.Adam(ref_p, lr=1e-3, betas=(0.9, 0.99), weight_decay=5e-4)
    target = torch.tensor([3, 77], device=dev)
    taps = {k: v.to(dev) for k, v in O.make_taps(4, B, Lg).items()}
    clipped = False
    for it in range(3):
        with torch.no_grad():
            fc7, _, fc14 = eng.forward(taps, train=False)
        loss = tr.loss_backward(fc7, fc14, target)
        # loss and dL/dlogits of the fused kernel
        t_rows = target.unsqueeze(1).repeat(1, tr.repeat).view(-1)               # train_off.py:133
        l7, l14 = fc7.clone().requires_grad_(True), fc14.clone().requires_grad_(True)
        want = F.cross_entropy(l7, t_rows) + F.cross_entropy(l14, t_rows)
        want.backward()
        assert abs(loss.item() - want.item()) < 1e-5 * max(1.0, abs(want.item()))
        assert torch.allclose(tr.g7, l7.grad, atol=1e-7, rtol=1e-5) and torch.allclose(tr.g14, l14.grad, atol=1e-7, rtol=1e-5)
        for p, n in zip(ref_p, names):
            p.grad = eng.grads[n].detach().clone()
        total = torch.nn.utils.clip_grad_norm_(ref_p, max_norm)
        clipped |= bool(total > max_norm)
        opt.step()
        tr.step()
        assert abs(tr.grad_norm() - total.item()) < 1e-5 * total.item()
        for p, n in zip(ref_p, names):
            assert torch.allclose(eng.params[n], p.detach(), atol=2e-7, rtol=1e-5), (it, n)
    # the RGB case (sum of two CE losses at random init: norm > 20) exercises the clipping branch, the other one the pass-through
    assert clipped == (variant == "rgb")
    assert torch.equal(eng.params["fc_action_motion_28.weight"], fc28_before)   # no gradient in the reference: Adam skips it


@pytest.mark.parametrize("B,Lg,precision", [(2, 3, "fp32"), (48, 3, "tf32"), (1, 3, "tf32")])
def test_cuda_graph_replay_is_bit_identical_to_eager_issue(dev, B, Lg, precision):
    """SURVEY 8f-1 / offk.h "safe under CUDA-graph capture": forward and backward captured as one CUDA graph each (all
    lanes, programmatic dependent launch included) replay the very kernels of the eager plan.  Everything that is not an
    atomic accumulation must be bit-identical; split-K outputs and weight gradients are fp32 atomics (order-dependent in the
    last bits) in BOTH modes, so they are compared against the run-to-run spread of the eager path itself."""
    from off_b200.engine import OFFEngine
    eng = OFFEngine(B, Lg, "rgb", dev, precision)
    torch.manual_seed(0)
    with torch.no_grad():
        for n, v in eng.params.items():
            v.uniform_(-0.05, 0.05)
    for t in eng.taps.values():
        t.copy_(torch.relu(torch.randn_like(t)))
    g7, g14 = torch.randn(eng.P, 101, device=dev), torch.randn(eng.P, 101, device=dev)

    def run(graph, seed):
        out = [x.clone() for x in eng.forward(train=True, seed=seed, graph=graph)]
        f28 = eng.buf["F28"].clone()
        grads = eng.backward(g7, g14, graph=graph)
        return out, f28, eng.grads_flat.clone()

    e1, e2 = run(False, 11), run(False, 11)
    g1, g2 = run(True, 11), run(True, 11)            # first call captures, second replays
    g3 = run(True, 12)                                # another seed through the SAME graph: new dropout masks
    torch.cuda.synchronize()
    assert torch.equal(e1[1], g1[1]) and torch.equal(g1[1], g2[1])          # unit GEMM + stencil: no atomics -> bit-exact
    assert not torch.equal(g3[1], g1[1])                                      # the device-resident seed reached the kernels
    spread = lambda a, b: max((x - y).abs().max().item() / max(x.abs().max().item(), 1e-30) for x, y in zip(a, b))
    tol = 10 * max(spread(e1[0], e2[0]), 1e-6)
    assert spread(e1[0], g1[0]) <= tol and spread(g1[0], g2[0]) <= tol
    gs = lambda a, b: ((a - b).norm() / a.norm()).item()
    assert gs(e1[2], g2[2]) <= 10 * max(gs(e1[2], e2[2]), 1e-6)
    assert eng.grads["fc_action_motion_28.weight"].abs().max().item() == 0
    # the nn.Module surface with use_graphs=True trains like the eager one
    from off_b200.modules import OFFSubNetwork
    if B == 2:
        prm = O.make_params(3, "rgb")
        taps = {k: v.to(dev) for k, v in O.make_taps(3, B, Lg).items()}
        outs = []
        for use in (False, True):
            net = OFFSubNetwork(B, Lg, "rgb", precision=precision, device=dev, use_graphs=use).eval()
            net.load_state_dict(prm)
            for _ in range(2):
                net.zero_grad(set_to_none=True)
                fc7, _, fc14 = net(taps)
                (fc7.sum() + fc14.sum()).backward()
            outs.append((fc7.detach().clone(), net.motion_conv_trans_28.weight.grad.clone()))
        assert _close(outs[0][0], outs[1][0], 1e-5) and _close(outs[0][1], outs[1][1], 1e-4)


def _close(a, b, rel):
    return ((a - b).norm() / a.norm()).item() <= rel


def test_eval_protocol_at_the_reference_shape(dev):
    """one video through the reference's evaluation shape -- 10 crops as the batch axis, 25 segments
    (test_rgb_off.py:24-25,184: 250 frames, 240 frame pairs) -- on the Flow / RGB_OFF_v2 surface (consensus over the 24 pairs
    inside the model), fused as mean_crops(rst1) + 2*mean_crops(rst2) + mean_crops(rst3); against the oracle's forward."""
    from off_b200 import Flow_OFF, evaluate as EV
    crops, segs = 10, 25
    model = Flow_OFF.bninception_off(101, crops, segs, device=dev, precision="fp32")
    prm = O.make_params(6, "flow")
    model.load_state_dict(prm, strict=False)
    g = torch.Generator(device=dev).manual_seed(6)
    taps = {t: torch.relu(torch.randn(crops * segs, cin, s, s, device=dev, generator=g)) for t, (cin, s) in O.LEVELS.items()}
    fused, r1, r2, r3 = EV.eval_video(model, taps, num_crops=crops)
    assert r1.shape == (crops, 101) and r3.shape == (crops, 101) and r2 is None and fused.shape == (1, 101)
    with torch.no_grad():
        ref = O.off_forward({k: v.double().cpu() for k, v in taps.items()}, {k: v.double() for k, v in prm.items()}, crops, segs, "flow")
    assert _close(torch.from_numpy(r1).double(), ref["fc7"], 1e-5) and _close(torch.from_numpy(r3).double(), ref["fc14"], 1e-5)
    want = ref["fc7"].mean(0) + ref["fc14"].mean(0)
    assert _close(torch.from_numpy(fused[0]), want, 1e-5)
    assert not model.training or model.off.training            # eval_video restored the caller's mode


@pytest.mark.parametrize("variant", ["rgb", "flow"])
def test_full_model_is_a_drop_in_of_the_reference_class(dev, variant):
    """Images in, the reference's return tuple out: BNInception_OFF(backbone='bninception') -- the built-in
    feature extractor producing the taps on the device + the OFF section on liboffk -- against the reference class itself run
    in fp64 on the same seeded FULL state_dict and images (tests/golden/backbone_*.npz, written by oracle/make_golden.py).
    RGB_OFF.RGB_OFF_forward -> (fc7 [P,101], score [N,101], fc14 [P,101]); Flow_OFF.forward -> consensus outputs [B,101]."""
    import numpy as np
    from off_b200 import RGB_OFF, Flow_OFF
    fix = np.load(os.path.join(GOLD, f"backbone_{variant}_b1_l3.npz"))
    B, Lg, seed = int(fix["batch"]), int(fix["length"]), int(fix["seed"])
    keys = [(l.split()[0], tuple(int(x) for x in l.split()[1:])) for l in open(os.path.join(GOLD, f"state_dict_keys_full_{variant}.txt"))]
    mod = RGB_OFF if variant == "rgb" else Flow_OFF
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        m = mod.bninception_off(101, B, Lg, device=dev, backbone="bninception", precision="fp32").eval()
        m.load_state_dict(O.make_state_dict(seed, keys), strict=True)
        x = O.hash_normal(seed, (B * Lg, 10 if variant == "flow" else 3, 224, 224)).to(dev)
        with torch.no_grad():
            out = m.RGB_OFF_forward(x) if variant == "rgb" else m(x)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    for got, key in zip(out[:3], ("fc7", "score", "fc14")):
        want = torch.from_numpy(fix[key])
        assert tuple(got.shape) == tuple(want.shape), key
        err = (got.double().cpu() - want).abs().max().item() / want.abs().max().item()
        assert err < 1e-4, (key, err)
