"""Shared helpers of the GPU parity tests."""
import off_oracle as O


def engine_gates(eng):
    """The engine's own ReLU sign patterns (site -> bool NCHW tensor on the CPU), keyed like off_oracle.GATE_SITES: the
    post-activation buffers are channels-last."""
    from off_b200 import spec as S
    nchw = lambda t: (t > 0).permute(0, 3, 1, 2).cpu()
    b = eng.buf
    g = {"gen_" + t: nchw(b["gd_" + t][..., :S.GEN_C]) for t in S.LEVELS}
    g["t28"] = nchw(b["t28"])
    for k in ("h1_28a", "h2_28a", "s28a", "h1_28b", "h2_28b", "s28b", "h1_28c", "h2_28c", "t14", "h1_14a", "h2_14a", "s14a",
              "h1_14b", "h2_14b", "h3_14b", "t7", "h1_7", "h2_7"):
        g[k] = nchw(b[k])
    g["s28c"] = nchw(b["F14"][..., 800:1056])
    g["s14b"] = nchw(b["F7"][..., 320:832])
    assert set(g) == set(O.GATE_SITES)
    return g
