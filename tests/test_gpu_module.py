"""GPU parity of the fused inner functions and the TFM-interface Mamba module against
(a) golden vectors produced by the unmodified reference module / mamba_inner_ref (oracle/gen_golden.py) and
(b) the torch oracle (oracle/torch_ref.py) run on CPU with the same parameters."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import torch_ref

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from mmunet_b200 import Mamba, ops
DEV = "cuda"


def close(name, got, ref, rtol=2e-3, atol=2e-3):
    got = got.detach().float().cpu().numpy().astype(np.float64)
    ref = np.asarray(ref.detach().cpu().numpy() if torch.is_tensor(ref) else ref, np.float64)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    atol = atol * max(1.0, float(np.abs(ref).max()))
    err = np.abs(got - ref)
    worst = float((err / (atol + rtol * np.abs(ref))).max())
    assert np.isfinite(got).all() and worst <= 1.0, f"{name}: max|err|={err.max():.3e} worst/bound={worst:.2f}"


def test_mamba_inner_fn_golden():
    c = np.load(os.path.join(GOLDEN, "mamba_inner.npz"))
    keys = ("xz", "conv_w", "conv_b", "x_proj_w", "dt_proj_w", "out_proj_w", "A", "D", "dt_bias")
    t = {k: torch.tensor(c[k], device=DEV, requires_grad=True) for k in keys}
    out = ops.mamba_inner_fn(t["xz"], t["conv_w"], t["conv_b"], t["x_proj_w"], t["dt_proj_w"], t["out_proj_w"], None,
                             t["A"], None, None, t["D"], t["dt_bias"], delta_softplus=True)
    close("out", out, c["out"])                       # the only assert the reference's own test makes (:221)
    out.backward(torch.tensor(c["dout"], device=DEV))
    for k in keys:
        close("d" + k, t[k].grad, c["d" + k], 5e-3, 5e-3)


@pytest.mark.parametrize("batch,d,L,N,R", [(2, 8, 48, 4, 2), (3, 6, 200, 16, 1), (1, 128, 256, 16, 4)])
@pytest.mark.parametrize("reverse", [False, True])
def test_inner_no_out_proj_vs_torch_oracle(batch, d, L, N, R, reverse):
    g = torch.Generator().manual_seed(1)
    p = dict(xz=torch.randn(batch, 2 * d, L, generator=g), conv_w=torch.randn(d, 1, 4, generator=g) * 0.5,
             conv_b=torch.randn(d, generator=g) * 0.5, x_proj_w=torch.randn(R + 2 * N, d, generator=g) * 0.3,
             dt_proj_w=torch.randn(d, R, generator=g) * 0.3, A=-0.5 * torch.rand(d, N, generator=g),
             D=torch.randn(d, generator=g), dt_bias=0.5 * torch.rand(d, generator=g))
    dout = torch.randn(batch, d, L, generator=g)
    cpu = {k: v.clone().requires_grad_() for k, v in p.items()}
    xz_in = cpu["xz"].flip(-1) if reverse else cpu["xz"]
    ref = torch_ref.mamba_inner(xz_in, cpu["conv_w"], cpu["conv_b"], cpu["x_proj_w"], cpu["dt_proj_w"], cpu["A"],
                                cpu["D"], cpu["dt_bias"], with_out_proj=False)
    if reverse:
        ref = ref.flip(-1)
    ref.backward(dout)
    gpu = {k: v.to(DEV).requires_grad_() for k, v in p.items()}
    # give xz the (l, b*l, 1) strides in_proj produces
    xz = gpu["xz"].detach().transpose(0, 1).contiguous().transpose(0, 1).requires_grad_()
    if reverse:
        out = ops.mamba_inner_fn_no_out_proj_reversed(xz, gpu["conv_w"], gpu["conv_b"], gpu["x_proj_w"], gpu["dt_proj_w"],
                                                      gpu["A"], gpu["D"], gpu["dt_bias"])
    else:
        out = ops.mamba_inner_fn_no_out_proj(xz, gpu["conv_w"], gpu["conv_b"], gpu["x_proj_w"], gpu["dt_proj_w"],
                                             gpu["A"], None, None, gpu["D"], gpu["dt_bias"])
    close("out", out, ref)
    out.backward(dout.to(DEV))
    close("dxz", xz.grad, cpu["xz"].grad, 5e-3, 5e-3)
    for k in ("conv_w", "conv_b", "x_proj_w", "dt_proj_w", "A", "D", "dt_bias"):
        close("d" + k, gpu[k].grad, cpu[k].grad, 5e-3, 5e-3)


@pytest.mark.parametrize("batch,d,L,N,R,e", [(2, 8, 48, 4, 2, 4), (3, 6, 203, 16, 1, 3), (1, 128, 256, 16, 4, 64)])
def test_bimamba_inner_fn_vs_torch_oracle(batch, d, L, N, R, e):
    """bimamba_inner_fn (selective_scan_interface.py:437-603, 616-624) against bimamba_inner_ref's formulation
    (:673-709, oracle/torch_ref.py:bimamba_inner): out = out_proj(scan(A) + flip(scan_flipped(A_b))), one conv / projection
    shared by both directions.  Forward and every gradient, including dA_b, dD, ddelta_bias and the out_proj pair; ragged L."""
    g = torch.Generator().manual_seed(11)
    p = dict(xz=torch.randn(batch, 2 * d, L, generator=g), conv_w=torch.randn(d, 1, 4, generator=g) * 0.5,
             conv_b=torch.randn(d, generator=g) * 0.5, x_proj_w=torch.randn(R + 2 * N, d, generator=g) * 0.3,
             dt_proj_w=torch.randn(d, R, generator=g) * 0.3, out_proj_w=torch.randn(e, d, generator=g) * 0.3,
             out_proj_b=torch.randn(e, generator=g) * 0.3, A=-0.5 * torch.rand(d, N, generator=g),
             A_b=-0.5 * torch.rand(d, N, generator=g), D=torch.randn(d, generator=g), dt_bias=0.5 * torch.rand(d, generator=g))
    dout = torch.randn(batch, L, e, generator=g)
    cpu = {k: v.clone().requires_grad_() for k, v in p.items()}
    ref = torch_ref.bimamba_inner(cpu["xz"], cpu["conv_w"], cpu["conv_b"], cpu["x_proj_w"], cpu["dt_proj_w"], cpu["out_proj_w"],
                                  cpu["out_proj_b"], cpu["A"], cpu["A_b"], cpu["D"], cpu["dt_bias"])
    ref.backward(dout)
    gpu = {k: v.to(DEV).requires_grad_() for k, v in p.items()}
    xz = gpu["xz"].detach().transpose(0, 1).contiguous().transpose(0, 1).requires_grad_()      # (l, b*l, 1) strides, as in_proj gives
    out = ops.bimamba_inner_fn(xz, gpu["conv_w"], gpu["conv_b"], gpu["x_proj_w"], gpu["dt_proj_w"], gpu["out_proj_w"],
                               gpu["out_proj_b"], gpu["A"], gpu["A_b"], None, None, gpu["D"], gpu["dt_bias"])
    close("out", out, ref)
    out.backward(dout.to(DEV))
    close("dxz", xz.grad, cpu["xz"].grad, 5e-3, 5e-3)
    for k in ("conv_w", "conv_b", "x_proj_w", "dt_proj_w", "out_proj_w", "out_proj_b", "A", "A_b", "D", "dt_bias"):
        close("d" + k, gpu[k].grad, cpu[k].grad, 5e-3, 5e-3)


def test_bimamba_inner_fn_bf16_autocast():
    g = torch.Generator().manual_seed(12)
    batch, d, L, N, R, e = 2, 64, 520, 16, 4, 32
    p = dict(xz=torch.randn(batch, 2 * d, L, generator=g), conv_w=torch.randn(d, 1, 4, generator=g) * 0.5,
             conv_b=torch.randn(d, generator=g) * 0.5, x_proj_w=torch.randn(R + 2 * N, d, generator=g) * 0.1,
             dt_proj_w=torch.randn(d, R, generator=g) * 0.3, out_proj_w=torch.randn(e, d, generator=g) * 0.1,
             A=-0.5 * torch.rand(d, N, generator=g), A_b=-0.5 * torch.rand(d, N, generator=g), D=torch.randn(d, generator=g),
             dt_bias=0.5 * torch.rand(d, generator=g))
    ref = torch_ref.bimamba_inner(p["xz"], p["conv_w"], p["conv_b"], p["x_proj_w"], p["dt_proj_w"], p["out_proj_w"], None,
                                  p["A"], p["A_b"], p["D"], p["dt_bias"])
    gpu = {k: v.to(DEV).requires_grad_() for k, v in p.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = ops.bimamba_inner_fn(gpu["xz"].to(torch.bfloat16), gpu["conv_w"], gpu["conv_b"], gpu["x_proj_w"], gpu["dt_proj_w"],
                                   gpu["out_proj_w"], None, gpu["A"], gpu["A_b"], None, None, gpu["D"], gpu["dt_bias"])
    assert out.dtype == torch.bfloat16
    close("out.bf16", out, ref, 3e-2, 5e-2)
    out.float().sum().backward()
    assert all(gpu[k].grad is not None and torch.isfinite(gpu[k].grad).all() for k in gpu if k != "xz")


def _load_golden_module(bimamba_type):
    c = np.load(os.path.join(GOLDEN, "tfm_mamba.npz"))
    m = Mamba(d_model=8, d_state=4, d_conv=4, expand=2, bimamba_type=bimamba_type, nslices=4)
    sd = {k[len("param."):]: torch.tensor(c[k]) for k in c.files if k.startswith("param.")}
    assert set(sd) == set(m.state_dict()), "state_dict names must match the reference module (SURVEY.md section 5)"
    m.load_state_dict(sd)
    return c, m.to(DEV)


def test_tfm_mamba_v3_golden():
    c, m = _load_golden_module("v3")
    x = torch.tensor(c["x"], device=DEV, requires_grad=True)
    out, o1, o2, o3 = m(x)
    for name, got in (("out", out), ("o1", o1), ("o2", o2), ("o3", o3)):
        close(name, got, c[name])
    out.backward(torch.tensor(c["dout"], device=DEV))
    close("dx", x.grad, c["dx"], 5e-3, 5e-3)
    for k, p in m.named_parameters():
        close("grad." + k, p.grad, c["grad." + k], 5e-3, 5e-3)


def test_tfm_mamba_v1_golden():
    """bimamba_type="v1": mamba_simple.py:304-318 evaluated with the same parameters (the reference module itself
    raises UnboundLocalError there - recorded in the fixture)."""
    c, m = _load_golden_module("v1")
    assert str(c["v1_status"]) == "UnboundLocalError"
    x = torch.tensor(c["x"], device=DEV, requires_grad=True)
    out, o1, o2, o3 = m(x)
    assert o1 is None and o2 is None and o3 is None
    close("v1.out", out, c["v1.out"])
    out.backward(torch.tensor(c["dout"], device=DEV))
    close("v1.dx", x.grad, c["v1.dx"], 5e-3, 5e-3)
    for k, p in m.named_parameters():
        ref = c["v1.grad." + k]
        if ref.size == 0:
            assert p.grad is None, k            # _b / _s sets are unused by v1
        else:
            close("v1.grad." + k, p.grad, ref, 5e-3, 5e-3)


@pytest.mark.parametrize("bimamba_type", ["v1", "v2", "v3"])
@pytest.mark.parametrize("d_model,L,ns", [(3, 256, 8), (64, 512, 16)])
def test_tfm_mamba_vs_torch_oracle(bimamba_type, d_model, L, ns):
    torch.manual_seed(3)
    m = Mamba(d_model=d_model, d_state=16, d_conv=4, expand=2, bimamba_type=bimamba_type, nslices=ns)
    x = torch.randn(2, L, d_model)
    xc = x.clone().requires_grad_()
    ref, r1, r2, r3 = torch_ref.mamba_forward(m, xc)
    dout = torch.randn_like(ref)
    ref.backward(dout)
    ref_grads = {k: (None if p.grad is None else p.grad.clone()) for k, p in m.named_parameters()}
    m.zero_grad()
    m = m.to(DEV)
    xg = x.to(DEV).requires_grad_()
    out, o1, o2, o3 = m(xg)
    close("out", out, ref)
    if bimamba_type == "v3":
        close("o1", o1, r1), close("o2", o2, r2), close("o3", o3, r3)
    out.backward(dout.to(DEV))
    close("dx", xg.grad, xc.grad, 5e-3, 5e-3)
    for k, p in m.named_parameters():
        if ref_grads[k] is None:
            assert p.grad is None, k
        else:
            close("grad." + k, p.grad, ref_grads[k], 5e-3, 5e-3)


@pytest.mark.parametrize("bimamba_type,d_model,L,ns,order", [("v3", 64, 512, 16, None), ("v3", 8, 1024, 64, None),
                                                            ("v1", 3, 48, 1, (3, 6, 8, 1)), ("v1", 3, 1024, 1, (3, 32, 32, 1)),
                                                            ("v1", 1, 160, 1, (3, 10, 16, 1))])
@pytest.mark.parametrize("autocast", [False, True])
def test_mamba_fused_scan_order_equals_explicit(bimamba_type, d_model, L, ns, order, autocast, monkeypatch):
    """The module with the scan order fused into the kernels (v3's nslices direction; MMConv's two-row order through the
    scan_order keyword) against the same module with MMU_FUSE=0 (explicit gather / scatter kernels): same outputs, same grads."""
    from mmunet_b200 import _lib
    monkeypatch.setenv("MMU_NARROW", "0")       # this test is about the generic inner functions (d_model = 3 has its own fused block)
    torch.manual_seed(9)
    m = Mamba(d_model=d_model, d_state=16, d_conv=4, expand=2, bimamba_type=bimamba_type, nslices=ns).to(DEV)
    x = torch.randn(2, L, d_model, device=DEV)
    dout = torch.randn(2, L, d_model, device=DEV)

    def run():
        m.zero_grad()
        xi = x.clone().requires_grad_()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = m(xi, scan_order=order)[0] if order is not None else m(xi)[0]
        out.float().backward(dout)
        return out.float().detach(), xi.grad.clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}

    od = order if order is not None else (_lib.ORDER_NSLICES, 1, L, ns)
    fusable = ops.order_fusable(od, 16, torch.bfloat16 if autocast else torch.float32)
    n0 = _lib.launch_count()
    o1, dx1, g1 = run()
    fused_launches = _lib.launch_count() - n0
    monkeypatch.setenv("MMU_FUSE", "0")
    n0 = _lib.launch_count()
    o2, dx2, g2 = run()
    if fusable:
        assert _lib.launch_count() - n0 >= fused_launches + 4      # the explicit path adds gather + scatter launches, both passes
    else:
        assert _lib.launch_count() - n0 == fused_launches          # same explicit path either way
    tol = dict(rtol=2e-2, atol=2e-2) if autocast else dict(rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(o1, o2, **tol)
    torch.testing.assert_close(dx1, dx2, rtol=tol["rtol"], atol=tol["atol"] * float(dx2.abs().max()))
    assert g1.keys() == g2.keys()
    for k in g1:
        torch.testing.assert_close(g1[k], g2[k], rtol=5 * tol["rtol"], atol=5 * tol["atol"] * max(1.0, float(g2[k].abs().max())), msg=lambda s_: f"{k}: {s_}")


def test_tfm_mamba_bf16_autocast():
    torch.manual_seed(5)
    m = Mamba(d_model=64, d_state=16, bimamba_type="v3", nslices=16)
    x = torch.randn(2, 1024, 64)
    ref, *_ = torch_ref.mamba_forward(m, x)
    m = m.to(DEV)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out, *_ = m(x.to(DEV).requires_grad_())
    assert out.dtype == torch.bfloat16
    close("out.bf16", out, ref.detach(), 2e-2, 5e-2)
    out.float().sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())


def test_selective_scan_fn_3d_BC_and_slow_path():
    torch.manual_seed(7)
    m = Mamba(d_model=8, d_state=4, bimamba_type="v1", use_fast_path=False)
    m2 = Mamba(d_model=8, d_state=4, bimamba_type="v1")
    m2.load_state_dict(m.state_dict())
    x = torch.randn(2, 70, 8, device=DEV)
    a, *_ = m.to(DEV)(x)
    b, *_ = m2.to(DEV)(x)
    close("slow==fast", a, b.detach(), 1e-4, 1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,order_kind", [(16, 16, "tworow"), (7, 10, "tworow"), (33, 24, "tworow"), (1, 300, None), (20, 52, None)])
@pytest.mark.parametrize("autocast", [False, True])
def test_narrow_mamba_fused_block_equals_generic_path(H, W, order_kind, autocast, monkeypatch):
    """MMConv's d_model = 3 Mamba through the fused narrow kernels (prologue -> scan -> epilogue, row f3) against the same module
    with MMU_NARROW=0 (in_proj / conv / x_proj / dt_proj / out_proj as separate ops, which the golden tests pin to the reference):
    same output, same input gradient, same gradient for every parameter; far fewer launches.  Odd H (tail row of the two-row
    order), W not a multiple of 4 and L not a multiple of the 256-token tile are covered."""
    from mmunet_b200 import _lib
    torch.manual_seed(11)
    L = H * W
    m = Mamba(d_model=3, d_state=16, d_conv=4, expand=2, bimamba_type="v1", nslices=4).to(DEV)
    with torch.no_grad():
        for p_ in m.parameters():
            if p_.dim() > 1:
                p_.mul_(2.0)                   # away from the tiny default init, so every gradient path is exercised
    x = torch.randn(3, L, 3, device=DEV)
    dout = torch.randn(3, L, 3, device=DEV)
    order = (_lib.ORDER_TWOROW, H, W, 1) if order_kind == "tworow" else None

    def run(ac=autocast):
        m.zero_grad()
        xi = x.clone().requires_grad_()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            out = m(xi, scan_order=order)[0]
        out.float().backward(dout)
        return out.float().detach(), xi.grad.clone(), {k: p_.grad.clone() for k, p_ in m.named_parameters() if p_.grad is not None}

    n0 = _lib.launch_count()
    o1, dx1, g1 = run()
    fused_launches = _lib.launch_count() - n0
    monkeypatch.setenv("MMU_NARROW", "0")
    n0 = _lib.launch_count()
    o2, dx2, g2 = run()
    generic_launches = _lib.launch_count() - n0
    assert fused_launches <= generic_launches + 2          # library launches: conv fwd / bwd -> prologue, epilogue and their adjoints
    assert g1.keys() == g2.keys()
    if not autocast:
        tol = dict(rtol=2e-4, atol=2e-4)
        torch.testing.assert_close(o1, o2, rtol=tol["rtol"], atol=tol["atol"] * max(1.0, float(o2.abs().max())))
        torch.testing.assert_close(dx1, dx2, rtol=tol["rtol"], atol=tol["atol"] * max(1.0, float(dx2.abs().max())))
        for k in g1:
            torch.testing.assert_close(g1[k], g2[k], rtol=5 * tol["rtol"], atol=5 * tol["atol"] * max(1.0, float(g2[k].abs().max())),
                                       msg=lambda s_: f"{k}: {s_}")
        return
    # bf16 autocast: the two pipelines round at different places (the fused block keeps fp32 between its stages), so both are
    # measured against the fp32 run of the generic path: the fused block must be no further from it than the generic bf16 path
    # (factor 2 + 0.5 % of the tensor's scale for noise), and inside north_star's 2e-2 band (of the tensor's scale) in the mean
    o0, dx0, g0 = run(ac=False)

    def no_worse(name, a, b, ref):
        scale = float(ref.abs().max())
        ea, eb = float((a - ref).abs().max()), float((b - ref).abs().max())
        assert ea <= 2.0 * eb + 5e-3 * scale, f"{name}: fused err {ea:.3e} vs generic bf16 err {eb:.3e} (scale {scale:.3e})"
        assert float((a - ref).abs().mean()) <= 2e-2 * max(1.0, scale), name

    no_worse("out", o1, o2, o0)
    no_worse("dx", dx1, dx2, dx0)
    for k in g1:
        no_worse(k, g1[k], g2[k], g0[k])


@pytest.mark.parametrize("H,W", [(16, 16), (9, 20)])
@pytest.mark.parametrize("autocast", [False, True])
def test_narrow_mamba_coordinate_epilogue(H, W, autocast):
    """MMConv's coordinate arithmetic (MMUNet.py:156-188) as the epilogue of the fused narrow block against the same block followed
    by the torch ops: y = clamp(softplus(altho), 0.01) * refined + row + snake_offsets(dy) * extend_scope, with the gradients of the
    offset map, of altho and of every block parameter."""
    from mmunet_b200 import _lib
    from mmunet_b200.mm_net import MMConv
    torch.manual_seed(13)
    K, B, scope = 3, 2, 1.5
    m = Mamba(d_model=K, d_state=16, d_conv=4, expand=2, bimamba_type="v1", nslices=4).to(DEV)
    with torch.no_grad():
        for p_ in m.parameters():
            if p_.dim() > 1:
                p_.mul_(2.0)
    altho = torch.nn.Parameter(torch.tensor(0.3, device=DEV))
    dy = torch.tanh(torch.randn(B, K, H, W, device=DEV))
    g = torch.randn(B, K, H, W, device=DEV)
    order = (_lib.ORDER_TWOROW, H, W, 1)
    helper = MMConv(in_channels=4, out_channels=4, kernel_size=K)           # for _snake_offsets only

    def run(fused):
        m.zero_grad()
        altho.grad = None
        x = dy.clone().requires_grad_()
        tokens = x.reshape(B, K, H * W).transpose(-1, -2)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            if fused:
                y = m(tokens, scan_order=order, coord_epilogue=(altho, scope, H, W))[0].transpose(-1, -2).reshape(B, K, H, W)
            else:
                refined = m(tokens, scan_order=order)[0].transpose(-1, -2).reshape(B, K, H, W)
                rows = torch.arange(H, dtype=torch.float32, device=DEV).view(1, 1, H, 1)
                gain = torch.clamp(torch.nn.functional.softplus(altho), min=0.01)
                y = gain * refined.float() + (rows + helper._snake_offsets(x).float() * scope)
        assert y.dtype == torch.float32
        y.backward(g)
        return y.detach(), x.grad.clone(), altho.grad.clone(), {k: p_.grad.clone() for k, p_ in m.named_parameters() if p_.grad is not None}

    y1, dx1, da1, g1 = run(True)
    y2, dx2, da2, g2 = run(False)
    tol = 3e-2 if autocast else 2e-4
    torch.testing.assert_close(y1, y2, rtol=tol, atol=tol * max(1.0, float((y2 - torch.arange(H, device=DEV).view(1, 1, H, 1)).abs().max())))
    torch.testing.assert_close(dx1, dx2, rtol=tol, atol=tol * max(1.0, float(dx2.abs().max())))
    torch.testing.assert_close(da1, da2, rtol=5 * tol, atol=5 * tol * max(1.0, float(da2.abs())))
    for k in g1:
        torch.testing.assert_close(g1[k], g2[k], rtol=5 * tol, atol=5 * tol * max(1.0, float(g2[k].abs().max())), msg=lambda s_: f"{k}: {s_}")


@pytest.mark.parametrize("d_model,L,ns,heads", [(48, 512, 64, 1), (96, 256, 32, 2)])
def test_v3_directional_outputs_feed_attention_like_consumer(d_model, L, ns, heads):
    """sibling users of the same boundary consume the three directional outputs, not only `out` -
    HWAUNETR's MFABlock turns (o_1, o_2, o_3) into q / k / v of a channel attention (src/model/HWAUNETR.py:202-270; dims 48 / 96,
    num_slices 64 / 32).  A consumer of that shape (restated here, not the reference module) must get the oracle's values and the
    oracle's gradients THROUGH o_1..o_3 as well as through out."""
    torch.manual_seed(17)
    m = Mamba(d_model=d_model, d_state=16, d_conv=4, expand=2, bimamba_type="v3", nslices=ns)
    x = torch.randn(2, L, d_model)
    w = torch.randn(2, L, d_model)

    def consumer(out, q, k, v):
        q, k, v = q.unsqueeze(1), k.unsqueeze(1), v.unsqueeze(1)             # (b, 1, d_inner, l)
        attn = (q.transpose(-2, -1) @ k / q.shape[-2]).softmax(-1)           # (b, 1, l, l)
        out_a = (v @ attn.transpose(-2, -1))[:, 0, :d_model]                 # (b, d_model, l)
        return ((out + out_a.transpose(-1, -2)) * w.to(out.device)).sum()

    xc = x.clone().requires_grad_()
    consumer(*torch_ref.mamba_forward(m, xc)).backward()
    ref_dx = xc.grad.clone()
    ref_grads = {k: (None if p.grad is None else p.grad.clone()) for k, p in m.named_parameters()}
    m.zero_grad()
    m = m.to(DEV)
    xg = x.to(DEV).requires_grad_()
    consumer(*m(xg)).backward()
    close("dx", xg.grad, ref_dx, 5e-3, 5e-3)
    for k, p in m.named_parameters():
        if ref_grads[k] is None:
            assert p.grad is None, k
        else:
            close("grad." + k, p.grad, ref_grads[k], 5e-3, 5e-3)
