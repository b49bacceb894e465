"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

Run:  python oracle/gen_golden.py            (needs /root/reference; CPU only)

The reference cannot travel to the GPU box, so its outputs are committed as small fixtures.
What is executed (all from /root/reference, loaded by path, nothing copied):
  * selective_scan_ref, mamba_inner_ref          requirements/Mamba/mamba/mamba_ssm/ops/selective_scan_interface.py
  * causal_conv1d_ref                            requirements/Mamba/causal-conv1d/causal_conv1d/causal_conv1d_interface.py
  * TFM Mamba (v3 forward+backward)              requirements/mamba_simple.py
  * MMConv.two_row_columnwise_flatten_grad_safe / inverse   src/UM_Net/MMUNet.py
  * MM_Net forward + backward (64x64, batch 2)    src/UM_Net/MMUNet.py
Harness-side patches only: stub modules `causal_conv1d_cuda` / `selective_scan_cuda`,
a stub `timm`, and a `mamba_ssm` namespace whose fused-op names are bound to the reference's own
*_ref functions (the reference has no CPU implementation of the fused ops other than those refs).
Input distributions follow the reference tests (tests/ops/test_selective_scan.py:58-88,
tests/test_causal_conv1d.py:39-50), seed 0.
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import torch
import torch.nn.functional as F

warnings.filterwarnings("ignore")
REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_reference():
    for n in ("causal_conv1d_cuda", "selective_scan_cuda", "timm"):
        sys.modules[n] = types.ModuleType(n)
    sys.path.insert(0, f"{REF}/requirements/Mamba/causal-conv1d")
    import causal_conv1d.causal_conv1d_interface as cci
    spec = importlib.util.spec_from_file_location(
        "ref_ssi", f"{REF}/requirements/Mamba/mamba/mamba_ssm/ops/selective_scan_interface.py")
    ssi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ssi)
    # CPU stand-ins for the fused ops = the reference's own refs
    ssi.causal_conv1d_fn = cci.causal_conv1d_ref
    ssi.selective_scan_fn = ssi.selective_scan_ref

    def inner_no_out_proj(xz, cw, cb, xw, dw, A, B=None, C=None, D=None, delta_bias=None,
                          B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
        d_inner = xz.shape[1] // 2
        eye = torch.eye(d_inner, dtype=xz.dtype)
        return ssi.mamba_inner_ref(xz, cw, cb, xw, dw, eye, None, A, B, C, D, delta_bias).transpose(1, 2)

    pkg = types.ModuleType("mamba_ssm")
    pkg.__path__ = []
    ops = types.ModuleType("mamba_ssm.ops")
    ops.__path__ = []
    iface = types.ModuleType("mamba_ssm.ops.selective_scan_interface")
    iface.selective_scan_fn = ssi.selective_scan_ref
    iface.mamba_inner_fn = ssi.mamba_inner_ref
    iface.bimamba_inner_fn = ssi.bimamba_inner_ref
    iface.mamba_inner_fn_no_out_proj = inner_no_out_proj
    sys.modules.update({"mamba_ssm": pkg, "mamba_ssm.ops": ops,
                        "mamba_ssm.ops.selective_scan_interface": iface})
    cpkg = sys.modules["causal_conv1d"]
    cpkg.causal_conv1d_fn = cci.causal_conv1d_ref
    cpkg.causal_conv1d_update = None
    spec = importlib.util.spec_from_file_location("ref_mamba_simple", f"{REF}/requirements/mamba_simple.py")
    ms = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ms)
    pkg.Mamba = ms.Mamba
    spec = importlib.util.spec_from_file_location("ref_mmunet", f"{REF}/src/UM_Net/MMUNet.py")
    mm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mm)
    return cci, ssi, ms, mm


def npy(t):
    return None if t is None else t.detach().cpu().numpy()


def gen_scan(ssi):
    cases = {}
    for name, (B, D, L, N, G, has_z, has_D, has_bias, softplus) in {
        "ref128": (2, 4, 128, 8, 1, True, True, True, True),      # reference test shape
        "odd37": (2, 3, 37, 4, 1, True, True, True, True),        # ragged L
        "groups2": (2, 4, 64, 8, 2, True, True, True, True),      # varBC_groups=2
        "plain": (1, 2, 50, 16, 1, False, False, False, False),   # no z / D / bias / softplus
        "n16_l300": (2, 6, 300, 16, 1, True, True, True, True),   # MMConv-like D=6, N=16
    }.items():
        torch.manual_seed(0)
        A = (-0.5 * torch.rand(D, N)).requires_grad_()
        shp = (B, N, L) if G == 1 else (B, G, N, L)
        Bm = torch.randn(*shp, requires_grad=True)
        Cm = torch.randn(*shp, requires_grad=True)
        Dp = torch.randn(D, requires_grad=True) if has_D else None
        z = torch.randn(B, D, L, requires_grad=True) if has_z else None
        bias = (0.5 * torch.rand(D)).requires_grad_() if has_bias else None
        u = torch.randn(B, D, L, requires_grad=True)
        delta = (0.5 * torch.rand(B, D, L)).requires_grad_()
        out, last = ssi.selective_scan_ref(u, delta, A, Bm, Cm, Dp, z, bias, softplus, True)
        g = torch.randn_like(out)
        out.backward(g)
        rec = dict(u=u, delta=delta, A=A, B=Bm, C=Cm, D=Dp, z=z, delta_bias=bias, dout=g, out=out,
                   last_state=last, du=u.grad, ddelta=delta.grad, dA=A.grad, dB=Bm.grad, dC=Cm.grad,
                   dD=None if Dp is None else Dp.grad, dz=None if z is None else z.grad,
                   ddelta_bias=None if bias is None else bias.grad)
        for k, v in rec.items():
            if v is not None:
                cases[f"{name}.{k}"] = npy(v)
        cases[f"{name}.softplus"] = np.array(int(softplus))
    np.savez_compressed(os.path.join(OUT, "selective_scan.npz"), **cases)


def gen_conv(cci):
    cases = {}
    for name, (B, D, L, W, silu, has_bias) in {
        "w4_silu": (2, 8, 151, 4, True, True),
        "w3_plain": (2, 8, 64, 3, False, True),
        "w2_nobias": (1, 5, 8, 2, True, False),
        "w4_short": (2, 4, 3, 4, True, True),          # L < width
    }.items():
        torch.manual_seed(0)
        x = torch.randn(B, D, L, requires_grad=True)
        w = torch.randn(D, W, requires_grad=True)
        b = torch.randn(D, requires_grad=True) if has_bias else None
        out = cci.causal_conv1d_ref(x, w, b, "silu" if silu else None)
        g = torch.randn_like(out)
        out.backward(g)
        rec = dict(x=x, w=w, bias=b, dout=g, out=out, dx=x.grad, dw=w.grad, dbias=None if b is None else b.grad)
        for k, v in rec.items():
            if v is not None:
                cases[f"{name}.{k}"] = npy(v)
        cases[f"{name}.silu"] = np.array(int(silu))
    np.savez_compressed(os.path.join(OUT, "causal_conv1d.npz"), **cases)


def gen_inner(ssi):
    """mamba_inner_ref forward + grads (the reference test asserts forward only, test_selective_scan.py:221)."""
    torch.manual_seed(0)
    Bsz, d, L, N, R, dm, W = 2, 8, 48, 4, 2, 4, 4
    t = dict(xz=torch.randn(Bsz, 2 * d, L), conv_w=torch.randn(d, 1, W), conv_b=torch.randn(d),
             x_proj_w=torch.randn(R + 2 * N, d) * 0.3, dt_proj_w=torch.randn(d, R) * 0.3,
             out_proj_w=torch.randn(dm, d) * 0.3, A=-0.5 * torch.rand(d, N), D=torch.randn(d),
             dt_bias=0.5 * torch.rand(d))
    for v in t.values():
        v.requires_grad_()
    out = ssi.mamba_inner_ref(t["xz"], t["conv_w"], t["conv_b"], t["x_proj_w"], t["dt_proj_w"], t["out_proj_w"],
                              None, t["A"], None, None, t["D"], t["dt_bias"], delta_softplus=True)
    g = torch.randn_like(out)
    out.backward(g)
    cases = {k: npy(v) for k, v in t.items()}
    cases.update({"d" + k: npy(v.grad) for k, v in t.items()})
    cases["out"] = npy(out)
    cases["dout"] = npy(g)
    np.savez_compressed(os.path.join(OUT, "mamba_inner.npz"), **cases)


def gen_module(ms):
    """TFM Mamba, v3, forward (4 outputs) + parameter/input grads; and the v1 branch (lines 304-318)
    evaluated through mamba_inner_ref with the same parameters (the shipped constructor asserts v3)."""
    torch.manual_seed(0)
    m = ms.Mamba(d_model=8, d_state=4, d_conv=4, expand=2, bimamba_type="v3", nslices=4)
    x = torch.randn(2, 32, 8, requires_grad=True)
    out, o1, o2, o3 = m(x)
    g = torch.randn_like(out)
    out.backward(g)
    cases = {"x": npy(x), "dout": npy(g), "out": npy(out), "o1": npy(o1), "o2": npy(o2), "o3": npy(o3),
             "dx": npy(x.grad)}
    for k, p in m.named_parameters():
        cases["param." + k] = npy(p)
        cases["grad." + k] = npy(p.grad) if p.grad is not None else np.zeros(0, np.float32)
    # v1 branch
    m.zero_grad()
    x1 = x.detach().clone().requires_grad_()
    m.bimamba_type = "v1"
    try:
        m(x1)
        v1_status = "ran"
    except UnboundLocalError:
        v1_status = "UnboundLocalError"
    cases["v1_status"] = np.array(v1_status)
    sys.modules["mamba_ssm.ops.selective_scan_interface"]
    from einops import rearrange
    xz = rearrange(m.in_proj.weight @ rearrange(x1, "b l d -> d (b l)"), "d (b l) -> b d l", l=32)
    A = -torch.exp(m.A_log.float())
    iface = sys.modules["mamba_ssm.ops.selective_scan_interface"]
    out1 = iface.mamba_inner_fn(xz, m.conv1d.weight, m.conv1d.bias, m.x_proj.weight, m.dt_proj.weight,
                                m.out_proj.weight, m.out_proj.bias, A, None, None, m.D.float(),
                                delta_bias=m.dt_proj.bias.float(), delta_softplus=True)
    m.zero_grad()
    out1.backward(g)
    cases["v1.out"] = npy(out1)
    cases["v1.dx"] = npy(x1.grad)
    for k, p in m.named_parameters():
        cases["v1.grad." + k] = npy(p.grad) if p.grad is not None else np.zeros(0, np.float32)
    np.savez_compressed(os.path.join(OUT, "tfm_mamba.npz"), **cases)


def gen_orders(mm):
    """Exact integer index maps: run the reference permutations on an arange tensor."""
    conv = mm.MMConv.__new__(mm.MMConv)        # methods use no module state
    cases = {}
    for (H, W) in [(4, 6), (5, 3), (1, 7), (2, 2), (16, 16), (7, 8)]:
        src = torch.arange(H * W, dtype=torch.float64).reshape(1, 1, H, W)
        flat = mm.MMConv.two_row_columnwise_flatten_grad_safe(conv, src)
        back = mm.MMConv.inverse_two_row_columnwise_flatten(conv, flat, H, W)
        assert torch.equal(back, src)
        cases[f"tworow.{H}x{W}"] = flat.reshape(-1).to(torch.int64).numpy()
    for (L, ns) in [(16, 4), (64, 16), (36, 6), (8, 1), (8, 8)]:
        src = torch.arange(L, dtype=torch.float64).reshape(1, 1, L)
        xs = torch.stack(src.chunk(ns, dim=-1), dim=-1).flatten(-2)                      # mamba_simple.py:245-247
        cases[f"nslices.{L}_{ns}"] = xs.reshape(-1).to(torch.int64).numpy()
        inv = xs.reshape(1, 1, L // ns, ns).permute(0, 1, 3, 2).flatten(-2)             # :263
        assert torch.equal(inv, src)
    for L in [1, 5, 16]:
        cases[f"flip.{L}"] = torch.arange(L).flip([-1]).numpy()                          # :230
    np.savez_compressed(os.path.join(OUT, "scan_orders.npz"), **cases)


def gen_mm_net(ms, mm):
    """The reference MM_Net (src/UM_Net/MMUNet.py:474-585), unmodified, forward + backward on CPU at 64x64.
    Harness-side patches only: `Mamba` is the reference class behind the v1 shim (constructs as
    "v3" - the shipped constructor asserts it - then runs lines :190-209 + :304-318 for "v1"); MMConv's device default
    -> "cpu"; Dropout2d p -> 0 so that train-mode outputs are deterministic."""
    class ShimMamba(ms.Mamba):
        def __init__(self, *a, bimamba_type="none", **kw):
            super().__init__(*a, bimamba_type="v3", **kw)
            self.bimamba_type = bimamba_type

        def forward(self, hidden_states, inference_params=None):
            if self.bimamba_type in ("v2", "v3"):
                return super().forward(hidden_states, inference_params)
            from einops import rearrange
            L = hidden_states.shape[1]
            xz = rearrange(self.in_proj.weight @ rearrange(hidden_states, "b l d -> d (b l)"), "d (b l) -> b d l", l=L)
            A = -torch.exp(self.A_log.float())
            iface = sys.modules["mamba_ssm.ops.selective_scan_interface"]
            out = iface.mamba_inner_fn(xz, self.conv1d.weight, self.conv1d.bias, self.x_proj.weight, self.dt_proj.weight,
                                       self.out_proj.weight, self.out_proj.bias, A, None, None, self.D.float(),
                                       delta_bias=self.dt_proj.bias.float(), delta_softplus=True)
            return out, None, None, None

    mm.Mamba = ShimMamba
    d = list(mm.MMConv.__init__.__defaults__)
    d[6] = "cpu"
    mm.MMConv.__init__.__defaults__ = tuple(d)
    # ---- single blocks (well conditioned: errors are not amplified through 50 layers) ---------------------------------
    blocks = {}
    for name, (cin, cout, K, H, W) in {"k3_even": (8, 8, 3, 12, 10), "k3_oddH": (8, 12, 3, 7, 9), "k1": (16, 8, 1, 6, 8),
                                        "k9": (4, 4, 9, 10, 16)}.items():
        torch.manual_seed(7)
        conv = mm.MMConv(cin, cout, kernel_size=K, num_slices=4)
        x = torch.randn(2, cin, H, W, requires_grad=True)
        out = conv(x)
        g = torch.randn_like(out)
        out.backward(g)
        rec = {"x": x, "dout": g, "out": out, "dx": x.grad, "shape": torch.tensor([cin, cout, K])}
        for k, p in conv.named_parameters():
            if p.grad is not None:
                rec["grad:" + k] = p.grad
        for k, v in rec.items():
            blocks[f"mmconv_{name}.{k}"] = npy(v)
    torch.manual_seed(7)
    rcg = mm.RCG(num_slices=4)
    rcg.train()
    pre = torch.randn(2, 1, 8, 8, requires_grad=True)
    edge = torch.randn(2, 64, 16, 16, requires_grad=True)
    f = torch.randn(2, 64, 8, 8, requires_grad=True)
    out = rcg(pre, edge, f)
    g = torch.randn_like(out)
    out.backward(g)
    rec = {"pre": pre, "edge": edge, "f": f, "dout": g, "out": out, "dpre": pre.grad, "dedge": edge.grad, "df": f.grad}
    for k in ("mamba.A_log", "mamba.A_b_log", "mamba.A_s_log", "mamba.x_proj_s.weight", "mamba.conv1d_b.weight",
              "mamba.dt_proj.bias", "mamba.D_s", "conv1.0.altho", "conv1.0.mamba.in_proj.weight", "mlp.0.weight"):
        rec["grad:" + k] = dict(rcg.named_parameters())[k].grad
    for k, v in rec.items():
        blocks[f"rcg.{k}"] = npy(v)
    np.savez_compressed(os.path.join(OUT, "mm_blocks.npz"), **blocks)

    torch.manual_seed(50)                                    # train.py:160
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        net = mm.MM_Net(num_classes=1)
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.p = 0.0
    net.train()
    torch.manual_seed(1)
    x = torch.randn(2, 3, 64, 64, requires_grad=True)
    g = torch.randn(2, 1, 64, 64)
    sd = {k: v.clone() for k, v in net.state_dict().items()}      # before the BatchNorm running stats move
    out = net(x)
    out.backward(g)
    cases = {"x": npy(x), "dout": npy(g), "out": npy(out), "dx": npy(x.grad)}
    cases["param_names"] = np.array(list(sd.keys()))
    cases["param_sums"] = np.array([float(v.double().sum()) for v in sd.values()])
    cases["param_abs_sums"] = np.array([float(v.double().abs().sum()) for v in sd.values()])
    grads = {k: p.grad for k, p in net.named_parameters()}
    cases["nograd_names"] = np.array([k for k, v in grads.items() if v is None])
    for k in ("encoder1.0.weight", "encoder2.0.block1.0.mamba.A_log", "encoder2.0.block1.0.mamba.x_proj.weight",
              "encoder2.0.block1.0.altho", "encoder2.0.block1.0.offset_conv.weight", "encoder2.2.block1.3.dsc_conv_x.weight",
              "down5.0.mamba.in_proj.weight", "rcg2.mamba.A_s_log", "rcg2.mamba.conv1d_b.weight", "rcg4.mamba.out_proj.weight",
              "rcg3.mamba.dt_proj_s.bias", "side2.conv2.weight", "decoder2.conv1.0.mamba.dt_proj.weight", "line_predict.weight"):
        cases["grad." + k] = npy(grads[k])
    np.savez_compressed(os.path.join(OUT, "mm_net.npz"), **cases)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    cci, ssi, ms, mm = load_reference()
    gen_scan(ssi)
    gen_conv(cci)
    gen_inner(ssi)
    gen_module(ms)
    gen_orders(mm)
    gen_mm_net(ms, mm)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
