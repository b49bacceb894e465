"""ORACLE (torch flavour) — test infrastructure only; never imported by the product package.

Pure-PyTorch restatement of the reference's CPU path, used where a test needs autograd through
a *composition* (inner functions, the Mamba module) and by ``bench.py --impl reference``.

Each function names the reference lines it restates:
  selective_scan      selective_scan_interface.py:86-152   (selective_scan_ref)
  causal_conv1d       causal_conv1d_interface.py:49-65     (causal_conv1d_ref)
  mamba_inner         selective_scan_interface.py:636-670  (mamba_inner_ref; out_proj optional, the
                      reference has no *_no_out_proj ref - it is the same body minus the last line)
  bimamba_inner       selective_scan_interface.py:673-709
  mamba_forward       requirements/mamba_simple.py:185-362 (TFM Mamba.forward, v1/v2/v3 branches)
  two_row_flatten / two_row_unflatten   src/UM_Net/MMUNet.py:68-121
  nslices_gather / nslices_scatter      requirements/mamba_simple.py:245-247, 263

The recurrence is the same sequential loop over L as the reference.  The only deliberate
difference: per-step slices are taken with ``unbind`` rather than ``x[:, :, i]`` inside the loop,
which avoids the O(L^2) autograd cost of the reference without changing a
single arithmetic operation.

Parity status: PINNED via tests/test_oracle_golden.py (golden vectors from the real reference).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def selective_scan(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                   return_last_state=False):
    dtype_in = u.dtype
    u = u.float()
    delta = delta.float()
    if delta_bias is not None:
        delta = delta + delta_bias[..., None].float()
    if delta_softplus:
        delta = F.softplus(delta)
    batch, dim, L = u.shape
    N = A.shape[1]
    B = B.float()
    C = C.float()
    if B.dim() == 3:
        B = B[:, None]
    if C.dim() == 3:
        C = C[:, None]
    B = B.repeat_interleave(dim // B.shape[1], dim=1)     # (b, d, n, l)
    C = C.repeat_interleave(dim // C.shape[1], dim=1)
    dA = torch.exp(delta[..., None] * A[None, :, None, :])            # (b, d, l, n)
    dBu = delta[..., None] * B.permute(0, 1, 3, 2) * u[..., None]     # (b, d, l, n)
    h = A.new_zeros((batch, dim, N))
    ys = []
    for a_t, bu_t, c_t in zip(dA.unbind(2), dBu.unbind(2), C.unbind(3)):
        h = a_t * h + bu_t
        ys.append((h * c_t).sum(-1))
    y = torch.stack(ys, dim=2)
    out = y if D is None else y + u * D[:, None]
    if z is not None:
        out = out * F.silu(z.float())
    out = out.to(dtype_in)
    return (out, h) if return_last_state else out


def causal_conv1d(x, weight, bias=None, activation=None):
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu, or swish")
    dtype_in = x.dtype
    L = x.shape[-1]
    dim, width = weight.shape
    out = F.conv1d(x.to(weight.dtype), weight[:, None], bias, padding=width - 1, groups=dim)[..., :L]
    return (out if activation is None else F.silu(out)).to(dtype_in)


def mamba_inner(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, D, delta_bias,
                out_proj_weight=None, out_proj_bias=None, with_out_proj=True, return_parts=False):
    """conv -> x_proj -> dt_proj -> scan [-> out_proj].  xz: (b, 2d, l)."""
    L = xz.shape[-1]
    R = delta_proj_weight.shape[1]
    N = A.shape[-1]
    x, z = xz.chunk(2, dim=1)
    x = causal_conv1d(x, conv1d_weight.reshape(conv1d_weight.shape[0], -1), conv1d_bias, "silu")
    x_dbl = F.linear(x.permute(0, 2, 1).reshape(-1, x.shape[1]), x_proj_weight)       # (b l, R+2N)
    delta = (delta_proj_weight @ x_dbl[:, :R].t()).reshape(-1, xz.shape[0], L).permute(1, 0, 2)
    Bm = x_dbl[:, R:R + N].reshape(xz.shape[0], L, N).permute(0, 2, 1).contiguous()
    Cm = x_dbl[:, -N:].reshape(xz.shape[0], L, N).permute(0, 2, 1).contiguous()
    y = selective_scan(x, delta, A, Bm, Cm, D, z=z, delta_bias=delta_bias, delta_softplus=True)
    if return_parts:
        return dict(conv_out=x, x_dbl=x_dbl, delta=delta, B=Bm, C=Cm, y=y)
    if not with_out_proj:
        return y
    return F.linear(y.permute(0, 2, 1), out_proj_weight, out_proj_bias)


def bimamba_inner(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight,
                  out_proj_bias, A, A_b, D, delta_bias):
    p = mamba_inner(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, D, delta_bias,
                    return_parts=True)
    z = xz.chunk(2, dim=1)[1]
    y_b = selective_scan(p["conv_out"].flip([-1]), p["delta"].flip([-1]), A_b, p["B"].flip([-1]),
                         p["C"].flip([-1]), D, z.flip([-1]), delta_bias, delta_softplus=True)
    y = p["y"] + y_b.flip([-1])
    return F.linear(y.permute(0, 2, 1), out_proj_weight, out_proj_bias)


# ---- scan orders -------------------------------------------------------------------------------

def two_row_flatten(x):
    """(B,C,H,W) -> (B,C,H*W): rows in pairs, column-interleaved; odd last row appended."""
    B, C, H, W = x.shape
    ev = H // 2 * 2
    main = x[:, :, :ev].reshape(B, C, ev // 2, 2, W).permute(0, 1, 2, 4, 3).reshape(B, C, -1)
    if H % 2:
        main = torch.cat([main, x[:, :, ev:].reshape(B, C, -1)], dim=2)
    return main


def two_row_unflatten(xf, H, W):
    B, C, _ = xf.shape
    ev = H // 2 * 2
    main = xf[:, :, :ev * W].reshape(B, C, ev // 2, W, 2).permute(0, 1, 2, 4, 3).reshape(B, C, ev, W)
    if H % 2:
        main = torch.cat([main, xf[:, :, ev * W:].reshape(B, C, 1, W)], dim=2)
    return main


def nslices_gather(x, ns):
    """x[..., s*(L/ns)+j] -> y[..., j*ns+s]."""
    return torch.stack(x.chunk(ns, dim=-1), dim=-1).flatten(-2)


def nslices_scatter(y, ns):
    L = y.shape[-1]
    return y.reshape(*y.shape[:-1], L // ns, ns).transpose(-1, -2).flatten(-2)


# ---- TFM Mamba.forward ---------------------------------------------------------------------------

def mamba_forward(m, hidden_states):
    """Functional restatement of TFM ``Mamba.forward`` over a module-like object ``m`` carrying the
    reference's parameter names.  Returns (out, o_1, o_2, o_3); o_* are None outside v3."""
    batch, L, _ = hidden_states.shape
    xz = (m.in_proj.weight @ hidden_states.permute(2, 0, 1).reshape(hidden_states.shape[2], -1))
    xz = xz.reshape(-1, batch, L).permute(1, 0, 2)
    if m.in_proj.bias is not None:
        xz = xz + m.in_proj.bias.to(xz.dtype)[:, None]
    A = -torch.exp(m.A_log.float())

    def inner(xz_, sfx, A_):
        conv = getattr(m, "conv1d" + sfx)
        return mamba_inner(xz_, conv.weight, conv.bias, getattr(m, "x_proj" + sfx).weight,
                           getattr(m, "dt_proj" + sfx).weight, A_, getattr(m, "D" + sfx).float(),
                           getattr(m, "dt_proj" + sfx).bias.float(), with_out_proj=False)

    if m.bimamba_type in ("v2", "v3"):
        out = inner(xz, "", A)
        out_b = inner(xz.flip([-1]), "_b", -torch.exp(m.A_b_log.float()))
        total = out + out_b.flip([-1])
        o3 = None
        if m.bimamba_type == "v3":
            out_s = inner(nslices_gather(xz, m.nslices), "_s", -torch.exp(m.A_s_log.float()))
            o3 = nslices_scatter(out_s, m.nslices)
            total = total + o3
        res = F.linear(total.permute(0, 2, 1), m.out_proj.weight, m.out_proj.bias)
        return (res, out, out_b, o3) if m.bimamba_type == "v3" else (res, None, None, None)
    y = inner(xz, "", A)
    return F.linear(y.permute(0, 2, 1), m.out_proj.weight, m.out_proj.bias), None, None, None
