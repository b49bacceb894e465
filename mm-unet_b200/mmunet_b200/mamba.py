"""`Mamba` module with the interface of MM-UNet's TFM variant (requirements/mamba_simple.py:34-362).

Same constructor arguments, same 23 parameter tensors with the same names / shapes / initialisers (created in
the same order, so a given torch seed yields the same weights as the reference), same forward contract:
``forward(hidden_states (B, L, d_model)) -> (out, o_1, o_2, o_3)``.

Differences, all forced or performance-only:
  * the reference constructor asserts ``bimamba_type == "v3"`` (:125) and its non-v2/v3 branch leaves o_1..o_3 unbound
    (:303-318, 362) although 47 of MM_Net's 50 Mamba instances are "v1" (SURVEY.md section 0.4).  Here every
    bimamba_type constructs; "v1"/"none" run lines :304-318 and return (out, None, None, None); "v2" returns
    (out, None, None, None) as well.
  * v2/v3 backward direction: ``xz.flip(-1)`` / ``out_b.flip(-1)`` (:230, :270) are not materialised - the conv runs
    anti-causally and the scan kernel walks the sequence right-to-left (bit-identical token order).
  * v3 slice direction: the chunk/stack/flatten interleave (:245-247) and its inverse (:263) are one gather / one
    scatter kernel with the closed-form index map.
  * ``o_2`` is, in the reference, ``out_b`` in flipped token order.  It is produced lazily: set
    ``module.return_directional = False`` (MM-UNet ignores o_1..o_3) to skip that extra copy.
  * decode (``inference_params`` / ``step``) is outside the training hot path and raises NotImplementedError.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops


class Mamba(nn.Module):
    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1,
                 dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True, bias=False, use_fast_path=True,
                 layer_idx=None, device=None, dtype=None, bimamba_type="none", nslices=5):
        fk = {"device": device, "dtype": dtype}
        super().__init__()
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.use_fast_path, self.layer_idx = use_fast_path, layer_idx
        self.bimamba_type, self.nslices = bimamba_type, nslices
        self.return_directional = True
        self.concurrent_directions = True      # v2 / v3: run the independent scan directions on side streams
        self._side_streams = {}
        self.activation = "silu"
        self.act = nn.SiLU()
        di, R, N = self.d_inner, self.dt_rank, d_state

        def conv():
            return nn.Conv1d(di, di, kernel_size=d_conv, groups=di, padding=d_conv - 1, bias=conv_bias, **fk)

        def a_log():
            A = torch.arange(1, N + 1, dtype=torch.float32, device=device).repeat(di, 1).contiguous()
            p = nn.Parameter(torch.log(A))          # S4D-real init, kept in fp32
            p._no_weight_decay = True
            return p

        def d_skip():
            p = nn.Parameter(torch.ones(di, device=device))
            p._no_weight_decay = True
            return p

        self.in_proj = nn.Linear(d_model, di * 2, bias=bias, **fk)
        self.conv1d = conv()
        self.x_proj = nn.Linear(di, R + 2 * N, bias=False, **fk)
        self.dt_proj = nn.Linear(R, di, bias=True, **fk)
        dt_init_std = R ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(self.dt_proj.weight, dt_init_std)
        elif dt_init == "random":
            nn.init.uniform_(self.dt_proj.weight, -dt_init_std, dt_init_std)
        else:
            raise NotImplementedError
        dt = torch.exp(torch.rand(di, **fk) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min)).clamp(min=dt_init_floor)
        inv_dt = dt + torch.log(-torch.expm1(-dt))   # inverse softplus
        with torch.no_grad():
            self.dt_proj.bias.copy_(inv_dt)
        self.dt_proj.bias._no_reinit = True
        self.A_log = a_log()
        self.D = d_skip()
        # backward ("_b") and slice ("_s") parameter sets are created unconditionally, as in the reference (:127-181)
        self.A_b_log = a_log()
        self.conv1d_b = conv()
        self.x_proj_b = nn.Linear(di, R + 2 * N, bias=False, **fk)
        self.dt_proj_b = nn.Linear(R, di, bias=True, **fk)
        self.D_b = d_skip()
        self.A_s_log = a_log()
        self.conv1d_s = conv()
        self.x_proj_s = nn.Linear(di, R + 2 * N, bias=False, **fk)
        self.dt_proj_s = nn.Linear(R, di, bias=True, **fk)
        self.D_s = d_skip()
        self.out_proj = nn.Linear(di, d_model, bias=bias, **fk)

    def _streams(self, device):
        key = (device.type, device.index)
        if key not in self._side_streams:
            self._side_streams[key] = (torch.cuda.Stream(device), torch.cuda.Stream(device))
        return self._side_streams[key]

    def _inner(self, xz, sfx, reverse=False):
        conv = getattr(self, "conv1d" + sfx)
        dtp = getattr(self, "dt_proj" + sfx)
        A = -torch.exp(getattr(self, "A" + sfx + "_log").float())
        args = (xz, conv.weight, conv.bias, getattr(self, "x_proj" + sfx).weight, dtp.weight, A)
        D, db = getattr(self, "D" + sfx).float(), dtp.bias.float()
        if reverse:
            return ops.mamba_inner_fn_no_out_proj_reversed(*args, D=D, delta_bias=db, delta_softplus=True)
        return ops.mamba_inner_fn_no_out_proj(*args, None, None, D, delta_bias=db, delta_softplus=True)

    def forward(self, hidden_states, inference_params=None, scan_order=None, coord_epilogue=None):
        """scan_order = (kind, H, W, nslices) (extension, single-direction types only): the tokens of hidden_states are in natural
        order and the block scans them in that order, returning natural order - what MMConv gets from flatten -> Mamba ->
        inverse flatten (src/UM_Net/MMUNet.py:178-183), with the permutation inside the conv / scan kernels' addressing."""
        if inference_params is not None:
            raise NotImplementedError("mmunet_b200.Mamba: decode / inference_params is outside the training hot path")
        if scan_order is not None and (not self.use_fast_path or self.bimamba_type in ("v2", "v3")):
            raise NotImplementedError("mmunet_b200.Mamba: scan_order is for the single-direction fast path")
        batch, seqlen, dim = hidden_states.shape
        # in_proj and BLD -> BDL in one go (mamba_simple.py:201-205).  MM-UNet's token tensors are transposed views of
        # channel-major (b, d_model, l) maps (MMUNet.py:180, 405): then xz = W @ X[b] is a contiguous (b, 2d, l) batched
        # matmul with no copy; otherwise the reference's "d (b l)" form, a (b, 2d, l) view with strides (l, b*l, 1).
        tokens_cm = hidden_states.transpose(1, 2)
        if self.narrow_block_available(hidden_states):
            # narrow block (MMConv's d_model = 3 Mamba): in_proj .. dt_proj and out_proj run inside two fused kernels around the scan
            altho, coord = (coord_epilogue[0], tuple(coord_epilogue[1:])) if coord_epilogue is not None else (None, None)
            out = ops.mamba_narrow_fn(tokens_cm, self.in_proj.weight, self.conv1d.weight, self.conv1d.bias, self.x_proj.weight,
                                      self.dt_proj.weight, self.out_proj.weight, -torch.exp(self.A_log.float()), self.D.float(),
                                      self.dt_proj.bias.float(), order=scan_order, altho=altho, coord=coord)
            return out.transpose(1, 2), None, None, None
        if coord_epilogue is not None:
            raise NotImplementedError("mmunet_b200.Mamba: coord_epilogue needs the fused narrow block (see narrow_block_available)")
        if tokens_cm.stride(-1) == 1 or seqlen == 1:
            xz = torch.matmul(self.in_proj.weight, tokens_cm)
        else:
            # token-major input (e.g. a channels-last map flattened to (b, l, c)): X^T enters the GEMM as a transposed operand,
            # no "d (b l)" copy (mamba_simple.py:201-205 materialises it)
            flat = hidden_states.reshape(batch * seqlen, dim)
            xz = (self.in_proj.weight @ flat.t()).view(-1, batch, seqlen).transpose(0, 1)
        if self.in_proj.bias is not None:
            xz = xz + self.in_proj.bias.to(dtype=xz.dtype)[:, None]
        o_1 = o_2 = o_3 = None
        if not self.use_fast_path:
            out = self._slow_path(xz, seqlen)
        elif self.bimamba_type in ("v2", "v3"):
            v3 = self.bimamba_type == "v3"
            ns = self.nslices
            if v3 and seqlen % ns != 0:
                raise RuntimeError(f"Mamba v3: sequence length {seqlen} is not divisible by nslices {ns} "
                                   "(torch.stack fails in the reference, mamba_simple.py:245-246)")

            def slice_direction():
                # xz_s = stack(xz.chunk(ns, -1), -1).flatten(-2); out_s scattered back (mamba_simple.py:245-247, 263): the index
                # map is applied by the conv / scan kernels' own loads and stores (explicit gather / scatter if not fusable)
                conv, dtp = self.conv1d_s, self.dt_proj_s
                return ops.mamba_inner_fn_no_out_proj_ordered(xz, conv.weight, conv.bias, self.x_proj_s.weight, dtp.weight,
                                                              -torch.exp(self.A_s_log.float()), self.D_s.float(), dtp.bias.float(), True,
                                                              order=(_lib.ORDER_NSLICES, 1, seqlen, ns))

            if self.concurrent_directions and xz.is_cuda:
                # The directions are independent until the sum: run them on side streams so that their kernels share the GPU
                # (one direction's scan is 256 CTAs at MM-UNet's RCG shapes, about half of the resident-CTA slots).  Autograd
                # replays each direction's backward on the stream its forward ran on; a captured CUDA graph keeps the fork.
                cur = torch.cuda.current_stream(xz.device)
                side = self._streams(xz.device)
                for st in side:
                    st.wait_stream(cur)
                out_f = self._inner(xz, "")
                with torch.cuda.stream(side[0]):
                    out_b = self._inner(xz, "_b", reverse=True)       # already in un-flipped token order
                    out_b.record_stream(cur)
                out_s = None
                if v3:
                    with torch.cuda.stream(side[1]):
                        out_s = slice_direction()
                        out_s.record_stream(cur)
                for st in side:
                    cur.wait_stream(st)
            else:
                out_f = self._inner(xz, "")
                out_b = self._inner(xz, "_b", reverse=True)
                out_s = slice_direction() if v3 else None
            total = out_f + out_b
            if v3:
                total = total + out_s
                if self.return_directional:
                    o_1, o_2, o_3 = out_f, out_b.flip([-1]), out_s
            out = ops._out_proj_autograd(total, self.out_proj.weight, self.out_proj.bias)
        else:
            A = -torch.exp(self.A_log.float())
            if scan_order is not None:
                out = ops.mamba_inner_fn_ordered(xz, self.conv1d.weight, self.conv1d.bias, self.x_proj.weight, self.dt_proj.weight,
                                                 self.out_proj.weight, self.out_proj.bias, A, self.D.float(), self.dt_proj.bias.float(), True,
                                                 order=scan_order)
            else:
                out = ops.mamba_inner_fn(xz, self.conv1d.weight, self.conv1d.bias, self.x_proj.weight, self.dt_proj.weight,
                                         self.out_proj.weight, self.out_proj.bias, A, None, None, self.D.float(),
                                         delta_bias=self.dt_proj.bias.float(), delta_softplus=True)
        return out, o_1, o_2, o_3

    def narrow_block_available(self, like) -> bool:
        """True when forward() runs `like`-shaped CUDA inputs through the fused narrow block (and so accepts coord_epilogue)."""
        return bool(self.use_fast_path and self.bimamba_type not in ("v2", "v3") and like.is_cuda and self.in_proj.bias is None
                    and self.out_proj.bias is None
                    and ops.mamba_narrow_supported(self.d_model, self.d_inner, self.d_state, self.dt_rank, self.d_conv,
                                                   ops._autocast_dtype() or like.dtype))

    def _slow_path(self, xz, seqlen):
        """use_fast_path=False (mamba_simple.py:319-361): un-fused ops, single direction."""
        x, z = xz.chunk(2, dim=1)
        x = ops.causal_conv1d_fn(x, self.conv1d.weight.reshape(self.d_inner, -1), self.conv1d.bias, self.activation)
        x_dbl = self.x_proj(x.transpose(1, 2).reshape(-1, self.d_inner))
        dt, B, C = torch.split(x_dbl, [self.dt_rank, self.d_state, self.d_state], dim=-1)
        dt = (self.dt_proj.weight @ dt.t()).view(self.d_inner, -1, seqlen).transpose(0, 1)
        B = B.reshape(-1, seqlen, self.d_state).transpose(1, 2).contiguous()
        C = C.reshape(-1, seqlen, self.d_state).transpose(1, 2).contiguous()
        y = ops.selective_scan_fn(x, dt, -torch.exp(self.A_log.float()), B, C, self.D.float(), z=z,
                                  delta_bias=self.dt_proj.bias.float(), delta_softplus=True)
        return self.out_proj(y.transpose(1, 2))

    def step(self, hidden_states, conv_state, ssm_state):
        raise NotImplementedError("mmunet_b200.Mamba.step: decode is outside the training hot path")
