"""MM_Net on the CUDA hot path vs the unmodified reference MM_Net run on CPU (tests/golden/mm_net.npz): same seed ->
same weights, same input -> logits, input gradient and parameter gradients must agree.  Also one optimisation step of
the synthetic-data trainer."""
import numpy as np
import pytest
import torch

import _mm_blocks
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _build():
    from mmunet_b200.mm_net import MM_Net
    torch.manual_seed(50)
    net = MM_Net(num_classes=1)
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.p = 0.0
    return net.cuda().train()


@pytest.fixture()
def no_tf32():
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("name", _mm_blocks.MMCONV_CASES)
def test_mmconv_block_matches_reference(no_tf32, name):
    from mmunet_b200 import mm_net
    _mm_blocks.check_mmconv(mm_net, name, "cuda", 1e-3)


def test_rcg_block_matches_reference(no_tf32):
    from mmunet_b200 import mm_net
    _mm_blocks.check_rcg(mm_net, "cuda", 1e-3)


def test_mm_net_matches_reference_fp32(no_tf32):
    """Whole model, random init, train-mode BatchNorm over 2x2 maps: a 1e-6 relative input perturbation already moves the
    reference's own logits by 3e-3 and its input gradient by 5 % (measured on CPU), so the end-to-end check is a loose
    forward tolerance plus gradient direction; tight parity is asserted block by block above."""
    g = np.load(f"{GOLDEN}/mm_net.npz")
    net = _build()
    x = torch.tensor(g["x"], device="cuda", requires_grad=True)
    out = net(x)
    ref = g["out"]
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref, rtol=1e-2, atol=1e-2 * float(np.abs(ref).max()))
    out.backward(torch.tensor(g["dout"], device="cuda"))

    def cos(a, b):
        a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
        return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))

    assert cos(x.grad.cpu().numpy(), g["dx"]) > 0.98
    grads = {k: p.grad for k, p in net.named_parameters()}
    for key in g.files:
        if key.startswith("grad."):
            assert cos(grads[key[5:]].cpu().numpy(), g[key]) > 0.8, key
    assert sorted(k for k, v in grads.items() if v is None) == sorted(str(k) for k in g["nograd_names"])


@pytest.mark.parametrize("dtype,channels_last,graph", [("fp32", False, False), ("bf16", False, False), ("bf16", True, False),
                                                       ("bf16", True, True)])
def test_trainer_steps_reduce_loss(dtype, channels_last, graph):
    """Eager and whole-step CUDA-graph training, NCHW and channels-last (the latter runs the NHWC sampler / GroupNorm kernels)."""
    from mmunet_b200.train import Trainer
    tr = Trainer(image_size=64, batch_per_rank=2, dtype=dtype, device="cuda:0", ddp=False, channels_last=channels_last, graph=graph)
    tr.set_epoch(2)                       # past the warm-up: lr = 1e-3
    x, y = tr.synthetic_batch()
    losses = [float(tr.step(x, y)) for _ in range(8)]
    assert all(np.isfinite(losses)), losses
    assert losses[-1] < losses[0], losses
    assert (tr.graph is not None) == graph


def test_mm_net_ragged_sizes_channels_last(no_tf32):
    """96x96 input (the model needs a multiple of 32, like the reference): feature maps 48, 24, 12, 6, 3 - an odd map (two-row
    flatten with a tail row) and L = 9 / 36 on the generic scan kernels next to L = 144 / 576 / 2304 on the v3 kernels - in
    channels-last bf16; and the channels-last model must agree with the NCHW model on the same weights (different kernels:
    NHWC sampler / GroupNorm vs NCHW sampler / ATen GroupNorm; TF32 off, it alone moves the logits by 3 %)."""
    from mmunet_b200.mm_net import MM_Net
    torch.manual_seed(50)
    net = MM_Net(num_classes=1).cuda().eval()          # eval: BatchNorm uses running stats -> well conditioned comparison
    x = torch.randn(2, 3, 64, 64, device="cuda")
    with torch.no_grad():
        ref = net(x)
        got = net.to(memory_format=torch.channels_last)(x.contiguous(memory_format=torch.channels_last))
    torch.testing.assert_close(got, ref, rtol=2e-3, atol=2e-3 * float(ref.abs().max()))
    net.train()
    xr = torch.randn(2, 3, 96, 96, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = net(xr)
    assert out.shape == (2, 1, 96, 96) and torch.isfinite(out).all()
    out.float().sum().backward()
    assert torch.isfinite(xr.grad).all()
    assert all(torch.isfinite(p.grad).all() for p in net.parameters() if p.grad is not None)


@pytest.mark.parametrize("shape", [(2, 8, 12, 10, 3), (1, 5, 7, 9, 3), (2, 16, 6, 8, 1), (1, 4, 10, 16, 9), (3, 32, 33, 17, 3)])
@pytest.mark.parametrize("dtypes", [("fp32", "fp32"), ("fp32", "bf16"), ("bf16", "bf16")])
@pytest.mark.parametrize("channels_last", [False, True])
def test_snake_sample_matches_grid_sample(no_tf32, shape, dtypes, channels_last):
    """The fused sampler vs the reference formulation (coordinate rescale + F.grid_sample, MMUNet.py:190-224) on the same
    inputs, forward and both gradients; y reaches well outside [0, H-1] so that the clamp and its gradient mask are hit."""
    from mmunet_b200 import mm_net, ops
    B, C, H, W, K = shape
    tin, tout = (torch.float32 if d == "fp32" else torch.bfloat16 for d in dtypes)
    torch.manual_seed(3)
    conv = mm_net.MMConv(C, 4, kernel_size=K).cuda()
    feat = torch.randn(B, C, H, W, device="cuda").to(tin)
    if channels_last:      # NHWC kernels when C = 4 * 2^k, else the op falls back to the NCHW kernels on a contiguous copy
        feat = feat.contiguous(memory_format=torch.channels_last)
    y = (torch.arange(H, device="cuda").view(1, 1, H, 1) + 1.7 * torch.randn(B, K, H, W, device="cuda")).float()
    f1, y1 = feat.clone().requires_grad_(), y.clone().requires_grad_()
    f2, y2 = feat.float().clone().requires_grad_(), y.clone().requires_grad_()
    got = ops.snake_sample(f1, y1, tout)
    ref = conv._grid_sample(f2, y2)
    assert got.dtype == tout and got.shape == ref.shape
    if channels_last and C in (4, 8, 16, 32):
        assert got.is_contiguous(memory_format=torch.channels_last)
    tol = 1e-5 if tout == torch.float32 else 1e-2
    torch.testing.assert_close(got.float(), ref, rtol=tol, atol=tol)
    g = torch.randn_like(ref)
    got.backward(g.to(tout))
    ref.backward(g.to(tout).float())
    gt = 1e-4 if tin == torch.float32 and tout == torch.float32 else 2e-2
    torch.testing.assert_close(f1.grad.float(), f2.grad, rtol=gt, atol=gt * float(f2.grad.abs().max()))
    # d/dy: piecewise constant in y, so a sample within rounding distance of an integer row may take the neighbouring slope
    bad = ((y1.grad - y2.grad).abs() > gt * (float(y2.grad.abs().max()) + y2.grad.abs())).float().mean()
    assert bad <= 2e-3, f"{bad:.4f} of d_y elements differ"


def test_snake_sample_nan_and_fp16(no_tf32):
    """A NaN row coordinate poisons the sample and its gradient (as torch.clamp + grid_sample do) instead of silently reading
    row 0; fp16 feature maps (not supported by the fused sampler) take the reference formulation instead of raising."""
    from mmunet_b200 import mm_net, ops
    feat = torch.randn(1, 4, 6, 5, device="cuda", requires_grad=True)
    y = torch.arange(6, device="cuda").view(1, 1, 6, 1).expand(1, 3, 6, 5).float().clone()
    y[0, 1, 2, 3] = float("nan")
    y.requires_grad_()
    out = ops.snake_sample(feat, y)
    assert torch.isnan(out[0, :, 2 * 3 + 1, 3]).all() and torch.isfinite(out[0, :, 0]).all()
    out.nan_to_num().sum().backward()
    conv = mm_net.MMConv(8, 8, kernel_size=3).cuda()
    with torch.autocast("cuda", dtype=torch.float16):      # fp16 autocast: the layer before hands MMConv an fp16 map
        o16 = conv(torch.randn(1, 8, 6, 6, device="cuda", dtype=torch.float16))
    assert o16.shape == (1, 8, 6, 6) and torch.isfinite(o16.float()).all()


@pytest.mark.parametrize("shape", [(2, 64, 12, 10), (3, 16, 7, 9), (1, 512, 4, 4), (2, 128, 33, 17), (2, 8, 5, 5)])
@pytest.mark.parametrize("dtypes", [("fp32", "fp32"), ("bf16", "bf16"), ("fp32", "bf16")])
def test_group_norm_nhwc_matches_torch(shape, dtypes):
    """Channels-last GroupNorm (4 channels per group) vs F.group_norm evaluated in fp32 on the same (rounded) inputs."""
    import torch.nn.functional as F
    from mmunet_b200 import ops
    B, C, H, W = shape
    G = C // 4
    tin, tout = (torch.float32 if d == "fp32" else torch.bfloat16 for d in dtypes)
    torch.manual_seed(5)
    x = (1.5 * torch.randn(B, C, H, W, device="cuda") + 0.7).to(tin).contiguous(memory_format=torch.channels_last)
    w = torch.randn(C, device="cuda")
    b = torch.randn(C, device="cuda")
    assert ops.group_norm_nhwc_supported(x, G)
    x1, w1, b1 = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    x2, w2, b2 = x.float().clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    y = ops.group_norm_nhwc(x1, G, w1, b1, 1e-5, tout)
    ref = F.group_norm(x2, G, w2, b2, 1e-5)
    assert y.dtype == tout and y.is_contiguous(memory_format=torch.channels_last)
    tol = 2e-5 if tout == torch.float32 else 2e-2
    torch.testing.assert_close(y.float(), ref, rtol=tol, atol=tol)
    g = torch.randn_like(ref)
    y.backward(g.to(tout))
    ref.backward(g.to(tout).float())
    gt = 2e-4 if (tin, tout) == (torch.float32, torch.float32) else 3e-2
    for got, want, name in ((x1.grad, x2.grad, "dx"), (w1.grad, w2.grad, "dgamma"), (b1.grad, b2.grad, "dbeta")):
        torch.testing.assert_close(got.float(), want, rtol=gt, atol=gt * float(want.abs().max()), msg=lambda m: f"{name}: {m}")
    # statistics must survive a large common offset (one-pass E[x^2] - mean^2 would lose every digit here)
    if tin == torch.float32:
        xo = (x.float() + 300.0).contiguous(memory_format=torch.channels_last)
        yo = ops.group_norm_nhwc(xo, G, w, b, 1e-5, torch.float32)
        torch.testing.assert_close(yo, F.group_norm(xo.double(), G, w.double(), b.double(), 1e-5).float(), rtol=2e-3, atol=2e-3)
    # unsupported layouts are refused (the model then falls back to nn.GroupNorm)
    assert not ops.group_norm_nhwc_supported(x.contiguous(), G) or H * W == 1
    assert not ops.group_norm_nhwc_supported(x, G * 2)
