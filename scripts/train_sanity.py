"""Longer end-to-end sanity run of the trainer (CUDA graph, channels-last, bf16): a fixed set of 4 synthetic batches whose masks are a
deterministic function of the image (so there is something to learn), 120 steps; prints the loss every 10 steps."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200"))
from mmunet_b200.train import Trainer
tr = Trainer(image_size=256, batch_per_rank=8, dtype="bf16", device="cuda:0", ddp=False, channels_last=True)
tr.set_epoch(2)
g = torch.Generator().manual_seed(0)
batches = []
for _ in range(4):
    x = torch.randn(8, 3, 256, 256, generator=g)
    smooth = torch.nn.functional.avg_pool2d(x.mean(1, keepdim=True), 9, 1, 4)
    y = (smooth > 0.25 * smooth.std()).to(torch.uint8)          # "vessels": a smooth threshold of the image itself
    batches.append((x.pin_memory(), y.pin_memory()))
losses = []
for i in range(120):
    losses.append(float(tr.step(*batches[i % 4])))
    if i % 10 == 9:
        print(f"step {i + 1:3d}  loss {sum(losses[-4:]) / 4:.4f}  graph={tr.graph is not None}", flush=True)
assert all(l == l for l in losses), "NaN"
assert sum(losses[-4:]) < 0.8 * sum(losses[:4]), (losses[:4], losses[-4:])
print("ok")
