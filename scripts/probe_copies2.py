"""Eager training step under the torch profiler with shapes: which ATen ops launch the big copy / add / reduce kernels.  Not a test."""
import os, sys, collections, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200"))
from mmunet_b200.train import Trainer
from torch.profiler import profile, ProfilerActivity
tr = Trainer(image_size=512, batch_per_rank=16, dtype="bf16", device="cuda:0", ddp=False, channels_last=True, graph=False)
tr.set_epoch(2)
x, y = tr.synthetic_batch()
for _ in range(3):
    tr.step(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as p:
    tr.step(x, y)
    torch.cuda.synchronize()
ka = p.key_averages(group_by_input_shape=True, group_by_stack_n=12)
rows = [k for k in ka if k.self_device_time_total > 0 and k.key.startswith("aten::")]
rows.sort(key=lambda k: -k.self_device_time_total)
for k in rows[:110]:
    if not any(t in k.key for t in ("copy_", "sum", "add", "mul", "fill", "zero", "clone", "contiguous", "amax", "to")):
        continue
    st = [s for s in k.stack if ".py" in s and "torch/" not in s][:4]
    print(f"{k.self_device_time_total/1e3:8.3f} ms {k.count:4d} x  {k.key:22s} {str(k.input_shapes)[:70]}  {' <- '.join(s.split('/')[-1][:50] for s in st)}")
