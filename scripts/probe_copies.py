"""Which tensors get copied (layout / dtype conversions) in one eager MM_Net training step?  Groups aten::copy_ by input shape."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200"))
os.environ["MMU_TRAIN_GRAPH"] = "0"
from mmunet_b200.train import Trainer
from torch.profiler import profile, ProfilerActivity
cl = "cl" in sys.argv
tr = Trainer(image_size=512, batch_per_rank=16, dtype="bf16", device="cuda:0", ddp=False, channels_last=cl)
tr.set_epoch(2)
x, y = tr.synthetic_batch()
for _ in range(3):
    tr.step(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as p:
    tr.step(x, y)
    torch.cuda.synchronize()
ka = p.key_averages(group_by_input_shape=True)
rows = [k for k in ka if k.key in ("aten::copy_",)]
rows.sort(key=lambda k: -k.self_device_time_total)
tot = sum(k.self_device_time_total for k in rows)
print(f"aten::copy_ total {tot/1e3:.1f} ms in {sum(k.count for k in rows)} calls (channels_last={cl})")
for k in rows[:40]:
    print(f"{k.self_device_time_total/1e3:8.3f} ms {k.count:4d} x  {k.input_shapes}")
for name in ("aten::native_group_norm", "aten::native_group_norm_backward", "aten::cudnn_batch_norm", "aten::cudnn_batch_norm_backward",
             "aten::native_batch_norm", "aten::native_batch_norm_backward", "aten::cat", "aten::upsample_bilinear2d", "aten::add", "aten::add_", "aten::mul", "aten::relu_", "aten::threshold_backward"):
    r = [k for k in ka if k.key == name]
    if r:
        print(f"{name}: {sum(k.self_device_time_total for k in r)/1e3:.2f} ms in {sum(k.count for k in r)} calls")
