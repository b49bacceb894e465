"""Time MM_Net training steps on one GPU and print where the time goes.  python scripts/probe_train.py [size] [batch] [dtype] [prof]"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200"))
from mmunet_b200.train import Trainer
size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dtype = sys.argv[3] if len(sys.argv) > 3 else "bf16"
prof = len(sys.argv) > 4 and sys.argv[4] == "prof"
cl = "cl" in sys.argv[4:]
tr = Trainer(image_size=size, batch_per_rank=batch, dtype=dtype, device="cuda:0", ddp=False, channels_last=cl)
tr.set_epoch(2)
x, y = tr.synthetic_batch()
for _ in range(tr.graph_warmup + 3 if tr.use_graph else 3):
    tr.step(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
t0 = time.perf_counter(); e0.record()
for _ in range(n):
    loss = tr.step(x, y)
e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"graph={tr.graph is not None} channels_last={cl} size {size} batch {batch} {dtype}: {e0.elapsed_time(e1)/n:.1f} ms/step (wall {(t1-t0)/n*1e3:.1f} ms), {batch*n/(e0.elapsed_time(e1)/1e3):.2f} img/s, "
      f"loss {float(loss):.4f}, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
if prof:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as p:
        tr.step(x, y)
        torch.cuda.synchronize()
    ka = p.key_averages()
    print(ka.table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
    kern = [k for k in ka if str(k.device_type).endswith("CUDA")]
    kern.sort(key=lambda k: -k.device_time_total)
    tot = sum(k.device_time_total for k in kern)
    print(f"--- kernels only: total {tot/1e3:.1f} ms, {sum(k.count for k in kern)} launches")
    for k in kern[:70]:
        print(f"{k.device_time_total/1e3:9.3f} ms {100*k.device_time_total/tot:5.1f}% {k.count:5d} x {k.device_time_total/k.count:9.1f} us  {k.key[:110]}")
    ours = sum(k.device_time_total for k in kern if "mmu::" in k.key)
    print(f"--- mmu:: kernels {ours/1e3:.1f} ms = {100*ours/tot:.1f}% of GPU time")
