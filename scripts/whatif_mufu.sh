# Timing probe: product library vs the MMU_S5_WHATIF build (half of the ring forward's exponentials skipped; results wrong).  Not a test.
for lib in "" mm-unet_b200/mmunet_b200/libmmunet_b200_whatif.so; do
  MMU_LIB=$lib python - <<'PY'
import os, sys, torch
sys.path.insert(0, "mm-unet_b200"); sys.path.insert(0, ".")
from mmunet_b200 import ops, _lib
from scripts.probe_scan import make, timeit
for (B, D, L) in ((8, 384, 4096), (16, 128, 4096)):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, 16, torch.float32)
    ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
    t = timeit(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True), warm=3, it=20)
    print(os.environ.get("MMU_LIB") or "product", f"B{B} D{D} L{L}: fwd {t:.0f} us", flush=True)
PY
done
