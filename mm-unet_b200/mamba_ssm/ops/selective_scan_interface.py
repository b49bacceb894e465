"""Same import path and names as requirements/Mamba/mamba/mamba_ssm/ops/selective_scan_interface.py."""
from mmunet_b200.ops import (BiMambaInnerFn, MambaInnerFn, MambaInnerFnNoOutProj, SelectiveScanFn,  # noqa: F401
                             bimamba_inner_fn, mamba_inner_fn, mamba_inner_fn_no_out_proj, selective_scan_fn)
