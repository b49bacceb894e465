"""Drop-in for the reference's `mamba_ssm` package (requirements/Mamba/mamba/mamba_ssm/__init__.py:1-5), restricted to
what MM-UNet imports: `from mamba_ssm import Mamba` (src/UM_Net/MMUNet.py:7) and the fused-op module."""
from mamba_ssm.ops.selective_scan_interface import (bimamba_inner_fn, mamba_inner_fn,  # noqa: F401
                                                    mamba_inner_fn_no_out_proj, selective_scan_fn)
from mmunet_b200.mamba import Mamba  # noqa: F401

__version__ = "1.0.1"
