"""mmunet_b200 - B200-native (sm_100a) implementation of MM-UNet's Mamba-block hot path.

Public surface (mirrors the reference's operator API):
    from mmunet_b200 import Mamba, selective_scan_fn, causal_conv1d_fn, mamba_inner_fn,
                            mamba_inner_fn_no_out_proj, bimamba_inner_fn
The sibling packages `mamba_ssm/` and `causal_conv1d/` in this directory re-export the same objects under the
reference's import paths, so `requirements/mamba_simple.py` and `src/UM_Net/MMUNet.py` run unchanged with
`mm-unet_b200/` on PYTHONPATH.
"""
from .ops import (bimamba_inner_fn, causal_conv1d_fn, mamba_inner_fn, mamba_inner_fn_no_out_proj,  # noqa: F401
                  scan_order_gather, scan_order_index, scan_order_scatter, selective_scan_fn,
                  two_row_flatten, two_row_unflatten)
from .mamba import Mamba  # noqa: F401

__version__ = "0.1.0"
