"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that include/mmunet_b200.h
declares, struct layouts agree with the header, and the product refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from conftest import ROOT
from mmunet_b200 import _lib, ops

HEADER = os.path.join(ROOT, "include", "mmunet_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mmunet_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS)
    assert lib.mmu_version() == 100


def test_struct_layout_matches_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mmunet_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(mmu_scan_fwd_params),sizeof(mmu_scan_bwd_params),sizeof(mmu_conv_params),'
                   'offsetof(mmu_scan_fwd_params,u_bs),offsetof(mmu_conv_params,dout),'
                   'sizeof(mmu_narrow_params),offsetof(mmu_narrow_params,dweights),offsetof(mmu_narrow_params,altho));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = list(map(int, subprocess.check_output([str(exe)]).split()))
    want = [ctypes.sizeof(_lib.ScanFwdParams), ctypes.sizeof(_lib.ScanBwdParams), ctypes.sizeof(_lib.ConvParams),
            _lib.ScanFwdParams.u_bs.offset, _lib.ConvParams.dout.offset,
            ctypes.sizeof(_lib.NarrowParams), _lib.NarrowParams.dweights.offset, _lib.NarrowParams.altho.offset]
    assert got == want


def test_workspace_queries_are_pure():
    lib = _lib.lib()
    a = lib.mmu_selective_scan_fwd_workspace(16, 6, 65536, 16)
    assert a == lib.mmu_selective_scan_fwd_workspace(16, 6, 65536, 16) and a > 0
    assert lib.mmu_selective_scan_bwd_workspace(8, 384, 4096, 16) > 0


def test_argument_errors_without_gpu():
    lib = _lib.lib()
    p = _lib.ScanFwdParams()
    assert lib.mmu_selective_scan_fwd(ctypes.byref(p), None) == -1          # empty shape -> MMU_ERR_INVALID
    assert b"empty shape" in lib.mmu_last_error()
    p.batch = p.dim = p.seqlen = 1
    p.dstate = 300
    assert lib.mmu_selective_scan_fwd(ctypes.byref(p), None) == -1
    assert b"256" in lib.mmu_last_error()
    c = _lib.ConvParams()
    c.batch = c.dim = c.seqlen = 1
    c.width = 7
    assert lib.mmu_causal_conv1d_fwd(ctypes.byref(c), None) == -1
    assert b"width" in lib.mmu_last_error()


def test_narrow_block_queries_and_argument_errors_without_gpu():
    """The narrow-block entry points: host-side shape queries, and argument validation before any CUDA call."""
    lib = _lib.lib()
    assert lib.mmu_mamba_narrow_supported(3, 6, 16, 1, 4, _lib.F32) == 1 and lib.mmu_mamba_narrow_supported(3, 6, 16, 1, 4, _lib.BF16) == 1
    assert lib.mmu_mamba_narrow_supported(64, 128, 16, 4, 4, _lib.BF16) == 0 and lib.mmu_mamba_narrow_supported(3, 6, 16, 1, 4, _lib.F16) == 0
    assert lib.mmu_mamba_narrow_rows(6, 16) == 3 * 6 + 2 * 16
    assert lib.mmu_mamba_narrow_weight_floats(3, 6, 16, 1, 4) == 12 * 3 + 6 * 4 + 6 + 33 * 6 + 6 + 3 * 6 + 1
    p = _lib.NarrowParams()
    p.d_model, p.d_inner, p.d_state, p.dt_rank, p.d_conv = 64, 128, 16, 4, 4
    assert lib.mmu_mamba_narrow_pre_fwd(ctypes.byref(p), None) != 0 and b"not a narrow block" in lib.mmu_last_error()
    p.d_model, p.d_inner, p.d_state, p.dt_rank, p.d_conv = 3, 6, 16, 1, 4
    assert lib.mmu_mamba_narrow_pre_fwd(ctypes.byref(p), None) == -1 and b"bad shape" in lib.mmu_last_error()
    p.batch, p.seqlen, p.order, p.order_h, p.order_w = 1, 12, _lib.ORDER_TWOROW, 3, 5          # 3 * 5 != 12
    assert lib.mmu_mamba_narrow_post_fwd(ctypes.byref(p), None) == -1 and b"scan order" in lib.mmu_last_error()
    p.order = 0
    assert lib.mmu_mamba_narrow_pre_bwd(ctypes.byref(p), None) == -1 and b"null tensor" in lib.mmu_last_error()
    assert ops.mamba_narrow_supported(3, 6, 16, 1, 4, torch.float16) is False


def test_no_cpu_fallback():
    u = torch.randn(1, 2, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.selective_scan_fn(u, u, -torch.rand(2, 4), torch.randn(1, 4, 8), torch.randn(1, 4, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.causal_conv1d_fn(u, torch.randn(2, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mm-unet_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_dropin_import_paths():
    from causal_conv1d import causal_conv1d_fn, causal_conv1d_update  # noqa: F401  (mamba_simple.py:13-16)
    from mamba_ssm.ops.selective_scan_interface import (bimamba_inner_fn, mamba_inner_fn,  # noqa: F401  (:18-21)
                                                        mamba_inner_fn_no_out_proj, selective_scan_fn)
    from mamba_ssm import Mamba  # (MMUNet.py:7)
    m = Mamba(d_model=3, d_state=16, d_conv=4, expand=2, bimamba_type="v1", nslices=64)   # MMUNet.py:27-35
    names = set(m.state_dict())
    want = {"A_log", "D", "A_b_log", "D_b", "A_s_log", "D_s", "in_proj.weight", "out_proj.weight"}
    for s in ("", "_b", "_s"):
        want |= {f"conv1d{s}.weight", f"conv1d{s}.bias", f"x_proj{s}.weight", f"dt_proj{s}.weight", f"dt_proj{s}.bias"}
    assert names == want and len(names) == 23
