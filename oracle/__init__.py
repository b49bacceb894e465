"""ORACLE — test infrastructure only; never imported by the product package.

ctypes/numpy front end of ``oracle/scan_oracle.c`` (plain-C restatement of the reference's CPU
algorithms; see that file's header for the reference file:line map) plus ``oracle.torch_ref``
(torch restatement used where autograd through a composition is needed).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.

Parity status: PINNED against golden vectors generated from the unmodified reference
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``, checked by ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_SRC = os.path.join(_HERE, "scan_oracle.c")


def build(force: bool = False) -> str:
    """Compile scan_oracle.c -> liboracle.so with gcc (OpenMP when available)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"]
        try:
            subprocess.check_call(cmd)
        except subprocess.CalledProcessError:
            cmd.remove("-fopenmp")
            subprocess.check_call(cmd)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _bc4(x, B, N, L):
    x = _f32(x)
    if x.ndim == 3:
        x = x.reshape(B, 1, N, L)
    return x


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads of the C oracle (torchrun exports OMP_NUM_THREADS=1 to its workers); returns the count now in effect."""
    lib().oracle_set_num_threads(int(n))
    return num_threads()


def selective_scan_fwd(u, delta, A, Bm, Cm, D=None, z=None, delta_bias=None, delta_softplus=False):
    """-> (out (B,D,L), last_state (B,D,N)).  selective_scan_interface.py:86-152."""
    u, delta, A = _f32(u), _f32(delta), _f32(A)
    B_, D_, L = u.shape
    N = A.shape[1]
    Bm, Cm = _bc4(Bm, B_, N, L), _bc4(Cm, B_, N, L)
    G = Bm.shape[1]
    D, z, delta_bias = _f32(D), _f32(z), _f32(delta_bias)
    out = np.empty_like(u)
    last = np.empty((B_, D_, N), np.float32)
    i64 = ctypes.c_int64
    rc = lib().oracle_selective_scan_fwd(_p(u), _p(delta), _p(A), _p(Bm), _p(Cm), _p(D), _p(z),
                                         _p(delta_bias), int(bool(delta_softplus)), i64(B_), i64(D_),
                                         i64(L), i64(N), i64(G), _p(out), _p(last))
    assert rc == 0, rc
    return out, last


def selective_scan_bwd(u, delta, A, Bm, Cm, D, z, delta_bias, dout, delta_softplus=False):
    """-> dict(du, ddelta, dA, dB, dC, dD, dz, ddelta_bias); closed-form backward."""
    u, delta, A, dout = _f32(u), _f32(delta), _f32(A), _f32(dout)
    B_, D_, L = u.shape
    N = A.shape[1]
    squeeze = np.ndim(Bm) == 3
    Bm, Cm = _bc4(Bm, B_, N, L), _bc4(Cm, B_, N, L)
    G = Bm.shape[1]
    D, z, delta_bias = _f32(D), _f32(z), _f32(delta_bias)
    du, dd = np.empty_like(u), np.empty_like(u)
    dA = np.empty_like(A)
    dB, dC = np.empty_like(Bm), np.empty_like(Cm)
    dD = np.empty(D_, np.float32)
    dbias = np.empty(D_, np.float32)
    dz = np.empty_like(u) if z is not None else None
    i64 = ctypes.c_int64
    rc = lib().oracle_selective_scan_bwd(_p(u), _p(delta), _p(A), _p(Bm), _p(Cm), _p(D), _p(z),
                                         _p(delta_bias), _p(dout), int(bool(delta_softplus)), i64(B_),
                                         i64(D_), i64(L), i64(N), i64(G), _p(du), _p(dd), _p(dA), _p(dB),
                                         _p(dC), _p(dD), _p(dz), _p(dbias))
    assert rc == 0, rc
    if squeeze:
        dB, dC = dB[:, 0], dC[:, 0]
    return dict(du=du, ddelta=dd, dA=dA, dB=dB, dC=dC, dD=dD if D is not None else None, dz=dz,
                ddelta_bias=dbias if delta_bias is not None else None)


def causal_conv1d_fwd(x, w, bias=None, silu=False):
    """causal_conv1d_interface.py:49-65."""
    x, w, bias = _f32(x), _f32(w), _f32(bias)
    B_, D_, L = x.shape
    out = np.empty_like(x)
    i64 = ctypes.c_int64
    rc = lib().oracle_causal_conv1d_fwd(_p(x), _p(w), _p(bias), int(bool(silu)), i64(B_), i64(D_), i64(L),
                                        i64(w.shape[1]), _p(out))
    assert rc == 0
    return out


def causal_conv1d_bwd(x, w, bias, dout, silu=False):
    x, w, bias, dout = _f32(x), _f32(w), _f32(bias), _f32(dout)
    B_, D_, L = x.shape
    dx, dw = np.empty_like(x), np.empty_like(w)
    db = np.empty(D_, np.float32)
    i64 = ctypes.c_int64
    rc = lib().oracle_causal_conv1d_bwd(_p(x), _p(w), _p(bias), _p(dout), int(bool(silu)), i64(B_), i64(D_),
                                        i64(L), i64(w.shape[1]), _p(dx), _p(dw), _p(db))
    assert rc == 0
    return dx, dw, (db if bias is not None else None)


ORDER_ROWMAJOR, ORDER_FLIP, ORDER_NSLICES, ORDER_TWOROW = 0, 1, 2, 3


def scan_order_index(order: int, H: int, W: int, nslices: int = 1) -> np.ndarray:
    """idx[l] = flat source position (h*W+w) of token l (int64, exact)."""
    idx = np.empty(H * W, np.int64)
    i64 = ctypes.c_int64
    rc = lib().oracle_scan_order_index(int(order), i64(H), i64(W), i64(nslices), _p(idx))
    if rc != 0:
        raise ValueError(f"scan_order_index: bad arguments (rc={rc})")
    return idx
