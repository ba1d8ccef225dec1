"""ctypes binding of oracle/_build/libtatva_oracle.so (TEST INFRASTRUCTURE ONLY, see tatva_oracle.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "libtatva_oracle.so")
_KIND = {"tri3": 0, "tet4": 1, "hex8": 2}
_MAT = {"linear_elastic": 0, "neo_hookean": 1}
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_PATH)
        _lib.oracle_fused.restype = C.c_int
        _lib.oracle_fused.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int64, C.c_int64] + [C.c_void_p] * 5
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_fused_pf.restype = C.c_int
        _lib.oracle_fused_pf.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 5
        _lib.oracle_blocks.restype = C.c_int
        _lib.oracle_blocks.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64] + [C.c_void_p] * 4
        _lib.oracle_set_num_threads.restype = None
        _lib.oracle_set_num_threads.argtypes = [C.c_int]
    return _lib


def num_threads() -> int:
    return _load().oracle_num_threads()


def set_num_threads(n: int) -> None:
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline legs ask for all host cores explicitly."""
    _load().oracle_set_num_threads(int(n))


def _run_pf(kind, mode, params, coords, conn, s, t=None):
    """Two-field (u, phi) law: s, t (N,4) = [ux,uy,uz,phi]; params = (mu, lambda, Gc, ell, k)."""
    L = _load()
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    s = np.ascontiguousarray(s, dtype=np.float64).reshape(-1, 4)
    tt = np.ascontiguousarray(t, dtype=np.float64).reshape(-1, 4) if t is not None else None
    prm = np.ascontiguousarray(params, dtype=np.float64)
    out = np.zeros(1) if mode == 0 else np.empty_like(s)
    rc = L.oracle_fused_pf(_KIND[kind], mode, prm.ctypes.data, coords.shape[0], conn.shape[0], coords.ctypes.data, conn.ctypes.data,
                           s.ctypes.data, tt.ctypes.data if tt is not None else None, out.ctypes.data)
    if rc != 0:
        raise ValueError("oracle_fused_pf: unsupported element")
    return float(out[0]) if mode == 0 else out


def energy_pf(kind, params, coords, conn, s):
    return _run_pf(kind, 0, params, coords, conn, s)


def residual_pf(kind, params, coords, conn, s):
    return _run_pf(kind, 1, params, coords, conn, s)


def hvp_pf(kind, params, coords, conn, s, t):
    return _run_pf(kind, 2, params, coords, conn, s, t)


def _run(kind, material, mode, params, coords, conn, u, v=None):
    L = _load()
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    u = np.ascontiguousarray(u, dtype=np.float64)
    vv = np.ascontiguousarray(v, dtype=np.float64) if v is not None else None
    out = np.zeros(1) if mode == 0 else np.empty_like(u)
    rc = L.oracle_fused(
        _KIND[kind], _MAT[material], mode, float(params[0]), float(params[1]), coords.shape[0], conn.shape[0],
        coords.ctypes.data, conn.ctypes.data, u.ctypes.data, vv.ctypes.data if vv is not None else None, out.ctypes.data,
    )
    if rc != 0:
        raise ValueError("oracle_fused: unsupported element/material")
    return float(out[0]) if mode == 0 else out


def energy(kind, params, coords, conn, u, material="neo_hookean"):
    return _run(kind, material, 0, params, coords, conn, u)


def residual(kind, params, coords, conn, u, material="neo_hookean"):
    return _run(kind, material, 1, params, coords, conn, u)


def hvp(kind, params, coords, conn, u, v, material="neo_hookean"):
    return _run(kind, material, 2, params, coords, conn, u, v)


def _blocks(kind, what, coords, conn, arr, nv, out_shape):
    L = _load()
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    a = np.ascontiguousarray(arr, dtype=np.float64) if arr is not None else None
    out = np.empty(out_shape, dtype=np.float64)
    rc = L.oracle_blocks(_KIND[kind], what, nv, coords.shape[0], conn.shape[0], coords.ctypes.data, conn.ctypes.data, a.ctypes.data if a is not None else None, out.ctypes.data)
    if rc != 0:
        raise ValueError("oracle_blocks: unsupported arguments")
    return out


_NQ = {"tri3": 1, "tet4": 1, "hex8": 8}


def op_grad(kind, coords, conn, u):
    """Operator.grad at the config sizes: (E, Q, nv, dim) (tatva/operator.py:379-397)."""
    u = np.asarray(u).reshape(len(coords), -1)
    return _blocks(kind, 0, coords, conn, u, u.shape[1], (len(conn), _NQ[kind], u.shape[1], np.shape(coords)[1]))


def op_grad_adjoint(kind, coords, conn, g):
    """Transpose of op_grad: (E, Q, nv, dim) -> (N, nv)."""
    g = np.asarray(g)
    return _blocks(kind, 1, coords, conn, g, g.shape[2], (len(coords), g.shape[2]))


def op_integration_weights(kind, coords, conn):
    return _blocks(kind, 2, coords, conn, None, 1, (len(conn), _NQ[kind]))


def op_gather(kind, coords, conn, u):
    u = np.asarray(u).reshape(len(coords), -1)
    return _blocks(kind, 3, coords, conn, u, u.shape[1], (len(conn), np.shape(conn)[1], u.shape[1]))
