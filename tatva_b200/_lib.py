"""ctypes binding of libtatva_b200.so (the C ABI declared in include/tatva_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtatva_b200.so")

# enums mirrored from include/tatva_b200.h
TRI3, TET4, HEX8, QUAD4, TRI6, QUAD8, LINE2, LINE3 = 0, 1, 2, 3, 4, 5, 6, 7
LINEAR_ELASTIC, NEO_HOOKEAN, NEO_HOOKEAN_PHASE_FIELD = 0, 1, 2
USER_LAW_BASE = 1000
PLAN_CACHE_WEIGHTS = 1
ABI_VERSION = 9  # == TATVA_B200_ABI_VERSION in include/tatva_b200.h
VARIANT_DEFAULT, VARIANT_GENERIC, VARIANT_MODAL = 0, 1, 2

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
vp = C.c_void_p

# name -> (restype, argtypes); every symbol declared in the header
SIGNATURES = {
    "tatva_error_string": (C.c_char_p, [C.c_int]),
    "tatva_abi_version": (C.c_int, []),
    "tatva_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "tatva_plan_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int64, C.c_int64, vp, vp, C.c_int, vp]),
    "tatva_plan_destroy": (C.c_int, [vp]),
    "tatva_plan_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), c_i64p, c_i64p]),
    "tatva_plan_rebind": (C.c_int, [vp, vp, vp]),
    "tatva_plan_set_variant": (C.c_int, [vp, C.c_int]),
    "tatva_law_register": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "tatva_law_compile_log": (C.c_int, [C.c_char_p, C.c_int]),
    "tatva_plan_set_quadrature": (C.c_int, [vp, C.c_int, c_f64p, c_f64p, vp]),
    "tatva_plan_set_tiles": (C.c_int, [vp, vp, vp, vp, C.c_int]),
    "tatva_halo_exchange": (C.c_int, [vp, vp, vp, c_i64p, vp, vp, c_i64p, vp, vp, C.c_int, vp]),
    "tatva_nccl_version": (C.c_int, [C.POINTER(C.c_int)]),
    "tatva_halo_comm_unique_id": (C.c_int, [vp]),
    "tatva_halo_comm_create": (C.c_int, [C.POINTER(vp), vp, C.c_int, C.c_int]),
    "tatva_halo_comm_destroy": (C.c_int, [vp]),
    "tatva_zero_release": (C.c_int, [vp, C.c_int64, vp]),
    "tatva_plan_cache_geometry": (C.c_int, [vp, C.c_int, vp]),
    "tatva_plan_set_node_schedule": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "tatva_plan_set_point_grid": (C.c_int, [vp, C.c_int, C.c_int, c_f64p, c_f64p, vp, vp]),
    "tatva_op_grad": (C.c_int, [vp, vp, C.c_int, vp, vp]),
    "tatva_op_grad_adjoint": (C.c_int, [vp, vp, C.c_int, vp, vp]),
    "tatva_op_eval": (C.c_int, [vp, vp, C.c_int, vp, vp]),
    "tatva_op_eval_adjoint": (C.c_int, [vp, vp, C.c_int, vp, vp]),
    "tatva_op_integration_weights": (C.c_int, [vp, vp, vp]),
    "tatva_op_integrate_quad": (C.c_int, [vp, vp, C.c_int, vp, vp]),
    "tatva_op_interpolate": (C.c_int, [vp, vp, C.c_int, vp, C.c_int64, vp, vp, vp]),
    "tatva_op_gather": (C.c_int, [vp, vp, C.c_int, vp, vp]),
    "tatva_op_gather_adjoint": (C.c_int, [vp, vp, C.c_int, vp, vp]),
    "tatva_op_sum_rows": (C.c_int, [vp, vp, C.c_int64, C.c_int, vp, vp]),
    "tatva_energy": (C.c_int, [vp, C.c_int, c_f64p, C.c_int, vp, vp, vp]),
    "tatva_residual": (C.c_int, [vp, C.c_int, c_f64p, C.c_int, vp, vp, vp]),
    "tatva_hvp": (C.c_int, [vp, C.c_int, c_f64p, C.c_int, vp, vp, vp, vp]),
    "tatva_hessian_diag": (C.c_int, [vp, C.c_int, c_f64p, C.c_int, vp, vp, vp]),
    "tatva_hvp_lifted": (C.c_int, [vp, C.c_int, c_f64p, C.c_int, vp, vp, vp, C.c_int64, vp, vp]),
    "tatva_hvp_elems": (C.c_int, [vp, C.c_int, c_f64p, C.c_int, vp, vp, vp, C.c_int64, C.c_int64, C.c_int, vp]),
    "tatva_residual_elems": (C.c_int, [vp, C.c_int, c_f64p, C.c_int, vp, vp, C.c_int64, C.c_int64, C.c_int, vp]),
    "tatva_csr_assemble": (C.c_int, [vp, C.c_int, c_f64p, C.c_int, vp, vp, vp, C.c_int64, vp, vp]),
    "tatva_csr_assemble_sym": (C.c_int, [vp, C.c_int, c_f64p, C.c_int, vp, vp, vp, vp, C.c_int64, vp, vp]),
    "tatva_csr_assemble_rows": (C.c_int, [vp, C.c_int, c_f64p, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "tatva_halo_pack": (C.c_int, [vp, vp, C.c_int64, vp, vp]),
    "tatva_halo_unpack_set": (C.c_int, [vp, vp, C.c_int64, vp, vp]),
    "tatva_halo_unpack_add": (C.c_int, [vp, vp, C.c_int64, vp, vp]),
    "tatva_peer_pull": (C.c_int, [vp, C.c_int64, C.c_int64, vp, vp, vp, vp]),
    "tatva_peer_push_add": (C.c_int, [vp, C.c_int64, C.c_int64, vp, vp, vp, vp]),
    "tatva_lift": (C.c_int, [vp, vp, vp, vp, C.c_int64, vp, vp]),
    "tatva_reduce_adjoint": (C.c_int, [vp, vp, vp, C.c_int64, vp, vp]),
    "tatva_cg_dot": (C.c_int, [vp, vp, C.c_int64, vp, vp, C.c_int, vp]),
    "tatva_cg_after_matvec": (C.c_int, [vp, vp, vp, vp, C.c_int64, vp, vp, vp]),
    "tatva_cg_update": (C.c_int, [vp, vp, vp, vp, vp, C.c_int64, vp, vp, vp]),
    "tatva_cg_direction": (C.c_int, [vp, vp, vp, C.c_int64, vp, vp]),
    "tatva_pcg_reciprocal": (C.c_int, [vp, C.c_int64, vp, vp]),
    "tatva_pcg_start": (C.c_int, [vp, vp, vp, C.c_int64, vp, vp, vp]),
    "tatva_pcg_after_matvec": (C.c_int, [vp, vp, vp, vp, vp, C.c_int64, vp, vp, vp]),
    "tatva_host_pattern_from_mesh": (C.c_int, [c_i32p, C.c_int64, C.c_int, C.c_int64, C.c_int, c_i32p, c_i32p, c_i64p]),
    "tatva_host_distance2_colors": (C.c_int, [c_i32p, c_i32p, C.c_int64, c_i32p, c_i32p]),
    "tatva_host_pattern_from_element_dofs": (C.c_int, [c_i32p, C.c_int64, C.c_int, c_i32p, C.c_int64, C.c_int64, c_i32p, c_i32p, C.POINTER(C.c_int64)]),
    "tatva_host_node_to_elements": (C.c_int, [c_i32p, C.c_int64, C.c_int, C.c_int64, c_i32p, c_i32p]),
    "tatva_host_build_tiles": (C.c_int, [c_i32p, C.c_int64, C.c_int, C.c_int, c_i32p, c_i32p, C.POINTER(C.c_uint16), c_i32p]),
    "tatva_host_node_schedule": (C.c_int, [c_i32p, C.c_int64, C.c_int, C.c_int, c_i32p, C.POINTER(C.c_uint8), c_i32p, c_i64p, c_i64p, c_i32p, c_i32p, C.POINTER(C.c_uint16)]),
    "tatva_host_csr_element_positions": (C.c_int, [c_i32p, C.c_int64, C.c_int, C.c_int, c_i32p, c_i32p, c_i32p]),
    "tatva_host_build_point_grid": (C.c_int, [c_f64p, C.c_int64, c_i32p, C.c_int64, C.c_int, C.c_int, C.c_int, c_f64p, c_f64p, c_i32p, c_i32p]),
    "tatva_probe_element": (C.c_int, [C.c_int, C.c_int, c_f64p, C.c_int, C.c_int, c_f64p, c_f64p, c_f64p, c_f64p]),
    "tatva_probe_hex8_nh_modal": (C.c_int, [C.c_int, c_f64p, c_f64p, c_f64p, C.c_double, C.c_double, c_f64p]),
    "tatva_probe_tet4_nh_ref": (C.c_int, [C.c_int, c_f64p, c_f64p, c_f64p, C.c_double, C.c_double, c_f64p]),
    "tatva_hvp_lifted_dot": (C.c_int, [vp, C.c_int, c_f64p, C.c_int, vp, vp, vp, C.c_int64, vp, C.c_int, vp, vp, C.c_int, C.c_int, vp]),
    "tatva_cg_after_dot": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_int64, vp, vp, vp]),
    "tatva_hvp_dot": (C.c_int, [vp, C.c_int, c_f64p, C.c_int, vp, vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, vp]),
    "tatva_host_csr_tile_schedule": (C.c_int, [c_i32p, C.c_int64, C.c_int, C.c_int, C.c_int, c_i32p, c_i32p, c_i32p, c_i64p, c_i64p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, C.POINTER(C.c_uint32)]),
    "tatva_csr_assemble_tiled": (C.c_int, [vp, C.c_int, c_f64p, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int64, vp, vp]),
    "tatva_fp64_peak_tflops": (C.c_int, [c_f64p, vp]),
}


class TatvaError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        from . import build as _build

        override = os.environ.get("TATVA_B200_LIB")  # measurement aid: load an experimental build of the same ABI
        if override:
            L = C.CDLL(override)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
            return _lib
        stale = os.path.exists(LIB_PATH) and _build.needs_build()
        if not os.path.exists(LIB_PATH) or stale:
            # not built yet (fresh checkout) or older than its sources: compile it now if nvcc is around (cheap when up
            # to date); there is no CPU fallback
            try:
                _build.build()
            except Exception as exc:  # noqa: BLE001
                if not os.path.exists(LIB_PATH):
                    raise TatvaError(
                        f"{LIB_PATH} not found and building it failed ({exc}); run `python -m tatva_b200.build` "
                        "(there is no CPU fallback)"
                    ) from exc
                # a box without nvcc (the GPU box uses the prebuilt library): the ABI check below decides
        L = C.CDLL(LIB_PATH)
        L.tatva_abi_version.restype = C.c_int
        have = L.tatva_abi_version()
        if have != ABI_VERSION:
            raise TatvaError(f"{LIB_PATH} has ABI version {have}, this package expects {ABI_VERSION}: rebuild with `python -m tatva_b200.build --force`")
        missing = [name for name in SIGNATURES if not hasattr(L, name)]
        if missing:
            raise TatvaError(f"{LIB_PATH} lacks {missing}: rebuild with `python -m tatva_b200.build --force`")
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def host_tables_only(what: str) -> None:
    """Gate of the few places that can run their INDEX algebra on host arrays (Lifter tables on NumPy vectors, exchange
    plans on CPU tensors over gloo): that mode exists for the world_size-2 `gloo` tests of the host logic, which opt in
    with TATVA_B200_HOST_TABLES=1 (tests/conftest.py).  Anywhere else a non-CUDA operand is an error, not a fallback."""
    if os.environ.get("TATVA_B200_HOST_TABLES") != "1":
        raise TatvaError(f"{what}: operands must be CUDA tensors (there is no CPU fallback; TATVA_B200_HOST_TABLES=1 enables the host index paths for the gloo tests)")


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = lib().tatva_error_string(code).decode()
        raise TatvaError(f"{what or 'tatva_b200 call'} failed: {msg} (code {code})")


_PARAMS = {}


def params_array(values) -> tuple[C.Array, int]:
    """ctypes view of a law's parameters; cached per value tuple (the hot calls pass it on every launch, and for the
    launch-sized configurations building it was a visible part of the call)."""
    key = tuple(float(v) for v in values)
    hit = _PARAMS.get(key)
    if hit is None:
        if len(_PARAMS) > 4096:
            _PARAMS.clear()
        hit = _PARAMS[key] = ((C.c_double * len(key))(*key), len(key))
    return hit
