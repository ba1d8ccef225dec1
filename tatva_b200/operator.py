"""Operator — the quadrature loop, mirror of tatva/operator.py on hand-written CUDA kernels.

Same constructor and methods as the reference (`Operator(mesh, element, batch_size, cache_weights)`;
`map`, `map_over_elements`, `eval`, `grad`, `integrate`, `integrate_per_element`,
`get_integration_weights`).  Arrays are torch CUDA float64 tensors (anything torch.as_tensor accepts
is moved to the operator's device).  Every method launches kernels of libtatva_b200.so through the
C ABI on torch's current stream; there is no CPU path.

Differentiation: the reference leaves residuals and Hessian-vector products to jax.grad / jax.jvp.
Here `grad`, `eval`, `integrate`, the gather of `map` and the fused `energy` are torch.autograd
Functions whose backward / forward-mode rules are the adjoint / tangent kernels, so
`torch.autograd.grad` (and double backward for H v) work on user energies, and the fused
`energy(material)` -> `residual(material)` -> `hvp(material)` chain is what autograd follows for the
laws in tatva_b200.materials.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Sequence

import numpy as np
import torch

from . import _lib
from .element import Element
from .mesh import Mesh


_HVP_CALLS = {"tatva_hvp", "tatva_hvp_elems", "tatva_hvp_dot"}
_FUSED_CALLS = {"tatva_energy", "tatva_residual", "tatva_hvp", "tatva_hvp_lifted", "tatva_hvp_lifted_dot", "tatva_hvp_dot", "tatva_hessian_diag", "tatva_csr_assemble", "tatva_csr_assemble_tiled", "tatva_csr_assemble_sym", "tatva_csr_assemble_rows"}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class Operator:
    def __init__(self, mesh: Mesh, element: Element, batch_size: int | None = None, cache_weights: bool = False, *, device=None, sort_elements: bool = False, stage_tiles: bool = False, node_schedule: bool | int | str = "auto", cache_geometry: bool = False):
        self.mesh = mesh
        self.element = element
        self.cache_weights = bool(cache_weights)
        self._check_init(mesh)
        if element.kind is None:
            raise NotImplementedError(f"{type(element).__name__} has no CUDA kernel (the eight reference elements are supported)")
        self._custom_rule = not getattr(element, "_default_rule", True)  # Element(quad_points, quad_weights), element/base.py:37-51
        if self._custom_rule and len(element.quad_weights) > 64:
            raise NotImplementedError("custom quadrature rules hold at most 64 points")
        if not torch.cuda.is_available():
            raise _lib.TatvaError("tatva_b200.Operator needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.coords = torch.as_tensor(_to_np_or_tensor(mesh.coords), dtype=torch.float64, device=self.device).contiguous()
        self.elements = torch.as_tensor(_to_np_or_tensor(mesh.elements), device=self.device).to(torch.int32).contiguous()
        self.n_nodes, self.dim = self.coords.shape
        self._line = getattr(element, "gradient_components", None) == 1  # arc-length gradient: no spatial axis
        expected = 2 if self._line else np.asarray(element.quad_points).shape[1]
        if self.dim != expected:
            raise ValueError(f"{type(element).__name__} needs {expected}-D node coordinates, the mesh has {self.dim}-D")
        self.gdim = 1 if self._line else self.dim
        self.n_elements, self.npe = self.elements.shape
        self.nq = len(element.quad_weights)
        self._batch_size_arg, self._sort_elements, self._node_schedule_arg = batch_size, bool(sort_elements), node_schedule
        # Hex8 x neo-Hookean HVP, opt-in: the mesh-only part of the Gauss-point arithmetic computed once and kept by the plan
        # (tatva_plan_cache_geometry, 512 bytes per element, built at the first such HVP).  Measured SLOWER on B200 at
        # config 3 (0.514 vs 0.396 ms: 21 % fewer FP64 instructions, but streaming 1.07 GB per application at 12 warps per
        # SM leaves too few bytes in flight to hide HBM latency), so the default re-derives the geometry every time.
        self._cache_geometry, self._geometry_cached = bool(cache_geometry), False
        self.batch_size = self.n_elements if batch_size is None else int(batch_size)  # operator.py:116-117
        self.quad_points = torch.as_tensor(element.quad_points, dtype=torch.float64, device=self.device)
        self._L = _lib.lib()
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(
                self._L.tatva_plan_create(
                    C.byref(handle), element.kind, self.n_nodes, self.n_elements, self.coords.data_ptr(),
                    self.elements.data_ptr(), _lib.PLAN_CACHE_WEIGHTS if cache_weights else 0, _stream(),
                ),
                "tatva_plan_create",
            )
        self._plan = handle
        if self._custom_rule:
            self._set_rule(handle)
        # Fused energy / residual / HVP / assembly are order-independent sums over elements, so they may run
        # on a locality-sorted copy of the connectivity (Morton order of centroids); the (E, Q, ...)-shaped
        # building blocks keep the caller's element order.
        self._plan_fused, self.elements_fused = self._plan, self.elements
        if sort_elements:
            from .mesh import locality_order

            perm = torch.as_tensor(locality_order(self.coords.cpu().numpy(), self.elements.cpu().numpy()), device=self.device)
            self.elements_fused = self.elements[perm].contiguous()
            h2 = C.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(self._L.tatva_plan_create(C.byref(h2), element.kind, self.n_nodes, self.n_elements, self.coords.data_ptr(), self.elements_fused.data_ptr(), 0, _stream()), "tatva_plan_create")
            self._plan_fused = h2
            if self._custom_rule:
                self._set_rule(h2)

        # Shared-memory staging tiles for gather-bound kernels (Tet4 x neo-Hookean): per CTA the unique nodes are
        # gathered once, coalesced, and elements read them through tile-local uint16 connectivity.
        self._point_grid = None
        self._tiles = None
        if stage_tiles:
            self._build_tiles()
        # Node schedule of the warp-cooperative residual / HVP kernels (Tri3, Tet4 with the default rule): one nodal-row
        # load per distinct node of a warp, one atomic add per distinct node of a 128-element tile.
        self._node_schedule = None
        # "auto" (default): Tet4 with the default rule, kept only when the element order has locality (>= 3 node references
        # per distinct node of a 128-element tile; measured on config 2: HVP 0.0535 -> 0.0472 ms, residual 0.0455 -> 0.0391;
        # on a shuffled element list the schedule loses and is dropped).  True forces it; an int > 1 also caps the
        # contributors per table entry (measured slower: more atomic adds).
        self._node_schedule_cap = 0
        if node_schedule == "auto":
            if element.kind == _lib.TET4 and not self._custom_rule:
                self._build_node_schedule(0)
                if self.node_schedule_stats["references_per_entry"] < 3.0:
                    self._drop_node_schedule()
        elif node_schedule:
            self._build_node_schedule(0 if node_schedule is True or int(node_schedule) == 1 else int(node_schedule))

    def _set_rule(self, plan):
        """Install the element's own quadrature rule in the plan: the generic kernels then evaluate the shape functions
        at these points (the specialised Hex8 / Tet4 kernels are default-rule only and are bypassed)."""
        pts = np.ascontiguousarray(self.element.quad_points, dtype=np.float64)
        wts = np.ascontiguousarray(self.element.quad_weights, dtype=np.float64)
        if pts.ndim != 2 or pts.shape[0] != wts.shape[0]:
            raise ValueError("quad_points must be (n_quad, reference dimension) and quad_weights (n_quad,)")
        with torch.cuda.device(self.device):
            _lib.check(self._L.tatva_plan_set_quadrature(plan, wts.shape[0], pts.ctypes.data_as(_lib.c_f64p), wts.ctypes.data_as(_lib.c_f64p), _stream()), "tatva_plan_set_quadrature")

    def _replace(self, **changes) -> "Operator":
        """A new Operator with the given constructor arguments changed (operator.py:497-504, `dataclasses.replace`);
        a new plan is created, nothing is shared with `self`."""
        kw = dict(mesh=self.mesh, element=self.element, batch_size=self._batch_size_arg, cache_weights=self.cache_weights, device=self.device, sort_elements=self._sort_elements, stage_tiles=self._tiles is not None, node_schedule=self._node_schedule_arg, cache_geometry=self._cache_geometry)
        unknown = set(changes) - set(kw)
        if unknown:
            raise TypeError(f"Operator._replace() got unexpected field(s): {sorted(unknown)}")
        kw.update(changes)
        return type(self)(kw.pop("mesh"), kw.pop("element"), kw.pop("batch_size"), kw.pop("cache_weights"), **kw)

    def _build_tiles(self):
        conn = np.ascontiguousarray(self.elements_fused.cpu().numpy(), dtype=np.int32)
        E, npe = conn.shape
        n_tiles = (E + 127) // 128
        tile_ptr = np.empty(n_tiles + 1, dtype=np.int32)
        mx = C.c_int32()
        i32 = lambda a: a.ctypes.data_as(_lib.c_i32p)  # noqa: E731
        _lib.check(self._L.tatva_host_build_tiles(i32(conn), E, npe, 128, i32(tile_ptr), None, None, C.byref(mx)), "tatva_host_build_tiles")
        tile_nodes = np.empty(int(tile_ptr[-1]), dtype=np.int32)
        local = np.empty((E, npe), dtype=np.uint16)
        _lib.check(self._L.tatva_host_build_tiles(i32(conn), E, npe, 128, i32(tile_ptr), i32(tile_nodes), local.ctypes.data_as(C.POINTER(C.c_uint16)), C.byref(mx)), "tatva_host_build_tiles")
        dev = lambda a: torch.as_tensor(a, device=self.device)  # noqa: E731
        self._tiles = (dev(tile_ptr), dev(tile_nodes), torch.as_tensor(local.view(np.int16), device=self.device), int(mx.value))
        tp, tn, tc, m = self._tiles
        _lib.check(self._L.tatva_plan_set_tiles(self._plan_fused, tp.data_ptr(), tn.data_ptr(), tc.data_ptr(), m), "tatva_plan_set_tiles")

    def _drop_node_schedule(self):
        _lib.check(self._L.tatva_plan_set_node_schedule(self._plan_fused, None, None, None, None, None, None), "tatva_plan_set_node_schedule")
        self._node_schedule = None

    def _build_node_schedule(self, cap: int = 0):
        self._node_schedule_cap = cap
        if self.npe > 8 or self._custom_rule:
            raise NotImplementedError("node_schedule: elements with at most 8 nodes and the default quadrature rule")
        conn = np.ascontiguousarray(self.elements_fused.cpu().numpy(), dtype=np.int32)
        E, npe = conn.shape
        n_warps, n_tiles = (E + 31) // 32, (E + 127) // 128
        i32 = lambda a: a.ctypes.data_as(_lib.c_i32p)  # noqa: E731
        u8, u16 = C.POINTER(C.c_uint8), C.POINTER(C.c_uint16)
        warp_nodes = np.empty(n_tiles * 128, dtype=np.int32)
        warp_local = np.empty(E * npe, dtype=np.uint8)
        ch_ptr = np.empty(n_tiles + 1, dtype=np.int32)
        n_ch, n_ell = C.c_int64(), C.c_int64()
        _lib.check(self._L.tatva_host_node_schedule(i32(conn), E, npe, cap, i32(warp_nodes), warp_local.ctypes.data_as(u8), i32(ch_ptr), C.byref(n_ch), C.byref(n_ell), None, None, None), "tatva_host_node_schedule")
        tn_node = np.empty(32 * n_ch.value, dtype=np.int32)
        ell_ptr = np.empty(n_ch.value + 1, dtype=np.int32)
        ell = np.empty(n_ell.value, dtype=np.uint16)
        _lib.check(self._L.tatva_host_node_schedule(i32(conn), E, npe, cap, None, None, i32(ch_ptr), C.byref(n_ch), C.byref(n_ell), i32(tn_node), i32(ell_ptr), ell.ctypes.data_as(u16)), "tatva_host_node_schedule")
        dev = lambda a: torch.as_tensor(a, device=self.device)  # noqa: E731
        # per-tile header {first chunk, chunks, first table entry, table entries}: one 16-byte load per tile
        hdr = np.stack([ch_ptr[:-1], np.diff(ch_ptr), ell_ptr[ch_ptr[:-1]], ell_ptr[ch_ptr[1:]] - ell_ptr[ch_ptr[:-1]]], axis=1).astype(np.int32)
        self._node_schedule = (dev(warp_nodes), dev(warp_local), dev(np.ascontiguousarray(hdr)), dev(tn_node), dev(ell_ptr), dev(ell.view(np.int16)))
        self.node_schedule_stats = dict(entries_per_tile=float((tn_node >= 0).sum()) / n_tiles, references_per_entry=float(E * npe) / max(1, int((tn_node >= 0).sum())), chunks_per_tile=float(n_ch.value) / n_tiles, direct_rows=int((warp_local == 255).sum()), distinct_per_warp=float((warp_nodes >= 0).sum()) / n_warps,
                                        ell_fill=float(E * npe) / max(1, n_ell.value))
        _lib.check(self._L.tatva_plan_set_node_schedule(self._plan_fused, *(t.data_ptr() for t in self._node_schedule)), "tatva_plan_set_node_schedule")

    def __del__(self):
        L = getattr(self, "_L", None)
        plans = {id(p): p for p in (getattr(self, "_plan", None), getattr(self, "_plan_fused", None)) if p is not None}
        for plan in plans.values():
            try:
                L.tatva_plan_destroy(plan)
            except Exception:
                pass
        self._plan = self._plan_fused = None

    # operator.py:132-170
    @staticmethod
    def _check_init(mesh):
        coords, elements = mesh.coords, mesh.elements
        if coords.ndim != 2:
            raise ValueError("Mesh coordinates must be a 2D array shaped (n_nodes, n_dim).")
        if coords.shape[0] == 0:
            raise ValueError("Mesh must contain at least one node.")
        if elements.ndim != 2:
            raise ValueError("Mesh elements must be a 2D array shaped (n_elements, n_nodes_per_element).")
        if elements.shape[0] == 0:
            raise ValueError("Mesh must contain at least one element.")
        el = _to_np_or_tensor(elements)
        is_int = (not el.dtype.is_floating_point and el.dtype != torch.bool) if isinstance(el, torch.Tensor) else np.issubdtype(el.dtype, np.integer)
        if not is_int:
            raise TypeError("Mesh element connectivity must contain integer indices.")
        if int(el.min()) < 0:
            raise ValueError("Mesh element connectivity contains negative node indices.")
        if int(el.max()) >= coords.shape[0]:
            raise ValueError("Mesh element connectivity references nodes outside the mesh coordinates array.")

    # -- helpers ------------------------------------------------------------------------------
    def _as_dev(self, x) -> torch.Tensor:
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
        if t.dtype != torch.float64 or t.device != self.device:
            t = t.to(device=self.device, dtype=torch.float64)
        return t

    # Hex8 x neo-Hookean HVP kernels kept for measurement; 0 = default.  8 / 9 are timing experiments
    # (no scatter / no gather) and do not compute the HVP.
    HEX8_NH_HVP_VARIANTS = (0, 1, 2, 3, 15, 16, 17, 20, 22, 23, 25, 26, 27, 28, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49)

    def hvp_variants(self) -> tuple:
        return self.HEX8_NH_HVP_VARIANTS

    def set_variant(self, variant: int) -> None:
        """Select the Hex8 neo-Hookean HVP kernel variant (benchmarking aid)."""
        for plan in {id(p): p for p in (self._plan, self._plan_fused)}.values():
            _lib.check(self._L.tatva_plan_set_variant(plan, int(variant)), "tatva_plan_set_variant")

    def _call(self, name, *args):
        plan = self._plan_fused if name in _FUSED_CALLS else self._plan
        if name in _HVP_CALLS and args:
            self._ensure_geometry(args[0])
        if torch.cuda.current_device() == self._dev_index:  # the usual case (one process per GPU): no device switch needed
            _lib.check(getattr(self._L, name)(plan, *args, _stream()), name)
            return
        with torch.cuda.device(self.device):
            _lib.check(getattr(self._L, name)(plan, *args, _stream()), name)

    def _ensure_geometry(self, material_id) -> None:
        """Build the plan's geometry cache before the first Hex8 x neo-Hookean HVP (set-up work: it allocates)."""
        if self._cache_geometry and not self._geometry_cached and self.element.kind == _lib.HEX8 and not self._custom_rule and material_id == _lib.NEO_HOOKEAN:
            with torch.cuda.device(self.device):
                _lib.check(self._L.tatva_plan_cache_geometry(self._plan_fused, 1, _stream()), "tatva_plan_cache_geometry")
            self._geometry_cached = True

    # raw (non-differentiable) kernel wrappers; nodal arrays are (N, nv) contiguous
    def _k_grad(self, u2):
        out = torch.empty((self.n_elements, self.nq, u2.shape[1], self.gdim), dtype=torch.float64, device=self.device)
        self._call("tatva_op_grad", u2.data_ptr(), u2.shape[1], out.data_ptr())
        return out

    def _k_grad_adj(self, g4):
        y = torch.empty((self.n_nodes, g4.shape[2]), dtype=torch.float64, device=self.device)
        self._call("tatva_op_grad_adjoint", g4.data_ptr(), g4.shape[2], y.data_ptr())
        return y

    def _k_eval(self, u2):
        out = torch.empty((self.n_elements, self.nq, u2.shape[1]), dtype=torch.float64, device=self.device)
        self._call("tatva_op_eval", u2.data_ptr(), u2.shape[1], out.data_ptr())
        return out

    def _k_eval_adj(self, g3):
        y = torch.empty((self.n_nodes, g3.shape[2]), dtype=torch.float64, device=self.device)
        self._call("tatva_op_eval_adjoint", g3.data_ptr(), g3.shape[2], y.data_ptr())
        return y

    def _k_gather(self, u2):
        out = torch.empty((self.n_elements, self.npe, u2.shape[1]), dtype=torch.float64, device=self.device)
        self._call("tatva_op_gather", u2.data_ptr(), u2.shape[1], out.data_ptr())
        return out

    def _k_gather_adj(self, g3):
        y = torch.empty((self.n_nodes, g3.shape[2]), dtype=torch.float64, device=self.device)
        self._call("tatva_op_gather_adjoint", g3.data_ptr(), g3.shape[2], y.data_ptr())
        return y

    def _k_integrate_quad(self, v3):
        out = torch.empty((self.n_elements, v3.shape[2]), dtype=torch.float64, device=self.device)
        self._call("tatva_op_integrate_quad", v3.data_ptr(), v3.shape[2], out.data_ptr())
        return out

    def _k_sum_rows(self, a2):
        nv = a2.shape[1]
        if nv > 64:  # the kernel sums up to 64 columns per launch; wider value shapes go in column chunks
            return torch.cat([self._k_sum_rows(a2[:, c : c + 64].contiguous()) for c in range(0, nv, 64)])
        out = torch.empty((nv,), dtype=torch.float64, device=self.device)
        self._call("tatva_op_sum_rows", a2.data_ptr(), a2.shape[0], nv, out.data_ptr())
        return out

    # -- reference API ------------------------------------------------------------------------
    def get_integration_weights(self) -> torch.Tensor:
        """det(J) * w per (element, quad point) — operator.py:172-192 (no abs)."""
        out = torch.empty((self.n_elements, self.nq), dtype=torch.float64, device=self.device)
        self._call("tatva_op_integration_weights", out.data_ptr())
        return out

    def grad(self, nodal_values) -> torch.Tensor:
        """(N, *v) -> (E, Q, *v, dim) — operator.py:379-397; line elements: (E, Q, *v) (element/base.py:169-173)."""
        u = self._as_dev(nodal_values)
        vshape = tuple(u.shape[1:])
        out = _LinearOp.apply(u.reshape(self.n_nodes, -1), self, "grad")
        return out.reshape((self.n_elements, self.nq) + vshape + (() if self._line else (self.dim,)))

    def eval(self, nodal_values) -> torch.Tensor:
        """(N, *v) -> (E, Q, *v) — operator.py:358-377."""
        u = self._as_dev(nodal_values)
        vshape = tuple(u.shape[1:])
        out = _LinearOp.apply(u.reshape(self.n_nodes, -1), self, "eval")
        return out.reshape((self.n_elements, self.nq) + vshape)

    def integrate_per_element(self, arg) -> torch.Tensor:
        """operator.py:321-356, with the reference's dispatch on arg.shape[0]."""
        if isinstance(arg, (int, float)) and not isinstance(arg, bool):
            # operator.py:335 gathers jnp.array([arg]) with clamped indices: the constant field
            vals = torch.full((self.n_nodes,), float(arg), dtype=torch.float64, device=self.device)
            return self.integrate_per_element(self.eval(vals))
        a = self._as_dev(arg)
        if a.shape[0] == self.n_elements:
            vshape = tuple(a.shape[2:])
            out = _LinearOp.apply(a.reshape(self.n_elements, self.nq, -1), self, "integrate")
            return out.reshape((self.n_elements,) + vshape)
        return self.integrate_per_element(self.eval(a))

    def integrate(self, arg) -> torch.Tensor:
        """operator.py:307-319: sum over elements of integrate_per_element."""
        res = self.integrate_per_element(arg)
        vshape = tuple(res.shape[1:])
        out = _LinearOp.apply(res.reshape(self.n_elements, -1), self, "sum")
        return out.reshape(vshape)

    def _gather(self, v) -> torch.Tensor:
        t = self._as_dev(v)
        vshape = tuple(t.shape[1:])
        out = _LinearOp.apply(t.reshape(self.n_nodes, -1), self, "gather")
        return out.reshape((self.n_elements, self.npe) + vshape)

    def map(self, func: Callable, *, element_quantity: Sequence[int] = ()) -> Callable:
        """operator.py:225-266: func(xi, *element_values, **kw) at every quadrature point of every
        element.  Nodal arguments are gathered by the CUDA gather kernel; `func` itself is user torch
        code, vectorised with torch.vmap in chunks of `batch_size` elements."""

        def _mapped(*values, **kwargs):
            xs = tuple(self._as_dev(v) if i in element_quantity else self._gather(v) for i, v in enumerate(values))

            def at_each_element(*el_values):
                return torch.vmap(lambda xi: func(xi, *el_values, **kwargs))(self.quad_points)

            return torch.vmap(at_each_element, chunk_size=self.batch_size)(*xs)

        return _mapped

    def map_over_elements(self, func: Callable, *, element_quantity: Sequence[int] = ()) -> Callable:
        """operator.py:268-305."""

        def _mapped(*values, **kwargs):
            xs = tuple(self._as_dev(v) if i in element_quantity else self._gather(v) for i, v in enumerate(values))
            return torch.vmap(lambda *el: func(*el, **kwargs), chunk_size=self.batch_size)(*xs)

        return _mapped

    # -- post-processing: point evaluation and L2 projection --------------------------------------
    def quads(self) -> torch.Tensor:
        """Quadrature points in physical coordinates, (E, Q, dim) — operator.py:506-516."""
        return self.eval(self.coords)

    def interpolate(self, arg, points) -> torch.Tensor:
        """Nodal values (N, *v) at physical points (P, 2) -> (P, *v) — operator.py:399-463.  The containing element is
        found as mesh.find_containing_polygons does (mesh.py:294-388), the reference point by ONE Newton step from
        the first quadrature point (exact for affine elements), all inside one kernel.  Raises RuntimeError if a
        point lies outside the mesh, like the reference outside a trace."""
        u = self._as_dev(arg)
        pts = self._as_dev(points).contiguous()
        if pts.ndim != 2 or pts.shape[1] != 2 or self.dim != 2 or self._line:
            raise NotImplementedError("interpolate: plane elements and (P, 2) points, as in the reference (mesh.py:303)")
        vshape = tuple(u.shape[1:])
        u2 = u.reshape(self.n_nodes, -1).contiguous()
        if self._point_grid is None:
            self._build_point_grid()
        out = torch.empty((pts.shape[0], u2.shape[1]), dtype=torch.float64, device=self.device)
        elem = torch.empty(pts.shape[0], dtype=torch.int32, device=self.device)
        if pts.shape[0]:
            self._call("tatva_op_interpolate", u2.data_ptr(), u2.shape[1], pts.data_ptr(), pts.shape[0], out.data_ptr(), elem.data_ptr())
            if bool((elem < 0).any()):
                raise RuntimeError("Some points are outside the mesh, revise the points")
        return out.reshape((pts.shape[0],) + vshape)

    def _build_point_grid(self):
        """Uniform background grid for point location, ~1 element per bin, built once on the host (C++) and attached
        to the plan; the search then visits only the elements whose bounding box overlaps the point's bin."""
        coords = np.ascontiguousarray(self.coords.cpu().numpy(), dtype=np.float64)
        conn = np.ascontiguousarray(self.elements.cpu().numpy(), dtype=np.int32)
        side = int(min(4096, max(1, round(np.sqrt(self.n_elements)))))
        lo, inv = np.zeros(2), np.zeros(2)
        ptr = np.zeros(side * side + 1, dtype=np.int32)
        f64 = lambda a: a.ctypes.data_as(_lib.c_f64p)  # noqa: E731
        i32 = lambda a: a.ctypes.data_as(_lib.c_i32p)  # noqa: E731
        args = (f64(coords), self.n_nodes, i32(conn), self.n_elements, self.npe, side, side, f64(lo), f64(inv), i32(ptr))
        _lib.check(self._L.tatva_host_build_point_grid(*args, None), "tatva_host_build_point_grid")
        elems = np.empty(max(int(ptr[-1]), 1), dtype=np.int32)
        _lib.check(self._L.tatva_host_build_point_grid(*args, i32(elems)), "tatva_host_build_point_grid")
        d_ptr, d_elems = torch.as_tensor(ptr, device=self.device), torch.as_tensor(elems, device=self.device)
        self._point_grid = (d_ptr, d_elems, lo, inv, side)  # keeps the device views alive
        _lib.check(self._L.tatva_plan_set_point_grid(self._plan, side, side, f64(lo), f64(inv), d_ptr.data_ptr(), d_elems.data_ptr()), "tatva_plan_set_point_grid")

    def project(self, field, colored_matrix=None, lifter=None, *, tol: float = 1e-12, maxiter: int = 2000, use_graph: bool = False) -> torch.Tensor:
        """L2 projection of a quadrature field (E, Q, *v) onto the nodal space, (N, *v) — operator.py:518-554,
        utils.py:118-258: M x = b with M_ab = int N_a N_b, b_a = int N_a f.  The reference assembles M through
        sparse.jacfwd and calls a direct sparse solve; here M is applied matrix-free (eval kernel, weights, eval-adjoint
        kernel) inside the device-resident Jacobi-preconditioned CG, all components at once.  `colored_matrix` only
        selects scalar (multi right-hand side) or coupled layout, which give the same nodal values; a `lifter` pins its
        Fixed DOFs (utils.py:233-236, :193-201) — constraints that tie DOFs together are not supported here.  As in the
        reference, M is integrated with the element's own rule: rules that under-integrate N_a N_b (the one-point
        Tri3 / Tetrahedron4 rules) give a singular M."""
        from .solver import ConjugateGradient

        f = self._as_dev(field)
        if f.ndim < 2 or tuple(f.shape[:2]) != (self.n_elements, self.nq):
            raise ValueError(f"field must be shaped (n_elements, n_quad, ...) = ({self.n_elements}, {self.nq}, ...)")
        vshape = tuple(f.shape[2:])
        K = int(np.prod(vshape)) if vshape else 1
        if lifter is not None:
            dim_s = max(int(lifter.size) // self.n_nodes, 1)
        else:
            dim_s = 1 if colored_matrix is None else max(int(colored_matrix.shape[0]) // self.n_nodes, 1)
        if lifter is not None and colored_matrix is not None and colored_matrix.shape[0] != lifter.size_reduced:
            raise ValueError(f"Colored matrix size does not match lifter reduced size. Expected {lifter.size_reduced}, got {colored_matrix.shape[0]}")
        if dim_s > 1 and K != dim_s:
            raise ValueError(f"a coupled projection with {dim_s} DOFs per node needs a field with {dim_s} components, got {K}")
        W = self.get_integration_weights()
        b = self._k_eval_adj((f.reshape(self.n_elements, self.nq, K) * W[:, :, None]).contiguous())  # (N, K)
        # Jacobi: the consistent diagonal M_aa = sum_q W N_a^2 (the lumped mass of quadratic elements is not positive)
        Nq = torch.as_tensor(np.stack([self.element.shape_function(xi) for xi in self.element.quad_points]), dtype=torch.float64, device=self.device)
        diag = self._k_gather_adj(torch.einsum("eq,qn->en", W, Nq * Nq)[:, :, None].contiguous())  # (N, 1)
        mask, shift = None, None
        if lifter is not None:
            want = self.n_nodes * dim_s
            if lifter.size != want:
                raise ValueError(f"lifter.size = {lifter.size}, expected {want}")
            m = lifter.dof_map()
            if (m >= 0).sum() != lifter.size_reduced or not np.array_equal(m[lifter.free_dofs], np.arange(lifter.size_reduced)):
                raise NotImplementedError("project: only lifters made of Fixed constraints are supported")
            free = torch.as_tensor(m >= 0, device=self.device)
            const = torch.as_tensor(np.asarray(lifter.lift_from_zeros(np.zeros(lifter.size_reduced))), dtype=torch.float64, device=self.device)
            if dim_s == 1:
                mask, shift = free[:, None].expand(self.n_nodes, K), const[:, None].expand(self.n_nodes, K)
            else:
                mask, shift = free.reshape(self.n_nodes, K), const.reshape(self.n_nodes, K)
            mask = mask.to(torch.float64).contiguous()
        n = self.n_nodes * K

        def matvec(x, out):
            x2 = x.view(self.n_nodes, K)
            if mask is not None:
                x2 = x2 * mask
            y = self._k_eval_adj((self._k_eval(x2.contiguous()) * W[:, :, None]).contiguous())
            if mask is not None:
                y = y * mask + x.view(self.n_nodes, K) * (1.0 - mask)  # identity on the pinned DOFs keeps the system SPD
            out.copy_(y.reshape(-1))
            return out

        rhs = b if mask is None else b * mask
        cg = ConjugateGradient(matvec, n, self.device, use_graph=use_graph, jacobi=True)  # graph capture costs ~0.5 s: only pays for very long solves
        cg.set_diagonal(diag.expand(self.n_nodes, K).contiguous().reshape(-1))
        x, info = cg.solve(rhs.reshape(-1).contiguous(), tol=tol, maxiter=maxiter, check_every=5)
        if not info["converged"]:
            raise RuntimeError(f"project: CG did not converge ({info})")
        x = x.view(self.n_nodes, K)
        if mask is not None:
            x = x * mask + shift * (1.0 - mask)
        return x.reshape((self.n_nodes,) + vshape)

    # -- fused energy / residual / HVP -----------------------------------------------------------
    def _fused_shape(self, material, u):
        dpn = material.dofs_per_node(self.dim)
        t = self._as_dev(u)
        if t.numel() != self.n_nodes * dpn:
            raise ValueError(f"expected {self.n_nodes}x{dpn} nodal values, got shape {tuple(t.shape)}")
        return t, dpn

    def energy(self, material) -> Callable:
        """u -> E(u) = op.integrate(psi(op.grad(u))) in one kernel; differentiable (grad -> residual kernel,
        second derivative -> HVP kernel)."""
        return lambda u: _Energy.apply(self._fused_shape(material, u)[0], self, material)

    def residual(self, material) -> "FusedResidual":
        """u -> dE/du (same shape as u) = jax.grad(E)(u) of the reference.  The returned callable is
        recognised by tatva_b200.sparse.jacfwd, which then assembles dR/du with one kernel."""
        return FusedResidual(self, material)

    def hvp(self, material) -> Callable:
        """(u, v) -> H(u) v = jax.jvp(jax.grad(E), (u,), (v,))[1] of the reference (sparse/base.py:264)."""

        def _hvp(u, v):
            return self._raw_hvp(material, self._as_dev(u), self._as_dev(v))

        return _hvp

    def _raw_hvp_lifted(self, material, u_full, v_red, dof_map, out):
        """out = reduce_adjoint(H(u_full) lift_0(v_red)) in one kernel (`tatva_hvp_lifted`); all arguments are
        contiguous CUDA tensors, `dof_map` = Lifter.dof_map(device)."""
        prm, n = _lib.params_array(material.params())
        self._call("tatva_hvp_lifted", material.material_id, prm, n, u_full.data_ptr(), v_red.data_ptr(), dof_map.data_ptr(), out.numel(), out.data_ptr())
        return out

    def hessian_diagonal(self, material, u, out=None) -> torch.Tensor:
        """diag(d2E/du2) at u, same shape as u: what `ColoredMatrix.diagonal()` would return after `sparse.jacfwd`
        (tatva/sparse/base.py:37-105), computed element-wise without the matrix (Jacobi preconditioner)."""
        prm, n = _lib.params_array(material.params())
        uc = self._as_dev(u).contiguous()
        if out is None:
            out = torch.empty_like(uc)
        self._call("tatva_hessian_diag", material.material_id, prm, n, uc.data_ptr(), out.data_ptr())
        return out

    def _raw_energy(self, material, u):
        prm, n = _lib.params_array(material.params())
        out = torch.empty((), dtype=torch.float64, device=self.device)
        uc = u.contiguous()
        self._call("tatva_energy", material.material_id, prm, n, uc.data_ptr(), out.data_ptr())
        return out

    def _raw_residual(self, material, u):
        prm, n = _lib.params_array(material.params())
        uc = u.contiguous()
        out = torch.empty_like(uc)
        self._call("tatva_residual", material.material_id, prm, n, uc.data_ptr(), out.data_ptr())
        return out

    def _raw_hvp(self, material, u, v, out=None):
        prm, n = _lib.params_array(material.params())
        uc, vc = u.contiguous(), v.contiguous()
        if out is None:
            out = torch.empty_like(uc)
        self._call("tatva_hvp", material.material_id, prm, n, uc.data_ptr(), vc.data_ptr(), out.data_ptr())
        return out


class FusedResidual:
    """Callable u -> r(u) backed by the fused residual kernel (differentiable: its JVP / VJP is the HVP kernel)."""

    def __init__(self, op: Operator, material):
        self.op, self.material = op, material

    def __call__(self, u):
        return _Residual.apply(self.op._fused_shape(self.material, u)[0], self.op, self.material)


def _to_np_or_tensor(a):
    return a if isinstance(a, torch.Tensor) else np.asarray(a)


# forward kernel / adjoint kernel of each linear building block
_LINEAR = {
    "grad": ("_k_grad", "_k_grad_adj"),
    "eval": ("_k_eval", "_k_eval_adj"),
    "gather": ("_k_gather", "_k_gather_adj"),
}


class _LinearOp(torch.autograd.Function):
    """y = A x for A in {grad, eval, gather, integrate, sum}; backward is the adjoint kernel, which is
    itself differentiable (its backward is A again), so reverse-over-reverse gives H v."""

    @staticmethod
    def forward(x, op, which):
        x = x.contiguous()
        if which in _LINEAR:
            return getattr(op, _LINEAR[which][0])(x)
        if which.endswith("_adj") and which[:-4] in _LINEAR:
            return getattr(op, _LINEAR[which[:-4]][1])(x)
        if which == "integrate":
            return op._k_integrate_quad(x)
        if which == "integrate_adj":  # (E, nv) -> (E, Q, nv): g[e,c] * W[e,q]
            return x[:, None, :] * op.get_integration_weights()[:, :, None]
        if which == "sum":
            return op._k_sum_rows(x)
        if which == "sum_adj":  # (nv,) -> (E, nv)
            return x[None, :].expand(op.n_elements, -1).contiguous()
        raise ValueError(which)

    @staticmethod
    def setup_context(ctx, inputs, output):
        _, ctx.op, ctx.which = inputs

    @staticmethod
    def backward(ctx, g):
        which = ctx.which
        adj = which[:-4] if which.endswith("_adj") else which + "_adj"
        return _LinearOp.apply(g.contiguous(), ctx.op, adj), None, None

    @staticmethod
    def jvp(ctx, t, *_):
        return _LinearOp.apply(t.contiguous(), ctx.op, ctx.which)


class _Energy(torch.autograd.Function):
    @staticmethod
    def forward(u, op, material):
        return op._raw_energy(material, u)

    @staticmethod
    def setup_context(ctx, inputs, output):
        u, ctx.op, ctx.material = inputs
        ctx.save_for_backward(u)
        ctx.save_for_forward(u)

    @staticmethod
    def backward(ctx, g):
        (u,) = ctx.saved_tensors
        return g * _Residual.apply(u, ctx.op, ctx.material), None, None

    @staticmethod
    def jvp(ctx, t, *_):
        (u,) = ctx.saved_tensors
        return (_Residual.apply(u, ctx.op, ctx.material) * t).sum()


class _Residual(torch.autograd.Function):
    @staticmethod
    def forward(u, op, material):
        return op._raw_residual(material, u)

    @staticmethod
    def setup_context(ctx, inputs, output):
        u, ctx.op, ctx.material = inputs
        ctx.save_for_backward(u)
        ctx.save_for_forward(u)

    @staticmethod
    def backward(ctx, g):
        (u,) = ctx.saved_tensors  # the Hessian is symmetric: J^T g = H g
        return ctx.op._raw_hvp(ctx.material, u, g.reshape(u.shape)).reshape(u.shape), None, None

    @staticmethod
    def jvp(ctx, t, *_):
        (u,) = ctx.saved_tensors
        return ctx.op._raw_hvp(ctx.material, u, t.reshape(u.shape)).reshape(u.shape)
