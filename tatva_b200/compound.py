"""Compound state layout — mirror of tatva.compound (tatva/compound/__init__.py, field.py, field_types.py).

A `Compound` subclass declares named fields over one flat array.  The layout rules are the
reference's: AUTO-sized full-nodal fields (two or more) are stacked on axis 1, i.e. node-interleaved
`[ux, uy, uz, phi]` per node (compound/__init__.py:166-199, :334-389; pinned by the reference's
tests/test_compound.py:134-147); a single such field is left contiguous (:184-188); all other fields
follow in declaration order.  Every field is an affine view (offset + strides) of the flat array, so
on torch tensors the views are zero-copy `as_strided` windows that the CUDA kernels read directly
(the fused phase-field kernels take the interleaved (N, 4) block as is).

The flat array may be a NumPy array or a torch tensor (CPU or CUDA).
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from enum import IntEnum
from math import prod
from typing import Any, Callable

import numpy as np

try:  # torch is the device-array library of this package, but layouts work on NumPy alone
    import torch
except ImportError:  # pragma: no cover
    torch = None


class CompoundError(ValueError):
    """Base error class for Compound-related errors."""


class CompoundStackError(CompoundError):
    pass


class FieldSize(IntEnum):
    AUTO = -1


# -- field types (compound/field_types.py) ---------------------------------------------------------


class _FieldType:
    def get(self):
        return self


@dataclass
class Nodal(_FieldType):
    node_ids: Any = None
    stack: bool = True


@dataclass
class Local(_FieldType):
    pass


@dataclass
class Shared(_FieldType):
    pass


class FieldType(IntEnum):
    LOCAL = 0
    NODAL = 1
    SHARED = 2

    def get(self) -> _FieldType:
        return {FieldType.LOCAL: Local, FieldType.NODAL: Nodal, FieldType.SHARED: Shared}[self]()


@dataclass(frozen=True)
class _FieldSpec:
    shape: tuple
    default_factory: Callable | None = None
    field_type: Any = FieldType.LOCAL


field = _FieldSpec


def _row_major(shape):
    out, s = [], 1
    for extent in reversed(shape):
        out.append(s)
        s *= extent
    return tuple(reversed(out))


def _is_torch(a):
    return torch is not None and isinstance(a, torch.Tensor)


class Field:
    """Descriptor: an affine window (offset, strides) of the flat array.

    `root_slice` / `root_shape` describe the contiguous parent block (the stacked block for stacked
    fields, the field itself otherwise) — used by pattern_from_compound and the MPI layout."""

    def __init__(self, shape, offset, strides, default_factory=None, field_type=FieldType.LOCAL, root_slice=None, root_shape=None, view_slice=None):
        self.shape = tuple(shape)
        self.size = int(prod(self.shape))
        self._base_offset = int(offset)
        self._strides = tuple(int(s) for s in strides)
        self.default_factory = default_factory
        self.field_type = field_type
        contiguous = self._strides == _row_major(self.shape)
        self._slice = slice(self._base_offset, self._base_offset + self.size) if contiguous else None
        self._root_slice = root_slice if root_slice is not None else self._slice
        self._root_shape = tuple(root_shape) if root_shape is not None else self.shape
        if view_slice is not None:
            self._view_slice = view_slice

    # -- flat DOF indices (compound/field.py:157-200) -------------------------------------------
    def indices(self, arg) -> np.ndarray:
        if not isinstance(arg, tuple):
            arg = (arg,)
        arg = arg + (slice(None),) * (len(self.shape) - len(arg))
        axes = []
        for sub, extent in zip(arg, self.shape, strict=True):
            if isinstance(sub, (int, np.integer)):
                axes.append(np.asarray([sub if sub >= 0 else sub + extent], dtype=np.int64))
            elif isinstance(sub, slice):
                axes.append(np.arange(*sub.indices(extent), dtype=np.int64))
            else:
                v = np.asarray(sub.cpu() if _is_torch(sub) else sub, dtype=np.int64).reshape(-1)
                axes.append(np.where(v < 0, v + extent, v))
        if not axes:
            return np.asarray([self._base_offset], dtype=np.int64)
        idx = np.full(tuple(len(a) for a in axes), self._base_offset, dtype=np.int64)
        for ax, (stride, vals) in enumerate(zip(self._strides, axes, strict=True)):
            shape = [1] * len(axes)
            shape[ax] = len(vals)
            idx = idx + (vals * stride).reshape(shape)
        return idx.reshape(-1)

    def __getitem__(self, arg):
        return self.indices(arg)

    # -- views -------------------------------------------------------------------------------------
    def _view(self, arr):
        if _is_torch(arr):
            return arr.as_strided(self.shape, self._strides, arr.storage_offset() + self._base_offset)
        a = np.asarray(arr)
        return np.lib.stride_tricks.as_strided(a[self._base_offset :], shape=self.shape, strides=tuple(s * a.itemsize for s in self._strides), writeable=False)

    def _set_in_array(self, arr, value):
        """Functional update: a new flat array with this field replaced (reference: arr.at[...].set)."""
        if _is_torch(arr):
            out = arr.clone()
            out.as_strided(self.shape, self._strides, out.storage_offset() + self._base_offset).copy_(
                torch.broadcast_to(torch.as_tensor(value, dtype=arr.dtype, device=arr.device), self.shape)  # Array | float, as arr.at[...].set
            )
            return out
        out = np.array(arr, copy=True)
        out[self.indices(slice(None))] = np.broadcast_to(np.asarray(value, dtype=out.dtype), self.shape).reshape(-1)
        return out

    def __get__(self, instance, owner=None):
        if instance is None:
            return self
        return self._view(instance.arr)


def _auto_nodal(spec):
    return len(spec.shape) > 0 and spec.shape[0] == FieldSize.AUTO


def _stack(items, offset, axis):
    """Descriptors for fields sharing one contiguous block, stacked along `axis`
    (compound/__init__.py:334-389)."""
    first = items[0][1].shape
    ax = axis % len(first) if len(first) > 0 else 0
    prefix = first[:ax]
    spans, width = [], 0
    for name, it in items:
        shp = it.shape
        if len(shp) < ax:
            raise CompoundStackError(f"Field {name} rank {len(shp)} is not compatible with stacking axis {ax}.")
        if shp[:ax] != prefix:
            raise CompoundStackError(f"Field {name} shape {shp} prefix does not match base shape along axis {ax}.")
        ext = int(prod(shp[ax:])) if shp[ax:] else 1
        spans.append((name, it, width, width + ext))
        width += ext
    root_shape = tuple(prefix) + (width,)
    total = int(prod(root_shape))
    root_slice = slice(offset, offset + total)
    root_strides = _row_major(root_shape)
    out = {}
    for name, it, a, b in spans:
        # window [..., a:b] of the root block, reshaped to the field's own trailing shape
        strides = root_strides[:ax] + tuple(s * root_strides[ax] for s in _row_major(it.shape[ax:]))
        out[name] = Field(
            it.shape, offset + a * root_strides[ax], strides, it.default_factory, it.field_type,
            root_slice=root_slice, root_shape=root_shape, view_slice=(slice(None),) * len(prefix) + (slice(a, b),),
        )
    return out, total


class _GlobalFieldIndices:
    """Natural (mesh-global) DOF ids of a field, indexed like the field itself (compound/mpi.py:113-202).  For a
    field on a node subset the first index is a GLOBAL node id, translated to its position in the subset."""

    def __init__(self, info):
        self.shape, self._base_offset, self._strides, self.global_subset = tuple(info.global_shape), int(info.global_base_offset), tuple(info.global_strides), info.global_subset

    def _subset_position(self, first):
        sub = self.global_subset
        if isinstance(first, (int, np.integer)):
            pos = int(np.searchsorted(sub, first))
            if pos >= len(sub) or sub[pos] != first:
                raise IndexError(f"Global node ID {first} not in field subset.")
            return pos
        if isinstance(first, slice):
            if first.step is not None and first.step != 1:
                raise NotImplementedError("Slicing with step on global subset not supported yet.")
            lo = int(np.searchsorted(sub, first.start)) if first.start is not None else 0
            hi = int(np.searchsorted(sub, first.stop)) if first.stop is not None else len(sub)
            return slice(lo, hi)
        ids = np.asarray(first.cpu() if _is_torch(first) else first)
        pos = np.searchsorted(sub, ids)
        ok = (pos < len(sub)) & (sub[np.minimum(pos, len(sub) - 1)] == ids)
        if not np.all(ok):
            raise IndexError(f"Global node IDs {ids[~ok]} not found in field subset.")
        return pos.astype(np.int64)

    def __getitem__(self, arg) -> np.ndarray:
        if not isinstance(arg, tuple):
            arg = (arg,)
        arg = arg + (slice(None),) * (len(self.shape) - len(arg))
        if self.global_subset is not None and arg:
            arg = (self._subset_position(arg[0]),) + arg[1:]
        # same affine index arithmetic as the local Field
        return Field(self.shape, self._base_offset, self._strides).indices(arg)


class _GlobalIndicesView:
    """`MyState._g.<field>[...]`: natural global DOF ids (compound/mpi.py:86-110)."""

    def __init__(self, compound_cls):
        self._compound = compound_cls

    def __getattr__(self, name):
        info = self._compound._global_field_info
        if info is None or name not in info:
            raise AttributeError(f"'{self._compound.__name__}' has no field '{name}' or layout not initialized.")
        return _GlobalFieldIndices(info[name])


class _GlobalDataView:
    """`state._g.<field>`: the field assembled over all ranks in natural global order (compound/mpi.py:205-276).
    Owned entries are placed at their natural ids in a zero vector of the global size and summed over the ranks
    (one all-reduce, cached per view)."""

    def __init__(self, instance):
        self._state = instance
        self._full = None

    def _gather(self):
        if self._full is None:
            from .mpi import _as_comm

            st = self._state
            layout, comm = st._layout, _as_comm(st._comm)
            arr = st.arr if _is_torch(st.arr) else torch.as_tensor(np.asarray(st.arr))
            owned = torch.as_tensor(layout.owned_mask, device=arr.device)
            l2g = torch.as_tensor(np.asarray(layout.natural_l2g, dtype=np.int64), device=arr.device)
            full = torch.zeros(int(layout.n_global), dtype=arr.dtype, device=arr.device)
            full[l2g[owned]] = arr.reshape(-1)[owned]
            if comm.size > 1:
                import torch.distributed as dist

                dist.all_reduce(full, group=comm.group)
            self._full = full if _is_torch(st.arr) else full.numpy()
        return self._full

    def __getattr__(self, name):
        st = self._state
        info = st._global_field_info
        if name.startswith("_") or info is None or name not in dict(st.fields):
            raise AttributeError(f"'{type(st).__name__}' has no field '{name}'")
        idx = _GlobalFieldIndices(info[name])[(slice(None),) * len(info[name].global_shape)]
        full = self._gather()
        picked = full[torch.as_tensor(idx, device=full.device)] if _is_torch(full) else full[idx]
        return picked.reshape(tuple(info[name].global_shape))


class _GlobalView:
    """Descriptor behind `Compound._g` (compound/mpi.py:58-83): indices on the class, gathered data on an instance."""

    def __get__(self, instance, owner):
        if owner._layout is None:
            raise ValueError(f"Compound class '{owner.__name__}' is missing a DOF layout. Global view requires a complete MPI layout.")
        return _GlobalIndicesView(owner) if instance is None else _GlobalDataView(instance)


class Compound:
    """Flat state with named fields; see the module docstring."""

    _g = _GlobalView()

    fields: tuple = ()
    size: int = 0
    _mesh = None
    _layout = None
    _global_field_info = None
    _comm = None

    def __init_subclass__(cls, *, mesh=None, partition_info=None, comm=None, **kwargs):
        super().__init_subclass__(**kwargs)
        inherited = []
        for base in cls.__mro__[1:]:
            if isinstance(base, type) and issubclass(base, Compound) and base is not Compound:
                inherited = list(base.fields)
                break
        offset = sum(int(prod(f.shape)) for _, f in inherited)
        specs = [(n, v) for n, v in cls.__dict__.items() if isinstance(v, _FieldSpec)]
        reserved = set(dir(Compound)) | {"arr"}
        for n, _ in specs:
            if n in reserved:
                raise CompoundError(f"Field name '{n}' is reserved and cannot be used in Compound class.")
        stacked, plain, resolved = [], [], {}
        for name, spec in specs:
            ft = spec.field_type.get()
            if _auto_nodal(spec):
                if isinstance(ft, Nodal) and ft.node_ids is not None:
                    n_items = len(ft.node_ids)
                else:
                    if mesh is None:
                        raise CompoundError(f"Mesh must be provided to resolve AUTO size for field '{name}'.")
                    n_items = mesh.coords.shape[0]
                spec = replace(spec, shape=(n_items, *spec.shape[1:]), field_type=ft if isinstance(ft, Nodal) else FieldType.NODAL)
            ft = spec.field_type.get()
            (stacked if isinstance(ft, Nodal) and ft.node_ids is None and ft.stack else plain).append((name, spec))
        if len(stacked) == 1:  # compound/__init__.py:184-188
            plain, stacked = stacked + plain, []
        if stacked:
            desc, total = _stack(stacked, offset, axis=1)
            resolved.update(desc)
            offset += total
        for name, spec in plain:
            n = int(prod(spec.shape))
            resolved[name] = Field(spec.shape, offset, _row_major(spec.shape), spec.default_factory, spec.field_type)
            offset += n
        all_fields = list(inherited)
        for name, _ in specs:  # declaration order
            setattr(cls, name, resolved[name])
            all_fields.append((name, resolved[name]))
        cls.fields = tuple(all_fields)
        cls.size = offset
        cls._mesh = mesh
        if partition_info is not None and comm is not None:
            from .mpi import layout_from_compound

            cls._comm = comm
            cls._layout, cls._global_field_info = layout_from_compound(cls, partition_info, comm)

    @classmethod
    def get_layout(cls):
        if cls._layout is None:
            raise CompoundError("Layout not set on Compound class.")
        return cls._layout

    def __init__(self, arr=None, **kwargs):
        if arr is not None:
            n = arr.numel() if _is_torch(arr) else np.size(arr)
            assert n == self.size, f"Array size {n} does not match expected size {self.size}."
            self.arr = arr
        else:
            self.arr = np.zeros(self.size, dtype=np.float64)
            for name, f in self.fields:
                if name in kwargs:
                    self.arr = f._set_in_array(self.arr, kwargs[name])
                elif f.default_factory is not None:
                    self.arr = f._set_in_array(self.arr, f.default_factory())

    def __len__(self):
        return len(self.fields)

    def __iter__(self):
        for name, _ in self.fields:
            yield getattr(self, name)

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(f'{n}={f.shape}' for n, f in self.fields)})"

    def __add__(self, other):
        return type(self)(self.arr + other.arr)

    def at(self, name):
        f = dict(self.fields).get(name)
        if f is None:
            raise AttributeError(f"Unknown field name: {name}")
        return _At(self, f)

    def flatten(self):
        return self.arr


class _At:
    def __init__(self, state, f):
        self.state, self.f = state, f

    def set(self, value):
        return type(self.state)(self.f._set_in_array(self.state.arr, value))


def stack_fields(*names, axis=-1):
    """Class decorator: lay the named fields out as one stacked block (compound/__init__.py:392-470)."""
    if not names:
        raise CompoundError("At least one field name is required.")

    def deco(cls):
        fmap = dict(cls.fields)
        for n in names:
            if n not in fmap:
                raise CompoundStackError(f"Unknown field name in stack_fields: {n}")
        involved = set(names)
        first = min(i for i, (n, _) in enumerate(cls.fields) if n in involved)
        base = cls.fields[first][1]._base_offset
        desc, total = _stack([(n, fmap[n]) for n in names], base, axis)
        out, offset = [], base + total
        for i, (n, f) in enumerate(cls.fields):
            if i < first:
                out.append((n, f))
            elif n in involved:
                setattr(cls, n, desc[n])
                out.append((n, desc[n]))
            else:
                nf = Field(f.shape, offset, _row_major(f.shape), f.default_factory, f.field_type)
                setattr(cls, n, nf)
                out.append((n, nf))
                offset += nf.size
        cls.fields = tuple(out)
        cls.size = offset
        return cls

    return deco


__all__ = ["Compound", "field", "stack_fields", "FieldSize", "FieldType", "Nodal", "Local", "Shared", "CompoundError", "CompoundStackError"]
