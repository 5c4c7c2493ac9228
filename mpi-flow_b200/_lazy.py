"""Deferred plane-xyz tensors for the reference's call sequence (utils/utils.py:304-331).

The reference materialises xyz_src and xyz_tgt ([B,S,3,H,W] each, 377 MB at 1280x384x64) only to hand them to
render_tgt_rgb_depth, which could derive them from the plane disparities.  The drop-in
get_src_xyz_from_plane_disparity / get_tgt_xyz_from_plane_disparity therefore return a `LazyTensor`: a torch.Tensor
subclass that knows its shape, dtype and device, carries the provenance render_tgt_rgb_depth needs to fuse, and runs
the kernel that produces its values only when something actually reads them.

  * metadata (shape, dtype, device, dim(), size(), is_contiguous() ...) never materialises;
  * dtype-only `.to(...)`, `.float()`, `.half()`, `.double()`, `.contiguous()`, `.detach()`, `.clone()` stay deferred and
    keep the provenance (the generator runs in fp16 and converts its xyz: utils/utils.py:304-309);
  * every other torch function or method materialises the operands first (once; the value is cached) and then runs on
    ordinary tensors -- the result is whatever the reference would have computed;
  * an in-place operation materialises and applies itself to the cached value; the value's version counter (shared
    with every view of it) tells `edited`, so a later render_tgt_rgb_depth reads the edited values through the
    un-fused operators instead of re-deriving the planes.
"""
import torch

_NO_SUBCLASS = torch._C.DisableTorchFunctionSubclass

# attribute reads and queries answered from the wrapper's own metadata
_META_NAMES = {"shape", "dtype", "device", "ndim", "is_cuda", "is_cpu", "layout", "requires_grad", "is_leaf", "grad",
               "grad_fn", "names", "is_sparse", "is_quantized", "is_meta", "is_complex", "is_floating_point", "_version",
               "size", "dim", "numel", "nelement", "stride", "is_contiguous", "element_size", "ndimension", "storage_offset",
               "__len__", "is_shared", "is_pinned", "get_device", "is_inference", "is_nested", "is_mkldnn", "is_xpu",
               "is_ipu", "is_maia", "is_mtia", "is_vulkan", "is_mps", "is_xla", "is_sparse_csr", "retains_grad", "itemsize",
               "nbytes", "output_nr", "is_signed", "__format__", "__repr__", "__hash__", "type"}


def _name(func):
    return getattr(func, "__name__", None) or getattr(getattr(func, "__self__", None), "__name__", "")


class LazyTensor(torch.Tensor):
    """A tensor whose values are produced by `thunk()` on first use.  `plane` is the provenance tag of a plane-derived
    xyz tensor (utils/mpi/mpi_rendering.py), None for anything else."""

    @staticmethod
    def __new__(cls, thunk, shape, dtype, device, plane=None):
        t = torch.Tensor._make_wrapper_subclass(cls, tuple(shape), dtype=dtype, device=device, requires_grad=False)
        t._thunk, t._value, t._mpiflow_plane = thunk, None, plane
        return t

    # ---- evaluation ---------------------------------------------------------------------------------------------
    @property
    def is_materialized(self):
        return self._value is not None

    def materialize(self):
        if self._value is None:
            with _NO_SUBCLASS():
                v = self._thunk()
            if isinstance(v, LazyTensor):
                v = v.materialize()
            assert tuple(v.shape) == tuple(self.shape) and v.dtype == self.dtype, "deferred tensor changed its metadata"
            self._value, self._thunk, self._version0 = v, None, v._version
        return self._value

    @property
    def edited(self):
        """the materialised values were modified in place (directly or through a view of them) since they were computed"""
        return self._value is not None and self._value._version != self._version0

    def _derived(self, fn, dtype=None):
        """still deferred: fn applied to the value when (if ever) it is needed; same provenance"""
        parent = self
        return LazyTensor(lambda: fn(parent.materialize()), self.shape, dtype or self.dtype, self.device, self._mpiflow_plane)

    # ---- interception -------------------------------------------------------------------------------------------
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = _name(func)
        self = args[0] if args and isinstance(args[0], LazyTensor) else None
        if self is not None:
            if name == "__get__" or name in _META_NAMES:
                if name == "type" and (len(args) > 1 or kwargs):          # t.type(dtype) converts
                    pass
                elif name in ("__repr__", "__format__"):
                    return f"LazyTensor(shape={tuple(self.shape)}, dtype={self.dtype}, device={self.device}, " \
                           f"materialized={self.is_materialized})"
                else:
                    with _NO_SUBCLASS():
                        return func(*args, **kwargs)
            if not self.is_materialized:
                stay = cls._stays_deferred(self, name, args[1:], kwargs)
                if stay is not None:
                    return stay
        # anything else reads values: run on ordinary tensors
        inplace = self is not None and name.endswith("_") and not name.endswith("__")
        real_args = _materialize_tree(args)
        real_kwargs = _materialize_tree(kwargs)
        with _NO_SUBCLASS():
            out = func(*real_args, **real_kwargs)
        if inplace or ("out" in kwargs and isinstance(kwargs["out"], LazyTensor)):
            target = self if inplace else kwargs["out"]
            target._mpiflow_plane = None          # edited: no longer what the plane geometry says
            return target
        return out

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        """safety net (C++ callers that skip __torch_function__): same rule, values first"""
        return func(*_materialize_tree(args), **_materialize_tree(kwargs or {}))

    @staticmethod
    def _stays_deferred(self, name, rest, kwargs):
        if name in ("contiguous", "detach", "clone", "requires_grad_") and not kwargs.get("memory_format"):
            return self if name != "clone" else self._derived(lambda v: v.clone())
        simple = {"float": torch.float32, "half": torch.float16, "double": torch.float64, "bfloat16": torch.bfloat16}
        if name in simple and not rest and not kwargs:
            return self if self.dtype == simple[name] else self._derived(lambda v, d=simple[name]: v.to(d), simple[name])
        if name in ("to", "type"):
            dtype, device = kwargs.get("dtype"), kwargs.get("device")
            for a in rest:
                if isinstance(a, torch.dtype):
                    dtype = a
                elif isinstance(a, (torch.device, str)) and not (name == "type"):
                    device = a
                elif isinstance(a, torch.Tensor):
                    dtype, device = a.dtype, a.device
                elif isinstance(a, bool):
                    continue                                        # non_blocking / copy flags
                else:
                    return None
            if device is not None and torch.device(device) != self.device:
                if not (torch.device(device).type == self.device.type and torch.device(device).index is None):
                    return None                                     # a real transfer: materialise
            if dtype is None or dtype == self.dtype:
                return self
            return self._derived(lambda v, d=dtype: v.to(d), dtype)
        return None


def _materialize_tree(x):
    if isinstance(x, LazyTensor):
        return x.materialize()
    if isinstance(x, (list, tuple)):
        return type(x)(_materialize_tree(v) for v in x)
    if isinstance(x, dict):
        return {k: _materialize_tree(v) for k, v in x.items()}
    return x


def materialize(t):
    """the ordinary tensor behind t (t itself when it is not deferred)"""
    return t.materialize() if isinstance(t, LazyTensor) else t
