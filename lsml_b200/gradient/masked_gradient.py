"""Drop-in for ``lsml.gradient.masked_gradient`` (reference: lsml/gradient/masked_gradient.py).

Same two public functions, same argument meaning and error behaviour:

* ``gradient_centered(arr, mask=None, dx=None, return_gradient_magnitude=True, normalize=False)``
  (reference :71-152)
* ``gradient_magnitude_osher_sethian(arr, nu, mask=None, dx=None)`` (reference :155-237)

``arr`` may be

* a float64 **numpy** array: the call goes through the six drop-in C symbols of
  ``liblsml_b200.so`` with host pointers, exactly like the reference's ctypes call at
  masked_gradient.py:235 (stage to the GPU, sm_100a kernel, copy back), or
* a CUDA **torch** tensor (float64 or float32): the device entry points run on the current
  stream with no host copy and torch tensors are returned.  A leading batch axis is allowed
  for tensors via ``batched=True``.
"""
import ctypes

import numpy as np
from numpy.ctypeslib import ndpointer

from .. import _lib

_c_double_arr = ndpointer(ctypes.c_double, flags="C_CONTIGUOUS")
_c_bool_arr = ndpointer(ctypes.c_bool, flags="C_CONTIGUOUS")


def _is_tensor(x):
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda")


def _host_gradient_centered_func(ndim):
    func = getattr(_lib.lib, "gradient_centered{:d}d".format(ndim))
    func.restype = None
    func.argtypes = ((ctypes.c_int,) * ndim + (_c_double_arr, _c_bool_arr) +
                     (_c_double_arr,) * ndim + (_c_double_arr,) +
                     (ctypes.c_double,) * ndim + (ctypes.c_int,))
    return func


def _host_gmag_os_func(ndim):
    func = getattr(_lib.lib, "gmag_os{:d}d".format(ndim))
    func.restype = None
    func.argtypes = ((ctypes.c_int,) * ndim + (_c_double_arr, _c_bool_arr, _c_double_arr,
                                               _c_double_arr) + (ctypes.c_double,) * ndim)
    return func


def _validate(arr, mask, dx, batched):
    ndim = arr.ndim - (1 if batched else 0)
    assert 1 <= ndim <= 3, "Only dimensions 1-3 supported."
    if not _is_tensor(arr) and arr.dtype != np.float64:      # checked second, as the reference
        raise ValueError("`arr` must be float type.")
    if mask is not None and mask.ndim != arr.ndim:
        raise ValueError("Shape mismatch between `mask` and `arr`.")
    if dx is not None:
        if len(dx) != ndim:
            raise ValueError("`dx` vector shape mismatch.")
    else:
        dx = np.ones(ndim, dtype=float)
    return ndim, np.asarray(dx, dtype=np.float64)


def _tensor_args(arr, mask, batched):
    import torch
    if not arr.is_cuda:
        raise ValueError("torch tensors must live on a CUDA device (no CPU path)")
    if arr.dtype not in (torch.float64, torch.float32):
        raise ValueError("`arr` must be float type.")
    arr = arr.contiguous()
    if mask is not None:
        if mask.dtype not in (torch.bool, torch.uint8):
            raise TypeError("`mask` must be a bool tensor")
        mask = mask.contiguous()
    shape = tuple(arr.shape[1:] if batched else arr.shape)
    batch = arr.shape[0] if batched else 1
    suffix = "f64" if arr.dtype == torch.float64 else "f32"
    return arr, mask, shape, batch, suffix


def gradient_centered(arr, mask=None, dx=None, return_gradient_magnitude=True,
                      normalize=False, batched=False):
    """Centered-difference gradient of ``arr`` where ``mask`` is true (reference :71-152)."""
    ndim, dx = _validate(arr, mask, dx, batched and _is_tensor(arr))
    if _is_tensor(arr):
        import torch
        arr, mask, shape, batch, suffix = _tensor_args(arr, mask, batched)
        grads = torch.zeros((ndim,) + tuple(arr.shape), dtype=arr.dtype, device=arr.device)
        gmag = torch.zeros_like(arr)
        with torch.cuda.device(arr.device):
            fn = getattr(_lib.lib, "lsml_gradient_centered_" + suffix)
            _lib.check(fn(ctypes.c_int(ndim), _lib.int_array(shape), _lib.c_i64(batch),
                          _lib.ptr(arr), _lib.ptr(mask), _lib.ptr(grads), _lib.ptr(gmag),
                          _lib.double_array(dx), ctypes.c_int(int(normalize)),
                          _lib.current_stream()), "gradient_centered")
        gradients = [grads[a] for a in range(ndim)]
    else:
        if arr.dtype != np.float64:
            raise ValueError("`arr` must be float type.")
        _lib.require_cuda()
        arr = np.ascontiguousarray(arr)
        mask = (np.ones(arr.shape, dtype=bool) if mask is None
                else np.ascontiguousarray(mask, dtype=bool))
        gradients = [np.zeros_like(arr) for _ in range(ndim)]
        gmag = np.zeros_like(arr)
        func = _host_gradient_centered_func(ndim)
        func(*(arr.shape + (arr, mask) + tuple(gradients) + (gmag,) + tuple(dx) +
               (int(normalize),)))
    if return_gradient_magnitude:
        return gradients, gmag
    return gradients


def gradient_magnitude_osher_sethian(arr, nu, mask=None, dx=None, batched=False):
    """Osher-Sethian upwind gradient magnitude for ``u_t = nu |Du|`` (reference :155-237)."""
    ndim, dx = _validate(arr, mask, dx, batched and _is_tensor(arr))
    if _is_tensor(arr):
        import torch
        arr, mask, shape, batch, suffix = _tensor_args(arr, mask, batched)
        nu = nu.to(dtype=arr.dtype).contiguous()
        gmag = torch.zeros_like(arr)
        with torch.cuda.device(arr.device):
            fn = getattr(_lib.lib, "lsml_gmag_os_" + suffix)
            _lib.check(fn(ctypes.c_int(ndim), _lib.int_array(shape), _lib.c_i64(batch),
                          _lib.ptr(arr), _lib.ptr(mask), _lib.ptr(nu), _lib.ptr(gmag),
                          _lib.double_array(dx), _lib.current_stream()), "gmag_os")
        return gmag
    if arr.dtype != np.float64:
        raise ValueError("`arr` must be float type.")
    _lib.require_cuda()
    arr = np.ascontiguousarray(arr)
    nu = np.ascontiguousarray(nu, dtype=np.float64)
    mask = (np.ones(arr.shape, dtype=bool) if mask is None
            else np.ascontiguousarray(mask, dtype=bool))
    gmag = np.zeros_like(arr)
    func = _host_gmag_os_func(ndim)
    func(*(arr.shape + (arr, mask, nu, gmag) + tuple(dx)))
    return gmag
