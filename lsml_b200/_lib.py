"""ctypes binding of ``liblsml_b200.so`` (the C ABI declared in ``include/lsml_b200.h``).

There is no CPU fallback: importing this module without the built library raises
``ImportError`` (build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C lsml_b200/csrc``), and every call on a box without a CUDA device raises
``RuntimeError`` with the library's own error string.
"""
import ctypes
import pathlib

_PKG = pathlib.Path(__file__).resolve().parent
LIB_PATH = _PKG / "liblsml_b200.so"

c_i64 = ctypes.c_longlong
c_p = ctypes.c_void_p

if not LIB_PATH.exists():
    raise ImportError(
        "lsml_b200: %s is missing -- the CUDA library must be built (make -C lsml_b200/csrc); "
        "there is no CPU fallback" % LIB_PATH)

lib = ctypes.CDLL(str(LIB_PATH))

lib.lsml_last_error.restype = ctypes.c_char_p
lib.lsml_launch_count.restype = c_i64
lib.lsml_compact_workspace_bytes.restype = c_i64
lib.lsml_compact_workspace_bytes.argtypes = [c_i64]


class LsmlError(RuntimeError):
    pass


def check(status, what=""):
    if status != 0:
        msg = lib.lsml_last_error().decode("utf-8", "replace")
        raise LsmlError("liblsml_b200 %s failed (status %d): %s" % (what, status, msg))


def launch_count():
    return int(lib.lsml_launch_count())


def int_array(values):
    return (ctypes.c_int * len(values))(*[int(v) for v in values])


def double_array(values):
    return (ctypes.c_double * len(values))(*[float(v) for v in values])


def ptr(t):
    """Device (or host) address of a torch tensor / numpy array, or NULL for None."""
    if t is None:
        return c_p(0)
    if hasattr(t, "data_ptr"):
        return c_p(t.data_ptr())
    return c_p(t.ctypes.data)


class BandSource(ctypes.Structure):
    """``lsml_band_source`` of include/lsml_b200.h: the feature program plus the device buffers
    it reads, handed to the ``lsml_*_predict_band`` entry points."""
    _fields_ = [("nfeat", ctypes.c_int), ("kind", c_p), ("a", c_p), ("b", c_p),
                ("ndim", ctypes.c_int), ("shape", c_p), ("batch", c_i64), ("dx", c_p),
                ("band_idx", c_p), ("planes", c_p), ("plane_stride", c_i64),
                ("item_feat", c_p), ("com", c_p), ("origin0", ctypes.c_int)]


def current_stream():
    import torch
    return c_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    n = lib.lsml_device_count()
    if n <= 0:
        raise LsmlError("lsml_b200 needs a CUDA device (found %d): %s" % (
            n, lib.lsml_last_error().decode("utf-8", "replace")))
    return n
