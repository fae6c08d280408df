"""ctypes binding of the stand-in problem builder (libdcp_harness.so, include/dcp_harness.h).

Test/bench infrastructure: it plays the role deal.II plays for the reference
(/root/reference/include/core/boussinesq_model.tpp:184-412 `setup_dofs`), producing the arrays the
device library takes.  Not part of the product path.
"""
import ctypes
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB = None

_DTYPES = {0: np.float64, 1: np.int32, 2: np.int64, 3: np.int8, 4: np.int16}


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_ROOT, "lib", "libdcp_harness.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'` or `make`")
        L = ctypes.CDLL(path)
        L.dcph_create.restype = ctypes.c_void_p
        L.dcph_create.argtypes = [ctypes.c_char_p]
        L.dcph_destroy.argtypes = [ctypes.c_void_p]
        L.dcph_array.restype = ctypes.c_int
        L.dcph_array.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p),
                                 ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int)]
        L.dcph_scalar.restype = ctypes.c_int64
        L.dcph_scalar.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.dcph_n_arrays.restype = ctypes.c_int
        L.dcph_n_arrays.argtypes = [ctypes.c_void_p]
        L.dcph_array_name.restype = ctypes.c_char_p
        L.dcph_array_name.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.dcph_n_scalars.restype = ctypes.c_int
        L.dcph_n_scalars.argtypes = [ctypes.c_void_p]
        L.dcph_scalar_name.restype = ctypes.c_char_p
        L.dcph_scalar_name.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.dcph_last_error.restype = ctypes.c_char_p
        _LIB = L
    return _LIB


class Problem:
    """One mesh + DoF + constraint + pattern set.  Arrays are zero-copy numpy views into the C++ object."""

    def __init__(self, **spec):
        self.spec = dict(spec)
        s = ",".join(f"{k}={v}" for k, v in spec.items())
        self._h = lib().dcph_create(s.encode())
        if not self._h:
            raise RuntimeError("dcph_create: " + lib().dcph_last_error().decode())
        self._cache = {}

    def close(self):
        if getattr(self, "_h", None):
            self._cache.clear()
            lib().dcph_destroy(self._h)
            self._h = None

    def __del__(self):
        if sys.is_finalizing():
            return
        try:
            self.close()
        except Exception:
            pass

    def names(self):
        L = lib()
        return [L.dcph_array_name(self._h, i).decode() for i in range(L.dcph_n_arrays(self._h))]

    def scalar_names(self):
        L = lib()
        return [L.dcph_scalar_name(self._h, i).decode() for i in range(L.dcph_n_scalars(self._h))]

    def has(self, name):
        return name in self.names()

    def scalar(self, name):
        v = lib().dcph_scalar(self._h, name.encode())
        if v < 0:
            raise KeyError(name)
        return int(v)

    def array(self, name):
        if name in self._cache:
            return self._cache[name]
        p = ctypes.c_void_p()
        n = ctypes.c_int64()
        dt = ctypes.c_int()
        if lib().dcph_array(self._h, name.encode(), ctypes.byref(p), ctypes.byref(n), ctypes.byref(dt)) != 0:
            raise KeyError(name)
        dtype = np.dtype(_DTYPES[dt.value])
        if n.value == 0:
            a = np.zeros(0, dtype=dtype)
        else:
            buf = (ctypes.c_char * (n.value * dtype.itemsize)).from_address(p.value)
            a = np.frombuffer(buf, dtype=dtype)
        self._cache[name] = a
        return a

    __getitem__ = array

    # convenience ------------------------------------------------------------------------------
    @property
    def dim(self):
        return self.scalar("dim")

    @property
    def n_cells(self):
        return self.scalar("n_cells")

    def csr(self, name):
        return self.array(name + ".rowptr"), self.array(name + ".col"), self.scalar(name + ".n_rows"), self.scalar(name + ".n_cols")
