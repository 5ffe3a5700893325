"""Test helper: drive rpkg/src/cv_rglue.c the way R does, through the stand-in runtime in
rpkg/rstub/ (R is not installed in this image).  `RSession.call(name, *args)` is `.Call(name, ...)`:
numpy arrays / scalars / lists / None go in as REALSXP / INTSXP / LGLSXP / VECSXP / NULL with R's
attributes (dim for matrices), the result comes back as numpy (matrices column-major) or a list."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NILSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP = 0, 10, 13, 14, 16, 19


class RError(RuntimeError):
    """An R-level error: what stop() / Rf_error() raise."""


class RSession:
    def __init__(self):
        sys.path.insert(0, os.path.join(ROOT, "rpkg"))
        import build_stub
        path = build_stub.build()
        self.dll = C.CDLL(path)
        d = self.dll
        vp = C.c_void_p
        d.rstub_load.restype = C.c_int
        d.rstub_load.argtypes = [vp]
        d.rstub_routine_name.restype = C.c_char_p
        d.rstub_routine_name.argtypes = [C.c_int]
        d.rstub_routine_nargs.restype = C.c_int
        d.rstub_routine_nargs.argtypes = [C.c_int]
        d.rstub_dynamic_symbols.restype = C.c_int
        d.rstub_nil.restype = vp
        d.rstub_new.restype = vp
        d.rstub_new.argtypes = [C.c_int, C.c_ssize_t]
        d.rstub_new_string.restype = vp
        d.rstub_new_string.argtypes = [C.c_char_p]
        d.rstub_set_dim.argtypes = [vp, C.c_int, C.c_int]
        d.rstub_type.restype = C.c_int
        d.rstub_type.argtypes = [vp]
        d.rstub_len.restype = C.c_ssize_t
        d.rstub_len.argtypes = [vp]
        d.rstub_data.restype = vp
        d.rstub_data.argtypes = [vp]
        d.rstub_get_dim.restype = C.c_int
        d.rstub_get_dim.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        d.rstub_elt.restype = vp
        d.rstub_elt.argtypes = [vp, C.c_ssize_t]
        d.rstub_set_elt.argtypes = [vp, C.c_ssize_t, vp]
        d.rstub_live_objects.restype = C.c_long
        d.rstub_dot_call.restype = vp
        d.rstub_dot_call.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp), C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        init = C.cast(d.R_init_chromVARb200, vp)
        self.n_routines = d.rstub_load(init)            # library(chromVARb200): R_init_<pkg>(dll)
        self.routines = {d.rstub_routine_name(i).decode(): d.rstub_routine_nargs(i) for i in range(self.n_routines)}

    # ---- numpy -> SEXP ----
    def to_sexp(self, v):
        d = self.dll
        if v is None:
            return d.rstub_nil()
        if isinstance(v, str):
            return d.rstub_new_string(v.encode())
        if isinstance(v, (list, tuple)):
            s = d.rstub_new(VECSXP, len(v))
            for i, e in enumerate(v):
                d.rstub_set_elt(s, i, self.to_sexp(e))
            return s
        if isinstance(v, (bool, np.bool_)):
            v = np.array([v], dtype=bool)
        elif isinstance(v, (int, np.integer)):
            v = np.array([v], dtype=np.int32)
        elif isinstance(v, (float, np.floating)):
            v = np.array([v], dtype=np.float64)
        a = np.asarray(v)
        if a.dtype == np.bool_:
            t, a = LGLSXP, a.astype(np.int32)
        elif np.issubdtype(a.dtype, np.integer):
            t, a = INTSXP, a.astype(np.int32)
        else:
            t, a = REALSXP, a.astype(np.float64)
        s = d.rstub_new(t, a.size)
        if s is None:
            raise MemoryError("rstub: allocation failed")
        flat = np.asfortranarray(a).ravel(order="F")     # R matrices are column-major
        if a.size:
            C.memmove(d.rstub_data(s), flat.ctypes.data, flat.nbytes)
        if a.ndim == 2:
            d.rstub_set_dim(s, a.shape[0], a.shape[1])
        return s

    # ---- SEXP -> numpy ----
    def from_sexp(self, s):
        d = self.dll
        if s is None or s == d.rstub_nil():
            return None
        t, n = d.rstub_type(s), d.rstub_len(s)
        if t == VECSXP:
            return [self.from_sexp(d.rstub_elt(s, i)) for i in range(n)]
        dt = {REALSXP: np.float64, INTSXP: np.int32, LGLSXP: np.int32}[t]
        out = np.empty(n, dt)
        if n:
            C.memmove(out.ctypes.data, d.rstub_data(s), out.nbytes)
        nr, nc = C.c_int(), C.c_int()
        if d.rstub_get_dim(s, C.byref(nr), C.byref(nc)):
            out = out.reshape((nr.value, nc.value), order="F")
        return out

    def call(self, name, *args):
        """.Call(name, ...) through the registered table."""
        sx = [self.to_sexp(a) for a in args]
        arr = (C.c_void_p * max(len(sx), 1))(*sx)
        err = C.create_string_buffer(1024)
        status = C.c_int()
        res = self.dll.rstub_dot_call(name.encode(), len(sx), arr, err, 1024, C.byref(status))
        try:
            if status.value == 1:
                raise RError(err.value.decode())
            if status.value == 2:
                raise AssertionError("PROTECT stack imbalance in %s: %s" % (name, err.value.decode()))
            return self.from_sexp(res)
        finally:
            self.dll.rstub_gc()


def dgc_slots(mats):
    """The (p, i, x) slot lists of one or several scipy CSC matrices, as the R shim passes them."""
    if not isinstance(mats, (list, tuple)):
        mats = [mats]
    return ([np.asarray(m.indptr, np.int32) for m in mats], [np.asarray(m.indices, np.int32) for m in mats],
            [np.asarray(m.data, np.float64) for m in mats])
