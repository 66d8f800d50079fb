"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. Nothing under stress-particle-sph_b200/ may import this module.
"""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "stress-particle-sph_b200"))
from spsph import _abi  # noqa: E402

_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(f"{_SO} missing: run `make -C oracle`")
        L = C.CDLL(_SO)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(_abi.Params), C.POINTER(_abi.State)]
        L.oracle_step.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_double]
        L.oracle_run.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_int32, C.POINTER(C.c_double)]
        L.oracle_download.argtypes = [C.c_void_p, C.POINTER(_abi.State)]
        L.oracle_pair_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64)] + [C.POINTER(C.c_int32)] * 3
        L.oracle_pairs.argtypes = [C.c_void_p, C.POINTER(C.c_int64)] + [C.c_void_p] * 6
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_last_error.argtypes = [C.c_void_p]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_get_list_capacity.restype = C.c_int64
        L.oracle_get_list_capacity.argtypes = [C.c_void_p]
        L.oracle_set_list_capacity.restype = None
        L.oracle_set_list_capacity.argtypes = [C.c_void_p, C.c_int64]
        _lib = L
    return _lib


class Oracle:
    def __init__(self, problem):
        self.p = _abi.copy_params(problem.params)
        st = problem.state()
        self.h = lib().oracle_create(C.byref(self.p), C.byref(st))
        if not self.h:
            raise RuntimeError("oracle_create failed")

    def step(self, itimestep, time_sph, dt):
        if lib().oracle_step(self.h, itimestep, time_sph, dt):
            raise RuntimeError(lib().oracle_last_error(self.h).decode())

    def run(self, first_itimestep, time_sph, dt, nsteps):
        t = C.c_double()
        if lib().oracle_run(self.h, first_itimestep, time_sph, dt, nsteps, C.byref(t)):
            raise RuntimeError(lib().oracle_last_error(self.h).decode())
        return t.value

    def download(self):
        st, arrays = _abi.alloc_state(self.p)
        lib().oracle_download(self.h, C.byref(st))
        return arrays

    def pair_stats(self):
        n = C.c_int64()
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        lib().oracle_pair_stats(self.h, n, a, b, c)
        return dict(npairs=n.value, maxiac=a.value, miniac=b.value, noiac=c.value)

    def pairs(self):
        n = C.c_int64()
        lib().oracle_pairs(self.h, C.byref(n), None, None, None, None, None, None)
        k = n.value
        out = dict(pair_i=np.zeros(k, np.int32), pair_j=np.zeros(k, np.int32), pint_type=np.zeros(k, np.int32),
                   w=np.zeros(k, np.float32), dwdx=np.zeros(k, np.float32), dwdy=np.zeros(k, np.float32))
        lib().oracle_pairs(self.h, C.byref(n), *[out[f].ctypes.data for f in
                                                 ("pair_i", "pair_j", "pint_type", "w", "dwdx", "dwdy")])
        return out

    def list_capacity(self):
        return lib().oracle_get_list_capacity(self.h)

    def set_list_capacity(self, m):
        lib().oracle_set_list_capacity(self.h, int(m))

    def close(self):
        if self.h:
            lib().oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
