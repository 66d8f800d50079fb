"""Host-side problem set-up (reader for input.txt / <name>.dat / <name>.pts) through libspsph_host.so.

Mirrors the reference's Init_sph -> problem_input_data (2_SPH_main_2018.f90:32-46): after load() the arrays
are exactly what the Fortran driver holds at the end of Init_sph, in the same layout and order.
"""
import ctypes as C
import os

import numpy as np

from . import _abi

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_HOST_SO = os.path.join(_PKG, "libspsph_host.so")
_lib = None


def host_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_HOST_SO):
            raise RuntimeError(f"{_HOST_SO} is missing: run `python __graft_entry__.py` (build()) first")
        lib = C.CDLL(_HOST_SO)
        lib.spsph_problem_load.restype = C.c_void_p
        lib.spsph_problem_load.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        lib.spsph_problem_params.restype = C.POINTER(_abi.Params)
        lib.spsph_problem_params.argtypes = [C.c_void_p]
        lib.spsph_problem_state.argtypes = [C.c_void_p, C.POINTER(_abi.State)]
        lib.spsph_problem_nblocks.argtypes = [C.c_void_p]
        lib.spsph_problem_block.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_double)] * 2 + [C.POINTER(C.c_int)] * 4
        lib.spsph_problem_name.restype = C.c_char_p
        lib.spsph_problem_name.argtypes = [C.c_void_p]
        lib.spsph_problem_free.argtypes = [C.c_void_p]
        lib.spsph_problem_gid_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
        lib.spsph_problem_gid_results.argtypes = [C.c_void_p, C.POINTER(_abi.State), C.c_double, C.c_char_p]
        _lib = lib
    return _lib


class Problem:
    """params (ctypes struct) + dict of numpy arrays in reference layout + time blocks."""

    def __init__(self, params, arrays, blocks, name=""):
        self.params = params
        self.arrays = arrays
        self.blocks = blocks  # list of dict(dt, time_end, maxtimestep, print_step, save_step, plot_step)
        self.name = name

    def state(self):
        return _abi.state_from_arrays(self.arrays)

    def copy(self):
        return Problem(_abi.copy_params(self.params), {k: v.copy() for k, v in self.arrays.items()},
                       [dict(b) for b in self.blocks], self.name)


def load(directory, variant):
    """Read input.txt + <name>.dat + <name>.pts in `directory`; variant in {'code','bui','vs','sl'}."""
    lib = host_lib()
    v = _abi.VARIANTS[variant] if isinstance(variant, str) else int(variant)
    err = C.create_string_buffer(512)
    h = lib.spsph_problem_load(os.fsencode(directory), v, err, 512)
    if not h:
        raise RuntimeError("spsph_problem_load: " + err.value.decode())
    try:
        p = _abi.copy_params(lib.spsph_problem_params(h).contents)
        if p.struct_bytes != C.sizeof(_abi.Params):
            raise RuntimeError("spsph_params layout mismatch between include/spsph.h and spsph/_abi.py")
        st = _abi.State()
        lib.spsph_problem_state(h, C.byref(st))
        arrays = {}
        for name, ctype, dt, shape in _abi.STATE_FIELDS:
            shp = shape(p)
            n = int(np.prod(shp))
            ptr = getattr(st, name)
            arrays[name] = np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shp).copy() if n else np.zeros(shp, dt)
        blocks = []
        for k in range(lib.spsph_problem_nblocks(h)):
            dt_, te = C.c_double(), C.c_double()
            a, b, c_, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            lib.spsph_problem_block(h, k, dt_, te, a, b, c_, d)
            blocks.append(dict(dt=dt_.value, time_end=te.value, maxtimestep=a.value, print_step=b.value,
                               save_step=c_.value, plot_step=d.value))
        name = lib.spsph_problem_name(h).decode()
    finally:
        lib.spsph_problem_free(h)
    return Problem(p, arrays, blocks, name)


class GidWriter:
    """<name>.post.msh / <name>.post.res of the reference driver (OutputMesh, OutputRes: host/sph_gid.cpp) for the deck
    in `directory`; frames are appended from downloaded states"""

    def __init__(self, directory, variant, path_prefix):
        lib = host_lib()
        v = _abi.VARIANTS[variant] if isinstance(variant, str) else int(variant)
        err = C.create_string_buffer(512)
        self._h = lib.spsph_problem_load(os.fsencode(directory), v, err, 512)
        if not self._h:
            raise RuntimeError("spsph_problem_load: " + err.value.decode())
        self.prefix = os.fsencode(path_prefix)

    def mesh(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if host_lib().spsph_problem_gid_mesh(self._h, x.ctypes.data, self.prefix):
            raise RuntimeError("gid mesh writer failed")

    def frame(self, arrays, time_sph):
        st = _abi.state_from_arrays(arrays)
        if host_lib().spsph_problem_gid_results(self._h, C.byref(st), float(time_sph), self.prefix):
            raise RuntimeError("gid result writer failed")

    def close(self):
        if self._h:
            host_lib().spsph_problem_free(self._h)
            self._h = None
