"""ctypes mirror of include/spsph.h (spsph_params, spsph_state). Keep in sync with the header."""
import ctypes as C

import numpy as np

MAX_TCURVES = 8
MAX_TCURVE_PTS = 128
MAX_BCS = 16
NPROP = 20
NINT_VARS = 10

VARIANT_CODE, VARIANT_BUI, VARIANT_VS, VARIANT_SL = 0, 1, 2, 3
VARIANTS = {"code": VARIANT_CODE, "bui": VARIANT_BUI, "vs": VARIANT_VS, "sl": VARIANT_SL}


class Params(C.Structure):
    _fields_ = (
        [("struct_bytes", C.c_int32), ("variant", C.c_int32)]
        + [(n, C.c_int32) for n in (
            "ndimn", "nstre", "nnode", "nstress", "ntotal", "ntotal2", "ndummy", "ndummy2", "npoints",
            "sp_sph", "inside_approach", "sph_shift", "vel_vector", "shift_update", "dummy_nodes",
            "skf", "sle", "cspm", "update_x", "xsph", "cont_density", "art_stress",
            "ntype_eco", "ncrit", "ntype_solid", "no_bcs", "ifsigman", "ic_grav", "tcurve_grav",
            "bc_loop_ntotal")]
        + [("ae_threshold", C.c_float), ("ntcurves", C.c_int32), ("nptstcurves", C.c_int32 * MAX_TCURVES)]
        + [(n, C.c_double) for n in ("dx", "dy", "sml", "r_x", "r_y", "disp_tol", "alpha", "beta", "damping", "ft_grav")]
        + [("cgrav", C.c_double * 2), ("props", C.c_double * NPROP)]
        + [(n, C.c_double) for n in ("D11", "D22", "D12", "D33", "D41", "D42")]
        + [("xmin_domain", C.c_double * 2), ("xmax_domain", C.c_double * 2), ("pi", C.c_double),
           ("ttcurves", (C.c_double * MAX_TCURVE_PTS) * MAX_TCURVES),
           ("ftcurves", (C.c_float * MAX_TCURVE_PTS) * MAX_TCURVES),
           ("bc_list", (C.c_double * 8) * MAX_BCS)]
    )


_D, _F, _I = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32)

# (name, ctype pointer, numpy dtype, shape as function of params)
STATE_FIELDS = [
    ("x", _D, np.float64, lambda p: (p.ntotal2, 2)),
    ("vel", _D, np.float64, lambda p: (p.ntotal2, 2)),
    ("stress", _D, np.float64, lambda p: (p.ntotal2, 4)),
    ("rho", _D, np.float64, lambda p: (p.ntotal2,)),
    ("mass", _D, np.float64, lambda p: (p.ntotal2,)),
    ("hsml", _D, np.float64, lambda p: (p.ntotal2,)),
    ("itype", _I, np.int32, lambda p: (p.ntotal2,)),
    ("internal_vars", _D, np.float64, lambda p: (p.ntotal, NINT_VARS)),
    ("f_drucker", _D, np.float64, lambda p: (p.ntotal,)),
    ("x00", _D, np.float64, lambda p: (p.ntotal2, 2)),
    ("displ", _D, np.float64, lambda p: (p.nnode, 2)),
    ("x_10", _D, np.float64, lambda p: (p.nnode, 2)),
    ("disp_10", _D, np.float64, lambda p: (p.nnode,)),
    ("wall_position", _F, np.float32, lambda p: (p.ntotal2,)),
    ("horizontal_or_not", _F, np.float32, lambda p: (p.ntotal2,)),
    ("n_int", _F, np.float32, lambda p: (p.nnode,)),
    ("bc_int", _I, np.int32, lambda p: (p.nnode,)),
    ("if_out_domain", _I, np.int32, lambda p: (p.ntotal2,)),
    ("bc_or_not", _I, np.int32, lambda p: (p.ntotal,)),
    ("bc_info", _I, np.int32, lambda p: (p.ntotal, 8)),
]


class State(C.Structure):
    _fields_ = [(n, t) for n, t, _, _ in STATE_FIELDS]


def alloc_state(p, pinned=False):
    """Allocate host arrays for every field; returns (State, dict of numpy arrays).

    Arrays are C-contiguous with the particle index slowest, which is byte-identical to the
    reference's Fortran column-major x(ndimn,ntotal2) etc.
    """
    arrays = {}
    st = State()
    for name, ctype, dt, shape in STATE_FIELDS:
        a = np.zeros(shape(p), dtype=dt)
        arrays[name] = a
        setattr(st, name, a.ctypes.data_as(ctype))
    return st, arrays


def state_from_arrays(arrays):
    st = State()
    for name, ctype, dt, _ in STATE_FIELDS:
        a = arrays.get(name)
        if a is None:
            setattr(st, name, ctype())
        else:
            assert a.dtype == dt and a.flags["C_CONTIGUOUS"], name
            setattr(st, name, a.ctypes.data_as(ctype))
    return st


def copy_params(p):
    q = Params()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(Params))
    return q
