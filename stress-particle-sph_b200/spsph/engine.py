"""Device engine: ctypes binding of the C-ABI in include/spsph.h (libspsph_cuda.so).

Engine mirrors the reference driver's three call sites (1_SPH_2018.f90:132,173,156): upload after set-up,
step == time_integration, download before the writers. No CPU fallback: without the CUDA library or a GPU
the constructor raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CUDA_SO = os.environ.get("SPSPH_CUDA_SO") or os.path.join(_PKG, "libspsph_cuda.so")  # override: kernel tuning experiments
_lib = None


def cuda_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_CUDA_SO):
            raise RuntimeError(f"{_CUDA_SO} is missing: the CUDA extension must be built (python __graft_entry__.py); "
                               "there is no CPU fallback")
        L = C.CDLL(_CUDA_SO)
        H = C.c_void_p
        L.spsph_create.argtypes = [C.POINTER(H), C.POINTER(_abi.Params), C.c_int]
        L.spsph_upload.argtypes = [H, C.POINTER(_abi.State)]
        L.spsph_step.argtypes = [H, C.c_int32, C.c_double, C.c_double]
        L.spsph_run.argtypes = [H, C.c_int32, C.c_double, C.c_double, C.c_int32, C.POINTER(C.c_double)]
        L.spsph_download.argtypes = [H, C.POINTER(_abi.State)]
        L.spsph_pair_stats.argtypes = [H, C.POINTER(C.c_int64)] + [C.POINTER(C.c_int32)] * 3
        L.spsph_pairs.argtypes = [H, C.POINTER(C.c_int64)] + [C.c_void_p] * 6
        L.spsph_last_run_ms.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_int64)]
        L.spsph_sync.argtypes = [H]
        L.spsph_get_list_capacity.argtypes = [H, C.POINTER(C.c_int64)]
        L.spsph_set_list_capacity.argtypes = [H, C.c_int64]
        L.spsph_path_counts.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.spsph_upload_rows.argtypes = [H, C.POINTER(_abi.State), C.c_void_p, C.c_int32]
        L.spsph_download_rows.argtypes = [H, C.POINTER(_abi.State), C.c_void_p, C.c_int32]
        L.spsph_download_frame.argtypes = [H, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.spsph_dist_unique_id.argtypes = [C.c_char_p]
        L.spsph_dist_init.argtypes = [H, C.c_int32, C.c_int32, C.c_char_p, C.POINTER(C.c_double), C.c_int32, C.c_int32]
        L.spsph_dist_flags.argtypes = [H, C.c_void_p]
        L.spsph_dist_set_planes.argtypes = [H, C.POINTER(C.c_double)]
        L.spsph_local_counts.argtypes = [H, C.c_void_p]
        L.spsph_profile.argtypes = [H, C.c_int]
        L.spsph_profile_get.argtypes = [H, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.spsph_destroy.argtypes = [H]
        L.spsph_last_error.restype = C.c_char_p
        L.spsph_last_error.argtypes = [H]
        L.spsph_version.restype = C.c_char_p
        _lib = L
    return _lib


EXPORTS = ["spsph_create", "spsph_upload", "spsph_step", "spsph_run", "spsph_download", "spsph_pair_stats",
           "spsph_pairs", "spsph_last_run_ms", "spsph_sync", "spsph_profile", "spsph_profile_get",
           "spsph_dist_unique_id", "spsph_dist_init", "spsph_dist_flags", "spsph_dist_set_planes", "spsph_local_counts", "spsph_get_list_capacity",
           "spsph_set_list_capacity", "spsph_path_counts", "spsph_upload_rows",
           "spsph_download_rows", "spsph_download_frame", "spsph_destroy", "spsph_last_error", "spsph_version"]


# column codes of spsph_download_frame (include/spsph.h, SPSPH_COL_*)
FRAME_COLS = {n: k for k, n in enumerate(["x", "y", "vx", "vy", "sxx", "syy", "sxy", "szz", "epsp", "disp_10", "rho",
                                          "hsml", "displ_x", "displ_y", "f_drucker", "bc_or_not"])}

# row sets of the time-varying state: (rows of all ids, of ids < ntotal, of ids < nnode)
_ROWS_ALL = ("x", "vel", "stress", "if_out_domain")
_ROWS_PART = ("internal_vars", "f_drucker", "bc_or_not")
_ROWS_NODE = ("displ", "x_10", "disp_10", "n_int", "bc_int")


def _row_selector(p, ids, key):
    if key in _ROWS_ALL:
        return ids
    if key in _ROWS_PART:
        return ids[ids < p.ntotal]
    if key in _ROWS_NODE:
        return ids[ids < p.nnode]
    raise KeyError(f"{key} is not part of the time-varying state")


def row_arrays(p, arrays, ids, keys=_ROWS_ALL + _ROWS_PART + _ROWS_NODE):
    """compact copies of the rows `ids` of full arrays, in the layout spsph_upload_rows expects"""
    ids = np.asarray(ids, dtype=np.int32)
    return {k: np.ascontiguousarray(arrays[k][_row_selector(p, ids, k)]) for k in keys if k in arrays}


def alloc_rows(p, ids, keys):
    ids = np.asarray(ids, dtype=np.int32)
    spec = {name: (dt, shape(p)) for name, _, dt, shape in _abi.STATE_FIELDS}
    return {k: np.zeros((len(_row_selector(p, ids, k)),) + tuple(spec[k][1][1:]), spec[k][0]) for k in keys}


def dist_unique_id():
    buf = C.create_string_buffer(128)
    if cuda_lib().spsph_dist_unique_id(buf):
        raise RuntimeError("ncclGetUniqueId failed (libnccl.so.2 not loadable?)")
    return buf.raw


class Engine:
    def __init__(self, problem, device=0, upload=True):
        self.L = cuda_lib()
        self.p = _abi.copy_params(problem.params)
        self.h = C.c_void_p()
        rc = self.L.spsph_create(C.byref(self.h), C.byref(self.p), device)
        if rc:
            msg = self.L.spsph_last_error(self.h).decode() if self.h else "spsph_create failed"
            if self.h:
                self.L.spsph_destroy(self.h)
                self.h = None
            raise RuntimeError(msg)
        if upload:
            self.upload(problem.arrays)

    def _chk(self, rc):
        if rc:
            raise RuntimeError(self.L.spsph_last_error(self.h).decode())

    def upload(self, arrays):
        st = _abi.state_from_arrays(arrays)
        self._chk(self.L.spsph_upload(self.h, C.byref(st)))

    def upload_rows(self, rows, ids):
        """time-varying state of the particles `ids` (ascending int32): rows = dict of compact arrays (see row_arrays)"""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        st = _abi.state_from_arrays(rows)
        self._chk(self.L.spsph_upload_rows(self.h, C.byref(st), ids.ctypes.data, len(ids)))

    def download_rows(self, ids, rows=None, keys=("x", "vel", "stress", "internal_vars", "displ")):
        """-> dict of compact arrays holding the rows of the particles `ids` (ascending int32)"""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        if rows is None:
            rows = alloc_rows(self.p, ids, keys)
        st = _abi.state_from_arrays(rows)
        self._chk(self.L.spsph_download_rows(self.h, C.byref(st), ids.ctypes.data, len(ids)))
        return rows

    def download_frame(self, cols, first=0, count=None, out=None):
        """output frame packed on the device: (count, len(cols)) float64 table of the columns `cols` (names or codes of
        FRAME_COLS, in the order a writer prints them) for the particles first .. first+count-1, one transfer"""
        codes = np.ascontiguousarray([FRAME_COLS[c] if isinstance(c, str) else int(c) for c in cols], dtype=np.int32)
        if count is None:
            count = self.p.ntotal2 - first
        if out is None:
            out = np.zeros((count, len(codes)), np.float64)
        assert out.dtype == np.float64 and out.flags["C_CONTIGUOUS"] and out.size == count * len(codes)
        self._chk(self.L.spsph_download_frame(self.h, codes.ctypes.data, len(codes), first, count, out.ctypes.data))
        return out

    def step(self, itimestep, time_sph, dt):
        self._chk(self.L.spsph_step(self.h, itimestep, time_sph, dt))

    def run(self, first_itimestep, time_sph, dt, nsteps):
        t = C.c_double()
        self._chk(self.L.spsph_run(self.h, first_itimestep, time_sph, dt, nsteps, C.byref(t)))
        return t.value

    def last_run(self):
        ms, n = C.c_float(), C.c_int64()
        self.L.spsph_last_run_ms(self.h, C.byref(ms), C.byref(n))
        return ms.value, n.value

    def profile(self, enable=True):
        self._chk(self.L.spsph_profile(self.h, 1 if enable else 0))

    def profile_get(self):
        """{kernel name: (total_ms, launches)} accumulated since profile(True)"""
        out = {}
        kid = 0
        while True:
            name, ms, n = C.c_char_p(), C.c_double(), C.c_int64()
            if self.L.spsph_profile_get(self.h, kid, C.byref(name), C.byref(ms), C.byref(n)):
                break
            out[name.value.decode()] = (ms.value, n.value)
            kid += 1
        return out

    def dist_init(self, rank, nranks, unique_id, plan):
        """join the x-slab decomposition (spsph.dist.plan_slabs); unique_id: 128 bytes from dist_unique_id()"""
        planes = np.ascontiguousarray(plan["planes"], dtype=np.float64)
        assert len(planes) == nranks + 1 and len(unique_id) == 128
        self.in_dist_mode = True
        self._chk(self.L.spsph_dist_init(self.h, rank, nranks, bytes(unique_id),
                                         planes.ctypes.data_as(C.POINTER(C.c_double)), int(plan["halo_cells"]),
                                         int(plan["halo_capacity"])))

    def set_planes(self, planes):
        """new slab planes for the following steps (dynamic re-slabbing; see spsph.dist.rebalance)"""
        planes = np.ascontiguousarray(planes, dtype=np.float64)
        self._chk(self.L.spsph_dist_set_planes(self.h, planes.ctypes.data_as(C.POINTER(C.c_double))))

    def local_counts(self):
        n = np.zeros(3, np.int32)
        self._chk(self.L.spsph_local_counts(self.h, n.ctypes.data))
        return [int(v) for v in n]

    def dist_flags(self):
        f = np.zeros(self.p.ntotal2, np.int32)
        self._chk(self.L.spsph_dist_flags(self.h, f.ctypes.data))
        return f

    def sync(self):
        self._chk(self.L.spsph_sync(self.h))

    def list_capacity(self):
        """length the reference's pair list has grown to (part of a checkpoint, see checkpoint.py)"""
        n = C.c_int64()
        self._chk(self.L.spsph_get_list_capacity(self.h, C.byref(n)))
        return n.value

    def path_counts(self):
        """(steps on the cell-tile kernels, steps on the id-list kernels) since the engine was created"""
        a, b = C.c_int64(), C.c_int64()
        self._chk(self.L.spsph_path_counts(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_list_capacity(self, m_pairs):
        self._chk(self.L.spsph_set_list_capacity(self.h, int(m_pairs)))

    def download(self, arrays=None):
        if arrays is None:
            st, arrays = _abi.alloc_state(self.p)
        else:
            st = _abi.state_from_arrays(arrays)
        self._chk(self.L.spsph_download(self.h, C.byref(st)))
        return arrays

    def pair_stats(self):
        n = C.c_int64()
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self._chk(self.L.spsph_pair_stats(self.h, n, a, b, c))
        return dict(npairs=n.value, maxiac=a.value, miniac=b.value, noiac=c.value)

    def pairs(self):
        n = C.c_int64()
        self._chk(self.L.spsph_pairs(self.h, C.byref(n), None, None, None, None, None, None))
        k = n.value
        out = dict(pair_i=np.zeros(k, np.int32), pair_j=np.zeros(k, np.int32), pint_type=np.zeros(k, np.int32),
                   w=np.zeros(k, np.float32), dwdx=np.zeros(k, np.float32), dwdy=np.zeros(k, np.float32))
        self._chk(self.L.spsph_pairs(self.h, C.byref(n), *[out[f].ctypes.data for f in
                                                            ("pair_i", "pair_j", "pint_type", "w", "dwdx", "dwdy")]))
        return out

    def close(self):
        if getattr(self, "h", None):
            self.L.spsph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
