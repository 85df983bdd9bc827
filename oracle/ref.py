"""ctypes front end of oracle/_ref/libschwz_ref.so: the reference's OWN SolverRAS
(/root/reference/source/*.cpp compiled unmodified against the stand-ins of
oracle/ref_shim/, see oracle/Makefile) run with one thread per MPI rank.

TEST INFRASTRUCTURE ONLY, same rules as oracle.py.  Used to pin oracle/schwz_oracle.cpp
(and through it the CUDA path) against outputs of the reference itself:
tests/test_ref_pinning.py, tests/golden/make_ref_golden.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libschwz_ref.so")
REFERENCE_ROOT = "/root/reference"

PARTITION = {"regular": 0, "metis": 1, "regular2d": 4}


class RefConfig(C.Structure):
    _fields_ = [
        ("num_subdomains", C.c_int32),
        ("laplacian_n", C.c_int32),
        ("matrix_file", C.c_char_p),
        ("partition", C.c_int32),
        ("overlap", C.c_int32),
        ("max_iters", C.c_int32),
        ("tolerance", C.c_double),
        ("local_tol", C.c_double),
        ("local_max_iters", C.c_int32),
        ("non_symmetric", C.c_int32),
        ("restart_iter", C.c_int32),
        ("enable_onesided", C.c_int32),
        ("remote_put", C.c_int32),
        ("one_by_one", C.c_int32),
        ("conv_tree", C.c_int32),
        ("enable_global_check", C.c_int32),
        ("decentralized_accumulate", C.c_int32),
        ("put_all_local_residual_norms", C.c_int32),
        ("use_mixed_precision", C.c_int32),
        ("local_precond", C.c_char_p),
        ("precond_max_block_size", C.c_int32),
        ("record_iterates", C.c_int32),
        ("run", C.c_int32),
        ("metis_objtype", C.c_char_p),
        ("enable_overlap", C.c_int32),
        ("flush_local", C.c_int32),
        ("lock_local", C.c_int32),
    ]


def available():
    """True when the library exists (prebuilt) or can be built (reference present)."""
    return os.path.exists(_LIB_PATH) or os.path.isdir(os.path.join(REFERENCE_ROOT, "source"))


def build():
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "source")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError("oracle/_ref/libschwz_ref.so is missing and /root/reference is not "
                           "here to build it from")
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.ref_run.restype = C.c_void_p
        L.ref_run.argtypes = [C.POINTER(RefConfig)]
        L.ref_error.restype = C.c_char_p
        L.ref_log.restype = C.c_char_p
        for f in ("ref_free", "ref_error", "ref_log"):
            getattr(L, f).argtypes = [C.c_void_p]
        _lib = L
    return _lib


_VEC = {
    "first_row": np.int32, "permutation": np.int32, "i_permutation": np.int32,
    "l2g": np.int32, "g2l": np.int32, "overlap_row": np.int32,
    "partition_indices": np.uint32,
    "local_rp": np.int32, "local_ci": np.int32, "local_v": np.float64,
    "interface_rp": np.int32, "interface_ci": np.int32, "interface_v": np.float64,
    "global_rp": np.int32, "global_ci": np.int32, "global_v": np.float64,
    "neighbors_in": np.int32, "neighbors_out": np.int32,
    "put_displacements": np.int32, "get_displacements": np.int32,
    "local_rhs": np.float64, "local_residuals": np.float64,
    "local_converged_resnorm": np.float64, "solution": np.float64,
}


class Run:
    """One execution of the reference (bench_ras flag semantics, benchmarking/bench_ras.cpp)."""

    def __init__(self, P, laplacian_n=0, matrix_file=None, partition="regular", overlap=2,
                 max_iters=100, tolerance=1e-6, local_tol=1e-12, local_max_iters=-1,
                 non_symmetric=False, restart_iter=1, enable_onesided=False,
                 remote_comm_type="get", enable_one_by_one=False,
                 global_convergence_type="centralized-tree", enable_global_check=False,
                 enable_accumulate=False, put_all_local_residual_norms=False,
                 use_mixed_precision=False, local_precond="null", precond_max_block_size=16,
                 record_iterates=False, run=True, metis_objtype="null", enable_comm_overlap=False,
                 flush_type="flush-all", lock_type="lock-all"):
        c = RefConfig()
        c.num_subdomains = P
        c.laplacian_n = laplacian_n
        c.matrix_file = (matrix_file or "null").encode()
        c.partition = PARTITION[partition]
        c.overlap = overlap
        c.max_iters = max_iters
        c.tolerance = tolerance
        c.local_tol = local_tol
        c.local_max_iters = local_max_iters
        c.non_symmetric = int(non_symmetric)
        c.restart_iter = restart_iter
        c.enable_onesided = int(enable_onesided)
        c.remote_put = int(remote_comm_type == "put")
        c.one_by_one = int(enable_one_by_one)
        c.conv_tree = int(global_convergence_type == "centralized-tree")
        c.enable_global_check = int(enable_global_check)
        c.decentralized_accumulate = int(enable_accumulate)
        c.put_all_local_residual_norms = int(put_all_local_residual_norms)
        c.use_mixed_precision = int(use_mixed_precision)
        c.local_precond = local_precond.encode()
        c.precond_max_block_size = precond_max_block_size
        c.record_iterates = int(record_iterates)
        c.run = int(run)
        c.metis_objtype = metis_objtype.encode()
        c.enable_overlap = int(enable_comm_overlap)
        c.flush_local = int(flush_type == "flush-local")
        c.lock_local = int(lock_type == "lock-local")
        self.P = P
        self._cfg = c
        self.h = C.c_void_p(lib().ref_run(C.byref(c)))
        err = lib().ref_error(self.h).decode()
        self.log = lib().ref_log(self.h).decode()
        if err:
            raise RuntimeError("reference run failed: " + err)

    def __del__(self):
        try:
            if self.h:
                lib().ref_free(self.h)
                self.h = None
        except Exception:
            pass

    def sizes(self, r):
        out = (C.c_int64 * 8)()
        lib().ref_sizes(self.h, C.c_int(r), out)
        keys = ("global_size", "local_size", "local_size_x", "overlap_size", "nnz_local",
                "nnz_interface", "num_neighbors_in", "num_neighbors_out")
        return dict(zip(keys, (int(x) for x in out)))

    def vec(self, name, r):
        f = getattr(lib(), "ref_" + name)
        n = f(self.h, C.c_int(r), None, C.c_int64(0))
        out = np.zeros(n, _VEC[name])
        if n:
            f(self.h, C.c_int(r), out.ctypes.data_as(C.c_void_p), C.c_int64(n))
        return out

    def _list(self, fn, r, j, dtype=np.int32):
        f = getattr(lib(), fn)
        n = f(self.h, C.c_int(r), C.c_int(j), None, C.c_int64(0))
        out = np.zeros(n, dtype)
        if n:
            f(self.h, C.c_int(r), C.c_int(j), out.ctypes.data_as(C.c_void_p), C.c_int64(n))
        return out

    def get_list(self, r, j):
        return self._list("ref_get_list", r, j)

    def put_list(self, r, j):
        return self._list("ref_put_list", r, j)

    def global_residuals(self, r, j):
        return self._list("ref_global_residuals", r, j, np.float64)

    def iter_count(self, r):
        return int(lib().ref_iter_count(self.h, C.c_int(r)))

    def run_seconds(self, r=0):
        """wall time of SolverRAS::run on rank r (the loop of source/schwarz_base.cpp:384-455
        plus, when it converged, the final residual computation)"""
        f = lib().ref_run_seconds
        f.restype = C.c_double
        return float(f(self.h, C.c_int(r)))

    def num_iterates(self, r):
        return int(lib().ref_num_iterates(self.h, C.c_int(r)))

    def iterate(self, r, k):
        return self._list("ref_iterate", r, k, np.float64)

    def local_matrix(self, r):
        return self.vec("local_rp", r), self.vec("local_ci", r), self.vec("local_v", r)

    def interface_matrix(self, r):
        return self.vec("interface_rp", r), self.vec("interface_ci", r), self.vec("interface_v", r)

    def global_matrix(self, r=0):
        return self.vec("global_rp", r), self.vec("global_ci", r), self.vec("global_v", r)


# ---- kernel-level probes of the Ginkgo stand-in (ref_shim/ref_driver.cpp) -------------------
def _ip(a):
    return a.ctypes.data_as(C.c_void_p)


class Precond:
    """A preconditioner built with the builder calls of source/solve.cpp:486-652 on a CSR."""

    def __init__(self, rp, ci, v, kind, max_block_size=16):
        L = lib()
        L.ref_precond_create.restype = C.c_void_p
        L.ref_precond_csr.restype = C.c_int64
        self.n = len(rp) - 1
        self.rp = np.ascontiguousarray(rp, np.int32)
        self.ci = np.ascontiguousarray(ci, np.int32)
        self.v = np.ascontiguousarray(v, np.float64)
        self.kind = kind
        self.h = C.c_void_p(L.ref_precond_create(C.c_int(self.n), _ip(self.rp), _ip(self.ci),
                                                 _ip(self.v), kind.encode(),
                                                 C.c_int(max_block_size)))

    def __del__(self):
        try:
            if self.h:
                lib().ref_precond_free(self.h)
                self.h = None
        except Exception:
            pass

    def apply(self, b):
        b = np.ascontiguousarray(b, np.float64)
        x = np.zeros(self.n)
        lib().ref_precond_apply(self.h, _ip(b), _ip(x))
        return x

    def block_ptrs(self):
        n = lib().ref_precond_block_ptrs(self.h, None, C.c_int64(0))
        out = np.zeros(n, np.int32)
        lib().ref_precond_block_ptrs(self.h, _ip(out), C.c_int64(n))
        return out

    def blocks(self):
        n = lib().ref_precond_blocks(self.h, None, C.c_int64(0))
        out = np.zeros(n, np.float64)
        lib().ref_precond_blocks(self.h, _ip(out), C.c_int64(n))
        return out

    def csr(self, which):
        """0 L, 1 U of the ILU; 2 / 3 approximate inverses of L / U (ISAI)."""
        nnz = lib().ref_precond_csr(self.h, C.c_int(which), None, None, None)
        if nnz < 0:
            raise ValueError("this preconditioner has no matrix %d" % which)
        rp = np.zeros(self.n + 1, np.int32)
        ci = np.zeros(nnz, np.int32)
        v = np.zeros(nnz, np.float64)
        lib().ref_precond_csr(self.h, C.c_int(which), _ip(rp), _ip(ci), _ip(v))
        return rp, ci, v


def krylov_solve(rp, ci, v, b, x0, max_iters, tol, gmres=False, restart=1, precond=None):
    """gko::solver::Cg / Gmres of the stand-in with the reference's stopping criteria."""
    rp = np.ascontiguousarray(rp, np.int32)
    ci = np.ascontiguousarray(ci, np.int32)
    v = np.ascontiguousarray(v, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    x = np.array(x0, np.float64, copy=True)
    lib().ref_krylov_solve(C.c_int(len(rp) - 1), _ip(rp), _ip(ci), _ip(v), _ip(b), _ip(x),
                           C.c_int(int(gmres)), C.c_int(restart), C.c_int(max_iters),
                           C.c_double(tol), precond.h if precond is not None else None)
    return x
