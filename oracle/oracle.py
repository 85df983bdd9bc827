"""ctypes front end of the CPU oracle (oracle/schwz_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libschwz_oracle.so")

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(_HERE, "schwz_oracle.cpp")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class OrcOptions(C.Structure):
    _fields_ = [
        ("tolerance", C.c_double),
        ("local_tol", C.c_double),
        ("max_iters", C.c_int32),
        ("local_max_iters", C.c_int32),
        ("non_symmetric", C.c_int32),
        ("restart_iter", C.c_int32),
        ("local_solver", C.c_int32),
        ("enable_onesided", C.c_int32),
        ("enable_put", C.c_int32),
        ("enable_one_by_one", C.c_int32),
        ("enable_global_check", C.c_int32),
        ("conv_tree", C.c_int32),
        ("conv_decentralized", C.c_int32),
        ("enable_accumulate", C.c_int32),
        ("iter_offset", C.c_int32),
        ("use_mixed_precision", C.c_int32),
        ("local_precond", C.c_int32),
        ("precond_max_block_size", C.c_int32),
        ("local_factorization", C.c_int32),
    ]


PRECOND = {"null": 0, "block-jacobi": 1, "ilu": 2, "isai": 3}


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_laplacian2d.restype = C.c_int64
        L.orc_laplacian3d.restype = C.c_int64
        L.orc_cholesky.restype = C.c_int64
        L.orc_global_matrix.restype = C.c_int64
        L.orc_factor.restype = C.c_int64
        L.orc_create.restype = C.c_void_p
        for f in ("orc_destroy", "orc_first_row", "orc_permutation",
                  "orc_global_matrix", "orc_sizes", "orc_l2g", "orc_g2l",
                  "orc_local_matrix", "orc_interface_matrix", "orc_neighbors",
                  "orc_get_list", "orc_put_list", "orc_displacements",
                  "orc_set_rhs", "orc_configure", "orc_step", "orc_run",
                  "orc_iter_count", "orc_x", "orc_local_solution",
                  "orc_local_rhs", "orc_rank_status", "orc_history",
                  "orc_final_residual", "orc_factor"):
            getattr(L, f).argtypes = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def set_threads(n):
    lib().orc_set_threads(int(n))


def max_threads():
    return int(lib().orc_max_threads())


def set_rank_threads(n):
    """Subdomains stepped side by side by Problem.step() (synchronous mode): 1 = one after the
    other with the whole OpenMP team each; n > 1 = n at once with set_threads() threads each."""
    lib().orc_set_rank_threads(int(n))


def laplacian2d(n):
    """source/initialization.cpp:214-265 -> (rowptr, col, val) int32/int32/f64."""
    N = n * n
    rp = np.zeros(N + 1, np.int32)
    ci = np.zeros(5 * N, np.int32)
    v = np.zeros(5 * N, np.float64)
    nnz = lib().orc_laplacian2d(C.c_int(n), _p(rp), _p(ci), _p(v))
    return rp, ci[:nnz].copy(), v[:nnz].copy()


def laplacian3d(n):
    N = n ** 3
    rp = np.zeros(N + 1, np.int32)
    ci = np.zeros(7 * N, np.int32)
    v = np.zeros(7 * N, np.float64)
    nnz = lib().orc_laplacian3d(C.c_int(n), _p(rp), _p(ci), _p(v))
    return rp, ci[:nnz].copy(), v[:nnz].copy()


def partition_regular2d(N, P):
    pi = np.zeros(N, np.uint32)
    lib().orc_partition_regular2d(C.c_int64(N), C.c_int(P), _p(pi))
    return pi


def partition_regular2d_rect(N, P, px=0, py=0):
    """px x py rectangular extension of regular2d (0, 0: most square factorisation, px <= py);
    equals partition_regular2d for perfect-square P."""
    pi = np.zeros(N, np.uint32)
    if lib().orc_partition_regular2d_rect(C.c_int64(N), C.c_int(P), C.c_int(px), C.c_int(py),
                                          _p(pi)) != 0:
        raise ValueError("regular2d: N must be a square and px * py == P")
    return pi


def spmv(rp, ci, v, x, alpha=1.0, beta=0.0, y=None):
    n = len(rp) - 1
    out = np.zeros(n) if y is None else np.array(y, dtype=np.float64)
    lib().orc_spmv(C.c_int32(n), _p(rp), _p(ci), _p(v), C.c_double(alpha),
                   _p(np.ascontiguousarray(x, np.float64)), C.c_double(beta),
                   _p(out))
    return out


class Precond:
    """Local preconditioner of the iterative local solve (source/solve.cpp:486-652):
    kind in PRECOND.  Restated Ginkgo semantics, pinned to oracle/_ref's stand-in."""

    def __init__(self, rp, ci, v, kind, max_block_size=16):
        L = lib()
        L.orc_precond_create.restype = C.c_void_p
        for f in ("orc_precond_block_ptrs", "orc_precond_blocks", "orc_precond_csr"):
            getattr(L, f).restype = C.c_int64
        self.n = len(rp) - 1
        self.kind = kind
        self.h = C.c_void_p(L.orc_precond_create(
            C.c_int32(self.n), _p(np.ascontiguousarray(rp, np.int32)),
            _p(np.ascontiguousarray(ci, np.int32)), _p(np.ascontiguousarray(v, np.float64)),
            C.c_int(PRECOND[kind]), C.c_int(max_block_size)))

    def __del__(self):
        try:
            if self.h:
                lib().orc_precond_free(self.h)
                self.h = None
        except Exception:
            pass

    def apply(self, r):
        r = np.ascontiguousarray(r, np.float64)
        z = np.zeros(self.n)
        lib().orc_precond_apply(self.h, C.c_int32(self.n), _p(r), _p(z))
        return z

    def block_ptrs(self):
        out = np.zeros(lib().orc_precond_block_ptrs(self.h, None), np.int32)
        lib().orc_precond_block_ptrs(self.h, _p(out))
        return out

    def blocks(self):
        out = np.zeros(lib().orc_precond_blocks(self.h, None), np.float64)
        lib().orc_precond_blocks(self.h, _p(out))
        return out

    def csr(self, which):
        """0 L, 1 U of the ILU; 2 / 3 approximate inverses of L / U (ISAI)."""
        nnz = lib().orc_precond_csr(self.h, C.c_int(which), None, None, None)
        rp = np.zeros(self.n + 1, np.int32)
        ci = np.zeros(nnz, np.int32)
        v = np.zeros(nnz, np.float64)
        lib().orc_precond_csr(self.h, C.c_int(which), _p(rp), _p(ci), _p(v))
        return rp, ci, v


def cg(rp, ci, v, b, x0, max_iters, factor, precond=None):
    n = len(rp) - 1
    x = np.array(x0, dtype=np.float64)
    it = lib().orc_cg_pc(C.c_int32(n), _p(rp), _p(ci), _p(v),
                         _p(np.ascontiguousarray(b, np.float64)), _p(x),
                         C.c_int(max_iters), C.c_double(factor),
                         precond.h if precond is not None else None)
    return x, int(it)


def gmres(rp, ci, v, b, x0, max_iters, factor, restart, precond=None):
    n = len(rp) - 1
    x = np.array(x0, dtype=np.float64)
    it = lib().orc_gmres_pc(C.c_int32(n), _p(rp), _p(ci), _p(v),
                            _p(np.ascontiguousarray(b, np.float64)), _p(x),
                            C.c_int(max_iters), C.c_double(factor),
                            C.c_int(restart), precond.h if precond is not None else None)
    return x, int(it)


def cholesky(rp, ci, v, perm):
    n = len(rp) - 1
    perm = np.ascontiguousarray(perm, np.int32)
    nnz = lib().orc_cholesky(C.c_int32(n), _p(rp), _p(ci), _p(v), _p(perm),
                             None, None, None)
    if nnz < 0:
        raise ValueError("matrix not SPD")
    Lrp = np.zeros(n + 1, np.int32)
    Lci = np.zeros(nnz, np.int32)
    Lv = np.zeros(nnz, np.float64)
    lib().orc_cholesky(C.c_int32(n), _p(rp), _p(ci), _p(v), _p(perm), _p(Lrp),
                       _p(Lci), _p(Lv))
    return Lrp, Lci, Lv


def trs(rp, ci, v, b, upper):
    n = len(rp) - 1
    x = np.zeros(n)
    lib().orc_trs(C.c_int32(n), _p(rp), _p(ci), _p(v), C.c_int(int(upper)),
                  _p(np.ascontiguousarray(b, np.float64)), _p(x))
    return x


class Problem:
    """All subdomains of one RAS problem (the oracle runs them in one process).

    partition: "regular" (1-D split, reference default), or "permute" with an
    explicit part-id vector (what metis / regular2d feed into
    source/restricted_schwarz.cpp:105-152).
    """

    def __init__(self, rp, ci, v, P, part=None, overlap=2):
        self.N = len(rp) - 1
        self.P = P
        self.rp = np.ascontiguousarray(rp, np.int32)
        self.ci = np.ascontiguousarray(ci, np.int32)
        self.v = np.ascontiguousarray(v, np.float64)
        kind = 0 if part is None else 1
        self.part = None if part is None else np.ascontiguousarray(part, np.uint32)
        self.h = C.c_void_p(lib().orc_create(
            C.c_int32(self.N), _p(self.rp), _p(self.ci), _p(self.v), C.c_int(P),
            C.c_int(kind), _p(self.part), C.c_int(overlap)))

    def __del__(self):
        try:
            if self.h:
                lib().orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- index sets -----------------------------------------------------
    def first_row(self):
        out = np.zeros(self.P + 1, np.int32)
        lib().orc_first_row(self.h, _p(out))
        return out

    def permutation(self):
        perm = np.zeros(self.N, np.int32)
        iperm = np.zeros(self.N, np.int32)
        ok = lib().orc_permutation(self.h, _p(perm), _p(iperm))
        return (perm, iperm) if ok else (None, None)

    def global_matrix(self):
        nnz = lib().orc_global_matrix(self.h, None, None, None)
        rp = np.zeros(self.N + 1, np.int32)
        ci = np.zeros(nnz, np.int32)
        v = np.zeros(nnz, np.float64)
        lib().orc_global_matrix(self.h, _p(rp), _p(ci), _p(v))
        return rp, ci, v

    def sizes(self, r):
        out = np.zeros(8, np.int64)
        lib().orc_sizes(self.h, C.c_int(r), _p(out))
        keys = ("local_size", "local_size_x", "overlap_size", "nnz_local",
                "nnz_interface", "n_halo", "num_neighbors_in",
                "num_neighbors_out")
        return dict(zip(keys, (int(x) for x in out)))

    def l2g(self, r):
        s = self.sizes(r)
        out = np.zeros(s["local_size_x"] + s["n_halo"], np.int32)
        lib().orc_l2g(self.h, C.c_int(r), _p(out))
        return out

    def g2l(self, r):
        out = np.zeros(self.N, np.int32)
        lib().orc_g2l(self.h, C.c_int(r), _p(out))
        return out

    def local_matrix(self, r):
        s = self.sizes(r)
        rp = np.zeros(s["local_size_x"] + 1, np.int32)
        ci = np.zeros(s["nnz_local"], np.int32)
        v = np.zeros(s["nnz_local"], np.float64)
        lib().orc_local_matrix(self.h, C.c_int(r), _p(rp), _p(ci), _p(v))
        return rp, ci, v

    def interface_matrix(self, r):
        s = self.sizes(r)
        rp = np.zeros(s["local_size_x"] + 1, np.int32)
        ci = np.zeros(s["nnz_interface"], np.int32)
        v = np.zeros(s["nnz_interface"], np.float64)
        n = lib().orc_interface_matrix(self.h, C.c_int(r), _p(rp), _p(ci), _p(v))
        if n == 0:
            return np.zeros(1, np.int32), ci, v
        return rp, ci, v

    def neighbors(self, r):
        s = self.sizes(r)
        nin = np.zeros(max(s["num_neighbors_in"], 1), np.int32)
        nout = np.zeros(max(s["num_neighbors_out"], 1), np.int32)
        lib().orc_neighbors(self.h, C.c_int(r), _p(nin), _p(nout))
        return nin[:s["num_neighbors_in"]], nout[:s["num_neighbors_out"]]

    def get_list(self, r, j):
        n = lib().orc_get_list(self.h, C.c_int(r), C.c_int(j), None)
        out = np.zeros(n, np.int32)
        lib().orc_get_list(self.h, C.c_int(r), C.c_int(j), _p(out))
        return out

    def put_list(self, r, j):
        n = lib().orc_put_list(self.h, C.c_int(r), C.c_int(j), None)
        out = np.zeros(n, np.int32)
        lib().orc_put_list(self.h, C.c_int(r), C.c_int(j), _p(out))
        return out

    def displacements(self, r):
        pd = np.zeros(self.P + 1, np.int32)
        gd = np.zeros(self.P + 1, np.int32)
        lib().orc_displacements(self.h, C.c_int(r), _p(pd), _p(gd))
        return pd, gd

    # ---- run -------------------------------------------------------------
    def set_rhs(self, rhs):
        rhs = np.ascontiguousarray(rhs, np.float64)
        assert rhs.shape[0] == self.N
        lib().orc_set_rhs(self.h, _p(rhs))

    def configure(self, tolerance=1e-6, local_tol=1e-12, max_iters=100,
                  local_max_iters=-1, non_symmetric=False, restart_iter=1,
                  local_solver="iterative-ginkgo", enable_onesided=False,
                  remote_comm_type="get", enable_one_by_one=False,
                  enable_global_check=False,
                  global_convergence_type="centralized-tree",
                  enable_accumulate=False, iter_offset=False, factor_perms=None,
                  use_mixed_precision=False, local_precond="null",
                  precond_max_block_size=16, local_factorization="cholmod"):
        o = OrcOptions()
        o.tolerance = tolerance
        o.local_tol = local_tol
        o.max_iters = max_iters
        o.local_max_iters = local_max_iters
        o.non_symmetric = int(non_symmetric)
        o.restart_iter = restart_iter
        o.local_solver = {"iterative-ginkgo": 2, "direct-ginkgo": 1}[local_solver]
        o.enable_onesided = int(enable_onesided)
        o.enable_put = int(remote_comm_type == "put")
        o.enable_one_by_one = int(enable_one_by_one)
        o.enable_global_check = int(enable_global_check)
        o.conv_tree = int(global_convergence_type == "centralized-tree")
        o.conv_decentralized = int(global_convergence_type == "decentralized")
        o.enable_accumulate = int(enable_accumulate)
        o.iter_offset = int(iter_offset)
        o.use_mixed_precision = int(use_mixed_precision)
        o.local_precond = PRECOND[local_precond]
        o.precond_max_block_size = precond_max_block_size
        o.local_factorization = {"cholmod": 0, "umfpack": 1}[local_factorization]
        perm_all = None
        if factor_perms is not None:
            perm_all = np.ascontiguousarray(np.concatenate(factor_perms), np.int32)
        rc = lib().orc_configure(self.h, C.byref(o), _p(perm_all))
        if rc != 0:
            raise RuntimeError("oracle configure failed (factorisation)")
        self.max_iters = max_iters

    def step(self):
        return int(lib().orc_step(self.h))

    def run(self):
        return int(lib().orc_run(self.h))

    def iter_count(self):
        return int(lib().orc_iter_count(self.h))

    def x(self, r):
        out = np.zeros(self.N)
        lib().orc_x(self.h, C.c_int(r), _p(out))
        return out

    def local_solution(self, r):
        out = np.zeros(self.sizes(r)["local_size_x"])
        lib().orc_local_solution(self.h, C.c_int(r), _p(out))
        return out

    def local_rhs(self, r):
        out = np.zeros(self.sizes(r)["local_size_x"])
        lib().orc_local_rhs(self.h, C.c_int(r), _p(out))
        return out

    def status(self, r):
        out = np.zeros(8)
        lib().orc_rank_status(self.h, C.c_int(r), _p(out))
        keys = ("resnorm", "resnorm0", "gres", "gres0", "num_converged",
                "finished", "finished_iter", "last_local_iters")
        return dict(zip(keys, out.tolist()))

    def history(self, r):
        n = lib().orc_history(self.h, C.c_int(r), None, None, None)
        res = np.zeros(n)
        gres = np.zeros(n)
        lib().orc_history(self.h, C.c_int(r), _p(res), _p(gres), None)
        return res, gres

    def final_residual(self):
        x = np.zeros(self.N)
        out = np.zeros(4)
        lib().orc_final_residual(self.h, _p(x), _p(out))
        return x, dict(residual_norm=out[0], rhs_norm=out[1], sol_norm=out[2],
                       relative=out[3])

    def factor(self, r):
        nnz = lib().orc_factor(self.h, C.c_int(r), None, None, None)
        n = self.sizes(r)["local_size_x"]
        rp = np.zeros(n + 1, np.int32)
        ci = np.zeros(nnz, np.int32)
        v = np.zeros(nnz, np.float64)
        lib().orc_factor(self.h, C.c_int(r), _p(rp), _p(ci), _p(v))
        return rp, ci, v
