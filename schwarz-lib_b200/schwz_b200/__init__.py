"""schwz_b200 — Python front end of libschwz_b200.so (the C ABI declared in
include/schwz_b200.h).

This module is plumbing: it loads the shared library, wraps handles and moves
numpy arrays across the boundary.  All computation happens in the hand-written
sm_100a kernels behind the C ABI; there is NO CPU fallback — if the library is
missing, loading fails loudly, and every device entry point fails when no GPU
is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libschwz_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "schwz_b200.h")


class SchwzError(RuntimeError):
    pass


class RasOptions(C.Structure):
    _fields_ = [("tolerance", C.c_double), ("local_tol", C.c_double),
                ("local_max_iters", C.c_int32), ("local_solver", C.c_int32),
                ("non_symmetric", C.c_int32), ("restart_iter", C.c_int32),
                ("overlap", C.c_int32), ("use_mixed_precision", C.c_int32),
                ("local_precond", C.c_int32), ("precond_max_block_size", C.c_int32)]


# metadata.local_precond of the reference (--local_precond of bench_ras)
PRECOND = {"null": 0, "block-jacobi": 1, "ilu": 2, "isai": 3}


class MailboxLayout(C.Structure):
    _fields_ = [("recv_stride", C.c_int64), ("flags_off", C.c_int64),
                ("conv_off", C.c_int64), ("err_off", C.c_int64), ("send_off", C.c_int64),
                ("slots_off", C.c_int64), ("x_off", C.c_int64), ("bytes", C.c_int64)]

    def as_tuple(self):
        return (self.recv_stride, self.flags_off, self.conv_off, self.err_off, self.send_off,
                self.slots_off, self.x_off, self.bytes)

    @classmethod
    def from_tuple(cls, t):
        return cls(*t)


class LoopOptions(C.Structure):
    _fields_ = [("num_subdomains", C.c_int32), ("max_iters", C.c_int32),
                ("tolerance", C.c_double), ("enable_onesided", C.c_int32),
                ("enable_global_check", C.c_int32), ("conv_decentralized", C.c_int32),
                ("iter_offset", C.c_int32), ("exchange_mode", C.c_int32),
                ("comm", C.c_void_p)]


class LoopResult(C.Structure):
    _fields_ = [("iters", C.c_int32), ("converged", C.c_int32),
                ("global_resnorm", C.c_double), ("global_resnorm0", C.c_double),
                ("elapsed_s", C.c_double), ("host_stream_syncs", C.c_int32),
                ("host_event_waits", C.c_int32)]


_lib = None


def load():
    """Load libschwz_b200.so; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SchwzError(
                "libschwz_b200.so not built (%s); run `make -C schwarz-lib_b200` or "
                "__graft_entry__.build()" % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        L.schwz_b200_last_error.restype = C.c_char_p
        for name in ("schwz_b200_launch_count", "schwz_b200_spmv_bytes",
                     "schwz_b200_ras_kernel_bytes",
                     "schwz_b200_host_cholesky", "schwz_b200_laplacian2d",
                     "schwz_b200_laplacian3d", "schwz_b200_precond_bytes_per_apply",
                     "schwz_b200_precond_block_ptrs", "schwz_b200_precond_blocks",
                     "schwz_b200_precond_csr"):
            getattr(L, name).restype = C.c_int64
        L.schwz_b200_host_free.restype = None
        _lib = L
    return _lib


def _chk(rc):
    if rc != 0:
        raise SchwzError(load().schwz_b200_last_error().decode())


def _p(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def device_count():
    n = C.c_int(0)
    _chk(load().schwz_b200_device_count(C.byref(n)))
    return n.value


def launch_count():
    return int(load().schwz_b200_launch_count())


# --------------------------------------------------------------------------
# host-side index sets (no GPU needed)
# --------------------------------------------------------------------------
def laplacian2d(n):
    N = n * n
    rp = np.zeros(N + 1, np.int32)
    ci = np.zeros(5 * N, np.int32)
    v = np.zeros(5 * N, np.float64)
    nnz = load().schwz_b200_laplacian2d(C.c_int32(n), _p(rp), _p(ci), _p(v))
    return rp, ci[:nnz].copy(), v[:nnz].copy()


def laplacian3d(n):
    N = n ** 3
    rp = np.zeros(N + 1, np.int32)
    ci = np.zeros(7 * N, np.int32)
    v = np.zeros(7 * N, np.float64)
    nnz = load().schwz_b200_laplacian3d(C.c_int32(n), _p(rp), _p(ci), _p(v))
    return rp, ci[:nnz].copy(), v[:nnz].copy()


def read_mtx(path):
    n = C.c_int32(0)
    nnz = C.c_int64(0)
    rp = C.POINTER(C.c_int32)()
    ci = C.POINTER(C.c_int32)()
    v = C.POINTER(C.c_double)()
    _chk(load().schwz_b200_read_mtx(path.encode(), C.byref(n), C.byref(nnz), C.byref(rp),
                                    C.byref(ci), C.byref(v)))
    try:
        rpa = np.ctypeslib.as_array(rp, shape=(n.value + 1,)).copy()
        cia = np.ctypeslib.as_array(ci, shape=(max(nnz.value, 1),))[:nnz.value].copy()
        va = np.ctypeslib.as_array(v, shape=(max(nnz.value, 1),))[:nnz.value].copy()
    finally:
        for q in (rp, ci, v):
            load().schwz_b200_host_free(q)
    return rpa, cia, va


def partition_regular2d(N, P):
    part = np.zeros(N, np.uint32)
    _chk(load().schwz_b200_partition_regular2d(C.c_int64(N), C.c_int32(P), _p(part)))
    return part


def partition_regular2d_rect(N, P, px=0, py=0):
    """px x py rectangular extension of regular2d (0, 0: most square factorisation)"""
    part = np.zeros(N, np.uint32)
    _chk(load().schwz_b200_partition_regular2d_rect(C.c_int64(N), C.c_int32(P), C.c_int32(px),
                                                    C.c_int32(py), _p(part)))
    return part


def partition_metis(rp, ci, P, objtype="null"):
    rp, ci = _i32(rp), _i32(ci)
    N = len(rp) - 1
    part = np.zeros(N, np.uint32)
    _chk(load().schwz_b200_partition_metis(C.c_int32(N), _p(rp), _p(ci), C.c_int32(P),
                                           objtype.encode(), _p(part)))
    return part


def host_cholesky(rp, ci, v, perm=None):
    rp, ci, v = _i32(rp), _i32(ci), _f64(v)
    n = len(rp) - 1
    pp = None if perm is None else _i32(perm)
    nnz = load().schwz_b200_host_cholesky(C.c_int32(n), _p(rp), _p(ci), _p(v), _p(pp),
                                          None, None, None)
    if nnz < 0:
        raise SchwzError(load().schwz_b200_last_error().decode())
    Lrp = np.zeros(n + 1, np.int32)
    Lci = np.zeros(nnz, np.int32)
    Lv = np.zeros(nnz, np.float64)
    load().schwz_b200_host_cholesky(C.c_int32(n), _p(rp), _p(ci), _p(v), _p(pp), _p(Lrp),
                                    _p(Lci), _p(Lv))
    return Lrp, Lci, Lv


def nd_ordering(rp, ci):
    rp, ci = _i32(rp), _i32(ci)
    n = len(rp) - 1
    perm = np.zeros(n, np.int32)
    _chk(load().schwz_b200_host_nd_ordering(C.c_int32(n), _p(rp), _p(ci), _p(perm)))
    return perm


class HostLu:
    """Sparse LU with threshold partial pivoting on the host, P A Q = L U (the UMFPACK branch of
    the reference, source/solve.cpp:145-171, 322-385).  col_perm = Q or None."""

    def __init__(self, rp, ci, v, col_perm=None, diag_pivot_tol=1e-3):
        rp, ci, v = _i32(rp), _i32(ci), _f64(v)
        self.n = len(rp) - 1
        self.col_perm = None if col_perm is None else _i32(col_perm)
        h = C.c_void_p()
        _chk(load().schwz_b200_host_lu_create(C.c_int32(self.n), _p(rp), _p(ci), _p(v),
                                              _p(self.col_perm), C.c_double(diag_pivot_tol),
                                              C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            load().schwz_b200_host_lu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def factors(self):
        """(L csr triple, U csr triple, row_perm)"""
        nl, nu = C.c_int64(0), C.c_int64(0)
        _chk(load().schwz_b200_host_lu_nnz(self.h, C.byref(nl), C.byref(nu)))
        L = (np.zeros(self.n + 1, np.int32), np.zeros(nl.value, np.int32), np.zeros(nl.value))
        U = (np.zeros(self.n + 1, np.int32), np.zeros(nu.value, np.int32), np.zeros(nu.value))
        p = np.zeros(self.n, np.int32)
        _chk(load().schwz_b200_host_lu_get(self.h, _p(L[0]), _p(L[1]), _p(L[2]), _p(U[0]),
                                           _p(U[1]), _p(U[2]), _p(p)))
        return L, U, p


class Setup:
    """Index sets of one RAS problem (SolverRAS::setup_local_matrices /
    setup_comm_buffers / setup_windows, source/restricted_schwarz.cpp:56-711).

    matrix: ("laplacian2d", n) | ("laplacian3d", n) | (rowptr, col, val)
    partition: None for the regular 1-D split, or a part-id vector (what the
    reference's regular2d / metis partitioners produce).
    """

    def __init__(self, matrix, P, part=None, overlap=2):
        self.P = P
        self.overlap = overlap
        h = C.c_void_p()
        kind = 0 if part is None else 1
        self._part = None if part is None else np.ascontiguousarray(part, np.uint32)
        if isinstance(matrix[0], str):
            mk = {"laplacian2d": 1, "laplacian3d": 2}[matrix[0]]
            n = int(matrix[1])
            self.N = n * n if mk == 1 else n ** 3
            _chk(load().schwz_b200_setup_create(C.c_int32(mk), C.c_int32(n), C.c_int32(self.N),
                                                None, None, None, C.c_int32(P), C.c_int32(kind),
                                                _p(self._part), C.c_int32(overlap), C.byref(h)))
        else:
            rp, ci, v = _i32(matrix[0]), _i32(matrix[1]), _f64(matrix[2])
            self.N = len(rp) - 1
            _chk(load().schwz_b200_setup_create(C.c_int32(0), C.c_int32(0), C.c_int32(self.N),
                                                _p(rp), _p(ci), _p(v), C.c_int32(P),
                                                C.c_int32(kind), _p(self._part),
                                                C.c_int32(overlap), C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                load().schwz_b200_setup_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def first_row(self):
        out = np.zeros(self.P + 1, np.int32)
        _chk(load().schwz_b200_setup_first_row(self.h, _p(out)))
        return out

    def permutation(self):
        perm = np.zeros(self.N, np.int32)
        iperm = np.zeros(self.N, np.int32)
        _chk(load().schwz_b200_setup_permutation(self.h, _p(perm), _p(iperm)))
        return perm, iperm

    def sizes(self, r):
        out = np.zeros(8, np.int64)
        _chk(load().schwz_b200_setup_sizes(self.h, C.c_int32(r), _p(out)))
        keys = ("local_size", "local_size_x", "overlap_size", "nnz_local", "nnz_interface",
                "n_halo", "num_neighbors_in", "num_neighbors_out")
        return dict(zip(keys, (int(x) for x in out)))

    def l2g(self, r):
        s = self.sizes(r)
        out = np.zeros(s["local_size_x"] + s["n_halo"], np.int32)
        _chk(load().schwz_b200_setup_l2g(self.h, C.c_int32(r), _p(out)))
        return out

    def local_matrix(self, r):
        s = self.sizes(r)
        rp = np.zeros(s["local_size_x"] + 1, np.int32)
        ci = np.zeros(s["nnz_local"], np.int32)
        v = np.zeros(s["nnz_local"], np.float64)
        _chk(load().schwz_b200_setup_local_matrix(self.h, C.c_int32(r), _p(rp), _p(ci), _p(v)))
        return rp, ci, v

    def interface_matrix(self, r):
        s = self.sizes(r)
        nrows = s["local_size_x"] if s["nnz_interface"] > 0 else 0
        rp = np.zeros(nrows + 1, np.int32)
        ci = np.zeros(s["nnz_interface"], np.int32)
        v = np.zeros(s["nnz_interface"], np.float64)
        _chk(load().schwz_b200_setup_interface_matrix(self.h, C.c_int32(r), _p(rp), _p(ci), _p(v)))
        return rp, ci, v

    def neighbors(self, r):
        s = self.sizes(r)
        nin = np.zeros(max(s["num_neighbors_in"], 1), np.int32)
        nout = np.zeros(max(s["num_neighbors_out"], 1), np.int32)
        _chk(load().schwz_b200_setup_neighbors(self.h, C.c_int32(r), _p(nin), _p(nout)))
        return nin[:s["num_neighbors_in"]], nout[:s["num_neighbors_out"]]

    def get_list(self, r, j):
        n = C.c_int32(0)
        _chk(load().schwz_b200_setup_get_count(self.h, C.c_int32(r), C.c_int32(j), C.byref(n)))
        out = np.zeros(n.value, np.int32)
        _chk(load().schwz_b200_setup_get_list(self.h, C.c_int32(r), C.c_int32(j), _p(out)))
        return out

    def put_list(self, r, j):
        n = C.c_int32(0)
        _chk(load().schwz_b200_setup_put_count(self.h, C.c_int32(r), C.c_int32(j), C.byref(n)))
        out = np.zeros(n.value, np.int32)
        _chk(load().schwz_b200_setup_put_list(self.h, C.c_int32(r), C.c_int32(j), _p(out)))
        return out

    def displacements(self, r):
        pd = np.zeros(self.P + 1, np.int32)
        gd = np.zeros(self.P + 1, np.int32)
        _chk(load().schwz_b200_setup_displacements(self.h, C.c_int32(r), _p(pd), _p(gd)))
        return pd, gd

    def release(self, r):
        _chk(load().schwz_b200_setup_release_rank(self.h, C.c_int32(r)))


# --------------------------------------------------------------------------
# device side
# --------------------------------------------------------------------------
class Context:
    def __init__(self, device=0):
        h = C.c_void_p()
        _chk(load().schwz_b200_ctx_create(C.c_int(device), C.byref(h)))
        self.h = h
        self.device = device
        self._bufs = []

    def close(self):
        if self.h:
            load().schwz_b200_ctx_destroy(self.h)
            self.h = None

    def stream(self):
        s = C.c_void_p()
        _chk(load().schwz_b200_ctx_stream(self.h, C.byref(s)))
        return s.value

    def sync(self):
        _chk(load().schwz_b200_ctx_sync(self.h))

    def malloc(self, nbytes):
        p = C.c_void_p()
        _chk(load().schwz_b200_malloc(self.h, C.c_size_t(nbytes), C.byref(p)))
        return p

    def free(self, p):
        _chk(load().schwz_b200_free(self.h, p))

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.malloc(max(arr.nbytes, 8))
        if arr.nbytes:
            _chk(load().schwz_b200_h2d(self.h, p, _p(arr), C.c_size_t(arr.nbytes)))
        return p

    def h2d(self, dev, arr):
        arr = np.ascontiguousarray(arr)
        _chk(load().schwz_b200_h2d(self.h, dev, _p(arr), C.c_size_t(arr.nbytes)))

    def to_host(self, dev, n, dtype=np.float64):
        out = np.zeros(n, dtype)
        if out.nbytes:
            _chk(load().schwz_b200_d2h(self.h, _p(out), dev, C.c_size_t(out.nbytes)))
        return out

    def zeros(self, n, dtype=np.float64):
        nbytes = max(n * np.dtype(dtype).itemsize, 8)
        p = self.malloc(nbytes)
        _chk(load().schwz_b200_memset(self.h, p, C.c_int(0), C.c_size_t(nbytes)))
        return p

    def timer_start(self):
        _chk(load().schwz_b200_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0)
        _chk(load().schwz_b200_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def ipc_export(self, dev):
        buf = (C.c_char * 64)()
        _chk(load().schwz_b200_ipc_export(self.h, dev, buf))
        return bytes(buf)

    def ipc_import(self, handle):
        p = C.c_void_p()
        buf = (C.c_char * 64).from_buffer_copy(handle)
        _chk(load().schwz_b200_ipc_import(self.h, buf, C.byref(p)))
        return p

    def dot(self, n, a, b):
        out = C.c_double(0)
        _chk(load().schwz_b200_dot(self.h, C.c_int64(n), a, b, C.byref(out)))
        return out.value

    def nrm2(self, n, a):
        out = C.c_double(0)
        _chk(load().schwz_b200_nrm2(self.h, C.c_int64(n), a, C.byref(out)))
        return out.value

    def axpy(self, n, alpha, x, y):
        _chk(load().schwz_b200_axpy(self.h, C.c_int64(n), C.c_double(alpha), x, y))

    def gather(self, n, idx, frm, into, op=1):
        _chk(load().schwz_b200_gather(self.h, C.c_int32(n), idx, frm, into, C.c_int(op)))

    def scatter(self, n, idx, frm, into, op=1):
        _chk(load().schwz_b200_scatter(self.h, C.c_int32(n), idx, frm, into, C.c_int(op)))

    def permute(self, n, perm, inverse, src, dst):
        _chk(load().schwz_b200_permute(self.h, C.c_int32(n), perm, C.c_int(int(inverse)), src, dst))


def enable_peers(ctxs):
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    _chk(load().schwz_b200_enable_peers(arr, C.c_int(len(ctxs))))


class Csr:
    """Device CSR (gko::matrix::Csr on the CUDA executor)."""

    def __init__(self, ctx, rp, ci, v, ncols=None):
        rp, ci, v = _i32(rp), _i32(ci), _f64(v)
        self.ctx = ctx
        self.nrows = len(rp) - 1
        self.ncols = self.nrows if ncols is None else ncols
        self.nnz = int(rp[-1])
        h = C.c_void_p()
        _chk(load().schwz_b200_csr_upload(ctx.h, C.c_int32(self.nrows), C.c_int32(self.ncols),
                                          _p(rp), _p(ci), _p(v), C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            load().schwz_b200_csr_destroy(self.h)
            self.h = None

    def spmv(self, x_dev, y_dev, alpha=1.0, beta=0.0):
        _chk(load().schwz_b200_spmv(self.ctx.h, self.h, C.c_double(alpha), x_dev,
                                    C.c_double(beta), y_dev))

    def spmv_bytes(self, beta_nonzero=False):
        return int(load().schwz_b200_spmv_bytes(self.h, C.c_int(int(beta_nonzero))))


class Precond:
    """Local preconditioner of CG / GMRES (source/solve.cpp:486-652): generated on the host
    from the host CSR, applied on the device.  kind in PRECOND.  ctx=None: generation only.
    The getters must be used before the handle is attached to a large Ras (which drops the
    host copy)."""

    def __init__(self, ctx, rp, ci, v, kind, max_block_size=16):
        self.n = len(rp) - 1
        self.kind = kind
        h = C.c_void_p()
        _chk(load().schwz_b200_precond_create(ctx.h if ctx is not None else None,
                                              C.c_int32(self.n), _p(_i32(rp)),
                                              _p(_i32(ci)), _p(_f64(v)),
                                              C.c_int32(PRECOND[kind]),
                                              C.c_int32(max_block_size), C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            load().schwz_b200_precond_destroy(self.h)
            self.h = None

    def apply(self, r_dev, z_dev, dot_dev=None):
        _chk(load().schwz_b200_precond_apply(self.h, r_dev, z_dev, dot_dev))

    def bytes_per_apply(self):
        return int(load().schwz_b200_precond_bytes_per_apply(self.h))

    def block_ptrs(self):
        out = np.zeros(load().schwz_b200_precond_block_ptrs(self.h, None), np.int32)
        load().schwz_b200_precond_block_ptrs(self.h, _p(out))
        return out

    def blocks(self):
        out = np.zeros(load().schwz_b200_precond_blocks(self.h, None), np.float64)
        load().schwz_b200_precond_blocks(self.h, _p(out))
        return out

    def csr(self, which):
        """0 L, 1 U of the ILU(0); 2 / 3 sparse approximate inverses of L / U (ISAI)."""
        nnz = load().schwz_b200_precond_csr(self.h, C.c_int32(which), None, None, None)
        rp = np.zeros(self.n + 1, np.int32)
        ci = np.zeros(nnz, np.int32)
        v = np.zeros(nnz, np.float64)
        load().schwz_b200_precond_csr(self.h, C.c_int32(which), _p(rp), _p(ci), _p(v))
        return rp, ci, v


class Cg:
    def __init__(self, ctx, A, precond=None):
        h = C.c_void_p()
        _chk(load().schwz_b200_cg_create(ctx.h, A.h, C.byref(h)))
        self.h = h
        self.precond = precond
        if precond is not None:
            _chk(load().schwz_b200_cg_set_precond(self.h, precond.h))

    def close(self):
        if self.h:
            load().schwz_b200_cg_destroy(self.h)
            self.h = None

    def solve(self, b_dev, x_dev, max_iters, tol):
        _chk(load().schwz_b200_cg_solve(self.h, b_dev, x_dev, C.c_int32(max_iters),
                                        C.c_double(tol)))

    def result(self):
        it = C.c_int32(0)
        rn = C.c_double(0)
        r0 = C.c_double(0)
        _chk(load().schwz_b200_cg_result(self.h, C.byref(it), C.byref(rn), C.byref(r0)))
        return it.value, rn.value, r0.value


class Gmres:
    def __init__(self, ctx, A, restart, precond=None):
        h = C.c_void_p()
        _chk(load().schwz_b200_gmres_create(ctx.h, A.h, C.c_int32(restart), C.byref(h)))
        self.h = h
        self.precond = precond
        if precond is not None:
            _chk(load().schwz_b200_gmres_set_precond(self.h, precond.h))

    def close(self):
        if self.h:
            load().schwz_b200_gmres_destroy(self.h)
            self.h = None

    def solve(self, b_dev, x_dev, max_iters, tol):
        _chk(load().schwz_b200_gmres_solve(self.h, b_dev, x_dev, C.c_int32(max_iters),
                                           C.c_double(tol)))

    def result(self):
        it = C.c_int32(0)
        rn = C.c_double(0)
        r0 = C.c_double(0)
        _chk(load().schwz_b200_gmres_result(self.h, C.byref(it), C.byref(rn), C.byref(r0)))
        return it.value, rn.value, r0.value


class Trs:
    def __init__(self, ctx, rp, ci, v, upper):
        rp, ci, v = _i32(rp), _i32(ci), _f64(v)
        h = C.c_void_p()
        _chk(load().schwz_b200_trs_analyze(ctx.h, C.c_int32(len(rp) - 1), _p(rp), _p(ci), _p(v),
                                           C.c_int(int(upper)), C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            load().schwz_b200_trs_destroy(self.h)
            self.h = None

    def solve(self, b_dev, x_dev):
        _chk(load().schwz_b200_trs_solve(self.h, b_dev, x_dev))

    def levels(self):
        n = C.c_int32(0)
        _chk(load().schwz_b200_trs_levels(self.h, C.byref(n)))
        return n.value

    def error(self):
        n = C.c_int32(0)
        _chk(load().schwz_b200_trs_error(self.h, C.byref(n)))
        return n.value


class Comm:
    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        _chk(load().schwz_b200_comm_unique_id(buf))
        return bytes(buf)

    def __init__(self, ctx, uid, nranks, rank):
        h = C.c_void_p()
        buf = (C.c_char * 128).from_buffer_copy(uid)
        _chk(load().schwz_b200_comm_create(ctx.h, buf, C.c_int32(nranks), C.c_int32(rank),
                                           C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            load().schwz_b200_comm_destroy(self.h)
            self.h = None


class Ras:
    """One subdomain of the RAS iteration on a device (the per-rank state of
    schwz::SolverRAS)."""

    def __init__(self, ctx, setup, rank, rhs=None, tolerance=1e-6, local_tol=1e-12,
                 local_max_iters=-1, local_solver="iterative-ginkgo", non_symmetric=False,
                 restart_iter=1, use_mixed_precision=False, local_precond="null",
                 precond_max_block_size=16):
        self.ctx = ctx
        self.rank = rank
        o = RasOptions(tolerance, local_tol, local_max_iters,
                       {"iterative-ginkgo": 2, "direct-ginkgo": 1}[local_solver],
                       int(non_symmetric), restart_iter, setup.overlap, int(use_mixed_precision),
                       PRECOND[local_precond], precond_max_block_size)
        h = C.c_void_p()
        rhs_arr = None if rhs is None else _f64(rhs)
        _chk(load().schwz_b200_ras_create(ctx.h, setup.h, C.c_int32(rank), _p(rhs_arr),
                                          C.byref(o), C.byref(h)))
        self.h = h
        info = np.zeros(6, np.int64)
        _chk(load().schwz_b200_ras_info(self.h, _p(info)))
        (self.n_in, self.n_out, self.local_size, self.local_size_x, self.n_halo,
         self.nnz_local) = (int(x) for x in info)

    def close(self):
        if self.h:
            load().schwz_b200_ras_destroy(self.h)
            self.h = None

    def neighbors(self):
        nin = np.zeros(max(self.n_in, 1), np.int32)
        nout = np.zeros(max(self.n_out, 1), np.int32)
        _chk(load().schwz_b200_ras_neighbors(self.h, _p(nin), _p(nout)))
        return nin[:self.n_in], nout[:self.n_out]

    def set_factors(self, Lrp, Lci, Lv, perm=None):
        Lrp, Lci, Lv = _i32(Lrp), _i32(Lci), _f64(Lv)
        pp = None if perm is None else _i32(perm)
        _chk(load().schwz_b200_ras_set_factors(self.h, _p(Lrp), _p(Lci), _p(Lv), _p(pp)))

    def set_local_max_iters(self, cap):
        _chk(load().schwz_b200_ras_set_local_max_iters(self.h, C.c_int32(cap)))

    def set_lu_factors(self, lu):
        """direct local solve Q U^-1 L^-1 P b with the factors of a HostLu"""
        _chk(load().schwz_b200_ras_set_lu_factors(self.h, lu.h, _p(lu.col_perm)))

    def mailbox(self):
        base = C.c_void_p()
        lay = MailboxLayout()
        _chk(load().schwz_b200_ras_mailbox(self.h, C.byref(base), C.byref(lay)))
        return base, lay

    def connect(self, j_out, peer_base, peer_layout, recv_off, flag_slot, same_process):
        _chk(load().schwz_b200_ras_connect(self.h, C.c_int32(j_out), peer_base,
                                           C.byref(peer_layout), C.c_int32(recv_off),
                                           C.c_int32(flag_slot), C.c_int32(int(same_process))))

    def connect_in(self, j_in, peer_base, peer_layout, send_off):
        _chk(load().schwz_b200_ras_connect_in(self.h, C.c_int32(j_in), peer_base,
                                              C.byref(peer_layout), C.c_int32(send_off)))

    def connect_conv(self, peer_rank, peer_base, peer_layout):
        _chk(load().schwz_b200_ras_connect_conv(self.h, C.c_int32(peer_rank), peer_base,
                                                C.byref(peer_layout)))

    def set_exchange_mode(self, mode):
        """EXCHANGE_MODES key or number (Settings::comm_settings put/get x one-by-one)."""
        _chk(load().schwz_b200_ras_set_exchange_mode(self.h, C.c_int32(_exchange_mode(mode))))

    def conv_tree(self, converged_all_local):
        _chk(load().schwz_b200_ras_conv_tree(self.h, C.c_int32(int(converged_all_local))))

    def conv_accumulate(self, converged_all_local):
        _chk(load().schwz_b200_ras_conv_accumulate(self.h, C.c_int32(int(converged_all_local))))

    def conv_set_local(self, converged_all_local):
        _chk(load().schwz_b200_ras_conv_set_local(self.h, C.c_int32(int(converged_all_local))))

    def conv_count(self):
        n = C.c_int32()
        _chk(load().schwz_b200_ras_conv_count(self.h, C.byref(n)))
        return int(n.value)

    # loop stages
    def set_onesided(self, on):
        _chk(load().schwz_b200_ras_set_onesided(self.h, C.c_int32(int(on))))

    def exchange_push(self, it):
        _chk(load().schwz_b200_ras_exchange_push(self.h, C.c_int32(it)))

    def exchange_unpack(self, it, wait_flags=False):
        _chk(load().schwz_b200_ras_exchange_unpack(self.h, C.c_int32(it), C.c_int32(int(wait_flags))))

    def wait_push_of(self, other):
        _chk(load().schwz_b200_ras_wait_push_of(self.h, other.h))

    def update_boundary(self):
        _chk(load().schwz_b200_ras_update_boundary(self.h))

    def local_residual(self):
        _chk(load().schwz_b200_ras_local_residual(self.h))

    def residual_norm(self):
        out = C.c_double(0)
        _chk(load().schwz_b200_ras_residual_norm(self.h, C.byref(out)))
        return out.value

    def local_solve(self):
        _chk(load().schwz_b200_ras_local_solve(self.h))

    def restrict(self):
        _chk(load().schwz_b200_ras_restrict(self.h))

    def last_local_iters(self):
        n = C.c_int32(0)
        _chk(load().schwz_b200_ras_last_local_iters(self.h, C.byref(n)))
        return n.value

    def sync(self):
        _chk(load().schwz_b200_ras_sync(self.h))

    def x(self):
        out = np.zeros(self.local_size_x + self.n_halo)
        _chk(load().schwz_b200_ras_get_x(self.h, _p(out)))
        return out

    def local_solution(self):
        out = np.zeros(self.local_size_x)
        _chk(load().schwz_b200_ras_get_local_solution(self.h, _p(out)))
        return out

    def set_x_own(self, arr):
        arr = _f64(arr)
        assert arr.shape[0] == self.local_size
        _chk(load().schwz_b200_ras_set_x_own(self.h, _p(arr)))

    def upload_rhs(self, host_ptr):
        _chk(load().schwz_b200_ras_upload_rhs(self.h, C.c_void_p(host_ptr)))

    def download_solution(self, host_ptr):
        _chk(load().schwz_b200_ras_download_solution(self.h, C.c_void_p(host_ptr)))

    def reset(self):
        _chk(load().schwz_b200_ras_reset(self.h))

    def kernel_time_ms(self, kind, reps=20):
        ms = C.c_float(0)
        _chk(load().schwz_b200_ras_kernel_time(self.h, C.c_int32(kind), C.c_int32(reps),
                                               C.byref(ms)))
        return ms.value

    def kernel_bytes(self, kind):
        return int(load().schwz_b200_ras_kernel_bytes(self.h, C.c_int32(kind)))

    def true_residual_sq(self):
        out = C.c_double(0)
        _chk(load().schwz_b200_ras_true_residual_sq(self.h, C.byref(out)))
        return out.value


EXCHANGE_MODES = {"put": 0, "get": 1, "put-one-by-one": 2, "get-one-by-one": 3}


def _exchange_mode(mode):
    return EXCHANGE_MODES[mode] if isinstance(mode, str) else int(mode)


def exchange_mode(remote_comm_type="put", enable_one_by_one=False):
    """bench_ras flags --remote_comm_type / --enable_one_by_one -> exchange mode."""
    return EXCHANGE_MODES[remote_comm_type + ("-one-by-one" if enable_one_by_one else "")]


def mailbox_layout(in_total, n_in, P, out_total=0, x_len=0):
    """Layout of a subdomain's peer-visible mailbox from its index-set sizes
    (host only; the same function the device side uses)."""
    lay = MailboxLayout()
    _chk(load().schwz_b200_mailbox_layout(C.c_int64(in_total), C.c_int32(n_in), C.c_int32(P),
                                          C.c_int64(out_total), C.c_int64(x_len), C.byref(lay)))
    return lay


def remote_connection_plan(setup, my_ranks, nbr_in_of):
    """Which peer mailboxes the subdomains of this process must be connected to.

    Replaces the MPI handshake + MPI_Alltoall of displacement tables
    (source/restricted_schwarz.cpp:400-472, 624-658) for the one-process-per-GPU
    launch: every process holds the index sets, so the plan is computed locally
    and only the mailbox handles travel.  nbr_in_of[q] = neighbors_in list of
    subdomain q (from its owner).  Returns tuples
    (rank, j_out, q, recv_offset_elems, flag_slot) for every out-neighbour q that
    lives in another process."""
    plan = []
    mine = set(my_ranks)
    for r in my_ranks:
        _, nout = setup.neighbors(r)
        pd, _ = setup.displacements(r)
        for j, q in enumerate(nout.tolist()):
            if q in mine:
                continue
            plan.append((r, j, q, int(pd[q]), list(nbr_in_of[q]).index(r)))
    return plan


def connect_local(subs, setup):
    arr = (C.c_void_p * len(subs))(*[s.h for s in subs])
    _chk(load().schwz_b200_ras_connect_local(arr, C.c_int32(len(subs)), setup.h))


def refresh_halo(subs, num_subdomains):
    """one synchronous halo exchange outside the loop (collective): x's overlap / halo entries
    become the neighbours' current values"""
    arr = (C.c_void_p * len(subs))(*[s.h for s in subs])
    _chk(load().schwz_b200_ras_refresh_halo(arr, C.c_int32(len(subs)), C.c_int32(num_subdomains)))


def ras_run(subs, num_subdomains, max_iters, tolerance=1e-6, enable_onesided=False,
            enable_global_check=True, conv_decentralized=False, iter_offset=False, comm=None,
            history=False, exchange="put", enable_accumulate=False):
    """The outer loop of SchwarzBase::run (source/schwarz_base.cpp:387-452) over
    the subdomains of this process.  One-sided runs: conv_decentralized selects the flag
    flooding protocol (else the centralised tree), `exchange` one of EXCHANGE_MODES."""
    arr = (C.c_void_p * len(subs))(*[s.h for s in subs])
    o = LoopOptions(num_subdomains, max_iters, tolerance, int(enable_onesided),
                    int(enable_global_check),
                    2 if (conv_decentralized and enable_accumulate) else int(conv_decentralized),
                    int(iter_offset),
                    _exchange_mode(exchange), comm.h if comm is not None else None)
    res = LoopResult()
    hist = np.zeros((max_iters, len(subs))) if history else None
    _chk(load().schwz_b200_ras_run(arr, C.c_int32(len(subs)), C.byref(o), C.byref(res), _p(hist)))
    out = dict(iters=res.iters, converged=bool(res.converged), global_resnorm=res.global_resnorm,
               global_resnorm0=res.global_resnorm0, elapsed_s=res.elapsed_s,
               host_stream_syncs=res.host_stream_syncs, host_event_waits=res.host_event_waits)
    if history:
        out["history"] = hist[:max(res.iters + (1 if res.converged else 0), 0)]
    return out
