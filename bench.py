#!/usr/bin/env python
"""bench.py — RAS outer iterations per second on BASELINE.json configs[1]:
2-D 5-pt Laplacian 8192x8192, fp64/int32, 8 subdomains (1-D strips,
--partition=regular), CG local solve with a fixed budget of --local-iters
iterations, synchronous halo exchange, global convergence check.

A "step" is one outer RAS iteration over all 8 subdomains: halo exchange ->
boundary update -> residual check (+ allgather) -> local CG solve ->
restriction (source/schwarz_base.cpp:387-452 of the reference).  The 8
subdomains are spread over the N GPUs (8/N per GPU), so the problem — and the
sequence of iterates — is the same at every N ("strong" scaling).

    python bench.py --gpus N --steps K --warmup W
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU arm (oracle port on the host cores)

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "schwarz-lib_b200"))

METRIC = "ras_outer_iters_per_s"
UNIT = "outer iters/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--size", dest="n", type=int, default=8192,
                    help="1-D Laplacian size (grid is n x n, or n^3 with --dim 3); use --size under torchrun")
    ap.add_argument("--subdomains", type=int, default=8)
    ap.add_argument("--local-iters", type=int, default=50, help="--local_max_iters of bench_ras")
    ap.add_argument("--dim", type=int, default=2, choices=[2, 3],
                    help="2: 5-pt Laplacian n x n (cfg2); 3: 7-pt Laplacian n^3 (cfg4, use --n 512)")
    ap.add_argument("--matrix", default="laplacian", choices=["laplacian", "ani4", "ani3"],
                    help="ani4 / ani3: tests/golden/ani{4,3}_crop.npz (the reference's "
                         "matrices/ani*_crop.mtx), METIS partition, GMRES(30) local solve to "
                         "local_tol (cfg3)")
    ap.add_argument("--partition", default="regular", choices=["regular", "regular2d"],
                    help="regular: 1-D strips (reference-exact); regular2d: px x py blocks - the "
                         "reference's rule for square counts, its documented rectangular "
                         "extension otherwise (8 -> 2 x 4)")
    ap.add_argument("--local-tol", type=float, default=1e-12, help="--local_tol of bench_ras")
    ap.add_argument("--onesided", action="store_true",
                    help="one-sided Put exchange + decentralised convergence flags (cfg4)")
    ap.add_argument("--to-tolerance", type=int, default=0, metavar="MAX_ITERS",
                    help="additionally run the outer loop of THIS workload from a zero start "
                         "until the global criterion (set_tol 1e-6) is met or MAX_ITERS, and "
                         "report it as time_to_solution_full (rhs upload and solution download "
                         "included)")
    ap.add_argument("--tts-size", type=int, default=1024,
                    help="grid size of the time_to_solution leg every run carries (same "
                         "workload shape at n x n, run from a zero start to set_tol 1e-6); "
                         "0 turns it off")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-breakdown", action="store_true",
                    help="diagnostic: synchronise between upload / run / download of the e2e leg "
                         "and report the three wall times (changes the e2e number slightly)")
    return ap.parse_args()


def make_setup(args, S):
    """host index sets of the workload (outside every timed region)"""
    if args.matrix in ("ani4", "ani3"):
        z = np.load(os.path.join(ROOT, "tests", "golden", "%s_crop.npz" % args.matrix))
        mat = (z["rowptr"], z["col"], z["val"])
        part = S.partition_metis(mat[0], mat[1], args.subdomains)
        return S.Setup(mat, args.subdomains, part=part)
    kind = "laplacian3d" if args.dim == 3 else "laplacian2d"
    if args.partition == "regular2d":
        assert args.dim == 2, "regular2d partitions the 2-D grid"
        part = S.partition_regular2d_rect(args.n * args.n, args.subdomains)
        return S.Setup((kind, args.n), args.subdomains, part=part)
    return S.Setup((kind, args.n), args.subdomains)


def ras_kwargs(args):
    if args.matrix in ("ani4", "ani3"):
        return dict(tolerance=1e-6, local_tol=1e-12, local_max_iters=-1, non_symmetric=True,
                    restart_iter=30)
    return dict(tolerance=1e-6, local_tol=args.local_tol, local_max_iters=args.local_iters)


def rect_factors(P):
    px = max(d for d in range(1, int(P ** 0.5) + 1) if P % d == 0)
    return px, P // px


def workload(args):
    if args.matrix in ("ani4", "ani3"):
        dims = {"ani4": "N=3081, nnz=20971", "ani3": "N=741, nnz=4951"}[args.matrix]
        return {
            "workload": "cfg3: matrices/%s_crop.mtx (%s), METIS partition into %d "
                        "subdomains, overlap 2, GMRES(30) local solve to local_tol=1e-12, "
                        "synchronous halo exchange, enable_global_check"
                        % (args.matrix, dims, args.subdomains),
            "subdomains": args.subdomains, "overlap": 2, "partition": "metis", "restart_iter": 30,
            "l2_policy": "latency-bound workload (whole problem is < 1 MB); no L2 flush",
        }
    if args.dim == 3:
        return {
            "workload": "cfg4: 3D 7-pt Laplacian %d^3 fp64/int32, %d subdomains (regular 1-D slabs, "
                        "overlap 2), CG local solve local_max_iters=%d local_tol=%g, %s"
                        % (args.n, args.subdomains, args.local_iters, args.local_tol,
                           "one-sided Put exchange, decentralised convergence flags" if args.onesided
                           else "synchronous halo exchange, enable_global_check"),
            "n": args.n, "dim": 3, "subdomains": args.subdomains,
            "local_max_iters": args.local_iters, "overlap": 2, "partition": "regular",
            "l2_policy": "inputs larger than L2 (each local CSR is ~1.7 GB vs 126 MB L2)",
        }
    if args.partition == "regular2d":
        shape = "regular2d %d x %d blocks" % rect_factors(args.subdomains)
    else:
        shape = "regular 1-D strips"
    return {
        "workload": "cfg2: 2D 5-pt Laplacian %dx%d fp64/int32, %d subdomains (%s, "
                    "overlap 2), CG local solve local_max_iters=%d local_tol=%g, synchronous "
                    "halo exchange, enable_global_check" % (args.n, args.n, args.subdomains, shape,
                                                            args.local_iters, args.local_tol),
        "n": args.n, "subdomains": args.subdomains, "local_max_iters": args.local_iters,
        "overlap": 2, "partition": args.partition,
        "l2_policy": "inputs larger than L2 (each local CSR is ~0.5 GB vs 126 MB L2)",
    }


# ----------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------
# CPU arm.  The reference cannot step cfg2 itself (every rank replicates the 4.3 GB global
# matrix and sets up through std::map walks), so the arm is the oracle port
# (oracle/schwz_oracle.cpp, pinned bit for bit to oracle/_ref = the reference's own sources on
# stand-ins for MPI / Ginkgo), kind = "port", stepping its REAL outer loop on the bench
# workload: exchange -> boundary update -> residual + global check -> local solve ->
# restriction for all subdomains (source/schwarz_base.cpp:387-452).  Nothing of the product is
# imported here: matrix, partition and index sets all come from the oracle.  Up to 2048^2 the
# reference's own SolverRAS::run (oracle/_ref) is timed instead, kind = "reference".
# ----------------------------------------------------------------------------
def metis_part_fixture(matrix, P):
    """the partition vector the reference's own METIS call produced (tests/golden/ref_cfg3_*,
    generated from oracle/_ref by tests/golden/make_ref_golden.py)"""
    f = os.path.join(ROOT, "tests", "golden", "ref_cfg3_%s_metis_P%d_gmres.npz"
                     % (matrix.replace("_crop", ""), P))
    if not os.path.exists(f):
        raise RuntimeError("no METIS partition fixture for %s at %d subdomains (%s)" % (matrix, P, f))
    return np.load(f)["partition_indices"]


def host_cores():
    """cores this process may run on - NOT omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1 to its workers"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class CpuArm:
    """the oracle's RAS loop on the bench workload, all host cores"""

    def __init__(self, args):
        self.args = args
        self.cores = host_cores()
        os.environ["OMP_NUM_THREADS"] = str(self.cores)   # before libgomp initialises
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O
        self.O = O
        P = args.subdomains
        t0 = time.perf_counter()
        if args.matrix in ("ani4", "ani3"):
            z = np.load(os.path.join(ROOT, "tests", "golden", "%s_crop.npz" % args.matrix))
            mat = (z["rowptr"], z["col"], z["val"])
            part = metis_part_fixture(args.matrix, P)
            self.ob = O.Problem(*mat, P, part=part)
        elif args.partition == "regular2d":
            mat = O.laplacian2d(args.n)
            part = O.partition_regular2d_rect(args.n * args.n, P)
            self.ob = O.Problem(*mat, P, part=part)
        else:
            mat = O.laplacian3d(args.n) if args.dim == 3 else O.laplacian2d(args.n)
            self.ob = O.Problem(*mat, P)
        del mat
        kw = dict(ras_kwargs(args))
        nonsym = kw.pop("non_symmetric", False)
        self.ob.configure(max_iters=100000, enable_global_check=not args.onesided,
                          enable_onesided=args.onesided, remote_comm_type="put",
                          global_convergence_type="decentralized" if args.onesided
                          else "centralized-tree", non_symmetric=nonsym, **kw)
        self.setup_s = time.perf_counter() - t0
        self.side = min(P, self.cores)
        self.modes = {"team": (1, self.cores),
                      "ranks": (self.side, max(1, self.cores // self.side))}
        self.tried = {}
        self.mode = None
        self.steps_done = 0
        self.cold = True

    def _step(self, mode):
        rt, th = self.modes[mode]
        self.O.set_rank_threads(1 if self.args.onesided else rt)
        self.O.set_threads(th)
        t0 = time.perf_counter()
        self.ob.step()
        self.steps_done += 1
        return time.perf_counter() - t0

    def step(self):
        """one outer iteration; the first call is cold (page faults, first touch), the next two
        try the two ways of occupying the cores (subdomains side by side as the reference's
        MPI ranks x OpenMP threads run / one subdomain after the other with every core on it),
        the faster is kept from then on"""
        if self.cold:
            self.cold = False
            return self._step("ranks")
        for m in ("ranks", "team"):
            if m not in self.tried and self.mode is None:
                self.tried[m] = self._step(m)
                if len(self.tried) == 2:
                    self.mode = min(self.tried, key=self.tried.get)
                return self.tried[m]
        return self._step(self.mode)

    def describe(self, times):
        rt, th = self.modes[self.mode or "ranks"]
        st = self.ob.status(0)
        return ("%d outer iteration(s) of the oracle's own RAS loop on the bench workload (all %d "
                "subdomains: exchange, boundary update, residual + global check, local solve, "
                "restriction), %.3g s each, %s; first tries: %s; global residual ratio after "
                "%d iterations %.6g"
                % (len(times), self.args.subdomains, float(np.mean(times)),
                   "%d subdomains side by side x %d thread(s)" % (rt, th) if rt > 1
                   else "one subdomain after the other x %d threads" % th,
                   ", ".join("%s %.3g s" % kv for kv in self.tried.items()),
                   self.steps_done, st["gres"] / st["gres0"] if st["gres0"] > 0 else float("nan")))


def ref_own_loop(args, cores, steps):
    """oracle/_ref = the reference's own SolverRAS::run (one thread per MPI rank, sequential
    stand-in Ginkgo per rank): seconds per outer iteration = wall time of run() (the window
    of source/schwarz_base.cpp:384-455) / iterations, max over the ranks.  Only for sizes the
    reference can set up in seconds."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libschwz_ref.so")):
        return None
    import ref as R
    K = max(args.subdomains + 2, steps)   # max_iters >= ranks (a window-sizing quirk upstream)
    rr = R.Run(args.subdomains, laplacian_n=args.n, max_iters=K,
               local_max_iters=args.local_iters, enable_global_check=True, tolerance=1e-30,
               local_tol=args.local_tol)
    return max(rr.run_seconds(r) for r in range(args.subdomains)) / K


def cpu_baseline(args, steps, warmup=2):
    """bounded sample for the b200 arm's cpu_baseline object"""
    arm = CpuArm(args)
    for _ in range(max(warmup, 3)):
        arm.step()
    times = [arm.step() for _ in range(steps)]
    t = float(np.mean(times))
    return {"value": 1.0 / t, "unit": UNIT, "cores": arm.cores, "kind": "port",
            "sample": arm.describe(times), "setup_s": arm.setup_s}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    small = (args.matrix == "laplacian" and args.dim == 2 and args.n <= 2048
             and args.partition == "regular" and not args.onesided)
    base = None
    if small:
        cores = host_cores()
        t = ref_own_loop(args, cores, args.steps + args.warmup)
        if t is not None:
            base = {"value": 1.0 / t, "unit": UNIT, "cores": min(cores, args.subdomains),
                    "kind": "reference",
                    "sample": "oracle/_ref: the reference's own SolverRAS::run, %d rank threads, "
                              "%.3g s per outer iteration (wall time of run() over %d iterations)"
                              % (args.subdomains, t, max(args.subdomains + 2,
                                                         args.steps + args.warmup))}
            steps = args.steps
    if base is None:
        arm = CpuArm(args)
        for _ in range(args.warmup):
            arm.step()
        if arm.mode is None:
            arm.mode = min(arm.tried, key=arm.tried.get) if arm.tried else "ranks"
        times = [arm.step() for _ in range(args.steps)]   # every requested step is really run
        steps = len(times)
        t = float(np.mean(times))
        base = {"value": 1.0 / t, "unit": UNIT, "cores": arm.cores, "kind": "port",
                "sample": arm.describe(times), "setup_s": arm.setup_s}
    value = base["value"]
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
           "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
           "ms_per_step": 1e3 / value, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload(args), "cpu_baseline": base,
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------
class StdoutToStderr:
    """Everything written to file descriptor 1 while this is active - by Python, torch, NCCL (its
    version banner) or any other library - goes to stderr, so that stdout carries exactly the one
    JSON line of the contract, printed after leaving the context."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    with StdoutToStderr():
        line = run_b200(args)
    if line is not None:
        print(line, flush=True)


def run_b200(args):
    # NCCL's banner goes to stderr (and so does anything else that prints, see StdoutToStderr)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as dist
    import schwz_b200 as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    P = args.subdomains
    assert P % world == 0, "subdomains must divide evenly over the GPUs"
    nl = P // world
    my = list(range(rank * nl, (rank + 1) * nl))
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)

    # ---- setup (host index sets, upload) — outside every timed region --------
    setup = make_setup(args, S)
    ctxs = [S.Context(dev) for _ in my]
    subs = []
    for c, r in zip(ctxs, my):
        subs.append(S.Ras(c, setup, r, **ras_kwargs(args)))
        setup.release(r)
    S.connect_local(subs, setup)
    def connect_remote(subs_, setup_, imported_, ctxs_):
        """exchange mailbox IPC handles + layouts + in-neighbour lists, open the peers' mailboxes,
        create the NCCL communicator of the residual-norm allgather"""
        mine = {}
        for s in subs_:
            base, lay = s.mailbox()
            mine[s.rank] = (s.ctx.ipc_export(base), lay.as_tuple(), s.neighbors()[0].tolist())
        allinfo = [None] * world
        dist.all_gather_object(allinfo, mine)
        info = {}
        for d in allinfo:
            info.update(d)
        by_rank = {s.rank: s for s in subs_}
        plan = S.remote_connection_plan(setup_, my, {q: v[2] for q, v in info.items()})
        for r, j, q, recv_off, slot in plan:
            handle, lay, _ = info[q]
            if q not in imported_:
                imported_[q] = by_rank[r].ctx.ipc_import(handle)
            by_rank[r].connect(j, imported_[q], S.MailboxLayout.from_tuple(lay), recv_off, slot,
                               same_process=False)
        uid = [S.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        return S.Comm(ctxs_[0], uid[0], world, rank)

    comm = None
    imported = {}
    if world > 1:
        comm = connect_remote(subs, setup, imported, ctxs)

    def barrier():
        for s in subs:
            s.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def run_steps(k):
        if args.onesided:
            return S.ras_run(subs, P, k, tolerance=1e-6, enable_onesided=True,
                             conv_decentralized=True, comm=comm)
        return S.ras_run(subs, P, k, tolerance=1e-6, enable_global_check=True, comm=comm)

    # ---- warm-up, then exactly K timed steps ---------------------------------
    barrier()
    if args.warmup > 0:
        run_steps(args.warmup)
    barrier()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    launches0 = S.launch_count()
    for c in ctxs:
        c.timer_start()                 # CUDA event on every subdomain's stream
    t0 = time.perf_counter()
    res = run_steps(args.steps)
    ms = max(c.timer_stop() for c in ctxs)   # device time, max over the local streams
    wall = time.perf_counter() - t0
    barrier()
    launches = S.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms, float(launches)], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms = float(tmax[0])
        launches = int(t[1])
    # exactly K outer iterations must have run: a run that met the criterion early (or found
    # every subdomain finished on entry) would overstate the rate
    if res["converged"] or res["iters"] != args.steps:
        raise RuntimeError("timed run executed %d of %d outer iterations (converged=%s): reset "
                           "the state or lower --steps" % (res["iters"], args.steps,
                                                          res["converged"]))
    value = res["iters"] / (ms * 1e-3)

    # ---- roofline of the dominant kernel (CG SpMV with fused dot) ------------
    s0 = subs[min(1, len(subs) - 1)]
    roof = {}
    kern = {}
    for kind, name in ((0, "csr_spmv_tma_kernel<EPI_DOT>"), (1, "cg_r_update_kernel"),
                       (2, "cg_xp_update_kernel"), (3, "csr_spmv_tma_kernel<EPI_NRM2>")):
        if args.matrix != "laplacian" and kind in (1, 2):
            continue                     # GMRES local solve: no CG vector kernels
        kms = s0.kernel_time_ms(kind, 20)
        kb = s0.kernel_bytes(kind)
        kern[name] = {"ms": kms, "bytes": kb, "GB/s": kb / (kms * 1e-3) / 1e9}
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    k0 = kern["csr_spmv_tma_kernel<EPI_DOT>"]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    per_step_spmv_ms = k0["ms"] * args.local_iters * nl if args.matrix == "laplacian" else None
    roof = {"bound": "hbm", "kernel": "csr_spmv_tma_kernel<EPI_DOT> (CG: q = A p, p.q fused)",
            "achieved": k0["GB/s"], "peak": peak, "unit": "GB/s", "frac": k0["GB/s"] / peak,
            "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": k0["bytes"],
            "traffic_note": "DRAM bytes per launch from ncu (profiles/spmv_traffic.json). Below the "
                            "algorithmic bytes since round 2: narrow row tiles stream their columns "
                            "as 16-bit offsets (10 instead of 12 B per non-zero); `achieved` keeps "
                            "the algorithmic definition of SURVEY 8(d), so frac can read ~1.0",
            "launch_ms": k0["ms"],
            "share_of_step": per_step_spmv_ms / (ms / args.steps) if per_step_spmv_ms else None,
            "other_kernels": {k: v for k, v in kern.items() if k != "csr_spmv_tma_kernel<EPI_DOT>"}}

    # ---- halo exchange: push (pack + peer stores) + unpack of the subdomain of this rank that
    # has a neighbour on another GPU when there is one (NVLink), else a same-GPU neighbour ----
    # (every subdomain of every rank runs the same number of exchanges so that the epoch
    # counters of neighbours stay in step for the e2e run below)
    sb = subs[-1] if (world > 1 and rank < world - 1) else subs[0]
    barrier()
    halo_all = {s.rank: s.kernel_time_ms(4, 20) for s in subs}
    barrier()
    push_all = {s.rank: s.kernel_time_ms(5, 20) for s in subs}
    barrier()
    unpack_all = {s.rank: s.kernel_time_ms(6, 20) for s in subs}
    barrier()
    halo_ms = halo_all[sb.rank]
    halo_bytes = sb.kernel_bytes(4)
    _, nout = setup.neighbors(sb.rank)
    out_bytes = [8 * len(setup.put_list(sb.rank, j)) for j in range(len(nout))]
    remote_bytes = sum(b for b, q in zip(out_bytes, nout) if int(q) // nl != rank)
    push_ms = push_all[sb.rank]
    halo = {"subdomain": sb.rank, "push_unpack_ms": halo_ms, "algorithmic_bytes": halo_bytes,
            "GB/s": halo_bytes / (halo_ms * 1e-3) / 1e9,
            "payload_bytes_out": sum(out_bytes),
            "push_ms": push_ms, "unpack_ms": unpack_all[sb.rank],
            # payload that leaves this GPU through NVLink peer stores in one push launch; the
            # launch also carries the same-GPU neighbours' blocks, so this is a lower bound of
            # the link rate (NVLink 5: 900 GB/s per direction)
            "nvlink": {"payload_bytes": remote_bytes,
                       "GB/s": remote_bytes / (push_ms * 1e-3) / 1e9 if remote_bytes else None,
                       "peak_GB/s": 900.0},
            "link": "nvlink peer stores to the next GPU + local" if world > 1 and rank < world - 1
                    else "same GPU"}

    # ---- e2e: the plugin call with HOST buffers ------------------------------
    # rhs (pinned host) -> device, zero initial state, K outer iterations (each
    # reads the residual norms back, as the reference does), solution -> host.
    e2e = None
    if not args.no_e2e:
        N = setup.N
        rhs = torch.ones(N, dtype=torch.float64).pin_memory()
        sol = torch.zeros(N, dtype=torch.float64).pin_memory()
        barrier()
        t0 = time.perf_counter()
        for s in subs:
            s.upload_rhs(rhs.data_ptr())
            s.reset()
        if world > 1:
            dist.barrier()          # reset everywhere before anybody's first push
        if args.e2e_breakdown:
            for s in subs:
                s.sync()
            t_up = time.perf_counter() - t0
        r2 = run_steps(args.steps)
        if args.e2e_breakdown:
            for s in subs:
                s.sync()
            t_run = time.perf_counter() - t0
        for s in subs:
            s.download_solution(sol.data_ptr())
        if args.e2e_breakdown:
            for s in subs:
                s.sync()
            t_down = time.perf_counter() - t0
        barrier()
        te = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([te], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t[0])
        h2d = sum(s.local_size_x for s in subs) * 8
        d2h = sum(s.local_size for s in subs) * 8
        if world > 1:
            t = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            h2d, d2h = int(t[0]), int(t[1])
        assert r2["iters"] == args.steps and not r2["converged"]
        e2e = {"value": args.steps / te, "unit": UNIT,
               "h2d_bytes_per_step": h2d / args.steps,
               "d2h_bytes_per_step": d2h / args.steps + 8 * P,
               "outer_iterations": r2["iters"],
               "note": "schwz_b200_ras_upload_rhs + ras_reset + ras_run(K) + "
                       "ras_download_solution with pinned host buffers; rhs/solution copies "
                       "amortised over the K steps; the loop state (residual norms, decision) "
                       "is copied to pinned host memory behind every step, the host waits for "
                       "it once per chunk of steps"}
        if args.e2e_breakdown:
            e2e["breakdown_ms_rank0"] = {"upload+reset": 1e3 * t_up, "run": 1e3 * (t_run - t_up),
                                         "download": 1e3 * (t_down - t_run),
                                         "final_barrier": 1e3 * (te - t_down)}

    # ---- time-to-solution: zero start -> global criterion, host buffers in and out ----------
    def to_tolerance(subs_, setup_, comm_, P_, max_iters):
        N = setup_.N
        rhs = torch.ones(N, dtype=torch.float64).pin_memory()
        sol = torch.zeros(N, dtype=torch.float64).pin_memory()
        for s in subs_:
            s.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for s in subs_:
            s.upload_rhs(rhs.data_ptr())
            s.reset()
        if world > 1:
            dist.barrier()
        if args.onesided:
            r3 = S.ras_run(subs_, P_, max_iters, tolerance=1e-6, enable_onesided=True,
                           conv_decentralized=True, comm=comm_)
        else:
            r3 = S.ras_run(subs_, P_, max_iters, tolerance=1e-6, enable_global_check=True,
                           comm=comm_)
        for s in subs_:
            s.download_solution(sol.data_ptr())
        for s in subs_:
            s.sync()
        if world > 1:
            dist.barrier()
        tt = time.perf_counter() - t0
        # the true residual of what came back: ||b - A x|| / ||b|| (distributed, own rows)
        S.refresh_halo(subs_, P_)
        rr = sum(s.true_residual_sq() for s in subs_)
        if world > 1:
            t = torch.tensor([tt, rr], dtype=torch.float64, device="cuda")
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            tt, rr = float(tm[0]), float(t[1])
        return {"time_to_solution_s": tt, "outer_iterations": r3["iters"],
                "converged": r3["converged"], "tolerance": 1e-6,
                "global_resnorm_ratio": (r3["global_resnorm"] / r3["global_resnorm0"]
                                         if r3["global_resnorm0"] > 0 else None),
                "true_relative_residual": float(np.sqrt(rr) / np.sqrt(N)),
                "outer_iters_per_s": r3["iters"] / tt}

    tts_full = None
    if args.to_tolerance > 0:
        tts_full = to_tolerance(subs, setup, comm, P, args.to_tolerance)

    # The headline workload needs ~7e4 outer iterations of ~10 ms to reach set_tol (one-level
    # Schwarz: the count grows with strip width / overlap), far beyond a bench run; the leg every
    # run carries is therefore the SAME workload shape at --tts-size (default 1024^2), run to
    # tolerance from a zero start with host buffers in and out.
    tts = None
    if args.tts_size > 0 and args.matrix == "laplacian" and args.dim == 2 and not args.onesided:
        import copy
        a2 = copy.copy(args)
        a2.n = args.tts_size
        setup2 = make_setup(a2, S)
        ctxs2 = [S.Context(dev) for _ in my]
        subs2 = []
        for c, r in zip(ctxs2, my):
            subs2.append(S.Ras(c, setup2, r, **ras_kwargs(a2)))
            setup2.release(r)
        S.connect_local(subs2, setup2)
        comm2, imported2 = None, {}
        if world > 1:
            comm2 = connect_remote(subs2, setup2, imported2, ctxs2)
        tts = to_tolerance(subs2, setup2, comm2, P, 200000)
        tts["workload"] = workload(a2)["workload"]
        tts["n"] = a2.n
        for s in subs2:
            s.close()
        if comm2 is not None:
            comm2.close()
        for c in ctxs2:
            c.close()

    # ---- CPU baseline beside it (rank 0, N == 1 only) ------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(args, steps=3)
            if tts is not None:
                import copy
                a2 = copy.copy(args)
                a2.n = args.tts_size
                c2 = cpu_baseline(a2, steps=10, warmup=3)
                cpu["tts_s_per_outer"] = 1.0 / c2["value"]
                cpu["tts_steps"] = 10
        except Exception as e:  # the oracle is only a reported baseline
            cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "port",
                   "sample": "failed: %r" % (e,)}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": workload(args),
               "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof,
               "cpu_baseline": cpu, "halo": halo, "wall_ms_per_step": 1e3 * wall / args.steps,
               # host-blocking CUDA calls of the timed ras_run (rank 0): every stream
               # synchronisation sits before the first / after the last enqueued iteration,
               # the event waits look at a loop-state snapshot two chunks of iterations old
               "host_waits": {"stream_syncs": res["host_stream_syncs"],
                              "event_waits": res["host_event_waits"], "steps": args.steps},
               "global_resnorm": res["global_resnorm"], "impl": "b200"}
        if tts is not None:
            if cpu and cpu.get("tts_s_per_outer"):
                # the CPU arm's time for the same run: its measured seconds per outer
                # iteration at this size x the iterations the run needs (stated extrapolation)
                tts["cpu_baseline_s"] = cpu["tts_s_per_outer"] * tts["outer_iterations"]
                tts["cpu_baseline_how"] = ("oracle port, %d cores: %.4g s per outer iteration "
                                           "measured over %d iterations at this size x %d "
                                           "iterations" % (cpu["cores"], cpu["tts_s_per_outer"],
                                                           cpu["tts_steps"], tts["outer_iterations"]))
            out["time_to_solution"] = tts
        if tts_full is not None:
            out["time_to_solution_full"] = tts_full
        line = json.dumps(out)
    else:
        line = None

    for s in subs:
        s.close()
    if comm is not None:
        comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


if __name__ == "__main__":
    main()
