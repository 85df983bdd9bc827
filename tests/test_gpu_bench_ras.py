"""The drop-in driver: schwarz-lib_b200/bin/bench_ras with the reference's flag
surface (benchmarking/bench_base.hpp:50-144) and output lines
(source/schwarz_base.cpp:247-250, 474-478, 493-497), checked against the oracle
and the Appendix E known answers."""
import csv
import json
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "schwarz-lib_b200", "bin", "bench_ras")
E = json.load(open(os.path.join(GOLDEN, "appendix_e.json")))


@pytest.fixture(autouse=True)
def _need_gpu(sz):
    if sz.device_count() < 1:
        pytest.skip("no CUDA device")
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "schwarz-lib_b200"), "-s", "bin/bench_ras"])


def _run(args, cwd):
    p = subprocess.run([BIN] + args, cwd=cwd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


def test_cfg1_flags_and_output_lines(tmp_path):
    """BASELINE.json configs[0]: 100x100, regular partition, 2 ranks, CG, two-sided."""
    out = _run(["--executor=cuda", "--explicit_laplacian", "--set_1d_laplacian_size=100",
                "--partition=regular", "--overlap=2", "--local_solver=iterative-ginkgo",
                "--enable_global_check", "--set_tol=1e-6", "--local_tol=1e-12", "--num_iters=300",
                "--num_subdomains=2", "--timings_file=timings", "--write_comm_data",
                "--write_iters_and_residuals"], tmp_path)
    assert "Laplacian 2D Matrix (generated in house)" in out
    assert " Regular 1D partition" in out
    for r in (0, 1):
        assert "Subdomain %d has local problem size 5100 with 25198 non-zeros" % r in out
        assert " Rank %d converged in %d iterations" % (r, E["cfg1"]["stop_iter"]) in out
    m = re.search(r"relative residual norm of solution ([0-9.eE+-]+)", out)
    assert m and float(m.group(1)) == pytest.approx(E["cfg1"]["final_relative_residual"], rel=1e-3)
    assert re.search(r" Time taken for solve [0-9.eE+-]+", out)
    # timings CSV: header + the five stages in id order + "other"
    rows = list(csv.reader(open(tmp_path / "timings_00.csv")))
    assert rows[0] == ["func", "total", "avg", "min", "med", "max"]
    assert [r[0] for r in rows[1:]] == ["boundary_exchange", "boundary_update", "convergence_check",
                                        "local_solve", "expand_local_vec", "other"]
    # comm data: 200 halo values each way (Appendix E)
    send = open(tmp_path / "num_send_00.csv").read().splitlines()
    assert send[0] == "subdomain 0 has 1 neighbors" and send[1] == "my_id,to_id,num_send"
    assert "0,1,200" in send
    # residual log reproduces the oracle history
    hist = list(csv.reader(open(tmp_path / "iter_res_00.csv")))
    assert hist[0] == ["iter", "resnorm", "localiter", "localresnorm", "timestamp"]
    fix = json.load(open(os.path.join(GOLDEN, "cfg1_history.json")))
    got = np.array([float(r[1]) for r in hist[1:]])
    np.testing.assert_allclose(got, fix["local_resnorm_rank0"][:len(got)], rtol=1e-5, atol=1e-9)


def test_regular2d_direct_and_oversubscription(tmp_path, orc):
    """configs[4] in small: regular2d, factorised local solve, 16 subdomains on
    however many GPUs the box has."""
    out = _run(["--executor=cuda", "--explicit_laplacian", "--set_1d_laplacian_size=32",
                "--partition=regular2d", "--local_solver=direct-ginkgo", "--enable_global_check",
                "--num_iters=600", "--num_subdomains=16", "--write_perm_data"], tmp_path)
    # --write_perm_data (source/solve.cpp:434-453): the factor ordering of every rank
    perm = np.loadtxt(tmp_path / "perm_3.csv", dtype=np.int64)
    assert sorted(perm) == list(range(len(perm))) and len(perm) >= 64
    assert np.array_equal(perm, np.loadtxt(tmp_path / "inv_perm_3.csv", dtype=np.int64))
    pv = orc.partition_regular2d(32 * 32, 16)
    ob = orc.Problem(*orc.laplacian2d(32), 16, part=pv)
    ob.configure(max_iters=600, enable_global_check=True, local_solver="direct-ginkgo")
    want = ob.run()
    assert " Regular 2D partition" in out
    assert " Rank 0 converged in %d iterations" % want in out
    m = re.search(r"relative residual norm of solution ([0-9.eE+-]+)", out)
    assert m and float(m.group(1)) < 2e-6


def test_metis_gmres_from_matrix_file(tmp_path, ani4, orc, sz):
    """configs[2]: ani4_crop.mtx, METIS partition, GMRES local solve."""
    rp, ci, v = ani4
    n = len(rp) - 1
    path = tmp_path / "ani4_crop.mtx"
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (n, n, len(ci)))
        rows = np.repeat(np.arange(n), np.diff(rp))
        for r, c, x in zip(rows, ci, v):
            f.write("%d %d %.17g\n" % (r + 1, c + 1, x))
    out = _run(["--executor=cuda", "--matrix_filename=%s" % path, "--partition=metis",
                "--non_symmetric_matrix", "--restart_iter=30", "--overlap=2", "--enable_global_check",
                "--num_iters=400", "--num_subdomains=4"], tmp_path)
    part = sz.partition_metis(rp, ci, 4)
    ob = orc.Problem(rp, ci, v, 4, part=part)
    ob.configure(max_iters=400, enable_global_check=True, non_symmetric=True, restart_iter=30)
    want = ob.run()
    assert " METIS partition" in out
    assert " Rank 3 converged in %d iterations" % want in out


def test_onesided_decentralized_3d(tmp_path):
    """configs[3] in small: 3-D 7-pt Laplacian, one-sided put, decentralised flags."""
    out = _run(["--executor=cuda", "--explicit_laplacian", "--laplacian_dim=3",
                "--set_1d_laplacian_size=12", "--enable_onesided", "--remote_comm_type=put",
                "--global_convergence_type=decentralized", "--num_iters=3000",
                "--num_subdomains=4"], tmp_path)
    assert "Laplacian 3D Matrix (generated in house)" in out
    assert len(re.findall(r" Rank \d converged in \d+ iterations", out)) == 4
    m = re.search(r"relative residual norm of solution ([0-9.eE+-]+)", out)
    assert m and float(m.group(1)) < 1e-4


@pytest.mark.parametrize("precond,banner", [
    ("block-jacobi", " Local Ginkgo iterative solve(CG) with Block-Jacobi preconditioning "),
    ("ilu", " Local Ginkgo iterative solve(CG) with ParILU preconditioning "),
    ("isai", " Local Ginkgo iterative solve(CG) with ISAIpreconditioning ")])
def test_local_precond_flag(tmp_path, orc, precond, banner):
    """--local_precond / --precond_max_block_size (bench_base.hpp:70-72; solve.cpp:572-652):
    truncated local solves, so the preconditioner decides the outer iteration count."""
    out = _run(["--executor=cuda", "--explicit_laplacian", "--set_1d_laplacian_size=48",
                "--enable_global_check", "--num_iters=2000", "--num_subdomains=4",
                "--local_max_iters=8", "--local_precond=%s" % precond,
                "--precond_max_block_size=8"], tmp_path)
    assert banner in out
    ob = orc.Problem(*orc.laplacian2d(48), 4)
    ob.configure(max_iters=2000, enable_global_check=True, local_max_iters=8,
                 local_precond=precond, precond_max_block_size=8)
    want = ob.run()
    ob0 = orc.Problem(*orc.laplacian2d(48), 4)
    ob0.configure(max_iters=2000, enable_global_check=True, local_max_iters=8)
    assert want < ob0.run()                     # it does precondition
    got = int(re.search(r" Rank 0 converged in (\d+) iterations", out).group(1))
    assert abs(got - want) <= 2                 # truncated CG: see DESIGN.md section 5
    m = re.search(r"relative residual norm of solution ([0-9.eE+-]+)", out)
    assert m and float(m.group(1)) < 5e-6


def test_int64_index_type(tmp_path):
    """SolverRAS<double, int64> (include/settings.hpp:533-537 instantiates it): same run, index sets
    handed out as int64."""
    args = ["--executor=cuda", "--explicit_laplacian", "--set_1d_laplacian_size=40",
            "--partition=regular2d", "--enable_global_check", "--num_iters=400",
            "--num_subdomains=4", "--write_comm_data"]
    a = _run(args, tmp_path)
    b = _run(args + ["--index_bits=64"], tmp_path)
    pick = lambda out: sorted(l for l in out.splitlines()                       # noqa: E731
                              if "converged in" in l or "local problem size" in l)
    assert pick(a) == pick(b) and len(pick(a)) == 8
    ra = re.search(r"relative residual norm of solution ([0-9.eE+-]+)", a).group(1)
    rb = re.search(r"relative residual norm of solution ([0-9.eE+-]+)", b).group(1)
    assert ra == rb


def test_error_convention(tmp_path):
    p = subprocess.run([BIN, "--executor=omp", "--explicit_laplacian", "--num_subdomains=1"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert p.returncode == 1
    assert "Exception on processing" in p.stderr and "Aborting!" in p.stderr
