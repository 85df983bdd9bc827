"""The drop-in driver without a GPU: it builds, carries the reference's flag surface
(benchmarking/bench_base.hpp:50-144, names / types / defaults), and FAILS LOUDLY when there is no
CUDA device or a CPU executor is asked for - there is no CPU compute path behind it."""
import os
import re
import subprocess

import pytest

from conftest import ROOT

BIN = os.path.join(ROOT, "schwarz-lib_b200", "bin", "bench_ras")

# flag -> (type, default) of the reference
REFERENCE_FLAGS = {
    "executor": ("string", "reference"), "set_tol": ("double", "1e-06"), "num_iters": ("uint32", "100"),
    "num_threads": ("uint32", "1"), "set_1d_laplacian_size": ("uint32", "16"),
    "enable_debug_write": ("bool", "false"), "write_perm_data": ("bool", "false"),
    "write_iters_and_residuals": ("bool", "false"), "print_matrices": ("bool", "false"),
    "shifted_iter": ("uint32", "1"), "enable_onesided": ("bool", "false"),
    "remote_comm_type": ("string", "get"), "enable_one_by_one": ("bool", "false"),
    "enable_comm_overlap": ("bool", "false"), "flush_type": ("string", "flush-all"),
    "lock_type": ("string", "lock-all"), "enable_put_all_local_residual_norms": ("bool", "false"),
    "enable_global_check_iter_offset": ("bool", "false"), "enable_global_check": ("bool", "false"),
    "global_convergence_type": ("string", "centralized-tree"),
    "enable_decentralized_accumulate": ("bool", "false"), "local_tol": ("double", "1e-12"),
    "local_precond": ("string", "null"), "local_max_iters": ("int32", "-1"),
    "non_symmetric_matrix": ("bool", "false"), "restart_iter": ("uint32", "1"),
    "precond_max_block_size": ("uint32", "16"), "matrix_filename": ("string", "null"),
    "explicit_laplacian": ("bool", "false"), "enable_random_rhs": ("bool", "false"),
    "overlap": ("uint32", "2"), "factor_ordering_natural": ("bool", "false"),
    "local_reordering": ("string", "none"), "local_factorization": ("string", "cholmod"),
    "partition": ("string", "regular"), "metis_objtype": ("string", "null"),
    "local_solver": ("string", "iterative-ginkgo"), "debug": ("bool", "false"),
    "print_config": ("bool", "true"), "timings_file": ("string", "null"),
    "write_comm_data": ("bool", "false"), "use_mixed_precision": ("bool", "false"),
    "stage_through_host": ("bool", "false"), "enable_logging": ("bool", "false"),
    "updated_max_iters": ("int32", "-1"), "reset_local_crit_iter": ("int32", "-1"),
    "enable_twosided": ("bool", "true"),
}


@pytest.fixture(scope="module")
def binary(sz):
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "schwarz-lib_b200"), "-s", "bin/bench_ras"])
    return BIN


def test_flag_surface_is_the_references(binary):
    out = subprocess.run([binary, "--help"], capture_output=True, text=True, timeout=60).stdout
    found = {m.group(1): (m.group(2), m.group(3))
             for m in re.finditer(r"--(\w+) \(.*?\)\s+type: (\w+)\s+default: (\S+)", out)}
    for name, (typ, default) in REFERENCE_FLAGS.items():
        assert name in found, name
        assert found[name][0] == typ, (name, found[name])
        if typ == "double":
            assert float(found[name][1]) == float(default), (name, found[name])
        else:
            assert found[name][1].strip('"') == default, (name, found[name])
    # what this driver adds on top: where the subdomains run, the 3-D generator, the index type
    assert set(found) - set(REFERENCE_FLAGS) == {"num_subdomains", "num_devices", "laplacian_dim",
                                                 "index_bits"}


def test_cpu_executor_is_rejected_with_the_references_banner(binary, tmp_path):
    p = subprocess.run([binary, "--executor=omp", "--explicit_laplacian", "--num_subdomains=1"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert p.returncode == 1
    assert "Exception on processing" in p.stderr and "Aborting!" in p.stderr
    assert "CUDA devices only" in p.stderr


def test_unknown_flag_is_an_error(binary, tmp_path):
    p = subprocess.run([binary, "--no_such_flag=1"], cwd=tmp_path, capture_output=True, text=True,
                       timeout=60)
    assert p.returncode == 1 and "unknown command line flag 'no_such_flag'" in p.stderr


def test_no_cuda_device_is_fatal_not_a_fallback(binary, sz, tmp_path):
    if sz.device_count() > 0:
        pytest.skip("this box has a GPU")
    p = subprocess.run([binary, "--executor=cuda", "--explicit_laplacian", "--num_subdomains=2"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert p.returncode != 0
    assert "No CUDA devices available" in p.stdout + p.stderr
    assert "converged" not in p.stdout


def test_flag_table_above_is_what_the_reference_defines():
    """only where /root/reference is mounted (this container): REFERENCE_FLAGS == the DEFINE_*
    statements of benchmarking/bench_base.hpp, name / type / default."""
    path = "/root/reference/benchmarking/bench_base.hpp"
    if not os.path.exists(path):
        pytest.skip("/root/reference not present")
    src = open(path).read()
    defs = re.findall(r"DEFINE_(\w+)\(\s*(\w+)\s*,\s*([^,]+?)\s*,", src)
    assert len(defs) == len(REFERENCE_FLAGS) == 47
    for typ, name, default in defs:
        want_type, want_default = REFERENCE_FLAGS[name]
        assert typ == want_type, name
        default = default.strip().strip('"')
        if typ == "double":
            assert float(default) == float(want_default), name
        else:
            assert default == want_default, name
