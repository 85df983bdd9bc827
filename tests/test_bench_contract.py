"""bench.py's reference arm (the CPU leg of the contract) runs without a GPU: one JSON line with the
keys the driver reads, rank 0 only under a multi-rank launch."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                           "--steps", "2", "--warmup", "1"] + extra,
                          capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    p = _run(["--size", "256"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ras_outer_iters_per_s"
    assert d["unit"] == "outer iters/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    # up to 2048^2 the reference's own loop (oracle/_ref) is what is timed
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1
    assert cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_steps_the_real_loop_with_every_core_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1 to its workers: the arm sizes its team from the CPU
    affinity mask instead, really runs every requested step of the oracle's outer loop (no
    extrapolation), and never loads the product library."""
    code = (
        "import os, sys, json, io, contextlib\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--size', '2560', '--steps', '3', '--warmup', '2']\n"
        "sys.path.insert(0, %r)\n"
        "import bench\n"
        "buf = io.StringIO()\n"
        "with contextlib.redirect_stdout(buf):\n"
        "    bench.main()\n"
        "d = json.loads(buf.getvalue())\n"
        "maps = open('/proc/self/maps').read()\n"
        "d['_product_loaded'] = 'libschwz_b200' in maps or 'schwz_b200' in sys.modules\n"
        "d['_oracle_loaded'] = 'libschwz_oracle' in maps\n"
        "print(json.dumps(d))\n" % ROOT)
    e = dict(os.environ)
    e["OMP_NUM_THREADS"] = "1"
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900,
                       env=e, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == len(os.sched_getaffinity(0))
    assert d["steps"] == 3 and "3 outer iteration(s) of the oracle's own RAS loop" in cb["sample"]
    assert d["_oracle_loaded"] and not d["_product_loaded"]


def test_reference_arm_other_ranks_stay_silent():
    p = _run(["--size", "256", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_reference_arm_cfg3():
    p = _run(["--matrix", "ani4"])
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and "cfg3" in d["config"]["workload"] and d["value"] > 0


def test_b200_arm_keeps_stdout_to_the_one_json_line():
    """whatever torch / NCCL / the library write to fd 1 while the bench runs (NCCL prints its version
    banner there under torchrun) ends up on stderr; stdout carries the JSON line alone.  The GPU
    part is stubbed out: this checks the plumbing around it."""
    code = (
        "import os, sys\n"
        "sys.argv = ['bench.py']\n"
        "sys.path.insert(0, %r)\n"
        "import bench\n"
        "def fake(args):\n"
        "    print('python-level noise')\n"
        "    os.write(1, b'NCCL version 2.28.9+cuda12.9\\n')\n"
        "    os.system('echo child-process noise')\n"
        "    return '{\"metric\": \"ras_outer_iters_per_s\"}'\n"
        "bench.run_b200 = fake\n"
        "bench.main()\n" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    assert p.stdout == '{"metric": "ras_outer_iters_per_s"}\n'
    for noise in ("python-level noise", "NCCL version", "child-process noise"):
        assert noise in p.stderr
