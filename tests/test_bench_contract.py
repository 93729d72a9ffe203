"""bench.py's reference arm runs on the CPU alone: its JSON line must carry the contract's keys and
must not inherit a single OpenMP thread from the launcher (torch.distributed.run exports
OMP_NUM_THREADS=1 to its workers; the round-1 scaling record was void because of that)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_reference_arm(extra_env):
    env = dict(os.environ, **extra_env)
    done = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8",
                           "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600,
                          cwd=ROOT, env=env)
    assert done.returncode == 0, done.stderr[-2000:]
    return done.stdout.strip().splitlines()


def test_reference_arm_line_and_thread_count():
    lines = run_reference_arm({"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "8"})
    line = json.loads(lines[-1])
    assert line["impl"] == "reference" and line["metric"] == "GStencil/s" and line["unit"] == "GStencil/s"
    assert line["n_gpus"] == 8 and line["steps"] == 1 and line["higher_is_better"] is True
    assert line["scaling"] == "strong" and line["vs_baseline"] is None
    assert "C5 Jacobi3D" in line["config"]["workload"] and line["config"]["cpu_grid"] == [256, 1024, 1024]
    assert line["e2e"] == {"value": line["value"], "unit": "GStencil/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    baseline = line["cpu_baseline"]
    usable = len(os.sched_getaffinity(0))
    assert baseline["kind"] == "port" and baseline["value"] == line["value"]
    assert baseline["omp_get_max_threads"] == usable == baseline["cores"]      # not the launcher's 1
    assert "omp_get_max_threads()={}".format(usable) in baseline["sample"]


def test_other_ranks_of_the_reference_arm_do_nothing():
    assert run_reference_arm({"RANK": "3", "WORLD_SIZE": "8"}) == []
