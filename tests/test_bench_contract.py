"""-m "not gpu": the bench.py contract on the arm that runs without a GPU (--impl reference = the reference's own
Ipopt + MUMPS binaries, or the C port when oracle/_ref is absent), and the static pieces of the GPU arm."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-per-core", "10"],
                         capture_output=True, text=True, check=True, cwd=ROOT).stdout.strip().splitlines()
    assert len(out) == 1
    d = json.loads(out[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "mpc_solves_per_sec" and d["unit"] == "solves/s"
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]
    # the config describes the workload (identical in both arms); what this arm really solved per step is stated beside it
    assert d["config"]["problems_per_step"] == 65536 and d["scaling"] == "strong"
    assert d["solved_per_step"] == 10 * d["cpu_baseline"]["cores"]


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2"],
                       capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_roofline_traffic_comes_from_the_committed_launch_list():
    sys.path.insert(0, ROOT)
    import bench
    t = bench.profiled_counters()["dram_bytes_per_step"]
    assert t is not None and 1e9 < t < 1e11
    assert bench.f_iter(25) == 62204
