"""bench.py's output contract, checked on CPU through the arm that needs no GPU (`--impl reference` =
the plain-C oracle port on the host cores): one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    return lines


def test_reference_arm_prints_one_contract_line():
    lines = run_bench("--impl", "reference", "--gpus", "1", "--steps", "3", "--warmup", "3", "--ref-python-steps", "300")
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["value"] > 0 and abs(d["value"] - d["config"]["envs_per_gpu"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    # the reference's OWN Python env, timed in the same run (here from /root/reference, on the GPU box from baseline/_ref)
    ref = cb["reference_python"]
    assert "unavailable" not in ref and "error" not in ref, ref
    assert 100 < cb["reference_python_value"] < 1e6 and cb["reference_python_all_cores_value"] > 0
    assert ref["pve_1_process"] > ref["pvp_1_process"] > 0            # pve solves one QP per step, pvp two
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_under_torchrun_env_only_rank0_prints():
    """N > 1: rank 0 alone runs and prints; the other ranks exit 0 without work (no rendezvous needed)."""
    base = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r1 = run_bench("--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3", "--ref-python-steps", "0",
                   env=dict(base, RANK="1", LOCAL_RANK="1"))
    assert r1 == []
    r0 = run_bench("--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3", "--ref-python-steps", "0",
                   env=dict(base, RANK="0", LOCAL_RANK="0"))
    assert len(r0) == 1 and json.loads(r0[0])["n_gpus"] == 2


def test_both_arms_word_the_workload_identically():
    """The driver compares the two arms' `config` dicts (`same_config`)."""
    sys.path.insert(0, ROOT)
    import bench
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": workload_config(n)') == 2          # the reference arm's line and the GPU arm's line
    c = bench.workload_config(1 << 20)
    assert "workload" in c and "l2" in c and c["envs_per_gpu"] == 1 << 20


import pytest


@pytest.mark.gpu
def test_gpu_arm_contract_line():
    """The product arm on one GPU, short run: every key of the contract, roofline consistent with the timing,
    e2e bytes counted from the copied tensors, launches counted."""
    lines = run_bench("--gpus", "1", "--steps", "48", "--warmup", "3", "--cpu-seconds", "1", "--flush-steps", "4",
                      "--rollout-k", "4", "--e2e-steps", "4", "--sustained-steps", "200", "--strong-envs", "65536",
                      "--policy-envs", "4096", "--ref-python-steps", "200")
    assert len(lines) == 1
    d = json.loads(lines[0])
    n = d["config"]["envs_per_gpu"]
    assert "impl" not in d and d["metric"] == "env_steps_per_sec" and d["n_gpus"] == 1 and d["steps"] == 48
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert abs(d["value"] - n / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - 156 * n / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert 0.3 < r["frac"] < 1.05
    m = d["measurement"]
    assert m["launches_per_step"] == 2 and d["gpu_launches"] == 48 * 2           # headline: MergeVecEnv(lanes=2)
    assert d["serialized"]["value"] > 0 and d["serialized"]["at_headline_steps"]["steps"] == 48
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 2 * n and e["d2h_bytes_per_step"] == 50 * n and 0 < e["value"] < d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert "l2" in d["config"] and "workload" in d["config"]
    assert d["overlapped_streams"]["value"] > 0 and d["l2_warm"]["value"] > 0
    assert d["sustained"]["steps"] == 200 and d["lean_no_returns"]["bytes_per_env_step"] == 124
    assert e["sync_value"] > 0 and e["rew_done_info_value"] > e["value"] and 0 < e["frac_of_ceiling"] <= 1.05
    st = d["strong_8m"]
    assert st["total_envs"] == 65536 and st["steps"] == 256 and len(st["stats_digest"]) == 16 and st["stats_totals"][0] > 0
    assert d["policy_in_loop"]["fused"]["value"] > 0 and 0 < d["policy_in_loop"]["tf32x3"]["env_share"] < 1
    assert d["policy_in_loop"]["tf32x3"]["fused_step"]["value"] > 0 and d["policy_in_loop"]["tf32x3_4096_envs"]["value"] > 0
    assert d["policy_in_loop"]["f16x3"]["value"] > 0 and d["policy_in_loop"]["f16x3"]["fused_step"]["value"] > 0
    assert d["policy_in_loop"]["f16x3_4096_envs"]["fused_step"]["value"] > 0
    assert cb["reference_python_value"] > 0
