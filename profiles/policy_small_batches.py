import json, sys
sys.path.insert(0, "/root/repo")
import bench_policy
for n in (4096, 16384, 65536):
    for be in ("tf32x3", "f16x3"):
        r = bench_policy.measure(n, "dqn", be, k=32, replays=8)
        print(json.dumps({"envs": n, "backend": be, "us_per_step": 1e3 * r["ms_per_step"], "policy_us": 1e3 * r["policy_ms"],
                          "fused_us": None if not r["fused_step"] else 1e3 * r["fused_step"]["ms_per_step"]}))
