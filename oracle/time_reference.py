#!/usr/bin/env python
"""TEST INFRASTRUCTURE — times the UNMODIFIED reference env (with oracle/ref_shims) in the build
container, where /root/reference exists: BASELINE.json configs[0] (pve, one env, random actions,
10 000 steps) plus the pvp variant, one process and one process per core.  The result is written to
profiles/r01_reference_cpu_container.json.  The reference cannot travel to the GPU box, so this is
the only place its own code can be timed; the shims are lighter than real pygame/shapely/quadprog,
so these figures over-estimate the reference's speed.
"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def run(args):
    pvp, steps, seed = args
    import warnings
    warnings.simplefilter("ignore")
    from oracle.ref_loader import load_reference_env, quiet
    env = load_reference_env()
    rng = np.random.default_rng(seed)
    acts = rng.integers(0, 5, (steps, 2)).tolist()
    env.reset()
    with quiet():
        t0 = time.perf_counter()
        for a1, a2 in acts:
            _, _, done, _ = env.step(a1, a2 if pvp else None)
            if done:
                env.reset()
        dt = time.perf_counter() - t0
    return steps / dt


def main():
    steps = 10000
    cores = os.cpu_count() or 1
    out = {"host": "build container", "cores": cores, "steps_per_process": steps,
           "note": "unmodified merging_env.py + helper.py against oracle/ref_shims (gym/pygame/shapely/qpsolvers "
                   "stand-ins); env-steps/s"}
    for name, pvp in (("pve", False), ("pvp", True)):
        out[f"{name}_1_process"] = run((pvp, steps, 0))
        with mp.get_context("spawn").Pool(cores) as pool:
            t0 = time.perf_counter()
            pool.map(run, [(pvp, steps, s) for s in range(cores)])
            out[f"{name}_{cores}_processes_aggregate"] = cores * steps / (time.perf_counter() - t0)
    path = os.path.join(os.path.dirname(HERE), "profiles", "r01_reference_cpu_container.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
