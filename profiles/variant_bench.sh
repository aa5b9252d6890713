#!/bin/bash
# Runs bench.py once per prebuilt kernel variant (build/variants/lib_*.so) and prints one line each:
# the laned headline (2 launches per step on 2 streams) and the serialized launch (one launch per step, one stream).
for lib in build/variants/lib_*.so; do
  MERGING_B200_LIB=$PWD/$lib python bench.py --steps 2000 --warmup 50 --no-cpu-baseline --policy-envs 0 --strong-envs 0 \
      --flush-steps 0 --lean 0 --rollout-k 0 --e2e-steps 3 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['serialized']['at_headline_steps']; print('$lib', 'laned %.2f us frac %.3f' % (d['ms_per_step']*1e3, d['roofline']['frac']), '| serialized %.2f us frac %.3f' % (s['ms_per_step']*1e3, s['frac_of_peak']), '| l2_warm %.2f us' % (d['l2_warm']['ms_per_step']*1e3), d['clocks']['sm_mhz'])"
done
