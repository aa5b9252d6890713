#!/bin/bash
# Runs bench.py once per prebuilt kernel variant (build/variants/lib_*.so) and prints one line each.
for lib in build/variants/lib_*.so; do
  MERGING_B200_LIB=$PWD/$lib python bench.py --steps 2000 --warmup 50 --no-cpu-baseline 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', '%.3e steps/s' % d['value'], 'warm %.3e' % d['l2_warm']['value'], '%.2f us' % (d['ms_per_step']*1e3), 'frac %.3f' % d['roofline']['frac'], d['clocks'])"
done
