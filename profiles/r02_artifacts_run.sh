set -x
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo rc=$?
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu_steps20.json 2>/dev/null; echo rc=$?
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; echo rc=$?
F="--steps 24 --warmup 3 --no-graph --no-cpu-baseline --rollout-k 0 --flush-steps 0 --policy-envs 0 --strong-envs 0 --lean 0 --sustained-steps 0 --e2e-steps 3 --overlap-streams 0"
python bench.py $F > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_steps24.csv python bench.py $F > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:merge_step -s 30 -c 8 -o gpurun_out/r02_step python bench.py $F > gpurun_out/ncu2.log 2>&1
python profiles/ncu_target.py mlp_tc > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mlp_act_tc -s 2 -c 1 -o gpurun_out/r02_mlp_tc python profiles/ncu_target.py mlp_tc > gpurun_out/ncu3.log 2>&1
python profiles/record_bench.py > gpurun_out/r02_record_bench.json 2>&1
python profiles/misc_kernels.py > gpurun_out/r02_misc_kernels.json 2>&1
python profiles/rollout_bench.py > gpurun_out/r02_rollout_bench.json 2>&1
