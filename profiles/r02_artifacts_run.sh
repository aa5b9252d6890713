set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_1gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_gpu_1gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo rc=$?
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu_steps20.json 2>/dev/null; echo rc=$?
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; echo rc=$?
python bench_policy.py --backend tf32x3 > gpurun_out/r02_bench_policy_tf32x3.json
python bench_policy.py --backend fused > gpurun_out/r02_bench_policy_fused.json
python bench_policy.py --backend f16x3 > gpurun_out/r02_bench_policy_f16x3.json
python bench_policy.py --backend f16x3 --policy hdqn > gpurun_out/r02_bench_policy_hdqn_f16x3.json
python profiles/record_bench.py > gpurun_out/r02_record_bench.json 2>/dev/null
python profiles/f16x3_check.py > gpurun_out/r02_f16x3_check.jsonl 2>/dev/null
./build/exp_tc16_trace > gpurun_out/r02_mlp_tc16_trace.log 2>&1
python bench_policy.py --backend tf32x3 --policy hdqn > gpurun_out/r02_bench_policy_hdqn_tf32x3.json
python profiles/graphed_rollout.py > gpurun_out/r02_graphed_rollout.json 2>/dev/null
F="--steps 24 --warmup 3 --no-graph --no-cpu-baseline --rollout-k 0 --flush-steps 0 --policy-envs 0 --strong-envs 0 --lean 0 --sustained-steps 0 --e2e-steps 3 --overlap-streams 0"
python bench.py $F > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_steps24.csv python bench.py $F > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:merge_step -s 30 -c 8 -o gpurun_out/r02_step python bench.py $F > gpurun_out/ncu2.log 2>&1
python profiles/ncu_target.py mlp_tc > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mlp_act_tc -s 2 -c 1 -o gpurun_out/r02_mlp_tc python profiles/ncu_target.py mlp_tc > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_act_tc -s 2 -c 1 -o gpurun_out/r02_policy_step_tc python profiles/ncu_target.py policy_step_tc > gpurun_out/ncu4.log 2>&1
TC_BACKEND=f16x3 ncu --set full --clock-control none --import-source on -k regex:tc16 -s 3 -c 1 -o gpurun_out/r02_mlp_tc16 -f python profiles/tc_time.py > gpurun_out/ncu5.log 2>&1
