set -x
python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02_pytest_multi.log 2>&1; echo "multi rc=$?"; tail -2 gpurun_out/r02_pytest_multi.log
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench.py --gpus $N --steps 20 --warmup 3 --policy-envs 0 --flush-steps 0 > gpurun_out/r02_bench_${N}gpu_steps20.json 2> gpurun_out/bench_n$N.err; echo "bench $N rc=$?"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29808 bench.py --gpus 8 --steps 2000 --warmup 100 --policy-envs 0 --flush-steps 0 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/bench_n8b.err; echo "bench 8 (2000) rc=$?"
