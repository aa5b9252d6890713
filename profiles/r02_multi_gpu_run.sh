set -x
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
nproc > gpurun_out/nproc8.txt; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)" >> gpurun_out/nproc8.txt
python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02_pytest_multi.log 2>&1; echo "multi rc=$?"; tail -3 gpurun_out/r02_pytest_multi.log
python profiles/d2h_ceiling.py --out gpurun_out/d2h_ceiling_n1.json > /dev/null 2>gpurun_out/d2h1.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) profiles/d2h_ceiling.py --out gpurun_out/d2h_ceiling_n$N.json > /dev/null 2> gpurun_out/d2h$N.err
done
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench.py --gpus $N --steps 20 --warmup 3 --policy-envs 0 --flush-steps 0 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench $N rc=$?"
done
