#!/bin/bash
# one multi-GPU box: the scaling evidence for profiles/ (bench at 2/4/8 GPUs, bit-identity check, config 5 at 2/4/8, config 4 at 8)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4 8; do
  timeout 240 $TR --nproc-per-node $n --master-port $((29600+n)) bench.py --gpus $n --steps 500 --warmup 10 --no-cpu-baseline 2> gpurun_out/bench_r2_n$n.err | grep '^{' > gpurun_out/bench_r2_n$n.json
  echo "bench n=$n: $(python -c "import json;d=json.load(open('gpurun_out/bench_r2_n$n.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'])" 2>&1)"
done
timeout 240 $TR --nproc-per-node 8 --master-port 29620 scripts/multi_gpu_check.py velodyne_30k_1m p2p 2>&1 | grep "^\[p2p\]" > gpurun_out/mg8_check.log; tail -2 gpurun_out/mg8_check.log
for n in 2 4 8; do
  timeout 400 $TR --nproc-per-node $n --master-port $((29640+n)) scripts/dense_check.py dense_200k_20m noref 2> gpurun_out/dense_r2_n$n.err | grep '^{' > gpurun_out/dense_r2_n$n.json
  echo "dense n=$n: $(cut -c1-400 gpurun_out/dense_r2_n$n.json)"
done
timeout 400 $TR --nproc-per-node 8 --master-port 29660 scripts/stream_bench.py 300 0 2> gpurun_out/stream_r2_n8.err | grep '^{' > gpurun_out/stream_r2_n8.json
echo "stream n=8: $(cut -c1-600 gpurun_out/stream_r2_n8.json)"
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --target-processes all --print-limit 5 $TR --nproc-per-node 2 --master-port 29680 scripts/multi_gpu_check.py tiny p2p > gpurun_out/sanitize_memcheck_p2p2.log 2>&1
echo "memcheck 2-rank p2p: $(grep -E 'ERROR SUMMARY' gpurun_out/sanitize_memcheck_p2p2.log | sort | uniq -c | tr '\n' ';')"
