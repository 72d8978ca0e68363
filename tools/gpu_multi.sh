#!/bin/bash
# Multi-GPU session (gpurun --gpus N): slab tests over the visible GPUs, batch-sharded bench under
# torchrun, slab-decomposed single large 2D transforms (p2p and NCCL exchange), reference arm.
N=${1:-2}
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 600 python -m pytest tests/test_slab_gpu.py -m gpu -q -x -k "256-128 or 2048-4096 or 4096-2048 or 1024-32768" > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/pytest_multi_$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$? t=$((SECONDS-T0))"; tail -1 gpurun_out/bench_n$N.json | cut -c1-600
timeout 300 $TR bench.py --gpus $N --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref arm rc=$? t=$((SECONDS-T0))"; tail -1 gpurun_out/bench_ref_n$N.json | cut -c1-400
: > gpurun_out/slab_n$N.jsonl
for wl in slab2d8k slab2d slab2d8k_r2c slab2d_r2c; do
  for ex in p2p nccl; do
    [ "$ex" = p2p ] && [[ $wl == *r2c ]] && continue
    res=$(MUFFT_SLAB_EXCHANGE=$ex timeout 300 $TR bench.py --gpus $N --workload $wl --steps 10 --warmup 3 2> gpurun_out/slab_err.log | tail -1)
    echo "{\"exchange\": \"$ex\", \"bench\": $res}" >> gpurun_out/slab_n$N.jsonl
    echo "$wl $ex: $(echo "$res" | python -c 'import json,sys; r=json.loads(sys.stdin.read()); print("%.3f ms/step %.0f GFLOP/s" % (r["ms_per_step"], r["value"]))' 2>&1 | tail -1)  t=$((SECONDS-T0))"
  done
done
tail -5 gpurun_out/slab_err.log
