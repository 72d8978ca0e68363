#!/bin/bash
# 8-GPU session: batch-sharded bench under torchrun and the slab-decomposed 32768^2 transforms.
N=${1:-8}
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1; nproc > gpurun_out/nproc_$N.txt; free -g >> gpurun_out/nproc_$N.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
: > gpurun_out/slab_n$N.jsonl
for spec in "slab2d p2p" "slab2d nccl" "slab2d_r2c nccl" "slab2d8k p2p"; do
  set -- $spec; wl=$1; ex=$2
  res=$(MUFFT_SLAB_EXCHANGE=$ex timeout 240 $TR bench.py --gpus $N --workload $wl --steps 10 --warmup 3 2> gpurun_out/slab_err_$N.log | tail -1)
  echo "{\"exchange\": \"$ex\", \"bench\": $res}" >> gpurun_out/slab_n$N.jsonl
  echo "$wl $ex: $(echo "$res" | python -c 'import json,sys; r=json.loads(sys.stdin.read()); print("%.3f ms/step %.0f GFLOP/s" % (r["ms_per_step"], r["value"]))' 2>&1 | tail -1)  t=$((SECONDS-T0))"
done
timeout 400 $TR bench.py --gpus $N --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$? t=$((SECONDS-T0))"; tail -1 gpurun_out/bench_n$N.json | cut -c1-400
tail -3 gpurun_out/slab_err_$N.log
