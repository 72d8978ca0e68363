#!/bin/bash
# round 2, GPU call K (8 GPUs): bench at N=8 with also.slab2d / slab2d_r2c; slab sweep; the C slab plan against the oracle on 8 devices
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02k_topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02k_bench_n8.json 2> gpurun_out/r02k_bench_n8.err; echo "rc=$?" >> gpurun_out/r02k_bench_n8.err
python - > gpurun_out/r02k_slab.jsonl 2>&1 <<'PY'
import json, sys
sys.path.insert(0, ".")
import bench, mufft_b200 as m
for ndev in (8, 4):
    for nx, ny, real in ((32768, 32768, False), (32768, 32768, True), (16384, 16384, False), (8192, 8192, False)):
        print(json.dumps(bench.slab_also(m, ndev, nx, ny, real=real)), flush=True)
PY
python -m pytest tests/test_slab_gpu.py -m gpu -q -p no:cacheprovider --timeout 900 -k "slab_plan" > gpurun_out/r02k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02k_pytest.log
grep -v "^NCCL" gpurun_out/r02k_bench_n8.json | cut -c1-3500; cat gpurun_out/r02k_slab.jsonl; tail -3 gpurun_out/r02k_pytest.log
