#!/bin/bash
# round 2, GPU call J (2 GPUs): slab plan with the four-step column kernel; slab tests; bench at N=2
mkdir -p gpurun_out
python - > gpurun_out/r02j_slab.jsonl 2>&1 <<'PY'
import json, sys
sys.path.insert(0, ".")
import bench, mufft_b200 as m
for nx, ny in ((32768, 32768), (16384, 16384), (8192, 8192)):
    print(json.dumps(bench.slab_also(m, 2, nx, ny)), flush=True)
print(json.dumps(bench.slab_also(m, 1, 16384, 16384)), flush=True)
PY
python -m pytest tests/test_slab_gpu.py tests/test_batch_gpu.py -m gpu -q -p no:cacheprovider --timeout 900 -k "slab or second_device or long_columns" > gpurun_out/r02j_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02j_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02j_bench_n2.json 2> gpurun_out/r02j_bench_n2.err; echo "rc=$?" >> gpurun_out/r02j_bench_n2.err
cat gpurun_out/r02j_slab.jsonl; tail -3 gpurun_out/r02j_pytest.log; cut -c1-400 gpurun_out/r02j_bench_n2.json
