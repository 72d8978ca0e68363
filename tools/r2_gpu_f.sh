#!/bin/bash
# round 2, GPU call F (2 GPUs): slab plan behind the C ABI and the torch front end, second device in one process, bench at N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02f_topo.txt 2>&1
python -m pytest tests/test_slab_gpu.py tests/test_batch_gpu.py -m gpu -q -p no:cacheprovider --timeout 900 -k "slab or second_device" > gpurun_out/r02f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02f_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err; echo "rc=$?" >> gpurun_out/r02f_bench_n2.err
python - > gpurun_out/r02f_slab.jsonl 2>&1 <<'PY'
import json, sys
sys.path.insert(0, ".")
import bench, mufft_b200 as m
for nx, ny in ((32768, 32768), (8192, 8192), (16384, 16384)):
    print(json.dumps(bench.slab_also(m, 2, nx, ny)), flush=True)
print(json.dumps(bench.slab_also(m, 1, 16384, 16384)), flush=True)
PY
tail -3 gpurun_out/r02f_pytest.log; cat gpurun_out/r02f_bench_n2.json | cut -c1-3000; cat gpurun_out/r02f_slab.jsonl
