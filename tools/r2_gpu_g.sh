#!/bin/bash
# round 2, GPU call G: long columns in L2-sized blocks with the intermediate discarded, against whole-slab sweeps
mkdir -p gpurun_out
bash tools/ab_fast.sh tools/ab_r2g.txt > gpurun_out/r02g_ab.txt 2>&1
python -m pytest tests/test_slab_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider --timeout 900 -k "slab or c2c_2d or r2c_2d or c2r_2d" > gpurun_out/r02g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02g_pytest.log
python - > gpurun_out/r02g_slab.jsonl 2>&1 <<'PY'
import json, sys
sys.path.insert(0, ".")
import bench, mufft_b200 as m
print(json.dumps(bench.slab_also(m, 1, 16384, 16384)), flush=True)
PY
cat gpurun_out/r02g_ab.txt; tail -3 gpurun_out/r02g_pytest.log; cat gpurun_out/r02g_slab.jsonl
