#!/bin/bash
# round 2, GPU call A: state of the tree on hardware (full suite without -x, bench of record, probes)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/r02a_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_pytest.log
MUFFT_TEST_EXPERIMENTAL=1 python -m pytest tests/test_zz_experimental_gpu.py -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/r02a_exp.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_exp.log
python bench.py --gpus 1 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "rc=$?" >> gpurun_out/r02a_bench.err
for p in real2d early latency managed; do timeout 120 python tools/probe_optin.py $p > gpurun_out/r02a_probe_$p.jsonl 2>&1; done
tail -3 gpurun_out/r02a_pytest.log; tail -3 gpurun_out/r02a_exp.log; wc -c gpurun_out/r02a_bench.json
