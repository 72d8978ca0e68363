#!/bin/bash
# multi-GPU session: slab tests, slab benches with both exchange paths
nvidia-smi -L
make -s -C mufft_b200/csrc -j32 >/dev/null 2>&1
timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -x -q 2>&1 | tail -15
N=${1:-2}
for ex in p2p nccl; do for wl in slab2d8k slab2d; do
MUFFT_SLAB_EXCHANGE=$ex timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $wl --steps 5 --warmup 3 --also '' 2> gpurun_out/slab${N}_${wl}_$ex.err | grep '^{' > gpurun_out/slab${N}_${wl}_$ex.json; python - <<PY
import json
try:
    r=json.loads(open('gpurun_out/slab${N}_${wl}_$ex.json').read()); print('$wl x$N $ex %.3f ms %.0f GFLOP/s per-GPU %.0f GB/s' % (r['ms_per_step'], r['value'], r['roofline']['achieved']))
except Exception as e:
    print('$wl $ex failed', e); print(open('gpurun_out/slab${N}_${wl}_$ex.err').read()[-2500:])
PY
done; done
