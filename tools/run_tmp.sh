nvidia-smi -L
timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -x -q 2>&1 | tail -3
for wl in slab2d8k slab2d; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $wl --steps 5 --warmup 3 --also '' 2> gpurun_out/slab2_$wl.err | grep '^{' > gpurun_out/slab2_$wl.json; python - <<PY
import json
try:
    r=json.loads(open('gpurun_out/slab2_$wl.json').read()); print('$wl x2 %.3f ms %.0f GFLOP/s per-GPU %.0f GB/s' % (r['ms_per_step'], r['value'], r['roofline']['achieved']))
except Exception as e:
    print('$wl failed', e); print(open('gpurun_out/slab2_$wl.err').read()[-1500:])
PY
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --also '' 2> gpurun_out/bench2.err | grep '^{' > gpurun_out/bench2.json; cut -c1-900 gpurun_out/bench2.json
