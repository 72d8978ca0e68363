"""PCIe probe on the GPU box: raw pinned copy rates (one way, both ways at once) and the e2e step of
bench.py's default workload for a few staging configurations."""
import os, subprocess, sys, time, json
import torch
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=4):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both(): h2d(); d2h()
print("H2D alone  %.1f GB/s" % (n / timed(h2d) / 1e9))
print("D2H alone  %.1f GB/s" % (n / timed(d2h) / 1e9))
t = timed(both); print("both ways  %.1f + %.1f GB/s" % (n / t / 1e9, n / t / 1e9), flush=True)
del h_in, h_out, d_a, d_b
for slots, mb in ((3, 32), (4, 32), (3, 64), (4, 16), (6, 16), (4, 64), (2, 128)):
    env = dict(os.environ, MUFFT_CUDA_STAGE_SLOTS=str(slots), MUFFT_CUDA_STAGE_CHUNK_MB=str(mb))
    out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-cpu", "--also", ""], env=env, capture_output=True, text=True).stdout
    r = json.loads(out.strip().splitlines()[-1])
    print("slots %d chunk %3d MiB: e2e %.1f ms/step  %.0f GFLOP/s" % (slots, mb, r["e2e"]["ms_per_step"], r["e2e"]["value"]), flush=True)
