"""Measures pinned host<->device copy bandwidth on this box (one direction and both at once); used to size the
end-to-end pipeline of bench.py."""
import torch, time
n = 32 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(f, reps=20):
    f(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps
def d2h():
    with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
def h2d():
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
def both(): d2h(); h2d()
print("D2H %.1f GB/s" % (n / run(d2h) / 1e9)); print("H2D %.1f GB/s" % (n / run(h2d) / 1e9))
t = run(both); print("both: %.1f GB/s each direction" % (n / t / 1e9))
