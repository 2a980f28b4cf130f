"""Pinned-memory copy bandwidth of the box (H2D, D2H, both at once): the e2e number's ceiling."""
import torch, time
n = 1 << 30
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(f, reps=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
print("H2D GB/s %.1f" % (n / run(h2d) / 1e9)); print("D2H GB/s %.1f" % (n / run(d2h) / 1e9))
print("both: each direction GB/s %.1f" % (n / run(both) / 1e9))
import os; print("cpus", os.cpu_count())
