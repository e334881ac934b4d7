"""Host <-> device copy bandwidth of the box (pinned memory), one direction at a time and both at once:
the bound of the host-buffer entry points (bench.py's e2e leg)."""
import torch, time
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / reps
    return n / dt / 1e9
run(True, True, 1)
print("H2D alone %.1f GB/s, D2H alone %.1f GB/s, both at once %.1f GB/s each" % (run(True, False), run(False, True), run(True, True)))
