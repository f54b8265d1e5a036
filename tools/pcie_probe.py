"""PCIe probe for DESIGN.md §7: pinned H2D / D2H bandwidth alone and concurrently (two streams), at the chunk sizes the
host pipeline uses (fixed base: 303104 items x 32 B in, x 65 B out) and at large sizes."""
import torch

def bw(fn, nbytes, reps=8):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return reps * nbytes / e0.elapsed_time(e1) / 1e6

for label, n_in, n_out in (("chunk", 303104 * 32, 303104 * 65), ("4 chunks", 4 * 303104 * 32, 4 * 303104 * 65),
                           ("256 MiB", 256 << 20, 256 << 20)):
    hi = torch.empty(n_in, dtype=torch.uint8).pin_memory(); di = torch.empty(n_in, dtype=torch.uint8, device="cuda")
    ho = torch.empty(n_out, dtype=torch.uint8).pin_memory(); do = torch.empty(n_out, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h2d = bw(lambda: di.copy_(hi, non_blocking=True), n_in)
    d2h = bw(lambda: ho.copy_(do, non_blocking=True), n_out)
    def both():
        with torch.cuda.stream(s1):
            di.copy_(hi, non_blocking=True)
        with torch.cuda.stream(s2):
            ho.copy_(do, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    both(); torch.cuda.synchronize()
    e0.record()
    for _ in range(8):
        both()
    s1.synchronize(); s2.synchronize()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    print(f"{label}: H2D {h2d:.1f} GB/s, D2H {d2h:.1f} GB/s, concurrent: {8 * n_in / t / 1e6:.1f} + {8 * n_out / t / 1e6:.1f} GB/s "
          f"-> {8 * (n_in // 32) / t / 1e3:.0f} M items/s bound for 32 B in / 65 B out")
