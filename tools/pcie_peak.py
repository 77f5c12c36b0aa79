"""Diagnostic: the PCIe ceiling of the host-buffer (e2e) RHS call.  Times pinned H2D, D2H and both at once for
buffers of the Brusselator 4096^2 state size (268 MB each way), CUDA events, best of 5."""
import torch
n = 2 * 4096 * 4096
dev = torch.device("cuda", 0)
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device=dev)
d_out = torch.empty(n, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn):
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)


gb = n * 8 / 1e9
t = timed(lambda: d_in.copy_(h_in, non_blocking=True)); print(f"H2D alone: {t:.2f} ms = {gb / t * 1e3:.1f} GB/s")
t = timed(lambda: h_out.copy_(d_out, non_blocking=True)); print(f"D2H alone: {t:.2f} ms = {gb / t * 1e3:.1f} GB/s")
t = timed(both); print(f"H2D + D2H concurrently: {t:.2f} ms = {2 * gb / t * 1e3:.1f} GB/s total -> ceiling of the e2e RHS call: "
                       f"{4096 * 4096 / (t * 1e-3):.3e} grid-point updates/s")
