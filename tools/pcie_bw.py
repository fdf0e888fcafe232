"""Host<->device copy bandwidth of the box (pinned memory), one direction and both at once: the bound
of bench.py's e2e figure (5.1 MB in + 8.5 MB out per step)."""
import torch
dev = torch.device("cuda:0")
n = 64 << 20
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(reps):
        fn()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


for name, fn, nbytes in (("h2d", h2d, n), ("d2h", d2h, n), ("both", both, 2 * n)):
    ms = timed(fn)
    print("%s: %.1f GB/s (%.3f ms per 64 MiB%s)" % (name, nbytes / ms / 1e6, ms, " each way" if name == "both" else ""))
