"""Raw pinned-memory PCIe bandwidth on this box: H2D alone, D2H alone, both at once (the e2e floor)."""
import torch
dev = torch.device("cuda:0")
nbytes = 134217728
h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
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
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    ms = timed(fn)
    print(f"{name}: {ms:.3f} ms per 134 MB  ->  {nbytes / ms / 1e6:.1f} GB/s per direction")
