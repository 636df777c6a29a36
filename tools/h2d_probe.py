"""How long does the upload of one frame take on this box? (pinned host -> device)
One stream: depth then colour back to back; two streams: depth and colour on different copy streams."""
import sys
import torch

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 480)
d = torch.empty(h, w, dtype=torch.float32).pin_memory()
c = torch.empty(h, w, 3, dtype=torch.uint8).pin_memory()
dd, cd = torch.empty_like(d, device="cuda"), torch.empty_like(c, device="cuda")
both = torch.empty(h * w * 7, dtype=torch.uint8).pin_memory()
bd = torch.empty_like(both, device="cuda")
s, s2 = torch.cuda.Stream(), torch.cuda.Stream()
nbytes = d.numel() * 4 + c.numel()


def timed(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(n):
        fn()
    s.wait_stream(s2)
    e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


def one_stream():
    with torch.cuda.stream(s):
        dd.copy_(d, non_blocking=True), cd.copy_(c, non_blocking=True)


def two_streams():
    with torch.cuda.stream(s):
        dd.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        cd.copy_(c, non_blocking=True)


def single_copy():
    with torch.cuda.stream(s):
        bd.copy_(both, non_blocking=True)


for name, fn in (("depth + colour on one stream", one_stream), ("depth and colour on two streams", two_streams), ("one copy of the same bytes", single_copy)):
    us = timed(fn)
    print(f"{w}x{h} {name}: {us:.1f} us per frame ({nbytes / us / 1e3:.1f} GB/s)")
