"""How long does the upload of one frame take on this box? (pinned host -> device, copy stream)"""
import sys
import torch

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 480)
d = torch.empty(h, w, dtype=torch.float32).pin_memory()
c = torch.empty(h, w, 3, dtype=torch.uint8).pin_memory()
dd, cd = torch.empty_like(d, device="cuda"), torch.empty_like(c, device="cuda")
s = torch.cuda.Stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(s):
    for _ in range(20):
        dd.copy_(d, non_blocking=True), cd.copy_(c, non_blocking=True)
    e0.record(s)
    for _ in range(200):
        dd.copy_(d, non_blocking=True), cd.copy_(c, non_blocking=True)
    e1.record(s)
s.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 200
print(f"{w}x{h}: {us:.1f} us per frame upload ({(d.numel() * 4 + c.numel()) / us / 1e3:.1f} GB/s), back to back on one stream")
