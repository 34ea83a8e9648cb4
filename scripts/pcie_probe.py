"""What the host link of this box gives: H2D alone, D2H alone, both at once (two streams), for the bench's byte counts.
usage: python scripts/pcie_probe.py [MB]  -> one JSON line (GB/s)"""
import json
import sys
import time

import torch

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 330
n = mb << 20
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for _ in range(2):
    run(True, True, 2)
t_up, t_down, t_both = run(True, False), run(False, True), run(True, True)
print(json.dumps({"mb_each_way": mb, "h2d_gbs": n / t_up / 1e9, "d2h_gbs": n / t_down / 1e9, "duplex_gbs_each_way": n / t_both / 1e9,
                  "ms": {"h2d": t_up * 1e3, "d2h": t_down * 1e3, "both": t_both * 1e3}}))
