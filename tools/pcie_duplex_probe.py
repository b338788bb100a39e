#!/usr/bin/env python
"""Host <-> device copy bandwidth alone and in both directions at once (two streams, pinned memory): tells whether the
box's PCIe path is full duplex for the frame scheduler's upload-ahead / download overlap."""
import json
import time

import torch

n = 512 << 20
h_up, h_dn = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_up, d_dn = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, dn, reps=8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d_up.copy_(h_up, non_blocking=True)
        if dn:
            with torch.cuda.stream(s2):
                h_dn.copy_(d_dn, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * n * (int(up) + int(dn)) / dt / 1e9


run(True, True, 2)
print(json.dumps({"h2d_GBs": run(True, False), "d2h_GBs": run(False, True), "both_total_GBs": run(True, True)}))

# does a device -> host copy issued AFTER a queue of host -> device copies wait for them?  (it must not, for upload-ahead to pay)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
with torch.cuda.stream(s1):
    s1.wait_event(e0)
    for _ in range(10):
        d_up.copy_(h_up, non_blocking=True)
    e1.record(s1)
with torch.cuda.stream(s2):
    s2.wait_event(e0)
    h_dn.copy_(d_dn, non_blocking=True)
    e2.record(s2)
torch.cuda.synchronize()
print(json.dumps({"ten_h2d_ms": e0.elapsed_time(e1), "one_d2h_issued_after_them_ms": e0.elapsed_time(e2)}))
