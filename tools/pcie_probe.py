#!/usr/bin/env python
"""What the PCIe link of this box gives: H2D alone, D2H alone, both at once (pinned, two streams) -- the ceiling of the
end-to-end frame pipeline (bench.py e2e).  python tools/pcie_probe.py [MiB]"""
import sys, time, torch
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = mb << 20
h_up, h_dn = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_up, d_dn = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, dn, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): d_up.copy_(h_up, non_blocking=True)
        if dn:
            with torch.cuda.stream(s2): h_dn.copy_(d_dn, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
run(True, True, 2)
a, b, c = run(True, False), run(False, True), run(True, True)
print(f"{mb} MiB: H2D alone {n/a/1e9:.1f} GB/s, D2H alone {n/b/1e9:.1f} GB/s, both at once {n/c/1e9:.1f} GB/s each direction")
