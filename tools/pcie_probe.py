#!/usr/bin/env python
"""Host link probe for the e2e analysis: pinned-memory H2D / D2H bandwidth of this box, one direction at a
time and both together (torch copies on two streams, CUDA-event timing)."""
import json

import torch


def main():
    n = 1 << 29  # 4 GiB of fp64
    h_in = torch.empty(n, dtype=torch.float64, pin_memory=True).fill_(1.0)
    h_out = torch.empty(n, dtype=torch.float64, pin_memory=True)
    d_a = torch.empty(n, dtype=torch.float64, device="cuda")
    d_b = torch.ones(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    gb = n * 8 / 1e9

    def timed(fn):
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            torch.cuda.synchronize()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        return best

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    t1, t2, t3 = timed(h2d), timed(d2h), timed(both)
    print(json.dumps({"bytes_each": n * 8, "h2d_GBps": gb / t1, "d2h_GBps": gb / t2, "both_directions_GBps_each": gb / t3,
                      "h2d_ms_for_4.36GB_triangle": 4.36 / (gb / t1) * 1e3, "gpu": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
