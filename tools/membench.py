#!/usr/bin/env python3
"""HBM write/copy ceilings on this GPU with plain torch kernels (context for the roofline:
the turn kernel's traffic is 95% writes)."""
import json

import torch

dev = torch.device("cuda:0")
n = 512 * 1024 * 1024  # floats: 2 GiB
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


res = {}
ms = timeit(lambda: a.fill_(1.0))
res["fill_2GiB_GBs"] = n * 4 / ms / 1e6
ms = timeit(lambda: a.zero_())
res["memset_2GiB_GBs"] = n * 4 / ms / 1e6
ms = timeit(lambda: b.copy_(a))
res["copy_2GiB_GBs_read_plus_write"] = 2 * n * 4 / ms / 1e6
ms = timeit(lambda: a.sum())
res["read_sum_2GiB_GBs"] = n * 4 / ms / 1e6
print(json.dumps(res))
