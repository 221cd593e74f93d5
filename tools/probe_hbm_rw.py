"""Write-only / read-only / copy bandwidth of this GPU (context for the rooflines: MEASURED_PEAKS.json's HBM figure is a
copy, i.e. half reads and half writes; the build kernel is 94 % writes, the lookup 80 % reads)."""
import json

import torch

n = 1 << 30                                        # 4 GiB of fp32
a = torch.empty(n, device="cuda")
b = torch.empty(n, device="cuda")


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


out = {}
out["memset_write_GBs"] = 4 * n / timed(lambda: a.zero_()) / 1e9
out["fill_kernel_write_GBs"] = 4 * n / timed(lambda: a.fill_(1.5)) / 1e9
out["copy_GBs_read_plus_write"] = 8 * n / timed(lambda: b.copy_(a)) / 1e9
out["sum_read_GBs"] = 4 * n / timed(lambda: a.sum()) / 1e9
out["max_read_GBs"] = 4 * n / timed(lambda: a.max()) / 1e9
print(json.dumps(out))
