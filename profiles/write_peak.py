#!/usr/bin/env python
"""What a WRITE-ONLY kernel can reach on this device, next to the copy peak of MEASURED_PEAKS.json: torch kernels over 1 GiB,
best of 10 with CUDA events -- constant fill (compressible), iota (every word different), copy (read + write), and a
two-plane / four-plane iota write (several output streams at the same relative offsets, like the layout passes)."""
import json
import torch

dev = torch.device("cuda", 0)
n = 1 << 28   # int32 elements = 1 GiB
a = torch.empty(n, dtype=torch.int32, device=dev)
b = torch.empty(n, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def best(fn, nbytes, reps=10):
    t = []
    for _ in range(reps):
        flush.random_(0, 255)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        t.append(s.elapsed_time(e))
    return round(nbytes / (min(t) * 1e-3) / 1e9, 1)


out = {"bytes": n * 4}
out["fill_constant_gbs"] = best(lambda: a.fill_(7), n * 4)
out["iota_gbs"] = best(lambda: torch.arange(n, out=a, dtype=torch.int32), n * 4)
r = torch.randint(0, 1 << 30, (n,), dtype=torch.int32, device=dev)
out["copy_gbs_read_plus_write"] = best(lambda: b.copy_(r), 2 * n * 4)
q = a.view(4, n // 4)
out["iota_4_planes_gbs"] = best(lambda: torch.add(r[: n // 4].view(1, -1), torch.arange(4, device=dev, dtype=torch.int32).view(4, 1), out=q), n * 4 + n)
out["xor_inplace_gbs_read_plus_write"] = best(lambda: r.bitwise_xor_(12345), 2 * n * 4)
print(json.dumps(out))
