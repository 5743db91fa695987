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
# hand-written write-only kernel with hashed data (libtsim's measurement aid): 1, 3 and 4 output planes
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trafficsimulation_b200 import _lib
lib = _lib.load()
big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for streams in (1, 3, 4):
    per = (1 << 30) // streams // 16 * 16
    for blocks in (148 * 8, 148 * 32):
        out[f"hashed_write_{streams}_planes_{blocks}_ctas_gbs"] = best(
            lambda: _lib.check(lib.tsim_debug_write_probe(C.c_void_p(big.data_ptr()), C.c_longlong(per), streams, blocks, st)), per * streams)
print(json.dumps(out))
