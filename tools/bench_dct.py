"""tools/bench_dct.py [nz] [rows] : device time of axb_dct2_rows / axb_dct3_rows (CUDA events, 20 launches each)"""
import sys

import torch

sys.path.insert(0, ".")
from pyaxisymflow_b200 import _lib, fd  # noqa: E402
from pyaxisymflow_b200.device import ptr, stream_ptr  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
tabs = torch.from_numpy(fd.dct_tables(n)).cuda()
a = torch.randn((rows, n), dtype=torch.float64, device="cuda")
b = torch.zeros_like(a)
for name, inv in (("dct2", 0), ("dct3", 1)):
    def run():
        if inv:
            _lib.call("axb_dct3_rows", rows, n, ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(tabs), stream_ptr())
        else:
            _lib.call("axb_dct2_rows", rows, n, ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(tabs), 1.0 / n, 2.0 / n,
                      stream_ptr())
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gbs = 16.0 * rows * n / (ms * 1e-3) / 1e9
    print(f"{name} {rows}x{n}: {ms:.4f} ms  {gbs:.0f} GB/s  ({gbs / 6543.1:.3f} of 6543 GB/s)")
