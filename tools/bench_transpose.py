"""torchrun --nproc-per-node P tools/bench_transpose.py : time the peer-memory transposes, their barrier, and NCCL"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from pyaxisymflow_b200.slab import PeerTranspose, SlabComm, SlabLayout  # noqa: E402

nr, nz = 4096, 16384
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
L = SlabLayout(nr, nz, world, rank)
pt = PeerTranspose(L)
comm = SlabComm(L)
slab = torch.randn((nr, L.nzl), dtype=torch.float64, device="cuda")
rows = torch.randn((L.nrl, nz), dtype=torch.float64, device="cuda")


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


res = {
    "put slab->rows + barrier": timed(lambda: pt.slab_to_rows(slab)),
    "put rows->slab + barrier": timed(lambda: pt.rows_to_slab(rows)),
    "barrier only": timed(lambda: pt.h_rows.barrier(channel=0)),
    "nccl slab->rows": timed(lambda: comm.slab_to_rows(slab, rows)),
    "nccl rows->slab": timed(lambda: comm.rows_to_slab(rows, slab)),
}
if rank == 0:
    mb = nr * nz * 8 / world / 1e6
    for k, v in res.items():
        print(f"{k:28s} {v:8.3f} ms   ({mb:.0f} MB per rank, {mb * (world - 1) / world:.0f} MB leave the GPU)")
dist.barrier()
dist.destroy_process_group()
