"""[torchrun --nproc-per-node P] tools/rowslab_phases.py [nz] [nr] : per-phase ms of the r-slab step on every rank"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from pyaxisymflow_b200.rowslab import RowSlabRigidFlowStepper  # noqa: E402

nz = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nr = int(sys.argv[2]) if len(sys.argv) > 2 else nz // 4
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
multi = "RANK" in os.environ
if multi:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if multi else 0
s = RowSlabRigidFlowStepper(nz, grid_size_r=nr, use_graph=bool(os.environ.get('AXB_GRAPH')))
s.seed_vorticity()
s.step(3)
torch.cuda.synchronize()
ph = s.phase_times(5)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if multi:
    dist.barrier()
torch.cuda.synchronize()
e0.record()
s.step(10)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"rank": rank, "world": s.L.world, "peer_halos": s.peer_halos, "ms_per_step": e0.elapsed_time(e1) / 10,
                  "sum_phases": round(sum(ph.values()), 4), "phases": ph}), flush=True)
if multi:
    dist.barrier()
    dist.destroy_process_group()
