"""torchrun --nproc-per-node P tools/check_slab.py [nz] [steps] [rows|rows-eager|rows-periodic|z] : slab stepper vs the single-GPU stepper
(rows = the r-slab stepper of rowslab.py, the default; z = the z-slab stepper of slab.py)"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from pyaxisymflow_b200.rowslab import RowSlabRigidFlowStepper  # noqa: E402
from pyaxisymflow_b200.slab import SlabRigidFlowStepper  # noqa: E402
from pyaxisymflow_b200.timestep import RigidFlowStepper  # noqa: E402

nz = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
mode = sys.argv[3] if len(sys.argv) > 3 else "rows"
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
nr = nz // 4
pk = {}
if mode == "rows-periodic":  # config C2's loop: periodic z with 2 ghost columns a side (inner width = the given nz)
    nz += 4
    pk = {"periodic": True, "r_sph": 0.075, "Z_cm": 0.85}
if mode == "z":
    s = SlabRigidFlowStepper(nz, grid_size_r=nr)
else:       # rows: flag-synchronised, replayed as a CUDA graph; rows-eager: the same launched kernel by kernel
    s = RowSlabRigidFlowStepper(nz, grid_size_r=nr, use_graph=(mode != "rows-eager"), **pk)
s.seed_vorticity()
s.step(steps)
w = s.gather_vorticity()
sc = s.scalars()
if rank == 0:
    ref = RigidFlowStepper(nz, grid_size_r=nr, **pk)
    ref.seed_vorticity()
    ref.step(steps)
    torch.cuda.synchronize()
    err = ((w - ref.vorticity).abs().max() / ref.vorticity.abs().max()).item()
    rs = ref.scalars()
    print(s.solve_kernel_note())
    print(f"{mode}-slab x{world} vs single GPU at {nr}x{nz}, {steps} steps: rel Linf {err:.3e}; "
          f"t {sc['t']:.12e} vs {rs['t']:.12e}; Cd {sc['Cd']:.10e} vs {rs['Cd']:.10e}")
    assert err < 1e-10, err
    assert abs(sc["t"] - rs["t"]) <= 1e-14 * abs(rs["t"])
if hasattr(s, "close"):
    s.close()
dist.barrier()
dist.destroy_process_group()
