"""tiny driver for ncu: build the C4 (or given) stepper and run a few steps"""
import sys
import torch
sys.path.insert(0, ".")
from pyaxisymflow_b200.timestep import RigidFlowStepper

nz = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
s = RigidFlowStepper(nz, grid_size_r=nz // 4)
s.seed_vorticity()
torch.cuda.synchronize()
s.step(steps)
torch.cuda.synchronize()
print(s.scalars())
