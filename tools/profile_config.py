"""tiny driver for ncu: a few steps of config C3 (soft sphere, 2048 x 8192) or C5 (one particle case, 1024 x 2048)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyaxisymflow_b200.timestep import ParticleFlowStepper, RigidFlowStepper, SoftSphereStepper  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if which == "c2":
    s = RigidFlowStepper(4096, grid_size_r=1024, periodic=True, r_sph=0.075, Z_cm=0.85)
    s.seed_vorticity()
    s.t = 0.0
elif which == "c4":
    s = RigidFlowStepper(16384, grid_size_r=4096)
    s.seed_vorticity()
    s.t = 0.0
elif which == "c3":
    s = SoftSphereStepper(8192, grid_size_r=2048, Z_cm=0.47, reinit_levelset=True)
elif which == "c3d":                                  # device-resident soft-sphere step (eager launches)
    s = SoftSphereStepper(8192, grid_size_r=2048, device_scalars=True)
elif which == "c5b":                                  # the batched, device-resident 8-member ensemble (eager launches)
    from pyaxisymflow_b200.timestep import ParticleEnsemble
    freqs = [4.0, 8.0, 12.0, 16.0, 20.0, 24.0, 28.0, 32.0]
    s = ParticleEnsemble.batched_ensemble([(f, 0.01) for f in freqs], 2048, 1024, use_graph=False)
    s.t = 0.0
else:
    s = ParticleFlowStepper(2048, grid_size_r=1024)
torch.cuda.synchronize()
s.step(steps)
torch.cuda.synchronize()
print(which, "steps", steps, "t", s.t)
