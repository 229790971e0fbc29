import sys
import numpy as np
sys.path.insert(0, ".")
from pyaxisymflow_b200.ops import gen_periodic_boundary_ghost_comm
from pyaxisymflow_b200.static_pde_extrapolation import StaticPDEExtrapolation
g = np.load("tests/golden/static_pde.npz")
nr, nz = g["box_phi"].shape
dx = float(g["box_dx"])
per = gen_periodic_boundary_ghost_comm(2)
eta2, phi2 = g["slab_eta0"].copy(), g["slab_phi"].copy()
s2 = StaticPDEExtrapolation(dx, nr, nz, float(g["slab_tol"]), float(g["slab_band"]), periodic=True, per_communicator_gen=per, per_communicator_eta=per)
s2.extrapolate(eta2, phi2)
ref = g["slab_eta"]
d = np.abs(eta2 - ref)
j, k = np.unravel_index(np.argmax(d), d.shape)
print("sweeps", s2.sweeps, "bounds", s2.r_start, s2.r_end, s2.z_start, s2.z_end, "max diff", d.max(), "at", j, k, eta2[j, k], ref[j, k])
print("nonzero mine/ref", np.count_nonzero(eta2), np.count_nonzero(ref), "nan", np.isnan(eta2).sum())
np.set_printoptions(linewidth=250, precision=4)
print("mine col", eta2[:, k]); print("ref col ", ref[:, k])
print("mine row", eta2[j, :12]); print("ref row ", ref[j, :12])
