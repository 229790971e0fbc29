"""Physics-level pin (the only reference-held number that exercises the whole rigid-flow loop,
ENO3 included): steady drag of a sphere at Re = 100 against the empirical curve the reference plots beside its own
results, Cd = 24/Re (1 + 0.15 Re^0.687) ~ 1.09 (examples/FlowPastSphere/post_processing.py:80-84), run at the
reference's CPU-runnable configuration C1 (128 x 256) to the reference's own convergence criterion
(examples/FlowPastSphere/flow_past_sphere.py:194-207: means over 30 steps differ by < 1e-5)."""
import pytest

pytestmark = pytest.mark.gpu


def test_sphere_drag_at_re_100_matches_the_empirical_curve():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pyaxisymflow_b200.timestep import RigidFlowStepper

    Re, sample_size, drag_diff = 100.0, 30, 1e-5
    s = RigidFlowStepper(256, grid_size_r=128, Re=Re, use_graph=True)
    t_end = 300 * s.r_sph / s.U_0                      # nondim_T = 300 (flow_past_sphere.py:47-48)
    k_cd = 2 * 2 * 3.141592653589793 * s.dx * s.dx * s.brink_lam / (3.141592653589793 * s.r_sph ** 2)
    acc = torch.zeros((), dtype=torch.float64, device="cuda")
    previous, current, converged, it = 0.0, 0.0, False, 0
    while it < 400000:
        acc.zero_()
        for _ in range(sample_size):
            s.step(1)
            acc.add_(s.state[7])                       # drag sum of the step just taken (device resident)
        it += sample_size
        vals = torch.stack([acc, s.state[0]]).cpu()
        current = k_cd * float(vals[0]) / sample_size
        if previous != 0.0 and abs(current - previous) < drag_diff and float(vals[1]) > s.T_ramp:
            converged = True
            break
        previous = current
        if float(vals[1]) >= t_end:
            break
    cd_emp = 24 / Re * (1 + 0.15 * Re ** 0.687)
    print(f"Cd = {current:.4f} after {it} steps (t = {float(vals[1]):.3f}, converged = {converged}); "
          f"empirical {cd_emp:.4f}")
    assert converged or float(vals[1]) >= t_end
    # 128 x 256 with the outer wall at 5 sphere radii (blockage) and 26 cells per radius: measured 1.198 (+9.8 %)
    assert abs(current - cd_emp) <= 0.12 * cd_emp, (current, cd_emp)
