"""Host-side set-up of the fast-diagonalisation solve (pyaxisymflow_b200/fd.py) checked on
the CPU: the factor set -- LAPACK and closed-form/symmetrised bases alike -- must reproduce
the reference solver outputs stored in tests/golden/fast_diag.npz."""
import numpy as np
import pytest

from conftest import assert_close, golden
from pyaxisymflow_b200 import fd

BCS = ("homogenous_neumann_along_z_and_r", "homogenous_neumann_along_r_and_periodic_along_z",
       "homogenous_dirichlet_along_r_and_periodic_along_z")


@pytest.mark.parametrize("basis", ["lapack", "analytic"])
def test_stokes_factors_match_reference(basis):
    g = golden("fast_diag")
    rhs, dx = g["rhs"], float(g["dx"])
    nr, nz = rhs.shape
    for bc in BCS[:2]:
        f = fd.build_factors("stokes", bc, nr, nz, dx, basis)
        assert_close(fd.apply_factors_host(f, rhs), g["stokes_" + bc], 1e-10, f"{basis} {bc}")
    f = fd.build_factors("stokes", BCS[1], nr, nz - 4, dx, basis)
    assert_close(fd.apply_factors_host(f, rhs[:, 2:-2]), g["stokes_periodic_inner"], 1e-10, "inner grid")


@pytest.mark.parametrize("basis", ["lapack", "analytic"])
def test_potential_and_implicit_factors_match_reference(basis):
    g = golden("fast_diag")
    rhs, dx = g["rhs"], float(g["dx"])
    nr, nz = rhs.shape
    # The all-Neumann potential operator is singular (constants are in its null space): the
    # reference's 1/lambda blows the null mode up to ~1e11 and what is left of the solution is
    # noise at the 1e-2 level (DESIGN.md "Reference defects").  Only the LAPACK path, which
    # repeats the reference's exact computation, can be compared; the closed-form basis is
    # checked through the implicit-diffusion flavour, which shares all of its code.
    if basis == "lapack":
        f = fd.build_factors("potential", BCS[0], nr, nz, dx, basis)
        assert_close(fd.apply_factors_host(f, rhs), g["potential"], 1e-6, "potential")
    f = fd.build_factors("implicit_diffusion", None, nr, nz, dx, basis, nu_dt=float(g["nu"]) * float(g["time_step"]))
    assert_close(fd.apply_factors_host(f, rhs), g["implicit_diffusion"], 1e-10, "implicit diffusion")


def test_dirichlet_r_even_nr_is_a_defective_operator():
    """Reference defect: with Dirichlet-r and even Nr the radial operator has the
    eigenvalue 2/dx^2 twice with one eigenvector; la.eig returns a basis with cond ~ 1e14 and the
    reference's own solution misses the equation by ~1e-2.  The LAPACK path mimics it to that
    accuracy only; the analytic path refuses."""
    g = golden("fast_diag")
    rhs, dx = g["rhs"], float(g["dx"])
    nr, nz = rhs.shape
    f = fd.build_factors("stokes", BCS[2], nr, nz, dx, "lapack")
    assert_close(fd.apply_factors_host(f, rhs), g["stokes_" + BCS[2]], 1e-2, "dirichlet lapack")
    with pytest.raises(ValueError, match="defective"):
        fd.build_factors("stokes", BCS[2], nr, nz, dx, "analytic")
    # the direct r solve needs no eigenvectors and solves the equation the reference misses
    t = fd.build_factors("stokes", BCS[2], nr, nz, dx, "analytic", r_method="tridiagonal")
    psi = fd.apply_factors_host(t, rhs)
    sub, diag, sup, r = fd.radial_tridiagonal("stokes", BCS[2], nr, dx)
    res = (fd.dense_from_tridiagonal(sub, diag, sup) @ psi + psi @ fd.dense_axial(*fd.axial_kind("stokes", BCS[2]), nz, dx).T
           - r[:, None] * rhs)
    assert np.max(np.abs(res)) <= 1e-9 * np.max(np.abs(r[:, None] * rhs))
    # odd Nr is fine
    f = fd.build_factors("stokes", BCS[2], nr - 1, nz, dx, "analytic")
    h = fd.build_factors("stokes", BCS[2], nr - 1, nz, dx, "lapack")
    assert_close(fd.apply_factors_host(f, rhs[:-1]), fd.apply_factors_host(h, rhs[:-1]), 1e-10, "odd Nr")


@pytest.mark.parametrize("nr,nz", [(96, 200), (128, 255)])
def test_analytic_basis_solves_the_operator(nr, nz):
    """residual check at sizes without a golden: A_r psi + psi A_z^T = r o rhs"""
    dx = 1.0 / nz
    rng = np.random.default_rng(1)
    rhs = rng.standard_normal((nr, nz))
    for bc in BCS[:2]:
        f = fd.build_factors("stokes", bc, nr, nz, dx, "analytic")
        psi = fd.apply_factors_host(f, rhs)
        sub, diag, sup, r = fd.radial_tridiagonal("stokes", bc, nr, dx)
        Ar = fd.dense_from_tridiagonal(sub, diag, sup)
        Az = fd.dense_axial(*fd.axial_kind("stokes", bc), nz, dx)
        res = Ar @ psi + psi @ Az.T - r[:, None] * rhs
        assert np.max(np.abs(res)) <= 1e-9 * np.max(np.abs(r[:, None] * rhs))
        g = fd.build_factors("stokes", bc, nr, nz, dx, "lapack")
        assert_close(psi, fd.apply_factors_host(g, rhs), 1e-10, "analytic vs lapack " + bc)


def test_parity_split_z_transform_equals_dense():
    """The folded [E_L | O_L | ... | O_1] leaf products reproduce the dense N x N transforms."""
    rng = np.random.default_rng(5)
    for kind, bc, nr, nz, kw in (("stokes", BCS[0], 40, 2048, {}), ("implicit_diffusion", None, 24, 1024, {"nu_dt": 1e-7})):
        dx = 1.0 / nz
        rhs = rng.standard_normal((nr, nz))
        dense = fd.build_factors(kind, bc, nr, nz, dx, "analytic", split=0, **kw)
        ref = fd.apply_factors_host(dense, rhs)
        assert dense["zsplit"] is None
        for split in (1, 2, "auto"):
            f = fd.build_factors(kind, bc, nr, nz, dx, "analytic", split=split, **kw)
            zs = f["zsplit"]
            assert zs is not None and sum(zs["leaf_n"]) == nz and zs["leaf_off"][0] == 0
            assert_close(fd.apply_factors_host(f, rhs), ref, 1e-13, f"{kind} split={split}")
            assert fd.solve_flops(nr, nz, f) < 0.6 * fd.solve_flops(nr, nz, dense)
    # fold / unfold are inverse up to the factor 2 carried by the orthonormal bases
    x = rng.standard_normal((3, 64))
    y = fd.fold_host(fd.fold_host(x, 64), 64, inverse=True)
    assert np.allclose(y, 2 * x)
    # no split for periodic z or odd sizes
    assert fd.build_factors("stokes", BCS[1], 24, 1024, 1 / 1024, "analytic")["zsplit"] is None
    assert fd.split_levels("neumann", 1022) == 0 and fd.split_levels("neumann", 16384) == 3


@pytest.mark.parametrize("n", [64, 128, 512, 2048])
def test_dct_kernel_model_and_tables(n):
    """tools/dct_model.py restates the FFT-based cosine transforms of zfft.cu step by step in NumPy,
    reading the same packed twiddle table: both directions against the dense cosine matrix"""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location(
        "dct_model", os.path.join(os.path.dirname(__file__), "..", "tools", "dct_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    V = fd.axial_natural_block("neumann", n, n, np.arange(n)).numpy()
    ck = np.full(n, np.sqrt(2.0 / n))
    ck[0] = np.sqrt(1.0 / n)
    assert_close(m.dct2(x), (x @ V) / ck, 1e-13, "DCT-II model")
    assert_close(m.dct3(x), V @ (x / ck), 1e-13, "DCT-III model")


def test_fft_factors_solve_the_operator():
    nr, nz = 40, 128
    dx = 1.0 / nz
    rng = np.random.default_rng(5)
    rhs = rng.standard_normal((nr, nz))
    ref = fd.apply_factors_host(fd.build_factors("stokes", BCS[0], nr, nz, dx, "analytic", split=0), rhs)
    f = fd.build_factors("stokes", BCS[0], nr, nz, dx, "analytic", r_method="tridiagonal", z_method="fft")
    assert f["zfft"] is not None and f["Rz"] is None and f["zsplit"] is None
    assert_close(fd.apply_factors_host(f, rhs), ref, 1e-12, "fft factor set")
    assert fd.build_factors("stokes", BCS[0], nr, 96, dx, "analytic", r_method="tridiagonal",
                            z_method="auto")["zfft"] is None
    with pytest.raises(ValueError):
        fd.build_factors("stokes", BCS[0], nr, nz, dx, "analytic", z_method="fft")      # needs the direct r solve
    # periodic z now has its own FFT path (real FFT rows, csrc/pfft.cu); sizes with a large prime factor do not
    assert fd.build_factors("stokes", BCS[1], nr, nz, dx, "analytic", r_method="tridiagonal",
                            z_method="fft")["zfft"]["family"] == "periodic"
    with pytest.raises(ValueError):
        fd.build_factors("stokes", BCS[1], nr, 2 * 73, dx, "analytic", r_method="tridiagonal", z_method="fft")


def test_periodic_fourstep_model():
    """tools/periodic_fourstep_model.py (preparation for an FFT-class periodic-z solve at Nz - 4 = 4092 = 62 x 66, see
    DESIGN.md section 10): the two small real-embedded DFT GEMMs + twiddle / transpose passes reproduce numpy.fft,
    round-trip, and solve the periodic operator like a dense solve."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import periodic_fourstep_model as pm

    for n in (12, 132, 4092):
        fs, ef, eb = pm._check(n, nr=3)
        assert ef <= 1e-13 and eb <= 1e-13, (n, ef, eb)
    assert (fs.n1, fs.n2) == (62, 66)
    assert pm._check_solve(6, 60) <= 1e-12


def test_periodic_rfft_model():
    """tools/rfft_model.py restates the index arithmetic of csrc/pfft.cu (Stockham passes with generic radices from
    one table of M-th roots, real-FFT untangling, half-complex layout); it must agree with numpy.fft."""
    import sys

    sys.path.insert(0, "tools")
    import rfft_model as rm

    rng = np.random.default_rng(0)
    for n in (4092, 252, 60, 8, 1020, 52, 4096, 16380 // 4 * 2):
        f = rm.factorize(n // 2)
        assert f == fd.rfft_factors(n // 2)
        tm, tn = rm.tables(n)
        tab = fd.rfft_tables(n)
        assert np.allclose(tab[:n // 2, 0] + 1j * tab[:n // 2, 1], tm, atol=1e-15)
        assert np.allclose(tab[n // 2:, 0] + 1j * tab[n // 2:, 1], tn, atol=1e-15)
        x = rng.standard_normal(n)
        h = rm.rfft_row(x, f, tm, tn)
        ref = np.fft.rfft(x)
        m = n // 2
        assert np.abs(h[:m + 1] - ref.real).max() <= 1e-13 * np.abs(ref).max()
        assert np.abs(h[m + 1:] - ref.imag[1:m]).max() <= 1e-13 * np.abs(ref).max()
        assert np.abs(rm.irfft_row(h, f, tm, tn) - x).max() <= 1e-13
        assert np.array_equal(rm.mode_of_column(n)[:m + 1], np.arange(m + 1))
        # the passes as the kernel runs them (twiddles applied at the previous pass's store, odd radices with paired
        # inputs / outputs): same FFT
        z = rng.standard_normal(m) + 1j * rng.standard_normal(m)
        want = np.fft.fft(z)
        assert np.abs(rm.cfft_passes_paired(z, f, tm) - want).max() <= 1e-13 * np.abs(want).max()
    assert fd.rfft_factors(2044 // 2) is None and not fd.rfft_supported(2044)      # 511 = 7 * 73
    assert fd.rfft_supported(4092) and fd.rfft_spectral_width(4092) == 4096


def test_periodic_fft_factor_set_matches_reference():
    """periodic z through real FFT rows + the tridiagonal r solve: same solution as the reference's eigen-decomposition
    (golden outputs of FastDiagonalisationStokesSolver with the periodic bc, full and inner grid)"""
    g = golden("fast_diag")
    rhs, dx = g["rhs"], float(g["dx"])
    nr, nz = rhs.shape
    f = fd.build_factors("stokes", BCS[1], nr, nz, dx, "analytic", r_method="tridiagonal", z_method="fft")
    assert f["zfft"]["family"] == "periodic" and f["lam_z"].shape[0] == f["zfft"]["nz_spec"] == 64
    assert_close(fd.apply_factors_host(f, rhs), g["stokes_" + BCS[1]], 1e-10, "periodic rfft path")
    f = fd.build_factors("stokes", BCS[1], nr, nz - 4, dx, "analytic", r_method="tridiagonal", z_method="fft")
    assert_close(fd.apply_factors_host(f, rhs[:, 2:-2]), g["stokes_periodic_inner"], 1e-10, "inner grid")
