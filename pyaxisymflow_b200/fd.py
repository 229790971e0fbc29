"""Fast-diagonalisation solvers (fusion group G-FD).

Host side of SURVEY.md row a16 / a16': the three solver classes of the reference
(kernels/FastDiagonalisationStokesSolver.py, FastDiagonalisationPotentialSolver.py,
implicit_diffusion_solver.py) with the same constructor / ``solve`` / ``step`` signatures.
Set-up (building the two 1-D finite-difference operators and diagonalising them) runs once
on the host; every ``solve`` is four FP64 tensor-core GEMMs on the GPU (``axb_fd_solve``).

Two ways of obtaining the eigenbases, both yielding the same solution operator
``psi = V f(Lambda) V^-1 rhs`` (basis independent):

``basis="lapack"``
    exactly what the reference does: dense matrices, ``numpy.linalg.eig`` sorted by descending
    eigenvalue, ``numpy.linalg.inv`` (FastDiagonalisationStokesSolver.py:109-128).  O(N^3) --
    fine up to N ~ 2-4 k, hopeless at Nz = 16384 (~20-40 min).
``basis="analytic"``
    z: the operators are Toeplitz with Neumann / periodic / plain ends, whose orthonormal
    eigenvectors are known in closed form (DCT-II cosines, Fourier pairs, DST-I sines), so
    ``V^-1 = V^T``; the table is generated with exact integer argument reduction.
    r: the tridiagonal ``A_r`` is made symmetric by a diagonal similarity (row 0 of the
    Stokes operator decouples because ``A_r[0,1] = 0``) and handed to
    ``scipy.linalg.eigh_tridiagonal``.  O(N^2).
``basis="auto"`` picks "lapack" below 1536 points per axis, "analytic" above.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import AxbFdPlan
from .device import Stage, ptr, stream_ptr

_BC_NEUMANN = "homogenous_neumann_along_z_and_r"
_BC_NEU_PER = "homogenous_neumann_along_r_and_periodic_along_z"
_BC_DIR_PER = "homogenous_dirichlet_along_r_and_periodic_along_z"


# --------------------------------------------------------------------------------------
# operator entries
# --------------------------------------------------------------------------------------
def radial_tridiagonal(kind, bc_type, nr, dx):
    """(sub, diag, sup) of A_r for the three solver flavours."""
    r = np.linspace(dx / 2, nr * dx - dx / 2, nr)
    i2, h = 1 / dx / dx, 1 / 2 / dx
    if kind == "stokes":
        # P_r - D_r, P = i2*tridiag(-1,2,-1), D[i,i-1] = +h/r_i, D[i,i+1] = -h/r_i
        sub = -i2 - h / r[1:]
        sup = -i2 + h / r[:-1]
        diag = np.full(nr, 2 * i2)
        if bc_type in (_BC_NEUMANN, _BC_NEU_PER):
            diag[-1] = i2
            sub[-1] = -i2       # D_r[-1,-2] = 0
    elif kind == "potential":
        sub = i2 - h / r[1:]
        sup = i2 + h / r[:-1]
        diag = np.full(nr, -2 * i2)
        if bc_type == _BC_NEUMANN:
            diag[-1] = -i2
            sub[-1] = i2
    elif kind == "implicit_diffusion":
        sub = i2 - h / r[1:]
        sup = i2 + h / r[:-1]
        diag = -2 * i2 - 1.0 / r ** 2
    else:
        raise ValueError(kind)
    return sub, diag, sup, r


def axial_kind(kind, bc_type):
    """(family, sign) of A_z: family in {"neumann", "periodic", "dirichlet"}"""
    sign = 1.0 if kind == "stokes" else -1.0
    if kind == "implicit_diffusion":
        return "dirichlet", sign
    if bc_type == _BC_NEUMANN:
        return "neumann", sign
    if kind == "stokes" and bc_type in (_BC_NEU_PER, _BC_DIR_PER):
        return "periodic", sign
    return "dirichlet", sign   # unknown bc string: the reference applies no modification


def dense_from_tridiagonal(sub, diag, sup):
    n = diag.size
    m = np.zeros((n, n))
    i = np.arange(n)
    m[i, i] = diag
    m[i[1:], i[:-1]] = sub
    m[i[:-1], i[1:]] = sup
    return m


def dense_axial(family, sign, nz, dx):
    i2 = 1 / dx / dx
    m = dense_from_tridiagonal(np.full(nz - 1, -i2), np.full(nz, 2 * i2), np.full(nz - 1, -i2))
    if family == "neumann":
        m[0, 0] = m[-1, -1] = i2
    elif family == "periodic":
        m[0, -1] = m[0, 1]
        m[-1, 0] = m[-1, -2]
    return sign * m


# --------------------------------------------------------------------------------------
# eigenbases
# --------------------------------------------------------------------------------------
def _eig_sorted(m):
    lam, V = np.linalg.eig(m)
    o = lam.argsort()[::-1]
    return lam[o], V[:, o]


class _ComplexBasis(Exception):
    """la.eig returned complex conjugate pairs for a (numerically) repeated real eigenvalue"""


def _real_or_raise(lam, V):
    if np.iscomplexobj(lam) or np.iscomplexobj(V):
        if np.max(np.abs(np.imag(lam))) > 0 or np.max(np.abs(np.imag(V))) > 0:
            raise _ComplexBasis()
        lam, V = np.real(lam), np.real(V)
    return lam, V


def radial_basis_lapack(sub, diag, sup):
    lam, V = _real_or_raise(*_eig_sorted(dense_from_tridiagonal(sub, diag, sup)))
    return lam, V, np.linalg.inv(V)


def _sym_tridiag_eig(sub, diag, sup):
    """eigen-decomposition of a tridiagonal with sub*sup > 0 via diagonal symmetrisation.
    Returns lam, V, V^-1 with A = V diag(lam) V^-1."""
    from scipy.linalg import eigh_tridiagonal

    prod = sub * sup
    if np.any(prod <= 0):
        raise ValueError("tridiagonal operator is not symmetrisable")
    off = np.sign(sup) * np.sqrt(prod)
    # d_{i+1}/d_i = sqrt(sub_i / sup_i); accumulate in log space to avoid overflow
    logd = np.concatenate([[0.0], np.cumsum(0.5 * (np.log(np.abs(sub)) - np.log(np.abs(sup))))])
    logd -= logd.mean()
    d = np.exp(logd)
    lam, Q = eigh_tridiagonal(diag, off)
    return lam, d[:, None] * Q, Q.T / d[None, :]


def radial_basis_analytic(sub, diag, sup):
    n = diag.size
    if abs(sup[0]) <= 1e-12 * abs(diag[0]):
        # row 0 decouples: A = [[d0, 0], [a e1, T']]
        lam_t, Vt, Vti = _sym_tridiag_eig(sub[1:], diag[1:], sup[1:])
        d0 = diag[0]
        # (d0 I - T') w = a e1, solved in T' eigen-coordinates
        e1 = np.zeros(n - 1)
        e1[0] = sub[0]
        gap = np.min(np.abs(d0 - lam_t))
        if gap <= 1e-9 * abs(d0):
            # e.g. the Dirichlet-r Stokes operator with even Nr: the zero-diagonal odd block
            # gives T' the eigenvalue 2/dx^2 = A_r[0,0] exactly, A_r has a Jordan block and is
            # NOT diagonalisable (the reference's la.eig basis then has cond ~ 1e14 and its
            # solve carries O(1e-2) residuals, see DESIGN.md "Reference defects").
            raise ValueError("A_r is defective (repeated eigenvalue %.6g): fast diagonalisation is "
                             "ill-posed for this grid/bc; basis='lapack' mimics the reference" % d0)
        w = Vt @ ((Vti @ e1) / (d0 - lam_t))
        lam = np.concatenate([[d0], lam_t])
        V = np.zeros((n, n))
        V[0, 0] = 1.0
        V[1:, 0] = w
        V[1:, 1:] = Vt
        Vi = np.zeros((n, n))
        Vi[0, 0] = 1.0
        Vi[1:, 0] = -(Vti @ w)
        Vi[1:, 1:] = Vti
    else:
        lam, V, Vi = _sym_tridiag_eig(sub, diag, sup)
    o = lam.argsort()[::-1]
    return lam[o], V[:, o], Vi[o, :]


def axial_basis_lapack(family, sign, nz, dx):
    lam, V = _real_or_raise(*_eig_sorted(dense_axial(family, sign, nz, dx)))
    return lam, V, np.linalg.inv(V)


def axial_basis_analytic(family, sign, nz, dx, device="cpu"):
    """orthonormal closed-form eigenvectors as a torch tensor V[j, k] on ``device``; V^-1 = V^T."""
    i2 = 1 / dx / dx
    j = torch.arange(nz, dtype=torch.int64, device=device)
    k = torch.arange(nz, dtype=torch.int64, device=device)
    if family == "neumann":
        # v_k[j] = cos(pi k (2j+1) / (2N)), lam_k = (2 - 2 cos(pi k / N)) / dx^2
        m = (k[None, :] * (2 * j[:, None] + 1)) % (4 * nz)
        V = torch.cos(m.to(torch.float64) * (np.pi / (2 * nz)))
        V *= np.sqrt(2.0 / nz)
        V[:, 0] = np.sqrt(1.0 / nz)
        lam = (2 - 2 * np.cos(np.pi * np.arange(nz) / nz)) * i2
    elif family == "periodic":
        # columns: 1, cos(2 pi q j/N), sin(2 pi q j/N) ..., (-1)^j
        V = torch.empty((nz, nz), dtype=torch.float64, device=device)
        lam = np.empty(nz)
        V[:, 0] = np.sqrt(1.0 / nz)
        lam[0] = 0.0
        col = 1
        q = 1
        while col < nz:
            m = ((q * j) % nz).to(torch.float64) * (2 * np.pi / nz)
            lq = (2 - 2 * np.cos(2 * np.pi * q / nz)) * i2
            if 2 * q == nz:
                V[:, col] = torch.cos(m) * np.sqrt(1.0 / nz)
                lam[col] = lq
                col += 1
            else:
                V[:, col] = torch.cos(m) * np.sqrt(2.0 / nz)
                V[:, col + 1] = torch.sin(m) * np.sqrt(2.0 / nz)
                lam[col] = lam[col + 1] = lq
                col += 2
            q += 1
    else:
        # plain Dirichlet-type Toeplitz: v_k[j] = sin(pi (k+1)(j+1)/(N+1))
        m = ((k[None, :] + 1) * (j[:, None] + 1)) % (2 * (nz + 1))
        V = torch.sin(m.to(torch.float64) * (np.pi / (nz + 1))) * np.sqrt(2.0 / (nz + 1))
        lam = (2 - 2 * np.cos(np.pi * (np.arange(nz) + 1) / (nz + 1))) * i2
    lam = sign * lam
    o = np.argsort(lam)[::-1].copy()
    V = V[:, torch.as_tensor(o, device=device)].contiguous()
    return lam[o], V


def axial_natural_block(family, nz, rows, modes, device="cpu"):
    """rows x len(modes) block V[j, k] (j < rows, k in modes, natural mode order) of the orthonormal
    closed-form z eigenvectors, generated with exact integer argument reduction."""
    j = torch.arange(rows, dtype=torch.int64, device=device)
    k = torch.as_tensor(modes, dtype=torch.int64, device=device)
    if family == "neumann":
        m = (k[None, :] * (2 * j[:, None] + 1)) % (4 * nz)
        V = torch.cos(m.to(torch.float64) * (np.pi / (2 * nz))) * np.sqrt(2.0 / nz)
        V[:, k == 0] = np.sqrt(1.0 / nz)
        return V
    if family == "dirichlet":
        m = ((k[None, :] + 1) * (j[:, None] + 1)) % (2 * (nz + 1))
        return torch.sin(m.to(torch.float64) * (np.pi / (nz + 1))) * np.sqrt(2.0 / (nz + 1))
    raise ValueError(family)


def axial_natural_eigenvalues(family, sign, nz, dx):
    i2 = 1 / dx / dx
    k = np.arange(nz)
    if family == "neumann":
        return sign * (2 - 2 * np.cos(np.pi * k / nz)) * i2
    return sign * (2 - 2 * np.cos(np.pi * (k + 1) / (nz + 1))) * i2


def split_levels(family, nz, requested="auto", min_leaf=256):
    """number of parity-split levels of the z transform (0 = dense N x N products)"""
    if family not in ("neumann", "dirichlet"):
        return 0
    cap = 3 if family == "neumann" else 1      # only the DCT-II even branch is self-similar
    if requested != "auto":
        cap = min(cap, int(requested))
        min_leaf = 4                            # explicit request: split as asked (tests use small grids)
    L = 0
    while L < cap and nz % (2 ** (L + 1) * 2) == 0 and nz // 2 ** (L + 1) >= min_leaf:
        L += 1
    return L


def axial_split_plan(family, sign, nz, dx, levels, device="cpu"):
    """Leaves of the parity-split z transform: column layout [E_L | O_L | ... | O_1].
    Returns dict(leaf_n, leaf_off, fold_len, fwd, bwd, lam_z (folded order), modes)."""
    lam_nat = axial_natural_eigenvalues(family, sign, nz, dx)
    L = levels
    nL = nz // 2 ** L
    leaves = [("E", L, np.arange(0, nz, 2 ** L), nL)]
    for lev in range(L, 0, -1):
        leaves.append(("O", lev, np.arange(2 ** (lev - 1), nz, 2 ** lev), nz // 2 ** lev))
    off, plan = 0, {"leaf_n": [], "leaf_off": [], "fwd": [], "bwd": [], "modes": []}
    lam = np.empty(nz)
    for _, _, modes, n in leaves:
        assert modes.size == n
        F = axial_natural_block(family, nz, n, modes, device=device)
        plan["leaf_n"].append(n)
        plan["leaf_off"].append(off)
        plan["fwd"].append(F.contiguous())
        plan["bwd"].append(F.t().contiguous())
        plan["modes"].append(modes)
        lam[off:off + n] = lam_nat[modes]
        off += n
    plan["fold_len"] = [nz // 2 ** lev for lev in range(L)]
    plan["lam_z"] = lam
    return plan


def rfft_factors(m, max_radix=64):
    """radices of the half-length complex FFT of csrc/pfft.cu (4s, a 2, then odd primes); None if a prime factor
    exceeds max_radix"""
    out, twos = [], 0
    while m % 2 == 0:
        m //= 2
        twos += 1
    out += [4] * (twos // 2) + [2] * (twos % 2)
    p = 3
    while p <= max_radix and m > 1:
        while m % p == 0:
            out.append(p)
            m //= p
        p += 2
    return out if m == 1 else None


def rfft_supported(nz):
    """periodic z: even nz whose half length has small prime factors only and fits shared memory (csrc/pfft.cu:
    two row buffers + the R x R weight tables of all passes)"""
    if nz < 4 or nz % 2:
        return False
    f = rfft_factors(nz // 2)
    return f is not None and (2 * (nz // 2) + sum(r * r for r in f)) * 16 <= 200 * 1024


def fft_eligible(family, nz):
    """the shared-memory FFT transforms cover the Neumann-z family at nz = 2^p, 64 <= nz <= 16384 (cosine
    transforms, csrc/zfft.cu) and the periodic family at even nz with small prime factors (real FFT, csrc/pfft.cu)"""
    if family == "periodic":
        return rfft_supported(nz)
    return family == "neumann" and 64 <= nz <= 16384 and (nz & (nz - 1)) == 0


def rfft_tables(nz):
    """tables of axb_rfft_rows / axb_irfft_rows as one (nz + 1, 2) float64 array:
    [exp(-2 pi i k / M), k < M | exp(-2 pi i k / nz), k <= M], M = nz / 2"""
    M = nz // 2
    tab = np.concatenate([np.exp(-2j * np.pi * np.arange(M) / M), np.exp(-2j * np.pi * np.arange(M + 1) / nz)])
    tab[0], tab[M], tab[M + M] = 1, 1, -1
    return np.ascontiguousarray(np.stack([tab.real, tab.imag], axis=1))


def rfft_spectral_width(nz):
    """pitch of the half-complex spectral rows: a multiple of 16 so that the factored TMA sweeps apply"""
    return (nz + 15) // 16 * 16


def rfft_column_eigenvalues(sign, nz, dx):
    """lam_z of every column of the half-complex layout [Re X_0 .. Re X_M | Im X_1 .. Im X_{M-1} | padding]:
    (2 - 2 cos(2 pi m / nz)) / dx^2 of the column's Fourier mode m; padding columns (zero data) take mode 1's value"""
    M = nz // 2
    modes = np.concatenate([np.arange(M + 1), np.arange(1, M)])
    lam = np.empty(rfft_spectral_width(nz))
    lam[:nz] = sign * (2 - 2 * np.cos(2 * np.pi * modes / nz)) / dx / dx
    lam[nz:] = sign * (2 - 2 * np.cos(2 * np.pi / nz)) / dx / dx
    return lam


def dct_tables(nz):
    """twiddle tables of axb_dct2_rows / axb_dct3_rows as one (3 nz / 2 + 2, 2) float64 array"""
    M = nz // 2
    km = np.arange(M)
    k = np.arange(M + 1)
    tab = np.concatenate([np.exp(-2j * np.pi * km / M), np.exp(-2j * np.pi * k / nz),
                          np.exp(-1j * np.pi * k / (2 * nz))])
    # exact values where cos / sin are known, so the quarter points carry no rounding
    tab[M // 4 * np.arange(4)] = [1, -1j, -1, 1j]
    tab[M + M // 2] = -1j
    return np.ascontiguousarray(np.stack([tab.real, tab.imag], axis=1))


def fold_host(x, n, inverse=False):
    """NumPy restatement of axb_fd_fold (CPU tests of the set-up only)"""
    h = n // 2
    y = x.copy()
    if not inverse:
        a, b = x[:, :h], x[:, n - 1:h - 1:-1] if h > 0 else x[:, :0]
        y[:, :h], y[:, h:n] = a + b, a - b
    else:
        e, o = x[:, :h], x[:, h:n]
        y[:, :h] = e + o
        y[:, n - 1:h - 1:-1] = e - o
    return y


def thomas_host(x, sub, diag, sup, lam, scale, c0, c1):
    """NumPy restatement of axb_tridiag_solve_columns (all columns at once; CPU tests only)"""
    nr = x.shape[0]
    cp = np.zeros_like(x)
    d = np.array(x, dtype=np.float64)
    cprev = np.zeros(x.shape[1])
    dprev = np.zeros(x.shape[1])
    for m in range(nr):
        b = c0 + c1 * (diag[m] + lam)
        a = c1 * sub[m - 1] if m > 0 else 0.0
        cu = c1 * sup[m] if m < nr - 1 else 0.0
        rhs = d[m] * (scale[m] if scale is not None else 1.0)
        den = b - a * cprev
        cprev = cu / den
        dprev = (rhs - a * dprev) / den
        cp[m], d[m] = cprev, dprev
    xn = np.zeros(x.shape[1])
    for m in range(nr - 1, -1, -1):
        xn = d[m] - cp[m] * xn
        d[m] = xn
    return d


def build_factors(kind, bc_type, nr, nz, dx, basis="auto", nu_dt=None, device="cpu", split="auto",
                  r_method="eigen", z_method="gemm"):
    """Factor set of the solve  sol = Lrb (((Lr rhs) Rz) o 1/(c0 + c1 (lam_z (+) lam_r))) Rzb  as
    float64 torch tensors on ``device`` (see include/axisym_b200.h, axb_fd_plan_t)."""
    if basis == "auto":
        basis = "lapack" if max(nr, nz) < 1536 else "analytic"
    sub, diag, sup, r = radial_tridiagonal(kind, bc_type, nr, dx)
    family, sign = axial_kind(kind, bc_type)
    zsplit = None
    if basis == "lapack":
        try:
            lam_r, Vr, Vri = radial_basis_lapack(sub, diag, sup)
            lam_z, Vz, Vzi = axial_basis_lapack(family, sign, nz, dx)
            Rz = torch.from_numpy(np.ascontiguousarray(Vzi.T)).to(device)
            Rzb = torch.from_numpy(np.ascontiguousarray(Vz.T)).to(device)
        except _ComplexBasis:
            # degenerate periodic pairs came back as complex conjugates (the reference then carries
            # complex arrays and drops the imaginary part on assignment); the solution operator is
            # basis independent, so use the real closed-form basis instead
            basis = "analytic"
    if basis == "lapack":
        pass
    elif basis == "analytic":
        if r_method == "tridiagonal":
            lam_r = Vr = Vri = None       # the r direction is solved directly, no eigen-decomposition
        else:
            lam_r, Vr, Vri = radial_basis_analytic(sub, diag, sup)
        if z_method == "auto":
            z_method = "fft" if (r_method == "tridiagonal" and fft_eligible(family, nz)) else "gemm"
        levels = split_levels(family, nz, split)
        if z_method == "fft":
            if r_method != "tridiagonal" or not fft_eligible(family, nz):
                raise ValueError("z_method='fft' needs r_method='tridiagonal' and either the Neumann-z family with "
                                 "nz = 2^p, 64 <= nz <= 16384, or periodic z with an even nz of small prime factors")
            if family == "periodic":
                lam_z, Rz, Rzb = rfft_column_eigenvalues(sign, nz, dx), None, None
            else:
                lam_z, Rz, Rzb = axial_natural_eigenvalues(family, sign, nz, dx), None, None
        elif levels > 0:
            zsplit = axial_split_plan(family, sign, nz, dx, levels, device=device)
            lam_z, Rz, Rzb = zsplit["lam_z"], None, None
        else:
            lam_z, Vz_t = axial_basis_analytic(family, sign, nz, dx, device=device)
            Rz = Vz_t                       # V^-T = V for an orthogonal basis
            Rzb = Vz_t.t().contiguous()
    else:
        raise ValueError(f"unknown basis {basis!r}")
    def dev(a):
        return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(device)

    if r_method == "tridiagonal" and basis != "analytic":
        raise ValueError("r_method='tridiagonal' needs basis='analytic'")
    if z_method == "auto":
        z_method = "gemm"
    if z_method not in ("gemm", "fft") or (z_method == "fft" and basis != "analytic"):
        raise ValueError(f"z_method {z_method!r} is not available with basis {basis!r}")
    zfft = None
    if z_method == "fft":
        zfft = {"tables": dev(rfft_tables(nz) if family == "periodic" else dct_tables(nz)), "family": family,
                "sign": sign, "dx": dx, "nz_spec": rfft_spectral_width(nz) if family == "periodic" else nz}
    tri = None
    if r_method == "tridiagonal":
        tri = {"sub": dev(sub), "diag": dev(diag), "sup": dev(sup), "scale": dev(r) if kind == "stokes" else None}
        Lr = None
    else:
        Lr = Vri * r[None, :] if kind == "stokes" else Vri   # fold  r o rhs  into the first factor
    f = {
        "Lr": dev(Lr), "Lrb": dev(Vr), "Rz": Rz, "Rzb": Rzb, "lam_r": dev(lam_r), "lam_z": dev(lam_z),
        "c0": 0.0, "c1": 1.0, "basis": basis, "zsplit": zsplit, "tri": tri, "r_method": r_method,
        "zfft": zfft, "z_method": z_method,
    }
    if kind == "implicit_diffusion":
        f["c0"], f["c1"] = 1.0, -float(nu_dt)
    return f


def apply_factors_host(f, rhs):
    """NumPy evaluation of the factorised solve (used only by the CPU tests of the set-up)."""
    g = {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in f.items()}
    zs = f.get("zsplit")
    if f.get("tri") is not None:
        tri = {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in f["tri"].items()}
        t = np.array(rhs, dtype=np.float64)
        if f.get("zfft") is not None and f["zfft"]["family"] == "periodic":
            # real FFT rows in the half-complex layout (host check only; numpy.fft stands in for csrc/pfft.cu)
            n = t.shape[1]
            M, ns = n // 2, f["zfft"]["nz_spec"]
            X = np.fft.rfft(t, axis=1)
            spec = np.zeros((t.shape[0], ns))
            spec[:, :M + 1], spec[:, M + 1:n] = X.real, X.imag[:, 1:M]
            spec = thomas_host(spec, tri["sub"], tri["diag"], tri["sup"], g["lam_z"], tri["scale"], g["c0"], g["c1"])
            Y = np.zeros((t.shape[0], M + 1), complex)
            Y.real, Y.imag[:, 1:M] = spec[:, :M + 1], spec[:, M + 1:n]
            return np.fft.irfft(Y, n=n, axis=1)
        if f.get("zfft") is not None:
            # DCT-II / DCT-III written as products with the orthonormal cosine basis (host check only)
            zf = f["zfft"]
            V = axial_natural_block(zf["family"], t.shape[1], t.shape[1], np.arange(t.shape[1])).numpy()
            spec = thomas_host(t @ V, tri["sub"], tri["diag"], tri["sup"], g["lam_z"], tri["scale"], g["c0"], g["c1"])
            return spec @ V.T
        if zs is None:
            spec = t @ g["Rz"]
        else:
            for n in zs["fold_len"]:
                t = fold_host(t, n)
            spec = np.empty_like(t)
            for n, off, F in zip(zs["leaf_n"], zs["leaf_off"], zs["fwd"]):
                spec[:, off:off + n] = t[:, off:off + n] @ F.cpu().numpy()
        spec = thomas_host(spec, tri["sub"], tri["diag"], tri["sup"], g["lam_z"], tri["scale"], g["c0"], g["c1"])
        if zs is None:
            return spec @ g["Rzb"]
        for n, off, B in zip(zs["leaf_n"], zs["leaf_off"], zs["bwd"]):
            t[:, off:off + n] = spec[:, off:off + n] @ B.cpu().numpy()
        for n in reversed(zs["fold_len"]):
            t = fold_host(t, n, inverse=True)
        return t
    scale = 1.0 / (g["c0"] + g["c1"] * (g["lam_z"][None, :] + g["lam_r"][:, None]))
    if zs is None:
        spec = ((g["Lr"] @ rhs) @ g["Rz"]) * scale
        return g["Lrb"] @ (spec @ g["Rzb"])
    t = g["Lr"] @ rhs
    for n in zs["fold_len"]:
        t = fold_host(t, n)
    spec = np.empty_like(t)
    for n, off, F in zip(zs["leaf_n"], zs["leaf_off"], zs["fwd"]):
        spec[:, off:off + n] = t[:, off:off + n] @ F.cpu().numpy()
    spec *= scale
    for n, off, B in zip(zs["leaf_n"], zs["leaf_off"], zs["bwd"]):
        t[:, off:off + n] = spec[:, off:off + n] @ B.cpu().numpy()
    for n in reversed(zs["fold_len"]):
        t = fold_host(t, n, inverse=True)
    return g["Lrb"] @ t


# --------------------------------------------------------------------------------------
# device plan + the three reference classes
# --------------------------------------------------------------------------------------
def factor_pivots(nr, nz, f):
    """reciprocal LU pivots of the nz tridiagonal r systems (axb_tridiag_factor_columns), kept with the
    factor set on the GPU; needs nz % 16 == 0 (otherwise the solve recomputes the pivots every time)"""
    tri = f.get("tri")
    if f.get("zfft") is not None:
        nz = f["zfft"]["nz_spec"]
    if tri is None or nz % 16 or not f["lam_z"].is_cuda:
        return
    inv = torch.empty((nr, nz), dtype=torch.float64, device=f["lam_z"].device)
    rc = torch.empty((nr, 4), dtype=torch.float64, device=f["lam_z"].device)
    _lib.call("axb_tridiag_factor_columns", nr, nz, ptr(tri["sub"]), ptr(tri["diag"]), ptr(tri["sup"]),
              ptr(f["lam_z"]), ptr(tri["scale"]), float(f["c0"]), float(f["c1"]), ptr(inv), ptr(rc), stream_ptr())
    tri["inv"], tri["row_coef"] = inv, rc


def make_plan(nr, nz, f, work):
    """axb_fd_plan_t over a factor set living on the GPU"""
    p = AxbFdPlan()
    p.nr, p.nz = nr, nz
    opt = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    p.Lr, p.Lrb = opt(f["Lr"]), opt(f["Lrb"])
    p.Rz, p.Rzb = opt(f["Rz"]), opt(f["Rzb"])
    p.lam_r, p.lam_z = opt(f["lam_r"]), f["lam_z"].data_ptr()
    tri = f.get("tri")
    p.r_tridiagonal = 0
    if tri is not None:
        p.r_tridiagonal = 1
        p.r_sub, p.r_diag, p.r_sup = tri["sub"].data_ptr(), tri["diag"].data_ptr(), tri["sup"].data_ptr()
        p.r_scale = opt(tri["scale"])
    p.c0, p.c1, p.work = f["c0"], f["c1"], work.data_ptr()
    p.r_inv_pivots = p.r_row_coef = None
    if tri is not None and tri.get("inv") is not None:
        p.r_inv_pivots, p.r_row_coef = tri["inv"].data_ptr(), tri["row_coef"].data_ptr()
    p.z_fft, p.nz_spec = 0, nz
    if f.get("zfft") is not None:
        p.z_fft, p.z_tables = (2 if f["zfft"]["family"] == "periodic" else 1), f["zfft"]["tables"].data_ptr()
        p.nz_spec = f["zfft"]["nz_spec"]
    zs = f.get("zsplit")
    p.n_leaves = p.n_folds = 0
    if zs is not None:
        p.n_leaves, p.n_folds = len(zs["leaf_n"]), len(zs["fold_len"])
        for i, (n, off, F, B) in enumerate(zip(zs["leaf_n"], zs["leaf_off"], zs["fwd"], zs["bwd"])):
            p.leaf_n[i], p.leaf_off[i] = n, off
            p.leaf_fwd[i], p.leaf_bwd[i] = F.data_ptr(), B.data_ptr()
        for i, n in enumerate(zs["fold_len"]):
            p.fold_len[i] = n
    return p


def solve_hbm_bytes(nr, nz, f):
    """algorithmic HBM bytes of one solve on the DCT + factored-sweep path ("Solve: HBM
    accounting"): DCT-II reads the right-hand side and writes the spectrum (16 B/pt), each sweep reads the
    field and the pivots and writes the field (24 B/pt, twice), DCT-III reads and writes (16 B/pt)."""
    if f.get("zfft") is None:
        return None
    tri = f.get("tri") or {}
    return (16 + 16 + (48 if tri.get("inv") is not None else 40)) * float(nr) * nz


def solve_flops(nr, nz, f):
    """floating-point operations one solve executes with this factor set"""
    zs = f.get("zsplit")
    if f.get("zfft") is not None:
        return 2 * (2.5 * nr * nz * np.log2(nz)) + 10.0 * nr * nz     # two real FFTs per row + Thomas
    z = 2.0 * nr * nz * nz if zs is None else sum(2.0 * nr * n * n for n in zs["leaf_n"])
    if f.get("tri") is not None:
        return 2 * z + 10.0 * nr * nz          # GEMM flops + the Thomas sweeps
    return 2 * (2.0 * nr * nr * nz) + 2 * z


class _FdBase:
    kind = None

    def _setup(self, grid_size_r, grid_size_z, dx, real_dtype, bc_type, basis, nu_dt=None, split="auto",
               r_method="auto", z_method="auto"):
        if real_dtype != np.float64:
            raise TypeError("libaxisym_b200 computes in float64 only")
        if not torch.cuda.is_available():
            raise _lib.AxbError("the fast-diagonalisation solve runs on the GPU only (no CPU fallback)")
        self.dx, self.grid_size_r, self.grid_size_z = dx, grid_size_r, grid_size_z
        self.real_dtype, self.bc_type = real_dtype, bc_type
        self.radial_coord = np.linspace(dx / 2, grid_size_r * dx - dx / 2, grid_size_r).reshape(grid_size_r, 1)
        # "auto": grids too large for la.eig (basis left to "auto", >= 1536 points a side) take the direct
        # r solve and, where the z family allows, the FFT z transforms; everything else follows the
        # reference's eigen-decomposition.
        if r_method == "auto":
            r_method = "tridiagonal" if (basis == "auto" and max(grid_size_r, grid_size_z) >= 1536) else "eigen"
        if r_method == "tridiagonal" and basis == "auto":
            basis = "analytic"
        if r_method != "tridiagonal" and z_method == "auto":
            z_method = "gemm"
        self.factors = build_factors(self.kind, bc_type, grid_size_r, grid_size_z, dx, basis, nu_dt, device="cuda",
                                     split=split, r_method=r_method, z_method=z_method)
        self.basis = self.factors["basis"]
        # spectral buffer of the reference (FastDiagonalisationStokesSolver.py:38-39) x 2
        f = self.factors
        nzs = f["zfft"]["nz_spec"] if f.get("zfft") is not None else grid_size_z
        self.work = torch.empty(2 * grid_size_r * nzs, dtype=torch.float64, device="cuda")
        factor_pivots(grid_size_r, grid_size_z, f)
        self.plan = make_plan(grid_size_r, grid_size_z, f, self.work)

    def fork(self):
        """same factor set, private work fields and plan: for solves that run concurrently on different streams
        (members of an ensemble)"""
        import copy

        o = copy.copy(self)
        o.work = torch.empty_like(self.work)
        o.plan = make_plan(self.grid_size_r, self.grid_size_z, self.factors, o.work)
        return o

    def _solve(self, solution_field, rhs_field):
        st = Stage()
        sol = st.dev(solution_field, out=True)
        rhs = st.dev(rhs_field)
        shape = (self.grid_size_r, self.grid_size_z)
        if tuple(sol.shape) != shape or tuple(rhs.shape) != shape:
            raise ValueError(f"shapes {tuple(sol.shape)}, {tuple(rhs.shape)} do not match the solver grid {shape}")
        wb = None
        if sol.stride(1) != 1:
            wb, sol = sol, sol.contiguous()
        if rhs.stride(1) != 1:
            rhs = rhs.contiguous()
        _lib.call("axb_fd_solve", ctypes.byref(self.plan), ptr(sol), sol.stride(0), ptr(rhs), rhs.stride(0),
                  stream_ptr())
        if wb is not None:
            wb.copy_(sol)
        st.finish()


class FastDiagonalisationStokesSolver(_FdBase):
    """kernels/FastDiagonalisationStokesSolver.py:6-156"""
    kind = "stokes"

    def __init__(self, grid_size_r, grid_size_z, dx, real_dtype=np.float64, bc_type=_BC_NEUMANN, basis="auto",
                 split="auto", r_method="auto", z_method="auto"):
        self._setup(grid_size_r, grid_size_z, dx, real_dtype, bc_type, basis, split=split, r_method=r_method,
                    z_method=z_method)

    def solve(self, solution_field, rhs_field):
        self._solve(solution_field, rhs_field)

    def flops(self):
        return solve_flops(self.grid_size_r, self.grid_size_z, self.factors)

    def hbm_bytes(self):
        """algorithmic HBM bytes per solve when the solve is bandwidth bound (DCT path), else None"""
        return solve_hbm_bytes(self.grid_size_r, self.grid_size_z, self.factors)

    def kernel_note(self):
        if self.factors.get("zfft") is not None and self.factors["zfft"]["family"] == "periodic":
            return ("fast-diagonalisation solve = k_rfft_rows + k_tri_sweep_ws<fwd> + k_tri_sweep_ws<bwd> + k_irfft_rows "
                    "(4 launches: shared-memory mixed-radix real FFTs along periodic z, radices "
                    f"{rfft_factors(self.grid_size_z // 2)}, factored tridiagonal sweeps along r; 80 algorithmic "
                    "B/grid-pt)")
        if self.factors.get("zfft") is not None:
            nz = self.grid_size_z
            dct = ("k_dct_rows_w<II> + {} + k_dct_rows_w<III>" if nz == 16384 else
                   "k_dct_rows_rr<II> + {} + k_dct_rows_rr<III>" if nz in (64, 1024) else "k_dct2_rows + {} + k_dct3_rows")
            return ("fast-diagonalisation solve = " + dct.format("k_tri_sweep_ws<fwd> + k_tri_sweep_ws<bwd>") +
                    " (4 launches: on-chip FFT cosine transforms along z, factored tridiagonal sweeps along r through a "
                    "TMA ring; 80 algorithmic B/grid-pt)")
        zs = self.factors.get("zsplit")
        tri = self.factors.get("tri") is not None
        rpart = "batched tridiagonal r solve" if tri else "2 r-transforms"
        if zs is None:
            return f"k_dgemm_tma ({2 if tri else 4} launches per solve: {rpart} + 2 dense z-transforms)"
        n = len(zs["leaf_n"])
        return (f"k_dgemm_tma ({(0 if tri else 2) + 2 * n} launches per solve: {rpart} + 2x{n} parity-split z leaves "
                f"{zs['leaf_n']}; flops = executed, {self.flops() / (4.0 * self.grid_size_r * self.grid_size_z * (self.grid_size_r + self.grid_size_z)):.3f} of the dense count)")


class FastDiagonalisationPotentialSolver(_FdBase):
    """kernels/FastDiagonalisationPotentialSolver.py:6-145"""
    kind = "potential"

    def __init__(self, grid_size_r, grid_size_z, dx, real_dtype=np.float64, bc_type=_BC_NEUMANN, basis="auto"):
        # the all-Neumann potential operator is singular; keep the reference's eigen-decomposition for it
        self._setup(grid_size_r, grid_size_z, dx, real_dtype, bc_type, basis, r_method="eigen", z_method="gemm")

    def solve(self, solution_field, rhs_field):
        self._solve(solution_field, rhs_field)


class ImplicitEulerDiffusionStepper(_FdBase):
    """kernels/implicit_diffusion_solver.py:6-143"""
    kind = "implicit_diffusion"

    def __init__(self, time_step, kinematic_viscosity, grid_size_r, grid_size_z, dx, real_dtype=np.float64,
                 basis="auto", split="auto", r_method="auto", z_method="auto"):
        self.time_step = time_step
        self.nu_times_dt = self.time_step * kinematic_viscosity
        self._setup(grid_size_r, grid_size_z, dx, real_dtype, None, basis, nu_dt=self.nu_times_dt, split=split,
                    r_method=r_method, z_method=z_method)

    def step(self, vorticity_field, dt):
        if dt != self.time_step:
            raise ValueError(
                "dt should be constant throughout the simulation if using "
                "implicit diffusion! Please use the value of dt being used "
                "to initialize the implicit diffusion stepper")
        # in place like the reference: the first GEMM reads the field, the last one rewrites it
        self._solve(vorticity_field, vorticity_field)
