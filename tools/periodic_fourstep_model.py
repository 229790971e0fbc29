"""NumPy model of a periodic-z transform for the fast-diagonalisation solve that needs no power-of-two length
(config C2: Nz - 4 = 4092 = 62 x 66 unknowns per row).  PREPARATION for the next round -- nothing in the
product calls this yet; it exists so that the index maps and the real-embedded matrices
are settled and checked on the CPU before a kernel is written.

Today the periodic solve applies the z eigenbasis as two dense (n x n) GEMMs per solve (K = 4092).  The periodic
second-difference operator is circulant, so its eigenbasis is the DFT; with n = n1 * n2 the DFT of a real row
factors into two small dense DFTs (the "four-step" FFT), i.e. GEMMs with K = n1 and K = 2 n2 over ALL rows at
once, plus two transposing passes -- exactly the kernels the library already has (axb_dgemm) and two simple
permutation / twiddle passes:

  forward   x (Nr, n) real
    T1   Xt[row, j2, j1] = x[row, j1 n2 + j2]
    G1   A  = Xt (Nr n2, n1) @ [C1 | S1]                       -> complex A[row, j2, k1] as [re | im]
    TW   B[row, k1, j2] = A[row, j2, k1] * w^(j2 k1), w = exp(-2 pi i / n)   (transpose + twiddle)
    G2   C  = [B_re | B_im] (Nr n1, 2 n2) @ M2 (2 n2, 2 (h2+1))  -> X^[k1 + k2 n1] for k2 = 0 .. h2 = n2 / 2
  spectral layout: (Nr, n1, 2, h2+1) real -- every column is one (k1, re/im, k2), its mode number k = k1 + k2 n1;
  the tridiagonal r solve runs per column with lambda_z(k) (columns with k > n/2 are mirror modes: redundant but
  harmless; they get weight 0 on the way back)
  inverse
    G3   B' = [C_re | C_im] @ M3(k1 == 0 ? a : b)   weights c_k in {1, 2, 0} folded into the two matrices
    TW'  A'[row, j2, k1] = B'[row, k1, j2] * w^(-j2 k1)
    G4   xt = [A'_re | A'_im] (Nr n2, 2 n1) @ [[C1], [-S1']] / n   (real part only)
    T8   x[row, j1 n2 + j2] = xt[row, j2, j1]

Run:  python tools/periodic_fourstep_model.py   (checks against numpy.fft and a dense periodic solve)
"""
import numpy as np


def split(n):
    """n = n1 * n2 with n2 even and n1 + n2 minimal"""
    best = None
    for n1 in range(2, n):
        if n % n1 == 0 and (n // n1) % 2 == 0:
            n2 = n // n1
            if best is None or n1 + n2 < best[0] + best[1]:
                best = (n1, n2)
    if best is None:
        raise ValueError(f"{n} has no factorisation with an even factor")
    return best


class FourStep:
    def __init__(self, n):
        self.n = n
        self.n1, self.n2 = n1, n2 = split(n)
        self.h2 = h2 = n2 // 2
        j1, k1 = np.arange(n1)[:, None], np.arange(n1)[None, :]
        a1 = 2 * np.pi * ((j1 * k1) % n1) / n1
        self.M1 = np.hstack([np.cos(a1), -np.sin(a1)])                      # (n1, 2 n1): x real -> [re | im]
        j2, k2 = np.arange(n2)[:, None], np.arange(h2 + 1)[None, :]
        a2 = 2 * np.pi * ((j2 * k2) % n2) / n2
        Fr, Fi = np.cos(a2), -np.sin(a2)
        self.M2 = np.vstack([np.hstack([Fr, Fi]), np.hstack([-Fi, Fr])])    # (2 n2, 2 (h2+1))
        jj, kk = np.arange(n2)[:, None], np.arange(n1)[None, :]
        at = 2 * np.pi * ((jj * kk) % n) / n
        self.tw = np.cos(at) - 1j * np.sin(at)                               # w^(j2 k1), shape (n2, n1)
        # inverse over k2 with the Hermitian weights: rows k1 == 0: [1, 2, ..., 2, 1]; rows k1 != 0: [2, ..., 2, 0]
        b2 = 2 * np.pi * ((np.arange(h2 + 1)[:, None] * np.arange(n2)[None, :]) % n2) / n2
        Gr, Gi = np.cos(b2), np.sin(b2)                                      # e^{+i ...}
        wa = np.full(h2 + 1, 2.0); wa[0] = wa[-1] = 1.0
        wb = np.full(h2 + 1, 2.0); wb[-1] = 0.0
        emb = lambda w: np.vstack([np.hstack([w[:, None] * Gr, w[:, None] * Gi]),      # noqa: E731
                                   np.hstack([-w[:, None] * Gi, w[:, None] * Gr])])   # (2 (h2+1), 2 n2)
        self.M3a, self.M3b = emb(wa), emb(wb)
        c1 = 2 * np.pi * ((k1.T * j1.T) % n1) / n1                           # (k1, j1)
        self.M4 = np.vstack([np.cos(c1), -np.sin(c1)]) / n                   # (2 n1, n1): Re((a+ib) e^{+i t}) / n
        k = np.arange(n1)[:, None] + np.arange(h2 + 1)[None, :] * n1         # mode number of column (k1, k2)
        self.mode = k

    # ---- the passes, written the way the kernels would do them (real arithmetic, whole-field GEMMs) ------------
    def forward(self, x):
        nr, n, n1, n2, h2 = x.shape[0], self.n, self.n1, self.n2, self.h2
        xt = x.reshape(nr, n1, n2).transpose(0, 2, 1).reshape(nr * n2, n1)                  # T1
        a = xt @ self.M1                                                                    # G1
        a = (a[:, :n1] + 1j * a[:, n1:]).reshape(nr, n2, n1) * self.tw[None]                # TW (twiddle ...
        b = a.transpose(0, 2, 1).reshape(nr * n1, n2)                                       #     ... + transpose)
        c = np.hstack([b.real, b.imag]) @ self.M2                                           # G2
        return c.reshape(nr, n1, 2, h2 + 1)

    def inverse(self, spec):
        nr, n, n1, n2, h2 = spec.shape[0], self.n, self.n1, self.n2, self.h2
        c = spec.reshape(nr * n1, 2 * (h2 + 1))
        b = np.empty((nr * n1, 2 * n2))
        b[0::n1] = c[0::n1] @ self.M3a                                                      # G3, rows k1 == 0
        for k1 in range(1, n1):                                                             # (one strided GEMM on
            b[k1::n1] = c[k1::n1] @ self.M3b                                                #  the GPU: lda = n1 pitch)
        bc = (b[:, :n2] + 1j * b[:, n2:]).reshape(nr, n1, n2).transpose(0, 2, 1) * np.conj(self.tw)[None]   # TW'
        a = bc.reshape(nr * n2, n1)
        xt = np.hstack([a.real, a.imag]) @ self.M4                                          # G4
        return xt.reshape(nr, n2, n1).transpose(0, 2, 1).reshape(nr, n)                     # T8

    def eigenvalues(self, dx):
        """lambda_z of the periodic second-difference operator (+2, -1, -1 wrap) / dx^2 for every spectral column"""
        lam = (2.0 - 2.0 * np.cos(2 * np.pi * self.mode / self.n)) / dx / dx                # (n1, h2+1)
        return np.broadcast_to(lam[:, None, :], (self.n1, 2, self.h2 + 1))


def _check(n, nr=5, seed=0):
    rng = np.random.default_rng(seed)
    fs = FourStep(n)
    x = rng.standard_normal((nr, n))
    spec = fs.forward(x)
    ref = np.fft.fft(x, axis=1)
    got = spec[:, :, 0, :] + 1j * spec[:, :, 1, :]                                           # [row, k1, k2]
    err_f = np.max(np.abs(got - ref[:, fs.mode % n])) / np.max(np.abs(ref))
    err_b = np.max(np.abs(fs.inverse(spec) - x)) / np.max(np.abs(x))
    return fs, err_f, err_b


def _check_solve(nr, n, seed=1):
    """periodic Stokes-type solve (A_r (x) I + I (x) A_z) psi = rhs with the transform above + a dense r solve per
    column, against the dense solution"""
    rng = np.random.default_rng(seed)
    dx = 1.0 / (n + 4)
    fs = FourStep(n)
    i2 = 1 / dx / dx
    r = np.linspace(dx / 2, nr * dx - dx / 2, nr)
    Pr = (np.diag(np.full(nr, 2.0)) - np.diag(np.ones(nr - 1), 1) - np.diag(np.ones(nr - 1), -1)) * i2
    Dr = (np.diag(np.ones(nr - 1), -1) - np.diag(np.ones(nr - 1), 1)) / (2 * dx) / r[:, None]   # sub +, super -
    Pr[-1, -1] = i2
    Dr[-1, -2] = 0
    Ar = Pr - Dr                                                                            # Neumann r (reference)
    Az = (np.diag(np.full(n, 2.0)) - np.diag(np.ones(n - 1), 1) - np.diag(np.ones(n - 1), -1)) * i2
    Az[0, -1] = Az[-1, 0] = -i2
    rhs = rng.standard_normal((nr, n))
    dense = np.linalg.solve(np.kron(Ar, np.eye(n)) + np.kron(np.eye(nr), Az), rhs.reshape(-1)).reshape(nr, n)
    spec = fs.forward(rhs)
    lam = fs.eigenvalues(dx)
    out = np.empty_like(spec)
    for k1 in range(fs.n1):
        for part in range(2):
            for k2 in range(fs.h2 + 1):
                out[:, k1, part, k2] = np.linalg.solve(Ar + lam[k1, part, k2] * np.eye(nr), spec[:, k1, part, k2])
    return np.max(np.abs(fs.inverse(out) - dense)) / np.max(np.abs(dense))


if __name__ == "__main__":
    for n in (12, 60, 132, 1020, 4092):
        fs, ef, eb = _check(n)
        print(f"n = {n:5d} = {fs.n1} x {fs.n2}: forward vs numpy.fft {ef:.2e}, round trip {eb:.2e}, "
              f"spectral columns {fs.n1 * 2 * (fs.h2 + 1)} (field {n})")
    print("periodic solve vs dense (nr = 10, n = 60):", f"{_check_solve(10, 60):.2e}")
    print("periodic solve vs dense (nr = 6, n = 132):", f"{_check_solve(6, 132):.2e}")
