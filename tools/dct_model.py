"""NumPy model of the in-shared-memory DCT-II / DCT-III kernels (index maths check)."""
import numpy as np


def radix_plan(M):
    m = int(np.log2(M))
    assert 1 << m == M
    plan = [16] * (m // 4)
    if m % 4:
        plan.append(1 << (m % 4))
    return plan


def fft_stockham(z, tabM):
    """forward complex FFT of length M, in-place Stockham passes exactly as the kernel does them"""
    M = z.size
    s = z.copy()
    Ns = 1
    for R in radix_plan(M):
        j = np.arange(M // R)
        v = np.stack([s[j + t * (M // R)] for t in range(R)])          # v[t, j]
        L = Ns * R
        jm = j % Ns
        for t in range(R):
            v[t] *= tabM[(jm * t * (M // L)) % M]
        # R-point DFT
        W = np.exp(-2j * np.pi * np.outer(np.arange(R), np.arange(R)) / R)
        o = W @ v
        base = (j // Ns) * L + jm
        out = np.empty_like(s)
        for t in range(R):
            out[base + t * Ns] = o[t]
        s = out
        Ns = L
    return s


def tables(N):
    """the packed table the kernels read (pyaxisymflow_b200.fd.dct_tables), split into its three parts"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from pyaxisymflow_b200.fd import dct_tables

    M = N // 2
    t = dct_tables(N)
    t = t[:, 0] + 1j * t[:, 1]
    return t[:M], t[M:2 * M + 1], t[2 * M + 1:]


def dct2(x):
    N = x.size
    M = N // 2
    tabM, tabN, tab4N = tables(N)
    z = np.empty(M, complex)
    n = np.arange(M // 2)
    z[n] = x[4 * n] + 1j * x[4 * n + 2]
    z[M - 1 - n] = x[4 * n + 3] + 1j * x[4 * n + 1]
    Z = fft_stockham(z, tabM)
    X = np.empty(N)
    X[0] = Z[0].real + Z[0].imag
    X[M] = (Z[0].real - Z[0].imag) * np.sqrt(0.5)
    k = np.arange(1, M // 2 + 1)
    Zk, Zm = Z[k], Z[M - k]
    E = (Zk + np.conj(Zm)) / 2
    D = (Zk - np.conj(Zm)) / 2
    P = tabN[k] * D
    Vk = (E.real + P.imag) + 1j * (E.imag - P.real)
    Vm = (E.real - P.imag) + 1j * (-E.imag - P.real)
    a = tab4N[k] * Vk
    b = tab4N[M - k] * Vm
    X[k], X[N - k] = a.real, -a.imag
    X[M - k], X[M + k] = b.real, -b.imag
    return X


def dct3(a):
    """y[j] = sum_k a[k] cos(pi k (2j+1) / (2N))"""
    N = a.size
    M = N // 2
    tabM, tabN, tab4N = tables(N)
    Z = np.empty(M, complex)
    Z[0] = (a[0] + a[M] * np.sqrt(0.5)) + 1j * (a[0] - a[M] * np.sqrt(0.5))
    k = np.arange(1, M // 2 + 1)
    Bk = np.conj(tab4N[k]) * (a[k] - 1j * a[N - k]) / 2
    Bm = np.conj(tab4N[M - k]) * (a[M - k] - 1j * a[M + k]) / 2
    S = Bk + np.conj(Bm)
    Dd = Bk - np.conj(Bm)
    Q = np.conj(tabN[k]) * Dd
    Z[k] = (S.real - Q.imag) + 1j * (S.imag + Q.real)
    Z[M - k] = (S.real + Q.imag) + 1j * (-S.imag + Q.real)
    sw = Z.imag + 1j * Z.real
    o = fft_stockham(sw, tabM)
    z = o.imag + 1j * o.real
    y = np.empty(N)
    n = np.arange(M // 2)
    y[4 * n], y[4 * n + 2] = z[n].real, z[n].imag
    y[4 * n + 3], y[4 * n + 1] = z[M - 1 - n].real, z[M - 1 - n].imag
    return y


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for N in (64, 128, 256, 512, 2048, 16384):
        x = rng.standard_normal(N)
        j = np.arange(N)
        if N <= 2048:
            C = np.cos(np.pi * np.outer(np.arange(N), 2 * j + 1) / (2 * N))
            print(N, np.abs(dct2(x) - C @ x).max(), np.abs(dct3(x) - C.T @ x).max())
        import scipy.fft as sf
        print(N, "scipy", np.abs(dct2(x) - sf.dct(x, 2) / 2).max(), np.abs(dct3(dct2(x)) * 2 / N - x - (dct2(x)[0]) / N).max())
