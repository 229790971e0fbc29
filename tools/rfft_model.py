"""NumPy model of csrc/pfft.cu: real FFT of an even-length row through a half-length mixed-radix complex FFT
(Stockham auto-sort passes, generic O(R^2) butterflies from one table of M-th roots), the half-complex spectral
layout [Re_0 .. Re_M | Im_1 .. Im_{M-1}] and its inverse.  The index arithmetic is the kernel's, so the CPU test
of this model (tests/test_fd_factors_cpu.py::test_periodic_rfft_model) pins it against numpy.fft."""
import numpy as np


def factorize(m, max_radix=64):
    """radices of the passes (4 instead of 2*2); None when a prime factor exceeds max_radix"""
    f, p = [], 2
    while m > 1:
        while m % p == 0:
            f.append(p)
            m //= p
        p += 1
        if p > max_radix and m > 1:
            return None
    out = []
    twos = f.count(2)
    out += [4] * (twos // 2) + [2] * (twos % 2)
    out += [x for x in f if x != 2]
    return out                                  # 4s, a 2, then the odd primes ascending: the kernel's order


def tables(n):
    m = n // 2
    k = np.arange(m)
    tab_m = np.exp(-2j * np.pi * k / m)           # M-th roots
    tab_n = np.exp(-2j * np.pi * np.arange(m + 1) / n)
    return tab_m, tab_n


def cfft_passes(z, factors, tab_m):
    """forward complex FFT of length M by the kernel's passes; z is a 1-D complex array"""
    m = z.size
    ns = 1
    cur = z.copy()
    for r in factors:
        L = m // r
        q = np.arange(m)
        jm = q % ns
        u = (q // ns) % r
        j = (q // (ns * r)) * ns + jm
        e = (jm * (m // (ns * r)) + u * (m // r)) % m
        acc = np.zeros(m, complex)
        idx = np.zeros(m, int)
        for t in range(r):
            acc += cur[j + t * L] * tab_m[idx]
            idx = idx + e
            idx = np.where(idx >= m, idx - m, idx)
        cur = acc
        ns *= r
    return cur


def rfft_row(x, factors, tab_m, tab_n):
    """x (N reals) -> half-complex layout of length N, unnormalised"""
    n = x.size
    m = n // 2
    z = x[0::2] + 1j * x[1::2]
    Z = cfft_passes(z, factors, tab_m)
    k = np.arange(m + 1)
    zk = Z[k % m]
    zm = np.conj(Z[(m - k) % m])
    E = 0.5 * (zk + zm)
    O = -0.5j * (zk - zm)
    X = E + tab_n[k] * O
    out = np.empty(n)
    out[:m + 1] = X.real
    out[m + 1:] = X.imag[1:m]
    return out


def irfft_row(h, factors, tab_m, tab_n):
    """inverse of rfft_row including the 1/N normalisation"""
    n = h.size
    m = n // 2
    X = np.zeros(m + 1, complex)
    X.real = h[:m + 1]
    X.imag[1:m] = h[m + 1:]
    k = np.arange(m)
    xk, xm = X[k], np.conj(X[m - k])
    E = 0.5 * (xk + xm)
    O = 0.5 * (xk - xm) * np.conj(tab_n[k])
    Z = E + 1j * O
    # inverse FFT through the forward passes: ifft(Z) = conj(fft(conj(Z))) / M
    z = np.conj(cfft_passes(np.conj(Z), factors, tab_m)) / m
    x = np.empty(n)
    x[0::2], x[1::2] = z.real, z.imag
    return x


def mode_of_column(n):
    """Fourier mode of every column of the half-complex layout"""
    m = n // 2
    return np.concatenate([np.arange(m + 1), np.arange(1, m)])


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n in (4092, 252, 60, 8, 1020, 4096):
        f = factorize(n // 2)
        tm, tn = tables(n)
        x = rng.standard_normal(n)
        h = rfft_row(x, f, tm, tn)
        ref = np.fft.rfft(x)
        m = n // 2
        err = max(np.abs(h[:m + 1] - ref.real).max(), np.abs(h[m + 1:] - ref.imag[1:m]).max()) / np.abs(ref).max()
        back = np.abs(irfft_row(h, f, tm, tn) - x).max()
        print(n, f, err, back)
