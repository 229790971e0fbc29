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


def cfft_passes_paired(z, factors, tab_m):
    """the passes as csrc/pfft.cu runs them since round 2 (same result as cfft_passes):
      * the pass twiddle W_M^(jm t M/(ns R)) of input (j, t) is applied when the PREVIOUS pass stores that element
        (every element is read by exactly one (j, t) of the next pass), so a butterfly reads plain values;
      * an odd radix R pairs the inputs t, R-t and the outputs u, R-u:
            S_t = x_t + x_(R-t),  D_t = x_t - x_(R-t),  A_u = sum_t S_t cos(2 pi u t / R),  B_u = sum_t D_t sin(..)
            X_u = x_0 + A_u - i B_u,   X_(R-u) = x_0 + A_u + i B_u,   X_0 = x_0 + sum_t S_t
        -- a quarter of the multiplications of the R x R sum;
      * radix 2 and 4 are the usual butterflies."""
    m = z.size
    ns = 1
    cur = z.astype(complex).copy()
    for p, r in enumerate(factors):
        L = m // r
        out = np.zeros(m, complex)
        j = np.arange(L)
        jm = j % ns
        base = (j - jm) * r + jm
        x = [cur[j + t * L] for t in range(r)]
        if r == 2:
            res = {0: x[0] + x[1], 1: x[0] - x[1]}
        elif r == 4:
            t0, t1, t2, t3 = x[0] + x[2], x[0] - x[2], x[1] + x[3], x[1] - x[3]
            res = {0: t0 + t2, 2: t0 - t2, 1: t1 - 1j * t3, 3: t1 + 1j * t3}
        else:
            h = (r - 1) // 2
            S = [x[t] + x[r - t] for t in range(1, h + 1)]
            D = [x[t] - x[r - t] for t in range(1, h + 1)]
            res = {0: x[0] + sum(S)}
            for u in range(1, h + 1):
                w = [tab_m[((u * t) % r) * L] for t in range(1, h + 1)]       # (cos, -sin) of 2 pi u t / R
                A = sum(S[t] * w[t].real for t in range(h))
                B = sum(D[t] * (-w[t].imag) for t in range(h))
                res[u] = x[0] + A - 1j * B
                res[r - u] = x[0] + A + 1j * B
        for u, v in res.items():
            out[base + u * ns] = v
        ns *= r
        if p + 1 < len(factors):                  # pre-twiddle for the next pass, by output position
            rn = factors[p + 1]
            Ln = m // rn
            pos = np.arange(m)
            tn = pos // Ln
            jn = pos - tn * Ln
            k = (jn % ns) * tn * (m // (ns * rn))
            assert k.max() < m
            out = out * tab_m[k]
        cur = out
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
        z = rng.standard_normal(m) + 1j * rng.standard_normal(m)
        pair = np.abs(cfft_passes_paired(z, f, tm) - np.fft.fft(z)).max() / np.abs(np.fft.fft(z)).max()
        print(n, f, err, back, pair)
