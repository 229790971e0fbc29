"""Thread-level NumPy model of csrc/zfft.cu's warp-local DCT-II / DCT-III kernels (k_dct_rows_w), N = 16384:
index maps, shared-memory addresses (bank conflicts counted per 16-lane group), twiddle bookkeeping.  Checked against
scipy.fft.dct.  The CUDA kernel follows this file line by line; tests/test_fd_factors_cpu.py runs it.

Row of N reals -> complex z[m], M = N/2 = 8192 points, m = n1 + 32 n2:
   phase 1 (half-warp local): the 256-point FFT over n2 of every residue n1 (two radix-16 passes, exchange X1)
   phase 2 (CTA wide, X2):    radix-16 over d, n1 = c + 2 d  -> E (c = 0) / O (c = 1), the half-length FFTs
   phase 3 (warp local, X3):  radix-2 + real-FFT untangling + quarter-wave rotation on quadruples
"""
import numpy as np

N = 16384
M = N // 2
H = M // 2
T = 512            # threads
RH = np.sqrt(0.5)


def W(n, k):
    return np.exp(-2j * np.pi * k / n)


def dft16(v):
    """v[..., 16] natural order in and out"""
    F = W(16, np.outer(np.arange(16), np.arange(16)))
    return v @ F.T


class Banks:
    """counts the worst bank multiplicity of 8-byte accesses per group of 16 consecutive lanes"""

    def __init__(self):
        self.worst = {}

    def check(self, name, addr):
        """addr[tid] in 8-byte words"""
        a = np.asarray(addr).reshape(-1, 16)
        worst = 1
        for g in a:
            _, cnt = np.unique(g % 16, return_counts=True)
            # identical addresses broadcast: count distinct addresses per bank
            per_bank = {}
            for x in g:
                per_bank.setdefault(x % 16, set()).add(x)
            worst = max(worst, max(len(s) for s in per_bank.values()))
        self.worst[name] = max(self.worst.get(name, 1), worst)


def ids():
    tid = np.arange(T)
    w, lane = tid >> 5, tid & 31
    hw = (lane >> 3) & 1
    j = (lane & 7) | ((lane >> 4) << 3)
    return tid, w, lane, hw, j


def sigma(n1):
    return 8 * ((n1 & 1) ^ (n1 >> 4))


def stage_word(s, word):
    """8-byte word address of word `word` (0..3) of sector s in the staged row (csrc/zfft.cu stage_addr8): the TMA box
    (16 doubles, 8 x 1 KB, 8 x 128 B, 16 x 8 KB) with the 128-byte swizzle puts source offset
    B = 8192 i3 + 1024 i1 + 128 i2 + 16 cc + .. into line 64 i3 + 8 i2 + i1 at chunk cc ^ i1"""
    h = word >> 1
    i1, i2, i3 = (s >> 5) & 7, (s >> 2) & 7, s >> 8
    return i3 * 1024 + i2 * 128 + i1 * 16 + 2 * (((2 * s + h) & 7) ^ i1) + (word & 1)


def stage_fill(x):
    """what the TMA load leaves in shared memory"""
    S = np.zeros(x.size)
    B = 16 * np.arange(x.size // 2)                      # byte offset of every 16-byte chunk of the row
    i3, i1, i2, cc = B >> 13, (B >> 10) & 7, (B >> 7) & 7, (B >> 4) & 7
    dst = ((i3 * 64 + i2 * 8 + i1) * 8 + (cc ^ i1)) * 2
    S[dst] = x[B // 8]
    S[dst + 1] = x[B // 8 + 1]
    return S


def fft_core(v, n1, a_role, kb_role, bk, tabM, sign=1.0, x1x=0, mirror=256):
    """phases 1b .. 2 shared by both transforms.  v[tid, b]: the 16 points z[n1 + 32 (a + 16 b)] of every thread.
    Returns vq[tid, e] = E/O[k2 + 256 e] in the phase-3 thread layout together with (k2, c, mu, G)."""
    tid, w, lane, hw, j = ids()
    d = n1 >> 1
    v = dft16(v)                                              # U[a][kb]
    # tw1 (with the b0 factor of tw2 folded in): w^kb, w = W_4096^(16 a + d)
    jw = 16 * a_role + d
    w1 = sign * tabM[2 * jw]                                    # sign = -1: the slots hold the points rotated by 8
    kb = np.arange(16)
    v = v * w1[:, None] ** kb[None, :]
    # X1: half-warp transpose through XB, element (a, kb) of half-warp region R at R + 16 a + (kb ^ a)
    XB = np.zeros(M, complex)
    R = 512 * w + 256 * hw
    for k in range(16):
        addr = R + 16 * a_role + (k ^ a_role ^ x1x)
        bk.check("X1 write", addr)
        XB[addr] = v[:, k]
    u = np.empty_like(v)
    for a in range(16):
        addr = R + 16 * a + (kb_role ^ a ^ x1x)
        bk.check("X1 read", addr)
        u[:, a] = XB[addr]
    u = dft16(u)                                              # Y[n1][kb + 16 ka] (times b0)
    # tw2: S[d][ka] = W_256^(d ka)
    ka = np.arange(16)
    u = u * W(256, d[:, None] * ka[None, :])
    # X2: CTA wide, A(n1, k2) = 256 n1 + (k2 ^ sigma(n1))
    XB = np.zeros(M, complex)
    for q in range(16):
        k2 = kb_role + 16 * q
        addr = 256 * n1 + (k2 ^ sigma(n1))
        bk.check("X2 write", addr)
        XB[addr] = u[:, q]
    mu, c, G = lane >> 4, (lane >> 3) & 1, lane & 7
    if mirror == 255:                                       # DCT-III: residues pair as k2 <-> 255 - k2
        k2 = np.where(mu == 0, 8 * w + G, 255 - 8 * w - G)
    else:                                                   # DCT-II: k2 <-> 256 - k2, 0 and 128 pair with themselves
        k2 = np.where(mu == 0, 8 * w + G, 256 - 8 * w - G)
        k2 = np.where((mu == 1) & (w == 0) & (G == 0), 128, k2)
    vq = np.empty_like(u)
    for dd in range(16):
        nn = c + 2 * dd
        addr = 256 * nn + (k2 ^ sigma(nn))
        bk.check("X2 read", addr)
        vq[:, dd] = XB[addr]
    vq = dft16(vq)                                            # E / O [k2 + 256 e]
    return vq, k2, c, mu, G


def x3_addr(w, G, m, e):
    return 512 * w + 64 * G + 16 * m + (e ^ ((2 * G + (m & 1)) & 15))


def dct2_w(x, tabs, bk=None):
    """DCT-II of one row: X[k] = sum_j x[j] cos(pi k (2j+1) / 2N)"""
    bk = bk or Banks()
    tabM, tabN, tabQ = tabs
    tid, w, lane, hw, j = ids()
    # ---- stage: the row as the TMA load leaves it
    S = stage_fill(x)
    # ---- phase 1a: load 16 points.  hw0: n1 = w, a = j, slot i holds b = i;  hw1: n1 = 31 - w, a = 15 - j,
    #      slot i holds b = i ^ 8 (the rotation by 8 is a sign (-1)^kb, folded into tw1)
    n1 = np.where(hw == 0, w, 31 - w)
    a_role = np.where(hw == 0, j, 15 - j)
    v = np.empty((T, 16), complex)
    for i in range(16):
        b = np.where(hw == 0, i, i ^ 8)
        m = n1 + 32 * (a_role + 16 * b)
        lowhalf = m < H
        s = np.where(lowhalf, m, M - 1 - m)
        w_re = np.where(lowhalf, 0, 3)
        w_im = np.where(lowhalf, 2, 1)
        a_re, a_im = stage_word(s, w_re), stage_word(s, w_im)
        bk.check("stage read", a_re)
        bk.check("stage read", a_im)
        v[:, i] = S[a_re] + 1j * S[a_im]
    # DFT of the rotated sequence: Y[k] = (-1)^k X[k] for hw1 -> tw1's base is negated
    sign = np.where(hw == 0, 1.0, -1.0)
    kb_role = np.where(hw == 0, j, j ^ 8)
    vq, k2, c, mu, G = fft_core(v, n1, a_role, kb_role, bk, tabM, sign)
    # ---- phase 3: X3 (warp local) + radix 2 + untangle.  member m = 2 mu + c of group G holds E/O of residue k2
    m = 2 * mu + c
    XB = np.zeros(M, complex)
    for e in range(16):
        addr = x3_addr(w, G, m, e)
        bk.check("X3 write", addr)
        XB[addr] = vq[:, e]
    X = np.zeros(N)

    def emit(k, zk, zm, q, tn):
        ex, ey = 0.5 * (zk.real + zm.real), 0.5 * (zk.imag - zm.imag)
        dd = 0.5 * (zk.real - zm.real) + 0.5j * (zk.imag + zm.imag)
        qm = RH * (q.real - q.imag) - 1j * RH * (q.real + q.imag)
        p = tn * dd
        vk = (ex + p.imag) + 1j * (ey - p.real)
        vm = (ex - p.imag) + 1j * (-ey - p.real)
        aa, bb = q * vk, qm * vm
        X[k], X[N - k], X[M - k], X[M + k] = aa.real, -aa.imag, bb.real, -bb.imag

    def quad(k, v0, v1, v2, v3):
        """E[k], O[k], E[H-k], O[H-k] -> the 8 outputs of {k, k+H, H-k, M-k}"""
        q = tabQ[k]
        tn = tabN[k]
        wk = tabM[k]
        wb, cwd = wk * v1, np.conj(wk) * v3
        z_k, z_kh, z_hk, z_mk = v0 + wb, v0 - wb, v2 - cwd, v2 + cwd
        q2 = tabQ[H - k]
        emit(k, z_k, z_mk, q, tn)
        emit(H - k, z_hk, z_kh, q2, -tn.imag - 1j * tn.real)

    special = (w == 0) & (G == 0)
    for t in range(T):
        ww, gg, mm = w[t], G[t], m[t]
        if not special[t]:
            kk2 = 8 * ww + gg
            for i in range(4):
                e = 4 * i + mm
                addrs = [x3_addr(ww, gg, 0, e), x3_addr(ww, gg, 1, e), x3_addr(ww, gg, 2, 15 - e),
                         x3_addr(ww, gg, 3, 15 - e)]
                quad(kk2 + 256 * e, *[XB[a_] for a_ in addrs])
        elif mm >= 2:                                   # residue 128: members 2, 3 pair with themselves
            cc = mm - 2
            for i in range(4):
                e = 2 * i + cc
                quad(128 + 256 * e, XB[x3_addr(0, 0, 2, e)], XB[x3_addr(0, 0, 3, e)], XB[x3_addr(0, 0, 2, 15 - e)],
                     XB[x3_addr(0, 0, 3, 15 - e)])
        else:                                           # residue 0: e = 1 .. 7 pair with 16 - e; k = 0 and k = H/2
            cc = mm
            es = (1, 3, 5, 7) if cc == 0 else (2, 4, 6)
            for e in es:
                quad(256 * e, XB[x3_addr(0, 0, 0, e)], XB[x3_addr(0, 0, 1, e)], XB[x3_addr(0, 0, 0, 16 - e)],
                     XB[x3_addr(0, 0, 1, 16 - e)])
            if cc == 1:
                e0, o0 = XB[x3_addr(0, 0, 0, 0)], XB[x3_addr(0, 0, 1, 0)]
                z0, zh = e0 + o0, e0 - o0
                X[0] = z0.real + z0.imag
                X[M] = (z0.real - z0.imag) * RH
                emit(H, zh, zh, tabQ[H], -1j)
                e8, o8 = XB[x3_addr(0, 0, 0, 8)], XB[x3_addr(0, 0, 1, 8)]
                k = H // 2
                wk = tabM[k]
                emit(k, e8 + wk * o8, e8 - wk * o8, tabQ[k], tabN[k])
    # bank check of the regular X3 reads
    for i in range(4):
        e = 4 * i + m
        for ms in range(4):
            ee = e if ms < 2 else 15 - e
            bk.check("X3 read", x3_addr(w, G, ms, ee))
    return X, bk


def stage_word3(k):
    """DCT-III stage: TMA box (16 doubles, 8 x 256 B, 2 x 128 B, 64 x 2 KB), 128-byte swizzle: source offset
    B = 8 k = 2048 i3 + 256 i1 + 128 i2 + 16 cc + .. lands in line 16 i3 + 8 i2 + i1 at chunk cc ^ i1"""
    cc, i2, i1, i3 = (k >> 1) & 7, (k >> 4) & 1, (k >> 5) & 7, k >> 8
    return (16 * i3 + 8 * i2 + i1) * 16 + 2 * (cc ^ i1) + (k & 1)


def stage_fill3(x):
    S = np.zeros(x.size)
    k = np.arange(x.size)
    S[stage_word3(k)] = x
    return S


def dct3_w(a_in, tabs, bk=None):
    """DCT-III of one row: y[j] = sum_k a[k] cos(pi k (2j+1) / 2N), through the same forward FFT on swapped (im, re)"""
    bk = bk or Banks()
    tabM, tabN, tabQ = tabs
    tid, w, lane, hw, j = ids()
    S = stage_fill3(a_in)
    # ---- phase-1 roles: hw0 does residue n1 = w, hw1 residue 32 - w (Z[k] and Z[M-k] come as pairs); warp 0: 0 and 16
    n1 = np.where(hw == 0, w, 32 - w)
    n1 = np.where((w == 0) & (hw == 1), 16, n1)
    a_role = j.copy()
    kb_role = j.copy()
    v = np.zeros((T, 16), complex)
    XB = np.zeros(M, complex)

    def pair(k):
        """(Z[k], Z[M-k]) swapped (im, re), 1 <= k < M/2"""
        ak, ank, amk, apk = (S[stage_word3(k)], S[stage_word3(N - k)], S[stage_word3(M - k)], S[stage_word3(M + k)])
        qk = tabQ[k]
        qm = RH * (qk.real - qk.imag) - 1j * RH * (qk.real + qk.imag)
        bkk = np.conj(qk) * (0.5 * ak - 0.5j * ank)
        bm = np.conj(qm) * (0.5 * amk - 0.5j * apk)
        sx, sy = bkk.real + bm.real, bkk.imag - bm.imag
        dd = (bkk.real - bm.real) + 1j * (bkk.imag + bm.imag)
        q = np.conj(tabN[k]) * dd
        return (sy + q.real) + 1j * (sx - q.imag), (-sy + q.real) + 1j * (sx + q.imag)

    for i in range(8):                                      # bank check of the four stage reads of own slot i
        k = n1 + 32 * (a_role + 16 * i)
        for idx in (k, N - k, M - k, M + k):
            bk.check("stage3 read", stage_word3(np.where((idx >= 0) & (idx < N), idx, 0)))
    for t in range(T):
        ww, hh, jj, nn, aa = w[t], hw[t], j[t], n1[t], a_role[t]
        base = 512 * ww
        for i in range(8):
            k = nn + 32 * (aa + 16 * i)
            if k == 0:
                a0, am = S[stage_word3(0)], S[stage_word3(M)] * RH
                v[t, 0] = (a0 - am) + 1j * (a0 + am)
                # Z[M/2] (its own mirror): slot 8 of the same thread
                ak, ank = S[stage_word3(H)], S[stage_word3(N - H)]
                amk, apk = S[stage_word3(M - H)], S[stage_word3(M + H)]
                qk = np.exp(-1j * np.pi / 8)
                qm = RH * (qk.real - qk.imag) - 1j * RH * (qk.real + qk.imag)
                bkk = np.conj(qk) * (0.5 * ak - 0.5j * ank)
                bm = np.conj(qm) * (0.5 * amk - 0.5j * apk)
                sx, sy = bkk.real + bm.real, bkk.imag - bm.imag
                dd = (bkk.real - bm.real) + 1j * (bkk.imag + bm.imag)
                q = np.conj(-1j) * dd
                v[t, 8] = (sy + q.real) + 1j * (sx - q.imag)
                continue
            zk, zm = pair(k)
            v[t, i] = zk
            # destination of Z[M-k]
            if ww > 0:
                dh, dj, ds = hh ^ 1, 15 - jj, 15 - i
            elif hh == 0:                                   # residue 0: n2 <-> 256 - n2
                if aa == 0:
                    dh, dj, ds = 0, 0, 16 - i
                else:
                    dh, dj, ds = 0, 16 - jj, 15 - i
            else:                                           # residue 16: n2 <-> 255 - n2
                dh, dj, ds = 1, 15 - jj, 15 - i
            XB[base + dh * 128 + (ds - 8) * 16 + (dj ^ (8 * dh))] = zm
    for t in range(T):
        ww, hh, jj = w[t], hw[t], j[t]
        for sl in range(8, 16):
            if ww == 0 and hh == 0 and jj == 0 and sl == 8:
                continue                                    # Z[M/2], already in place
            v[t, sl] = XB[512 * ww + hh * 128 + (sl - 8) * 16 + (jj ^ (8 * hh))]
    x1x = np.where(hw == 0, 0, 8)
    vq, k2, c, mu, G = fft_core(v, n1, a_role, kb_role, bk, tabM, 1.0, x1x, mirror=255)
    # ---- phase 3: quadruples {n, n + M/2, M/2 - 1 - n, M - 1 - n}, n = (8 w + G) + 256 e
    m = 2 * mu + c
    XB = np.zeros(M, complex)
    for e in range(16):
        addr = x3_addr(w, G, m, e)
        bk.check("X3 write", addr)
        XB[addr] = vq[:, e]
    y = np.zeros(N)
    wm1 = tabM[1]
    for t in range(T):
        ww, gg, mm = w[t], G[t], m[t]
        for i in range(4):
            e = 4 * i + mm
            n = 8 * ww + gg + 256 * e
            n2 = H - 1 - n
            v0, v1 = XB[x3_addr(ww, gg, 0, e)], XB[x3_addr(ww, gg, 1, e)]
            v2, v3 = XB[x3_addr(ww, gg, 2, 15 - e)], XB[x3_addr(ww, gg, 3, 15 - e)]
            wn = tabM[n]
            w1 = wn * wm1
            wb, wd = wn * v1, (-np.conj(w1)) * v3
            z_n, z_nh, z_c, z_m = v0 + wb, v0 - wb, v2 + wd, v2 - wd
            y[4 * n:4 * n + 4] = z_n.imag, z_m.real, z_n.real, z_m.imag
            y[4 * n2:4 * n2 + 4] = z_c.imag, z_nh.real, z_c.real, z_nh.imag
    return y, bk


def tables():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from pyaxisymflow_b200.fd import dct_tables

    t = dct_tables(N)
    t = t[:, 0] + 1j * t[:, 1]
    return t[:M], t[M:2 * M + 1], t[2 * M + 1:]


if __name__ == "__main__":
    import scipy.fft as sf

    rng = np.random.default_rng(0)
    x = rng.standard_normal(N)
    X, bk = dct2_w(x, tables())
    ref = sf.dct(x, 2) / 2
    print("dct2_w vs scipy:", np.abs(X - ref).max() / np.abs(ref).max())
    print("worst bank multiplicity per 16-lane group:", bk.worst)
    a = rng.standard_normal(N)
    y, bk3 = dct3_w(a, tables())
    ref3 = sf.dct(a, 3) / 2 + a[0] / 2                      # scipy's DCT-III halves the k = 0 term
    print("dct3_w vs scipy:", np.abs(y - ref3).max() / np.abs(ref3).max())
    print("worst bank multiplicity per 16-lane group:", bk3.worst)
