"""Model of the GPU narrow-band re-initialisation (csrc/reinit.cu) in plain Python/NumPy.

The fast-marching method accepts cells one at a time from a heap; the GPU instead iterates the
SAME upwind update to its fixed point, all band cells at once:

  step 1  front cells (4-neighbourhood straddles the zero contour): distance from the linear
          crossings -- identical to the marcher's first step, purely local;
  step 2  repeat until nothing changes: every other cell recomputes its value from the neighbours
          that the marcher would have frozen before it -- front cells and cells whose current
          |value| <= narrow, and *causally* smaller than the result (a dimension whose upwind
          value is not below the 2-D result is dropped and the 1-D result used).  The iteration is
          organised like the kernel: 32x32 tiles, up to 64 Jacobi iterations per tile and launch with
          the other tiles' values frozen;
  step 3  cells not accepted (|value| > narrow) that touch an accepted cell get the marcher's
          tentative value (update from all accepted neighbours, no causality filter), everything
          else stays masked.

tests/test_widen_oracle_cpu.py checks this model against the heap-based restatement
(oracle.fmm_distance) for equality; the CUDA kernels repeat this arithmetic operation by operation.
This file is test code (it imports the oracle), not product code.
"""
import numpy as np

MAXD = np.finfo(np.float64).max
EPS = np.finfo(np.float64).eps


def _quadratic(a, b, c, positive):
    c = c - 1
    det = b * b - 4 * a * c
    if det < 0:
        return None
    if positive:
        return (-b + np.sqrt(det)) / 2.0 / a
    return (-b - np.sqrt(det)) / 2.0 / a


def _dim_terms(v1, v2, idx2):
    aa = 9.0 / 4.0
    if v2 < MAXD:
        tp = (1.0 / 3.0) * (4 * v1 - v2)
        return idx2 * aa, -(idx2 * 2 * aa * tp), idx2 * aa * (tp * tp)
    return idx2, -(idx2 * 2 * v1), idx2 * (v1 * v1)


def _upwind(d, ok, j, k, dim, order):
    """(value1, value2) of one dimension from the usable neighbours (distance_marcher's selection)."""
    nr, nz = d.shape
    v1 = v2 = MAXD
    for s in (-1, 1):
        jj, kk = (j + s, k) if dim == 0 else (j, k + s)
        if not (0 <= jj < nr and 0 <= kk < nz) or not ok[jj, kk]:
            continue
        if abs(d[jj, kk]) < abs(v1):
            v1 = d[jj, kk]
            j2, k2 = (j + 2 * s, k) if dim == 0 else (j, k + 2 * s)
            if order == 2 and 0 <= j2 < nr and 0 <= k2 < nz and ok[j2, k2]:
                d2 = d[j2, k2]
                if (d2 <= v1 and v1 >= 0) or (d2 >= v1 and v1 <= 0):
                    v2 = d2
    return v1, v2


def update_cell(d, ok, phi, j, k, dx, order, causal):
    """new value of cell (j,k); None = no usable neighbour (or, non-causal form, negative discriminant)."""
    idx2 = 1 / dx / dx
    pos = phi[j, k] > EPS
    t = [_upwind(d, ok, j, k, dim, order) for dim in (0, 1)]
    have = [t[0][0] < MAXD, t[1][0] < MAXD]
    if not (have[0] or have[1]):
        return None
    if have[0] and have[1]:
        a0, b0, c0 = _dim_terms(*t[0], idx2)
        a1, b1, c1 = _dim_terms(*t[1], idx2)
        # accumulate like the marcher: a = 0 + a0 + a1, b = 0 - x0 - x1, c = 0 + c0 + c1
        r = _quadratic(a0 + a1, b0 + b1, c0 + c1, pos)
        if not causal:
            return r
        big = max(abs(t[0][0]), abs(t[1][0]))
        if r is not None and abs(r) > big:
            return r
        dim = 0 if abs(t[0][0]) <= abs(t[1][0]) else 1
    else:
        dim = 0 if have[0] else 1
    a, b, c = _dim_terms(*t[dim], idx2)
    return _quadratic(a, b, c, pos)


TH = TW = 32      # csrc/reinit.cu: tile shape, on-chip iterations per launch
INNER = 64


def launch_bounds(shape, narrow, dx):
    """(free_launch, max_launch) exactly as axb_reinit_distance computes them"""
    w = int(np.ceil(narrow / dx))
    chain = 2 * w + int(np.ceil(2.0 * np.sqrt(2.0 * max(shape) * w)))
    free_launch = 8 + 2 * ((chain + TW - 1) // TW)
    return free_launch, 2 * free_launch


def reinit(phi, dx, narrow, order=2, free_launch=None, max_launch=None):
    """returns (distance, unmasked, launches).  distance is MAXD where masked.

    One launch = every tile iterates up to INNER times on its own cells with the values of the other tiles
    frozen (what a thread block does in shared memory), ending early when the tile is stationary.  The
    iteration ends when a launch changes no tile.  After ``free_launch`` launches values may only decrease in
    magnitude: where two fronts collide inside the band, or neighbours are exactly tied, the second-order
    update can otherwise flip between two upwind selections for ever."""
    from oracle.axisym_oracle import fmm_initial_front

    phi = np.asarray(phi, dtype=np.float64)
    nr, nz = phi.shape
    fl, ml = launch_bounds(phi.shape, narrow, dx)
    free_launch = fl if free_launch is None else free_launch
    max_launch = ml if max_launch is None else max_launch
    d, front = fmm_initial_front(phi, dx)
    launch, converged = 0, False
    while launch < max_launch and not converged:
        launch += 1
        monotone = launch > free_launch
        new = d.copy()
        for tj in range(0, nr, TH):
            for tk in range(0, nz, TW):
                ja, jb, ka, kb = max(tj - 2, 0), min(tj + TH + 2, nr), max(tk - 2, 0), min(tk + TW + 2, nz)
                if not np.any(front[ja:jb, ka:kb] | (np.abs(d[ja:jb, ka:kb]) <= narrow)):
                    continue                      # nothing usable within reach: every cell stays / becomes MAXD
                loc = d.copy()                    # the tile's view: other tiles frozen at the launch start
                own = np.zeros_like(front)
                own[tj:tj + TH, tk:tk + TW] = True
                cells = list(zip(*np.nonzero(own & ~front)))
                for _ in range(INNER):
                    ok = front | (np.abs(loc) <= narrow)
                    nxt = loc.copy()
                    changed = False
                    for j, k in cells:
                        r = update_cell(loc, ok, phi, j, k, dx, order, causal=True)
                        r = MAXD if r is None else r
                        if monotone and not abs(r) < abs(loc[j, k]):
                            r = loc[j, k]
                        if r != loc[j, k]:
                            changed = True
                        nxt[j, k] = r
                    loc = nxt
                    if not changed:
                        break
                new[tj:tj + TH, tk:tk + TW] = loc[tj:tj + TH, tk:tk + TW]
        converged = np.array_equal(new, d)
        d = new
    if not converged:
        raise RuntimeError("reinit model: no fixed point within the launch bound")
    acc = front | (np.abs(d) <= narrow)
    out = np.where(acc, d, MAXD)
    reach = np.zeros_like(acc)
    reach[1:] |= acc[:-1]; reach[:-1] |= acc[1:]; reach[:, 1:] |= acc[:, :-1]; reach[:, :-1] |= acc[:, 1:]
    ring = reach & ~acc
    for j, k in zip(*np.nonzero(ring)):
        r = update_cell(d, acc, phi, j, k, dx, order, causal=False)
        if r is None:
            raise RuntimeError("Negative discriminant in distance marcher quadratic.")
        out[j, k] = r
    return out, acc | ring, launch
