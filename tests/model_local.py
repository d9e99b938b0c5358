"""Numpy model of the LOCAL traced batch: the per-cell code of K2's local instantiation
(gotoh_stream.cuh, LT), the per-lane end-cell tracking, and K4's local walk with Waterman-Eggert
boxes and local preprofile counts (traceback.cu).  Executable specification of the encoding
argument (stop code in the free 'second gap state' bit), tested against the oracle on the CPU."""
import numpy as np

NEG = -np.inf


def fill(a, b, S, go, ge, boxes=(), K=2):
    """Returns (nibbles [L1+1, L2+1], key) with key = (best, y, x) reduced over lanes like atomicMax."""
    L1, L2 = len(a), len(b)
    go, ge = np.float32(go), np.float32(ge)
    M = np.full((L1 + 1, L2 + 1), NEG, np.float32)
    U = np.full((L1 + 1, L2 + 1), NEG, np.float32)
    Lm = np.full((L1 + 1, L2 + 1), NEG, np.float32)
    M[0, 0] = 0
    U[0, 0] = go - ge
    Lm[0, 0] = go - ge
    for y in range(1, L1 + 1):
        U[y, 0] = np.float32((y - 1) * ge + go)
    for x in range(1, L2 + 1):
        Lm[0, x] = np.float32((x - 1) * ge + go)
    nib = np.zeros((L1 + 1, L2 + 1), np.uint8)
    nl = (L2 + K - 1) // K
    best = np.zeros(nl, np.float32)
    by = np.zeros(nl, np.int64)
    bx = np.zeros(nl, np.int64)
    sign = lambda v: 1 if np.signbit(v) and not np.isnan(v) else 0
    for y in range(1, L1 + 1):
        rb = np.zeros(nl, np.float32)
        rk = np.full(nl, -1)
        for x in range(1, L2 + 1):
            d = max(M[y - 1, x - 1], U[y - 1, x - 1], Lm[y - 1, x - 1])
            mraw = np.float32(d + S[a[y - 1], b[x - 1]])
            m = max(mraw, np.float32(0))
            mo_up, ue = np.float32(M[y - 1, x] + go), np.float32(U[y - 1, x] + ge)
            ml, le = np.float32(M[y, x - 1] + go), np.float32(Lm[y, x - 1] + ge)
            u, l = max(mo_up, ue), max(ml, le)
            if any(ylo <= y <= yhi and xlo <= x <= xhi for (ylo, yhi, xlo, xhi) in boxes):
                m = u = l = np.float32(0)
            with np.errstate(invalid="ignore"):
                ul = max(u, l)
                nm = np.float32(m - ul)
                b0 = sign(nm)
                b1 = sign(np.float32(u - l)) if b0 else sign(mraw)
                b2 = sign(np.float32(mo_up - ue))
                b3 = sign(np.float32(ml - le))
            nib[y, x] = b0 | (b1 << 1) | (b2 << 2) | (b3 << 3)
            M[y, x], U[y, x], Lm[y, x] = m, u, l
            lane = (x - 1) // K
            if m > rb[lane]:
                rb[lane] = m
                rk[lane] = x
        for lane in range(nl):
            if rb[lane] > best[lane]:
                best[lane], by[lane], bx[lane] = rb[lane], y, rk[lane]
    # atomicMax over (value, ~(y << 11 | x))
    order = sorted(range(nl), key=lambda i: (-best[i], by[i], bx[i]))
    i = order[0]
    key = (float(best[i]), int(by[i]), int(bx[i])) if best[i] > 0 else (0.0, 0, 0)
    return nib, key


def walk(nib, key, boxes=()):
    """K4 local mode: path (front to back) in reference coordinates."""
    v, y, x = key
    s = 0
    path = []
    while True:
        path.append((y, x))
        if y == 0 and x == 0:
            break
        if x == 0:
            if s == 1:
                y -= 1
                continue
            break
        if y == 0:
            if s == 2:
                x -= 1
                continue
            break
        n = int(nib[y, x])
        if any(ylo <= y <= yhi and xlo <= x <= xhi for (ylo, yhi, xlo, xhi) in boxes):
            break
        if s == 0:
            if (n & 3) == 2:
                break
            y, x = y - 1, x - 1
            if y == 0 and x == 0:
                s = 0          # code00 of the local borders: M (0 > open - extend)
            elif y == 0:
                s = 2
            elif x == 0:
                s = 1
            else:
                c = int(nib[y, x])
                s = 0 if not (c & 1) else (2 if (c & 2) else 1)
        elif s == 1:
            s = 1 if (n >> 2) & 1 else 0
            y -= 1
        else:
            s = 2 if (n >> 3) & 1 else 0
            x -= 1
    path.reverse()
    return np.asarray(path, np.int32)


def counts_from_path(path, slave, L1, A, counts):
    """K4's local preprofile rule on a front-to-back path."""
    first = {}
    for (y, x) in path:
        if y not in first:
            first[y] = x
    y0, ye = int(path[0][0]), int(path[-1][0])
    for y in range(y0 + 1, ye + 1):
        if first[y] > first[y - 1]:
            counts[y - 1, slave[first[y] - 1]] += 1
    if y0 >= 1:
        x0 = int(path[0][1])
        counts[y0 - 1, slave[x0 - 1 if x0 >= 1 else len(slave) - 1]] += 1
