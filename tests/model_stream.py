"""numpy lane-model of the inter-task streaming kernel (praline_b200/csrc/gotoh_stream.cu).

One "warp" = 32 lanes held as numpy vectors; a resident sequence is laid across the lanes
(K columns each) and a concatenated stream of row sequences is pushed through the systolic
pipeline: lane l works on stream position t - l at step t.  The model mirrors the kernel's
data flow statement by statement (ring words, reset/emit flags, shuffle of the strip edge,
packed 4-bit traceback in [step/8][k][lane] words) so that indexing logic can be checked
against the oracle on the CPU; the CUDA kernel is a transcription of it.
"""
import numpy as np

FLAG_LAST = 1 << 31
FLAG_EMIT = 1 << 30
NINF = np.float32(-np.inf)


def borders(mode, go, ge, maxlen, transposed=False):
    """topD[x], leftD[y] (kernel orientation) from reference component/align.py:367-385.
    mode in reference numbering; returns also code00, top_ramp, left_ramp flags."""
    go = np.float32(go)
    ge = np.float32(ge)
    n = np.arange(maxlen)
    ramp = np.empty(maxlen + 1, np.float32)
    ramp[1:] = n * np.full(maxlen, ge, np.float32) + go   # int64 * f32 -> f64, + f32 scalar, cast on store
    u_zero = mode in (2, 3)          # semiglobal_both / one : U border (column 0) is 0
    l_zero = mode in (2, 4)          # semiglobal_both / two : L border (row 0) is 0
    u00 = np.float32(0) if u_zero else np.float32(go - ge)
    l00 = np.float32(0) if l_zero else np.float32(go - ge)
    colU = np.zeros(maxlen + 1, np.float32) if u_zero else ramp.copy()   # D(y,0), y>=1
    rowL = np.zeros(maxlen + 1, np.float32) if l_zero else ramp.copy()   # D(0,x), x>=1
    vals = [np.float32(0), u00, l00]
    d00 = max(vals)
    code00_ref = int(np.argmax(vals))            # first argmax in (M, U, L) order
    colU[0] = d00
    rowL[0] = d00
    if not transposed:
        return dict(topD=rowL, leftD=colU, code00=code00_ref, top_ramp=not l_zero, left_ramp=not u_zero)
    # kernel columns = sequence one: kernel top border = reference column-0 (U) border
    code00_k = {0: 0, 1: 2, 2: 1}[code00_ref]
    return dict(topD=colU, leftD=rowL, code00=code00_k, top_ramp=not u_zero, left_ramp=not l_zero)


def run_warp(K, resident, streams, S, go, ge, mode, transposed=False, want_tb=False):
    """Returns per streamed sequence: dict(score / rowkey / colkey, emit_t) and the tb words."""
    lanes = np.arange(32)
    L2 = len(resident)
    assert L2 <= 32 * K
    local = mode == 1
    semi = mode in (2, 3, 4)
    maxlen = max([32 * K] + [len(s) for s in streams]) + 1
    B = borders(mode, go, ge, maxlen, transposed)
    topD, leftD = B["topD"], B["leftD"]
    go = np.float32(go)
    ge = np.float32(ge)
    # profile: prof[a][lane][k]
    A = S.shape[0]
    prof = np.full((A, 32, K), NINF if local else 0, np.float32)
    for x in range(L2):
        b = resident[x]
        prof[:, x // K, x % K] = S[b, :] if transposed else S[:, b]
    lr, klast = (L2 - 1) // K, (L2 - 1) % K
    # stream words
    words = [FLAG_LAST]  # dummy row
    for s in streams:
        for p, sym in enumerate(s):
            words.append(int(sym) | (FLAG_LAST | FLAG_EMIT if p == len(s) - 1 else 0))
    total = len(words)
    T = (total + 31 + 7) // 8 * 8
    words = words + [0] * (T + 64)

    def reset_state(mask, Mo, U, D, Dleft_prev, y, colbest):
        for k in range(K):
            Mo[mask, k] = NINF
            U[mask, k] = NINF
            D[mask, k] = topD[lanes[mask] * K + k + 1]
        Dleft_prev[mask] = topD[lanes[mask] * K]
        y[mask] = 0
        colbest[mask] = topD[L2]

    Mo = np.zeros((32, K), np.float32)
    U = np.zeros((32, K), np.float32)
    D = np.zeros((32, K), np.float32)
    Dleft_prev = np.zeros(32, np.float32)
    Mo_last = np.zeros(32, np.float32)
    L_last = np.zeros(32, np.float32)
    D_last = np.zeros(32, np.float32)
    y = np.zeros(32, np.int64)
    q = np.zeros(32, np.int64)
    best = np.zeros(32, np.float32)
    colbest = np.zeros(32, np.float32)
    colbest_y = np.zeros(32, np.int64)
    acc = np.zeros((32, K), np.uint32)
    tb = np.zeros((T // 8, K, 32), np.uint32)
    out = [dict(score=None, rowkey=(NINF, -1), colkey=(NINF, -1), emit_t=None) for _ in streams]
    if local:
        for o in out:
            o["score"] = np.float32(0)

    with np.errstate(invalid="ignore"):
        for t in range(T):
            r = t - lanes
            w = np.array([words[ri] if ri >= 0 else 0 for ri in r], np.uint64)
            sym = (w & 0xFFFF).astype(np.int64)
            y += 1
            # receive strip edge from the left lane
            Ml = np.roll(Mo_last, 1)
            Ll = np.roll(L_last, 1)
            Dn = np.roll(D_last, 1)
            Ml[0] = NINF
            Ll[0] = NINF
            Dn[0] = leftD[min(y[0], maxlen)]
            diag = Dleft_prev.copy()
            Dleft_prev = Dn.copy()
            for k in range(K):
                s = prof[sym, lanes, k]
                m = diag + s
                if local:
                    m = np.maximum(m, np.float32(0))
                    best = np.maximum(best, m)
                ue = U[:, k] + ge
                u = np.maximum(Mo[:, k], ue)
                le = Ll + ge
                l = np.maximum(Ml, le)
                diag = D[:, k].copy()
                ul = np.maximum(u, l)
                d = np.maximum(m, ul)
                if want_tb:
                    # four sign bits: 1 = the second operand won strictly (ties keep the priority)
                    not_m = (m < ul).astype(np.int64)
                    second = ((l < u) if transposed else (u < l)).astype(np.int64)
                    nib = not_m | (second << 1) | ((Mo[:, k] < ue).astype(np.int64) << 2) | ((Ml < le).astype(np.int64) << 3)
                    acc[:, k] = ((acc[:, k].astype(np.uint64) << 4) & 0xFFFFFFFF).astype(np.uint32) | nib.astype(np.uint32)
                mo = m + go
                Mo[:, k] = mo
                U[:, k] = u
                D[:, k] = d
                Ml = mo
                Ll = l
            Mo_last, L_last, D_last = Ml.copy(), Ll.copy(), D[:, K - 1].copy()
            if want_tb and (t & 7) == 7:
                tb[t >> 3] = acc.T
            if semi:
                # running max of the last real column (largest y wins ties)
                dl = D[lr, klast]
                if dl >= colbest[lr]:
                    colbest[lr] = dl
                    colbest_y[lr] = y[lr]
            last = (w & FLAG_LAST) != 0
            emit = (w & FLAG_EMIT) != 0
            for ln in np.nonzero(emit)[0]:
                p = q[ln]
                o = out[p]
                if semi:
                    # last-row max over my valid columns, largest x wins ties; lane 0 adds x = 0
                    for k in range(K):
                        x = ln * K + k + 1
                        if x <= L2 and D[ln, k] >= o["rowkey"][0] and (D[ln, k] > o["rowkey"][0] or x > o["rowkey"][1]):
                            o["rowkey"] = (D[ln, k], x)
                    if ln == 0:
                        v = leftD[y[0]]
                        if v > o["rowkey"][0]:
                            o["rowkey"] = (v, 0)
                    if ln == lr:
                        o["colkey"] = (colbest[lr], int(colbest_y[lr]))
                        o["emit_t"] = t
                        o["L1"] = int(y[lr])
                elif local:
                    o["score"] = max(o["score"], best[ln])
                    if ln == lr:
                        o["emit_t"] = t
                else:
                    if ln == lr:
                        o["score"] = D[ln, klast]
                        o["emit_t"] = t
                q[ln] += 1
            if last.any():
                reset_state(last, Mo, U, D, Dleft_prev, y, colbest)
                colbest_y[last] = 0
                best[last] = 0
    return out, tb, dict(T=T, lr=lr, klast=klast, B=B)


def fetch_nib(tb, K, emit_t, lr, L1, yk, xk):
    lane, k = (xk - 1) // K, (xk - 1) % K
    step = emit_t - (L1 - yk) - (lr - lane)
    w = int(tb[step >> 3, k, lane])
    return (w >> (4 * (7 - (step & 7)))) & 15


def traceback(tb, K, info, L1, L2, emit_t, start, transposed):
    """Pointer chase in kernel coordinates; returns the path in REFERENCE (y, x) order.
    start = (yk, xk, state) with kernel states 0 = M, 1 = from-above, 2 = from-left."""
    B, lr = info["B"], info["lr"]
    yk, xk, s = start
    path = []

    def code_at(y_, x_):
        if y_ == 0 and x_ == 0:
            return B["code00"]
        if y_ == 0:
            return 2
        if x_ == 0:
            return 1
        nb = fetch_nib(tb, K, emit_t, lr, L1, y_, x_)
        if not nb & 1:
            return 0
        if transposed:
            return 1 if nb & 2 else 2
        return 2 if nb & 2 else 1

    while True:
        path.append((xk, yk) if transposed else (yk, xk))
        if yk == 0 and xk == 0:
            break
        if xk == 0:
            if s == 1 and B["left_ramp"]:
                yk -= 1
                continue
            break
        if yk == 0:
            if s == 2 and B["top_ramp"]:
                xk -= 1
                continue
            break
        nib = fetch_nib(tb, K, emit_t, lr, L1, yk, xk)
        if s == 0:
            yk, xk = yk - 1, xk - 1
            s = code_at(yk, xk)
        elif s == 1:
            s = 1 if (nib >> 2) & 1 else 0
            yk -= 1
        else:
            s = 2 if (nib >> 3) & 1 else 0
            xk -= 1
    path.reverse()
    return path
