"""Host side of the B200 alignment core: tiling, device buffers, C-ABI calls.

PyTorch is used for device memory, streams and (in parallel.py) torch.distributed only; all
arithmetic of the hot path runs in libpraline_b200.so.  Mirrors what the reference does around
its C extension in praline/component/align.py (array prep :119-221, border init :357-385,
end-cell choice and traceback :401-433), but for whole batches of pairs at once.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib

MODES = {"global": 0, "local": 1, "semiglobal_both": 2, "semiglobal_one": 3, "semiglobal_two": 4}

TILE_DTYPE = np.dtype([("resident", np.int32), ("stream_begin", np.int32), ("stream_end", np.int32),
                       ("resident2", np.int32), ("out_base", np.int64), ("out_base2", np.int64),
                       ("b_skip", np.int32), ("reserved", np.int32)])


def _gaps(gap_series):
    """component/align.py:182-189: one value = linear, two = affine, more is an error."""
    gs = [float(g) for g in gap_series]
    if len(gs) == 1:
        return np.float32(gs[0]), np.float32(gs[0])
    if len(gs) == 2:
        return np.float32(gs[0]), np.float32(gs[1])
    raise ValueError("the fast aligner only supports linear and affine gap penalties at the moment")


def borders(mode, go, ge, maxlen, transposed=False):
    """max(M, U, L) along row 0 and column 0 in the kernel's orientation, built with the same
    numpy arithmetic as the reference (component/align.py:367-385): int64 arange times f32
    array is f64, plus an f32 scalar, cast to f32 on store."""
    go, ge = np.float32(go), np.float32(ge)
    ramp = np.empty(maxlen + 1, np.float32)
    ramp[1:] = np.arange(maxlen) * np.full(maxlen, ge, np.float32) + go
    u_zero = mode in (2, 3)
    l_zero = mode in (2, 4)
    u00 = np.float32(0) if u_zero else np.float32(go - ge)
    l00 = np.float32(0) if l_zero else np.float32(go - ge)
    col_u = np.zeros(maxlen + 1, np.float32) if u_zero else ramp.copy()
    row_l = np.zeros(maxlen + 1, np.float32) if l_zero else ramp.copy()
    vals = [np.float32(0), u00, l00]
    d00 = max(vals)
    code00 = int(np.argmax(vals))
    col_u[0] = d00
    row_l[0] = d00
    if not transposed:
        return dict(topD=row_l, leftD=col_u, code00=code00, top_ramp=int(not l_zero), left_ramp=int(not u_zero),
                    left0=0.0 if u_zero else float(go), left1=0.0 if u_zero else float(ge))
    return dict(topD=col_u, leftD=row_l, code00={0: 0, 1: 2, 2: 1}[code00], top_ramp=int(not u_zero),
                left_ramp=int(not l_zero), left0=0.0 if l_zero else float(go), left1=0.0 if l_zero else float(ge))


class SeqBatch(object):
    """A set of index sequences resident on the device (uint8 symbols + int64 offsets)."""

    def __init__(self, engine, seqs):
        self.lens = np.fromiter(map(len, seqs), np.int64, len(seqs))
        if len(seqs) == 0 or (self.lens <= 0).any():
            raise ValueError("empty sequences cannot be aligned")
        # (the packing is part of every end-to-end step of bench.py: one concatenate over the caller's arrays)
        flat = np.concatenate(seqs if all(type(s) is np.ndarray for s in seqs) else [np.asarray(s) for s in seqs])
        if flat.min() < 0 or flat.max() > 63:
            raise ValueError("symbol indices must lie in 0..63")
        self.n = len(seqs)
        self.offs = np.zeros(self.n + 1, np.int64)
        np.cumsum(self.lens, out=self.offs[1:])
        flat8 = flat if flat.dtype == np.uint8 else flat.astype(np.uint8)
        self.flat_host = torch.from_numpy(flat8).pin_memory() if engine.pin else torch.from_numpy(flat8)
        self.offs_host = torch.from_numpy(self.offs)
        self.flat_dev = self.flat_host.to(engine.device, non_blocking=True)
        self.offs_dev = self.offs_host.to(engine.device, non_blocking=True)
        self.max_sym = int(flat.max())

    @property
    def h2d_bytes(self):
        return self.flat_host.numel() + self.offs_host.numel() * 8


ROWBLOCK_DTYPE = np.dtype([("row0", np.int64), ("src0", np.int64), ("rows", np.int32), ("res", np.int32),
                           ("dummy", np.int32), ("reserved", np.int32)])


def plan_profile_wave(wt, nw, cs, lens_s, str_s, res_s, offs, rows_per_block=32):
    """Row layout of one wave of a profile batch.  Every (tile, warp) region holds one dummy row
    followed by the profile rows of its streamed sequences in stream order, so that the matrix
    row of a stream position IS that position.  The score-row kernels get the wave as blocks of
    <= 32 consecutive matrix rows of ONE streamed sequence (the region's dummy row rides in front
    of its first sequence).  Returns (first row per region, row blocks, number of rows)."""
    tb_, te_ = wt["stream_begin"].astype(np.int64), wt["stream_end"].astype(np.int64)
    per = (te_ - tb_ + nw - 1) // nw
    w = np.arange(nw)[None, :]
    sb = tb_[:, None] + w * per[:, None]
    se = np.minimum(sb + per[:, None], te_[:, None])
    sb = np.minimum(sb, se)
    reg_rows = np.where(se > sb, cs[se] - cs[sb] + 1, 0).ravel()      # +1: the dummy row
    mrow_base = np.zeros(reg_rows.size, np.int64)
    np.cumsum(reg_rows[:-1], out=mrow_base[1:])
    n_rows = int(reg_rows.sum())
    e_lo, e_hi = int(tb_[0]), int(te_[-1])
    el = lens_s[e_lo:e_hi]
    sbf = sb.ravel()
    reg_of = np.repeat(np.arange(reg_rows.size), (se - sb).ravel())   # region of every stream element
    first = (np.arange(e_lo, e_hi) == sbf[reg_of]).astype(np.int64)   # element opens its region
    # the element's rows in the matrix, the dummy row included for region openers
    e_row0 = mrow_base[reg_of] + 1 + (cs[e_lo:e_hi] - cs[sbf[reg_of]]) - first
    e_rows = el + first
    e_src0 = offs[str_s[e_lo:e_hi]] - first
    nblk = (e_rows + rows_per_block - 1) // rows_per_block
    eb = np.repeat(np.arange(len(el)), nblk)
    bfirst = np.cumsum(nblk) - nblk
    within = (np.arange(int(nblk.sum())) - bfirst[eb]) * rows_per_block
    blocks = np.zeros(len(eb), ROWBLOCK_DTYPE)
    blocks["row0"] = e_row0[eb] + within
    blocks["src0"] = e_src0[eb] + within
    blocks["rows"] = np.minimum(rows_per_block, e_rows[eb] - within)
    blocks["res"] = res_s[e_lo:e_hi][eb]
    blocks["dummy"] = (first[eb] == 1) & (within == 0)
    return mrow_base, blocks, n_rows


QUAD_DTYPE = np.dtype([("q0", np.int64), ("Lr", np.int32), ("nblk", np.int32), ("row0", np.int64, 4), ("src0", np.int64, 4),
                       ("rows", np.int32, 4), ("dummy", np.int32, 4), ("bcan", np.int64), ("reserved", np.int64),
                       ("can0", np.int64, 4)])


def row_block_quads(blocks, offs=None, padoff=None):
    """128-row tiles of the tensor-core score-row kernel (pgpu_quad): groups of <= 4 consecutive row
    blocks with one resident, each with the resident's first profile row and length.  Without `offs`
    returns just int32 [n_quads x 2] = (first block, number of blocks)."""
    res = np.ascontiguousarray(blocks["res"])
    n = len(res)
    if n == 0:
        return np.zeros((0, 2), np.int32) if offs is None else np.zeros(0, QUAD_DTYPE)
    new_run = np.ones(n, bool)
    new_run[1:] = res[1:] != res[:-1]
    run_start = np.maximum.accumulate(np.where(new_run, np.arange(n), 0))
    first = np.flatnonzero((np.arange(n) - run_start) % 4 == 0)
    count = np.diff(np.append(first, n))
    if offs is None:
        return np.stack([first, count], axis=1).astype(np.int32)
    # field views of a structured array are strided: gather from contiguous copies, build every quad field as a
    # plain 2-D array and assign each field once (this runs per wave of a profile batch, on the host)
    q = np.zeros(len(first), QUAD_DTYPE)
    r = res[first]
    q["q0"] = offs[r]
    q["Lr"] = offs[r + 1] - offs[r]
    q["nblk"] = count
    idx = np.minimum(first[:, None] + np.arange(4)[None, :], n - 1)
    ok = np.arange(4)[None, :] < count[:, None]
    src0 = np.ascontiguousarray(blocks["src0"], np.int64)
    dummy = np.ascontiguousarray(blocks["dummy"])
    q["row0"] = np.where(ok, np.ascontiguousarray(blocks["row0"])[idx], 0)
    q["src0"] = np.where(ok, src0[idx], 0)
    q["rows"] = np.where(ok, np.ascontiguousarray(blocks["rows"])[idx], 0)
    q["dummy"] = np.where(ok, dummy[idx], 0)
    if padoff is not None:
        q["bcan"] = padoff[r]
        # streamed side: a block whose first profile row sits on an 8-row group of its sequence's pre-split
        # rows (every block but the ones that carry a region's dummy row in front) can be fetched by TMA
        seq = np.clip(np.searchsorted(offs, src0, side="right") - 1, 0, len(offs) - 2)
        rel = src0 - offs[seq]
        can = np.where((dummy == 0) & (rel >= 0) & (rel % 8 == 0), padoff[seq] + rel, -1)
        q["can0"] = np.where(ok, can[idx], -1)
    else:
        q["can0"] = -1
    return q


def dense_symbol_count(sym_used, nnz_total, rows, A):
    """Number of symbols with a nonzero entry anywhere in a profile batch when its rows are dense enough for the
    symbols-in-use kernel (score_rows_x2.cu) to pay -- its work per cell is symbols_in_use x entries of the streamed
    row, against entries x entries of the compacted walk of k_build_rows_t --, else 0."""
    u = int(np.count_nonzero(sym_used))
    if A > 32 or u < 1 or rows < 1:
        return 0
    return u if nnz_total >= 0.55 * u * rows else 0


def plan_profile_wave_native(wt, nw, cs, lens_s, str_s, res_s, offs, want_quads=False, padoff=None, rows_per_block=32):
    """plan_profile_wave (+ row_block_quads) in one pass of the library's host planner (csrc/host_plan.cu,
    pgpu_plan_profile_wave): the same (first row per region, row blocks, number of rows[, quads]) -- the numpy
    functions above are its specification (tests/test_host_cpu.py compares them)."""
    lib = _lib.load()
    tb = np.ascontiguousarray(wt["stream_begin"], np.int64)
    te = np.ascontiguousarray(wt["stream_end"], np.int64)
    arrs = [np.ascontiguousarray(a, np.int64) for a in (cs, lens_s, str_s, res_s, offs)]
    n_el = int(te[-1] - tb[0])
    rows = int(arrs[0][te[-1]] - arrs[0][tb[0]])
    cap = n_el + (rows + n_el) // rows_per_block + 1
    blocks = np.empty(cap, ROWBLOCK_DTYPE)
    quads = np.empty(cap if want_quads else 0, QUAD_DTYPE)
    mrow_base = np.empty(len(tb) * nw, np.int64)
    n_rows = ctypes.c_int64(0)
    n_quads = ctypes.c_longlong(0)
    pad = np.ascontiguousarray(padoff, np.int64) if padoff is not None else None
    nb = lib.pgpu_plan_profile_wave(len(tb), tb.ctypes.data, te.ctypes.data, int(nw), arrs[0].ctypes.data, arrs[1].ctypes.data,
                                    arrs[2].ctypes.data, arrs[3].ctypes.data, arrs[4].ctypes.data, int(rows_per_block),
                                    mrow_base.ctypes.data, blocks.ctypes.data, cap, ctypes.byref(n_rows), int(want_quads),
                                    pad.ctypes.data if pad is not None else None, quads.ctypes.data if want_quads else None,
                                    cap, ctypes.byref(n_quads))
    if nb < 0:
        raise _lib.PralineGpuError("libpraline_b200: %s" % lib.pgpu_last_error().decode())
    if want_quads:
        return mrow_base, blocks[:nb], int(n_rows.value), quads[:n_quads.value]
    return mrow_base, blocks[:nb], int(n_rows.value)


class ProfileBatch(object):
    """A set of f32 profiles [L x A] resident on the device: one [rows x A] array plus int64 row
    offsets (the analogue of SeqBatch for ProfileTrack inputs, component/align.py:171-172)."""

    def __init__(self, engine, profiles):
        profiles = [np.ascontiguousarray(p, np.float32) for p in profiles]
        self.lens = np.asarray([p.shape[0] for p in profiles], np.int64)
        if len(profiles) == 0 or (self.lens <= 0).any():
            raise ValueError("empty sequences cannot be aligned")
        self.A = int(profiles[0].shape[1])
        if any(p.shape[1] != self.A for p in profiles):
            raise ValueError("profiles of one batch must share an alphabet")
        self.n = len(profiles)
        self.offs = np.zeros(self.n + 1, np.int64)
        np.cumsum(self.lens, out=self.offs[1:])
        flat = np.concatenate(profiles, axis=0)
        self.prof_dev = engine.dev(flat)
        self.offs_dev = engine.dev(self.offs)
        self.flat_dev = None
        self.max_sym = self.A - 1
        # symbols in use and entries per row: dense batches get the packed f32x2 score rows (score_rows_x2.cu)
        nzm = flat != 0
        self.sym_used = nzm.any(axis=0)
        self.nnz_total = int(nzm.sum())

    def dense_syms(self):
        """Number of symbols in use when the batch is dense enough for the packed f32x2 score rows, else 0."""
        return dense_symbol_count(self.sym_used, self.nnz_total, int(self.offs[-1]), self.A)


class GrowingProfileBatch(ProfileBatch):
    """ProfileBatch with room to append: the ad-hoc MSA adds one merged profile per round and
    never needs the old ones moved (dead profiles simply stay behind)."""

    def __init__(self, engine, profiles, cap_rows=0):
        ProfileBatch.__init__(self, engine, profiles)
        self._engine = engine
        rows = int(self.offs[-1])
        cap = max(int(cap_rows), rows)
        self._store = torch.empty((cap, self.A), dtype=torch.float32, device=engine.device)
        self._store[:rows].copy_(self.prof_dev)
        self.prof_dev = self._store[:rows]

    def append(self, profile):
        """Adds one [L x A] f32 profile; returns its index."""
        p = np.ascontiguousarray(profile, np.float32)
        if p.ndim != 2 or p.shape[1] != self.A or p.shape[0] < 1:
            raise ValueError("profile does not fit this batch")
        r0 = int(self.offs[-1])
        r1 = r0 + p.shape[0]
        if r1 > self._store.shape[0]:
            bigger = torch.empty((max(2 * self._store.shape[0], r1), self.A), dtype=torch.float32, device=self._store.device)
            bigger[:r0].copy_(self._store[:r0])
            self._store = bigger
        self._store[r0:r1].copy_(torch.from_numpy(p), non_blocking=False)
        self.prof_dev = self._store[:r1]
        nzm = p != 0
        self.sym_used = self.sym_used | nzm.any(axis=0)
        self.nnz_total += int(nzm.sum())
        self.lens = np.append(self.lens, p.shape[0])
        self.offs = np.append(self.offs, r1)
        self.offs_dev = self._engine.dev(self.offs)
        self.n += 1
        return self.n - 1


def tb_words16r(tiles, K, cs, nw):
    """Traceback words per (tile, warp) of the paired-resident traced kernel: four steps per word row of
    32 x (K | 1) words, T = the warp's stream rows + dummy row + pipeline drain, rounded to 32
    (gotoh_stream16r.cuh)."""
    tb_, te_ = tiles["stream_begin"].astype(np.int64), tiles["stream_end"].astype(np.int64)
    per = (te_ - tb_ + nw - 1) // nw
    w = np.arange(nw)[None, :]
    sb = tb_[:, None] + w * per[:, None]
    se = np.minimum(sb + per[:, None], te_[:, None])
    sb = np.minimum(sb, se)
    rows = cs[se] - cs[sb]
    T = np.where(se > sb, (rows + 1 + 31 + 31) // 32 * 32, 0)
    return (T // 4) * ((K | 1) * 32)


class DualWavePlan(object):
    """Host plan of the symmetric traced all-vs-all (Engine.allpairs_dual), pure numpy (tested on the CPU).

    Units = resident pairs (2p, 2p+1) with their stream j = 2p+1 .. n-1, as Engine.allpairs_tiles(paired=True); the
    shard (rank, world) is a contiguous unit range of equal DP cells.  waves() yields launches of at most `budget`
    traceback words, one columns-per-lane class each, planned `chunk` units at a time so that the host plans the next
    chunk while the device fills this one (393,000 tiles at BASELINE config 3 take 0.5 s in one go).  Slots are
    wave-local and the two residents of a tile interleave: A of stream element e at out_base + 2 e, B at
    out_base + 2 e + 1 -- the four walkers of a stream element (2 residents x 2 orientations) are neighbouring threads
    and share the sectors they read; an element without a B pair leaves a hole (emit_t stays -1)."""

    def __init__(self, lens, kcls, nw, tile, budget, shard=(0, 1), chunk=24):
        self.lens = lens = np.asarray(lens, np.int64)
        self.kcls, self.nw, self.tile, self.budget, self.chunk = np.asarray(kcls), int(nw), int(tile), int(budget), max(1, int(chunk))
        n = self.n = len(lens)
        cs = self.cs = np.zeros(n + 1, np.int64)
        np.cumsum(lens, out=cs[1:])
        rank, world = shard
        ui = self.ui = np.arange(0, n - 1, 2, dtype=np.int64)
        uhb = self.uhb = ui + 1 <= n - 2
        ucells = lens[ui] * (cs[n] - cs[ui + 1]) + np.where(uhb, lens[np.minimum(ui + 1, n - 1)] * (cs[n] - cs[np.minimum(ui + 2, n)]), 0)
        ucum = self.ucum = np.concatenate([[0], np.cumsum(ucells)])
        ucuts = np.searchsorted(ucum, ucum[-1] * np.arange(world + 1) / world, side="left")
        ucuts[0], ucuts[-1] = 0, len(ui)
        ucuts = np.maximum.accumulate(ucuts)
        self.u_lo, self.u_hi = int(ucuts[rank]), int(ucuts[rank + 1])
        self.cells = int(ucum[self.u_hi] - ucum[self.u_lo])

    def unit_tiles(self, u0, u1):
        """Tile records of units [u0, u1), grouped by columns-per-lane class."""
        n, tile, kcls = self.n, self.tile, self.kcls
        gi_u, hb_u = self.ui[u0:u1], self.uhb[u0:u1]
        cnt = n - 1 - gi_u
        ntile = (cnt + tile - 1) // tile
        unit = np.repeat(np.arange(len(gi_u)), ntile)
        first = np.cumsum(ntile) - ntile
        gi = gi_u[unit]
        tb_ = gi + 1 + (np.arange(int(ntile.sum())) - first[unit]) * tile
        te_ = np.minimum(tb_ + tile, n)
        hb = hb_u[unit]
        t = np.zeros(len(tb_), TILE_DTYPE)
        t["resident"] = gi
        t["resident2"] = np.where(hb, gi + 1, -1)
        t["stream_begin"] = tb_
        t["stream_end"] = te_
        t["b_skip"] = np.where(hb & (tb_ == gi + 1), 1, 0)
        kk = np.maximum(kcls[gi], np.where(hb, kcls[np.minimum(gi + 1, n - 1)], 0))
        return {int(K): t[kk == K] for K in np.unique(kk)}

    def waves(self):
        """-> (K, tile records, word offset per (tile, warp), slots, words, stream elements per tile for A, for B)."""
        pending = {}
        budget = self.budget
        for u0 in list(range(self.u_lo, self.u_hi, self.chunk)) + [None]:
            if u0 is not None:
                for K, t in self.unit_tiles(u0, min(u0 + self.chunk, self.u_hi)).items():
                    w = tb_words16r(t, K, self.cs, self.nw)
                    if K in pending:
                        pending[K] = (np.concatenate([pending[K][0], t]), np.concatenate([pending[K][1], w]))
                    else:
                        pending[K] = (t, w)
            for K in list(pending):
                tiles, words = pending[K]
                wcum = np.concatenate([[0], np.cumsum(words.sum(axis=1))])
                lo = 0
                while lo < len(tiles):
                    hi = max(lo + 1, int(np.searchsorted(wcum, wcum[lo] + budget, side="right")) - 1)
                    if hi >= len(tiles) and u0 is not None and wcum[-1] - wcum[lo] < budget:
                        break                      # a partial wave: wait for the next chunk's tiles
                    hi = min(hi, len(tiles))
                    wt = tiles[lo:hi].copy()
                    na = (wt["stream_end"] - wt["stream_begin"]).astype(np.int64)
                    nb = np.where(wt["resident2"] >= 0, na - wt["b_skip"], 0)
                    base = np.zeros(len(wt), np.int64)
                    np.cumsum((2 * na)[:-1], out=base[1:])
                    wt["out_base"] = base
                    wt["out_base2"] = base + 1 + 2 * wt["b_skip"]
                    wt["reserved"] = 2                      # PgTile.slot_stride
                    wbase = np.zeros(words[lo:hi].size, np.int64)
                    np.cumsum(words[lo:hi].ravel()[:-1], out=wbase[1:])
                    yield K, wt, wbase, int(2 * na.sum()), int(wcum[hi] - wcum[lo]), na, nb
                    lo = hi
                if lo >= len(tiles):
                    del pending[K]
                else:
                    pending[K] = (tiles[lo:], words[lo:])


class Engine(object):
    def __init__(self, device=0, pin=True):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.PralineGpuError("no CUDA device visible: praline_b200 has no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        _lib.check(self.lib.pgpu_init(device))
        self.pin = pin
        self.nw = self.lib.pgpu_warps_per_tile()
        self.k_set = [k for k in range(1, 65) if self.lib.pgpu_supported_k(k)]
        self.launches = 0          # kernels of ours launched (bench.py reports it)
        self._stage = None
        self.trace_on = os.environ.get("PGPU_TRACE", "") not in ("", "0")
        self._traces = []
        # traceback words per wave: 16 GiB of the 180 GB (more pairs per wave = more walkers in flight in K4)
        self.tb_budget_words = int(float(os.environ.get("PGPU_TB_GIB", "16")) * (1 << 28))
        self._borders = {}
        self.use_s16 = os.environ.get("PGPU_NO_S16", "") == ""
        self.m_budget_floats = 1 << 31     # 8 GiB of match scores per wave of a profile batch
        self.keep_mwave = False
        self.tc_tma = os.environ.get("PGPU_TC_TMA", "1") not in ("", "0")      # B tiles by TMA bulk copy
        self.fast_tc = os.environ.get("PGPU_FAST_TC", "1") not in ("", "0")   # tolerance-mode score rows on tcgen05
        self.rows_x2 = os.environ.get("PGPU_ROWS_X2", "1") not in ("", "0")   # exact score rows of dense batches: packed f32x2

    # -- helpers -------------------------------------------------------------------------------
    def _trace_event(self, name, chain=None):
        """PGPU_TRACE=1: CUDA events between the C-ABI calls of a pass; dump_trace() prints the device
        time of every section.  A no-op otherwise."""
        if not self.trace_on:
            return None
        chain = chain if chain is not None else []
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        chain.append((name, e))
        if name is None:
            self._traces.append(chain)
        return chain

    def take_trace(self):
        """[(section name, device ms)] of everything traced since the last call (PGPU_TRACE / trace_on)."""
        torch.cuda.synchronize()
        out = [(n0, e0.elapsed_time(e1)) for chain in self._traces
               for (n0, e0), (_, e1) in zip(chain[:-1], chain[1:])]
        self._traces = []
        return out

    def dump_trace(self):
        for name, ms in self.take_trace():
            print("  [trace] %-40s %8.3f ms" % (name, ms))

    def stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # (slot bytes, slots) of the staging rings; pinning a slot costs about 0.3 ms per MB, once
    STAGE_TIERS = ((256 << 10, 8), (2 << 20, 4), (8 << 20, 4), (32 << 20, 2))

    def dev(self, arr):
        """Host array -> device tensor, asynchronously on the current stream.  Arrays between 16 KB and
        32 MB go through rings of persistent pinned staging buffers (a fresh pin_memory() costs
        milliseconds per call, a pageable copy makes the host wait for the stream); slots are pinned on
        first use and reused only after the copy that read them has completed."""
        a = np.ascontiguousarray(arr)
        t = torch.from_numpy(a)
        nbytes = a.nbytes
        if not self.pin or nbytes < (1 << 14):
            return t.to(self.device, non_blocking=True)
        tier = next((k for k, (size, _) in enumerate(self.STAGE_TIERS) if nbytes <= size), None)
        if tier is None:
            return t.pin_memory().to(self.device, non_blocking=True)
        if self._stage is None:
            self._stage = [[[None, None] for _ in range(slots)] for _, slots in self.STAGE_TIERS]
            self._stage_next = [0] * len(self.STAGE_TIERS)
        ring = self._stage[tier]
        slot = ring[self._stage_next[tier]]
        self._stage_next[tier] = (self._stage_next[tier] + 1) % len(ring)
        if slot[0] is None:
            slot[0] = torch.empty(self.STAGE_TIERS[tier][0], dtype=torch.uint8).pin_memory()
        if slot[1] is not None:
            slot[1].synchronize()
        src = slot[0][:nbytes]
        src.numpy()[:] = a.reshape(-1).view(np.uint8)
        out = torch.empty(a.shape, dtype=t.dtype, device=self.device)
        out.view(torch.uint8).reshape(-1).copy_(src, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        slot[1] = ev
        return out

    @staticmethod
    def ptr(t):
        return ctypes.c_void_p(t.data_ptr()) if t is not None else None

    def k_for(self, length):
        for k in self.k_set:
            if 32 * k >= length:
                return k
        return None

    def k_classes(self, lens):
        """k_for over an array of lengths: columns-per-lane class per sequence, -1 beyond the limit."""
        ks = np.asarray(self.k_set, np.int64)
        pos = np.searchsorted(32 * ks, np.asarray(lens, np.int64), side="left")
        return np.where(pos < len(ks), ks[np.minimum(pos, len(ks) - 1)], -1)

    def batch(self, seqs):
        return SeqBatch(self, seqs)

    def profile_batch(self, profiles, cap_rows=None):
        """cap_rows: build a batch that profiles can be appended to (GrowingProfileBatch)."""
        if cap_rows is not None:
            return GrowingProfileBatch(self, profiles, cap_rows)
        return ProfileBatch(self, profiles)

    # -- inter-task batch ----------------------------------------------------------------------
    def _make_tiles(self, res_sorted, n_stream_total, tile):
        """res_sorted: resident id per stream element, grouped.  -> structured tile array whose
        out_base equals stream_begin (slot == stream element)."""
        n = n_stream_total
        change = np.flatnonzero(np.diff(res_sorted)) + 1
        starts = np.concatenate([[0], change]).astype(np.int64)
        ends = np.concatenate([change, [n]]).astype(np.int64)
        ntile = (ends - starts + tile - 1) // tile
        grp = np.repeat(np.arange(len(starts)), ntile)
        first = np.cumsum(ntile) - ntile
        tb = starts[grp] + (np.arange(int(ntile.sum())) - first[grp]) * tile
        te = np.minimum(tb + tile, ends[grp])
        tiles = np.zeros(len(tb), TILE_DTYPE)
        tiles["resident"] = res_sorted[tb]
        tiles["stream_begin"] = tb
        tiles["stream_end"] = te
        tiles["out_base"] = tb
        return tiles

    def _tb_words(self, tiles, K, cs):
        """Traceback words per (tile, warp); cs = prefix sum of stream-element lengths."""
        nw = self.nw
        tb_, te_ = tiles["stream_begin"].astype(np.int64), tiles["stream_end"].astype(np.int64)
        per = (te_ - tb_ + nw - 1) // nw
        w = np.arange(nw)[None, :]
        sb = tb_[:, None] + w * per[:, None]
        se = np.minimum(sb + per[:, None], te_[:, None])
        sb = np.minimum(sb, se)
        rows = cs[se] - cs[sb]
        T = np.where(se > sb, (rows + 1 + 31 + 31) // 32 * 32, 0)
        return (T // 8) * (K * 32)

    def _tb_words16(self, tiles, K, cs):
        """Traceback words per (tile, warp) of the packed traced kernel: a warp cuts its slice into two
        halves that run in lock step, four steps per word."""
        nw = self.nw
        tb_, te_ = tiles["stream_begin"].astype(np.int64), tiles["stream_end"].astype(np.int64)
        per = (te_ - tb_ + nw - 1) // nw
        w = np.arange(nw)[None, :]
        sb = tb_[:, None] + w * per[:, None]
        se = np.minimum(sb + per[:, None], te_[:, None])
        sb = np.minimum(sb, se)
        mid = sb + (se - sb + 1) // 2
        rows = np.maximum(cs[mid] - cs[sb], cs[se] - cs[mid])
        T = np.where(se > sb, (rows + 1 + 31 + 31) // 32 * 32, 0)
        return (T // 4) * (K * 32)

    def _pick_tile(self, n_pairs):
        # enough tiles to fill 148 SMs several times over, but streams of >= 2 sequences per warp
        if os.environ.get("PGPU_TILE"):
            return int(os.environ["PGPU_TILE"])
        spw = 16
        while spw > 2 and n_pairs // (self.nw * spw) < 148 * 6:
            spw //= 2
        return self.nw * spw

    def run_tiles(self, mode, K, transposed, batch, stream_ids_dev, tiles, n_slots, S_dev, A, go, ge,
                  scores_dev, cs=None, slot_res_dev=None, slot_str_dev=None, want_paths=False, caps=None,
                  tiles_dev=None, mwave_dev=None, mrow_base_dev=None, counts=None, S_host=None, paired=False,
                  out_shift=0):
        """Launch K2 (+K4) for one K class.  Returns list of (slot_lo, slot_hi, path_off, path_buf,
        path_start, path_len) per wave when want_paths.  out_shift (score-only launches): slot s is
        written at scores_dev[s + out_shift] (the rank slices of parallel.ShardedCondensed)."""
        lib = self.lib
        scores_ptr = ctypes.c_void_p(scores_dev.data_ptr() + 4 * int(out_shift)) if out_shift else self.ptr(scores_dev)
        md = MODES[mode] if isinstance(mode, str) else mode
        maxlen = max(32 * K, int(batch.lens.max())) + 2
        bkey = (md, float(go), float(ge), maxlen, bool(transposed))
        if bkey not in self._borders:
            B = borders(md, go, ge, maxlen, transposed)
            self._borders[bkey] = (B, self.dev(B["topD"]), self.dev(B["leftD"]))
        B, top_dev, left_dev = self._borders[bkey]
        semi = md != 0   # local and semiglobal results are reduced through the keys scratch
        out = []
        if not want_paths:
            if tiles_dev is None:
                tiles_dev = self.dev(tiles.view(np.uint8))
            neg = self.fits_s16(S_host, go, ge, batch.lens) if (
                md == 0 and mwave_dev is None and self.use_s16) else None
            if neg is not None:   # packed 16-bit DPX kernel: two streamed sequences per warp
                _lib.check(lib.pgpu_align_tiles16(K, int(paired), int(transposed), self.ptr(batch.flat_dev), self.ptr(batch.offs_dev),
                                                  self.ptr(stream_ids_dev), self.ptr(tiles_dev), len(tiles),
                                                  self.ptr(S_dev), A, int(go), int(ge), neg, self.ptr(top_dev),
                                                  int(B["left0"]), int(B["left1"]), maxlen + 1,
                                                  scores_ptr, self.stream()))
                self.launches += 1
                return out
            keys = torch.empty(2 * n_slots, dtype=torch.int64, device=self.device) if semi else None
            _lib.check(lib.pgpu_align_tiles(md, K, int(transposed), self.ptr(batch.flat_dev), self.ptr(batch.offs_dev),
                                            self.ptr(stream_ids_dev), self.ptr(tiles_dev), len(tiles), n_slots,
                                            self.ptr(S_dev), A, float(go), float(ge), self.ptr(top_dev),
                                            self.ptr(left_dev), B["left0"], B["left1"], maxlen + 1,
                                            scores_ptr, self.ptr(keys),
                                            None, None, None, None, self.ptr(mwave_dev), self.ptr(mrow_base_dev),
                                            self.stream()))
            self.launches += 1 + int(semi)
            return out
        # traced: waves bounded by the traceback budget; tiles are in slot order
        neg16 = self.fits_s16(S_host, go, ge, batch.lens, limit=16000) if (
            md == 0 and mwave_dev is None and self.use_s16) else None
        fmt = 1 if neg16 is not None else 0
        self.last_traced_fmt = fmt      # 1: packed int16 traced kernel, 0: f32 (tests look at it)
        words = self._tb_words16(tiles, K, cs) if fmt else self._tb_words(tiles, K, cs)
        per_tile = words.sum(axis=1)
        lo = 0
        while lo < len(tiles):
            acc = np.cumsum(per_tile[lo:])
            hi = lo + max(1, int(np.searchsorted(acc, self.tb_budget_words, side="right")))
            wt = tiles[lo:hi].copy()
            s_lo, s_hi = int(wt["stream_begin"][0]), int(wt["stream_end"][-1])
            ns = s_hi - s_lo
            wt["out_base"] -= s_lo
            wbase = np.zeros(words[lo:hi].size, np.int64)
            np.cumsum(words[lo:hi].ravel()[:-1], out=wbase[1:])
            tb = torch.empty(int(words[lo:hi].sum()), dtype=torch.int32, device=self.device)
            tiles_dev, wbase_dev = self.dev(wt.view(np.uint8)), self.dev(wbase)
            emit_t = torch.empty(ns, dtype=torch.int32, device=self.device)
            pair_tb = torch.empty(ns, dtype=torch.int64, device=self.device)
            keys = torch.empty(2 * ns, dtype=torch.int64, device=self.device) if semi else None
            sc = scores_dev[s_lo:s_hi]
            if fmt:     # packed int16 traced kernel (global mode, integer scores inside +-16000)
                _lib.check(lib.pgpu_align_tiles16_traced(K, int(transposed), self.ptr(batch.flat_dev), self.ptr(batch.offs_dev),
                                                         self.ptr(stream_ids_dev), self.ptr(tiles_dev), len(wt),
                                                         self.ptr(S_dev), A, int(go), int(ge), neg16, self.ptr(top_dev),
                                                         int(B["left0"]), int(B["left1"]), maxlen + 1, self.ptr(sc),
                                                         self.ptr(tb), self.ptr(wbase_dev), self.ptr(emit_t),
                                                         self.ptr(pair_tb), self.stream()))
            else:
                _lib.check(lib.pgpu_align_tiles(md, K, int(transposed), self.ptr(batch.flat_dev), self.ptr(batch.offs_dev),
                                                self.ptr(stream_ids_dev), self.ptr(tiles_dev), len(wt), ns,
                                                self.ptr(S_dev), A, float(go), float(ge), self.ptr(top_dev),
                                                self.ptr(left_dev), B["left0"], B["left1"], maxlen + 1, self.ptr(sc),
                                                self.ptr(keys),
                                                self.ptr(tb), self.ptr(wbase_dev), self.ptr(emit_t), self.ptr(pair_tb),
                                                None, None, self.stream()))
            if counts is not None:
                # preprofile mode: the walk adds into the masters' count tables, no path leaves the device
                cnt_dev, cnt_off_dev, thr = counts
                _lib.check(lib.pgpu_traceback_tiles(md, K, int(transposed), self.ptr(batch.offs_dev),
                                                    self.ptr(slot_res_dev[s_lo:s_hi]), self.ptr(slot_str_dev[s_lo:s_hi]),
                                                    ns, self.ptr(keys), self.ptr(tb), self.ptr(emit_t), self.ptr(pair_tb),
                                                    B["code00"], B["top_ramp"], B["left_ramp"], None, None, None, None,
                                                    self.ptr(batch.flat_dev), self.ptr(cnt_dev),
                                                    self.ptr(cnt_off_dev[s_lo:s_hi]), A, self.ptr(sc),
                                                    int(thr is not None), float(thr if thr is not None else 0.0),
                                                    fmt, self.stream()))
                self.launches += 2 + int(semi)
                lo = hi
                continue
            cap = caps[s_lo:s_hi]
            poff = np.zeros(ns, np.int64)
            np.cumsum(cap[:-1], out=poff[1:])
            poff_dev = self.dev(poff)
            pbuf = torch.empty((int(cap.sum()), 2), dtype=torch.int32, device=self.device)
            pstart = torch.empty(ns, dtype=torch.int32, device=self.device)
            plen = torch.empty(ns, dtype=torch.int32, device=self.device)
            _lib.check(lib.pgpu_traceback_tiles(md, K, int(transposed), self.ptr(batch.offs_dev),
                                                self.ptr(slot_res_dev[s_lo:s_hi]), self.ptr(slot_str_dev[s_lo:s_hi]),
                                                ns, self.ptr(keys), self.ptr(tb), self.ptr(emit_t), self.ptr(pair_tb),
                                                B["code00"], B["top_ramp"], B["left_ramp"], self.ptr(poff_dev),
                                                self.ptr(pbuf), self.ptr(pstart), self.ptr(plen),
                                                None, None, None, A, None, 0, 0.0, fmt, self.stream()))
            self.launches += 2 + int(semi)
            out.append((s_lo, s_hi, poff, pbuf, pstart, plen))
            lo = hi
        return out

    def fits_s16(self, S, go, ge, lens, limit=32000):
        """Sentinel for the packed int16 kernel, or None when the batch must stay in f32: integer
        scores, and |go| + (|ge| + max|S|) * (L1 + L2) plus the sentinel arithmetic inside int16
        (limit 32000), or inside +-16000 for the traced variant whose tie tests subtract two values."""
        if S is None:
            return None
        vals = np.concatenate([S.ravel().astype(np.float64), [float(go), float(ge)]])
        if not np.all(vals == np.round(vals)):
            return None
        smax = float(np.abs(S).max())
        lmax = int(np.max(lens))
        v = abs(float(go)) + (abs(float(ge)) + smax) * 2 * lmax
        neg = -(v + smax + 1)
        if neg - abs(float(go)) - 2 * abs(float(ge)) - smax < -limit or smax * lmax > limit:
            return None
        return int(neg)

    def integer_exact(self, S, go, ge, maxlen):
        """Traced batches derive tie flags from unrounded operands: exact iff all scores are
        integers small enough that every f32 sum is exact (see gotoh_stream.cu)."""
        S = np.asarray(S, np.float64)
        vals = np.concatenate([S.ravel(), [float(go), float(ge)]])
        if not np.all(vals == np.round(vals)):
            return False
        bound = np.abs(S).max() * maxlen + abs(float(go)) + abs(float(ge)) * 2 * maxlen
        return bound < 2 ** 23

    def align_pairs(self, batch, pi, pj, S, gap_series, mode="global", want_paths=False, resident=None,
                    device_only=False, counts=None):
        """Scores (and reference-format paths) of pairs (sequence_one = pi[k], sequence_two = pj[k]).

        Returns np.float32 scores [n] and, if want_paths, a list of int32 [rows, 2] arrays."""
        md = MODES[mode]
        go, ge = _gaps(gap_series)
        pi = np.asarray(pi, np.int64)
        pj = np.asarray(pj, np.int64)
        n = len(pi)
        S = np.ascontiguousarray(S, np.float32)
        A = S.shape[0]
        if batch.max_sym >= A:
            raise ValueError("sequence symbol outside the score matrix")
        if n == 0:
            return np.zeros(0, np.float32), ([] if want_paths else None)
        if md == 1 and not (go <= ge <= 0):
            raise _lib.PralineGpuError("batched local scores need open <= extend <= 0 (the border cell open - extend takes "
                                       "part in the reference's argmax): use align_general")
        if want_paths:
            if md == 1:
                raise _lib.PralineGpuError("batched local traceback is served by the general kernel")
            if not self.integer_exact(S, go, ge, int(batch.lens.max())):
                raise _lib.PralineGpuError("non-integer scores: use align_general for traced alignments")
        if resident is None:  # share the side with fewer distinct sequences
            resident = "one" if len(np.unique(pi)) < len(np.unique(pj)) else "two"
        transposed = resident == "one"
        res, strm = (pi, pj) if transposed else (pj, pi)
        kcls = self.k_classes(batch.lens)
        if (kcls[res] < 0).any():
            raise _lib.PralineGpuError("resident sequence longer than %d: use the general kernel" % (32 * self.k_set[-1]))
        kres = kcls[res]
        if n < 2 or ((res[1:] >= res[:-1]).all() and (kres[1:] >= kres[:-1]).all()):
            order = np.arange(n)            # already grouped (master-slave lists are): no sort
            res_s, str_s = res, strm
        else:
            order = np.lexsort((np.arange(n), res, kres))
            res_s, str_s = res[order], strm[order]
        S_dev = self.dev(S)
        stream_ids_dev = self.dev(str_s.astype(np.int32))
        scores_dev = torch.empty(n, dtype=torch.float32, device=self.device)
        cs = np.zeros(n + 1, np.int64)
        np.cumsum(batch.lens[str_s], out=cs[1:])
        caps = (batch.lens[res_s] + batch.lens[str_s] + 2).astype(np.int64)
        slot_res_dev = self.dev(res_s.astype(np.int32)) if want_paths else None
        paths_sorted = [None] * n if want_paths else None
        counts_ctx = None
        if counts is not None:   # (count tables on the device, offset of the master's table per pair, threshold)
            cnt_dev, cnt_off, thr = counts
            counts_ctx = (cnt_dev, self.dev(np.asarray(cnt_off, np.int64)[order]), thr)
        kk = kcls[res_s]
        bounds = np.flatnonzero(np.diff(kk)) + 1
        tile = self._pick_tile(n)
        pending = []
        for a, b in zip(np.concatenate([[0], bounds]), np.concatenate([bounds, [n]])):
            K = int(kk[a])
            tiles = self._make_tiles(res_s[a:b], b - a, tile)
            tiles["stream_begin"] += a
            tiles["stream_end"] += a
            tiles["out_base"] += a
            waves = self.run_tiles(md, K, transposed, batch, stream_ids_dev, tiles, n, S_dev, A, go, ge,
                                   scores_dev, cs=cs, slot_res_dev=slot_res_dev, slot_str_dev=stream_ids_dev,
                                   want_paths=want_paths, caps=caps, counts=counts_ctx, S_host=S)
            pending.extend(waves)
        if device_only:
            return scores_dev, order, pending
        scores_sorted = scores_dev.cpu().numpy()
        scores = np.empty(n, np.float32)
        scores[order] = scores_sorted
        if not want_paths:
            return scores, None
        for (s_lo, s_hi, poff, pbuf, pstart, plen) in pending:
            pb, ps, pl = pbuf.cpu().numpy(), pstart.cpu().numpy(), plen.cpu().numpy()
            for k in range(s_hi - s_lo):
                o = poff[k] + ps[k]
                paths_sorted[s_lo + k] = pb[o:o + pl[k]]
        paths = [None] * n
        for k, o in enumerate(order):
            paths[o] = paths_sorted[k]
        return scores, paths

    def preprofile_counts(self, batch, masters, slaves, S, gap_series, threshold=None):
        """Count tables of global master-slave preprofiles, entirely on the device.

        masters / slaves: sequence ids of every (master, slave) pair (sequence_one = master,
        preprofile.py:131-137).  Returns (counts int64 [sum L_master x A] on the host, offsets per
        distinct master id, scores per pair).  Replaces N(N-1) x (PairwiseAligner + compress_path +
        Alignment.merge) and the per-master get_frequencies of ProfileBuilder (profile.py:56)."""
        masters = np.asarray(masters, np.int64)
        slaves = np.asarray(slaves, np.int64)
        S = np.ascontiguousarray(S, np.float32)
        A = S.shape[0]
        cnt, off, uniq, lens, slot_of = self._own_counts(batch, masters, A)
        scores_dev, order, _ = self.align_pairs(batch, masters, slaves, S, gap_series, mode="global", want_paths=True,
                                                resident="one", device_only=True,
                                                counts=(cnt, off[slot_of[masters]], threshold))
        scores = np.empty(len(masters), np.float32)
        scores[order] = scores_dev.cpu().numpy()
        return cnt.cpu().numpy().astype(np.int64), {int(u): (int(off[k]), int(lens[k])) for k, u in enumerate(uniq)}, scores

    def preprofile_stage(self, batch, S, gap_series, threshold=None, masters=None, mode="global", iterations=2,
                         chunk_pairs=1 << 20, shard=(0, 1)):
        """The whole preprofile stage of the workflow (workflow.py:139-161 + :211-224): every master
        against ALL other sequences of the batch, count tables on the device.  Masters are processed in
        chunks of about chunk_pairs pairs; nothing synchronises between chunks, so the host plans
        chunk c + 1 while the device traces chunk c.  mode: "global" (GlobalMasterSlaveAligner) or
        "local" (LocalMasterSlaveAligner with `iterations` Waterman-Eggert iterations).

        Global mode with every sequence a master, a symmetric integer matrix and the packed range: the
        symmetric path (allpairs_dual) fills every UNORDERED pair once and walks it in both orientations;
        shard = (rank, world) then splits the pair list (tables must be summed over ranks), shard = None
        forces the per-master path.

        Returns (count tables as ONE int32 device tensor, {master id: (offset, length)}, DP cells)."""
        S = np.ascontiguousarray(S, np.float32)
        A = S.shape[0]
        n = batch.n
        all_masters = masters is None
        masters = np.arange(n, dtype=np.int64) if masters is None else np.unique(np.asarray(masters, np.int64))
        cnt, off, uniq, lens, slot_of = self._own_counts(batch, masters, A)
        if mode == "global" and shard is not None and (all_masters or len(masters) == n) and n >= 2 and \
                self.dual_traced_ok(S, gap_series, batch.lens) is not None:
            # one fill per UNORDERED pair feeds both master-slave walks; shard = (rank, world) cuts the pair list by
            # DP cells, every rank holds full-size tables and the caller sums them (parallel.allreduce_counts)
            rank, world = shard
            if rank != 0:
                cnt.zero_()             # the masters' own residues are counted once, on rank 0
            _, _, cells = self.allpairs_dual(batch, S, gap_series, counts=(cnt, self.dev(off[:-1].astype(np.int64)), threshold),
                                             shard=shard)
            return cnt, {int(u): (int(off[k]), int(lens[k])) for k, u in enumerate(uniq)}, 2 * cells
        per = max(1, int(chunk_pairs) // max(n - 1, 1))
        everyone = np.arange(n, dtype=np.int64)
        cells = 0
        total_len = int(batch.lens.sum())
        for c0 in range(0, len(masters), per):
            chunk = masters[c0:c0 + per]
            grid = np.broadcast_to(everyone, (len(chunk), n))
            s_arr = grid[grid != chunk[:, None]]              # row-major: each master's slaves in id order
            m_arr = np.repeat(chunk, n - 1)
            cells += int((batch.lens[chunk] * (total_len - batch.lens[chunk])).sum())
            ctx = (cnt, off[slot_of[m_arr]], threshold)
            if mode == "global":
                self.align_pairs(batch, m_arr, s_arr, S, gap_series, mode="global", want_paths=True, resident="one",
                                 device_only=True, counts=ctx)
            elif mode == "local":
                self.local_pairs(batch, m_arr, s_arr, S, gap_series, iterations=iterations, counts=ctx, device_only=True)
            else:
                raise ValueError("preprofile mode must be 'global' or 'local'")
        if mode == "local":
            cells *= iterations
        return cnt, {int(u): (int(off[k]), int(lens[k])) for k, u in enumerate(uniq)}, cells

    # -- symmetric traced all-vs-all: one fill per unordered pair, both master-slave walks -------------
    def dual_traced_ok(self, S, gap_series, lens):
        """One fill can serve the alignments (a, b) and (b, a) when the substitution matrix is symmetric
        (constant gaps are what PairwiseAligner always builds, component/align.py:212-217) and the packed
        traced range holds; returns the int16 sentinel or None."""
        S = np.asarray(S, np.float32)
        go, ge = _gaps(gap_series)
        if not self.use_s16 or os.environ.get("PGPU_NO_DUAL", "") != "" or not np.array_equal(S, S.T):
            return None
        return self.fits_s16(S, go, ge, lens, limit=16000)

    def _tb_words16r(self, tiles, K, cs):
        return tb_words16r(tiles, K, cs, self.nw)

    def _dual_state(self, n_streams):
        """Streams and traceback buffers of the dual path live in the engine: allocating 2 x 16 GiB per call
        costs more than a small stage takes."""
        st = self.__dict__.setdefault("_dual", {"streams": [], "bufs": []})
        while len(st["streams"]) < n_streams:
            st["streams"].append(torch.cuda.Stream(self.device))
            st["bufs"].append(None)
        return st

    def _dual_bufs(self, st, k, n_words, n_slots):
        b = st["bufs"][k]
        if b is None or b[0].numel() < n_words or b[1].numel() < n_slots:
            if b is not None:
                st["streams"][k].synchronize()
            st["bufs"][k] = b = None
            nw_, ns_ = max(n_words, 1 << 20), max(n_slots, 1 << 12)
            with torch.cuda.stream(st["streams"][k]):
                b = (torch.empty(nw_, dtype=torch.int32, device=self.device),       # traceback words
                     torch.empty(ns_, dtype=torch.int32, device=self.device),       # emit_t
                     torch.empty(ns_, dtype=torch.int64, device=self.device),       # pair_tb
                     torch.empty(ns_, dtype=torch.int32, device=self.device),       # slot_res
                     torch.empty(ns_, dtype=torch.int32, device=self.device),       # slot_str
                     torch.empty(ns_, dtype=torch.float32, device=self.device))     # scores
            st["bufs"][k] = b
        return b

    def allpairs_dual(self, batch, S, gap_series, counts=None, want_paths=False, shard=(0, 1), n_streams=None,
                      tile=None):
        """All unordered pairs (i < j) of the batch, global mode, traced ONCE per pair on the paired-resident
        packed kernel; the walk then runs both orientations (sequence_one = i and sequence_one = j) over the
        same traceback words (pgpu_traceback_dual).  counts = (count tables int32 on the device, offset of
        every sequence's table as a device int64 tensor (< 0: not a master), threshold): preprofile mode --
        what N x GlobalMasterSlaveAligner + ProfileBuilder produce (preprofile.py:127-154, profile.py:56).
        want_paths: returns {(i, j): path, (j, i): path} in reference format (tests; small batches).
        Waves alternate between n_streams streams with their own traceback buffers, so that the walk of one
        wave and the tail of its fill overlap the fill of the next.

        Returns (scores dict or None, paths dict or None, DP cells filled)."""
        go, ge = _gaps(gap_series)
        S = np.ascontiguousarray(S, np.float32)
        A = S.shape[0]
        neg16 = self.dual_traced_ok(S, gap_series, batch.lens)
        if neg16 is None:
            raise _lib.PralineGpuError("the dual traced path needs a symmetric integer matrix inside the packed range")
        if batch.max_sym >= A:
            raise ValueError("sequence symbol outside the score matrix")
        lib = self.lib
        n = batch.n
        n_streams = n_streams or int(os.environ.get("PGPU_DUAL_STREAMS", "2"))
        tile = tile or self._pick_tile_dual(n * (n - 1) // 2)
        lens = batch.lens
        cs = np.zeros(n + 1, np.int64)
        np.cumsum(lens, out=cs[1:])
        kcls = self.k_classes(lens)
        if (kcls < 0).any():
            raise _lib.PralineGpuError("sequence longer than %d: use the general kernel" % (32 * self.k_set[-1]))
        budget = max(1 << 20, self.tb_budget_words)
        plan = DualWavePlan(lens, kcls, self.nw, tile, budget, shard, chunk=int(os.environ.get("PGPU_DUAL_CHUNK", "24")))
        cells = plan.cells
        ucum, u_lo, u_hi = plan.ucum, plan.u_lo, plan.u_hi
        plan_waves = plan.waves
        S_dev = self.dev(S)
        maxlen_all = int(lens.max())

        cur = torch.cuda.current_stream(self.device)
        state = self._dual_state(n_streams)
        streams = state["streams"][:n_streams]
        for st in streams:
            st.wait_stream(cur)             # inputs and count tables were written on the caller's stream
        # buffers sized once: a full wave (or everything, for small jobs)
        # (a cell is half a byte of traceback = 1/8 word; _dual_bufs grows a buffer when a wave needs more)
        est_words = int(min(budget + (1 << 22), (ucum[u_hi] - ucum[u_lo]) / 8 * 1.3 + (1 << 22)))
        est_slots = int(est_words * 8 / max(float(lens.mean()) ** 2, 1.0) * 1.5) + (1 << 12)
        sc_out, path_out = ({}, {}) if want_paths else (None, None)
        cnt_dev, seq_off_dev, thr = counts if counts is not None else (None, None, None)
        trace = [] if os.environ.get("PGPU_DUAL_TRACE", "") != "" else None
        for wave_no, (K, wt, wbase, ns, n_words, na, nb) in enumerate(plan_waves()):
            maxlen = max(32 * K, maxlen_all) + 2
            bkey = (0, float(go), float(ge), maxlen, True)
            if bkey not in self._borders:
                B = borders(0, go, ge, maxlen, True)
                with torch.cuda.stream(cur):
                    self._borders[bkey] = (B, self.dev(B["topD"]), self.dev(B["leftD"]))
                for st in streams:
                    st.wait_stream(cur)
            B, top_dev, _ = self._borders[bkey]
            k = wave_no % n_streams
            st = streams[k]
            tb, emit_t, pair_tb, slot_res, slot_str, sc = self._dual_bufs(state, k, max(est_words, n_words), max(est_slots, ns))
            with torch.cuda.stream(st):
                tiles_dev, wbase_dev = self.dev(wt.view(np.uint8)), self.dev(wbase)
                sptr = ctypes.c_void_p(st.cuda_stream)
                emit_t[:ns].fill_(-1)           # holes of the interleaved numbering stay -1: the walk skips them
                tr = None
                if trace is not None:       # PGPU_DUAL_TRACE: device timeline of the waves (fill / walk per stream)
                    tr = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                    tr[0].record(st)
                _lib.check(lib.pgpu_align_tiles16_paired_traced(
                    K, self.ptr(batch.flat_dev), self.ptr(batch.offs_dev), self.ptr(tiles_dev), len(wt),
                    self.ptr(S_dev), A, int(go), int(ge), neg16, self.ptr(top_dev), int(B["left0"]), int(B["left1"]),
                    maxlen + 1, self.ptr(sc), self.ptr(tb), self.ptr(wbase_dev), self.ptr(emit_t), self.ptr(pair_tb),
                    self.ptr(slot_res), self.ptr(slot_str), sptr))
                poff_dev = pbuf = pstart = plen = None
                if want_paths:
                    # host copy of what the kernel records per slot: (resident, streamed) ids
                    res_h = np.full(ns, -1, np.int64)
                    str_h = np.full(ns, -1, np.int64)
                    for a_, b_, r1, r2, t0, sk, ob in zip(na, nb, wt["resident"], wt["resident2"], wt["stream_begin"],
                                                          wt["b_skip"], wt["out_base"]):
                        res_h[ob:ob + 2 * a_:2] = r1
                        str_h[ob:ob + 2 * a_:2] = np.arange(t0, t0 + a_)
                        if b_ > 0:
                            res_h[ob + 1 + 2 * sk:ob + 2 * a_:2] = r2
                            str_h[ob + 1 + 2 * sk:ob + 2 * a_:2] = np.arange(t0 + sk, t0 + a_)
                    live = res_h >= 0
                    cap = np.repeat(np.where(live, batch.lens[np.maximum(res_h, 0)] + batch.lens[np.maximum(str_h, 0)] + 2, 0), 2).astype(np.int64)
                    poff = np.zeros(2 * ns, np.int64)
                    np.cumsum(cap[:-1], out=poff[1:])
                    poff_dev = self.dev(poff)
                    pbuf = torch.empty((int(cap.sum()), 2), dtype=torch.int32, device=self.device)
                    pstart = torch.empty(2 * ns, dtype=torch.int32, device=self.device)
                    plen = torch.empty(2 * ns, dtype=torch.int32, device=self.device)
                if tr is not None:
                    tr[1].record(st)
                if os.environ.get("PGPU_DUAL_SKIP_WALK", "") == "":     # (timing experiments only: fill without the walk)
                  _lib.check(lib.pgpu_traceback_dual(
                    K, self.ptr(batch.offs_dev), self.ptr(slot_res), self.ptr(slot_str), ns, self.ptr(tb),
                    self.ptr(emit_t), self.ptr(pair_tb), B["code00"], B["top_ramp"], B["left_ramp"],
                    self.ptr(batch.flat_dev), self.ptr(cnt_dev), self.ptr(seq_off_dev), A, self.ptr(sc),
                    int(thr is not None), float(thr if thr is not None else 0.0), self.ptr(poff_dev), self.ptr(pbuf),
                    self.ptr(pstart), self.ptr(plen), sptr))
                if tr is not None:
                    tr[2].record(st)
                    trace.append((k, tr))
                # the records were allocated under this stream: the caching allocator hands their memory out again
                # only to later work of the same stream, so dropping the references here is safe
                del tiles_dev, wbase_dev
                self.launches += 2
                if want_paths:
                    st.synchronize()
                    assert np.array_equal(slot_res[:ns].cpu().numpy()[live], res_h[live])
                    assert np.array_equal(slot_str[:ns].cpu().numpy()[live], str_h[live])
                    assert np.array_equal(emit_t[:ns].cpu().numpy() >= 0, live)
                    pb, ps_, pl, sch = pbuf.cpu().numpy(), pstart.cpu().numpy(), plen.cpu().numpy(), sc[:ns].cpu().numpy()
                    for q in np.flatnonzero(live):
                        i, j = int(res_h[q]), int(str_h[q])
                        sc_out[(i, j)] = sch[q]
                        o = poff[2 * q] + ps_[2 * q]
                        path_out[(i, j)] = pb[o:o + pl[2 * q]].copy()
                        o = poff[2 * q + 1] + ps_[2 * q + 1]
                        path_out[(j, i)] = pb[o:o + pl[2 * q + 1]].copy()
        for st in streams:
            cur.wait_stream(st)
        if trace:
            torch.cuda.synchronize(self.device)
            t0 = trace[0][1][0]
            self.dual_trace = [(k, t0.elapsed_time(e[0]), t0.elapsed_time(e[1]), t0.elapsed_time(e[2])) for k, e in trace]
        return sc_out, path_out, cells

    def _pick_tile_dual(self, n_pairs):
        # traced waves hold a few hundred tiles only: shorter streams per warp keep the last round of CTAs full
        if os.environ.get("PGPU_TILE_DUAL"):
            return int(os.environ["PGPU_TILE_DUAL"])
        return self.nw * 8 if n_pairs >= 148 * 6 * self.nw * 8 else self._pick_tile(n_pairs)

    NBOX = 3    # PGPU_NBOX: boxes per pair -> up to NBOX + 1 Waterman-Eggert iterations on the device

    def local_batchable(self, S, gap_series, lens):
        """Local traced batches (pgpu_align_tiles_local): integer-exact scores, gap penalties <= 0,
        residents within the K classes, streamed sequences below 2^20 rows (end-cell key)."""
        go, ge = _gaps(gap_series)
        lens = np.asarray(lens, np.int64)
        # go <= ge: the border cell o[0,0,1] = open - extend takes part in the reference's argmax (align.py:371, :401-403)
        # and must not be positive, since the batched reduction looks at interior cells only
        return bool(go <= ge <= 0 and len(lens) and lens.min() >= 1 and lens.max() < (1 << 20)
                    and self.k_for(int(lens.max())) is not None
                    and self.integer_exact(S, go, ge, int(lens.max())))

    def local_pairs(self, batch, pi, pj, S, gap_series, iterations=1, want_paths=False, counts=None, device_only=False):
        """Local alignments of pairs (sequence_one = pi[k], sequence_two = pj[k]) with Waterman-Eggert
        iterations, the inner loop of LocalMasterSlaveAligner (preprofile.py:227-267): iteration n is a
        local alignment whose zero_idxs are the bounding boxes of iterations 0 .. n-1 of the same pair.

        Returns (scores f32 [iterations x n], paths or None, boxes int32 [n x NBOX x 4]); paths is a
        list per iteration of int32 [rows, 2] arrays (not extended).  counts = (count tables on the
        device, table offset per pair, threshold): local preprofile mode of the walk."""
        if not 1 <= iterations <= self.NBOX + 1:
            raise _lib.PralineGpuError("1..%d Waterman-Eggert iterations are batched" % (self.NBOX + 1))
        go, ge = _gaps(gap_series)
        pi = np.asarray(pi, np.int64)
        pj = np.asarray(pj, np.int64)
        n = len(pi)
        S = np.ascontiguousarray(S, np.float32)
        A = S.shape[0]
        if batch.max_sym >= A:
            raise ValueError("sequence symbol outside the score matrix")
        if not self.local_batchable(S, gap_series, batch.lens):
            raise _lib.PralineGpuError("local batch outside the batched path: use align_general")
        if n == 0:
            return np.zeros((iterations, 0), np.float32), ([[] for _ in range(iterations)] if want_paths else None), \
                np.zeros((0, self.NBOX, 4), np.int32)
        res, strm = pj, pi                      # reference orientation: sequence two lies across the lanes
        kcls = self.k_classes(batch.lens)
        kres = kcls[res]
        if n < 2 or ((res[1:] >= res[:-1]).all() and (kres[1:] >= kres[:-1]).all()):
            order = np.arange(n)
        else:
            order = np.lexsort((np.arange(n), res, kres))
        res_s, str_s = res[order], strm[order]
        S_dev = self.dev(S)
        stream_ids_dev = self.dev(str_s.astype(np.int32))
        slot_res_dev = self.dev(res_s.astype(np.int32))
        scores_dev = torch.empty((iterations, n), dtype=torch.float32, device=self.device)
        empty = np.tile(np.array([1, 0, 1, 0], np.int32), (n, self.NBOX, 1))
        boxes_dev = self.dev(empty)
        cs = np.zeros(n + 1, np.int64)
        np.cumsum(batch.lens[str_s], out=cs[1:])
        caps = (batch.lens[res_s] + batch.lens[str_s] + 2).astype(np.int64)
        cnt_dev = cnt_off_dev = thr = None
        if counts is not None:
            cnt_dev, cnt_off, thr = counts
            cnt_off_dev = self.dev(np.asarray(cnt_off, np.int64)[order])
        maxlen = max(32 * self.k_set[-1], int(batch.lens.max())) + 2
        bkey = (1, float(go), float(ge), maxlen, False)
        if bkey not in self._borders:
            B = borders(1, go, ge, maxlen, False)
            self._borders[bkey] = (B, self.dev(B["topD"]), self.dev(B["leftD"]))
        B, top_dev, _ = self._borders[bkey]
        kk = kcls[res_s]
        bounds = np.flatnonzero(np.diff(kk)) + 1
        tile = self._pick_tile(n)
        pending = []
        for a, b in zip(np.concatenate([[0], bounds]), np.concatenate([bounds, [n]])):
            K = int(kk[a])
            tiles = self._make_tiles(res_s[a:b], b - a, tile)
            for f in ("stream_begin", "stream_end", "out_base"):
                tiles[f] += a
            words = self._tb_words(tiles, K, cs)
            per_tile = words.sum(axis=1)
            lo = 0
            while lo < len(tiles):
                acc = np.cumsum(per_tile[lo:])
                hi = lo + max(1, int(np.searchsorted(acc, self.tb_budget_words, side="right")))
                wt = tiles[lo:hi].copy()
                s_lo, s_hi = int(wt["stream_begin"][0]), int(wt["stream_end"][-1])
                ns = s_hi - s_lo
                wt["out_base"] -= s_lo
                wbase = np.zeros(words[lo:hi].size, np.int64)
                np.cumsum(words[lo:hi].ravel()[:-1], out=wbase[1:])
                tb = torch.empty(int(words[lo:hi].sum()), dtype=torch.int32, device=self.device)
                tiles_dev, wbase_dev = self.dev(wt.view(np.uint8)), self.dev(wbase)
                emit_t = torch.empty(ns, dtype=torch.int32, device=self.device)
                pair_tb = torch.empty(ns, dtype=torch.int64, device=self.device)
                keys = torch.empty(2 * ns, dtype=torch.int64, device=self.device)
                wboxes = boxes_dev[s_lo:s_hi]
                poff = poff_dev = None
                if want_paths:
                    cap = caps[s_lo:s_hi]
                    poff = np.zeros(ns, np.int64)
                    np.cumsum(cap[:-1], out=poff[1:])
                    poff_dev = self.dev(poff)
                for it in range(iterations):
                    sc = scores_dev[it, s_lo:s_hi]
                    ev = self._trace_event("local fill it%d K%d (%d slots)" % (it, K, ns))
                    _lib.check(self.lib.pgpu_align_tiles_local(
                        K, self.ptr(batch.flat_dev), self.ptr(batch.offs_dev), self.ptr(stream_ids_dev), self.ptr(tiles_dev),
                        len(wt), ns, self.ptr(S_dev), A, float(go), float(ge), self.ptr(top_dev), B["left0"], B["left1"],
                        maxlen + 1, self.ptr(sc), self.ptr(keys), self.ptr(tb), self.ptr(wbase_dev), self.ptr(emit_t),
                        self.ptr(pair_tb), self.ptr(wboxes) if it > 0 else None, self.stream()))
                    self._trace_event("local walk it%d" % it, ev)
                    pbuf = pstart = plen = None
                    if want_paths:
                        pbuf = torch.empty((int(cap.sum()), 2), dtype=torch.int32, device=self.device)
                        pstart = torch.empty(ns, dtype=torch.int32, device=self.device)
                        plen = torch.empty(ns, dtype=torch.int32, device=self.device)
                    _lib.check(self.lib.pgpu_traceback_tiles_local(
                        K, self.ptr(batch.offs_dev), self.ptr(slot_res_dev[s_lo:s_hi]), self.ptr(stream_ids_dev[s_lo:s_hi]),
                        ns, self.ptr(keys), self.ptr(tb), self.ptr(emit_t), self.ptr(pair_tb), B["code00"],
                        self.ptr(poff_dev), self.ptr(pbuf), self.ptr(pstart), self.ptr(plen),
                        self.ptr(batch.flat_dev), self.ptr(cnt_dev), self.ptr(cnt_off_dev[s_lo:s_hi]) if cnt_dev is not None else None,
                        A, self.ptr(sc), int(thr is not None), float(thr if thr is not None else 0.0),
                        self.ptr(wboxes) if it > 0 else None, self.ptr(wboxes) if it < self.NBOX else None,
                        min(it, self.NBOX - 1), self.stream()))
                    self.launches += 3
                    self._trace_event(None, ev)
                    if want_paths:
                        pending.append((it, s_lo, s_hi, poff, pbuf, pstart, plen))
                lo = hi
        if device_only:        # nothing read back: the caller keeps queueing work
            return scores_dev, order, boxes_dev
        scores = np.empty((iterations, n), np.float32)
        scores[:, order] = scores_dev.cpu().numpy()
        boxes = np.empty((n, self.NBOX, 4), np.int32)
        boxes[order] = boxes_dev.cpu().numpy()
        paths = None
        if want_paths:
            paths = [[None] * n for _ in range(iterations)]
            for (it, s_lo, s_hi, poff, pbuf, pstart, plen) in pending:
                pb, ps, pl = pbuf.cpu().numpy(), pstart.cpu().numpy(), plen.cpu().numpy()
                for k in range(s_hi - s_lo):
                    o = poff[k] + ps[k]
                    paths[it][order[s_lo + k]] = pb[o:o + pl[k]]
        return scores, paths, boxes

    def local_preprofile_counts(self, batch, masters, slaves, S, gap_series, iterations=2, threshold=None):
        """Count tables of local master-slave preprofiles (LocalMasterSlaveAligner, preprofile.py:160-
        267, followed by ProfileBuilder, profile.py:56), entirely on the device.  Same return value as
        preprofile_counts; scores are [iterations x pairs]."""
        masters = np.asarray(masters, np.int64)
        slaves = np.asarray(slaves, np.int64)
        S = np.ascontiguousarray(S, np.float32)
        A = S.shape[0]
        cnt, off, uniq, lens, slot_of = self._own_counts(batch, masters, A)
        scores, _, _ = self.local_pairs(batch, masters, slaves, S, gap_series, iterations=iterations,
                                        counts=(cnt, off[slot_of[masters]], threshold))
        return cnt.cpu().numpy().astype(np.int64), {int(u): (int(off[k]), int(lens[k])) for k, u in enumerate(uniq)}, scores

    def _own_counts(self, batch, masters, A):
        """Zeroed count tables of the distinct masters with every master's own residues counted once
        (the master occupies every column of its own alignment, util/align.py:205-211)."""
        uniq = np.unique(masters)
        lens = batch.lens[uniq]
        off = np.zeros(len(uniq) + 1, np.int64)
        np.cumsum(lens * A, out=off[1:])
        slot_of = np.full(batch.n, -1, np.int64)
        slot_of[uniq] = np.arange(len(uniq))
        cnt = torch.zeros(int(off[-1]), dtype=torch.int32, device=self.device)
        if len(lens):
            # position of every master residue in the concatenated tables, without a Python loop over the masters
            starts = np.zeros(len(lens), np.int64)
            np.cumsum(lens[:-1], out=starts[1:])
            rows = np.arange(int(lens.sum()), dtype=np.int64) - np.repeat(starts, lens)
            base = np.repeat(off[:-1], lens)
            src = np.repeat(np.asarray(batch.offs, np.int64)[uniq], lens) + rows
            syms = batch.flat_host.numpy()[src].astype(np.int64)
            cnt[self.dev(base + rows * A + syms)] = 1
        return cnt, off, uniq, lens, slot_of

    def align_profile_pairs(self, pbatch, pi, pj, S, gap_series, mode="global", resident=None, fast=False):
        """Scores of profile x profile pairs (sequence_one = pi[k], sequence_two = pj[k]), one track
        set, constant gaps: K1 rows in the reference's evaluation order feed the streaming kernel.
        Score only (GuideTreeBuilder, ad-hoc rounds); traced profile alignments use align_general.

        fast=True factors the contraction through W = P . S^T (A fused multiply-adds per cell instead
        of nnz1 x nnz2 terms): scores within 1e-5 relative of the reference, not bit-identical."""
        md = MODES[mode]
        if md == 1 and not (_gaps(gap_series)[0] <= _gaps(gap_series)[1] <= 0):
            raise _lib.PralineGpuError("batched local scores need open <= extend <= 0 (the border cell open - extend takes "
                                       "part in the reference's argmax): use align_general")
        go, ge = _gaps(gap_series)
        pi = np.asarray(pi, np.int64)
        pj = np.asarray(pj, np.int64)
        n = len(pi)
        S = np.ascontiguousarray(S, np.float32)
        A = S.shape[0]
        if A != pbatch.A:
            raise ValueError("profile alphabet does not match the score matrix")
        if n == 0:
            return np.zeros(0, np.float32)
        if resident is None:
            # exact score rows are cheapest with sequence one resident (k_build_rows_t tabulates the
            # first product of every term per streamed row); tolerance mode does not care
            resident = "one" if (not fast or len(np.unique(pi)) < len(np.unique(pj))) else "two"
        transposed = resident == "one"
        res, strm = (pi, pj) if transposed else (pj, pi)
        kcls = self.k_classes(pbatch.lens)
        if (kcls[res] < 0).any():
            raise _lib.PralineGpuError("resident profile longer than %d: use the general kernel" % (32 * self.k_set[-1]))
        order = np.lexsort((np.arange(n), res, kcls[res]))
        res_s, str_s = res[order], strm[order]
        S_dev = self.dev(S)
        wres = None
        if fast:
            wres = torch.empty_like(pbatch.prof_dev)
            _lib.check(self.lib.pgpu_profile_times_matrix(self.ptr(pbatch.prof_dev), self.ptr(S_dev), A,
                                                          int(pbatch.prof_dev.shape[0]), int(transposed), self.ptr(wres),
                                                          self.stream()))
            self.launches += 1
        whi = wlo = phi = plo = padoff = None
        if fast and self.fast_tc and A <= 32 and self.tc_tma:
            # resident side pre-split once per batch into the tensor core's operand layout: the score-row
            # kernel then fetches its B tiles with TMA bulk copies
            pad = np.where(kcls > 0, 32 * kcls, (pbatch.lens + 31) // 32 * 32).astype(np.int64)
            padoff = np.zeros(pbatch.n + 1, np.int64)
            np.cumsum(pad, out=padoff[1:])
            whi = torch.empty(int(padoff[-1]) * 128, dtype=torch.uint8, device=self.device)
            wlo = torch.empty_like(whi)
            padoff_dev = self.dev(padoff)
            _lib.check(self.lib.pgpu_split_residents(self.ptr(wres), self.ptr(pbatch.offs_dev), self.ptr(padoff_dev), pbatch.n, A,
                                                     self.ptr(whi), self.ptr(wlo), self.stream()))
            # ... and the streamed side (the profiles themselves), 32 rows of slack: a bulk copy always takes 32 rows
            phi = torch.zeros((int(padoff[-1]) + 32) * 128, dtype=torch.uint8, device=self.device)
            plo = torch.zeros_like(phi)
            _lib.check(self.lib.pgpu_split_residents(self.ptr(pbatch.prof_dev), self.ptr(pbatch.offs_dev), self.ptr(padoff_dev),
                                                     pbatch.n, A, self.ptr(phi), self.ptr(plo), self.stream()))
            self.launches += 2
        stream_ids_dev = self.dev(str_s.astype(np.int32))
        scores_dev = torch.empty(n, dtype=torch.float32, device=self.device)
        lens_s = pbatch.lens[str_s]
        cs = np.zeros(n + 1, np.int64)
        np.cumsum(lens_s, out=cs[1:])
        kk = kcls[res_s]
        bounds = np.flatnonzero(np.diff(kk)) + 1
        tile = self._pick_tile(n)
        nw = self.nw
        for a, b in zip(np.concatenate([[0], bounds]), np.concatenate([bounds, [n]])):
            K = int(kk[a])
            width = 32 * K
            tiles = self._make_tiles(res_s[a:b], b - a, tile)
            for f in ("stream_begin", "stream_end", "out_base"):
                tiles[f] += a
            rows_per_tile = (cs[tiles["stream_end"]] - cs[tiles["stream_begin"]]) + nw
            cum_floats = np.cumsum(rows_per_tile) * width
            lo = 0
            while lo < len(tiles):   # waves bounded by the matrix budget
                done = int(cum_floats[lo - 1]) if lo else 0
                hi = max(lo + 1, int(np.searchsorted(cum_floats, done + self.m_budget_floats, side="right")))
                wt = tiles[lo:hi]
                use_tc = fast and self.fast_tc and A <= 32
                if use_tc:
                    mrow_base, blocks, n_rows, quads = plan_profile_wave_native(wt, nw, cs, lens_s, str_s, res_s, pbatch.offs,
                                                                                want_quads=True, padoff=padoff)
                else:
                    mrow_base, blocks, n_rows = plan_profile_wave_native(wt, nw, cs, lens_s, str_s, res_s, pbatch.offs)
                # + 4 floats: the matrix-fed kernel reads whole aligned float4s around a lane's K scores
                mwave = torch.empty(n_rows * width + 4, dtype=torch.float32, device=self.device)[:n_rows * width]
                # keep every device temporary referenced until the launches that read it are queued:
                # a tensor freed right after data_ptr() is handed to the next allocation
                blocks_dev = self.dev(blocks.view(np.uint8))
                if use_tc:
                    # tensor-core score rows (tcgen05 tf32 with a hi/lo split): 128-row tiles = quads of
                    # consecutive row blocks that share a resident
                    quads_dev = self.dev(quads.view(np.uint8))
                ev = self._trace_event("score rows %s (%d B)" % ("tc" if use_tc else "fma" if fast else "exact",
                                                               n_rows * width * 4))
                if use_tc:
                    _lib.check(self.lib.pgpu_build_rows_tc(self.ptr(pbatch.prof_dev), self.ptr(wres), A, self.ptr(quads_dev),
                                                           len(quads), width, int(md == 1), self.ptr(mwave),
                                                           self.ptr(whi), self.ptr(wlo), self.ptr(phi), self.ptr(plo),
                                                           self.stream()))
                elif fast:
                    _lib.check(self.lib.pgpu_build_rows_fast(self.ptr(pbatch.prof_dev), self.ptr(wres), self.ptr(pbatch.offs_dev),
                                                             A, self.ptr(blocks_dev), len(blocks), width, int(md == 1),
                                                             self.ptr(mwave), self.stream()))
                else:
                    _lib.check(self.lib.pgpu_build_rows(self.ptr(pbatch.prof_dev), self.ptr(pbatch.offs_dev), A,
                                                        self.ptr(S_dev), self.ptr(blocks_dev), len(blocks), width,
                                                        int(transposed), int(md == 1),
                                                        getattr(pbatch, "dense_syms", lambda: 0)() if self.rows_x2 else 0,
                                                        self.ptr(mwave), self.stream()))
                self.launches += 1
                self._trace_event("matrix-fed stream", ev)
                if self.keep_mwave:      # tests compare the score rows of the three builders element by element
                    self.last_mwave = mwave
                self.run_tiles(md, K, transposed, pbatch, stream_ids_dev, wt, n, S_dev, A, go, ge, scores_dev,
                               mwave_dev=mwave, mrow_base_dev=self.dev(mrow_base))
                self._trace_event(None, ev)
                lo = hi
        out = np.empty(n, np.float32)
        out[order] = scores_dev.cpu().numpy()
        return out

    def allpairs_tiles(self, batch, shard=(0, 1), tile=None, paired=False):
        """All unordered pairs (i < j), sequence_one = i resident, sequence_two = j streamed, slots
        in condensed (np.triu_indices) order.  Returns (tiles per K, this shard's slot range, its
        DP cells, slot cut per rank, device cache, paired).

        paired=True lays residents out two by two (2p, 2p+1) for the packed int16 kernel whose
        register halves carry two residents over one shared stream (pgpu_align_tiles16, paired=1):
        the stream of a pair starts at j = 2p+1, where the high half has no partner yet (b_skip)."""
        n = batch.n
        rank, world = shard
        tile = tile or self._pick_tile(n * (n - 1) // 2)
        n_pairs = n * (n - 1) // 2
        lens = batch.lens
        cs = np.zeros(n + 1, np.int64)
        np.cumsum(lens, out=cs[1:])
        kcls = self.k_classes(lens)
        if (kcls[:max(n - 1, 1)] < 0).any():
            raise _lib.PralineGpuError("sequence longer than %d: use the general kernel" % (32 * self.k_set[-1]))
        cond = lambda i, j: i * n - i * (i + 1) // 2 + (j - i - 1)
        step = 2 if paired else 1
        i = np.arange(0, n - 1, step, dtype=np.int64)            # first resident of every unit
        has_b = (i + 1 <= n - 2) if paired else np.zeros(len(i), bool)
        cnt = n - 1 - i                                           # stream elements j = i+1 .. n-1
        ntile = (cnt + tile - 1) // tile
        unit = np.repeat(np.arange(len(i)), ntile)
        first = np.cumsum(ntile) - ntile
        gi = i[unit]
        tb = gi + 1 + (np.arange(int(ntile.sum())) - first[unit]) * tile
        te = np.minimum(tb + tile, n)
        hb = has_b[unit]
        rows = cs[te] - cs[tb]
        cells = lens[gi] * rows + np.where(hb, lens[np.minimum(gi + 1, n - 1)] * (cs[te] - cs[np.maximum(tb, gi + 2)]), 0)
        # shard at unit boundaries so that every rank owns whole condensed rows
        ucells = np.zeros(len(i) + 1, np.int64)
        np.add.at(ucells, unit + 1, cells)
        ucum = np.cumsum(ucells)
        ucuts = np.searchsorted(ucum, ucum[-1] * np.arange(world + 1) / world, side="left")
        ucuts[0], ucuts[-1] = 0, len(i)
        ucuts = np.maximum.accumulate(ucuts)
        sel = (unit >= ucuts[rank]) & (unit < ucuts[rank + 1])
        tiles = np.zeros(int(sel.sum()), TILE_DTYPE)
        tiles["resident"] = gi[sel]
        tiles["resident2"] = np.where(hb[sel], gi[sel] + 1, -1)
        tiles["stream_begin"] = tb[sel]
        tiles["stream_end"] = te[sel]
        tiles["out_base"] = cond(gi[sel], tb[sel])
        tiles["out_base2"] = np.where(hb[sel], cond(gi[sel] + 1, np.maximum(tb[sel], gi[sel] + 2)), 0)
        tiles["b_skip"] = np.where(hb[sel] & (tb[sel] == gi[sel] + 1), 1, 0)
        row_cut = lambda u: int(cond(i[u], i[u] + 1)) if u < len(i) else n_pairs
        slot_cuts = [row_cut(int(u)) for u in ucuts]
        kk = kcls[tiles["resident"]]
        if paired:
            kk = np.maximum(kk, np.where(tiles["resident2"] >= 0, kcls[np.maximum(tiles["resident2"], 0)], 0))
        by_k = {int(K): tiles[kk == K] for K in np.unique(kk)}
        return by_k, (slot_cuts[rank], slot_cuts[rank + 1]), int(cells[sel].sum()), slot_cuts, {}, bool(paired)

    def allpairs_plan(self, batch, S_host, gap_series, mode="global", shard=(0, 1)):
        """allpairs_tiles for this batch, cached in the engine per (sequence lengths, shard, kernel
        choice): the plan depends on nothing else, so repeated all-vs-all calls over sequences of the
        same lengths (every step of a job, a re-uploaded batch) plan once and reuse the tile records
        already on the device.  The cache holds the last few plans."""
        md = MODES[mode]
        go, ge = _gaps(gap_series)
        paired = self.wants_paired(S_host, go, ge, md, batch)
        key = (hash(batch.lens.tobytes()), int(batch.n), int(shard[0]), int(shard[1]), bool(paired))
        cache = self.__dict__.setdefault("_plan_cache", {})
        plan = cache.pop(key, None)
        if plan is None:
            plan = self.allpairs_tiles(batch, shard, paired=paired)
            while len(cache) >= 4:
                cache.pop(next(iter(cache)))
        cache[key] = plan          # most recently used last
        return plan

    def allpairs_scores(self, batch, S_dev, A, gap_series, mode="global", shard=(0, 1), out=None, plan=None,
                        S_host=None):
        """Condensed all-vs-all score vector on the device (this shard's slots filled).  out: a plain
        condensed f32 tensor, or a parallel.ShardedCondensed built on the plan's slot cuts -- then this
        rank's scores land directly in its slice of the all-gather buffer."""
        md = MODES[mode]
        go, ge = _gaps(gap_series)
        n_pairs = batch.n * (batch.n - 1) // 2
        if plan is None:
            plan = self.allpairs_plan(batch, S_host, gap_series, mode, shard)
        by_k, rng, cells = plan[0], plan[1], plan[2]
        cache = plan[4] if len(plan) > 4 else {}
        paired = bool(plan[5]) if len(plan) > 5 else False
        if paired and not self.wants_paired(S_host, go, ge, md, batch):
            raise _lib.PralineGpuError("a paired-resident plan needs the packed int16 kernel (global mode, integer scores)")
        if out is None:
            out = torch.empty(n_pairs, dtype=torch.float32, device=self.device)
        out_dev, shift = out, 0
        if hasattr(out, "layout_dev"):
            if list(out.cuts) != [int(c) for c in plan[3]]:
                raise _lib.PralineGpuError("the sharded score buffer was built on other slot cuts than this plan")
            out_dev, shift = out.buf, out.shift[shard[0]]
        for K, tiles in by_k.items():
            if K not in cache:
                cache[K] = self.dev(tiles.view(np.uint8))
            self.run_tiles(md, K, True, batch, None, tiles, n_pairs, S_dev, A, go, ge, out_dev, tiles_dev=cache[K],
                           S_host=S_host, paired=paired, out_shift=shift)
        return out, rng, cells

    def wants_paired(self, S_host, go, ge, md, batch):
        """All-vs-all runs on the paired-resident int16 kernel when the packed path applies."""
        return bool(self.use_s16 and md == 0 and S_host is not None and
                    self.fits_s16(np.asarray(S_host), go, ge, batch.lens) is not None)

    # -- general single alignment ----------------------------------------------------------------
    def padded_matrix(self, L1, L2):
        """[L1 x L2] f32 view whose pitch covers whole 128-column strips: the layout the lean
        wavefront kernel wants (pad columns are computed and ignored)."""
        pitch = (L2 + 127) // 128 * 128
        return torch.empty((L1, pitch), dtype=torch.float32, device=self.device)[:, :L2]

    def build_scores(self, P1s, P2s, Ss):
        """m = sum_sets P1 . S . P2^T on the device, reference evaluation order (cext.c:308-455)."""
        n = len(P1s)
        d1 = [self.dev(np.asarray(p, np.float32)) for p in P1s]
        d2 = [self.dev(np.asarray(p, np.float32)) for p in P2s]
        ds = [self.dev(np.asarray(s, np.float32)) for s in Ss]
        L1, L2 = d1[0].shape[0], d2[0].shape[0]
        m = self.padded_matrix(L1, L2)
        arr = ctypes.c_void_p * n
        A = (ctypes.c_int * n)(*[int(p.shape[1]) for p in d1])
        _lib.check(self.lib.pgpu_build_scores(n, arr(*[t.data_ptr() for t in d1]), arr(*[t.data_ptr() for t in d2]),
                                              arr(*[t.data_ptr() for t in ds]), A, L1, L2, self.ptr(m), int(m.stride(0)),
                                              self.stream()))
        self.launches += 1
        return m

    def align_general(self, mode, m, g1, g2, zero_idxs=None, want_path=True, want_matrices=False):
        """One RawPairwiseAligner call (component/align.py:302-447).  m may be a device tensor."""
        md = MODES[mode]
        if isinstance(m, torch.Tensor):
            m_dev = m
        else:
            mh = np.asarray(m, np.float32)
            m_dev = self.padded_matrix(mh.shape[0], mh.shape[1])
            m_dev.copy_(torch.from_numpy(np.ascontiguousarray(mh)), non_blocking=False)
        L1, L2 = int(m_dev.shape[0]), int(m_dev.shape[1])
        if L1 < 1 or L2 < 1:
            raise ValueError("empty sequences cannot be aligned")
        g1h = np.asarray(g1, np.float32).reshape(L1, 2)
        g2h = np.asarray(g2, np.float32).reshape(L2, 2)
        var_gaps = int(not ((g1h == g1h[0]).all() and (g2h == g2h[0]).all()))
        g1_dev, g2_dev = self.dev(g1h), self.dev(g2h)
        z_dev = None
        if zero_idxs is not None and len(zero_idxs):
            z = np.zeros((L1 + 1, L2 + 1), np.uint8)
            zi = np.asarray(zero_idxs, np.int64).reshape(-1, 2)     # Waterman-Eggert boxes: tens of thousands of cells
            z[zi[:, 0], zi[:, 1]] = 1
            z_dev = self.dev(z)
        ws = torch.empty(int(self.lib.pgpu_general_workspace_bytes(L1, L2)), dtype=torch.uint8, device=self.device)
        # one int32 buffer: [score | cell y x k | path start | path len | pad 2] + path rows
        nrows = (L1 + L2 + 2) if want_path else 0
        outb = torch.zeros(8 + 2 * nrows, dtype=torch.int32, device=self.device)
        o = t = None
        if want_matrices:
            o = torch.zeros((L1 + 1, L2 + 1, 3), dtype=torch.float32, device=self.device)
            t = torch.zeros((L1 + 1, L2 + 1, 3), dtype=torch.uint8, device=self.device)
        base = outb.data_ptr()
        _lib.check(self.lib.pgpu_align_general(md, L1, L2, self.ptr(m_dev), int(m_dev.stride(0)), self.ptr(g1_dev),
                                               self.ptr(g2_dev), var_gaps, self.ptr(z_dev), L2 + 1, self.ptr(ws),
                                               ctypes.c_void_p(base), ctypes.c_void_p(base + 4),
                                               ctypes.c_void_p(base + 32) if want_path else None,
                                               ctypes.c_void_p(base + 16) if want_path else None,
                                               ctypes.c_void_p(base + 20) if want_path else None,
                                               self.ptr(o), self.ptr(t), self.stream()))
        self.launches += 3 + int(want_path)
        h = outb.cpu().numpy()          # the single device -> host read of this alignment
        score = float(h[:1].view(np.float32)[0])
        if score != score:      # k_gen_finalize: a strip hand-off timed out (GPU shared / preempted); the fill is invalid
            raise _lib.PralineGpuError("wavefront kernel: a strip-to-strip hand-off timed out; no result was produced")
        res = dict(score=score, cell=tuple(int(v) for v in h[1:4]))
        if want_path:
            st, ln = int(h[4]), int(h[5])
            res["path"] = h[8:].reshape(-1, 2)[st:st + ln].copy()
        if want_matrices:
            res["o"], res["t"] = o.cpu().numpy(), t.cpu().numpy()
        return res

    def align_profile_pair(self, mode, p1, p2, S, g1, g2, want_path=True):
        """One profile x profile alignment of ONE track set, K1 + K3 in one library call (pgpu_align_profile_long):
        the same result as build_scores + align_general; long global / semiglobal alignments with constant gap
        pairs (BASELINE config 5) have the score matrix built beside the wavefront fill."""
        md = MODES[mode]
        d1, d2 = self.dev(np.asarray(p1, np.float32)), self.dev(np.asarray(p2, np.float32))
        S_dev = self.dev(np.asarray(S, np.float32))
        L1, L2, A = int(d1.shape[0]), int(d2.shape[0]), int(d1.shape[1])
        if L1 < 1 or L2 < 1:
            raise ValueError("empty sequences cannot be aligned")
        if int(d2.shape[1]) != A or tuple(S_dev.shape) != (A, A):
            raise ValueError("profile alphabets do not match the score matrix")
        g1h = np.asarray(g1, np.float32).reshape(L1, 2)
        g2h = np.asarray(g2, np.float32).reshape(L2, 2)
        var_gaps = int(not ((g1h == g1h[0]).all() and (g2h == g2h[0]).all()))
        g1_dev, g2_dev = self.dev(g1h), self.dev(g2h)
        m = self.padded_matrix(L1, L2)
        ws = torch.empty(int(self.lib.pgpu_general_workspace_bytes(L1, L2)), dtype=torch.uint8, device=self.device)
        nrows = (L1 + L2 + 2) if want_path else 0
        outb = torch.zeros(8 + 2 * nrows, dtype=torch.int32, device=self.device)
        base = outb.data_ptr()
        _lib.check(self.lib.pgpu_align_profile_long(md, self.ptr(d1), self.ptr(d2), self.ptr(S_dev), A, L1, L2, self.ptr(m),
                                                    int(m.stride(0)), self.ptr(g1_dev), self.ptr(g2_dev), var_gaps, self.ptr(ws),
                                                    ctypes.c_void_p(base), ctypes.c_void_p(base + 4),
                                                    ctypes.c_void_p(base + 32) if want_path else None,
                                                    ctypes.c_void_p(base + 16) if want_path else None,
                                                    ctypes.c_void_p(base + 20) if want_path else None, self.stream()))
        self.launches += 4 + int(want_path)
        h = outb.cpu().numpy()          # the single device -> host read of this alignment
        score = float(h[:1].view(np.float32)[0])
        if score != score:
            raise _lib.PralineGpuError("wavefront kernel: a hand-off timed out; no result was produced")
        res = dict(score=score, cell=tuple(int(v) for v in h[1:4]))
        if want_path:
            st, ln = int(h[4]), int(h[5])
            res["path"] = h[8:].reshape(-1, 2)[st:st + ln].copy()
        return res

    # -- progressive merge: count tables stay on the device, one guide-tree level per call -----------
    def merge_level(self, jobs, S, gap_series, mode, n_streams=8):
        """The independent merges of ONE guide-tree level (TreeMultipleSequenceAligner, component/msa.py:124-237):
        jobs = [(counts_one, rows_one, counts_two, rows_two)] with int32 [rows x A] count tables on the device.
        Per job, on one of n_streams streams: probability profiles from the counts (pgpu_counts_to_profile =
        ProfileTrack.profile), K1 score matrix in the reference's evaluation order, K3 wavefront fill + walk, and the
        merged count table gathered along the path on the device (pgpu_merge_counts = ProfileTrack.merge).  One
        synchronisation and ONE device -> host read per level bring back the paths (Alignment.merge is host
        bookkeeping the output needs anyway).  Returns [(merged counts tensor, rows, path int32 [rows + 1, 2], score)]."""
        md = MODES[mode]
        go, ge = _gaps(gap_series)
        S = np.ascontiguousarray(S, np.float32)
        A = S.shape[0]
        if not jobs:
            return []
        lib = self.lib
        cur = torch.cuda.current_stream(self.device)
        state = self._dual_state(n_streams)
        streams = state["streams"][:n_streams]
        for st in streams:
            st.wait_stream(cur)
        S_dev = self.dev(S)
        maxlen = max(max(j[1], j[3]) for j in jobs)
        gkey = ("gapconst", float(go), float(ge))
        cache = self.__dict__.setdefault("_gap_cache", {})
        if gkey not in cache or cache[gkey].shape[0] < maxlen:
            g = np.empty((max(maxlen, 1024) * 2, 2), np.float32)
            g[:] = (go, ge)
            cache[gkey] = self.dev(g)
        g_dev = cache[gkey]
        for st in streams:
            st.wait_stream(cur)
        # one int32 block per job: [score | cell y x k | path start | path len | pad 2] + path rows; all in one tensor.
        # Every buffer of the level is ONE allocation carved by offsets (six torch.empty per merge, most of them a
        # cudaMalloc on a fresh stream pool, cost more than the kernels of a 400 x 400 merge).
        nj = len(jobs)
        offs = np.zeros(nj + 1, np.int64)
        p_off = np.zeros(nj + 1, np.int64)      # profile rows (both operands)
        m_off = np.zeros(nj + 1, np.int64)      # match-score floats
        w_off = np.zeros(nj + 1, np.int64)      # workspace bytes
        o_off = np.zeros(nj + 1, np.int64)      # merged count rows
        pitch = [0] * nj
        for k, (c1, r1, c2, r2) in enumerate(jobs):
            offs[k + 1] = offs[k] + 8 + 2 * (r1 + r2 + 2)
            p_off[k + 1] = p_off[k] + r1 + r2
            pitch[k] = (r2 + 127) // 128 * 128
            m_off[k + 1] = m_off[k] + r1 * pitch[k]
            w_off[k + 1] = w_off[k] + (int(lib.pgpu_general_workspace_bytes(r1, r2)) + 255) // 256 * 256
            o_off[k + 1] = o_off[k] + r1 + r2 + 1
        outs = torch.zeros(int(offs[-1]), dtype=torch.int32, device=self.device)
        prof = torch.empty((int(p_off[-1]), A), dtype=torch.float32, device=self.device)
        mbuf = torch.empty(int(m_off[-1]), dtype=torch.float32, device=self.device)
        wbuf = torch.empty(int(w_off[-1]), dtype=torch.uint8, device=self.device)
        obuf = torch.empty((int(o_off[-1]), A), dtype=torch.int32, device=self.device)
        for st in streams:
            st.wait_stream(cur)
        merged = []
        arr1 = ctypes.c_void_p * 1
        cA = (ctypes.c_int * 1)(A)
        for k, (c1, r1, c2, r2) in enumerate(jobs):
            st = streams[k % n_streams]
            sptr = ctypes.c_void_p(st.cuda_stream)
            p1 = prof.data_ptr() + 4 * A * int(p_off[k])
            p2 = p1 + 4 * A * r1
            mp = mbuf.data_ptr() + 4 * int(m_off[k])
            _lib.check(lib.pgpu_counts_to_profile(self.ptr(c1), r1, A, ctypes.c_void_p(p1), sptr))
            _lib.check(lib.pgpu_counts_to_profile(self.ptr(c2), r2, A, ctypes.c_void_p(p2), sptr))
            _lib.check(lib.pgpu_build_scores(1, arr1(p1), arr1(p2), arr1(S_dev.data_ptr()), cA, r1, r2, ctypes.c_void_p(mp),
                                             pitch[k], sptr))
            base = outs.data_ptr() + 4 * int(offs[k])
            _lib.check(lib.pgpu_align_general(md, r1, r2, ctypes.c_void_p(mp), pitch[k], self.ptr(g_dev), self.ptr(g_dev),
                                              0, None, r2 + 1, ctypes.c_void_p(wbuf.data_ptr() + int(w_off[k])),
                                              ctypes.c_void_p(base), ctypes.c_void_p(base + 4),
                                              ctypes.c_void_p(base + 32), ctypes.c_void_p(base + 16),
                                              ctypes.c_void_p(base + 20), None, None, sptr))
            out = obuf[int(o_off[k]):int(o_off[k + 1])]
            _lib.check(lib.pgpu_merge_counts(self.ptr(c1), self.ptr(c2), A, ctypes.c_void_p(base), self.ptr(out),
                                             r1 + r2 + 1, sptr))
            merged.append(out)
            self.launches += 7
        keep = (prof, mbuf, wbuf)
        for st in streams:
            cur.wait_stream(st)
        h = outs.cpu().numpy()          # the single device -> host read of the level
        del keep
        res = []
        for k, (c1, r1, c2, r2) in enumerate(jobs):
            b = h[offs[k]:offs[k + 1]]
            score = float(b[:1].view(np.float32)[0])
            if score != score:
                raise _lib.PralineGpuError("wavefront kernel: a strip-to-strip hand-off timed out; no result was produced")
            st_, ln = int(b[4]), int(b[5])
            path = b[8:].reshape(-1, 2)[st_:st_ + ln].copy()
            res.append((merged[k][:ln - 1], ln - 1, path, score))
        return res

    def build_scores_seq(self, batch, i, j, S_dev, A):
        """m[y][x] = S[a_y][b_x] for sequences i (rows) and j (columns) of a batch."""
        L1, L2 = int(batch.lens[i]), int(batch.lens[j])
        m = self.padded_matrix(L1, L2)
        base = batch.flat_dev.data_ptr()
        _lib.check(self.lib.pgpu_build_scores_seq(ctypes.c_void_p(base + int(batch.offs[i])),
                                                  ctypes.c_void_p(base + int(batch.offs[j])), self.ptr(S_dev), A,
                                                  L1, L2, self.ptr(m), int(m.stride(0)), self.stream()))
        self.launches += 1
        return m

    def align_seq_pair_general(self, batch, i, j, S, gap_series, mode, zero_idxs=None, want_matrices=False):
        """One PairwiseAligner call through the general kernels (any gaps, masks, local paths)."""
        go, ge = _gaps(gap_series)
        S = np.ascontiguousarray(S, np.float32)
        m = self.build_scores_seq(batch, i, j, self.dev(S), S.shape[0])
        L1, L2 = int(batch.lens[i]), int(batch.lens[j])
        g1 = np.empty((L1, 2), np.float32)
        g2 = np.empty((L2, 2), np.float32)
        g1[:] = (go, ge)
        g2[:] = (go, ge)
        return self.align_general(mode, m, g1, g2, zero_idxs=zero_idxs, want_matrices=want_matrices)

    def fill_debug(self, mode, m, g1, g2, z=None):
        """B3 shim: host arrays in, the reference's full o and t arrays out."""
        m = np.ascontiguousarray(m, np.float32)
        g1 = np.ascontiguousarray(g1, np.float32)
        g2 = np.ascontiguousarray(g2, np.float32)
        L1, L2 = m.shape
        o = np.zeros((L1 + 1, L2 + 1, 3), np.float32)
        t = np.zeros((L1 + 1, L2 + 1, 3), np.uint8)
        zc = np.ascontiguousarray(z, np.uint8) if z is not None else None
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _lib.check(self.lib.pgpu_fill_debug(MODES[mode], vp(m), vp(g1), vp(g2), vp(o), vp(t),
                                            vp(zc) if zc is not None else None, L1, L2))
        return o, t

    # -- guide-tree clustering ---------------------------------------------------------------------
    LINKAGES = {"single": 0, "complete": 1, "average": 2}

    def tree_distance(self, scores, n):
        """Condensed pair scores (np.triu_indices order; device tensor, array, or a
        parallel.ShardedCondensed holding the rank slices) -> the f32 distance matrix of
        GuideTreeBuilder on the device (component/tree.py:92-147): d has 0 on the diagonal,
        dist = (-d) + d.max(), all in f32 like the reference's numpy expression
        (pgpu_tree_distance: a max reduction and a tiled symmetric fill, cluster.cu)."""
        cuts_dev = shift_dev = None
        n_cuts = 0
        if hasattr(scores, "layout_dev"):
            if scores.world > 1:
                cuts_dev, shift_dev = scores.layout_dev()
                n_cuts = scores.world
            cond = scores.buf
        else:
            cond = scores if isinstance(scores, torch.Tensor) else self.dev(np.ascontiguousarray(scores, np.float32))
            cond = cond.to(torch.float32).contiguous()
            if cond.numel() != n * (n - 1) // 2:
                raise ValueError("condensed vector of %d sequences has %d entries" % (n, n * (n - 1) // 2))
        d = torch.empty((n, n), dtype=torch.float32, device=self.device)
        scratch = torch.empty(1, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.pgpu_tree_distance(n, self.ptr(cond), n_cuts, self.ptr(cuts_dev), self.ptr(shift_dev),
                                               self.ptr(d), self.ptr(scratch), self.stream()))
        self.launches += 2
        return d

    def cluster_merge_order(self, dist, linkage="average"):
        """HierarchicalClusteringAlgorithm(dist).merge_order(linkage) (util/cluster.py:27-57) on the
        device; dist is an [n x n] f32 array or device tensor.  Returns a list of (one, two) tuples."""
        if linkage not in self.LINKAGES:
            raise ValueError("unknown linkage method '%s'" % linkage)
        dd = dist if isinstance(dist, torch.Tensor) else self.dev(np.ascontiguousarray(dist, np.float32))
        dd = dd.to(torch.float32).contiguous()
        n = int(dd.shape[0])
        if dd.dim() != 2 or int(dd.shape[1]) != n:
            raise ValueError("distance matrix must be square")
        if n < 2:
            return []
        ws = torch.empty(int(self.lib.pgpu_cluster_workspace_bytes(n)), dtype=torch.uint8, device=self.device)
        merges = torch.empty((n - 1, 2), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.pgpu_cluster_merge_order(n, self.LINKAGES[linkage], self.ptr(dd), self.ptr(ws),
                                                     self.ptr(merges), self.stream()))
        self.launches += 2
        return [(int(a), int(b)) for a, b in merges.cpu().numpy().tolist()]

    def microbench(self):
        out = (ctypes.c_double * 14)()
        _lib.check(self.lib.pgpu_microbench(out, 14))
        names = ["fadd", "fmnmx", "fmnmx3", "cell_mix", "viaddmnmx_s32", "viaddmnmx_s16x2", "shfl", "lds128", "sm_mhz",
                 "shf", "imad", "lop3", "iadd3", "cell_mix16"]
        return dict(zip(names, [float(v) for v in out]))


_ENGINES = {}


def get_engine(device=0):
    if device not in _ENGINES:
        _ENGINES[device] = Engine(device)
    return _ENGINES[device]
