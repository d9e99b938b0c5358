#!/usr/bin/env python
"""bench.py -- GCUPS of the all-vs-all affine-gap DP (BASELINE.json metric) on N B200s.

A "step" is one pass of the hot path over the whole pair list of BASELINE configs[2] (the Target
of north_star): 10,000 synthetic 400-residue proteins (seed 3 family), all
49,995,000 unordered pairs = 8.0e12 DP cells, BLOSUM62, gaps [-11, -1], global mode, score per
pair -- the all-vs-all GuideTreeBuilder queues (reference: praline/component/tree.py:99-147).
STRONG scaling: the pair list is fixed and sharded by DP cells over the N ranks; the condensed
score vector is assembled by one in-place NCCL all-gather per step (praline_b200/parallel.py).
Every line carries a parity check: rank 0 compares >= 256 random slots of EVERY rank's slice of
the gathered vector with the oracle, every rank's copy must hash to rank 0's, and the integer
checksums of the whole vector must equal the recorded 1-GPU values; a mismatch exits non-zero.

Sub-records (under "configs"): BASELINE configs[1] (1,000 x 300 aa, packed int16 and f32
kernels, N = 1), the f32 kernel on the Target, the preprofile stage of configs[2] (traced
alignments -> count tables, sharded by master), configs[3] (progressive merge workflow, N = 1),
configs[4] (20 kb x 20 kb DNA profiles, N = 1) and the MSA wall time of configs[0].

    python bench.py --gpus 1 --steps 5 --warmup 3
    python bench.py --impl reference ...    # the reference's own C path on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from praline_b200 import matrices, synth  # noqa: E402

C3 = dict(name="c3", n_seqs=10000, length=400, seed=3, gaps=[-11.0, -1.0], mode="global")
C2 = dict(name="c2", n_seqs=1000, length=300, seed=2, gaps=[-11.0, -1.0], mode="global")
W_FLOPS_PER_CELL = 11.0   # 7 add + 4 max of the reference recurrence
# Integer checksums of the C3 condensed score vector (scores are integers, so int64 sums are exact and
# independent of the sharding): sum(score) and sum(score * (slot % 1000003)).  Recorded from the 1-GPU
# run whose sample was oracle-checked (profiles/r02_bench_1gpu_a.json); every N must reproduce them.
C3_CHECKSUMS = (42018717167, 21000152364599771)


def workload(cfg):
    return synth.family(cfg["seed"], cfg["n_seqs"], cfg["length"]), matrices.blosum62()


def config_dict(cfg, world, extra=None):
    n = cfg["n_seqs"]
    c = {"workload": "all-vs-all pairwise scoring (guide tree), %d synthetic %d-aa proteins (%d pairs), BLOSUM62 affine "
                     "[-11,-1], global, score per pair (BASELINE configs[%d]%s)"
                     % (n, cfg["length"], n * (n - 1) // 2, 2 if cfg is C3 else 1,
                        ", the Target of north_star" if cfg is C3 else ""),
         "n_seqs": n, "seq_len": cfg["length"], "pairs": n * (n - 1) // 2,
         "l2": "256 MiB buffer written between timed steps; the 200 MB score vector alone exceeds L2"}
    if extra:
        c.update(extra)
    return c


# ---- clocks sampling -------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region (NVML, every ~3 ms)."""

    def __init__(self, index=0):
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop = False
        self.th = threading.Thread(target=self._run, daemon=True)
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        if nv is None:
            return
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        while not self.stop:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.mx.append(mx)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.003)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": float(max(self.mx)) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ---- CPU baselines ---------------------------------------------------------------------------
def _ref_worker(args):
    """One worker of the reference arm: cext_build_scores + cext_align_global per pair, i.e. the
    reference's own compiled C path (oracle/_ref), exactly the arrays PairwiseAligner builds."""
    import oracle
    seqs, S, pairs, gaps = args
    cx = oracle.ref_cext()
    cells = 0
    scores = []
    onehot = {}
    for i, j in pairs:
        for k in (i, j):
            if k not in onehot:
                p = np.zeros((len(seqs[k]), S.shape[0]), np.float32)
                p[np.arange(len(seqs[k])), seqs[k]] = 1.0
                nz = np.full(p.shape, -1, np.intp)
                nz[:, 0] = seqs[k]
                onehot[k] = (p, nz)
        (p1, nz1), (p2, nz2) = onehot[i], onehot[j]
        L1, L2 = p1.shape[0], p2.shape[0]
        m = np.zeros((L1, L2), np.float32)
        cx.cext_build_scores([p1], [p2], [nz1], [nz2], [S], m)
        g1, g2 = oracle.gap_arrays(L1, L2, gaps)
        o, t = oracle.ref_init_borders("global", g1, g2, L1, L2)
        z = np.zeros((L1 + 1, L2 + 1), np.uint8)
        cx.cext_align_global(m, g1, g2, o, t, z)
        scores.append(float(o[L1, L2].max()))
        cells += L1 * L2
    return cells, scores


def reference_arm(args, emit=None):
    """--impl reference: the reference's CPU implementation of the path on all host cores, on the
    same workload (the C3 family), each step a bounded random sample of its pairs."""
    import multiprocessing as mp
    import oracle
    seqs, S = workload(C3)
    n = len(seqs)
    cores = os.cpu_count() or 1
    kind = "reference" if oracle.ref_cext() is not None else "port"
    rng = np.random.default_rng(0)
    per_step = 600 * cores                      # bounded sample: ~1.5 s of work per core and step
    flat, offs = synth.pack(seqs)

    def one_step(pool, k):
        i, j = rng.integers(0, n, per_step), rng.integers(0, n, per_step)
        j = np.where(i == j, (j + 1) % n, j)
        pi, pj = np.minimum(i, j).astype(np.int32), np.maximum(i, j).astype(np.int32)
        sub = []
        if kind == "reference":
            for c in np.array_split(np.arange(per_step), cores):        # ship only the sequences a worker needs
                ids = np.unique(np.concatenate([pi[c], pj[c]]))
                remap = {int(g): k2 for k2, g in enumerate(ids)}
                sub.append(([seqs[g] for g in ids], S, [(remap[int(a)], remap[int(b)]) for a, b in zip(pi[c], pj[c])],
                            C3["gaps"]))
        t0 = time.perf_counter()
        if kind == "reference":
            res = pool.map(_ref_worker, sub)
            cells = sum(r[0] for r in res)
        else:
            oracle.align_batch("global", flat, offs, pi, pj, S, C3["gaps"])
            cells = int((offs[1:] - offs[:-1])[pi] @ (offs[1:] - offs[:-1])[pj])
        return cells, time.perf_counter() - t0

    with mp.get_context("fork").Pool(cores) as pool:
        for k in range(args.warmup):
            one_step(pool, k)
        tot_c, tot_t = 0, 0.0
        for k in range(args.steps):
            c, t = one_step(pool, k)
            tot_c += c
            tot_t += t
    gcups = tot_c / tot_t / 1e9
    used = cores if kind == "reference" else 1
    line = {"impl": "reference", "metric": "GCUPS all-vs-all affine DP", "value": gcups, "unit": "GCUPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(C3, max(args.gpus, 1)),
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": used, "kind": kind,
                             "sample": "%d random pairs of the workload per step x %d steps, cext_build_scores + "
                                       "cext_align_global per pair" % (per_step, args.steps)},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    (emit or (lambda o: print(json.dumps(o))))(line)


def cpu_baseline_port(seqs, S, gaps, seconds=12.0):
    """The oracle's scalar C loop on one core over a bounded sample of the same pairs."""
    import oracle
    flat, offs = synth.pack(seqs)
    n = len(seqs)
    rng = np.random.default_rng(1)

    def draw(k):
        i = rng.integers(0, n, k)
        j = rng.integers(0, n, k)
        keep = i != j
        return np.minimum(i, j)[keep].astype(np.int32), np.maximum(i, j)[keep].astype(np.int32)

    pi, pj = draw(200)
    t0 = time.perf_counter()
    oracle.align_batch("global", flat, offs, pi, pj, S, gaps)
    dt = time.perf_counter() - t0
    pi, pj = draw(int(max(200, 200 * seconds / max(dt, 1e-3))))
    lens = offs[1:] - offs[:-1]
    t0 = time.perf_counter()
    oracle.align_batch("global", flat, offs, pi, pj, S, gaps)
    dt = time.perf_counter() - t0
    cells = int((lens[pi] * lens[pj]).sum())
    return {"value": cells / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": "port",
            "sample": "%d random pairs of the workload (%.1f s), oracle/praline_oracle.c scalar loop incl. traceback"
                      % (len(pi), dt)}


def measured_peaks():
    """MEASURED_PEAKS.json (driver-written: measured copy bandwidth and bf16 GEMM rate of this pool's B200s)."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def _tool(script, argv, timeout=900, env=None):
    """Run a tools/ script in its own process and return the JSON lines it printed."""
    try:
        e = dict(os.environ)
        e.update(env or {})
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            e.pop(k, None)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", script)] + [str(a) for a in argv],
                             capture_output=True, text=True, timeout=timeout, env=e)
        lines = [json.loads(l) for l in out.stdout.strip().splitlines() if l.startswith("{")]
        if not lines:
            return [{"error": (out.stderr or "no output")[-300:]}]
        return lines
    except Exception as e:   # the DP numbers stand on their own
        return [{"error": str(e)[:300]}]


def msa_e2e(n=50, length=300, preprofile="global", msa="tree", arms=("gpu", "cpu1", "cpun")):
    """Second half of the BASELINE metric: wall time of the reference's MSA workflow
    (`praline --preprofile-global --msa-tree`) on the GPU manager, on the reference's stock Manager
    (`-t 1`) and on its ParallelExecutionManager with all host cores (`-t $(nproc)`, cmd.py:41-44);
    outputs must be byte-identical (SHA-256 of the FASTA).  Needs the reference package
    (baseline/_ref); returns None when it is absent."""
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "praline")):
        return None
    res = {"n_seqs": n, "length": length, "preprofile": preprofile, "msa": msa}
    shas = set()
    for arm in arms:
        r = _tool("msa_e2e.py", [n, length, preprofile, msa, arm])[-1]
        if "error" in r:
            res[arm] = r
            continue
        shas.add(r["sha256"])
        res[arm] = {k: r[k] for k in ("wall_s", "first_s", "threads", "batched_requests") if k in r}
        res["cores"] = r.get("cores")
    res["identical"] = len(shas) == 1
    g = res.get("gpu", {}).get("wall_s")
    for arm in ("cpu1", "cpun"):
        if g and res.get(arm, {}).get("wall_s"):
            res["speedup_vs_" + arm] = res[arm]["wall_s"] / g
    return res


# ---- our arm ---------------------------------------------------------------------------------
def timed_allpairs(eng, batch, S, S_dev, gaps, mode, rank, world, steps, warmup, s_host, flush, sync, sampler=None):
    """W warm-up + K timed steps of the sharded all-vs-all with the in-place all-gather; CUDA events
    per step on the launching stream, L2 flushed between steps.  Returns (ms per step of this rank,
    score buffer, plan, launches)."""
    import torch
    from praline_b200 import parallel
    plan = eng.allpairs_plan(batch, s_host, gaps, mode, (rank, world))
    sc = parallel.ShardedCondensed(plan[3], eng.device)

    def step():
        eng.allpairs_scores(batch, S_dev, S.shape[0], gaps, mode=mode, shard=(rank, world), out=sc, plan=plan,
                            S_host=s_host)
        sc.allgather(rank)

    for _ in range(max(warmup, 0)):
        step()
    sync()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    l0 = eng.launches

    def run():
        sync()
        for k in range(steps):
            flush.fill_(k & 0xff)            # L2 flush, outside the timed events
            torch.cuda.synchronize(eng.device)
            ev[k][0].record()
            step()
            ev[k][1].record()
        sync()

    if sampler is not None:
        with sampler:
            run()
    else:
        run()
    ms = sum(a.elapsed_time(b) for a, b in ev) / max(steps, 1)
    return ms, sc, plan, eng.launches - l0


def parity_check(eng, sc, seqs, S, gaps, rank, world, expect=None, per_rank=256):
    """Rank 0 oracle-checks `per_rank` random slots of every rank's slice of the gathered vector;
    every rank's copy must carry the same integer checksums (and `expect`, when recorded)."""
    import torch
    import torch.distributed as dist
    n = len(seqs)
    vec = sc.condensed()
    iv = vec.to(torch.int64)
    idx = torch.arange(vec.numel(), device=vec.device, dtype=torch.int64) % 1000003
    sums = torch.stack([iv.sum(), (iv * idx).sum(), (vec != iv.to(torch.float32)).sum().to(torch.int64)])
    allsums = [sums.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(allsums, sums)
    res = {"slots_checked": 0, "oracle_mismatches": 0, "ranks_identical": True, "checksums": None, "ok": True}
    if rank == 0:
        import oracle
        ref = allsums[0].cpu().tolist()
        res["ranks_identical"] = all(a.cpu().tolist() == ref for a in allsums)
        res["checksums"] = ref[:2]
        res["non_integer_scores"] = int(ref[2])
        if expect is not None:
            res["matches_recorded_1gpu_checksums"] = (list(expect) == ref[:2])
        rng = np.random.default_rng(12345)
        slots = np.concatenate([rng.integers(sc.cuts[r], max(sc.cuts[r + 1], sc.cuts[r] + 1), per_rank)
                                for r in range(world) if sc.cuts[r + 1] > sc.cuts[r]])
        # condensed slot -> (i, j), i < j (np.triu_indices order)
        i = (n - 2 - np.floor(np.sqrt(-8.0 * slots + 4.0 * n * (n - 1) - 7) / 2.0 - 0.5)).astype(np.int64)
        j = (slots + i + 1 - n * (n - 1) // 2 + (n - i) * ((n - i) - 1) // 2).astype(np.int64)
        flat, offs = synth.pack(seqs)
        want = oracle.align_batch("global", flat, offs, i.astype(np.int32), j.astype(np.int32), S, gaps)
        got = sc.buf[torch.from_numpy(sc.where(slots)).to(sc.buf.device)].cpu().numpy()
        res["slots_checked"] = int(len(slots))
        res["oracle_mismatches"] = int((got != want).sum())
        res["ok"] = bool(res["oracle_mismatches"] == 0 and res["ranks_identical"] and res["non_integer_scores"] == 0
                         and res.get("matches_recorded_1gpu_checksums", True))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline + roofline only (profiling runs)")
    ap.add_argument("--full-msa", action="store_true", help="also time the MSA workflow at 500 sequences")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # stdout carries exactly ONE JSON line: anything a library prints there meanwhile (NCCL's version
    # banner, for one) goes to stderr instead; emit() restores the descriptor for the line itself
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(_real_stdout, 1)
        print(json.dumps(obj))
        sys.stdout.flush()
        os.dup2(2, 1)

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, emit)
        return

    import torch
    import torch.distributed as dist
    from praline_b200 import get_engine, parallel

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = get_engine(local_rank)
    dev = eng.device
    seqs, S = workload(C3)
    gaps, mode = C3["gaps"], C3["mode"]
    n = len(seqs)
    n_pairs = n * (n - 1) // 2

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident-input arm: sequences, matrix and tile plan already in HBM ------------------
    batch = eng.batch(seqs)
    S_dev = eng.dev(S)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lens = batch.lens
    total_cells = int((int(lens.sum()) ** 2 - int((lens ** 2).sum())) // 2)
    clk = ClockSampler(local_rank)
    ms_rank, sc, plan, launches = timed_allpairs(eng, batch, S, S_dev, gaps, mode, rank, world, args.steps, args.warmup,
                                                 S, flush, sync, clk)
    my_cells = plan[2]
    ms_per_step = allmax(ms_rank)
    gcups = total_cells / (ms_per_step * 1e-3) / 1e9
    check = parity_check(eng, sc, seqs, S, gaps, rank, world, expect=C3_CHECKSUMS)

    # ---- end-to-end arm: host buffers in, host scores out, every step --------------------------
    # The call a user of the library makes: sequences on the host -> Engine.batch (pinned H2D) ->
    # allpairs_plan (cached per lengths/shard) -> kernels -> in-place all-gather -> every rank copies ITS
    # slice of the vector into one pinned host vector shared by the ranks of the node (SharedHostVector).
    host = parallel.SharedHostVector(n_pairs, rank, world)
    h2d = d2h = 0

    def step_e2e():
        nonlocal h2d, d2h
        b = eng.batch(seqs)                                  # pinned host -> device
        sd = eng.dev(S)
        pl = eng.allpairs_plan(b, S, gaps, mode, (rank, world))
        tiles_new = 0 if pl[4] else sum(t.nbytes for t in pl[0].values())
        o, (lo, hi), _ = eng.allpairs_scores(b, sd, S.shape[0], gaps, mode=mode, shard=(rank, world), out=sc, plan=pl,
                                             S_host=S)
        host.tensor[lo:hi].copy_(sc.slice_of(rank), non_blocking=True)   # device -> shared pinned host, own slice
        sc.allgather(rank)                                   # every rank also holds the full vector on its device
        torch.cuda.synchronize(dev)
        h2d = b.h2d_bytes + S.nbytes + tiles_new
        d2h = (hi - lo) * 4

    for _ in range(max(args.warmup, 1)):
        step_e2e()
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    sync()
    e2e_s = allmax(time.perf_counter() - t0)
    e2e_gcups = total_cells * args.steps / e2e_s / 1e9
    tb = torch.tensor([h2d, d2h], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tb)
    h2d_all, d2h_all = int(tb[0].item()), int(tb[1].item())
    if rank == 0:      # the host vector assembled from the ranks' slices is the same vector
        pick = np.random.default_rng(5).integers(0, n_pairs, 100000)
        dev_vals = sc.buf[torch.from_numpy(sc.where(pick)).to(dev)].cpu().numpy()
        check["e2e_host_vector_matches"] = bool(np.array_equal(host.array[pick], dev_vals))
        check["ok"] = bool(check["ok"] and check["e2e_host_vector_matches"])

    # ---- sub-records on every N: the f32 kernel on the Target, the preprofile stage of configs[2] ----
    configs = {}
    eng.use_s16 = False
    f32_ms, sc32, plan32, _ = timed_allpairs(eng, batch, S, S_dev, gaps, mode, rank, world, 1, 1, None, flush, sync)
    eng.use_s16 = True
    f32_ms = allmax(f32_ms)
    same32 = bool(torch.equal(sc32.condensed(), sc.condensed()))
    configs["c3_f32_kernel"] = {"kernel": "k_stream<13,global,score-only> (f32, what non-integer inputs get)",
                                "ms_per_step": f32_ms, "gcups": total_cells / (f32_ms * 1e-3) / 1e9,
                                "identical_to_int16_vector": same32, "steps": 1, "warmup": 1}
    del sc32
    if not args.no_extras:
        # the preprofile stage of configs[2]: every master against all other sequences (99,990,000 ordered pairs), traced
        # on the device into count tables.  Symmetric path: every UNORDERED pair is filled once on the paired-resident
        # traced kernel and walked in both orientations; the pair list is cut by DP cells over the ranks and the
        # full-size tables are summed by one all-reduce.  Checked at full size against the per-master path (one fill
        # per ordered pair, the round-1 path) on a sample of masters taken from every rank's part of the id range.
        small = eng.batch(seqs[:64])
        eng.preprofile_stage(small, S, gaps)                               # warm-up (both kernels, both streams)
        del small
        sync()
        l0 = eng.launches
        t0 = time.perf_counter()
        cnt_dev, where, cells_mine = eng.preprofile_stage(batch, S, gaps, shard=(rank, world))
        cnt_all = parallel.allreduce_counts(cnt_dev)
        sync()
        pre_s = allmax(time.perf_counter() - t0)
        counts_sum = int(cnt_all.sum(dtype=torch.int64).item())
        sample = np.unique(np.concatenate([np.arange(r * n // world, r * n // world + 3) for r in range(world)] + [[n - 1]]))
        ref_cnt, ref_where, _ = eng.preprofile_stage(batch, S, gaps, masters=sample, shard=None)
        same = True
        for mid in sample:
            o1, l1 = where[int(mid)]
            o2, l2 = ref_where[int(mid)]
            same = same and l1 == l2 and bool(torch.equal(cnt_all[o1:o1 + l1 * S.shape[0]], ref_cnt[o2:o2 + l2 * S.shape[0]]))
        pre_cells = float(lens.sum()) ** 2 - float((lens ** 2).sum())
        configs["c3_preprofile"] = {"stage": "preprofile stage of BASELINE configs[2]: 99,990,000 global master-slave alignments "
                                             "-> count tables on the device (preprofile.py:127-154, util/align.py:187-232); "
                                             "one traced fill per unordered pair, two walks (symmetric S, constant gaps)",
                                    "wall_s": pre_s, "gcups": pre_cells / pre_s / 1e9, "cells": pre_cells,
                                    "cells_filled": pre_cells / 2, "gcups_filled": pre_cells / 2 / pre_s / 1e9,
                                    "counts_sum": counts_sum, "counts_sum_expected": 39090321290,
                                    "identical_counts_sum": counts_sum == 39090321290,
                                    "sample_masters_identical_to_per_master_path": bool(same),
                                    "sample_masters": [int(m) for m in sample],
                                    "gpu_launches_rank0": int(eng.launches - l0), "n_gpus": world}
        check["ok"] = bool(check["ok"] and counts_sum == 39090321290 and same)
        del cnt_dev, cnt_all, ref_cnt

    # ---- roofline of the dominant kernel, timed live with CUDA events ---------------------------
    roof = None
    cpu = None
    clk_sum = clk.summary()
    if rank == 0:
        mb = eng.microbench()
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
        for a, b in kev:
            a.record()
            eng.allpairs_scores(batch, S_dev, S.shape[0], gaps, mode=mode, shard=(rank, world), out=sc, plan=plan, S_host=S)
            b.record()
        torch.cuda.synchronize(dev)
        kms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
        rate = my_cells / (kms * 1e-3)                     # cells per second of this rank's launch
        # The binding unit of the packed kernel is the half-rate ALU pipe that executes the DPX instructions:
        # 3 per TWO cells (2 VIADDMNMX.S16x2 + 1 VIMNMX3.S16x2).  peak = the measured VIADDMNMX.S16x2 rate of the
        # box (pgpu_microbench, warp-instr/ns/SM) x 32 lanes x SMs; achieved = 1.5 lane-instructions per cell at
        # the kernel's cell rate.  ncu's sm__inst_executed_pipe_alu of the same kernel adds the few non-DPX ALU
        # instructions (profiles/r02_kstream16r_*).
        alu_peak = mb["viaddmnmx_s16x2"] * 32.0 * sms * 1e9
        mix_peak = mb["cell_mix16"] * 32.0 * sms * 1e9      # bare 5-instruction recurrence of two cells
        issue_peak = mb["fadd"] * 32.0 * sms * 1e9          # full-rate FP32 pipe (FADD), measured: ~3.6 warp-instr/clk/SM
        slots_peak = 4.0 * mb["sm_mhz"] * 1e6 * 32.0 * sms  # issue slots: 4 warp-instructions per clock and SM at the measured clock
        # ALU-pipe lane-instructions per cell of the shipped kernel, from the committed ncu capture of the same build
        # (tools/roofline_counters.py -> profiles/r02_roofline_counters.json): 1.5 DPX + the kernel's other ALU-pipe
        # instructions.  frac = that x the cell rate timed HERE / the ALU-pipe rate measured HERE: it reproduces ncu's
        # sm__inst_executed_pipe_alu percentage of the capture when the cell rates agree.
        cnt = {}
        try:
            with open(os.path.join(ROOT, "profiles", "r02_roofline_counters.json")) as f:
                cnt = json.load(f).get("k_stream16r_13", {})
        except Exception:
            pass
        alu_per_cell = float(cnt.get("alu_lane_instr_per_cell", 1.5))
        alu_ach = rate * alu_per_cell
        roof = {"bound": "cuda_core_alu_pipe", "achieved": alu_ach / 1e12, "peak": alu_peak / 1e12, "unit": "Tlane-op/s",
                "frac": alu_ach / alu_peak,
                "frac_note": "of measured: utilisation of the binding unit, the half-rate ALU pipe that executes the DPX "
                             "instructions = ALU-pipe lane-instructions per cell of this kernel (%.3f: 1.5 DPX -- 2 VIADDMNMX.S16x2 "
                             "+ 1 VIMNMX3.S16x2 per two cells -- plus its other ALU-pipe instructions, from the committed ncu "
                             "capture) x the cell rate timed in this run / the VIADDMNMX.S16x2 rate of this box (pgpu_microbench, "
                             "%.2f warp-instr/ns/SM); ncu's sm__inst_executed_pipe_alu of the capture: %.1f %%"
                             % (alu_per_cell, mb["viaddmnmx_s16x2"], float(cnt.get("alu_pipe_pct", float("nan")))),
                "frac_recurrence_only": rate * 1.5 / alu_peak,
                "frac_of_bare_recurrence_mix": rate * 2.5 / mix_peak,
                "frac_of_issue": rate * float(cnt.get("lane_instr_per_cell", 2.5)) / slots_peak,
                "frac_of_issue_recurrence_only": rate * 2.5 / slots_peak,
                "frac_notes": "recurrence_only: the 1.5 DPX lane-instructions per cell alone on the ALU pipe; bare mix: 2.5 "
                              "lane-instructions per cell against the bare 5-instruction packed recurrence (no loads, shuffles or "
                              "loop) measured on this box; issue: all issued lane-instructions per cell (ncu capture) against 4 "
                              "issue slots per clock and SM at the clock pgpu_microbench measured = ncu's issue-slot utilisation; "
                              "issue_recurrence_only: the 2.5 recurrence instructions alone",
                "ncu_capture": {k: cnt.get(k) for k in ("capture", "workload", "ncu_duration_ms", "plain_run_kernel_ms",
                                                        "gcups_under_ncu", "alu_pipe_pct", "fma_pipe_pct", "fmaheavy_cycles_pct",
                                                        "issue_active_pct", "lane_instr_per_cell", "alu_lane_instr_per_cell",
                                                        "recurrence_share_of_issued", "dpx_share_of_alu_pipe")},
                # SURVEY 8d's W = 11 f32 add/max per cell against the measured f32 add issue rate: a note only -- above 1
                # for the packed kernel because one DPX instruction does two cells and fuses add+max
                "survey_w11_note": {"w_ops_per_cell": W_FLOPS_PER_CELL, "achieved": rate * W_FLOPS_PER_CELL / 1e12,
                                    "peak": issue_peak / 1e12, "frac": rate * W_FLOPS_PER_CELL / issue_peak},
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel in the ncu --set full capture, scaled by
                # cells to this launch; algorithmic bytes per launch = sequences + tiles in, 4 B per pair out
                "traffic": (float(cnt["dram_bytes"]) * my_cells / float(cnt["cells"])) if cnt.get("dram_bytes") else None,
                "algorithmic_bytes": int(batch.h2d_bytes + sum(t.nbytes for t in plan[0].values()) + 4 * (plan[1][1] - plan[1][0])),
                "kernel": "k_stream16r<13> (packed s16x2 DPX, paired residents)", "kernel_ms": kms,
                "gcups_kernel": rate / 1e9, "pipe_rates": mb,
                "note": "HBM is not the bound (4 B per pair out, 8e4 cells per pair); tensor cores have nothing to contract"}

    if rank == 0 and not args.no_extras and world == 1:
        # BASELINE configs[1]: 1,000 x 300 aa on one B200, packed int16 and f32 kernels
        seqs2, _ = workload(C2)
        b2 = eng.batch(seqs2)
        cells2 = int((int(b2.lens.sum()) ** 2 - int((b2.lens ** 2).sum())) // 2)
        nosync = lambda: torch.cuda.synchronize(dev)
        ms16, sc16, _, _ = timed_allpairs(eng, b2, S, S_dev, gaps, mode, 0, 1, 5, 3, S, flush, nosync)
        eng.use_s16 = False
        ms32, sc32b, _, _ = timed_allpairs(eng, b2, S, S_dev, gaps, mode, 0, 1, 3, 2, None, flush, nosync)
        eng.use_s16 = True
        configs["c2"] = {"workload": config_dict(C2, 1)["workload"], "cells": cells2,
                         "int16_kernel": {"ms_per_step": ms16, "gcups": cells2 / (ms16 * 1e-3) / 1e9, "kernel": "k_stream16r<10>"},
                         "f32_kernel": {"ms_per_step": ms32, "gcups": cells2 / (ms32 * 1e-3) / 1e9, "kernel": "k_stream<10>"},
                         "identical": bool(torch.equal(sc16.condensed(), sc32b.condensed()))}
        # traced variant on the same workload (fill with packed traceback + K4 walk, device time)
        tpi, tpj = synth.all_pairs(C2["n_seqs"])
        tpi, tpj = tpi[:120000], tpj[:120000]
        tcells = int((b2.lens[tpi] * b2.lens[tpj]).sum())
        eng.align_pairs(b2, tpi, tpj, S, gaps, mode=mode, want_paths=True, resident="one", device_only=True)
        torch.cuda.synchronize(dev)
        ta, tb_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ta.record()
        eng.align_pairs(b2, tpi, tpj, S, gaps, mode=mode, want_paths=True, resident="one", device_only=True)
        tb_.record()
        torch.cuda.synchronize(dev)
        configs["c2"]["traced_gcups"] = tcells / (ta.elapsed_time(tb_) * 1e-3) / 1e9
        # the caller of the scores: distance matrix + clustering kernel on the device, C3 size
        ga, gb, gc = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        ga.record()
        dmat = eng.tree_distance(sc, n)
        gb.record()
        merges = eng.cluster_merge_order(dmat, "average")
        gc.record()
        torch.cuda.synchronize(dev)
        configs["c3_guide_tree"] = {"distance_ms": ga.elapsed_time(gb), "cluster_ms_incl_d2h": gb.elapsed_time(gc),
                                    "merges": len(merges), "note": "average linkage, merge order of util/cluster.py on %d sequences" % n}
        del dmat
        # the score-matrix kernel on tensor cores (north star (1)): HBM GB/s of k_build_rows_tc against the
        # measured copy bandwidth, on one wave of 60 depth-50 profiles of length 400 (tolerance mode)
        try:
            profs = [synth.profile_from_counts(synth.count_profile(2000 + k, 400, 50, 20, 27)) for k in range(60)]
            pbat = eng.profile_batch(profs)
            ppi, ppj = synth.all_pairs(len(profs))
            eng.align_profile_pairs(pbat, ppi, ppj, S, gaps, mode=mode, fast=True)
            eng.take_trace()
            eng.trace_on = True
            eng.align_profile_pairs(pbat, ppi, ppj, S, gaps, mode=mode, fast=True)
            eng.trace_on = False
            tr = eng.take_trace()
            tc = [(nm, ms) for nm, ms in tr if nm.startswith("score rows tc")]
            fed = [ms for nm, ms in tr if nm == "matrix-fed stream"]
            nbytes = sum(int(nm.split("(")[1].split(" ")[0]) for nm, _ in tc)
            tc_ms = sum(ms for _, ms in tc)
            hbm_meas = measured_peaks().get("hbm_gbs")
            hbm_peak = hbm_meas or 6650.0     # B200_PROFILING.md's fallback when the driver's file is absent
            pcells = float((pbat.lens[ppi] * pbat.lens[ppj]).sum())
            roof["score_rows_tc"] = {"kernel": "k_build_rows_tc (tcgen05 kind::tf32, hi/lo split, TMEM accumulator)",
                                     "bound": "hbm", "bytes_written": nbytes, "kernel_ms": tc_ms,
                                     "achieved": nbytes / (tc_ms * 1e-3) / 1e9 if tc_ms else None,
                                     "peak": hbm_peak, "unit": "GB/s",
                                     "frac": (nbytes / (tc_ms * 1e-3) / 1e9 / hbm_peak) if tc_ms and hbm_peak else None,
                                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if hbm_meas else "6650 GB/s (of fallback)",
                                     "matrix_fed_stream_ms": sum(fed),
                                     "profile_batch_gcups_device": pcells / ((tc_ms + sum(fed)) * 1e-3) / 1e9 if tc_ms else None}
        except Exception as e:   # the headline numbers stand on their own
            roof["score_rows_tc"] = {"error": str(e)[:200]}
        # the exact-order score rows (what bounds BASELINE configs[3] in exact mode): FP32-pipe utilisation of
        # k_build_rows_x2 on 120 dense depth-2000 preprofiles of length 400.  Algorithmic work = 2 individually rounded
        # f32 operations (one multiply, one add; cext.c:89) per (nonzero of row y) x (nonzero of row x) per cell;
        # peak = the FADD rate of this box (pgpu_microbench) x 32 lanes.
        try:
            profs = [synth.profile_from_counts(synth.count_profile(3000 + k, 400, 2000, 20, 27)) for k in range(120)]
            pbat = eng.profile_batch(profs)
            ppi, ppj = synth.all_pairs(len(profs))
            eng.align_profile_pairs(pbat, ppi, ppj, S, gaps, mode=mode, fast=False)
            eng.take_trace()
            eng.trace_on = True
            eng.align_profile_pairs(pbat, ppi, ppj, S, gaps, mode=mode, fast=False)
            eng.trace_on = False
            tr = eng.take_trace()
            ex_ms = sum(ms for nm, ms in tr if nm.startswith("score rows exact"))
            fed_ms = sum(ms for nm, ms in tr if nm == "matrix-fed stream")
            nnz = [(p_ != 0).sum(axis=1).astype(np.float64) for p_ in profs]
            tot = np.array([v.sum() for v in nnz])
            lane_ops = 2.0 * float((tot[ppi] * tot[ppj]).sum())
            pcells = float((pbat.lens[ppi] * pbat.lens[ppj]).sum())
            fp_peak = mb["fadd"] * 1e9 * 32 * sms / 1e12 if mb.get("fadd") else None
            roof["score_rows_exact"] = {"kernel": "k_build_rows_x2 (packed f32x2: FFMA2 with a -0 addend + FADD2, reference evaluation order)",
                                        "bound": "cuda_core_fp32_pipe", "kernel_ms": ex_ms,
                                        "achieved": lane_ops / (ex_ms * 1e-3) / 1e12 if ex_ms else None,
                                        "peak": fp_peak, "unit": "Tlane-op/s",
                                        "frac": (lane_ops / (ex_ms * 1e-3) / 1e12 / fp_peak) if ex_ms and fp_peak else None,
                                        "frac_note": "useful individually rounded f32 operations per second / measured FADD lane rate; "
                                                     "the kernel's own pipe time also covers the 416-column padding of 400-column "
                                                     "residents (ncu capture profiles/r02_kbuildrows_x2_*: fmaheavy pipe 82.8 % busy)",
                                        "dense_syms": pbat.dense_syms(), "matrix_fed_stream_ms": fed_ms,
                                        "profile_batch_gcups_device": pcells / ((ex_ms + fed_ms) * 1e-3) / 1e9 if ex_ms else None}
        except Exception as e:
            roof["score_rows_exact"] = {"error": str(e)[:200]}
        # BASELINE configs[4]: 20 kb x 20 kb DNA profiles on the intra-task wavefront path (own process)
        c5 = _tool("run_c5.py", [20000, 3], timeout=600)
        configs["c5"] = {"workload": "long nucleotide profile x profile alignment, 20 kb x 20 kb, depth-8 count profiles (BASELINE configs[4])",
                         "cases": c5}
        if not args.no_cpu_baseline:
            cpu = cpu_baseline_port(seqs, S, gaps)
            # BASELINE configs[3]: progressive merge of 2,000 x 400 aa through the reference's workflow (own process)
            configs["c4"] = {"workload": "progressive profile-profile merge, 2,000 x 400 aa, global preprofiles, guide tree, "
                                         "merge_mode='semiglobal' (BASELINE configs[3]); whole workflow wall time on GpuBatchManager",
                             "exact_profile_scores": _tool("run_c4.py", [2000, 400, 40], timeout=900)[-1],
                             "tolerance_profile_scores": _tool("run_c4.py", [2000, 400, 0], timeout=900,
                                                               env={"PGPU_FAST_PROFILES": "1"})[-1]}
            c4 = configs["c4"]
            # does the tolerance mode (profile scores within 1e-5, north_star) change the guide tree or any column?
            c4["tolerance_mode_alignment_identical_to_exact"] = bool(
                c4["exact_profile_scores"].get("fasta_md5") is not None and
                c4["exact_profile_scores"].get("fasta_md5") == c4["tolerance_profile_scores"].get("fasta_md5"))
            c4["bound_note"] = ("exact mode is bound by the reference's evaluation order of the profile score rows in the "
                                "guide-tree stage: 2.0e6 dense profile pairs x 1.6e5 cells x ~400 individually rounded "
                                "mul+add terms (cext.c:63-95) = 2.6e14 f32 operations >= 6.9 s at the FP32 pipe peak of one B200 "
                                "on top of the ~4 s the rest of the workflow takes; roofline.score_rows_exact gives the pipe "
                                "fraction the packed f32x2 kernel reaches (DESIGN.md)")
            # BASELINE configs[0] / second half of the metric: MSA wall time next to the host-CPU reference
            configs["msa_e2e"] = {"tree_50": msa_e2e(50, 300, "global", "tree"),
                                  "cli_default_50": msa_e2e(50, 300, "dummy", "ad_hoc"),
                                  "larger_runs": "profiles/r02_msa_e2e_full.json (200 and 500 sequences, --full-msa)"}
            if args.full_msa:
                configs["msa_e2e"]["tree_200"] = msa_e2e(200, 300, "global", "tree", arms=("gpu", "cpun"))
                configs["msa_e2e"]["tree_500"] = msa_e2e(500, 300, "global", "tree", arms=("gpu", "cpun"))

    if rank == 0:
        line = {"metric": "GCUPS all-vs-all affine DP", "value": gcups, "unit": "GCUPS", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None,
                "dtype": "i16" if plan[5] else "f32",
                "data": "synthetic",
                "config": config_dict(C3, world),
                "parallelism": "pairs sharded by DP cells over %d rank(s)%s"
                               % (world, ", in-place NCCL all-gather of the score slices" if world > 1 else ""),
                "clocks": clk_sum,
                "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all,
                        "note": "host sequences -> pinned H2D -> plan (cached per lengths) -> kernel -> all-gather; every rank "
                                "copies its slice into one pinned host vector shared by the ranks (bytes are sums over ranks)"},
                "gpu_launches": int(launches), "parity_check": check["ok"], "parity": check,
                "roofline": roof, "cpu_baseline": cpu, "configs": configs}
        emit(line)
    ok = torch.tensor([1 if (rank != 0 or check["ok"]) else 0], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    host.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if int(ok.item()) == 0:
        sys.stderr.write("bench.py: PARITY CHECK FAILED: %s\n" % json.dumps(check))
        sys.exit(3)


if __name__ == "__main__":
    main()
