#!/usr/bin/env python
"""bench.py -- GCUPS of the all-vs-all affine-gap DP (BASELINE.json metric) on N B200s.

A "step" is one pass of the hot path over the whole pair list of BASELINE config 2:
1,000 synthetic 300-residue proteins (seed 2 family), all 499,500 unordered
pairs, BLOSUM62, gaps [-11, -1], global mode, score per pair (what GuideTreeBuilder needs).
With N > 1 the pair list is sharded by DP cells over the ranks and the condensed score vector
is assembled with one NCCL all-gather per step.

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

WORKLOAD = dict(n_seqs=1000, length=300, seed=2, gaps=[-11.0, -1.0], mode="global")
W_FLOPS_PER_CELL = 11.0   # 7 add + 4 max of the reference recurrence


def n_seqs_for(world):
    """Weak scaling: the pair count (the per-GPU work) stays that of configs[1] per rank, so the
    sequence count grows with sqrt(world): 1000, 1414, 2000, 2828 at 1, 2, 4, 8 GPUs."""
    return int(round(WORKLOAD["n_seqs"] * np.sqrt(world)))


def workload(world=1):
    seqs = synth.family(WORKLOAD["seed"], n_seqs_for(world), WORKLOAD["length"])
    return seqs, matrices.blosum62()


def config_dict(extra=None, world=1):
    n = n_seqs_for(world)
    c = {"workload": "all-vs-all pairwise scoring, %d synthetic 300-aa proteins (%d pairs), "
                     "BLOSUM62 affine [-11,-1], global, score per pair (BASELINE configs[1]%s)"
                     % (n, n * (n - 1) // 2, "" if world == 1 else ", pair count scaled x%d" % world),
         "n_seqs": n, "seq_len": WORKLOAD["length"], "pairs": n * (n - 1) // 2,
         "l2": "256 MiB buffer written between timed steps (inputs are smaller than L2)"}
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
    """--impl reference: the reference's CPU implementation of the path on all host cores."""
    import multiprocessing as mp
    import oracle
    seqs, S = workload(max(args.gpus, 1))
    pi, pj = synth.all_pairs(len(seqs))
    cores = os.cpu_count() or 1
    kind = "reference" if oracle.ref_cext() is not None else "port"
    rng = np.random.default_rng(0)
    per_step = 250 * cores                      # bounded sample: ~3 s of work per core and step
    flat, offs = synth.pack(seqs)

    def one_step(pool, k):
        pick = rng.choice(len(pi), per_step, replace=False)
        t0 = time.perf_counter()
        if kind == "reference":
            chunks = np.array_split(pick, cores)
            res = pool.map(_ref_worker, [(seqs, S, list(zip(pi[c], pj[c])), WORKLOAD["gaps"]) for c in chunks])
            cells = sum(r[0] for r in res)
        else:
            oracle.align_batch("global", flat, offs, pi[pick], pj[pick], S, WORKLOAD["gaps"])
            cells = int((offs[1:] - offs[:-1])[pi[pick]] @ (offs[1:] - offs[:-1])[pj[pick]])
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
            "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict({"sample": "%d random pairs of the workload per step" % per_step},
                                  max(args.gpus, 1)),
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": used, "kind": kind,
                             "sample": "%d random pairs per step x %d steps, cext_build_scores + cext_align_global"
                                       % (per_step, args.steps)},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    (emit or (lambda o: print(json.dumps(o))))(line)


def cpu_baseline_port(seqs, S, seconds=12.0):
    """The oracle's scalar C loop on one core over a bounded sample of the same pairs."""
    import oracle
    flat, offs = synth.pack(seqs)
    pi, pj = synth.all_pairs(len(seqs))
    rng = np.random.default_rng(1)
    pick = rng.choice(len(pi), 400, replace=False)
    t0 = time.perf_counter()
    oracle.align_batch("global", flat, offs, pi[pick], pj[pick], S, WORKLOAD["gaps"])
    dt = time.perf_counter() - t0
    n = int(min(len(pi), max(400, 400 * seconds / max(dt, 1e-3))))
    pick = rng.choice(len(pi), n, replace=False)
    lens = offs[1:] - offs[:-1]
    t0 = time.perf_counter()
    oracle.align_batch("global", flat, offs, pi[pick], pj[pick], S, WORKLOAD["gaps"])
    dt = time.perf_counter() - t0
    cells = int((lens[pi[pick]] * lens[pj[pick]]).sum())
    return {"value": cells / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": "port",
            "sample": "%d random pairs of the workload (%.1f s), oracle/praline_oracle.c scalar loop incl. traceback"
                      % (n, dt)}


def measured_peaks():
    """MEASURED_PEAKS.json (driver-written: measured copy bandwidth and bf16 GEMM rate of this pool's B200s)."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def msa_e2e(n=50, length=300, preprofile="global", msa="tree"):
    """Second half of the BASELINE metric: wall time of the reference's MSA workflow
    (`praline --preprofile-global --msa-tree`, 50 x 300 aa, BASELINE configs[0] scale) on the
    GPU manager and on the reference's own single-process Manager; outputs must be identical.
    Needs the reference package (baseline/_ref); returns None when it is absent."""
    import subprocess
    tool = os.path.join(ROOT, "tools", "msa_e2e.py")
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "praline")):
        return None
    try:
        out = subprocess.run([sys.executable, tool, str(n), str(length), preprofile, msa], capture_output=True,
                             text=True, timeout=600)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:   # the DP numbers above stand on their own
        return {"error": str(e)[:200]}


# ---- our arm ---------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    # NCCL prints its version banner on stdout at some debug levels: keep stdout for the one JSON line
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
    seqs, S = workload(world)
    gaps, mode = WORKLOAD["gaps"], WORKLOAD["mode"]
    n = len(seqs)
    n_pairs = n * (n - 1) // 2

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- resident-input arm: sequences, matrix and tile plan already in HBM ------------------
    batch = eng.batch(seqs)
    S_dev = eng.dev(S)
    go_, ge_ = float(gaps[0]), float(gaps[-1])
    plan = eng.allpairs_tiles(batch, (rank, world), paired=eng.wants_paired(S, go_, ge_, 0, batch))
    slot_cuts = plan[3]
    my_cells = plan[2]
    out = torch.empty(n_pairs, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    total_cells = int(sum(len(seqs[i]) for i in range(n)) ** 2 - sum(len(s) ** 2 for s in seqs)) // 2

    def step_resident():
        eng.allpairs_scores(batch, S_dev, S.shape[0], gaps, mode=mode, shard=(rank, world), out=out, plan=plan,
                            S_host=S)
        if world > 1:
            parallel.allgather_condensed(out, slot_cuts)

    for _ in range(max(args.warmup, 0)):
        step_resident()
    sync()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = eng.launches
    with ClockSampler(local_rank) as clk:
        sync()
        for k in range(args.steps):
            flush.fill_(k & 0xff)            # L2 flush, outside the timed events
            torch.cuda.synchronize(dev)
            ev[k][0].record()
            step_resident()
            ev[k][1].record()
        sync()
    launches = eng.launches - l0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / max(args.steps, 1)
    gcups = total_cells / (ms_per_step * 1e-3) / 1e9

    # ---- end-to-end arm: host buffers in, host scores out, every step --------------------------
    host_out = torch.empty(n_pairs, dtype=torch.float32).pin_memory()
    h2d = d2h = 0

    def step_e2e():
        nonlocal h2d, d2h
        b = eng.batch(seqs)                                  # pinned host -> device
        sd = eng.dev(S)
        pl = eng.allpairs_tiles(b, (rank, world), paired=eng.wants_paired(S, go_, ge_, 0, b))
        o, (lo, hi), _ = eng.allpairs_scores(b, sd, S.shape[0], gaps, mode=mode, shard=(rank, world), plan=pl,
                                             S_host=S)
        if world > 1:
            parallel.allgather_condensed(o, pl[3])
            lo, hi = 0, n_pairs
        host_out[lo:hi].copy_(o[lo:hi], non_blocking=True)   # device -> pinned host
        torch.cuda.synchronize(dev)
        h2d = b.h2d_bytes + S.nbytes + sum(t.nbytes for t in pl[0].values())
        d2h = (hi - lo) * 4

    for _ in range(max(args.warmup, 1)):
        step_e2e()
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    sync()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_gcups = total_cells * args.steps / float(te.item()) / 1e9

    # ---- roofline of the dominant kernel (k_stream), timed live with CUDA events ----------------
    roof = None
    cpu = None
    if rank == 0:
        mb = eng.microbench()
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        clk_sum = clk.summary()
        mhz = clk_sum["sm_mhz"] or mb["sm_mhz"]
        # peak f32 add/max lane-ops per second on this box: the measured
        # full-rate f32 issue (FADD, warp-instructions / clk / SM) x 32 lanes x SMs x SM clock
        # sampled during the timed region
        peak = mb["fadd"] * 32 * sms * 1e9     # warp-instr/ns/SM x lanes x SMs -> lane-ops/s, wall clock
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
        for a, b in kev:
            a.record()
            eng.allpairs_scores(batch, S_dev, S.shape[0], gaps, mode=mode, shard=(rank, world), out=out, plan=plan,
                            S_host=S)
            b.record()
        torch.cuda.synchronize(dev)
        kms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
        achieved = my_cells * W_FLOPS_PER_CELL / (kms * 1e-3)
        # The roofline of this kernel is the CUDA-core pipe its recurrence runs on (not HBM, not
        # tensor).  peak = the BARE recurrence (same instructions, no loads / shuffles / loop) measured on
        # this box by pgpu_microbench: 5 packed DPX/add instructions per 2 cells (cell_mix16) or the
        # 7-instruction f32 cell (cell_mix); achieved = the same instruction count at the kernel's cell rate.
        ipc_cell = 2.5 if plan[5] else 7.0
        mix = mb["cell_mix16"] if plan[5] else mb["cell_mix"]
        ach_ops = my_cells / (kms * 1e-3) * ipc_cell
        peak_ops = mix * 32.0 * sms * 1e9
        roof = {"bound": "cuda_core_issue", "achieved": ach_ops / 1e12, "peak": peak_ops / 1e12, "unit": "Tlane-op/s",
                "frac": ach_ops / peak_ops,
                "frac_note": "of measured: recurrence lane-instructions per second of this kernel (%.1f per cell) / the rate "
                             "of the bare recurrence instruction mix on this box (pgpu_microbench, %.2f warp-instr/ns/SM)"
                             % (ipc_cell, mix),
                # SURVEY 8d's DP-cell roofline (the figure BASELINE's '>= 50 %% of the DP-cell roofline' refers to):
                # algorithmic W = 11 f32 add/max per cell against the measured f32 add issue rate; above 1 for the
                # packed kernel because one DPX instruction does two cells
                "dp_cell_roofline": {"w_ops_per_cell": W_FLOPS_PER_CELL, "achieved": achieved / 1e12, "peak": peak / 1e12,
                                     "unit": "Tlane-op/s", "frac": achieved / peak},
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one ncu --set full capture
                # (profiles/r01_kstream16r_raw.csv: 427,520 B read, 0 B written -- the 2 MB of scores stay in
                # L2; f32 kernel, profiles/r01_kstream_v2_raw.csv: 443,392 B); algorithmic bytes per launch:
                # 0.3 MB sequences + 0.1 MB tiles in, 2.0 MB scores out
                "traffic": 427520 if plan[5] else 443392,
                "note": "DP-cell roofline (SURVEY 8d): algorithmic 11 f32 add/max per cell (the kernel issues 7); "
                        "peak = measured f32 add issue rate (%.2f warp-instr/ns/SM, wall clock) x 32 lanes x %d SMs "
                        "(of measured; SM clock %.0f MHz during the run); f32 max / compare / shift / integer ops "
                        "issue at half that rate on this part (pipe_rates, warp-instr/ns/SM); HBM is not the "
                        "bound: 4 B/pair out" % (mb["fadd"], sms, mhz),
                "issue_bound_frac": (my_cells / (kms * 1e-3)) * (2.5 if plan[5] else 7.0) / 32.0 / (mb["fadd"] * sms * 1e9),
                "issue_bound_note": "fraction of the instruction-issue bound of this kernel's own recurrence at the "
                                    "measured full issue rate: 5 packed DPX/add instructions per 2 cells (int16 "
                                    "kernel) or 7 per cell (f32 kernel); the DPX and max instructions themselves "
                                    "issue at half rate (pipe_rates), which is the binding pipe",
                "kernel": "k_stream16r<10> (packed s16x2, paired residents)" if plan[5] else "k_stream<10,global,score-only>", "kernel_ms": kms,
                "gcups_kernel": my_cells / (kms * 1e-3) / 1e9, "pipe_rates": mb}
        # traced variant (what the preprofile master-slave alignments need): K2 with packed traceback
        # + K4 walk, device time only, on the first 120k pairs of the same workload
        tpi, tpj = synth.all_pairs(n)
        tpi, tpj = tpi[:120000], tpj[:120000]
        tcells = int((batch.lens[tpi] * batch.lens[tpj]).sum())
        eng.align_pairs(batch, tpi, tpj, S, gaps, mode=mode, want_paths=True, resident="one", device_only=True)
        torch.cuda.synchronize(dev)
        ta, tb_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ta.record()
        eng.align_pairs(batch, tpi, tpj, S, gaps, mode=mode, want_paths=True, resident="one", device_only=True)
        tb_.record()
        torch.cuda.synchronize(dev)
        roof["traced_gcups"] = tcells / (ta.elapsed_time(tb_) * 1e-3) / 1e9
        roof["traced_note"] = "120000 pairs, fill with 4-bit traceback + per-pair path walk, device time incl. plan upload"
        # the caller of the scores: distance matrix + clustering kernel on the device
        ga, gb, gc = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        eng.cluster_merge_order(eng.tree_distance(out, n), "average")
        ga.record()
        dmat = eng.tree_distance(out, n)
        gb.record()
        merges = eng.cluster_merge_order(dmat, "average")
        gc.record()
        torch.cuda.synchronize(dev)
        roof["guide_tree"] = {"distance_ms": ga.elapsed_time(gb), "cluster_ms_incl_d2h": gb.elapsed_time(gc),
                              "merges": len(merges), "note": "average linkage, merge order of util/cluster.py on %d sequences" % n}
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
        if not args.no_cpu_baseline:
            cpu = cpu_baseline_port(seqs, S)
            roof["msa_e2e"] = msa_e2e()
            roof["msa_e2e_cli_default"] = msa_e2e(preprofile="dummy", msa="ad_hoc")   # praline in.fa out.aln

    if rank == 0:
        line = {"metric": "GCUPS all-vs-all affine DP", "value": gcups, "unit": "GCUPS", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "i16" if (eng.use_s16 and eng.fits_s16(S, gaps[0], gaps[1], batch.lens) is not None) else "f32",
                "data": "synthetic",
                "config": config_dict({"parallelism": "pairs sharded by DP cells over %d rank(s)%s"
                                       % (world, ", NCCL all-gather of scores" if world > 1 else "")}, world),
                "clocks": clk_sum,
                "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
