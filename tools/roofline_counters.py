"""Derives the per-cell instruction figures bench.py's `roofline` uses from committed ncu captures (no GPU):

    python tools/roofline_counters.py > profiles/r02_roofline_counters.json

For every capture: kernel duration, cells (from the plain run's log of the same command), total / ALU-pipe / FMA-pipe
lane-instructions per cell, issue-slot utilisation, DRAM bytes.  ALU-pipe warp-instructions = pct_of_peak x 2 per clock
and SM (the half-rate pipe: 4 sub-partitions x 0.5) x active cycles x SMs."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMS = 148


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def raw(path):
    rows = list(csv.reader(open(os.path.join(ROOT, path))))
    m = {h: v for h, v in zip(rows[0], rows[2])}
    m["_units"] = {h: u for h, u in zip(rows[0], rows[1])}
    return m


def nbytes(m, key):
    return float(m[key]) * UNIT.get(m["_units"][key], 1.0)


def cells_from_log(path):
    d = json.loads(open(os.path.join(ROOT, path)).read().strip().splitlines()[-1])
    return d["gcups_allpairs"] * 1e9 * d["allpairs_ms"] * 1e-3, d


out = {}
for key, rawf, logf, recur in (("k_stream16r_13", "profiles/r02_kstream16r13_raw.csv", "profiles/r02_tree3000.log", 2.5),
                               ("k_stream16r_10", "profiles/r02_kstream16r10_raw.csv", "profiles/r02_tree1000.log", 2.5)):
    m = raw(rawf)
    cells, log = cells_from_log(logf)
    cyc = float(m["sm__cycles_active.avg"])
    alu = float(m["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]) / 100 * 2.0 * cyc * SMS
    fma = float(m["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]) / 100 * 4.0 * cyc * SMS
    tot = float(m["smsp__inst_executed.sum"])
    dur = float(m["gpu__time_duration.sum"])
    out[key] = {
        "capture": rawf, "plain_run_log": logf, "workload": "%d x %d aa all-vs-all" % (log["n_seqs"], log["length"]),
        "ncu_duration_ms": dur, "plain_run_kernel_ms": log["allpairs_ms"], "cells": cells,
        "gcups_under_ncu": cells / dur / 1e6,
        "alu_pipe_pct": float(m["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]),
        "fma_pipe_pct": float(m["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]),
        "fmaheavy_cycles_pct": float(m["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"]),
        "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
        "lane_instr_per_cell": tot * 32 / cells, "alu_lane_instr_per_cell": alu * 32 / cells,
        "fma_lane_instr_per_cell": fma * 32 / cells, "recurrence_lane_instr_per_cell": recur,
        "recurrence_share_of_issued": recur / (tot * 32 / cells),
        "dpx_share_of_alu_pipe": 1.5 / (alu * 32 / cells),
        "dram_bytes": nbytes(m, "dram__bytes_read.sum") + nbytes(m, "dram__bytes_write.sum"),
    }
print(json.dumps(out, indent=1))
