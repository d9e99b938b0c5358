"""Compact summary of one kernel from an ncu report (read here, no GPU):
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [row]
prints duration, registers, occupancy, pipe utilisation, issue rate, stall samples, DRAM bytes."""
import csv, subprocess, sys, io
rep = sys.argv[1]
row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rows[0], rows[1], rows[2 + row]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
def g(k):
    return m.get(k, ("?", ""))
print("kernel:", g("Kernel Name")[0][:120])
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "sm__cycles_active.avg"]
for k in keys:
    if k in m:
        print("  %-72s %s %s" % (k, m[k][0], m[k][1]))
st = {}
for h, (v, u) in m.items():
    if "pcsamp_warps_issue_stalled_" in h and "not_issued" not in h:
        try:
            st[h.split("stalled_")[1]] = float(v)
        except ValueError:
            pass
tot = sum(st.values()) or 1.0
print("  stall samples:", ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(st.items(), key=lambda x: -x[1])[:9]))
