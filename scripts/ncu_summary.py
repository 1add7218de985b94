"""Summarise an .ncu-rep (raw page) into the handful of numbers we track: duration, pipe utilisation,
issue rate, stall mix, memory traffic.  Usage: python scripts/ncu_summary.py report.ncu-rep [kernel-substr]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "gpc__cycles_elapsed.avg.per_second",
        "launch__registers_per_thread", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active", "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum "]
ik = hdr.index("Kernel Name")
for r in rows[2:]:
    if sub and sub not in r[ik]:
        continue
    print("==", r[ik][:100])
    for i, h in enumerate(hdr):
        hh = h + " "
        if any(k in hh for k in KEYS) and ".min" not in h and ".max" not in h and "Not Issued" not in h:
            try:
                v = float(r[i])
            except ValueError:
                continue
            if "issue_stalled" in h and v < 0.15:
                continue
            print(f"  {h} [{units[i]}] = {r[i]}")
