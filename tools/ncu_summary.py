"""Key metrics of an ncu report as `name [unit] = value` lines (what the profiles/ summaries hold):
python tools/ncu_summary.py report.ncu-rep [extra-metric-substring ...]"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum\t",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum ",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__m_l1tex2xbar_write_bytes.sum ", "launch__cluster_size",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__registers_per_thread ", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max ", "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.max.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.max.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "smsp__inst_executed_op_shared", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    for h, u, v in zip(hdr, units, r):
        hh = h + " "
        if any(hh.startswith(k) or h == k.strip() for k in KEYS) or any(e in h for e in extra):
            print(f"{h} [{u}] = {v}")
    stalls = []
    for h, u, v in zip(hdr, units, r):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try:
                stalls.append((float(v.replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    print("warp stall reasons (warps per issue-active cycle): " + ", ".join(f"{n} {x:.2f}" for x, n in sorted(stalls, reverse=True)[:8]))
    print()
