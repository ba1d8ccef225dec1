"""Summarise one kernel of an .ncu-rep as a markdown table (+ the DRAM traffic JSON bench.py reads).
usage: python tools/ncu_summary.py gpurun_out/prof_hvp.ncu-rep profiles/r01_hvp_ncu_full_summary.md [profiles/r01_hvp_traffic.json]"""
import csv, io, json, subprocess, sys

rep, out_md = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
WANT = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sass__inst_executed_local_loads",
    "sass__inst_executed_local_stores", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
] + [f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio" for k in ("wait", "long_scoreboard", "math_pipe_throttle", "lg_throttle", "not_selected", "dispatch_stall", "no_instruction", "short_scoreboard", "branch_resolving", "drain")]
get = lambda k: (units[hdr.index(k)], vals[hdr.index(k)]) if k in hdr else ("", "n/a")  # noqa: E731
lines = ["| metric | unit | value |", "|---|---|---|"] + [f"| {k} | {get(k)[0]} | {get(k)[1]} |" for k in WANT]
open(out_md, "w").write("\n".join(lines) + "\n")
if len(sys.argv) > 3:
    def to_bytes(k):
        u, v = get(k)
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    num = lambda k: float(get(k)[1]) if get(k)[1] not in ("n/a", "") else None  # noqa: E731
    grid, block = num("launch__grid_size"), num("launch__block_size")
    json.dump({"kernel": get("Kernel Name")[1], "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"), "traffic": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
               # pipe occupancy of the same capture (bench.py copies these into roofline.ncu_same_kernel)
               "fp64_pipe_active_pct_at_2_cycles_per_instruction": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
               "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "registers_per_thread": num("launch__registers_per_thread"),
               "instructions_per_thread": num("smsp__inst_executed.sum") / (grid * block / 32) if grid and block else None,  # element-per-thread kernels: per element
               "kernel_us_under_ncu": num("gpu__time_duration.sum"),
               "source": rep}, open(sys.argv[3], "w"))
print(open(out_md).read())
