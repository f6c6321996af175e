"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of numbers the roofline uses.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/rN_xxx.txt]
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu_dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_lsu_wavefronts_pct"),
    ("sm__cycles_elapsed.max", "sm_cycles"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb_per_issue"),
    ("smsp__inst_executed.sum", "inst"),
    ("sm__inst_executed_pipe_fp64.sum", "fp64_inst"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for row in data:
        name = row[col["Kernel Name"]]
        print("kernel: %s  grid=%s block=%s" % (name, row[col["Grid Size"]], row[col["Block Size"]]))
        vals = {}
        for metric, short in WANT:
            if metric in col:
                vals[short] = (row[col[metric]], units[col[metric]])
                print("    %-28s %s %s" % (short, row[col[metric]], units[col[metric]]))
        try:
            def tobytes(v, u):
                v = float(v.replace(",", ""))
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            def tosec(v, u):
                v = float(v.replace(",", ""))
                return v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[u]
            traffic = tobytes(*vals["dram_read"]) + tobytes(*vals["dram_write"])
            dur = tosec(*vals["duration"])
            print("    %-28s %.0f bytes  (%.1f GB/s over the profiled duration)" % ("dram_traffic", traffic, traffic / dur / 1e9))
        except Exception as e:  # pragma: no cover
            print("    (traffic not derivable: %s)" % e)


if __name__ == "__main__":
    main(sys.argv[1])
