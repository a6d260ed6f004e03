"""Turn an .ncu-rep (ncu --set full) into a short text summary for profiles/ (run here, no GPU needed)."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("kernel: %s" % name)
        for k in KEYS:
            if k in hdr:
                print("  %-72s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "_not_issued" not in h:
                try:
                    stalls[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        tot = sum(stalls.values()) or 1.0
        print("  warp-state samples (share): " + ", ".join(
            "%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v / tot > 0.01))
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
            wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
            ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            print("  dram traffic per launch: %.3f MB" % ((rd * scale[ur] + wr * scale[uw]) / 1e6))
        except Exception:
            pass
        print()


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("==== %s" % p)
        main(p)
