"""Turns an ncu report (.ncu-rep) into a small text summary that can be committed under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.md"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of ncu peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "blocks/SM (register limit)"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print("# ncu summary of `%s`\n" % path.split("/")[-1])
    print("Captured with `ncu --set full --clock-control none --import-source on` (see the command in profiles/README.md).")
    print("Per-launch values; durations under ncu are cold-cache and serialised.\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("## %s\n" % name)
        print("| metric | value |")
        print("|---|---|")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print("| %s (`%s`) | %s %s |" % (label, key, r[i], units[i]))
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
            ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tot = rd * scale[ur] + wr * scale[uw]
            dur = float(r[hdr.index("gpu__time_duration.sum")]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[units[hdr.index("gpu__time_duration.sum")]]
            print("| **DRAM traffic per launch** | %.3f GB |" % (tot / 1e9))
            print("| **DRAM bandwidth (traffic / duration)** | %.0f GB/s |" % (tot / dur / 1e9))
        except Exception:
            pass
        print()


if __name__ == "__main__":
    main(sys.argv[1])
