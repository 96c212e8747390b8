"""Summarise ncu captures into committed text files under profiles/.
    python tools/summarize_ncu.py <tag>      # reads gpurun_out/prof_softras.ncu-rep + launches.csv
"""
import csv
import io
import os
import re
import subprocess
import sys

KEEP = r"^(gpu__time_duration\.sum|launch__(registers_per_thread|grid_size|block_size|occupancy_limit_\w+|waves_per_multiprocessor)|" \
       r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__issue_active\.avg\.pct_of_peak_sustained_active|" \
       r"smsp__thread_inst_executed_per_inst_executed\.ratio|inst_executed|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|" \
       r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|dram__bytes_(read|write)\.sum|lts__t_bytes\.sum|" \
       r"sm__inst_executed_pipe_(alu|fma|fp64|lsu|xu|adu|cbu)\.avg\.pct_of_peak_sustained_active|" \
       r"sm__cycles_active\.avg|sm__cycles_elapsed\.max|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|" \
       r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|" \
       r"smsp__warps_eligible\.avg\.per_cycle_active)$"


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs("profiles", exist_ok=True)
    rep = "gpurun_out/prof_softras.ncu-rep"
    title = "bench.py c3 workload, 4 images 1024^2, 39 200 faces"
    name_out = "ncu_full_softras"
    if len(sys.argv) > 3:      # python tools/summarize_ncu.py <tag> <report.ncu-rep> <name> ["title"]
        rep, name_out = sys.argv[2], sys.argv[3]
        title = sys.argv[4] if len(sys.argv) > 4 else rep
    if os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        ki = hdr.index("Kernel Name")
        with open("profiles/%s_%s.md" % (tag, name_out), "w") as f:
            f.write("# ncu --set full --clock-control none (%s)\n\n" % title)
            f.write("Per-launch values under the profiler (serialised, cold caches): use for SHARES and pipe/stall structure, not for bench numbers.\n\n")
            for r in rows[2:]:
                name = re.sub(r"\(.*", "", r[ki])
                f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % name)
                for i, h in enumerate(hdr):
                    if re.search(KEEP, h):
                        f.write("| %s | %s | %s |\n" % (h, r[i], units[i]))
                f.write("\n")
        print("wrote profiles/%s_%s.md" % (tag, name_out))
        if len(sys.argv) > 3:
            return
    lc = "gpurun_out/launches.csv"
    if os.path.exists(lc):
        lines = [l for l in open(lc) if l.startswith('"')]
        rows = list(csv.reader(lines))
        hdr = rows[0]
        ni, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
        agg = {}
        for r in rows[1:]:
            if r[mi] != "gpu__time_duration.sum":
                continue
            n = re.sub(r"\(.*", "", r[ni])
            n = re.sub(r"void ", "", n)
            t = float(r[vi].replace(",", ""))
            a = agg.setdefault(n, [0, 0.0])
            a[0] += 1
            a[1] += t
        tot = sum(a[1] for a in agg.values())
        unit = rows[1][hdr.index("Metric Unit")]
        with open("profiles/%s_launches.md" % tag, "w") as f:
            f.write("# ncu launch list: `ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 1`\n\n")
            f.write("All kernels of the process (warm-up, timed, profiler and e2e steps), time unit %s; shares, not absolutes.\n\n" % unit)
            f.write("| kernel | launches | total | share |\n|---|---|---|---|\n")
            for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
                f.write("| %s | %d | %.1f | %.1f%% |\n" % (n[:110], a[0], a[1], 100 * a[1] / tot))
        print("wrote profiles/%s_launches.md" % tag)


if __name__ == "__main__":
    main()
