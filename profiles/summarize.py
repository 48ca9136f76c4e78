"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python profiles/summarize.py launches <launches.csv> <out.txt> "<command line that produced it>"
  python profiles/summarize.py full <report.ncu-rep> <out.txt> "<command line>"
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
           "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def short(name):
    m = re.search(r"(k_[a-z_0-9]+)", name)
    return "gzpb::" + m.group(1) if m else "torch (setup fills/copies)"


def launches(path, out, cmd):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
        k = short(r[ki])
        tot[k] = tot.get(k, 0.0) + v
        cnt[k] += 1
    ours = sum(v for k, v in tot.items() if k.startswith("gzpb"))
    with open(out, "w") as f:
        f.write("ncu launch list: %s (serialised, cold cache; compare SHARES)\n" % cmd)
        f.write("%-28s %8s %12s %8s\n" % ("kernel", "launches", "total_ms", "share"))
        for k, v in tot.items():
            f.write("%-28s %8d %12.3f %8.3f\n" % (k, cnt[k], v, v / ours if k.startswith("gzpb") else 0.0))


def full(path, out, cmd):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("%s\n" % cmd)
        f.write("units: %s\n" % {m: units[hdr.index(m)] for m in METRICS if m in hdr})
        for r in rows[2:]:
            d = {"Kernel Name": short(r[hdr.index("Kernel Name")])}
            for m in METRICS:
                if m in hdr:
                    d[m] = r[hdr.index(m)]
            f.write("%s\n" % d)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
