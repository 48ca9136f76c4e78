"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python profiles/summarize.py launches <launches.csv> <out.txt> "<command line that produced it>"
  python profiles/summarize.py full <report.ncu-rep> <out.txt> "<command line>"
  python profiles/summarize.py source <report.ncu-rep> <out.txt> "<command line>" [top_n]
      hottest CUDA source lines of every kernel in a report captured with --import-source on (needs -lineinfo):
      instructions executed, warp-stall samples and lanes per instruction, per line
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


def source(rep, out, cmd, top=25):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    kernels = collections.OrderedDict()          # kernel -> {(file, line): [source, inst, thread_inst, samples]}
    fname = func = None
    col = {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1]
        elif len(r) == 2 and r[0] == "Function Name":
            func = r[1]
        elif r and r[0] == "Line No":
            col = {n: i for i, n in enumerate(r)}
        elif col and len(r) > 8 and r[0].strip().isdigit():
            def num(name):
                try:
                    return float(r[col[name]].replace(",", ""))
                except (KeyError, ValueError, IndexError):
                    return 0.0
            d = kernels.setdefault(func, collections.OrderedDict())
            key = (fname, int(r[0]))
            e = d.setdefault(key, [r[1].strip(), 0.0, 0.0, 0.0])
            e[1] += num("Instructions Executed"); e[2] += num("Thread Instructions Executed"); e[3] += num("Warp Stall Sampling (All Samples)")
    with open(out, "w") as f:
        f.write("%s\nhottest CUDA source lines (ncu --page source --print-source cuda,sass); inst = warp instructions executed, lanes = thread instructions / inst\n" % cmd)
        for k, d in kernels.items():
            ti = sum(e[1] for e in d.values()) or 1.0
            ts = sum(e[3] for e in d.values()) or 1.0
            f.write("\n== %s: %.3g warp instructions, %d stall samples\n" % (short(k), ti, ts))
            f.write("%6s %6s %5s  %-34s %s\n" % ("inst%", "stall%", "lanes", "file:line", "source"))
            for (fn, ln), e in sorted(d.items(), key=lambda kv: -kv[1][1])[:top]:
                f.write("%6.2f %6.2f %5.1f  %-34s %s\n" % (100 * e[1] / ti, 100 * e[3] / ts, e[2] / e[1] if e[1] else 0.0,
                                                       "%s:%d" % (fn.split("/")[-1], ln), e[0][:110]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "source":
        source(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else 25)
        sys.exit(0)
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
