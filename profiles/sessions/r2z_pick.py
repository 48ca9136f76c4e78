"""Session r2z: choose the kernel-variant combination from the A/B lines (gpurun_out/r2z_kernels.jsonl).
Prints `<lib> <inflight>`; every candidate must reproduce the reference build's packed stream (sha1) with status ok."""
import json
import sys

rows = {}
for ln in open(sys.argv[1]):
    ln = ln.strip()
    if ln.startswith("{"):
        d = json.loads(ln)
        rows.setdefault(d["label"], []).append(d)


def best(label):
    r = rows.get(label)
    return min(r, key=lambda d: d["ms_best"]) if r else None


head = best("head@3256")
if not head or not head["status_ok"]:
    print("head 3256"); sys.exit(0)
ok3256 = lambda d: d and d["status_ok"] and d["sha1"] == head["sha1"]
# link variant: smallest split+link time
link = [(best(l + "0m1@3256")["kernel_ms"]["chain"], l) for l in "ABD" if ok3256(best(l + "0m1@3256"))]
if not link:
    print("head 3256"); sys.exit(0)
L = min(link)[1]
# match variant: pipelined loads only if they are faster than the reference build's k_match
M = "m1" if best("A0m1@3256") and ok3256(best("A0m1@3256")) and best("A0m1@3256")["kernel_ms"]["match"] < 0.995 * head["kernel_ms"]["match"] else "m0"
# emit variant and batch size: highest whole-pipeline throughput among the valid candidates
cands = []
a0 = best("A0m1@3256")
if ok3256(a0):
    cands.append((a0["GiB/s"], "0", 3256))
a0b = best("A0m1@4736")
a2b, a2 = best("A2m1@4736"), best("A2m1@3256")
if a0b and a0b["status_ok"]:
    cands.append((a0b["GiB/s"], "0", 4736))
    if a2b and a2b["status_ok"] and a2b["sha1"] == a0b["sha1"] and ok3256(a2):
        cands.append((a2b["GiB/s"], "2", 4736))
        cands.append((a2["GiB/s"], "2", 3256))
g, E, N = max(cands)
print("%s%s%s %d" % (L, E, M, N))
