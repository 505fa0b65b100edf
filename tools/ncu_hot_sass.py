"""Top stalled SASS instructions of one kernel from an .ncu-rep (source page): where the warps wait and why.
usage: python tools/ncu_hot_sass.py <rep> <kernel-regex> [top-n]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# first kernel instance only
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
body = []
for r in rows[start + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ci["# Samples"]]) for r in body)
print("instructions %d, samples %d" % (len(body), tot))
agg = {s: sum(int(r[ci[s]]) for r in body) for s in stalls}
print("by reason:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ci["# Samples"]]))[:topn]
for i in sorted(order):
    r = body[i]
    why = sorted(((int(r[ci[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print("%5d %6.2f%% exec %8s  %-60s %s" % (i, 100.0 * int(r[ci["# Samples"]]) / tot, r[ci["Instructions Executed"]], r[ci["Source"]].strip()[:60],
                                        " ".join("%s:%d" % (n, c) for c, n in why if c)))
