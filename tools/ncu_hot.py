"""hot SASS instructions of one kernel from an ncu report: python tools/ncu_hot.py file.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = []
for idx, r in enumerate(rows[2:]):
    if len(r) < len(hdr):
        continue
    data.append((idx, int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]]), float(r[ci["Avg. Threads Executed"]] or 0), r))
tot_i = sum(d[1] for d in data)
tot_s = sum(d[2] for d in data)
print("warp instructions %d, samples %d, SASS lines %d" % (tot_i, tot_s, len(data)))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(d[4][ci[h]] or 0) for d in data) for h in stalls}
print("stall samples:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot_s * 0.01})
print("-- by samples")
for idx, n, s, thr, r in sorted(data, key=lambda d: -d[2])[:top]:
    st = {h[6:]: int(r[ci[h]]) for h in stalls if int(r[ci[h]] or 0) > s * 0.2}
    print("%5d %5.1f%%smp %5.1f%%ins thr%5.1f  %-60s %s" % (idx, 100 * s / tot_s, 100 * n / tot_i, thr, r[ci["Source"]].strip()[:60], st))
