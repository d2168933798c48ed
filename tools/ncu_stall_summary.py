"""Summarise `ncu -i X.ncu-rep --page source --csv --kernel-name regex:<k>` into per-stall-reason shares and the
hottest instructions (warp-state sampling).  Usage: python tools/ncu_stall_summary.py <rep> <kernel regex> [top]"""
import csv, io, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r][0]
print(rows[0][1] if len(rows[0]) > 1 else kern)
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 2 and r[idx["# Samples"]].isdigit()]
# the source page lists every instruction once per view (SASS / source-correlated): keep the first copy
first = data[0][idx["Address"]]
dup = [i for i, r in enumerate(data) if r[idx["Address"]] == first]
if len(dup) > 1:
    data = data[:dup[1]]
tot = sum(int(r[idx["# Samples"]]) for r in data)
print("instructions: %d   warp-state samples: %d" % (len(data), tot))
agg = {c: 0 for c in stall}
for r in data:
    for c in stall:
        try:
            agg[c] += int(r[idx[c]] or 0)
        except ValueError:
            pass
print("\nstall reasons (share of samples):")
for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print("  %-24s %6.1f %%" % (c, 100.0 * v / tot))
print("\nhottest instructions:")
for i, r in sorted(sorted(enumerate(data), key=lambda ir: -int(ir[1][idx["# Samples"]]))[:top_n]):
    st = {}
    for c in stall:
        try:
            st[c] = int(r[idx[c]] or 0)
        except ValueError:
            st[c] = 0
    main = max(st.items(), key=lambda kv: kv[1])
    print("  #%-5d %5.1f %%  %-22s %s" % (i, 100.0 * int(r[idx["# Samples"]]) / tot, main[0], r[idx["Source"]].strip()[:80]))
