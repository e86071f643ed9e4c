"""Summarise `ncu -i X.ncu-rep --page source --csv` output: stall samples per reason, per code region
(split at BAR.SYNC) and the hottest instructions.  Usage: python tools/ncu_source_summary.py file.csv [kernel-index]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
# the csv holds one block per kernel launch: "Kernel Name" line, header line, instruction lines
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
b = blocks[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
hdr, data = b["hdr"], b["data"]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
num = lambda r, k: int(r[ci[k]] or 0)
tot = collections.Counter()
for r in data:
    for s in stalls:
        tot[s] += num(r, s)
allsum = sum(tot.values()) or 1
print(b["name"][:80]); print("total samples", allsum)
for s, v in tot.most_common(10):
    print(f"  {s:28s} {v:7d} {100*v/allsum:5.1f}%")
bounds = [0] + [i + 1 for i, r in enumerate(data) if "BAR.SYNC" in r[ci["Source"]]] + [len(data)]
for a, e in zip(bounds[:-1], bounds[1:]):
    seg = data[a:e]
    samples = sum(num(r, "# Samples") for r in seg)
    instr = sum(num(r, "Instructions Executed") for r in seg)
    st = collections.Counter()
    for r in seg:
        for s in stalls:
            st[s] += num(r, s)
    print(f"region instr[{a}:{e}]: samples {samples} ({100*samples/allsum:.1f}%), warp-instr {instr/1e6:.2f}M; "
          f"top: {[(k.replace('stall_', ''), v) for k, v in st.most_common(5)]}")
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:16]:
    print(num(r, "# Samples"), r[ci["Source"]].strip()[:80])
