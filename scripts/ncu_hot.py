"""Hot regions of a kernel from `ncu --page source --csv` output (SASS view)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 60
hdr = next(r for r in rows if r and r[0] == "Address")
data = []
for r in rows[rows.index(hdr) + 1:]:
    if r and r[0] == "Address":
        break  # next kernel section
    if len(r) == len(hdr) and r[0].startswith("0x"):
        data.append(r)
isrc, ie, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[ie]) for r in data); tots = sum(int(r[iss]) for r in data)
print("total inst", tot, "samples", tots, "sass instructions", len(data))
for b in range(0, len(data), blk):
    seg = data[b:b + blk]
    e = sum(int(r[ie]) for r in seg); s = sum(int(r[iss]) for r in seg)
    ops = {}
    for r in seg:
        toks = r[isrc].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[ie])
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:7]
    print(f"{b:5d} inst {100*e/tot:5.1f}% stall {100*s/tots:5.1f}%  ", " ".join(f"{k}:{100*v/tot:.1f}" for k, v in top))
