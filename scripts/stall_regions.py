"""Per-region warp-stall sampling of a kernel from an ncu --set full --import-source on capture.

    ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv
    python scripts/stall_regions.py src.csv [kernel substring]

Splits the kernel's SASS at marker instructions (barriers, the MUFU.EX2 chains of the tap staging, the
lock CAS / EXCH of the add-out, the FFMA2 blocks of the point slots, global reductions / stores) and
prints, per region, the share of stall samples, of executed warp instructions, and the top stall reasons.
"""
import csv
import sys
from collections import Counter, OrderedDict

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else "spread_reg_kernel"
rows = list(csv.reader(open(path)))
# the file holds one table per kernel: "Kernel Name", name / header / instructions ...
tables, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "ins": []}
        tables.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["ins"].append(r)
tab = [t for t in tables if want in t["name"]][0]
h = {k: i for i, k in enumerate(tab["hdr"])}
stall_cols = [k for k in tab["hdr"] if k.startswith("stall_") and "Not Issued" not in k]


def region_of(idx, text, state):
    op = text.split()[0] if not text.strip().startswith("@") else text.split()[1]
    if "MUFU.EX2" in text:
        state["r"] = "stage (tap evaluation)"
    elif "ATOMS.CAS" in text:
        state["r"] = "add-out (lock .. unlock)"
    elif "ATOMS.EXCH" in text:
        state["next"] = "slide / loop control"
    elif op.startswith("FFMA2"):
        state["r"] = "point slots (window loads + FFMA2)"
    elif op.startswith("RED") or op.startswith("ATOMG"):
        state["r"] = "flush (global reductions)"
    elif op.startswith("LDGSTS"):
        state["r"] = "tile load (cp.async)"
    elif op.startswith("STG"):
        state["r"] = "store y"
    elif op.startswith("BAR"):
        state["r"] = "after barrier %d" % state["bar"]
        state["bar"] += 1
    r = state["r"]
    if "next" in state:
        state["r"] = state.pop("next")
    return r


# two passes: a region label applies from its first marker BACKWARDS to the previous marker's end is not
# knowable, so label forwards and additionally pull the window loads that precede an FFMA2 block into it
state = {"r": "prologue", "bar": 0}
labels = []
for i, r in enumerate(tab["ins"]):
    labels.append(region_of(i, r[h["Source"]], state))
for i in range(len(labels) - 1, 0, -1):  # LDS / FMUL directly before an FFMA2 block belong to the slot
    if labels[i].startswith("point slots") and not labels[i - 1].startswith("point slots"):
        j = i - 1
        while j >= 0 and tab["ins"][j][h["Source"]].split()[0].split(".")[0] in ("LDS", "FMUL", "ISETP", "BRA", "BRX", "LDC", "IMAD", "SHF", "HFMA2", "VIADDMNMX", "BSYNC", "BREAK") and i - j < 40:
            labels[j] = labels[i]
            j -= 1

agg = OrderedDict()
for lab, r in zip(labels, tab["ins"]):
    a = agg.setdefault(lab, {"samples": 0, "inst": 0, "n": 0, "stalls": Counter()})
    a["samples"] += int(r[h["# Samples"]] or 0)
    a["inst"] += int(r[h["Instructions Executed"]] or 0)
    a["n"] += 1
    for k in stall_cols:
        v = int(r[h[k]] or 0)
        if v:
            a["stalls"][k[6:]] += v
ts = sum(a["samples"] for a in agg.values()) or 1
ti = sum(a["inst"] for a in agg.values()) or 1
print(f"# {tab['name'][:90]}\n# {len(tab['ins'])} SASS instructions, {ts} stall samples, {ti} warp instructions executed")
print(f"{'region':42s} {'sass':>5s} {'samples':>8s} {'inst':>7s}  top stall reasons")
for lab, a in agg.items():
    top = ", ".join(f"{k}={100 * v / max(a['samples'], 1):.0f}%" for k, v in a["stalls"].most_common(4))
    print(f"{lab:42s} {a['n']:5d} {100 * a['samples'] / ts:7.1f}% {100 * a['inst'] / ti:6.1f}%  {top}")
