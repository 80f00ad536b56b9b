"""Summarise `ncu -i X.ncu-rep --page source --csv` (SASS view): dynamic instruction mix, stall reasons,
and the hottest straight-line regions (split at branch targets / backward branches) with their
executed-instruction share.  Usage: python tools/ncu_source_summary.py src.csv [n_regions]"""
import csv
import re
import sys
from collections import Counter


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    ex = col["Instructions Executed"]
    smp = col["# Samples"]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_i = sum(int(r[ex] or 0) for r in data)
    tot_s = sum(int(r[smp] or 0) for r in data)
    print(f"instructions executed (warp-level): {tot_i:.4g}   samples: {tot_s}")
    st = Counter()
    for r in data:
        for h in stall_cols:
            st[h] += int(r[col[h]] or 0)
    print("stalls: " + "  ".join(f"{k[6:]} {100*v/max(1,sum(st.values())):.1f}%" for k, v in st.most_common(10)))
    ops = Counter()
    ops_s = Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[col["Source"]])
        op = m.group(2) if m else "?"
        ops[op] += int(r[ex] or 0)
        ops_s[op] += int(r[smp] or 0)
    print("opcodes: " + "  ".join(f"{k} {100*v/tot_i:.1f}%/{100*ops_s[k]/tot_s:.1f}%" for k, v in ops.most_common(28)))
    # regions: consecutive instructions with the same execution count (within 2%) are one block
    regions = []
    cur = None
    for idx, r in enumerate(data):
        n = int(r[ex] or 0)
        if cur is None or abs(n - cur["n"]) > 0.02 * max(n, cur["n"], 1):
            cur = {"start": idx, "n": n, "len": 0, "inst": 0, "smp": 0}
            regions.append(cur)
        cur["len"] += 1
        cur["inst"] += n
        cur["smp"] += int(r[smp] or 0)
    regions.sort(key=lambda g: -g["inst"])
    print(f"top {top} regions by executed instructions:")
    for g in regions[:top]:
        seg = data[g["start"]:g["start"] + g["len"]]
        mix = Counter()
        for r in seg:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[col["Source"]])
            mix[m.group(2) if m else "?"] += 1
        mixs = " ".join(f"{k}:{v}" for k, v in mix.most_common(8))
        print(f"  sass[{g['start']:5d}+{g['len']:4d}] exec/inst {g['n']:.3g}  inst {100*g['inst']/tot_i:5.1f}%  samples {100*g['smp']/tot_s:5.1f}%  {mixs}")


if __name__ == "__main__":
    main()
