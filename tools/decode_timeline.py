"""Decode the SDES_FL_TIMELINE output of the fused gradient kernel (two lines 'TIMELINE control: ...' / 'TIMELINE epilogue: ...'):
per item of CTA 0, the control warp's (before wait, after wait) stamps of its 8 waits for the epilogue and epilogue warp 2's
stamps (arrive X | per hop: before wait, after wait, arrive), in SM cycles relative to the item's start.
usage: python tools/decode_timeline.py file [first_item] [n_items]"""
import sys
lines = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith("TIMELINE")]
c = [int(x) for x in lines[-2].split()[2:]]
e = [int(x) for x in lines[-1].split()[2:]]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
for it in range(first, first + n):
    seg = c[it * 16:(it + 1) * 16]
    base = seg[0]
    print(f"item {it}: length {c[(it + 1) * 16] - base} cycles")
    print("  control  (wait from, to):", [(seg[2 * i] - base, seg[2 * i + 1] - base) for i in range(8)])
    print("  epilogue:", [x - base for x in e[it * 22:(it + 1) * 22]])
