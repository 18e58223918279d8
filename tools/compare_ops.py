"""Compare two tools/profile_ops.py tables signature by signature.  usage: python tools/compare_ops.py A.txt B.txt [family]"""
import re
import sys


def load(path):
    out = {}
    for line in open(path):
        m = re.match(r"\s*([\d.]+)\s+[\d.]+%\s+(\d+)\s+([\d.]+)\s+([\d.]+)\s+(\(.*?\))\s*(.*)", line)
        if m:
            out[m.group(5)] = (float(m.group(1)), int(m.group(2)), float(m.group(3)), m.group(6))
    return out


a, b = load(sys.argv[1]), load(sys.argv[2])
fam = sys.argv[3] if len(sys.argv) > 3 else ""
ta = tb = tbest = 0.0
rows = []
for sig, (tot, n, us, extra) in a.items():
    if fam and fam not in sig:
        continue
    if sig in b:
        rows.append((tot, n, us, b[sig][2], sig, extra, b[sig][3]))
        ta += tot
        tb += b[sig][0]
        tbest += n * min(us, b[sig][2])
for tot, n, us, us2, sig, e1, e2 in sorted(rows, reverse=True):
    print(f"{n:3d} x {us:8.2f} -> {us2:8.2f}  ({us2 / us:5.2f})  {sig}  {e1} | {e2}")
print(f"total A {ta:.1f} us, B {tb:.1f} us, best-of-both {tbest:.1f} us")
