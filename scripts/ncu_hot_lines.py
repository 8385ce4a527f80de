"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda` (stdin or file): stall reasons and
the hottest CUDA source lines by warp-stall samples.  Usage: python scripts/ncu_hot_lines.py page.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
lines, stall = [], collections.Counter()
for r in rows:
    if len(r) > 3 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0] not in ("", "-") and r[2] == "-":          # a CUDA source line: aggregated over its SASS
        d = dict(zip(hdr[4:], r[4:]))
        try:
            s = int(d.get("# Samples", "0"))
        except ValueError:
            continue
        lines.append((s, int(r[0]), r[1].strip(), d))
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k:
                try:
                    stall[k] += int(v)
                except ValueError:
                    pass
tot = sum(s for s, *_ in lines)
print(f"# warp-stall samples over CUDA lines: {tot}")
print("== stall reasons (all samples)")
ts = sum(stall.values())
for k, v in stall.most_common(9):
    print(f"{k:28s}{v:9d} {100.0 * v / ts:5.1f}%")
print("== hottest source lines (samples, share, line, avg threads/inst, source)")
for s, ln, src, d in sorted(lines, key=lambda t: -t[0])[:top]:
    try:
        thr = float(d["Thread Instructions Executed"]) / max(1.0, float(d["Instructions Executed"]))
    except Exception:
        thr = float("nan")
    print(f"{s:8d} {100.0 * s / tot:5.1f}%  :{ln:<5d} thr/inst={thr:5.1f}  {src[:110]}")
