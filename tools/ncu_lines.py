"""Development tool: per-source-line instruction and stall-sample totals from
`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.  Usage: python tools/ncu_lines.py file.csv [top]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
fil = None; rows = []
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": fil = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No"): continue
    if r[0] and r[0].isdigit() and len(r) > 8 and r[7].isdigit():
        rows.append((fil, int(r[0]), r[1].strip()[:100], int(r[6]) if r[6].isdigit() else 0, int(r[7]), int(r[8])))
ti = sum(x[4] for x in rows); ts = sum(x[3] for x in rows)
print("total warp-instructions", ti, "samples", ts)
byfile = collections.Counter(); sfile = collections.Counter()
for f, l, s, sm, ins, th in rows: byfile[f] += ins; sfile[f] += sm
for f, n in byfile.most_common(): print("  %-28s %6.2f %% inst %6.2f %% samples" % (f, 100.0 * n / ti, 100.0 * sfile[f] / ts))
for f, l, s, sm, ins, th in sorted(rows, key=lambda x: -x[4])[:top]:
    print("%5.2f%% i %5.2f%% s  lanes %4.1f  %s:%d  %s" % (100.0 * ins / ti, 100.0 * sm / ts, th / max(ins, 1), f, l, s))
