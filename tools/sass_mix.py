"""Development tool: instruction mix (by SASS mnemonic) and stall samples of an `ncu --page source --csv` export."""
import csv, sys, collections
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    names = rows[h]
    ci, cs, ct = names.index("Instructions Executed"), names.index("# Samples"), names.index("Thread Instructions Executed")
    mix, smp, thr = collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows[h + 1:]:
        if len(r) <= ct: continue
        toks = r[1].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        mix[op] += int(r[ci]); smp[op] += int(r[cs]); thr[op] += int(r[ct])
    tot, stot = sum(mix.values()), sum(smp.values())
    print(path, "warp-instructions", tot, "samples", stot, "lanes/inst %.2f" % (sum(thr.values()) / tot))
    for op, n in mix.most_common(28):
        print("  %-10s %6.2f %% of inst  %6.2f %% of samples  lanes %.1f" % (op, 100.0 * n / tot, 100.0 * smp[op] / stot, thr[op] / max(n, 1)))
