#!/usr/bin/env python3
"""Opcode mix of a kernel from an `ncu --page source --csv` dump: thread-instructions per opcode, first kernel (or
the kernel whose name contains argv[2]).  Usage: tools/ncu_opmix.py prof_source.csv [name-substring] [units]
`units` (e.g. number of elements / permutations processed) prints per-unit counts."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
units = float(sys.argv[3]) if len(sys.argv) > 3 else None
i = 0; done = False
while i < len(rows) and not done:
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2
        ops = collections.Counter(); stall = collections.Counter(); tot = 0
        ci = hdr.index("Thread Instructions Executed"); si = hdr.index("Source")
        ss = hdr.index("# Samples")
        while j < len(rows) and rows[j] and rows[j][0] != "Kernel Name":
            r = rows[j]
            op = r[si].strip().split()
            if op and op[0].startswith("@"): op = op[1:]
            o = op[0].rstrip(";") if op else "?"
            n = int(r[ci] or 0); ops[o] += n; tot += n; stall[o] += int(r[ss] or 0)
            j += 1
        if want in name:
            print(name[:120]); print("thread-instructions: %.4g" % tot + (" = %.1f per unit" % (tot / units) if units else ""))
            ts = sum(stall.values())
            for o, n in ops.most_common(28):
                print("  %-22s %6.2f %%  %s   samples %5.1f %%" % (o, 100.0 * n / tot, ("%9.1f/unit" % (n / units)) if units else "", 100.0 * stall[o] / max(ts, 1)))
            done = True
        i = j
    else:
        i += 1
