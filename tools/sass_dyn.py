#!/usr/bin/env python3
"""Dynamic opcode mix of a kernel from its static SASS: instruction counts per address range times the trip counts given
on the command line, with the FMA-heavy / ALU pipe cycles of tools/ubench/int_pipes2.cu (IMAD.WIDE and IMAD.HI 4 clk per
warp instruction per SMSP, other IMAD forms 2, ALU-pipe instructions 2).
  python tools/sass_dyn.py <obj-or-cubin> <function-substring> [lo-hi:mult ...]     (hex addresses, inclusive)
Without ranges: prints the branch structure so that the ranges can be chosen."""
import re, subprocess, sys, collections

def load(obj, fn):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for part in re.split(r"Function : ", txt)[1:]:
        if fn in part.split()[0]:
            return re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", part)
    raise SystemExit("function not found")

HEAVY4 = ("IMAD.WIDE", "IMAD.HI")
def cls(op):
    if op.startswith(HEAVY4): return "heavy4"
    if op.startswith("IMAD"): return "heavy2"
    if op.startswith(("IADD3", "LOP3", "SHF", "SEL", "MOV", "LEA", "ISETP", "PRMT", "VIADD", "IABS", "ICMP", "SHL", "SHR")): return "alu"
    if op.startswith(("LDG", "STG", "LDL", "STL", "LDS", "STS", "LDC", "LDCU")): return "mem"
    return "other"

if __name__ == "__main__":
    ins = load(sys.argv[1], sys.argv[2])
    if len(sys.argv) == 3:
        print(len(ins), "instructions, last address", ins[-1][0])
        for a, t in ins:
            if re.search(r"\b(BRA|CALL|RET|EXIT)", t): print(a, t[:90])
        sys.exit(0)
    tot = collections.Counter()
    for spec in sys.argv[3:]:
        rng, mult = spec.split(":"); lo, hi = [int(x, 16) for x in rng.split("-")]; mult = float(mult)
        for a, t in ins:
            if lo <= int(a, 16) <= hi:
                op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
                tot[op] += mult
    n = sum(tot.values())
    pipes = collections.Counter()
    for op, c in tot.items(): pipes[cls(op)] += c
    for op, c in tot.most_common(24): print("%-22s %9.0f  %5.1f%%" % (op, c, 100 * c / n))
    print("total thread-instructions %.0f" % n)
    print("classes", {k: round(v) for k, v in pipes.items()})
    print("FMA-heavy clk %.0f   ALU clk %.0f   issue clk %.0f" % (4 * pipes["heavy4"] + 2 * pipes["heavy2"], 2 * pipes["alu"], n))
