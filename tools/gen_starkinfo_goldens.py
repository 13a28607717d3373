#!/usr/bin/env python3
"""Goldens for the PIL codegen (eigen_zkvm_b200/starkinfo.py, the port of starky/src/starkinfo*.rs): the emitted step
programs and the StarkInfo summary of the six reference fixtures, rendered as text, one op per line.

The reference ships no starkinfo goldens; these pin the codegen independently of the prover/oracle pair that consumes it
(both read the same programs, so a mis-port would otherwise be invisible to every byte-identical proof test).
tests/test_starkinfo_goldens.py compares a fresh run with these files and checks the Fibonacci programs against the
listing in SURVEY.md Appendix C, which was derived by an independent throw-away port.

    python tools/gen_starkinfo_goldens.py            # rewrites tests/golden/starkinfo/*.txt
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G = os.path.join(ROOT, "tests", "golden")
FIXTURES = [  # name, pil, starkStruct
    ("fib.gl", "fib.pil.json.gl", "starkStruct.json.gl"),
    ("plookup.gl", "plookup.pil.json.gl", "starkStruct.json.gl"),
    ("fib.bn128", "fib.pil.json", "starkStruct.json"),
    ("plookup.bn128", "plookup.pil.json", "starkStruct.json"),
    ("pe.bn128", "pe.pil.json", "starkStruct.json"),
    ("connection.bn128", "connection.pil.json", "starkStruct.json"),
]
PROGRAMS = ["publics_code", "step2prev", "step3prev", "step3", "step42ns", "step52ns", "verifier_code", "verifier_query_code"]
SYM = {"add": "+", "sub": "-", "mul": "*"}


def ref(r):
    t = r["type_"]
    p = "'" if r.get("prime") else ""
    if t == "number": return "number(%s)" % r["value"]
    if t in ("x", "Zi", "Z", "xDivXSubXi", "xDivXSubWXi"): return t
    s = "%s%d%s" % (t, r["id"], p)
    if r.get("dim", 1) != 1: s += ":%d" % r["dim"]
    return s


def render_segment(code):
    out = []
    for op in code:
        d, s = ref(op["dest"]), [ref(x) for x in op["src"]]
        if op["op"] == "copy": out.append("%s = %s" % (d, s[0]))
        elif op["op"] == "muladd": out.append("%s = %s * %s + %s" % (d, s[0], s[1], s[2]))
        else: out.append("%s = %s %s %s" % (d, s[0], SYM[op["op"]], s[1]))
    return out


def render(pil_name, ss_name):
    from eigen_zkvm_b200 import starkinfo as si
    pil = si.load_pil(os.path.join(G, pil_name))
    ss = json.load(open(os.path.join(G, ss_name)))
    info, prog = si.new_starkinfo(pil, ss)
    lines = ["# %s + %s" % (pil_name, ss_name)]
    lines.append("n_cm1=%d n_cm2=%d n_cm3=%d n_cm4=%d n_q=%d q_deg=%d q_dim=%d n_constants=%d n_publics=%d" % (
        info.n_cm1, info.n_cm2, info.n_cm3, info.n_cm4, info.n_q, info.q_deg, info.q_dim, info.n_constants, info.n_publics))
    lines.append("sections " + " ".join("%s=%d" % (k, v) for k, v in sorted(info.map_sectionsN.items())))
    lines.append("ev_map " + " ".join("%s%d%s" % (e["type_"], e["id"], "'" if e["prime"] else "") for e in info.ev_map))
    for name in PROGRAMS:
        if name == "publics_code":          # one segment per public (empty for publics that are plain trace cells)
            for i, pc in enumerate(prog[name]):
                seg = pc["first"] if isinstance(pc, dict) else pc
                lines.append("[publics_code %d] %d ops" % (i, len(seg)))
                lines += render_segment(seg)
            continue
        seg = prog[name]["first"]
        lines.append("[%s] %d ops" % (name, len(seg)))
        lines += render_segment(seg)
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    out = os.path.join(G, "starkinfo")
    os.makedirs(out, exist_ok=True)
    for name, pil, ss in FIXTURES:
        open(os.path.join(out, name + ".txt"), "w").write(render(pil, ss))
        print("wrote", name)
