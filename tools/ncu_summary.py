#!/usr/bin/env python3
"""Digest the `ncu --set full` captures made by tools/prof_round.sh (gpurun_out/prof_*_<tag>_{raw,source}.csv) into
  profiles/ncu_summary_<round>.json   per-kernel DRAM bytes and thread-instructions per algorithmic byte (read by bench.py)
  profiles/ncu_<tag>.md               key metrics + opcode mix per kernel (the evidence DESIGN.md cites)
Usage: tools/ncu_summary.py <tag> [round]      e.g. tools/ncu_summary.py r1c r1
The algorithmic bytes of each profiled launch are those of tools/prof_round.sh's fixed shapes (SURVEY.md 8d figures)."""
import csv, json, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]; rnd = sys.argv[2] if len(sys.argv) > 2 else "r1"
G = os.path.join(ROOT, "gpurun_out")
# capture -> (kernel substring, bench timing names, algorithmic bytes per launch, units per launch, unit name)
CAPS = {
    # round 2: the capture is the SECOND level of a 2^22-leaf tree (2^20 permutations, the generic zero-capacity kernel); the first level of
    # width-2 leaves runs the variant that also skips the four zero-padded lanes
    "merkle": ("k_merkle_level", ["merkle_level"], 96.0 * (1 << 20), 1 << 20, "permutation"),
    "lh": ("k_linearhash", ["linearhash_leaves"], (8.0 * 48 + 32) * (1 << 20), 10 << 20, "permutation"),
    "ntt": ("k_ntt3", ["ntt_pass", "intt_pass", "lde_ntt_pass", "lde_intt_pass"], 16.0 * (1 << 25), 1 << 25, "element-pass"),
    # round 2: tools/prof_kernels.py msm runs the TABLE mode (13 shifted copies of 2^22 bases, window 20 bits): 13 mixed additions per point
    "msm": ("k_msm_accumulate", ["msm_accumulate"], 96.0 * (1 << 22), int(os.environ.get("MSM_WINDOWS", "13")) << 22, "mixed addition"),
    "eval": ("k_eval", ["step_program"], None, None, "row"),
}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m.get(unit, 1)


def opmix(path, want):
    rows = list(csv.reader(open(path)))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1]; hdr = rows[i + 1]; j = i + 2
            ci = hdr.index("Thread Instructions Executed"); si = hdr.index("Source")
            ops = collections.Counter(); tot = 0
            while j < len(rows) and rows[j] and rows[j][0] != "Kernel Name":
                op = rows[j][si].strip().split()
                if op and op[0].startswith("@"): op = op[1:]
                n = int(rows[j][ci] or 0); ops[op[0].rstrip(";") if op else "?"] += n; tot += n; j += 1
            if want in name: return tot, ops
            i = j
        else:
            i += 1
    return 0, {}


summary = {}; md = ["# ncu captures `%s` (`ncu --set full --clock-control none --import-source on`, one B200; tools/prof_round.sh)\n" % tag]
for cap, (kname, tnames, abytes, units, uname) in CAPS.items():
    raw = os.path.join(G, "prof_%s_%s_raw.csv" % (cap, tag)); src = os.path.join(G, "prof_%s_%s_source.csv" % (cap, tag))
    if not os.path.exists(raw): continue
    rows = list(csv.reader(open(raw))); hdr, un = rows[0], rows[1]
    launches = [dict(zip(hdr, r)) for r in rows[2:] if kname in dict(zip(hdr, r)).get("Kernel Name", "")]
    if not launches: continue
    md.append("\n## %s (%d launch(es) captured)\n\n| metric | " % (kname, len(launches)) + " | ".join("launch %d" % i for i in range(len(launches))) + " |\n|---|" + "---|" * len(launches))
    for k in KEYS:
        if k in hdr: md.append("| %s [%s] | " % (k, un[hdr.index(k)]) + " | ".join(l[k] for l in launches) + " |")
    dram = sum(to_bytes(l["dram__bytes_read.sum"], un[hdr.index("dram__bytes_read.sum")]) + to_bytes(l["dram__bytes_write.sum"], un[hdr.index("dram__bytes_write.sum")]) for l in launches) / len(launches)
    winstr = sum(float(l["smsp__inst_executed.sum"]) for l in launches) / len(launches)
    tot, ops = opmix(src, kname) if os.path.exists(src) else (0, {})
    if abytes:
        tinstr = tot if tot else winstr * 32
        avg = lambda k: sum(float(l[k]) for l in launches) / len(launches) if k in hdr else None
        for t in tnames:
            summary[t] = {"kernel": kname, "dram_bytes_per_algo_byte": dram / abytes, "thread_instr_per_algo_byte": tinstr / abytes,
                          "thread_instr_per_unit": tinstr / units, "unit_name": uname, "capture": "prof_%s_%s" % (cap, tag),
                          "issue_active_pct": avg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                          "alu_pipe_pct": avg("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                          "fmaheavy_pipe_pct": avg("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed")}
        md.append("\nDRAM traffic %.1f MB for %.1f MB algorithmic (x%.2f); %.1f thread-instructions per %s." % (dram / 1e6, abytes / 1e6, dram / abytes, tinstr / units, uname))
    if tot:
        md.append("\nOpcode mix of the first captured launch (share of thread-instructions): " + ", ".join("%s %.1f%%" % (o, 100.0 * n / tot) for o, n in ops.most_common(14)))
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
if "msm_accumulate" in summary: summary["msm_accumulate"]["windows"] = int(os.environ.get("MSM_WINDOWS", "13"))
json.dump(summary, open(os.path.join(ROOT, "profiles", "ncu_summary_%s.json" % rnd), "w"), indent=1)
open(os.path.join(ROOT, "profiles", "ncu_%s.md" % tag), "w").write("\n".join(md) + "\n")
print(json.dumps(summary, indent=1)[:1500])
