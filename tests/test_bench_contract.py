"""bench.py's output contract, checked on the CPU box: the reference arm prints exactly one JSON line with the required
keys, and the committed GPU bench lines under profiles/ carry every key the driver reads."""
import glob, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e"]


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--log-n", "12", "--cpu-sample-log-n", "12"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in BASE_KEYS + ["impl", "cpu_baseline"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is False and d["unit"] == "s/proof"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    # the line says what it ran: full size measured here, so the workload string is our arm's and nothing is extrapolated
    assert d["steps_measured"] == 1 and d["config"]["reference_rows_log2"] == 12
    assert "measured: 1 genuine stark_gen run(s) at 2^12 rows" in d["config"]["reference_run"] and "scalar" in d["cpu_baseline"]["sample"]
    assert "EXTRAPOLATED" not in d["config"]["workload"]


def test_reference_arm_labels_an_extrapolated_run():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--log-n", "14", "--cpu-sample-log-n", "12"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.strip()][0])
    assert "EXTRAPOLATED" in d["config"]["workload"] and "2^12 rows" in d["config"]["workload"] and d["config"]["reference_rows_log2"] == 12
    assert abs(d["value"] - d["ms_per_step"] / 1e3 * (2 ** 14 * 15) / (2 ** 12 * 13)) < 1e-9


def test_committed_gpu_bench_lines_carry_the_contract_keys():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r[0-9]_n*_v*.json")))
    assert files
    latest = [f for f in files if "_n1_" in f][-1]
    d = json.load(open(latest))
    for k in BASE_KEYS + ["clocks", "gpu_launches", "roofline", "cpu_baseline"]:
        assert k in d, k
    assert d["metric"] == "stark_proof_gen_seconds" and d["n_gpus"] == 1 and d["gpu_launches"] > 0 and d["data"] == "synthetic"
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 16 << 24 and e["d2h_bytes_per_step"] > 0 and e["value"] >= d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and "sample" in c
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_headline_verification_helper_on_an_oracle_proof():
    """bench.verify_headline_proof (what bench.py runs on the 2^24 GPU proof) accepts a genuine proof, and its closed forms
    (F_N by matrix powers, ISLAST at xi and on the coset) agree with the oracle's own setup at 2^12 rows."""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    from oracle import stark_oracle as so
    nbits = 12
    ss = bench.stark_struct(nbits)
    cm, const = so.fibonacci_inputs(nbits)
    setup = so.stark_setup(const, bench.fib_pil(nbits), ss)
    js = so.proof_to_json(so.stark_gen(cm, const, setup, ss))
    rep = bench.verify_headline_proof(js, nbits, ss, setup["const_root"])
    assert rep["accepted"] and rep["tampered_rejected"] == 3 and len(rep["closed_form_checks"]) == 3
    import pytest
    with pytest.raises(AssertionError):
        bench.verify_headline_proof(js, nbits, ss, [1, 2, 3, 4])        # wrong constant root
