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


def test_committed_gpu_bench_lines_carry_the_contract_keys():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r1_n*_v*.json")))
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
