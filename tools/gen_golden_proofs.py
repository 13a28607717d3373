#!/usr/bin/env python3
"""Generate tests/golden/*.proof.json + derived_goldens.json with the CPU oracle (oracle/stark_oracle.py).

The reference has no golden proof bytes (its tests only assert `stark_verify == true`,
starky/src/stark_gen.rs:1176-1194); these files are the oracle's own outputs, each accepted by the restated
verifier, and serve (i) as regression pins for the oracle and (ii) as the bit-exact target of the CUDA path.
"""
import json, os, sys, hashlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import stark_oracle as so
from eigen_zkvm_b200 import starkinfo as si
G = os.path.join(ROOT, "tests", "golden")
out = {}
PROVER_ADDR = "273030697313060285579891744179749754319274977764"      # the address the reference's tests pass (stark_gen.rs:1014)
def run(name, pil, ss, cm, const):
    setup = so.stark_setup(const, pil, ss)
    proof = so.stark_gen(cm, const, setup, ss)
    assert so.stark_verify(proof, setup["const_root"], setup["starkinfo"], ss, setup["program"])
    js = so.proof_to_json(proof, PROVER_ADDR)
    out[name] = {"rootC": [str(x) for x in proof["rootC"]], "root1": [str(x) for x in proof["root1"]], "root2": [str(x) for x in proof["root2"]],
                 "root3": [str(x) for x in proof["root3"]], "root4": [str(x) for x in proof["root4"]],
                 "publics": [str(x) for x in proof["publics"]], "evals": [[str(x) for x in e] for e in proof["evals"]],
                 "finalPol0": [str(x) for x in proof["fri"]["last"][0]], "proof_sha256": hashlib.sha256(js.encode()).hexdigest(), "proof_len": len(js)}
    return js
ss10 = json.load(open(os.path.join(G, "starkStruct.json.gl")))
cm = np.fromfile(os.path.join(G, "fib.cm.gl"), dtype="<u8"); const = np.fromfile(os.path.join(G, "fib.const.gl"), dtype="<u8")
open(os.path.join(G, "fib10.proof.json"), "w").write(run("fib10", si.load_pil(os.path.join(G, "fib.pil.json.gl")), ss10, cm, const))
cm = np.fromfile(os.path.join(G, "plookup.cm.gl"), dtype="<u8"); const = np.fromfile(os.path.join(G, "plookup.const.gl"), dtype="<u8")
open(os.path.join(G, "plookup10.proof.json"), "w").write(run("plookup10", si.load_pil(os.path.join(G, "plookup.pil.json.gl")), ss10, cm, const))
for nm in ("pe", "connection"):      # permutation / connection fixtures (stark_gen.rs:1023-1148), here with the GL hash
    cm = np.fromfile(os.path.join(G, nm + ".cm"), dtype="<u8"); const = np.fromfile(os.path.join(G, nm + ".const"), dtype="<u8")
    open(os.path.join(G, nm + "10.proof.json"), "w").write(run(nm + "10", si.load_pil(os.path.join(G, nm + ".pil.json")), ss10, cm, const))
# BN128 / BLS12-381 hash back-ends: the reference's own fixtures (stark_gen.rs:981-1022 fib, :1093-1148 plookup; starkStruct.json{,.bls12381})
for nm, struct, tag in (("fib", "starkStruct.json", "bn128"), ("fib", "starkStruct.json.bls12381", "bls12381"), ("plookup", "starkStruct.json", "bn128")):
    ssb = json.load(open(os.path.join(G, struct)))
    cm = np.fromfile(os.path.join(G, nm + ".cm"), dtype="<u8"); const = np.fromfile(os.path.join(G, nm + ".const"), dtype="<u8")
    open(os.path.join(G, "%s10.%s.proof.json" % (nm, tag)), "w").write(run("%s10_%s" % (nm, tag), si.load_pil(os.path.join(G, nm + ".pil.json")), ssb, cm, const))
ss12 = {"nBits": 12, "nBitsExt": 13, "nQueries": 8, "verificationHashType": "GL", "steps": [{"nBits": 13}, {"nBits": 9}, {"nBits": 5}]}
cm, const = so.fibonacci_inputs(12)
run("fib12", so.fibonacci_pil(os.path.join(G, "fib.pil.json.gl"), 12), ss12, cm, const)   # sha only (proof is 90 KB)
json.dump(out, open(os.path.join(G, "derived_goldens.json"), "w"), indent=1)
print({k: v["proof_sha256"][:16] for k, v in out.items()})
