import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g; g.build()
from eigen_zkvm_b200 import starky, starkinfo as si, synthetic as syn
from oracle import stark_oracle as so
nb = 8
pil = syn.wide_fib_pil(nb, 6, 31)
for hash_type in ("BN128",):
    ss = {"nBits": nb, "nBitsExt": nb + 1, "nQueries": 8, "verificationHashType": hash_type, "steps": [{"nBits": 9}, {"nBits": 5}, {"nBits": 2}]}
    cm, const = syn.wide_fib_trace(nb, 6, 31)
    setup = starky.StarkSetup.new(const, si.load_pil(pil), ss)
    js = starky.StarkProof.stark_gen(cm, setup, "0x1")
    osetup = so.stark_setup(const, si.load_pil(pil), ss)
    oproof = so.stark_gen(cm, const, osetup, ss)
    a = json.loads(js); b = json.loads(so.proof_to_json(oproof, "0x1"))
    print(list(a.keys()) == list(b.keys()))
    for k in a:
        if a[k] != b.get(k): print("DIFF", k, str(a[k])[:200], "|||", str(b.get(k))[:200])
    x = a["s2_siblings"]; y = b["s2_siblings"]
    print(len(x), len(y), len(x[0]), len(y[0]), [len(v) for v in x[0]], [len(v) for v in y[0]])
    for q in range(len(x)):
        for d in range(len(x[q])):
            if x[q][d] != y[q][d]:
                print("q", q, "d", d); print(" gpu", [v[-6:] for v in x[q][d]]); print(" ora", [v[-6:] for v in y[q][d]]); break
        else: continue
        break
