"""Small-size walk through every kernel family for compute-sanitizer (racecheck / memcheck) runs:
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from eigen_zkvm_b200 import starky as sk, starkinfo as si, groth16 as g16, merklehash_big as mb
G = os.path.join(ROOT, "tests", "golden")
rng = np.random.default_rng(1)
P = 0xFFFFFFFF00000001
for bits, w in [(3, 2), (6, 1), (9, 2), (10, 3), (11, 5), (12, 2), (13, 3), (14, 1), (15, 2), (16, 1), (19, 1)]:       # >= 12: the TMA-staged kernels (k_ntt3), 2 and 3 passes
    a = rng.integers(0, P, size=(1 << bits) * w, dtype=np.uint64)
    f = sk.fft(a, w, bits); b = sk.ifft(f, w, bits)
    assert (b == a).all(), (bits, w)
    sk.interpolate(a, w, bits, bits + 1)
for name, struct in [("fib", "starkStruct.json.gl"), ("plookup", "starkStruct.json.gl")]:
    pil = si.load_pil(os.path.join(G, name + ".pil.json.gl")); ss = json.load(open(os.path.join(G, struct)))
    cm = np.fromfile(os.path.join(G, name + ".cm.gl"), dtype="<u8"); const = np.fromfile(os.path.join(G, name + ".const.gl"), dtype="<u8")
    setup = sk.StarkSetup.new(const, pil, ss)
    assert sk.StarkProof.stark_gen(cm, setup) == open(os.path.join(G, name + "10.proof.json")).read()
pil = si.load_pil(os.path.join(G, "fib.pil.json")); ss = json.load(open(os.path.join(G, "starkStruct.json")))
cm = np.fromfile(os.path.join(G, "fib.cm"), dtype="<u8"); const = np.fromfile(os.path.join(G, "fib.const"), dtype="<u8")
setup = sk.StarkSetup.new(const, pil, ss)
assert sk.StarkProof.stark_gen(cm, setup, "273030697313060285579891744179749754319274977764") == open(os.path.join(G, "fib10.bn128.proof.json")).read()
t = mb.MerkleTree("BLS12381"); t.merkelize(rng.integers(0, P, size=(40, 7), dtype=np.uint64), 7, 40)
from oracle import curves as C
for cid, c in ((0, C.BN254_G1), (3, C.BLS381_G2)):
    pts = [c.mul(3 + i, c.gen) for i in range(40)]
    bases = np.array([c.affine_to_words(p) for p in pts], dtype=np.uint64)
    sc = rng.integers(0, 2**62, size=(40, 4), dtype=np.uint64)
    g16.multiexp(bases, sc, cid)
x = rng.integers(0, 2**60, size=(1 << 14, 4), dtype=np.uint64)        # 2^14: local block + radix-8 pass + radix-2 pass
assert (g16.fr_fft(g16.fr_fft(x, 0, g16.FFT), 0, g16.IFFT) == x).all()
g16.groth16_h(x, x, x, 1)
a = rng.integers(0, P, size=(1 << 15) * 5, dtype=np.uint64)
t = sk.MerkleTreeGL(); t.merkelize(a, 5, 1 << 15)            # thread-per-node levels, warp-per-node levels, fused top
print("sanitize walk ok")
