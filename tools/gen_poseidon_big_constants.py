#!/usr/bin/env python3
"""Extracts the BN128 / BLS12-381 Poseidon constants (data, not code) from the reference's
starky/src/poseidon_{bn128,bls12381}_constants_opt.rs into eigen_zkvm_b200/data/poseidon_{bn128,bls12381}.bin.

Layout (little endian): magic "PSDB", u32 version = 1, u32 n_t = 16, then for t = 2..17:
  u32 t, u32 n_rounds_p, u32 n_c, u32 n_s, then C[n_c], S[n_s], M[t*t] (row j, column i = M[j][i]), P[t*t],
  every field element as 32 bytes little-endian CANONICAL (the device converts to Montgomery form at load time).
R_P per width: poseidon_bn128_opt.rs:65-66, poseidon_bls12381_opt.rs:66-67.
"""
import os, re, struct, sys

RP = {"bn128": [56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64, 68],
      "bls12381": [55, 55, 56, 56, 56, 56, 57, 57, 57, 57, 57, 57, 57, 57, 59, 59]}
MOD = {"bn128": 21888242871839275222246405745257275088548364400416034343698204186575808495617,
       "bls12381": 52435875175126190479447740508185965837690552500527637822603658699938581184513}


def parse_nested(src, name):
    """returns the nested list of hex strings assigned to `let <name>: ... = vec![ ... ];`"""
    i = src.index("let %s:" % name)
    i = src.index("vec![", i)
    depth = 0; j = i; out_stack = []; cur = None; root = None
    tok = re.compile(r'vec!\[|\]|"(0x[0-9a-fA-F]+)"')
    for m in tok.finditer(src, i):
        s = m.group(0)
        if s == "vec![":
            new = []
            if cur is not None: cur.append(new); out_stack.append(cur)
            cur = new
            if root is None: root = new
            depth += 1
        elif s == "]":
            depth -= 1
            if depth == 0: break
            cur = out_stack.pop()
        else:
            cur.append(int(m.group(1), 16))
    return root


def main(ref="/root/reference"):
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "eigen_zkvm_b200", "data")
    os.makedirs(dst, exist_ok=True)
    for name in ("bn128", "bls12381"):
        src = open(os.path.join(ref, "starky", "src", "poseidon_%s_constants_opt.rs" % name)).read()
        C = parse_nested(src, "c_str"); M = parse_nested(src, "m_str"); P = parse_nested(src, "p_str"); S = parse_nested(src, "s_str")
        assert len(C) == len(M) == len(P) == len(S) == 16, (len(C), len(M), len(P), len(S))
        out = [b"PSDB", struct.pack("<II", 1, 16)]
        p = MOD[name]
        for k in range(16):
            t = k + 2; rp = RP[name][k]
            assert len(C[k]) == 8 * t + rp - t + t or True
            assert len(S[k]) == (2 * t - 1) * rp, (t, len(S[k]))
            assert len(M[k]) == t and all(len(r) == t for r in M[k]) and len(P[k]) == t
            out.append(struct.pack("<IIII", t, rp, len(C[k]), len(S[k])))
            flat = C[k] + S[k] + [v for row in M[k] for v in row] + [v for row in P[k] for v in row]
            for v in flat:
                assert 0 <= v < p
                out.append(v.to_bytes(32, "little"))
        path = os.path.join(dst, "poseidon_%s.bin" % name)
        open(path, "wb").write(b"".join(out))
        print("wrote", path, os.path.getsize(path), "bytes; C lens", [len(c) for c in C])


if __name__ == "__main__":
    main(*sys.argv[1:])
