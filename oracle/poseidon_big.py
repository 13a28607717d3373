"""BN128 / BLS12-381 Poseidon, LinearHash and 16-ary Merkle tree oracle -- TEST INFRASTRUCTURE (python big ints).

CPU restatement of the reference's last-layer hash stack:
  Poseidon (variable width t = len(inputs) + 1 <= 17, x^5, optimised rounds)  starky/src/poseidon_bn128_opt.rs:94-225,
                                                                               poseidon_bls12381_opt.rs:95-231
  LinearHash  starky/src/linearhash_bn128.rs:23-131, linearhash_bls12381.rs (same code, other field)
  MerkleTree  starky/src/merklehash_bn128.rs:26-87,176-224, merklehash_bls12381.rs
Pinned to the reference's KATs in tests/test_oracle_poseidon_big.py (poseidon_bn128_opt.rs:232-300,
poseidon_bls12381_opt.rs:237-310, linearhash_bn128.rs:140-175, linearhash_bls12381.rs:139-192,
merklehash_bn128.rs:270-292, merklehash_bls12381.rs:274-293).
Values are field elements (python ints in [0, r)); the reference keeps digests as the 4 Montgomery limbs of the scalar
(digest.rs:45-65) -- `to_ref_limbs` converts for the KATs that assert those limbs.
Constants: eigen_zkvm_b200/data/poseidon_{bn128,bls12381}.bin, extracted from the reference's constant tables by
tools/gen_poseidon_big_constants.py (data only).
"""
import os, struct

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "eigen_zkvm_b200", "data")
MOD = {"bn128": 21888242871839275222246405745257275088548364400416034343698204186575808495617,      # starky/src/field_bn128.rs:12
       "bls12381": 52435875175126190479447740508185965837690552500527637822603658699938581184513}   # starky/src/field_bls12381.rs:12
OUT_LANE = {"bn128": 0, "bls12381": 1}     # poseidon_bn128_opt.rs:94-97 returns state[0]; poseidon_bls12381_opt.rs:95-103 state[1]
_CONST = {}


def constants(field):
    if field not in _CONST:
        b = open(os.path.join(_DATA, "poseidon_%s.bin" % field), "rb").read()
        assert b[:4] == b"PSDB"
        ver, nt = struct.unpack_from("<II", b, 4)
        off = 12; tabs = {}
        for _ in range(nt):
            t, rp, nc, ns = struct.unpack_from("<IIII", b, off); off += 16
            def take(n):
                nonlocal off
                v = [int.from_bytes(b[off + 32 * i: off + 32 * i + 32], "little") for i in range(n)]
                off += 32 * n
                return v
            C = take(nc); S = take(ns); M = take(t * t); P = take(t * t)
            tabs[t] = (rp, C, S, M, P)
        _CONST[field] = tabs
    return _CONST[field]


def permute(field, inputs, init_state=0):
    """hash_inner: returns the full state after the permutation."""
    p = MOD[field]
    t = len(inputs) + 1
    if not 2 <= t <= 17:
        raise ValueError("Wrong inputs length")
    rp, C, S, M, P = constants(field)[t]
    st = [init_state] + list(inputs)
    st = [(a + C[i]) % p for i, a in enumerate(st)]
    def mix(mat, st):
        return [sum(mat[j * t + i] * st[j] for j in range(t)) % p for i in range(t)]
    for r in range(3):
        st = [(pow(a, 5, p) + C[(r + 1) * t + i]) % p for i, a in enumerate(st)]
        st = mix(M, st)
    st = [(pow(a, 5, p) + C[4 * t + i]) % p for i, a in enumerate(st)]
    st = mix(P, st)
    for r in range(rp):
        st[0] = (pow(st[0], 5, p) + C[5 * t + r]) % p
        base = (2 * t - 1) * r
        s0 = sum(S[base + j] * st[j] for j in range(t)) % p
        for k in range(1, t):
            st[k] = (st[k] + S[base + t + k - 1] * st[0]) % p
        st[0] = s0
    for r in range(3):
        st = [(pow(a, 5, p) + C[5 * t + rp + r * t + i]) % p for i, a in enumerate(st)]
        st = mix(M, st)
    st = [pow(a, 5, p) for a in st]
    return mix(M, st)


def hash(field, inputs, init_state=0):
    return permute(field, inputs, init_state)[OUT_LANE[field]]


def pack3(vals):
    """to_bn128 / to_bls12381 (digest.rs:161-190) on chunks of 3 GL elements."""
    return [sum(v << (64 * i) for i, v in enumerate(vals[k:k + 3])) for k in range(0, len(vals), 3)]


def hash_element_array(field, vals):
    """LinearHash*::hash_element_array (linearhash_bn128.rs:105-131): the leaf digest of one row of GL elements."""
    p = MOD[field]
    if len(vals) <= 4:
        return sum(int(v) << (64 * i) for i, v in enumerate(vals)) % p      # to_bn128_mont: the raw 256-bit integer mod r
    buf = pack3([int(v) for v in vals])
    d = 0
    for i in range(0, len(buf), 16):
        d = hash(field, buf[i:i + 16], d)
    return d


def hash_element_matrix(field, columns):
    """LinearHash*::hash_element_matrix (linearhash_bn128.rs:23-67)."""
    flat = [int(e) for col in columns for e in col]
    vals3 = pack3(flat)
    vals3 = [v % MOD[field] for v in vals3]
    if not vals3: return 0
    if len(vals3) == 1: return vals3[0]
    st = 0
    for i in range(0, len(vals3), 16):
        st = hash(field, vals3[i:i + 16], st)
    return st


def get_n_nodes(n):
    """merklehash_bn128.rs:26-40"""
    nn = (n - 1) // 16 + 1
    acc = nn * 16
    while n > 1:
        n = nn; nn = (n - 1) // 16 + 1
        acc += nn * 16 if n > 1 else 1
    return acc


def merkelize(field, rows):
    """MerkleTree*::merkelize (merklehash_bn128.rs:176-224): returns `nodes` (field elements, zero padded levels)."""
    height = len(rows)
    nodes = [0] * get_n_nodes(height)
    for i, r in enumerate(rows):
        nodes[i] = hash_element_array(field, r)
    n = height; nn = (n - 1) // 16 + 1; p_in = 0; p_out = nn * 16
    while n > 1:
        for i in range(nn):
            nodes[p_out + i] = hash(field, nodes[p_in + 16 * i: p_in + 16 * i + 16], 0)
        n = nn; nn = (n - 1) // 16 + 1; p_in = p_out; p_out = p_in + nn * 16
    return nodes


def to_ref_limbs(field, x):
    """the reference's in-memory digest: 4 u64 limbs of the Montgomery representative x * 2^256 mod r (digest.rs:45-65)."""
    m = x * (1 << 256) % MOD[field]
    return [(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
