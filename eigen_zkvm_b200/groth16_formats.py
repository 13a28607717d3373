"""Disk formats at the groth16 boundary: bellman's `VerifyingKey::read` / `Proof::read` binary layouts and
their JSON twins (groth16/src/json_utils.rs:20-120,200-345; round-trip tests :350-429 on groth16/test-vectors/*), plus the
conversion of affine points to the in-memory Montgomery limbs that `b200_msm*` takes as bases.

verification_key.bin  = alpha_g1, beta_g1, beta_g2, gamma_g2, delta_g1, delta_g2 (uncompressed), u32 BE count, IC[count]
   G1 uncompressed = x || y big-endian (32 B each for BN254, 48 B for BLS12-381); G2 = x.c1 || x.c0 || y.c1 || y.c0
proof.bin (BN254)     = A (G1 compressed, 32 B), B (G2 compressed, 64 B), C (G1 compressed): big-endian x with the two
   spare top bits of byte 0 as flags: 0x40 = point at infinity, 0x80 = y is the lexicographically larger root.
Host-side integer code only; no field arithmetic of the hot path lives here."""
import json

FIELDS = {
    "BN128": dict(q=21888242871839275222246405745257275088696311157297823662689037894645226208583, nbytes=32, limbs64=4, b=3),
    "BLS12381": dict(q=0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab, nbytes=48, limbs64=6, b=4),
}


def _be(b): return int.from_bytes(b, "big")


class _Reader:
    def __init__(self, data): self.d, self.o = bytes(data), 0
    def take(self, n):
        if self.o + n > len(self.d): raise ValueError("unexpected end of file")
        v = self.d[self.o:self.o + n]; self.o += n; return v


def _read_g1(r, f):
    n = FIELDS[f]["nbytes"]
    return (_be(r.take(n)), _be(r.take(n)))


def _read_g2(r, f):
    n = FIELDS[f]["nbytes"]
    xc1, xc0, yc1, yc0 = (_be(r.take(n)) for _ in range(4))
    return ((xc0, xc1), (yc0, yc1))


def read_vk_bin(data, curve="BN128"):
    """bellman `VerifyingKey::read` -> the dict `serialize_vk` writes (json_utils.rs: `VerifyingKeyFile`)."""
    r = _Reader(data)
    a1 = _read_g1(r, curve); b1 = _read_g1(r, curve); b2 = _read_g2(r, curve); g2 = _read_g2(r, curve); d1 = _read_g1(r, curve); d2 = _read_g2(r, curve)
    n_ic = _be(r.take(4))
    ic = [_read_g1(r, curve) for _ in range(n_ic)]
    if r.o != len(r.d): raise ValueError("trailing bytes in verification key")
    g1j = lambda p: {"x": str(p[0]), "y": str(p[1])}
    g2j = lambda p: {"x": [str(p[0][0]), str(p[0][1])], "y": [str(p[1][0]), str(p[1][1])]}
    return {"protocol": "groth16", "curve": curve, "vk_alpha_1": g1j(a1), "vk_beta_1": g1j(b1), "vk_beta_2": g2j(b2), "vk_gamma_2": g2j(g2),
            "vk_delta_1": g1j(d1), "vk_delta_2": g2j(d2), "IC": [g1j(p) for p in ic]}


# ---- decompression (BN254: q = 3 mod 4) -----------------------------------------------------------------------------
def _sqrt_fq(a, q):
    y = pow(a, (q + 1) // 4, q)
    return y if y * y % q == a % q else None


def _f2_mul(a, b, q): return ((a[0] * b[0] - a[1] * b[1]) % q, (a[0] * b[1] + a[1] * b[0]) % q)


def _f2_pow(a, e, q):
    r = (1, 0)
    while e:
        if e & 1: r = _f2_mul(r, a, q)
        a = _f2_mul(a, a, q); e >>= 1
    return r


def _sqrt_fq2(a, q):
    """square root in Fq[u]/(u^2+1) for q = 3 mod 4 (Adj--Rodriguez-Henriquez, alg. 9)"""
    if a == (0, 0): return (0, 0)
    a1 = _f2_pow(a, (q - 3) // 4, q)
    alpha = _f2_mul(_f2_mul(a1, a1, q), a, q)
    x0 = _f2_mul(a1, a, q)
    if alpha == (q - 1, 0):
        x = ((-x0[1]) % q, x0[0])                    # u * x0
    else:
        b = _f2_pow(((alpha[0] + 1) % q, alpha[1]), (q - 1) // 2, q)
        x = _f2_mul(b, x0, q)
    return x if _f2_mul(x, x, q) == (a[0] % q, a[1] % q) else None


def _decompress_g1(b, curve):
    f = FIELDS[curve]; q = f["q"]
    flags, x = b[0] & 0xC0, _be(bytes([b[0] & 0x3F]) + b[1:])
    if flags & 0x40: return None
    y = _sqrt_fq((x * x * x + f["b"]) % q, q)
    if y is None: raise ValueError("not on the curve")
    if (y > q - y) != bool(flags & 0x80): y = q - y
    return (x, y)


def _decompress_g2(b, curve):
    f = FIELDS[curve]; q = f["q"]; n = f["nbytes"]
    flags = b[0] & 0xC0
    if flags & 0x40: return None
    xc1 = _be(bytes([b[0] & 0x3F]) + b[1:n]); xc0 = _be(b[n:2 * n])
    x = (xc0, xc1)
    if curve != "BN128": raise NotImplementedError("compressed G2 is only needed for BN254 proofs here")
    inv = pow((9 * 9 + 1) % q, q - 2, q)             # b' = 3 / (9 + u)
    bt = (3 * 9 * inv % q, (-3) * inv % q)
    x3 = _f2_mul(_f2_mul(x, x, q), x, q)
    y = _sqrt_fq2(((x3[0] + bt[0]) % q, (x3[1] + bt[1]) % q), q)
    if y is None: raise ValueError("not on the curve")
    neg = ((-y[0]) % q, (-y[1]) % q)
    larger = (y[1], y[0]) > (neg[1], neg[0])         # lexicographic: c1 first, then c0
    if larger != bool(flags & 0x80): y = neg
    return (x, y)


def read_proof_bin(data, curve="BN128"):
    """bellman `Proof::read` (A, B, C compressed) -> the dict `serialize_proof` writes (`ProofFile`)."""
    n = FIELDS[curve]["nbytes"]
    r = _Reader(data)
    a = _decompress_g1(r.take(n), curve); b = _decompress_g2(r.take(2 * n), curve); c = _decompress_g1(r.take(n), curve)
    if r.o != len(r.d): raise ValueError("trailing bytes in proof")
    return {"pi_a": {"x": str(a[0]), "y": str(a[1])}, "pi_b": {"x": [str(b[0][0]), str(b[0][1])], "y": [str(b[1][0]), str(b[1][1])]},
            "pi_c": {"x": str(c[0]), "y": str(c[1])}, "protocol": "groth16", "curve": curve}


# ---- in-memory bases for b200_msm* -------------------------------------------------------------------------------------
def g1_to_words(points, curve="BN128"):
    """[(x, y) | None] -> (n, 2 * limbs64) uint64: Montgomery limbs, all-zero = infinity (include/b200zk.h)."""
    import numpy as np
    f = FIELDS[curve]; q, L = f["q"], f["limbs64"]; R = 1 << (64 * L)
    out = np.zeros((len(points), 2 * L), dtype=np.uint64)
    for i, p in enumerate(points):
        if p is None: continue
        for k, v in enumerate(p):
            m = v * R % q
            for j in range(L): out[i, k * L + j] = (m >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def g2_to_words(points, curve="BN128"):
    """[((x.c0, x.c1), (y.c0, y.c1)) | None] -> (n, 4 * limbs64) uint64: x.c0 || x.c1 || y.c0 || y.c1 Montgomery limbs."""
    import numpy as np
    f = FIELDS[curve]; q, L = f["q"], f["limbs64"]; R = 1 << (64 * L)
    out = np.zeros((len(points), 4 * L), dtype=np.uint64)
    for i, p in enumerate(points):
        if p is None: continue
        for k, v in enumerate((p[0][0], p[0][1], p[1][0], p[1][1])):
            m = v * R % q
            for j in range(L): out[i, k * L + j] = (m >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def vk_points(vk):
    """the G1 / G2 points of a parsed verification key as integer tuples (for g1_to_words / g2_to_words)"""
    g1 = lambda p: (int(p["x"]), int(p["y"]))
    g2 = lambda p: ((int(p["x"][0]), int(p["x"][1])), (int(p["y"][0]), int(p["y"][1])))
    return {"g1": [g1(vk[k]) for k in ("vk_alpha_1", "vk_beta_1", "vk_delta_1")] + [g1(p) for p in vk["IC"]],
            "g2": [g2(vk[k]) for k in ("vk_beta_2", "vk_gamma_2", "vk_delta_2")]}
