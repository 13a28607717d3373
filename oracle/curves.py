"""BN254 / BLS12-381 G1 and G2 oracle -- TEST INFRASTRUCTURE (python big ints, affine double-and-add).

PARITY UNPINNED by the reference: the groth16 MSMs live in un-vendored crates (bellman_ce 0.3.2 / pairing_ce for
BN254, bellperson 0.26 / blstrs 0.7.1 for BLS12-381; Cargo.lock:668-670,731-733), and proofs are randomised, so no
reference test fixes an MSM output.  Pinned here instead:
  * BN254 Fq, Fr moduli: groth16/src/api.rs:636, starky/src/field_bn128.rs:12; BLS12-381 Fr: starky/src/field_bls12381.rs:12
  * curve equations and twists by the on-curve G1/G2 points the reference ships in groth16/test-vectors/
    verification_key.json (BN254) and verification_key_bls12381.json (tests/test_oracle_curves.py)
  * generators: r * G = O on all four groups.
Fp2 = Fp[u]/(u^2 + 1); an Fp2 element is a tuple (c0, c1); G2 twists: y^2 = x^3 + 3/(9+u) (BN254, D-type),
y^2 = x^3 + 4(1+u) (BLS12-381, M-type).
"""


class Field:
    """Fp (deg = 1: ints) or Fp2 (deg = 2: tuples) arithmetic over a prime p."""

    def __init__(self, p, deg):
        self.p, self.deg = p, deg
        self.zero = 0 if deg == 1 else (0, 0)
        self.one = 1 if deg == 1 else (1, 0)

    def add(self, a, b):
        p = self.p
        return (a + b) % p if self.deg == 1 else ((a[0] + b[0]) % p, (a[1] + b[1]) % p)

    def sub(self, a, b):
        p = self.p
        return (a - b) % p if self.deg == 1 else ((a[0] - b[0]) % p, (a[1] - b[1]) % p)

    def neg(self, a):
        return self.sub(self.zero, a)

    def mul(self, a, b):
        p = self.p
        if self.deg == 1:
            return a * b % p
        return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

    def muli(self, a, k):
        p = self.p
        return a * k % p if self.deg == 1 else (a[0] * k % p, a[1] * k % p)

    def inv(self, a):
        p = self.p
        if self.deg == 1:
            return pow(a, p - 2, p)
        n = pow((a[0] * a[0] + a[1] * a[1]) % p, p - 2, p)
        return (a[0] * n % p, (-a[1] * n) % p)

    def is_zero(self, a):
        return a == self.zero


class Curve:
    def __init__(self, name, p, r, deg, b, gen, limbs32):
        self.name, self.p, self.r, self.F, self.b, self.gen = name, p, r, Field(p, deg), b, gen
        self.deg, self.limbs32 = deg, limbs32          # limbs32: 32-bit limbs per Fp element
        self.mont_r = (1 << (32 * limbs32)) % p

    def is_on_curve(self, P):
        if P is None: return True
        F = self.F
        x, y = P
        return F.sub(F.mul(y, y), F.add(F.mul(F.mul(x, x), x), self.b)) == F.zero

    def add(self, P, Q):
        F = self.F
        if P is None: return Q
        if Q is None: return P
        if P[0] == Q[0]:
            if F.is_zero(F.add(P[1], Q[1])): return None
            l = F.mul(F.muli(F.mul(P[0], P[0]), 3), F.inv(F.muli(P[1], 2)))
        else:
            l = F.mul(F.sub(Q[1], P[1]), F.inv(F.sub(Q[0], P[0])))
        x = F.sub(F.sub(F.mul(l, l), P[0]), Q[0])
        return (x, F.sub(F.mul(l, F.sub(P[0], x)), P[1]))

    def neg(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))

    def mul(self, k, P):
        acc = None
        while k:
            if k & 1: acc = self.add(acc, P)
            P = self.add(P, P); k >>= 1
        return acc

    def msm_naive(self, points, scalars):
        acc = None
        for P, s in zip(points, scalars):
            acc = self.add(acc, self.mul(s % self.r, P))
        return acc

    # ---- boundary forms (include/b200zk.h): little-endian u64 limbs, Montgomery, (0,..,0) = infinity ----
    def _fp_to_u64(self, v):
        v = v * self.mont_r % self.p
        return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(self.limbs32 // 2)]

    def _f_to_u64(self, a):
        return self._fp_to_u64(a) if self.deg == 1 else self._fp_to_u64(a[0]) + self._fp_to_u64(a[1])

    def _fp_from_u64(self, w):
        v = sum(int(x) << (64 * i) for i, x in enumerate(w))
        return v * pow(self.mont_r, -1, self.p) % self.p

    def _f_from_u64(self, w):
        n = self.limbs32 // 2
        return self._fp_from_u64(w[:n]) if self.deg == 1 else (self._fp_from_u64(w[:n]), self._fp_from_u64(w[n:2 * n]))

    @property
    def f_words(self):          # u64 words per coordinate
        return self.limbs32 // 2 * self.deg

    def affine_to_words(self, P):
        if P is None: return [0] * (2 * self.f_words)
        return self._f_to_u64(P[0]) + self._f_to_u64(P[1])

    def affine_from_words(self, w):
        w = [int(x) for x in w]
        if not any(w): return None
        n = self.f_words
        return (self._f_from_u64(w[:n]), self._f_from_u64(w[n:2 * n]))

    def jacobian_from_words(self, w):
        """(X, Y, Z) Montgomery words -> affine point (None = infinity)."""
        w = [int(x) for x in w]
        n = self.f_words
        F = self.F
        X, Y, Z = self._f_from_u64(w[:n]), self._f_from_u64(w[n:2 * n]), self._f_from_u64(w[2 * n:3 * n])
        if F.is_zero(Z): return None
        zi = F.inv(Z); zi2 = F.mul(zi, zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))


BN254_Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
BN254_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
BLS381_Q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
BLS381_R = 52435875175126190479447740508185965837690552500527637822603658699938581184513

_f2 = Field(BN254_Q, 2)
BN254_G1 = Curve("bn254_g1", BN254_Q, BN254_R, 1, 3, (1, 2), 8)
BN254_G2 = Curve("bn254_g2", BN254_Q, BN254_R, 2, _f2.mul((3, 0), _f2.inv((9, 1))),
                 ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
                   11559732032986387107991004021392285783925812861821192530917403151452391805634),
                  (8495653923123431417604973247489272438418190587263600148770280649306958101930,
                   4082367875863433681332203403145435568316851327593401208105741076214120093531)), 8)
BLS381_G1 = Curve("bls12381_g1", BLS381_Q, BLS381_R, 1, 4,
                  (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
                   0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1), 12)
BLS381_G2 = Curve("bls12381_g2", BLS381_Q, BLS381_R, 2, (4, 4),
                  ((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
                    0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
                   (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
                    0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be)), 12)
CURVES = {c.name: c for c in (BN254_G1, BN254_G2, BLS381_G1, BLS381_G2)}


# ---- C Pippenger over the same boundary forms (oracle/curves_oracle.c): arbitrates at sizes python cannot reach ---------------
_CLIB = None
SCALAR_BITS = {"bn254_g1": 254, "bn254_g2": 254, "bls12381_g1": 255, "bls12381_g2": 255}


def clib():
    global _CLIB
    if _CLIB is None:
        import ctypes, os, subprocess
        here = os.path.dirname(os.path.abspath(__file__))
        so = os.path.join(here, "libcurves_oracle.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(here, "curves_oracle.c")):
            subprocess.check_call(["make", "-s", "-C", here, "libcurves_oracle.so"])
        _CLIB = ctypes.CDLL(so)
        _CLIB.cv_msm.restype = ctypes.c_int
        _CLIB.cv_num_threads.restype = ctypes.c_int
    return _CLIB


def msm_c(curve, bases_words, scalars_words):
    """bases_words: (n, 2 * f_words) uint64 Montgomery affine (all-zero = infinity); scalars_words: (n, 4) uint64 canonical.
    Returns the affine result as (2 * f_words,) uint64 Montgomery words (all-zero = infinity)."""
    import ctypes
    import numpy as np
    b = np.ascontiguousarray(bases_words, dtype=np.uint64).reshape(-1, 2 * curve.f_words)
    s = np.ascontiguousarray(scalars_words, dtype=np.uint64).reshape(-1, 4)
    assert b.shape[0] == s.shape[0]
    limbs = curve.limbs32 // 2
    p = np.array([(curve.p >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(limbs)], dtype=np.uint64)
    out = np.zeros(2 * curve.f_words, dtype=np.uint64)
    rc = clib().cv_msm(limbs, curve.deg, p.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p),
                       ctypes.c_size_t(b.shape[0]), SCALAR_BITS[curve.name], out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out
