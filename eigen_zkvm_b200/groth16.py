"""Host-side mirror of the groth16 final-layer hot path: the G1/G2 multiexps behind `Groth16::prove`
(groth16/src/groth16.rs:88-96 -> bellman_ce::groth16::create_random_proof -> bellman_ce::multiexp for BN254;
groth16.rs:45-57 -> bellperson + blstrs for BLS12-381).

`multiexp(bases, scalars, curve)` mirrors bellman's `multiexp(pool, (bases, 0), FullDensity, exponents)`: bases are
affine points, exponents canonical scalar representations; the result is a projective point.  Buffers use the
libraries' in-memory forms (Montgomery limbs for Fq, canonical limbs for Fr) -- see include/b200zk.h.
"""
import ctypes
import numpy as np
from . import _lib

BN254_G1, BN254_G2, BLS12381_G1, BLS12381_G2 = 0, 1, 2, 3
CURVE_NAMES = {BN254_G1: "bn254_g1", BN254_G2: "bn254_g2", BLS12381_G1: "bls12381_g1", BLS12381_G2: "bls12381_g2"}


def point_words(curve):
    """u64 words of one affine base (8 / 16 / 12 / 24); a Jacobian result has 1.5x as many."""
    b = _lib.lib().b200_msm_point_bytes(curve)
    if not b:
        raise ValueError("unknown curve id %r" % (curve,))
    return b // 8


def multiexp(bases, scalars4, curve=BN254_G1):
    """bases: (n, point_words) uint64 = x||y Montgomery limbs, all-zero = infinity; scalars4: (n, 4) uint64 canonical.
    Returns (1.5 * point_words,) uint64 = Jacobian (X, Y, Z) Montgomery limbs."""
    pw = point_words(curve)
    b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, pw)
    s = np.ascontiguousarray(scalars4, dtype=np.uint64).reshape(-1, 4)
    if b.shape[0] != s.shape[0]:
        raise ValueError("bases and exponents differ in length")      # bellman: assert_eq!(query_size, exponents.len())
    out = np.zeros(pw * 3 // 2, dtype=np.uint64)
    _lib.check(_lib.lib().b200_msm(curve, b.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p), b.shape[0], out.ctypes.data_as(ctypes.c_void_p)))
    return out


def multiexp_dev(d_bases_ptr, d_scalars_ptr, n, curve=BN254_G1):
    out = np.zeros(point_words(curve) * 3 // 2, dtype=np.uint64)
    _lib.check(_lib.lib().b200_msm_dev(curve, ctypes.c_void_p(d_bases_ptr), ctypes.c_void_p(d_scalars_ptr), n, out.ctypes.data_as(ctypes.c_void_p)))
    return out


def point_add(a, b, curve=BN254_G1):
    a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.zeros(point_words(curve) * 3 // 2, dtype=np.uint64)
    _lib.check(_lib.lib().b200_point_add(curve, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p)))
    return out


def g1_add(a12, b12):
    return point_add(a12, b12, BN254_G1)


def random_points_dev(d_bases_ptr, n, seed, curve=BN254_G1):
    _lib.check(_lib.lib().b200_random_points_dev(curve, ctypes.c_void_p(d_bases_ptr), n, seed))


def jacobian_to_affine_mont(j, curve=BN254_G1):
    """(X, Y, Z) with Z in {0, R}: returns the affine Montgomery pair (all-zero for infinity)."""
    j = np.asarray(j, dtype=np.uint64)
    pw = point_words(curve)
    if not j[pw:].any():
        return np.zeros(pw, dtype=np.uint64)
    return j[:pw].copy()


# ---- scalar-field domain: bellman_ce `EvaluationDomain` and the quotient of `create_random_proof` -------------------------
FR_BN254, FR_BLS12381 = 0, 1
FFT, IFFT, COSET_FFT, ICOSET_FFT = 0, 1, 2, 3


def fr_fft(vals4, field=FR_BN254, mode=FFT):
    """vals4: (2^k, 4) uint64 Montgomery `Fr`s; returns the transformed copy (fft / ifft / coset_fft / icoset_fft)."""
    a = np.ascontiguousarray(vals4, dtype=np.uint64).reshape(-1, 4).copy()
    n = a.shape[0]
    if n == 0 or n & (n - 1):
        raise ValueError("domain size must be a power of two")
    _lib.check(_lib.lib().b200_fr_fft(field, a.ctypes.data_as(ctypes.c_void_p), n.bit_length() - 1, mode))
    return a


def groth16_h(a4, b4, c4, field=FR_BN254):
    """a, b, c: (m, 4) uint64 Montgomery evaluations of the QAP polynomials on the domain; returns (m - 1, 4) canonical
    coefficients of H = (A * B - C) / Z, the exponents of the `h` multiexp."""
    a = np.ascontiguousarray(a4, dtype=np.uint64).reshape(-1, 4); b = np.ascontiguousarray(b4, dtype=np.uint64).reshape(-1, 4)
    c = np.ascontiguousarray(c4, dtype=np.uint64).reshape(-1, 4)
    m = a.shape[0]
    if m == 0 or m & (m - 1) or b.shape[0] != m or c.shape[0] != m:
        raise ValueError("a, b, c must have the same power-of-two length")
    out = np.zeros((max(m - 1, 1), 4), dtype=np.uint64)
    _lib.check(_lib.lib().b200_groth16_h(field, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), c.ctypes.data_as(ctypes.c_void_p),
                                         m.bit_length() - 1, out.ctypes.data_as(ctypes.c_void_p)))
    return out[:m - 1]
