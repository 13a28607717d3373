"""Host-side mirror of the groth16 final-layer hot path: the BN254 G1 multiexp behind `Groth16::prove`
(groth16/src/groth16.rs:88-96 -> bellman_ce::groth16::create_random_proof -> bellman_ce::multiexp).

`multiexp(bases, scalars)` mirrors bellman's `multiexp(pool, (bases, 0), FullDensity, exponents)`: bases are affine
points, exponents canonical scalar representations; the result is a projective point.  Buffers use bellman's
in-memory forms (Montgomery limbs for Fq, canonical limbs for Fr) -- see include/b200zk.h.
"""
import ctypes
import numpy as np
from . import _lib


def multiexp(bases8, scalars4):
    """bases8: (n, 8) uint64 = x||y Montgomery limbs, (0,0) = infinity; scalars4: (n, 4) uint64 canonical.
    Returns (12,) uint64 = Jacobian (X, Y, Z) Montgomery limbs."""
    b = np.ascontiguousarray(bases8, dtype=np.uint64).reshape(-1, 8)
    s = np.ascontiguousarray(scalars4, dtype=np.uint64).reshape(-1, 4)
    if b.shape[0] != s.shape[0]:
        raise ValueError("bases and exponents differ in length")      # bellman: assert_eq!(query_size, exponents.len())
    out = np.zeros(12, dtype=np.uint64)
    _lib.check(_lib.lib().b200_msm_bn254_g1(b.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p), b.shape[0], out.ctypes.data_as(ctypes.c_void_p)))
    return out


def multiexp_dev(d_bases_ptr, d_scalars_ptr, n):
    out = np.zeros(12, dtype=np.uint64)
    _lib.check(_lib.lib().b200_msm_bn254_g1_dev(ctypes.c_void_p(d_bases_ptr), ctypes.c_void_p(d_scalars_ptr), n, out.ctypes.data_as(ctypes.c_void_p)))
    return out


def g1_add(a12, b12):
    a = np.ascontiguousarray(a12, dtype=np.uint64); b = np.ascontiguousarray(b12, dtype=np.uint64)
    out = np.zeros(12, dtype=np.uint64)
    _lib.check(_lib.lib().b200_bn254_g1_add(a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p)))
    return out


def random_points_dev(d_bases_ptr, n, seed):
    _lib.check(_lib.lib().b200_bn254_g1_random_points_dev(ctypes.c_void_p(d_bases_ptr), n, seed))


def jacobian_to_affine_mont(j12):
    """(X, Y, Z) with Z in {0, R}: returns the 8-limb affine Montgomery pair ((0,0) for infinity)."""
    j = np.asarray(j12, dtype=np.uint64)
    if not j[8:].any():
        return np.zeros(8, dtype=np.uint64)
    return j[:8].copy()
