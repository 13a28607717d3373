"""BN254 (alt_bn128) G1 oracle -- TEST INFRASTRUCTURE.  An independent python big-int implementation
(affine double-and-add) plus a ctypes front-end of the C Pippenger in bn254_oracle.c (the CPU baseline).
PARITY UNPINNED by the reference (its MSM lives in un-vendored bellman_ce 0.3.2); see bn254_oracle.c header.
Constants: Fq modulus groth16/src/api.rs:636; Fr modulus starky/src/field_bn128.rs:12; curve y^2 = x^3 + 3,
generator (1, 2) (the Solidity verifier template in groth16/src/api.rs uses the same group).
"""
import ctypes, os, subprocess
import numpy as np

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
G1 = (1, 2)
MONT_R = (1 << 256) % Q
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libbn254_oracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        L = ctypes.CDLL(so)
        L.bn_is_on_curve.restype = ctypes.c_int
        L.bn_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


# ---- python big-int reference (affine, None = infinity) -----------------------------------------------
def is_on_curve(p):
    return p is None or (p[1] * p[1] - p[0] ** 3 - 3) % Q == 0


def add(p, q):
    if p is None: return q
    if q is None: return p
    if p[0] == q[0]:
        if (p[1] + q[1]) % Q == 0: return None
        l = 3 * p[0] * p[0] * pow(2 * p[1], Q - 2, Q) % Q
    else:
        l = (q[1] - p[1]) * pow(q[0] - p[0], Q - 2, Q) % Q
    x = (l * l - p[0] - q[0]) % Q
    return (x, (l * (p[0] - x) - p[1]) % Q)


def mul(k, p):
    acc = None
    while k:
        if k & 1: acc = add(acc, p)
        p = add(p, p); k >>= 1
    return acc


def msm_naive(points, scalars):
    acc = None
    for p, s in zip(points, scalars):
        acc = add(acc, mul(s % R, p))
    return acc


# ---- packing: affine Montgomery 8 x u64 per point, (0,0) = infinity -------------------------------------
def _limbs(v): return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
def _unlimbs(a): return sum(int(a[i]) << (64 * i) for i in range(4))


def pack_points(points):
    out = np.zeros((len(points), 8), dtype=np.uint64)
    for i, p in enumerate(points):
        if p is not None:
            out[i, :4] = _limbs(p[0] * MONT_R % Q); out[i, 4:] = _limbs(p[1] * MONT_R % Q)
    return out


def unpack_point(a8):
    x = _unlimbs(a8[:4]); y = _unlimbs(a8[4:])
    if x == 0 and y == 0: return None
    ri = pow(MONT_R, Q - 2, Q)
    return (x * ri % Q, y * ri % Q)


def pack_scalars(scalars):
    out = np.zeros((len(scalars), 4), dtype=np.uint64)
    for i, s in enumerate(scalars):
        out[i] = _limbs(s % R)
    return out


def msm_c(bases8, scalars4, naive=False):
    b = np.ascontiguousarray(bases8, dtype=np.uint64); s = np.ascontiguousarray(scalars4, dtype=np.uint64)
    out = np.zeros(8, dtype=np.uint64)
    f = lib().bn_msm_naive if naive else lib().bn_msm
    f(b.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(b.shape[0]), out.ctypes.data_as(ctypes.c_void_p))
    return out


def on_curve_c(a8):
    a = np.ascontiguousarray(a8, dtype=np.uint64)
    return bool(lib().bn_is_on_curve(a.ctypes.data_as(ctypes.c_void_p)))
