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


class MsmTable:
    """Device-resident bases of one multiexp of a circuit plus their shifted copies 2^(c w) P (b200_msm_table_new): built once per
    proving key, every later `run` moves only the scalars."""
    def __init__(self, bases=None, curve=BN254_G1, device_ptr=None, n=None):
        self.curve = curve; self._h = ctypes.c_void_p()
        if device_ptr is not None:
            _lib.check(_lib.lib().b200_msm_table_new(curve, ctypes.c_void_p(device_ptr), n, 1, ctypes.byref(self._h)))
        else:
            b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, point_words(curve))
            _lib.check(_lib.lib().b200_msm_table_new(curve, b.ctypes.data_as(ctypes.c_void_p), b.shape[0], 0, ctypes.byref(self._h)))
        c = ctypes.c_uint(); w = ctypes.c_uint(); nn = ctypes.c_size_t()
        _lib.check(_lib.lib().b200_msm_table_info(self._h, ctypes.byref(c), ctypes.byref(w), ctypes.byref(nn)))
        self.window_bits, self.windows, self.n = c.value, w.value, nn.value

    def run(self, scalars4):
        s = np.ascontiguousarray(scalars4, dtype=np.uint64).reshape(-1, 4)
        if s.shape[0] != self.n:
            raise ValueError("bases and exponents differ in length")
        out = np.zeros(point_words(self.curve) * 3 // 2, dtype=np.uint64)
        _lib.check(_lib.lib().b200_msm_table_run(self._h, s.ctypes.data_as(ctypes.c_void_p), 0, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def run_dev(self, d_scalars_ptr):
        out = np.zeros(point_words(self.curve) * 3 // 2, dtype=np.uint64)
        _lib.check(_lib.lib().b200_msm_table_run(self._h, ctypes.c_void_p(d_scalars_ptr), 1, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def set_partial_output(self, on=True):
        """un-normalised results (no field inversion): for per-GPU partial sums that points_sum_dev combines"""
        _lib.check(_lib.lib().b200_msm_table_set_partial_output(self._h, 1 if on else 0))

    def free(self):
        if self._h:
            _lib.lib().b200_msm_table_free(self._h); self._h = ctypes.c_void_p()

    def __del__(self):
        try: self.free()
        except Exception: pass


def points_sum_dev(d_points_ptr, count, curve=BN254_G1):
    """sum of `count` Jacobian triples in device memory (the all-gathered per-GPU partial sums)"""
    out = np.zeros(point_words(curve) * 3 // 2, dtype=np.uint64)
    _lib.check(_lib.lib().b200_points_sum_dev(curve, ctypes.c_void_p(d_points_ptr), count, out.ctypes.data_as(ctypes.c_void_p)))
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


# ---- `Groth16::prove` in one call ----------------------------------------------------------------------------------------
BN128, BLS12381 = 0, 1
FR_MODULUS = {BN128: 21888242871839275222246405745257275088548364400416034343698204186575808495617,
              BLS12381: 52435875175126190479447740508185965837690552500527637822603658699938581184513}


def _canon_words(vals):
    return np.array([[(int(v) >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for v in vals], dtype=np.uint64).reshape(len(vals), 4)


def fr_to_mont_words(vals, curve=BN128):
    """python ints -> the in-memory `Fr` (4 x u64 Montgomery limbs, R = 2^256)"""
    p = FR_MODULUS[curve]
    return _canon_words([int(v) % p * (1 << 256) % p for v in vals])


def read_wtns(data, curve=BN128):
    """`load_witness_from_bin_reader` (algebraic/src/reader.rs:87-138): the witness as python ints"""
    buf = (ctypes.c_char * len(data)).from_buffer_copy(data)
    n = ctypes.c_size_t()
    _lib.check(_lib.lib().b200_wtns_read(buf, len(data), curve, None, 0, ctypes.byref(n)))
    out = np.zeros((n.value, 4), dtype=np.uint64)
    _lib.check(_lib.lib().b200_wtns_read(buf, len(data), curve, out.ctypes.data_as(ctypes.c_void_p), n.value, ctypes.byref(n)))
    return [sum(int(out[i, k]) << (64 * k) for k in range(4)) for i in range(n.value)]


class Parameters:
    """`Parameters<E>` read with `Parameters::read(reader, false)` (groth16/src/api.rs:161,545-550) and kept on the device."""
    def __init__(self, data, curve=BN128):
        self.curve = curve; self._h = ctypes.c_void_p()
        buf = (ctypes.c_char * len(data)).from_buffer_copy(data)
        _lib.check(_lib.lib().b200_groth16_pk_read(curve, buf, len(data), ctypes.byref(self._h)))
        cnt = (ctypes.c_size_t * 6)()
        _lib.check(_lib.lib().b200_groth16_pk_info(self._h, cnt))
        self.n_h, self.n_l, self.n_a, self.n_b_g1, self.n_b_g2, self.n_ic = [int(x) for x in cnt]

    def free(self):
        if self._h:
            _lib.lib().b200_groth16_pk_free(self._h); self._h = ctypes.c_void_p()

    def __del__(self):
        try: self.free()
        except Exception: pass


def prove(params, a, b, c, inputs, aux, a_aux_density, b_input_density, b_aux_density, r, s):
    """`Groth16::prove` after synthesis (see include/b200zk.h).  a, b, c, inputs, aux: python ints (canonical values); densities: bools.
    Returns (A, B, C) as word arrays in the in-memory Montgomery form: A, C = (2 * w,) and B = (4 * w,) uint64, w = 4 or 6."""
    cv = params.curve
    w = point_words(BN254_G1 if cv == BN128 else BLS12381_G1) // 2
    am, bm, cm = fr_to_mont_words(a, cv), fr_to_mont_words(b, cv), fr_to_mont_words(c, cv)
    iw, xw = _canon_words(inputs), _canon_words(aux)
    da = np.array(a_aux_density, dtype=np.uint8); dbi = np.array(b_input_density, dtype=np.uint8); dba = np.array(b_aux_density, dtype=np.uint8)
    rw, sw = _canon_words([r])[0], _canon_words([s])[0]
    out = np.zeros(8 * w, dtype=np.uint64)
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    _lib.check(_lib.lib().b200_groth16_prove(params._h, p(am), p(bm), p(cm), len(a), p(iw), len(inputs), p(xw), len(aux), p(da), p(dbi), p(dba), p(rw), p(sw), p(out)))
    return out[:2 * w].copy(), out[2 * w:6 * w].copy(), out[6 * w:].copy()
