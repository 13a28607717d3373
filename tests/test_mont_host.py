"""csrc/mont.cuh compiled for the HOST (carry flag emulated) against python big ints: the same source runs on the GPU."""
import ctypes, os, random, subprocess, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gen_curve_params import FIELDS  # noqa: E402

NAMES = list(FIELDS)


@pytest.fixture(scope="module")
def L():
    out = os.path.join(ROOT, "tests", "_build"); os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libmont_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "eigen_zkvm_b200", "csrc"),
                           "-o", so, os.path.join(ROOT, "tests", "mont_host.cpp")])
    return ctypes.CDLL(so)


def tolimbs(x, n): return (ctypes.c_uint32 * n)(*[(x >> (32 * i)) & 0xffffffff for i in range(n)])
def fromlimbs(a): return sum(int(v) << (32 * i) for i, v in enumerate(a))


@pytest.mark.parametrize("fi", range(4))
def test_fp(L, fi):
    random.seed(fi)
    p = FIELDS[NAMES[fi]]; n = (p.bit_length() + 31) // 32; n += n & 1; R = 1 << (32 * n); Ri = pow(R, -1, p)
    edge = [0, 1, 2, p - 1, p - 2, R % p, (1 << (p.bit_length() - 1))]
    cases = [(a, b) for a in edge for b in edge] + [(random.randrange(p), random.randrange(p)) for _ in range(1500)]
    o = (ctypes.c_uint32 * n)()
    for it, (a, b) in enumerate(cases):
        A, B = tolimbs(a, n), tolimbs(b, n)
        L.fp_op(fi, 0, A, B, o); assert fromlimbs(o) == a * b * Ri % p
        L.fp_op(fi, 1, A, B, o); assert fromlimbs(o) == (a + b) % p
        L.fp_op(fi, 2, A, B, o); assert fromlimbs(o) == (a - b) % p
        L.fp_op(fi, 6, A, B, o); assert fromlimbs(o) == (-a) % p
        L.fp_op(fi, 4, A, B, o); assert fromlimbs(o) == a * R % p
        L.fp_op(fi, 5, A, B, o); assert fromlimbs(o) == a * Ri % p
        if it % 100 == 0:
            L.fp_op(fi, 3, A, B, o)
            am = a * Ri % p
            assert fromlimbs(o) == ((pow(am, -1, p) * R) % p if am else 0)


@pytest.mark.parametrize("fi", [0, 2])
def test_fp2(L, fi):
    random.seed(10 + fi)
    p = FIELDS[NAMES[fi]]; n = (p.bit_length() + 31) // 32; n += n & 1; R = 1 << (32 * n); Ri = pow(R, -1, p)
    o = (ctypes.c_uint32 * (2 * n))(); o2 = (ctypes.c_uint32 * (2 * n))()
    split = lambda r: (r & (R - 1), r >> (32 * n))
    for it in range(300):
        a0, a1, b0, b1 = [random.randrange(p) for _ in range(4)]
        if it == 0: a1 = b1 = 0
        if it == 1: a0 = b0 = 0
        A = tolimbs(a0 | (a1 << (32 * n)), 2 * n); B = tolimbs(b0 | (b1 << (32 * n)), 2 * n)
        L.fp2_op(fi, 0, A, B, o); assert split(fromlimbs(o)) == ((a0 * b0 - a1 * b1) * Ri % p, (a0 * b1 + a1 * b0) * Ri % p)
        L.fp2_op(fi, 4, A, B, o); assert split(fromlimbs(o)) == ((a0 * a0 - a1 * a1) * Ri % p, (2 * a0 * a1) * Ri % p)
        L.fp2_op(fi, 1, A, B, o); assert split(fromlimbs(o)) == ((a0 + b0) % p, (a1 + b1) % p)
        L.fp2_op(fi, 2, A, B, o); assert split(fromlimbs(o)) == ((a0 - b0) % p, (a1 - b1) % p)
        if it < 6 and (a0 or a1):
            L.fp2_op(fi, 3, A, B, o)
            L.fp2_op(fi, 0, A, o, o2); assert split(fromlimbs(o2)) == (R % p, 0)


@pytest.mark.parametrize("fi", range(4))
def test_fp_lazy_dot(L, fi):
    """FpWide: sum of up to 17 unreduced products (+ addend), one Montgomery reduction, LOGK conditional subtractions."""
    random.seed(40 + fi)
    p = FIELDS[NAMES[fi]]; n = (p.bit_length() + 31) // 32; n += n & 1; R = 1 << (32 * n); Ri = pow(R, -1, p)
    o = (ctypes.c_uint32 * n)()
    def pack(xs): return (ctypes.c_uint32 * (n * len(xs)))(*[(x >> (32 * i)) & 0xffffffff for x in xs for i in range(n)])
    for it in range(400):
        terms = random.choice([1, 2, 3, 5, 16, 17])
        if it < 40:      # extremes: all operands p - 1 (the bound), zeros
            a = [p - 1] * terms; b = [p - 1] * terms; c = p - 1 if it % 2 else None
            if it >= 20: a = [0] * terms
        else:
            a = [random.randrange(p) for _ in range(terms)]; b = [random.randrange(p) for _ in range(terms)]
            c = random.randrange(p) if it % 3 == 0 else None
        # the reduced value is < p (terms p / R + 2): pick the smallest admissible LOGK, and also a larger one
        bound = terms * p / R + 2
        logk = 1
        while (1 << logk) < bound: logk += 1
        want = (sum(x * y for x, y in zip(a, b)) * Ri + (c or 0)) % p
        for lk in {logk, 4}:
            if lk > 4: continue
            L.fp_dot(fi, lk, terms, pack(a), pack(b), tolimbs(c, n) if c is not None else None, o)
            assert fromlimbs(o) == want, (fi, it, terms, lk)
