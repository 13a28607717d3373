"""csrc/field.cuh (Goldilocks carry-chain primitives) compiled for the HOST against python ints: canonical results are
exact, weak results are congruent; inputs include the weak range [p, 2^64) wherever the contract allows it."""
import ctypes, itertools, os, random, subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 0xFFFFFFFF00000001
M64 = (1 << 64) - 1
U64 = ctypes.c_uint64


@pytest.fixture(scope="module", params=[0, 1], ids=["carry-chain products", "plain IMAD products"])
def L(request):
    # both formulations of the 64x64 product (GL_PLAIN_IMAD, field.cuh) must give the same numbers
    out = os.path.join(ROOT, "tests", "_build"); os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libfield_host_%d.so" % request.param)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DGL_PLAIN_IMAD=%d" % request.param,
                           "-I", os.path.join(ROOT, "eigen_zkvm_b200", "csrc"), "-I", "/usr/local/cuda/include",
                           "-o", so, os.path.join(ROOT, "tests", "field_host.cpp")])
    lib = ctypes.CDLL(so)
    for n in ["t_gl_add", "t_gl_sub", "t_gl_addw", "t_gl_mul", "t_gl_mulw", "t_gl_red128", "t_gl_red128w"]:
        getattr(lib, n).restype = U64; getattr(lib, n).argtypes = [U64, U64]
    lib.t_gl_maddw.restype = U64; lib.t_gl_maddw.argtypes = [U64, U64, U64]
    for n in ["t_gl_red96", "t_gl_red96w"]:
        getattr(lib, n).restype = U64; getattr(lib, n).argtypes = [U64, ctypes.c_uint32]
    for n in ["t_gl_canon", "t_gl_inv"]:
        getattr(lib, n).restype = U64; getattr(lib, n).argtypes = [U64]
    lib.t_gl_mulwide_add.argtypes = [U64, U64, U64, ctypes.POINTER(U64), ctypes.POINTER(U64)]
    return lib


EDGE_CANON = [0, 1, 2, 0xFFFFFFFF, 0x100000000, 0x100000001, 0xFFFFFFFE00000001, P - 2, P - 1, 1 << 63, (1 << 63) - 1, 0xFFFFFFFF00000000]
EDGE_WEAK = EDGE_CANON + [P, P + 1, M64 - 1, M64, 0xFFFFFFFF80000000]


def _pairs(edge, rnd, n, hi=M64):
    return list(itertools.product(edge, edge)) + [(rnd.randrange(hi + 1), rnd.randrange(hi + 1)) for _ in range(n)]


def test_add_sub(L):
    rnd = random.Random(1)
    for a, b in _pairs(EDGE_CANON, rnd, 20000, P - 1):
        assert L.t_gl_add(a, b) == (a + b) % P
        assert L.t_gl_sub(a, b) == (a - b) % P
    for a, b in _pairs(EDGE_WEAK, rnd, 20000):
        if b <= P:                      # contract: a weak, b <= p
            r = L.t_gl_sub(a, b); assert r % P == (a - b) % P
        if b < P:                       # contract: a weak, b canonical
            r = L.t_gl_addw(a, b); assert r % P == (a + b) % P
        assert L.t_gl_canon(a) == a % P


def test_mul_and_reductions(L):
    rnd = random.Random(2)
    lo = U64(); hi = U64()
    for a, b in _pairs(EDGE_WEAK, rnd, 30000):
        c = rnd.choice(EDGE_WEAK) if rnd.random() < 0.3 else rnd.randrange(1 << 64)
        L.t_gl_mulwide_add(a, b, c, ctypes.byref(lo), ctypes.byref(hi))
        assert lo.value | (hi.value << 64) == a * b + c
        assert L.t_gl_mul(a, b) == a * b % P
        assert L.t_gl_mulw(a, b) % P == a * b % P
        assert L.t_gl_maddw(a, b, c) % P == (a * b + c) % P
        # arbitrary 128-bit and 96-bit inputs
        assert L.t_gl_red128(a, b) == (a | (b << 64)) % P
        assert L.t_gl_red128w(a, b) % P == (a | (b << 64)) % P
        h32 = b & 0xFFFFFFFF
        assert L.t_gl_red96(a, h32) == (a | (h32 << 64)) % P
        assert L.t_gl_red96w(a, h32) % P == (a | (h32 << 64)) % P


def test_inv_and_f3(L):
    rnd = random.Random(3)
    for a in EDGE_CANON[1:] + [rnd.randrange(1, P) for _ in range(50)]:
        assert L.t_gl_inv(a) == pow(a, P - 2, P)
    # f3g.rs:619-624 KAT: (1,2,3) * (4,5,p-1) = (17,23,18)
    A = (U64 * 3)(1, 2, 3); B = (U64 * 3)(4, 5, P - 1); O = (U64 * 3)()
    L.t_f3_mul(A, B, O)
    assert list(O) == [17, 23, 18]
    # the lazy schoolbook product (three reductions for ten multiply-accumulates) against GL[x]/(x^3 - x - 1) in python ints, incl. the
    # largest canonical operands (every unreduced sum at its bound)
    def ref(a, b):
        c = [0] * 5
        for i in range(3):
            for j in range(3): c[i + j] += a[i] * b[j]
        # x^3 = x + 1, x^4 = x^2 + x
        return [(c[0] + c[3]) % P, (c[1] + c[3] + c[4]) % P, (c[2] + c[4]) % P]
    cases = [([P - 1] * 3, [P - 1] * 3), ([0, 0, P - 1], [0, 0, P - 1]), ([P - 1, 0, 0], [0, P - 1, P - 1]), ([0] * 3, [5, 6, 7])]
    cases += [([rnd.randrange(P) for _ in range(3)], [rnd.randrange(P) for _ in range(3)]) for _ in range(500)]
    for a, b in cases:
        A = (U64 * 3)(*a); B = (U64 * 3)(*b)
        L.t_f3_mul(A, B, O)
        assert list(O) == ref(a, b), (a, b)
