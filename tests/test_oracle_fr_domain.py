"""Pins oracle/fr_domain.py by definition (the reference does not vendor bellman's domain code): constants by their
defining properties, transforms against the direct DFT, and the H computation against polynomial arithmetic."""
import random
from oracle import fr_domain as fd


def test_constants():
    for f in ("bn254", "bls12381"):
        r = fd.MOD[f]; s = fd.S[f]
        assert (r - 1) % (1 << s) == 0 and ((r - 1) >> s) % 2 == 1          # r - 1 = 2^S * t, t odd
        assert pow(fd.GENERATOR, (r - 1) // 2, r) == r - 1                      # 7 is a quadratic non-residue
        w = fd.root_of_unity(f)
        assert pow(w, 1 << s, r) == 1 and pow(w, 1 << (s - 1), r) == r - 1      # order exactly 2^S
        assert fd.omega(f, 3) == pow(w, 1 << (s - 3), r)


def test_transforms_match_the_direct_dft():
    rnd = random.Random(1)
    for f in ("bn254", "bls12381"):
        p = fd.MOD[f]
        for lg in (0, 1, 2, 5, 7):
            a = [rnd.randrange(p) for _ in range(1 << lg)]
            assert fd.fft(f, a) == fd.dft_naive(f, a)
            assert fd.ifft(f, fd.fft(f, a)) == a
            assert fd.icoset_fft(f, fd.coset_fft(f, a)) == a
            # coset_fft evaluates the polynomial on g * w^k
            if lg <= 5:
                w = fd.omega(f, lg)
                ev = [sum(c * pow(fd.GENERATOR * pow(w, k, p) % p, j, p) for j, c in enumerate(a)) % p for k in range(1 << lg)]
                assert fd.coset_fft(f, a) == ev


def test_h_is_the_quotient_polynomial():
    # pick A, B of degree < m, set C = A * B mod Z + a multiple trick: choose H freely, C := A*B - H*Z evaluated on the domain
    rnd = random.Random(2)
    for f in ("bn254", "bls12381"):
        p = fd.MOD[f]; lg = 5; m = 1 << lg; w = fd.omega(f, lg)
        A = [rnd.randrange(p) for _ in range(m)]; B = [rnd.randrange(p) for _ in range(m)]
        H = [rnd.randrange(p) for _ in range(m - 1)]
        def ev(poly, x): 
            acc = 0
            for c in reversed(poly): acc = (acc * x + c) % p
            return acc
        pts = [pow(w, k, p) for k in range(m)]
        a = [ev(A, x) for x in pts]; b = [ev(B, x) for x in pts]
        # on the domain Z vanishes, so c = a * b there; A*B - C = H*Z requires C = A*B - H*Z as polynomials, deg C may reach 2m-2;
        # bellman only ever sees c's evaluations on the domain, i.e. C mod Z.  Build C_low = (A*B - H*Z) mod (x^m - 1) by evaluation.
        c = [x * y % p for x, y in zip(a, b)]
        # with these inputs (C := interpolation of a*b) the quotient is H' = (A*B - C)/Z; check the defining identity on a fresh point
        h = fd.groth16_h(f, a, b, c)
        C = fd.ifft(f, c)
        x0 = rnd.randrange(2, p)
        lhs = (ev(A, x0) * ev(B, x0) - ev(C, x0)) % p
        rhs = ev(h, x0) * ((pow(x0, m, p) - 1) % p) % p
        assert lhs == rhs and len(h) == m - 1
