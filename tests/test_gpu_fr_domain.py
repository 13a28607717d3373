"""GPU parity of the scalar-field domain (fft / ifft / coset_fft / icoset_fft) and of the groth16 quotient H through the
C-ABI, against oracle/fr_domain.py (pinned by definition in tests/test_oracle_fr_domain.py)."""
import random
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
R256 = 1 << 256


@pytest.fixture(scope="module")
def g16():
    import __graft_entry__ as g
    g.build()
    from eigen_zkvm_b200 import groth16
    return groth16


def _mont(vals, p):
    return np.array([[((v * R256 % p) >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for v in vals], dtype=np.uint64).reshape(len(vals), 4)


def _unmont(a, p):
    ri = pow(R256, -1, p)
    return [sum(int(r[i]) << (64 * i) for i in range(4)) * ri % p for r in np.asarray(a).reshape(-1, 4)]


def _canon(a):
    return [sum(int(r[i]) << (64 * i) for i in range(4)) for r in np.asarray(a).reshape(-1, 4)]


@pytest.mark.parametrize("field,fid", [("bn254", 0), ("bls12381", 1)])
def test_transforms_against_oracle(g16, field, fid):
    from oracle import fr_domain as fd
    p = fd.MOD[field]; rnd = random.Random(fid)
    for lg in (0, 1, 2, 3, 8, 9, 10, 11, 12, 13, 14, 16):
        a = [rnd.randrange(p) for _ in range(1 << lg)]
        if lg >= 2: a[0] = 0; a[1] = p - 1
        am = _mont(a, p)
        assert _unmont(g16.fr_fft(am, fid, g16.FFT), p) == fd.fft(field, a), (field, lg)
        assert _unmont(g16.fr_fft(am, fid, g16.IFFT), p) == fd.ifft(field, a)
        assert _unmont(g16.fr_fft(am, fid, g16.COSET_FFT), p) == fd.coset_fft(field, a)
        assert _unmont(g16.fr_fft(am, fid, g16.ICOSET_FFT), p) == fd.icoset_fft(field, a)
    with pytest.raises(ValueError):
        g16.fr_fft(np.zeros((3, 4), dtype=np.uint64), fid)


@pytest.mark.parametrize("field,fid", [("bn254", 0), ("bls12381", 1)])
def test_h_against_oracle(g16, field, fid):
    from oracle import fr_domain as fd
    p = fd.MOD[field]; rnd = random.Random(10 + fid)
    for lg in (0, 1, 5, 10, 11, 13):
        m = 1 << lg
        a = [rnd.randrange(p) for _ in range(m)]; b = [rnd.randrange(p) for _ in range(m)]
        c = [x * y % p for x, y in zip(a, b)] if lg != 5 else [rnd.randrange(p) for _ in range(m)]     # satisfied and unsatisfied witnesses
        h = g16.groth16_h(_mont(a, p), _mont(b, p), _mont(c, p), fid)
        assert _canon(h) == fd.groth16_h(field, a, b, c), (field, lg)


@pytest.mark.parametrize("field,fid", [("bn254", 0), ("bls12381", 1)])
def test_large_size_properties(g16, field, fid):
    """2^20 points (the BN128 final layer runs at 2^22; the oracle is too slow there): round trips, and the quotient identity
    A(x) B(x) - C(x) = H(x) (x^m - 1) at a random point, with A, B, C interpolated by the device ifft."""
    from oracle import fr_domain as fd
    p = fd.MOD[field]; lg = 20; m = 1 << lg
    rng = np.random.default_rng(7 + fid)
    def rand_mont():
        x = rng.integers(0, 2**63, size=(m, 4), dtype=np.uint64); x[:, 3] &= np.uint64((1 << 60) - 1)      # < 2^252 < r: valid Montgomery limbs
        return x
    am, bm = rand_mont(), rand_mont()
    back = g16.fr_fft(g16.fr_fft(am, fid, g16.FFT), fid, g16.IFFT)
    assert (back == am).all()
    back = g16.fr_fft(g16.fr_fft(am, fid, g16.COSET_FFT), fid, g16.ICOSET_FFT)
    assert (back == am).all()
    # a satisfied witness: c_i = a_i * b_i on the domain, so that Z divides A * B - C
    av, bv = _unmont(am, p), _unmont(bm, p)
    cm = _mont([x * y % p for x, y in zip(av, bv)], p)
    h = _canon(g16.groth16_h(am, bm, cm, fid))
    A = _unmont(g16.fr_fft(am, fid, g16.IFFT), p); B = _unmont(g16.fr_fft(bm, fid, g16.IFFT), p); Cc = _unmont(g16.fr_fft(cm, fid, g16.IFFT), p)
    x0 = 0x1234567890ABCDEF1234567890ABCDEF % p
    def ev(poly):
        acc = 0
        for cf in reversed(poly): acc = (acc * x0 + cf) % p
        return acc
    eA, eB, eC, eH = ev(A), ev(B), ev(Cc), ev(h)
    assert (eA * eB - eC) % p == eH * ((pow(x0, m, p) - 1) % p) % p
