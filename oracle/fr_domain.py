"""Scalar-field evaluation domain and the groth16 `H` computation -- TEST INFRASTRUCTURE (python big ints).

CPU restatement of bellman_ce's `domain::EvaluationDomain` (fft / ifft / coset_fft / icoset_fft / mul_assign /
sub_assign / divide_by_z_on_coset) and of the quotient computation inside `groth16::create_random_proof`
(a = ifft(a).coset_fft(), same for b, c; a = (a * b - c) / Z on the coset; h = icoset_fft(a) without its last
coefficient), reached from `Groth16::prove` (groth16/src/groth16.rs:88-96; bellperson twin :45-57).
PARITY UNPINNED by the reference: bellman_ce 0.3.2 / bellperson 0.26 are un-vendored (Cargo.lock:668-670,731-733) and no
reference test fixes these values.  Pinned here by definition instead: the transforms are compared with the direct O(n^2)
evaluation of the DFT they implement, and the constants by their defining properties (tests/test_oracle_fr_domain.py):
  r - 1 = 2^S * t with S = 28 (BN254 Fr, starky/src/field_bn128.rs:12) / S = 32 (BLS12-381 Fr, field_bls12381.rs:12);
  multiplicative generator g = 7 (a quadratic non-residue in both fields), ROOT_OF_UNITY = g^t has order exactly 2^S;
  omega_m = ROOT_OF_UNITY^(2^(S - log2 m)); cosets are g * <omega>.
"""

MOD = {"bn254": 21888242871839275222246405745257275088548364400416034343698204186575808495617,
       "bls12381": 52435875175126190479447740508185965837690552500527637822603658699938581184513}
S = {"bn254": 28, "bls12381": 32}
GENERATOR = 7


def root_of_unity(field):
    r = MOD[field]
    return pow(GENERATOR, (r - 1) >> S[field], r)


def omega(field, log_m):
    if log_m > S[field]:
        raise ValueError("PolynomialDegreeTooLarge")
    return pow(root_of_unity(field), 1 << (S[field] - log_m), MOD[field])


def _bitrev(a):
    n = len(a); lg = n.bit_length() - 1
    out = list(a)
    for k in range(n):
        rk = int(format(k, "0%db" % lg)[::-1], 2) if lg else 0
        if k < rk: out[k], out[rk] = out[rk], out[k]
    return out


def _fft(a, w, p):
    """serial_fft of bellman: bit reversal + iterative butterflies; a[k] <- sum_j a[j] w^(jk)."""
    a = _bitrev(a); n = len(a); m = 1
    while m < n:
        wm = pow(w, n // (2 * m), p)
        for k in range(0, n, 2 * m):
            ww = 1
            for j in range(m):
                t = a[k + j + m] * ww % p
                a[k + j + m] = (a[k + j] - t) % p
                a[k + j] = (a[k + j] + t) % p
                ww = ww * wm % p
        m *= 2
    return a


def fft(field, a):
    lg = len(a).bit_length() - 1
    return _fft(a, omega(field, lg), MOD[field])


def ifft(field, a):
    p = MOD[field]; lg = len(a).bit_length() - 1
    minv = pow(len(a), p - 2, p)
    return [x * minv % p for x in _fft(a, pow(omega(field, lg), p - 2, p), p)]


def distribute_powers(field, a, g):
    p = MOD[field]; out = []; u = 1
    for x in a:
        out.append(x * u % p); u = u * g % p
    return out


def coset_fft(field, a):
    return fft(field, distribute_powers(field, a, GENERATOR))


def icoset_fft(field, a):
    p = MOD[field]
    return distribute_powers(field, ifft(field, a), pow(GENERATOR, p - 2, p))


def dft_naive(field, a):
    p = MOD[field]; n = len(a); w = omega(field, n.bit_length() - 1)
    return [sum(a[j] * pow(w, j * k, p) for j in range(n)) % p for k in range(n)]


def groth16_h(field, a, b, c):
    """a, b, c: evaluations of the QAP polynomials on the domain (length m = 2^k).  Returns the m - 1 coefficients of
    H = (A * B - C) / Z that the prover feeds to the `h` multiexp."""
    p = MOD[field]; m = len(a)
    A = coset_fft(field, ifft(field, a)); B = coset_fft(field, ifft(field, b)); C = coset_fft(field, ifft(field, c))
    zinv = pow((pow(GENERATOR, m, p) - 1) % p, p - 2, p)          # divide_by_z_on_coset: Z(g w^i) = g^m - 1
    q = [((x * y - z) % p) * zinv % p for x, y, z in zip(A, B, C)]
    h = icoset_fft(field, q)
    return h[:m - 1]
