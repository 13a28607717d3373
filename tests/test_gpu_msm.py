"""GPU parity for the BN254 G1 MSM through the C-ABI, against the CPU oracle (python big-int + C Pippenger).
Parity is at group-element level (the prover's proofs are randomised: groth16/src/api.rs:154,173)."""
import random
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g16():
    import __graft_entry__ as g
    g.build()
    from eigen_zkvm_b200 import groth16
    return groth16


def _scalars(n, seed):
    from oracle import bn254 as bn
    rnd = random.Random(seed)
    return [rnd.randrange(bn.R) for _ in range(n)]


def test_small_cases_against_python_bigint(g16):
    from oracle import bn254 as bn
    rnd = random.Random(9)
    G = bn.G1
    # [k]G for small k and edge scalars
    for k in [0, 1, 2, 3, 5, 255, 256, 65535, 65536, 2**128 + 1, bn.R - 1]:
        out = g16.multiexp(bn.pack_points([G]), bn.pack_scalars([k]))
        assert bn.unpack_point(g16.jacobian_to_affine_mont(out)) == bn.mul(k % bn.R, G), k
    for n in (2, 7, 33, 200):
        pts = [bn.mul(rnd.randrange(1, bn.R), G) for _ in range(n)]
        sc = [rnd.randrange(bn.R) for _ in range(n)]
        if n >= 7:
            pts[3] = None; sc[4] = 0; sc[5] = bn.R - 1; pts[6] = pts[2]
        out = g16.multiexp(bn.pack_points(pts), bn.pack_scalars(sc))
        assert bn.unpack_point(g16.jacobian_to_affine_mont(out)) == bn.msm_naive(pts, sc)
    # all equal points and scalars (exercises the doubling branch inside buckets), cancellation to infinity
    p = bn.mul(777, G)
    out = g16.multiexp(bn.pack_points([p] * 64), bn.pack_scalars([3] * 64))
    assert bn.unpack_point(g16.jacobian_to_affine_mont(out)) == bn.mul(192, p)
    out = g16.multiexp(bn.pack_points([p, p]), bn.pack_scalars([5, bn.R - 5]))
    assert bn.unpack_point(g16.jacobian_to_affine_mont(out)) is None
    assert bn.unpack_point(g16.jacobian_to_affine_mont(g16.multiexp(np.zeros((0, 8), dtype=np.uint64), np.zeros((0, 4), dtype=np.uint64)))) is None
    with pytest.raises(ValueError):
        g16.multiexp(bn.pack_points([p, p]), bn.pack_scalars([5]))


@pytest.mark.parametrize("logn", [10, 13, 16, 18])
def test_random_points_against_c_pippenger(g16, logn):
    import torch
    from oracle import bn254 as bn
    n = 1 << logn
    d_b = torch.empty(n * 8, dtype=torch.int64, device="cuda")
    g16.random_points_dev(d_b.data_ptr(), n, 0xB254)
    bases = d_b.cpu().numpy().view(np.uint64).reshape(n, 8)
    for i in (0, 1, n // 2, n - 1):
        assert bn.on_curve_c(bases[i])
    rng = np.random.default_rng(logn)
    sc = rng.integers(0, 2**63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    sc[:, 3] &= np.uint64((1 << 61) - 1)           # < 2^253 < r: canonical
    out = g16.multiexp(bases, sc)
    assert (g16.jacobian_to_affine_mont(out) == bn.msm_c(bases, sc)).all()
    d_s = torch.from_numpy(sc.view(np.int64)).cuda()
    assert (g16.multiexp_dev(d_b.data_ptr(), d_s.data_ptr(), n) == out).all()


def test_large_linearity_and_chunk_combination(g16):
    # size-independent properties at BASELINE's n = 2^22: MSM(s) + MSM(t) = MSM(s + t) and chunked partial sums combine
    import torch
    from oracle import bn254 as bn
    n = 1 << 22
    d_b = torch.empty(n * 8, dtype=torch.int64, device="cuda")
    g16.random_points_dev(d_b.data_ptr(), n, 0xB254)
    rng = np.random.default_rng(1)
    def rs():
        a = rng.integers(0, 2**63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 59) - 1)        # < 2^251 so that s + t < r without reduction
        return a
    s, t = rs(), rs()
    st = np.zeros_like(s); carry = np.zeros(n, dtype=np.uint64)
    for l in range(4):
        x = s[:, l] + t[:, l]; c1 = x < s[:, l]; y = x + carry; c2 = y < x
        st[:, l] = y; carry = (c1 | c2).astype(np.uint64)
    up = lambda a: torch.from_numpy(a.view(np.int64)).cuda()
    ds, dt, dst = up(s), up(t), up(st)
    ms = g16.multiexp_dev(d_b.data_ptr(), ds.data_ptr(), n); mt = g16.multiexp_dev(d_b.data_ptr(), dt.data_ptr(), n)
    mst = g16.multiexp_dev(d_b.data_ptr(), dst.data_ptr(), n)
    assert (g16.g1_add(ms, mt) == mst).all()
    # 8 chunks (the multi-GPU split) combine to the same point
    acc = None
    for k in range(8):
        lo = k * (n // 8)
        part = g16.multiexp_dev(d_b.data_ptr() + lo * 64, dst.data_ptr() + lo * 32, n // 8)
        acc = part if acc is None else g16.g1_add(acc, part)
    assert (acc == mst).all()
    assert bn.on_curve_c(g16.jacobian_to_affine_mont(mst))


# ---- BN254 G2, BLS12-381 G1 / G2 (the groth16 `B_g2` query and the BLS12-381 final layer) against oracle/curves.py ----
CURVE_IDS = {"bn254_g1": 0, "bn254_g2": 1, "bls12381_g1": 2, "bls12381_g2": 3}


def _pack(c, pts):
    return np.array([c.affine_to_words(p) for p in pts], dtype=np.uint64).reshape(len(pts), -1)


def _scal(sc):
    return np.array([[(s >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for s in sc], dtype=np.uint64).reshape(len(sc), 4)


@pytest.mark.parametrize("name", ["bn254_g1", "bn254_g2", "bls12381_g1", "bls12381_g2"])
def test_all_curves_small_cases_against_python_bigint(g16, name):
    from oracle import curves as C
    c = C.CURVES[name]; cid = CURVE_IDS[name]
    assert g16.point_words(cid) == 2 * c.f_words
    rnd = random.Random(11)
    G = c.gen
    for k in [0, 1, 2, 3, 65535, 65536, 2**128 + 1, c.r - 1]:
        out = g16.multiexp(_pack(c, [G]), _scal([k]), cid)
        assert c.jacobian_from_words(out) == c.mul(k % c.r, G), k
    for n in (2, 7, 40):
        pts = [c.mul(rnd.randrange(1, 1 << 70), G) for _ in range(n)]
        sc = [rnd.randrange(c.r) for _ in range(n)]
        if n >= 7:
            pts[3] = None; sc[4] = 0; sc[5] = c.r - 1; pts[6] = pts[2]
        out = g16.multiexp(_pack(c, pts), _scal(sc), cid)
        assert c.jacobian_from_words(out) == c.msm_naive(pts, sc)
    p = c.mul(777, G)
    assert c.jacobian_from_words(g16.multiexp(_pack(c, [p] * 64), _scal([3] * 64), cid)) == c.mul(192, p)       # doubling branch in a bucket
    assert c.jacobian_from_words(g16.multiexp(_pack(c, [p, p]), _scal([5, c.r - 5]), cid)) is None               # cancellation
    assert c.jacobian_from_words(g16.multiexp(np.zeros((0, 2 * c.f_words), dtype=np.uint64), np.zeros((0, 4), dtype=np.uint64), cid)) is None
    # point_add on Jacobian triples, including infinity
    a = g16.multiexp(_pack(c, [G]), _scal([5]), cid); b = g16.multiexp(_pack(c, [G]), _scal([9]), cid)
    inf = g16.multiexp(_pack(c, [G]), _scal([0]), cid)
    assert c.jacobian_from_words(g16.point_add(a, b, cid)) == c.mul(14, G)
    assert c.jacobian_from_words(g16.point_add(a, inf, cid)) == c.mul(5, G)
    assert c.jacobian_from_words(g16.point_add(a, a, cid)) == c.mul(10, G)


@pytest.mark.parametrize("name,logn", [("bn254_g2", 12), ("bls12381_g1", 12), ("bls12381_g2", 11), ("bn254_g2", 18), ("bls12381_g1", 18), ("bls12381_g2", 16)])
def test_all_curves_random_points_and_linearity(g16, name, logn):
    """Device-generated points are on the curve; MSM(s) + MSM(t) = MSM(s + t); chunked partial sums combine; and for the
    small sizes the result equals a python big-int Pippenger-free sum over a sample (c = 12 window path)."""
    import torch
    from oracle import curves as C
    c = C.CURVES[name]; cid = CURVE_IDS[name]
    n = 1 << logn; pw = 2 * c.f_words
    d_b = torch.empty(n * pw, dtype=torch.int64, device="cuda")
    g16.random_points_dev(d_b.data_ptr(), n, 0xB254, cid)
    bases = d_b.cpu().numpy().view(np.uint64).reshape(n, pw)
    for i in (0, 1, n // 2, n - 1):
        P = c.affine_from_words(bases[i])
        assert P is not None and c.is_on_curve(P)
    rng = np.random.default_rng(logn)
    def rs():
        a = rng.integers(0, 2**63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 59) - 1)
        return a
    s, t = rs(), rs()
    st = np.zeros_like(s); carry = np.zeros(n, dtype=np.uint64)
    for l in range(4):
        x = s[:, l] + t[:, l]; c1 = x < s[:, l]; y = x + carry; c2 = y < x
        st[:, l] = y; carry = (c1 | c2).astype(np.uint64)
    up = lambda a: torch.from_numpy(a.view(np.int64)).cuda()
    ds, dt, dst = up(s), up(t), up(st)
    ms = g16.multiexp_dev(d_b.data_ptr(), ds.data_ptr(), n, cid); mt = g16.multiexp_dev(d_b.data_ptr(), dt.data_ptr(), n, cid)
    mst = g16.multiexp_dev(d_b.data_ptr(), dst.data_ptr(), n, cid)
    assert (g16.point_add(ms, mt, cid) == mst).all()
    assert c.is_on_curve(c.jacobian_from_words(mst))
    acc = None
    for k in range(4):
        lo = k * (n // 4)
        part = g16.multiexp_dev(d_b.data_ptr() + lo * pw * 8, dst.data_ptr() + lo * 32, n // 4, cid)
        acc = part if acc is None else g16.point_add(acc, part, cid)
    assert (acc == mst).all()
    assert (g16.multiexp(bases, st, cid) == mst).all()          # host-buffer entry point
    if logn <= 12:
        m = 96      # exact check on a prefix (python big ints)
        pts = [c.affine_from_words(bases[i]) for i in range(m)]
        sc = [sum(int(st[i, l]) << (64 * l) for l in range(4)) for i in range(m)]
        out = g16.multiexp(bases[:m], st[:m], cid)
        assert c.jacobian_from_words(out) == c.msm_naive(pts, sc)


def test_skewed_scalars_load_balanced_buckets(g16):
    """groth16 witnesses are mostly 0 / 1 / small values: buckets with 10^4..10^6 points must neither be slow nor wrong.
    All-equal scalars put every point into one bucket per window (the chunked path, ch = 256 here); compared with the C
    oracle, and timed against a uniform MSM of the same size (must stay within 3x)."""
    import time, torch
    from oracle import bn254 as bn
    n = 1 << 16
    d_b = torch.empty(n * 8, dtype=torch.int64, device="cuda")
    g16.random_points_dev(d_b.data_ptr(), n, 0x5EED)
    bases = d_b.cpu().numpy().view(np.uint64).reshape(n, 8)
    rng = np.random.default_rng(4)
    uniform = rng.integers(0, 2**63, size=(n, 4), dtype=np.uint64); uniform[:, 3] &= np.uint64((1 << 60) - 1)
    cases = {}
    ones = np.zeros((n, 4), dtype=np.uint64); ones[:, 0] = 1
    cases["all ones"] = ones
    same = np.tile(uniform[:1], (n, 1)); cases["all equal (one bucket per window)"] = same
    mixed = np.zeros((n, 4), dtype=np.uint64)
    kind = rng.integers(0, 10, size=n)
    mixed[kind < 4, 0] = 1                                   # 40 % ones, 30 % zeros, 20 % bytes, 10 % full width
    small = (kind >= 7) & (kind < 9); mixed[small, 0] = rng.integers(0, 256, size=int(small.sum()), dtype=np.uint64)
    mixed[kind == 9] = uniform[kind == 9]
    cases["witness-like mix"] = mixed
    g16.multiexp(bases, uniform)                             # warm-up (workspace growth)
    t0 = time.perf_counter(); g16.multiexp(bases, uniform); t_uniform = time.perf_counter() - t0
    for name, sc in cases.items():
        t0 = time.perf_counter(); out = g16.multiexp(bases, sc); dt = time.perf_counter() - t0
        assert (g16.jacobian_to_affine_mont(out) == bn.msm_c(bases, sc)).all(), name
        assert dt < 3 * t_uniform + 0.01, (name, dt, t_uniform)
    # the other groups: all-equal scalars against [k * n] applied to the sum of points (python big ints, small n)
    from oracle import curves as C
    for cname, cid in (("bn254_g2", 1), ("bls12381_g1", 2), ("bls12381_g2", 3)):
        c = C.CURVES[cname]; m = 600
        pts = [c.mul(3 + 7 * i, c.gen) for i in range(m)]
        k = 0x1234567890ABCDEF1234567
        s = c.gen; tot = None
        for p in pts: tot = c.add(tot, p)
        out = g16.multiexp(_pack(c, pts), _scal([k] * m), cid)
        assert c.jacobian_from_words(out) == c.mul(k, tot), cname


# ---- round 2: the C oracle for all four groups at 2^13 .. 2^16 points, and the per-circuit table mode ------------------------
@pytest.mark.parametrize("name,logn", [("bn254_g2", 14), ("bls12381_g1", 16), ("bls12381_g2", 13), ("bn254_g1", 16)])
def test_all_curves_against_the_c_oracle_at_scale(g16, name, logn):
    """beyond n = 600 the G2 / BLS12-381 MSMs were only checked by linearity, which cannot see an error shared by
    all windows.  oracle/curves_oracle.c (unsigned windows, Jacobian, 64-bit CIOS; itself pinned to python big ints) gives the
    exact point at 2^13 .. 2^16; the GPU must match it in BOTH modes: plain Pippenger and the shifted-base table."""
    import torch
    from oracle import curves as C
    c = C.CURVES[name]; cid = CURVE_IDS[name]
    n = 1 << logn; pw = 2 * c.f_words
    d_b = torch.empty(n * pw, dtype=torch.int64, device="cuda")
    g16.random_points_dev(d_b.data_ptr(), n, 0xC0FFEE, cid)
    bases = d_b.cpu().numpy().view(np.uint64).reshape(n, pw).copy()
    bases[5] = 0                                                     # a point at infinity among the bases
    d_b = torch.from_numpy(bases.view(np.int64).reshape(-1)).cuda()
    rng = np.random.default_rng(logn + 100)
    sc = rng.integers(0, 2**63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    sc[:, 3] &= np.uint64((1 << 60) - 1)           # < 2^252 < r for both curves
    sc[7] = 0; sc[8] = [1, 0, 0, 0]
    rm1 = c.r - 1; sc[9] = [(rm1 >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]          # the largest canonical scalar: top window + signed-digit carry
    want = C.msm_c(c, bases, sc)
    assert want.any() and c.is_on_curve(c.affine_from_words(want))
    d_s = torch.from_numpy(sc.view(np.int64).reshape(-1)).cuda()
    plain = g16.multiexp_dev(d_b.data_ptr(), d_s.data_ptr(), n, cid)
    assert (g16.jacobian_to_affine_mont(plain, cid) == want).all()
    tab = g16.MsmTable(device_ptr=d_b.data_ptr(), n=n, curve=cid)
    assert tab.n == n and tab.windows * tab.window_bits >= C.SCALAR_BITS[name] + 1
    assert (tab.run_dev(d_s.data_ptr()) == plain).all()
    assert (tab.run(sc) == plain).all()                              # host scalars
    assert (tab.run_dev(d_s.data_ptr()) == plain).all()              # reusable
    tab.free()


def test_table_mode_small_and_edge_cases(g16):
    from oracle import bn254 as bn
    rnd = random.Random(21)
    G = bn.G1
    for n in (1, 3, 50):
        pts = [bn.mul(rnd.randrange(1, bn.R), G) for _ in range(n)]
        sc = [rnd.randrange(bn.R) for _ in range(n)]
        if n >= 3:
            pts[1] = None; sc[2] = bn.R - 1
        tab = g16.MsmTable(bn.pack_points(pts))
        out = tab.run(bn.pack_scalars(sc))
        assert bn.unpack_point(g16.jacobian_to_affine_mont(out)) == bn.msm_naive(pts, sc)
        assert bn.unpack_point(g16.jacobian_to_affine_mont(tab.run(bn.pack_scalars([0] * n)))) is None
        with pytest.raises(ValueError):
            tab.run(bn.pack_scalars(sc + [1]))
    # a 2-torsion-free curve has no finite point whose 2^k multiple is infinity, but the all-zero (infinity) base must stay infinity in every window
    tab = g16.MsmTable(bn.pack_points([None, G]))
    assert bn.unpack_point(g16.jacobian_to_affine_mont(tab.run(bn.pack_scalars([bn.R - 1, 2])))) == bn.mul(2, G)


def test_table_mode_2_22_matches_plain_and_device_sum(g16):
    import torch
    n = 1 << 22
    d_b = torch.empty(n * 8, dtype=torch.int64, device="cuda")
    g16.random_points_dev(d_b.data_ptr(), n, 0xB254)
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    d_s = torch.randint(0, 2**62, (n * 4,), dtype=torch.int64, device="cuda", generator=gen)
    d_s.view(-1, 4)[:, 3] &= (1 << 59) - 1
    plain = g16.multiexp_dev(d_b.data_ptr(), d_s.data_ptr(), n)
    tab = g16.MsmTable(device_ptr=d_b.data_ptr(), n=n)
    assert (tab.run_dev(d_s.data_ptr()) == plain).all()
    tab.free()
    # 8 per-rank tables (the multi-GPU split): partial sums gathered in one device buffer and added by one kernel
    parts = []
    for k in range(8):
        lo = k * (n // 8)
        t = g16.MsmTable(device_ptr=d_b.data_ptr() + lo * 64, n=n // 8)
        norm = t.run_dev(d_s.data_ptr() + lo * 32)
        if k % 2:                 # un-normalised partial sums (what the ranks of a multi-GPU run produce): another triple of the same point
            t.set_partial_output(True)
            raw = t.run_dev(d_s.data_ptr() + lo * 32)
            assert (raw != norm).any() and (raw[8:12] != norm[8:12]).any()
            one = torch.from_numpy(raw.view(np.int64).copy()).cuda()
            assert (g16.points_sum_dev(one.data_ptr(), 1) == norm).all()
            parts.append(raw)
        else:
            parts.append(norm)
        t.free()
    gathered = torch.from_numpy(np.concatenate(parts).view(np.int64)).cuda()
    assert (g16.points_sum_dev(gathered.data_ptr(), 8) == plain).all()
    # infinity stays infinity in the un-normalised form
    from oracle import bn254 as bn
    t = g16.MsmTable(bn.pack_points([bn.G1, bn.G1])); t.set_partial_output(True)
    assert bn.unpack_point(g16.jacobian_to_affine_mont(t.run(bn.pack_scalars([5, bn.R - 5])))) is None
