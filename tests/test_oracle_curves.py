"""Pins the generic curve oracle (oracle/curves.py): moduli vs the reference's sources, curve/twist equations vs the
on-curve points the reference ships in groth16/test-vectors/, generators by r * G = O, and against oracle/bn254.py."""
import json, os, random
from oracle import curves as C, bn254 as bn


def test_moduli_match_reference_sources():
    # groth16/src/api.rs:636, starky/src/field_bn128.rs:12, starky/src/field_bls12381.rs:12
    assert C.BN254_Q == bn.Q and C.BN254_R == bn.R
    assert C.BLS381_R == 52435875175126190479447740508185965837690552500527637822603658699938581184513
    assert C.BLS381_Q % 4 == 3 and C.BLS381_Q.bit_length() == 381


def test_generators():
    for c in C.CURVES.values():
        assert c.is_on_curve(c.gen) and c.mul(c.r, c.gen) is None and c.mul(c.r - 1, c.gen) == c.neg(c.gen)


def _g1(p): return (int(p["x"]), int(p["y"])) if isinstance(p, dict) else (int(p[0]), int(p[1]))
def _g2(p): return (tuple(int(v) for v in p["x"]), tuple(int(v) for v in p["y"])) if isinstance(p, dict) else (tuple(int(v) for v in p[0]), tuple(int(v) for v in p[1]))


def test_reference_vk_points_on_curve_and_in_subgroup(golden_dir):
    # groth16/test-vectors/verification_key{,_bls12381}.json; the JSON twin stores G2 as x: [c0, c1] (groth16/src/json_utils.rs:36-40)
    for fn, g1, g2 in (("groth16_verification_key.json", C.BN254_G1, C.BN254_G2), ("groth16_verification_key_bls12381.json", C.BLS381_G1, C.BLS381_G2)):
        vk = json.load(open(os.path.join(golden_dir, fn)))
        p1 = [_g1(vk[k]) for k in ("vk_alpha_1", "vk_beta_1", "vk_delta_1") if k in vk] + [_g1(p) for p in vk["IC"]]
        p2 = [_g2(vk[k]) for k in ("vk_beta_2", "vk_gamma_2", "vk_delta_2") if k in vk]
        assert len(p1) >= 3 and len(p2) >= 2
        for p in p1:
            assert g1.is_on_curve(p) and g1.mul(g1.r, p) is None
        for p in p2:
            assert g2.is_on_curve(p) and g2.mul(g2.r, p) is None


def test_matches_bn254_oracle_and_word_round_trip():
    rnd = random.Random(3)
    c = C.BN254_G1
    for _ in range(5):
        k = rnd.randrange(1, c.r)
        assert c.mul(k, c.gen) == bn.mul(k, bn.G1)
    for c in C.CURVES.values():
        p = c.mul(rnd.randrange(1, 1 << 64), c.gen)
        assert c.affine_from_words(c.affine_to_words(p)) == p and c.affine_from_words(c.affine_to_words(None)) is None
        one = c._f_to_u64(c.F.one)
        assert c.jacobian_from_words(c.affine_to_words(p) + one) == p
        # linearity of the naive MSM (sanity of add/mul): sum k_i P = (sum k_i) P for equal points
        ks = [rnd.randrange(c.r) for _ in range(4)]
        assert c.msm_naive([p] * 4, ks) == c.mul(sum(ks) % c.r, p)


def test_c_pippenger_matches_python_bigint_on_all_four_groups():
    """oracle/curves_oracle.c (unsigned-window Pippenger, Jacobian, 64-bit CIOS) == python big-int double-and-add."""
    import numpy as np
    rnd = random.Random(5)
    for c in C.CURVES.values():
        for n in ((1, 2, 37, 120) if c.deg == 1 else (1, 2, 37)):      # python big ints on G2 are slow: keep the CPU suite short
            pts = [c.mul(rnd.randrange(1, 1 << 40), c.gen) for _ in range(n)]
            sc = [rnd.randrange(c.r) for _ in range(n)]
            if n >= 37:
                pts[3] = None; sc[4] = 0; sc[5] = c.r - 1; pts[6] = pts[2]; sc[7] = 1; pts[8] = c.neg(pts[9]); sc[8] = sc[9]      # infinity, 0, -1, repeats, cancellation
            bw = np.array([c.affine_to_words(p) for p in pts], dtype=np.uint64)
            sw = np.array([[(s >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for s in sc], dtype=np.uint64)
            assert c.affine_from_words(C.msm_c(c, bw, sw)) == c.msm_naive(pts, sc), (c.name, n)
        p = c.mul(12345, c.gen)
        bw = np.array([c.affine_to_words(p)] * 2, dtype=np.uint64)
        sw = np.array([[5, 0, 0, 0], [(c.r - 5) & 0xFFFFFFFFFFFFFFFF, ((c.r - 5) >> 64) & 0xFFFFFFFFFFFFFFFF, ((c.r - 5) >> 128) & 0xFFFFFFFFFFFFFFFF, (c.r - 5) >> 192]], dtype=np.uint64)
        assert c.affine_from_words(C.msm_c(c, bw, sw)) is None


def test_c_pippenger_matches_the_bn254_c_oracle():
    import numpy as np
    rnd = random.Random(6)
    n = 3000
    pts = [bn.mul(rnd.randrange(1, 1 << 30), bn.G1) for _ in range(40)]
    pts = [pts[rnd.randrange(40)] for _ in range(n)]
    sc = [rnd.randrange(bn.R) for _ in range(n)]
    B = bn.pack_points(pts); S = bn.pack_scalars(sc)
    assert (C.msm_c(C.BN254_G1, B, S) == bn.msm_c(B, S)).all()
