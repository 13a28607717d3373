"""Pins the BN254 MSM oracle: curve constants vs the reference's own data, C Pippenger vs python big-int."""
import json, os, random
import numpy as np
from oracle import bn254 as bn


def test_constants_match_reference_sources():
    # groth16/src/api.rs:636 (Fq modulus in the Solidity template), starky/src/field_bn128.rs:12 (Fr modulus)
    assert bn.Q == 21888242871839275222246405745257275088696311157297823662689037894645226208583
    assert bn.R == 21888242871839275222246405745257275088548364400416034343698204186575808495617
    assert bn.is_on_curve(bn.G1) and bn.mul(bn.R, bn.G1) is None
    # 2G, the well known EIP-196 value
    assert bn.add(bn.G1, bn.G1) == (1368015179489954701390400359078579693043519447331113978918064868415326638035,
                                    9918110051302171585080402603319702774565515993150576347155970296011118125764)


def test_reference_test_vector_points_are_on_curve(golden_dir):
    # groth16/test-vectors/{verification_key,proof}.json (groth16/src/json_utils.rs:350-429 round-trips them)
    vk = json.load(open(os.path.join(golden_dir, "groth16_verification_key.json")))
    pr = json.load(open(os.path.join(golden_dir, "groth16_proof.json")))
    pts = [vk["vk_alpha_1"], vk["vk_beta_1"], vk["vk_delta_1"]] + vk["IC"] + [pr["pi_a"], pr["pi_c"]] if "vk_delta_1" in vk else [vk["vk_alpha_1"], vk["vk_beta_1"]] + vk["IC"] + [pr["pi_a"], pr["pi_c"]]
    assert len(pts) >= 5
    for p in pts:
        q = (int(p["x"]) if isinstance(p, dict) else int(p[0]), int(p["y"]) if isinstance(p, dict) else int(p[1]))
        assert bn.is_on_curve(q)
        assert bn.on_curve_c(bn.pack_points([q])[0])
        assert bn.mul(bn.R, q) is None


def test_c_pippenger_matches_python_naive():
    rnd = random.Random(5)
    for n in (1, 2, 7, 33, 200):
        pts = [bn.mul(rnd.randrange(1, bn.R), bn.G1) for _ in range(n)]
        sc = [rnd.randrange(bn.R) for _ in range(n)]
        if n >= 7:
            pts[3] = None; sc[4] = 0; sc[5] = bn.R - 1; pts[6] = pts[2]
        exp = bn.msm_naive(pts, sc)
        B = bn.pack_points(pts); S = bn.pack_scalars(sc)
        assert bn.unpack_point(bn.msm_c(B, S)) == exp
        assert bn.unpack_point(bn.msm_c(B, S, naive=True)) == exp
    # cancellation: P + (-P) = infinity
    p = bn.mul(12345, bn.G1)
    assert bn.unpack_point(bn.msm_c(bn.pack_points([p, p]), bn.pack_scalars([5, bn.R - 5]))) is None
