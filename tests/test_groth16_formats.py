"""The reference's own serde fixtures for the groth16 boundary (groth16/src/json_utils.rs:350-429 round-trips
groth16/test-vectors/{verification_key,verification_key_bls12381,proof}.{bin,json}): parsing the .bin files must give the
JSON twins the reference wrote, and the base-packing helpers must agree with the curve oracle."""
import json, os
import numpy as np
from eigen_zkvm_b200 import groth16_formats as gf


def _load(golden_dir, name): return open(os.path.join(golden_dir, name), "rb").read()


def test_verification_key_bin_equals_json_twin(golden_dir):
    for binf, jsonf, curve in (("groth16_verification_key.bin", "groth16_verification_key.json", "BN128"),
                               ("groth16_verification_key_bls12381.bin", "groth16_verification_key_bls12381.json", "BLS12381")):
        vk = gf.read_vk_bin(_load(golden_dir, binf), curve)
        twin = json.loads(_load(golden_dir, jsonf))
        assert vk == twin, curve
    assert len(_load(golden_dir, "groth16_verification_key.bin")) == 708 and len(_load(golden_dir, "groth16_verification_key_bls12381.bin")) == 1060


def test_proof_bin_equals_json_twin(golden_dir):
    pr = gf.read_proof_bin(_load(golden_dir, "groth16_proof.bin"), "BN128")
    assert pr == json.loads(_load(golden_dir, "groth16_proof.json"))


def test_truncated_and_trailing_input_rejected(golden_dir):
    import pytest
    d = _load(golden_dir, "groth16_verification_key.bin")
    with pytest.raises(ValueError):
        gf.read_vk_bin(d[:-1], "BN128")
    with pytest.raises(ValueError):
        gf.read_vk_bin(d + b"\x00", "BN128")


def test_base_packing_matches_curve_oracle(golden_dir):
    from oracle import curves as C
    for jsonf, curve, c1, c2 in (("groth16_verification_key.json", "BN128", C.BN254_G1, C.BN254_G2),
                                 ("groth16_verification_key_bls12381.json", "BLS12381", C.BLS381_G1, C.BLS381_G2)):
        pts = gf.vk_points(json.loads(_load(golden_dir, jsonf)))
        w1 = gf.g1_to_words(pts["g1"] + [None], curve); w2 = gf.g2_to_words(pts["g2"] + [None], curve)
        for p, row in zip(pts["g1"] + [None], w1):
            assert list(row) == c1.affine_to_words(p) and c1.affine_from_words(row) == p
        for p, row in zip(pts["g2"] + [None], w2):
            assert list(row) == c2.affine_to_words(p) and c2.affine_from_words(row) == p
