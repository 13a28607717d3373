"""End-to-end checks of the CPU oracle's stark_gen: same acceptance criterion as the reference's own
tests (starky/src/stark_gen.rs:1149-1195: setup -> stark_gen -> stark_verify == true), plus regression
pins and the independently derived roots of SURVEY.md Appendix C."""
import hashlib, json, os
import numpy as np
import pytest
from oracle import stark_oracle as so
from eigen_zkvm_b200 import starkinfo as si


def _load(golden_dir, name):
    pil = si.load_pil(os.path.join(golden_dir, name + ".pil.json.gl"))
    ss = json.load(open(os.path.join(golden_dir, "starkStruct.json.gl")))
    cm = np.fromfile(os.path.join(golden_dir, name + ".cm.gl"), dtype="<u8")
    const = np.fromfile(os.path.join(golden_dir, name + ".const.gl"), dtype="<u8")
    return pil, ss, cm, const


def test_fibonacci_generator_reproduces_fixture(golden_dir):
    cm, const = so.fibonacci_inputs(10)
    assert (cm.reshape(-1) == np.fromfile(os.path.join(golden_dir, "fib.cm.gl"), dtype="<u8")).all()
    assert (const.reshape(-1) == np.fromfile(os.path.join(golden_dir, "fib.const.gl"), dtype="<u8")).all()


def test_codegen_shapes_fibonacci(golden_dir):
    pil, ss, _, _ = _load(golden_dir, "fib")
    info, prog = si.new_starkinfo(pil, ss)
    assert (info.n_cm1, info.n_cm2, info.n_cm3, info.n_cm4, info.q_deg, info.q_dim) == (2, 0, 0, 1, 1, 3)
    assert len(prog["step42ns"]["first"]) == 17 and prog["step42ns"]["tmp_used"] == 16
    assert len(prog["step52ns"]["first"]) == 25
    assert [(e["type_"], e["id"], e["prime"]) for e in info.ev_map] == [
        ("const", 0, False), ("cm", 0, True), ("cm", 1, False), ("cm", 0, False), ("cm", 1, True), ("cm", 2, False)]
    json.loads(si.setup_json(info, prog, ss))


def test_codegen_shapes_plookup(golden_dir):
    pil, ss, _, _ = _load(golden_dir, "plookup")
    info, prog = si.new_starkinfo(pil, ss)
    assert (info.n_cm1, info.n_cm2, info.n_cm3, info.n_cm4, info.q_deg, info.q_dim) == (4, 2, 3, 2, 2, 3)
    assert [len(prog[k]["first"]) for k in ("step2prev", "step3prev", "step3", "step42ns", "step52ns")] == [22, 53, 42, 67, 86]
    assert len(info.ev_map) == 21
    assert info.map_sectionsN["cm2_n"] == 6 and info.map_sectionsN["cm3_n"] == 9 and info.map_sectionsN["tmpexp_n"] == 6


@pytest.mark.parametrize("name", ["pe", "connection"])
def test_prove_verify_permutation_connection(golden_dir, name):
    # stark_gen.rs:1023-1148 fixtures (there with the BN128 hash); same acceptance criterion
    pil = si.load_pil(os.path.join(golden_dir, name + ".pil.json"))
    ss = json.load(open(os.path.join(golden_dir, "starkStruct.json.gl")))
    cm = np.fromfile(os.path.join(golden_dir, name + ".cm"), dtype="<u8"); const = np.fromfile(os.path.join(golden_dir, name + ".const"), dtype="<u8")
    setup = so.stark_setup(const, pil, ss)
    proof = so.stark_gen(cm, const, setup, ss)
    assert so.stark_verify(proof, setup["const_root"], setup["starkinfo"], ss, setup["program"])
    assert so.proof_to_json(proof) == open(os.path.join(golden_dir, name + "10.proof.json")).read()


@pytest.mark.parametrize("name", ["fib", "plookup"])
def test_prove_verify_fixture(golden_dir, name):
    pil, ss, cm, const = _load(golden_dir, name)
    setup = so.stark_setup(const, pil, ss)
    proof = so.stark_gen(cm, const, setup, ss)
    assert so.stark_verify(proof, setup["const_root"], setup["starkinfo"], ss, setup["program"])
    js = so.proof_to_json(proof)
    golden = open(os.path.join(golden_dir, name + "10.proof.json")).read()
    assert js == golden                                       # regression pin (oracle output is deterministic)
    back = so.proof_from_json(js)                             # serde round trip (stark_gen.rs:1180-1187)
    assert so.stark_verify(back, setup["const_root"], setup["starkinfo"], ss, setup["program"])
    ev0 = back["evals"][0]
    back["evals"][0] = (ev0[0] ^ 1, ev0[1], ev0[2])           # tamper one lane -> rejected
    assert not so.stark_verify(back, setup["const_root"], setup["starkinfo"], ss, setup["program"])
    if name == "fib":
        # stark_setup.rs:100-116 KAT + SURVEY.md Appendix C (derived by an independent restatement)
        assert setup["const_root"] == [15302509084042343527, 985081440042889555, 14692153289195851822, 1611894784155222896]
        assert proof["root1"] == [6591581766092552436, 3248318708045465285, 11214734650546771418, 4874274597592687832]
        assert proof["root2"] == proof["root3"] == [10191288259157808067, 944536249556834531, 16268598854718968908, 2417244819673331317]
        assert proof["root4"] == [13656481478292468445, 4033401606716075800, 3200995078452154898, 5366720658569992306]
        assert proof["publics"] == [11696381471667068125]
        assert proof["evals"][5] == (9457068024692792230, 5024045632355647482, 12403197773770060969)
        assert proof["fri"]["last"][0] == (1681700097388906458, 11069517619086642873, 6283742259489242253)
    else:
        assert setup["const_root"] == [3211021716450354219, 12729857658698525015, 363862269707841568, 1828601116351836526]
        assert proof["root3"] == [8229705030647647208, 13494235484931506177, 5791251872178438624, 17613038797463224100]
        assert proof["root4"] == [7643301529802113906, 15434704686344870993, 7869556029450542646, 17769937482814441768]


def test_fibonacci_2_12_config1(golden_dir):
    # BASELINE.json config 1 (synthetic 2^12, steps 13/9/5, nQueries 8); values of SURVEY.md Appendix C
    ss = {"nBits": 12, "nBitsExt": 13, "nQueries": 8, "verificationHashType": "GL", "steps": [{"nBits": 13}, {"nBits": 9}, {"nBits": 5}]}
    cm, const = so.fibonacci_inputs(12)
    setup = so.stark_setup(const, so.fibonacci_pil(os.path.join(golden_dir, "fib.pil.json.gl"), 12), ss)
    proof = so.stark_gen(cm, const, setup, ss)
    assert so.stark_verify(proof, setup["const_root"], setup["starkinfo"], ss, setup["program"])
    assert setup["const_root"] == [6214561362095951581, 5719160207234536665, 7910044398877437422, 13567121903497174917]
    assert proof["root1"] == [18430399238323009296, 12719397174109614227, 15109511881276515236, 8825810024238079302]
    assert proof["root4"] == [10821576761769313892, 7457168644129333439, 18441848496933695346, 6410940867872620433]
    assert proof["publics"] == [15228958502552419041]
    assert len(proof["fri"]["last"]) == 32
    d = json.load(open(os.path.join(golden_dir, "derived_goldens.json")))["fib12"]
    assert hashlib.sha256(so.proof_to_json(proof).encode()).hexdigest() == d["proof_sha256"]


PROVER_ADDR = "273030697313060285579891744179749754319274977764"


@pytest.mark.parametrize("name,struct,tag", [("fib", "starkStruct.json", "bn128"), ("fib", "starkStruct.json.bls12381", "bls12381"), ("plookup", "starkStruct.json", "bn128")])
def test_prove_verify_big_hash_fixtures(golden_dir, name, struct, tag):
    """The reference's BN128 / BLS12-381 end-to-end tests (stark_gen.rs:981-1022 fib, :1093-1148 plookup; same criterion:
    setup -> stark_gen -> serde round trip -> stark_verify == true) and its const-root KAT (stark_setup.rs:83-98)."""
    pil = si.load_pil(os.path.join(golden_dir, name + ".pil.json"))
    ss = json.load(open(os.path.join(golden_dir, struct)))
    cm = np.fromfile(os.path.join(golden_dir, name + ".cm"), dtype="<u8"); const = np.fromfile(os.path.join(golden_dir, name + ".const"), dtype="<u8")
    setup = so.stark_setup(const, pil, ss)
    if (name, tag) == ("fib", "bn128"):
        assert setup["const_root"] == [4658128321472362347225942316135505030498162093259225938328465623672244875764]
    proof = so.stark_gen(cm, const, setup, ss)
    js = so.proof_to_json(proof, PROVER_ADDR)
    assert js == open(os.path.join(golden_dir, "%s10.%s.proof.json" % (name, tag))).read()        # regression pin
    assert json.loads(js)["proverAddr"] == PROVER_ADDR and list(json.loads(js))[-1] == "proverAddr"   # serializer.rs:262-266
    back = so.proof_from_json(js, ss["verificationHashType"])
    assert so.stark_verify(back, setup["const_root"], setup["starkinfo"], ss, setup["program"])
    ev0 = back["evals"][0]
    back["evals"][0] = (ev0[0] ^ 1, ev0[1], ev0[2])
    assert not so.stark_verify(back, setup["const_root"], setup["starkinfo"], ss, setup["program"])
    back = so.proof_from_json(js, ss["verificationHashType"])
    q = back["fri"]["queries"][0]["pol_queries"][0][0]
    q[1][0][3] ^= 1                                                                                # one sibling of tree1's first level
    assert not so.stark_verify(back, setup["const_root"], setup["starkinfo"], ss, setup["program"])
    if name == "fib":
        assert proof["publics"] == [1, 2, 74469561660084004]      # in1 is an `imP` public (calculate_exp_at_point, stark_gen.rs:559-572)
