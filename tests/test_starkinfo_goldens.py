"""The PIL codegen (eigen_zkvm_b200/starkinfo.py, port of starky/src/starkinfo*.rs) pinned independently of the prover / oracle
pair that consumes it: (1) the step programs of the six reference fixtures equal the committed renderings under
tests/golden/starkinfo/ (tools/gen_starkinfo_goldens.py); (2) the Fibonacci programs equal the listing in SURVEY.md Appendix C,
typed here by hand from that text (it was derived by a separate throw-away port during the survey); (3) the plookup fixture has
the shapes Appendix C records (program lengths, section widths, number of evaluations)."""
import os, re, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_starkinfo_goldens as gg   # noqa: E402


@pytest.mark.parametrize("name,pil,ss", gg.FIXTURES)
def test_emitted_programs_equal_the_goldens(name, pil, ss):
    want = open(os.path.join(gg.G, "starkinfo", name + ".txt")).read()
    assert gg.render(pil, ss) == want


def _section(text, name):
    m = re.search(r"^\[%s\] (\d+) ops\n((?:(?!\[).*\n)*)" % re.escape(name), text, re.M)
    return int(m.group(1)), [l for l in m.group(2).splitlines() if l]


def _canon(lines):
    """rename temporaries in order of first definition (the allocation order differs between ports, the dataflow must not)"""
    names = {}
    def sub(m):
        return names.setdefault(m.group(0), "t%d" % len(names))
    return [re.sub(r"tmp\d+", sub, l) for l in lines]


# SURVEY.md Appendix C, `step42ns` and `step52ns` of fib.pil.json.gl (vc = challenge 4, vf1 = 5, vf2 = 6; ' = next row)
APPENDIX_C_STEP42NS = """
t0 = number(1) - const0
t1 = cm0' - cm1
t2 = t0 * t1
t3 = t2 - number(0)
t4 = number(1) - const0
t5 = cm0 + cm1
t6 = cm1' - t5
t7 = t4 * t6
t8 = t7 - number(0)
t9 = cm1 - public0
t10 = const0 * t9
t11 = t10 - number(0)
t12:3 = challenge4:3 * t3
t13:3 = t12:3 + t8
t14:3 = challenge4:3 * t13:3
t15:3 = t14:3 + t11
q0:3 = t15:3 * Zi
""".strip().splitlines()

APPENDIX_C_STEP52NS = """
t0:3 = challenge5:3 * cm0
t1:3 = t0:3 + cm1
t2:3 = challenge5:3 * t1:3
t3:3 = t2:3 + cm2:3
t4:3 = challenge5:3 * t3:3
t5:3 = const0 - eval0:3
t6:3 = t5:3 * challenge6:3
t7:3 = cm1 - eval2:3
t8:3 = t6:3 + t7:3
t9:3 = t8:3 * challenge6:3
t10:3 = cm0 - eval3:3
t11:3 = t9:3 + t10:3
t12:3 = t11:3 * challenge6:3
t13:3 = cm2:3 - eval5:3
t14:3 = t12:3 + t13:3
t15:3 = t14:3 * xDivXSubXi
t16:3 = t4:3 + t15:3
t17:3 = challenge5:3 * t16:3
t18:3 = cm0 - eval1:3
t19:3 = t18:3 * challenge6:3
t20:3 = cm1 - eval4:3
t21:3 = t19:3 + t20:3
t22:3 = t21:3 * xDivXSubWXi
t23:3 = t17:3 + t22:3
f0:3 = t23:3
""".strip().splitlines()


def test_fibonacci_programs_equal_survey_appendix_c():
    text = gg.render("fib.pil.json.gl", "starkStruct.json.gl")
    assert "n_cm1=2 n_cm2=0 n_cm3=0 n_cm4=1" in text and "q_deg=1 q_dim=3" in text
    assert "cm1_n=2" in text and "cm1_2ns=2" in text and "cm4_2ns=3" in text and "q_2ns=3" in text and "f_2ns=3" in text
    assert "ev_map const0 cm0' cm1 cm0 cm1' cm2\n" in text            # order fixes `evals` in the proof and the transcript
    n, lines = _section(text, "step42ns")
    assert n == 17 and _canon(lines) == APPENDIX_C_STEP42NS
    n, lines = _section(text, "step52ns")
    assert n == 25 and _canon(lines) == APPENDIX_C_STEP52NS
    for empty in ("step2prev", "step3prev", "step3"):
        assert _section(text, empty)[0] == 0


def test_plookup_shapes_equal_survey_appendix_c():
    text = gg.render("plookup.pil.json.gl", "starkStruct.json.gl")
    assert "n_cm1=4 n_cm2=2 n_cm3=3 n_cm4=2" in text and "q_deg=2 q_dim=3" in text
    for w in ("cm2_n=6", "cm3_n=9", "cm4_2ns=6", "tmpexp_n=6"):
        assert w in text
    assert len(re.search(r"^ev_map (.*)$", text, re.M).group(1).split()) == 21
    assert [_section(text, s)[0] for s in ("step2prev", "step3prev", "step3", "step42ns", "step52ns")] == [22, 53, 42, 67, 86]
