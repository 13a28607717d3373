"""PIL -> (StarkInfo, Program): the host-side "compiler" that feeds `stark_gen`.

In the reference this is `StarkInfo::new` (starky/src/starkinfo.rs:170-275) and its helpers
(starkinfo_codegen.rs, starkinfo_Z.rs, starkinfo_cp_prover.rs, starkinfo_cp_ver.rs,
starkinfo_fri_prover.rs, starkinfo_fri_ver.rs, starkinfo_map.rs, expressionops.rs).  It runs on the
*caller's* side of the `StarkProof::stark_gen(cm_pols, const_pols, const_tree, &starkinfo, &program, ..)`
boundary (starky/src/stark_gen.rs:193-202): milliseconds of serial host work whose output (the step
programs) drives the device evaluator.  A Rust caller would hand our C-ABI the serde_json of its own
`StarkInfo`/`Program`; because rustc is not available in this image, this module re-derives the same
structures from the PIL JSON and emits them with the serde field names (`to_json`), so the C++ side
parses exactly what `serde_json::to_string(&setup.starkinfo)` would contain.

Data model: plain dicts/lists mirroring the Rust structs (Node, Section, Segment, PolType, PCCTX, ...).
"""
import copy
import json
import sys

CHALLENGE_MAP = {"u": 0, "defVal": 1, "gamma": 2, "beta": 3, "vc": 4, "vf1": 5, "vf2": 6, "xi": 7}  # constant.rs:39-50
GLOBAL_L1 = "Global.L1"      # constant.rs:118
GL_P = 0xFFFFFFFF00000001
KS0 = 12275445934081160404   # helper.rs:16-23 (get_ks)

SECTION_NAMES = ["cm1_n", "cm1_2ns", "cm2_n", "cm2_2ns", "cm3_n", "cm3_2ns", "cm4_n", "cm4_2ns", "tmpexp_n", "q_2ns", "f_2ns"]


# ----------------------------------------------------------------------------------------------
# expressions (types.rs:34-96, expressionops.rs)
# ----------------------------------------------------------------------------------------------
def _expr(op, id=None, value=None, values=None, next=None):
    return {"op": op, "deg": 0, "id": id, "next": next, "value": value, "values": values,
            "keep": None, "keep2ns": None, "idQ": None, "const_": None}


def _norm_expr(e):
    for k in ("id", "next", "value", "values", "keep", "keep2ns", "idQ"):
        e.setdefault(k, None)
    if "const_" not in e:
        e["const_"] = e.get("const")
    e.setdefault("deg", 0)
    if e["values"]:
        for v in e["values"]:
            _norm_expr(v)
    return e


class E:
    add = staticmethod(lambda a, b: _expr("add", values=[copy.deepcopy(a), copy.deepcopy(b)]))
    sub = staticmethod(lambda a, b: _expr("sub", values=[copy.deepcopy(a), copy.deepcopy(b)]))
    mul = staticmethod(lambda a, b: _expr("mul", values=[copy.deepcopy(a), copy.deepcopy(b)]))
    exp = staticmethod(lambda id, next=None: _expr("exp", id=id, next=next))
    cm = staticmethod(lambda id, next=None: _expr("cm", id=id, next=next))
    const_ = staticmethod(lambda id, next=None: _expr("const", id=id, next=next))
    q = staticmethod(lambda id, next=None: _expr("q", id=id, next=next))
    challenge = staticmethod(lambda name: _expr("challenge", id=CHALLENGE_MAP[name]))
    number = staticmethod(lambda n: _expr("number", value=str(n)))
    eval = staticmethod(lambda n: _expr("eval", id=n))
    xDivXSubXi = staticmethod(lambda: _expr("xDivXSubXi"))
    xDivXSubWXi = staticmethod(lambda: _expr("xDivXSubWXi"))
    x = staticmethod(lambda: _expr("x"))
    nop = staticmethod(lambda: _expr("nop"))
    is_nop = staticmethod(lambda e: e["op"] == "nop")


def _next(e):
    return bool(e.get("next"))


def load_pil(obj):
    """Accepts a path, a JSON string or a dict; returns a normalised deep copy (types.rs:141-166)."""
    if isinstance(obj, str):
        obj = json.loads(obj) if obj.lstrip().startswith("{") else json.load(open(obj))
    pil = copy.deepcopy(obj)
    for e in pil["expressions"]:
        _norm_expr(e)
    pil.setdefault("permutationIdentities", None)
    pil.setdefault("connectionIdentities", None)
    pil["cm_dims"] = []
    pil["q2exp"] = []
    return pil


# ----------------------------------------------------------------------------------------------
# code generation (starkinfo_codegen.rs)
# ----------------------------------------------------------------------------------------------
def node(type_, id=0, value=None, dim=0, prime=False, tree_pos=0):
    return {"type_": type_, "id": id, "value": value, "dim": dim, "prime": bool(prime), "tree_pos": tree_pos, "p": 0, "exp_id": 0}


class Context:
    def __init__(self):
        self.exp_id = 0
        self.tmp_used = 0
        self.code = []          # list of {exp_id, prime, code:[Section], tmp_used, idQ}
        self.calculated = {}    # (("exps"|"expsPrime"), exp_id) -> bool ; membership matters (codegen.rs:306)


def _eval_exp(cc, exp, prime):
    """starkinfo_codegen.rs:387-421 + eval_single_op: left-to-right post-order, iterative."""
    out = []                    # result stack
    stack = [(exp, False)]
    while stack:
        e, done = stack.pop()
        vals = e["values"] or []
        if not done and vals:
            stack.append((e, True))
            for v in reversed(vals):
                stack.append((v, False))
            continue
        k = len(vals)
        args = out[len(out) - k:] if k else []
        if k:
            del out[len(out) - k:]
        out.append(_eval_single(cc, e, prime, args))
    assert len(out) == 1
    return out[0]


def _eval_single(cc, exp, prime, values):
    op = exp["op"]
    if op in ("add", "sub", "mul", "muladd"):
        r = node("tmp", cc["tmp_used"]); cc["tmp_used"] += 1
        cc["code"].append({"op": op, "dest": dict(r), "src": list(values)})
        return r
    if op in ("addc", "mulc"):
        a = values[0]; b = node("number", 0, str(exp["const_"]))
        r = node("tmp", cc["tmp_used"]); cc["tmp_used"] += 1
        cc["code"].append({"op": "add" if op == "addc" else "mul", "dest": dict(r), "src": [a, b]})
        return r
    if op == "neg":
        a = node("number", 0, "0"); b = values[0]
        r = node("tmp", cc["tmp_used"]); cc["tmp_used"] += 1
        cc["code"].append({"op": "sub", "dest": dict(r), "src": [a, b]})
        return r
    if op in ("cm", "const", "exp", "q"):
        if _next(exp) and prime:
            raise ValueError("Double Prime")
        return node(op, exp["id"], None, 0, _next(exp) or prime)
    if op == "number":
        return node("number", 0, exp["value"])
    if op in ("public", "challenge", "eval"):
        return node(op, exp["id"])
    if op in ("xDivXSubXi", "xDivXSubWXi", "x"):
        return node(op, 0)
    raise ValueError("InvalidOperator: eval_exp: %s" % op)


def _find_muladd(exp):
    """starkinfo_codegen.rs:343-370."""
    if exp["values"] is not None:
        values = exp["values"]
        if exp["op"] == "add" and values[0]["op"] == "mul":
            vv = values[0]["values"]
            return _expr("muladd", values=[_find_muladd(vv[0]), _find_muladd(vv[1]), _find_muladd(values[1])])
        if exp["op"] == "add" and values[1]["op"] == "mul":
            vv = values[1]["values"]
            return _expr("muladd", values=[_find_muladd(vv[0]), _find_muladd(vv[1]), _find_muladd(values[0])])
        r = dict(exp)
        mv = [_find_muladd(v) for v in values]
        if mv:
            r["values"] = mv
        return r
    return dict(exp)


def _calculate_deps(ctx, pil, expr, prime, exp_id, muladd):
    """Pre-order walk (codegen.rs:547-568): an `exp` node triggers code generation of its target first."""
    stack = [expr]
    while stack:
        e = stack.pop()
        if e["op"] == "exp":
            if prime and _next(e):
                raise ValueError("Double prime")
            pil_code_gen(ctx, pil, e["id"], prime or _next(e), "", 0, muladd)
        if e["values"] is not None:
            for v in reversed(e["values"]):
                stack.append(v)


def pil_code_gen(ctx, pil, exp_id, prime, res_type, res_id, muladd):
    """starkinfo_codegen.rs:294-341."""
    prime_idx = "expsPrime" if prime else "exps"
    if (prime_idx, exp_id) in ctx.calculated:
        if res_type:
            c = next(x for x in ctx.code if x["exp_id"] == exp_id and x["prime"] == prime)
            dest = node(res_type, res_id, None, 0, prime)
            c["code"].append({"op": "copy", "dest": dest, "src": [dict(c["code"][-1]["dest"])]})
        return
    exp = pil["expressions"][exp_id]
    _calculate_deps(ctx, pil, exp, prime, exp_id, False)
    cc = {"exp_id": exp_id, "tmp_used": ctx.tmp_used, "code": []}
    e2 = pil["expressions"][exp_id]
    if muladd:
        e2 = _find_muladd(e2)
    ret = _eval_exp(cc, e2, prime)
    if ret["type_"] == "tmp":
        cc["code"][-1]["dest"] = node("exp", exp_id, None, 0, prime)
        cc["tmp_used"] -= 1
    else:
        cc["code"].append({"op": "copy", "dest": node("exp", exp_id, None, 0, prime), "src": [ret]})
    if res_type:
        if prime:
            raise ValueError("Prime in retType")
        cc["code"].append({"op": "copy", "dest": node(res_type, res_id, None, 0, prime), "src": [node("exp", exp_id, None, 0, prime)]})
    ctx.code.append({"exp_id": exp_id, "prime": prime, "code": cc["code"], "tmp_used": 0, "idQ": None})
    ctx.calculated[(prime_idx, exp_id)] = True
    if cc["tmp_used"] > ctx.tmp_used:
        ctx.tmp_used = cc["tmp_used"]


def _exp_and_expprimes(ctx, pil):
    calc = {}
    for c in ctx.code:
        e = pil["expressions"][c["exp_id"]]
        if e["idQ"] is not None or e["keep"] is not None or e["keep2ns"] is not None:
            calc[c["exp_id"]] = calc.get(c["exp_id"], 0) | (2 if c["prime"] else 1)
    return {k: v == 3 for k, v in calc.items()}


def _build_linear_code(ctx, pil, loop_pos):
    ee = _exp_and_expprimes(ctx, pil) if loop_pos in ("i", "last") else {}
    res = []
    for i, c in enumerate(ctx.code):
        # NB: the reference indexes the map by the *position* i, not by exp_id (codegen.rs:615); kept as is.
        ep = ee.get(i)
        if ep and ((loop_pos == "i" and not c["prime"]) or loop_pos == "last"):
            continue
        res.extend(copy.deepcopy(c["code"]))
    return res


def build_code(ctx, pil):
    """starkinfo_codegen.rs:588-605."""
    seg = {"first": _build_linear_code(ctx, pil, "first"), "i": _build_linear_code(ctx, pil, "i"),
           "last": _build_linear_code(ctx, pil, "last"), "tmp_used": ctx.tmp_used}
    for i, e in enumerate(pil["expressions"]):
        if e["keep"] is None and e["idQ"] is None:
            ctx.calculated[("exps", i)] = False
            ctx.calculated[("expsPrime", i)] = False
    ctx.code = []
    return seg


def iterate_code(seg, f):
    for part in ("first", "i", "last"):
        for c in seg[part]:
            for s in c["src"]:
                f(s)
            f(c["dest"])


def _segment():
    return {"first": [], "i": [], "last": [], "tmp_used": 0}


def _pcctx(**kw):
    d = dict(f_exp_id=0, t_exp_id=0, h1_id=0, h2_id=0, z_id=0, c1_id=0, c2_id=0, num_id=0, den_id=0)
    d.update(kw)
    return d


# ----------------------------------------------------------------------------------------------
# degree analysis for intermediate polynomials (starkinfo_cp_prover.rs:138-288)
# ----------------------------------------------------------------------------------------------
def _deg(pil, exp):
    op = exp["op"]; values = exp["values"] or []
    if op in ("add", "sub", "addc", "mulc", "neg"):
        md = 1
        for v in values:
            md = max(md, _deg(pil, v))
        return md
    if op == "mul":
        return _deg(pil, values[0]) + _deg(pil, values[1])
    if op == "muladd":
        return max(_deg(pil, values[0]) + _deg(pil, values[1]), _deg(pil, values[2]))
    if op in ("cm", "const", "x"):
        return 1
    if op == "exp":
        return _deg(pil, pil["expressions"][exp["id"]])
    if op in ("number", "public", "challenge", "eval"):
        return 0
    raise ValueError("Exp op not defined: %s" % op)


def _calc_im(pil, exp, im, max_deg, abs_max, st):
    if im is None:
        return None, -1
    op = exp["op"]
    if op in ("add", "sub", "addc", "mulc", "neg"):
        md = 0; im_e = dict(im)
        for v in exp["values"]:
            im_e, d = _calc_im(pil, v, im_e, max_deg, abs_max, st)
            if d > md:
                md = d
        return im_e, md
    if op in ("number", "public", "challenge"):
        return dict(im), 0
    if op in ("x", "const", "cm"):
        if max_deg < 1:
            return None, -1
        return dict(im), 1
    if op == "mul":
        eb = None; ed = -1
        values = exp["values"]
        if values[0]["op"] in ("number", "public", "challenge"):
            return _calc_im(pil, values[1], im, max_deg, abs_max, st)
        if values[1]["op"] in ("number", "public", "challenge"):
            return _calc_im(pil, values[0], im, max_deg, abs_max, st)
        here = _deg(pil, exp)
        if here <= max_deg:
            return dict(im), here
        for l in range(0, max_deg + 1):
            r = max_deg - l
            e1, d1 = _calc_im(pil, values[0], im, l, abs_max, st)
            e2, d2 = _calc_im(pil, values[1], e1, r, abs_max, st)
            if e2 is not None:
                if eb is None or len(e2) < len(eb):
                    eb = e2; ed = d1 + d2
            if eb is not None and len(eb) == len(im):
                return eb, ed
        return eb, ed
    if op == "exp":
        if max_deg < 1:
            return None, -1
        if exp["id"] in im:
            return dict(im), 1
        e, d = _calc_im(pil, pil["expressions"][exp["id"]], im, abs_max, abs_max, st)
        if e is None:
            return None, -1
        if d > max_deg:
            e[exp["id"]] = True
            if d > st[0]:
                st[0] = d
            return e, 1
        return e, d
    raise ValueError("Exp op not defined: %s" % op)


def calculate_im_pols(pil, exp, max_deg):
    st = [0]
    re, rd = _calc_im(pil, exp, {}, max_deg, max_deg, st)
    return re, max(rd, st[0]) - 1


# ----------------------------------------------------------------------------------------------
# StarkInfo::new
# ----------------------------------------------------------------------------------------------
class StarkInfo:
    def __init__(self):
        self.var_pol_map = []
        self.n_cm1 = self.n_cm2 = self.n_cm3 = self.n_cm4 = self.n_q = 0
        self.pu_ctx, self.pe_ctx, self.ci_ctx = [], [], []
        self.n_constants = self.n_publics = self.c_exp = 0
        self.im_exps = {}
        self.q_deg = self.q_dim = 0
        self.im_exps_list = []
        self.im_exp2cm = {}
        self.qs, self.exps_2ns, self.exps_n = [], [], []
        self.ev_map = []
        self.fri_exp_id = self.n_exps = 0
        self.cm_n, self.cm_2ns, self.tmpexp_n, self.q_2ns, self.f_2ns = [], [], [], [], []
        self.map_sections = {k: [] for k in SECTION_NAMES}
        self.map_sectionsN1 = {k: 0 for k in SECTION_NAMES}
        self.map_sectionsN3 = {k: 0 for k in SECTION_NAMES}
        self.map_sectionsN = {k: 0 for k in SECTION_NAMES}
        self.map_offsets = {k: 0 for k in SECTION_NAMES}
        self.map_deg = {k: 0 for k in SECTION_NAMES}
        self.map_total_n = 0
        self.exp2pol = {}
        self.publics = []
        self.ev_idx = {"cm": {}, "const_": {}}

    # -- serde-compatible dump ------------------------------------------------------------------
    def to_dict(self):
        d = {k: getattr(self, k) for k in (
            "var_pol_map", "n_cm1", "n_cm2", "n_cm3", "n_cm4", "n_q", "pu_ctx", "pe_ctx", "ci_ctx", "n_constants", "n_publics",
            "c_exp", "q_deg", "q_dim", "im_exps_list", "qs", "exps_2ns", "exps_n", "ev_map", "fri_exp_id", "n_exps", "cm_n",
            "cm_2ns", "tmpexp_n", "q_2ns", "f_2ns", "map_sections", "map_sectionsN1", "map_sectionsN3", "map_sectionsN",
            "map_offsets", "map_deg", "map_total_n", "publics")}
        d["im_exps"] = {str(k): v for k, v in self.im_exps.items()}
        d["im_exp2cm"] = {str(k): v for k, v in self.im_exp2cm.items()}
        d["exp2pol"] = {str(k): v for k, v in self.exp2pol.items()}
        d["ev_idx"] = {"cm": [[[p, i], v] for (p, i), v in self.ev_idx["cm"].items()],
                       "const_": [[[p, i], v] for (p, i), v in self.ev_idx["const_"].items()]}
        return d

    def _ev_get(self, type_, p, id):
        return self.ev_idx["cm" if type_ == "cm" else "const_"].get((p, id))

    def _ev_set(self, type_, p, id, idx):
        self.ev_idx["cm" if type_ == "cm" else "const_"][(p, id)] = idx


def new_starkinfo(pil, stark_struct, global_l1=None):
    """StarkInfo::new (starkinfo.rs:170-275).  `pil` (from load_pil) is MUTATED exactly like the
    reference mutates its `&mut PIL` (new expressions, nCommitments, nQ, polIdentities)."""
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))
    pil_deg = next(iter(pil["references"].values()))["polDeg"]
    if (1 << stark_struct["nBits"]) != pil_deg:
        raise ValueError("stark_deg != pil_deg")
    if stark_struct["nBitsExt"] != stark_struct["steps"][0]["nBits"]:
        raise ValueError("MustEqualDegreeError: stark_struct.nBitsExt != stark_struct.steps[0].nBits")
    info = StarkInfo()
    info.n_constants = pil["nConstants"]
    info.n_publics = len(pil["publics"])
    program = {"publics_code": [], "step2prev": _segment(), "step3prev": _segment(), "step3": _segment(),
               "step42ns": _segment(), "step52ns": _segment(), "verifier_code": _segment(), "verifier_query_code": _segment()}

    _generate_public_calculators(info, pil, program)
    info.n_cm1 = pil["nCommitments"]
    ctx = Context(); ctx2ns = Context()
    _generate_step2(info, ctx, pil, program)
    _generate_step3(info, ctx, pil, program, global_l1)
    _generate_constraint_polynomial(info, ctx, ctx2ns, pil, stark_struct, program)
    ctx = Context()
    for k, v in info.im_exps.items():
        ctx.calculated[("exps", k)] = v
        ctx.calculated[("expsPrime", k)] = v
    _generate_constraint_polynomial_verifier(info, ctx, pil, program)
    _generate_fri_polynomial(info, ctx2ns, pil, program)
    ctx = Context()
    _generate_fri_verifier(info, ctx, pil, program)
    _map(info, pil, stark_struct, program)
    info.publics = copy.deepcopy(pil["publics"])
    return info, program


def _fix_exp_to_tmp(seg):
    """closure used by generate_public_calculators (starkinfo.rs:291-305)."""
    exp_map = {}
    st = {"tmp_used": seg["tmp_used"]}

    def f(r):
        p = 1 if r["prime"] else 0
        if r["type_"] == "exp":
            if (p, r["id"]) not in exp_map:
                exp_map[(p, r["id"])] = st["tmp_used"]; st["tmp_used"] += 1
            r["prime"] = False
            r["type_"] = "tmp"
            r["id"] = exp_map[(p, r["id"])]
    iterate_code(seg, f)
    seg["tmp_used"] = st["tmp_used"]


def _generate_public_calculators(info, pil, program):
    for p in list(pil["publics"]):
        if p["polType"] == "imP":
            ctx = Context()
            pil_code_gen(ctx, pil, p["polId"], False, "", 0, False)
            seg = build_code(ctx, pil)
            _fix_exp_to_tmp(seg)
            program["publics_code"].append(seg)


def _generate_step2(info, ctx, pil, program):
    """starkinfo.rs:324-408 (plookup f/t folding, h1/h2 commitments)."""
    for pi in list(pil["plookupIdentities"]):
        u = E.challenge("u"); def_val = E.challenge("defVal")
        t_exp = E.nop()
        for j in pi["t"]:
            e = E.exp(j)
            t_exp = e if E.is_nop(t_exp) else E.add(E.mul(u, t_exp), e)
        if pi.get("selT") is not None:
            t_exp = E.sub(t_exp, def_val)
            t_exp = E.mul(t_exp, E.exp(pi["selT"]))
            t_exp = E.add(t_exp, def_val)
            t_exp["idQ"] = pil["nQ"]; pil["nQ"] += 1
        t_exp_id = len(pil["expressions"])
        t_exp["keep"] = True
        pil["expressions"].append(t_exp)

        f_exp = E.nop()
        for j in pi["f"]:
            e = E.exp(j)
            # `f_exp == E::nop()` compares (op,deg,id) only (types.rs:70-74)
            f_exp = e if (f_exp["op"], f_exp["deg"], f_exp["id"]) == ("nop", 0, None) else E.add(E.mul(f_exp, u), e)
        if pi.get("selF") is not None:
            f_exp = E.sub(f_exp, E.exp(t_exp_id))
            f_exp = E.mul(f_exp, E.exp(pi["selF"]))
            f_exp = E.add(f_exp, E.exp(t_exp_id))
            f_exp["idQ"] = pil["nQ"]; pil["nQ"] += 1
        f_exp_id = len(pil["expressions"])
        f_exp["keep"] = True
        pil["expressions"].append(f_exp)

        pil_code_gen(ctx, pil, f_exp_id, False, "", 0, False)
        pil_code_gen(ctx, pil, t_exp_id, False, "", 0, False)
        h1_id = pil["nCommitments"]; pil["nCommitments"] += 1
        h2_id = pil["nCommitments"]; pil["nCommitments"] += 1
        info.pu_ctx.append(_pcctx(f_exp_id=f_exp_id, t_exp_id=t_exp_id, h1_id=h1_id, h2_id=h2_id))
    program["step2prev"] = build_code(ctx, pil)
    ctx.calculated.clear()
    info.n_cm2 = pil["nCommitments"] - info.n_cm1


def _add_identity(pil, e_id):
    pil["polIdentities"].append({"e": e_id, "line": 0, "fileName": ""})


def _generate_step3(info, ctx, pil, program, global_l1):
    """starkinfo_Z.rs."""
    l1name = global_l1 or GLOBAL_L1
    # -- generate_permutation_LC
    for pi in list(pil.get("permutationIdentities") or []):
        u = E.challenge("u"); def_val = E.challenge("defVal")
        t_exp = E.nop()
        for j in pi["t"]:
            e = E.exp(j)
            t_exp = e if E.is_nop(t_exp) else E.add(E.mul(u, t_exp), e)
        if pi.get("selT") is not None:
            t_exp = E.sub(t_exp, def_val); t_exp = E.mul(t_exp, E.exp(pi["selT"])); t_exp = E.add(t_exp, def_val)
            t_exp["idQ"] = pil["nQ"]; pil["nQ"] += 1
        t_exp_id = len(pil["expressions"]); pil["expressions"].append(t_exp)
        f_exp = E.nop()
        for j in pi["f"]:
            e = E.exp(j)
            f_exp = e if E.is_nop(f_exp) else E.add(E.mul(f_exp, u), e)
        if pi.get("selF") is not None:
            f_exp = E.sub(f_exp, def_val); f_exp = E.mul(f_exp, E.exp(pi["selF"])); f_exp = E.add(f_exp, def_val)
            f_exp["idQ"] = pil["nQ"]; pil["nQ"] += 1
        f_exp_id = len(pil["expressions"]); pil["expressions"].append(f_exp)
        info.pe_ctx.append(_pcctx(f_exp_id=f_exp_id, t_exp_id=t_exp_id))

    def l1_expr():
        if l1name not in pil["references"]:
            raise ValueError("%s must be defined" % l1name)
        return E.const_(pil["references"][l1name]["id"])

    one = E.number("1")
    # -- generate_plookup_Z
    for i in range(len(pil["plookupIdentities"])):
        pu = info.pu_ctx[i]
        pu["z_id"] = pil["nCommitments"]; pil["nCommitments"] += 1
        h1 = E.cm(pu["h1_id"]); h2 = E.cm(pu["h2_id"]); h1p = E.cm(pu["h1_id"], True)
        f = E.exp(pu["f_exp_id"]); t = E.exp(pu["t_exp_id"]); tp = E.exp(pu["t_exp_id"], True)
        z = E.cm(pu["z_id"]); zp = E.cm(pu["z_id"], True)
        c1 = E.mul(l1_expr(), E.sub(z, one)); c1["deg"] = 2
        pu["c1_id"] = len(pil["expressions"]); pil["expressions"].append(c1); _add_identity(pil, pu["c1_id"])
        gamma = E.challenge("gamma"); beta = E.challenge("beta")
        num_exp = E.mul(E.mul(E.add(f, gamma), E.add(E.add(t, E.mul(tp, beta)), E.mul(gamma, E.add(one, beta)))), E.add(one, beta))
        num_exp["idQ"] = pil["nQ"]; pil["nQ"] += 1; num_exp["keep"] = True
        pu["num_id"] = len(pil["expressions"]); pil["expressions"].append(num_exp)
        den_exp = E.mul(E.add(E.add(h1, E.mul(h2, beta)), E.mul(gamma, E.add(one, beta))),
                        E.add(E.add(h2, E.mul(h1p, beta)), E.mul(gamma, E.add(one, beta))))
        den_exp["idQ"] = pil["nQ"]; pil["nQ"] += 1
        pu["den_id"] = len(pil["expressions"]); den_exp["keep"] = True; pil["expressions"].append(den_exp)
        c2 = E.sub(E.mul(zp, E.exp(pu["den_id"])), E.mul(z, E.exp(pu["num_id"]))); c2["deg"] = 2
        pu["c2_id"] = len(pil["expressions"]); pil["expressions"].append(c2); _add_identity(pil, pu["c2_id"])
        pil_code_gen(ctx, pil, pu["num_id"], False, "", 0, False)
        pil_code_gen(ctx, pil, pu["den_id"], False, "", 0, False)
    # -- generate_permutation_Z
    for i in range(len(pil.get("permutationIdentities") or [])):
        pe = info.pe_ctx[i]
        pe["z_id"] = pil["nCommitments"]; pil["nCommitments"] += 1
        f = E.exp(pe["f_exp_id"]); t = E.exp(pe["t_exp_id"]); z = E.cm(pe["z_id"]); zp = E.cm(pe["z_id"], True)
        c1 = E.mul(l1_expr(), E.sub(z, one)); c1["deg"] = 2
        pe["c1_id"] = len(pil["expressions"]); pil["expressions"].append(c1); _add_identity(pil, pe["c1_id"])
        beta = E.challenge("beta")
        num_exp = E.add(f, beta); pe["num_id"] = len(pil["expressions"]); num_exp["keep"] = True; pil["expressions"].append(num_exp)
        den_exp = E.add(t, beta); pe["den_id"] = len(pil["expressions"]); den_exp["keep"] = True; pil["expressions"].append(den_exp)
        c2 = E.sub(E.mul(zp, E.exp(pe["den_id"])), E.mul(z, E.exp(pe["num_id"]))); c2["deg"] = 2
        pe["c2_id"] = len(pil["expressions"]); pil["expressions"].append(c2); _add_identity(pil, pe["c2_id"])
        pil_code_gen(ctx, pil, pe["num_id"], False, "", 0, False)
        pil_code_gen(ctx, pil, pe["den_id"], False, "", 0, False)
    # -- generate_connections_Z
    for ci in list(pil.get("connectionIdentities") or []):
        pols = ci["pols"]; conns = ci["connections"]
        cc = _pcctx(z_id=pil["nCommitments"]); pil["nCommitments"] += 1
        gamma = E.challenge("gamma"); beta = E.challenge("beta")
        num_exp = E.add(E.add(E.exp(pols[0]), E.mul(beta, E.x())), gamma)
        den_exp = E.add(E.add(E.exp(pols[0]), E.mul(beta, E.exp(conns[0]))), gamma)
        cc["num_id"] = len(pil["expressions"]); num_exp["keep"] = True; pil["expressions"].append(num_exp)
        cc["den_id"] = len(pil["expressions"]); den_exp["keep"] = True; pil["expressions"].append(den_exp)
        ks = [KS0]
        for _ in range(1, len(pols) - 1):
            ks.append(ks[-1] * KS0 % GL_P)
        for i in range(1, len(pols)):
            num_exp = E.mul(E.exp(cc["num_id"]), E.add(E.add(E.exp(pols[i]), E.mul(E.mul(beta, E.number(ks[i - 1])), E.x())), gamma))
            num_exp["idQ"] = pil["nQ"]; pil["nQ"] += 1
            den_exp = E.mul(E.exp(cc["den_id"]), E.add(E.add(E.exp(pols[i]), E.mul(beta, E.exp(conns[i]))), gamma))
            den_exp["idQ"] = pil["nQ"]; pil["nQ"] += 1
            cc["num_id"] = len(pil["expressions"]); pil["expressions"].append(num_exp)
            cc["den_id"] = len(pil["expressions"]); pil["expressions"].append(den_exp)
        z = E.cm(cc["z_id"]); zp = E.cm(cc["z_id"], True)
        c1 = E.mul(l1_expr(), E.sub(z, one)); c1["deg"] = 2
        cc["c1_id"] = len(pil["expressions"]); pil["expressions"].append(c1); _add_identity(pil, cc["c1_id"])
        c2 = E.sub(E.mul(zp, E.exp(cc["den_id"])), E.mul(z, E.exp(cc["num_id"]))); c2["deg"] = 2
        cc["c2_id"] = len(pil["expressions"]); pil["expressions"].append(c2); _add_identity(pil, cc["c2_id"])
        pil_code_gen(ctx, pil, cc["num_id"], False, "", 0, False)
        pil_code_gen(ctx, pil, cc["den_id"], False, "", 0, False)
        info.ci_ctx.append(cc)
    program["step3prev"] = build_code(ctx, pil)
    ctx.calculated.clear()


def _generate_constraint_polynomial(info, ctx, ctx2ns, pil, stark_struct, program):
    """starkinfo_cp_prover.rs:11-135."""
    vc = E.challenge("vc")
    c_exp = E.nop()
    for pi in pil["polIdentities"]:
        e = E.exp(pi["e"])
        c_exp = E.add(E.mul(vc, c_exp), e) if not E.is_nop(c_exp) else e
    info.q_deg = 0
    max_deg = (1 << (stark_struct["nBitsExt"] - stark_struct["nBits"])) + 1
    for d in range(2, max_deg + 1):
        im_exps, q_deg = calculate_im_pols(pil, c_exp, d)
        if im_exps is not None and (info.q_deg == 0 or (len(im_exps) + q_deg < len(info.im_exps) + info.q_deg)):
            info.q_deg = q_deg
            info.im_exps = im_exps
    info.im_exps_list = sorted(info.im_exps.keys())
    info.im_exp2cm = {}
    for k in info.im_exps_list:
        info.im_exp2cm[k] = pil["nCommitments"]; pil["nCommitments"] += 1
        lhs = copy.deepcopy(pil["expressions"][k])
        rhs = _expr("cm", id=pil["nCommitments"] - 1)
        e = _expr("sub", values=[lhs, rhs])
        c_exp = E.add(E.mul(vc, c_exp), e) if not E.is_nop(c_exp) else e
    info.c_exp = len(pil["expressions"])
    pil["expressions"].append(c_exp)
    info.n_cm3 = pil["nCommitments"] - info.n_cm1 - info.n_cm2
    info.qs = [0] * info.q_deg
    for i in range(info.q_deg):
        info.qs[i] = pil["nCommitments"]; pil["nCommitments"] += 1
    for k in info.im_exps_list:
        pil_code_gen(ctx, pil, k, False, "", 0, False)
    program["step3"] = build_code(ctx, pil)
    for k, v in info.im_exps.items():
        ctx2ns.calculated[("exps", k)] = v
        ctx2ns.calculated[("expsPrime", k)] = v
    pil_code_gen(ctx2ns, pil, info.c_exp, False, "", 0, False)
    code = ctx2ns.code[-1]["code"]
    code.append({"op": "mul", "dest": node("q", 0), "src": [dict(code[-1]["dest"]), node("Zi", 0)]})
    program["step42ns"] = build_code(ctx2ns, pil)
    info.n_cm4 = info.q_deg


def _generate_constraint_polynomial_verifier(info, ctx, pil, program):
    """starkinfo_cp_ver.rs."""
    pil_code_gen(ctx, pil, info.c_exp, False, "", 0, True)
    code = build_code(ctx, pil)
    exp_map = {}
    st = {"tmp_used": code["tmp_used"]}

    def to_eval(r, p):
        if info._ev_get(r["type_"], p, r["id"]) is None:
            info._ev_set(r["type_"], p, r["id"], len(info.ev_map))
            info.ev_map.append(node(r["type_"], r["id"], None, 0, r["prime"]))
        r["prime"] = False
        r["id"] = info._ev_get(r["type_"], p, r["id"])
        r["type_"] = "eval"

    def f(r):
        p = 1 if r["prime"] else 0
        t = r["type_"]
        if t == "exp":
            if r["id"] in info.im_exps_list:
                r["type_"] = "cm"
                r["id"] = info.im_exp2cm[r["id"]]
                to_eval(r, p)
            else:
                if (p, r["id"]) not in exp_map:
                    exp_map[(p, r["id"])] = st["tmp_used"]; st["tmp_used"] += 1
                r["type_"] = "tmp"
                r["exp_id"] = r["id"]
                r["id"] = exp_map[(p, r["id"])]
        elif t in ("cm", "const"):
            to_eval(r, p)
        elif t in ("number", "challenge", "public", "tmp", "Z", "x", "eval"):
            pass
        else:
            raise ValueError("Invalid reference type: %r" % r)
    iterate_code(code, f)
    for i in range(info.q_deg):
        info._ev_set("cm", 0, info.qs[i], len(info.ev_map))
        info.ev_map.append(node("cm", info.qs[i]))
    code["tmp_used"] = st["tmp_used"]
    program["verifier_code"] = code


def _generate_fri_polynomial(info, ctx, pil, program):
    """starkinfo_fri_prover.rs:10-98."""
    vf1 = E.challenge("vf1"); vf2 = E.challenge("vf2")
    fri_exp = E.nop()
    for i in range(pil["nCommitments"]):
        fri_exp = E.cm(i) if E.is_nop(fri_exp) else E.add(E.mul(vf1, fri_exp), E.cm(i))
    fri1 = E.nop(); fri2 = E.nop()
    for i, ev in enumerate(info.ev_map):
        cur = fri2 if ev["prime"] else fri1
        e = {"cm": E.cm, "q": E.q, "const": E.const_}[ev["type_"]](ev["id"])
        cur = E.add(E.mul(cur, vf2), E.sub(e, E.eval(i))) if not E.is_nop(cur) else E.sub(e, E.eval(i))
        if ev["prime"]:
            fri2 = cur
        else:
            fri1 = cur
    if not E.is_nop(fri_exp):       # sic: the reference tests fri_exp here, not fri1_exp
        fri1 = E.mul(fri1, E.xDivXSubXi())
        fri_exp = E.add(E.mul(vf1, fri_exp), fri1) if not E.is_nop(fri_exp) else fri1
    if not E.is_nop(fri2):
        fri2 = E.mul(fri2, E.xDivXSubWXi())
        fri_exp = E.add(E.mul(vf1, fri_exp), fri2) if not E.is_nop(fri_exp) else fri2
    info.fri_exp_id = len(pil["expressions"])
    fri_exp["keep2ns"] = True
    pil["expressions"].append(fri_exp)
    pil_code_gen(ctx, pil, info.fri_exp_id, False, "f", 0, False)
    ctx.code[-1]["code"][-1]["dest"] = node("f", 0)
    program["step52ns"] = build_code(ctx, pil)


def _generate_fri_verifier(info, ctx, pil, program):
    pil_code_gen(ctx, pil, info.fri_exp_id, False, "", 0, True)
    program["verifier_query_code"] = build_code(ctx, pil)
    info.n_exps = len(pil["expressions"])


# ----------------------------------------------------------------------------------------------
# map (starkinfo_map.rs)
# ----------------------------------------------------------------------------------------------
def get_exp_dim(pil, exp):
    """starkinfo_map.rs:517-541 (field-extension dimension, 1 or 3)."""
    op = exp["op"]
    if op in ("add", "sub", "mul", "muladd", "addc", "mulc", "neg"):
        md = 1
        for v in exp["values"]:
            md = max(md, get_exp_dim(pil, v))
        return md
    if op == "cm":
        return pil["cm_dims"][exp["id"]]
    if op == "const":
        return 1
    if op == "exp":
        return get_exp_dim(pil, pil["expressions"][exp["id"]])
    if op == "q":
        return get_exp_dim(pil, pil["expressions"][pil["q2exp"][exp["id"]]])
    if op in ("number", "public", "x"):
        return 1
    if op in ("challenge", "eval", "xDivXSubXi", "xDivXSubWXi"):
        return 3
    raise ValueError("Exp op not defined: %s" % op)


def _map(info, pil, stark_struct, program):
    def add_pol(section, dim):
        info.var_pol_map.append({"section": section, "section_pos": 0, "dim": dim, "exp_id": 0})
        return len(info.var_pol_map) - 1

    tmpexps = {}

    def im_none(i):
        return (i not in info.im_exps) or (not info.im_exps[i])

    def add_tmpexp(exp_id, dim):
        if im_none(exp_id) and exp_id not in tmpexps:
            tmpexps[exp_id] = len(info.tmpexp_n)
            pp = add_pol("tmpexp_n", dim)
            info.tmpexp_n.append(pp); info.map_sections["tmpexp_n"].append(pp)
            info.exp2pol[exp_id] = pp

    def add_cm(sec, dim):
        a = add_pol(sec + "_n", dim); b = add_pol(sec + "_2ns", dim)
        info.cm_n.append(a); info.cm_2ns.append(b)
        info.map_sections[sec + "_n"].append(a); info.map_sections[sec + "_2ns"].append(b)
        return a

    pil["cm_dims"] = [0] * (info.n_cm1 + info.n_cm2 + info.n_cm3 + info.n_cm4)
    for i in range(info.n_cm1):
        add_cm("cm1", 1); pil["cm_dims"][i] = 1
    for i, pu in enumerate(info.pu_ctx):
        dim = max(get_exp_dim(pil, pil["expressions"][pu["f_exp_id"]]), get_exp_dim(pil, pil["expressions"][pu["t_exp_id"]]))
        add_cm("cm2", dim); pil["cm_dims"][info.n_cm1 + i * 2] = dim
        add_cm("cm2", dim); pil["cm_dims"][info.n_cm1 + i * 2 + 1] = dim
        add_tmpexp(pu["f_exp_id"], dim)
        add_tmpexp(pu["t_exp_id"], dim)
    allz = info.pu_ctx + info.pe_ctx + info.ci_ctx
    for i, o in enumerate(allz):
        add_cm("cm3", 3); pil["cm_dims"][info.n_cm1 + info.n_cm2 + i] = 3
        add_tmpexp(o["num_id"], 3)
        add_tmpexp(o["den_id"], 3)
    for i, k in enumerate(info.im_exps_list):
        dim = get_exp_dim(pil, pil["expressions"][k])
        a = add_cm("cm3", dim)
        pil["cm_dims"][info.n_cm1 + info.n_cm2 + i] = dim      # sic (starkinfo_map.rs:172): indexes by i, not nZ+i
        info.exp2pol[k] = a
    info.q_dim = get_exp_dim(pil, pil["expressions"][info.c_exp])
    for i in range(info.q_deg):
        add_cm("cm4", info.q_dim); pil["cm_dims"][info.n_cm1 + info.n_cm2 + info.n_cm3 + i] = info.q_dim
    info.q_2ns.append(add_pol("q_2ns", info.q_dim))
    info.f_2ns.append(add_pol("f_2ns", 3))

    # map_section (starkinfo_map.rs:488-515): dim-1 columns first, then dim-3 ones
    for s in ["cm1_n", "cm1_2ns", "cm2_n", "cm2_2ns", "cm3_n", "cm3_2ns", "cm4_n", "cm4_2ns", "q_2ns", "f_2ns", "tmpexp_n"]:
        p = 0
        for e in (1, 2, 3):
            for pp in info.var_pol_map:
                if pp["section"] == s and pp["dim"] == e:
                    pp["section_pos"] = p; p += e
            if e == 1:
                info.map_sectionsN1[s] = p
            if e == 3:
                info.map_sectionsN[s] = p
        info.map_sectionsN3[s] = (info.map_sectionsN[s] - info.map_sectionsN1[s]) // 3

    N = 1 << stark_struct["nBits"]; Next = 1 << stark_struct["nBitsExt"]
    mo = info.map_offsets; sn = info.map_sectionsN
    mo["cm1_n"] = 0
    mo["cm2_n"] = mo["cm1_n"] + N * sn["cm1_n"]
    mo["cm3_n"] = mo["cm2_n"] + N * sn["cm2_n"]
    mo["cm4_n"] = mo["cm3_n"] + N * sn["cm3_n"]
    mo["tmpexp_n"] = mo["cm4_n"] + N * sn["cm4_n"]
    mo["cm1_2ns"] = mo["tmpexp_n"] + N * sn["tmpexp_n"]
    mo["cm2_2ns"] = mo["cm1_2ns"] + Next * sn["cm1_2ns"]
    mo["cm3_2ns"] = mo["cm2_2ns"] + Next * sn["cm2_2ns"]
    mo["cm4_2ns"] = mo["cm3_2ns"] + Next * sn["cm3_2ns"]
    mo["q_2ns"] = mo["cm4_2ns"] + Next * sn["cm4_2ns"]
    mo["f_2ns"] = mo["q_2ns"] + Next * sn["q_2ns"]
    info.map_total_n = mo["f_2ns"] + Next * sn["f_2ns"]
    for k in SECTION_NAMES:
        info.map_deg[k] = N if k.endswith("_n") else Next

    for seg in program["publics_code"]:
        _fix_prover_code(info, seg, "n", pil, tmpexps)
    _fix_prover_code(info, program["step2prev"], "n", pil, tmpexps)
    _fix_prover_code(info, program["step3prev"], "n", pil, tmpexps)
    _fix_prover_code(info, program["step3"], "n", pil, tmpexps)
    _fix_prover_code(info, program["step42ns"], "2ns", pil, tmpexps)
    _fix_prover_code(info, program["step52ns"], "2ns", pil, tmpexps)
    _fix_prover_code(info, program["verifier_query_code"], "2ns", pil, tmpexps)

    def fix_tree(r):
        if r["type_"] == "cm":
            p1 = info.var_pol_map[info.cm_2ns[r["id"]]]
            r["type_"] = {"cm1_2ns": "tree1", "cm2_2ns": "tree2", "cm3_2ns": "tree3", "cm4_2ns": "tree4"}[p1["section"]]
            r["tree_pos"] = p1["section_pos"]
            r["dim"] = p1["dim"]
    iterate_code(program["verifier_query_code"], fix_tree)

    for i in range(info.n_publics):
        if i < len(program["publics_code"]) and _seg_is_some(program["publics_code"][i]):
            _set_code_dimensions(info, program["publics_code"][i], 1)
    for name in ("step2prev", "step3prev", "step3", "step42ns", "step52ns"):
        _set_code_dimensions(info, program[name], 1)
    _set_code_dimensions(info, program["verifier_code"], 3)
    _set_code_dimensions(info, program["verifier_query_code"], 1)


def _seg_is_some(seg):
    return bool(seg["first"] or seg["i"] or seg["last"])


def _fix_prover_code(info, seg, dom, pil, tmpexps):
    exp_map = {}
    st = {"tmp_used": seg["tmp_used"]}

    def f(r):
        t = r["type_"]
        if t == "cm":
            r["p"] = info.cm_n[r["id"]] if dom == "n" else info.cm_2ns[r["id"]]
        elif t == "exp":
            if r["id"] in info.im_exps_list:
                r["type_"] = "cm"
                r["id"] = info.im_exp2cm[r["id"]]
            elif r["id"] in tmpexps and dom == "n":
                r["type_"] = "tmpExp"
                r["dim"] = get_exp_dim(pil, pil["expressions"][r["id"]])
                r["id"] = tmpexps[r["id"]]
            else:
                p = 1 if r["prime"] else 0
                if (p, r["id"]) not in exp_map:
                    exp_map[(p, r["id"])] = st["tmp_used"]; st["tmp_used"] += 1
                r["type_"] = "tmp"
                r["exp_id"] = r["id"]
                r["id"] = exp_map[(p, r["id"])]
        elif t in ("const", "number", "challenge", "public", "tmp", "Zi", "xDivXSubXi", "xDivXSubWXi", "eval", "x", "q", "f", "tmpExp"):
            pass
        else:
            raise ValueError("Invalid reference type %s" % t)
    iterate_code(seg, f)
    seg["tmp_used"] = st["tmp_used"]


def _set_code_dimensions(info, seg, dim_x):
    tmp_dim = {}

    def get_dim(r):
        t = r["type_"]
        if t == "tmp":
            d = tmp_dim[r["id"]]
        elif t in ("tree1", "tree2", "tree3", "tree4", "tmpExp"):
            d = r["dim"]
        elif t == "cm":
            d = info.var_pol_map[info.cm_2ns[r["id"]]]["dim"]
        elif t == "q":
            d = info.var_pol_map[info.qs[r["id"]]]["dim"]
        elif t in ("const", "number", "public", "Zi"):
            d = 1
        elif t in ("eval", "challenge", "Z"):
            d = 3
        elif t in ("xDivXSubXi", "xDivXSubWXi", "x"):
            d = dim_x
        else:
            raise ValueError("Invalid reference type get %s" % t)
        if d == 0:
            raise ValueError("Invalid dim")
        r["dim"] = d
        return d

    def set_dim(r, dim):
        if r["type_"] == "tmp":
            tmp_dim[r["id"]] = dim
            r["dim"] = dim
        elif r["type_"] in ("exp", "cm", "q", "tmpExp", "f"):
            r["dim"] = dim
        else:
            raise ValueError("Invalid reference type set %s" % r["type_"])

    for part in ("first", "i", "last"):
        for c in seg[part]:
            if c["op"] in ("add", "sub", "mul"):
                nd = max(get_dim(c["src"][0]), get_dim(c["src"][1]))
            elif c["op"] == "muladd":
                nd = max(max(get_dim(c["src"][0]), get_dim(c["src"][1])), get_dim(c["src"][2]))
            elif c["op"] == "copy":
                nd = get_dim(c["src"][0])
            else:
                raise ValueError("Invalid op: %s" % c["op"])
            set_dim(c["dest"], nd)


# ----------------------------------------------------------------------------------------------
def setup_json(info, program, stark_struct):
    """The JSON blob the C-ABI consumes: {"starkinfo": serde(StarkInfo), "program": serde(Program),
    "stark_struct": serde(StarkStruct)}."""
    return json.dumps({"starkinfo": info.to_dict(), "program": program, "stark_struct": stark_struct}, separators=(",", ":"))
