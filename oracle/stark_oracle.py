"""CPU ORACLE of `StarkSetup::new`, `StarkProof::stark_gen`, `FRI::prove`, `stark_verify`, `FRI::verify`
(GL, BN128 and BLS12381 hash back-ends) -- TEST INFRASTRUCTURE, not product code.

Restates starky/src/stark_setup.rs:27-66, stark_gen.rs:193-557 & 575-963, fri.rs:84-297,
transcript.rs:8-103, stark_verify.rs:21-213 and serializer.rs:137-270 on top of the C primitives in
gl_oracle.c.  The heavy O(N) loops run in C/OpenMP (so the same code doubles as the timed CPU baseline);
the protocol skeleton is Python.  The step programs come from eigen_zkvm_b200.starkinfo (the host-side
port of the reference's PIL codegen, which sits on the caller's side of the stark_gen boundary).

Pinning: the primitives are pinned to the reference KATs (tests/test_oracle_kats.py); a full proof has no
golden bytes in the reference -- its own tests only assert `stark_verify == true`
(stark_gen.rs:1176-1194).  We assert the same through the restated verifier, reproduce the const-root
KAT (stark_setup.rs:100-116) and the self-verified roots listed in SURVEY.md Appendix C.
"""
import ctypes
import json
import time
import numpy as np
from . import gl
from .gl import P, f3_add, f3_sub, f3_mul, f3_muls, f3_inv, f3_div, f3_pow

SEC_ORDER = ["cm1_n", "cm2_n", "cm3_n", "cm4_n", "tmpexp_n", "const_n", "cm1_2ns", "cm2_2ns", "cm3_2ns", "cm4_2ns",
             "const_2ns", "q_2ns", "f_2ns", "xDivXSubXi", "xDivXSubWXi"]
SEC_IDX = {n: i for i, n in enumerate(SEC_ORDER)}


class _Sec(ctypes.Structure):
    _fields_ = [("base", ctypes.c_void_p), ("width", ctypes.c_size_t)]


def parse_pil_number(s):            # types.rs:221-233
    v = int(s, 16) if s.startswith("0x") else int(s)
    return v % P


# ---------------------------------------------------------------------------------------------
class TranscriptGL:                 # transcript.rs:8-103
    def __init__(self):
        self.state = [0, 0, 0, 0]; self.pending = []; self.out = []

    def _update(self):
        while len(self.pending) < 8:
            self.pending.append(0)
        self.out = gl.poseidon(self.pending, self.state)
        self.pending = []
        self.state = self.out[0:4]

    def put(self, elems):
        for e in elems:
            self.out = []
            self.pending.append(int(e) % P)
            if len(self.pending) == 8:
                self._update()

    def get_fields1(self):
        while not self.out:
            self._update()
        return self.out.pop(0)

    def get_field(self):
        return (self.get_fields1(), self.get_fields1(), self.get_fields1())

    def get_permutations(self, n, nbits):
        total = n * nbits
        nf = (total - 1) // 63 + 1
        fields = [self.get_fields1() for _ in range(nf)]
        res = []; cf = 0; cb = 0
        for _ in range(n):
            a = 0
            for j in range(nbits):
                if (fields[cf] >> cb) & 1:
                    a += 1 << j
                cb += 1
                if cb == 63:
                    cb = 0; cf += 1
            res.append(a)
        return res


class TranscriptBig:                # transcript_bn128.rs:14-135, transcript_bls12381.rs (same code over the other field)
    """Poseidon sponge over the BN128 / BLS12-381 scalar field: rate 16, state = out[0]; every 254-bit output yields three
    64-bit limbs reduced mod p_GL (helper.rs:61-65); query indices take 253 bits per output."""
    def __init__(self, field):
        from oracle import poseidon_big as pb
        self.pb, self.field = pb, field
        self.state = 0; self.pending = []; self.out = []; self.out3 = []

    def _update(self):
        while len(self.pending) < 16:
            self.pending.append(0)
        self.out = self.pb.permute(self.field, self.pending, self.state)      # hash_ex(.., 17)
        self.out3 = []
        self.pending = []
        self.state = self.out[0]

    def put(self, elems):               # each element: a GL value or a digest (field element), transcript_bn128.rs:89-102
        for e in elems:
            self.out = []               # NOTE: out3 is not cleared here, exactly like add_1 (transcript_bn128.rs:33-40)
            self.pending.append(int(e) % self.pb.MOD[self.field])
            if len(self.pending) == 16:
                self._update()

    def get_fields1(self):
        while True:
            if self.out3:
                return self.out3.pop(0)
            if self.out:
                v = self.out.pop(0)
                self.out3 += [(v & 0xFFFFFFFFFFFFFFFF) % P, ((v >> 64) & 0xFFFFFFFFFFFFFFFF) % P, ((v >> 128) & 0xFFFFFFFFFFFFFFFF) % P]
                continue
            self._update()

    def get_field(self):
        return (self.get_fields1(), self.get_fields1(), self.get_fields1())

    def _get_fields253(self):
        while not self.out:
            self._update()
        return self.out.pop(0)

    def get_permutations(self, n, nbits):
        total = n * nbits
        nf = (total - 1) // 253 + 1
        fields = [self._get_fields253() for _ in range(nf)]
        res = []; cf = 0; cb = 0
        for _ in range(n):
            a = 0
            for j in range(nbits):
                if (fields[cf] >> cb) & 1:
                    a += 1 << j
                cb += 1
                if cb == 253:
                    cb = 0; cf += 1
            res.append(a)
        return res


class TreeBig:
    """MerkleTreeBN128 / MerkleTreeBLS12381 (merklehash_bn128.rs): 16-ary, digests are field elements; root() is a
    one-element list so that transcripts and serializers treat it like any other element list."""
    def __init__(self, field, elements, width, height):
        from oracle import poseidon_big as pb
        self.pb, self.field, self.width, self.height = pb, field, width, height
        self.elements = np.ascontiguousarray(elements, dtype=np.uint64).reshape(-1)
        rows = self.elements.reshape(height, width).tolist() if width else [[] for _ in range(height)]
        if width == 0:
            # empty buffer: leaves stay zero digests, levels are still hashed (merklehash_bn128.rs:191-224)
            self.nodes = [0] * pb.get_n_nodes(height)
            n = height; nn = (n - 1) // 16 + 1; p_in = 0; p_out = nn * 16
            while n > 1:
                cache = {}
                for i in range(nn):
                    ch = tuple(self.nodes[p_in + 16 * i: p_in + 16 * i + 16])
                    if ch not in cache: cache[ch] = pb.hash(field, list(ch), 0)
                    self.nodes[p_out + i] = cache[ch]
                n = nn; nn = (n - 1) // 16 + 1; p_in = p_out; p_out = p_in + nn * 16
        else:
            self.nodes = pb.merkelize(field, rows)

    def root(self):
        return [self.nodes[-1]]

    def group_proof(self, idx):         # merklehash_bn128.rs:89-106,226-243
        vals = [int(x) for x in self.elements[idx * self.width:(idx + 1) * self.width]]
        mp = []; n = self.height; off = 0
        while n > 1:
            si = idx & ~0xF
            mp.append(list(self.nodes[off + si: off + si + 16]))
            nn = (n - 1) // 16 + 1
            off += nn * 16; idx >>= 4; n = nn
        return vals, mp


def _big_field(hash_type):
    return {"BN128": "bn128", "BLS12381": "bls12381"}[hash_type]


def make_tree(hash_type, elements, width, height):
    return Tree(elements, width, height) if hash_type == "GL" else TreeBig(_big_field(hash_type), elements, width, height)


def make_transcript(hash_type):
    return TranscriptGL() if hash_type == "GL" else TranscriptBig(_big_field(hash_type))


def verify_group_proof_any(hash_type, root, sibs, idx, vals):
    if hash_type == "GL":
        return verify_group_proof(root, sibs, idx, vals)
    from oracle import poseidon_big as pb
    field = _big_field(hash_type)
    cur = pb.hash_element_array(field, [int(v) for v in vals])        # merklehash_bn128.rs:108-129,245-254
    for lvl in sibs:
        lvl = [int(x) for x in lvl]
        if len(lvl) != 16 or lvl[idx & 15] != cur:
            return False
        cur = pb.hash(field, lvl, 0); idx >>= 4
    return [cur] == [int(x) for x in root]


class Tree:
    """MerkleTreeGL (merklehash.rs): keeps elements (row-major) + nodes."""
    def __init__(self, elements, width, height):
        self.width, self.height = width, height
        self.elements = np.ascontiguousarray(elements, dtype=np.uint64).reshape(-1)
        self.nodes = gl.merkelize(self.elements, width, height)

    def root(self):
        return [int(x) for x in self.nodes[-1]]

    def group_proof(self, idx):
        vals = [int(x) for x in self.elements[idx * self.width:(idx + 1) * self.width]]
        sib = gl.merkle_proof(self.nodes, self.height, idx)
        return vals, [[int(x) for x in r] for r in sib]


def verify_group_proof(root, sibs, idx, vals):
    s = np.array(sibs, dtype=np.uint64).reshape(-1, 4) if len(sibs) else np.zeros((0, 4), dtype=np.uint64)
    return gl.merkle_root_from_proof(np.array(vals, dtype=np.uint64), s, idx) == list(root)


# ---------------------------------------------------------------------------------------------
class Ctx:
    pass


def _compile(prog_first, info, dom, n_consts):
    """interpreter.rs:183-283 (compile_code/get_ref/set_ref) -> flat int64 words for ora_eval_program."""
    consts = []

    def cidx(v):
        consts.append(v % P); return len(consts) - 1

    def mem_from_pol(pol_id, prime):
        p = info.var_pol_map[pol_id]
        return [1, SEC_IDX[p["section"]], p["section_pos"], 1 if prime else 0, p["dim"]]

    def ref(r, publics):
        t = r["type_"]
        if t == "tmp":
            return [0, r["id"], 0, 0, 0]
        if t == "const":
            return [1, SEC_IDX["const_n" if dom == "n" else "const_2ns"], r["id"], 1 if r["prime"] else 0, 1]
        if t == "cm":
            return mem_from_pol(info.cm_n[r["id"]] if dom == "n" else info.cm_2ns[r["id"]], r["prime"])
        if t == "tmpExp":
            assert dom == "n"
            return mem_from_pol(info.tmpexp_n[r["id"]], r["prime"])
        if t == "number":
            return [2, cidx(parse_pil_number(r["value"])), 0, 0, 1]
        if t == "public":
            return [2, cidx(publics[r["id"]]), 0, 0, 1]
        if t == "challenge":
            return [3, r["id"], 0, 0, 3]
        if t == "eval":
            return [3, 8 + r["id"], 0, 0, 3]
        if t == "xDivXSubXi":
            return [1, SEC_IDX["xDivXSubXi"], 0, 0, 3]
        if t == "xDivXSubWXi":
            return [1, SEC_IDX["xDivXSubWXi"], 0, 0, 3]
        if t == "x":
            return [4, 0, 0, 0, 1]
        if t == "Zi":
            return [5, 0, 0, 0, 1]
        if t == "q":
            assert dom == "2ns"
            return [1, SEC_IDX["q_2ns"], r["id"], 0, info.q_dim]
        if t == "f":
            assert dom == "2ns"
            return [1, SEC_IDX["f_2ns"], r["id"], 0, 3]
        raise ValueError("Invalid reference type %s" % t)

    def build(publics):
        words = []
        for c in prog_first:
            op = {"add": 0, "sub": 1, "mul": 2, "copy": 3}[c["op"]]
            srcs = [ref(s, publics) for s in c["src"]]
            while len(srcs) < 2:
                srcs.append([0, 0, 0, 0, 0])
            words += [op] + ref(c["dest"], publics) + srcs[0] + srcs[1]
        return np.array(words, dtype=np.int64), np.array(consts if consts else [0], dtype=np.uint64)
    return build


def _run_program(ctx, info, seg, dom):
    """calculate_exps_parallel (stark_gen.rs:786-963): run seg.first on every row of the domain."""
    if not seg["first"]:
        return
    words, consts = _compile(seg["first"], info, dom, info.n_constants)(ctx.publics)
    n = ctx.N if dom == "n" else ctx.Next
    nxt = 1 if dom == "n" else (1 << ctx.ext_bits)
    secs = (_Sec * len(SEC_ORDER))()
    for i, name in enumerate(SEC_ORDER):
        arr, width = ctx.sec.get(name, (np.zeros(0, dtype=np.uint64), 0))
        secs[i].base = arr.ctypes.data if arr.size else 0
        secs[i].width = width
    f3c = np.zeros((8 + max(1, len(ctx.evals))) * 3, dtype=np.uint64)
    for i, c in enumerate(ctx.challenge):
        f3c[3 * i:3 * i + 3] = c
    for i, e in enumerate(ctx.evals):
        f3c[3 * (8 + i):3 * (8 + i) + 3] = e
    x = ctx.x_n if dom == "n" else ctx.x_2ns
    gl.lib().ora_eval_program(words.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(seg["first"])), ctypes.c_size_t(seg["tmp_used"]),
                              secs, consts.ctypes.data_as(ctypes.c_void_p), f3c.ctypes.data_as(ctypes.c_void_p),
                              x.ctypes.data_as(ctypes.c_void_p), ctx.zi.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(ctx.zi.size),
                              ctypes.c_size_t(n), ctypes.c_size_t(nxt))


def _calculate_exp_at_point(ctx, info, seg, idx):
    """StarkProof::calculate_exp_at_point (stark_gen.rs:559-572): run a (base-field) public calculator at one row of the
    "n" domain and return the value of its last operation (compile_code(.., ret = true), interpreter.rs:183-215)."""
    N = ctx.N
    tmp = {}

    def val(r):
        t = r["type_"]
        if t == "tmp": return tmp[r["id"]]
        if t == "number": return parse_pil_number(r["value"])
        if t == "public": return ctx.publics[r["id"]]
        row = (idx + (1 if r["prime"] else 0)) % N
        if t == "const":
            return int(ctx.sec["const_n"][0][row * info.n_constants + r["id"]])
        if t == "cm":
            pm = info.var_pol_map[info.cm_n[r["id"]]]
            if pm["dim"] != 1: raise NotImplementedError("extension-field public calculators")
            arr, w = ctx.sec[pm["section"]]
            return int(arr[row * w + pm["section_pos"]])
        raise NotImplementedError("public calculator operand " + t)

    res = None
    for op in seg["first"]:
        a = [val(x) for x in op["src"]]
        o = op["op"]
        res = a[0] if o == "copy" else (a[0] + a[1]) % P if o == "add" else (a[0] - a[1]) % P if o == "sub" else a[0] * a[1] % P
        if op["dest"]["type_"] != "tmp": raise NotImplementedError("public calculator destination")
        tmp[op["dest"]["id"]] = res
    return res


def _x_table(n, start, w):
    out = np.zeros(n, dtype=np.uint64)
    gl.lib().ora_x_table(out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n), ctypes.c_uint64(start), ctypes.c_uint64(w))
    return out


# ---------------------------------------------------------------------------------------------
def stark_setup(const_rowmajor, pil, stark_struct):
    """StarkSetup::new (stark_setup.rs:27-66). Returns dict(const_tree, const_root, starkinfo, program)."""
    from eigen_zkvm_b200 import starkinfo as si
    nb, nbe = stark_struct["nBits"], stark_struct["nBitsExt"]
    nc = pil["nConstants"]
    ext = gl.lde(const_rowmajor, nc, nb, nbe)
    tree = make_tree(stark_struct["verificationHashType"], ext, nc, 1 << nbe)
    info, program = si.new_starkinfo(pil, stark_struct)
    return {"const_tree": tree, "const_root": tree.root(), "starkinfo": info, "program": program}


def _get_pol(ctx, info, pol_id):
    p = info.var_pol_map[pol_id]
    arr, width = ctx.sec[p["section"]]
    a = arr.reshape(-1, width)
    if p["dim"] == 1:
        return [(int(v), 0, 0) for v in a[:, p["section_pos"]]], 1
    return [tuple(int(x) for x in r) for r in a[:, p["section_pos"]:p["section_pos"] + 3]], 3


def _set_pol(ctx, info, pol_id, vals):
    p = info.var_pol_map[pol_id]
    arr, width = ctx.sec[p["section"]]
    a = arr.reshape(-1, width)
    if p["dim"] == 1:
        a[:, p["section_pos"]] = np.array([v[0] for v in vals], dtype=np.uint64)
    else:
        a[:, p["section_pos"]:p["section_pos"] + 3] = np.array(vals, dtype=np.uint64)


def _calculate_H1H2(f, t):          # stark_gen.rs:625-651
    idx_t = {}
    s = []
    for i, e in enumerate(t):
        idx_t[e] = i; s.append((e, i))
    for e in f:
        if e not in idx_t:
            raise ValueError("Number not included: %r" % (e,))
        s.append((e, idx_t[e]))
    s.sort(key=lambda a: a[1])      # python sort is stable, like Rust sort_by
    return [s[2 * i][0] for i in range(len(f))], [s[2 * i + 1][0] for i in range(len(f))]


def _calculate_Z(num, den):         # stark_gen.rs:653-666
    n = len(num)
    d = np.array(den, dtype=np.uint64)
    di = gl.f3_batch_inverse(d).reshape(-1, 3)
    z = [(1, 0, 0)]
    for i in range(1, n):
        z.append(f3_mul(z[i - 1], f3_mul(num[i - 1], tuple(int(x) for x in di[i - 1]))))
    chk = f3_mul(z[n - 1], f3_mul(num[n - 1], tuple(int(x) for x in di[n - 1])))
    assert chk == (1, 0, 0), "calculate_Z wrap-around check failed"
    return z


def stark_gen(cm_rowmajor, const_rowmajor, setup, stark_struct, timings=None):
    """StarkProof::stark_gen::<TranscriptGL> for MerkleTreeGL (stark_gen.rs:193-557)."""
    info, program, const_tree = setup["starkinfo"], setup["program"], setup["const_tree"]
    L = gl.lib()
    T0 = time.perf_counter()
    tm = {}

    def tick(name, t0):
        tm[name] = tm.get(name, 0.0) + (time.perf_counter() - t0)

    ctx = Ctx()
    ctx.nbits, ctx.nbits_ext = stark_struct["nBits"], stark_struct["nBitsExt"]
    ctx.N, ctx.Next = 1 << ctx.nbits, 1 << ctx.nbits_ext
    ctx.ext_bits = ctx.nbits_ext - ctx.nbits
    N, Next = ctx.N, ctx.Next
    sn = info.map_sectionsN
    z = lambda rows, w: np.zeros(rows * w, dtype=np.uint64)
    ctx.sec = {
        "cm1_n": (np.ascontiguousarray(cm_rowmajor, dtype=np.uint64).reshape(-1).copy(), sn["cm1_n"]),
        "cm2_n": (z(N, sn["cm2_n"]), sn["cm2_n"]), "cm3_n": (z(N, sn["cm3_n"]), sn["cm3_n"]), "cm4_n": (z(0, 0), sn["cm4_n"]),
        "tmpexp_n": (z(N, sn["tmpexp_n"]), sn["tmpexp_n"]),
        "const_n": (np.ascontiguousarray(const_rowmajor, dtype=np.uint64).reshape(-1), info.n_constants),
        "const_2ns": (const_tree.elements, info.n_constants),
        "q_2ns": (z(Next, info.q_dim), info.q_dim), "f_2ns": (z(Next, 3), 3),
        "xDivXSubXi": (z(0, 0), 3), "xDivXSubWXi": (z(0, 0), 3),
    }
    assert ctx.sec["cm1_n"][0].size == N * sn["cm1_n"]
    ctx.challenge = [(0, 0, 0)] * 8
    ctx.evals = []
    t0 = time.perf_counter()
    ctx.x_n = _x_table(N, 1, gl.root(ctx.nbits))
    ctx.x_2ns = _x_table(Next, gl.SHIFT, gl.root(ctx.nbits_ext))
    ctx.zi = np.zeros(1 << ctx.ext_bits, dtype=np.uint64)
    L.ora_zh_inv(ctx.zi.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint(ctx.nbits), ctypes.c_uint(ctx.ext_bits))
    tick("tables", t0)

    # publics (stark_gen.rs:256-270)
    ctx.publics = []
    pc = 0
    for pe in info.publics:
        if pe["polType"] == "cmP":
            ctx.publics.append(int(ctx.sec["cm1_n"][0][pe["idx"] * sn["cm1_n"] + pe["polId"]]))
        elif pe["polType"] == "imP":
            ctx.publics.append(_calculate_exp_at_point(ctx, info, program["publics_code"][len(ctx.publics)], pe["idx"]))
        else:
            raise ValueError("Invalid public type")
    hash_type = stark_struct["verificationHashType"]
    tr = make_transcript(hash_type)
    for p in ctx.publics:
        tr.put([p])

    def extend_and_merkelize(name):         # stark_gen.rs:710-732
        t0 = time.perf_counter()
        arr, w = ctx.sec[name + "_n"]
        ext = gl.lde(arr, w, ctx.nbits, ctx.nbits_ext)
        tick("lde", t0); t0 = time.perf_counter()
        tree = make_tree(hash_type, ext, w, Next)
        tick("merkle", t0)
        ctx.sec[name + "_2ns"] = (tree.elements, w)
        return tree

    n_cm = info.n_cm1
    tree1 = extend_and_merkelize("cm1")
    tr.put(tree1.root())
    ctx.challenge[0] = tr.get_field(); ctx.challenge[1] = tr.get_field()
    t0 = time.perf_counter(); _run_program(ctx, info, program["step2prev"], "n"); tick("eval_n", t0)
    for pu in info.pu_ctx:
        f, _ = _get_pol(ctx, info, info.exp2pol[pu["f_exp_id"]]); t, _ = _get_pol(ctx, info, info.exp2pol[pu["t_exp_id"]])
        h1, h2 = _calculate_H1H2(f, t)
        _set_pol(ctx, info, info.cm_n[n_cm], h1); n_cm += 1
        _set_pol(ctx, info, info.cm_n[n_cm], h2); n_cm += 1
    tree2 = extend_and_merkelize("cm2")
    tr.put(tree2.root())
    ctx.challenge[2] = tr.get_field(); ctx.challenge[3] = tr.get_field()
    t0 = time.perf_counter(); _run_program(ctx, info, program["step3prev"], "n"); tick("eval_n", t0)
    for o in info.pu_ctx + info.pe_ctx + info.ci_ctx:
        num, _ = _get_pol(ctx, info, info.exp2pol[o["num_id"]]); den, _ = _get_pol(ctx, info, info.exp2pol[o["den_id"]])
        _set_pol(ctx, info, info.cm_n[n_cm], _calculate_Z(num, den)); n_cm += 1
    t0 = time.perf_counter(); _run_program(ctx, info, program["step3"], "n"); tick("eval_n", t0)
    tree3 = extend_and_merkelize("cm3")
    tr.put(tree3.root())
    ctx.challenge[4] = tr.get_field()
    t0 = time.perf_counter(); _run_program(ctx, info, program["step42ns"], "2ns"); tick("eval_q", t0)

    # quotient split (stark_gen.rs:375-396)
    t0 = time.perf_counter()
    qd, qg = info.q_dim, info.q_deg
    qq1 = gl.intt(ctx.sec["q_2ns"][0], qd, ctx.nbits_ext)
    qq2 = np.zeros(Next * qd * qg, dtype=np.uint64)
    if qg > 0:
        L.ora_quotient_split(qq1.ctypes.data_as(ctypes.c_void_p), qq2.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(N), ctypes.c_size_t(Next),
                             ctypes.c_size_t(qd), ctypes.c_size_t(qg), ctypes.c_uint(ctx.nbits))
        cm4 = gl.ntt(qq2, qd * qg, ctx.nbits_ext)
    else:
        cm4 = qq2
    tick("quotient", t0); t0 = time.perf_counter()
    tree4 = make_tree(hash_type, cm4, sn["cm4_2ns"], Next)
    tick("merkle", t0)
    ctx.sec["cm4_2ns"] = (tree4.elements, sn["cm4_2ns"])
    tr.put(tree4.root())
    ctx.challenge[7] = tr.get_field()       # xi

    # evaluations (stark_gen.rs:416-466)
    t0 = time.perf_counter()
    xi = ctx.challenge[7]
    shift_inv = gl.inv(gl.SHIFT)
    w_n = gl.root(ctx.nbits)
    xis = f3_muls(xi, shift_inv)
    wxis = f3_muls(f3_muls(xi, w_n), shift_inv)

    def lev(base):
        pw = np.zeros(N * 3, dtype=np.uint64)
        L.ora_f3_powers(np.array(base, dtype=np.uint64).ctypes.data_as(ctypes.c_void_p), pw.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(N))
        return gl.intt(pw, 3, ctx.nbits)
    LEv, LpEv = lev(xis), lev(wxis)
    ctx.evals = []
    for ev in info.ev_map:
        if ev["type_"] == "const":
            buf, size, off, dim = ctx.sec["const_2ns"][0], info.n_constants, ev["id"], 1
        elif ev["type_"] == "cm":
            p = info.var_pol_map[info.cm_2ns[ev["id"]]]
            buf, size = ctx.sec[p["section"]]
            off, dim = p["section_pos"], p["dim"]
        else:
            raise ValueError("Invalid ev type")
        lv = LpEv if ev["prime"] else LEv
        out = np.zeros(3, dtype=np.uint64)
        L.ora_eval_dot(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(size), ctypes.c_size_t(off), ctypes.c_int(dim), ctypes.c_uint(ctx.ext_bits),
                       lv.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(N), out.ctypes.data_as(ctypes.c_void_p))
        ctx.evals.append(tuple(int(x) for x in out))
    tick("evals", t0)
    for e in ctx.evals:
        tr.put(list(e))
    ctx.challenge[5] = tr.get_field(); ctx.challenge[6] = tr.get_field()

    # xDivXSubXi tables (stark_gen.rs:481-522)
    t0 = time.perf_counter()
    wxi = f3_muls(xi, w_n)
    for name, pt in (("xDivXSubXi", xi), ("xDivXSubWXi", wxi)):
        out = np.zeros(Next * 3, dtype=np.uint64)
        L.ora_xdivxsub(ctx.x_2ns.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(Next), np.array(pt, dtype=np.uint64).ctypes.data_as(ctypes.c_void_p),
                       out.ctypes.data_as(ctypes.c_void_p))
        ctx.sec[name] = (out, 3)
    tick("xdivxsub", t0)
    t0 = time.perf_counter(); _run_program(ctx, info, program["step52ns"], "2ns"); tick("eval_f", t0)

    trees = [tree1, tree2, tree3, tree4, const_tree]
    t0 = time.perf_counter()
    fri = fri_prove(tr, ctx.sec["f_2ns"][0], stark_struct, lambda idx: [t.group_proof(idx) for t in trees], tm)
    tick("fri", t0)
    proof = {"rootC": const_tree.root(), "root1": tree1.root(), "root2": tree2.root(), "root3": tree3.root(), "root4": tree4.root(),
             "evals": ctx.evals, "publics": ctx.publics, "fri": fri}
    tm["total"] = time.perf_counter() - T0
    if timings is not None:
        timings.update(tm)
    return proof


def fri_prove(tr, pol, stark_struct, query_pol, tm=None):
    """FRI::prove (fri.rs:84-184). pol: Next x 3 u64 (AoS)."""
    L = gl.lib()
    steps = [s["nBits"] for s in stark_struct["steps"]]
    pol = np.ascontiguousarray(pol, dtype=np.uint64).reshape(-1)
    pol_bits = stark_struct["nBitsExt"]
    assert pol.size == 3 << pol_bits
    shift_inv = gl.inv(gl.SHIFT)
    trees = []
    queries = [{"root": None, "pol_queries": []} for _ in steps]
    for si, nb in enumerate(steps):
        red = pol_bits - nb
        pol2_n = 1 << (pol_bits - red)
        special_x = tr.get_field()
        if si == 0:
            pol2 = pol[:3 * pol2_n].copy()
        else:
            pol2 = np.zeros(3 * pol2_n, dtype=np.uint64)
            L.ora_fri_fold(pol.ctypes.data_as(ctypes.c_void_p), pol2.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint(pol_bits), ctypes.c_uint(red),
                           ctypes.c_uint64(shift_inv), np.array(special_x, dtype=np.uint64).ctypes.data_as(ctypes.c_void_p))
        if si < len(steps) - 1:
            n_groups = 1 << steps[si + 1]
            group_size = (1 << nb) // n_groups
            tb = np.zeros(pol2.size, dtype=np.uint64)
            L.ora_fri_transpose(pol2.ctypes.data_as(ctypes.c_void_p), tb.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(pol2_n), ctypes.c_uint(steps[si + 1]))
            t = make_tree(stark_struct["verificationHashType"], tb, 3 * group_size, n_groups)
            trees.append(t)
            queries[si + 1]["root"] = t.root()
            tr.put(t.root())
        else:
            for v in pol2:
                tr.put([int(v)])
        pol = pol2
        pol_bits -= red
        for _ in range(red):
            shift_inv = shift_inv * shift_inv % P
    last = [tuple(int(x) for x in pol[3 * i:3 * i + 3]) for i in range(pol.size // 3)]
    ys = tr.get_permutations(stark_struct["nQueries"], steps[0])
    for si in range(len(steps)):
        for y in ys:
            if si == 0:
                queries[si]["pol_queries"].append(query_pol(y))
            else:
                queries[si]["pol_queries"].append([trees[si - 1].group_proof(y)])
        if si < len(steps) - 1:
            ys = [y % (1 << steps[si + 1]) for y in ys]
    return {"queries": queries, "last": last}


# ---------------------------------------------------------------------------------------------
def _digest_json(d):                # digest.rs:84-111
    d = [int(x) for x in d]
    if len(d) == 1:                 # BN128 / BLS12-381 digest: the scalar as a decimal string
        return str(d[0])
    if d[1] == 0 and d[2] == 0 and d[3] == 0:
        return str(d[0])
    return [str(x) for x in d]


def proof_to_json(proof, prover_addr=None):
    """serde_json::to_string(&StarkProof<M>) (serializer.rs:137-270); compact, insertion ordered.  `prover_addr` is
    serialized (last) for the non-GL back-ends only (serializer.rs:262-266)."""
    q = proof["fri"]["queries"]
    o = {}
    o["rootC"] = _digest_json(proof["rootC"])
    for k in ("root1", "root2", "root3", "root4"):
        o[k] = _digest_json(proof[k])
    o["evals"] = [[str(x) for x in e] for e in proof["evals"]]
    sib_json = lambda sibs: [[str(x) for x in lvl] for lvl in sibs]     # from_basefield(lane) -> [lane,0,0,0] -> one string
    for i in range(1, len(q)):
        o["s%d_root" % i] = _digest_json(q[i]["root"])
        o["s%d_vals" % i] = [[str(x) for x in pq[0][0]] for pq in q[i]["pol_queries"]]
        o["s%d_siblings" % i] = [sib_json(pq[0][1]) for pq in q[i]["pol_queries"]]
    names = ["1", "2", "3", "4", "C"]
    vals = {n: [[str(x) for x in pq[j][0]] for pq in q[0]["pol_queries"]] for j, n in enumerate(names)}
    sibs = {n: [sib_json(pq[j][1]) for pq in q[0]["pol_queries"]] for j, n in enumerate(names)}
    o["s0_vals1"] = vals["1"]
    if vals["2"]: o["s0_vals2"] = vals["2"]
    if vals["3"]: o["s0_vals3"] = vals["3"]
    o["s0_vals4"] = vals["4"]; o["s0_valsC"] = vals["C"]
    o["s0_siblings1"] = sibs["1"]
    if sibs["2"]: o["s0_siblings2"] = sibs["2"]
    if sibs["3"]: o["s0_siblings3"] = sibs["3"]
    o["s0_siblings4"] = sibs["4"]; o["s0_siblingsC"] = sibs["C"]
    o["finalPol"] = [[str(x) for x in e] for e in proof["fri"]["last"]]
    o["publics"] = [str(p) for p in proof["publics"]]
    if prover_addr is not None and len(proof["root1"]) == 1:
        o["proverAddr"] = prover_addr
    return json.dumps(o, separators=(",", ":"))


def proof_from_json(s, hash_type="GL"):
    """Inverse of proof_to_json (serializer.rs:277-520), enough for stark_verify."""
    o = json.loads(s) if isinstance(s, str) else s
    if hash_type == "GL":
        dg = lambda v: [int(v), 0, 0, 0] if isinstance(v, str) else [int(x) for x in v]
    else:
        dg = lambda v: [int(v)]
    nsteps = 1 + sum(1 for k in o if k.startswith("s") and k.endswith("_root"))
    nq = len(o["s0_vals1"])
    queries = [{"root": None, "pol_queries": []} for _ in range(nsteps)]
    names = ["1", "2", "3", "4", "C"]
    for qi in range(nq):
        pq = []
        for n in names:
            v = o.get("s0_vals" + n); s_ = o.get("s0_siblings" + n)
            vals = [int(x) for x in v[qi]] if v is not None else []
            sib = [[int(x) for x in lvl] for lvl in s_[qi]] if s_ is not None else []
            pq.append((vals, sib))
        queries[0]["pol_queries"].append(pq)
    for i in range(1, nsteps):
        queries[i]["root"] = dg(o["s%d_root" % i])
        for qi in range(nq):
            queries[i]["pol_queries"].append([([int(x) for x in o["s%d_vals" % i][qi]], [[int(x) for x in lvl] for lvl in o["s%d_siblings" % i][qi]])])
    return {"rootC": dg(o["rootC"]), "root1": dg(o["root1"]), "root2": dg(o["root2"]), "root3": dg(o["root3"]), "root4": dg(o["root4"]),
            "evals": [tuple(int(x) for x in e) for e in o["evals"]], "publics": [int(p) for p in o["publics"]],
            "fri": {"queries": queries, "last": [tuple(int(x) for x in e) for e in o["finalPol"]]}}


# ---------------------------------------------------------------------------------------------
def _exec_verifier_code(code, ctxv):
    """stark_verify.rs:123-213 (execute_code); values are (f3, dim)."""
    tmp = {}

    def ext(arr, pos, dim):
        return ((arr[pos], 0, 0), 1) if dim == 1 else ((arr[pos], arr[pos + 1], arr[pos + 2]), 3)

    def get(r):
        t = r["type_"]
        if t == "tmp": return tmp[r["id"]]
        if t in ("tree1", "tree2", "tree3", "tree4"): return ext(ctxv[t], r["tree_pos"], r["dim"])
        if t == "const": return ((ctxv["consts"][r["id"]], 0, 0), 1)
        if t == "eval": return (ctxv["evals"][r["id"]], 3)
        if t == "number": return ((parse_pil_number(r["value"]), 0, 0), 1)
        if t == "public": return ((ctxv["publics"][r["id"]], 0, 0), 1)
        if t == "challenge": return (ctxv["challenge"][r["id"]], 3)
        if t == "xDivXSubXi": return (ctxv["xDivXSubXi"], 3)
        if t == "xDivXSubWXi": return (ctxv["xDivXSubWXi"], 3)
        if t == "x": return (ctxv["challenge"][7], 3)
        if t == "Z": return (ctxv["Zp"] if r["prime"] else ctxv["Z"], 3)
        raise ValueError("Invalid reference type, get: %s" % t)

    def vmul(a, b):
        (x, dx), (y, dy) = a, b
        if dx == 1 and dy == 1: return ((x[0] * y[0] % P, 0, 0), 1)
        if dx == 3 and dy == 1: return (f3_muls(x, y[0]), 3)
        if dx == 1 and dy == 3: return (f3_muls(y, x[0]), 3)
        return (f3_mul(x, y), 3)
    vadd = lambda a, b: (f3_add(a[0], b[0]), max(a[1], b[1]))
    vsub = lambda a, b: (f3_sub(a[0], b[0]), max(a[1], b[1]))
    for ci in code:
        src = [get(s) for s in ci["src"]]
        op = ci["op"]
        if op == "add": res = vadd(src[0], src[1])
        elif op == "sub": res = vsub(src[0], src[1])
        elif op == "mul": res = vmul(src[0], src[1])
        elif op == "muladd": res = vadd(vmul(src[0], src[1]), src[2])
        elif op == "copy": res = src[0]
        else: raise ValueError("Invalid op")
        assert ci["dest"]["type_"] == "tmp"
        tmp[ci["dest"]["id"]] = res
    return get(code[-1]["dest"])[0]


def stark_verify(proof, const_root, info, stark_struct, program, reason=None, trace=None):
    """stark_verify (stark_verify.rs:21-121) + FRI::verify (fri.rs:187-297). Returns bool.
    trace (optional dict) receives the evaluation point `xi` and, per query, (index, constant-tree leaf, tree-1 leaf): enough for a caller
    that knows its polynomials in closed form (bench.py at 2^24 rows) to check openings without rebuilding the trees."""
    why = reason if reason is not None else []
    nb, nbe = stark_struct["nBits"], stark_struct["nBitsExt"]
    ext_bits = nbe - nb
    N = 1 << nb
    hash_type = stark_struct["verificationHashType"]
    tr = make_transcript(hash_type)
    for p in proof["publics"]:
        tr.put([p])
    ch = [(0, 0, 0)] * 8
    tr.put(proof["root1"]); ch[0] = tr.get_field(); ch[1] = tr.get_field()
    tr.put(proof["root2"]); ch[2] = tr.get_field(); ch[3] = tr.get_field()
    tr.put(proof["root3"]); ch[4] = tr.get_field()
    tr.put(proof["root4"]); ch[7] = tr.get_field()
    for e in proof["evals"]:
        tr.put(list(e))
    ch[5] = tr.get_field(); ch[6] = tr.get_field()
    if trace is not None:
        trace["xi"] = ch[7]
    x_n = f3_pow(ch[7], N)
    Z = f3_sub(x_n, (1, 0, 0))
    Zp = f3_sub(f3_pow(f3_muls(ch[7], gl.root(nb)), N), (1, 0, 0))
    ctxv = {"evals": proof["evals"], "publics": proof["publics"], "challenge": ch, "Z": Z, "Zp": Zp}
    res = _exec_verifier_code(program["verifier_code"]["first"], ctxv)
    x_acc = (1, 0, 0); q = (0, 0, 0)
    for i in range(info.q_deg):
        q = f3_add(q, f3_mul(x_acc, proof["evals"][info._ev_get("cm", 0, info.qs[i])]))
        x_acc = f3_mul(x_acc, x_n)
    if res != f3_mul(q, Z):
        why.append("Q != C*Z"); return False

    steps = [s["nBits"] for s in stark_struct["steps"]]
    fp = proof["fri"]

    def check_query(query, idx):
        roots = [proof["root1"], proof["root2"], proof["root3"], proof["root4"], const_root]
        for j in range(5):
            if not verify_group_proof_any(hash_type, roots[j], query[j][1], idx, query[j][0]):
                why.append("merkle s0 tree %d idx %d" % (j, idx)); return None
        if trace is not None:
            trace.setdefault("queries", []).append((idx, list(query[4][0]), list(query[0][0])))
        cq = {"tree1": query[0][0], "tree2": query[1][0], "tree3": query[2][0], "tree4": query[3][0], "consts": query[4][0],
              "evals": proof["evals"], "publics": proof["publics"], "challenge": ch}
        x = (gl.SHIFT * pow(gl.root(nb + ext_bits), idx, P) % P, 0, 0)
        cq["xDivXSubXi"] = f3_div(x, f3_sub(x, ch[7]))
        cq["xDivXSubWXi"] = f3_div(x, f3_sub(x, f3_muls(ch[7], gl.root(nb))))
        return [_exec_verifier_code(program["verifier_query_code"]["first"], cq)]

    # FRI::verify
    assert len(fp["queries"]) == len(steps)
    special_x = []
    for si in range(len(steps)):
        special_x.append(tr.get_field())
        if si < len(steps) - 1:
            tr.put(fp["queries"][si + 1]["root"])
        else:
            for e in fp["last"]:
                tr.put(list(e))
    nq = stark_struct["nQueries"]
    ys = tr.get_permutations(nq, steps[0])
    pol_bits = nbe
    shift = gl.SHIFT

    def small_ifft(vals):       # fft.rs:72-83 on F3G
        n = len(vals)
        bits = n.bit_length() - 1
        a = np.array(vals, dtype=np.uint64).reshape(-1)
        return [tuple(int(x) for x in r) for r in gl.intt(a, 3, bits).reshape(-1, 3)] if n > 1 else list(vals)

    for si, nbs in enumerate(steps):
        item = fp["queries"][si]
        red = pol_bits - nbs
        for i in range(nq):
            if si == 0:
                pg = check_query(item["pol_queries"][i], ys[i])
                if pg is None:
                    return False
            else:
                qv, qs_ = item["pol_queries"][i][0]
                if not verify_group_proof_any(hash_type, item["root"], qs_, ys[i], qv):
                    why.append("merkle fri step %d" % si); return False
                pg = [tuple(qv[k:k + 3]) for k in range(0, len(qv), 3)]
            pc = small_ifft(pg)
            sinv = gl.inv(shift * pow(gl.root(pol_bits), ys[i], P) % P)
            xx = f3_muls(special_x[si], sinv)
            ev = pc[-1]
            for k in range(len(pc) - 2, -1, -1):
                ev = f3_add(f3_mul(ev, xx), pc[k])
            if si < len(steps) - 1:
                nng = 1 << steps[si + 1]
                gi = ys[i] // nng
                nxt = fp["queries"][si + 1]["pol_queries"][i][0][0]
                if ev != tuple(nxt[3 * gi:3 * gi + 3]):
                    why.append("fri fold mismatch step %d" % (si + 1)); return False
            elif ev != fp["last"][ys[i]]:
                why.append("fri last mismatch"); return False
        pol_bits = nbs
        for _ in range(red):
            shift = shift * shift % P
        if si < len(steps) - 1:
            ys = [y % (1 << steps[si + 1]) for y in ys]
    max_deg = 0 if pol_bits < (nbe - nb) else 1 << (pol_bits - (nbe - nb))
    lc = small_ifft(fp["last"])
    for i in range(max_deg + 1, len(lc)):
        if lc[i] != (0, 0, 0):
            why.append("final pol degree"); return False
    return True


# ---------------------------------------------------------------------------------------------
def fibonacci_inputs(nbits):
    """The generator that reproduces starky/data/fib.{cm,const}.gl: cm row i = (F_i, F_{i+1}),
    F_0=1, F_1=2 mod p; const ISLAST[i] = (i == N-1)."""
    N = 1 << nbits
    cm = np.zeros((N, 2), dtype=np.uint64)
    a, b = 1, 2
    # vectorised in blocks would need 128-bit adds; a python loop is fine up to 2^20, C handles larger via fib_fill
    for i in range(N):
        cm[i, 0] = a; cm[i, 1] = b
        a, b = b, (a + b) % P
    const = np.zeros((N, 1), dtype=np.uint64); const[N - 1, 0] = 1
    return cm, const


def fibonacci_pil(golden_pil_path, nbits):
    from eigen_zkvm_b200 import starkinfo as si
    pil = si.load_pil(golden_pil_path)
    N = 1 << nbits
    for r in pil["references"].values():
        r["polDeg"] = N
    pil["publics"][0]["idx"] = N - 1
    return pil


def zkvm_steps(nbits_ext):
    """zkvm/src/lib.rs:128-139: steps (2..=nBitsExt).rev().step_by(4)."""
    return [{"nBits": b} for b in range(nbits_ext, 1, -4)]
