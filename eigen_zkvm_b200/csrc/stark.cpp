// Host orchestration of the eSTARK prover on one B200: the device-side equivalent of
// `StarkSetup::new` (starky/src/stark_setup.rs:27-66), `StarkProof::stark_gen` (stark_gen.rs:193-557),
// `FRI::prove` (fri.rs:84-184), `TranscriptGL` (transcript.rs:8-103) and the proof serializer
// (serializer.rs:137-270).  Everything O(N) runs in the CUDA kernels of ntt.cu / merkle.cu / evaluator.cu;
// this file sequences them, keeps the Fiat-Shamir transcript (its Poseidon permutations run on the device
// too) and assembles the proof JSON.  No CPU fallback exists: without a CUDA device every entry point fails.
#include "b200_internal.h"
#include "mini_json.h"
#include "stark.h"
#include "transcript.h"
#include <cstring>
#include <sstream>
#include <algorithm>
#include <set>
#include <map>
#include <mutex>
#include <tuple>

namespace b200 {

static const char* SEC_NAMES[15] = {"cm1_n", "cm2_n", "cm3_n", "cm4_n", "tmpexp_n", "const_n", "cm1_2ns", "cm2_2ns", "cm3_2ns", "cm4_2ns",
                                    "const_2ns", "q_2ns", "f_2ns", "xDivXSubXi", "xDivXSubWXi"};
enum { S_CM1N = 0, S_CM2N, S_CM3N, S_CM4N, S_TMPEXP, S_CONSTN, S_CM1E, S_CM2E, S_CM3E, S_CM4E, S_CONSTE, S_Q, S_F, S_XDX, S_XDWX, S_COUNT };
static int sec_index(const std::string& s) {
    for (int i = 0; i < S_COUNT; i++) if (s == SEC_NAMES[i]) return i;
    throw std::runtime_error("unknown section " + s);
}

// ------------------------------------------------------------------------------------------------ parsed setup
struct Node { std::string type; size_t id = 0; std::string value; bool prime = false; u32 dim = 0; };
struct Section { std::string op; Node dest; std::vector<Node> src; };
struct Segment { std::vector<Section> first; size_t tmp_used = 0; };
struct PolType { int sec; size_t pos; u32 dim; };
struct EvMap { std::string type; size_t id; bool prime; };
struct Public { std::string polType; size_t polId, idx; };

static Node parse_node(const mj::Value& v) {
    Node n; n.type = v.at("type_").as_str(); n.id = v.at("id").as_size(); n.prime = v.at("prime").as_bool(); n.dim = (u32)v.at("dim").as_int();
    if (v.has("value") && !v.at("value").is_null()) n.value = v.at("value").as_str();
    return n;
}
static Segment parse_segment(const mj::Value& v) {
    Segment s; s.tmp_used = v.at("tmp_used").as_size();
    const mj::Value& f = v.at("first");
    for (size_t i = 0; i < f.size(); i++) {
        Section c; c.op = f[i].at("op").as_str(); c.dest = parse_node(f[i].at("dest"));
        const mj::Value& src = f[i].at("src");
        for (size_t j = 0; j < src.size(); j++) c.src.push_back(parse_node(src[j]));
        s.first.push_back(c);
    }
    return s;
}
static std::vector<size_t> parse_usize_vec(const mj::Value& v) { std::vector<size_t> r; for (size_t i = 0; i < v.size(); i++) r.push_back(v[i].as_size()); return r; }

static u64 parse_pil_number(const std::string& s) {     // types.rs:221-233
    bool neg = false; size_t i = 0; unsigned __int128 v = 0;
    if (s.size() > 2 && s[0] == '0' && s[1] == 'x') { for (i = 2; i < s.size(); i++) { char c = s[i]; int d = c <= '9' ? c - '0' : (c | 32) - 'a' + 10; v = (v * 16 + d) % GL_P_HOST; } }
    else { if (s[0] == '-') { neg = true; i = 1; } for (; i < s.size(); i++) v = (v * 10 + (s[i] - '0')) % GL_P_HOST; }
    u64 r = (u64)v;
    return neg && r ? GL_P_HOST - r : r;
}

struct Setup {
    int device = 0;
    unsigned nbits = 0, nbits_ext = 0, n_queries = 0;
    std::vector<unsigned> steps;
    size_t n_cm1 = 0, n_constants = 0, q_deg = 0, q_dim = 0;
    size_t secN[S_COUNT] = {0};                 // map_sectionsN (+ const, x tables)
    std::vector<PolType> var_pol_map;
    std::vector<size_t> cm_n, cm_2ns, tmpexp_n;
    std::vector<EvMap> ev_map;
    std::vector<Public> publics;
    struct ArgCtx { size_t f_exp_id, t_exp_id, num_id, den_id; };
    std::vector<ArgCtx> pu_ctx, pe_ctx, ci_ctx;           // plookup / permutation / connection arguments (starkinfo.rs:14-26)
    std::map<size_t, size_t> exp2pol;
    Segment step2prev, step3prev, step3, step42ns, step52ns;
    std::vector<Segment> publics_code;          // one per public; non-empty `first` for imP publics (starkinfo.rs:274-322)
    int hash = 0;                               // verificationHashType: 0 = GL, 1 = BN128, 2 = BLS12381 (prove.rs:52-89)
    // device-resident, built once per circuit (stark_setup.rs:38-57)
    u64* d_const_n = nullptr;                   // [n_constants][N]
    u64* d_const_2ns = nullptr;                 // [n_constants][Next]
    u64* d_const_nodes = nullptr;
    DevTree const_tree;
    Arena arena;                                // per-proof workspace, reused across proofs
    std::string last_timing_json;
    std::string json;                           // the setup JSON this was built from (kept for setup_export)
    bool self_verify = false;                   // prove.rs:124-132: verify every proof before returning it (setup_set_self_verify / B200_SELF_VERIFY=1)
    std::mutex mu;                              // one proof at a time per setup (the arena is shared); other setups may run concurrently
};

// ------------------------------------------------------------------------------------------------ host F3 helpers (a handful of scalars per proof)
static void hf3_muls(const u64 a[3], u64 s, u64 r[3]) { for (int i = 0; i < 3; i++) r[i] = h_mul(a[i], s); }

// ------------------------------------------------------------------------------------------------ step-program compiler
// interpreter.rs:183-283 (compile_code / get_ref / set_ref) -> EvProgram, with static dims and slot allocation.
static EvProgram compile_program(const Setup& S, const Segment& seg, bool dom_ext, const std::vector<u64>& publics) {
    EvProgram P;
    struct Tmp { u32 dim = 0; u32 slot = 0; long last_use = -1; };
    std::map<size_t, long> last_use;                         // tmp id -> last op index reading it
    for (size_t k = 0; k < seg.first.size(); k++) for (auto& s : seg.first[k].src) if (s.type == "tmp") last_use[s.id] = (long)k;
    std::map<size_t, Tmp> tmps;
    std::vector<bool> used;                                   // slot occupancy
    auto alloc = [&](u32 dim) -> u32 {
        for (u32 s = 0; s + dim <= used.size(); s++) { bool ok = true; for (u32 j = 0; j < dim; j++) if (used[s + j]) { ok = false; break; } if (ok) { for (u32 j = 0; j < dim; j++) used[s + j] = true; return s; } }
        u32 s = (u32)used.size(); while (s > 0 && !used[s - 1]) s--;   // extend from the last free run at the end
        while (used.size() < s + dim) used.push_back(false);
        for (u32 j = 0; j < dim; j++) used[s + j] = true;
        return s;
    };
    auto release = [&](const Tmp& t) { for (u32 j = 0; j < t.dim; j++) used[t.slot + j] = false; };
    auto mem_from_pol = [&](size_t pol_id, bool prime) { const PolType& p = S.var_pol_map.at(pol_id); EvOperand o{1, (u32)p.sec, (u32)p.pos, prime ? 1u : 0u, p.dim}; return o; };
    auto cst = [&](u64 v) { P.consts.push_back(v); return (u32)(P.consts.size() - 1); };
    auto get_ref = [&](const Node& r) -> EvOperand {
        const std::string& t = r.type;
        if (t == "tmp") { auto it = tmps.find(r.id); if (it == tmps.end()) throw std::runtime_error("step program reads an unset tmp"); return EvOperand{0, it->second.slot, 0, 0, it->second.dim}; }
        if (t == "const") return EvOperand{1, (u32)(dom_ext ? S_CONSTE : S_CONSTN), (u32)r.id, r.prime ? 1u : 0u, 1};
        if (t == "cm") return mem_from_pol(dom_ext ? S.cm_2ns.at(r.id) : S.cm_n.at(r.id), r.prime);
        if (t == "tmpExp") { if (dom_ext) throw std::runtime_error("tmpExp in 2ns"); return mem_from_pol(S.tmpexp_n.at(r.id), r.prime); }
        if (t == "number") return EvOperand{2, cst(parse_pil_number(r.value)), 0, 0, 1};
        if (t == "public") return EvOperand{2, cst(publics.at(r.id)), 0, 0, 1};
        if (t == "challenge") return EvOperand{3, (u32)r.id, 0, 0, 3};
        if (t == "eval") return EvOperand{3, (u32)(8 + r.id), 0, 0, 3};
        if (t == "xDivXSubXi") return EvOperand{1, S_XDX, 0, 0, 3};
        if (t == "xDivXSubWXi") return EvOperand{1, S_XDWX, 0, 0, 3};
        if (t == "x") return EvOperand{4, 0, 0, 0, 1};
        if (t == "Zi") return EvOperand{5, 0, 0, 0, 1};
        throw std::runtime_error("Invalid reference type get, " + t);
    };
    for (size_t k = 0; k < seg.first.size(); k++) {
        const Section& c = seg.first[k];
        EvOp op{};
        if (c.op == "add") op.opc = 0; else if (c.op == "sub") op.opc = 1; else if (c.op == "mul") op.opc = 2; else if (c.op == "copy") op.opc = 3;
        else throw std::runtime_error("Invalid op " + c.op);
        op.s0 = get_ref(c.src.at(0));
        if (op.opc != 3) op.s1 = get_ref(c.src.at(1));
        u32 rd = op.opc == 3 ? op.s0.dim : std::max(op.s0.dim, op.s1.dim);
        // sources whose live range ends here free their slots before the destination is placed
        for (auto& s : c.src) if (s.type == "tmp" && last_use[s.id] == (long)k) { auto it = tmps.find(s.id); if (it != tmps.end()) { release(it->second); tmps.erase(it); } }
        const Node& d = c.dest;
        if (d.type == "tmp") {
            auto it = tmps.find(d.id); if (it != tmps.end()) { release(it->second); tmps.erase(it); }
            Tmp t; t.dim = rd; t.slot = alloc(rd); tmps[d.id] = t;
            op.d = EvOperand{0, t.slot, 0, 0, rd};
            if (!last_use.count(d.id) || last_use[d.id] < (long)k) { release(t); tmps.erase(d.id); }    // dead store: slot is scratch
        } else if (d.type == "q") { if (!dom_ext) throw std::runtime_error("Accessing q in domain n"); op.d = EvOperand{1, S_Q, (u32)d.id, 0, rd}; }
        else if (d.type == "f") { if (!dom_ext) throw std::runtime_error("Accessing f in domain n"); op.d = EvOperand{1, S_F, (u32)d.id, 0, rd}; }
        else if (d.type == "cm") { op.d = mem_from_pol(dom_ext ? S.cm_2ns.at(d.id) : S.cm_n.at(d.id), d.prime); op.d.dim = rd; }
        else if (d.type == "tmpExp") { if (dom_ext) throw std::runtime_error("tmpExp in 2ns"); op.d = mem_from_pol(S.tmpexp_n.at(d.id), d.prime); op.d.dim = rd; }
        else throw std::runtime_error("Invalid reference type set " + d.type);
        P.ops.push_back(op);
    }
    P.n_slots = (u32)used.size();
    return P;
}

// algorithmic bytes of one program launch: every distinct (section, column, prime) read or written once per row
static double program_bytes(const EvProgram& P, size_t n) {
    std::set<std::tuple<u32, u32, u32>> cols;
    auto add = [&](const EvOperand& o) { if (o.kind == 1) for (u32 l = 0; l < o.dim; l++) cols.insert(std::make_tuple(o.a, o.b + l, o.prime)); };
    for (auto& op : P.ops) { add(op.d); add(op.s0); if (op.opc != 3) add(op.s1); }
    return 8.0 * (double)n * (double)cols.size();
}

// ------------------------------------------------------------------------------------------------ setup
static std::string u256_dec(const u64 d[4]) {       // a 256-bit little-endian integer in decimal
    u64 v[4] = {d[0], d[1], d[2], d[3]};
    std::string out;
    for (;;) {
        unsigned __int128 rem = 0; bool nz = false;
        for (int i = 3; i >= 0; i--) { unsigned __int128 cur = (rem << 64) | v[i]; v[i] = (u64)(cur / 10000000000000000000ULL); rem = cur % 10000000000000000000ULL; nz |= v[i] != 0; }
        std::string part = std::to_string((u64)rem);
        if (nz) part = std::string(19 - part.size(), '0') + part;
        out = part + out;
        if (!nz) break;
    }
    return out;
}
static void json_digest(std::ostringstream& o, const u64 d[4], int hash = 0) {      // digest.rs:84-111
    if (hash != 0) { o << '"' << u256_dec(d) << '"'; return; }      // BN128 / BLS12-381: the scalar as one decimal string
    if (d[1] == 0 && d[2] == 0 && d[3] == 0) o << '"' << d[0] << '"';
    else o << "[\"" << d[0] << "\",\"" << d[1] << "\",\"" << d[2] << "\",\"" << d[3] << "\"]";
}

// host-only part of the setup: the serde JSON of StarkInfo + Program + StarkStruct -> Setup (no CUDA call)
static std::unique_ptr<Setup> parse_setup(const std::string& setup_json) {
    mj::P root = mj::Parser::parse(setup_json);
    const mj::Value& si = root->at("starkinfo"); const mj::Value& pr = root->at("program"); const mj::Value& ss = root->at("stark_struct");
    std::unique_ptr<Setup> S(new Setup());
    { const char* sv = getenv("B200_SELF_VERIFY"); S->self_verify = sv && atoi(sv) != 0; }
    S->nbits = (unsigned)ss.at("nBits").as_int(); S->nbits_ext = (unsigned)ss.at("nBitsExt").as_int(); S->n_queries = (unsigned)ss.at("nQueries").as_int();
    {
        const std::string ht = ss.at("verificationHashType").as_str();
        if (ht == "GL") S->hash = 0; else if (ht == "BN128") S->hash = 1; else if (ht == "BLS12381") S->hash = 2;
        else throw std::runtime_error("verificationHashType " + ht + " is not supported");
    }
    for (size_t i = 0; i < ss.at("steps").size(); i++) S->steps.push_back((unsigned)ss.at("steps")[i].at("nBits").as_int());
    if (S->steps.empty() || S->steps[0] != S->nbits_ext) throw std::runtime_error("MustEqualDegreeError: nBitsExt != steps[0].nBits");
    S->n_cm1 = si.at("n_cm1").as_size(); S->n_constants = si.at("n_constants").as_size(); S->q_deg = si.at("q_deg").as_size(); S->q_dim = si.at("q_dim").as_size();
    const mj::Value& sn = si.at("map_sectionsN");
    const char* real[11] = {"cm1_n", "cm2_n", "cm3_n", "cm4_n", "tmpexp_n", "cm1_2ns", "cm2_2ns", "cm3_2ns", "cm4_2ns", "q_2ns", "f_2ns"};
    for (auto nm : real) S->secN[sec_index(nm)] = sn.at(nm).as_size();
    S->secN[S_CONSTN] = S->secN[S_CONSTE] = S->n_constants; S->secN[S_XDX] = S->secN[S_XDWX] = 3;
    const mj::Value& vpm = si.at("var_pol_map");
    for (size_t i = 0; i < vpm.size(); i++) S->var_pol_map.push_back(PolType{sec_index(vpm[i].at("section").as_str()), vpm[i].at("section_pos").as_size(), (u32)vpm[i].at("dim").as_int()});
    S->cm_n = parse_usize_vec(si.at("cm_n")); S->cm_2ns = parse_usize_vec(si.at("cm_2ns")); S->tmpexp_n = parse_usize_vec(si.at("tmpexp_n"));
    for (size_t i = 0; i < si.at("ev_map").size(); i++) { const mj::Value& e = si.at("ev_map")[i]; S->ev_map.push_back(EvMap{e.at("type_").as_str(), e.at("id").as_size(), e.at("prime").as_bool()}); }
    for (size_t i = 0; i < si.at("publics").size(); i++) { const mj::Value& e = si.at("publics")[i]; S->publics.push_back(Public{e.at("polType").as_str(), e.at("polId").as_size(), e.at("idx").as_size()}); }
    auto parse_ctx = [&](const char* key, std::vector<Setup::ArgCtx>& out) {
        const mj::Value& a = si.at(key);
        for (size_t i = 0; i < a.size(); i++) out.push_back(Setup::ArgCtx{a[i].at("f_exp_id").as_size(), a[i].at("t_exp_id").as_size(), a[i].at("num_id").as_size(), a[i].at("den_id").as_size()});
    };
    parse_ctx("pu_ctx", S->pu_ctx); parse_ctx("pe_ctx", S->pe_ctx); parse_ctx("ci_ctx", S->ci_ctx);
    for (auto& kv : si.at("exp2pol").obj) S->exp2pol[(size_t)std::stoull(kv.first)] = kv.second->as_size();
    S->step2prev = parse_segment(pr.at("step2prev")); S->step3prev = parse_segment(pr.at("step3prev")); S->step3 = parse_segment(pr.at("step3"));
    S->step42ns = parse_segment(pr.at("step42ns")); S->step52ns = parse_segment(pr.at("step52ns"));
    if (pr.has("publics_code")) for (size_t i = 0; i < pr.at("publics_code").size(); i++) S->publics_code.push_back(parse_segment(pr.at("publics_code")[i]));
    return S;
}

// debug / test hook (host only): the CUDA source the JIT would compile for one step program of this setup
std::string step_program_source(const std::string& setup_json, const std::string& which) {
    std::unique_ptr<Setup> S = parse_setup(setup_json);
    const Segment* seg = which == "step2prev" ? &S->step2prev : which == "step3prev" ? &S->step3prev : which == "step3" ? &S->step3
                       : which == "step42ns" ? &S->step42ns : which == "step52ns" ? &S->step52ns : nullptr;
    if (!seg) throw std::invalid_argument("unknown step program " + which);
    std::vector<u64> publics(S->publics.size(), 0);
    EvProgram P = compile_program(*S, *seg, which == "step42ns" || which == "step52ns", publics);
    return eval_jit_source(P);
}

// u64 words of the `nodes` array of a tree of `height` leaves for this setup's hash: the binary GL layout (2h - 1 digests) is LARGER than
// the 16-ary one for h >= 16 but SMALLER below (a 4-leaf 16-ary tree is 16 + 1 digests): always take the maximum
static size_t tree_nodes_u64(size_t height) { return std::max(merkle_n_nodes(height), big_merkle_n_nodes(height)) * 4; }
Setup* setup_new(const std::string& setup_json, const u64* const_rowmajor, bool const_on_device, size_t n_rows, size_t n_consts) {
    std::unique_ptr<Setup> S = parse_setup(setup_json);
    S->json = setup_json;
    B200_CUDA_CHECK(cudaGetDevice(&S->device));
    if (n_consts != S->n_constants) throw std::runtime_error("const_pol.nPols != pil.nConstants");
    const size_t N = (size_t)1 << S->nbits, Ne = (size_t)1 << S->nbits_ext;
    if (n_rows != N) throw std::runtime_error("constant polynomial height != 2^nBits");
    if (n_consts && !const_rowmajor) throw std::invalid_argument("null constant polynomials");

    // const polynomials: row-major in (polsarray.rs write_buff) -> column-major, LDE, Merkle (stark_setup.rs:38-57)
    size_t nc = S->n_constants;
    B200_CUDA_CHECK(cudaMalloc(&S->d_const_n, std::max<size_t>(1, nc * N) * 8));
    B200_CUDA_CHECK(cudaMalloc(&S->d_const_2ns, std::max<size_t>(1, nc * Ne) * 8));
    if (nc) {
        u64* d_rm = nullptr;
        if (const_on_device) d_rm = const_cast<u64*>(const_rowmajor);
        else { B200_CUDA_CHECK(cudaMalloc(&d_rm, nc * N * 8)); B200_CUDA_CHECK(cudaMemcpyAsync(d_rm, const_rowmajor, nc * N * 8, cudaMemcpyHostToDevice, stream())); }
        transpose_rm_to_cm(d_rm, S->d_const_n, N, nc);
        B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
        if (!const_on_device) B200_CUDA_CHECK(cudaFree(d_rm));
        lde_cols(S->d_const_n, S->d_const_2ns, nc, S->nbits, S->nbits_ext);
    }
    B200_CUDA_CHECK(cudaMalloc(&S->d_const_nodes, tree_nodes_u64(Ne) * 8));
    S->const_tree.hash = S->hash;
    merkelize(S->const_tree, colview_plain(S->d_const_2ns, Ne), nc, Ne, S->d_const_nodes);
    B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
    return S.release();
}

// ---- serialized StarkSetup (stark_setup.rs:13-19: const_tree, const_root, starkinfo, program are serde-serializable, so that the
// constant LDE and tree are built once per CIRCUIT, not once per process).  File: "B2SU", version, the setup JSON, the shape, then the
// constant polynomials, their extension and the tree nodes exactly as they sit in device memory.
static const size_t IO_CHUNK = (size_t)8 << 20;       // u64 per staging chunk (64 MiB)
static void dev_to_file(FILE* f, const u64* d, size_t n, u64* pin) {
    for (size_t o = 0; o < n; o += IO_CHUNK) {
        const size_t k = std::min(IO_CHUNK, n - o);
        B200_CUDA_CHECK(cudaMemcpy(pin, d + o, k * 8, cudaMemcpyDeviceToHost));
        if (fwrite(pin, 8, k, f) != k) throw std::runtime_error("setup export: short write");
    }
}
static void file_to_dev(FILE* f, u64* d, size_t n, u64* pin) {
    for (size_t o = 0; o < n; o += IO_CHUNK) {
        const size_t k = std::min(IO_CHUNK, n - o);
        if (fread(pin, 8, k, f) != k) throw std::runtime_error("setup import: truncated file");
        B200_CUDA_CHECK(cudaMemcpy(d + o, pin, k * 8, cudaMemcpyHostToDevice));
    }
}
static size_t const_nodes_u64(const Setup& S, size_t Ne) { return (S.hash == 0 ? merkle_n_nodes(Ne) : big_merkle_n_nodes(Ne)) * 4; }
void setup_export(const Setup* S, const char* path) {
    if (S->json.empty()) throw std::runtime_error("setup export: this setup was not built from JSON");
    const size_t N = (size_t)1 << S->nbits, Ne = (size_t)1 << S->nbits_ext, nc = S->n_constants;
    FILE* f = fopen(path, "wb");
    if (!f) throw std::invalid_argument(std::string("cannot create ") + path);
    u64* pin = nullptr;
    try {
        B200_CUDA_CHECK(cudaSetDevice(S->device));
        B200_CUDA_CHECK(cudaMallocHost(&pin, IO_CHUNK * 8));
        const u64 hdr[8] = {0x0000000155533242ull /* "B2SU", version 1 */, S->json.size(), N, Ne, nc, nc ? const_nodes_u64(*S, Ne) : 0, (u64)S->hash, 0};
        if (fwrite(hdr, 8, 8, f) != 8 || fwrite(S->json.data(), 1, S->json.size(), f) != S->json.size()) throw std::runtime_error("setup export: short write");
        if (fwrite(S->const_tree.root, 8, 4, f) != 4) throw std::runtime_error("setup export: short write");
        if (nc) { dev_to_file(f, S->d_const_n, nc * N, pin); dev_to_file(f, S->d_const_2ns, nc * Ne, pin); dev_to_file(f, S->d_const_nodes, hdr[5], pin); }
    } catch (...) { fclose(f); if (pin) cudaFreeHost(pin); throw; }
    cudaFreeHost(pin);
    if (fclose(f)) throw std::runtime_error("setup export: close failed");
}
Setup* setup_import(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) throw std::invalid_argument(std::string("cannot open ") + path);
    u64* pin = nullptr; std::unique_ptr<Setup> S;
    try {
        u64 hdr[8];
        if (fread(hdr, 8, 8, f) != 8 || hdr[0] != 0x0000000155533242ull) throw std::runtime_error("setup import: not a b200 setup file (or another version)");
        if (hdr[1] > ((u64)1 << 32)) throw std::runtime_error("setup import: bad header");
        std::string js(hdr[1], 0);
        if (fread(&js[0], 1, js.size(), f) != js.size()) throw std::runtime_error("setup import: truncated file");
        S = parse_setup(js); S->json = js;
        const size_t N = (size_t)1 << S->nbits, Ne = (size_t)1 << S->nbits_ext, nc = S->n_constants;
        if (hdr[2] != N || hdr[3] != Ne || hdr[4] != nc || hdr[6] != (u64)S->hash || hdr[5] != (nc ? const_nodes_u64(*S, Ne) : 0)) throw std::runtime_error("setup import: shape does not match the embedded setup JSON");
        u64 root[4];
        if (fread(root, 8, 4, f) != 4) throw std::runtime_error("setup import: truncated file");
        B200_CUDA_CHECK(cudaGetDevice(&S->device));
        B200_CUDA_CHECK(cudaMallocHost(&pin, IO_CHUNK * 8));
        B200_CUDA_CHECK(cudaMalloc(&S->d_const_n, std::max<size_t>(1, nc * N) * 8));
        B200_CUDA_CHECK(cudaMalloc(&S->d_const_2ns, std::max<size_t>(1, nc * Ne) * 8));
        B200_CUDA_CHECK(cudaMalloc(&S->d_const_nodes, tree_nodes_u64(Ne) * 8));
        S->const_tree.hash = S->hash;
        if (nc) {
            file_to_dev(f, S->d_const_n, nc * N, pin); file_to_dev(f, S->d_const_2ns, nc * Ne, pin); file_to_dev(f, S->d_const_nodes, hdr[5], pin);
            DevTree& t = S->const_tree;
            t.cols = colview_plain(S->d_const_2ns, Ne); t.width = nc; t.height = Ne; t.nodes = S->d_const_nodes; t.degenerate = false;
            memcpy(t.root, root, 32);
        } else {
            merkelize(S->const_tree, colview_plain(S->d_const_2ns, Ne), 0, Ne, S->d_const_nodes);       // no constants: the degenerate tree is a handful of permutations
            if (memcmp(S->const_tree.root, root, 32)) throw std::runtime_error("setup import: constant root mismatch");
        }
        char extra;
        if (fread(&extra, 1, 1, f) == 1) throw std::runtime_error("setup import: trailing bytes");
    } catch (...) {
        fclose(f); if (pin) cudaFreeHost(pin);
        if (S) { cudaFree(S->d_const_n); cudaFree(S->d_const_2ns); cudaFree(S->d_const_nodes); }
        throw;
    }
    fclose(f); cudaFreeHost(pin);
    return S.release();
}

void setup_free(Setup* S) {
    if (!S) return;
    cudaFree(S->d_const_n); cudaFree(S->d_const_2ns); cudaFree(S->d_const_nodes); S->arena.release();
    delete S;
}
void setup_set_self_verify(Setup* S, bool on) { S->self_verify = on; }
void setup_const_root(const Setup* S, u64 out4[4]) { memcpy(out4, S->const_tree.root, 32); }
void setup_shape(const Setup* S, size_t out[4]) { out[0] = S->nbits; out[1] = S->nbits_ext; out[2] = S->n_cm1; out[3] = S->n_constants; }

static size_t arena_need(const Setup& S) {
    const size_t N = (size_t)1 << S.nbits, Ne = (size_t)1 << S.nbits_ext;
    size_t wn = S.secN[S_CM1N] + S.secN[S_CM2N] + S.secN[S_CM3N] + S.secN[S_TMPEXP];
    size_t we = S.secN[S_CM1E] + S.secN[S_CM2E] + S.secN[S_CM3E] + S.secN[S_CM4E] + S.q_dim + 3 + 6;
    size_t u = N * (wn + S.n_cm1 /* row-major staging */ + 6 /* LEv */ + S.q_dim * S.q_deg) + Ne * (we + S.q_dim /* qq1 */);
    size_t trees = 0; for (int s : {S_CM1E, S_CM2E, S_CM3E, S_CM4E}) if (S.secN[s]) trees++;
    u += trees * merkle_n_nodes(Ne) * 4;
    if (!S.pu_ctx.empty() || !S.pe_ctx.empty() || !S.ci_ctx.empty()) u += calculate_Z_tmp_u64(N);
    u += 3 * Ne / 4 + 2 * merkle_n_nodes(Ne) * 4 / 4;        // FRI layers and their trees (geometric, generous)
    return u * 8 + (size_t)S.steps.size() * 4096 + (64u << 20);
}

// ------------------------------------------------------------------------------------------------ stark_gen
struct ProofParts {
    u64 root[5][4];                 // root1..4, rootC
    std::vector<std::array<u64, 3>> evals;
    std::vector<u64> publics;
    struct Opening { std::vector<u64> vals, sibs; size_t depth = 0, width = 0; };
    std::vector<std::array<Opening, 5>> s0;            // per query: tree1..4, const
    struct FriStep { u64 root[4]; Opening op; };       // op holds all queries
    std::vector<FriStep> fri;                          // steps 1..
    std::vector<u64> last;                             // finalPol lanes (AoS)
    int hash = 0; std::string prover_addr;
};

static void write_opening_vals(std::ostringstream& o, const ProofParts::Opening& op, size_t q) {
    o << '[';
    for (size_t c = 0; c < op.width; c++) { if (c) o << ','; o << '"' << op.vals[q * op.width + c] << '"'; }
    o << ']';
}
static void write_opening_sibs(std::ostringstream& o, const ProofParts::Opening& op, size_t q, int hash) {
    if (hash != 0) {        // [level][16] decimal scalars (merklehash_bn128.rs:89-106, serializer.rs:160-172)
        o << '[';
        for (size_t d = 0; d < op.depth; d++) { if (d) o << ','; o << '['; for (int k = 0; k < 16; k++) { if (k) o << ','; o << '"' << u256_dec(&op.sibs[((q * op.depth + d) * 16 + k) * 4]) << '"'; } o << ']'; }
        o << ']';
        return;
    }
    o << '[';
    for (size_t d = 0; d < op.depth; d++) { if (d) o << ','; o << '['; for (int k = 0; k < 4; k++) { if (k) o << ','; o << '"' << op.sibs[(q * op.depth + d) * 4 + k] << '"'; } o << ']'; }
    o << ']';
}
static std::string proof_json(const ProofParts& P, size_t n_queries) {     // serializer.rs:137-270
    std::ostringstream o;
    o << "{\"rootC\":"; json_digest(o, P.root[4], P.hash);
    for (int i = 0; i < 4; i++) { o << ",\"root" << (i + 1) << "\":"; json_digest(o, P.root[i], P.hash); }
    o << ",\"evals\":[";
    for (size_t i = 0; i < P.evals.size(); i++) { if (i) o << ','; o << "[\"" << P.evals[i][0] << "\",\"" << P.evals[i][1] << "\",\"" << P.evals[i][2] << "\"]"; }
    o << ']';
    for (size_t s = 0; s < P.fri.size(); s++) {
        o << ",\"s" << (s + 1) << "_root\":"; json_digest(o, P.fri[s].root, P.hash);
        o << ",\"s" << (s + 1) << "_vals\":[";
        for (size_t q = 0; q < n_queries; q++) { if (q) o << ','; write_opening_vals(o, P.fri[s].op, q); }
        o << "],\"s" << (s + 1) << "_siblings\":[";
        for (size_t q = 0; q < n_queries; q++) { if (q) o << ','; write_opening_sibs(o, P.fri[s].op, q, P.hash); }
        o << ']';
    }
    const char* nm[5] = {"1", "2", "3", "4", "C"};
    for (int pass = 0; pass < 2; pass++)
        for (int t = 0; t < 5; t++) {
            o << (pass == 0 ? ",\"s0_vals" : ",\"s0_siblings") << nm[t] << "\":[";
            for (size_t q = 0; q < n_queries; q++) { if (q) o << ','; if (pass == 0) write_opening_vals(o, P.s0[q][t], 0); else write_opening_sibs(o, P.s0[q][t], 0, P.hash); }
            o << ']';
        }
    o << ",\"finalPol\":[";
    for (size_t i = 0; i < P.last.size() / 3; i++) { if (i) o << ','; o << "[\"" << P.last[3 * i] << "\",\"" << P.last[3 * i + 1] << "\",\"" << P.last[3 * i + 2] << "\"]"; }
    o << "],\"publics\":[";
    for (size_t i = 0; i < P.publics.size(); i++) { if (i) o << ','; o << '"' << P.publics[i] << '"'; }
    o << "]";
    if (P.hash != 0) { o << ",\"proverAddr\":\""; for (char c : P.prover_addr) { if (c == '"' || c == '\\') o << '\\'; o << c; } o << '"'; }    // serializer.rs:262-266
    o << "}";
    return o.str();
}

// StarkProof::calculate_exp_at_point (stark_gen.rs:559-572): an `imP` public is the value of a small base-field program
// (program.publics_code[i], starkinfo.rs:274-322) at one row of the "n" domain; a handful of scalar reads from the device.
static u64 public_at_point(const Setup& S, const Segment& seg, size_t idx, const EvSection* sec, const std::vector<u64>& publics) {
    const size_t N = (size_t)1 << S.nbits;
    std::map<size_t, u64> tmp;
    auto fetch = [&](const u64* d) { u64 v; B200_CUDA_CHECK(cudaMemcpyAsync(&v, d, 8, cudaMemcpyDeviceToHost, stream())); B200_CUDA_CHECK(cudaStreamSynchronize(stream())); return v; };
    auto val = [&](const Node& r) -> u64 {
        if (r.type == "tmp") return tmp.at(r.id);
        if (r.type == "number") return parse_pil_number(r.value);
        if (r.type == "public") return publics.at(r.id);
        const size_t row = (idx + (r.prime ? 1 : 0)) % N;
        if (r.type == "const") return fetch(S.d_const_n + r.id * N + row);
        if (r.type == "cm") {
            const PolType& p = S.var_pol_map.at(S.cm_n.at(r.id));
            if (p.dim != 1) throw std::runtime_error("extension-field public calculators are not supported");
            return fetch(sec[p.sec].base + p.pos * N + row);
        }
        throw std::runtime_error("public calculator operand " + r.type + " is not supported");
    };
    u64 res = 0;
    for (auto& op : seg.first) {
        u64 a = val(op.src.at(0)), b = op.src.size() > 1 ? val(op.src[1]) : 0;
        if (op.op == "copy") res = a; else if (op.op == "add") res = h_add(a, b); else if (op.op == "sub") res = h_sub(a, b); else if (op.op == "mul") res = h_mul(a, b);
        else throw std::runtime_error("public calculator op " + op.op + " is not supported");
        if (op.dest.type != "tmp") throw std::runtime_error("public calculator destination is not supported");
        tmp[op.dest.id] = res;
    }
    return res;
}

std::string stark_gen(Setup* Sp, const u64* cm_rowmajor, bool cm_on_device, size_t n_rows, size_t n_cols, const char* prover_addr) {
    Setup& S = *Sp;
    std::lock_guard<std::mutex> setup_lock(S.mu);
    const size_t N = (size_t)1 << S.nbits, Ne = (size_t)1 << S.nbits_ext;
    const unsigned ext_bits = S.nbits_ext - S.nbits;
    if (n_rows != N || n_cols != S.n_cm1) throw std::runtime_error("cm_pols shape does not match the setup (rows " + std::to_string(n_rows) + " cols " + std::to_string(n_cols) + ")");
    B200_CUDA_CHECK(cudaSetDevice(S.device));
    S.arena.reserve(arena_need(S));
    Arena& A = S.arena; A.reset();
    cudaStream_t st = stream();

    EvSection sec[S_COUNT];
    for (int i = 0; i < S_COUNT; i++) sec[i] = EvSection{nullptr, 0};
    auto mk = [&](int s, size_t rows, bool zero) { size_t w = S.secN[s]; u64* p = A.alloc_u64(std::max<size_t>(1, w * rows)); if (zero && w) B200_CUDA_CHECK(cudaMemsetAsync(p, 0, w * rows * 8, st)); sec[s] = EvSection{p, rows}; return p; };
    // trace in: row-major (reference layout) -> column-major
    u64* cm1_n = mk(S_CM1N, N, false);
    {
        const u64* d_rm = cm_rowmajor;
        if (!cm_on_device) { u64* stg = A.alloc_u64(N * S.n_cm1); B200_CUDA_CHECK(cudaMemcpyAsync(stg, cm_rowmajor, N * S.n_cm1 * 8, cudaMemcpyHostToDevice, st)); d_rm = stg; }
        transpose_rm_to_cm(d_rm, cm1_n, N, S.n_cm1);
    }
    mk(S_CM2N, N, true); mk(S_CM3N, N, true); mk(S_TMPEXP, N, true);
    sec[S_CONSTN] = EvSection{S.d_const_n, N}; sec[S_CONSTE] = EvSection{S.d_const_2ns, Ne};
    u64* cm_e[4] = {mk(S_CM1E, Ne, false), mk(S_CM2E, Ne, false), mk(S_CM3E, Ne, false), mk(S_CM4E, Ne, false)};
    u64* q_2ns = mk(S_Q, Ne, true);
    S.secN[S_F] = 3; u64* f_2ns = mk(S_F, Ne, true);
    DevPowTab x_n_tab = powtab(h_root(S.nbits), S.nbits), x_e_tab = powtab(h_root(S.nbits_ext), S.nbits_ext);
    u64* d_zi = A.alloc_u64((size_t)1 << ext_bits);
    zh_inv_table(d_zi, S.nbits, ext_bits);

    ProofParts PP;
    PP.hash = S.hash; PP.prover_addr = prover_addr ? prover_addr : "";
    // publics (stark_gen.rs:256-270)
    for (size_t i = 0; i < S.publics.size(); i++) {
        const Public& pe = S.publics[i];
        u64 v;
        if (pe.polType == "cmP") { B200_CUDA_CHECK(cudaMemcpyAsync(&v, cm1_n + pe.polId * N + pe.idx, 8, cudaMemcpyDeviceToHost, st)); B200_CUDA_CHECK(cudaStreamSynchronize(st)); }
        else if (pe.polType == "imP") {
            if (i >= S.publics_code.size() || S.publics_code[i].first.empty()) throw std::runtime_error("imP public without a calculator in program.publics_code");
            v = public_at_point(S, S.publics_code[i], pe.idx, sec, PP.publics);
        } else throw std::runtime_error("Invalid public type " + pe.polType);
        PP.publics.push_back(v);
    }
    Transcript tr(S.hash);
    for (u64 p : PP.publics) tr.put1(p);

    std::vector<u64> f3c((8 + S.ev_map.size()) * 3, 0);       // challenges then evals
    auto challenge = [&](int i) { tr.get_field(&f3c[3 * i]); };
    auto run = [&](const Segment& seg, bool dom_ext) {
        if (seg.first.empty()) return;
        EvProgram P = compile_program(S, seg, dom_ext, PP.publics);
        size_t n = dom_ext ? Ne : N;
        eval_program(P, sec, S_COUNT, f3c.data(), (int)(f3c.size() / 3), dom_ext ? x_e_tab : x_n_tab, dom_ext ? 49 : 1, d_zi, (u32)((1u << ext_bits) - 1), n, dom_ext ? ((size_t)1 << ext_bits) : 1, program_bytes(P, n));
    };
    DevTree trees[4];
    for (auto& t : trees) t.hash = S.hash;
    auto extend_and_merkelize = [&](int k) {       // stark_gen.rs:710-732
        int sn_ = S_CM1N + k, se = S_CM1E + k; size_t w = S.secN[sn_];
        lde_cols(sec[sn_].base, cm_e[k], w, S.nbits, S.nbits_ext);
        u64* nodes = w ? A.alloc_u64(tree_nodes_u64(Ne)) : nullptr;
        merkelize(trees[k], colview_plain(cm_e[k], Ne), w, Ne, nodes);
        memcpy(PP.root[k], trees[k].root, 32);
        tr.put_digest(trees[k].root);
        (void)se;
    };
    struct DevPol { u64* p; u32 dim; };
    auto pol_n = [&](size_t pol_id) { const PolType& p = S.var_pol_map.at(pol_id); if (sec[p.sec].rows != N) throw std::runtime_error("polynomial is not in an n-domain section"); return DevPol{sec[p.sec].base + p.pos * N, p.dim}; };
    auto exp_pol = [&](size_t exp_id) { auto it = S.exp2pol.find(exp_id); if (it == S.exp2pol.end()) throw std::runtime_error("exp2pol has no entry for expression " + std::to_string(exp_id)); return pol_n(it->second); };
    size_t n_cm = S.n_cm1;
    extend_and_merkelize(0);
    challenge(0); challenge(1);
    run(S.step2prev, false);
    for (auto& pu : S.pu_ctx) {        // stark_gen.rs:300-308
        DevPol f = exp_pol(pu.f_exp_id), t = exp_pol(pu.t_exp_id), h1 = pol_n(S.cm_n.at(n_cm)), h2 = pol_n(S.cm_n.at(n_cm + 1));
        if (f.dim != t.dim || h1.dim != f.dim || h2.dim != f.dim) throw std::runtime_error("plookup polynomials of mixed dimension are not supported");
        calculate_H1H2(f.p, t.p, f.dim, N, h1.p, h2.p);
        n_cm += 2;
    }
    extend_and_merkelize(1);
    challenge(2); challenge(3);
    run(S.step3prev, false);
    {
        u64* ztmp = nullptr;
        auto do_z = [&](const Setup::ArgCtx& o) {  // stark_gen.rs:323-353
            if (!ztmp) ztmp = A.alloc_u64(calculate_Z_tmp_u64(N));
            DevPol num = exp_pol(o.num_id), den = exp_pol(o.den_id), z = pol_n(S.cm_n.at(n_cm));
            calculate_Z(num.p, num.dim, den.p, den.dim, z.p, z.dim, N, ztmp);
            n_cm++;
        };
        for (auto& o : S.pu_ctx) do_z(o);
        for (auto& o : S.pe_ctx) do_z(o);
        for (auto& o : S.ci_ctx) do_z(o);
    }
    run(S.step3, false);
    extend_and_merkelize(2);
    challenge(4);
    run(S.step42ns, true);

    // quotient: iNTT, split by degree chunks, forward NTT (stark_gen.rs:375-405)
    {
        u64* qq1 = A.alloc_u64(std::max<size_t>(1, S.q_dim * Ne));
        ntt_cols(q_2ns, qq1, S.q_dim, S.nbits_ext, true);
        size_t w4 = S.q_dim * S.q_deg;
        if (S.q_deg > 0) {
            u64* qq2 = A.alloc_u64(w4 * N);
            quotient_split(qq1, qq2, N, Ne, S.q_dim, S.q_deg, S.nbits);
            ntt_cols_padded(qq2, N, cm_e[3], w4, S.nbits_ext);
        }
        u64* nodes = w4 ? A.alloc_u64(tree_nodes_u64(Ne)) : nullptr;
        merkelize(trees[3], colview_plain(cm_e[3], Ne), S.secN[S_CM4E], Ne, nodes);
        memcpy(PP.root[3], trees[3].root, 32);
        tr.put_digest(trees[3].root);
    }
    challenge(7);   // xi

    // evaluations at xi and w*xi (stark_gen.rs:416-466)
    const u64* xi = &f3c[3 * 7];
    u64 shift_inv = h_inv(49), w_n = h_root(S.nbits);
    {
        u64 xis[3], wxis[3], t[3];
        hf3_muls(xi, shift_inv, xis);
        hf3_muls(xi, w_n, t); hf3_muls(t, shift_inv, wxis);
        u64* LEv = A.alloc_u64(3 * N); u64* LpEv = A.alloc_u64(3 * N);
        // LEv = iNTT(powers of xi / shift), LpEv likewise for w xi / shift: closed form of the geometric sums (evaluator.cu lagrange_row)
        lagrange_row(x_n_tab, S.nbits, xis, LEv); lagrange_row(x_n_tab, S.nbits, wxis, LpEv);
        // the evaluations at xi share LEv, those at w xi share LpEv: groups of up to 8 per launch read their Lagrange row once
        const size_t n_ev = S.ev_map.size();
        std::vector<std::array<u64, 3>> evs(n_ev);
        for (int prime = 0; prime < 2; prime++) {
            std::vector<size_t> idx;
            for (size_t i = 0; i < n_ev; i++) if ((S.ev_map[i].prime ? 1 : 0) == prime) idx.push_back(i);
            for (size_t g0 = 0; g0 < idx.size(); g0 += 8) {
                const int m = (int)std::min<size_t>(8, idx.size() - g0);
                const u64* cols[8]; size_t strides[8]; int dims[8]; u64 out[24];
                for (int e = 0; e < m; e++) {
                    const EvMap& ev = S.ev_map[idx[g0 + e]];
                    strides[e] = Ne;
                    if (ev.type == "const") { cols[e] = S.d_const_2ns + ev.id * Ne; dims[e] = 1; }
                    else if (ev.type == "cm") { const PolType& p = S.var_pol_map.at(S.cm_2ns.at(ev.id)); cols[e] = sec[p.sec].base + p.pos * Ne; dims[e] = (int)p.dim; }
                    else throw std::runtime_error("Invalid ev type: " + ev.type);
                }
                eval_dot_multi(cols, strides, dims, m, ext_bits, prime ? LpEv : LEv, N, out);
                for (int e = 0; e < m; e++) evs[idx[g0 + e]] = {out[3 * e], out[3 * e + 1], out[3 * e + 2]};
            }
        }
        for (size_t i = 0; i < n_ev; i++) { PP.evals.push_back(evs[i]); memcpy(&f3c[3 * (8 + i)], evs[i].data(), 24); }
    }
    for (auto& e : PP.evals) tr.put(e.data(), 3);
    challenge(5); challenge(6);

    // x/(x - xi), x/(x - w xi) (stark_gen.rs:481-522)
    {
        u64 wxi[3]; hf3_muls(xi, w_n, wxi);
        u64* a = A.alloc_u64(3 * Ne); u64* b = A.alloc_u64(3 * Ne);
        xdivxsub(x_e_tab, 49, Ne, xi, a); xdivxsub(x_e_tab, 49, Ne, wxi, b);
        sec[S_XDX] = EvSection{a, Ne}; sec[S_XDWX] = EvSection{b, Ne};
    }
    run(S.step52ns, true);

    // FRI (fri.rs:84-184)
    {
        const size_t nsteps = S.steps.size();
        unsigned pol_bits = S.nbits_ext;
        u64 sinv = shift_inv;
        const u64* pol = f_2ns; size_t pol_n = Ne;
        std::vector<DevTree> ftrees(nsteps > 0 ? nsteps - 1 : 0);
        for (auto& t : ftrees) t.hash = S.hash;
        PP.fri.resize(nsteps > 0 ? nsteps - 1 : 0);
        for (size_t si_ = 0; si_ < nsteps; si_++) {
            unsigned red = pol_bits - S.steps[si_];
            size_t pol2_n = (size_t)1 << (pol_bits - red);
            u64 sx[3]; tr.get_field(sx);
            const u64* pol2 = pol;
            if (si_ > 0) { u64* o = A.alloc_u64(3 * pol2_n); fri_fold(pol, o, pol_bits, red, sinv, sx); pol2 = o; }
            else if (red != 0) throw std::runtime_error("steps[0].nBits must equal nBitsExt");
            if (si_ + 1 < nsteps) {
                size_t n_groups = (size_t)1 << S.steps[si_ + 1], group_size = ((size_t)1 << S.steps[si_]) / n_groups;
                ColView cv{pol2, 3u, (u64)n_groups, (u64)pol2_n};   // column 3j+l of row r = lane l of pol2[j*n_groups + r] (fri.rs:299-317)
                u64* nodes = A.alloc_u64(tree_nodes_u64(n_groups));
                merkelize(ftrees[si_], cv, 3 * group_size, n_groups, nodes);
                memcpy(PP.fri[si_].root, ftrees[si_].root, 32);
                tr.put_digest(ftrees[si_].root);
            } else {
                std::vector<u64> h(3 * pol2_n);
                B200_CUDA_CHECK(cudaMemcpyAsync(h.data(), pol2, 3 * pol2_n * 8, cudaMemcpyDeviceToHost, st)); B200_CUDA_CHECK(cudaStreamSynchronize(st));
                PP.last.resize(3 * pol2_n);
                for (size_t i = 0; i < pol2_n; i++) for (int l = 0; l < 3; l++) PP.last[3 * i + l] = h[(size_t)l * pol2_n + i];
                tr.put(PP.last.data(), PP.last.size());
            }
            pol = pol2; pol_n = pol2_n; pol_bits -= red;
            for (unsigned j = 0; j < red; j++) sinv = h_mul(sinv, sinv);
        }
        (void)pol_n;
        std::vector<u64> ys = tr.get_permutations(S.n_queries, S.steps[0]);
        // step 0: openings of tree1..4 and the constant tree
        PP.s0.resize(ys.size());
        const DevTree* t0[5] = {&trees[0], &trees[1], &trees[2], &trees[3], &S.const_tree};
        for (int t = 0; t < 5; t++) {
            std::vector<u64> vals, sibs; size_t depth;
            merkle_open(*t0[t], ys, vals, sibs, depth);
            size_t w = t0[t]->width;
            for (size_t q = 0; q < ys.size(); q++) {
                ProofParts::Opening& op = PP.s0[q][t];
                op.width = w; op.depth = depth;
                op.vals.assign(vals.begin() + q * w, vals.begin() + (q + 1) * w);
                const size_t per = depth * (S.hash ? 64 : 4);
                op.sibs.assign(sibs.begin() + q * per, sibs.begin() + (q + 1) * per);
            }
        }
        for (size_t si_ = 1; si_ < nsteps; si_++) {
            for (auto& y : ys) y %= (u64)1 << S.steps[si_];
            ProofParts::Opening& op = PP.fri[si_ - 1].op;
            merkle_open(ftrees[si_ - 1], ys, op.vals, op.sibs, op.depth);
            op.width = ftrees[si_ - 1].width;
        }
    }
    memcpy(PP.root[4], S.const_tree.root, 32);
    std::string js = proof_json(PP, S.n_queries);
    if (S.self_verify) {          // the reference asserts stark_verify after every proof (prove.rs:124-132)
        ScopedTimer tv("self_verify_host");
        std::string why;
        if (!stark_verify(S.json, S.const_tree.root, js, why)) throw std::runtime_error("self-verification failed: the verifier rejects the generated proof (" + why + ")");
    }
    return js;
}

}  // namespace b200
