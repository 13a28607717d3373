// Row-parallel kernels of stark_gen other than NTT and hashing: the step-program evaluator
// (starky/src/stark_gen.rs:752-963 + interpreter.rs:91-283), Zi table (:575-592), quotient split (:375-396),
// LEv powers and evaluation dot-products (:416-466), xDivXSubXi tables (:481-522, polutils.rs:35-53) and the
// FRI fold (fri.rs:101-151).  All sections are column-major on the device; one thread owns one row, so every
// global access is coalesced across the warp.  Roofline class: HBM (the programs are short), except the
// batch inversion which is INT bound.
#include "b200_internal.h"
#include "field.cuh"
#include <map>
#include <set>
#include <vector>
#include <mutex>
#include <string>
#include <cstdio>
#include <cstdlib>

namespace b200 {

// ------------------------------------------------------------------------------------------------ evaluator
// Operand kinds (host compiler in stark.cpp mirrors interpreter.rs get_ref/set_ref):
//   0 TMP   a = first slot (dim consecutive u64 slots in shared memory, [slot][thread])
//   1 MEM   a = section, b = first column, prime -> row (i + next) mod n, dim in {1,3} = consecutive columns
//   2 CONST a = index into consts[]        (numbers, publics; dim 1)
//   3 F3C   a = index into f3consts[] * 3  (challenges, evals; dim 3)
//   4 X     x[i] = x_start * w^i           (dim 1)
//   5 ZI    zi[i & zi_mask]                (dim 1)
// opc: 0 add, 1 sub, 2 mul, 3 copy.  Value dims follow F3G's runtime `dim` (f3g.rs:13-18) but are inferred
// statically by the host compiler; a dim-1 value stored to a wider destination writes lane 0 only
// (interpreter.rs:146-166).
#define EV_MAX_SECS 16
struct EvSecs { EvSection s[EV_MAX_SECS]; };

GL_D f3 ev_load(const EvOperand& o, const EvSecs& secs, const u64* __restrict__ consts, const u64* __restrict__ f3c,
                const u64* slots, u32 nthreads, u32 tid, size_t i, size_t n, size_t next, PowTab xtab, u64 x_start, const u64* __restrict__ zi, u32 zi_mask) {
    f3 v = f3_make(0, 0, 0);
    switch (o.kind) {
    case 0:
        v.c[0] = slots[(size_t)o.a * nthreads + tid];
        if (o.dim == 3) { v.c[1] = slots[(size_t)(o.a + 1) * nthreads + tid]; v.c[2] = slots[(size_t)(o.a + 2) * nthreads + tid]; }
        break;
    case 1: {
        size_t row = i + ((o.prime & 1) ? next : 0); if (row >= n) row -= n;
        const u64* p = secs.s[o.a].base + (size_t)o.b * secs.s[o.a].rows + row;
        if (o.prime & 2) {      // the program also WRITES this column: a non-coherent (ld.global.nc) load would be undefined
            const volatile u64* q = p;
            v.c[0] = q[0];
            if (o.dim == 3) { v.c[1] = q[secs.s[o.a].rows]; v.c[2] = q[2 * secs.s[o.a].rows]; }
        } else {
            v.c[0] = __ldg(p);
            if (o.dim == 3) { v.c[1] = __ldg(p + secs.s[o.a].rows); v.c[2] = __ldg(p + 2 * secs.s[o.a].rows); }
        }
        break; }
    case 2: v.c[0] = __ldg(consts + o.a); break;
    case 3: v.c[0] = __ldg(f3c + 3 * o.a); v.c[1] = __ldg(f3c + 3 * o.a + 1); v.c[2] = __ldg(f3c + 3 * o.a + 2); break;
    case 4: v.c[0] = gl_mul(x_start, powtab_get(xtab, i)); break;
    case 5: v.c[0] = __ldg(zi + (i & zi_mask)); break;
    }
    return v;
}

__global__ void __launch_bounds__(128) k_eval(const EvOp* __restrict__ ops, u32 n_ops, EvSecs secs, const u64* __restrict__ consts, const u64* __restrict__ f3c,
                                               PowTab xtab, u64 x_start, const u64* __restrict__ zi, u32 zi_mask, size_t n, size_t next) {
    extern __shared__ u64 slots[];
    const u32 tid = threadIdx.x, nt = blockDim.x;
    size_t i = (size_t)blockIdx.x * nt + tid;
    if (i >= n) return;
    for (u32 k = 0; k < n_ops; k++) {
        EvOp op = ops[k];
        f3 a = ev_load(op.s0, secs, consts, f3c, slots, nt, tid, i, n, next, xtab, x_start, zi, zi_mask);
        f3 r; u32 rd;
        if (op.opc == 3) { r = a; rd = op.s0.dim; }
        else {
            f3 b = ev_load(op.s1, secs, consts, f3c, slots, nt, tid, i, n, next, xtab, x_start, zi, zi_mask);
            u32 da = op.s0.dim, db = op.s1.dim;
            rd = (da == 3 || db == 3) ? 3 : 1;
            if (op.opc == 2) {
                if (da == 1 && db == 1) r = f3_make(gl_mul(a.c[0], b.c[0]), 0, 0);
                else if (da == 3 && db == 1) r = f3_muls(a, b.c[0]);
                else if (da == 1 && db == 3) r = f3_muls(b, a.c[0]);
                else r = f3_mul(a, b);
            } else if (rd == 1) {
                r = f3_make(op.opc == 0 ? gl_add(a.c[0], b.c[0]) : gl_sub(a.c[0], b.c[0]), 0, 0);
            } else {
                r = op.opc == 0 ? f3_add(a, b) : f3_sub(a, b);
            }
        }
        if (op.d.kind == 0) {
            slots[(size_t)op.d.a * nt + tid] = r.c[0];
            if (rd == 3) { slots[(size_t)(op.d.a + 1) * nt + tid] = r.c[1]; slots[(size_t)(op.d.a + 2) * nt + tid] = r.c[2]; }
        } else {
            size_t row = i + ((op.d.prime & 1) ? next : 0); if (row >= n) row -= n;
            u64* p = secs.s[op.d.a].base + (size_t)op.d.b * secs.s[op.d.a].rows + row;
            p[0] = r.c[0];
            if (rd == 3) { p[secs.s[op.d.a].rows] = r.c[1]; p[2 * secs.s[op.d.a].rows] = r.c[2]; }
        }
    }
}

// ------------------------------------------------------------------------------------------------ JIT specialisation
// The interpreter above decodes a 64-byte op per row and keeps temporaries in shared memory.  For the proofs that matter
// (many rows, the same program every time) the host turns the program into straight-line CUDA -- temporaries become
// registers, operand kinds and dimensions become code -- and compiles it once with NVRTC (jit.cpp); the cubin is cached by
// the hash of the op list.  Same arithmetic (field.cuh), same load / store order, so results are bit-identical; when NVRTC
// is unavailable or B200_JIT=0 the interpreter runs instead (still on the GPU).
// (section, column) pairs the program stores to.  Step programs do read back what they wrote (plookup step3: tmpExp and cm
// columns are written and read in one program), so loads of those columns must be ordinary coherent loads; only columns the
// program never writes may go through the read-only path (__ldg / ld.global.nc).
static std::set<std::pair<u32, u32>> written_columns(const std::vector<EvOp>& ops) {
    std::set<std::pair<u32, u32>> w;
    for (auto& op : ops) if (op.d.kind == 1) for (u32 l = 0; l < 3; l++) w.insert({op.d.a, op.d.b + l});     // a dim-1 result touches lane 0 only; over-approximate
    return w;
}
static bool reads_written(const std::set<std::pair<u32, u32>>& w, const EvOperand& x) {
    if (x.kind != 1) return false;
    for (u32 l = 0; l < x.dim; l++) if (w.count({x.a, x.b + l})) return true;
    return false;
}

std::string eval_jit_source(const EvProgram& p) {
    std::string o;
    const auto wcols = written_columns(p.ops);
    o += "#include \"field.cuh\"\n";
    o += "struct EvSection { u64* base; u64 rows; };\nstruct EvSecs { EvSection s[16]; };\n";
    // long programs: keep the extension-field products out of line so that the kernel stays within the instruction cache
    const bool big = p.ops.size() > 20;
    o += big ? "__device__ __noinline__ f3 j_f3_mul(f3 a, f3 b) { return f3_mul(a, b); }\n__device__ __noinline__ f3 j_f3_muls(f3 a, u64 b) { return f3_muls(a, b); }\n"
             : "#define j_f3_mul f3_mul\n#define j_f3_muls f3_muls\n";
    o += "extern \"C\" __global__ void __launch_bounds__(128) k_step(EvSecs secs, const u64* __restrict__ consts, const u64* __restrict__ f3c,\n"
         "        PowTab xtab, u64 x_start, const u64* __restrict__ zi, u32 zi_mask, size_t n, size_t next) {\n"
         "    const size_t i = (size_t)blockIdx.x * 128 + threadIdx.x;\n    if (i >= n) return;\n"
         "    size_t ip = i + next; if (ip >= n) ip -= n;\n";
    bool use_x = false, use_zi = false;
    for (auto& op : p.ops) for (const EvOperand* s : {&op.s0, &op.s1}) { if (op.opc == 3 && s == &op.s1) continue; use_x |= s->kind == 4; use_zi |= s->kind == 5; }
    for (u32 k = 0; k < p.n_slots; k++) o += "    u64 s" + std::to_string(k) + " = 0;\n";
    if (use_x) o += "    const u64 xval = gl_mul(x_start, powtab_get(xtab, i));\n";
    if (use_zi) o += "    const u64 zival = __ldg(zi + (i & zi_mask));\n";
    auto mem = [&](const EvOperand& x, u32 lane) {
        return std::string("secs.s[") + std::to_string(x.a) + "].base + (size_t)" + std::to_string(x.b + lane) + " * secs.s[" + std::to_string(x.a) + "].rows + " + ((x.prime & 1) ? "ip" : "i");
    };
    // expression of lane `lane` of operand x (lanes beyond its dimension are 0, like ev_load)
    auto lane_expr = [&](const EvOperand& x, u32 lane) -> std::string {
        if (lane >= x.dim) return "0ull";
        switch (x.kind) {
        case 0: return "s" + std::to_string(x.a + lane);
        case 1: return reads_written(wcols, x) ? "(*(const volatile u64*)(" + mem(x, lane) + "))" : "__ldg(" + mem(x, lane) + ")";
        case 2: return "__ldg(consts + " + std::to_string(x.a) + ")";
        case 3: return "__ldg(f3c + " + std::to_string(3 * x.a + lane) + ")";
        case 4: return "xval";
        case 5: return "zival";
        }
        throw std::runtime_error("bad operand kind");
    };
    auto f3_expr = [&](const EvOperand& x) { return "f3_make(" + lane_expr(x, 0) + ", " + lane_expr(x, 1) + ", " + lane_expr(x, 2) + ")"; };
    for (size_t k = 0; k < p.ops.size(); k++) {
        const EvOp& op = p.ops[k];
        u32 rd; std::string body;
        if (op.opc == 3) {
            rd = op.s0.dim;
            if (rd == 1) body = "const u64 r0 = " + lane_expr(op.s0, 0) + ";";
            else body = "const f3 r = " + f3_expr(op.s0) + ";";
        } else {
            const u32 da = op.s0.dim, db = op.s1.dim;
            rd = (da == 3 || db == 3) ? 3 : 1;
            if (rd == 1) {
                const char* fn = op.opc == 0 ? "gl_add" : (op.opc == 1 ? "gl_sub" : "gl_mul");
                body = std::string("const u64 r0 = ") + fn + "(" + lane_expr(op.s0, 0) + ", " + lane_expr(op.s1, 0) + ");";
            } else if (op.opc == 2) {
                if (da == 3 && db == 1) body = "const f3 r = j_f3_muls(" + f3_expr(op.s0) + ", " + lane_expr(op.s1, 0) + ");";
                else if (da == 1 && db == 3) body = "const f3 r = j_f3_muls(" + f3_expr(op.s1) + ", " + lane_expr(op.s0, 0) + ");";
                else body = "const f3 r = j_f3_mul(" + f3_expr(op.s0) + ", " + f3_expr(op.s1) + ");";
            } else {
                body = std::string("const f3 r = ") + (op.opc == 0 ? "f3_add(" : "f3_sub(") + f3_expr(op.s0) + ", " + f3_expr(op.s1) + ");";
            }
        }
        o += "    { " + body + " ";
        for (u32 l = 0; l < rd; l++) {
            std::string val = rd == 1 ? "r0" : "r.c[" + std::to_string(l) + "]";
            if (op.d.kind == 0) o += "s" + std::to_string(op.d.a + l) + " = " + val + "; ";
            else o += "*(" + mem(op.d, l) + ") = " + val + "; ";
        }
        o += "}\n";
    }
    o += "}\n";
    return o;
}

struct JitKernel;
bool jit_enabled();
std::string jit_compile_cubin(const std::string& src, std::string& err);
JitKernel* jit_load(const std::string& cubin, const char* name, std::string& err);
void jit_launch(JitKernel* k, unsigned grid, unsigned block, cudaStream_t st, void** args);

static JitKernel* jit_for(const EvProgram& p) {
    static std::map<std::pair<int, u64>, JitKernel*> cache;     // (device, hash of the op list) -> kernel, nullptr = failed once
    static std::mutex mu;
    if (!jit_enabled() || p.ops.size() > 20000) return nullptr;
    u64 h = 1469598103934665603ull;
    const unsigned char* b = reinterpret_cast<const unsigned char*>(p.ops.data());
    for (size_t i = 0; i < p.ops.size() * sizeof(EvOp); i++) { h ^= b[i]; h *= 1099511628211ull; }
    h ^= p.n_slots; h *= 1099511628211ull;
    int dev = current_device();
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find({dev, h});
    if (it != cache.end()) return it->second;
    std::string err;
    JitKernel* k = nullptr;
    std::string cubin = jit_compile_cubin(eval_jit_source(p), err);
    if (!cubin.empty()) k = jit_load(cubin, "k_step", err);
    if (!k && getenv("B200_JIT_VERBOSE")) fprintf(stderr, "b200zk: step-program JIT unavailable, using the interpreter kernel: %s\n", err.c_str());
    cache[{dev, h}] = k;
    return k;
}

void eval_program(const EvProgram& p, const EvSection* secs, int n_secs, const u64* h_f3consts, int n_f3,
                  DevPowTab x_tab, u64 x_start, const u64* d_zi, u32 zi_mask, size_t n, size_t next, double algo_bytes) {
    if (p.ops.empty() || n == 0) return;
    if (n_secs > EV_MAX_SECS) throw std::runtime_error("too many sections");
    EvSecs s{};
    for (int i = 0; i < n_secs; i++) s.s[i] = secs[i];
    // program + constants travel once per launch (a few KB)
    size_t ops_bytes = p.ops.size() * sizeof(EvOp), c_bytes = (p.consts.size() + 1) * 8, f_bytes = (size_t)(n_f3 + 1) * 24;
    static char* g_buf[16] = {nullptr}; static size_t g_cap[16] = {0};
    int dev0 = current_device();
    size_t need = ops_bytes + c_bytes + f_bytes + 64;
    if (g_cap[dev0] < need) { if (g_buf[dev0]) { B200_CUDA_CHECK(cudaStreamSynchronize(stream())); B200_CUDA_CHECK(cudaFree(g_buf[dev0])); } size_t cap = need < (1u << 20) ? (1u << 20) : need; B200_CUDA_CHECK(cudaMalloc(&g_buf[dev0], cap)); g_cap[dev0] = cap; }
    char* d = g_buf[dev0];
    EvOp* d_ops = reinterpret_cast<EvOp*>(d);
    u64* d_c = reinterpret_cast<u64*>(d + ((ops_bytes + 15) & ~(size_t)15));
    u64* d_f = d_c + p.consts.size() + 1;
    // interpreter copy of the op list: sources that read a column this program writes get bit 1 of `prime` (coherent load)
    static thread_local std::vector<EvOp> h_ops;
    h_ops = p.ops;
    { const auto wcols = written_columns(p.ops);
      for (auto& op : h_ops) { if (reads_written(wcols, op.s0)) op.s0.prime |= 2; if (op.opc != 3 && reads_written(wcols, op.s1)) op.s1.prime |= 2; } }
    B200_CUDA_CHECK(cudaMemcpyAsync(d_ops, h_ops.data(), ops_bytes, cudaMemcpyHostToDevice, stream()));
    if (!p.consts.empty()) B200_CUDA_CHECK(cudaMemcpyAsync(d_c, p.consts.data(), p.consts.size() * 8, cudaMemcpyHostToDevice, stream()));
    if (n_f3) B200_CUDA_CHECK(cudaMemcpyAsync(d_f, h_f3consts, (size_t)n_f3 * 24, cudaMemcpyHostToDevice, stream()));
    const u32 nt = 128;
    size_t smem = (size_t)(p.n_slots ? p.n_slots : 1) * nt * 8;
    if (smem > 200 * 1024) throw std::runtime_error("step program needs too many live temporaries");
    static bool attr[16] = {false};
    int dev = current_device();
    if (!attr[dev]) { B200_CUDA_CHECK(cudaFuncSetAttribute(k_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr[dev] = true; }
    JitKernel* jk = jit_for(p);
    {
        ScopedTimer t("step_program", algo_bytes);
        PowTab xt{x_tab.lo, x_tab.hi};
        if (jk) {
            const u64* cc = d_c; const u64* ff = d_f; const u64* zz = d_zi; u64 xs = x_start; u32 zm = zi_mask; size_t nn = n, nx = next;
            void* args[] = {&s, &cc, &ff, &xt, &xs, &zz, &zm, &nn, &nx};
            jit_launch(jk, (unsigned)((n + nt - 1) / nt), nt, stream(), args);
        } else {
            k_eval<<<(unsigned)((n + nt - 1) / nt), nt, smem, stream()>>>(d_ops, (u32)p.ops.size(), s, d_c, d_f, xt, x_start, d_zi, zi_mask, n, next);
        }
        launch_count_add(1);
    }
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaStreamSynchronize(stream()));     // the staging buffer is reused by the next launch
}

// ------------------------------------------------------------------------------------------------ Zi
__global__ void k_zh_inv(u64* out, u64 sn, u64 w, u32 m) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = gl_inv(gl_sub(gl_mul(sn, gl_pow(w, i)), 1));
}
void zh_inv_table(u64* d_out, unsigned nbits, unsigned ext_bits) {
    u64 sn = 49; for (unsigned i = 0; i < nbits; i++) sn = h_mul(sn, sn);
    u32 m = 1u << ext_bits;
    k_zh_inv<<<(m + 127) / 128, 128, 0, stream()>>>(d_out, sn, h_root(ext_bits), m);
    launch_count_add(1);
    B200_CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ quotient split
// qq1: [q_dim][n_ext] coefficients; qq2: [q_dim*q_deg][n] with qq2[p*q_dim+k][i] = qq1[k][p*n+i] * (49^-n)^p
__global__ void k_qsplit(const u64* __restrict__ qq1, u64* __restrict__ qq2, size_t n, size_t n_ext, u32 q_dim, u32 q_deg, u64 s_inv_n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 col = blockIdx.y, p = col / q_dim, k = col % q_dim;
    u64 v = qq1[(size_t)k * n_ext + (size_t)p * n + i];
    u64 s = 1; for (u32 e = 0; e < p; e++) s = gl_mul(s, s_inv_n);
    qq2[(size_t)col * n + i] = p ? gl_mul(v, s) : v;
}
void quotient_split(const u64* d_qq1, u64* d_qq2, size_t n, size_t n_ext, size_t q_dim, size_t q_deg, unsigned nbits) {
    if (q_deg == 0) return;
    u64 s_inv_n = h_pow(h_inv(49), 1ull << nbits);
    ScopedTimer t("quotient_split", 16.0 * (double)n * q_dim * q_deg);
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)(q_dim * q_deg));
    k_qsplit<<<grid, 256, 0, stream()>>>(d_qq1, d_qq2, n, n_ext, (u32)q_dim, (u32)q_deg, s_inv_n);
    launch_count_add(1);
    B200_CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ F3 powers
// out[e] = base^e for e < n (stark_gen.rs:416-427, LEv / LpEv before their iNTT).  Two-level: a table of base^j (j < 2^PW_LO) and
// one of base^(j 2^PW_LO) are built by a small kernel, the main kernel is one extension-field product per element with
// coalesced stores -- 24 B written per element and nothing read from HBM (both tables stay in L1/L2).
GL_D f3 f3_pow_dev(f3 a, u64 e) { f3 r = f3_make(1, 0, 0); while (e) { if (e & 1) r = f3_mul(r, a); a = f3_mul(a, a); e >>= 1; } return r; }
#define PW_LO 10
__global__ void k_f3_pow_tables(f3 base, u64* __restrict__ lo /* 3 x 2^PW_LO */, u64* __restrict__ hi /* 3 x n_hi */, u32 n_hi) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 n_lo = 1u << PW_LO;
    if (i < n_lo) { f3 v = f3_pow_dev(base, i); lo[i] = v.c[0]; lo[n_lo + i] = v.c[1]; lo[2 * n_lo + i] = v.c[2]; }
    else if (i < n_lo + n_hi) { u32 j = i - n_lo; f3 v = f3_pow_dev(base, (u64)j << PW_LO); hi[j] = v.c[0]; hi[n_hi + j] = v.c[1]; hi[2 * (size_t)n_hi + j] = v.c[2]; }
}
__global__ void __launch_bounds__(256) k_f3_powers(const u64* __restrict__ lo, const u64* __restrict__ hi, u32 n_hi, u64* __restrict__ out, size_t n) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const u32 n_lo = 1u << PW_LO, l = (u32)e & (n_lo - 1), h = (u32)(e >> PW_LO);
    f3 a = f3_make(__ldg(lo + l), __ldg(lo + n_lo + l), __ldg(lo + 2 * n_lo + l));
    f3 b = f3_make(__ldg(hi + h), __ldg(hi + n_hi + h), __ldg(hi + 2 * (size_t)n_hi + h));
    f3 r = f3_mul(a, b);
    out[e] = r.c[0]; out[n + e] = r.c[1]; out[2 * n + e] = r.c[2];
}
void f3_powers(const u64 base3[3], u64* d_out, size_t n) {
    f3 b; b.c[0] = base3[0]; b.c[1] = base3[1]; b.c[2] = base3[2];
    if (n == 0) return;
    const u32 n_lo = 1u << PW_LO, n_hi = (u32)((n + n_lo - 1) >> PW_LO);
    static u64* g_tab[16] = {nullptr}; static size_t g_cap[16] = {0};
    int dev = current_device();
    size_t need = 3 * ((size_t)n_lo + n_hi);
    if (g_cap[dev] < need) { if (g_tab[dev]) { B200_CUDA_CHECK(cudaStreamSynchronize(stream())); B200_CUDA_CHECK(cudaFree(g_tab[dev])); } B200_CUDA_CHECK(cudaMalloc(&g_tab[dev], need * 8)); g_cap[dev] = need; }
    u64* lo = g_tab[dev]; u64* hi = lo + 3 * (size_t)n_lo;
    ScopedTimer t("f3_powers", 24.0 * (double)n);
    k_f3_pow_tables<<<(n_lo + n_hi + 127) / 128, 128, 0, stream()>>>(b, lo, hi, n_hi);
    k_f3_powers<<<(unsigned)((n + 255) / 256), 256, 0, stream()>>>(lo, hi, n_hi, d_out, n);
    launch_count_add(2);
    B200_CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ eval dot
// acc = sum_k pol[k << ext_bits] * L[k];  pol: `dim` consecutive columns (stride col_stride), L: [3][n]
__global__ void __launch_bounds__(256) k_eval_dot(const u64* __restrict__ col0, size_t col_stride, int dim, unsigned ext_bits, const u64* __restrict__ L, size_t n, u64* __restrict__ partial) {
    __shared__ u64 red[3][256];
    f3 acc = f3_make(0, 0, 0);
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        f3 l = f3_make(L[k], L[n + k], L[2 * n + k]);
        size_t r = k << ext_bits;
        if (dim == 1) acc = f3_add(acc, f3_muls(l, __ldg(col0 + r)));
        else acc = f3_add(acc, f3_mul(f3_make(__ldg(col0 + r), __ldg(col0 + col_stride + r), __ldg(col0 + 2 * col_stride + r)), l));
    }
    red[0][threadIdx.x] = acc.c[0]; red[1][threadIdx.x] = acc.c[1]; red[2][threadIdx.x] = acc.c[2];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) for (int l = 0; l < 3; l++) red[l][threadIdx.x] = gl_add(red[l][threadIdx.x], red[l][threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) for (int l = 0; l < 3; l++) partial[3 * blockIdx.x + l] = red[l][0];
}
// Up to EVD_MAX evaluations that share one Lagrange row in ONE launch: the row (24 B per point, the larger operand) is read once per
// group instead of once per evaluation, and the group costs one device-to-host read and one synchronisation instead of one each.
#define EVD_MAX 8
struct EvalDotArgs { const u64* col0[EVD_MAX]; u64 stride[EVD_MAX]; int dim[EVD_MAX]; int m; };
template <int M> __global__ void __launch_bounds__(256) k_eval_dot_multi(EvalDotArgs a, unsigned ext_bits, const u64* __restrict__ L, size_t n, u64* __restrict__ partial) {
    __shared__ u64 red[8][EVD_MAX * 3];
    f3 acc[M];
#pragma unroll
    for (int e = 0; e < M; e++) acc[e] = f3_make(0, 0, 0);
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const f3 l = f3_make(__ldg(L + k), __ldg(L + n + k), __ldg(L + 2 * n + k));
        const size_t r = k << ext_bits;
#pragma unroll
        for (int e = 0; e < M; e++) {
            const u64* c = a.col0[e];
            if (a.dim[e] == 1) acc[e] = f3_add(acc[e], f3_muls(l, __ldg(c + r)));
            else acc[e] = f3_add(acc[e], f3_mul(f3_make(__ldg(c + r), __ldg(c + a.stride[e] + r), __ldg(c + 2 * a.stride[e] + r)), l));
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < M; e++) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            u64 v = acc[e].c[c];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const u64 o = gl_pack(__shfl_down_sync(0xffffffffu, (u32)v, off), __shfl_down_sync(0xffffffffu, (u32)(v >> 32), off));
                v = gl_add(v, o);
            }
            if (lane == 0) red[warp][3 * e + c] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x < M * 3) {
        u64 v = red[0][threadIdx.x];
        for (int w = 1; w < 8; w++) v = gl_add(v, red[w][threadIdx.x]);
        partial[(size_t)blockIdx.x * EVD_MAX * 3 + threadIdx.x] = v;
    }
}
// cols[e] / strides[e] / dims[e], e < m <= EVD_MAX; out: m x 3
void eval_dot_multi(const u64* const* d_col0, const size_t* col_stride, const int* dim, int m, unsigned ext_bits, const u64* d_L, size_t n, u64* out) {
    if (m <= 0 || m > EVD_MAX) throw std::runtime_error("eval_dot_multi: bad group size");
    unsigned blocks = (unsigned)((n + 256 * 16 - 1) / (256 * 16)); if (blocks > 1024) blocks = 1024; if (blocks == 0) blocks = 1;
    static u64* d_part[B200_MAX_DEVICES] = {nullptr};
    int dev = current_device();
    if (!d_part[dev]) B200_CUDA_CHECK(cudaMalloc(&d_part[dev], 1024 * EVD_MAX * 24));
    EvalDotArgs a; a.m = m; double bytes = 0;
    for (int e = 0; e < EVD_MAX; e++) { a.col0[e] = e < m ? d_col0[e] : nullptr; a.stride[e] = e < m ? col_stride[e] : 0; a.dim[e] = e < m ? dim[e] : 0; if (e < m) bytes += 8.0 * dim[e]; }
    {
        ScopedTimer t("eval_dot", (double)n * (bytes + 24.0));
        switch (m) {
        case 1: k_eval_dot_multi<1><<<blocks, 256, 0, stream()>>>(a, ext_bits, d_L, n, d_part[dev]); break;
        case 2: k_eval_dot_multi<2><<<blocks, 256, 0, stream()>>>(a, ext_bits, d_L, n, d_part[dev]); break;
        case 3: k_eval_dot_multi<3><<<blocks, 256, 0, stream()>>>(a, ext_bits, d_L, n, d_part[dev]); break;
        case 4: k_eval_dot_multi<4><<<blocks, 256, 0, stream()>>>(a, ext_bits, d_L, n, d_part[dev]); break;
        case 5: k_eval_dot_multi<5><<<blocks, 256, 0, stream()>>>(a, ext_bits, d_L, n, d_part[dev]); break;
        case 6: k_eval_dot_multi<6><<<blocks, 256, 0, stream()>>>(a, ext_bits, d_L, n, d_part[dev]); break;
        case 7: k_eval_dot_multi<7><<<blocks, 256, 0, stream()>>>(a, ext_bits, d_L, n, d_part[dev]); break;
        default: k_eval_dot_multi<8><<<blocks, 256, 0, stream()>>>(a, ext_bits, d_L, n, d_part[dev]); break;
        }
        launch_count_add(1);
    }
    std::vector<u64> h((size_t)blocks * EVD_MAX * 3);
    B200_CUDA_CHECK(cudaMemcpyAsync(h.data(), d_part[dev], h.size() * 8, cudaMemcpyDeviceToHost, stream()));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
    for (int e = 0; e < m; e++) for (int l = 0; l < 3; l++) {
        u64 v = 0;
        for (unsigned b = 0; b < blocks; b++) v = h_add(v, h[(size_t)b * EVD_MAX * 3 + 3 * e + l]);
        out[3 * e + l] = v;
    }
}
void eval_dot(const u64* d_col0, size_t col_stride, int dim, unsigned ext_bits, const u64* d_L, size_t n, u64 out3[3]) {
    unsigned blocks = (unsigned)((n + 256 * 16 - 1) / (256 * 16)); if (blocks > 1024) blocks = 1024; if (blocks == 0) blocks = 1;
    static u64* d_part[16] = {nullptr};
    int dev = current_device();
    if (!d_part[dev]) B200_CUDA_CHECK(cudaMalloc(&d_part[dev], 1024 * 24));
    {
        ScopedTimer t("eval_dot", (double)n * (8.0 * dim + 24.0));
        k_eval_dot<<<blocks, 256, 0, stream()>>>(d_col0, col_stride, dim, ext_bits, d_L, n, d_part[dev]);
        launch_count_add(1);
    }
    std::vector<u64> h(blocks * 3);
    B200_CUDA_CHECK(cudaMemcpyAsync(h.data(), d_part[dev], blocks * 24, cudaMemcpyDeviceToHost, stream()));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
    u64 a[3] = {0, 0, 0};
    for (unsigned b = 0; b < blocks; b++) for (int l = 0; l < 3; l++) a[l] = h_add(a[l], h[3 * b + l]);
    out3[0] = a[0]; out3[1] = a[1]; out3[2] = a[2];
}

// ------------------------------------------------------------------------------------------------ x / (x - pt)
// out[k] = x_k / (x_k - pt) over the coset, x_k in the base field, pt in GF(p^3) (stark_gen.rs:481-522).  The reference
// inverts N_ext extension-field denominators with two serial batch_inverse calls; field inverses are unique, so any
// exact method gives the same values.  Here: with a = x - pt0, b = -pt1, c = -pt2 (b, c constant over the coset) the inverse
// of (a, b, c) is (i1, i2, i3) / t (f3g.rs:207-235) where
//     t  = ((k2 - a) a + k1) a + k0          k2 = -2c, k1 = 3bc + bb - cc, k0 = -bbb + bcc - ccc      (the norm, a base-field cubic)
//     i1 = (-a - 2c) a + m1                  m1 = bc + bb - cc
//     i2 = b a - cc,   i3 = c a + cc - bb
// so only BASE-field inversions remain: Montgomery batch inversion of t over XB elements per thread (strided, coalesced).
// 12 base-field products per element + one gl_inv per XB, against 27 + an extension-field inverse in the first version.
#define XB 16
struct XdivConsts { u64 p0, b, c, c2, k2, k1, k0, m1, ncc, ccmbb; u64 sc[3]; u32 has_scale; };
__global__ void __launch_bounds__(128) k_xdivxsub(PowTab xtab, u64 x_start, size_t n_ext, XdivConsts q, u64* __restrict__ out) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    u64 pre[XB], xs[XB], ts[XB];
    u64 acc = 1;
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < XB; j++) {
        size_t k = t + (size_t)j * stride;
        if (k < n_ext) {
            // weak representatives (any u64 of the right residue) wherever the consumer is a product: the additive constants ride on the
            // products' 128-bit sums (gl_maddw) and nothing is canonicalised until the three output coordinates
            u64 x = gl_mul(x_start, powtab_getw(xtab, (u32)k));
            u64 a = gl_sub(x, q.p0);
            u64 tn = gl_maddw(gl_maddw(gl_sub(q.k2, a), a, q.k1), a, q.k0);
            xs[j] = x; ts[j] = tn;
            pre[j] = acc;                                   // product of the previous norms
            acc = gl_mulw(acc, tn);
            cnt = j + 1;
        }
    }
    if (cnt == 0) return;
    u64 inv = gl_inv(acc);
#pragma unroll
    for (int j = XB - 1; j >= 0; j--) {
        if (j < cnt) {
            size_t k = t + (size_t)j * stride;
            u64 ti = gl_mulw(inv, pre[j]);                   // 1 / t_j
            inv = gl_mulw(inv, ts[j]);
            u64 x = xs[j], a = gl_sub(x, q.p0);
            u64 s = gl_mulw(x, ti);                          // x / t
            u64 i1 = gl_maddw(gl_sub(gl_neg(a), q.c2), a, q.m1);
            u64 i2 = gl_maddw(q.b, a, q.ncc);
            u64 i3 = gl_maddw(q.c, a, q.ccmbb);
            f3 r;
            if (q.has_scale) r = f3_mul(f3_make(gl_mulw(i1, s), gl_mulw(i2, s), gl_mulw(i3, s)), f3_make(q.sc[0], q.sc[1], q.sc[2]));
            else r = f3_make(gl_mul(i1, s), gl_mul(i2, s), gl_mul(i3, s));
            out[k] = r.c[0]; out[n_ext + k] = r.c[1]; out[2 * n_ext + k] = r.c[2];
        }
    }
}
// scale3 != nullptr: every value is multiplied by that GF(p^3) constant (lagrange_row below)
void xdivxsub(DevPowTab x_tab, u64 x_start, size_t n_ext, const u64 pt3[3], u64* d_out, const u64* scale3, const char* timer_name) {
    XdivConsts q;
    q.has_scale = scale3 ? 1u : 0u; for (int i = 0; i < 3; i++) q.sc[i] = scale3 ? scale3[i] : 0;
    const u64 b = h_sub(0, pt3[1]), c = h_sub(0, pt3[2]);
    const u64 bb = h_mul(b, b), cc = h_mul(c, c), bc = h_mul(b, c);
    q.p0 = pt3[0]; q.b = b; q.c = c; q.c2 = h_add(c, c); q.ncc = h_sub(0, cc); q.ccmbb = h_sub(cc, bb);
    q.k2 = h_sub(0, q.c2);
    q.k1 = h_sub(h_add(h_add(bc, h_add(bc, bc)), bb), cc);
    q.k0 = h_sub(h_sub(h_mul(bc, c), h_mul(bb, b)), h_mul(cc, c));
    q.m1 = h_sub(h_add(bc, bb), cc);
    size_t nthreads = (n_ext + XB - 1) / XB;
    unsigned blocks = (unsigned)((nthreads + 127) / 128);
    ScopedTimer t(timer_name, 24.0 * (double)n_ext);
    PowTab xt{x_tab.lo, x_tab.hi};
    k_xdivxsub<<<blocks, 128, 0, stream()>>>(xt, x_start, n_ext, q, d_out);
    launch_count_add(1);
    B200_CUDA_CHECK(cudaGetLastError());
}

// L[k] = iNTT((y^i)_i)[k] for y in GF(p^3) \ GF(p), i < n = 2^nbits -- what the reference obtains by filling LEv[i] = (xi / shift)^i and
// running a size-n inverse transform over extension elements (stark_gen.rs:416-436, fft.rs:72-83).  The geometric sum has a closed form,
//     L[k] = (1/n) sum_i (y w^-k)^i = (y^n - 1) / (n (y w^-k - 1)) = -c * x_k / (x_k - y),   x_k = w^k,  c = (y^n - 1) / n,
// i.e. the x / (x - pt) kernel above on the un-shifted domain times one constant: 24 B written per element instead of a table fill
// and three NTT passes over three columns.  Same field elements, so the evaluations (and the proof) are bit-identical.
void lagrange_row(DevPowTab x_n_tab, unsigned nbits, const u64 y3[3], u64* d_out) {
    const size_t n = (size_t)1 << nbits;
    f3 y = f3_make(y3[0], y3[1], y3[2]), yn = y;
    for (unsigned i = 0; i < nbits; i++) yn = f3_mul(yn, yn);
    f3 c = f3_muls(f3_sub(yn, f3_make(1, 0, 0)), h_inv(n % GL_P));
    const u64 negc[3] = {gl_neg(c.c[0]), gl_neg(c.c[1]), gl_neg(c.c[2])};
    xdivxsub(x_n_tab, 1, n, y3, d_out, negc, "lagrange_row");
}

// ------------------------------------------------------------------------------------------------ FRI fold
// out[g] = eval( iNTT_{n_x}( pol[i*pol2_n + g] )_i * (sinv0 * winv^g)^i , special_x )      (fri.rs:112-126)
#define FRI_MAX_NX 64
__global__ void __launch_bounds__(128) k_fri_fold(const u64* __restrict__ pol, u64* __restrict__ out, u32 pol_bits, u32 red_bits, u64 sinv0, f3 sx,
                                                   PowTab wi_tab, const u64* __restrict__ stage_wi /* w_{n_x}^-e */, u64 nx_inv) {
    const size_t n = (size_t)1 << pol_bits, n_x = (size_t)1 << red_bits, pol2_n = n >> red_bits;
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= pol2_n) return;
    f3 pp[FRI_MAX_NX];
    for (u32 i = 0; i < n_x; i++) {
        size_t a = (size_t)i * pol2_n + g;
        pp[i] = f3_make(pol[a], pol[n + a], pol[2 * n + a]);
    }
    // inverse DFT: radix-2 DIF with inverse roots (natural in, bit-reversed out), then unscramble on read
    for (int s = (int)red_bits - 1; s >= 0; s--) {
        u32 half = 1u << s;
        for (u32 pi = 0; pi < n_x / 2; pi++) {
            u32 j = pi & (half - 1), p = ((pi >> s) << (s + 1)) | j;
            f3 u = pp[p], v = pp[p + half];
            pp[p] = f3_add(u, v);
            f3 d = f3_sub(u, v);
            u32 te = j << (red_bits - 1 - s);
            pp[p + half] = te ? f3_muls(d, stage_wi[te]) : d;
        }
    }
    u64 acc = gl_mul(sinv0, powtab_get(wi_tab, g));
    // coefficients c_i = pp[brev(i)] / n_x ; res = sum_i c_i acc^i sx^i via Horner from the top
    f3 res = f3_make(0, 0, 0);
    u64 r = gl_pow(acc, n_x - 1);
    u64 acc_inv_step = 0; (void)acc_inv_step;
    // Horner needs c_i * acc^i for descending i: recompute powers ascending into a small array
    u64 pw[FRI_MAX_NX];
    pw[0] = nx_inv; for (u32 i = 1; i < n_x; i++) pw[i] = gl_mul(pw[i - 1], acc);
    (void)r;
    for (int i = (int)n_x - 1; i >= 0; i--) {
        u32 bi = red_bits ? (__brev((u32)i) >> (32 - red_bits)) : 0;
        f3 c = f3_muls(pp[bi], pw[i]);
        res = (i == (int)n_x - 1) ? c : f3_add(f3_mul(res, sx), c);
    }
    out[g] = res.c[0]; out[pol2_n + g] = res.c[1]; out[2 * pol2_n + g] = res.c[2];
}
// The same fold for reductions of 1..4 bits per step (every step of the reference's starkStructs) with everything in registers: the
// generic kernel above indexes its arrays at run time, which puts them in local memory.  The evaluation at special_x is not a Horner chain
// of GF(p^3) products but ONE lazy dot product  sum_i (c_i acc^i) * sx^i  with the powers of special_x as kernel arguments: ten
// multiply-accumulates per term and three reductions in all (field.cuh gl_acc).
struct FriSx { u64 s[16][3]; };
__host__ __device__ constexpr u32 fri_brev(u32 i, int bits) { u32 r = 0; for (int b = 0; b < bits; b++) if (i & (1u << b)) r |= 1u << (bits - 1 - b); return r; }
template <int RB> __global__ void __launch_bounds__(128) k_fri_fold_t(const u64* __restrict__ pol, u64* __restrict__ out, u32 pol_bits, u64 sinv0, FriSx SX,
                                                                       PowTab wi_tab, const u64* __restrict__ stage_wi /* w_{n_x}^-e */, u64 nx_inv) {
    constexpr int NX = 1 << RB;
    const size_t n = (size_t)1 << pol_bits, pol2_n = n >> RB;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= pol2_n) return;
    f3 pp[NX];
#pragma unroll
    for (int i = 0; i < NX; i++) { const size_t a = (size_t)i * pol2_n + g; pp[i] = f3_make(__ldg(pol + a), __ldg(pol + n + a), __ldg(pol + 2 * n + a)); }
#pragma unroll
    for (int s = RB - 1; s >= 0; s--) {
#pragma unroll
        for (int pi = 0; pi < NX / 2; pi++) {
            const int half = 1 << s, j = pi & (half - 1), p = ((pi >> s) << (s + 1)) | j, te = j << (RB - 1 - s);
            const f3 u = pp[p], v = pp[p + half];
            pp[p] = f3_add(u, v);
            const f3 d = f3_sub(u, v);
            pp[p + half] = te ? f3_muls(d, __ldg(stage_wi + te)) : d;
        }
    }
    const u64 acc = gl_mul(sinv0, powtab_get(wi_tab, g));
    gl_acc A0, A1, A2, AS;          // c0 = A0 + AS, c1 = A1 + AS, c2 = A2 with AS = d1 s2 + d2 s1
    u64 pw = nx_inv;                // acc^i / n_x
#pragma unroll
    for (int i = 0; i < NX; i++) {
        const f3 c = pp[fri_brev((u32)i, RB)];
        const u64 d0 = gl_mulw(c.c[0], pw), d1 = gl_mulw(c.c[1], pw), d2 = gl_mulw(c.c[2], pw);
        const u64 s0 = SX.s[i][0], s1 = SX.s[i][1], s2 = SX.s[i][2];
        if (i == 0) { A0 = gl_acc_mul(d0, s0); AS = gl_acc_mul(d1, s2); A1 = gl_acc_mul(d0, s1); A2 = gl_acc_mul(d0, s2); }
        else { gl_acc_mad(A0, d0, s0); gl_acc_mad(AS, d1, s2); gl_acc_mad(A1, d0, s1); gl_acc_mad(A2, d0, s2); }
        gl_acc_mad(AS, d2, s1);
        gl_acc_mad(A1, d1, s0); gl_acc_mad(A1, d2, s2);
        gl_acc_mad(A2, d1, s1); gl_acc_mad(A2, d2, s0); gl_acc_mad(A2, d2, s2);
        if (i + 1 < NX) pw = gl_mulw(pw, acc);
    }
    gl_acc_add(A0, AS); gl_acc_add(A1, AS);
    out[g] = gl_acc_red(A0); out[pol2_n + g] = gl_acc_red(A1); out[2 * pol2_n + g] = gl_acc_red(A2);
}
void fri_fold(const u64* d_pol, u64* d_out, unsigned pol_bits, unsigned red_bits, u64 sinv0, const u64 sx3[3]) {
    if (red_bits > 6) throw std::runtime_error("fri_fold: reduction of more than 6 bits per step is not supported");
    size_t pol2_n = (size_t)1 << (pol_bits - red_bits);
    f3 sx; sx.c[0] = sx3[0]; sx.c[1] = sx3[1]; sx.c[2] = sx3[2];
    DevPowTab wt = powtab(h_root_inv(pol_bits), pol_bits);
    DevPowTab st = powtab(h_root_inv(red_bits), red_bits > 0 ? red_bits : 1);   // lo table holds w_{n_x}^-e for e < 4096
    ScopedTimer t("fri_fold", 24.0 * ((double)((size_t)1 << pol_bits) + (double)pol2_n));
    PowTab w{wt.lo, wt.hi};
    const unsigned blocks = (unsigned)((pol2_n + 127) / 128);
    const u64 nx_inv = h_inv((1ull << red_bits) % GL_P);
    if (red_bits >= 1 && red_bits <= 4) {
        FriSx SX; f3 pw = f3_make(1, 0, 0);
        for (unsigned i = 0; i < 16; i++) { for (int l = 0; l < 3; l++) SX.s[i][l] = pw.c[l]; pw = f3_mul(pw, sx); }
        switch (red_bits) {
        case 1: k_fri_fold_t<1><<<blocks, 128, 0, stream()>>>(d_pol, d_out, pol_bits, sinv0, SX, w, st.lo, nx_inv); break;
        case 2: k_fri_fold_t<2><<<blocks, 128, 0, stream()>>>(d_pol, d_out, pol_bits, sinv0, SX, w, st.lo, nx_inv); break;
        case 3: k_fri_fold_t<3><<<blocks, 128, 0, stream()>>>(d_pol, d_out, pol_bits, sinv0, SX, w, st.lo, nx_inv); break;
        default: k_fri_fold_t<4><<<blocks, 128, 0, stream()>>>(d_pol, d_out, pol_bits, sinv0, SX, w, st.lo, nx_inv); break;
        }
    } else
    k_fri_fold<<<blocks, 128, 0, stream()>>>(d_pol, d_out, pol_bits, red_bits, sinv0, sx, w, st.lo, nx_inv);
    launch_count_add(1);
    B200_CUDA_CHECK(cudaGetLastError());
}


// ------------------------------------------------------------------------------------------------ synthetic trace
// Fibonacci trace of the reference fixtures (starky/data/fib.cm.gl: row i = (F_i, F_{i+1}), F_0 = 1, F_1 = 2 mod p),
// row-major N x 2, generated on the device for benchmarks: thread t jumps to row t*CH with a 2x2 matrix power.
#define FIB_CH 1024
__global__ void k_fib_trace(u64* __restrict__ out, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t r0 = t * FIB_CH;
    if (r0 >= n) return;
    // M = [[0,1],[1,1]];  (F_r, F_{r+1}) = M^r (1, 2)
    u64 m00 = 1, m01 = 0, m10 = 0, m11 = 1, b00 = 0, b01 = 1, b10 = 1, b11 = 1;
    for (size_t e = r0; e; e >>= 1) {
        if (e & 1) { u64 a = gl_add(gl_mul(m00, b00), gl_mul(m01, b10)), b = gl_add(gl_mul(m00, b01), gl_mul(m01, b11)),
                         c = gl_add(gl_mul(m10, b00), gl_mul(m11, b10)), d = gl_add(gl_mul(m10, b01), gl_mul(m11, b11)); m00 = a; m01 = b; m10 = c; m11 = d; }
        u64 a = gl_add(gl_mul(b00, b00), gl_mul(b01, b10)), b = gl_add(gl_mul(b00, b01), gl_mul(b01, b11)),
            c = gl_add(gl_mul(b10, b00), gl_mul(b11, b10)), d = gl_add(gl_mul(b10, b01), gl_mul(b11, b11)); b00 = a; b01 = b; b10 = c; b11 = d;
    }
    u64 x = gl_add(m00, gl_dbl(m01)), y = gl_add(m10, gl_dbl(m11));
    for (size_t r = r0; r < r0 + FIB_CH && r < n; r++) { out[2 * r] = x; out[2 * r + 1] = y; u64 z = gl_add(x, y); x = y; y = z; }
}
void fib_trace(u64* d_out_rowmajor, size_t n) {
    size_t nt = (n + FIB_CH - 1) / FIB_CH;
    k_fib_trace<<<(unsigned)((nt + 63) / 64), 64, 0, stream()>>>(d_out_rowmajor, n);
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
}

}  // namespace b200
