/*
 * gl_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's (0xEigenLabs/eigen-zkvm, crate `starky`) CPU algorithms
 * for the Goldilocks STARK hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the checker or as the
 * timed CPU baseline -- never on the product path.
 *
 * Parity status: PINNED.  The functions below reproduce every known-answer vector the reference's
 * own tests carry for this path (tests/test_oracle_kats.py):
 *   Poseidon        starky/src/poseidon_opt.rs:219-262
 *   LinearHash      starky/src/linearhash.rs:311-362
 *   MerkleTreeGL    starky/src/merklehash.rs:469-497, 519-545
 *   F3G             starky/src/f3g.rs:619-624, 642-652
 *   LDE+Merkle      starky/src/stark_setup.rs:100-116 (const root of data/fib.const.gl)
 *   blocked-vs-simple NTT equivalence  starky/src/fft_p.rs:372-477 (we restate the simple one)
 *
 * All values are canonical u64 in [0,p), p = 2^64 - 2^32 + 1 (fields/src/field_gl.rs:12).  The
 * reference keeps Montgomery form internally (field_gl.rs:525-538) but every file / proof / KAT
 * carries canonical values (as_int(), field_gl.rs:542-544), which is what this oracle computes on.
 *
 * Matrices are ROW-MAJOR [row][col] exactly like the reference buffers (polsarray.rs:219-227).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "poseidon_gl_params.h"

typedef uint64_t u64;
typedef unsigned __int128 u128;
#define GLP 0xFFFFFFFF00000001ULL

/* ---------------------------------------------------------------- field_gl.rs:385-463 (add/sub/mul) */
static inline u64 gl_add(u64 a, u64 b) { u64 s = a + b; if (s < a || s >= GLP) s -= GLP; return s; }
static inline u64 gl_sub(u64 a, u64 b) { return a >= b ? a - b : a + (GLP - b); }
static inline u64 gl_neg(u64 a) { return a ? GLP - a : 0; }
/* 128-bit product reduced with 2^64 = 2^32-1, 2^96 = -1 (mod p); same value as Montgomery mul + as_int */
static inline u64 gl_red128(u128 x) {
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hh = hi >> 32, hl = hi & 0xFFFFFFFFULL;
    u64 t = lo - hh; if (lo < hh) t -= 0xFFFFFFFFULL;           /* borrow: +p == -(2^32-1) mod 2^64 */
    u64 m = hl * 0xFFFFFFFFULL;                                 /* hl*(2^32-1) < 2^64 */
    u64 r = t + m; if (r < t) r += 0xFFFFFFFFULL;               /* carry: 2^64 == 2^32-1 */
    if (r >= GLP) r -= GLP;
    return r;
}
static inline u64 gl_mul(u64 a, u64 b) { return gl_red128((u128)a * b); }
static u64 gl_pow(u64 a, u64 e) { u64 r = 1; while (e) { if (e & 1) r = gl_mul(r, a); a = gl_mul(a, a); e >>= 1; } return r; }
static inline u64 gl_inv(u64 a) { return gl_pow(a, GLP - 2); }

u64 ora_gl_add(u64 a, u64 b) { return gl_add(a % GLP, b % GLP); }
u64 ora_gl_sub(u64 a, u64 b) { return gl_sub(a % GLP, b % GLP); }
u64 ora_gl_mul(u64 a, u64 b) { return gl_mul(a % GLP, b % GLP); }
u64 ora_gl_mul_slow(u64 a, u64 b) { return (u64)(((u128)(a % GLP) * (b % GLP)) % GLP); }
u64 ora_gl_inv(u64 a) { return gl_inv(a % GLP); }
u64 ora_gl_pow(u64 a, u64 e) { return gl_pow(a % GLP, e); }

/* ---------------------------------------------------------------- constant.rs:52-68 (SHIFT, MG) */
#define GL_SHIFT 49ULL
static u64 MG_W[33], MG_WI[33];
static int mg_ready = 0;
static void mg_init(void) {
    if (mg_ready) return;
    MG_W[32] = gl_pow(7, 0xFFFFFFFFULL);          /* 7^(2^32-1): primitive 2^32-th root */
    MG_WI[32] = gl_inv(MG_W[32]);
    for (int n = 31; n >= 0; n--) { MG_W[n] = gl_mul(MG_W[n + 1], MG_W[n + 1]); MG_WI[n] = gl_mul(MG_WI[n + 1], MG_WI[n + 1]); }
    mg_ready = 1;
}
u64 ora_root(unsigned k) { mg_init(); return MG_W[k]; }
u64 ora_root_inv(unsigned k) { mg_init(); return MG_WI[k]; }

/* ---------------------------------------------------------------- f3g.rs:207-235, 407-449 */
typedef struct { u64 c[3]; } f3;
static inline f3 f3_add(f3 a, f3 b) { f3 r = {{gl_add(a.c[0], b.c[0]), gl_add(a.c[1], b.c[1]), gl_add(a.c[2], b.c[2])}}; return r; }
static inline f3 f3_sub(f3 a, f3 b) { f3 r = {{gl_sub(a.c[0], b.c[0]), gl_sub(a.c[1], b.c[1]), gl_sub(a.c[2], b.c[2])}}; return r; }
static inline f3 f3_mul(f3 a, f3 b) {   /* x^3 = x + 1 ; f3g.rs:419-431 */
    u64 A = gl_mul(gl_add(a.c[0], a.c[1]), gl_add(b.c[0], b.c[1]));
    u64 B = gl_mul(gl_add(a.c[0], a.c[2]), gl_add(b.c[0], b.c[2]));
    u64 C = gl_mul(gl_add(a.c[1], a.c[2]), gl_add(b.c[1], b.c[2]));
    u64 D = gl_mul(a.c[0], b.c[0]), E = gl_mul(a.c[1], b.c[1]), F = gl_mul(a.c[2], b.c[2]);
    u64 G = gl_sub(D, E);
    f3 r = {{gl_sub(gl_add(C, G), F), gl_sub(gl_sub(gl_sub(gl_add(A, C), E), E), D), gl_sub(B, G)}};
    return r;
}
static inline f3 f3_muls(f3 a, u64 s) { f3 r = {{gl_mul(a.c[0], s), gl_mul(a.c[1], s), gl_mul(a.c[2], s)}}; return r; }
static f3 f3_inv(f3 x) {                /* f3g.rs:207-235 */
    u64 a = x.c[0], b = x.c[1], c = x.c[2];
    u64 aa = gl_mul(a, a), ac = gl_mul(a, c), ba = gl_mul(b, a), bb = gl_mul(b, b), bc = gl_mul(b, c), cc = gl_mul(c, c);
    u64 aaa = gl_mul(aa, a), aac = gl_mul(aa, c), abc = gl_mul(ba, c), abb = gl_mul(ba, b), acc = gl_mul(ac, c);
    u64 bbb = gl_mul(bb, b), bcc = gl_mul(bc, c), ccc = gl_mul(cc, c);
    u64 t = gl_neg(aaa);
    t = gl_sub(t, aac); t = gl_sub(t, aac); t = gl_add(t, abc); t = gl_add(t, abc); t = gl_add(t, abc);
    t = gl_add(t, abb); t = gl_sub(t, acc); t = gl_sub(t, bbb); t = gl_add(t, bcc); t = gl_sub(t, ccc);
    u64 ti = gl_inv(t);
    u64 i1 = gl_neg(aa); i1 = gl_sub(i1, ac); i1 = gl_sub(i1, ac); i1 = gl_add(i1, bc); i1 = gl_add(i1, bb); i1 = gl_sub(i1, cc);
    u64 i2 = gl_sub(ba, cc);
    u64 i3 = gl_add(gl_add(gl_neg(bb), ac), cc);
    f3 r = {{gl_mul(i1, ti), gl_mul(i2, ti), gl_mul(i3, ti)}};
    return r;
}
void ora_f3_mul(const u64 *a, const u64 *b, u64 *o) { f3 x = {{a[0], a[1], a[2]}}, y = {{b[0], b[1], b[2]}}; f3 r = f3_mul(x, y); memcpy(o, r.c, 24); }
void ora_f3_inv(const u64 *a, u64 *o) { f3 x = {{a[0], a[1], a[2]}}; f3 r = f3_inv(x); memcpy(o, r.c, 24); }

/* ---------------------------------------------------------------- poseidon_opt.rs:80-200 */
static u64 PC[118], PM[144], PP[144], PS[506];
static int pos_ready = 0;
static void pos_init(void) {
    if (pos_ready) return;
    for (int i = 0; i < 118; i++) PC[i] = ORA_POS_C[i] % GLP;
    for (int i = 0; i < 144; i++) { PM[i] = ORA_POS_M[i] % GLP; PP[i] = ORA_POS_P[i] % GLP; }
    for (int i = 0; i < 506; i++) PS[i] = ORA_POS_S[i] % GLP;
    pos_ready = 1;
}
static inline u64 pow7(u64 x) { u64 x2 = gl_mul(x, x), x3 = gl_mul(x2, x), x6 = gl_mul(x3, x3); return gl_mul(x6, x); }
static inline void mat12(const u64 *Mx, u64 *st) {         /* st'[i] = sum_j Mx[j][i]*st[j] */
    u64 t[12];
    for (int i = 0; i < 12; i++) { u64 acc = 0; for (int j = 0; j < 12; j++) acc = gl_add(acc, gl_mul(Mx[j * 12 + i], st[j])); t[i] = acc; }
    memcpy(st, t, sizeof t);
}
/* state = in[0..8] || cap[0..4]; out = 12 lanes (callers take the first 4 when hashing) */
static void poseidon12(const u64 *in8, const u64 *cap4, u64 *out12) {
    u64 st[12];
    for (int i = 0; i < 8; i++) st[i] = in8[i];
    for (int i = 0; i < 4; i++) st[8 + i] = cap4[i];
    for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], PC[i]);
    for (int r = 0; r < 3; r++) {
        for (int i = 0; i < 12; i++) st[i] = gl_add(pow7(st[i]), PC[(r + 1) * 12 + i]);
        mat12(PM, st);
    }
    for (int i = 0; i < 12; i++) st[i] = gl_add(pow7(st[i]), PC[4 * 12 + i]);
    mat12(PP, st);
    for (int r = 0; r < 22; r++) {
        st[0] = gl_add(pow7(st[0]), PC[5 * 12 + r]);
        u64 s0 = 0;
        for (int j = 0; j < 12; j++) s0 = gl_add(s0, gl_mul(PS[23 * r + j], st[j]));
        for (int k = 1; k < 12; k++) st[k] = gl_add(st[k], gl_mul(PS[23 * r + 12 + k - 1], st[0]));
        st[0] = s0;
    }
    for (int r = 0; r < 3; r++) {
        for (int i = 0; i < 12; i++) st[i] = gl_add(pow7(st[i]), PC[5 * 12 + 22 + r * 12 + i]);
        mat12(PM, st);
    }
    for (int i = 0; i < 12; i++) st[i] = pow7(st[i]);
    mat12(PM, st);
    memcpy(out12, st, sizeof st);
}
void ora_poseidon(const u64 *in8, const u64 *cap4, u64 *out12) { pos_init(); poseidon12(in8, cap4, out12); }

/* ---------------------------------------------------------------- linearhash.rs:79-145 */
static void lh_inner(const u64 *v, size_t n, u64 *out4) {     /* _hash */
    u64 st[4] = {0, 0, 0, 0};
    if (n <= 4) { for (size_t i = 0; i < n; i++) st[i] = v[i]; memcpy(out4, st, 32); return; }
    u64 blk[8], o[12];
    size_t i = 0;
    while (i < n) {
        size_t m = n - i < 8 ? n - i : 8;
        for (size_t k = 0; k < 8; k++) blk[k] = k < m ? v[i + k] : 0;
        poseidon12(blk, st, o); memcpy(st, o, 32);
        i += m;
    }
    memcpy(out4, st, 32);
}
static void lh_hash(const u64 *v, size_t n, u64 *out4) {      /* hash(batch_size = 0) */
    u64 st[4] = {0, 0, 0, 0};
    if (n <= 4) { for (size_t i = 0; i < n; i++) st[i] = v[i]; memcpy(out4, st, 32); return; }
    size_t bs = (n + 3) / 4; if (bs < 8) bs = 8;
    size_t hsz = (n + bs - 1) / bs;
    u64 hashes[16];
    for (size_t c = 0; c < hsz; c++) { size_t m = n - c * bs < bs ? n - c * bs : bs; lh_inner(v + c * bs, m, hashes + 4 * c); }
    if (hsz * 4 <= 4) { memcpy(out4, hashes, 32); return; }
    lh_inner(hashes, hsz * 4, out4);
}
void ora_linearhash(const u64 *v, size_t n, u64 *out4) { pos_init(); lh_hash(v, n, out4); }

/* ---------------------------------------------------------------- merklehash.rs:47-61, 79-134, 293-346 */
size_t ora_merkle_n_nodes(size_t n_) {
    size_t n = n_, next_n = (n - 1) / 2 + 1, acc = next_n * 2;
    while (n > 1) { n = next_n; next_n = (n - 1) / 2 + 1; if (n > 1) acc += next_n * 2; else acc += 1; }
    return acc;
}
/* nodes: ora_merkle_n_nodes(height)*4 u64, zero-initialised by us; leaves row-major height x width (width may be 0) */
void ora_merkelize(const u64 *leaves, size_t width, size_t height, u64 *nodes) {
    pos_init();
    size_t nn = ora_merkle_n_nodes(height);
    memset(nodes, 0, nn * 32);
    if (width > 0) {
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)height; i++) lh_hash(leaves + (size_t)i * width, width, nodes + 4 * (size_t)i);
    }
    size_t n64 = height, next = (n64 - 1) / 2 + 1, p_in = 0, p_out = next * 2;
    while (n64 > 1) {
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)next; i++) {
            u64 o[12]; static const u64 z[4] = {0, 0, 0, 0};
            poseidon12(nodes + 4 * (p_in + 2 * (size_t)i), z, o);      /* left||right are adjacent */
            memcpy(nodes + 4 * (p_out + (size_t)i), o, 32);
        }
        n64 = next; next = (n64 - 1) / 2 + 1; p_in = p_out; p_out = p_in + next * 2;
    }
}
/* merkle_gen_merkle_proof (merklehash.rs:64-77): siblings bottom-up, 4 lanes each; returns depth */
size_t ora_merkle_proof(const u64 *nodes, size_t height, size_t idx, u64 *sib_out) {
    size_t n = height, offset = 0, d = 0;
    while (n > 1) {
        size_t si = idx ^ 1;
        memcpy(sib_out + 4 * d, nodes + 4 * (offset + si), 32);
        size_t next_n = (n - 1) / 2 + 1;
        offset += next_n * 2; idx >>= 1; n = next_n; d++;
    }
    return d;
}
/* merklehash.rs:184-233: recompute the root from a group opening */
void ora_merkle_root_from_proof(const u64 *vals, size_t width, const u64 *sibs, size_t depth, size_t idx, u64 *root4) {
    pos_init();
    u64 cur[4]; lh_hash(vals, width, cur);
    static const u64 z[4] = {0, 0, 0, 0};
    for (size_t d = 0; d < depth; d++) {
        u64 in[8], o[12];
        if ((idx & 1) == 0) { memcpy(in, cur, 32); memcpy(in + 4, sibs + 4 * d, 32); }
        else { memcpy(in, sibs + 4 * d, 32); memcpy(in + 4, cur, 32); }
        poseidon12(in, z, o); memcpy(cur, o, 32); idx >>= 1;
    }
    memcpy(root4, cur, 32);
}

/* ---------------------------------------------------------------- fft.rs:39-83 (textbook DIT), batched over columns */
static unsigned brev(unsigned x, unsigned bits) { unsigned r = 0; for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r; }
/* in-place natural-order NTT of every column of a row-major n x w matrix */
static void ntt_cols(u64 *buf, size_t w, unsigned bits) {
    mg_init();
    size_t n = (size_t)1 << bits;
    /* bit-reverse rows */
    for (size_t i = 0; i < n; i++) { size_t r = brev((unsigned)i, bits); if (r > i) for (size_t c = 0; c < w; c++) { u64 t = buf[i * w + c]; buf[i * w + c] = buf[r * w + c]; buf[r * w + c] = t; } }
    for (unsigned s = 1; s <= bits; s++) {
        size_t m = (size_t)1 << s, half = m >> 1;
        u64 winc = MG_W[s];
        u64 *tw = (u64 *)malloc(half * sizeof(u64));
        tw[0] = 1; for (size_t j = 1; j < half; j++) tw[j] = gl_mul(tw[j - 1], winc);
#pragma omp parallel for schedule(static)
        for (long kk = 0; kk < (long)(n / m); kk++) {
            size_t k = (size_t)kk * m;
            for (size_t j = 0; j < half; j++) {
                u64 wv = tw[j];
                u64 *pu = buf + (k + j) * w, *pt = buf + (k + j + half) * w;
                for (size_t c = 0; c < w; c++) { u64 t = gl_mul(wv, pt[c]); u64 u = pu[c]; pu[c] = gl_add(u, t); pt[c] = gl_sub(u, t); }
            }
        }
        free(tw);
    }
}
/* fft_p.rs:242-253 semantics (natural in / natural out); out may alias in */
void ora_ntt(const u64 *in, u64 *out, size_t w, unsigned bits) {
    size_t n = (size_t)1 << bits;
    if (out != in) memcpy(out, in, n * w * 8);
    ntt_cols(out, w, bits);
}
/* fft.rs:72-83: ifft = fft, reversed index, times 1/n */
void ora_intt(const u64 *in, u64 *out, size_t w, unsigned bits) {
    size_t n = (size_t)1 << bits;
    u64 *q = (u64 *)malloc(n * w * 8);
    memcpy(q, in, n * w * 8);
    ntt_cols(q, w, bits);
    u64 ninv = gl_inv((u64)n % GLP);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) {
        size_t src = ((size_t)i == 0) ? 0 : n - (size_t)i;
        for (size_t c = 0; c < w; c++) out[(size_t)i * w + c] = gl_mul(q[src * w + c], ninv);
    }
    free(q);
}
/* fft_p.rs:255-355 == polutils.rs:24-33 (extend_pol): coset LDE, shift 49; out is n_ext x w */
void ora_lde(const u64 *in, u64 *out, size_t w, unsigned bits, unsigned bits_ext) {
    if (w == 0) return;
    size_t n = (size_t)1 << bits, ne = (size_t)1 << bits_ext;
    u64 *co = (u64 *)malloc(n * w * 8);
    ora_intt(in, co, w, bits);
    memset(out, 0, ne * w * 8);
    u64 r = 1;
    for (size_t i = 0; i < n; i++) { for (size_t c = 0; c < w; c++) out[i * w + c] = gl_mul(co[i * w + c], r); r = gl_mul(r, GL_SHIFT); }
    free(co);
    ntt_cols(out, w, bits_ext);
}

/* ---------------------------------------------------------------- polutils.rs:35-53 (batch_inverse over F3G) */
void ora_f3_batch_inverse(const u64 *in, u64 *out, size_t n) {
    if (!n) return;
    f3 *tmp = (f3 *)malloc(n * sizeof(f3));
    const f3 *e = (const f3 *)in; f3 *res = (f3 *)out;
    tmp[0] = e[0];
    for (size_t i = 1; i < n; i++) tmp[i] = f3_mul(e[i], tmp[i - 1]);
    f3 z = f3_inv(tmp[n - 1]);
    for (size_t i = n - 1; i >= 1; i--) { f3 ei = e[i]; res[i] = f3_mul(z, tmp[i - 1]); z = f3_mul(z, ei); }
    res[0] = z;
    free(tmp);
}
/* a CPU prover would chunk this across threads; identical values (field inverses are unique) */
void ora_f3_batch_inverse_par(const u64 *in, u64 *out, size_t n) {
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    size_t chunk = (n + nt - 1) / nt;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < nt; t++) { size_t a = (size_t)t * chunk; if (a < n) { size_t m = n - a < chunk ? n - a : chunk; ora_f3_batch_inverse(in + 3 * a, out + 3 * a, m); } }
}

/* ---------------------------------------------------------------- stark_gen.rs:752-783 + interpreter.rs
 * Straight-line step program over rows.  Encoding (int64 words), produced by oracle/stark_oracle.py
 * from the codegen Sections:  per op: [opcode(0 add,1 sub,2 mul,3 copy), dest(5 words), srcA(5), srcB(5)]
 * operand = [kind, a, b, c, dim]:
 *   kind 0 tmp:     a = tmp id
 *   kind 1 mem:     a = section index, b = offset(col), c = prime(0/1); dim 1|3; address = off + ((i+next*prime)%n)*width
 *   kind 2 const64: a = index into consts[] (dim 1)
 *   kind 3 f3 const:a = index into f3consts[] (3 lanes; challenges / evals), dim 3
 *   kind 4 x:       x[i]  (dim 1)        kind 5 Zi: zi[i % zi_len] (dim 1)
 * Values carry a runtime dim like the reference F3G (f3g.rs:13-18); writes follow interpreter.rs:146-166:
 * a dim-1 value touches only lane 0 of a wider destination.
 */
typedef struct { u64 *base; size_t width; } ora_section;
typedef struct { f3 v; int dim; } val;
static inline val v_op(int op, val a, val b) {
    val r;
    if (op == 2) {
        if (a.dim == 1 && b.dim == 1) { r.dim = 1; r.v.c[0] = gl_mul(a.v.c[0], b.v.c[0]); r.v.c[1] = r.v.c[2] = 0; }
        else if (a.dim == 3 && b.dim == 1) { r.dim = 3; r.v = f3_muls(a.v, b.v.c[0]); }
        else if (a.dim == 1 && b.dim == 3) { r.dim = 3; r.v = f3_muls(b.v, a.v.c[0]); }
        else { r.dim = 3; r.v = f3_mul(a.v, b.v); }
        return r;
    }
    r.dim = (a.dim == 3 || b.dim == 3) ? 3 : 1;
    if (op == 0) r.v = f3_add(a.v, b.v); else r.v = f3_sub(a.v, b.v);     /* lanes 1,2 of a dim-1 value are 0 */
    return r;
}
void ora_eval_program(const int64_t *prog, size_t n_ops, size_t n_tmp,
                      ora_section *secs, const u64 *consts, const u64 *f3consts,
                      const u64 *x, const u64 *zi, size_t zi_len,
                      size_t n, size_t next) {
#pragma omp parallel
    {
        val *tmp = (val *)calloc(n_tmp ? n_tmp : 1, sizeof(val));
#pragma omp for schedule(static)
        for (long ii = 0; ii < (long)n; ii++) {
            size_t i = (size_t)ii;
            for (size_t k = 0; k < n_ops; k++) {
                const int64_t *o = prog + 16 * k;
                val s[2]; int nsrc = (o[0] == 3) ? 1 : 2;
                for (int q = 0; q < nsrc; q++) {
                    const int64_t *p = o + 6 + 5 * q; val v; v.v.c[0] = v.v.c[1] = v.v.c[2] = 0; v.dim = 1;
                    switch (p[0]) {
                    case 0: v = tmp[p[1]]; break;
                    case 1: { ora_section *sc = &secs[p[1]]; size_t row = (i + (p[3] ? next : 0)) % n; const u64 *a = sc->base + (size_t)p[2] + row * sc->width;
                              v.dim = (int)p[4]; v.v.c[0] = a[0]; if (v.dim == 3) { v.v.c[1] = a[1]; v.v.c[2] = a[2]; } break; }
                    case 2: v.v.c[0] = consts[p[1]]; break;
                    case 3: v.dim = 3; v.v.c[0] = f3consts[3 * p[1]]; v.v.c[1] = f3consts[3 * p[1] + 1]; v.v.c[2] = f3consts[3 * p[1] + 2]; break;
                    case 4: v.v.c[0] = x[i]; break;
                    case 5: v.v.c[0] = zi[i % zi_len]; break;
                    }
                    s[q] = v;
                }
                val r = (o[0] == 3) ? s[0] : v_op((int)o[0], s[0], s[1]);
                const int64_t *d = o + 1;
                if (d[0] == 0) tmp[d[1]] = r;
                else { ora_section *sc = &secs[d[1]]; size_t row = (i + (d[3] ? next : 0)) % n; u64 *a = sc->base + (size_t)d[2] + row * sc->width;
                       a[0] = r.v.c[0]; if (r.dim == 3) { a[1] = r.v.c[1]; a[2] = r.v.c[2]; } }
            }
        }
        free(tmp);
    }
}

/* ---------------------------------------------------------------- stark_gen.rs:236-247 x tables, :575-592 Zi */
void ora_x_table(u64 *out, size_t n, u64 start, u64 w) {    /* out[k] = start * w^k */
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    size_t chunk = (n + nt - 1) / nt;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < nt; t++) { size_t a = (size_t)t * chunk; if (a >= n) continue; size_t b = a + chunk < n ? a + chunk : n;
        u64 v = gl_mul(start, gl_pow(w, a)); for (size_t k = a; k < b; k++) { out[k] = v; v = gl_mul(v, w); } }
}
void ora_zh_inv(u64 *out, unsigned nbits, unsigned ext_bits) {
    mg_init();
    u64 sn = GL_SHIFT; for (unsigned i = 0; i < nbits; i++) sn = gl_mul(sn, sn);
    u64 w = 1; size_t m = (size_t)1 << ext_bits;
    for (size_t i = 0; i < m; i++) { out[i] = gl_inv(gl_sub(gl_mul(sn, w), 1)); w = gl_mul(w, MG_W[ext_bits]); }
}

/* ---------------------------------------------------------------- stark_gen.rs:375-396 (quotient split) */
/* qq1: n_ext x q_dim (coefficients); qq2: n_ext x (q_dim*q_deg), zero beyond row n */
void ora_quotient_split(const u64 *qq1, u64 *qq2, size_t n, size_t n_ext, size_t q_dim, size_t q_deg, unsigned nbits) {
    memset(qq2, 0, n_ext * q_dim * q_deg * 8);
    u64 shift_inv_n = gl_pow(gl_inv(GL_SHIFT), (u64)1 << nbits), cur = 1;
    for (size_t p = 0; p < q_deg; p++) {
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)n; i++) for (size_t k = 0; k < q_dim; k++)
            qq2[(size_t)i * q_dim * q_deg + q_dim * p + k] = gl_mul(qq1[p * n * q_dim + (size_t)i * q_dim + k], cur);
        cur = gl_mul(cur, shift_inv_n);
    }
}

/* ---------------------------------------------------------------- stark_gen.rs:416-466 (LEv + evals) */
/* powers: out[i] = base^i as F3G, n entries (3 lanes each) */
void ora_f3_powers(const u64 *base3, u64 *out, size_t n) {
    f3 b = {{base3[0], base3[1], base3[2]}}, cur = {{1, 0, 0}};
    for (size_t i = 0; i < n; i++) { memcpy(out + 3 * i, cur.c, 24); cur = f3_mul(cur, b); }
}
/* acc = sum_k pol[(k<<ext_bits)*size + off (+lane)] * L[k]  ; L is n x 3 */
void ora_eval_dot(const u64 *buf, size_t size, size_t off, int dim, unsigned ext_bits, const u64 *L, size_t n, u64 *out3) {
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    f3 *part = (f3 *)calloc(nt, sizeof(f3));
    size_t chunk = (n + nt - 1) / nt;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < nt; t++) {
        f3 acc = {{0, 0, 0}};
        size_t a = (size_t)t * chunk, b = a + chunk < n ? a + chunk : n;
        for (size_t k = a; k < b; k++) {
            const u64 *p = buf + (k << ext_bits) * size + off; f3 l = {{L[3 * k], L[3 * k + 1], L[3 * k + 2]}};
            if (dim == 1) acc = f3_add(acc, f3_muls(l, p[0])); else { f3 v = {{p[0], p[1], p[2]}}; acc = f3_add(acc, f3_mul(v, l)); }
        }
        part[t] = acc;
    }
    f3 acc = {{0, 0, 0}}; for (int t = 0; t < nt; t++) acc = f3_add(acc, part[t]);
    free(part); memcpy(out3, acc.c, 24);
}

/* ---------------------------------------------------------------- stark_gen.rs:481-522 (xDivXSubXi tables) */
void ora_xdivxsub(const u64 *x, size_t n_ext, const u64 *xi3, u64 *out /* n_ext x 3 */) {
    u64 *den = (u64 *)malloc(n_ext * 24);
#pragma omp parallel for schedule(static)
    for (long k = 0; k < (long)n_ext; k++) { den[3 * k] = gl_sub(x[k], xi3[0]); den[3 * k + 1] = gl_neg(xi3[1]); den[3 * k + 2] = gl_neg(xi3[2]); }
    ora_f3_batch_inverse_par(den, out, n_ext);
#pragma omp parallel for schedule(static)
    for (long k = 0; k < (long)n_ext; k++) { f3 v = {{out[3 * k], out[3 * k + 1], out[3 * k + 2]}}; v = f3_muls(v, x[k]); memcpy(out + 3 * k, v.c, 24); }
    free(den);
}

/* ---------------------------------------------------------------- fri.rs:101-151 (one fold step) */
/* pol: 2^pol_bits F3G (AoS 3 lanes); out: pol2_n = 2^(pol_bits-red_bits) F3G.  sinv0 = 49^-(2^folded) */
void ora_fri_fold(const u64 *pol, u64 *out, unsigned pol_bits, unsigned red_bits, u64 sinv0, const u64 *special_x3) {
    mg_init();
    size_t n = (size_t)1 << pol_bits, n_x = (size_t)1 << red_bits, pol2_n = n >> red_bits;
    u64 wi = MG_WI[pol_bits];
    f3 sx = {{special_x3[0], special_x3[1], special_x3[2]}};
    u64 ninv = gl_inv((u64)n_x);
#pragma omp parallel
    {
        f3 *pp = (f3 *)malloc(n_x * sizeof(f3)), *qq = (f3 *)malloc(n_x * sizeof(f3));
#pragma omp for schedule(static)
        for (long gg = 0; gg < (long)pol2_n; gg++) {
            size_t g = (size_t)gg;
            /* gather + textbook FFT (fft.rs:39-70) on F3G with base-field twiddles, then ifft reversal */
            for (size_t i = 0; i < n_x; i++) { size_t r = brev((unsigned)i, red_bits); memcpy(pp[r].c, pol + 3 * (i * pol2_n + g), 24); }
            for (unsigned s = 1; s <= red_bits; s++) {
                size_t m = (size_t)1 << s, half = m >> 1; u64 winc = MG_W[s];
                for (size_t k = 0; k < n_x; k += m) { u64 w = 1; for (size_t j = 0; j < half; j++) {
                    f3 t = f3_muls(pp[k + j + half], w), u = pp[k + j]; pp[k + j] = f3_add(u, t); pp[k + j + half] = f3_sub(u, t); w = gl_mul(w, winc); } }
            }
            for (size_t i = 0; i < n_x; i++) qq[i] = f3_muls(pp[i == 0 ? 0 : n_x - i], ninv);
            /* pol_mul_axi(c, 1, sinv*wi^g) then eval_pol at special_x (polutils.rs:5-22) */
            u64 acc = gl_mul(sinv0, gl_pow(wi, g)), r = 1;
            for (size_t i = 0; i < n_x; i++) { qq[i] = f3_muls(qq[i], r); r = gl_mul(r, acc); }
            f3 res = qq[n_x - 1];
            for (size_t i = n_x - 1; i-- > 0;) res = f3_add(f3_mul(res, sx), qq[i]);
            memcpy(out + 3 * g, res.c, 24);
        }
        free(pp); free(qq);
    }
}
/* fri.rs:299-317 get_transposed_buffer: pol (n F3G) -> (w=2^tbits rows) x (h*3) */
void ora_fri_transpose(const u64 *pol, u64 *out, size_t n, unsigned tbits) {
    size_t w = (size_t)1 << tbits, h = n / w;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)w; i++) for (size_t j = 0; j < h; j++) memcpy(out + ((size_t)i * h + j) * 3, pol + 3 * (j * w + (size_t)i), 24);
}

int ora_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void ora_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
