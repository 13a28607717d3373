/*
 * curves_oracle.c -- CPU ORACLE (test infrastructure) for multi-scalar multiplication on the four groups of the groth16 final
 * layer: BN254 G1 / G2 and BLS12-381 G1 / G2.  Generic over the base-field size (4 or 6 64-bit limbs) and the extension degree
 * (1 = G1 over Fp, 2 = G2 over Fp2 = Fp[u]/(u^2 + 1)); the moduli come from the caller (oracle/curves.py, which pins them
 * against the reference's sources).
 *
 * PARITY UNPINNED by the reference: its MSMs live in un-vendored crates (bellman_ce 0.3.2 / pairing_ce for BN254, bellperson
 * 0.26 + blstrs 0.7.1 for BLS12-381; Cargo.lock:668-670,731-733) and no reference test fixes an MSM output (proofs are
 * randomised, groth16/src/api.rs:154,173).  This file restates the published bellman `multiexp`: UNSIGNED c-bit windows,
 * 2^c - 1 buckets per window, running-sum bucket fold, Horner over windows with c doublings -- in Jacobian coordinates with
 * 64-bit-limb CIOS Montgomery arithmetic.  It shares no code and no design with the CUDA kernels (signed windows, XYZZ,
 * 32-bit limbs, sorted buckets, shifted-base tables), and is itself checked against python big-int double-and-add
 * (tests/test_oracle_curves.py), so it can arbitrate at sizes python cannot reach (2^14 .. 2^16 points).
 *
 * Boundary forms = include/b200zk.h: affine (x, y) little-endian MONTGOMERY limbs (G2 coordinate = c0 || c1), all-zero =
 * infinity; scalars canonical 4 x u64.
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif
typedef uint64_t u64;
typedef unsigned __int128 u128;
#define MAXL 6
typedef struct { int n, deg; u64 p[MAXL], pinv, one[MAXL], r2[MAXL]; } ctx_t;
typedef struct { u64 l[MAXL]; } fp;
typedef struct { fp c[2]; } fe;
typedef struct { fe x, y, z; } jac;      /* z = 0: infinity */

static int geq(const ctx_t* C, const u64* a, const u64* b) { for (int i = C->n - 1; i >= 0; i--) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; } return 1; }
static void subn(const ctx_t* C, u64* r, const u64* a, const u64* b) { u64 bw = 0; for (int i = 0; i < C->n; i++) { u128 d = (u128)a[i] - b[i] - bw; r[i] = (u64)d; bw = (u64)(d >> 64) & 1; } }
static void fp_add(const ctx_t* C, fp* r, const fp* a, const fp* b) {
    u128 c = 0; u64 t[MAXL];
    for (int i = 0; i < C->n; i++) { c += (u128)a->l[i] + b->l[i]; t[i] = (u64)c; c >>= 64; }
    if (c || geq(C, t, C->p)) subn(C, r->l, t, C->p); else memcpy(r->l, t, 8 * C->n);
}
static void fp_sub(const ctx_t* C, fp* r, const fp* a, const fp* b) {
    if (geq(C, a->l, b->l)) subn(C, r->l, a->l, b->l); else { u64 t[MAXL]; subn(C, t, b->l, a->l); subn(C, r->l, C->p, t); }
}
static void fp_mul(const ctx_t* C, fp* r, const fp* a, const fp* b) {      /* CIOS Montgomery */
    const int n = C->n; u64 t[MAXL + 2]; memset(t, 0, sizeof t);
    for (int i = 0; i < n; i++) {
        u128 c = 0;
        for (int j = 0; j < n; j++) { c += (u128)a->l[j] * b->l[i] + t[j]; t[j] = (u64)c; c >>= 64; }
        c += t[n]; t[n] = (u64)c; t[n + 1] = (u64)(c >> 64);
        u64 m = t[0] * C->pinv;
        c = ((u128)m * C->p[0] + t[0]) >> 64;
        for (int j = 1; j < n; j++) { c += (u128)m * C->p[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
        c += t[n]; t[n - 1] = (u64)c; t[n] = t[n + 1] + (u64)(c >> 64);
    }
    if (t[n] || geq(C, t, C->p)) subn(C, r->l, t, C->p); else memcpy(r->l, t, 8 * n);
}
static int fp_is_zero(const ctx_t* C, const fp* a) { u64 o = 0; for (int i = 0; i < C->n; i++) o |= a->l[i]; return o == 0; }
static void fp_pow_pm2(const ctx_t* C, fp* r, const fp* a) {   /* a^(p-2) */
    u64 e[MAXL]; memcpy(e, C->p, 8 * C->n);
    u64 bw = 2; for (int i = 0; i < C->n && bw; i++) { u64 o = e[i]; e[i] -= bw; bw = o < bw ? 1 : 0; }
    fp acc; memcpy(acc.l, C->one, 8 * C->n);
    for (int i = C->n * 64 - 1; i >= 0; i--) { fp t; fp_mul(C, &t, &acc, &acc); acc = t; if ((e[i >> 6] >> (i & 63)) & 1) { fp_mul(C, &t, &acc, a); acc = t; } }
    *r = acc;
}
/* ---- Fp or Fp2 */
static void fe_add(const ctx_t* C, fe* r, const fe* a, const fe* b) { for (int k = 0; k < C->deg; k++) fp_add(C, &r->c[k], &a->c[k], &b->c[k]); }
static void fe_sub(const ctx_t* C, fe* r, const fe* a, const fe* b) { for (int k = 0; k < C->deg; k++) fp_sub(C, &r->c[k], &a->c[k], &b->c[k]); }
static void fe_mul(const ctx_t* C, fe* r, const fe* a, const fe* b) {
    if (C->deg == 1) { fp t; fp_mul(C, &t, &a->c[0], &b->c[0]); r->c[0] = t; return; }
    fp t0, t1, t2, t3, o0, o1;
    fp_mul(C, &t0, &a->c[0], &b->c[0]); fp_mul(C, &t1, &a->c[1], &b->c[1]); fp_mul(C, &t2, &a->c[0], &b->c[1]); fp_mul(C, &t3, &a->c[1], &b->c[0]);
    fp_sub(C, &o0, &t0, &t1); fp_add(C, &o1, &t2, &t3); r->c[0] = o0; r->c[1] = o1;
}
static int fe_is_zero(const ctx_t* C, const fe* a) { for (int k = 0; k < C->deg; k++) if (!fp_is_zero(C, &a->c[k])) return 0; return 1; }
static void fe_inv(const ctx_t* C, fe* r, const fe* a) {
    if (C->deg == 1) { fp_pow_pm2(C, &r->c[0], &a->c[0]); return; }
    fp n0, n1, n, ni, z; fp_mul(C, &n0, &a->c[0], &a->c[0]); fp_mul(C, &n1, &a->c[1], &a->c[1]); fp_add(C, &n, &n0, &n1); fp_pow_pm2(C, &ni, &n);
    memset(&z, 0, sizeof z);
    fp_mul(C, &r->c[0], &a->c[0], &ni); fp t; fp_mul(C, &t, &a->c[1], &ni); fp_sub(C, &r->c[1], &z, &t);
}
/* ---- Jacobian, a = 0 */
static void j_dbl(const ctx_t* C, jac* r, const jac* p) {       /* dbl-2009-l */
    if (fe_is_zero(C, &p->z)) { *r = *p; return; }
    fe A, B, Cc, D, E, F, t, X3, Y3, Z3;
    fe_mul(C, &A, &p->x, &p->x); fe_mul(C, &B, &p->y, &p->y); fe_mul(C, &Cc, &B, &B);
    fe_add(C, &t, &p->x, &B); fe_mul(C, &D, &t, &t); fe_sub(C, &D, &D, &A); fe_sub(C, &D, &D, &Cc); fe_add(C, &D, &D, &D);
    fe_add(C, &E, &A, &A); fe_add(C, &E, &E, &A); fe_mul(C, &F, &E, &E);
    fe_sub(C, &X3, &F, &D); fe_sub(C, &X3, &X3, &D);
    fe_sub(C, &t, &D, &X3); fe_mul(C, &Y3, &E, &t); fe c8; fe_add(C, &c8, &Cc, &Cc); fe_add(C, &c8, &c8, &c8); fe_add(C, &c8, &c8, &c8); fe_sub(C, &Y3, &Y3, &c8);
    fe_mul(C, &Z3, &p->y, &p->z); fe_add(C, &Z3, &Z3, &Z3);
    r->x = X3; r->y = Y3; r->z = Z3;
}
static void j_add(const ctx_t* C, jac* r, const jac* p, const jac* q) {     /* add-2007-bl, with the doubling / infinity cases */
    if (fe_is_zero(C, &p->z)) { *r = *q; return; }
    if (fe_is_zero(C, &q->z)) { *r = *p; return; }
    fe Z1Z1, Z2Z2, U1, U2, S1, S2, H, R, t;
    fe_mul(C, &Z1Z1, &p->z, &p->z); fe_mul(C, &Z2Z2, &q->z, &q->z);
    fe_mul(C, &U1, &p->x, &Z2Z2); fe_mul(C, &U2, &q->x, &Z1Z1);
    fe_mul(C, &t, &q->z, &Z2Z2); fe_mul(C, &S1, &p->y, &t); fe_mul(C, &t, &p->z, &Z1Z1); fe_mul(C, &S2, &q->y, &t);
    fe_sub(C, &H, &U2, &U1); fe_sub(C, &R, &S2, &S1);
    if (fe_is_zero(C, &H)) { if (fe_is_zero(C, &R)) { j_dbl(C, r, p); return; } memset(r, 0, sizeof *r); return; }
    fe HH, HHH, V, X3, Y3, Z3;
    fe_mul(C, &HH, &H, &H); fe_mul(C, &HHH, &HH, &H); fe_mul(C, &V, &U1, &HH);
    fe_mul(C, &X3, &R, &R); fe_sub(C, &X3, &X3, &HHH); fe_sub(C, &X3, &X3, &V); fe_sub(C, &X3, &X3, &V);
    fe_sub(C, &t, &V, &X3); fe_mul(C, &Y3, &R, &t); fe_mul(C, &t, &S1, &HHH); fe_sub(C, &Y3, &Y3, &t);
    fe_mul(C, &t, &p->z, &q->z); fe_mul(C, &Z3, &t, &H);
    r->x = X3; r->y = Y3; r->z = Z3;
}
static void load_fe(const ctx_t* C, fe* r, const u64* w) { memset(r, 0, sizeof *r); for (int k = 0; k < C->deg; k++) memcpy(r->c[k].l, w + k * C->n, 8 * C->n); }
static void store_fe(const ctx_t* C, u64* w, const fe* a) { for (int k = 0; k < C->deg; k++) memcpy(w + k * C->n, a->c[k].l, 8 * C->n); }

static void ctx_init(ctx_t* C, int n, int deg, const u64* p) {
    memset(C, 0, sizeof *C); C->n = n; C->deg = deg; memcpy(C->p, p, 8 * n);
    u64 inv = 1; for (int i = 0; i < 6; i++) inv *= 2 - p[0] * inv;     /* p^-1 mod 2^64 (Newton) */
    C->pinv = (u64)0 - inv;
    /* R mod p and R^2 mod p by doubling 1 modulo p: 64 n and 128 n times */
    fp x; memset(&x, 0, sizeof x); x.l[0] = 1;
    for (int i = 0; i < 128 * n; i++) { fp t; fp_add(C, &t, &x, &x); x = t; if (i == 64 * n - 1) memcpy(C->one, x.l, 8 * n); }
    memcpy(C->r2, x.l, 8 * n);
}

/* out: affine words (all-zero = infinity).  bases: n x (2 * deg * limbs) u64; scalars: n x 4 u64 canonical. */
int cv_msm(int limbs, int deg, const u64* p, const u64* bases, const u64* scalars, size_t n, int scalar_bits, u64* out) {
    ctx_t C; ctx_init(&C, limbs, deg, p);
    const int fw = limbs * deg, pw = 2 * fw;
    int c = n < 32 ? 3 : 0; if (!c) { size_t t = n; while (t >>= 1) c++; c = (c * 69 + 99) / 100; if (c < 3) c = 3; if (c > 16) c = 16; }      /* ~ ln n */
    const int nwin = (scalar_bits + c - 1) / c;
    jac* wsum = (jac*)calloc(nwin, sizeof(jac));
    fe one; memset(&one, 0, sizeof one); memcpy(one.c[0].l, C.one, 8 * limbs);
#pragma omp parallel for schedule(dynamic, 1)
    for (int w = 0; w < nwin; w++) {
        const size_t nbk = ((size_t)1 << c) - 1;
        jac* bk = (jac*)calloc(nbk, sizeof(jac));
        for (size_t i = 0; i < n; i++) {
            const u64* s = scalars + 4 * i;
            int off = w * c, limb = off >> 6, sh = off & 63;
            u64 v = s[limb] >> sh; if (sh + c > 64 && limb < 3) v |= s[limb + 1] << (64 - sh);
            v &= ((u64)1 << c) - 1;
            if (!v) continue;
            const u64* b = bases + (size_t)pw * i;
            int any = 0; for (int k = 0; k < pw; k++) any |= b[k] != 0;
            if (!any) continue;
            jac q; load_fe(&C, &q.x, b); load_fe(&C, &q.y, b + fw); q.z = one;
            jac t; j_add(&C, &t, &bk[v - 1], &q); bk[v - 1] = t;
        }
        jac run, acc; memset(&run, 0, sizeof run); memset(&acc, 0, sizeof acc);
        for (size_t k = nbk; k-- > 0;) { jac t; j_add(&C, &t, &run, &bk[k]); run = t; j_add(&C, &t, &acc, &run); acc = t; }
        wsum[w] = acc;
        free(bk);
    }
    jac tot; memset(&tot, 0, sizeof tot);
    for (int w = nwin - 1; w >= 0; w--) { for (int k = 0; k < c; k++) { jac t; j_dbl(&C, &t, &tot); tot = t; } jac t; j_add(&C, &t, &tot, &wsum[w]); tot = t; }
    free(wsum);
    memset(out, 0, 8 * pw);
    if (fe_is_zero(&C, &tot.z)) return 0;
    fe zi, zi2, zi3, x, y; fe_inv(&C, &zi, &tot.z); fe_mul(&C, &zi2, &zi, &zi); fe_mul(&C, &zi3, &zi2, &zi);
    fe_mul(&C, &x, &tot.x, &zi2); fe_mul(&C, &y, &tot.y, &zi3);
    store_fe(&C, out, &x); store_fe(&C, out + fw, &y);
    return 0;
}
int cv_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
