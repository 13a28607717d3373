/*
 * bn254_oracle.c -- CPU ORACLE / CPU baseline for the BN254 G1 multi-scalar multiplication (test infrastructure).
 *
 * The reference's groth16 prover delegates to `bellman_ce::groth16::create_random_proof`
 * (groth16/src/groth16.rs:88-96); its MSM (`bellman_ce::multiexp`, bellman_ce 0.3.2 =
 * matter-labs/bellman@beta 416f79d3, pulled through franklin-crypto 0.0.5@beta 9e3c2a12, Cargo.lock:668-670,
 * 2259-2261) is NOT vendored under /root/reference, and no reference test pins an MSM output (proofs are
 * randomised, groth16/src/api.rs:154,173).  PARITY UNPINNED by the reference; we pin it ourselves:
 *   - curve constants cross-checked against in-repo sources (Fq modulus groth16/src/api.rs:636, Fr modulus
 *     starky/src/field_bn128.rs:12) and on-curve sample points from groth16/test-vectors/*.json;
 *   - this C Pippenger == an independent Python big-int double-and-add (oracle/bn254.py) on small inputs.
 *
 * Algorithm restated (published bellman multiexp): split scalars into c-bit windows (c = ln n, 3 for n < 32),
 * per window accumulate points into 2^c - 1 buckets, fold buckets with a running sum, combine windows by
 * Horner with c doublings.  Windows run in parallel (OpenMP) like bellman's Worker pool.
 *
 * Representation: Fq elements as 4 x u64 little-endian limbs in MONTGOMERY form (R = 2^256), which is what
 * pairing_ce's `Fq` holds in memory; affine points (x, y), infinity encoded as (0, 0); scalars canonical 4 x u64.
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
typedef uint64_t u64;
typedef unsigned __int128 u128;

static const u64 Q[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static u64 QINV;          /* -q^-1 mod 2^64 */
static u64 R1[4], R2[4];  /* R mod q, R^2 mod q */
static u64 B3[4];         /* curve b = 3 in Montgomery form */
static int ready = 0;

static int geq(const u64 *a, const u64 *b) { for (int i = 3; i >= 0; i--) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; } return 1; }
static void sub_n(u64 *r, const u64 *a, const u64 *b) { u128 bw = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)a[i] - b[i] - bw; r[i] = (u64)d; bw = (d >> 64) & 1; } }
static void fq_add(u64 *r, const u64 *a, const u64 *b) {
    u128 c = 0; u64 t[4];
    for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; t[i] = (u64)c; c >>= 64; }
    if (c || geq(t, Q)) sub_n(r, t, Q); else memcpy(r, t, 32);
}
static void fq_sub(u64 *r, const u64 *a, const u64 *b) {
    if (geq(a, b)) sub_n(r, a, b); else { u64 t[4]; sub_n(t, b, a); sub_n(r, Q, t); }
}
static void fq_mul(u64 *r, const u64 *a, const u64 *b) {   /* CIOS Montgomery */
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (u64)c; c >>= 64; }
        c += t[4]; t[4] = (u64)c; t[5] = (u64)(c >> 64);
        u64 m = t[0] * QINV;
        c = ((u128)m * Q[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * Q[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
        c += t[4]; t[3] = (u64)c; t[4] = t[5] + (u64)(c >> 64);
    }
    if (t[4] || geq(t, Q)) sub_n(r, t, Q); else memcpy(r, t, 32);
}
static int fq_is_zero(const u64 *a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static void fq_pow(u64 *r, const u64 *a, const u64 *e) {
    u64 acc[4]; memcpy(acc, R1, 32);
    for (int i = 255; i >= 0; i--) { fq_mul(acc, acc, acc); if ((e[i / 64] >> (i % 64)) & 1) fq_mul(acc, acc, a); }
    memcpy(r, acc, 32);
}
static void fq_inv(u64 *r, const u64 *a) { u64 e[4]; static const u64 two[4] = {2, 0, 0, 0}; sub_n(e, Q, two); fq_pow(r, a, e); }
static void init(void) {
    if (ready) return;
    u64 inv = 1; for (int i = 0; i < 63; i++) { inv *= inv; inv *= Q[0]; } QINV = (u64)0 - inv;      /* q^-1 = q^(2^63-1) mod 2^64 */
    /* R mod q by doubling 1, 256 times */
    u64 x[4] = {1, 0, 0, 0};
    for (int i = 0; i < 256; i++) fq_add(x, x, x);
    memcpy(R1, x, 32);
    for (int i = 0; i < 256; i++) fq_add(x, x, x);
    memcpy(R2, x, 32);
    u64 three[4] = {3, 0, 0, 0}; ready = 1; fq_mul(B3, three, R2);
}
void bn_to_mont(const u64 *a, u64 *r) { init(); fq_mul(r, a, R2); }
void bn_from_mont(const u64 *a, u64 *r) { init(); static const u64 one[4] = {1, 0, 0, 0}; fq_mul(r, a, one); }

/* Jacobian points (X, Y, Z), Z = 0 <=> infinity */
typedef struct { u64 x[4], y[4], z[4]; } jac;
static void jac_zero(jac *p) { memset(p, 0, sizeof *p); memcpy(p->x, R1, 32); memcpy(p->y, R1, 32); }
static void jac_dbl(jac *r, const jac *p) {
    if (fq_is_zero(p->z)) { *r = *p; return; }
    u64 a[4], b[4], c[4], d[4], e[4], f[4], t[4];
    fq_mul(a, p->x, p->x); fq_mul(b, p->y, p->y); fq_mul(c, b, b);
    fq_add(t, p->x, b); fq_mul(t, t, t); fq_sub(t, t, a); fq_sub(t, t, c); fq_add(d, t, t);
    fq_add(e, a, a); fq_add(e, e, a); fq_mul(f, e, e);
    u64 z3[4]; fq_mul(z3, p->y, p->z); fq_add(z3, z3, z3);
    u64 x3[4]; fq_sub(x3, f, d); fq_sub(x3, x3, d);
    u64 y3[4]; fq_sub(t, d, x3); fq_mul(y3, e, t); u64 c8[4]; fq_add(c8, c, c); fq_add(c8, c8, c8); fq_add(c8, c8, c8); fq_sub(y3, y3, c8);
    memcpy(r->x, x3, 32); memcpy(r->y, y3, 32); memcpy(r->z, z3, 32);
}
static void jac_add_affine(jac *r, const jac *p, const u64 *qx, const u64 *qy) {   /* q != infinity */
    if (fq_is_zero(p->z)) { memcpy(r->x, qx, 32); memcpy(r->y, qy, 32); memcpy(r->z, R1, 32); return; }
    u64 z1z1[4], u2[4], s2[4], h[4], hh[4], i[4], j[4], rr[4], v[4], t[4];
    fq_mul(z1z1, p->z, p->z); fq_mul(u2, qx, z1z1); fq_mul(s2, qy, p->z); fq_mul(s2, s2, z1z1);
    if (!memcmp(u2, p->x, 32)) { if (!memcmp(s2, p->y, 32)) { jac_dbl(r, p); return; } jac_zero(r); memset(r->z, 0, 32); return; }
    fq_sub(h, u2, p->x); fq_mul(hh, h, h); fq_add(i, hh, hh); fq_add(i, i, i); fq_mul(j, h, i);
    fq_sub(rr, s2, p->y); fq_add(rr, rr, rr); fq_mul(v, p->x, i);
    u64 x3[4], y3[4], z3[4];
    fq_mul(x3, rr, rr); fq_sub(x3, x3, j); fq_sub(x3, x3, v); fq_sub(x3, x3, v);
    fq_sub(t, v, x3); fq_mul(y3, rr, t); fq_mul(t, p->y, j); fq_add(t, t, t); fq_sub(y3, y3, t);
    fq_add(z3, p->z, h); fq_mul(z3, z3, z3); fq_sub(z3, z3, z1z1); fq_sub(z3, z3, hh);
    memcpy(r->x, x3, 32); memcpy(r->y, y3, 32); memcpy(r->z, z3, 32);
}
static void jac_add(jac *r, const jac *p, const jac *q) {
    if (fq_is_zero(p->z)) { *r = *q; return; }
    if (fq_is_zero(q->z)) { *r = *p; return; }
    u64 z1z1[4], z2z2[4], u1[4], u2[4], s1[4], s2[4], h[4], i[4], j[4], rr[4], v[4], t[4];
    fq_mul(z1z1, p->z, p->z); fq_mul(z2z2, q->z, q->z); fq_mul(u1, p->x, z2z2); fq_mul(u2, q->x, z1z1);
    fq_mul(s1, p->y, q->z); fq_mul(s1, s1, z2z2); fq_mul(s2, q->y, p->z); fq_mul(s2, s2, z1z1);
    if (!memcmp(u1, u2, 32)) { if (!memcmp(s1, s2, 32)) { jac_dbl(r, p); return; } jac_zero(r); memset(r->z, 0, 32); return; }
    fq_sub(h, u2, u1); fq_add(i, h, h); fq_mul(i, i, i); fq_mul(j, h, i); fq_sub(rr, s2, s1); fq_add(rr, rr, rr); fq_mul(v, u1, i);
    u64 x3[4], y3[4], z3[4];
    fq_mul(x3, rr, rr); fq_sub(x3, x3, j); fq_sub(x3, x3, v); fq_sub(x3, x3, v);
    fq_sub(t, v, x3); fq_mul(y3, rr, t); fq_mul(t, s1, j); fq_add(t, t, t); fq_sub(y3, y3, t);
    fq_add(z3, p->z, q->z); fq_mul(z3, z3, z3); fq_sub(z3, z3, z1z1); fq_sub(z3, z3, z2z2); fq_mul(z3, z3, h);
    memcpy(r->x, x3, 32); memcpy(r->y, y3, 32); memcpy(r->z, z3, 32);
}
static void jac_to_affine(const jac *p, u64 *out8) {    /* (x,y) Montgomery; infinity -> (0,0) */
    if (fq_is_zero(p->z)) { memset(out8, 0, 64); return; }
    u64 zi[4], zi2[4], zi3[4];
    fq_inv(zi, p->z); fq_mul(zi2, zi, zi); fq_mul(zi3, zi2, zi);
    fq_mul(out8, p->x, zi2); fq_mul(out8 + 4, p->y, zi3);
}
int bn_is_on_curve(const u64 *pt8) {   /* Montgomery affine */
    init();
    if (fq_is_zero(pt8) && fq_is_zero(pt8 + 4)) return 1;
    u64 l[4], r[4];
    fq_mul(l, pt8 + 4, pt8 + 4); fq_mul(r, pt8, pt8); fq_mul(r, r, pt8); fq_add(r, r, B3);
    return !memcmp(l, r, 32);
}
/* out = a + b (affine Montgomery, (0,0) = infinity) */
void bn_add_affine(const u64 *a8, const u64 *b8, u64 *out8) {
    init();
    jac p; jac_zero(&p); memset(p.z, 0, 32);
    if (!(fq_is_zero(a8) && fq_is_zero(a8 + 4))) jac_add_affine(&p, &p, a8, a8 + 4);
    if (!(fq_is_zero(b8) && fq_is_zero(b8 + 4))) jac_add_affine(&p, &p, b8, b8 + 4);
    jac_to_affine(&p, out8);
}
static unsigned get_bits(const u64 *s, unsigned off, unsigned c) {
    if (off >= 256) return 0;
    unsigned limb = off / 64, sh = off % 64;
    u64 v = s[limb] >> sh;
    if (sh + c > 64 && limb < 3) v |= s[limb + 1] << (64 - sh);
    return (unsigned)(v & ((1ull << c) - 1));
}
/* bases: n x 8 u64 (x, y Montgomery), scalars: n x 4 u64 canonical; out: affine Montgomery 8 u64 */
void bn_msm(const u64 *bases, const u64 *scalars, size_t n, u64 *out8) {
    init();
    unsigned c = n < 32 ? 3 : (unsigned)ceil(log((double)n));
    unsigned nw = (254 + c - 1) / c;
    jac *wsum = (jac *)malloc(nw * sizeof(jac));
#pragma omp parallel for schedule(dynamic)
    for (int w = 0; w < (int)nw; w++) {
        size_t nb = ((size_t)1 << c) - 1;
        jac *bk = (jac *)malloc(nb * sizeof(jac));
        for (size_t b = 0; b < nb; b++) { jac_zero(&bk[b]); memset(bk[b].z, 0, 32); }
        for (size_t i = 0; i < n; i++) {
            unsigned d = get_bits(scalars + 4 * i, (unsigned)w * c, c);
            const u64 *p = bases + 8 * i;
            if (d && !(fq_is_zero(p) && fq_is_zero(p + 4))) jac_add_affine(&bk[d - 1], &bk[d - 1], p, p + 4);
        }
        jac run, acc; jac_zero(&run); memset(run.z, 0, 32); acc = run;
        for (size_t b = nb; b-- > 0;) { jac_add(&run, &run, &bk[b]); jac_add(&acc, &acc, &run); }
        wsum[w] = acc;
        free(bk);
    }
    jac tot; jac_zero(&tot); memset(tot.z, 0, 32);
    for (int w = (int)nw - 1; w >= 0; w--) { for (unsigned k = 0; k < c; k++) jac_dbl(&tot, &tot); jac_add(&tot, &tot, &wsum[w]); }
    free(wsum);
    jac_to_affine(&tot, out8);
}
/* naive double-and-add reference for cross-checking the Pippenger above */
void bn_msm_naive(const u64 *bases, const u64 *scalars, size_t n, u64 *out8) {
    init();
    jac tot; jac_zero(&tot); memset(tot.z, 0, 32);
    for (size_t i = 0; i < n; i++) {
        const u64 *p = bases + 8 * i, *s = scalars + 4 * i;
        if (fq_is_zero(p) && fq_is_zero(p + 4)) continue;
        jac acc; jac_zero(&acc); memset(acc.z, 0, 32);
        for (int b = 255; b >= 0; b--) { jac_dbl(&acc, &acc); if ((s[b / 64] >> (b % 64)) & 1) jac_add_affine(&acc, &acc, p, p + 4); }
        jac_add(&tot, &tot, &acc);
    }
    jac_to_affine(&tot, out8);
}
int bn_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
