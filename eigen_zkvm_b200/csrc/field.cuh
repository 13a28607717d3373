// Goldilocks field (p = 2^64 - 2^32 + 1) and its cubic extension GF(p^3) = GL[x]/(x^3 - x - 1) for sm_100a.
//
// Reference semantics: fields/src/field_gl.rs:9-30,385-463 (the reference keeps Montgomery form internally;
// every buffer that crosses our boundary is canonical, field_gl.rs:542-544) and starky/src/f3g.rs:207-235,
// 407-449.  Here values live in registers as canonical u64; products are reduced with the special form
// 2^64 = 2^32 - 1, 2^96 = -1 (mod p) instead of Montgomery -- integer-exact, so results are bit-identical.
#pragma once
#ifdef __CUDACC_RTC__          /* NVRTC (the step-program JIT, jit.cpp) has no standard headers */
typedef unsigned long long uint64_t;
typedef unsigned int uint32_t;
#else
#include <stdint.h>
#include <cuda_runtime.h>
#endif

typedef uint64_t u64;
typedef uint32_t u32;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL /* 2^64 mod p */

#ifndef GL_RED_ALU
#define GL_RED_ALU 0   /* measured on B200: the IMAD.HI form is 5 % faster for Poseidon and the NTT than the all-IADD3 form */
#endif
#include "mont.cuh"     // mp_* carry primitives (PTX add.cc / mad.lo.cc chains; emulated on the host for unit tests)

#ifdef __CUDACC__
#define GL_HD __host__ __device__ __forceinline__
#define GL_D __device__ __forceinline__
#else
#define GL_HD inline
#define GL_D inline
#endif

GL_HD u64 gl_pack(u32 lo, u32 hi) { return (u64)lo | ((u64)hi << 32); }

// ---------------------------------------------------------------------------------------------------------
// Representatives.  "canonical" = in [0, p).  "weak" = any u64 of the right residue.  All arithmetic below is written
// as PTX carry chains (ptxas fuses mad.lo.cc/madc.hi.cc pairs into one IMAD.WIDE with carry; the compiler's own
// lowering of 64-bit compares costs ISETP/SEL pairs and register moves that nearly double the ALU-pipe work, see
// profiles/README.md).  Identities used: 2^64 = 2^32 - 1, 2^96 = -1 (mod p).
//
// gl_sub: a weak, b <= p.  Result == a - b (mod p); canonical whenever a is canonical.            5 instructions
GL_HD u64 gl_sub(u64 a, u64 b) {
    u32 d0 = mp_sub_cc((u32)a, (u32)b), d1 = mp_subc_cc((u32)(a >> 32), (u32)(b >> 32));
    u32 bw = mp_subc(0, 0);                 // 0xffffffff on borrow: true value is d - 2^64 == d - (2^32-1); cannot borrow again for b <= p
    u32 r0 = mp_sub_cc(d0, bw), r1 = mp_subc(d1, 0);
    return gl_pack(r0, r1);
}
// a, b canonical: a + b = a - (p - b), and p - b lies in [1, p]
GL_HD u64 gl_add(u64 a, u64 b) { return gl_sub(a, GL_P - b); }
GL_HD u64 gl_neg(u64 a) { return a ? GL_P - a : 0; }
GL_HD u64 gl_dbl(u64 a) { return gl_add(a, a); }
// (r1:r0) + c * 2^64 == (r1:r0) + c * (2^32 - 1) for a carry bit c; the caller guarantees that this sum does not wrap.  NOTE: never feed the carry of an add chain into subc (or a borrow into addc): ptxas keeps
// the subtract flag inverted, the PTX-documented semantics do not hold across the two families.
GL_HD u64 gl_fold_carry(u32 r0, u32 r1, u32 c) {
    // r - c + (c << 32) with plain adds (IADD3 issues at twice the rate of IMAD and four times that of IMAD.HI on sm_100a,
    // see tools/ubench/int_pipes.cu)
    u32 q0 = mp_sub_cc(r0, c), q1 = mp_subc(r1, 0);
    return gl_pack(q0, q1 + c);
}
// weak sum of a weak and a CANONICAL value (a + b - 2^64 < p, so one fold of the carry suffices).      5 instructions
GL_HD u64 gl_addw(u64 a, u64 b) {
    u32 s0 = mp_add_cc((u32)a, (u32)b), s1 = mp_addc_cc((u32)(a >> 32), (u32)(b >> 32));
    u32 c = mp_addc(0, 0);
    return gl_fold_carry(s0, s1, c);
}

// (hi:lo) = a * b + c for any u64 a, b, c (< 2^128).  The even limb products (a0 b0, a1 b1) and the odd ones
// (a0 b1 + a1 b0) are accumulated separately, each on aligned register pairs, and merged with one carry chain:
// 4 IMAD.WIDE + 4 adds, no register moves.
#ifndef GL_PLAIN_IMAD
#define GL_PLAIN_IMAD 0   /* measured on B200: the plain-IMAD variant below makes the Poseidon kernels 24 % slower; kept as a tested alternative */
#endif
#if GL_PLAIN_IMAD
// Variant: the four partial products are PLAIN IMAD.WIDE (2 issue cycles each on the FMA-heavy pipe; a multiply-add with
// carry-out costs 4, tools/ubench/int_pipes.cu) and every carry is propagated by IADD3 chains, which issue on the other pipes.
GL_HD void gl_mulwide_add(u64 a, u64 b, u64 c, u64& lo, u64& hi) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    const u64 p00 = mp_mul_wide(a0, b0), p11 = mp_mul_wide(a1, b1), p01 = mp_mul_wide(a0, b1), p10 = mp_mul_wide(a1, b0);
    u32 e0 = mp_add_cc((u32)p00, (u32)c), e1 = mp_addc_cc((u32)(p00 >> 32), (u32)(c >> 32));
    u32 e2 = mp_addc_cc((u32)p11, 0), e3 = mp_addc((u32)(p11 >> 32), 0);
    u32 m0 = mp_add_cc((u32)p01, (u32)p10), m1 = mp_addc_cc((u32)(p01 >> 32), (u32)(p10 >> 32)), m2 = mp_addc(0, 0);
    e1 = mp_add_cc(e1, m0); e2 = mp_addc_cc(e2, m1); e3 = mp_addc(e3, m2);
    lo = gl_pack(e0, e1); hi = gl_pack(e2, e3);
}
GL_HD void gl_mulwide(u64 a, u64 b, u64& lo, u64& hi) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    const u64 p00 = mp_mul_wide(a0, b0), p11 = mp_mul_wide(a1, b1), p01 = mp_mul_wide(a0, b1), p10 = mp_mul_wide(a1, b0);
    u32 m0 = mp_add_cc((u32)p01, (u32)p10), m1 = mp_addc_cc((u32)(p01 >> 32), (u32)(p10 >> 32)), m2 = mp_addc(0, 0);
    u32 e1 = mp_add_cc((u32)(p00 >> 32), m0), e2 = mp_addc_cc((u32)p11, m1), e3 = mp_addc((u32)(p11 >> 32), m2);
    lo = gl_pack((u32)p00, e1); hi = gl_pack(e2, e3);
}
#else
GL_HD void gl_mulwide_add(u64 a, u64 b, u64 c, u64& lo, u64& hi) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u32 e0 = mp_mad_lo_cc(a0, b0, (u32)c), e1 = mp_madc_hi_cc(a0, b0, (u32)(c >> 32));
    u32 e2 = mp_madc_lo_cc(a1, b1, 0), e3 = mp_madc_hi(a1, b1, 0);       // a1 b1 + carry < 2^64
    u32 o0 = mp_mul_lo(a0, b1), o1 = mp_mul_hi(a0, b1);
    o0 = mp_mad_lo_cc(a1, b0, o0); o1 = mp_madc_hi_cc(a1, b0, o1);
    u32 o2 = mp_addc(0, 0);
    e1 = mp_add_cc(e1, o0); e2 = mp_addc_cc(e2, o1); e3 = mp_addc(e3, o2);
    lo = gl_pack(e0, e1); hi = gl_pack(e2, e3);
}
GL_HD void gl_mulwide(u64 a, u64 b, u64& lo, u64& hi) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u32 e0 = mp_mul_lo(a0, b0), e1 = mp_mul_hi(a0, b0), e2 = mp_mul_lo(a1, b1), e3 = mp_mul_hi(a1, b1);
    u32 o0 = mp_mul_lo(a0, b1), o1 = mp_mul_hi(a0, b1);
    o0 = mp_mad_lo_cc(a1, b0, o0); o1 = mp_madc_hi_cc(a1, b0, o1);
    u32 o2 = mp_addc(0, 0);
    e1 = mp_add_cc(e1, o0); e2 = mp_addc_cc(e2, o1); e3 = mp_addc(e3, o2);
    lo = gl_pack(e0, e1); hi = gl_pack(e2, e3);
}
#endif
GL_HD void gl_sqrwide(u64 a, u64& lo, u64& hi) { gl_mulwide(a, a, lo, hi); }

// (hi:lo) mod p as a WEAK representative: x = lo + hl*2^64 + hh*2^96 == (lo - hh) + hl*(2^32-1).            9 instructions
GL_HD u64 gl_red128w(u64 lo, u64 hi) {
    const u32 hl = (u32)hi, hh = (u32)(hi >> 32);
    u32 t0 = mp_sub_cc((u32)lo, hh), t1 = mp_subc_cc((u32)(lo >> 32), 0);
    u32 bw = mp_subc(0, 0);
    t0 = mp_sub_cc(t0, bw); t1 = mp_subc(t1, 0);                          // t == lo - hh, weak
#if GL_RED_ALU
    // hl * (2^32 - 1) = (hl << 32) - hl built with two subtractions: keeps IMAD.HI (a quarter-rate instruction on the
    // FMA-heavy pipe, the binding resource of the Poseidon kernels) out of every reduction
    u32 m0 = mp_sub_cc(0, hl), m1 = mp_subc(hl, 0);
    u32 r0 = mp_add_cc(t0, m0), r1 = mp_addc_cc(t1, m1);
#else
    u32 r0 = mp_mad_lo_cc(hl, 0xffffffffu, t0), r1 = mp_madc_hi_cc(hl, 0xffffffffu, t1);
#endif
    u32 c = mp_addc(0, 0);
    return gl_fold_carry(r0, r1, c);                                      // the sum cannot wrap twice
}
// lo + hi32 * 2^64, weak
GL_HD u64 gl_red96w(u64 lo, u32 hi32) {
#if GL_RED_ALU
    u32 m0 = mp_sub_cc(0, hi32), m1 = mp_subc(hi32, 0);
    u32 r0 = mp_add_cc((u32)lo, m0), r1 = mp_addc_cc((u32)(lo >> 32), m1);
#else
    u32 r0 = mp_mad_lo_cc(hi32, 0xffffffffu, (u32)lo), r1 = mp_madc_hi_cc(hi32, 0xffffffffu, (u32)(lo >> 32));
#endif
    u32 c = mp_addc(0, 0);
    return gl_fold_carry(r0, r1, c);
}
GL_HD u64 gl_canon(u64 r) {          // [0, 2^64) -> [0, p):  r >= p  <=>  r + (2^32-1) carries, and then r - p is that sum
    u32 s0 = mp_add_cc((u32)r, 0xffffffffu), s1 = mp_addc_cc((u32)(r >> 32), 0);
    u32 c = mp_addc(0, 0);
    return c ? gl_pack(s0, s1) : r;
}
GL_HD u64 gl_red128(u64 lo, u64 hi) { return gl_canon(gl_red128w(lo, hi)); }
GL_HD u64 gl_red96(u64 lo, u32 hi32) { return gl_canon(gl_red96w(lo, hi32)); }
GL_HD u64 gl_mulw(u64 a, u64 b) { u64 lo, hi; gl_mulwide(a, b, lo, hi); return gl_red128w(lo, hi); }
GL_HD u64 gl_sqrw(u64 a) { return gl_mulw(a, a); }
GL_HD u64 gl_mul(u64 a, u64 b) { return gl_canon(gl_mulw(a, b)); }
GL_HD u64 gl_sqr(u64 a) { return gl_mul(a, a); }
// a * b + c, weak (c any u64)
GL_HD u64 gl_maddw(u64 a, u64 b, u64 c) { u64 lo, hi; gl_mulwide_add(a, b, c, lo, hi); return gl_red128w(lo, hi); }
GL_HD u64 gl_pow(u64 a, u64 e) {
    u64 r = 1;
    while (e) { if (e & 1) r = gl_mul(r, a); a = gl_sqr(a); e >>= 1; }
    return r;
}
// a^(p-2) with a fixed addition chain: p-2 = 2^64 - 2^32 - 1 = (2^32-1)*2^32 + (2^32 - 1)
GL_HD u64 gl_inv(u64 a) {
    // t_k = a^(2^k - 1)
    u64 t2 = gl_mul(gl_sqr(a), a);                         // 2^2-1
    u64 t4 = t2; for (int i = 0; i < 2; i++) t4 = gl_sqr(t4); t4 = gl_mul(t4, t2);      // 2^4-1
    u64 t8 = t4; for (int i = 0; i < 4; i++) t8 = gl_sqr(t8); t8 = gl_mul(t8, t4);      // 2^8-1
    u64 t16 = t8; for (int i = 0; i < 8; i++) t16 = gl_sqr(t16); t16 = gl_mul(t16, t8); // 2^16-1
    u64 t32 = t16; for (int i = 0; i < 16; i++) t32 = gl_sqr(t32); t32 = gl_mul(t32, t16); // 2^32-1
    // exponent p-2 = (2^32-1)*2^32 + (2^32-1) - 0 ... check: (2^32-1)*2^32 + 2^32 - 1 = 2^64 - 1; we need 2^64-2^32-1
    // p-2 = (2^32-2)*2^32 + (2^32-1):  a^(2^32-2) = t32 / a = (t31)^2 where t31 = a^(2^31-1)
    u64 t31 = t16; for (int i = 0; i < 15; i++) t31 = gl_sqr(t31);                      // a^((2^16-1)*2^15)
    // a^(2^15-1) = t8^(2^7) * a^(2^7-1); build 2^7-1 from t4: t4^(2^3) * a^(2^3-1)
    u64 t3 = gl_mul(gl_sqr(t2), a);                        // 2^3-1
    u64 t7 = t4; for (int i = 0; i < 3; i++) t7 = gl_sqr(t7); t7 = gl_mul(t7, t3);      // 2^7-1
    u64 t15 = t8; for (int i = 0; i < 7; i++) t15 = gl_sqr(t15); t15 = gl_mul(t15, t7); // 2^15-1
    t31 = gl_mul(t31, t15);                                // 2^31-1
    u64 hi = gl_sqr(t31);                                  // a^(2^32-2)
    for (int i = 0; i < 32; i++) hi = gl_sqr(hi);          // a^((2^32-2)*2^32)
    return gl_mul(hi, t32);
}

// ---------------------------------------------------------------------------------------------------------
// GF(p^3): (c0, c1, c2) = c0 + c1 x + c2 x^2, x^3 = x + 1   (starky/src/f3g.rs:419-431)
struct f3 { u64 c[3]; };
GL_HD f3 f3_make(u64 a, u64 b, u64 c) { f3 r; r.c[0] = a; r.c[1] = b; r.c[2] = c; return r; }
GL_HD f3 f3_add(f3 a, f3 b) { return f3_make(gl_add(a.c[0], b.c[0]), gl_add(a.c[1], b.c[1]), gl_add(a.c[2], b.c[2])); }
GL_HD f3 f3_sub(f3 a, f3 b) { return f3_make(gl_sub(a.c[0], b.c[0]), gl_sub(a.c[1], b.c[1]), gl_sub(a.c[2], b.c[2])); }
GL_HD f3 f3_muls(f3 a, u64 s) { return f3_make(gl_mul(a.c[0], s), gl_mul(a.c[1], s), gl_mul(a.c[2], s)); }
// Unreduced sums of 64 x 64-bit products (a few thousand terms at most: e4 and o2 count the carries of the 128-bit rows): the even limb products on (e0..e4), the odd ones on (o0..o2), each term four
// multiply-adds on aligned register pairs and three carry adds; one reduction per SUM instead of one per product.
struct gl_acc { u32 e0, e1, e2, e3, e4, o0, o1, o2; };
GL_HD gl_acc gl_acc_mul(u64 a, u64 b) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    gl_acc A;
    A.e0 = mp_mul_lo(a0, b0); A.e1 = mp_mul_hi(a0, b0); A.e2 = mp_mul_lo(a1, b1); A.e3 = mp_mul_hi(a1, b1); A.e4 = 0;
    A.o0 = mp_mul_lo(a0, b1); A.o1 = mp_mul_hi(a0, b1);
    A.o0 = mp_mad_lo_cc(a1, b0, A.o0); A.o1 = mp_madc_hi_cc(a1, b0, A.o1); A.o2 = mp_addc(0, 0);
    return A;
}
GL_HD void gl_acc_mad(gl_acc& A, u64 a, u64 b) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    A.e0 = mp_mad_lo_cc(a0, b0, A.e0); A.e1 = mp_madc_hi_cc(a0, b0, A.e1); A.e2 = mp_madc_lo_cc(a1, b1, A.e2); A.e3 = mp_madc_hi_cc(a1, b1, A.e3); A.e4 = mp_addc(A.e4, 0);
    A.o0 = mp_mad_lo_cc(a0, b1, A.o0); A.o1 = mp_madc_hi_cc(a0, b1, A.o1); A.o2 = mp_addc(A.o2, 0);
    A.o0 = mp_mad_lo_cc(a1, b0, A.o0); A.o1 = mp_madc_hi_cc(a1, b0, A.o1); A.o2 = mp_addc(A.o2, 0);
}
GL_HD void gl_acc_add(gl_acc& A, const gl_acc& B) {
    A.e0 = mp_add_cc(A.e0, B.e0); A.e1 = mp_addc_cc(A.e1, B.e1); A.e2 = mp_addc_cc(A.e2, B.e2); A.e3 = mp_addc_cc(A.e3, B.e3); A.e4 = mp_addc(A.e4, B.e4);
    A.o0 = mp_add_cc(A.o0, B.o0); A.o1 = mp_addc_cc(A.o1, B.o1); A.o2 = mp_addc(A.o2, B.o2);
}
GL_HD u64 gl_acc_red(gl_acc A) {         // canonical; total = E + 2^32 O < 2^20 * 2^128 (e4 << 32 stays below p), and 2^128 = -2^32 (mod p)
    u32 e1 = mp_add_cc(A.e1, A.o0), e2 = mp_addc_cc(A.e2, A.o1), e3 = mp_addc_cc(A.e3, A.o2), e4 = mp_addc(A.e4, 0);
    return gl_canon(gl_sub(gl_red128w(gl_pack(A.e0, e1), gl_pack(e2, e3)), (u64)e4 << 32));
}
#ifndef F3_MUL_LAZY
#define F3_MUL_LAZY 1   /* schoolbook with x^3 = x + 1 folded in, ten multiply-accumulates and THREE reductions (142 instructions) instead of six
                           Karatsuba products with canonical sums and differences around them (235); same canonical values */
#endif
#if F3_MUL_LAZY
GL_HD f3 f3_mul(f3 a, f3 b) {
    // c0 = a0 b0 + (a1 b2 + a2 b1);  c1 = a0 b1 + a1 b0 + (a1 b2 + a2 b1) + a2 b2;  c2 = a0 b2 + a1 b1 + a2 b0 + a2 b2
    gl_acc S = gl_acc_mul(a.c[1], b.c[2]); gl_acc_mad(S, a.c[2], b.c[1]);
    gl_acc A0 = S; gl_acc_mad(A0, a.c[0], b.c[0]);
    gl_acc A1 = S; gl_acc_mad(A1, a.c[0], b.c[1]); gl_acc_mad(A1, a.c[1], b.c[0]); gl_acc_mad(A1, a.c[2], b.c[2]);
    gl_acc A2 = gl_acc_mul(a.c[0], b.c[2]); gl_acc_mad(A2, a.c[1], b.c[1]); gl_acc_mad(A2, a.c[2], b.c[0]); gl_acc_mad(A2, a.c[2], b.c[2]);
    return f3_make(gl_acc_red(A0), gl_acc_red(A1), gl_acc_red(A2));
}
#else
GL_HD f3 f3_mul(f3 a, f3 b) {
    u64 A = gl_mul(gl_add(a.c[0], a.c[1]), gl_add(b.c[0], b.c[1]));
    u64 B = gl_mul(gl_add(a.c[0], a.c[2]), gl_add(b.c[0], b.c[2]));
    u64 C = gl_mul(gl_add(a.c[1], a.c[2]), gl_add(b.c[1], b.c[2]));
    u64 D = gl_mul(a.c[0], b.c[0]), E = gl_mul(a.c[1], b.c[1]), F = gl_mul(a.c[2], b.c[2]);
    u64 G = gl_sub(D, E);
    return f3_make(gl_sub(gl_add(C, G), F), gl_sub(gl_sub(gl_sub(gl_add(A, C), E), E), D), gl_sub(B, G));
}
#endif
GL_HD f3 f3_inv(f3 x) {   // f3g.rs:207-235
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
    return f3_make(gl_mul(i1, ti), gl_mul(i2, ti), gl_mul(i3, ti));
}

// ---------------------------------------------------------------------------------------------------------
// g^e from a two-level table: lo[e & (2^LO_BITS-1)] * hi[e >> LO_BITS]
#define POW_LO_BITS 12
struct PowTab { const u64* lo; const u64* hi; };
#ifdef __CUDACC__
__device__ __forceinline__ u64 powtab_get(PowTab t, u64 e) {
    u64 l = __ldg(t.lo + (e & ((1u << POW_LO_BITS) - 1)));
    u64 h = __ldg(t.hi + (e >> POW_LO_BITS));
    return gl_mul(l, h);
}
// weak variant for consumers that only multiply by the result (the NTT twiddles)
__device__ __forceinline__ u64 powtab_getw(PowTab t, u32 e) {
    u64 l = __ldg(t.lo + (e & ((1u << POW_LO_BITS) - 1)));
    u64 h = __ldg(t.hi + (e >> POW_LO_BITS));
    return gl_mulw(l, h);
}
#endif
