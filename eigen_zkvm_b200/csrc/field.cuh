// Goldilocks field (p = 2^64 - 2^32 + 1) and its cubic extension GF(p^3) = GL[x]/(x^3 - x - 1) for sm_100a.
//
// Reference semantics: fields/src/field_gl.rs:9-30,385-463 (the reference keeps Montgomery form internally;
// every buffer that crosses our boundary is canonical, field_gl.rs:542-544) and starky/src/f3g.rs:207-235,
// 407-449.  Here values live in registers as canonical u64; products are reduced with the special form
// 2^64 = 2^32 - 1, 2^96 = -1 (mod p) instead of Montgomery -- integer-exact, so results are bit-identical.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

typedef uint64_t u64;
typedef uint32_t u32;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL /* 2^64 mod p */

#ifdef __CUDACC__
#define GL_HD __host__ __device__ __forceinline__
#define GL_D __device__ __forceinline__
#else
#define GL_HD inline
#define GL_D inline
#endif

// ---------------------------------------------------------------------------------------------------------
// canonical add / sub / neg, written as PTX carry chains (the compiler otherwise lowers the 64-bit compares
// to ISETP/SEL pairs that nearly double the ALU-pipe work; see profiles/README.md).
//
// gl_sub: a any u64, b <= p.  Result == a - b (mod p); canonical whenever a is canonical.
GL_D u64 gl_sub(u64 a, u64 b) {
    u64 d; u32 bw;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;" : "=l"(d), "=r"(bw) : "l"(a), "l"(b));
    return d - (u64)bw;            // on borrow bw = 2^32-1: true value is d - 2^64 == d - (2^32-1) (mod p); cannot borrow again for b <= p
}
// a, b canonical: a + b = a - (p - b), and p - b lies in [1, p]
GL_D u64 gl_add(u64 a, u64 b) { return gl_sub(a, GL_P - b); }
GL_D u64 gl_neg(u64 a) { return a ? GL_P - a : 0; }
GL_D u64 gl_dbl(u64 a) { return gl_add(a, a); }

// full 64x64 -> 128 product from four 32x32 -> 64 multiply-adds (IMAD.WIDE.U32), no carries needed:
// each partial sum below is < 2^64 by construction.
GL_D void gl_mulwide(u64 a, u64 b, u64& lo, u64& hi) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u64 p00 = (u64)a0 * b0;
    u64 mid = (u64)a0 * b1 + (p00 >> 32);
    u64 mid2 = (u64)a1 * b0 + (u32)mid;
    hi = (u64)a1 * b1 + (mid >> 32) + (mid2 >> 32);
    lo = (mid2 << 32) | (u32)p00;
}
GL_D void gl_sqrwide(u64 a, u64& lo, u64& hi) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32);
    u64 p00 = (u64)a0 * a0, p01 = (u64)a0 * a1;
    u64 mid = p01 + (p00 >> 32);
    u64 mid2 = p01 + (u32)mid;
    hi = (u64)a1 * a1 + (mid >> 32) + (mid2 >> 32);
    lo = (mid2 << 32) | (u32)p00;
}

// (hi:lo) mod p, hi,lo arbitrary 64-bit.  x = lo + hl*2^64 + hh*2^96 == lo + hl*(2^32-1) - hh.  Canonical result.
GL_D u64 gl_red128(u64 lo, u64 hi) {
    u32 hh = (u32)(hi >> 32), hl = (u32)hi;
    u64 t = gl_sub(lo, (u64)hh);                 // any representative of lo - hh
    u64 m = ((u64)hl << 32) - (u64)hl;           // hl * (2^32 - 1) < 2^64
    u64 r, r2; u32 c, c2;
    asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(r), "=r"(c) : "l"(t), "l"(m));
    r += (u64)(0u - c);                          // carry: 2^64 == 2^32-1; the sum cannot wrap twice (see oracle/gl_oracle.c)
    asm("add.cc.u64 %0, %2, 0xffffffff;\n\taddc.u32 %1, 0, 0;" : "=l"(r2), "=r"(c2) : "l"(r));
    return c2 ? r2 : r;                          // r >= p  <=>  r + (2^32-1) carries, and then r - p = r2
}
// "weak" reductions: same as above without the last conditional subtraction; the result is a representative in
// [0, 2^64) of the right residue.  Safe wherever the consumer accepts any u64: both operands of a product, the left
// operand of gl_sub/gl_add, the MDS accumulators.  NOT safe as the right operand of gl_sub/gl_add (needs <= p).
GL_D u64 gl_red128w(u64 lo, u64 hi) {
    u32 hh = (u32)(hi >> 32), hl = (u32)hi;
    u64 t = gl_sub(lo, (u64)hh);
    u64 m = ((u64)hl << 32) - (u64)hl;
    u64 r; u32 c;
    asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(r), "=r"(c) : "l"(t), "l"(m));
    return r + (u64)(0u - c);
}
GL_D u64 gl_red96w(u64 lo, u32 hi32) {
    u64 m = ((u64)hi32 << 32) - (u64)hi32;
    u64 r; u32 c;
    asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(r), "=r"(c) : "l"(lo), "l"(m));
    return r + (u64)(0u - c);
}
GL_D u64 gl_canon(u64 r) {          // [0, 2^64) -> [0, p)
    u64 r2; u32 c2;
    asm("add.cc.u64 %0, %2, 0xffffffff;\n\taddc.u32 %1, 0, 0;" : "=l"(r2), "=r"(c2) : "l"(r));
    return c2 ? r2 : r;
}
GL_D u64 gl_mulw(u64 a, u64 b) { u64 lo, hi; gl_mulwide(a, b, lo, hi); return gl_red128w(lo, hi); }
GL_D u64 gl_sqrw(u64 a) { u64 lo, hi; gl_sqrwide(a, lo, hi); return gl_red128w(lo, hi); }
GL_D u64 gl_mul(u64 a, u64 b) { u64 lo, hi; gl_mulwide(a, b, lo, hi); return gl_red128(lo, hi); }
GL_D u64 gl_sqr(u64 a) { u64 lo, hi; gl_sqrwide(a, lo, hi); return gl_red128(lo, hi); }
// small-constant multiply-accumulate support: value = lo + hi32 * 2^64 with hi32 < 2^32
GL_D u64 gl_red96(u64 lo, u32 hi32) {
    u64 m = ((u64)hi32 << 32) - (u64)hi32;
    u64 r, r2; u32 c, c2;
    asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(r), "=r"(c) : "l"(lo), "l"(m));
    r += (u64)(0u - c);
    asm("add.cc.u64 %0, %2, 0xffffffff;\n\taddc.u32 %1, 0, 0;" : "=l"(r2), "=r"(c2) : "l"(r));
    return c2 ? r2 : r;
}
GL_D u64 gl_pow(u64 a, u64 e) {
    u64 r = 1;
    while (e) { if (e & 1) r = gl_mul(r, a); a = gl_sqr(a); e >>= 1; }
    return r;
}
// a^(p-2) with a fixed addition chain: p-2 = 2^64 - 2^32 - 1 = (2^32-1)*2^32 + (2^32 - 1)
GL_D u64 gl_inv(u64 a) {
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
GL_D f3 f3_make(u64 a, u64 b, u64 c) { f3 r; r.c[0] = a; r.c[1] = b; r.c[2] = c; return r; }
GL_D f3 f3_add(f3 a, f3 b) { return f3_make(gl_add(a.c[0], b.c[0]), gl_add(a.c[1], b.c[1]), gl_add(a.c[2], b.c[2])); }
GL_D f3 f3_sub(f3 a, f3 b) { return f3_make(gl_sub(a.c[0], b.c[0]), gl_sub(a.c[1], b.c[1]), gl_sub(a.c[2], b.c[2])); }
GL_D f3 f3_muls(f3 a, u64 s) { return f3_make(gl_mul(a.c[0], s), gl_mul(a.c[1], s), gl_mul(a.c[2], s)); }
GL_D f3 f3_mul(f3 a, f3 b) {
    u64 A = gl_mul(gl_add(a.c[0], a.c[1]), gl_add(b.c[0], b.c[1]));
    u64 B = gl_mul(gl_add(a.c[0], a.c[2]), gl_add(b.c[0], b.c[2]));
    u64 C = gl_mul(gl_add(a.c[1], a.c[2]), gl_add(b.c[1], b.c[2]));
    u64 D = gl_mul(a.c[0], b.c[0]), E = gl_mul(a.c[1], b.c[1]), F = gl_mul(a.c[2], b.c[2]);
    u64 G = gl_sub(D, E);
    return f3_make(gl_sub(gl_add(C, G), F), gl_sub(gl_sub(gl_sub(gl_add(A, C), E), E), D), gl_sub(B, G));
}
GL_D f3 f3_inv(f3 x) {   // f3g.rs:207-235
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
GL_D u64 powtab_get(PowTab t, u64 e) {
    u64 l = __ldg(t.lo + (e & ((1u << POW_LO_BITS) - 1)));
    u64 h = __ldg(t.hi + (e >> POW_LO_BITS));
    return gl_mul(l, h);
}
