// Poseidon over Goldilocks, t = 12, x^7, R_F = 8, R_P = 22, "optimised" round structure.
// Reference: starky/src/poseidon_opt.rs:80-200 (constants: poseidon_constants_opt.rs).
//
// One permutation per thread, state in 12 x u64 registers, loops fully unrolled so indices are static.
// Round constants / sparse matrices sit in __constant__ memory: every lane of a warp reads the same
// word at the same time, i.e. a constant-bank broadcast, no LSU traffic.
// The MDS matrix M has entries <= 41 (+8 on the diagonal head), so st' = M^T st is accumulated as two
// 64-bit sums over the 32-bit halves of the state and reduced once per lane (no 64x64 products).
#pragma once
#include "field.cuh"

// Filled by merkle.cu (the only translation unit that includes this header); canonical (< p) values.
static __constant__ u64 cPOS_C[118];
static __constant__ u64 cPOS_P[144];
static __constant__ u64 cPOS_S[506];
static __constant__ u32 cPOS_M[144];   // small entries
static __constant__ u64 cPOS_RC[96];   // additive constants of the 8 full rounds (see poseidon12)

// State lanes are kept as WEAK representatives (any u64 of the right residue, see field.cuh) between layers; every
// consumer below accepts them, and the caller canonicalises the lanes it stores.
// x^7 + c: the round constant rides on the last product's 128-bit sum (no separate modular addition)
GL_D u64 pos_pow7_c(u64 x, u64 c) {
    u64 x2 = gl_sqrw(x), x3 = gl_mulw(x2, x), x6 = gl_sqrw(x3);
    return gl_maddw(x6, x, c);
}

// lane value = lo + hi * 2^32 for two sums lo, hi < 2^42 of small-entry products: three 32-bit words, one weak reduction
GL_D u64 pos_mds_combine(u64 lo, u64 hi) {
    u32 v1 = mp_add_cc((u32)(lo >> 32), (u32)hi), v2 = mp_addc((u32)(hi >> 32), 0);
    return gl_red96w(gl_pack((u32)lo, v1), v2);
}
// st'[i] = sum_j M[j][i] * st[j] with 32-bit-small M
GL_D void pos_mds_small(u64* st) {
    u64 lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { lo[i] = 0; hi[i] = 0; }
#pragma unroll
    for (int j = 0; j < 12; j++) {
        const u32 sl = (u32)st[j], sh = (u32)(st[j] >> 32);
#pragma unroll
        for (int i = 0; i < 12; i++) {
            const u32 m = cPOS_M[j * 12 + i];
            lo[i] = mp_mad_wide(m, sl, lo[i]);            // < 12 * 49 * 2^32 < 2^42
            hi[i] = mp_mad_wide(m, sh, hi[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 12; i++) st[i] = pos_mds_combine(lo[i], hi[i]);
}

// sum_j coef[j*stride] * st[j] with full-width entries.  The even limb products (a0 b0 + 2^64 a1 b1) and the odd ones
// (a0 b1 + a1 b0, weight 2^32) of all 12 terms are accumulated in two separate multi-word sums -- every product is
// one IMAD.WIDE with carry on an aligned register pair, 7 instructions per term -- merged and reduced once.
GL_D u64 pos_dot12(const u64* __restrict__ coef, int stride, const u64* st) {
    u32 e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0, o0 = 0, o1 = 0, o2 = 0;
#pragma unroll
    for (int j = 0; j < 12; j++) {
        const u64 a = coef[j * stride], b = st[j];
        const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
#if GL_PLAIN_IMAD
        const u64 p00 = mp_mul_wide(a0, b0), p11 = mp_mul_wide(a1, b1), p01 = mp_mul_wide(a0, b1), p10 = mp_mul_wide(a1, b0);
        e0 = mp_add_cc(e0, (u32)p00); e1 = mp_addc_cc(e1, (u32)(p00 >> 32)); e2 = mp_addc_cc(e2, (u32)p11); e3 = mp_addc_cc(e3, (u32)(p11 >> 32)); e4 = mp_addc(e4, 0);
        o0 = mp_add_cc(o0, (u32)p01); o1 = mp_addc_cc(o1, (u32)(p01 >> 32)); o2 = mp_addc(o2, 0);
        o0 = mp_add_cc(o0, (u32)p10); o1 = mp_addc_cc(o1, (u32)(p10 >> 32)); o2 = mp_addc(o2, 0);
#else
        e0 = mp_mad_lo_cc(a0, b0, e0); e1 = mp_madc_hi_cc(a0, b0, e1); e2 = mp_madc_lo_cc(a1, b1, e2); e3 = mp_madc_hi_cc(a1, b1, e3); e4 = mp_addc(e4, 0);
        o0 = mp_mad_lo_cc(a0, b1, o0); o1 = mp_madc_hi_cc(a0, b1, o1); o2 = mp_addc(o2, 0);
        o0 = mp_mad_lo_cc(a1, b0, o0); o1 = mp_madc_hi_cc(a1, b0, o1); o2 = mp_addc(o2, 0);
#endif
    }
    // total = E + 2^32 O  (< 12 * 2^128: five 32-bit words and a small sixth)
    e1 = mp_add_cc(e1, o0); e2 = mp_addc_cc(e2, o1); e3 = mp_addc_cc(e3, o2); e4 = mp_addc(e4, 0);
    // 2^128 = 2^64 (2^32 - 1) = 2^96 - 2^64 = -1 - (2^32 - 1) = -2^32 (mod p)
    u64 r = gl_red128w(gl_pack(e0, e1), gl_pack(e2, e3));
    return gl_sub(r, (u64)e4 << 32);         // e4 <= 12, so the subtrahend is < p; weak in, weak out
}

#ifndef POS_LOOPED_LAYERS
#define POS_LOOPED_LAYERS 1
#endif
// compact (looped) forms of the two dense layers: one output lane per iteration, results staged through a small
// local array (L1-resident) because registers cannot be indexed dynamically.  Same arithmetic, ~10x less code.
GL_D void pos_mds_small_looped(u64* st) {
    u32 sl[12], sh[12]; u64 t[12];
#pragma unroll
    for (int j = 0; j < 12; j++) { sl[j] = (u32)st[j]; sh[j] = (u32)(st[j] >> 32); }
#pragma unroll 1
    for (int i = 0; i < 12; i++) {
        u64 lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < 12; j++) { const u32 m = cPOS_M[j * 12 + i]; lo = mp_mad_wide(m, sl[j], lo); hi = mp_mad_wide(m, sh[j], hi); }
        t[i] = pos_mds_combine(lo, hi);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) st[i] = t[i];
}
GL_D void pos_dense_looped(const u64* __restrict__ Mx, u64* st) {
    u64 t[12];
#pragma unroll 1
    for (int i = 0; i < 12; i++) t[i] = pos_dot12(Mx + i, 12, st);
#pragma unroll
    for (int i = 0; i < 12; i++) st[i] = t[i];
}

// S-box + round constant; kept out of line: the permutation is ~10^4 instructions when everything is inlined,
// far beyond the 32 KB instruction cache (ncu: stall_no_instruction 2.8 per issue), and a call costs a few cycles.
__device__ __noinline__ u64 pos_sbox_c_call(u64 x, u64 c) { return pos_pow7_c(x, c); }
template <bool CALL> __device__ __forceinline__ u64 pos_sbox_c(u64 x, u64 c) { if (CALL) return pos_sbox_c_call(x, c); return pos_pow7_c(x, c); }

// in/out: st[12] = inp[0..8] || cap[0..4]  ->  full 12-lane output (first 4 = digest)
// Round schedule of poseidon_opt.rs:80-200 folded into ONE loop over the 8 full rounds so that each code block
// (S-box layer, small-MDS layer, dense P layer, partial round) exists once:
//   r = 0..2: sbox, +C[12(r+1)+i], MDS      r = 3: sbox, +C[48+i], P, then the 22 partial rounds
//   r = 4..6: sbox, +C[82+12(r-4)+i], MDS   r = 7: sbox, MDS
// cPOS_RC[r][i] holds those additive constants (zeros for r = 7).
template <bool CALL = true> __device__ __forceinline__ void poseidon12(u64* st) {
#pragma unroll
    for (int i = 0; i < 12; i++) st[i] = gl_addw(st[i], cPOS_C[i]);
#pragma unroll 1
    for (int r = 0; r < 8; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) st[i] = pos_sbox_c<CALL>(st[i], cPOS_RC[r * 12 + i]);
#if POS_LOOPED_LAYERS
        if (r != 3) { pos_mds_small_looped(st); continue; }
        pos_dense_looped(cPOS_P, st);
#else
        if (r != 3) { pos_mds_small(st); continue; }
        {
            u64 t[12];
#pragma unroll
            for (int i = 0; i < 12; i++) t[i] = pos_dot12(cPOS_P + i, 12, st);
#pragma unroll
            for (int i = 0; i < 12; i++) st[i] = t[i];
        }
#endif
#pragma unroll 1
        for (int q = 0; q < 22; q++) {
            const u64* S = cPOS_S + 23 * q;
            u64 x0 = pos_sbox_c<CALL>(st[0], cPOS_C[60 + q]);
            st[0] = x0;
            u64 s0 = pos_dot12(S, 1, st);
#pragma unroll
            for (int k = 1; k < 12; k++) st[k] = gl_maddw(S[11 + k], x0, st[k]);
            st[0] = s0;
        }
    }
}
