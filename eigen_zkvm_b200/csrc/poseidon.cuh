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
static __constant__ u64 cPOS_K0[12];   // (C[i])^7 + RC[0][i]: what the first S-box layer leaves in a lane whose input is 0
static __constant__ u32 cPOS_MT[12 * 12];   // M transposed: cPOS_MT[i * 12 + j] = M[j][i] (one 48-byte row per output lane)

// Blocked partial rounds (POS_BLOCKED): coefficients for 5 blocks of 4 rounds + 1 block of 2, see pos_partial_block.
//   per block of B rounds: for i < B: [S_r[0], S_r[1..11], W_(i,0) .. W_(i,i-1)]  (12 + i words), then for lane j = 1..11: [v_r0[j] .. v_(r0+B-1)[j]]  (B words)
#define POS_BLK_WORDS(B) (12 * (B) + (B) * ((B) - 1) / 2 + 11 * (B))
static __constant__ u64 cPOS_BLK[5 * POS_BLK_WORDS(4) + POS_BLK_WORDS(2)];
#ifndef POS_BLOCKED
#define POS_BLOCKED 0      /* measured on B200 (profiles/ab_poseidon_r2.txt): 7 % fewer instructions but 2.5 % SLOWER -- the unrolled block bodies take the kernel from 36 to 62 KB of code; kept as a tested option */
#endif
#ifndef POS_MDS3
#define POS_MDS3 1      /* small-MDS layers on three 22/21/21-bit limbs with 32-bit IMADs (see pos_mds3) */
#endif
#ifndef POS_SQR3
#define POS_SQR3 1      /* squarings with three partial products */
#endif

// State lanes are kept as WEAK representatives (any u64 of the right residue, see field.cuh) between layers; every
// consumer below accepts them, and the caller canonicalises the lanes it stores.
// x^7 + c: the round constant rides on the last product's 128-bit sum (no separate modular addition)
// a^2 as (hi:lo): a0^2 + 2^33 a0 a1 + 2^64 a1^2 -- three IMAD.WIDE, the doubling and the merge on the add pipes
GL_D u64 pos_sqrw(u64 a) {
#if POS_SQR3
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32);
    const u64 p00 = mp_mul_wide(a0, a0), p01 = mp_mul_wide(a0, a1), p11 = mp_mul_wide(a1, a1);
    u32 d0 = mp_add_cc((u32)p01, (u32)p01), d1 = mp_addc_cc((u32)(p01 >> 32), (u32)(p01 >> 32)), d2 = mp_addc(0, 0);
    u32 e1 = mp_add_cc((u32)(p00 >> 32), d0), e2 = mp_addc_cc((u32)p11, d1), e3 = mp_addc((u32)(p11 >> 32), d2);
    return gl_red128w(gl_pack((u32)p00, e1), gl_pack(e2, e3));
#else
    return gl_sqrw(a);
#endif
}
GL_D u64 pos_pow7_c(u64 x, u64 c) {
    u64 x2 = pos_sqrw(x), x3 = gl_mulw(x2, x), x6 = pos_sqrw(x3);
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

// sum_{j < N} coef[j] * val[j] + addend (any u64), weak.  Same accumulation as pos_dot12; N <= 16 (the top word counts the carries).
template <int N> GL_D u64 pos_dotn(const u64* __restrict__ coef, const u64* val, u64 addend) {
    u32 e0 = (u32)addend, e1 = (u32)(addend >> 32), e2 = 0, e3 = 0, e4 = 0, o0 = 0, o1 = 0, o2 = 0;
#pragma unroll
    for (int j = 0; j < N; j++) {
        const u64 a = coef[j], b = val[j];
        const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
        e0 = mp_mad_lo_cc(a0, b0, e0); e1 = mp_madc_hi_cc(a0, b0, e1); e2 = mp_madc_lo_cc(a1, b1, e2); e3 = mp_madc_hi_cc(a1, b1, e3); e4 = mp_addc(e4, 0);
        o0 = mp_mad_lo_cc(a0, b1, o0); o1 = mp_madc_hi_cc(a0, b1, o1); o2 = mp_addc(o2, 0);
        o0 = mp_mad_lo_cc(a1, b0, o0); o1 = mp_madc_hi_cc(a1, b0, o1); o2 = mp_addc(o2, 0);
    }
    e1 = mp_add_cc(e1, o0); e2 = mp_addc_cc(e2, o1); e3 = mp_addc_cc(e3, o2); e4 = mp_addc(e4, 0);
    u64 r = gl_red128w(gl_pack(e0, e1), gl_pack(e2, e3));
    return gl_sub(r, (u64)e4 << 32);         // 2^128 = -2^32 (mod p); e4 <= N + 1
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

// ---- small-MDS layer on 22/21/21-bit limbs ----------------------------------------------------------------
// M's entries are <= 41, so a lane is sum_j m_j * (l0_j + 2^22 l1_j + 2^43 l2_j) with three 32-bit sums
// (12 * 41 * 2^22 < 2^31): 36 plain 32-bit IMADs per lane instead of 24 IMAD.WIDE (whose 64-bit results take
// twice the FMA-heavy pipe time) plus the carry adds that join the wide products.
GL_D u64 pos_mds3_combine(u32 A0, u32 A1, u32 A2) {
    // A0 + 2^22 A1 + 2^43 A2 as three words, then one weak reduction
    u32 w0 = mp_add_cc(A0, A1 << 22);
    u32 w1 = mp_addc_cc(A1 >> 10, A2 << 11);
    u32 w2 = mp_addc(A2 >> 21, 0);
    return gl_red96w(gl_pack(w0, w1), w2);
}
GL_D u32 pos_mac32(u32 s, u32 m, u32 acc) {
#ifdef __CUDA_ARCH__
    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(s), "r"(m)); return acc;
#else
    return acc + s * m;
#endif
}
// st' = M^T st for the first n_out lanes (the last layer of a hash needs only the digest lanes).
// UNROLL: all 12 lanes as straight-line code with the 13 distinct entries in (uniform) registers -- no constant loads or local
// staging in the layer, 0.6 k instructions of code; measured on B200: 3 % faster for k_merkle_level (S-boxes inlined, one
// permutation per thread), 5 % slower for k_linearhash, whose sponge loop is already larger (profiles/ab_poseidon_r2.txt).
template <bool UNROLL> GL_D void pos_mds3(u64* st, int n_out) {
    u32 l0[12], l1[12], l2[12];
#pragma unroll
    for (int j = 0; j < 12; j++) { l0[j] = (u32)st[j] & 0x3fffffu; l1[j] = (u32)(st[j] >> 22) & 0x1fffffu; l2[j] = (u32)(st[j] >> 43); }
    if (UNROLL) {
    u32 c[13];
#pragma unroll
    for (int k = 0; k < 12; k++) c[k] = cPOS_MT[k * 12];          // circulant: M[j][i] = c[(i - j) mod 12], except M[0][0]
    c[12] = cPOS_MT[1 * 12 + 1];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        if (i < n_out) {
            u32 A0 = 0, A1 = 0, A2 = 0;
#pragma unroll
            for (int j = 0; j < 12; j++) {
                const u32 m = (i == 0 && j == 0) ? c[0] : (i == j ? c[12] : c[(i - j + 12) % 12]);
                A0 = pos_mac32(l0[j], m, A0); A1 = pos_mac32(l1[j], m, A1); A2 = pos_mac32(l2[j], m, A2);
            }
            st[i] = pos_mds3_combine(A0, A1, A2);
        }
    }
    } else {
    u64 t[12];
#pragma unroll 1
    for (int i = 0; i < n_out; i++) {
        const uint4* row = reinterpret_cast<const uint4*>(cPOS_MT + i * 12);
        const uint4 m0 = row[0], m1 = row[1], m2 = row[2];
        const u32 m[12] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w, m2.x, m2.y, m2.z, m2.w};
        u32 A0 = 0, A1 = 0, A2 = 0;
#pragma unroll
        for (int j = 0; j < 12; j++) { A0 = pos_mac32(l0[j], m[j], A0); A1 = pos_mac32(l1[j], m[j], A1); A2 = pos_mac32(l2[j], m[j], A2); }
        t[i] = pos_mds3_combine(A0, A1, A2);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) st[i] = t[i];
    }
}

// S-box + round constant; kept out of line: the permutation is ~10^4 instructions when everything is inlined,
// far beyond the 32 KB instruction cache (ncu: stall_no_instruction 2.8 per issue), and a call costs a few cycles.
__device__ __noinline__ u64 pos_sbox_c_call(u64 x, u64 c) { return pos_pow7_c(x, c); }
template <bool CALL> __device__ __forceinline__ u64 pos_sbox_c(u64 x, u64 c) { if (CALL) return pos_sbox_c_call(x, c); return pos_pow7_c(x, c); }

// B partial rounds at once (poseidon_opt.rs:148-168 restated).  In the reference every round updates all eleven passive lanes,
// st[k] += S_r[11 + k] * x_r, and reduces each of them (11 reductions per round) although the only consumer inside the partial rounds
// is the next round's dot product.  Within a block the passive lanes are left untouched: round i of the block takes
//     s0_i = S_r[0] x_i + sum_{j>=1} S_r[j] u_j + sum_{m<i} W_(i,m) x_m,      W_(i,m) = sum_{j>=1} S_r[j] S_(r0+m)[11 + j]   (host, pos_init)
// and the eleven lanes are brought up to date ONCE per block, u_j += sum_m S_(r0+m)[11 + j] x_m: one reduction per lane per block
// instead of per round, for i extra dot terms in round i.  Exact (all mod p), so the permutation is unchanged.
template <int B, bool CALL> GL_D void pos_partial_block(u64* st, const u64* __restrict__ tab, const u64* __restrict__ rc) {
    u64 x[B];
    u64 cur = st[0];
#pragma unroll
    for (int i = 0; i < B; i++) {
        x[i] = pos_sbox_c<CALL>(cur, rc[i]);
        u64 val[12 + B];
        val[0] = x[i];
#pragma unroll
        for (int j = 1; j < 12; j++) val[j] = st[j];
#pragma unroll
        for (int m = 0; m < i; m++) val[12 + m] = x[m];
        if (i == 0) cur = pos_dotn<12>(tab, val, 0);
        else if (i == 1) cur = pos_dotn<13>(tab, val, 0);
        else if (i == 2) cur = pos_dotn<14>(tab, val, 0);
        else cur = pos_dotn<15>(tab, val, 0);
        tab += 12 + i;
    }
#pragma unroll
    for (int j = 1; j < 12; j++) { st[j] = pos_dotn<B>(tab, x, st[j]); tab += B; }
    st[0] = cur;
}

// in/out: st[12] = inp[0..8] || cap[0..4]  ->  the first NOUT lanes of the output (4 = digest, 12 = transcript)
// Round schedule of poseidon_opt.rs:80-200: the first S-box layer is peeled off (out-of-line S-boxes), then ONE loop over
// the 8 linear layers so that each code block (small-MDS layer, dense P layer + partial rounds, S-box layer) exists once:
//   peeled: +C[i], sbox, +RC[0]      r = 0..2: MDS, sbox, +RC[r+1]      r = 3: P, the 22 partial rounds, sbox, +RC[4]
//   r = 4..6: MDS, sbox, +RC[r+1]    r = 7: MDS (first NOUT lanes only)
// cPOS_RC[r][i] = C[12(r+1)+i] for r < 4, C[82+12(r-4)+i] for r = 4..6, 0 for r = 7: the constant that follows S-box layer r.
// ZMASK: bit i set = the caller guarantees input lane i is zero (zero capacity, zero-padded leaves): that lane leaves
// the first S-box layer as the constant cPOS_K0[i] and costs nothing.
// cap_zero: the same promise for the four capacity lanes, made at run time (first block of a sponge).
template <bool CALL = true, int NOUT = 12, u32 ZMASK = 0, bool MDS_UNROLL = false> __device__ __forceinline__ void poseidon12(u64* st, bool cap_zero = false) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if ((ZMASK >> i) & 1) st[i] = cPOS_K0[i];
        else st[i] = pos_sbox_c<true>(gl_addw(st[i], cPOS_C[i]), cPOS_RC[i]);
    }
    if (((ZMASK >> 8) & 15) == 15 || cap_zero) {
#pragma unroll
        for (int i = 8; i < 12; i++) st[i] = cPOS_K0[i];
    } else {
#pragma unroll
        for (int i = 8; i < 12; i++) st[i] = pos_sbox_c<true>(gl_addw(st[i], cPOS_C[i]), cPOS_RC[i]);
    }
#pragma unroll 1
    for (int r = 0; r < 8; r++) {
        if (r != 3) {
#if POS_MDS3
            pos_mds3<MDS_UNROLL>(st, r == 7 ? NOUT : 12);
#else
            pos_mds_small_looped(st);
#endif
        } else {
            pos_dense_looped(cPOS_P, st);
#if POS_BLOCKED
#pragma unroll 1
            for (int blk = 0; blk < 5; blk++) pos_partial_block<4, CALL>(st, cPOS_BLK + blk * POS_BLK_WORDS(4), cPOS_C + 60 + 4 * blk);
            pos_partial_block<2, CALL>(st, cPOS_BLK + 5 * POS_BLK_WORDS(4), cPOS_C + 80);
#else
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
#endif
        }
        if (r == 7) break;
#pragma unroll
        for (int i = 0; i < 12; i++) st[i] = pos_sbox_c<CALL>(st[i], cPOS_RC[(r + 1) * 12 + i]);
    }
}

// ---- warp-resident permutation (lane i < 12 holds state element i) -----------------------------------------------------------
// One permutation on one thread is ~2 * 10^4 DEPENDENT instructions: 45 us.  The last ten levels of every tree are chains of such
// permutations with fewer nodes than the machine has SMs, i.e. pure latency (3.2 ms of a 2^24-row proof).  With the state spread
// over 12 lanes the S-boxes of a full round run in parallel, a linear layer is 12 shuffled elements per lane, and a partial round is
// one S-box on lane 0 plus one 64 x 64 product per lane and a 4-step shuffle reduction of the 128-bit terms: ~4x lower latency.
// Throughput per instruction is far worse (20 idle lanes, serial partial S-boxes), so only the small levels use it.
GL_D u64 pos_shfl(u64 v, int src) {
    return gl_pack(__shfl_sync(0xffffffffu, (u32)v, src), __shfl_sync(0xffffffffu, (u32)(v >> 32), src));
}
// all 32 lanes must call; lanes >= 12 carry zeros and are ignored.  s: this lane's element (weak in, weak out).
// Every lane needs ITS OWN constants (row li of the matrices): from the constant bank that is a lane-divergent access, which the
// hardware serialises (12 replays per read -- the first version spent 36 us per permutation that way, no faster than one thread).
// They come from a global-memory copy instead (gPOS_WT, filled by pos_init): consecutive lanes read consecutive words, the loads
// hit L1 and are issued a round ahead of their use.  In a partial round the eleven passive products and their shuffle reduction do
// not depend on the round's S-box, so they are written first and every lane evaluates the S-box (only lane 0's result is used):
// one basic block, the two chains overlap, and the critical path per round is S-box -> one product -> one reduction.
#define POS_WT_C 0
#define POS_WT_RC 12
#define POS_WT_P 108
#define POS_WT_S 252
#define POS_WT_MT 758          /* 144 u32 */
#define POS_WT_WORDS 830
static __device__ __align__(16) u64 gPOS_WT[POS_WT_WORDS];
__device__ __noinline__ u64 poseidon12_warp(u64 s) {
    const int lane = threadIdx.x & 31;
    const bool act = lane < 12;
    const int li = act ? lane : 0;
    const u64* __restrict__ wt = gPOS_WT;
    if (!act) s = 0;
    u32 mrow[12];                       // row li of M transposed: used by the seven small-MDS layers
    {
        const uint4* mr = reinterpret_cast<const uint4*>(reinterpret_cast<const u32*>(wt + POS_WT_MT) + li * 12);
#pragma unroll
        for (int q = 0; q < 3; q++) { const uint4 v = __ldg(mr + q); mrow[4 * q] = v.x; mrow[4 * q + 1] = v.y; mrow[4 * q + 2] = v.z; mrow[4 * q + 3] = v.w; }
    }
    s = pos_pow7_c(gl_addw(s, __ldg(wt + POS_WT_C + li)), __ldg(wt + POS_WT_RC + li));
#pragma unroll 1
    for (int r = 0; r < 8; r++) {
        const u64 rc_next = __ldg(wt + POS_WT_RC + (r < 7 ? r + 1 : 7) * 12 + li);
        if (r != 3) {
            // st'[i] = sum_j M[j][i] st[j]: two 64-bit sums over the 32-bit halves, entries <= 41
            u64 lo = 0, hi = 0;
#pragma unroll
            for (int j = 0; j < 12; j++) {
                const u64 v = pos_shfl(s, j);
                lo = mp_mad_wide(mrow[j], (u32)v, lo); hi = mp_mad_wide(mrow[j], (u32)(v >> 32), hi);
            }
            s = pos_mds_combine(lo, hi);
        } else {
            // dense P layer: lane i takes column i
            u64 st[12];
#pragma unroll
            for (int j = 0; j < 12; j++) st[j] = pos_shfl(s, j);
            s = pos_dot12(wt + POS_WT_P + li, 12, st);
            const u64* S = wt + POS_WT_S;
            u64 Sa = __ldg(S + li), Sb = __ldg(S + 11 + li);
#pragma unroll 1
            for (int q = 0; q < 22; q++) {
                const int qn = q < 21 ? q + 1 : 21;
                const u64 Sa_n = __ldg(S + 23 * qn + li), Sb_n = __ldg(S + 23 * qn + 11 + li);
                // (1) passive terms S[lane] * st[lane], lanes 1..11, as four 32-bit words; summed over the lanes with a fifth word of headroom
                u64 plo = 0, phi = 0;
                gl_mulwide((act && lane >= 1) ? Sa : 0, s, plo, phi);
                u32 w0 = (u32)plo, w1 = (u32)(plo >> 32), w2 = (u32)phi, w3 = (u32)(phi >> 32), w4 = 0;
#pragma unroll
                for (int off = 8; off > 0; off >>= 1) {
                    const u32 o0 = __shfl_down_sync(0xffffffffu, w0, off), o1 = __shfl_down_sync(0xffffffffu, w1, off), o2 = __shfl_down_sync(0xffffffffu, w2, off),
                              o3 = __shfl_down_sync(0xffffffffu, w3, off), o4 = __shfl_down_sync(0xffffffffu, w4, off);
                    w0 = mp_add_cc(w0, o0); w1 = mp_addc_cc(w1, o1); w2 = mp_addc_cc(w2, o2); w3 = mp_addc_cc(w3, o3); w4 = mp_addc(w4, o4);
                }
                // (2) the S-box, on every lane (no divergent branch); lane 0's is the round's x0
                const u64 x0 = pos_shfl(pos_pow7_c(s, cPOS_C[60 + q]), 0);
                // (3) lane 0: S[0] x0 + the passive sum (< 12 * 2^128); 2^128 = -2^32 (mod p)
                u64 tlo, thi; gl_mulwide(Sa, x0, tlo, thi);
                w0 = mp_add_cc(w0, (u32)tlo); w1 = mp_addc_cc(w1, (u32)(tlo >> 32)); w2 = mp_addc_cc(w2, (u32)thi); w3 = mp_addc_cc(w3, (u32)(thi >> 32)); w4 = mp_addc(w4, 0);
                const u64 s0 = gl_sub(gl_red128w(gl_pack(w0, w1), gl_pack(w2, w3)), (u64)w4 << 32);
                // (4) passive lanes: st[i] += S[11 + i] x0
                const u64 sn = gl_maddw(Sb, x0, s);
                s = lane == 0 ? s0 : (act ? sn : 0);
                Sa = Sa_n; Sb = Sb_n;
            }
        }
        if (r == 7) break;
        s = pos_pow7_c(s, rc_next);
    }
    return s;
}
