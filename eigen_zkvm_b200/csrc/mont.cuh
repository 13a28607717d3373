// Montgomery prime fields of 8 or 12 x 32-bit limbs (BN254 Fq/Fr, BLS12-381 Fq/Fr) and their quadratic extension
// Fp2 = Fp[u]/(u^2 + 1), for sm_100a.  Used by the MSM kernels (msm.cu: curve coordinates) and by the BN128 /
// BLS12-381 Poseidon Merkle back-ends (merkle_big.cu: scalar fields).
//
// Reference semantics: the values are the in-memory forms of the reference's field types -- `Fq`/`Fr` of
// pairing_ce / ff_ce (bellman_ce) for BN254 and blstrs for BLS12-381: little-endian limbs in Montgomery form with
// R = 2^(32 N); starky/src/field_bn128.rs:12 and field_bls12381.rs:12 give the two scalar moduli.
//
// Multiplication: word-serial Montgomery with the products of the even and the odd limbs of `a` accumulated in two
// separate limb arrays, so that every 32x32->64 product lands on an aligned (lo, hi) register pair and one carry
// chain (mad.lo.cc / madc.hi.cc, fused by ptxas into IMAD.WIDE with carry) runs through a whole row.  The arrays
// swap roles after each division by 2^32; they are merged once at the end.  Requires N even and 2p(1 + 2^-31) < 2^(32N)
// (checked by tools/gen_curve_params.py for all four fields), so no intermediate overflows N limbs.  Never mix the two
// carry families: the flag left by an add chain must not feed subc and vice versa (ptxas keeps the subtract flag inverted).
//
// Everything here also compiles for the host (the carry flag is emulated), which is how the arithmetic is unit
// tested on the CPU-only build box before it runs on a GPU (tests/test_mont_host.py).
#pragma once
#ifdef __CUDACC_RTC__          /* NVRTC (the step-program JIT, jit.cpp) has no standard headers */
typedef unsigned long long uint64_t;
typedef unsigned int uint32_t;
#else
#include <stdint.h>
#endif

#ifndef B200_U32_TYPES
#define B200_U32_TYPES
typedef uint64_t u64;
typedef uint32_t u32;
#endif

#ifdef __CUDACC__
#define MP_HD __host__ __device__ __forceinline__
#define MP_NOINLINE __host__ __device__ __noinline__
#ifdef FP2_INLINE
#define MP_FP2 __host__ __device__ __forceinline__
#else
#define MP_FP2 __host__ __device__ __noinline__
#endif
#else
#define MP_FP2
#define MP_HD inline
#define MP_NOINLINE
#endif

// ---------------------------------------------------------------------------------------------- carry primitives
#ifndef __CUDA_ARCH__
static thread_local u32 mp_cf_;   // host emulation of PTX CC.CF
#endif
MP_HD u32 mp_add_cc(u32 a, u32 b) {
#ifdef __CUDA_ARCH__
    u32 r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    u64 s = (u64)a + b; mp_cf_ = (u32)(s >> 32); return (u32)s;
#endif
}
MP_HD u32 mp_addc_cc(u32 a, u32 b) {
#ifdef __CUDA_ARCH__
    u32 r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    u64 s = (u64)a + b + mp_cf_; mp_cf_ = (u32)(s >> 32); return (u32)s;
#endif
}
MP_HD u32 mp_addc(u32 a, u32 b) {
#ifdef __CUDA_ARCH__
    u32 r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return a + b + mp_cf_;
#endif
}
MP_HD u32 mp_sub_cc(u32 a, u32 b) {
#ifdef __CUDA_ARCH__
    u32 r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    u64 d = (u64)a - b; mp_cf_ = (u32)(d >> 63); return (u32)d;
#endif
}
MP_HD u32 mp_subc_cc(u32 a, u32 b) {
#ifdef __CUDA_ARCH__
    u32 r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    u64 d = (u64)a - b - mp_cf_; mp_cf_ = (u32)(d >> 63); return (u32)d;
#endif
}
MP_HD u32 mp_subc(u32 a, u32 b) {
#ifdef __CUDA_ARCH__
    u32 r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return a - b - mp_cf_;
#endif
}
MP_HD u32 mp_mul_lo(u32 a, u32 b) { return a * b; }
MP_HD u32 mp_mul_hi(u32 a, u32 b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (u32)(((u64)a * b) >> 32);
#endif
}
MP_HD u32 mp_mad_lo_cc(u32 a, u32 b, u32 c) {
#ifdef __CUDA_ARCH__
    u32 r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    u64 s = (u64)(u32)(a * b) + c; mp_cf_ = (u32)(s >> 32); return (u32)s;
#endif
}
MP_HD u32 mp_madc_lo_cc(u32 a, u32 b, u32 c) {
#ifdef __CUDA_ARCH__
    u32 r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    u64 s = (u64)(u32)(a * b) + c + mp_cf_; mp_cf_ = (u32)(s >> 32); return (u32)s;
#endif
}
MP_HD u32 mp_madc_hi_cc(u32 a, u32 b, u32 c) {
#ifdef __CUDA_ARCH__
    u32 r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    u64 s = (((u64)a * b) >> 32) + c + mp_cf_; mp_cf_ = (u32)(s >> 32); return (u32)s;
#endif
}
MP_HD u32 mp_madc_hi(u32 a, u32 b, u32 c) {
#ifdef __CUDA_ARCH__
    u32 r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    return (u32)((((u64)a * b) >> 32) + c + mp_cf_);
#endif
}

// a * b as one plain IMAD.WIDE.U32 (no carry in or out: half the FMA-heavy pipe time of the mad.lo.cc/madc.hi.cc pair)
MP_HD u64 mp_mul_wide(u32 a, u32 b) {
#ifdef __CUDA_ARCH__
    u64 r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r;
#else
    return (u64)a * b;
#endif
}
// a * b + c with a 64-bit accumulator and no carry out (one IMAD.WIDE.U32); the caller guarantees it cannot overflow
MP_HD u64 mp_mad_wide(u32 a, u32 b, u64 c) {
#ifdef __CUDA_ARCH__
    u64 r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r;
#else
    return (u64)a * b + c;
#endif
}

// ---------------------------------------------------------------------------------------------- Fp<P>
// P supplies: N (limbs, even), M0 = -p^-1 mod 2^32, and constexpr limb accessors mod(i), one(i) (= R mod p),
// r2(i) (= R^2 mod p); see curve_params.h (generated by tools/gen_curve_params.py).
template <class P> struct Fp {
    static constexpr int N = P::N;
    u32 l[P::N];

    MP_HD static Fp zero() { Fp r; for (int i = 0; i < N; i++) r.l[i] = 0; return r; }
    MP_HD static Fp one() { Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::one(i); return r; }
    MP_HD static Fp modulus() { Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::mod(i); return r; }
    MP_HD bool is_zero() const { u32 o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= l[i]; return o == 0; }
    MP_HD bool operator==(const Fp& b) const { u32 o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= l[i] ^ b.l[i]; return o == 0; }

    // a < 2p: r = a - p if a >= p
    MP_HD static Fp cond_sub(const Fp& a) {
        Fp t;
        t.l[0] = mp_sub_cc(a.l[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < N; i++) t.l[i] = mp_subc_cc(a.l[i], P::mod(i));
        u32 bw = mp_subc(0, 0);                 // all-ones on borrow
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = bw ? a.l[i] : t.l[i];
        return r;
    }
    MP_HD friend Fp operator+(const Fp& a, const Fp& b) {       // a, b < p < 2^(32N-2): no carry out
        Fp s;
        s.l[0] = mp_add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) s.l[i] = mp_addc_cc(a.l[i], b.l[i]);
        s.l[N - 1] = mp_addc(a.l[N - 1], b.l[N - 1]);
        return cond_sub(s);
    }
    MP_HD friend Fp operator-(const Fp& a, const Fp& b) {
        Fp d;
        d.l[0] = mp_sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) d.l[i] = mp_subc_cc(a.l[i], b.l[i]);
        u32 bw = mp_subc(0, 0);
        Fp r;
        r.l[0] = mp_add_cc(d.l[0], P::mod(0) & bw);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = mp_addc_cc(d.l[i], P::mod(i) & bw);
        r.l[N - 1] = mp_addc(d.l[N - 1], P::mod(N - 1) & bw);
        return r;
    }
    MP_HD Fp neg() const { return is_zero() ? *this : zero() - *this; }
    MP_HD Fp dbl() const { return *this + *this; }

    // acc[j+1 : j] = a[j] * bi for j = 0, 2, .. (products of every other limb of a, starting at a[0])
    MP_HD static void mul_n(u32* acc, const u32* a, u32 bi) {
#pragma unroll
        for (int j = 0; j < N; j += 2) { acc[j] = mp_mul_lo(a[j], bi); acc[j + 1] = mp_mul_hi(a[j], bi); }
    }
    // acc += sum_{j even} a[j] * bi * 2^(32 j); leaves the carry out of acc[N-1] in CC.CF
    MP_HD static void cmad_n(u32* acc, const u32* a, u32 bi) {
        acc[0] = mp_mad_lo_cc(a[0], bi, acc[0]);
        acc[1] = mp_madc_hi_cc(a[0], bi, acc[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) { acc[j] = mp_madc_lo_cc(a[j], bi, acc[j]); acc[j + 1] = mp_madc_hi_cc(a[j], bi, acc[j + 1]); }
    }
    // the same with the modulus as the multiplicand (limbs off, off+2, ..)
    template <int OFF> MP_HD static void cmad_mod(u32* acc, u32 mi) {
        acc[0] = mp_mad_lo_cc(P::mod(OFF), mi, acc[0]);
        acc[1] = mp_madc_hi_cc(P::mod(OFF), mi, acc[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) { acc[j] = mp_madc_lo_cc(P::mod(OFF + j), mi, acc[j]); acc[j + 1] = mp_madc_hi_cc(P::mod(OFF + j), mi, acc[j + 1]); }
    }
    // acc = (acc >> 64) + sum_{j even} a[j] * bi * 2^(32 j), continuing the carry chain that is live in CC.CF
    MP_HD static void madc_n_rshift(u32* acc, const u32* a, u32 bi) {
#pragma unroll
        for (int j = 0; j < N - 2; j += 2) { acc[j] = mp_madc_lo_cc(a[j], bi, acc[j + 2]); acc[j + 1] = mp_madc_hi_cc(a[j], bi, acc[j + 3]); }
        acc[N - 2] = mp_madc_lo_cc(a[N - 2], bi, 0);
        acc[N - 1] = mp_madc_hi(a[N - 2], bi, 0);
    }
    // One word of b: T <- (T + a * bi + m * p) / 2^32.  On entry T = ev + (od >> 32) with od[0] == 0 (T = 0 when
    // `first`); on exit T = (ev >> 32) + od with ev[0] == 0, so the caller swaps the two arrays for the next word.
    // The limb ev[1] of weight 1 is folded in by the leading add.cc, whose carry (weight 2^32) continues into the
    // chain that rebuilds od = (od >> 64) + a_odd * bi.
    MP_HD static void mad_redc(u32* ev, u32* od, const u32* a, u32 bi, bool first) {
        if (first) {
            mul_n(od, a + 1, bi);
            mul_n(ev, a, bi);
        } else {
            ev[0] = mp_add_cc(ev[0], od[1]);
            madc_n_rshift(od, a + 1, bi);
            cmad_n(ev, a, bi);
            od[N - 1] = mp_addc(od[N - 1], 0);
        }
        u32 mi = ev[0] * P::M0;
        cmad_mod<1>(od, mi);
        cmad_mod<0>(ev, mi);
        od[N - 1] = mp_addc(od[N - 1], 0);
    }
    MP_HD friend Fp operator*(const Fp& a, const Fp& b) {
        u32 ev[N], od[N];
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            mad_redc(ev, od, a.l, b.l[i], i == 0);
            mad_redc(od, ev, a.l, b.l[i + 1], false);
        }
        // result = (od >> 32) + ev
        Fp r;
        r.l[0] = mp_add_cc(ev[0], od[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = mp_addc_cc(ev[i], od[i + 1]);
        r.l[N - 1] = mp_addc(ev[N - 1], 0);
        return cond_sub(r);
    }
    MP_HD Fp sqr() const { return *this * *this; }
    // e: little-endian limbs, n_limbs of them
    MP_HD Fp pow(const u32* e, int n_limbs) const {
        Fp acc = one();
        for (int i = n_limbs * 32 - 1; i >= 0; i--) { acc = acc.sqr(); if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * *this; }
        return acc;
    }
    MP_HD Fp inv() const {      // a^(p-2); 0 -> 0
        u32 e[N];
        for (int i = 0; i < N; i++) e[i] = P::pm2(i);
        return pow(e, N);
    }
    MP_HD Fp to_mont() const { Fp r2;
#pragma unroll
        for (int i = 0; i < N; i++) r2.l[i] = P::r2(i); return *this * r2; }
    MP_HD Fp from_mont() const { Fp o = zero(); o.l[0] = 1; return *this * o; }
};

// ---------------------------------------------------------------------------------------------- FpWide<P>: lazy dot products
// sum_k a_k * b_k (+ c) with ONE Montgomery reduction: the 2N-limb products are accumulated unreduced, so a term costs
// the N^2 multiplies of a_k * b_k alone instead of 2 N^2 (product + reduction).  Poseidon's linear layers over the 254 / 255-bit
// scalar fields are dot products of t <= 17 terms (merkle_big.cu).  Same alignment trick as Fp::operator*: products whose
// limb index sum i + j is even land on the aligned pairs (e[k], e[k+1]) (weight 2^(32 k), k even), the odd ones on
// (o[m], o[m+1]) (weight 2^(32 (m + 1)), m even), each row one carry chain; the carry OUT of a row cannot be added into the
// next limb (it has no headroom in a sum of many products), so it is counted in cy[q] (weight 2^(32 (N + q))).
// Bounds: `terms` products and an addend c < p give T < terms p^2 + p R, so the reduced value is < p (terms p / R + 2); the
// caller passes LOGK with that < 2^LOGK p and the result is brought below p by LOGK conditional subtractions of 2^s p.
template <class P> struct FpWide {
    static constexpr int N = P::N;
    u32 e[2 * N], o[2 * N], cy[N + 1];

    MP_HD void clear() {
#pragma unroll
        for (int i = 0; i < 2 * N; i++) { e[i] = 0; o[i] = 0; }
#pragma unroll
        for (int i = 0; i <= N; i++) cy[i] = 0;
    }
    // T = c * R: after the reduction the result is c + (the dot product)
    MP_HD void set_addend(const Fp<P>& c) {
        clear();
#pragma unroll
        for (int i = 0; i < N; i++) e[N + i] = c.l[i];
    }
    // one row chain: acc[0..N) += sum_{j in 0, 2, ..} x[j] * w * 2^(32 j); `cin`: continue the carry that is live in CC.CF
    template <bool CIN> MP_HD static void row(u32* acc, const u32* x, u32 w) {
        acc[0] = CIN ? mp_madc_lo_cc(x[0], w, acc[0]) : mp_mad_lo_cc(x[0], w, acc[0]);
        acc[1] = mp_madc_hi_cc(x[0], w, acc[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) { acc[j] = mp_madc_lo_cc(x[j], w, acc[j]); acc[j + 1] = mp_madc_hi_cc(x[j], w, acc[j + 1]); }
    }
    template <int I> MP_HD void mad_row(const u32* a, u32 bi) {
        if constexpr (I % 2 == 0) {
            row<false>(e + I, a, bi);         cy[I] = mp_addc(cy[I], 0);             // a_even * b_I: weights I .. I+N-1
            row<false>(o + I, a + 1, bi);     cy[I + 1] = mp_addc(cy[I + 1], 0);     // a_odd * b_I:  weights I+1 .. I+N
        } else {
            row<false>(e + I + 1, a + 1, bi); cy[I + 1] = mp_addc(cy[I + 1], 0);     // a_odd * b_I:  weights I+1 .. I+N
            row<false>(o + I - 1, a, bi);     cy[I] = mp_addc(cy[I], 0);             // a_even * b_I: weights I .. I+N-1
        }
    }
    template <int I> MP_HD void mad_rows(const u32* a, const u32* b) { if constexpr (I < N) { mad_row<I>(a, b[I]); mad_rows<I + 1>(a, b); } }
    // T += a * b
    MP_HD void mad(const Fp<P>& a, const Fp<P>& b) { mad_rows<0>(a.l, b.l); }

    struct ModLimbs { u32 l[P::N]; };
    template <int I> MP_HD void redc_row(const ModLimbs& p) {
        if constexpr (I % 2 == 0) {
            if constexpr (I > 0) e[I] = mp_add_cc(e[I], o[I - 1]);                             // the odd array's word of weight I; carry -> weight I+1
            const u32 m = e[I] * P::M0;
            if constexpr (I > 0) row<true>(o + I, p.l + 1, m); else row<false>(o + I, p.l + 1, m);
            cy[I + 1] = mp_addc(cy[I + 1], 0);
            row<false>(e + I, p.l, m);        cy[I] = mp_addc(cy[I], 0);             // e[I] becomes 0
        } else {
            o[I - 1] = mp_add_cc(o[I - 1], e[I]);
            const u32 m = o[I - 1] * P::M0;
            row<true>(e + I + 1, p.l + 1, m); cy[I + 1] = mp_addc(cy[I + 1], 0);
            row<false>(o + I - 1, p.l, m);    cy[I] = mp_addc(cy[I], 0);             // o[I-1] becomes 0
        }
    }
    template <int I> MP_HD void redc_rows(const ModLimbs& p) { if constexpr (I < N) { redc_row<I>(p); redc_rows<I + 1>(p); } }
    // limb i of 2^s p (N + 1 limbs)
    MP_HD static constexpr u32 shifted_mod(int s, int i) {
        return (i < N ? (P::mod(i) << s) : 0u) | ((s > 0 && i > 0) ? (P::mod(i - 1) >> (32 - s)) : 0u);
    }
    template <int S> MP_HD static void cond_sub_shifted(u32* r) {
        u32 t[N + 1];
        t[0] = mp_sub_cc(r[0], shifted_mod(S, 0));
#pragma unroll
        for (int i = 1; i <= N; i++) t[i] = mp_subc_cc(r[i], shifted_mod(S, i));
        const u32 bw = mp_subc(0, 0);
#pragma unroll
        for (int i = 0; i <= N; i++) r[i] = bw ? r[i] : t[i];
    }
    template <int S> MP_HD static void cond_subs(u32* r) { if constexpr (S >= 0) { cond_sub_shifted<S>(r); cond_subs<S - 1>(r); } }
    // T / R mod p; requires T / R + p < 2^LOGK p
    template <int LOGK> MP_HD Fp<P> reduce() {
        ModLimbs p;
#pragma unroll
        for (int i = 0; i < N; i++) p.l[i] = P::mod(i);
        redc_rows<0>(p);
        u32 r[N + 1];
        r[0] = mp_add_cc(e[N], o[N - 1]);
#pragma unroll
        for (int q = 1; q <= N - 2; q++) r[q] = mp_addc_cc(e[N + q], o[N + q - 1]);
        r[N - 1] = mp_addc_cc(e[2 * N - 1], 0);
        r[N] = mp_addc(0, 0);
        r[0] = mp_add_cc(r[0], cy[0]);
#pragma unroll
        for (int q = 1; q < N; q++) r[q] = mp_addc_cc(r[q], cy[q]);
        r[N] = mp_addc(r[N], cy[N]);
        cond_subs<LOGK - 1>(r);
        Fp<P> out;
#pragma unroll
        for (int i = 0; i < N; i++) out.l[i] = r[i];
        return out;
    }
};

// ---------------------------------------------------------------------------------------------- Fp2 = Fp[u]/(u^2+1)
// Fp2 products are kept out of line (one copy per field): a G2 point operation has 10-14 of them, and inlining
// 3 Montgomery products at every site makes the kernels too large to compile in reasonable time.
template <class P> struct Fp2;
template <class P> MP_FP2 void fp2_mul_nl(const Fp2<P>* a, const Fp2<P>* b, Fp2<P>* r);
template <class P> MP_FP2 void fp2_sqr_nl(const Fp2<P>* a, Fp2<P>* r);
template <class P> struct Fp2 {
    typedef Fp<P> F;
    static constexpr int N = 2 * P::N;
    F c0, c1;
    MP_HD static Fp2 zero() { Fp2 r; r.c0 = F::zero(); r.c1 = F::zero(); return r; }
    MP_HD static Fp2 one() { Fp2 r; r.c0 = F::one(); r.c1 = F::zero(); return r; }
    MP_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    MP_HD bool operator==(const Fp2& b) const { return c0 == b.c0 && c1 == b.c1; }
    MP_HD friend Fp2 operator+(const Fp2& a, const Fp2& b) { Fp2 r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; return r; }
    MP_HD friend Fp2 operator-(const Fp2& a, const Fp2& b) { Fp2 r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; return r; }
    MP_HD Fp2 neg() const { Fp2 r; r.c0 = c0.neg(); r.c1 = c1.neg(); return r; }
    MP_HD Fp2 dbl() const { Fp2 r; r.c0 = c0.dbl(); r.c1 = c1.dbl(); return r; }
    MP_HD friend Fp2 operator*(const Fp2& a, const Fp2& b) { Fp2 r; fp2_mul_nl<P>(&a, &b, &r); return r; }
    MP_HD Fp2 sqr() const { Fp2 r; fp2_sqr_nl<P>(this, &r); return r; }
    MP_HD Fp2 inv() const {                                       // conj / (c0^2 + c1^2)
        F n = (c0.sqr() + c1.sqr()).inv();
        Fp2 r; r.c0 = c0 * n; r.c1 = (c1 * n).neg(); return r;
    }
};
template <class P> MP_FP2 void fp2_mul_nl(const Fp2<P>* a, const Fp2<P>* b, Fp2<P>* r) {     // Karatsuba, 3 base products
    Fp<P> a0 = a->c0, a1 = a->c1, b0 = b->c0, b1 = b->c1;
    Fp<P> aa = a0 * b0, bb = a1 * b1, s = (a0 + a1) * (b0 + b1);
    r->c0 = aa - bb; r->c1 = s - aa - bb;
}
template <class P> MP_FP2 void fp2_sqr_nl(const Fp2<P>* a, Fp2<P>* r) {                        // (c0+c1)(c0-c1), 2 c0 c1
    Fp<P> a0 = a->c0, a1 = a->c1;
    Fp<P> p = a0 * a1;
    r->c0 = (a0 + a1) * (a0 - a1); r->c1 = p.dbl();
}
