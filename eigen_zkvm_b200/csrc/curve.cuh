// Short-Weierstrass curves y^2 = x^3 + b (a = 0) over F = Fp<P> (G1) or Fp2<P> (G2): BN254 and BLS12-381.
// Accumulators are in XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; ZZ = 0 <=> infinity), the cheapest
// form for the mixed additions that dominate a bucket MSM (EFD: madd-2008-s, add-2008-s, dbl-2008-s-1).
// The boundary forms are those of the reference's curve libraries (pairing_ce for BN254, blstrs for BLS12-381):
// affine (x, y) in Montgomery limbs with the all-zero pair as the point at infinity, Jacobian (X, Y, Z) results.
#pragma once
#include "curve_params.h"

template <class F> struct Affine { F x, y; };            // (0, 0) = infinity
template <class F> struct Jacobian { F x, y, z; };       // z = 0 = infinity
template <class F> struct Xyzz;
// out-of-line point operations (one copy per curve) for everything except the bucket-accumulation inner loop
template <class F> MP_NOINLINE void xyzz_add_nl(const Xyzz<F>* a, const Xyzz<F>* b, Xyzz<F>* r);
template <class F> MP_NOINLINE void xyzz_dbl_nl(const Xyzz<F>* a, Xyzz<F>* r);
template <class F> MP_NOINLINE void xyzz_dbl_affine_nl(const F* x, const F* y, Xyzz<F>* r);
template <class F> struct Xyzz {
    F x, y, zz, zzz;
    MP_HD static Xyzz inf() { Xyzz p; p.x = F::zero(); p.y = F::zero(); p.zz = F::zero(); p.zzz = F::zero(); return p; }
    MP_HD bool is_inf() const { return zz.is_zero(); }
    MP_HD Xyzz neg() const { Xyzz p = *this; p.y = y.neg(); return p; }
    MP_HD static Xyzz from_affine(const F& x, const F& y) { Xyzz r; r.x = x; r.y = y; r.zz = F::one(); r.zzz = F::one(); return r; }
    MP_HD static Xyzz dbl_affine(const F& x, const F& y) { Xyzz r; xyzz_dbl_affine_nl<F>(&x, &y, &r); return r; }
    MP_HD Xyzz dbl() const { Xyzz r; xyzz_dbl_nl<F>(this, &r); return r; }
    MP_HD Xyzz add(const Xyzz& o) const { Xyzz r; xyzz_add_nl<F>(this, &o, &r); return r; }
    MP_HD static Xyzz dbl_affine_inl(const F& x, const F& y) {       // mdbl-2008-s-1
        F u = y.dbl(), v = u.sqr(), w = u * v, s = x * v;
        F x2 = x.sqr(), m = x2.dbl() + x2;
        Xyzz r;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * y;
        r.zz = v; r.zzz = w;
        return r;
    }
    MP_HD Xyzz dbl_inl() const {                                      // dbl-2008-s-1
        if (is_inf()) return *this;
        F u = y.dbl(), v = u.sqr(), w = u * v, s = x * v;
        F x2 = x.sqr(), m = x2.dbl() + x2;
        Xyzz r;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * y;
        r.zz = v * zz; r.zzz = w * zzz;
        return r;
    }
    MP_HD Xyzz add_affine(const F& x2, const F& y2) const {           // madd-2008-s; (x2, y2) finite
        if (is_inf()) return from_affine(x2, y2);
        F u2 = x2 * zz, s2 = y2 * zzz;
        F p_ = u2 - x, r_ = s2 - y;
        if (p_.is_zero()) { if (r_.is_zero()) return dbl_affine(x2, y2); return inf(); }
        F pp = p_.sqr(), ppp = p_ * pp, q = x * pp;
        Xyzz r;
        r.x = r_.sqr() - ppp - q.dbl();
        r.y = r_ * (q - r.x) - y * ppp;
        r.zz = zz * pp; r.zzz = zzz * ppp;
        return r;
    }
    MP_HD Xyzz add_inl(const Xyzz& o) const {                          // add-2008-s
        if (is_inf()) return o;
        if (o.is_inf()) return *this;
        F u1 = x * o.zz, u2 = o.x * zz, s1 = y * o.zzz, s2 = o.y * zzz;
        F p_ = u2 - u1, r_ = s2 - s1;
        if (p_.is_zero()) { if (r_.is_zero()) return dbl(); return inf(); }
        F pp = p_.sqr(), ppp = p_ * pp, q = u1 * pp;
        Xyzz r;
        r.x = r_.sqr() - ppp - q.dbl();
        r.y = r_ * (q - r.x) - s1 * ppp;
        r.zz = zz * o.zz * pp; r.zzz = zzz * o.zzz * ppp;
        return r;
    }
    MP_HD Xyzz mul_small(u32 k) const {     // k * p, double-and-add
        Xyzz acc = inf(), p = *this;
        while (k) { if (k & 1) acc = acc.add(p); p = p.dbl(); k >>= 1; }
        return acc;
    }
    // normalised Jacobian triple: (x, y, 1), or (0, 1, 0) for the point at infinity
    MP_HD Jacobian<F> to_jacobian() const {
        Jacobian<F> o;
        if (is_inf()) { o.x = F::zero(); o.y = F::one(); o.z = F::zero(); return o; }
        F zi = zzz.inv();                 // 1 / z^3
        F zinv = zi * zz;                 // 1 / z
        o.x = x * zinv.sqr(); o.y = y * zi; o.z = F::one();
        return o;
    }
    // the same point as an UN-normalised Jacobian triple (X ZZ ZZZ^2, Y ZZ^3 ZZZ^2, ZZ ZZZ): eight products and no field inversion (a
    // single thread's inversion is 380 dependent Montgomery products, 0.17 ms) -- for partial sums that are added up again anyway
    MP_HD Jacobian<F> to_jacobian_raw() const {
        Jacobian<F> o;
        if (is_inf()) { o.x = F::zero(); o.y = F::one(); o.z = F::zero(); return o; }
        F t = zzz.sqr(), zz2 = zz.sqr();
        o.x = x * (zz * t); o.y = y * ((zz2 * zz) * t); o.z = zz * zzz;
        return o;
    }
    MP_HD static Xyzz from_jacobian(const Jacobian<F>& j) {
        if (j.z.is_zero()) return inf();
        Xyzz r; r.x = j.x; r.y = j.y; r.zz = j.z.sqr(); r.zzz = r.zz * j.z; return r;
    }
};
template <class F> MP_NOINLINE void xyzz_add_nl(const Xyzz<F>* a, const Xyzz<F>* b, Xyzz<F>* r) { Xyzz<F> x = *a, y = *b; *r = x.add_inl(y); }
template <class F> MP_NOINLINE void xyzz_dbl_nl(const Xyzz<F>* a, Xyzz<F>* r) { Xyzz<F> x = *a; *r = x.dbl_inl(); }
template <class F> MP_NOINLINE void xyzz_dbl_affine_nl(const F* x, const F* y, Xyzz<F>* r) { F a = *x, b = *y; *r = Xyzz<F>::dbl_affine_inl(a, b); }
