// Host build of csrc/mont.cuh (the PTX carry flag is emulated) so that the Montgomery arithmetic used by the MSM and
// the BN128 / BLS12-381 Poseidon kernels is checked against python big ints on the CPU-only box.  Test harness only.
#include "curve_params.h"
#include <cstring>
template <class P> static void ops(int op, const u32* a, const u32* b, u32* o) {
    Fp<P> x, y, r; memcpy(x.l, a, 4 * P::N); memcpy(y.l, b, 4 * P::N);
    switch (op) { case 0: r = x * y; break; case 1: r = x + y; break; case 2: r = x - y; break; case 3: r = x.inv(); break;
                  case 4: r = x.to_mont(); break; case 5: r = x.from_mont(); break; case 6: r = x.neg(); break; default: r = x; }
    memcpy(o, r.l, 4 * P::N);
}
template <class P> static void ops2(int op, const u32* a, const u32* b, u32* o) {
    Fp2<P> x, y, r; memcpy(&x, a, 8 * P::N); memcpy(&y, b, 8 * P::N);
    switch (op) { case 0: r = x * y; break; case 1: r = x + y; break; case 2: r = x - y; break; case 3: r = x.inv(); break; case 4: r = x.sqr(); break; default: r = x; }
    memcpy(o, &r, 8 * P::N);
}
extern "C" void fp_op(int field, int op, const u32* a, const u32* b, u32* o) {
    switch (field) { case 0: ops<Bn254Fq>(op, a, b, o); break; case 1: ops<Bn254Fr>(op, a, b, o); break; case 2: ops<Bls381Fq>(op, a, b, o); break; case 3: ops<Bls381Fr>(op, a, b, o); break; }
}
extern "C" void fp2_op(int field, int op, const u32* a, const u32* b, u32* o) {
    switch (field) { case 0: ops2<Bn254Fq>(op, a, b, o); break; case 2: ops2<Bls381Fq>(op, a, b, o); break; }
}
// lazy dot product: o = sum_k a_k * b_k / R (+ c) with one reduction; n terms, limbs consecutive
template <class P, int LOGK> static void dotk(int n, const u32* a, const u32* b, const u32* c, u32* o) {
    FpWide<P> w;
    if (c) { Fp<P> cc; memcpy(cc.l, c, 4 * P::N); w.set_addend(cc); } else w.clear();
    for (int k = 0; k < n; k++) { Fp<P> x, y; memcpy(x.l, a + k * P::N, 4 * P::N); memcpy(y.l, b + k * P::N, 4 * P::N); w.mad(x, y); }
    Fp<P> r = w.template reduce<LOGK>();
    memcpy(o, r.l, 4 * P::N);
}
template <class P> static void dot(int logk, int n, const u32* a, const u32* b, const u32* c, u32* o) {
    switch (logk) { case 1: dotk<P, 1>(n, a, b, c, o); break; case 2: dotk<P, 2>(n, a, b, c, o); break; case 3: dotk<P, 3>(n, a, b, c, o); break; default: dotk<P, 4>(n, a, b, c, o); }
}
extern "C" void fp_dot(int field, int logk, int n, const u32* a, const u32* b, const u32* c, u32* o) {
    switch (field) { case 0: dot<Bn254Fq>(logk, n, a, b, c, o); break; case 1: dot<Bn254Fr>(logk, n, a, b, c, o); break; case 2: dot<Bls381Fq>(logk, n, a, b, c, o); break; case 3: dot<Bls381Fr>(logk, n, a, b, c, o); break; }
}
