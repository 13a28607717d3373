// Poseidon-GL on the HOST, for the Fiat-Shamir transcript and the log2(N) digests of a width-0 tree only.
// Reference: starky/src/poseidon_opt.rs:80-200 (the same optimised round structure as the device kernels in poseidon.cuh),
// called from TranscriptGL (starky/src/transcript.rs:45-75).  The transcript is a chain of single permutations whose inputs are
// 32-byte roots and a handful of evaluations: in the reference it is host code too, and running each one as a 1-thread kernel
// cost a launch, two copies and a stream synchronisation (23+ per proof, VERDICT r1 "host-side drags").  Everything that hashes
// DATA (leaves, tree levels, FRI layers) stays on the device; this file is part of the product library, not of oracle/.
#include "b200_internal.h"
#include "field.cuh"
#include "poseidon_gl_params.h"

namespace b200 {

static inline u64 pow7(u64 x) { u64 x2 = gl_mul(x, x), x3 = gl_mul(x2, x), x6 = gl_mul(x3, x3); return gl_mul(x6, x); }

void poseidon12_host(const u64 in12[12], u64 out12[12]) {
    u64 st[12], t[12];
    for (int i = 0; i < 12; i++) st[i] = gl_add(in12[i] % GL_P, POS_C[i] % GL_P);
    auto mix = [&](const uint64_t* Mx) {           // st'[i] = sum_j Mx[j][i] st[j]
        for (int i = 0; i < 12; i++) { u64 acc = 0; for (int j = 0; j < 12; j++) acc = gl_add(acc, gl_mul(Mx[j * 12 + i] % GL_P, st[j])); t[i] = acc; }
        for (int i = 0; i < 12; i++) st[i] = t[i];
    };
    for (int r = 0; r < 3; r++) {
        for (int i = 0; i < 12; i++) st[i] = gl_add(pow7(st[i]), POS_C[(r + 1) * 12 + i] % GL_P);
        mix(POS_M);
    }
    for (int i = 0; i < 12; i++) st[i] = gl_add(pow7(st[i]), POS_C[4 * 12 + i] % GL_P);
    mix(POS_P);
    for (int r = 0; r < 22; r++) {
        const uint64_t* S = POS_S + 23 * r;
        st[0] = gl_add(pow7(st[0]), POS_C[5 * 12 + r] % GL_P);
        u64 s0 = 0;
        for (int j = 0; j < 12; j++) s0 = gl_add(s0, gl_mul(S[j] % GL_P, st[j]));
        for (int k = 1; k < 12; k++) st[k] = gl_add(st[k], gl_mul(S[11 + k] % GL_P, st[0]));
        st[0] = s0;
    }
    for (int r = 0; r < 3; r++) {
        for (int i = 0; i < 12; i++) st[i] = gl_add(pow7(st[i]), POS_C[5 * 12 + 22 + r * 12 + i] % GL_P);
        mix(POS_M);
    }
    for (int i = 0; i < 12; i++) st[i] = pow7(st[i]);
    mix(POS_M);
    for (int i = 0; i < 12; i++) out12[i] = st[i];
}

}  // namespace b200
