// Fiat-Shamir transcripts of the STARK prover and verifier (host side): TranscriptGL (starky/src/transcript.rs:8-103) and
// TranscriptBN128 / TranscriptBLS12381 (transcript_bn128.rs:14-135).  Shared by stark.cpp (prover) and verify.cpp (verifier).
#pragma once
#include "b200_internal.h"
#include "stark.h"
#include <cstring>
#include <vector>
#include <array>

namespace b200 {

// ------------------------------------------------------------------------------------------------ transcript
struct Transcript {
    // hash == 0: TranscriptGL (transcript.rs:8-103): rate 8, capacity 4, outputs are GL lanes.
    // hash != 0: TranscriptBN128 / TranscriptBLS12381 (transcript_bn128.rs:14-135): rate 16, state = out[0]; each
    //            254-bit output yields three 64-bit limbs reduced mod p_GL; query indices take 253 bits per output.
    // The permutations run on the device either way.
    int hash = 0;
    u64 state[4] = {0, 0, 0, 0};
    std::vector<u64> pending, out;                       // GL: one u64 per element
    std::vector<std::array<u64, 4>> bpending, bout;      // big: canonical 4 x u64 per element
    std::vector<u64> out3;
    explicit Transcript(int h = 0) : hash(h) {}
    void update() {
        if (hash == 0) {
            while (pending.size() < 8) pending.push_back(0);
            u64 in[12], o[12];
            for (int i = 0; i < 8; i++) in[i] = pending[i];
            for (int i = 0; i < 4; i++) in[8 + i] = state[i];
            poseidon_perm_host(in, o);
            out.assign(o, o + 12); pending.clear();
            memcpy(state, o, 32);
        } else {
            while (bpending.size() < 16) bpending.push_back({0, 0, 0, 0});
            u64 in[17 * 4], o[17 * 4];
            memcpy(in, state, 32);
            for (int i = 0; i < 16; i++) memcpy(in + 4 * (1 + i), bpending[i].data(), 32);
            big_poseidon_host(hash - 1, in, 17, o);                       // hash_ex(.., 17)
            bout.clear(); for (int i = 0; i < 17; i++) { std::array<u64, 4> e; memcpy(e.data(), o + 4 * i, 32); bout.push_back(e); }
            out3.clear(); bpending.clear();
            memcpy(state, o, 32);
        }
    }
    void put_elem(const std::array<u64, 4>& e) { bout.clear(); bpending.push_back(e); if (bpending.size() == 16) update(); }   // add_1: out3 is NOT cleared (transcript_bn128.rs:33-40)
    void put1(u64 e) {
        if (hash == 0) { out.clear(); pending.push_back(e); if (pending.size() == 8) update(); }
        else put_elem({e, 0, 0, 0});
    }
    void put(const u64* e, size_t n) { for (size_t i = 0; i < n; i++) put1(e[i]); }
    void put_digest(const u64 d[4]) { if (hash == 0) put(d, 4); else put_elem({d[0], d[1], d[2], d[3]}); }
    u64 get1() {
        if (hash == 0) { while (out.empty()) update(); u64 v = out.front(); out.erase(out.begin()); return v; }
        for (;;) {
            if (!out3.empty()) { u64 v = out3.front(); out3.erase(out3.begin()); return v; }
            if (!bout.empty()) {
                std::array<u64, 4> v = bout.front(); bout.erase(bout.begin());
                for (int k = 0; k < 3; k++) out3.push_back(v[k] >= GL_P_HOST ? v[k] - GL_P_HOST : v[k]);     // helper.rs:61-65
                continue;
            }
            update();
        }
    }
    void get_field(u64 f[3]) { f[0] = get1(); f[1] = get1(); f[2] = get1(); }
    std::vector<u64> get_permutations(size_t n, size_t nbits) {
        std::vector<u64> res;
        if (hash == 0) {
            size_t total = n * nbits, nf = (total - 1) / 63 + 1;
            std::vector<u64> fields; for (size_t i = 0; i < nf; i++) fields.push_back(get1());
            size_t cf = 0, cb = 0;
            for (size_t i = 0; i < n; i++) { u64 a = 0; for (size_t j = 0; j < nbits; j++) { if ((fields[cf] >> cb) & 1) a += 1ull << j; if (++cb == 63) { cb = 0; cf++; } } res.push_back(a); }
            return res;
        }
        size_t total = n * nbits, nf = (total - 1) / 253 + 1;
        std::vector<std::array<u64, 4>> fields;
        for (size_t i = 0; i < nf; i++) { while (bout.empty()) update(); fields.push_back(bout.front()); bout.erase(bout.begin()); }     // get_fields253
        size_t cf = 0, cb = 0;
        for (size_t i = 0; i < n; i++) { u64 a = 0; for (size_t j = 0; j < nbits; j++) { if ((fields[cf][cb >> 6] >> (cb & 63)) & 1) a += 1ull << j; if (++cb == 253) { cb = 0; cf++; } } res.push_back(a); }
        return res;
    }
};


}  // namespace b200
