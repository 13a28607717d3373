// C++ entry points of the STARK prover (wrapped by the C-ABI in capi.cpp).
#pragma once
#include "b200_internal.h"
#define GL_P_HOST 0xFFFFFFFF00000001ULL
namespace b200 {
struct Setup;
// setup_json = {"starkinfo": serde(StarkInfo), "program": serde(Program), "stark_struct": serde(StarkStruct)}
Setup* setup_new(const std::string& setup_json, const u64* const_rowmajor, bool const_on_device, size_t n_rows, size_t n_consts);
void setup_free(Setup* s);
void setup_export(const Setup* s, const char* path);       // serialized StarkSetup (stark_setup.rs:13-19)
Setup* setup_import(const char* path);
void setup_const_root(const Setup* s, u64 out4[4]);
void setup_shape(const Setup* s, size_t out[4]);             // nBits, nBitsExt, committed columns of stage 1, constant columns
std::string step_program_source(const std::string& setup_json, const std::string& which);   // host only (JIT debug / tests)
std::string stark_gen(Setup* s, const u64* cm_rowmajor, bool cm_on_device, size_t n_rows, size_t n_cols, const char* prover_addr);
// verify.cpp: stark_verify (stark_verify.rs:21-121) + FRI::verify (fri.rs:187-297) on the host; `why` says what failed
bool stark_verify(const std::string& setup_json, const u64 const_root[4], const std::string& proof_json, std::string& why);
// prove.rs:124-132: when on, stark_gen verifies the proof it is about to return and fails if the verifier rejects it
void setup_set_self_verify(Setup* s, bool on);
}  // namespace b200
