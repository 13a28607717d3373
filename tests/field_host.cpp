// Host build of csrc/field.cuh (PTX carry flag emulated, see mont.cuh) for CPU-side unit tests.  Test harness only.
#include "field.cuh"
extern "C" {
uint64_t t_gl_add(uint64_t a, uint64_t b) { return gl_add(a, b); }
uint64_t t_gl_sub(uint64_t a, uint64_t b) { return gl_sub(a, b); }
uint64_t t_gl_addw(uint64_t a, uint64_t b) { return gl_addw(a, b); }
uint64_t t_gl_mul(uint64_t a, uint64_t b) { return gl_mul(a, b); }
uint64_t t_gl_mulw(uint64_t a, uint64_t b) { return gl_mulw(a, b); }
uint64_t t_gl_maddw(uint64_t a, uint64_t b, uint64_t c) { return gl_maddw(a, b, c); }
uint64_t t_gl_red128(uint64_t lo, uint64_t hi) { return gl_red128(lo, hi); }
uint64_t t_gl_red128w(uint64_t lo, uint64_t hi) { return gl_red128w(lo, hi); }
uint64_t t_gl_red96(uint64_t lo, uint32_t hi) { return gl_red96(lo, hi); }
uint64_t t_gl_red96w(uint64_t lo, uint32_t hi) { return gl_red96w(lo, hi); }
uint64_t t_gl_canon(uint64_t a) { return gl_canon(a); }
uint64_t t_gl_inv(uint64_t a) { return gl_inv(a); }
void t_gl_mulwide_add(uint64_t a, uint64_t b, uint64_t c, uint64_t* lo, uint64_t* hi) { gl_mulwide_add(a, b, c, *lo, *hi); }
void t_f3_mul(const uint64_t* a, const uint64_t* b, uint64_t* o) { f3 r = f3_mul(f3_make(a[0], a[1], a[2]), f3_make(b[0], b[1], b[2])); o[0] = r.c[0]; o[1] = r.c[1]; o[2] = r.c[2]; }
}
