/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).  Exported wrappers of the inline
 * Goldilocks helpers so the Python tests can cross-check them against big-int arithmetic. */
#include "oracle.h"
uint64_t orc_gl_mul(uint64_t a, uint64_t b) { return gl_mul(a, b); }
uint64_t orc_gl_add(uint64_t a, uint64_t b) { return gl_add(a, b); }
uint64_t orc_gl_sub(uint64_t a, uint64_t b) { return gl_sub(a, b); }
uint64_t orc_gl_inv(uint64_t a) { return gl_inv(a); }
