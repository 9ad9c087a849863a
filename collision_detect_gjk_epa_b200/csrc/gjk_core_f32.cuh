// gjk_core_f32.cuh -- the flag path of gjk_core.cuh with binary32 operands (namespace gjk32): SURVEY.md 8(f)2, "same
// kernels templated on float".  The text is gjk_core_body.inc, the one the binary64 product path is built from; here
// real = float, products and sums are __fmul_rn / __fadd_rn / __fsub_rn (not contracted, like the binary64 build),
// sqrt and '/' are the IEEE binary32 operators.  The reference's thresholds are absolute numbers chosen for binary64 at unit
// scale -- 1e-8 ~ sqrt(eps) for its on-plane / duplicate bands (F:106,123,140,157,199,203,1248), 1e-12 for its zero-vector
// tests (F:1349,1388), 1e-15 for 'positive' (F:1305) -- and sit below binary32's rounding noise (a dot product of unit-scale
// float vectors is uncertain by ~1e-6): kept as they are, separated pairs never take the coplanar exit F:203 and 1 % of
// them walk on into a flat tetrahedron that 'contains' the origin (measured, tests/test_host_sim.py history).  The
// binary32 build scales them to its own precision: 1e-8 -> 1e-5, 1e-12 -> 1e-7, 1e-15 -> 1e-9, E1's margin 1e-6 -> 1e-4.
// With that the walk takes the same exits after the same number of iterations as the binary64 one, and its flags differ only
// for pairs that touch to within the band: tests/test_host_sim.py (CPU), tests/test_gpu_f32_mode.py and bench.py report the rate.
#pragma once
#include "gjk_core.cuh"

#define GJK_NS gjk32
#define GJK_REAL float
#define GJK_REAL_F64 0
#define GJK_REAL_MAX FLT_MAX
#define GJK_T8 1.0e-5f
#define GJK_T12 1.0e-7f
#define GJK_T15 1.0e-9f
#define GJK_TE1 1.0e-4f
#include "gjk_core_body.inc"
#undef GJK_NS
#undef GJK_REAL
#undef GJK_REAL_F64
#undef GJK_REAL_MAX
#undef GJK_T8
#undef GJK_T12
#undef GJK_T15
#undef GJK_TE1
