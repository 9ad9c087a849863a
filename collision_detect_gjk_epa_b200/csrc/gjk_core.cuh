// gjk_core.cuh -- GJK intersection test for one convex-hull pair, run by a sub-warp of L lanes.
//
// Path: GJKEPA driver (reference src/GCLIB_GJKEPA.f90:76-236, "F:<line>") with support_mapping
// (F:1030-1062), update_simplex_GJK (F:1070-1157), isPointInSimplex (F:1217-1265), IS_INSIDE_PF
// (F:1271-1337), UTZVEC/UNINML/DIST_PF_SIGN/VEC_PL (F:1343-1440).  This is not a translation:
//  * one pair is worked by L in {1,2,4} lanes; lane l owns tetrahedron faces l, l+L, ... and hull
//    vertices l, l+L, ...; the support function is a sub-warp (value,index) max-dot reduction;
//  * the four face normals are computed ONCE per simplex and shared by the containment test of
//    this walk step and the face pick of the next (the reference computes them twice), and the
//    degeneracy tests reuse face (1,2,3)'s cross product;
//  * the two opposite initial supports come from one pass (first arg-max and first arg-min).
// Flags must be bit-exact with the reference arithmetic, so every product/sum that feeds a
// comparison goes through mul()/add() (never contracted to FMA), divisions and square roots are
// IEEE, scans keep the first index on ties and reductions break ties towards the lower index.
//
// The L=1 instantiation is plain C++ and is also compiled by g++ (tests/host_sim) to check the
// shared logic against the oracle on a machine without a GPU.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define GJK_HD __host__ __device__ __forceinline__
#define GJK_HD_NOINLINE __host__ __device__ __noinline__
#else
#define GJK_HD inline
#define GJK_HD_NOINLINE
#endif

#include "fp64_exact.cuh"

#define GJK_NS gjk
#define GJK_REAL double
#define GJK_REAL_F64 1
#define GJK_REAL_MAX DBL_MAX
// the reference's thresholds: on-plane / duplicate-point bands F:106,123,140,157,199,203,1248; zero-vector tests F:1349,1388;
// 'positive' in IS_INSIDE_PF F:1305; the margin of shortcut E1
#define GJK_T8 1.0e-8
#define GJK_T12 1.0e-12
#define GJK_T15 1.0e-15
#define GJK_TE1 1.0e-6
#define GJK_C(x) ((real)(x))
#include "gjk_core_body.inc"
#undef GJK_NS
#undef GJK_REAL
#undef GJK_REAL_F64
#undef GJK_REAL_MAX
#undef GJK_T8
#undef GJK_T12
#undef GJK_T15
#undef GJK_TE1
