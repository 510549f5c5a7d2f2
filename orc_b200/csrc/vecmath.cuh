// vecmath.cuh — 3-vector / 3x3-tensor arithmetic with the reference's operator semantics
// (src/lib.rs:223-606): one IEEE rounding per operator, left-associative, no FMA contraction
// (device code is compiled with -fmad=false, host code with -ffp-contract=off).
// Shared by the host mesh builder and the device assembly kernels.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define ORC_HD __host__ __device__ __forceinline__
#else
#define ORC_HD inline
#endif

namespace orc {

struct V3 {
    double x, y, z;
};
struct T3 {
    V3 x, y, z;
};

ORC_HD V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
ORC_HD V3 vzero() { return v3(0., 0., 0.); }
ORC_HD V3 vadd(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }   // lib.rs:368-378
ORC_HD V3 vadds(V3 a, double s) { return v3(a.x + s, a.y + s, a.z + s); }     // lib.rs:356-366
ORC_HD V3 vsub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }   // lib.rs:402-412
ORC_HD V3 vsubs(V3 a, double s) { return v3(a.x - s, a.y - s, a.z - s); }     // lib.rs:390-400
ORC_HD V3 vneg(V3 a) { return v3(-a.x, -a.y, -a.z); }                          // lib.rs:529-538
ORC_HD V3 vmuls(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }     // lib.rs:479-508 `Vector * Float`
// lib.rs:540-549 `Float * Vector`: the z component is built from rhs.y. Kept on purpose:
// parity with the reference's CPU path means reproducing it at every site that uses this operator.
ORC_HD V3 smulv_q1(double s, V3 a) { return v3(a.x * s, a.y * s, a.y * s); }
ORC_HD V3 vdivs(V3 a, double s) { return v3(a.x / s, a.y / s, a.z / s); }     // lib.rs:429-447
ORC_HD V3 vdivv(V3 a, V3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }   // lib.rs:450-459
ORC_HD double vdot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // lib.rs:240-242
ORC_HD V3 vcross(V3 a, V3 b) {                                                 // lib.rs:254-260
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
ORC_HD double vnorm(V3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }  // lib.rs:262-264, powi(2) == x*x
ORC_HD V3 vunit(V3 a) { double l = vnorm(a); return v3(a.x / l, a.y / l, a.z / l); }  // lib.rs:266-273
ORC_HD T3 vouter(V3 a, V3 b) {                                                 // lib.rs:275-293
    T3 t;
    t.x = v3(a.x * b.x, a.x * b.y, a.x * b.z);
    t.y = v3(a.y * b.x, a.y * b.y, a.y * b.z);
    t.z = v3(a.z * b.x, a.z * b.y, a.z * b.z);
    return t;
}
ORC_HD T3 tadd(T3 a, T3 b) { T3 t; t.x = vadd(a.x, b.x); t.y = vadd(a.y, b.y); t.z = vadd(a.z, b.z); return t; }  // lib.rs:608-617
ORC_HD V3 tinner(T3 t, V3 v) { return v3(vdot(t.x, v), vdot(t.y, v), vdot(t.z, v)); }  // lib.rs:584-590

}  // namespace orc
