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


// ---- nalgebra's closed-form inverses (linalg/inverse.rs try_inverse_mut, dimensions 1-3): cofactors over the determinant,
// false when the determinant is exactly zero. Row-major m[i][j]. Used by the least-squares gradients (solver.rs:803-869,
// 903-947, 624-693), whose normal equations nalgebra accumulates entry by entry in ascending inner index. ----
ORC_HD bool inv_n(int n, const double (&m)[3][3], double (&o)[3][3]) {
    if (n == 1) {
        if (m[0][0] == 0.) return false;
        o[0][0] = 1. / m[0][0];
        return true;
    }
    if (n == 2) {
        const double m11 = m[0][0], m12 = m[0][1], m21 = m[1][0], m22 = m[1][1];
        const double det = m11 * m22 - m21 * m12;
        if (det == 0.) return false;
        o[0][0] = m22 / det; o[0][1] = -m12 / det;
        o[1][0] = -m21 / det; o[1][1] = m11 / det;
        return true;
    }
    const double m11 = m[0][0], m12 = m[0][1], m13 = m[0][2], m21 = m[1][0], m22 = m[1][1], m23 = m[1][2], m31 = m[2][0], m32 = m[2][1],
                 m33 = m[2][2];
    const double minor_m12_m23 = m22 * m33 - m32 * m23;
    const double minor_m11_m23 = m21 * m33 - m31 * m23;
    const double minor_m11_m22 = m21 * m32 - m31 * m22;
    const double det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
    if (det == 0.) return false;
    o[0][0] = minor_m12_m23 / det;
    o[0][1] = (m13 * m32 - m33 * m12) / det;
    o[0][2] = (m12 * m23 - m22 * m13) / det;
    o[1][0] = -minor_m11_m23 / det;
    o[1][1] = (m11 * m33 - m31 * m13) / det;
    o[1][2] = (m13 * m21 - m23 * m11) / det;
    o[2][0] = minor_m11_m22 / det;
    o[2][1] = (m12 * m31 - m32 * m11) / det;
    o[2][2] = (m11 * m22 - m21 * m12) / det;
    return true;
}
// normal equations of one least-squares fit: C = A^T A, rhs_q = A^T b_q for NB right-hand sides, accumulated row by row
template <int NB>
struct Lsq3 {
    double c[3][3];
    double r[NB][3];
    int rows;
    ORC_HD void clear() {
        rows = 0;
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) c[i][j] = 0.; for (int q = 0; q < NB; ++q) r[q][i] = 0.; }
    }
    ORC_HD void add(V3 x, const double (&b)[NB]) {   // gemv_uninit: the first term is assigned, later ones are added to the sum
        const double a[3] = {x.x, x.y, x.z};
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) { const double t = a[i] * a[j]; c[i][j] = rows == 0 ? t : t + c[i][j]; }
            for (int q = 0; q < NB; ++q) { const double t = a[i] * b[q]; r[q][i] = rows == 0 ? t : t + r[q][i]; }
        }
        ++rows;
    }
    // g_q = C^-1 rhs_q; false when C is singular (the reference unwraps a None there)
    ORC_HD bool solve(V3 (&g)[NB]) const {
        double inv[3][3];
        if (!inv_n(3, c, inv)) return false;
        for (int q = 0; q < NB; ++q) {
            double o[3];
            for (int i = 0; i < 3; ++i) { double acc = inv[i][0] * r[q][0]; acc = inv[i][1] * r[q][1] + acc; acc = inv[i][2] * r[q][2] + acc; o[i] = acc; }
            g[q] = v3(o[0], o[1], o[2]);
        }
        return true;
    }
};

}  // namespace orc
