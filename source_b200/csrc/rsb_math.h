// rsb_math.h -- scalar fp64 building blocks of the ray/scene hot path.
//
// Every expression here is written in the exact operation order of the reference
// (file:line cited per function) and the library is compiled with -fmad=false, so the
// results are bit-identical to the reference's x86-64 (no-FMA) build for +,-,*,/,sqrt.
// The header compiles both as CUDA device code and as plain host C++ (tests build a
// host harness from the same source to pin parity without a GPU; the product library
// never exports a CPU path).
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define RSB_HD __host__ __device__ __forceinline__
#define RSB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define RSB_HD inline
#define RSB_HD_NOINLINE inline
#endif

#ifndef RSB_INF
#define RSB_INF ((double)INFINITY)
#endif

namespace rsb {

struct V3 {
    double x, y, z;
};

RSB_HD V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
RSB_HD double v3_get(const V3& v, int axis) { return axis == 0 ? v.x : (axis == 1 ? v.y : v.z); }

// raysect/core/math/vector.pyx:~280 (dot): x*x' + y*y' + z*z', summed left to right
RSB_HD double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// raysect/core/math/vector.pyx:306-310 (cross)
RSB_HD V3 cross(const V3& a, const V3& b) {
    return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}

// raysect/core/math/_vec3.pyx:152 (get_length)
RSB_HD double length(const V3& a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }

// raysect/core/math/vector.pyx:313-337, normal.pyx:~205-219 (normalise): t = 1/sqrt(|v|^2), then scale
RSB_HD V3 normalise(const V3& a) {
    double t = a.x * a.x + a.y * a.y + a.z * a.z;
    t = 1.0 / sqrt(t);
    return v3(a.x * t, a.y * t, a.z * t);
}

// raysect/core/math/point.pyx:253-281 (Point3D.transform).  m = rows 0..2 of an affine 4x4 (row-major 3x4) followed
// by m[12] = 1.0 / m33.  The reference forms w = m30*x + m31*y + m32*z + m33, then w = 1.0 / w, and multiplies every
// component by it; the bottom row of an affine matrix is (0, 0, 0, m33), so w is the per-matrix constant m33 -- exactly
// 1.0 for matrices built from translations, rotations and scalings, but AffineMatrix3D.inverse() of a non-rigid chain
// can leave m33 = 1 - 1 ulp.  The reciprocal is formed once per matrix on the host (the same IEEE division) and the
// three multiplies are kept: x * 1.0 is x, anything else is what the reference computes.
#define RSB_MAT_WORDS 13
RSB_HD V3 xform_point(const double* m, const V3& p) {
    const double w = m[12];
    return v3((m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3]) * w,
              (m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7]) * w,
              (m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]) * w);
}

// raysect/core/math/vector.pyx:339-366 (Vector3D.transform)
RSB_HD V3 xform_vector(const double* m, const V3& v) {
    return v3(m[0] * v.x + m[1] * v.y + m[2] * v.z,
              m[4] * v.x + m[5] * v.y + m[6] * v.z,
              m[8] * v.x + m[9] * v.y + m[10] * v.z);
}

// raysect/core/math/normal.pyx:222-248 (Normal3D.transform / transform_with_inverse):
// multiply by the TRANSPOSE of the inverse matrix.  minv = rows 0..2 of the inverse (3x4).
RSB_HD V3 xform_normal_with_inverse(const double* minv, const V3& n) {
    return v3(minv[0] * n.x + minv[4] * n.y + minv[8] * n.z,
              minv[1] * n.x + minv[5] * n.y + minv[9] * n.z,
              minv[2] * n.x + minv[6] * n.y + minv[10] * n.z);
}

// 3x3 transforms used by the surface-space code (row-major 3x3)
RSB_HD V3 xform_vector33(const double* m, const V3& v) {
    return v3(m[0] * v.x + m[1] * v.y + m[2] * v.z,
              m[3] * v.x + m[4] * v.y + m[5] * v.z,
              m[6] * v.x + m[7] * v.y + m[8] * v.z);
}

// raysect/core/math/vector.pyx:442-470, normal.pyx:346-370 (orthogonal)
RSB_HD V3 orthogonal(const V3& a) {
    V3 n = normalise(a);
    V3 v = v3(1, 0, 0);
    if (fabs(dot(n, v)) > 0.5) v = v3(0, 1, 0);
    double m = dot(n, v);
    v = v3(v.x - m * n.x, v.y - m * n.y, v.z - m * n.z);
    return normalise(v);
}

// raysect/core/math/cython/utility.pyx:376-420 (solve_quadratic, PBRT-stable form)
RSB_HD bool solve_quadratic(double a, double b, double c, double* t0, double* t1) {
    double d = b * b - 4 * a * c;
    if (d < 0) return false;
    double q;
    if (b < 0) q = -0.5 * (b - sqrt(d));
    else q = -0.5 * (b + sqrt(d));
    *t0 = q / a;
    *t1 = c / q;
    return true;
}

// ---- quartic roots (raysect/core/math/cython/utility.pyx:423-733), for the torus ------------------------------------------
// Transcribed operation for operation -- including one_newton_step's derivative, which is not the derivative of the quartic
// (utility.pyx:676: ((4x + 3b)x + 2d)x + e) -- because the roots' bits are the hit distances.
#ifndef RSB_PI
#define RSB_PI 3.14159265358979323846
#endif
#define RSB_EQN_EPS 1.0e-9
RSB_HD bool eqn_is_zero(double v) { return v < RSB_EQN_EPS && v > -RSB_EQN_EPS; }     // utility.pxd:91-92
RSB_HD void swap_dbl(double* a, double* b) { double t = *a; *a = *b; *b = t; }
RSB_HD void sort_three_doubles(double* a, double* b, double* c) {                       // utility.pxd:69-75
    if (*a > *b) swap_dbl(a, b);
    if (*b > *c) {
        swap_dbl(b, c);
        if (*a > *c) swap_dbl(a, c);
    }
}
RSB_HD void sort_four_doubles(double* a, double* b, double* c, double* d) {             // utility.pxd:77-89
    if (*a > *b) swap_dbl(a, b);
    if (*b > *c) swap_dbl(b, c);
    if (*c > *d) swap_dbl(c, d);
    if (*a > *b) swap_dbl(a, b);
    if (*b > *c) swap_dbl(b, c);
    if (*a > *b) swap_dbl(a, b);
}

// utility.pyx:423-497
RSB_HD int solve_cubic(double a, double b, double c, double d, double* t0, double* t1, double* t2) {
    b /= a;
    c /= a;
    d /= a;
    double sq_b = b * b;
    double q = (3.0 * c - sq_b) / 9.0;
    double r = (c * b - 3.0 * d) / 6.0 - b * sq_b / 27.0;
    double cb_q = q * q * q;
    double D = cb_q + r * r;
    if (D > 0) {
        double A = cbrt(fabs(r) + sqrt(D));
        double z0;
        if (r < 0) z0 = q / A - A;
        else z0 = A - q / A;
        *t0 = z0 - b / 3.0;
        *t1 = -0.5 * z0 - b / 3.0;
        *t2 = 0.5 * sqrt(3.0) * (A + q / A);
        return 1;
    }
    double phi;
    if (eqn_is_zero(q)) phi = 0.0;
    else phi = acos(r / sqrt(-cb_q)) / 3.0;
    double u = 2.0 * sqrt(-q);
    *t0 = u * cos(phi) - b / 3.0;
    *t1 = -u * cos(phi + RSB_PI / 3.0) - b / 3.0;
    *t2 = -u * cos(phi - RSB_PI / 3.0) - b / 3.0;
    return 3;
}

// utility.pyx:500-555
RSB_HD int solve_biquadratic(double a, double c, double e, double* t0, double* t1, double* t2, double* t3) {
    double s0, s1;
    if (!solve_quadratic(a, c, e, &s0, &s1)) return 0;
    if (s0 > s1) swap_dbl(&s0, &s1);
    if (s0 >= 0) {
        double sx0 = sqrt(s0), sx1 = sqrt(s1);
        *t0 = -sx1; *t1 = -sx0; *t2 = sx0; *t3 = sx1;
        return 4;
    } else if (s1 >= 0) {
        double sx1 = sqrt(s1);
        *t0 = -sx1; *t1 = sx1;
        return 2;
    }
    return 0;
}

// utility.pyx:558-664 (_solve_depressed_quartic): x^4 + p x^2 + q x + r through the resolvent cubic (Van der Waerden)
RSB_HD int solve_depressed_quartic(double p, double q, double r, double* t0, double* t1, double* t2, double* t3) {
    double sigma = q > 0 ? 1.0 : -1.0;
    if (eqn_is_zero(q)) return solve_biquadratic(1.0, p, r, t0, t1, t2, t3);
    int num = solve_cubic(1.0, -2.0 * p, p * p - 4.0 * r, q * q, t0, t1, t2);
    double A, B;
    if (num > 1) {
        sort_three_doubles(t0, t1, t2);
        if (!(*t0 <= 0)) return 0;
        double s0 = sqrt(-*t0);
        A = -*t1 - *t2 - 2.0 * sigma * sqrt(*t1 * *t2);
        B = -*t1 - *t2 + 2.0 * sigma * sqrt(*t1 * *t2);
        if (A >= 0 && B >= 0) {
            double sq_A = sqrt(A), sq_B = sqrt(B);
            *t0 = 0.5 * (s0 + sq_A); *t1 = 0.5 * (s0 - sq_A); *t2 = 0.5 * (-s0 + sq_B); *t3 = 0.5 * (-s0 - sq_B);
            return 4;
        } else if (A < 0 && B >= 0) {
            double sq_B = sqrt(B);
            *t0 = 0.5 * (-s0 + sq_B); *t1 = 0.5 * (-s0 - sq_B);
            return 2;
        } else if (A >= 0 && B < 0) {
            double sq_A = sqrt(A);
            *t0 = 0.5 * (s0 + sq_A); *t1 = 0.5 * (s0 - sq_A);
            return 2;
        }
        return 0;
    }
    if (!(*t0 <= 0)) return 0;
    double s0 = sqrt(-*t0);
    A = -2.0 * *t1 - 2.0 * sigma * sqrt(*t1 * *t1 + *t2 * *t2);
    B = -2.0 * *t1 + 2.0 * sigma * sqrt(*t1 * *t1 + *t2 * *t2);
    if (A >= 0 && B >= 0) {
        double sq_A = sqrt(A), sq_B = sqrt(B);
        *t0 = 0.5 * (s0 + sq_A); *t1 = 0.5 * (s0 - sq_A); *t2 = 0.5 * (-s0 + sq_B); *t3 = 0.5 * (-s0 - sq_B);
        return 4;
    } else if (A < 0 && B >= 0) {
        double sq_B = sqrt(B);
        *t0 = 0.5 * (-s0 + sq_B); *t1 = 0.5 * (-s0 - sq_B);
        return 2;
    } else if (A >= 0 && B < 0) {
        double sq_A = sqrt(A);
        *t0 = 0.5 * (s0 + sq_A); *t1 = 0.5 * (s0 - sq_A);
        return 2;
    }
    return 0;
}

// utility.pyx:667-680
RSB_HD void one_newton_step(double b, double c, double d, double e, double* x) {
    double dfx = ((4.0 * *x + 3 * b) * *x + 2.0 * d) * *x + e;
    if (!eqn_is_zero(dfx)) {
        double fx = (((*x + b) * *x + c) * *x + d) * *x + e;
        *x = *x - fx / dfx;
    }
}

// utility.pyx:683-733.  Roots the reference leaves undefined are carried as 0.0 here (it reads and shifts uninitialised
// memory; no caller looks at them).
RSB_HD int solve_quartic(double a, double b, double c, double d, double e, double* t0, double* t1, double* t2, double* t3) {
    b /= a;
    c /= a;
    d /= a;
    e /= a;
    double sq_b = b * b;
    double p = c - 3 * sq_b / 8.0;
    double q = sq_b * b / 8.0 - 0.5 * b * c + d;
    double r = -3.0 * sq_b * sq_b / 256.0 + sq_b * c / 16.0 - b * d / 4.0 + e;
    int num;
    *t0 = 0.0; *t1 = 0.0; *t2 = 0.0; *t3 = 0.0;
    if (eqn_is_zero(r)) {
        *t0 = 0;
        num = 1 + solve_cubic(1.0, 0.0, p, q, t1, t2, t3);
    } else {
        num = solve_depressed_quartic(p, q, r, t0, t1, t2, t3);
    }
    *t0 -= b / 4.0;
    *t1 -= b / 4.0;
    *t2 -= b / 4.0;
    *t3 -= b / 4.0;
    if (num > 0) {
        one_newton_step(b, c, d, e, t0);
        one_newton_step(b, c, d, e, t1);
    }
    if (num > 2) one_newton_step(b, c, d, e, t2);
    if (num > 3) one_newton_step(b, c, d, e, t3);
    return num;
}

// raysect/core/boundingbox.pyx:200-245 (_slab)
RSB_HD void box_slab(double origin, double direction, double lower, double upper, double* front, double* back) {
    double tmin, tmax;
    if (direction != 0.0) {
        double reciprocal = 1.0 / direction;
        if (direction > 0) {
            tmin = (lower - origin) * reciprocal;
            tmax = (upper - origin) * reciprocal;
        } else {
            tmin = (upper - origin) * reciprocal;
            tmax = (lower - origin) * reciprocal;
        }
    } else {
        if (origin < lower) { tmin = -RSB_INF; tmax = -RSB_INF; }
        else if (origin > upper) { tmin = RSB_INF; tmax = RSB_INF; }
        else { tmin = -RSB_INF; tmax = RSB_INF; }
    }
    if (tmin > *front) *front = tmin;
    if (tmax < *back) *back = tmax;
}

// raysect/core/boundingbox.pyx:180-198 (intersect).  box = lower xyz, upper xyz.
RSB_HD bool box_intersect(const double* box, const V3& o, const V3& d, double* front, double* back) {
    double f = -RSB_INF, b = RSB_INF;
    box_slab(o.x, d.x, box[0], box[3], &f, &b);
    box_slab(o.y, d.y, box[1], box[4], &f, &b);
    box_slab(o.z, d.z, box[2], box[5], &f, &b);
    *front = f;
    *back = b;
    if (f > b) return false;
    if ((f < 0.0) && (b < 0.0)) return false;
    return true;
}

// Same slab test with the reciprocal `1.0 / direction` supplied by the caller: every world-space AABB test of
// one ray divides by the same three direction components, so the quotients are computed once per ray and
// reused -- bit-identical to recomputing them per box (boundingbox.pyx:208-209).
RSB_HD void box_slab_inv(double origin, double direction, double reciprocal, double lower, double upper, double* front, double* back) {
    double tmin, tmax;
    if (direction != 0.0) {
        if (direction > 0) {
            tmin = (lower - origin) * reciprocal;
            tmax = (upper - origin) * reciprocal;
        } else {
            tmin = (upper - origin) * reciprocal;
            tmax = (lower - origin) * reciprocal;
        }
    } else {
        if (origin < lower) { tmin = -RSB_INF; tmax = -RSB_INF; }
        else if (origin > upper) { tmin = RSB_INF; tmax = RSB_INF; }
        else { tmin = -RSB_INF; tmax = RSB_INF; }
    }
    if (tmin > *front) *front = tmin;
    if (tmax < *back) *back = tmax;
}

RSB_HD V3 ray_reciprocals(const V3& d) { return v3(1.0 / d.x, 1.0 / d.y, 1.0 / d.z); }

// ---- exact division through a precomputed reciprocal ------------------------------------------------------
// For r = RN(1/d), q = x*r followed by the FMA residual correction q += r * fma(-d, q, x) yields the correctly
// rounded x/d (Markstein 1990) -- i.e. the very bits of the IEEE division the reference executes -- unless d's
// significand is all ones or an intermediate leaves the normal range.  The kd traversal divides by the same
// three direction components at every branch node (kdtree3d.pyx:672), so the reciprocals are formed once per
// ray; where exactness cannot be guaranteed the recipe falls back to a true division.
RSB_HD double exact_recip(double d) {
    double r = 1.0 / d;
    unsigned long long bits;
    memcpy(&bits, &d, 8);
    double ar = fabs(r);
    if ((bits & 0xFFFFFFFFFFFFFULL) == 0xFFFFFFFFFFFFFULL || !(ar > 1e-250 && ar < 1e250)) return 0.0;   // 0 = "divide"
    return r;
}

// A true division on a path that is (almost) never taken.  On the device its operands pass through an empty volatile
// asm, which pins the division inside the branch that guards it (the compiler may otherwise if-convert
// `rare ? x / d : shortcut` and run div.rn.f64's inline sequence on every lane).  Measured: nvcc 12.9 did not if-convert
// these sites -- k_wf_trace's SASS is identical with and without the pin (6,152 instructions) -- so this is a guard, not
// a speed-up; the 14 % of k_wf_trace's instructions that ncu attributes to div_recip1's guard line are the exponent
// window test and the branch themselves, executed once per kd branch node.
RSB_HD double div_rare(double x, double d) {
#ifdef __CUDA_ARCH__
    asm volatile("" : "+d"(x), "+d"(d));
#endif
    return x / d;
}

RSB_HD double div_exact(double x, double d, double r) {
    double ax = fabs(x);
    // +-0 / d for a positive finite d is x itself.  (Left to the division it costs ~100 instructions: CUDA's
    // div.rn.f64 sends a zero numerator down its slow path, and the per-sample Welford update of a dark bin divides
    // zero twice -- 31 % of k_wf_finalize's instructions in the round-1 profile.)
    if (ax == 0.0 && d > 0.0 && d < 1e300) return x;
    if (r == 0.0 || !(ax > 1e-250 && ax < 1e250)) return div_rare(x, d);
    double q = x * r;
    double e = fma(-d, q, x);
    q = fma(e, r, q);
    e = fma(-d, q, x);
    q = fma(e, r, q);
    return q;
}

// Single-correction form used by the kd traversal (one plane distance per branch node, kdtree3d.pyx:672):
// q = RN(x*r); q' = RN(q + r * (x - d*q)), the residual being exact in one FMA.  With r = RN(1/d) this is the
// correctly rounded quotient whenever d's significand is not all ones and nothing leaves the normal range
// (Markstein); `unsafe` bit k flags a direction component that fails that test, and |x| is windowed here, so
// that every other case takes a true division.  The dependent chain is 3 instructions instead of the ~10 of
// div.rn.f64 -- the traversal is bound by exactly that chain.  Checked against x / d on 4e8 adversarial pairs
// (tests/test_cabi_symbols.py::test_kd_plane_distance_through_reciprocal_is_the_ieee_quotient checks that form; 4e8 more pairs were run offline).
// Non-zero when some NON-ZERO direction component cannot take the reciprocal shortcut (all-ones significand, or a
// reciprocal outside 1e-140 .. 1e140).  Zero components never reach the division (kdtree3d.pyx:660-668).
RSB_HD int recip_unsafe_mask(const V3& d, const V3& r) {
    int m = 0;
    const double c[3] = {d.x, d.y, d.z}, q[3] = {r.x, r.y, r.z};
    for (int k = 0; k < 3; ++k) {
        unsigned long long bits;
        memcpy(&bits, &c[k], 8);
        double ar = fabs(q[k]);
        if (c[k] != 0.0 && ((bits & 0xFFFFFFFFFFFFFULL) == 0xFFFFFFFFFFFFFULL || !(ar > 1e-140 && ar < 1e140))) m |= 1 << k;
    }
    return m;
}

// `unsafe`: the ray has a flagged direction component (then every plane distance of the ray is a true division).
// The numerator's magnitude is windowed on its exponent field alone: biased exponent in [559, 1488], i.e.
// 2^-464 <= |x| < 2^466 (about 2e-140 .. 2e140); subnormal, infinite and NaN numerators divide.
RSB_HD double div_recip1(double x, double d, double r, bool unsafe) {
    unsigned long long bits;
    memcpy(&bits, &x, 8);
    unsigned int ex = ((unsigned int)(bits >> 52) & 0x7FFu) - 559u;
    // (a zero numerator -- an origin exactly on the split plane, 1.2 % of the Cornell box's branch visits -- stays on
    // the shortcut: q = +-0 and both corrections keep it zero; only the SIGN of a zero quotient may differ from the
    // division's, and the traversal compares plane distances, it never looks at the sign of a zero)
    if (unsafe || ex > 929u) {
        if (unsafe || x != 0.0) return div_rare(x, d);
    }
    double q = x * r;
    double e = fma(-d, q, x);
    return fma(e, r, q);
}

RSB_HD V3 exact_reciprocals(const V3& d) { return v3(exact_recip(d.x), exact_recip(d.y), exact_recip(d.z)); }

RSB_HD bool box_intersect_inv(const double* box, const V3& o, const V3& d, const V3& inv, double* front, double* back) {
    double f = -RSB_INF, b = RSB_INF;
    box_slab_inv(o.x, d.x, inv.x, box[0], box[3], &f, &b);
    box_slab_inv(o.y, d.y, inv.y, box[1], box[4], &f, &b);
    box_slab_inv(o.z, d.z, inv.z, box[2], box[5], &f, &b);
    *front = f;
    *back = b;
    if (f > b) return false;
    if ((f < 0.0) && (b < 0.0)) return false;
    return true;
}

RSB_HD bool box_hit_inv(const double* box, const V3& o, const V3& d, const V3& inv) {
    double f, b;
    return box_intersect_inv(box, o, d, inv, &f, &b);
}

// raysect/core/boundingbox.pyx:146-158 (hit)
RSB_HD bool box_hit(const double* box, const V3& o, const V3& d) {
    double f, b;
    return box_intersect(box, o, d, &f, &b);
}

// raysect/core/boundingbox.pyx:247-263 (contains)
RSB_HD bool box_contains(const double* box, const V3& p) {
    if ((p.x < box[0]) || (p.x > box[3])) return false;
    if ((p.y < box[1]) || (p.y > box[4])) return false;
    if ((p.z < box[2]) || (p.z > box[5])) return false;
    return true;
}

// raysect/core/math/cython/utility.pyx:40-94 (find_index): bisection over a monotonic array
RSB_HD int find_index(const double* x, int n, double v) {
    if (v < x[0]) return -1;
    int top = n - 1;
    if (v >= x[top]) return top;
    int bottom = 0;
    int bis = top / 2;
    while ((top - bottom) != 1) {
        if (v >= x[bis]) bottom = bis;
        else top = bis;
        bis = (top + bottom) / 2;
    }
    return bottom;
}

}  // namespace rsb
