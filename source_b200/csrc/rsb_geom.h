// rsb_geom.h -- ray/primitive intersection, kd-tree traversal, World.hit / World.contains.
//
// Restates, in iterative GPU form, the reference call stack of SURVEY 3.4:
//   World.hit -> KDTree.hit -> KDTree3DCore._trace/_trace_branch -> _PrimitiveKDTree._trace_leaf
//   -> BoundPrimitive.hit -> {Sphere,Box,Cylinder,Cone,CSG,Mesh}.hit
// The reference recurses and allocates an Intersection per candidate; here the leaf loop keeps
// only (t, primitive, face code) for the running closest candidate and the full intersection
// geometry is generated once, for the winner, by re-running the same arithmetic.
#pragma once
#include "rsb_scene.h"

namespace rsb {

#define RSB_KD_STACK 64       // far-child stack entries shared by the world tree and a nested mesh tree
#define RSB_CSG_MAX_EVENTS 12 // surface crossings materialised per CSG node (<= 6 convex leaves)
#define RSB_CSG_MAX_DEPTH 4   // nesting of CSG operators below a world-level CSG primitive

// face / surface-type codes carried from hit() to _generate_intersection()
//  box      : code = axis*2 + (face==UPPER)         raysect/primitive/box.pyx:42-55
//  cylinder : 0 = CYLINDER body, 1 = SLAB lower, 2 = SLAB upper   cylinder.pyx:44-55
//  cone     : 0 = CONE, 1 = BASE                                    cone.pyx:44-47
struct Crossing {
    double t;
    int32_t code;
};

struct KdStackEntry {
    double tmax;
    int32_t node;
    int32_t pad;
};

// Full intersection record in the hit primitive's local space
// (raysect/core/intersection.pxd:37-53 + mesh.pxd:37-41).
struct Isect {
    double t;
    V3 hit, inside, outside, normal;
    int32_t prim;       // world-level primitive row
    int32_t leaf;       // row of the analytic leaf that was hit (== prim unless CSG)
    int32_t code;       // face code | triangle id
    int32_t exiting;
    float u, v, w;      // mesh barycentrics (MeshIntersection)
};

// Compact result of the closest-hit search
struct HitRec {
    double t;
    int32_t prim;
    int32_t leaf;
    int32_t code;
    int32_t flip;       // bit0: CSG Subtract parity (normal negated, inside/outside swapped); bit1: CSG exiting flag
    int32_t node;       // world kd-tree leaf (reference node id) in which the hit was accepted
    int32_t mesh_node;  // mesh kd-tree leaf for mesh hits, else -1
    float u, v, w;
};

struct TraverseStats {   // roofline counters; only touched when a kernel is built with counting on
    unsigned long long branches, leaves, items, prim_tests, tri_tests;
};

// ---------------------------------------------------------------------------------------------
// Analytic primitives.  Each *_crossings returns the 0, 1 or 2 surface crossings that the
// reference's hit() followed by next_intersection() would report for a local-space ray
// (o, d, max_distance), in that order.
// ---------------------------------------------------------------------------------------------

// raysect/primitive/sphere.pyx:115-163
RSB_HD int sphere_crossings(const double* params, const V3& o, const V3& d, double max_distance, Crossing* out) {
    double radius = params[0];
    double a = d.x * d.x + d.y * d.y + d.z * d.z;
    double b = 2 * (d.x * o.x + d.y * o.y + d.z * o.z);
    double c = o.x * o.x + o.y * o.y + o.z * o.z - radius * radius;
    double t0, t1;
    if (!solve_quadratic(a, b, c, &t0, &t1)) return 0;
    if (t0 > t1) { double tmp = t0; t0 = t1; t1 = tmp; }
    if (t0 > max_distance || t1 < 0.0) return 0;
    if (t0 >= 0.0) {
        out[0].t = t0; out[0].code = 0;
        if (t1 <= max_distance) { out[1].t = t1; out[1].code = 0; return 2; }
        return 1;
    } else if (t1 <= max_distance) {
        out[0].t = t1; out[0].code = 0;
        return 1;
    }
    return 0;
}

// raysect/primitive/sphere.pyx:165-200 (_generate_intersection)
RSB_HD void sphere_geometry(const V3& o, const V3& d, double t, Isect* is) {
    const double EPSILON = 1e-9;
    V3 hit = v3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z);
    V3 n = normalise(hit);
    double dx = EPSILON * n.x, dy = EPSILON * n.y, dz = EPSILON * n.z;
    is->hit = hit;
    is->normal = n;
    is->inside = v3(hit.x - dx, hit.y - dy, hit.z - dz);
    is->outside = v3(hit.x + dx, hit.y + dy, hit.z + dz);
    is->exiting = dot(d, n) >= 0.0;
}

// raysect/primitive/box.pyx:232-287 (_slab with face/axis tracking)
RSB_HD void prim_box_slab(int axis, double origin, double direction, double lower, double upper,
                          double* near_t, double* far_t, int* near_code, int* far_code) {
    double tmin, tmax;
    int fmin, fmax;   // -1 NO_FACE, 0 LOWER_FACE, 1 UPPER_FACE
    if (direction != 0.0) {
        double reciprocal = 1.0 / direction;
        if (direction > 0) {
            tmin = (lower - origin) * reciprocal;
            tmax = (upper - origin) * reciprocal;
            fmin = 0; fmax = 1;
        } else {
            tmin = (upper - origin) * reciprocal;
            tmax = (lower - origin) * reciprocal;
            fmin = 1; fmax = 0;
        }
    } else {
        if (origin < lower) { tmin = -RSB_INF; tmax = -RSB_INF; }
        else if (origin > upper) { tmin = RSB_INF; tmax = RSB_INF; }
        else { tmin = -RSB_INF; tmax = RSB_INF; }
        fmin = -1; fmax = -1;
    }
    // code = axis*2 + upper; NO_FACE is encoded as "upper" because the reference's normal
    // selection is `-1 if face == LOWER_FACE else +1` (box.pyx:303-306)
    if (tmin > *near_t) { *near_t = tmin; *near_code = axis * 2 + (fmin == 0 ? 0 : 1); }
    if (tmax < *far_t) { *far_t = tmax; *far_code = axis * 2 + (fmax == 0 ? 0 : 1); }
}

// raysect/primitive/box.pyx:157-219
RSB_HD int box_crossings(const double* params, const V3& o, const V3& d, double max_distance, Crossing* out) {
    double near_t = -RSB_INF, far_t = RSB_INF;
    int near_code = 0, far_code = 0;
    prim_box_slab(0, o.x, d.x, params[0], params[3], &near_t, &far_t, &near_code, &far_code);
    prim_box_slab(1, o.y, d.y, params[1], params[4], &near_t, &far_t, &near_code, &far_code);
    prim_box_slab(2, o.z, d.z, params[2], params[5], &near_t, &far_t, &near_code, &far_code);
    if (near_t > far_t) return 0;
    if (near_t > max_distance || far_t < 0.0) return 0;
    if (near_t >= 0.0) {
        out[0].t = near_t; out[0].code = near_code;
        if (far_t <= max_distance) { out[1].t = far_t; out[1].code = far_code; return 2; }
        return 1;
    } else if (far_t <= max_distance) {
        out[0].t = far_t; out[0].code = far_code;
        return 1;
    }
    return 0;
}

// raysect/primitive/box.pyx:330-342 (_interior_offset)
RSB_HD double box_interior_offset(double hit, double lower, double upper) {
    const double EPSILON = 1e-9;
    if (fabs(hit - lower) < EPSILON) return EPSILON;
    else if (fabs(hit - upper) < EPSILON) return -EPSILON;
    return 0.0;
}

// raysect/primitive/box.pyx:289-328 (_generate_intersection)
RSB_HD void box_geometry(const double* params, const V3& o, const V3& d, double t, int code, Isect* is) {
    const double EPSILON = 1e-9;
    V3 hit = v3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z);
    int axis = code >> 1;
    double s = (code & 1) ? 1.0 : -1.0;
    V3 n = v3(0, 0, 0);
    if (axis == 0) n.x = s; else if (axis == 1) n.y = s; else n.z = s;
    is->hit = hit;
    is->normal = n;
    is->inside = v3(hit.x + box_interior_offset(hit.x, params[0], params[3]),
                    hit.y + box_interior_offset(hit.y, params[1], params[4]),
                    hit.z + box_interior_offset(hit.z, params[2], params[5]));
    is->outside = v3(hit.x + EPSILON * n.x, hit.y + EPSILON * n.y, hit.z + EPSILON * n.z);
    is->exiting = dot(d, n) >= 0.0;
}

// raysect/primitive/cylinder.pyx:148-271
RSB_HD int cylinder_crossings(const double* params, const V3& o, const V3& d, double max_distance, Crossing* out) {
    double radius = params[0], height = params[1];
    double near_t, far_t;
    int near_code, far_code;   // 0 body, 1 lower slab, 2 upper slab
    if (d.x == 0 && d.y == 0) {
        // ray parallel to the axis: inside the infinite cylinder or a miss (cylinder.pyx:166-181)
        if ((o.x * o.x + o.y * o.y) <= (radius * radius)) {
            near_t = -RSB_INF; far_t = RSB_INF;
            // NO_TYPE/NO_FACE: the reference's normal selection falls to "slab, not LOWER_FACE" => +z
            near_code = 2; far_code = 2;
        } else {
            return 0;
        }
    } else {
        double a = d.x * d.x + d.y * d.y;
        double b = 2.0 * (d.x * o.x + d.y * o.y);
        double c = o.x * o.x + o.y * o.y - radius * radius;
        double t0, t1;
        if (!solve_quadratic(a, b, c, &t0, &t1)) return 0;
        if (t0 > t1) { double tmp = t0; t0 = t1; t1 = tmp; }
        near_t = t0; far_t = t1;
        near_code = 0; far_code = 0;
    }
    if (d.z != 0.0) {
        double temp = 1.0 / d.z;
        double t0, t1;
        int f0, f1;
        if (d.z > 0) {
            t0 = -o.z * temp;
            t1 = (height - o.z) * temp;
            f0 = 1; f1 = 2;
        } else {
            t0 = (height - o.z) * temp;
            t1 = -o.z * temp;
            f0 = 2; f1 = 1;
        }
        if (t0 > near_t) { near_t = t0; near_code = f0; }
        if (t1 < far_t) { far_t = t1; far_code = f1; }
    }
    if (near_t > far_t) return 0;
    if (near_t > max_distance || far_t < 0.0) return 0;
    if (near_t >= 0.0) {
        out[0].t = near_t; out[0].code = near_code;
        if (far_t <= max_distance) { out[1].t = far_t; out[1].code = far_code; return 2; }
        return 1;
    } else if (far_t <= max_distance) {
        out[0].t = far_t; out[0].code = far_code;
        return 1;
    }
    return 0;
}

// raysect/primitive/cylinder.pyx:282-349 (_generate_intersection, _interior_offset)
RSB_HD void cylinder_geometry(const double* params, const V3& o, const V3& d, double t, int code, Isect* is) {
    const double EPSILON = 1e-9;
    double radius = params[0], height = params[1];
    V3 hit = v3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z);
    V3 n;
    if (code == 0) n = normalise(v3(hit.x, hit.y, 0));
    else if (code == 1) n = v3(0, 0, -1);
    else n = v3(0, 0, 1);
    double x, y, z;
    if (code == 0) {
        x = -EPSILON * n.x;
        y = -EPSILON * n.y;
    } else {
        x = 0; y = 0;
        if (hit.x != 0.0 && hit.y != 0.0) {
            double len = sqrt(hit.x * hit.x + hit.y * hit.y);
            if ((len - radius) < EPSILON) {
                len = 1.0 / len;
                x = -EPSILON * len * hit.x;
                y = -EPSILON * len * hit.y;
            }
        }
    }
    if (fabs(hit.z) < EPSILON) z = EPSILON;
    else if (fabs(hit.z - height) < EPSILON) z = -EPSILON;
    else z = 0;
    is->hit = hit;
    is->normal = n;
    is->inside = v3(hit.x + x, hit.y + y, hit.z + z);
    is->outside = v3(hit.x + EPSILON * n.x, hit.y + EPSILON * n.y, hit.z + EPSILON * n.z);
    is->exiting = dot(d, n) >= 0.0;
}

// raysect/primitive/cone.pyx:142-262.  Codes: 0 CONE, 1 BASE.
RSB_HD int cone_crossings(const double* params, const V3& o, const V3& d, double max_distance, Crossing* out) {
    double radius = params[0], height = params[1];
    double k = radius / height;
    k = k * k;
    double a = d.x * d.x + d.y * d.y - k * d.z * d.z;
    double b = 2 * (d.x * o.x + d.y * o.y - k * d.z * (o.z - height));
    double c = o.x * o.x + o.y * o.y - k * (o.z - height) * (o.z - height);
    double t0, t1;
    int t0_type, t1_type;
    if (!solve_quadratic(a, b, c, &t0, &t1)) return 0;
    if (t0 == t1) {
        // ray passes through the tip (cone.pyx:176-191, including the reference's
        // `direction.y**2` placement at :185, reproduced verbatim)
        t0 = -b / (2.0 * a);
        t0_type = 0;
        k = -o.z / d.z;
        double ex = o.x + k * d.x;
        double r2 = ex * ex + (o.y + k * (d.y * d.y));
        if (r2 <= (radius * radius)) { t1 = k; t1_type = 1; }
        else { t1 = t0; t1_type = t0_type; }
    } else {
        double t0_z = o.z + t0 * d.z;
        double t1_z = o.z + t1 * d.z;
        bool t0_outside = t0_z < 0 || t0_z > height;
        bool t1_outside = t1_z < 0 || t1_z > height;
        if (t0_outside && t1_outside) return 0;
        else if (!t0_outside && t1_outside) { t0_type = 0; t1 = -o.z / d.z; t1_type = 1; }
        else if (t0_outside && !t1_outside) { t0_type = 1; t0 = -o.z / d.z; t1_type = 0; }
        else { t0_type = 0; t1_type = 0; }
    }
    if (t0 > t1) {
        double tmp = t0; t0 = t1; t1 = tmp;
        int ti = t0_type; t0_type = t1_type; t1_type = ti;
    }
    if (t0 > max_distance || t1 < 0.0) return 0;
    if (t0 >= 0.0) {
        out[0].t = t0; out[0].code = t0_type;
        if (t1 <= max_distance) { out[1].t = t1; out[1].code = t1_type; return 2; }
        return 1;
    } else if (t1 <= max_distance) {
        out[0].t = t1; out[0].code = t1_type;
        return 1;
    }
    return 0;
}

// raysect/primitive/cone.pyx:273-357 (_generate_intersection, _interior_point)
RSB_HD void cone_geometry(const double* params, const V3& o, const V3& d, double t, int code, Isect* is) {
    const double EPSILON = 1e-9;
    double radius = params[0], height = params[1];
    V3 hit = v3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z);
    V3 n;
    if (code == 1) {
        n = v3(0, 0, -1);
    } else if (hit.z >= height) {
        n = v3(0, 0, 1);
    } else {
        double a = hit.y / hit.x;
        double b = height / sqrt(1 + a * a);
        b = hit.x < 0 ? -b : b;
        n = normalise(v3(b, b * a, radius));
    }
    // _interior_point
    V3 inside;
    {
        double x = hit.x - EPSILON * n.x;
        double y = hit.y - EPSILON * n.y;
        double z = hit.z - EPSILON * n.z;
        double k = radius / height;
        double inner_height = height - EPSILON * sqrt(1 + k * k) / k;
        if (z > inner_height) {
            inside = v3(0, 0, inner_height);
        } else if (z < EPSILON) {
            double inner_radius = k * (height - EPSILON) - EPSILON * sqrt(1 + k * k);
            double scale = inner_radius / sqrt(hit.x * hit.x + hit.y * hit.y);
            inside = v3(scale * hit.x, scale * hit.y, EPSILON);
        } else {
            inside = v3(x, y, z);
        }
    }
    is->hit = hit;
    is->normal = n;
    is->inside = inside;
    is->outside = v3(hit.x + EPSILON * n.x, hit.y + EPSILON * n.y, hit.z + EPSILON * n.z);
    is->exiting = dot(d, n) >= 0.0;
}

// raysect/primitive/parabola.pyx:141-257.  Codes: 0 PARABOLA, 1 BASE.  The paraboloid z = height - k (x^2 + y^2),
// k = height / radius^2, tip at z = height, closed by the base disc at z = 0.
RSB_HD int parabola_crossings(const double* params, const V3& o, const V3& d, double max_distance, Crossing* out) {
    double radius = params[0], height = params[1];
    double k = height / (radius * radius);
    double a = k * (d.x * d.x + d.y * d.y);
    double b = 2 * k * (d.x * o.x + d.y * o.y) + d.z;
    double c = k * (o.x * o.x + o.y * o.y) - (height - o.z);
    double t0, t1;
    int t0_type, t1_type;
    if (!solve_quadratic(a, b, c, &t0, &t1)) return 0;
    if (t0 == t1) {
        // parabola.pyx:176-190, with the same `direction.y**2` placement as the cone (:184), reproduced verbatim
        t0 = -b / (2.0 * a);
        t0_type = 0;
        k = -o.z / d.z;
        double ex = o.x + k * d.x;
        double r2 = ex * ex + (o.y + k * (d.y * d.y));
        if (r2 <= (radius * radius)) { t1 = k; t1_type = 1; }
        else { t1 = t0; t1_type = t0_type; }
    } else {
        double t0_z = o.z + t0 * d.z;
        double t1_z = o.z + t1 * d.z;
        bool t0_outside = t0_z < 0;
        bool t1_outside = t1_z < 0;
        if (t0_outside && t1_outside) return 0;
        else if (!t0_outside && t1_outside) { t0_type = 0; t1 = -o.z / d.z; t1_type = 1; }
        else if (t0_outside && !t1_outside) { t0_type = 1; t0 = -o.z / d.z; t1_type = 0; }
        else { t0_type = 0; t1_type = 0; }
    }
    if (t0 > t1) {
        double tmp = t0; t0 = t1; t1 = tmp;
        int ti = t0_type; t0_type = t1_type; t1_type = ti;
    }
    if (t0 > max_distance || t1 < 0.0) return 0;
    if (t0 >= 0.0) {
        out[0].t = t0; out[0].code = t0_type;
        if (t1 <= max_distance) { out[1].t = t1; out[1].code = t1_type; return 2; }
        return 1;
    } else if (t1 <= max_distance) {
        out[0].t = t1; out[0].code = t1_type;
        return 1;
    }
    return 0;
}

// raysect/primitive/parabola.pyx:270-342 (_generate_intersection, _interior_point)
RSB_HD void parabola_geometry(const double* params, const V3& o, const V3& d, double t, int code, Isect* is) {
    const double EPSILON = 1e-9;
    double radius = params[0], height = params[1];
    V3 hit = v3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z);
    V3 n;
    if (code == 1) {
        n = v3(0, 0, -1);
    } else {
        double k = 2 * height / (radius * radius);
        n = normalise(v3(k * hit.x, k * hit.y, 1));
    }
    double x = hit.x, y = hit.y, z = hit.z;
    double inner_radius = radius - EPSILON;
    double hit_radius_sqr = hit.x * hit.x + hit.y * hit.y;
    if (hit_radius_sqr > (inner_radius * inner_radius)) {
        double scale = inner_radius / sqrt(hit_radius_sqr);
        x = scale * hit.x;
        y = scale * hit.y;
    }
    if (hit.z < EPSILON) {
        z = EPSILON;
    } else {
        x = hit.x - n.x * EPSILON;
        y = hit.y - n.y * EPSILON;
        z = hit.z - n.z * EPSILON;
    }
    is->hit = hit;
    is->normal = n;
    is->inside = v3(x, y, z);
    is->outside = v3(hit.x + EPSILON * n.x, hit.y + EPSILON * n.y, hit.z + EPSILON * n.z);
    is->exiting = dot(d, n) >= 0.0;
}

// raysect/primitive/torus.pyx:156-262 (hit).  params: major radius, minor radius.  The quartic's sorted real roots; the
// first one inside [0, max_distance] is the hit, the next one (if inside the range) is what next_intersection() would
// hand out.  Tori are world-level primitives here (not CSG operands: those need all four crossings).
RSB_HD int torus_crossings(const double* params, const V3& o, const V3& d, double max_distance, Crossing* out) {
    const double major = params[0], minor = params[1];
    double sq_origin_xy = o.x * o.x + o.y * o.y;
    double sq_direction_xy = d.x * d.x + d.y * d.y;
    double sq_origin = sq_origin_xy + o.z * o.z;
    double sq_direction = sq_direction_xy + d.z * d.z;
    double origin_direction_xy = o.x * d.x + o.y * d.y;
    double origin_dot_direction = origin_direction_xy + o.z * d.z;
    double sq_r = minor * minor;
    double sq_R = major * major;
    double R2_r2 = sq_R - sq_r;
    double a = sq_direction * sq_direction;
    double b = 4.0 * sq_direction * origin_dot_direction;
    double c = 2.0 * (2.0 * origin_dot_direction * origin_dot_direction + sq_direction * (sq_origin + R2_r2)) - 4.0 * sq_R * sq_direction_xy;
    double dd = 4.0 * origin_dot_direction * (sq_origin + R2_r2) - 8.0 * sq_R * origin_direction_xy;
    double e = (sq_origin + R2_r2) * (sq_origin + R2_r2) - 4.0 * sq_R * sq_origin_xy;
    double t[4];
    int num = solve_quartic(a, b, c, dd, e, &t[0], &t[1], &t[2], &t[3]);
    if (num == 0) return 0;
    if (num == 1) {
        if (t[0] > max_distance || t[0] < 0.0) return 0;
        out[0].t = t[0]; out[0].code = 0;
        return 1;
    }
    if (num == 2) {
        if (t[0] > t[1]) swap_dbl(&t[0], &t[1]);
        t[2] = t[1]; t[3] = t[1];
    } else if (num == 3) {
        sort_three_doubles(&t[0], &t[1], &t[2]);
        t[3] = t[2];
    } else if (num == 4) {
        sort_four_doubles(&t[0], &t[1], &t[2], &t[3]);
    } else {
        return 0;
    }
    if (t[0] > max_distance || t[3] < 0.0) return 0;
    for (int i = 0; i < num - 1; ++i) {
        if (t[i] >= 0.0) {
            out[0].t = t[i]; out[0].code = 0;
            if (t[i + 1] <= max_distance) { out[1].t = t[i + 1]; out[1].code = 0; return 2; }
            return 1;
        }
    }
    if (t[num - 1] <= max_distance) {
        out[0].t = t[num - 1]; out[0].code = 0;
        return 1;
    }
    return 0;
}

// torus.pyx:294-327 (_generate_intersection)
RSB_HD void torus_geometry(const double* params, const V3& o, const V3& d, double t, Isect* is) {
    const double EPSILON = 1e-9;
    V3 hit = v3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z);
    double alpha = params[0] / hypot(hit.x, hit.y);
    V3 n = normalise(v3((1.0 - alpha) * hit.x, (1.0 - alpha) * hit.y, hit.z));
    double delta_x = EPSILON * n.x, delta_y = EPSILON * n.y, delta_z = EPSILON * n.z;
    is->hit = hit;
    is->normal = n;
    is->inside = v3(hit.x - delta_x, hit.y - delta_y, hit.z - delta_z);
    is->outside = v3(hit.x + delta_x, hit.y + delta_y, hit.z + delta_z);
    is->exiting = dot(d, n) >= 0.0;
}

// TORUS: compile the torus (and its quartic solver) in.  Only the full-featured kernel instantiations do: in the lean ones the
// solver's code alone cost the Cornell trace kernel 1.5 % (0.223 -> 0.227 ms per wave) without a torus in sight.
template <bool TORUS = false>
RSB_HD int analytic_crossings(int type, const double* params, const V3& o, const V3& d, double max_distance, Crossing* out) {
    if (TORUS && type == PRIM_TORUS) return torus_crossings(params, o, d, max_distance, out);
    switch (type) {
        case PRIM_PARABOLA: return parabola_crossings(params, o, d, max_distance, out);
        case PRIM_SPHERE: return sphere_crossings(params, o, d, max_distance, out);
        case PRIM_BOX: return box_crossings(params, o, d, max_distance, out);
        case PRIM_CYLINDER: return cylinder_crossings(params, o, d, max_distance, out);
        case PRIM_CONE: return cone_crossings(params, o, d, max_distance, out);
        default: return 0;
    }
}

template <bool TORUS = false>
RSB_HD void analytic_geometry(int type, const double* params, const V3& o, const V3& d, double t, int code, Isect* is) {
    if (TORUS && type == PRIM_TORUS) { torus_geometry(params, o, d, t, is); return; }
    switch (type) {
        case PRIM_SPHERE: sphere_geometry(o, d, t, is); break;
        case PRIM_BOX: box_geometry(params, o, d, t, code, is); break;
        case PRIM_CYLINDER: cylinder_geometry(params, o, d, t, code, is); break;
        case PRIM_PARABOLA: parabola_geometry(params, o, d, t, code, is); break;
        default: cone_geometry(params, o, d, t, code, is); break;
    }
}

// point-in-primitive tests on a LOCAL-space point
// sphere.pyx:202-214, box.pyx:344-359, cylinder.pyx:351-367, cone.pyx:359-380
RSB_HD bool analytic_contains(int type, const double* params, const V3& p) {
    switch (type) {
        case PRIM_SPHERE: {
            double d2 = p.x * p.x + p.y * p.y + p.z * p.z;
            return d2 <= params[0] * params[0];
        }
        case PRIM_BOX: {
            if ((p.x < params[0]) || (p.x > params[3])) return false;
            if ((p.y < params[1]) || (p.y > params[4])) return false;
            if ((p.z < params[2]) || (p.z > params[5])) return false;
            return true;
        }
        case PRIM_CYLINDER: {
            bool slab = (0.0 <= p.z) && (p.z <= params[1]);
            return slab && ((p.x * p.x + p.y * p.y) <= (params[0] * params[0]));
        }
        case PRIM_TORUS: {      // torus.pyx:329-343
            double distance_xy = p.x * p.x + p.y * p.y;
            double distance_sqr = distance_xy + p.z * p.z;
            double sq_R = params[0] * params[0];
            double R2_r2 = sq_R - params[1] * params[1];
            double discriminant = distance_sqr * distance_sqr + 2.0 * distance_sqr * R2_r2 + R2_r2 * R2_r2 - 4.0 * sq_R * distance_xy;
            return discriminant <= 0.0;
        }
        case PRIM_PARABOLA: {   // parabola.pyx:344-363
            double radius = params[0], height = params[1];
            if (p.z < 0 || p.z > height) return false;
            double parabola_radius = radius * sqrt((height - p.z) / height);
            double point_radius = sqrt(p.x * p.x + p.y * p.y);
            return point_radius <= parabola_radius;
        }
        case PRIM_CONE: {
            double radius = params[0], height = params[1];
            if (p.z < 0 || p.z > height) return false;
            double pr2 = p.x * p.x + p.y * p.y;
            double cr = (height - p.z) * radius / height;
            cr *= cr;
            return pr2 <= cr;
        }
        default: return false;
    }
}

// ---------------------------------------------------------------------------------------------
// kd-tree traversal (raysect/core/math/spatial/kdtree3d.pyx:589-700), recursion unrolled onto an
// explicit far-child stack.  `leaf(item_offset, item_count, max_range)` returns true on a hit.
// When the near subtree finishes without a hit the far child resumes with
// min_range = plane_distance, which equals the max_range of the last leaf visited (front-to-back
// order), so only (node, max_range) is stacked.
// ---------------------------------------------------------------------------------------------
struct KdCursor {
    double min_range, max_range;
    int32_t node, sp;
};

// Per-ray storage the traversal indexes by split axis: element k of {origin xyz, direction xyz, 1/direction xyz}
// lives at p[k * S].  Selecting a component of a register-resident vector by a run-time axis costs a chain of
// predicated moves per 64-bit value (14 % of k_wf_trace's instructions in the round-1 profile); one indexed load
// replaces it.  On the device S = the CTA size and p points into shared memory at the thread's own column
// (bank = thread index, whatever the axis: conflict free); host code and the slow paths use S = 1 over a local
// array.  `unsafe` = recip_unsafe_mask: direction components whose plane distances must use a true division.
template <int S>
struct RayAx {
    double* p;
    int32_t unsafe;
    RSB_HD double o(int a) const { return p[a * S]; }
    RSB_HD double d(int a) const { return p[(3 + a) * S]; }
    RSB_HD double r(int a) const { return p[(6 + a) * S]; }
    RSB_HD V3 O() const { return v3(p[0], p[S], p[2 * S]); }
    RSB_HD V3 D() const { return v3(p[3 * S], p[4 * S], p[5 * S]); }
    RSB_HD V3 R() const { return v3(p[6 * S], p[7 * S], p[8 * S]); }
    // the reciprocals are the very quotients BoundingBox3D._slab forms (boundingbox.pyx:208-209)
    RSB_HD void set(double* storage, const V3& o, const V3& d) {
        p = storage;
        V3 rc = ray_reciprocals(d);
        p[0] = o.x; p[S] = o.y; p[2 * S] = o.z;
        p[3 * S] = d.x; p[4 * S] = d.y; p[5 * S] = d.z;
        p[6 * S] = rc.x; p[7 * S] = rc.y; p[8 * S] = rc.z;
        unsafe = recip_unsafe_mask(d, rc);
    }
};
#define RSB_AX_WORDS 18   // doubles of RayAx storage per thread: the world-space ray + one nested mesh-local ray

enum KdResult : int32_t { KD_MISS = 0, KD_HIT = 1, KD_MORE = 2 };

// KDTree3DCore._trace (kdtree3d.pyx:589-607): clip the ray against the tree bounds
template <int S>
RSB_HD bool kd_begin(const KdTree& tree, const RayAx<S>& ax, KdCursor& c) {
    c.node = 0;
    c.sp = 0;
    return box_intersect_inv(tree.bounds, ax.O(), ax.D(), ax.R(), &c.min_range, &c.max_range);
}

// One unit of traversal: descend from the cursor to the next leaf in front-to-back order
// (_trace_branch, kdtree3d.pyx:626-700) and run the leaf test (_trace_leaf).  "While-while" form: every
// lane of a warp first reaches its next leaf (cheap, uniform code) and only then are leaves processed,
// so the expensive item tests run with the warp converged.  KD_MORE: no hit in that leaf, the cursor
// points at the next subtree; callers loop.
// The branch step is written without control flow: the three cases of _trace_branch (near only / far only /
// both) become predicates, the plane distance (n.split - origin) / direction comes from the per-ray reciprocal
// (div_recip1: correctly rounded, 3 dependent instructions).  A zero direction component gives a NaN or
// infinite quotient that every comparison below rejects, which is the reference's `direction == 0` branch:
// go to the child on the origin's side.
// One 16-byte load per node visit (the address space -- shared for a staged world tree, global for mesh trees --
// is inferred by the compiler from the kernel's template flags)
RSB_HD KdNode kd_load_node(const KdNode* p) {
#ifdef __CUDA_ARCH__
    int4 v = *reinterpret_cast<const int4*>(p);
    KdNode n;
    n.split = __hiloint2double(v.y, v.x);
    n.upper = v.z;
    n.axis = v.w;
    return n;
#else
    return *p;
#endif
}

// _trace_branch (kdtree3d.pyx:626-700) from `node` down to the next leaf in front-to-back order; far children
// that must be visited later are stacked with the max_range they resume with.  Returns the leaf node.
template <int S, class Stats>
RSB_HD KdNode kd_descend(const KdTree& tree, const RayAx<S>& ax, KdStackEntry* stack, int& node, int& sp, double min_range,
                         double& max_range, Stats& stats) {
    const bool unsafe = ax.unsafe != 0;
    KdNode n = kd_load_node(tree.nodes + node);
    while (n.axis >= 0) {
        stats.branch();
        const int axis = n.axis;
        const double origin = ax.o(axis), direction = ax.d(axis);
#ifdef RSB_KD_TRUE_DIVIDE
        const double plane_distance = (n.split - origin) / direction;
#else
        const double plane_distance = div_recip1(n.split - origin, direction, ax.r(axis), unsafe);
#endif
        const bool below_split = origin < n.split || (origin == n.split && direction < 0);
        const int lower_id = node + 1, upper_id = n.upper;
        const int near_id = below_split ? lower_id : upper_id;
        const int far_id = below_split ? upper_id : lower_id;
        const bool only_near = direction == 0 || plane_distance > max_range || plane_distance <= 0;
        const bool only_far = !only_near && plane_distance < min_range;
        if (!only_near && !only_far) {
            stack[sp].node = far_id;
            stack[sp].tmax = max_range;
            ++sp;
            max_range = plane_distance;
        }
        node = only_far ? far_id : near_id;
        n = kd_load_node(tree.nodes + node);
    }
    return n;
}

template <int S, class LeafFn, class Stats>
RSB_HD int kd_advance(const KdTree& tree, const RayAx<S>& ax, KdStackEntry* stack, KdCursor& c, LeafFn& leaf,
                      Stats& stats, int* hit_node) {
    int node = c.node, sp = c.sp;
    double max_range = c.max_range;
    KdNode n = kd_descend(tree, ax, stack, node, sp, c.min_range, max_range, stats);
    stats.leaf(n.leaf.item_count);
    if (n.leaf.item_count > 0 && leaf(n.leaf.item_offset, n.leaf.item_count, max_range)) {
        *hit_node = node;
        return KD_HIT;
    }
    if (sp == 0) return KD_MISS;
    --sp;
    // the far child resumes with min_range = the plane distance = max_range of the leaf just left
    c.node = stack[sp].node;
    c.min_range = max_range;
    c.max_range = stack[sp].tmax;
    c.sp = sp;
    return KD_MORE;
}

template <int S, class LeafFn, class Stats>
RSB_HD bool kd_trace(const KdTree& tree, const RayAx<S>& ax, KdStackEntry* stack, LeafFn& leaf, Stats& stats, int* hit_node) {
    KdCursor c;
    if (!kd_begin(tree, ax, c)) return false;
    int r;
    do { r = kd_advance(tree, ax, stack, c, leaf, stats, hit_node); } while (r == KD_MORE);
    return r == KD_HIT;
}

// raysect/core/math/spatial/kdtree3d.pyx:736-792 (_items_containing*): descend to the one leaf
// holding the point.  Returns false when the point is outside the tree bounds.
template <class Stats>
RSB_HD bool kd_locate(const KdTree& tree, const V3& p, int* item_offset, int* item_count, Stats& stats) {
    stats.contains_query();
    if (!box_contains(tree.bounds, p)) return false;
    int node = 0;
    for (;;) {
        KdNode n = tree.nodes[node];
        if (n.axis < 0) {
            stats.leaf(n.leaf.item_count);
            *item_offset = n.leaf.item_offset;
            *item_count = n.leaf.item_count;
            return true;
        }
        stats.branch();
        node = (v3_get(p, n.axis) < n.split) ? node + 1 : n.upper;
    }
}

struct NoStats {
    RSB_HD void branch() {}
    RSB_HD void leaf(int) {}
    RSB_HD void prim_test() {}
    RSB_HD void tri_test() {}
    RSB_HD void contains_query() {}
    RSB_HD void table_read() {}
};

struct CountStats {
    unsigned long long branches = 0, leaves = 0, items = 0, prim_tests = 0, tri_tests = 0, contains = 0, tables = 0;
    RSB_HD void contains_query() { ++contains; }
    RSB_HD void table_read() { ++tables; }
    RSB_HD void branch() { ++branches; }
    RSB_HD void leaf(int n) { ++leaves; items += (unsigned long long)n; }
    RSB_HD void prim_test() { ++prim_tests; }
    RSB_HD void tri_test() { ++tri_tests; }
};

// ---------------------------------------------------------------------------------------------
// Triangle mesh (raysect/primitive/mesh/mesh.pyx:506-713): Woop-Benthin-Wald watertight test in
// float32 with the reference's exact double promotions.
// ---------------------------------------------------------------------------------------------
struct RaySpace {
    int ix, iy, iz;
    float sx, sy, sz;
};

// mesh.pyx:566-610 (_calc_rayspace_transform)
RSB_HD RaySpace mesh_rayspace(const V3& d) {
    RaySpace rs;
    int ix, iy, iz;
    if (fabs(d.x) > fabs(d.y) && fabs(d.x) > fabs(d.z)) { ix = 1; iy = 2; iz = 0; }
    else if (fabs(d.y) > fabs(d.x) && fabs(d.y) > fabs(d.z)) { ix = 2; iy = 0; iz = 1; }
    else { ix = 0; iy = 1; iz = 2; }
    float rdz = (float)v3_get(d, iz);
    if (rdz < 0.0f) { int tmp = ix; ix = iy; iy = tmp; }
    rs.sz = (float)(1.0 / (double)rdz);
    rs.sx = (float)(v3_get(d, ix) * (double)rs.sz);
    rs.sy = (float)(v3_get(d, iy) * (double)rs.sz);
    rs.ix = ix; rs.iy = iy; rs.iz = iz;
    return rs;
}

RSB_HD float f3_get(const float* v, int axis) { return axis == 0 ? v[0] : (axis == 1 ? v[1] : v[2]); }

// mesh.pyx:616-713 (_hit_triangle).  out = (u, v, w, t) as float32.
RSB_HD bool mesh_hit_triangle(const F4* tri, const V3& o, double max_distance, const RaySpace& rs, float* out) {
    F4 q0 = tri[0], q1 = tri[1], q2 = tri[2];
    // centre on the ray origin: float vertex minus double origin, rounded back to float (mesh.pyx:639-649)
    float v1[3], v2[3], v3_[3];
    v1[0] = (float)((double)q0.x - o.x); v1[1] = (float)((double)q0.y - o.y); v1[2] = (float)((double)q0.z - o.z);
    v2[0] = (float)((double)q0.w - o.x); v2[1] = (float)((double)q1.x - o.y); v2[2] = (float)((double)q1.y - o.z);
    v3_[0] = (float)((double)q1.z - o.x); v3_[1] = (float)((double)q1.w - o.y); v3_[2] = (float)((double)q2.x - o.z);
    float sx = rs.sx, sy = rs.sy, sz = rs.sz;
    float v1z = f3_get(v1, rs.iz), v2z = f3_get(v2, rs.iz), v3z = f3_get(v3_, rs.iz);
    float x1 = f3_get(v1, rs.ix) - sx * v1z;
    float x2 = f3_get(v2, rs.ix) - sx * v2z;
    float x3 = f3_get(v3_, rs.ix) - sx * v3z;
    float y1 = f3_get(v1, rs.iy) - sy * v1z;
    float y2 = f3_get(v2, rs.iy) - sy * v2z;
    float y3 = f3_get(v3_, rs.iy) - sy * v3z;
    float u = x3 * y2 - y3 * x2;
    float v = x1 * y3 - y1 * x3;
    float w = x2 * y1 - y2 * x1;
    if (u == 0.0f || v == 0.0f || w == 0.0f) {
        u = (float)((double)x3 * (double)y2 - (double)y3 * (double)x2);
        v = (float)((double)x1 * (double)y3 - (double)y1 * (double)x3);
        w = (float)((double)x2 * (double)y1 - (double)y2 * (double)x1);
    }
    if ((u < 0.0f || v < 0.0f || w < 0.0f) && (u > 0.0f || v > 0.0f || w > 0.0f)) return false;
    float det = u + v + w;
    if (det == 0.0f) return false;
    float z1 = sz * v1z;
    float z2 = sz * v2z;
    float z3 = sz * v3z;
    float t = u * z1 + v * z2 + w * z3;
    if (det > 0.0f) {
        if (t < 0.0f || (double)t > max_distance * (double)det) return false;
    } else {
        if (t > 0.0f || (double)t < max_distance * (double)det) return false;
    }
    // mesh.pyx:705: the reciprocal is formed in double and rounded to float.  A double holds more than 2 * 24 + 2
    // significand bits, so that double rounding is the correctly rounded float quotient (Figueroa 1995) as long as the
    // quotient is a normal float: on the device one IEEE float division replaces the ~40-instruction double one.
#ifdef __CUDA_ARCH__
    const float adet = fabsf(det);
    float det_reciprocal = (adet > 1e-30f && adet < 1e30f) ? __fdiv_rn(1.0f, det) : (float)(1.0 / (double)det);
#else
    float det_reciprocal = (float)(1.0 / (double)det);
#endif
    out[0] = u * det_reciprocal;
    out[1] = v * det_reciprocal;
    out[2] = w * det_reciprocal;
    out[3] = t * det_reciprocal;
    return true;
}

struct MeshHit {
    double t;       // == (double)(float)t, mesh.pyx:557-561
    int32_t tri;
    int32_t node;   // mesh kd-tree leaf that produced the hit
    float u, v, w;
};

template <class Stats>
struct MeshLeaf {
    const Mesh* mesh;
    V3 o;
    double max_distance;
    RaySpace rs;
    MeshHit* result;
    Stats* stats;
    // mesh.pyx:520-563 (_trace_leaf): strict `t < distance`, first of equal-t triangles wins
    RSB_HD bool operator()(int offset, int count, double max_range) {
        double distance = max_distance < max_range ? max_distance : max_range;   // min(ray.max_distance, max_range)
        int closest = -1;
        float cu = 0, cv = 0, cw = 0;
        for (int i = 0; i < count; ++i) {
            int tri = mesh->tree.items[offset + i];
            float h[4];
            stats->tri_test();
            if (mesh_hit_triangle(mesh->tri + 3 * (size_t)tri, o, max_distance, rs, h)) {
                double t = (double)h[3];
                if (t < distance) {
                    distance = t;
                    closest = tri;
                    cu = h[0]; cv = h[1]; cw = h[2];
                }
            }
        }
        if (closest < 0) return false;
        result->t = (double)(float)distance;
        result->tri = closest;
        result->u = cu; result->v = cv; result->w = cw;
        return true;
    }
};

// mesh.pyx:506-518 (MeshData.trace) on a mesh-local ray; `axbuf` = 9*S doubles of RayAx storage
template <int S, class Stats>
RSB_HD bool mesh_trace(const Mesh& mesh, const V3& o, const V3& d, double max_distance, KdStackEntry* stack, MeshHit* out, Stats& stats,
                       double* axbuf) {
    RayAx<S> ax;
    ax.set(axbuf, o, d);
    MeshLeaf<Stats> leaf;
    leaf.mesh = &mesh;
    leaf.o = o;
    leaf.max_distance = max_distance;
    leaf.rs = mesh_rayspace(d);
    leaf.result = out;
    leaf.stats = &stats;
    return kd_trace(mesh.tree, ax, stack, leaf, stats, &out->node);
}

// mesh.pyx:718-800 (calc_intersection, _intersection_normal) in mesh-local space
RSB_HD void mesh_geometry(const Mesh& mesh, const V3& o, const V3& d, const MeshHit& h, Isect* is) {
    const double EPSILON = 1e-6;
    F4 q2 = mesh.tri[3 * (size_t)h.tri + 2];
    V3 fn = v3((double)q2.y, (double)q2.z, (double)q2.w);
    double t = h.t;
    V3 hit = v3(o.x + d.x * t, o.y + d.y * t, o.z + d.z * t);
    is->hit = hit;
    is->inside = v3(hit.x - fn.x * EPSILON, hit.y - fn.y * EPSILON, hit.z - fn.z * EPSILON);
    is->outside = v3(hit.x + fn.x * EPSILON, hit.y + fn.y * EPSILON, hit.z + fn.z * EPSILON);
    V3 n;
    if (mesh.smoothing && mesh.vnormals != nullptr) {
        const int32_t* row = mesh.tri_idx + (size_t)h.tri * mesh.idx_stride;
        const float* n1 = mesh.vnormals + 3 * (size_t)row[3];
        const float* n2 = mesh.vnormals + 3 * (size_t)row[4];
        const float* n3 = mesh.vnormals + 3 * (size_t)row[5];
        // float32 arithmetic, left to right (mesh.pyx:788-792)
        float nx = h.u * n1[0] + h.v * n2[0] + h.w * n3[0];
        float ny = h.u * n1[1] + h.v * n2[1] + h.w * n3[1];
        float nz = h.u * n1[2] + h.v * n2[2] + h.w * n3[2];
        n = normalise(v3((double)nx, (double)ny, (double)nz));
    } else {
        n = normalise(fn);
    }
    is->normal = n;
    is->exiting = dot(d, fn) > 0.0;   // strict, mesh.pyx:756
    is->code = h.tri;
    is->u = h.u; is->v = h.v; is->w = h.w;
}

// mesh.pyx:805-830 (MeshData.contains): +z ray, orientation of the first face hit
template <class Stats>
RSB_HD bool mesh_contains(const Mesh& mesh, const V3& p, KdStackEntry* stack, Stats& stats) {
    MeshHit h;
    double axbuf[9];
    if (!mesh_trace<1>(mesh, p, v3(0, 0, 1), RSB_INF, stack, &h, stats, axbuf)) return false;
    return mesh.tri[3 * (size_t)h.tri + 2].w > 0.0f;
}

// ---------------------------------------------------------------------------------------------
// CSG (raysect/primitive/csg.pyx:132-241 + operator rules :326-348, :424-446, :526-568).
// The reference pulls crossings lazily from two stateful children; all analytic leaves are convex
// (<= 2 crossings), so each node's full, ordered crossing list is materialised bottom-up and the
// operator automaton is run over the two child lists.  The emitted sequence is exactly what
// hit() + repeated next_intersection() would yield.
// ---------------------------------------------------------------------------------------------
struct CsgEvent {
    double t;
    int32_t leaf;       // analytic leaf row
    int16_t code;
    int8_t exiting;
    int8_t flip;
};

RSB_HD bool csg_valid(int op, bool inside_a, bool inside_b, bool closest_is_a) {
    if (op == PRIM_UNION) {
        if (!inside_a && !inside_b) return true;
        else if (inside_a && !inside_b && closest_is_a) return true;
        else if (!inside_a && inside_b && !closest_is_a) return true;
        return false;
    } else if (op == PRIM_INTERSECT) {
        if (inside_a && inside_b) return true;
        else if (inside_a && !inside_b && !closest_is_a) return true;
        else if (!inside_a && inside_b && closest_is_a) return true;
        return false;
    } else {
        if (!inside_a && !inside_b && closest_is_a) return true;
        else if (inside_a && !inside_b) return true;
        else if (inside_a && inside_b && !closest_is_a) return true;
        return false;
    }
}

// Crossings of primitive row `id` for a ray (o, d) given in the row's PARENT space, behind the
// BoundPrimitive AABB pre-test (raysect/core/acceleration/boundprimitive.pyx:42-51).
// Children always see max_distance = +inf (csg.pyx:142-144).
template <int DEPTH>
struct CsgEval {
    static RSB_HD_NOINLINE int run(const Scene& sc, int id, const V3& o, const V3& d, CsgEvent* out) {
        const Prim& p = sc.prims[id];
        if (!box_hit(p.bbox, o, d)) return 0;
        V3 lo = xform_point(p.to_local, o);
        V3 ld = xform_vector(p.to_local, d);
        if (p.type <= PRIM_CONE) {
            Crossing c[2];
            int n = analytic_crossings(p.type, p.params, lo, ld, RSB_INF, c);
            for (int i = 0; i < n; ++i) {
                Isect is;
                analytic_geometry(p.type, p.params, lo, ld, c[i].t, c[i].code, &is);
                out[i].t = c[i].t;
                out[i].leaf = id;
                out[i].code = (int16_t)c[i].code;
                out[i].exiting = (int8_t)is.exiting;
                out[i].flip = 0;
            }
            return n;
        }
        if (p.type == PRIM_MESH) return 0;   // rejected at scene creation
        CsgEvent a[RSB_CSG_MAX_EVENTS], b[RSB_CSG_MAX_EVENTS];
        int na = CsgEval<DEPTH - 1>::run(sc, p.child_a, lo, ld, a);
        if (na == 0 && p.type != PRIM_UNION) return 0;   // terminate_early (csg.pyx:421-422, 523-524)
        int nb = CsgEval<DEPTH - 1>::run(sc, p.child_b, lo, ld, b);
        int ia = 0, ib = 0, n = 0;
        while (ia < na || ib < nb) {
            bool has_a = ia < na, has_b = ib < nb;
            bool closest_is_a = has_a && (!has_b || a[ia].t < b[ib].t);   // _closest_intersection, csg.pyx:226-234
            bool inside_a = has_a && a[ia].exiting;
            bool inside_b = has_b && b[ib].exiting;
            if (csg_valid(p.type, inside_a, inside_b, closest_is_a) && n < RSB_CSG_MAX_EVENTS) {
                CsgEvent e = closest_is_a ? a[ia] : b[ib];
                if (p.type == PRIM_SUBTRACT && !closest_is_a) {   // _modify_intersection, csg.pyx:550-568
                    e.exiting = !e.exiting;
                    e.flip ^= 1;
                }
                out[n++] = e;
            }
            if (closest_is_a) ++ia; else ++ib;
        }
        return n;
    }
};

template <>
struct CsgEval<0> {
    static RSB_HD int run(const Scene& sc, int id, const V3& o, const V3& d, CsgEvent* out) {
        const Prim& p = sc.prims[id];
        if (p.type > PRIM_CONE) return 0;   // nesting deeper than RSB_CSG_MAX_DEPTH is rejected at scene creation
        if (!box_hit(p.bbox, o, d)) return 0;
        V3 lo = xform_point(p.to_local, o);
        V3 ld = xform_vector(p.to_local, d);
        Crossing c[2];
        int n = analytic_crossings(p.type, p.params, lo, ld, RSB_INF, c);
        for (int i = 0; i < n; ++i) {
            Isect is;
            analytic_geometry(p.type, p.params, lo, ld, c[i].t, c[i].code, &is);
            out[i].t = c[i].t;
            out[i].leaf = id;
            out[i].code = (int16_t)c[i].code;
            out[i].exiting = (int8_t)is.exiting;
            out[i].flip = 0;
        }
        return n;
    }
};

// CSGPrimitive.hit for a world-level CSG row: first valid crossing, provided it lies within
// ray.max_distance (csg.pyx:181-212); the AABB pre-test was done by the caller.
RSB_HD_NOINLINE bool csg_first_hit(const Scene& sc, int id, const V3& o, const V3& d, double max_distance, CsgEvent* ev) {
    const Prim& p = sc.prims[id];
    V3 lo = xform_point(p.to_local, o);
    V3 ld = xform_vector(p.to_local, d);
    CsgEvent a[RSB_CSG_MAX_EVENTS], b[RSB_CSG_MAX_EVENTS];
    int na = CsgEval<RSB_CSG_MAX_DEPTH - 1>::run(sc, p.child_a, lo, ld, a);
    if (na == 0 && p.type != PRIM_UNION) return false;
    int nb = CsgEval<RSB_CSG_MAX_DEPTH - 1>::run(sc, p.child_b, lo, ld, b);
    int ia = 0, ib = 0;
    while (ia < na || ib < nb) {
        bool has_a = ia < na, has_b = ib < nb;
        bool closest_is_a = has_a && (!has_b || a[ia].t < b[ib].t);
        bool inside_a = has_a && a[ia].exiting;
        bool inside_b = has_b && b[ib].exiting;
        if (csg_valid(p.type, inside_a, inside_b, closest_is_a)) {
            CsgEvent e = closest_is_a ? a[ia] : b[ib];
            if (!(e.t <= max_distance)) return false;
            if (p.type == PRIM_SUBTRACT && !closest_is_a) {
                e.exiting = !e.exiting;
                e.flip ^= 1;
            }
            *ev = e;
            return true;
        }
        if (closest_is_a) ++ia; else ++ib;
    }
    return false;
}

// CSG contains (csg.pyx:350-352, 448-450, 570-572) behind BoundPrimitive.contains; p in parent space
template <int DEPTH>
struct CsgContains {
    static RSB_HD_NOINLINE bool run(const Scene& sc, int id, const V3& pt) {
        const Prim& p = sc.prims[id];
        if (!box_contains(p.bbox, pt)) return false;
        V3 lp = xform_point(p.to_local, pt);
        if (p.type <= PRIM_CONE) return analytic_contains(p.type, p.params, lp);
        if (p.type == PRIM_MESH) return false;
        bool ca = CsgContains<DEPTH - 1>::run(sc, p.child_a, lp);
        if (p.type == PRIM_UNION) return ca || CsgContains<DEPTH - 1>::run(sc, p.child_b, lp);
        if (p.type == PRIM_INTERSECT) return ca && CsgContains<DEPTH - 1>::run(sc, p.child_b, lp);
        return ca && !CsgContains<DEPTH - 1>::run(sc, p.child_b, lp);
    }
};
template <>
struct CsgContains<0> {
    static RSB_HD bool run(const Scene& sc, int id, const V3& pt) {
        const Prim& p = sc.prims[id];
        if (p.type > PRIM_CONE) return false;
        if (!box_contains(p.bbox, pt)) return false;
        return analytic_contains(p.type, p.params, xform_point(p.to_local, pt));
    }
};

// Geometry of a CSG crossing, re-expressed level by level up to the world-level CSG primitive's
// local space exactly as _identify_intersection does (csg.pyx:200-208): points by each child's
// to_root, the normal by the inverse-transpose of that same matrix.
RSB_HD_NOINLINE void csg_geometry(const Scene& sc, int top, const CsgEvent& ev, const V3& o, const V3& d, Isect* is) {
    // chain of rows from the leaf up to (excluding) the world-level CSG primitive
    int chain[RSB_CSG_MAX_DEPTH + 1];
    int n = 0;
    for (int id = ev.leaf; id != top && n <= RSB_CSG_MAX_DEPTH; id = sc.prims[id].parent) chain[n++] = id;
    // ray into the leaf's local space, through every level top-down
    V3 lo = xform_point(sc.prims[top].to_local, o);
    V3 ld = xform_vector(sc.prims[top].to_local, d);
    for (int i = n - 1; i >= 0; --i) {
        const Prim& p = sc.prims[chain[i]];
        lo = xform_point(p.to_local, lo);
        ld = xform_vector(p.to_local, ld);
    }
    const Prim& leaf = sc.prims[ev.leaf];
    analytic_geometry(leaf.type, leaf.params, lo, ld, ev.t, ev.code, is);
    for (int i = 0; i < n; ++i) {
        const Prim& p = sc.prims[chain[i]];
        is->hit = xform_point(p.to_root, is->hit);
        is->inside = xform_point(p.to_root, is->inside);
        is->outside = xform_point(p.to_root, is->outside);
        is->normal = xform_normal_with_inverse(p.root_inv, is->normal);
    }
    if (ev.flip) {
        V3 tmp = is->inside; is->inside = is->outside; is->outside = tmp;
        is->normal = v3(-is->normal.x, -is->normal.y, -is->normal.z);
    }
    is->exiting = ev.exiting;
    is->code = ev.code;
}

// ---------------------------------------------------------------------------------------------
// World.hit  (raysect/core/scenegraph/world.pyx:125-146 -> acceleration/kdtree.pyx:73-175)
// ---------------------------------------------------------------------------------------------
// FEAT: scene features the instantiation must handle (bit 0: Mesh primitives, bit 1: CSG primitives).  The
// wavefront kernels are instantiated per feature set, so that a scene of analytic primitives does not pay the
// registers and instruction footprint of the nested mesh traversal and the CSG evaluator.
#define RSB_FEAT_MESH 1
#define RSB_FEAT_CSG 2
#define RSB_FEAT_ALL 3
// Conductor and the volume emitters ride in the full-featured instantiation (the one that also carries the CSG
// evaluator): scenes that use them are dispatched to RSB_FEAT_ALL, and the lean analytic-only / mesh-only kernels
// do not pay their registers (measured: 3 % of the Cornell bench when compiled in unconditionally).
#define RSB_FEAT_RARE_MATERIALS RSB_FEAT_CSG
#define RSB_FEAT_STAGED 4   // kernels only: world tree, item list and primitive table are in shared memory

// AABB pre-test of world leaf item k (BoundPrimitive.hit, boundprimitive.pyx:42-51): id of the primitive, true when its box is hit
template <int FEAT>
RSB_HD bool world_item_pretest(const Scene& sc, int k, const V3& o, const V3& d, const V3& inv, int* id) {
    if (!(FEAT & RSB_FEAT_STAGED) && sc.world_rows != nullptr) {
        const LeafRow& r = sc.world_rows[k];
        *id = r.id;
        return box_hit_inv(r.bbox, o, d, inv);
    }
    *id = sc.world.items[k];
    return box_hit_inv(sc.prims[*id].bbox, o, d, inv);
}

template <class Stats, int FEAT = RSB_FEAT_ALL, int S = 1>
struct WorldLeaf {
    const Scene* sc;
    RayAx<S> ax;                // world-space ray: origin, direction, 1.0 / direction (shared by every AABB test)
    double max_distance;
    KdStackEntry* mesh_stack;   // stack space above the world traversal's own entries
    double* mesh_axbuf;         // RayAx storage of a nested mesh traversal
    HitRec* best;
    Stats* stats;

    // one candidate primitive that passed its AABB pre-test (BoundPrimitive.hit, boundprimitive.pyx:42-51)
    RSB_HD void test(int id, double& distance, bool& found) {
        const Prim& p = sc->prims[id];
        const V3 o = ax.O(), d = ax.D();
        if (p.type <= PRIM_CONE) {
            V3 lo = xform_point(p.to_local, o);
            V3 ld = xform_vector(p.to_local, d);
            Crossing c[2];
            if (analytic_crossings<(FEAT & RSB_FEAT_CSG) != 0>(p.type, p.params, lo, ld, max_distance, c) > 0 && c[0].t <= distance) {
                distance = c[0].t;
                best->t = c[0].t; best->prim = id; best->leaf = id; best->code = c[0].code; best->flip = 0;
                best->mesh_node = -1;
                found = true;
            }
        } else if ((FEAT & RSB_FEAT_MESH) && p.type == PRIM_MESH) {
            V3 lo = xform_point(p.to_local, o);
            V3 ld = xform_vector(p.to_local, d);
            MeshHit mh;
            if (mesh_trace<S>(sc->meshes[p.mesh], lo, ld, max_distance, mesh_stack, &mh, *stats, mesh_axbuf) && mh.t <= distance) {
                distance = mh.t;
                best->t = mh.t; best->prim = id; best->leaf = id; best->code = mh.tri; best->flip = 0;
                best->mesh_node = mh.node;
                best->u = mh.u; best->v = mh.v; best->w = mh.w;
                found = true;
            }
        } else if ((FEAT & RSB_FEAT_CSG) && p.type >= PRIM_UNION) {
            CsgEvent ev;
            if (csg_first_hit(*sc, id, o, d, max_distance, &ev) && ev.t <= distance) {
                distance = ev.t;
                best->t = ev.t; best->prim = id; best->leaf = ev.leaf; best->code = ev.code;
                best->flip = ev.flip | (ev.exiting << 1);
                best->mesh_node = -1;
                found = true;
            }
        }
    }

    // _PrimitiveKDTree._trace_leaf (acceleration/kdtree.pyx:73-122): `<=` => last of equal-t items wins.
    // Two phases per chunk of items: the cheap AABB pre-tests first, collecting the survivors (in item
    // order, which the tie rule depends on), then the primitive tests -- so lanes whose items were all
    // rejected do not sit through other lanes' quadratic solves one item at a time.
    RSB_HD bool operator()(int offset, int count, double max_range) {
        double distance = max_distance < max_range ? max_distance : max_range;
        bool found = false;
        for (int base = 0; base < count; base += 4) {
            int cand[4];
            int nc = 0;
            int end = count - base < 4 ? count - base : 4;
            {
                const V3 o = ax.O(), d = ax.D(), inv = ax.R();
                for (int i = 0; i < end; ++i) {
                    int id;
                    stats->prim_test();
                    if (world_item_pretest<FEAT>(*sc, offset + base + i, o, d, inv, &id)) cand[nc++] = id;
                }
            }
            for (int i = 0; i < nc; ++i) test(cand[i], distance, found);
        }
        return found;
    }
};

// World.hit for scenes that contain meshes: ONE loop over both levels of the hierarchy.
//
// The reference descends the world tree, and wherever a leaf lists a Mesh it runs that mesh's whole kd traversal
// from inside the leaf test (Mesh.hit -> MeshData.trace).  Done literally on a GPU, every lane of a warp reaches
// its mesh candidate at a different point of the world-level control flow (another leaf, another item of the
// leaf), so the long nested traversals run one or two lanes at a time: the round-1 profile of the 1.3 M-triangle
// sweep shows 3.2 of 32 lanes active per instruction, 1.8 in the triangle tests.  Here a lane that meets a mesh
// candidate SUSPENDS its world-level leaf (state: leaf, chunk, candidate index, running closest distance) and
// switches to the MESH state; the warp then advances all lanes that are inside a mesh by one mesh traversal unit
// per trip, whatever world-level path brought them there.  Per ray the sequence of AABB tests, primitive tests,
// mesh leaf visits and triangle tests -- and therefore every comparison and tie rule -- is exactly that of
// world_hit_ax / _PrimitiveKDTree._trace_leaf; only the interleaving BETWEEN rays changes.
// The traversal is a resumable object so that kernels can interleave it with lane refill (k_hit_sweep):
//   begin(o, d)  world-level work up to the first mesh unit; false: the ray is finished already
//   step()       ONE mesh traversal unit (descend to the next mesh leaf + its triangle tests), then the world-level
//                work that follows it; false: the ray is finished
//   finish()     closest hit found?
template <int FEAT, int S, class Stats>
struct NestedTraversal {
    enum { ST_WORLD = 0, ST_MESH = 1, ST_DONE = 2 };
    const Scene* sc;
    Stats* stats;
    KdStackEntry* stack;
    HitRec* rec;
    WorldLeaf<Stats, FEAT, S> leaf;
    int node, sp, leaf_node;
    double min_range, max_range;
    int item_offset, item_count, item_base;   // the world leaf being processed, and how far
    int cand[4];
    int nc, ci;
    bool have_leaf, found;
    double distance;
    int state;
    // nested mesh traversal of candidate cand[ci]
    RayAx<S> max;
    KdCursor mc;
    MeshLeaf<Stats> mleaf;
    MeshHit mh;
    // split pipeline (rsb_trav.cuh): the last Mesh.hit answer of this ray.  Mesh.hit(ray) depends on the ray and its
    // max_distance only (mesh.pyx:506-518), and the reference calls it again from every world leaf that lists the mesh
    // until a hit is accepted; the repeated calls return the very same intersection, so it is kept instead.
    int memo_prim;
    bool memo_hit;
    MeshHit memo;

    RSB_HD void init(const Scene& scene, double max_distance, KdStackEntry* stack_, HitRec* rec_, Stats& stats_, double* axbuf) {
        sc = &scene;
        stats = &stats_;
        stack = stack_;
        rec = rec_;
        leaf.sc = &scene;
        leaf.ax.p = axbuf;
        leaf.ax.unsafe = 0;
        leaf.max_distance = max_distance;
        leaf.mesh_stack = stack_ + (RSB_KD_STACK / 2);
        leaf.mesh_axbuf = axbuf + 9 * S;
        leaf.best = rec_;
        leaf.stats = &stats_;
        max.p = leaf.mesh_axbuf;
        max.unsafe = 0;
        mc.node = 0; mc.sp = 0; mc.min_range = 0; mc.max_range = 0;
        mleaf.mesh = nullptr;
        mleaf.max_distance = max_distance;
        mleaf.result = &mh;
        mleaf.stats = &stats_;
        mh.node = -1;
        state = ST_DONE;
        found = false;
        memo_prim = -1;
        memo_hit = false;
        memo.t = 0.0; memo.tri = -1; memo.node = -1; memo.u = memo.v = memo.w = 0.0f;
        node = sp = 0; leaf_node = -1;
        min_range = max_range = distance = 0.0;
        item_offset = item_count = item_base = nc = ci = 0;
        have_leaf = false;
    }

    // Mesh.hit's answer for candidate cand[ci] against the running closest distance of the world leaf
    // (_PrimitiveKDTree._trace_leaf, acceleration/kdtree.pyx:103-122: `<=`)
    RSB_HD void accept_mesh(bool hit, const MeshHit& m) {
        if (hit && m.t <= distance) {
            const int id = cand[ci];
            distance = m.t;
            rec->t = m.t; rec->prim = id; rec->leaf = id; rec->code = m.tri; rec->flip = 0;
            rec->mesh_node = m.node;
            rec->u = m.u; rec->v = m.v; rec->w = m.w;
            found = true;
        }
    }

    RSB_HD void run_world() { run_world_t<false>(); }

    // SPLIT: the world-level walk stops in front of every Mesh.hit call (state ST_MESH, nothing of the mesh touched yet);
    // a separate kernel answers it and resume_split() carries on.
    template <bool SPLIT>
    RSB_HD void run_world_t() {
        while (state == ST_WORLD) {
            if (ci < nc) {
                const int id = cand[ci];
                const Prim& p = sc->prims[id];
                if (SPLIT && p.type == PRIM_MESH) {
                    if (id == memo_prim) { accept_mesh(memo_hit, memo); ++ci; }
                    else state = ST_MESH;
                } else if (p.type == PRIM_MESH) {
                    const V3 wo = leaf.ax.O(), wd = leaf.ax.D();
                    const V3 lo = xform_point(p.to_local, wo);
                    const V3 ld = xform_vector(p.to_local, wd);
                    max.set(leaf.mesh_axbuf, lo, ld);
                    mleaf.mesh = &sc->meshes[p.mesh];
                    mleaf.o = lo;
                    mleaf.rs = mesh_rayspace(ld);
                    if (kd_begin(mleaf.mesh->tree, max, mc)) state = ST_MESH;   // MeshData.trace, mesh.pyx:506-518
                    else ++ci;
                } else {
                    leaf.test(id, distance, found);
                    ++ci;
                }
            } else if (item_base < item_count) {
                // next chunk of (up to) four items: AABB pre-tests, survivors in item order
                const V3 ro = leaf.ax.O(), rd = leaf.ax.D(), inv = leaf.ax.R();
                const int end = item_count - item_base < 4 ? item_count - item_base : 4;
                nc = 0;
                ci = 0;
                for (int i = 0; i < end; ++i) {
                    int id;
                    stats->prim_test();
                    if (world_item_pretest<FEAT>(*sc, item_offset + item_base + i, ro, rd, inv, &id)) cand[nc++] = id;
                }
                item_base += end;
            } else if (have_leaf) {
                // leaf finished: report, or resume at the nearest stacked far child with min_range = the plane
                // distance = max_range of the leaf just left
                have_leaf = false;
                if (found || sp == 0) state = ST_DONE;
                else { --sp; node = stack[sp].node; min_range = max_range; max_range = stack[sp].tmax; }
            } else {
                KdNode n = kd_descend(sc->world, leaf.ax, stack, node, sp, min_range, max_range, *stats);
                stats->leaf(n.leaf.item_count);
                leaf_node = node;
                item_offset = n.leaf.item_offset;
                item_count = n.leaf.item_count;
                item_base = 0;
                nc = 0;
                ci = 0;
                distance = leaf.max_distance < max_range ? leaf.max_distance : max_range;
                found = false;
                have_leaf = true;
            }
        }
    }

    RSB_HD bool begin(const V3& o, const V3& d) { return begin_t<false>(o, d); }

    template <bool SPLIT>
    RSB_HD bool begin_t(const V3& o, const V3& d) {
        leaf.ax.set(leaf.ax.p, o, d);
        memo_prim = -1;
        rec->u = rec->v = rec->w = 0.0f;
        rec->node = -1;
        rec->mesh_node = -1;
        found = false;
        have_leaf = false;
        nc = ci = item_base = item_count = 0;
        node = 0;
        sp = 0;
        KdCursor c;
        if (!kd_begin(sc->world, leaf.ax, c)) { state = ST_DONE; return false; }
        min_range = c.min_range;
        max_range = c.max_range;
        state = ST_WORLD;
        run_world_t<SPLIT>();
        return state != ST_DONE;
    }

    // split pipeline: Mesh.hit of candidate cand[ci] was answered elsewhere
    RSB_HD bool resume_split(bool hit, const MeshHit& m) {
        memo_prim = cand[ci];
        memo_hit = hit;
        memo = m;
        accept_mesh(hit, m);
        ++ci;
        state = ST_WORLD;
        run_world_t<true>();
        return state != ST_DONE;
    }

    // a walk suspended by the split pipeline in front of Mesh.hit, carried on by the nested loop (step()): MeshData.trace's
    // preamble for candidate cand[ci] (mesh.pyx:506-518)
    RSB_HD bool enter_mesh() {
        const Prim& p = sc->prims[cand[ci]];
        const V3 wo = leaf.ax.O(), wd = leaf.ax.D();
        const V3 lo = xform_point(p.to_local, wo);
        const V3 ld = xform_vector(p.to_local, wd);
        max.set(leaf.mesh_axbuf, lo, ld);
        mleaf.mesh = &sc->meshes[p.mesh];
        mleaf.o = lo;
        mleaf.rs = mesh_rayspace(ld);
        if (kd_begin(mleaf.mesh->tree, max, mc)) state = ST_MESH;
        else { ++ci; state = ST_WORLD; run_world(); }
        return state != ST_DONE;
    }

    // step() in three parts, so that kernels can run the triangle tests of a whole warp's leaves cooperatively
    // (rsb_kernels.cuh mesh_leaf_coop) between the first and the last:
    //   step_descend()        _trace_branch down to the next mesh leaf: m_off / m_cnt / mc.max_range describe it
    //   (leaf test)           MeshData._trace_leaf over the leaf's triangles -> mh, true on a hit
    //   step_resolve(hit)     the rest of kd_advance, and the world-level work that follows
    int m_off, m_cnt;
    RSB_HD void step_descend() {
        int mnode = mc.node, msp = mc.sp;
        double mr = mc.max_range;
        KdNode n = kd_descend(mleaf.mesh->tree, max, leaf.mesh_stack, mnode, msp, mc.min_range, mr, *stats);
        stats->leaf(n.leaf.item_count);
        mc.node = mnode;
        mc.sp = msp;
        mc.max_range = mr;
        m_off = n.leaf.item_offset;
        m_cnt = n.leaf.item_count;
    }

    RSB_HD bool step_resolve(bool leaf_hit) {
        bool finished;
        if (leaf_hit) {
            mh.node = mc.node;
            finished = true;
        } else if (mc.sp == 0) {
            finished = true;
        } else {
            // the far child resumes with min_range = the plane distance = max_range of the leaf just left
            --mc.sp;
            mc.node = leaf.mesh_stack[mc.sp].node;
            mc.min_range = mc.max_range;
            mc.max_range = leaf.mesh_stack[mc.sp].tmax;
            finished = false;
        }
        if (finished) {
            accept_mesh(leaf_hit, mh);
            ++ci;
            state = ST_WORLD;
            run_world();
        }
        return state != ST_DONE;
    }

    RSB_HD bool step() {
        step_descend();
        const bool leaf_hit = m_cnt > 0 && mleaf(m_off, m_cnt, mc.max_range);
        return step_resolve(leaf_hit);
    }

    RSB_HD bool finish() {
        if (found) rec->node = leaf_node;
        return found;
    }
};

template <int FEAT, int S, class Stats>
RSB_HD bool world_hit_nested(const Scene& sc, const V3& o, const V3& d, double max_distance, KdStackEntry* stack, HitRec* rec, Stats& stats,
                             double* axbuf) {
    NestedTraversal<FEAT, S, Stats> t;
    t.init(sc, max_distance, stack, rec, stats, axbuf);
    if (t.begin(o, d)) {
        while (t.step()) {}
    }
    return t.finish();
}

// Closest hit of a world-space ray.  `stack` must hold RSB_KD_STACK entries, `axbuf` RSB_AX_WORDS * S doubles
// (element k of the calling thread at axbuf[k * S]).
// (Postponing the primitive tests -- a lane keeps descending and pre-testing AABBs until it holds candidates, and
// the warp runs the primitive tests only when every lane has some or has finished -- was measured SLOWER on the
// Cornell scene: 176 vs 156 us per 1M-ray wave, 9.3 vs 11.2 active lanes per instruction; profiles/README.md.)
template <int FEAT, int S, class Stats>
RSB_HD bool world_hit_ax(const Scene& sc, const V3& o, const V3& d, double max_distance, KdStackEntry* stack, HitRec* rec, Stats& stats,
                         double* axbuf) {
#ifndef RSB_NO_NESTED_LOOP
    if ((FEAT & RSB_FEAT_MESH) != 0) return world_hit_nested<FEAT, S>(sc, o, d, max_distance, stack, rec, stats, axbuf);
#endif
    WorldLeaf<Stats, FEAT, S> leaf;
    leaf.sc = &sc;
    leaf.ax.set(axbuf, o, d);
    leaf.max_distance = max_distance;
    leaf.mesh_stack = stack + (RSB_KD_STACK / 2);
    leaf.mesh_axbuf = axbuf + 9 * S;
    leaf.best = rec;
    leaf.stats = &stats;
    rec->u = rec->v = rec->w = 0.0f;
    rec->node = -1;
    rec->mesh_node = -1;
    KdCursor c;
    if (!kd_begin(sc.world, leaf.ax, c)) return false;
    int r;
    do { r = kd_advance(sc.world, leaf.ax, stack, c, leaf, stats, &rec->node); } while (r == KD_MORE);
    return r == KD_HIT;
}

// The same over thread-local storage (host builds, and device code outside the traversal kernels)
template <int FEAT = RSB_FEAT_ALL, class Stats>
RSB_HD bool world_hit(const Scene& sc, const V3& o, const V3& d, double max_distance, KdStackEntry* stack, HitRec* rec, Stats& stats) {
    double axbuf[RSB_AX_WORDS];
    return world_hit_ax<FEAT, 1>(sc, o, d, max_distance, stack, rec, stats, axbuf);
}

// Intersection geometry for a HitRec, in the world-level primitive's local space.
template <int FEAT = RSB_FEAT_ALL>
RSB_HD void world_hit_geometry(const Scene& sc, const V3& o, const V3& d, const HitRec& rec, Isect* is) {
    const Prim& p = sc.prims[rec.prim];
    is->t = rec.t;
    is->prim = rec.prim;
    is->leaf = rec.leaf;
    is->code = rec.code;
    is->u = rec.u; is->v = rec.v; is->w = rec.w;
    if (p.type <= PRIM_CONE) {
        V3 lo = xform_point(p.to_local, o);
        V3 ld = xform_vector(p.to_local, d);
        analytic_geometry<(FEAT & RSB_FEAT_CSG) != 0>(p.type, p.params, lo, ld, rec.t, rec.code, is);
    } else if ((FEAT & RSB_FEAT_MESH) && p.type == PRIM_MESH) {
        V3 lo = xform_point(p.to_local, o);
        V3 ld = xform_vector(p.to_local, d);
        MeshHit mh;
        mh.t = rec.t; mh.tri = rec.code; mh.node = rec.mesh_node; mh.u = rec.u; mh.v = rec.v; mh.w = rec.w;
        mesh_geometry(sc.meshes[p.mesh], lo, ld, mh, is);
    } else if ((FEAT & RSB_FEAT_CSG) && p.type >= PRIM_UNION) {
        CsgEvent ev;
        ev.t = rec.t; ev.leaf = rec.leaf; ev.code = (int16_t)rec.code;
        ev.flip = (int8_t)(rec.flip & 1);
        ev.exiting = (int8_t)((rec.flip >> 1) & 1);
        csg_geometry(sc, rec.prim, ev, o, d, is);
    }
}

// Primitive.contains for a world-level row behind BoundPrimitive.contains (boundprimitive.pyx:62-66)
template <int FEAT = RSB_FEAT_ALL, class Stats>
RSB_HD bool prim_contains(const Scene& sc, int id, const V3& pt, KdStackEntry* stack, Stats& stats) {
    const Prim& p = sc.prims[id];
    if (!box_contains(p.bbox, pt)) return false;
    if (p.type <= PRIM_CONE) return analytic_contains(p.type, p.params, xform_point(p.to_local, pt));
    if (!(FEAT & RSB_FEAT_MESH) && p.type == PRIM_MESH) return false;
    if (!(FEAT & RSB_FEAT_CSG) && p.type >= PRIM_UNION) return false;
    if (p.type == PRIM_MESH) {
        const Mesh& m = sc.meshes[p.mesh];
        if (!m.closed) return false;   // mesh.pyx:1290-1292
        return mesh_contains(m, xform_point(p.to_local, pt), stack, stats);
    }
    V3 lp = xform_point(p.to_local, pt);
    bool ca = CsgContains<RSB_CSG_MAX_DEPTH - 1>::run(sc, p.child_a, lp);
    if (p.type == PRIM_UNION) return ca || CsgContains<RSB_CSG_MAX_DEPTH - 1>::run(sc, p.child_b, lp);
    if (p.type == PRIM_INTERSECT) return ca && CsgContains<RSB_CSG_MAX_DEPTH - 1>::run(sc, p.child_b, lp);
    return ca && !CsgContains<RSB_CSG_MAX_DEPTH - 1>::run(sc, p.child_b, lp);
}

}  // namespace rsb
