/*
 * rs_oracle.c -- TEST INFRASTRUCTURE ONLY: a plain-C, scalar, recursive CPU restatement of Raysect's
 * ray/scene intersection + spectral trace path, written to mirror the REFERENCE's structure (recursive kd
 * traversal, stateful next_intersection() iterators for CSG, recursive Ray.trace with per-bin multiplies on
 * the unwind, per-sample Welford) rather than the product's (iterative traversal, event lists, log + replay).
 * It shares no code with source_b200/ apart from the C-ABI input structs of include/raysect_b200.h.
 *
 * Pinned: tests/test_c_oracle.py checks it against the golden vectors produced by the compiled reference
 * (tests/golden/, incl. the reference's own RNG known-answer vector) -- bit for bit on this container's libm.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may build or call it.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference/raysect).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/raysect_b200.h"

typedef struct { double x, y, z; } v3;
static v3 V(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static double comp(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
static double dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* core/math/vector.pyx:306-310 */
static v3 cross3(v3 a, v3 b) { return V(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
/* core/math/vector.pyx:313-337 */
static v3 norm3(v3 a) { double t = a.x * a.x + a.y * a.y + a.z * a.z; t = 1.0 / sqrt(t); return V(a.x * t, a.y * t, a.z * t); }
static double len3(v3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
/* core/math/point.pyx:253-281; m = rows 0..2 then m33 (the bottom row of an affine matrix is (0, 0, 0, m33)) */
static v3 xpoint(const double* m, v3 p) {
    double w = 0.0 * p.x + 0.0 * p.y + 0.0 * p.z + m[12];
    w = 1.0 / w;
    return V((m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3]) * w, (m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7]) * w,
             (m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]) * w);
}
/* core/math/vector.pyx:339-366 */
static v3 xvec(const double* m, v3 v) {
    return V(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
/* core/math/normal.pyx:222-248: transpose of the inverse */
static v3 xnormal_inv(const double* mi, v3 n) {
    return V(mi[0] * n.x + mi[4] * n.y + mi[8] * n.z, mi[1] * n.x + mi[5] * n.y + mi[9] * n.z, mi[2] * n.x + mi[6] * n.y + mi[10] * n.z);
}
/* core/math/vector.pyx:442-470 */
static v3 orthogonal3(v3 a) {
    v3 n = norm3(a), v = V(1, 0, 0);
    if (fabs(dot3(n, v)) > 0.5) v = V(0, 1, 0);
    double m = dot3(n, v);
    return norm3(V(v.x - m * n.x, v.y - m * n.y, v.z - m * n.z));
}

/* ------------------------------------------------------------------ RNG: core/math/random.pyx:99-265 */
#define NN 312
#define MM 156
typedef struct { uint64_t mt[NN]; int mti; } mt_t;
static void init_genrand64(mt_t* g, uint64_t seed) {               /* :110-122 */
    g->mt[0] = seed;
    for (int i = 1; i < NN; ++i) g->mt[i] = 6364136223846793005ULL * (g->mt[i - 1] ^ (g->mt[i - 1] >> 62)) + (uint64_t)i;
    g->mti = NN;
}
static void init_by_array64(mt_t* g, const uint64_t* key, uint64_t klen) {   /* :125-164 */
    init_genrand64(g, 19650218ULL);
    unsigned i = 1, j = 0;
    uint64_t k = NN > klen ? NN : klen;
    for (; k; --k) {
        g->mt[i] = (g->mt[i] ^ ((g->mt[i - 1] ^ (g->mt[i - 1] >> 62)) * 3935559000370003845ULL)) + key[j] + j;
        ++i; ++j;
        if (i >= NN) { g->mt[0] = g->mt[NN - 1]; i = 1; }
        if (j >= klen) j = 0;
    }
    for (k = NN - 1; k; --k) {
        g->mt[i] = (g->mt[i] ^ ((g->mt[i - 1] ^ (g->mt[i - 1] >> 62)) * 2862933555777941757ULL)) - i;
        ++i;
        if (i >= NN) { g->mt[0] = g->mt[NN - 1]; i = 1; }
    }
    g->mt[0] = 1ULL << 63;
}
static void rs_seed(mt_t* g, uint64_t d) {                          /* :215-243: d.to_bytes(8*312,'big') */
    uint64_t key[NN];
    memset(key, 0, sizeof(key));
    key[NN - 1] = d;
    init_by_array64(g, key, NN);
}
static uint64_t rand_u64(mt_t* g) {                                 /* :167-212 (batch refill, as the reference) */
    static const uint64_t mag01[2] = {0ULL, 0xB5026F5AA96619E9ULL};
    if (g->mti >= NN) {
        int i;
        uint64_t x;
        for (i = 0; i < NN - MM; ++i) {
            x = (g->mt[i] & 0xFFFFFFFF80000000ULL) | (g->mt[i + 1] & 0x7FFFFFFFULL);
            g->mt[i] = g->mt[i + MM] ^ (x >> 1) ^ mag01[x & 1];
        }
        for (; i < NN - 1; ++i) {
            x = (g->mt[i] & 0xFFFFFFFF80000000ULL) | (g->mt[i + 1] & 0x7FFFFFFFULL);
            g->mt[i] = g->mt[i + (MM - NN)] ^ (x >> 1) ^ mag01[x & 1];
        }
        x = (g->mt[NN - 1] & 0xFFFFFFFF80000000ULL) | (g->mt[0] & 0x7FFFFFFFULL);
        g->mt[NN - 1] = g->mt[MM - 1] ^ (x >> 1) ^ mag01[x & 1];
        g->mti = 0;
    }
    uint64_t x = g->mt[g->mti++];
    x ^= (x >> 29) & 0x5555555555555555ULL;
    x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
    x ^= (x << 37) & 0xFFF7EEE000000000ULL;
    x ^= (x >> 43);
    return x;
}
static double uniform(mt_t* g) { return (double)(rand_u64(g) >> 11) * (1.0 / 9007199254740992.0); }   /* :247-265 */

/* ------------------------------------------------------------------ scene (parsed from RsbSceneDesc) */
typedef struct { int type; double split; int count; const int32_t* items; } kdnode;   /* kdtree3d.pxd:38-43 */
typedef struct { kdnode* nodes; int n; double bounds[6]; int32_t* item_store; } kdtree;
typedef struct {
    const RsbMeshDesc* d;
    float* fnorm;       /* face normals, mesh.pyx:428-462 */
    kdtree tree;
    /* last-hit state of MeshData (mesh.pxd:56-60) */
    int ix, iy, iz; float sx, sy, sz; float u, v, w, t; int i;
    int leaf;           /* node id of the leaf that produced the last hit (kd-node parity; not reference state) */
} mesh_t;
typedef struct {
    const RsbSceneDesc* d;
    kdtree world;
    mesh_t* meshes;
    double* cdf;
    double imp_total;
} scene_t;

static int parse_tree(const uint8_t* p, int64_t size, kdtree* t) {   /* kdtree3d.pyx:914-984 (load) */
    const uint8_t* e = p + size;
    p += 4 + 4 + 8 + 8;
    memcpy(t->bounds, p, 48); p += 48;
    int32_t n; memcpy(&n, p, 4); p += 4;
    t->n = n;
    t->nodes = (kdnode*)calloc((size_t)n, sizeof(kdnode));
    t->item_store = (int32_t*)malloc((size_t)size);
    int32_t* store = t->item_store;
    for (int i = 0; i < n && p < e; ++i) {
        int32_t type; memcpy(&type, p, 4); p += 4;
        t->nodes[i].type = type;
        if (type == -1) {
            int32_t c; memcpy(&c, p, 4); p += 4;
            t->nodes[i].count = c;
            t->nodes[i].items = store;
            memcpy(store, p, (size_t)c * 4); p += (size_t)c * 4; store += c;
        } else {
            memcpy(&t->nodes[i].split, p, 8); p += 8;
            int32_t c; memcpy(&c, p, 4); p += 4;
            t->nodes[i].count = c;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ intersection record: core/intersection.pxd:37-53 */
typedef struct {
    int hit; double t; int prim; int exiting; v3 p, pin, pout, n;
    const double* w2p; const double* p2w;
    int tri; float u, v, w;
    int mesh_leaf;      /* mesh kd leaf of a mesh hit, else -1 */
} isect;

typedef struct { v3 o, d; double maxd; } ray_t;

/* core/boundingbox.pyx:200-245 */
static void slab(double o, double d, double lo, double hi, double* f, double* b) {
    double tmin, tmax;
    if (d != 0.0) {
        double r = 1.0 / d;
        if (d > 0) { tmin = (lo - o) * r; tmax = (hi - o) * r; } else { tmin = (hi - o) * r; tmax = (lo - o) * r; }
    } else {
        if (o < lo) { tmin = -INFINITY; tmax = -INFINITY; } else if (o > hi) { tmin = INFINITY; tmax = INFINITY; } else { tmin = -INFINITY; tmax = INFINITY; }
    }
    if (tmin > *f) *f = tmin;
    if (tmax < *b) *b = tmax;
}
/* core/boundingbox.pyx:180-198 */
static int box_intersect(const double* bx, const ray_t* r, double* f, double* b) {
    *f = -INFINITY; *b = INFINITY;
    slab(r->o.x, r->d.x, bx[0], bx[3], f, b);
    slab(r->o.y, r->d.y, bx[1], bx[4], f, b);
    slab(r->o.z, r->d.z, bx[2], bx[5], f, b);
    if (*f > *b) return 0;
    if (*f < 0.0 && *b < 0.0) return 0;
    return 1;
}
static int box_contains(const double* bx, v3 p) {                   /* boundingbox.pyx:247-263 */
    if (p.x < bx[0] || p.x > bx[3]) return 0;
    if (p.y < bx[1] || p.y > bx[4]) return 0;
    if (p.z < bx[2] || p.z > bx[5]) return 0;
    return 1;
}
/* core/math/cython/utility.pyx:376-420 */
static int solve_quadratic(double a, double b, double c, double* t0, double* t1) {
    double d = b * b - 4 * a * c;
    if (d < 0) return 0;
    double q = (b < 0) ? -0.5 * (b - sqrt(d)) : -0.5 * (b + sqrt(d));
    *t0 = q / a; *t1 = c / q;
    return 1;
}

/* ------------------------------------------------------------------ analytic primitives as iterators
 * A primitive instance keeps the reference's next_intersection() cache: (further, next_t, next_code, o, d). */
typedef struct prim_state {
    int further; double next_t; int next_code; v3 o, d;     /* analytic cache (sphere.pyx:146-153 etc.) */
    /* csg cache (csg.pyx:60-64) */
    int cache_invalid; isect ca, cb; int last_is_a; ray_t cache_ray;
    int tested;                                              /* BoundPrimitive._primitive_tested */
} prim_state;

typedef struct { scene_t* s; prim_state* st; int hit_leaf; /* world kd leaf in which the last World.hit accepted its hit */ } ctx_t;

static const double* P_params(scene_t* s, int id) { return s->d->prim_params + 6 * (size_t)id; }
static const double* P_tl(scene_t* s, int id) { return s->d->prim_to_local + 13 * (size_t)id; }
static const double* P_tr(scene_t* s, int id) { return s->d->prim_to_root + 13 * (size_t)id; }
static const double* P_ri(scene_t* s, int id) { return s->d->prim_root_inv + 13 * (size_t)id; }
static const double* P_bb(scene_t* s, int id) { return s->d->prim_bbox + 6 * (size_t)id; }

static void finish(isect* it, scene_t* s, int id, double t, v3 d, v3 hit, v3 in, v3 out, v3 n, int exiting_ge) {
    it->hit = 1; it->t = t; it->prim = id; it->p = hit; it->pin = in; it->pout = out; it->n = n;
    it->exiting = exiting_ge; (void)d;
    it->w2p = P_tl(s, id); it->p2w = P_tr(s, id); it->tri = -1; it->u = it->v = it->w = 0;
}

/* sphere.pyx:165-200 */
static void sphere_gen(scene_t* s, int id, v3 o, v3 d, double t, isect* it) {
    v3 h = V(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z), n = norm3(h);
    double dx = 1e-9 * n.x, dy = 1e-9 * n.y, dz = 1e-9 * n.z;
    finish(it, s, id, t, d, h, V(h.x - dx, h.y - dy, h.z - dz), V(h.x + dx, h.y + dy, h.z + dz), n, dot3(d, n) >= 0.0);
}
/* box.pyx:289-342; code = axis*2 + upper */
static double box_off(double h, double lo, double hi) { if (fabs(h - lo) < 1e-9) return 1e-9; else if (fabs(h - hi) < 1e-9) return -1e-9; return 0.0; }
static void box_gen(scene_t* s, int id, v3 o, v3 d, double t, int code, isect* it) {
    const double* q = P_params(s, id);
    v3 h = V(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z), n = V(0, 0, 0);
    double sg = (code & 1) ? 1.0 : -1.0;
    if ((code >> 1) == 0) n.x = sg; else if ((code >> 1) == 1) n.y = sg; else n.z = sg;
    v3 in = V(h.x + box_off(h.x, q[0], q[3]), h.y + box_off(h.y, q[1], q[4]), h.z + box_off(h.z, q[2], q[5]));
    finish(it, s, id, t, d, h, in, V(h.x + 1e-9 * n.x, h.y + 1e-9 * n.y, h.z + 1e-9 * n.z), n, dot3(d, n) >= 0.0);
}
/* cylinder.pyx:282-349; code 0 body, 1 lower, 2 upper */
static void cyl_gen(scene_t* s, int id, v3 o, v3 d, double t, int code, isect* it) {
    const double* q = P_params(s, id);
    double radius = q[0], height = q[1];
    v3 h = V(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z), n;
    if (code == 0) n = norm3(V(h.x, h.y, 0)); else if (code == 1) n = V(0, 0, -1); else n = V(0, 0, 1);
    double x, y, z;
    if (code == 0) { x = -1e-9 * n.x; y = -1e-9 * n.y; }
    else {
        x = 0; y = 0;
        if (h.x != 0.0 && h.y != 0.0) {
            double l = sqrt(h.x * h.x + h.y * h.y);
            if ((l - radius) < 1e-9) { l = 1.0 / l; x = -1e-9 * l * h.x; y = -1e-9 * l * h.y; }
        }
    }
    if (fabs(h.z) < 1e-9) z = 1e-9; else if (fabs(h.z - height) < 1e-9) z = -1e-9; else z = 0;
    finish(it, s, id, t, d, h, V(h.x + x, h.y + y, h.z + z), V(h.x + 1e-9 * n.x, h.y + 1e-9 * n.y, h.z + 1e-9 * n.z), n, dot3(d, n) >= 0.0);
}
/* cone.pyx:273-357; code 0 cone, 1 base */
static void cone_gen(scene_t* s, int id, v3 o, v3 d, double t, int code, isect* it) {
    const double* q = P_params(s, id);
    double radius = q[0], height = q[1];
    v3 h = V(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z), n;
    if (code == 1) n = V(0, 0, -1);
    else if (h.z >= height) n = V(0, 0, 1);
    else { double a = h.y / h.x, b = height / sqrt(1 + a * a); b = h.x < 0 ? -b : b; n = norm3(V(b, b * a, radius)); }
    double x = h.x - 1e-9 * n.x, y = h.y - 1e-9 * n.y, z = h.z - 1e-9 * n.z, k = radius / height;
    double inner_h = height - 1e-9 * sqrt(1 + k * k) / k;
    v3 in;
    if (z > inner_h) in = V(0, 0, inner_h);
    else if (z < 1e-9) {
        double inner_r = k * (height - 1e-9) - 1e-9 * sqrt(1 + k * k);
        double sc = inner_r / sqrt(h.x * h.x + h.y * h.y);
        in = V(sc * h.x, sc * h.y, 1e-9);
    } else in = V(x, y, z);
    finish(it, s, id, t, d, h, in, V(h.x + 1e-9 * n.x, h.y + 1e-9 * n.y, h.z + 1e-9 * n.z), n, dot3(d, n) >= 0.0);
}
/* parabola.pyx:270-342; code 0 parabola, 1 base */
static void parabola_gen(scene_t* s, int id, v3 o, v3 d, double t, int code, isect* it) {
    const double* q = P_params(s, id);
    double radius = q[0], height = q[1];
    v3 h = V(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z), n;
    if (code == 1) n = V(0, 0, -1);
    else { double k = 2 * height / (radius * radius); n = norm3(V(k * h.x, k * h.y, 1)); }
    double x = h.x, y = h.y, z = h.z, inner_r = radius - 1e-9, hr2 = h.x * h.x + h.y * h.y;   /* _interior_point */
    if (hr2 > inner_r * inner_r) { double sc = inner_r / sqrt(hr2); x = sc * h.x; y = sc * h.y; }
    if (h.z < 1e-9) z = 1e-9;
    else { x = h.x - n.x * 1e-9; y = h.y - n.y * 1e-9; z = h.z - n.z * 1e-9; }
    finish(it, s, id, t, d, h, V(x, y, z), V(h.x + 1e-9 * n.x, h.y + 1e-9 * n.y, h.z + 1e-9 * n.z), n, dot3(d, n) >= 0.0);
}
static void analytic_gen(scene_t* s, int id, v3 o, v3 d, double t, int code, isect* it) {
    switch (s->d->prim_type[id]) {
        case RSB_PRIM_PARABOLA: parabola_gen(s, id, o, d, t, code, it); break;
        case RSB_PRIM_SPHERE: sphere_gen(s, id, o, d, t, it); break;
        case RSB_PRIM_BOX: box_gen(s, id, o, d, t, code, it); break;
        case RSB_PRIM_CYLINDER: cyl_gen(s, id, o, d, t, code, it); break;
        default: cone_gen(s, id, o, d, t, code, it); break;
    }
}

/* near/far selection shared by the four hit() methods (sphere.pyx:137-163 etc.) */
static int select_hit(ctx_t* c, int id, v3 o, v3 d, double maxd, double t0, int c0, double t1, int c1, isect* it) {
    prim_state* st = &c->st[id];
    if (t0 > maxd || t1 < 0.0) return 0;
    if (t0 >= 0.0) {
        if (t1 <= maxd) { st->further = 1; st->next_t = t1; st->next_code = c1; st->o = o; st->d = d; }
        analytic_gen(c->s, id, o, d, t0, c0, it);
        return 1;
    } else if (t1 <= maxd) {
        analytic_gen(c->s, id, o, d, t1, c1, it);
        return 1;
    }
    return 0;
}

static void box_slab2(int axis, double o, double d, double lo, double hi, double* nt, double* ft, int* nc, int* fc) {   /* box.pyx:232-287 */
    double tmin, tmax; int fmin, fmax;
    if (d != 0.0) {
        double r = 1.0 / d;
        if (d > 0) { tmin = (lo - o) * r; tmax = (hi - o) * r; fmin = 0; fmax = 1; }
        else { tmin = (hi - o) * r; tmax = (lo - o) * r; fmin = 1; fmax = 0; }
    } else {
        if (o < lo) { tmin = -INFINITY; tmax = -INFINITY; } else if (o > hi) { tmin = INFINITY; tmax = INFINITY; } else { tmin = -INFINITY; tmax = INFINITY; }
        fmin = -1; fmax = -1;
    }
    if (tmin > *nt) { *nt = tmin; *nc = axis * 2 + (fmin == 0 ? 0 : 1); }
    if (tmax < *ft) { *ft = tmax; *fc = axis * 2 + (fmax == 0 ? 0 : 1); }
}

static int prim_hit(ctx_t* c, int id, const ray_t* ray, isect* it);
static int prim_next(ctx_t* c, int id, isect* it);

/* hit() of the analytic primitives; ray given in the primitive's PARENT space */
static int analytic_hit(ctx_t* c, int id, const ray_t* ray, isect* it) {
    scene_t* s = c->s;
    prim_state* st = &c->st[id];
    st->further = 0;
    v3 o = xpoint(P_tl(s, id), ray->o), d = xvec(P_tl(s, id), ray->d);
    const double* q = P_params(s, id);
    double t0, t1;
    switch (s->d->prim_type[id]) {
        case RSB_PRIM_SPHERE: {                                 /* sphere.pyx:115-163 */
            double a = d.x * d.x + d.y * d.y + d.z * d.z;
            double b = 2 * (d.x * o.x + d.y * o.y + d.z * o.z);
            double cc = o.x * o.x + o.y * o.y + o.z * o.z - q[0] * q[0];
            if (!solve_quadratic(a, b, cc, &t0, &t1)) return 0;
            if (t0 > t1) { double tmp = t0; t0 = t1; t1 = tmp; }
            return select_hit(c, id, o, d, ray->maxd, t0, 0, t1, 0, it);
        }
        case RSB_PRIM_BOX: {                                    /* box.pyx:157-219 */
            double nt = -INFINITY, ft = INFINITY; int nc = 0, fc = 0;
            box_slab2(0, o.x, d.x, q[0], q[3], &nt, &ft, &nc, &fc);
            box_slab2(1, o.y, d.y, q[1], q[4], &nt, &ft, &nc, &fc);
            box_slab2(2, o.z, d.z, q[2], q[5], &nt, &ft, &nc, &fc);
            if (nt > ft) return 0;
            return select_hit(c, id, o, d, ray->maxd, nt, nc, ft, fc, it);
        }
        case RSB_PRIM_CYLINDER: {                               /* cylinder.pyx:148-271 */
            double radius = q[0], height = q[1], nt, ft; int nc, fc;
            if (d.x == 0 && d.y == 0) {
                if ((o.x * o.x + o.y * o.y) <= (radius * radius)) { nt = -INFINITY; ft = INFINITY; nc = 2; fc = 2; }
                else return 0;
            } else {
                double a = d.x * d.x + d.y * d.y, b = 2.0 * (d.x * o.x + d.y * o.y), cc = o.x * o.x + o.y * o.y - radius * radius;
                if (!solve_quadratic(a, b, cc, &t0, &t1)) return 0;
                if (t0 > t1) { double tmp = t0; t0 = t1; t1 = tmp; }
                nt = t0; ft = t1; nc = 0; fc = 0;
            }
            if (d.z != 0.0) {
                double tmp = 1.0 / d.z; int f0, f1;
                if (d.z > 0) { t0 = -o.z * tmp; t1 = (height - o.z) * tmp; f0 = 1; f1 = 2; }
                else { t0 = (height - o.z) * tmp; t1 = -o.z * tmp; f0 = 2; f1 = 1; }
                if (t0 > nt) { nt = t0; nc = f0; }
                if (t1 < ft) { ft = t1; fc = f1; }
            }
            if (nt > ft) return 0;
            return select_hit(c, id, o, d, ray->maxd, nt, nc, ft, fc, it);
        }
        case RSB_PRIM_PARABOLA: {                               /* parabola.pyx:141-257 */
            double radius = q[0], height = q[1], k = height / (radius * radius); int y0, y1;
            double a = k * (d.x * d.x + d.y * d.y);
            double b = 2 * k * (d.x * o.x + d.y * o.y) + d.z;
            double cc = k * (o.x * o.x + o.y * o.y) - (height - o.z);
            if (!solve_quadratic(a, b, cc, &t0, &t1)) return 0;
            if (t0 == t1) {
                t0 = -b / (2.0 * a); y0 = 0;
                k = -o.z / d.z;
                double ex = o.x + k * d.x;
                double r2 = ex * ex + (o.y + k * (d.y * d.y));       /* parabola.pyx:184, `**2` placement as written */
                if (r2 <= radius * radius) { t1 = k; y1 = 1; } else { t1 = t0; y1 = y0; }
            } else {
                int o0 = (o.z + t0 * d.z) < 0, o1 = (o.z + t1 * d.z) < 0;
                if (o0 && o1) return 0;
                else if (!o0 && o1) { y0 = 0; t1 = -o.z / d.z; y1 = 1; }
                else if (o0 && !o1) { y0 = 1; t0 = -o.z / d.z; y1 = 0; }
                else { y0 = 0; y1 = 0; }
            }
            if (t0 > t1) { double tmp = t0; t0 = t1; t1 = tmp; int ti = y0; y0 = y1; y1 = ti; }
            return select_hit(c, id, o, d, ray->maxd, t0, y0, t1, y1, it);
        }
        default: {                                              /* cone.pyx:142-262 */
            double radius = q[0], height = q[1], k = radius / height; int y0, y1;
            k = k * k;
            double a = d.x * d.x + d.y * d.y - k * d.z * d.z;
            double b = 2 * (d.x * o.x + d.y * o.y - k * d.z * (o.z - height));
            double cc = o.x * o.x + o.y * o.y - k * (o.z - height) * (o.z - height);
            if (!solve_quadratic(a, b, cc, &t0, &t1)) return 0;
            if (t0 == t1) {
                t0 = -b / (2.0 * a); y0 = 0;
                k = -o.z / d.z;
                double ex = o.x + k * d.x;
                double r2 = ex * ex + (o.y + k * (d.y * d.y));       /* cone.pyx:185, `**2` placement as written */
                if (r2 <= radius * radius) { t1 = k; y1 = 1; } else { t1 = t0; y1 = y0; }
            } else {
                double z0 = o.z + t0 * d.z, z1 = o.z + t1 * d.z;
                int o0 = z0 < 0 || z0 > height, o1 = z1 < 0 || z1 > height;
                if (o0 && o1) return 0;
                else if (!o0 && o1) { y0 = 0; t1 = -o.z / d.z; y1 = 1; }
                else if (o0 && !o1) { y0 = 1; t0 = -o.z / d.z; y1 = 0; }
                else { y0 = 0; y1 = 0; }
            }
            if (t0 > t1) { double tmp = t0; t0 = t1; t1 = tmp; int ti = y0; y0 = y1; y1 = ti; }
            return select_hit(c, id, o, d, ray->maxd, t0, y0, t1, y1, it);
        }
    }
}
static int analytic_next(ctx_t* c, int id, isect* it) {           /* sphere.pyx:156-163 etc. */
    prim_state* st = &c->st[id];
    if (!st->further) return 0;
    st->further = 0;
    analytic_gen(c->s, id, st->o, st->d, st->next_t, st->next_code, it);
    return 1;
}

/* ------------------------------------------------------------------ CSG: primitive/csg.pyx:132-241 + rules */
static int bound_hit(ctx_t* c, int id, const ray_t* ray, isect* it) {     /* boundprimitive.pyx:42-51 */
    double f, b;
    if (box_intersect(P_bb(c->s, id), ray, &f, &b)) { c->st[id].tested = 1; return prim_hit(c, id, ray, it); }
    c->st[id].tested = 0;
    return 0;
}
static int bound_next(ctx_t* c, int id, isect* it) {                     /* boundprimitive.pyx:53-60 */
    if (c->st[id].tested) return prim_next(c, id, it);
    return 0;
}
static int csg_valid(int op, const isect* a, const isect* b, int closest_is_a) {   /* :326-348, :424-446, :526-548 */
    int ia = a->hit && a->exiting, ib = b->hit && b->exiting;
    if (op == RSB_PRIM_UNION) {
        if (!ia && !ib) return 1; else if (ia && !ib && closest_is_a) return 1; else if (!ia && ib && !closest_is_a) return 1;
        return 0;
    } else if (op == RSB_PRIM_INTERSECT) {
        if (ia && ib) return 1; else if (ia && !ib && !closest_is_a) return 1; else if (!ia && ib && closest_is_a) return 1;
        return 0;
    }
    if (!ia && !ib && closest_is_a) return 1; else if (ia && !ib) return 1; else if (ia && ib && !closest_is_a) return 1;
    return 0;
}
/* _identify_intersection, csg.pyx:181-224 */
static int csg_identify(ctx_t* c, int id, const ray_t* ray, isect a, isect b, isect* out) {
    scene_t* s = c->s;
    prim_state* st = &c->st[id];
    int op = s->d->prim_type[id], ca = s->d->prim_child_a[id], cb = s->d->prim_child_b[id];
    for (;;) {
        if (!a.hit && !b.hit) return 0;
        int closest_is_a = a.hit && (!b.hit || a.t < b.t);          /* _closest_intersection :226-234 */
        const isect* closest = closest_is_a ? &a : &b;
        if (csg_valid(op, &a, &b, closest_is_a)) {
            if (closest->t <= ray->maxd) {
                st->cache_ray = *ray; st->ca = a; st->cb = b; st->last_is_a = closest_is_a; st->cache_invalid = 0;
                isect r = *closest;
                if (op == RSB_PRIM_SUBTRACT && !closest_is_a) {     /* _modify_intersection :550-568 */
                    v3 tmp = r.pin; r.pin = r.pout; r.pout = tmp;
                    r.n = V(-r.n.x, -r.n.y, -r.n.z);
                    r.exiting = !r.exiting;
                }
                /* re-express in this CSG primitive's space (:200-208); the child row that produced `closest` is
                   the operand (a or b) of THIS node: its to_root / inverse are rows ca / cb */
                int child = closest_is_a ? ca : cb;
                r.p = xpoint(P_tr(s, child), r.p);
                r.pin = xpoint(P_tr(s, child), r.pin);
                r.pout = xpoint(P_tr(s, child), r.pout);
                r.n = xnormal_inv(P_ri(s, child), r.n);
                r.w2p = P_tl(s, id); r.p2w = P_tr(s, id); r.prim = id;
                *out = r;
                return 1;
            }
            return 0;
        }
        if (closest_is_a) { if (!bound_next(c, ca, &a)) a.hit = 0; }
        else { if (!bound_next(c, cb, &b)) b.hit = 0; }
    }
}
static int csg_hit(ctx_t* c, int id, const ray_t* ray, isect* out) {       /* csg.pyx:132-155 */
    scene_t* s = c->s;
    prim_state* st = &c->st[id];
    st->cache_invalid = 1;
    ray_t lr;
    lr.o = xpoint(P_tl(s, id), ray->o); lr.d = xvec(P_tl(s, id), ray->d); lr.maxd = INFINITY;
    int op = s->d->prim_type[id];
    isect a, b;
    a.hit = bound_hit(c, s->d->prim_child_a[id], &lr, &a);
    if (!a.hit && op != RSB_PRIM_UNION) return 0;                          /* terminate_early */
    b.hit = bound_hit(c, s->d->prim_child_b[id], &lr, &b);
    return csg_identify(c, id, ray, a, b, out);
}
static int csg_next(ctx_t* c, int id, isect* out) {                        /* csg.pyx:160-179 */
    scene_t* s = c->s;
    prim_state* st = &c->st[id];
    if (st->cache_invalid) return 0;
    isect a = st->ca, b = st->cb;
    if (st->last_is_a) { if (!bound_next(c, s->d->prim_child_a[id], &a)) a.hit = 0; }
    else { if (!bound_next(c, s->d->prim_child_b[id], &b)) b.hit = 0; }
    ray_t r = st->cache_ray;
    return csg_identify(c, id, &r, a, b, out);
}

/* ------------------------------------------------------------------ mesh: primitive/mesh/mesh.pyx */
static float vtx(const mesh_t* m, int i, int k) { return m->d->vertices[3 * (size_t)i + k]; }
static int hit_triangle(mesh_t* m, int i, const ray_t* ray, float* out) {   /* :616-713 */
    const int32_t* row = m->d->triangles + (size_t)i * m->d->tri_stride;
    float v1[3], v2[3], v3_[3];
    double oo[3] = {ray->o.x, ray->o.y, ray->o.z};
    for (int k = 0; k < 3; ++k) {
        v1[k] = (float)(vtx(m, row[0], k) - oo[k]);
        v2[k] = (float)(vtx(m, row[1], k) - oo[k]);
        v3_[k] = (float)(vtx(m, row[2], k) - oo[k]);
    }
    int ix = m->ix, iy = m->iy, iz = m->iz;
    float sx = m->sx, sy = m->sy, sz = m->sz;
    float x1 = v1[ix] - sx * v1[iz], x2 = v2[ix] - sx * v2[iz], x3 = v3_[ix] - sx * v3_[iz];
    float y1 = v1[iy] - sy * v1[iz], y2 = v2[iy] - sy * v2[iz], y3 = v3_[iy] - sy * v3_[iz];
    float u = x3 * y2 - y3 * x2, v = x1 * y3 - y1 * x3, w = x2 * y1 - y2 * x1;
    if (u == 0.0 || v == 0.0 || w == 0.0) {
        u = (float)((double)x3 * (double)y2 - (double)y3 * (double)x2);
        v = (float)((double)x1 * (double)y3 - (double)y1 * (double)x3);
        w = (float)((double)x2 * (double)y1 - (double)y2 * (double)x1);
    }
    if ((u < 0.0 || v < 0.0 || w < 0.0) && (u > 0.0 || v > 0.0 || w > 0.0)) return 0;
    float det = u + v + w;
    if (det == 0.0) return 0;
    float z1 = sz * v1[iz], z2 = sz * v2[iz], z3 = sz * v3_[iz];
    float t = u * z1 + v * z2 + w * z3;
    if (det > 0.0) { if (t < 0.0 || t > ray->maxd * det) return 0; }
    else { if (t > 0.0 || t < ray->maxd * det) return 0; }
    float dr = 1.0 / det;
    out[0] = u * dr; out[1] = v * dr; out[2] = w * dr; out[3] = t * dr;
    return 1;
}
static int mesh_leaf(mesh_t* m, const kdnode* n, const ray_t* ray, double max_range) {   /* :520-563 */
    double distance = ray->maxd < max_range ? ray->maxd : max_range;
    int closest = -1; double u = 0, v = 0, w = 0;
    for (int k = 0; k < n->count; ++k) {
        float h[4];
        if (hit_triangle(m, n->items[k], ray, h)) {
            double t = h[3];
            if (t < distance) { distance = t; closest = n->items[k]; u = h[0]; v = h[1]; w = h[2]; }
        }
    }
    if (closest < 0) return 0;
    m->u = (float)u; m->v = (float)v; m->w = (float)w; m->t = (float)distance; m->i = closest;
    return 1;
}

/* recursive traversal, core/math/spatial/kdtree3d.pyx:609-700; leaf handler selected by `mesh` */
static int world_leaf(ctx_t* c, const kdnode* n, const ray_t* ray, double max_range, isect* best);
static int trace_node(ctx_t* c, const kdtree* t, mesh_t* mesh, int id, const ray_t* ray, double min_range, double max_range, isect* best) {
    const kdnode* n = &t->nodes[id];
    if (n->type == -1) {
        int found = mesh ? mesh_leaf(mesh, n, ray, max_range) : world_leaf(c, n, ray, max_range, best);
        if (found) { if (mesh) mesh->leaf = id; else c->hit_leaf = id; }
        return found;
    }
    int axis = n->type;
    double split = n->split, origin = comp(ray->o, axis), direction = comp(ray->d, axis);
    int lower_id = id + 1, upper_id = n->count;
    if (direction == 0) {
        if (origin < split) return trace_node(c, t, mesh, lower_id, ray, min_range, max_range, best);
        return trace_node(c, t, mesh, upper_id, ray, min_range, max_range, best);
    }
    double plane_distance = (split - origin) / direction;
    int below = origin < split || (origin == split && direction < 0);
    int near_id = below ? lower_id : upper_id, far_id = below ? upper_id : lower_id;
    if (plane_distance > max_range || plane_distance <= 0) return trace_node(c, t, mesh, near_id, ray, min_range, max_range, best);
    if (plane_distance < min_range) return trace_node(c, t, mesh, far_id, ray, min_range, max_range, best);
    if (trace_node(c, t, mesh, near_id, ray, min_range, plane_distance, best)) return 1;
    return trace_node(c, t, mesh, far_id, ray, plane_distance, max_range, best);
}
static int mesh_trace(ctx_t* c, mesh_t* m, const ray_t* ray) {              /* :506-518, :566-610 */
    m->u = m->v = m->w = -1.0f; m->t = INFINITY; m->i = -1;
    int ix, iy, iz;
    if (fabs(ray->d.x) > fabs(ray->d.y) && fabs(ray->d.x) > fabs(ray->d.z)) { ix = 1; iy = 2; iz = 0; }
    else if (fabs(ray->d.y) > fabs(ray->d.x) && fabs(ray->d.y) > fabs(ray->d.z)) { ix = 2; iy = 0; iz = 1; }
    else { ix = 0; iy = 1; iz = 2; }
    float rdz = (float)comp(ray->d, iz);
    if (rdz < 0.0) { int tmp = ix; ix = iy; iy = tmp; }
    m->sz = 1.0 / rdz;
    m->sx = comp(ray->d, ix) * m->sz;
    m->sy = comp(ray->d, iy) * m->sz;
    m->ix = ix; m->iy = iy; m->iz = iz;
    double f, b;
    if (!box_intersect(m->tree.bounds, ray, &f, &b)) return 0;
    return trace_node(c, &m->tree, m, 0, ray, f, b, NULL);
}
static int mesh_hit(ctx_t* c, int id, const ray_t* ray, isect* it) {        /* :1178-1211, :718-800 */
    scene_t* s = c->s;
    mesh_t* m = &s->meshes[s->d->prim_mesh[id]];
    ray_t lr;
    lr.o = xpoint(P_tl(s, id), ray->o); lr.d = xvec(P_tl(s, id), ray->d); lr.maxd = ray->maxd;
    if (!mesh_trace(c, m, &lr)) return 0;
    double t = m->t;
    const float* f = m->fnorm + 3 * (size_t)m->i;
    v3 fn = V(f[0], f[1], f[2]);
    v3 h = V(lr.o.x + lr.d.x * t, lr.o.y + lr.d.y * t, lr.o.z + lr.d.z * t);
    v3 n;
    if (m->d->smoothing && m->d->vertex_normals && m->d->tri_stride == 6) {
        const int32_t* row = m->d->triangles + (size_t)m->i * 6;
        const float* a = m->d->vertex_normals + 3 * (size_t)row[3];
        const float* b = m->d->vertex_normals + 3 * (size_t)row[4];
        const float* cc = m->d->vertex_normals + 3 * (size_t)row[5];
        float nx = m->u * a[0] + m->v * b[0] + m->w * cc[0];
        float ny = m->u * a[1] + m->v * b[1] + m->w * cc[1];
        float nz = m->u * a[2] + m->v * b[2] + m->w * cc[2];
        n = norm3(V(nx, ny, nz));
    } else n = norm3(fn);
    finish(it, s, id, t, lr.d, h, V(h.x - fn.x * 1e-6, h.y - fn.y * 1e-6, h.z - fn.z * 1e-6),
           V(h.x + fn.x * 1e-6, h.y + fn.y * 1e-6, h.z + fn.z * 1e-6), n, dot3(lr.d, fn) > 0.0);
    it->tri = m->i; it->u = m->u; it->v = m->v; it->w = m->w;
    it->mesh_leaf = m->leaf;
    return 1;
}

static int prim_hit(ctx_t* c, int id, const ray_t* ray, isect* it) {
    int t = c->s->d->prim_type[id];
    if (t <= RSB_PRIM_CONE) return analytic_hit(c, id, ray, it);
    if (t == RSB_PRIM_MESH) return mesh_hit(c, id, ray, it);
    return csg_hit(c, id, ray, it);
}
static int prim_next(ctx_t* c, int id, isect* it) {
    int t = c->s->d->prim_type[id];
    if (t <= RSB_PRIM_CONE) return analytic_next(c, id, it);
    if (t == RSB_PRIM_MESH) return 0;
    return csg_next(c, id, it);
}

/* _PrimitiveKDTree._trace_leaf, core/acceleration/kdtree.pyx:73-122 */
static int world_leaf(ctx_t* c, const kdnode* n, const ray_t* ray, double max_range, isect* best) {
    double distance = ray->maxd < max_range ? ray->maxd : max_range;
    int found = 0;
    for (int k = 0; k < n->count; ++k) {
        isect it;
        if (bound_hit(c, n->items[k], ray, &it) && it.t <= distance) { distance = it.t; *best = it; found = 1; }
    }
    return found;
}
static int world_hit(ctx_t* c, const ray_t* ray, isect* best) {            /* world.pyx:125-146, kdtree3d.pyx:589-607 */
    double f, b;
    best->hit = 0;
    if (!box_intersect(c->s->world.bounds, ray, &f, &b)) return 0;
    if (!trace_node(c, &c->s->world, NULL, 0, ray, f, b, best)) return 0;
    best->hit = 1;
    return 1;
}

/* ------------------------------------------------------------------ contains */
static int prim_contains(ctx_t* c, int id, v3 p);
static int bound_contains(ctx_t* c, int id, v3 p) { return box_contains(P_bb(c->s, id), p) ? prim_contains(c, id, p) : 0; }
static int prim_contains(ctx_t* c, int id, v3 p) {
    scene_t* s = c->s;
    int t = s->d->prim_type[id];
    const double* q = P_params(s, id);
    if (t == RSB_PRIM_MESH) {                                               /* mesh.pyx:1277-1295, :805-830 */
        mesh_t* m = &s->meshes[s->d->prim_mesh[id]];
        if (!m->d->closed) return 0;
        ray_t r; r.o = xpoint(P_tl(s, id), p); r.d = V(0, 0, 1); r.maxd = INFINITY;
        if (!mesh_trace(c, m, &r)) return 0;
        return m->fnorm[3 * (size_t)m->i + 2] > 0.0;
    }
    v3 l = xpoint(P_tl(s, id), p);
    switch (t) {
        case RSB_PRIM_SPHERE: return (l.x * l.x + l.y * l.y + l.z * l.z) <= q[0] * q[0];
        case RSB_PRIM_BOX:
            if (l.x < q[0] || l.x > q[3]) return 0;
            if (l.y < q[1] || l.y > q[4]) return 0;
            if (l.z < q[2] || l.z > q[5]) return 0;
            return 1;
        case RSB_PRIM_CYLINDER: return (0.0 <= l.z && l.z <= q[1]) && ((l.x * l.x + l.y * l.y) <= (q[0] * q[0]));
        case RSB_PRIM_PARABOLA:                                                 /* parabola.pyx:344-363 */
            if (l.z < 0 || l.z > q[1]) return 0;
            return sqrt(l.x * l.x + l.y * l.y) <= q[0] * sqrt((q[1] - l.z) / q[1]);
        case RSB_PRIM_CONE: {
            if (l.z < 0 || l.z > q[1]) return 0;
            double pr = l.x * l.x + l.y * l.y, cr = (q[1] - l.z) * q[0] / q[1];
            cr *= cr;
            return pr <= cr;
        }
        case RSB_PRIM_UNION: return bound_contains(c, s->d->prim_child_a[id], l) || bound_contains(c, s->d->prim_child_b[id], l);
        case RSB_PRIM_INTERSECT: return bound_contains(c, s->d->prim_child_a[id], l) && bound_contains(c, s->d->prim_child_b[id], l);
        default: return bound_contains(c, s->d->prim_child_a[id], l) && !bound_contains(c, s->d->prim_child_b[id], l);
    }
}
/* kdtree3d.pyx:736-792 + acceleration/kdtree.pyx:124-160 */
static int world_contains(ctx_t* c, v3 p, int32_t* out, int cap) {
    kdtree* t = &c->s->world;
    if (!box_contains(t->bounds, p)) return 0;
    int id = 0;
    while (t->nodes[id].type != -1) id = (comp(p, t->nodes[id].type) < t->nodes[id].split) ? id + 1 : t->nodes[id].count;
    int n = 0;
    for (int k = 0; k < t->nodes[id].count; ++k) {
        int pid = t->nodes[id].items[k];
        if (bound_contains(c, pid, p)) { if (n < cap) out[n] = pid; ++n; }
    }
    return n;
}

/* ------------------------------------------------------------------ optical: Ray.trace recursion */
typedef struct {
    ctx_t* c; const RsbRayConfig* cfg; const RsbSpectral* sp; mt_t* rng; uint64_t rays;
} tracer;

static int find_index(const double* x, int n, double v) {                  /* utility.pyx:40-94 */
    if (v < x[0]) return -1;
    int top = n - 1;
    if (v >= x[top]) return top;
    int bottom = 0, bis = top / 2;
    while ((top - bottom) != 1) { if (v >= x[bis]) bottom = bis; else top = bis; bis = (top + bottom) / 2; }
    return bottom;
}
static double max0(double x) { return x > 0 ? x : 0.0; }
static v3 vector_sphere(mt_t* g) {                                          /* random.pyx:375-387 */
    double z = 1.0 - 2.0 * uniform(g), r = sqrt(max0(1.0 - z * z)), phi = 2.0 * M_PI * uniform(g);
    return V(r * cos(phi), r * sin(phi), z);
}
static v3 vector_cone_uniform(mt_t* g, double theta) {                     /* random.pyx:425-445 */
    theta *= 0.017453292519943295;
    double phi = 2.0 * M_PI * uniform(g), ct = cos(theta), z = uniform(g) * (1 - ct) + ct, r = sqrt(max0(1.0 - z * z));
    return V(r * cos(phi), r * sin(phi), z);
}
static v3 hemisphere_cosine(mt_t* g) {                                      /* solidangle.pyx:228-232 */
    double r = sqrt(uniform(g)), phi = 2.0 * M_PI * uniform(g), x = r * cos(phi), y = r * sin(phi);
    return V(x, y, sqrt(max0(1.0 - x * x - y * y)));
}
static v3 important_sample(tracer* T, v3 origin) {                          /* optical/scenegraph/world.pyx:155-200 */
    scene_t* s = T->c->s;
    int idx = find_index(s->cdf, s->d->n_important, uniform(T->rng)) + 1;
    const double* sp = s->d->imp_sphere + 4 * idx;
    v3 dir = V(sp[0] - origin.x, sp[1] - origin.y, sp[2] - origin.z);
    double distance = len3(dir);
    if (distance == 0 || distance < sp[3]) return vector_sphere(T->rng);
    double ang = asin(sp[3] / distance);
    v3 smp = vector_cone_uniform(T->rng, ang * 180 / M_PI);
    dir = norm3(dir);
    v3 up = orthogonal3(dir), right = cross3(up, dir);                       /* core/math/cython/transform.pyx:45-70 */
    return V(right.x * smp.x + up.x * smp.y + dir.x * smp.z, right.y * smp.x + up.y * smp.y + dir.y * smp.z,
             right.z * smp.x + up.z * smp.y + dir.z * smp.z);
}
static double important_pdf(tracer* T, v3 origin, v3 direction) {           /* world.pyx:203-253 */
    scene_t* s = T->c->s;
    double pdf_all = 0;
    for (int i = 0; i < s->d->n_important; ++i) {
        const double* sp = s->d->imp_sphere + 4 * i;
        v3 axis = V(sp[0] - origin.x, sp[1] - origin.y, sp[2] - origin.z);
        double distance = len3(axis), solid;
        if (distance == 0 || distance < sp[3]) solid = 4 * M_PI;
        else {
            double t = sp[3] / distance, ac = sqrt(1 - t * t);
            axis = norm3(axis);
            if (dot3(direction, axis) < ac) continue;
            solid = 2 * M_PI * (1 - ac);
        }
        pdf_all += (s->d->imp_weight[i] / s->imp_total) * (1 / solid);
    }
    return pdf_all;
}

/* RoughConductor helpers, optical/material/conductor.pyx:203-247, 292-310 */
static double rough_d(v3 h, double r) { double r2 = r * r, h2 = h.z * h.z, k = h2 * (r2 - 1) + 1; return r2 / (M_PI * k * k); }
static double rough_g1(v3 v, double r) { double r2 = r * r; return 2 * v.z / (v.z + sqrt(r2 + (1 - r2) * (v.z * v.z))); }
static double rough_pdf(v3 si, v3 so, double r) {
    v3 h = V(si.x + so.x, si.y + so.y, si.z + so.z);
    if (len3(h) == 0.0) return 0.0;
    h = norm3(h);
    return 0.25 * rough_d(h, r) * fabs(h.z / dot3(so, h));
}
static v3 rough_sample(mt_t* g, v3 si, double r) {
    double e1 = uniform(g), e2 = uniform(g);
    double theta = atan(r * sqrt(e1) / sqrt(1 - e1)), phi = 2 * M_PI * e2;
    v3 f = V(cos(phi) * sin(theta), sin(phi) * sin(theta), cos(theta));
    double t = 2 * dot3(si, f);
    return V(t * f.x - si.x, t * f.y - si.y, t * f.z - si.z);
}

/* Ray.trace, optical/ray.pyx:338-401: fills spectrum[bins]; recursion through the material */
static void trace_ka(tracer* T, ray_t ray, int depth, double* spectrum, int keep_alive);
static void trace(tracer* T, ray_t ray, int depth, double* spectrum) { trace_ka(T, ray, depth, spectrum, 0); }

static void trace_ka(tracer* T, ray_t ray, int depth, double* spectrum, int keep_alive) {
    const RsbRayConfig* cfg = T->cfg;
    const RsbSceneDesc* d = T->c->s->d;
    int bins = cfg->bins;
    memset(spectrum, 0, sizeof(double) * (size_t)bins);
    double normalisation;
    if (keep_alive || depth < cfg->extinction_min_depth) normalisation = 1.0;     /* ray.pyx:382 */
    else {
        if (depth >= cfg->max_depth || uniform(T->rng) < cfg->extinction_prob) return;
        normalisation = 1 / (1 - cfg->extinction_prob);
    }
    isect it;
    if (!world_hit(T->c, &ray, &it)) return;
    int mat = d->prim_material[it.prim], mtype = d->mat_type[mat];
    const double* table = T->sp->tables + (size_t)mat * bins;
    if (mtype == RSB_MAT_EMITTER) {                                          /* emitter/uniform.pyx:67-81 */
        for (int i = 0; i < bins; ++i) spectrum[i] = table[i] * T->sp->scale[mat];
    } else if (mtype == RSB_MAT_LAMBERT || mtype == RSB_MAT_ROUGH_CONDUCTOR) {  /* ContinuousBSDF: material.pyx:291-361; lambert.pyx:77-105 | conductor.pyx:157-344 */
        const int rough = mtype == RSB_MAT_ROUGH_CONDUCTOR;
        const double rgh = T->sp->scale[mat];
        v3 n = it.n, refl_o;
        if (it.exiting) { refl_o = xpoint(it.p2w, it.pin); n = V(-n.x, -n.y, -n.z); } else refl_o = xpoint(it.p2w, it.pout);
        v3 tg = orthogonal3(n), bt = cross3(n, tg);
        double p2s[3][3] = {{tg.x, tg.y, tg.z}, {bt.x, bt.y, bt.z}, {n.x, n.y, n.z}};
        double s2p[3][3] = {{tg.x, bt.x, n.x}, {tg.y, bt.y, n.y}, {tg.z, bt.z, n.z}};
        double w2s[3][3], s2w[3][3];
        for (int r = 0; r < 3; ++r)
            for (int cc = 0; cc < 3; ++cc) {
                w2s[r][cc] = p2s[r][0] * it.w2p[cc] + p2s[r][1] * it.w2p[4 + cc] + p2s[r][2] * it.w2p[8 + cc] + 0.0 * 0.0;
                s2w[r][cc] = it.p2w[4 * r] * s2p[0][cc] + it.p2w[4 * r + 1] * s2p[1][cc] + it.p2w[4 * r + 2] * s2p[2][cc] + it.p2w[4 * r + 3] * 0.0;
            }
        v3 so, wo;
        double pdf;
        v3 si = V(0, 0, 0);                                                   /* s_incoming = (world_to_surface . d).neg() */
        if (rough) {
            v3 t = V(w2s[0][0] * ray.d.x + w2s[0][1] * ray.d.y + w2s[0][2] * ray.d.z, w2s[1][0] * ray.d.x + w2s[1][1] * ray.d.y + w2s[1][2] * ray.d.z,
                     w2s[2][0] * ray.d.x + w2s[2][1] * ray.d.y + w2s[2][2] * ray.d.z);
            si = V(-t.x, -t.y, -t.z);
        }
        if (cfg->importance_sampling && T->c->s->imp_total > 0) {
            v3 wh = xpoint(it.p2w, it.p);
            if (uniform(T->rng) < cfg->important_path_weight) {
                wo = important_sample(T, wh);
                so = V(w2s[0][0] * wo.x + w2s[0][1] * wo.y + w2s[0][2] * wo.z, w2s[1][0] * wo.x + w2s[1][1] * wo.y + w2s[1][2] * wo.z,
                       w2s[2][0] * wo.x + w2s[2][1] * wo.y + w2s[2][2] * wo.z);
            } else {
                so = rough ? rough_sample(T->rng, si, rgh) : hemisphere_cosine(T->rng);
                wo = V(s2w[0][0] * so.x + s2w[0][1] * so.y + s2w[0][2] * so.z, s2w[1][0] * so.x + s2w[1][1] * so.y + s2w[1][2] * so.z,
                       s2w[2][0] * so.x + s2w[2][1] * so.y + s2w[2][2] * so.z);
            }
            double pi = important_pdf(T, wh, wo), pb = rough ? rough_pdf(si, so, rgh) : (so.z >= 0.0 ? M_1_PI * so.z : 0.0);
            pdf = cfg->important_path_weight * pi + (1 - cfg->important_path_weight) * pb;
        } else {
            so = rough ? rough_sample(T->rng, si, rgh) : hemisphere_cosine(T->rng);
            pdf = rough ? rough_pdf(si, so, rgh) : (so.z >= 0.0 ? M_1_PI * so.z : 0.0);
        }
        double pc = so.z >= 0.0 ? M_1_PI * so.z : 0.0;
        if (rough) {                                                          /* evaluate_shading, conductor.pyx:249-289 */
            if (so.z > 0 && si.z != 0) {
                v3 h = norm3(V(si.x + so.x, si.y + so.y, si.z + so.z));
                ray_t dr;
                dr.o = refl_o;
                dr.d = V(s2w[0][0] * so.x + s2w[0][1] * so.y + s2w[0][2] * so.z, s2w[1][0] * so.x + s2w[1][1] * so.y + s2w[1][2] * so.z,
                         s2w[2][0] * so.x + s2w[2][1] * so.y + s2w[2][2] * so.z);
                dr.maxd = ray.maxd;
                T->rays += 1;
                trace(T, dr, depth + 1, spectrum);
                double f = rough_d(h, rgh) * (rough_g1(si, rgh) * rough_g1(so, rgh)) / (4 * si.z);
                for (int i = 0; i < bins; ++i) spectrum[i] *= f;
                double ci = dot3(h, so);                                      /* _f, conductor.pyx:312-328 */
                const double* kk = T->sp->tables + (size_t)T->sp->table2[mat] * bins;
                for (int i = 0; i < bins; ++i) {
                    double nn = table[i], k = kk[i];
                    double ci2 = ci * ci, k0 = nn * nn + k * k, k1 = k0 * ci2 + 1, k2 = 2 * nn * ci, k3 = k0 + ci2;
                    spectrum[i] *= 0.5 * ((k1 - k2) / (k1 + k2) + (k3 - k2) / (k3 + k2));
                }
            }
        } else if (pc != 0.0) {
            ray_t dr;
            dr.o = refl_o;
            dr.d = V(s2w[0][0] * so.x + s2w[0][1] * so.y + s2w[0][2] * so.z, s2w[1][0] * so.x + s2w[1][1] * so.y + s2w[1][2] * so.z,
                     s2w[2][0] * so.x + s2w[2][1] * so.y + s2w[2][2] * so.z);
            dr.maxd = ray.maxd;
            T->rays += 1;
            trace(T, dr, depth + 1, spectrum);
            for (int i = 0; i < bins; ++i) spectrum[i] *= table[i];
            for (int i = 0; i < bins; ++i) spectrum[i] *= pc;
        }
        double rc = 1.0 / pdf;
        for (int i = 0; i < bins; ++i) spectrum[i] *= rc;
    } else if (mtype == RSB_MAT_DIELECTRIC) {                                /* dielectric.pyx:153-308 */
        v3 inc = norm3(xvec(it.w2p, ray.d)), n = norm3(it.n);
        double c1 = -dot3(n, inc), n1, n2;
        if (c1 < 0.0) { n1 = T->sp->index_in[mat]; n2 = T->sp->index_out[mat]; } else { n1 = T->sp->index_out[mat]; n2 = T->sp->index_in[mat]; }
        double gamma = n1 / n2, c2s = 1 - (gamma * gamma) * (1 - c1 * c1);
        int reflect, dead = 0;
        v3 tr = V(0, 0, 0);
        if (c2s <= 0) { if (d->mat_transmission_only[mat]) dead = 1; reflect = 1; }
        else {
            double tmp = (c1 < 0.0) ? gamma * c1 + sqrt(c2s) : gamma * c1 - sqrt(c2s);
            tr = V(gamma * inc.x + tmp * n.x, gamma * inc.y + tmp * n.y, gamma * inc.z + tmp * n.z);
            double ci = c1, ct = -dot3(n, tr);
            double ra = (n1 * ci - n2 * ct) / (n1 * ci + n2 * ct), rb = (n1 * ct - n2 * ci) / (n1 * ct + n2 * ci);
            double transmission = 1 - 0.5 * (ra * ra + rb * rb);
            reflect = !(d->mat_transmission_only[mat] || uniform(T->rng) < transmission);
        }
        if (!dead) {
            ray_t dr;
            dr.maxd = ray.maxd;
            if (reflect) {
                double tmp = 2 * c1;
                dr.d = xvec(it.p2w, V(inc.x + tmp * n.x, inc.y + tmp * n.y, inc.z + tmp * n.z));
                dr.o = (c1 < 0.0) ? xpoint(it.p2w, it.pin) : xpoint(it.p2w, it.pout);
            } else {
                dr.d = xvec(it.p2w, tr);
                dr.o = (c1 < 0.0) ? xpoint(it.p2w, it.pout) : xpoint(it.p2w, it.pin);
            }
            T->rays += 1;
            trace(T, dr, depth + 1, spectrum);
        }
    }
    else if (mtype == RSB_MAT_VOLUME_EMITTER) {                            /* NullSurface.evaluate_surface, material.pyx:126-147 */
        ray_t dr;
        dr.maxd = ray.maxd;
        dr.d = ray.d;
        dr.o = it.exiting ? xpoint(it.p2w, it.pout) : xpoint(it.p2w, it.pin);
        T->rays += 1;
        trace_ka(T, dr, depth, spectrum, 1);                                 /* daughter.depth -= 1; keep_alive=True */
    }
    else if (mtype == RSB_MAT_CONDUCTOR) {                                 /* conductor.pyx:75-147 */
        v3 inc = norm3(xvec(it.w2p, ray.d)), n = norm3(it.n);
        double ci = dot3(n, inc), tmp = 2 * ci;
        ray_t dr;
        dr.maxd = ray.maxd;
        dr.d = xvec(it.p2w, V(inc.x - tmp * n.x, inc.y - tmp * n.y, inc.z - tmp * n.z));
        dr.o = (ci > 0.0) ? xpoint(it.p2w, it.pin) : xpoint(it.p2w, it.pout);   /* the side comes from ci, not from `exiting` */
        T->rays += 1;
        trace(T, dr, depth + 1, spectrum);
        ci = fabs(ci);
        const double* kk = T->sp->tables + (size_t)T->sp->table2[mat] * bins;     /* extinction k; `table` is the index n */
        for (int i = 0; i < bins; ++i) {                                          /* _fresnel, conductor.pyx:132-145 */
            double nn = table[i], k = kk[i];
            double ci2 = ci * ci, k0 = nn * nn + k * k, k1 = k0 * ci2 + 1, k2 = 2 * nn * ci, k3 = k0 + ci2;
            spectrum[i] *= 0.5 * ((k1 - k2) / (k1 + k2) + (k3 - k2) / (k3 + k2));
        }
    }
    /* _sample_volumes, ray.pyx:422-455 */
    int32_t inside[16];
    int n_in = world_contains(T->c, ray.o, inside, 16);
    if (n_in > 0) {
        v3 start = xpoint(it.p2w, it.p);
        for (int k = 0; k < n_in && k < 16; ++k) {
            int m2 = d->prim_material[inside[k]];
            if (d->mat_type[m2] == RSB_MAT_VOLUME_EMITTER) {                  /* homogeneous.pyx:66-91, uniform.pyx:129-131 */
                const double* w2p = d->prim_to_local + 13 * (size_t)inside[k];
                v3 ls = xpoint(w2p, start), le = xpoint(w2p, ray.o);
                double len = len3(V(ls.x - le.x, ls.y - le.y, ls.z - le.z));   /* end.vector_to(start).get_length() */
                if (len == 0) continue;
                const double* em = T->sp->tables + (size_t)m2 * bins;
                for (int i = 0; i < bins; ++i) {
                    double e = 0.0 + em[i] * T->sp->scale[m2];
                    spectrum[i] += e * len;
                }
                continue;
            }
            if (d->mat_type[m2] != RSB_MAT_DIELECTRIC) continue;
            double length = len3(V(ray.o.x - start.x, ray.o.y - start.y, ray.o.z - start.z));
            const double* tt = T->sp->tables + (size_t)m2 * bins;
            for (int i = 0; i < bins; ++i) spectrum[i] *= pow(tt[i], length);   /* dielectric.pyx:313-330 */
        }
    }
    for (int i = 0; i < bins; ++i) spectrum[i] *= normalisation;
}

/* ------------------------------------------------------------------ exported entry points */
static scene_t* scene_new(const RsbSceneDesc* d) {
    scene_t* s = (scene_t*)calloc(1, sizeof(scene_t));
    s->d = d;
    parse_tree(d->world_kdtree, d->world_kdtree_bytes, &s->world);
    s->meshes = (mesh_t*)calloc((size_t)(d->n_meshes > 0 ? d->n_meshes : 1), sizeof(mesh_t));
    for (int m = 0; m < d->n_meshes; ++m) {
        mesh_t* ms = &s->meshes[m];
        ms->d = &d->meshes[m];
        parse_tree(ms->d->kdtree, ms->d->kdtree_bytes, &ms->tree);
        ms->fnorm = (float*)malloc(sizeof(float) * 3 * (size_t)ms->d->n_triangles);
        for (int t = 0; t < ms->d->n_triangles; ++t) {                       /* mesh.pyx:428-462 */
            const int32_t* row = ms->d->triangles + (size_t)t * ms->d->tri_stride;
            v3 p1 = V(vtx(ms, row[0], 0), vtx(ms, row[0], 1), vtx(ms, row[0], 2));
            v3 p2 = V(vtx(ms, row[1], 0), vtx(ms, row[1], 1), vtx(ms, row[1], 2));
            v3 p3 = V(vtx(ms, row[2], 0), vtx(ms, row[2], 1), vtx(ms, row[2], 2));
            v3 n = norm3(cross3(V(p2.x - p1.x, p2.y - p1.y, p2.z - p1.z), V(p3.x - p1.x, p3.y - p1.y, p3.z - p1.z)));
            ms->fnorm[3 * t] = (float)n.x; ms->fnorm[3 * t + 1] = (float)n.y; ms->fnorm[3 * t + 2] = (float)n.z;
        }
    }
    s->imp_total = 0;
    if (d->n_important > 0) {                                                /* world.pyx:88-132 */
        s->cdf = (double*)malloc(sizeof(double) * (size_t)d->n_important);
        for (int i = 0; i < d->n_important; ++i) s->imp_total += d->imp_weight[i];
        for (int i = 0; i < d->n_important; ++i) s->cdf[i] = (i == 0) ? d->imp_weight[0] : s->cdf[i - 1] + d->imp_weight[i];
        for (int i = 0; i < d->n_important; ++i) s->cdf[i] /= s->imp_total;
    }
    return s;
}
static void scene_free(scene_t* s) {
    free(s->world.nodes); free(s->world.item_store);
    for (int m = 0; m < s->d->n_meshes; ++m) { free(s->meshes[m].tree.nodes); free(s->meshes[m].tree.item_store); free(s->meshes[m].fnorm); }
    free(s->meshes); free(s->cdf); free(s);
}

int ro_uniform(uint64_t seed, int64_t n, double* out) {
    mt_t g;
    rs_seed(&g, seed);
    for (int64_t i = 0; i < n; ++i) out[i] = uniform(&g);
    return 0;
}

int ro_hit(const RsbSceneDesc* d, int64_t n, const double* o, const double* dir, const double* maxd, int32_t* prim, double* t,
           int32_t* sub, uint8_t* exiting, double* geom, float* uvw, int32_t* node /* [n][2] world leaf, mesh leaf; or NULL */) {
    scene_t* s = scene_new(d);
    ctx_t c;
    c.s = s;
    c.st = (prim_state*)calloc((size_t)d->n_primitives, sizeof(prim_state));
    for (int64_t i = 0; i < n; ++i) {
        ray_t r;
        r.o = V(o[3 * i], o[3 * i + 1], o[3 * i + 2]); r.d = V(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
        r.maxd = maxd ? maxd[i] : INFINITY;
        isect it;
        if (world_hit(&c, &r, &it)) {
            prim[i] = it.prim; t[i] = it.t; sub[i] = it.tri; exiting[i] = (uint8_t)(it.exiting ? 1 : 0);
            double* g = geom + 12 * i;
            g[0] = it.p.x; g[1] = it.p.y; g[2] = it.p.z; g[3] = it.pin.x; g[4] = it.pin.y; g[5] = it.pin.z;
            g[6] = it.pout.x; g[7] = it.pout.y; g[8] = it.pout.z; g[9] = it.n.x; g[10] = it.n.y; g[11] = it.n.z;
            uvw[3 * i] = it.u; uvw[3 * i + 1] = it.v; uvw[3 * i + 2] = it.w;
            if (node) { node[2 * i] = c.hit_leaf; node[2 * i + 1] = d->prim_type[it.prim] == RSB_PRIM_MESH ? it.mesh_leaf : -1; }
        } else {
            if (node) { node[2 * i] = -1; node[2 * i + 1] = -1; }
            prim[i] = -1; t[i] = INFINITY; sub[i] = -1; exiting[i] = 0;
            memset(geom + 12 * i, 0, 96); uvw[3 * i] = uvw[3 * i + 1] = uvw[3 * i + 2] = 0;
        }
    }
    free(c.st);
    scene_free(s);
    return 0;
}

int ro_contains(const RsbSceneDesc* d, int64_t n, const double* p, int32_t cap, int32_t* count, int32_t* prims) {
    scene_t* s = scene_new(d);
    ctx_t c;
    c.s = s;
    c.st = (prim_state*)calloc((size_t)d->n_primitives, sizeof(prim_state));
    for (int64_t i = 0; i < n; ++i) {
        for (int k = 0; k < cap; ++k) prims[i * cap + k] = -1;
        count[i] = world_contains(&c, V(p[3 * i], p[3 * i + 1], p[3 * i + 2]), prims + i * cap, cap);
    }
    free(c.st);
    scene_free(s);
    return 0;
}

/* Observer._render_pixel (observer.pyx:363-419) for the listed pixels (NULL = all), MT19937-64 streams seeded
 * per pixel with seed + y*nx + x; PinholeCamera._generate_rays (pinhole.pyx:169-204); Welford (statsarray.pyx:743-777) */
/* StatsArray3D.combine_samples (raysect/core/math/statsarray.pyx:612-670) over a whole frame: merges n_new samples
 * with statistics (mean, variance) into the stored (fmean, fvar, fsamples), element by element, through
 * _combine_samples (:780-857): larger set first, Chan's pooled formula when both sets hold more than one sample,
 * else the special cases; a single new sample goes through _add_sample (:743-777).  This is what
 * SpectralPowerPipeline2D.update does with every pixel of an observe() pass (power.pyx:424-437). */
static void add_one(double sample, double* m, double* v, int* n) {
    if (*n == 0) { *n = 1; *m = sample; *v = 0; return; }
    double pm = *m, pv = *v;
    int pn = *n > 1 ? *n : 2;
    *n += 1;
    *m = pm + (sample - pm) / *n;
    *v = (pv * (pn - 1) + (sample - pm) * (sample - *m)) / (*n - 1);
}

int ro_combine(int64_t n, const double* mean, const double* variance, int32_t n_new, double* fmean, double* fvar, int32_t* fsamples) {
    if (n_new < 1) return 1;
    for (int64_t i = 0; i < n; ++i) {
        double mb = mean[i], vb = variance[i] < 0 ? 0 : variance[i];
        double mx = fmean[i], vx = fvar[i], my = mb, vy = vb, mt = 0, vt = 0;
        int nx = fsamples[i], ny = n_new, nt = 0;
        if (nx < ny) {
            int ti = nx; nx = ny; ny = ti;
            double td = mx; mx = my; my = td;
            td = vx; vx = vy; vy = td;
        }
        if (nx > 1 && ny > 1) {
            nt = nx + ny;
            mt = (nx * mx + ny * my) / (double)nt;
            vx = (nx - 1) * vx / (double)nx;
            vy = (ny - 1) * vy / (double)ny;
            vt = (nx * (mx * mx + vx) + ny * (my * my + vy)) / (double)nt - mt * mt;
            vt = nt * vt / (double)(nt - 1);
        } else if (nx == 0 && ny == 0) {
            nt = 0; mt = 0; vt = 0;
        } else if (nx == 1) {
            if (ny == 0) { nt = 1; mt = mx; vt = 0; }
            else {
                nt = 2;
                mt = 0.5 * (mx + my);
                double temp = mx - mt;
                vt = 2 * temp * temp;
            }
        } else if (nx > 1) {
            nt = nx; mt = mx; vt = vx;
            if (ny == 1) add_one(my, &mt, &vt, &nt);
        }
        fmean[i] = mt; fvar[i] = vt; fsamples[i] = nt;
    }
    return 0;
}

int ro_render(const RsbSceneDesc* d, const RsbCamera* cam, const RsbRayConfig* cfg, const RsbSpectral* sp, uint64_t seed,
              int64_t n_pixels, const int32_t* pixels, double* mean, double* variance, uint64_t* ray_count) {
    scene_t* s = scene_new(d);
    ctx_t c;
    c.s = s;
    c.st = (prim_state*)calloc((size_t)d->n_primitives, sizeof(prim_state));
    int bins = cfg->bins, spp = cam->pixel_samples;
    double* spectrum = (double*)malloc(sizeof(double) * (size_t)bins);
    double* jit = (double*)malloc(sizeof(double) * 2 * (size_t)spp);
    mt_t g;
    tracer T;
    T.c = &c; T.cfg = cfg; T.sp = sp; T.rng = &g; T.rays = 0;
    if (!pixels) n_pixels = (int64_t)cam->nx * cam->ny;
    for (int64_t w = 0; w < n_pixels; ++w) {
        int px = pixels ? pixels[2 * w] : (int)(w / cam->ny), py = pixels ? pixels[2 * w + 1] : (int)(w % cam->ny);
        rs_seed(&g, seed + (uint64_t)((int64_t)py * cam->nx + px));
        for (int k = 0; k < 2 * spp; ++k) jit[k] = uniform(&g);             /* RectangleSampler3D.samples(n), surface3d.pyx:94-110 */
        double pixel_x = cam->image_start_x - cam->image_delta * (px + 0.5);
        double pixel_y = cam->image_start_y - cam->image_delta * (py + 0.5);
        double half = 0.5 * cam->image_delta;
        double* m = mean + ((size_t)px * cam->ny + py) * bins;
        double* v = variance + ((size_t)px * cam->ny + py) * bins;
        for (int k = 0; k < spp; ++k) {
            /* C argument evaluation order of new_point3d(uniform()*w - ow, uniform()*h - oh, 0): right to left */
            double jy = jit[2 * k] * cam->image_delta - half, jx = jit[2 * k + 1] * cam->image_delta - half;
            v3 dir = norm3(V(jx + pixel_x, jy + pixel_y, 0.0 + 1.0));
            ray_t r;
            r.o = xpoint(cam->to_root, V(0, 0, 0)); r.d = xvec(cam->to_root, dir); r.maxd = cfg->max_distance;
            if (cam->kind == RSB_CAMERA_ORTHOGRAPHIC) {                     /* orthographic.pyx:139-167 */
                const double p2l[13] = {1, 0, 0, pixel_x, 0, 1, 0, pixel_y, 0, 0, 1, 0, 1};   /* translate(pixel_x, pixel_y, 0) */
                dir = V(0, 0, 1);                                           /* dir.z = 1: projection weight 1 */
                r.o = xpoint(cam->to_root, xpoint(p2l, V(jx, jy, 0)));
                r.d = xvec(cam->to_root, dir);
            }
            T.rays += 1;
            trace(&T, r, 0, spectrum);
            for (int i = 0; i < bins; ++i) {
                double x = spectrum[i] * dir.z;
                x = x * cam->sensitivity;
                if (k == 0) { m[i] = x; v[i] = 0; }
                else {
                    double pm = m[i], pv = v[i];
                    int pn = k > 1 ? k : 2, nn = k + 1;
                    m[i] = pm + (x - pm) / nn;
                    v[i] = (pv * (pn - 1) + (x - pm) * (x - m[i])) / (nn - 1);
                }
            }
        }
    }
    *ray_count += T.rays;
    free(spectrum); free(jit); free(c.st);
    scene_free(s);
    return 0;
}
