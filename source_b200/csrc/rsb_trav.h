// rsb_trav.h -- the kd-tree walk as ONE NODE VISIT per call, and Mesh.hit as a stand-alone query.
//
// kd_descend / kd_advance (rsb_geom.h) run "down to the next leaf, test it, pop" as one unit.  On the device that unit
// has a very uneven length -- a ray's first descent of a 1M-triangle tree is ~25 levels, every later one 1-5, and 40 %
// of the leaves reached are empty -- so a warp that advances all its lanes by one unit per trip runs every trip at the
// length of its longest lane (round-1 profile of the mesh sweep: 8.7 of 32 lanes active).  kd_visit is the same walk
// cut into single node visits: a branch step, or a leaf arrival (an EMPTY leaf is popped on the spot), so kernels can
// give every lane a bounded number of visits per trip and collect the lanes that stand at a non-empty leaf.
// Per ray the sequence of nodes, plane distances, comparisons, pushes and pops is exactly that of
// KDTree3DCore._trace_branch (kdtree3d.pyx:626-700) -- it is kd_descend's loop body and kd_advance's pop.
//
// Compiles for the device and for the host (the CPU test-suite pins it against the reference's goldens without a GPU).
#pragma once
#include "rsb_geom.h"

namespace rsb {

enum KdVisit : int32_t { VISIT_MORE = 0, VISIT_LEAF = 1, VISIT_DONE = 2 };

// far-child stack over a plain array (host code, slow paths)
struct LocalKdStack {
    KdStackEntry* e;
    RSB_HD void push(int sp, int node, double tmax) { e[sp].node = node; e[sp].tmax = tmax; }
    RSB_HD int node(int sp) const { return e[sp].node; }
    RSB_HD double tmax(int sp) const { return e[sp].tmax; }
};

// the far child resumes with min_range = the plane distance = max_range of the leaf just left (kd_advance)
template <class Stack>
RSB_HD bool kd_pop(const Stack& stk, int& node, int& sp, double& tmin, double& tmax) {
    if (sp == 0) return false;
    --sp;
    node = stk.node(sp);
    tmin = tmax;
    tmax = stk.tmax(sp);
    return true;
}

// One node visit.  VISIT_LEAF: `node` is a non-empty leaf (off, cnt), to be tested with max_range = tmax; the caller
// pops (kd_pop) when the leaf yields no hit.  VISIT_DONE: the walk ended without a hit.
template <int S, class Stack, class Stats>
RSB_HD int kd_visit(const KdNode* nodes, const RayAx<S>& ax, Stack& stk, int& node, int& sp, double& tmin, double& tmax, int& off,
                    int& cnt, Stats& stats) {
    const KdNode n = kd_load_node(nodes + node);
    if (n.axis >= 0) {
        stats.branch();
        const int axis = n.axis;
        const double origin = ax.o(axis), direction = ax.d(axis);
#ifdef RSB_KD_TRUE_DIVIDE
        const double plane_distance = (n.split - origin) / direction;
#else
        const double plane_distance = div_recip1(n.split - origin, direction, ax.r(axis), ax.unsafe != 0);
#endif
        const bool below_split = origin < n.split || (origin == n.split && direction < 0);
        const int lower_id = node + 1, upper_id = n.upper;
        const int near_id = below_split ? lower_id : upper_id;
        const int far_id = below_split ? upper_id : lower_id;
        const bool only_near = direction == 0 || plane_distance > tmax || plane_distance <= 0;
        const bool only_far = !only_near && plane_distance < tmin;
        if (!only_near && !only_far) {
#if defined(__CUDA_ARCH__) && !defined(RSB_NO_PREFETCH_FAR)
            // the far child is visited after the whole near subtree: start its trip from L2 now
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + far_id));
#endif
            stk.push(sp, far_id, tmax);
            ++sp;
            tmax = plane_distance;
        }
        node = only_far ? far_id : near_id;
        return VISIT_MORE;
    }
    stats.leaf(n.leaf.item_count);
    if (n.leaf.item_count > 0) {
        off = n.leaf.item_offset;
        cnt = n.leaf.item_count;
        return VISIT_LEAF;
    }
    return kd_pop(stk, node, sp, tmin, tmax) ? VISIT_MORE : VISIT_DONE;
}

// Mesh.hit (mesh.pyx:1255-1279 -> MeshData.trace :506-518) of world-level primitive row `prim` for a world-space ray,
// written over kd_visit: the sequential form of what k_rq_mesh does with pooled triangle tests.
template <class Stats>
RSB_HD bool mesh_query(const Scene& sc, int prim, const V3& o, const V3& d, double max_distance, KdStackEntry* stack, MeshHit* out,
                       Stats& stats) {
    const Prim& p = sc.prims[prim];
    const Mesh& mesh = sc.meshes[p.mesh];
    const V3 lo = xform_point(p.to_local, o);
    const V3 ld = xform_vector(p.to_local, d);
    double axbuf[9];
    RayAx<1> ax;
    ax.set(axbuf, lo, ld);
    MeshLeaf<Stats> leaf;
    leaf.mesh = &mesh;
    leaf.o = lo;
    leaf.max_distance = max_distance;
    leaf.rs = mesh_rayspace(ld);
    leaf.result = out;
    leaf.stats = &stats;
    KdCursor c;
    if (!kd_begin(mesh.tree, ax, c)) return false;
    LocalKdStack stk;
    stk.e = stack;
    int node = 0, sp = 0, off = 0, cnt = 0;
    double tmin = c.min_range, tmax = c.max_range;
    for (;;) {
        const int r = kd_visit(mesh.tree.nodes, ax, stk, node, sp, tmin, tmax, off, cnt, stats);
        if (r == VISIT_DONE) return false;
        if (r == VISIT_LEAF) {
            if (leaf(off, cnt, tmax)) { out->node = node; return true; }
            if (!kd_pop(stk, node, sp, tmin, tmax)) return false;
        }
    }
}

// World.hit through the split pipeline, run serially: the world-level walk stops in front of every Mesh.hit, the
// query is answered by mesh_query, the walk resumes (NestedTraversal::begin_t<true> / resume_split).  The device runs
// the three steps as separate kernels (rsb_trav.cuh); this is the same logic for host-side parity tests.
template <int FEAT, class Stats>
RSB_HD bool world_hit_split(const Scene& sc, const V3& o, const V3& d, double max_distance, KdStackEntry* stack, HitRec* rec, Stats& stats) {
    double axbuf[RSB_AX_WORDS];
    NestedTraversal<FEAT, 1, Stats> t;
    t.init(sc, max_distance, stack, rec, stats, axbuf);
    bool alive = t.template begin_t<true>(o, d);
    while (alive) {
        MeshHit mh;
        mh.t = 0.0; mh.tri = -1; mh.node = -1; mh.u = mh.v = mh.w = 0.0f;
        const bool hit = mesh_query(sc, t.cand[t.ci], o, d, max_distance, stack + (RSB_KD_STACK / 2), &mh, stats);
        alive = t.resume_split(hit, mh);
    }
    return t.finish();
}

// World.hit of a mesh-free scene over kd_visit (what k_rq_world runs, one lane)
template <int FEAT, class Stats>
RSB_HD bool world_hit_visits(const Scene& sc, const V3& o, const V3& d, double max_distance, KdStackEntry* stack, HitRec* rec, Stats& stats) {
    double axbuf[RSB_AX_WORDS];
    WorldLeaf<Stats, FEAT, 1> leaf;
    leaf.sc = &sc;
    leaf.ax.set(axbuf, o, d);
    leaf.max_distance = max_distance;
    leaf.mesh_stack = stack + (RSB_KD_STACK / 2);
    leaf.mesh_axbuf = axbuf + 9;
    leaf.best = rec;
    leaf.stats = &stats;
    rec->u = rec->v = rec->w = 0.0f;
    rec->node = -1;
    rec->mesh_node = -1;
    KdCursor c;
    if (!kd_begin(sc.world, leaf.ax, c)) return false;
    LocalKdStack stk;
    stk.e = stack;
    int node = 0, sp = 0, off = 0, cnt = 0;
    double tmin = c.min_range, tmax = c.max_range;
    for (;;) {
        const int r = kd_visit(sc.world.nodes, leaf.ax, stk, node, sp, tmin, tmax, off, cnt, stats);
        if (r == VISIT_DONE) return false;
        if (r == VISIT_LEAF) {
            if (leaf(off, cnt, tmax)) { rec->node = node; return true; }
            if (!kd_pop(stk, node, sp, tmin, tmax)) return false;
        }
    }
}

}  // namespace rsb
