// rsb_trav.cuh -- World.hit over ray arrays as a pipeline of small kernels (sm_100a).
//
// Round 1 ran the whole two-level query (world kd-tree -> BoundPrimitive -> Mesh.hit -> mesh kd-tree -> triangles) in
// one kernel: 118 registers (4 CTAs of 128 threads per SM), a 1 KB per-thread stack in local memory, and every lane of
// a warp at a different point of the hierarchy (profiles/r1b_ncu_full_k_hit_sweep_mesh.txt: 8.7 of 32 lanes active,
// 23 % of the warps an SM can hold, long-scoreboard stalls 54 %).  Here the query is cut where the reference's call
// graph is cut:
//
//   k_rq_begin    the world-level walk (world tree, AABB pre-tests, analytic / CSG primitives) up to the first
//                 Mesh.hit call; rays that meet none finish here, the others are PARKED (walk state -> HBM) and queued
//   k_rq_mesh     Mesh.hit for the queue: nothing but the mesh kd-tree walk and the triangle tests, persistent lanes
//                 refilled from the queue, far-child stack in shared memory, a bounded number of node visits per trip
//                 (rsb_trav.h kd_visit), triangle tests of all lanes standing at a leaf pooled over the warp
//   k_rq_resume   parked walks carry on with Mesh.hit's answer (kept as the ray's memo: the reference repeats the very
//                 same Mesh.hit call from every world leaf that lists the mesh)
//   k_rq_world    mesh-free scenes: the world walk in the same visit-per-trip form, one kernel
//
// Per ray the sequence of node visits, AABB tests, primitive tests and triangle tests is the reference's
// (kdtree3d.pyx:589-700, acceleration/kdtree.pyx:73-122, mesh.pyx:506-713); rsb_trav.h holds the serial statement of
// the same pipeline, which the CPU suite pins against the reference's goldens.
#pragma once
#include "rsb_kernels.cuh"
#include "rsb_trav.h"

namespace rsb {

#define RQ_THREADS RSB_RENDER_THREADS
#ifndef RQ_SCAP
#define RQ_SCAP 10           // far-child stack entries kept in shared memory (deeper ones spill to local memory)
#endif
#ifndef RQ_VISITS
#define RQ_VISITS 6          // node visits per lane per trip
#endif
#ifndef RQ_LEAF_MIN
#define RQ_LEAF_MIN 8        // lanes standing at a leaf that trigger the leaf phase while other lanes still descend (k_rq_world)
#endif
#ifndef RQ_MESH_LEAF_MIN
#define RQ_MESH_LEAF_MIN 4   // the same for k_rq_mesh, whose pooled leaf phase is cheap to enter: 4 against 8 measured +3 % on the
                             // 1.3 M-triangle sweep (sorted 862 -> 890, morton 904 -> 923 Mrays/s) and -1.5 % in k_rq_world
#endif
#ifndef RQ_REFILL
#define RQ_REFILL 8          // idle lanes that trigger a refill from the queue
#endif
#ifndef RQ_MESH_BLOCKS
#define RQ_MESH_BLOCKS 7     // resident CTAs per SM the register allocation of k_rq_mesh must allow
#endif
#ifndef RQ_WORLD_BLOCKS
#define RQ_WORLD_BLOCKS 5
#endif
#ifndef RQ_WALK_BLOCKS
#define RQ_WALK_BLOCKS 4     // k_rq_walk (world-level walk of scenes with meshes)
#endif
#define RQ_WORLD_STACK (RSB_KD_STACK / 2)   // world-level far-child entries a parked walk carries: the world half of the stack
#define RQ_MAX_ROUNDS 2      // Mesh.hit rounds before the last resume finishes whatever is left in place

// A parked world-level walk (NestedTraversal in state ST_MESH, nothing of the mesh touched yet) + the Mesh.hit answer
struct __align__(16) RqSusp {
    double min_range, max_range;
    double distance, best_t;
    int32_t node, sp, leaf_node, item_offset;
    int32_t item_count, item_base, nc, ci;
    int32_t cand[4];
    int32_t flags, best_prim, best_leaf, best_code;         // flags: 1 have_leaf, 2 found
    int32_t best_flip, best_mesh_node, memo_prim, memo_hit;
    float best_u, best_v, best_w;
    int32_t pad0;
    double memo_t;                                          // written by k_rq_mesh from here on
    int32_t memo_tri, memo_node;
    float memo_u, memo_v, memo_w;
    int32_t pad1;
};

// Device arrays of one batch of queries (capacity `cap`)
struct RqBuf {
    const double* ray;        // [6][ray_stride] origin xyz, direction xyz
    const double* md;         // [n] ray.max_distance, or null: md_all
    double md_all;
    long long ray_stride;
    double* hit_t;            // [cap]
    int4* hit_a;              // [cap] primitive (-1: miss), leaf row, code, flip
    float4* hit_uvw;          // [cap] u, v, w, mesh kd leaf (int bits)
    int32_t* hit_node;        // [cap] world kd leaf, or null
    RqSusp* susp;             // [cap]
    KdStackEntry* susp_stack; // [cap][RQ_WORLD_STACK]
    int2* queue;              // [RQ_MAX_ROUNDS + 1][cap] (query, mesh primitive row)
    unsigned int* ctr;        // [0 .. RQ_MAX_ROUNDS] queue lengths, [4 ..] fetch cursors of the mesh rounds, [8] of k_rq_world
    long long cap;
    const int32_t* perm;      // queries were reordered (rq_reorder): slot j holds the caller's query perm[j]; null = as given
};

template <int CAP, int T>
struct SmemKdStack {
    double* t;             // the thread's column: entry k at t[k * T]
    int32_t* n;
    KdStackEntry* spill;   // entries CAP.. (local memory; touched by the rare deep walk only)
    __device__ __forceinline__ void push(int sp, int node, double tmax) {
        if (sp < CAP) { t[sp * T] = tmax; n[sp * T] = node; }
        else { spill[sp - CAP].tmax = tmax; spill[sp - CAP].node = node; }
    }
    __device__ __forceinline__ int node(int sp) const { return sp < CAP ? n[sp * T] : spill[sp - CAP].node; }
    __device__ __forceinline__ double tmax(int sp) const { return sp < CAP ? t[sp * T] : spill[sp - CAP].tmax; }
};

// ---- clients: where a query's ray comes from and where its answer goes -------------------------------------
// fetch: RQ_SKIP no query in this place, RQ_TRACE trace it, RQ_MISS report a miss without tracing
enum RqFetch : int32_t { RQ_SKIP = 0, RQ_TRACE = 1, RQ_MISS = 2 };

struct RqArrayClient {
    RqBuf b;
    __device__ __forceinline__ int fetch(long long i, V3& o, V3& d, double& md) const {
        const long long s = b.ray_stride;
        o = v3(b.ray[i], b.ray[s + i], b.ray[2 * s + i]);
        d = v3(b.ray[3 * s + i], b.ray[4 * s + i], b.ray[5 * s + i]);
        md = b.md ? b.md[i] : b.md_all;
        return RQ_TRACE;
    }
    // reached by all 32 lanes together
    __device__ __forceinline__ void commit(long long i, bool finished, bool hit, const HitRec& rec) const {
        if (!finished) return;
        if (hit) {
            b.hit_t[i] = rec.t;
            b.hit_a[i] = make_int4(rec.prim, rec.leaf, rec.code, rec.flip);
            b.hit_uvw[i] = make_float4(rec.u, rec.v, rec.w, __int_as_float(rec.mesh_node));
            if (b.hit_node) b.hit_node[i] = rec.node;
        } else {
            b.hit_a[i] = make_int4(-1, -1, -1, 0);
        }
    }
};

// The wavefront renderer's slots as queries (k_wf_trace's prologue and epilogue, rsb_kernels.cuh): query i = slot i;
// live paths whose Russian roulette ended them (norm == 0) go straight to the ended list, hits are filed under the
// material family of the primitive they hit.
struct RqWfClient {
    WfArgs a;
    __device__ __forceinline__ int fetch(long long i, V3& o, V3& d, double& md) const {
        if (i >= a.n_slots || a.st.status[i] != SLOT_ALIVE) return RQ_SKIP;
        if (a.st.norm[i] == 0.0) return RQ_MISS;
        const size_t P = (size_t)a.n_slots;
        o = v3(a.st.ray[i], a.st.ray[P + i], a.st.ray[2 * P + i]);
        d = v3(a.st.ray[3 * P + i], a.st.ray[4 * P + i], a.st.ray[5 * P + i]);
        md = a.cfg.max_distance;
        return RQ_TRACE;
    }
    // reached by all 32 lanes together
    __device__ __forceinline__ void commit(long long i, bool finished, bool hit, const HitRec& rec) const {
        int list = -1;
        const int slot = (int)i;
        if (finished) {
            if (hit) {
                a.st.hit_t[slot] = rec.t;
                a.st.hit_a[slot] = make_int4(rec.prim, rec.leaf, rec.code, rec.flip);
                a.st.hit_uvw[slot] = make_float4(rec.u, rec.v, rec.w, __int_as_float(rec.mesh_node));
                a.st.status[slot] = SLOT_HIT;
                list = a.sp.mats[a.sc.prims[rec.prim].material].type;
                if (list >= MAT_CONDUCTOR) list = (list == MAT_ROUGH_CONDUCTOR) ? MAT_LAMBERT : MAT_DIELECTRIC;
            } else {
                a.st.status[slot] = SLOT_ENDED_ZERO;
                list = 4;
            }
        }
        __syncwarp();
        wf_append_lists(a, list, slot);
    }
};

// ---- parking ----------------------------------------------------------------------------------------------
template <class NT>
__device__ __forceinline__ void rq_park(const NT& t, const HitRec& rec, const KdStackEntry* stack, RqSusp* out, KdStackEntry* out_stack) {
    RqSusp s;
    s.min_range = t.min_range; s.max_range = t.max_range;
    s.distance = t.distance; s.best_t = rec.t;
    s.node = t.node; s.sp = t.sp; s.leaf_node = t.leaf_node; s.item_offset = t.item_offset;
    s.item_count = t.item_count; s.item_base = t.item_base; s.nc = t.nc; s.ci = t.ci;
    s.cand[0] = t.cand[0]; s.cand[1] = t.cand[1]; s.cand[2] = t.cand[2]; s.cand[3] = t.cand[3];
    s.flags = (t.have_leaf ? 1 : 0) | (t.found ? 2 : 0);
    s.best_prim = rec.prim; s.best_leaf = rec.leaf; s.best_code = rec.code;
    s.best_flip = rec.flip; s.best_mesh_node = rec.mesh_node;
    s.memo_prim = t.cand[t.ci];      // the primitive k_rq_mesh answers for; the answer becomes the memo
    s.memo_hit = 0;
    s.best_u = rec.u; s.best_v = rec.v; s.best_w = rec.w;
    s.pad0 = 0;
    int4* dst = reinterpret_cast<int4*>(out);
    const int4* src = reinterpret_cast<const int4*>(&s);
#pragma unroll
    for (int k = 0; k < 8; ++k) dst[k] = src[k];
    for (int k = 0; k < t.sp; ++k) reinterpret_cast<int4*>(out_stack)[k] = reinterpret_cast<const int4*>(stack)[k];
}

template <class NT>
__device__ __forceinline__ void rq_unpark(NT& t, HitRec& rec, KdStackEntry* stack, const RqSusp* in, const KdStackEntry* in_stack, bool* memo_hit,
                                          MeshHit* memo) {
    RqSusp s;
    int4* dst = reinterpret_cast<int4*>(&s);
    const int4* src = reinterpret_cast<const int4*>(in);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(RqSusp) / 16); ++k) dst[k] = src[k];
    t.min_range = s.min_range; t.max_range = s.max_range;
    t.distance = s.distance; rec.t = s.best_t;
    t.node = s.node; t.sp = s.sp; t.leaf_node = s.leaf_node; t.item_offset = s.item_offset;
    t.item_count = s.item_count; t.item_base = s.item_base; t.nc = s.nc; t.ci = s.ci;
    t.cand[0] = s.cand[0]; t.cand[1] = s.cand[1]; t.cand[2] = s.cand[2]; t.cand[3] = s.cand[3];
    t.have_leaf = (s.flags & 1) != 0;
    t.found = (s.flags & 2) != 0;
    rec.prim = s.best_prim; rec.leaf = s.best_leaf; rec.code = s.best_code;
    rec.flip = s.best_flip; rec.mesh_node = s.best_mesh_node; rec.node = -1;
    rec.u = s.best_u; rec.v = s.best_v; rec.w = s.best_w;
    t.state = NT::ST_MESH;
    *memo_hit = s.memo_hit != 0;
    memo->t = s.memo_t; memo->tri = s.memo_tri; memo->node = s.memo_node;
    memo->u = s.memo_u; memo->v = s.memo_v; memo->w = s.memo_w;
    for (int k = 0; k < s.sp; ++k) reinterpret_cast<int4*>(stack)[k] = reinterpret_cast<const int4*>(in_stack)[k];
}

// Warp-aggregated append of the lanes with `want` to a queue: returns the lane's position (valid where want)
__device__ __forceinline__ unsigned int rq_queue_slot(unsigned int* counter, bool want) {
    const unsigned m = __ballot_sync(RSB_FULL_MASK, want);
    if (m == 0) return 0;
    const int lane = threadIdx.x & 31;
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned int)__popc(m));
    base = __shfl_sync(RSB_FULL_MASK, base, 0);
    return base + __popc(m & ((1u << lane) - 1));
}

// ---- k_rq_begin / k_rq_resume ---------------------------------------------------------------------------
// RESUME: walks parked by the previous round carry on with Mesh.hit's answer; LAST: a walk that meets yet another
// mesh is finished in place by the nested loop instead of being parked again.
template <bool COUNT, int FEAT, class Client, bool RESUME, bool LAST>
__global__ void __launch_bounds__(RQ_THREADS, (FEAT & RSB_FEAT_CSG) ? 3 : RQ_WALK_BLOCKS)
k_rq_walk(Scene sc, int n_items, Client cl, RqBuf b, long long n, int round, DevCounters* counters) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* axbuf = ax_storage<(FEAT & RSB_FEAT_STAGED) != 0>(smem, sc, n_items);
    stage_scene<(FEAT & RSB_FEAT_STAGED) != 0>(sc, smem, n_items);
    typedef typename StatsSel<COUNT>::type Stats;
    Stats stats;
    KdStackEntry stack[RSB_KD_STACK];
    HitRec rec;
    NestedTraversal<FEAT, RQ_THREADS, Stats> t;
    const long long total = RESUME ? (long long)b.ctr[round] : n;
    const int2* queue_in = b.queue + (size_t)round * b.cap;
    int2* queue_out = b.queue + (size_t)(RESUME ? round + 1 : 0) * b.cap;
    unsigned int* ctr_out = b.ctr + (RESUME ? round + 1 : 0);
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned long long traced = 0;
    for (long long k0 = (long long)blockIdx.x * blockDim.x; k0 < total; k0 += stride) {
        const long long k = k0 + threadIdx.x;
        long long i = k;
        bool valid = k < total;
        if (RESUME && valid) i = queue_in[k].x;
        V3 o, d;
        double md = RSB_INF;
        int what = RQ_SKIP;
        if (valid) what = cl.fetch(i, o, d, md);
        valid = what != RQ_SKIP;
        bool alive = false, hit = false;
        if (what == RQ_TRACE) {
            t.init(sc, md, stack, &rec, stats, axbuf);
            if (RESUME) {
                bool mhit;
                MeshHit memo;
                rq_unpark(t, rec, stack, b.susp + i, b.susp_stack + (size_t)i * RQ_WORLD_STACK, &mhit, &memo);
                t.leaf.ax.set(axbuf, o, d);
                alive = t.resume_split(mhit, memo);
            } else {
                traced = traced + 1;
                alive = t.template begin_t<true>(o, d);
            }
            if (LAST && alive) {
                alive = t.enter_mesh();
                while (alive) alive = t.step();
            }
            if (!alive) hit = t.finish();
        }
        const unsigned int q = rq_queue_slot(ctr_out, alive);
        if (alive) {
            rq_park(t, rec, stack, b.susp + i, b.susp_stack + (size_t)i * RQ_WORLD_STACK);
            queue_out[q] = make_int2((int)i, t.cand[t.ci]);
        }
        cl.commit(i, valid && !alive, hit, rec);
    }
    if (COUNT) {
        __syncwarp();
        traced = warp_sum(traced);
        if ((threadIdx.x & 31) == 0 && traced) atomicAdd(&counters->rays, traced);
        flush_stats(stats, counters);
    }
}

// ---- k_rq_mesh --------------------------------------------------------------------------------------------
struct MeshPool {
    int32_t* rs_pack;      // [T] ix | iy << 2 | iz << 4 of the lane's ray-space permutation (mesh.pyx:566-610)
    float* rs_s;           // [3][T] sx, sy, sz
    int32_t* mesh_idx;     // [T]
    int32_t* lane_of_rank; // warp: [32] k-th lane (in lane order) that holds a leaf
    const double* ax0;     // RayAx storage of thread 0: rows 0..2 mesh-local origin, row 9 ray.max_distance
};

#define RQ_AX_ROWS 10
#define RQ_POOL_WARP_BYTES 128
#define RQ_MESH_SMEM (RQ_AX_ROWS * 8 * RQ_THREADS + RQ_SCAP * 12 * RQ_THREADS + 20 * RQ_THREADS + (RQ_THREADS / 32) * RQ_POOL_WARP_BYTES)

// MeshData._trace_leaf (mesh.pyx:520-563) for every lane with `in_leaf`, the (ray, triangle) pairs of all those leaves
// dealt out evenly over the warp, 32 pairs per round, everything in registers and shuffles:
//   * pair g belongs to the lane whose range [excl, excl + cnt) of the prefix sum brackets it.  The ranges that START
//     inside the round's window are flagged in one mask (redux.or), so the rank of g's owner among the leaf-holding
//     lanes is a popcount, and a 32-entry table maps ranks to lanes;
//   * the tester fetches the triangle and runs _hit_triangle with the OWNER's mesh-local origin, max_distance and
//     ray-space shear (left in shared memory by the owner when it picked the ray up);
//   * a segmented min-scan over the lanes of one owner leaves the owner's closest triangle of the round in the last lane
//     of its segment -- on equal t the EARLIER pair, i.e. the first in leaf order -- and the owner applies the
//     reference's `t < distance` to it.  Taking the minimum first and comparing once accepts exactly the triangle the
//     sequential loop ends up with: the loop keeps the first of the smallest t below the starting distance.
// (Round 2's first version went through shared memory with per-owner loops for the pair list and for the result scan:
// 22 % of the kernel's warp instructions ran with ~3 of 32 lanes.)  All 32 lanes.
template <class Stats>
__device__ __forceinline__ bool mesh_leaf_pool(const Scene& sc, const MeshPool& cs, bool in_leaf, int off, int cnt, double d0, MeshHit* mh,
                                               Stats& stats) {
    constexpr int T = RQ_THREADS;
    const int lane = threadIdx.x & 31;
    const int tid0 = threadIdx.x & ~31;
    const int c = in_leaf ? cnt : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(RSB_FULL_MASK, incl, o);
        if (lane >= o) incl += v;
    }
    const int excl = incl - c;
    const int total = __shfl_sync(RSB_FULL_MASK, incl, 31);
    if (total == 0) return false;
    const unsigned hold = __ballot_sync(RSB_FULL_MASK, c > 0);
    __syncwarp();
    if (c > 0) cs.lane_of_rank[__popc(hold & ((1u << lane) - 1))] = lane;
    __syncwarp();
    const float INF = __int_as_float(0x7f800000);
    double distance = d0;
    int closest = -1;
    float cu = 0, cv = 0, cw = 0;
    for (int base = 0; base < total; base += 32) {
        const int g = base + lane;
        const bool valid = g < total;
        const int rel = excl - base;
        const unsigned heads = __reduce_or_sync(RSB_FULL_MASK, (c > 0 && rel >= 0 && rel < 32) ? (1u << rel) : 0u);
        const int n_before = __popc(__ballot_sync(RSB_FULL_MASK, c > 0 && rel < 0));
        const int rank = n_before + __popc(heads & (0xffffffffu >> (31 - lane))) - 1;
        const int owner = valid ? cs.lane_of_rank[rank] : lane;
        const int o_excl = __shfl_sync(RSB_FULL_MASK, excl, owner);
        const int o_off = __shfl_sync(RSB_FULL_MASK, off, owner);
        float t = INF, u = 0.f, v = 0.f, w = 0.f;
        int tri = -1;
        if (valid) {
            const int ot = tid0 + owner;           // the owner's thread index within the CTA
            const Mesh& m = sc.meshes[cs.mesh_idx[ot]];
            tri = m.tree.items[o_off + (g - o_excl)];
            const double* ax = cs.ax0 + ot;
            const V3 o = v3(ax[0], ax[T], ax[2 * T]);
            const double md = ax[9 * T];
            RaySpace rs;
            const int pk = cs.rs_pack[ot];
            rs.ix = pk & 3; rs.iy = (pk >> 2) & 3; rs.iz = (pk >> 4) & 3;
            rs.sx = cs.rs_s[ot]; rs.sy = cs.rs_s[T + ot]; rs.sz = cs.rs_s[2 * T + ot];
            float h[4];
            stats.tri_test();
            if (mesh_hit_triangle(m.tri + 3 * (size_t)tri, o, md, rs, h)) { t = h[3]; u = h[0]; v = h[1]; w = h[2]; }
        }
        // segmented inclusive min-scan: lanes of one owner are contiguous
        const int key = valid ? owner : -1 - lane;
        float ts = t;
        int src = lane;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
            const float t2 = __shfl_up_sync(RSB_FULL_MASK, ts, dlt);
            const int s2 = __shfl_up_sync(RSB_FULL_MASK, src, dlt);
            const int k2 = __shfl_up_sync(RSB_FULL_MASK, key, dlt);
            if (lane >= dlt && k2 == key && t2 <= ts) { ts = t2; src = s2; }
        }
        // the owner reads the last lane of its segment in this window
        int e = -1;
        if (c > 0 && rel < 32 && rel + c > 0) e = (rel + c < 32 ? rel + c : 32) - 1;
        const int from = e >= 0 ? e : lane;
        const float wt = __shfl_sync(RSB_FULL_MASK, ts, from);
        const int wsrc = __shfl_sync(RSB_FULL_MASK, src, from);
        const int wtri = __shfl_sync(RSB_FULL_MASK, tri, wsrc);
        const float wu = __shfl_sync(RSB_FULL_MASK, u, wsrc);
        const float wv = __shfl_sync(RSB_FULL_MASK, v, wsrc);
        const float ww = __shfl_sync(RSB_FULL_MASK, w, wsrc);
        if (e >= 0 && (double)wt < distance) {
            distance = (double)wt;
            closest = wtri;
            cu = wu; cv = wv; cw = ww;
        }
    }
    if (closest < 0) return false;
    mh->t = (double)(float)distance;
    mh->tri = closest;
    mh->u = cu; mh->v = cv; mh->w = cw;
    return true;
}

__device__ __forceinline__ void rq_answer(RqSusp* s, bool hit, const MeshHit& mh) {
    if (hit) {
        int4 a;
        a.x = __double2loint(mh.t); a.y = __double2hiint(mh.t); a.z = mh.tri; a.w = mh.node;
        reinterpret_cast<int4*>(&s->memo_t)[0] = a;
        reinterpret_cast<float4*>(&s->memo_u)[0] = make_float4(mh.u, mh.v, mh.w, 0.f);
        s->memo_hit = 1;
    } else {
        s->memo_hit = 0;
    }
}

// Mesh.hit for the queries of one round.  Persistent lanes: a lane picks the next query when its own is answered.  Every
// trip gives each travelling lane a bounded number of node visits, then the triangles of the leaves the lanes hold are
// tested in one pooled pass.  A lane that reaches a non-empty leaf does not wait for that pass: it PARKS the leaf
// (offset, count, search distance, node id) and walks on towards its next leaf as if the parked one held no hit -- true
// for ~90 % of the leaves a ray visits -- and only stops at the following leaf while one is still parked.  When the
// pooled pass finds a hit in the parked leaf the walk beyond it is simply dropped (MeshData.trace returns at the first
// leaf with a hit, kdtree3d.pyx:692-700): the answer is the same, a lane just never idles through other lanes' descents.
// (The counting build walks without looking ahead, so that its counters are exactly the reference's visits.)
template <bool COUNT>
__global__ void __launch_bounds__(RQ_THREADS, RQ_MESH_BLOCKS)
k_rq_mesh(Scene sc, RqBuf b, int round, DevCounters* counters) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int T = RQ_THREADS;
    constexpr bool AHEAD = !COUNT;
    typedef typename StatsSel<COUNT>::type Stats;
    Stats stats;
    const int tid = threadIdx.x, lane = tid & 31;
    double* ax0 = reinterpret_cast<double*>(smem);
    RayAx<T> ax;
    ax.p = ax0 + tid;
    ax.unsafe = 0;
    KdStackEntry spill[RSB_KD_STACK - RQ_SCAP];
    SmemKdStack<RQ_SCAP, T> stk;
    stk.t = ax0 + RQ_AX_ROWS * T + tid;
    stk.n = reinterpret_cast<int32_t*>(ax0 + (RQ_AX_ROWS + RQ_SCAP) * T) + tid;
    stk.spill = spill;
    MeshPool cs;
    {
        unsigned char* base = smem + RQ_AX_ROWS * 8 * T + RQ_SCAP * 12 * T;
        cs.rs_pack = reinterpret_cast<int32_t*>(base);
        cs.rs_s = reinterpret_cast<float*>(base + 4 * T);
        cs.mesh_idx = reinterpret_cast<int32_t*>(base + 16 * T);
        cs.lane_of_rank = reinterpret_cast<int32_t*>(base + 20 * T + (tid >> 5) * RQ_POOL_WARP_BYTES);
        cs.ax0 = ax0;
    }
    const unsigned int n_q = b.ctr[round];
    const int2* queue = b.queue + (size_t)round * b.cap;
    unsigned int* cursor = b.ctr + 4 + round;
    enum { IDLE = 0, DESCEND = 1, LEAF = 2, END = 3 };   // END: the walk ran out of nodes (a parked leaf may still be open)
    int st = IDLE, q = 0, node = 0, sp = 0, off = 0, cnt = 0;
    double tmin = 0.0, tmax = 0.0, md = 0.0;
    const KdNode* nodes = nullptr;
    bool parked = false;                                 // a leaf whose triangles are waiting for the pooled pass
    int p_off = 0, p_cnt = 0, p_node = 0;
    double p_d0 = 0.0;
    bool exhausted = n_q == 0;
    for (;;) {
        // ---- refill: idle lanes take the next queries, a batch at a time so that the set-up runs with a filled warp
        const unsigned idle = __ballot_sync(RSB_FULL_MASK, st == IDLE);
        if (idle == RSB_FULL_MASK && exhausted) break;
        if (!exhausted && (idle == RSB_FULL_MASK || __popc(idle) >= RQ_REFILL)) {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(cursor, (unsigned int)__popc(idle));
            base = __shfl_sync(RSB_FULL_MASK, base, 0);
            if (base + (unsigned int)__popc(idle) >= n_q) exhausted = true;
            const unsigned int k = base + __popc(idle & ((1u << lane) - 1));
            if (st == IDLE && k < n_q) {
                const int2 e = queue[k];
                q = e.x;
                const Prim& P = sc.prims[e.y];
                const long long s = b.ray_stride;
                const V3 o = v3(b.ray[q], b.ray[s + q], b.ray[2 * s + q]);
                const V3 d = v3(b.ray[3 * s + q], b.ray[4 * s + q], b.ray[5 * s + q]);
                md = b.md ? b.md[q] : b.md_all;
                // Mesh.hit: ray into mesh space, MeshData.trace's ray-space shear and bounds clip (mesh.pyx:506-518, 566-610)
                const V3 lo = xform_point(P.to_local, o);
                const V3 ld = xform_vector(P.to_local, d);
                ax.set(ax.p, lo, ld);
                ax.p[9 * T] = md;
                const Mesh& m = sc.meshes[P.mesh];
                const RaySpace rs = mesh_rayspace(ld);
                cs.rs_pack[tid] = rs.ix | (rs.iy << 2) | (rs.iz << 4);
                cs.rs_s[tid] = rs.sx; cs.rs_s[T + tid] = rs.sy; cs.rs_s[2 * T + tid] = rs.sz;
                cs.mesh_idx[tid] = P.mesh;
                KdCursor c;
                if (kd_begin(m.tree, ax, c)) {
                    nodes = m.tree.nodes;
                    node = 0; sp = 0;
                    tmin = c.min_range; tmax = c.max_range;
                    st = DESCEND;
                } else {
                    MeshHit none;
                    rq_answer(b.susp + q, false, none);
                }
            }
        }
        // ---- visit phase: a bounded number of node visits for every lane that is on its way down
#pragma unroll 1
        for (int v = 0; v < RQ_VISITS; ++v) {
            if (st == DESCEND) {
                const int r = kd_visit(nodes, ax, stk, node, sp, tmin, tmax, off, cnt, stats);
                if (r == VISIT_LEAF) st = LEAF;
                else if (r == VISIT_DONE) st = END;
            }
            if (st == LEAF && !parked) {
                parked = true;
                p_off = off; p_cnt = cnt; p_node = node;
                p_d0 = md < tmax ? md : tmax;            // min(ray.max_distance, max_range), mesh.pyx:535
                if (AHEAD) st = kd_pop(stk, node, sp, tmin, tmax) ? DESCEND : END;
            }
            if (!__any_sync(RSB_FULL_MASK, st == DESCEND)) break;
        }
        // ---- leaf phase: pooled triangle tests once enough lanes hold a leaf (or nobody can move)
        const unsigned holding = __ballot_sync(RSB_FULL_MASK, parked);
        const unsigned moving = __ballot_sync(RSB_FULL_MASK, st == DESCEND);
        if (holding != 0 && (__popc(holding) >= RQ_MESH_LEAF_MIN || moving == 0)) {
            MeshHit mh;
            const bool leaf_hit = mesh_leaf_pool(sc, cs, parked, p_off, p_cnt, p_d0, &mh, stats);
            if (parked) {
                parked = false;
                if (leaf_hit) {
                    mh.node = p_node;
                    rq_answer(b.susp + q, true, mh);
                    st = IDLE;
                } else if (!AHEAD) {
                    st = kd_pop(stk, node, sp, tmin, tmax) ? DESCEND : END;
                }
            }
        }
        if (st == END && !parked) {
            MeshHit none;
            rq_answer(b.susp + q, false, none);
            st = IDLE;
        }
    }
    if (COUNT) {
        __syncwarp();
        flush_stats(stats, counters);
    }
}

// ---- k_rq_world -------------------------------------------------------------------------------------------
// Mesh-free scenes: World.hit as visits per trip + a leaf phase (WorldLeaf: AABB pre-tests, then the primitive tests).
#define RQ_WORLD_SMEM (9 * 8 * RQ_THREADS + RQ_SCAP * 12 * RQ_THREADS)
template <bool COUNT, int FEAT, class Client>
__global__ void __launch_bounds__(RQ_THREADS, (FEAT & RSB_FEAT_CSG) ? 3 : RQ_WORLD_BLOCKS)
k_rq_world(Scene sc, int n_items, Client cl, RqBuf b, long long n, DevCounters* counters) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int T = RQ_THREADS;
    constexpr bool STAGED = (FEAT & RSB_FEAT_STAGED) != 0;
    typedef typename StatsSel<COUNT>::type Stats;
    Stats stats;
    const int tid = threadIdx.x, lane = tid & 31;
    double* axbuf = ax_storage<STAGED>(smem, sc, n_items);
    stage_scene<STAGED>(sc, smem, n_items);
    KdStackEntry spill[RSB_KD_STACK - RQ_SCAP];
    SmemKdStack<RQ_SCAP, T> stk;
    stk.t = axbuf + 9 * T;
    stk.n = reinterpret_cast<int32_t*>(axbuf - tid + (9 + RQ_SCAP) * T) + tid;
    stk.spill = spill;
    HitRec rec;
    WorldLeaf<Stats, FEAT & ~RSB_FEAT_MESH, T> leaf;
    leaf.sc = &sc;
    leaf.ax.p = axbuf;
    leaf.ax.unsafe = 0;
    leaf.max_distance = RSB_INF;
    leaf.mesh_stack = nullptr;
    leaf.mesh_axbuf = nullptr;
    leaf.best = &rec;
    leaf.stats = &stats;
    unsigned int* cursor = b.ctr + 8;
    // Two variations measured SLOWER on the 10,000-sphere field and left out: walking on past a parked leaf as k_rq_mesh does
    // (951 vs 1045 Mrays/s: a leaf of analytic primitives holds the ray's hit far more often than a mesh leaf, so the
    // look-ahead is mostly thrown away), and pooling the (ray, item) pairs of the leaf phase over the warp like the
    // triangle tests (917 vs 1045: 1.3 items per non-empty leaf do not pay for the pool's bookkeeping).
    constexpr bool AHEAD = false;
    enum { IDLE = 0, DESCEND = 1, LEAF = 2, END = 3 };
    int st = IDLE, node = 0, sp = 0, off = 0, cnt = 0;
    long long q = 0;
    double tmin = 0.0, tmax = 0.0;
    bool parked = false;
    int p_off = 0, p_cnt = 0, p_node = 0;
    double p_tmax = 0.0;
    bool exhausted = n <= 0;
    unsigned long long traced = 0;
    for (;;) {
        const unsigned idle = __ballot_sync(RSB_FULL_MASK, st == IDLE);
        if (idle == RSB_FULL_MASK && exhausted) break;
        if (!exhausted && (idle == RSB_FULL_MASK || __popc(idle) >= RQ_REFILL)) {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(cursor, (unsigned int)__popc(idle));
            base = __shfl_sync(RSB_FULL_MASK, base, 0);
            if ((long long)base + __popc(idle) >= n) exhausted = true;
            const long long k = (long long)base + __popc(idle & ((1u << lane) - 1));
            bool done = false;
            if (st == IDLE && k < n) {
                q = k;
                V3 o, d;
                double md;
                const int what = cl.fetch(q, o, d, md);
                if (what == RQ_MISS) done = true;
                if (what == RQ_TRACE) {
                    traced = traced + 1;
                    leaf.ax.set(axbuf, o, d);
                    leaf.max_distance = md;
                    rec.u = rec.v = rec.w = 0.0f;
                    rec.node = -1;
                    rec.mesh_node = -1;
                    KdCursor c;
                    if (kd_begin(sc.world, leaf.ax, c)) {
                        node = 0; sp = 0;
                        tmin = c.min_range; tmax = c.max_range;
                        st = DESCEND;
                    } else {
                        done = true;
                    }
                }
            }
            cl.commit(q, done, false, rec);
        }
#pragma unroll 1
        for (int v = 0; v < RQ_VISITS; ++v) {
            if (st == DESCEND) {
                const int r = kd_visit(sc.world.nodes, leaf.ax, stk, node, sp, tmin, tmax, off, cnt, stats);
                if (r == VISIT_LEAF) st = LEAF;
                else if (r == VISIT_DONE) st = END;
            }
            if (st == LEAF && !parked) {
                parked = true;
                p_off = off; p_cnt = cnt; p_node = node; p_tmax = tmax;
                if (AHEAD) st = kd_pop(stk, node, sp, tmin, tmax) ? DESCEND : END;
            }
            if (!__any_sync(RSB_FULL_MASK, st == DESCEND)) break;
        }
        const unsigned holding = __ballot_sync(RSB_FULL_MASK, parked);
        const unsigned moving = __ballot_sync(RSB_FULL_MASK, st == DESCEND);
        bool done = false, hit = false;
        if (holding != 0 && (__popc(holding) >= RQ_LEAF_MIN || moving == 0)) {
            if (parked) {
                parked = false;
                if (leaf(p_off, p_cnt, p_tmax)) {
                    rec.node = p_node;
                    done = hit = true;
                    st = IDLE;
                } else if (!AHEAD) {
                    st = kd_pop(stk, node, sp, tmin, tmax) ? DESCEND : END;
                }
            }
        }
        if (st == END && !parked) {
            done = true;
            st = IDLE;
        }
        if (__any_sync(RSB_FULL_MASK, done)) cl.commit(q, done, hit, rec);
    }
    if (COUNT) {
        __syncwarp();
        traced = warp_sum(traced);
        if ((threadIdx.x & 31) == 0 && traced) atomicAdd(&counters->rays, traced);
        flush_stats(stats, counters);
    }
}

// ---- ray sources / sinks of the batch and sweep entry points ------------------------------------------------
// [n][3] origins, directions (+ max_distance) -> the pipeline's SoA rows
// slot j of the pipeline takes the caller's query perm[j] (reordered batches) or j
__global__ void k_rq_batch_in(long long n, const double* __restrict__ origins, const double* __restrict__ directions,
                              const double* __restrict__ max_distance, const int32_t* __restrict__ perm, double* __restrict__ ray,
                              long long stride, double* __restrict__ md) {
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += step) {
        const long long i = perm ? perm[j] : j;
        ray[j] = origins[3 * i]; ray[stride + j] = origins[3 * i + 1]; ray[2 * stride + j] = origins[3 * i + 2];
        ray[3 * stride + j] = directions[3 * i]; ray[4 * stride + j] = directions[3 * i + 1]; ray[5 * stride + j] = directions[3 * i + 2];
        md[j] = max_distance ? max_distance[i] : RSB_INF;
    }
}

// the pipeline's hit records -> the arrays of rsb_hit_batch (intersection geometry generated once, for the winner)
template <int FEAT>
__global__ void __launch_bounds__(128)
k_rq_batch_out(Scene sc, RqBuf b, long long n, int32_t* __restrict__ out_prim, double* __restrict__ out_t, int32_t* __restrict__ out_sub,
               uint8_t* __restrict__ out_flags, int32_t* __restrict__ out_node, double* __restrict__ out_geom, float* __restrict__ out_uvw) {
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += step) {
        const int4 a = b.hit_a[j];
        const long long i = b.perm ? b.perm[j] : j;      // the caller's index of the query in slot j
        if (a.x >= 0) {
            const long long s = b.ray_stride;
            const V3 o = v3(b.ray[j], b.ray[s + j], b.ray[2 * s + j]);
            const V3 d = v3(b.ray[3 * s + j], b.ray[4 * s + j], b.ray[5 * s + j]);
            const float4 uvw = b.hit_uvw[j];
            HitRec rec;
            rec.t = b.hit_t[j];
            rec.prim = a.x; rec.leaf = a.y; rec.code = a.z; rec.flip = a.w;
            rec.u = uvw.x; rec.v = uvw.y; rec.w = uvw.z;
            rec.mesh_node = __float_as_int(uvw.w);
            rec.node = b.hit_node ? b.hit_node[j] : -1;
            Isect is;
            world_hit_geometry<FEAT>(sc, o, d, rec, &is);
            out_prim[i] = rec.prim;
            out_t[i] = rec.t;
            out_sub[i] = rec.code;
            out_flags[i] = (uint8_t)(is.exiting ? 1 : 0);
            if (out_node) { out_node[2 * i] = rec.node; out_node[2 * i + 1] = rec.mesh_node; }
            if (out_geom) {
                double* g = out_geom + 12 * i;
                g[0] = is.hit.x; g[1] = is.hit.y; g[2] = is.hit.z;
                g[3] = is.inside.x; g[4] = is.inside.y; g[5] = is.inside.z;
                g[6] = is.outside.x; g[7] = is.outside.y; g[8] = is.outside.z;
                g[9] = is.normal.x; g[10] = is.normal.y; g[11] = is.normal.z;
            }
            if (out_uvw) { out_uvw[3 * i] = rec.u; out_uvw[3 * i + 1] = rec.v; out_uvw[3 * i + 2] = rec.w; }
        } else {
            out_prim[i] = -1;
            out_t[i] = RSB_INF;
            out_sub[i] = -1;
            out_flags[i] = 0;
            if (out_node) { out_node[2 * i] = -1; out_node[2 * i + 1] = -1; }
            if (out_geom) { double* g = out_geom + 12 * i; for (int k = 0; k < 12; ++k) g[k] = 0.0; }
            if (out_uvw) { out_uvw[3 * i] = 0; out_uvw[3 * i + 1] = 0; out_uvw[3 * i + 2] = 0; }
        }
    }
}

// ---- reordering of incoherent query batches (rq_reorder) ------------------------------------------------------------------
// Rays that arrive in no particular order (after a diffuse bounce; a user's batch) put 32 unrelated walks into every warp:
// each lane chases its own nodes through L2 (19 of 32 lanes active, 25-28 % of the HBM roofline on 10,000 spheres, against
// 50-78 % for the same rays in Morton order).  Before the traversal the batch is therefore sorted on a 20-bit key -- origin
// cell and octahedral direction cell, both normalised to the extent THIS batch covers, Morton-interleaved -- with one
// counting sort: histogram, two-level exclusive scan, scatter.  The traversal runs on the permuted copy; the output stage
// writes every answer back at the caller's index.  Answers do not depend on the order: every query is independent.
#define RQ_KEY_BITS 20
#define RQ_KEY_BINS (1 << RQ_KEY_BITS)

__device__ __forceinline__ unsigned int ro_encode(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // monotone: float order = unsigned order
}
__device__ __forceinline__ float ro_decode(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// the five sort coordinates of a ray: origin x, y, z and the octahedral image (u, v) of its direction
__device__ __forceinline__ void ro_coords(const V3& o, const V3& d, float* c) {
    c[0] = (float)o.x; c[1] = (float)o.y; c[2] = (float)o.z;
    const float dx = (float)d.x, dy = (float)d.y, dz = (float)d.z;
    const float l1 = fabsf(dx) + fabsf(dy) + fabsf(dz);
    float u = dx / l1, v = dy / l1;
    if (dz < 0.f) {
        const float fu = (1.f - fabsf(v)) * (u >= 0.f ? 1.f : -1.f), fv = (1.f - fabsf(u)) * (v >= 0.f ? 1.f : -1.f);
        u = fu; v = fv;
    }
    c[3] = u; c[4] = v;
}

// bounds[0..4] = minima, bounds[5..9] = maxima (ro_encode'd); initialised to 0xFFFFFFFF / 0 by the host
template <class Source>
__global__ void k_ro_bounds(long long n, Source src, unsigned int* __restrict__ bounds) {
    float lo[5], hi[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) { lo[k] = 3.4e38f; hi[k] = -3.4e38f; }
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        float c[5];
        V3 o, d;
        src.get(i, o, d);
        ro_coords(o, d, c);
#pragma unroll
        for (int k = 0; k < 5; ++k)
            if (fabsf(c[k]) < 1e30f) { lo[k] = fminf(lo[k], c[k]); hi[k] = fmaxf(hi[k], c[k]); }      // (NaN and infinities sort first)
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(RSB_FULL_MASK, lo[k], s));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(RSB_FULL_MASK, hi[k], s));
        }
        if ((threadIdx.x & 31) == 0 && lo[k] <= hi[k]) {
            atomicMin(bounds + k, ro_encode(lo[k]));
            atomicMax(bounds + 5 + k, ro_encode(hi[k]));
        }
    }
}

__device__ __forceinline__ unsigned int ro_spread2(unsigned int x) {      // 10 bits -> every second bit
    x &= 0x3FFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

// key of every query + histogram.  A batch whose origins coincide (a camera, a probe point: extent below 1e-4 of the
// direction-independent scale) spends all 20 bits on the direction (1024 x 1024 cells); otherwise 2 bits per origin axis
// lead and 7 + 7 direction bits follow.
template <class Source>
__global__ void k_ro_keys(long long n, Source src, const unsigned int* __restrict__ bounds, unsigned int* __restrict__ key,
                          unsigned int* __restrict__ hist) {
    float lo[5], inv[5];
    float origin_extent = 0.f, origin_scale = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        lo[k] = ro_decode(bounds[k]);
        const float hi = ro_decode(bounds[5 + k]);
        const float ext = hi - lo[k];
        inv[k] = ext > 0.f ? 1.f / ext : 0.f;
        if (k < 3) { origin_extent = fmaxf(origin_extent, ext); origin_scale = fmaxf(origin_scale, fmaxf(fabsf(hi), fabsf(lo[k]))); }
    }
    const bool shared_origin = !(origin_extent > 1e-4f * origin_scale);
    const int dir_bits = shared_origin ? 10 : 7;
    const float dir_cells = (float)(1 << dir_bits);
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        float c[5];
        V3 o, d;
        src.get(i, o, d);
        ro_coords(o, d, c);
        int q[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const float cells = k < 3 ? 4.f : dir_cells;
            int v = (int)((c[k] - lo[k]) * inv[k] * cells);       // NaN -> 0
            q[k] = v < 0 ? 0 : (v >= (int)cells ? (int)cells - 1 : v);
        }
        unsigned int kd = ro_spread2((unsigned int)q[3]) | (ro_spread2((unsigned int)q[4]) << 1);
        unsigned int k20 = shared_origin ? kd : ((((unsigned int)q[0] << 4) | ((unsigned int)q[1] << 2) | (unsigned int)q[2]) << 14) | kd;
        k20 &= (RQ_KEY_BINS - 1);
        key[i] = k20;
        atomicAdd(hist + k20, 1u);
    }
}

// exclusive scan of the histogram, level 1: each block scans its 1024 bins in place and reports its total
__global__ void __launch_bounds__(1024) k_ro_scan_bins(unsigned int* __restrict__ hist, unsigned int* __restrict__ block_sums) {
    __shared__ unsigned int warp_tot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int v = hist[blockIdx.x * 1024 + threadIdx.x];
    unsigned int x = v;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const unsigned int y = __shfl_up_sync(RSB_FULL_MASK, x, s);
        if (lane >= s) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned int w = warp_tot[lane];
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const unsigned int y = __shfl_up_sync(RSB_FULL_MASK, w, s);
            if (lane >= s) w += y;
        }
        warp_tot[lane] = w;
    }
    __syncthreads();
    const unsigned int before = warp ? warp_tot[warp - 1] : 0u;
    hist[blockIdx.x * 1024 + threadIdx.x] = before + x - v;
    if (threadIdx.x == 1023) block_sums[blockIdx.x] = before + x;
}

// level 2: exclusive scan of the 1024 block totals (one block)
__global__ void __launch_bounds__(1024) k_ro_scan_blocks(unsigned int* __restrict__ block_sums) {
    __shared__ unsigned int warp_tot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int v = block_sums[threadIdx.x];
    unsigned int x = v;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const unsigned int y = __shfl_up_sync(RSB_FULL_MASK, x, s);
        if (lane >= s) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned int w = warp_tot[lane];
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const unsigned int y = __shfl_up_sync(RSB_FULL_MASK, w, s);
            if (lane >= s) w += y;
        }
        warp_tot[lane] = w;
    }
    __syncthreads();
    block_sums[threadIdx.x] = (warp ? warp_tot[warp - 1] : 0u) + x - v;
}

// slot of every query: the bin's running cursor starts at its exclusive offset.  Only the 4-byte permutation is scattered;
// the input stage then READS the caller's rays (or generates the sweep's) in slot order and writes them coalesced.
// (First version: this kernel also moved the six SoA components of every ray -- 24 M scattered 8-byte stores per 4 M rays,
// 0.81 ms, a fifth of the whole sorted sweep.)
__global__ void k_ro_scatter(long long n, const unsigned int* __restrict__ key, unsigned int* __restrict__ hist,
                             const unsigned int* __restrict__ block_sums, int32_t* __restrict__ perm) {
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const unsigned int k = key[i];
        perm[(long long)block_sums[k >> 10] + atomicAdd(hist + k, 1u)] = (int32_t)i;
    }
}

__device__ __forceinline__ unsigned int morton_compact(unsigned long long x) {
    x &= 0x5555555555555555ULL;
    x = (x | (x >> 1)) & 0x3333333333333333ULL;
    x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0FULL;
    x = (x | (x >> 4)) & 0x00FF00FF00FF00FFULL;
    x = (x | (x >> 8)) & 0x0000FFFF0000FFFFULL;
    x = (x | (x >> 16)) & 0x00000000FFFFFFFFULL;
    return (unsigned int)x;
}

// Ray `index` of a sweep: from `origin` toward target + (jx, jy, 0) * half_window.  order_log2 = 0: (jx, jy) uniform
// over the whole window, independently per ray (incoherent, like rays after a diffuse bounce).  order_log2 = g > 0: the
// window is a 2^g x 2^g grid of cells walked along the Morton curve, ray `index` jitters inside cell index mod 4^g --
// neighbouring indices are neighbouring cells, the way an observer hands out the pixels of an image (primary rays).
struct SweepSource {
    long long first_index;
    unsigned long long seed;
    double ox, oy, oz, tx, ty, tz, half_window;
    int order_log2;
    __device__ __forceinline__ void get(long long i, V3& o, V3& d) const {
        const unsigned long long index = (unsigned long long)(first_index + i);
        Philox4x32 px;
        px.init(seed, index, 0u);
        double u1 = (double)(px.next_u64() >> 11) * (1.0 / 9007199254740992.0);
        double u2 = (double)(px.next_u64() >> 11) * (1.0 / 9007199254740992.0);
        if (order_log2 > 0) {
            const unsigned long long cell = index & ((1ULL << (2 * order_log2)) - 1ULL);
            const double inv = 1.0 / (double)(1ULL << order_log2);
            u1 = ((double)morton_compact(cell) + u1) * inv;
            u2 = ((double)morton_compact(cell >> 1) + u2) * inv;
        }
        const V3 p = v3(tx + (2.0 * u1 - 1.0) * half_window, ty + (2.0 * u2 - 1.0) * half_window, tz);
        o = v3(ox, oy, oz);
        d = normalise(v3(p.x - ox, p.y - oy, p.z - oz));
    }
};

// the caller's arrays of rsb_hit_batch
struct BatchSource {
    const double* origins;
    const double* directions;
    __device__ __forceinline__ void get(long long i, V3& o, V3& d) const {
        o = v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        d = v3(directions[3 * i], directions[3 * i + 1], directions[3 * i + 2]);
    }
};

// slot j of the pipeline takes ray perm[j] of the sweep (reordered) or ray j
__global__ void k_rq_sweep_gen(long long n, SweepSource src, const int32_t* __restrict__ perm, double* __restrict__ ray, long long stride) {
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += step) {
        V3 o, d;
        src.get(perm ? perm[j] : j, o, d);
        ray[j] = o.x; ray[stride + j] = o.y; ray[2 * stride + j] = o.z;
        ray[3 * stride + j] = d.x; ray[4 * stride + j] = d.y; ray[5 * stride + j] = d.z;
    }
}

__global__ void k_rq_sweep_reduce(RqBuf b, long long n, long long first_index, unsigned long long* out_hits, double* out_sum_t,
                                  unsigned long long* out_xor_prim) {
    unsigned long long hits = 0, xr = 0;
    double sum_t = 0.0;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const int prim = b.hit_a[i].x;
        if (prim >= 0) {
            hits += 1;
            sum_t += b.hit_t[i];
            xr ^= (unsigned long long)(unsigned)prim * 0x9E3779B97F4A7C15ULL + (unsigned long long)(first_index + (b.perm ? b.perm[i] : i));
        }
    }
    __syncwarp();
    hits = warp_sum(hits);
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
        sum_t += __shfl_down_sync(RSB_FULL_MASK, sum_t, k);
        xr ^= __shfl_down_sync(RSB_FULL_MASK, xr, k);
    }
    if ((threadIdx.x & 31) == 0 && hits) {
        atomicAdd(out_hits, hits);
        atomicAdd(out_sum_t, sum_t);
        atomicXor(out_xor_prim, xr);
    }
}

}  // namespace rsb
