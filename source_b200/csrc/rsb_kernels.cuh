// rsb_kernels.cuh -- sm_100a kernels of the hot path.
//
//  (World.hit over ray batches: rsb_trav.cuh)
//  k_contains_batch World.contains over a point batch
//  k_render         Observer._render_pixel for a pinhole camera: persistent threads, 1 thread = 1 pixel
//                   stream, one path SEGMENT per loop trip (no lane waits for a long path), finished
//                   paths are folded into the pixel's spectral statistics by the whole warp
//                   (lane = spectral bin) using ballot/shuffle
//  k_frame_combine  StatsArray3D.combine_samples over a frame slice
//  k_rng_uniform    MT19937-64 known-answer stream
#pragma once
#include <cuda_runtime.h>

#include "rsb_path.h"

namespace rsb {

#define RSB_FULL_MASK 0xffffffffu
#ifndef RSB_TRACE_MIN_BLOCKS
#define RSB_TRACE_MIN_BLOCKS 3      // k_wf_trace: CTAs of 128 threads per SM the register allocation must allow
#endif
#ifndef RSB_SHADE_MIN_BLOCKS
#define RSB_SHADE_MIN_BLOCKS 4
#endif
#ifndef RSB_MESH_MIN_BLOCKS
#define RSB_MESH_MIN_BLOCKS 4       // traversal kernels of scenes with meshes: resident CTAs per SM the registers must allow
#endif
#define RSB_RENDER_THREADS 128
#ifndef RSB_REFILL_LANES
#define RSB_REFILL_LANES 8          // persistent-lane kernels: idle lanes that trigger a refill
#endif

struct DevCounters {   // mirrors RsbCounters
    unsigned long long rays, branches, leaves, items, prim_tests, tri_tests, paths, contains, table_reads,
        contains_nodes, contains_items, contains_prim_tests;
};

template <bool COUNT>
struct StatsSel { typedef NoStats type; };
template <>
struct StatsSel<true> { typedef CountStats type; };

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(RSB_FULL_MASK, v, o);
    return v;
}

__device__ __forceinline__ void flush_stats(const NoStats&, DevCounters*, bool = false) {}
// shade: the traversal counters of this kernel belong to World.contains queries, not World.hit
__device__ __forceinline__ void flush_stats(const CountStats& s, DevCounters* c, bool shade = false) {
    // all lanes of the warp reach here together (kernel epilogue)
    unsigned long long b = warp_sum(s.branches), l = warp_sum(s.leaves), i = warp_sum(s.items),
                       p = warp_sum(s.prim_tests), t = warp_sum(s.tri_tests), cq = warp_sum(s.contains),
                       tb = warp_sum(s.tables);
    if ((threadIdx.x & 31) == 0 && shade) {
        atomicAdd(&c->contains, cq);
        atomicAdd(&c->table_reads, tb);
        atomicAdd(&c->contains_nodes, b + l);
        atomicAdd(&c->contains_items, i);
        atomicAdd(&c->contains_prim_tests, p);
        atomicAdd(&c->tri_tests, t);
        return;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&c->contains, cq);
        atomicAdd(&c->table_reads, tb);
        atomicAdd(&c->branches, b);
        atomicAdd(&c->leaves, l);
        atomicAdd(&c->items, i);
        atomicAdd(&c->prim_tests, p);
        atomicAdd(&c->tri_tests, t);
    }
}

// ---------------------------------------------------------------------------------------------
// Shared-memory staging of the world-level acceleration data.  The world kd-tree, its leaf item
// list and the primitive table are read by every ray on every segment; for scene-graph sized
// worlds (Cornell box: 89 nodes + 8 primitives = 4.5 KB; 10k spheres does not fit and stays in
// L2) they are copied once per CTA into shared memory and the Scene's pointers are redirected.
// ---------------------------------------------------------------------------------------------
struct StageLayout {
    int32_t nodes_bytes, items_bytes, prims_bytes, total;
};

__host__ __device__ inline StageLayout stage_layout(int n_nodes, int n_items, int n_prims) {
    StageLayout l;
    l.nodes_bytes = n_nodes * (int)sizeof(KdNode);
    l.items_bytes = ((n_items * 4 + 15) / 16) * 16;
    l.prims_bytes = n_prims * (int)sizeof(Prim);
    l.total = l.nodes_bytes + l.items_bytes + l.prims_bytes;
    return l;
}

// 1-D bulk copies global -> shared through the TMA engine (cp.async.bulk, completion counted on an mbarrier in bytes):
// one elected thread issues up to a few descriptors-free copies and the CTA waits on the barrier's phase, while the
// loads the threads issued before (slot state, queue entries) stay in flight.  Sizes and addresses are multiples of 16.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct BulkStage {
    uint64_t* bar;
    // all threads of the CTA; `bar` is 8 bytes of shared memory
    __device__ __forceinline__ void init(uint64_t* b) {
        bar = b;
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    // thread 0: announce the total, then issue the copies
    __device__ __forceinline__ void expect(uint32_t total_bytes) const {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(total_bytes) : "memory");
    }
    __device__ __forceinline__ void copy(void* dst, const void* src, uint32_t bytes) const {
        if (bytes == 0) return;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                     "l"(src), "r"(bytes), "r"(smem_addr(bar))
                     : "memory");
    }
    // all threads: until phase `parity` of the barrier has completed
    __device__ __forceinline__ void wait(uint32_t parity) const {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "RSB_BULK_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra RSB_BULK_DONE;\n"
            "bra RSB_BULK_WAIT;\n"
            "RSB_BULK_DONE:\n"
            "}\n" ::"r"(smem_addr(bar)),
            "r"(parity)
            : "memory");
    }
};

// n_items: number of leaf item ids of the world tree.  STAGED is a compile-time flag (RSB_FEAT_STAGED) so that
// every load through the redirected pointers is compiled as a shared-memory load, not a generic one.
template <bool STAGED>
__device__ __forceinline__ void stage_scene(Scene& sc, unsigned char* smem, int n_items) {
    if (!STAGED) return;
    __shared__ __align__(8) uint64_t stage_bar;
    StageLayout l = stage_layout(sc.world.n_nodes, n_items, sc.n_prims);
    BulkStage bs;
    bs.init(&stage_bar);
    if (threadIdx.x == 0) {
        bs.expect((uint32_t)l.total);
        bs.copy(smem, sc.world.nodes, (uint32_t)l.nodes_bytes);
        // item list: whole 16-B words (the allocation is padded to 16 B)
        bs.copy(smem + l.nodes_bytes, sc.world.items, (uint32_t)l.items_bytes);
        bs.copy(smem + l.nodes_bytes + l.items_bytes, sc.prims, (uint32_t)l.prims_bytes);
    }
    bs.wait(0);
    sc.world.nodes = reinterpret_cast<const KdNode*>(smem);
    sc.world.items = reinterpret_cast<const int32_t*>(smem + l.nodes_bytes);
    sc.prims = reinterpret_cast<const Prim*>(smem + l.nodes_bytes + l.items_bytes);
}

// one region, staged the same way (material rows of k_wf_shade, spectral tables of k_wf_finalize)
__device__ __forceinline__ void stage_region(uint64_t* bar, void* dst, const void* src, int bytes) {
    BulkStage bs;
    bs.init(bar);
    if (threadIdx.x == 0) {
        bs.expect((uint32_t)bytes);
        bs.copy(dst, src, (uint32_t)bytes);
    }
    bs.wait(0);
}

// RayAx storage of the calling thread (rsb_geom.h): RSB_AX_WORDS columns of blockDim.x doubles behind the staged
// scene.  Kernels that use it are launched with RSB_RENDER_THREADS threads and ax_bytes(FEAT) extra shared memory.
#define RSB_COOP_CAP 64      // (ray, triangle) pairs a warp tests per cooperative round
#define RSB_COOP_WARP_BYTES ((36 + 32 + RSB_COOP_CAP) * 4 + RSB_COOP_CAP * 16)
#define RSB_COOP_BYTES (5 * 4 * RSB_RENDER_THREADS + (RSB_RENDER_THREADS / 32) * RSB_COOP_WARP_BYTES)
__host__ __device__ inline int ax_bytes(int feat) {
    return (feat & RSB_FEAT_MESH) ? RSB_AX_WORDS * 8 * RSB_RENDER_THREADS + RSB_COOP_BYTES : 9 * 8 * RSB_RENDER_THREADS;
}

template <bool STAGED>
__device__ __forceinline__ double* ax_storage(unsigned char* smem, const Scene& sc, int n_items) {
    int off = STAGED ? stage_layout(sc.world.n_nodes, n_items, sc.n_prims).total : 0;
    return reinterpret_cast<double*>(smem + off) + threadIdx.x;
}

// ---------------------------------------------------------------------------------------------
// Cooperative triangle tests.  In the mesh traversal every lane reaches a leaf with its own number of triangles
// (0 for ~40 % of the leaves, a dozen for some); looping over them per lane ran the Woop test -- 45 % of the
// kernel's instructions -- with 1.8 of 32 lanes active (round-1 profile of the 1.3 M-triangle sweep).  Here the
// warp pools the (ray, triangle) pairs of all its lanes' current leaves and deals them out evenly: pair g belongs
// to the lane whose exclusive prefix sum of leaf sizes brackets g, the testing lane reads that lane's mesh-local
// origin and ray-space shear from shared memory, and the owner then replays MeshData._trace_leaf's comparison
// (`t < distance`, first of equal-t triangles wins, mesh.pyx:520-563) over its own results in leaf order -- the
// same values through the same comparisons as the sequential loop.
struct CoopSmem {
    int32_t* rs_pack;     // [T] ix | iy << 2 | iz << 4 of the thread's ray-space permutation
    float* rs_s;          // [3][T] sx, sy, sz
    int32_t* mesh_idx;    // [T]
    int32_t* prefix;      // warp: [33]
    int32_t* leaf_off;    // warp: [32]
    int32_t* res_tri;     // warp: [CAP]
    float4* res;          // warp: [CAP] (t, u, v, w); t = NaN: no hit
    const double* ax_cta; // RayAx storage of thread 0 of the CTA
};

__device__ __forceinline__ CoopSmem coop_carve(double* axbuf_thread) {
    CoopSmem cs;
    double* ax0 = axbuf_thread - threadIdx.x;
    cs.ax_cta = ax0;
    unsigned char* base = reinterpret_cast<unsigned char*>(ax0 + RSB_AX_WORDS * RSB_RENDER_THREADS);
    cs.rs_pack = reinterpret_cast<int32_t*>(base);
    cs.rs_s = reinterpret_cast<float*>(base + 4 * RSB_RENDER_THREADS);
    cs.mesh_idx = reinterpret_cast<int32_t*>(base + 16 * RSB_RENDER_THREADS);
    unsigned char* w = base + 20 * RSB_RENDER_THREADS + (threadIdx.x >> 5) * RSB_COOP_WARP_BYTES;
    cs.res = reinterpret_cast<float4*>(w);
    cs.prefix = reinterpret_cast<int32_t*>(w + RSB_COOP_CAP * 16);
    cs.leaf_off = cs.prefix + 36;
    cs.res_tri = cs.leaf_off + 32;
    return cs;
}

// publish the calling thread's ray-space transform and mesh for the lanes that will test its triangles
__device__ __forceinline__ void coop_publish(const CoopSmem& cs, const RaySpace& rs, int mesh_index) {
    const int t = threadIdx.x;
    cs.rs_pack[t] = rs.ix | (rs.iy << 2) | (rs.iz << 4);
    cs.rs_s[t] = rs.sx;
    cs.rs_s[RSB_RENDER_THREADS + t] = rs.sy;
    cs.rs_s[2 * RSB_RENDER_THREADS + t] = rs.sz;
    cs.mesh_idx[t] = mesh_index;
}

// All 32 lanes.  in_leaf: this lane stands at a mesh leaf (off, cnt) and accepts hits closer than d0 =
// min(ray.max_distance, leaf max_range).  true: *mh holds the leaf's closest triangle.
template <class Stats>
__device__ __forceinline__ bool mesh_leaf_coop(const Scene& sc, const CoopSmem& cs, bool in_leaf, int off, int cnt, double d0,
                                               double max_distance, MeshHit* mh, Stats& stats) {
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
    __syncwarp();
    cs.prefix[lane] = excl;
    if (lane == 31) cs.prefix[32] = total;
    cs.leaf_off[lane] = off;
    double distance = d0;
    int closest = -1;
    float cu = 0, cv = 0, cw = 0;
    for (int base = 0; base < total; base += RSB_COOP_CAP) {
        __syncwarp();
        const int lim = total - base < RSB_COOP_CAP ? total - base : RSB_COOP_CAP;
        for (int p = lane; p < lim; p += 32) {
            const int g = base + p;
            int lo = 0, hi = 32;           // prefix[lo] <= g < prefix[hi]
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                int mid = (lo + hi) >> 1;
                if (cs.prefix[mid] <= g) lo = mid; else hi = mid;
            }
            const int ot = tid0 + lo;      // the owner's thread index within the CTA
            const Mesh& m = sc.meshes[cs.mesh_idx[ot]];
            const int tri = m.tree.items[cs.leaf_off[lo] + (g - cs.prefix[lo])];
            const double* ax = cs.ax_cta + ot;
            V3 o = v3(ax[9 * RSB_RENDER_THREADS], ax[10 * RSB_RENDER_THREADS], ax[11 * RSB_RENDER_THREADS]);
            RaySpace rs;
            const int pk = cs.rs_pack[ot];
            rs.ix = pk & 3; rs.iy = (pk >> 2) & 3; rs.iz = (pk >> 4) & 3;
            rs.sx = cs.rs_s[ot]; rs.sy = cs.rs_s[RSB_RENDER_THREADS + ot]; rs.sz = cs.rs_s[2 * RSB_RENDER_THREADS + ot];
            float h[4];
            stats.tri_test();
            const bool hit = mesh_hit_triangle(m.tri + 3 * (size_t)tri, o, max_distance, rs, h);
            cs.res_tri[p] = tri;
            cs.res[p] = hit ? make_float4(h[3], h[0], h[1], h[2]) : make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);
        }
        __syncwarp();
        if (c > 0) {
            const int j0 = base > excl ? base - excl : 0;
            const int j1 = base + RSB_COOP_CAP - excl < c ? base + RSB_COOP_CAP - excl : c;
            for (int j = j0; j < j1; ++j) {
                const int p = excl + j - base;
                const float4 r = cs.res[p];
                if (r.x == r.x) {
                    const double t = (double)r.x;
                    if (t < distance) {
                        distance = t;
                        closest = cs.res_tri[p];
                        cu = r.y; cv = r.z; cw = r.w;
                    }
                }
            }
        }
    }
    __syncwarp();
    if (closest < 0) return false;
    mh->t = (double)(float)distance;
    mh->tri = closest;
    mh->u = cu; mh->v = cv; mh->w = cw;
    return true;
}

// (World.hit over ray batches -- rsb_hit_batch, rsb_hit_sweep -- is the kernel pipeline of rsb_trav.cuh)

__global__ void __launch_bounds__(128)
k_contains_batch(Scene sc, long long n, const double* __restrict__ points, int cap, int32_t* __restrict__ out_count,
                 int32_t* __restrict__ out_prims) {
    NoStats stats;
    KdStackEntry stack[RSB_KD_STACK];
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        V3 p = v3(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
        int offset, count, found = 0;
        if (kd_locate(sc.world, p, &offset, &count, stats)) {
            for (int k = 0; k < count; ++k) {
                int id = sc.world.items[offset + k];
                if (prim_contains(sc, id, p, stack, stats)) {
                    if (found < cap) out_prims[i * cap + found] = id;
                    ++found;
                }
            }
        }
        out_count[i] = found;
    }
}

__global__ void k_rng_uniform(unsigned long long seed, long long n, unsigned long long* state, double* out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Rng rng;
    rng.mode = RNG_MT19937_64;
    rng.mt.mt = reinterpret_cast<uint64_t*>(state);
    rng.mt.stride = 1;
    rng.mt.seed(seed);
    for (long long i = 0; i < n; ++i) out[i] = rng.uniform();
}

// ---------------------------------------------------------------------------------------------
// Wavefront renderer.  Observer._render_pixel (observer.pyx:363-419) runs one path at a time per pixel
// and the reference's RNG stream makes a pixel's samples inherently sequential, so the unit of
// parallelism is the PIXEL STREAM: P slots, each owning one pixel at a time (dynamic assignment from a
// work counter) and carrying exactly one path.  A "wave" advances every live path by one segment through
// three small kernels, so that all warps of an SM execute the same few KB of code (the monolithic
// per-thread path tracer measured 68 % instruction-fetch stalls and 6.7 of 32 lanes active, see
// profiles/):
//   k_wf_trace     roulette + World.hit                 1 thread = 1 slot
//   k_wf_shade     geometry, material, volumes, spawn   1 thread = 1 slot
//   k_wf_finalize  ended paths: log replay + Welford into the frame (1 warp = 1 path, lane = bin),
//                  then the slot's next sample / next pixel is generated
// Slot state is SoA in HBM; ended slots are handed to k_wf_finalize through a compacted list.
// ---------------------------------------------------------------------------------------------
enum SlotStatus : int32_t { SLOT_IDLE = 0, SLOT_ALIVE = 1, SLOT_HIT = 2, SLOT_ENDED_ZERO = 3, SLOT_ENDED_EMIT = 4 };

struct WfSlots {
    double* ray;            // [6][P] origin xyz, direction xyz
    double* weight;         // [P] projection weight of the sample in flight
    double* norm;           // [P] roulette normalisation of the segment in flight
    double* hit_t;          // [P]
    int4* hit_a;            // [P] prim, leaf, code, flip
    float4* hit_uvw;        // [P] mesh barycentrics
    int32_t* depth;         // [P]
    uint32_t* rays;         // [P] reference ray counter of the path in flight
    int32_t* sample;        // [P] index of the sample in flight
    int32_t* px;            // [P]
    int32_t* py;            // [P]
    int32_t* status;        // [P]
    int32_t* log_n;         // [P]
    uint32_t* philox_idx;   // [P] draws so far (Philox) / cursor of the pixel's path stream (MT19937-64)
    int32_t* work;          // [P] index (within the current pixel chunk) of the pixel the slot is rendering
    int32_t* group;         // [P] stream group of that work item: pass * n_slices + slice
    int32_t* additive;      // [P] the path in flight has logged an emitting volume (only touched when has_additive)
    int32_t* pix_mti;       // [n_chunk] MT19937-64 cursor of every stream of the chunk after its 2*spp jitter draws
    unsigned long long* pix_mt;   // [n_chunk][312] their state words (seeded up front by k_wf_seed)
    double* pix_jitter;     // [n_chunk][2*spp] ([4*spp] for a CCD) the jitter draws of every stream: uniform() number 0 .. 2*spp-1 of seed(...),
                            // which RectangleSampler3D.samples(spp) consumes before any tracing (pinhole.pyx:183)
    const unsigned long long* mt_table;   // [312] seed-independent part of seed(d) (rsb_rng.h mt_seed_table)
    LogEntry* log;          // [P][log_capacity]
    int32_t* ended;         // [2][P] compacted lists of ended slots (double buffered by wave parity)
    unsigned int* n_ended;  // [2]
    int32_t* hit_list;      // [4][P] slots with a hit, by material type of the hit primitive
    unsigned int* n_hit;    // [4]
};

struct WfArgs {
    Scene sc;
    Spectral sp;
    RayConfig cfg;
    Camera cam;
    WfSlots st;
    long long n_pixels;         // work items (pass, pixel) of the current chunk
    long long item_base;        // global index of the chunk's first work item; item g = pass * n_pix_pass + task
    long long n_pix_pass;       // pixel tasks per pass (length of `pixels`, or nx * ny)
    const int32_t* pixels;      // the task list [n_pix_pass][2], or null = the whole frame
    double* mean;               // pass 0 accumulates straight into the caller's frame
    double* variance;
    double* pass_mean;          // [n_passes - 1][frame_elems] frames of passes 1.., merged by k_pass_combine
    double* pass_variance;
    long long frame_elems;      // nx * ny * bins
    // RGBPipeline2D (rgb.pyx:216-290): every sample's spectrum is also projected on the CIE XYZ curves and the three
    // tristimulus values get their own per-work-item Welford statistics.  Null xyz_mean = no RGB pipeline; null mean = no
    // spectral pipeline (the per-bin statistics are then not kept at all).
    // Generalised to "projection channels": every channel sums one curve times the sample's spectrum over the bins, in the
    // operation order of the pixel processor it stands for (proj_mode: RSB_PROJ_XYZ rgb.pyx:550-558 / colour.pyx:178-186;
    // RSB_PROJ_POWER mono/power.pyx:768-779; RSB_PROJ_RADIANCE mono/radiance.pyx:184-195).
    const double* xyz_tab;      // [n_slices][bins][proj_channels]: the curves resampled on every slice's wavelength range
    const double* xyz_delta;    // [n_slices]: Spectrum.delta_wavelength of every slice
    double* xyz_mean;           // [work items][proj_channels], work item g = group * n_pix_pass + task (group = pass * n_slices + slice)
    double* xyz_variance;
    int32_t proj_channels;      // 1 .. RSB_PROJ_MAX
    int32_t proj_mode[8];
    unsigned long long seed_stride;   // pass p draws from streams seeded seed + p * seed_stride + y * nx + x
    int32_t n_passes;
    int32_t n_slices;           // spectral slices rendered concurrently; `sp` describes slice 0, slice k follows at strides
    int32_t frame_bins;         // bins of a frame row = n_slices * sp.bins
    int32_t has_additive;       // the scene has volume emitters: dark path ends may still have to be replayed
    unsigned long long* ray_count;
    unsigned long long* work_counter;
    unsigned int* n_idle;
    int32_t* overflow_flag;
    DevCounters* counters;
    unsigned long long seed;
    int32_t n_slots;
    int32_t log_capacity;
    int32_t n_items;
    int32_t staged;
    int32_t tables_staged;
    int32_t wave;
};

// Stream group of the work item a slot is rendering: (accumulated observe() call) * n_slices + (spectral slice).  Every
// group is an independent set of pixel streams, seeded seed + group * seed_stride + y * nx + x.
__device__ __forceinline__ int wf_group_of(const WfArgs& a, int slot) {
    if (a.n_passes * a.n_slices <= 1) return 0;
    return a.st.group[slot];
}
__device__ __forceinline__ int wf_pass_of(const WfArgs& a, int slot) { return a.n_slices > 1 ? wf_group_of(a, slot) / a.n_slices : wf_group_of(a, slot); }
__device__ __forceinline__ int wf_slice_of(const WfArgs& a, int slot) { return a.n_slices > 1 ? wf_group_of(a, slot) % a.n_slices : 0; }

// Per-slice material rows and spectral tables: slice k's block follows slice 0's ([n_materials] rows; [2][n_tables][bins]
// doubles: the tables, then their logs)
__device__ __forceinline__ Spectral wf_slice_spectral(const Spectral& sp, int slice) {
    Spectral s = sp;
    if (slice > 0) {
        s.mats = sp.mats + (size_t)slice * sp.n_materials;
        s.tables = sp.tables + (size_t)slice * 2 * sp.n_tables * sp.bins;
        s.tables_ln = s.tables + (size_t)sp.n_tables * sp.bins;
    }
    return s;
}

// The cursors of the pixel stream a slot is rendering are kept per SLOT (copied from the pixel's seeded cursors
// when the slot picks the pixel up), so that they arrive with the rest of the slot state instead of behind a
// second dependent load through the pixel index.
template <int RNGMODE>
__device__ __forceinline__ void wf_load_rng(const WfArgs& a, int slot, Rng& rng) {
    rng.mode = RNGMODE;
    if (RNGMODE == RNG_MT19937_64) {
        size_t w = (size_t)a.st.work[slot];
        rng.mt.mt = reinterpret_cast<uint64_t*>(a.st.pix_mt) + w * RSB_MT_NN;
        rng.mt.stride = 1;
        rng.mt.mti = (int)a.st.philox_idx[slot];
    } else {
        long long pixel_id = (long long)a.st.py[slot] * a.cam.nx + a.st.px[slot];
        rng.px.init(a.seed + (unsigned long long)wf_group_of(a, slot) * a.seed_stride, (unsigned long long)pixel_id,
                    (uint32_t)a.st.sample[slot]);
        rng.px.idx = a.st.philox_idx[slot];
    }
}

// Fetch the words the next RSB_MT_WIN_DRAWS draws of `g` will read into the thread's shared-memory column
// (rsb_rng.h): 13 independent loads in flight at once.  (A prefetch.global.L1 of the same words made no
// difference: the kernel is bound by the NUMBER of dependent round trips, and the prefetch sat behind the same
// cursor load.)
__device__ __forceinline__ void mt_preload(Mt19937_64& g, uint64_t* col) {
    const int i = g.mti >= RSB_MT_NN ? 0 : g.mti;
    uint64_t v[RSB_MT_WIN_WORDS];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        int j = i + k;
        v[k] = g.mt[j < RSB_MT_NN ? j : j - RSB_MT_NN];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        int j = i + k + RSB_MT_MM;
        v[7 + k] = g.mt[j < RSB_MT_NN ? j : j - RSB_MT_NN];
    }
#pragma unroll
    for (int k = 0; k < RSB_MT_WIN_WORDS; ++k) col[k * RSB_MT_WIN_STRIDE] = v[k];
    g.win = col;
    g.win_i0 = i;
    g.win_n = RSB_MT_WIN_DRAWS;
}

template <int RNGMODE>
__device__ __forceinline__ void wf_store_rng(const WfArgs& a, int slot, const Rng& rng) {
    if (RNGMODE == RNG_MT19937_64) a.st.philox_idx[slot] = (uint32_t)rng.mt.mti;
    else a.st.philox_idx[slot] = rng.px.idx;
}

__device__ __forceinline__ void wf_push_ended(const WfArgs& a, int slot) {
    int par = a.wave & 1;
    unsigned int k = atomicAdd(&a.st.n_ended[par], 1u);
    a.st.ended[(size_t)par * a.n_slots + k] = slot;
}

// Pixel (x, y) of work item w of the current chunk (FullFrameSampler2D task list, or the whole frame); returns
// the stream group (pass * n_slices + slice) the item belongs to
__device__ __forceinline__ int wf_pixel_of(const WfArgs& a, unsigned long long w, int* px, int* py) {
    unsigned long long g = w + (unsigned long long)a.item_base;
    int pass = 0;
    if (a.n_passes * a.n_slices > 1) {
        pass = (int)(g / (unsigned long long)a.n_pix_pass);
        g -= (unsigned long long)pass * (unsigned long long)a.n_pix_pass;
    }
    if (a.pixels) { *px = a.pixels[2 * g]; *py = a.pixels[2 * g + 1]; }
    else {
        *px = (int)(g / (unsigned long long)a.cam.ny);
        *py = (int)(g % (unsigned long long)a.cam.ny);
    }
    return pass;
}

// 1 warp = 32 streams of the chunk: seed(seed + group*stride + y*nx + x) for every stream ahead of time (seeding is a
// dependent chain; done lazily inside the wave loop it put the latency of one chain on the critical path of every
// wave).  Round 1 ran the full 935-step init_by_array64 twice per stream (a jitter cursor and a path cursor) with
// every thread writing its own 5 KB row word by word: 119 ms per 8 M streams, all of it uncoalesced 8-byte stores.
// Now: 312 steps per stream (rsb_rng.h mt_seed_fast: the rest is a seed-independent table), ONE state per stream, the
// words of 32 streams transposed through shared memory so that every store instruction writes 256 contiguous bytes of
// one stream, and the 2*spp jitter draws taken right away into a small per-stream table.
#define RSB_SEED_WARPS 4
__global__ void __launch_bounds__(32 * RSB_SEED_WARPS) k_wf_seed(const __grid_constant__ WfArgs a) {
    __shared__ uint64_t tile_all[RSB_SEED_WARPS][32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t (*tile)[33] = tile_all[warp];
    const long long w0 = ((long long)blockIdx.x * RSB_SEED_WARPS + warp) * 32;
    if (w0 >= a.n_pixels) return;
    const long long w = w0 + lane;
    const bool valid = w < a.n_pixels;
    const uint64_t* T = reinterpret_cast<const uint64_t*>(a.st.mt_table);
    uint64_t d = 1;
    if (valid) {
        int px, py;
        const int group = wf_pixel_of(a, (unsigned long long)w, &px, &py);
        d = a.seed + (unsigned long long)group * a.seed_stride + (unsigned long long)((long long)py * a.cam.nx + px);
    }
    uint64_t* state = reinterpret_cast<uint64_t*>(a.st.pix_mt);
    const int rows = (int)((a.n_pixels - w0) < 32 ? (a.n_pixels - w0) : 32);
    const uint64_t m1 = mt_seed_first(T, d);
    uint64_t prev = m1;
    for (int i0 = 0; i0 < RSB_MT_NN; i0 += 32) {
        const int n = RSB_MT_NN - i0 < 32 ? RSB_MT_NN - i0 : 32;
        for (int k = 0; k < n; ++k) {
            const int i = i0 + k;
            uint64_t v;
            if (i == 0) v = 9223372036854775808ULL;
            else if (i == 1) v = 0;                       // written last: it needs word 311
            else { prev = mt_seed_step(__ldg(T + i), prev, i); v = prev; }
            tile[lane][k] = v;
        }
        __syncwarp();
        // row r of the tile = words i0 .. i0+n-1 of stream w0 + r: one coalesced store per stream
        for (int r = 0; r < rows; ++r)
            if (lane < n) state[(size_t)(w0 + r) * RSB_MT_NN + i0 + lane] = tile[r][lane];
        __syncwarp();
    }
    if (!valid) return;
    uint64_t* mine = state + (size_t)w * RSB_MT_NN;
    mine[1] = mt_seed_last(m1, prev);
    // the jitter draws: the first 2*spp outputs of the stream, regenerating its words lazily in place like every draw
    Mt19937_64 g;
    g.mt = mine;
    g.stride = 1;
    g.mti = RSB_MT_NN;
    const int nj = 2 * camera_jitter_pairs(a.cam.kind) * a.cam.pixel_samples;
    double* jit = a.st.pix_jitter + (size_t)w * nj;
    if (a.cam.kind == 3) {
        int px, py;
        wf_pixel_of(a, (unsigned long long)w, &px, &py);
        if (!camera_pixel_draws(a.cam, px, py)) { a.st.pix_mti[w] = g.mti; return; }     // an edge pixel of a VectorCamera draws nothing
    }
    for (int k = 0; k < nj; ++k) jit[k] = (double)(g.next_u64() >> 11) * (1.0 / 9007199254740992.0);
    a.st.pix_mti[w] = g.mti;
}

// Next sample of the slot's pixel, or the next pixel from the work counter; PinholeCamera._generate_rays for
// that sample.  Runs on one thread.
template <int RNGMODE>
__device__ __forceinline__ void wf_regenerate(const WfArgs& a, int slot) {
    const int spp = a.cam.pixel_samples;
    int s = a.st.sample[slot];
    int px = a.st.px[slot], py = a.st.py[slot];
    if (s >= spp) {
        unsigned long long w = atomicAdd(a.work_counter, 1ULL);
        if (w >= (unsigned long long)a.n_pixels) {
            a.st.status[slot] = SLOT_IDLE;
            atomicAdd(a.n_idle, 1u);
            return;
        }
        const int group = wf_pixel_of(a, w, &px, &py);
        a.st.px[slot] = px;
        a.st.py[slot] = py;
        a.st.work[slot] = (int32_t)w;
        if (a.n_passes * a.n_slices > 1) a.st.group[slot] = group;
        if (RNGMODE == RNG_MT19937_64) a.st.philox_idx[slot] = (uint32_t)a.st.pix_mti[w];
        s = 0;
    }
    a.st.sample[slot] = s;
    Rng jit;
    jit.mode = RNGMODE;
    double u1, u2, u3 = 0.0, u4 = 0.0;
    const int pairs = camera_jitter_pairs(a.cam.kind);
    if (RNGMODE == RNG_MT19937_64) {
        // draws 2s and 2s+1 of the stream, taken when it was seeded (k_wf_seed); a CCD's direction draws follow the point
        // draws of ALL the task's samples: 2*spp + 2s, 2*spp + 2s + 1
        const double* jit_row = a.st.pix_jitter + (size_t)a.st.work[slot] * spp * 2 * pairs;
        const double2 j2 = *reinterpret_cast<const double2*>(jit_row + 2 * (size_t)s);
        u1 = j2.x;
        u2 = j2.y;
        if (pairs == 2) {
            const double2 j4 = *reinterpret_cast<const double2*>(jit_row + 2 * (size_t)spp + 2 * (size_t)s);
            u3 = j4.x;
            u4 = j4.y;
        }
    } else {
        long long pixel_id = (long long)py * a.cam.nx + px;
        jit.px.init(a.seed + (unsigned long long)wf_group_of(a, slot) * a.seed_stride, (unsigned long long)pixel_id, (uint32_t)s);
        if (camera_pixel_draws(a.cam, px, py)) {
            u1 = jit.uniform();
            u2 = jit.uniform();
        } else {
            u1 = u2 = 0.0;
        }
        if (pairs == 2) { u3 = jit.uniform(); u4 = jit.uniform(); }
        a.st.philox_idx[slot] = jit.px.idx;
    }
    V3 o, d;
    double weight;
    pinhole_ray(a.cam, px, py, u1, u2, &o, &d, &weight, u3, u4);
    const size_t P = (size_t)a.n_slots;
    a.st.ray[0 * P + slot] = o.x; a.st.ray[1 * P + slot] = o.y; a.st.ray[2 * P + slot] = o.z;
    a.st.ray[3 * P + slot] = d.x; a.st.ray[4 * P + slot] = d.y; a.st.ray[5 * P + slot] = d.z;
    a.st.weight[slot] = weight;
    a.st.depth[slot] = 0;
    a.st.rays[slot] = 1;
    a.st.log_n[slot] = 0;
    if (a.has_additive) a.st.additive[slot] = 0;
    a.st.status[slot] = SLOT_ALIVE;
    // roulette of the primary segment (see k_wf_trace); with the usual extinction_min_depth > 0 there is no draw
    double normalisation = 1.0;
    if (a.cfg.extinction_min_depth <= 0) {
        Rng rng;
        wf_load_rng<RNGMODE>(a, slot, rng);
        if (!path_roulette(a.cfg, 0, rng, &normalisation)) normalisation = 0.0;
        wf_store_rng<RNGMODE>(a, slot, rng);
    }
    a.st.norm[slot] = normalisation;
}

template <int RNGMODE>
__global__ void __launch_bounds__(128) k_wf_init(const __grid_constant__ WfArgs a) {
    int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= a.n_slots) return;
    a.st.sample[slot] = a.cam.pixel_samples;   // forces a pixel fetch
    a.st.px[slot] = 0;
    a.st.py[slot] = 0;
    a.st.work[slot] = 0;
    wf_regenerate<RNGMODE>(a, slot);
}

// Warp-aggregated append of every lane's slot to list `list` (0..3: per-material hit lists, 4: ended list, < 0:
// nothing): ballots first, then lanes 0..4 issue the (up to) five atomics of the warp together, then every lane
// takes its place.  Must be reached by all 32 lanes.
__device__ __forceinline__ void wf_append_lists(const WfArgs& a, int list, int slot) {
    const int lane = threadIdx.x & 31;
    const size_t P = (size_t)a.n_slots;
    unsigned masks[5];
#pragma unroll
    for (int l = 0; l < 5; ++l) masks[l] = __ballot_sync(RSB_FULL_MASK, list == l);
    unsigned int base = 0;
    if (lane < 5) {
        unsigned mine = lane == 0 ? masks[0] : lane == 1 ? masks[1] : lane == 2 ? masks[2] : lane == 3 ? masks[3] : masks[4];
        if (mine) {
            unsigned int* ctr = (lane < 4) ? &a.st.n_hit[lane] : &a.st.n_ended[a.wave & 1];
            base = atomicAdd(ctr, (unsigned int)__popc(mine));
        }
    }
    unsigned int b = __shfl_sync(RSB_FULL_MASK, base, list < 0 ? 0 : list);
    if (list >= 0) {
        unsigned m = list == 0 ? masks[0] : list == 1 ? masks[1] : list == 2 ? masks[2] : list == 3 ? masks[3] : masks[4];
        unsigned int k = b + __popc(m & ((1u << lane) - 1));
        if (list < 4) a.st.hit_list[(size_t)list * P + k] = slot;
        else a.st.ended[(size_t)(a.wave & 1) * P + k] = slot;
    }
}

// 1 thread = 1 slot.  (Round 2 tried the visit-per-trip walk of rsb_trav.cuh -- k_rq_world with the slots as queries -- in
// this kernel's place for the Cornell scene: 0.294 vs 0.220 ms per wave, 1,199 vs 1,471 Mrays/s for the frame: a staged
// 89-node tree is too short a walk to pay for persistent lanes, and the lists it emits are less ordered.)
// (Persistent lanes with batched refill -- the form k_hit_sweep uses -- were tried here twice.
// Round-1 first attempt: 208 vs 129 us per 262k-ray wave.  Second attempt, after the kernel had lost its RNG
// state and most of its instructions: 14.4 instead of 10.9 active lanes per instruction and 18 % fewer warp
// instructions, yet 143 vs 131 us per 1M-ray wave, and the lists it emits are less ordered, which cost
// k_wf_shade / k_wf_regen another 15-30 %.  A Cornell ray is ~3 traversal units long: too short to amortise the
// flush + refill.  See profiles/README.md.)
template <int RNGMODE, bool COUNT, int FEAT>
__global__ void __launch_bounds__(128, (FEAT & RSB_FEAT_CSG) ? RSB_TRACE_MIN_BLOCKS : ((FEAT & RSB_FEAT_MESH) ? RSB_MESH_MIN_BLOCKS : 5)) k_wf_trace(const __grid_constant__ WfArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    Scene sc = a.sc;
    double* axbuf = ax_storage<(FEAT & RSB_FEAT_STAGED) != 0>(smem, sc, a.n_items);
    typename StatsSel<COUNT>::type stats;
    const int P = a.n_slots;
    int slot = blockIdx.x * blockDim.x + threadIdx.x;
    // The Russian roulette of this segment was already played by the kernel that produced the ray (k_wf_shade /
    // k_wf_regen, right after their own draws -- the reference's draw order, ray.pyx:380-388): norm == 0 marks a
    // path it ended.  The trace kernel therefore touches no RNG state: nothing but slot-indexed, coalesced loads
    // stand before the traversal.
    const size_t PP = (size_t)P;
    const bool in_range = slot < P;
    const int cs = in_range ? slot : 0;
    const int status = a.st.status[cs];
    // Sparse waves (the drain at the end of a frame, where a few long paths are left -- a third of the waves of an
    // 8-GPU frame): a CTA without a live slot leaves before it stages the scene.  9,472 CTAs staging 4.5 KB each were
    // the floor of a nearly empty wave (0.125 ms for an eighth of the rays of a 0.220 ms wave).
    if (!__syncthreads_or(in_range && status == SLOT_ALIVE)) return;
    const double normalisation = a.st.norm[cs];
    PathState ps;
    ps.o = v3(a.st.ray[0 * PP + cs], a.st.ray[1 * PP + cs], a.st.ray[2 * PP + cs]);
    ps.d = v3(a.st.ray[3 * PP + cs], a.st.ray[4 * PP + cs], a.st.ray[5 * PP + cs]);
    // (the slot loads above are in flight while the CTA stages the scene)
    stage_scene<(FEAT & RSB_FEAT_STAGED) != 0>(sc, smem, a.n_items);
    const bool active = in_range && status == SLOT_ALIVE;
    unsigned long long hits = 0;
    int list = -1;
    KdStackEntry stack[RSB_KD_STACK];
    HitRec rec;
    bool hit = false;
#ifdef RSB_NO_COOP
    if constexpr (false) {
#else
    if constexpr ((FEAT & RSB_FEAT_MESH) != 0) {
#endif
        // scenes with meshes: the two-level traversal advances in warp-wide trips, one mesh traversal unit per trip,
        // with the triangle tests of all lanes' leaves pooled (mesh_leaf_coop); every lane of the warp takes part
        // in the pooled tests, whether or not it still has a ray of its own
        NestedTraversal<FEAT, RSB_RENDER_THREADS, typename StatsSel<COUNT>::type> t;
        t.init(sc, a.cfg.max_distance, stack, &rec, stats, axbuf);
        const CoopSmem cs = coop_carve(axbuf);
        bool tracing = false;
        if (active && normalisation != 0.0) {
            hits = 1;
            tracing = t.begin(ps.o, ps.d);
        }
        while (__any_sync(RSB_FULL_MASK, tracing)) {
            if (tracing) {
                t.step_descend();
                coop_publish(cs, t.mleaf.rs, (int)(t.mleaf.mesh - sc.meshes));
            }
            const double d0 = a.cfg.max_distance < t.mc.max_range ? a.cfg.max_distance : t.mc.max_range;
            const bool leaf_hit = mesh_leaf_coop(sc, cs, tracing, t.m_off, t.m_cnt, d0, a.cfg.max_distance, &t.mh, stats);
            if (tracing) tracing = t.step_resolve(leaf_hit);
        }
        if (hits) hit = t.finish();
    } else {
        if (active && normalisation != 0.0) {
            hits = 1;
            hit = world_hit_ax<FEAT, RSB_RENDER_THREADS>(sc, ps.o, ps.d, a.cfg.max_distance, stack, &rec, stats, axbuf);
        }
    }
    if (active) {
        if (hit) {
            a.st.hit_t[slot] = rec.t;
            a.st.hit_a[slot] = make_int4(rec.prim, rec.leaf, rec.code, rec.flip);
            a.st.hit_uvw[slot] = make_float4(rec.u, rec.v, rec.w, __int_as_float(rec.mesh_node));
            a.st.status[slot] = SLOT_HIT;
            list = a.sp.mats[sc.prims[rec.prim].material].type;     // 0..3: per-material hit lists
            // conductors and null surfaces share the specular list, the rough conductor (a ContinuousBSDF) Lambert's
            if (list >= MAT_CONDUCTOR) list = (list == MAT_ROUGH_CONDUCTOR) ? MAT_LAMBERT : MAT_DIELECTRIC;
        } else {
            a.st.status[slot] = SLOT_ENDED_ZERO;
            list = 4;                                                // ended list
        }
    }
    // compact into the five lists (one atomic per warp and list)
    __syncwarp();
    wf_append_lists(a, list, slot);
    if (COUNT) {
        __syncwarp();
        hits = warp_sum(hits);
        if ((threadIdx.x & 31) == 0 && hits) atomicAdd(&a.counters->rays, hits);
        flush_stats(stats, a.counters);
    }
}

// 1 thread = 1 hit slot.  The trace kernel compacts hit slots into one list per material family; every CTA
// walks the lists in turn (Lambert, dielectric, emitter, absorber), so a warp executes ONE BSDF at a time
// while the whole GPU stays busy (separate launches per family left the smaller lists under-occupied).
template <int RNGMODE, bool COUNT, int MAT, int FEAT>
__device__ __forceinline__ void wf_shade_list(const WfArgs& a, const Scene& sc, const Spectral& sp, KdStackEntry* stack,
                                              typename StatsSel<COUNT>::type& stats, uint64_t* mt_col) {
    const size_t P = (size_t)a.n_slots;
    const unsigned int stride = gridDim.x * blockDim.x;
    const unsigned int n = a.st.n_hit[MAT];
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        int slot = a.st.hit_list[(size_t)MAT * P + k];
        Rng rng;
        wf_load_rng<RNGMODE>(a, slot, rng);
        PathState ps;
        ps.o = v3(a.st.ray[0 * P + slot], a.st.ray[1 * P + slot], a.st.ray[2 * P + slot]);
        ps.d = v3(a.st.ray[3 * P + slot], a.st.ray[4 * P + slot], a.st.ray[5 * P + slot]);
        ps.depth = a.st.depth[slot];
        ps.rays = a.st.rays[slot];
        HitRec rec;
        rec.t = a.st.hit_t[slot];
        int4 h = a.st.hit_a[slot];
        float4 uvw = a.st.hit_uvw[slot];
        rec.prim = h.x; rec.leaf = h.y; rec.code = h.z; rec.flip = h.w;
        rec.u = uvw.x; rec.v = uvw.y; rec.w = uvw.z;
        rec.mesh_node = __float_as_int(uvw.w);
        rec.node = -1;
        // (after the slot-state loads above, so that those are already in flight)
        if (RNGMODE == RNG_MT19937_64 && (MAT == MAT_LAMBERT || MAT == MAT_DIELECTRIC)) mt_preload(rng.mt, mt_col);
        PathLog log;
        log.base = a.st.log + (size_t)slot * a.log_capacity;
        log.stride = 1;
        log.capacity = a.log_capacity;
        log.n = a.st.log_n[slot];
        log.overflow = 0;
        log.additive = 0;
        // (concurrent spectral slices: the slot's slice has its own material rows -- refractive indices -- and tables)
        const Spectral sps = a.n_slices > 1 ? wf_slice_spectral(sp, wf_slice_of(a, slot)) : sp;
        int r = path_shade<MAT, FEAT>(sc, sps, a.cfg, ps, rec, a.st.norm[slot], rng, stack, log, stats);
        a.st.log_n[slot] = log.n;
        if (log.overflow) atomicExch(a.overflow_flag, 1);
        if ((FEAT & RSB_FEAT_RARE_MATERIALS) && log.additive) a.st.additive[slot] = 1;
        if (r == PATH_CONTINUE) {
            // the daughter's Russian roulette, played here so that k_wf_trace needs no RNG state: it is the next
            // draw of the stream in the reference too (spawn_daughter -> daughter.trace -> roulette, ray.pyx:380-388);
            // a daughter spawned by a NullSurface is traced with keep_alive=True: no roulette (material.pyx:147)
            double normalisation;
            if ((FEAT & RSB_FEAT_RARE_MATERIALS) && ps.keep_alive) normalisation = 1.0;
            else if (!path_roulette(a.cfg, ps.depth, rng, &normalisation)) normalisation = 0.0;
            wf_store_rng<RNGMODE>(a, slot, rng);
            a.st.norm[slot] = normalisation;
            a.st.ray[0 * P + slot] = ps.o.x; a.st.ray[1 * P + slot] = ps.o.y; a.st.ray[2 * P + slot] = ps.o.z;
            a.st.ray[3 * P + slot] = ps.d.x; a.st.ray[4 * P + slot] = ps.d.y; a.st.ray[5 * P + slot] = ps.d.z;
            a.st.depth[slot] = ps.depth;
            a.st.rays[slot] = ps.rays;
            a.st.status[slot] = SLOT_ALIVE;
        } else {
            wf_store_rng<RNGMODE>(a, slot, rng);
            a.st.status[slot] = (r == PATH_EMITTED) ? SLOT_ENDED_EMIT : SLOT_ENDED_ZERO;
            wf_push_ended(a, slot);
        }
    }
}

template <int RNGMODE, bool COUNT, int FEAT>
__global__ void __launch_bounds__(128, RSB_SHADE_MIN_BLOCKS) k_wf_shade(const __grid_constant__ WfArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    {
        // sparse waves (the tail of a frame, small per-rank shares): a CTA whose first index lies past the end of
        // all four lists has nothing to shade -- leave before staging the scene (uniform across the CTA)
        const unsigned int first = blockIdx.x * blockDim.x;
        if (first >= a.st.n_hit[0] && first >= a.st.n_hit[1] && first >= a.st.n_hit[2] && first >= a.st.n_hit[3]) return;
    }
    Scene sc = a.sc;
    Spectral sp = a.sp;
    constexpr bool STAGED = (FEAT & RSB_FEAT_STAGED) != 0;
    stage_scene<STAGED>(sc, smem, a.n_items);
    int smem_used = STAGED ? stage_layout(a.sc.world.n_nodes, a.n_items, a.sc.n_prims).total : 0;
    if (a.tables_staged) {
        unsigned char* base = smem + smem_used;
        int mat_bytes = ((sp.n_materials * (int)sizeof(Material) + 15) / 16) * 16;
        __shared__ __align__(8) uint64_t mat_bar;
        stage_region(&mat_bar, base, a.sp.mats, mat_bytes);
        sp.mats = reinterpret_cast<const Material*>(base);
        smem_used += mat_bytes;
    }
    // the thread's column of the MT19937-64 state window (mt_preload)
    uint64_t* mt_col = reinterpret_cast<uint64_t*>(smem + smem_used) + threadIdx.x;
    typename StatsSel<COUNT>::type stats;
    KdStackEntry stack[RSB_KD_STACK];
    wf_shade_list<RNGMODE, COUNT, MAT_LAMBERT, FEAT>(a, sc, sp, stack, stats, mt_col);
    wf_shade_list<RNGMODE, COUNT, MAT_DIELECTRIC, FEAT>(a, sc, sp, stack, stats, mt_col);
    wf_shade_list<RNGMODE, COUNT, MAT_EMITTER, FEAT>(a, sc, sp, stack, stats, mt_col);
    wf_shade_list<RNGMODE, COUNT, MAT_ABSORBER, FEAT>(a, sc, sp, stack, stats, mt_col);
    if (COUNT) {
        __syncwarp();
        flush_stats(stats, a.counters, true);
    }
}

// The term a bin contributes to every projection channel, in the pixel processor's own operation order:
//   XYZ       delta * sample * curve                      (spectrum_to_ciexyz, colour.pyx:182-184)
//   POWER     sample * filter * sensitivity * delta       (PowerPixelProcessor.add_sample, mono/power.pyx:776-777)
//   RADIANCE  sample * filter * delta                     (RadiancePixelProcessor.add_sample, mono/radiance.pyx:192-193)
template <int NCH>
__device__ __forceinline__ void proj_terms(const WfArgs& a, int nch, double x, double delta, const double* __restrict__ curve,
                                           double* __restrict__ out) {
    if constexpr (NCH == 3) {
        const double dx = delta * x;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) out[ch] = dx * __ldg(curve + ch);
    } else {
        for (int ch = 0; ch < nch; ++ch) {
            const double cv = __ldg(curve + ch);
            const int mode = a.proj_mode[ch];
            double t;
            if (mode == RSB_PROJ_XYZ) t = (delta * x) * cv;
            else if (mode == RSB_PROJ_POWER) t = ((x * cv) * a.cam.sensitivity) * delta;
            else t = (x * cv) * delta;
            out[ch] = t;
        }
    }
}

// 1 warp = 1 ended path: the reference's unwind (per-bin multiplies), projection weight, sensitivity and
// PixelProcessor.add_sample (Welford) for bins lane, lane+32, ...
// NCH: projection channels compiled in -- 0: none (the spectral pipelines' kernel, nothing of the projection code in it: the
// generic form cost it 14 registers and 5 % of its time), 3: the three CIE XYZ channels of an RGB pipeline alone (unrolled),
// -1: a.proj_channels channels of any mode.
template <int RNGMODE, bool COUNT, int NCH>
__global__ void __launch_bounds__(128) k_wf_finalize(const __grid_constant__ WfArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    Spectral sp = a.sp;
    int tab_bytes = 0;
    if (a.tables_staged) {
        // tables and their logs are contiguous in HBM: [n_tables][bins] x 2
        tab_bytes = ((sp.n_tables * sp.bins * 16 + 15) / 16) * 16;
        __shared__ __align__(8) uint64_t tab_bar;
        stage_region(&tab_bar, smem, a.sp.tables, tab_bytes);
        sp.tables = reinterpret_cast<const double*>(smem);
        sp.tables_ln = sp.tables + (size_t)sp.n_tables * sp.bins;
    }
    // 32 log entries per warp, staged in shared memory and read back as one broadcast LDS.128 per entry
    LogEntry* wlog = reinterpret_cast<LogEntry*>(smem + tab_bytes) + (threadIdx.x >> 5) * 32;
    // projections (RGB / mono pipelines): the sample's per-bin terms of every channel, one row of channels * bins doubles
    // per warp, behind the log windows
    const int nch = NCH >= 0 ? NCH : a.proj_channels;
    double* wspec = reinterpret_cast<double*>(smem + tab_bytes + (blockDim.x >> 5) * 32 * sizeof(LogEntry)) + (size_t)(threadIdx.x >> 5) * nch * sp.bins;
    const bool keep_bins = a.mean != nullptr, keep_xyz = NCH != 0;
    const int par = a.wave & 1;
    const int lane = threadIdx.x & 31;
    const unsigned int n = a.st.n_ended[par];
    if (blockIdx.x == 0 && threadIdx.x == 0) a.st.n_ended[par ^ 1] = 0;   // list of the next wave
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int bins = sp.bins;
    for (unsigned int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += warps) {
        int slot = a.st.ended[(size_t)par * a.n_slots + k];
        int status = a.st.status[slot];
        int s = a.st.sample[slot];
        double w = a.st.weight[slot];
        PathLog log;
        log.base = a.st.log + (size_t)slot * a.log_capacity;
        log.stride = 1;
        log.capacity = a.log_capacity;
        log.n = a.st.log_n[slot];
        log.overflow = 0;
        // frame rows hold frame_bins = n_slices * bins values: slice k owns bins [k * bins, (k + 1) * bins)
        const int group = wf_group_of(a, slot);
        const int pass = a.n_slices > 1 ? group / a.n_slices : group;
        const int slice = a.n_slices > 1 ? group % a.n_slices : 0;
        size_t row = ((size_t)a.st.px[slot] * a.cam.ny + a.st.py[slot]) * a.frame_bins + (size_t)slice * bins;
        const Spectral sps = a.n_slices > 1 ? wf_slice_spectral(sp, slice) : sp;
        double* m = keep_bins ? (pass ? a.pass_mean + (size_t)(pass - 1) * a.frame_elems : a.mean) + row : nullptr;
        double* v = keep_bins ? (pass ? a.pass_variance + (size_t)(pass - 1) * a.frame_elems : a.variance) + row : nullptr;
        const double r_nn = 1.0 / (double)(s + 1), r_nn1 = s > 0 ? 1.0 / (double)s : 0.0;
        // (with emitting volumes in the scene a path that ended dark may still carry what they added on the way)
        const bool emit = status == SLOT_ENDED_EMIT || (a.has_additive && a.st.additive[slot] != 0);
        const double xyz_delta = keep_xyz ? a.xyz_delta[slice] : 0.0;
        const double* xyz_curve = keep_xyz ? a.xyz_tab + (size_t)slice * bins * nch : nullptr;
        for (int b0 = 0; b0 < bins; b0 += 64) {
            // two bins per lane per pass; the statistics rows are fetched before the replay so that their
            // latency overlaps it
            const int ba = b0 + lane, bb = b0 + 32 + lane;
            const bool ha = ba < bins, hb = bb < bins;
            double ma = 0, va = 0, mb = 0, vb = 0;
            if (s > 0 && keep_bins) {
                if (ha) { ma = m[ba]; va = v[ba]; }
                if (hb) { mb = m[bb]; vb = v[bb]; }
            }
            double xa = 0.0, xb = 0.0;
            if (emit) {
                // the log is read 32 entries at a time, one entry per lane (coalesced), and handed round through
                // shared memory: every lane applies every entry, newest first, to its own bins
                for (int top = log.n; top > 0; top -= 32) {
                    int cnt = top < 32 ? top : 32;
                    __syncwarp();
                    if (lane < cnt) wlog[lane] = log.get(top - 1 - lane);
                    __syncwarp();
                    for (int j = 0; j < cnt; ++j) {
                        const LogEntry e = wlog[j];
                        if (ha) xa = apply_entry(xa, e.op, e.table, e.v, sps, ba);
                        if (hb) xb = apply_entry(xb, e.op, e.table, e.v, sps, bb);
                    }
                }
            }
            if (ha) {
                double x = xa * w;              // spectrum.mul_scalar(projection_weight), observer.pyx:408
                // the terms of the pixel processors' sums, formed by the bin's own lane; only the additions below are a
                // serial chain
                if (keep_xyz) proj_terms<NCH>(a, nch, x, xyz_delta, xyz_curve + (size_t)nch * ba, wspec + (size_t)nch * ba);
                if (keep_bins) {
                    x = x * a.cam.sensitivity;      // add_sample(spectrum, sensitivity), power.pyx:478-481
                    welford_add_r(x, ma, va, s, r_nn, r_nn1, m + ba, v + ba);
                }
            }
            if (hb) {
                double x = xb * w;
                if (keep_xyz) proj_terms<NCH>(a, nch, x, xyz_delta, xyz_curve + (size_t)nch * bb, wspec + (size_t)nch * bb);
                if (keep_bins) {
                    x = x * a.cam.sensitivity;
                    welford_add_r(x, mb, vb, s, r_nn, r_nn1, m + bb, v + bb);
                }
            }
        }
        if (keep_xyz) {
            // XYZPixelProcessor.add_sample (rgb.pyx:550-558): spectrum_to_ciexyz sums delta * sample * curve over the bins
            // in index order (colour.pyx:178-186) -- a serial chain of additions, so lanes 0..channels-1 walk one channel each
            // (the terms are in shared memory already: the chain is bins x one DADD) -- then the tristimulus value times
            // the sensitivity enters the channel's running statistics
            __syncwarp();
            if (lane < nch) {
                double acc = 0.0;
#pragma unroll 8
                for (int i = 0; i < bins; ++i) acc += wspec[nch * i + lane];
                const size_t item = ((size_t)a.item_base + (size_t)a.st.work[slot]) * nch + lane;
                double pm = 0, pv = 0;
                if (s > 0) { pm = a.xyz_mean[item]; pv = a.xyz_variance[item]; }
                // (XYZ: the tristimulus value times the sensitivity, rgb.pyx:556-558; the mono processors fold the
                // sensitivity into their terms or ignore it)
                const double value = (NCH == 3 || a.proj_mode[lane] == RSB_PROJ_XYZ) ? acc * a.cam.sensitivity : acc;
                welford_add_r(value, pm, pv, s, r_nn, r_nn1, a.xyz_mean + item, a.xyz_variance + item);
            }
            __syncwarp();
        }
    }
}

// RGBPipeline2D / PowerPipeline2D / RadiancePipeline2D .update + finalise for every listed pixel (rgb.pyx:249-290,
// mono/power.pyx:516-556): the per-slice (mean, variance) of a pass are summed in slice order into the pass's working
// values, which are merged into the frame with combine_samples; the passes of a pixel are merged in order by the same
// thread, one thread per (pixel, channel).  The frame holds `nc` channels per pixel, channels [c0, c0 + nc) of the `nch`
// projection channels of the render.
// `bayer`: BayerPipeline2D (pipeline/bayer.pyx:339-349): ONE value per pixel, taken from channel c0 + mosaic[(x % 2) + 2 (y % 2)]
// of the three filter channels, mosaic = (0, 1, 1, 2) = red, green / green, blue (bayer.pyx:109).
__global__ void k_xyz_combine(long long n_pixels, const int32_t* __restrict__ pixels, int ny, int n_passes, int n_slices, int samples,
                              int nch, int c0, int nc, int bayer, const double* __restrict__ xyz_mean, const double* __restrict__ xyz_variance,
                              double* __restrict__ fmean, double* __restrict__ fvar, int32_t* __restrict__ fsamples) {
    long long total = n_pixels * nc;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        long long p = i / nc;
        int ch = (int)(i % nc);
        long long row = pixels ? ((long long)pixels[2 * p] * ny + pixels[2 * p + 1]) : p;
        long long dst = row * nc + ch;
        if (bayer) {
            const int x = (int)(row / ny), y = (int)(row % ny);
            const int index = (x % 2) + 2 * (y % 2);
            ch = index == 0 ? 0 : (index == 3 ? 2 : 1);
        }
        double fm = fmean[dst], fv = fvar[dst];
        int fn = fsamples[dst];
        for (int pass = 0; pass < n_passes; ++pass) {
            double wm = 0.0, wv = 0.0;
            for (int k = 0; k < n_slices; ++k) {
                long long src = (((long long)pass * n_slices + k) * n_pixels + p) * nch + c0 + ch;
                wm += xyz_mean[src];
                wv += xyz_variance[src];
            }
            if (wv < 0) wv = 0;   // statsarray.pyx:647-650
            double mt, vt;
            int nt;
            stats_combine(fm, fv, fn, wm, wv, samples, &mt, &vt, &nt);
            fm = mt; fv = vt; fn = nt;
        }
        fmean[dst] = fm;
        fvar[dst] = fv;
        fsamples[dst] = fn;
    }
}

// 1 thread = 1 ended slot: the ray counter of the finished path (observer.pyx:414), then the slot's next
// sample or next pixel.  Separate from the accumulate kernel so that ray generation runs converged.
template <int RNGMODE, bool COUNT>
__global__ void __launch_bounds__(128) k_wf_regen(const __grid_constant__ WfArgs a) {
    const int par = a.wave & 1;
    const unsigned int n = a.st.n_ended[par];
    const unsigned int stride = gridDim.x * blockDim.x;
    if (blockIdx.x == 0 && threadIdx.x < 4) a.st.n_hit[threadIdx.x] = 0;   // consumed by this wave's shade kernels
    unsigned long long rays = 0, paths = 0;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        int slot = a.st.ended[(size_t)par * a.n_slots + k];
        rays += a.st.rays[slot];
        paths += 1;
        a.st.sample[slot] = a.st.sample[slot] + 1;
        wf_regenerate<RNGMODE>(a, slot);
    }
    __syncwarp();
    rays = warp_sum(rays);
    paths = warp_sum(paths);
    if ((threadIdx.x & 31) == 0 && paths) {
        atomicAdd(a.ray_count, rays);
        if (COUNT) atomicAdd(&a.counters->paths, paths);
    }
}

// Several GPUs driven from one process (rsb_comm_gather_slices): the root device pulls the frame rows of a PEER's listed
// pixels out of the peer's memory -- plain loads through the peer mapping, i.e. over NVLink / NVSwitch -- and drops them in
// place in its own slice.  One thread per 16 bytes; rows are copied, so the root holds bit for bit what the owner wrote.
__global__ void k_gather_peer_rows(long long n_pixels, const int32_t* __restrict__ peer_pixels, int ny, int bins,
                                   const double* __restrict__ peer_mean, const double* __restrict__ peer_variance,
                                   double* __restrict__ mean, double* __restrict__ variance) {
    const long long total = n_pixels * bins;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long p = i / bins;
        const int b = (int)(i % bins);
        const long long at = ((long long)peer_pixels[2 * p] * ny + peer_pixels[2 * p + 1]) * bins + b;
        mean[at] = peer_mean[at];
        variance[at] = peer_variance[at];
    }
}

// StatsArray3D.combine_samples over the listed pixels of a slice (statsarray.pyx:780-857, power.pyx:424-437)
__global__ void k_frame_combine(long long n_pixels, const int32_t* __restrict__ pixels, int ny, int frame_bins, int slice_offset,
                                int slice_bins, const double* __restrict__ mean, const double* __restrict__ variance, int samples,
                                double* __restrict__ fmean, double* __restrict__ fvar, int32_t* __restrict__ fsamples) {
    long long total = n_pixels * slice_bins;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        long long p = i / slice_bins;
        int b = (int)(i % slice_bins);
        long long row = pixels ? ((long long)pixels[2 * p] * ny + pixels[2 * p + 1]) : p;
        long long src = row * slice_bins + b;
        long long dst = row * frame_bins + slice_offset + b;
        double mt, vt;
        int nt;
        // frame.combine_samples(x, y, z, mean, variance, samples): set x = frame, set y = new slice
        double vb = variance[src];
        if (vb < 0) vb = 0;   // statsarray.pyx:647-650
        stats_combine(fmean[dst], fvar[dst], fsamples[dst], mean[src], vb, samples, &mt, &vt, &nt);
        fmean[dst] = mt;
        fvar[dst] = vt;
        fsamples[dst] = nt;
    }
}

// P accumulated observe() calls rendered concurrently: pass p (p >= 1) of every listed pixel is merged into the
// caller's frame, which holds passes 0..p-1 (n_acc samples), exactly as SpectralPowerPipeline2D.update does
// between observe() calls (power.pyx:424-437 -> statsarray.pyx:780-857).  The passes of one pixel are merged
// in order by the same thread.
__global__ void k_pass_combine(long long n_pixels, const int32_t* __restrict__ pixels, int ny, int bins, int n_passes, int samples,
                               long long frame_elems, const double* __restrict__ pass_mean, const double* __restrict__ pass_variance,
                               double* __restrict__ fmean, double* __restrict__ fvar) {
    long long total = n_pixels * bins;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        long long p = i / bins;
        int b = (int)(i % bins);
        long long row = pixels ? ((long long)pixels[2 * p] * ny + pixels[2 * p + 1]) : p;
        long long e = row * bins + b;
        // the first update merges pass 0 into the empty frame (n = 0), like every later one
        double m = 0.0, v = 0.0;
        int n = 0;
        for (int q = 0; q < n_passes; ++q) {
            double mb = q ? pass_mean[(long long)(q - 1) * frame_elems + e] : fmean[e];
            double vb = q ? pass_variance[(long long)(q - 1) * frame_elems + e] : fvar[e];
            if (vb < 0) vb = 0;   // statsarray.pyx:647-650
            double mt, vt;
            int nt;
            stats_combine(m, v, n, mb, vb, samples, &mt, &vt, &nt);
            m = mt; v = vt; n = nt;
        }
        fmean[e] = m;
        fvar[e] = v;
    }
}

}  // namespace rsb
