// rsb_kernels.cuh -- sm_100a kernels of the hot path.
//
//  k_hit_batch      World.hit over a ray batch (1 thread = 1 ray, grid-stride)
//  k_hit_sweep      same, rays generated on device from (seed, index), results reduced on device
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
#define RSB_RENDER_THREADS 128

struct DevCounters {   // mirrors RsbCounters
    unsigned long long rays, branches, leaves, items, prim_tests, tri_tests, paths, contains, table_reads, reserved[3];
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

__device__ __forceinline__ void flush_stats(const NoStats&, DevCounters*) {}
__device__ __forceinline__ void flush_stats(const CountStats& s, DevCounters* c) {
    // all lanes of the warp reach here together (kernel epilogue)
    unsigned long long b = warp_sum(s.branches), l = warp_sum(s.leaves), i = warp_sum(s.items),
                       p = warp_sum(s.prim_tests), t = warp_sum(s.tri_tests), cq = warp_sum(s.contains),
                       tb = warp_sum(s.tables);
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

__device__ __forceinline__ void copy16(void* dst, const void* src, int bytes) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) d[i] = s[i];
}

// n_items: number of leaf item ids of the world tree; staged == 0 leaves the scene untouched
__device__ __forceinline__ void stage_scene(Scene& sc, unsigned char* smem, int n_items, int staged) {
    if (!staged) return;
    StageLayout l = stage_layout(sc.world.n_nodes, n_items, sc.n_prims);
    copy16(smem, sc.world.nodes, l.nodes_bytes);
    // item list: copy whole 16-B words (the allocation is padded to 16 B)
    copy16(smem + l.nodes_bytes, sc.world.items, l.items_bytes);
    copy16(smem + l.nodes_bytes + l.items_bytes, sc.prims, l.prims_bytes);
    __syncthreads();
    sc.world.nodes = reinterpret_cast<const KdNode*>(smem);
    sc.world.items = reinterpret_cast<const int32_t*>(smem + l.nodes_bytes);
    sc.prims = reinterpret_cast<const Prim*>(smem + l.nodes_bytes + l.items_bytes);
}

// ---------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(128)
k_hit_batch(Scene sc, int n_items, int staged, long long n, const double* __restrict__ origins,
            const double* __restrict__ directions, const double* __restrict__ max_distance,
            int32_t* __restrict__ out_prim, double* __restrict__ out_t, int32_t* __restrict__ out_sub,
            uint8_t* __restrict__ out_flags, int32_t* __restrict__ out_node, double* __restrict__ out_geom,
            float* __restrict__ out_uvw, DevCounters* counters) {
    extern __shared__ __align__(16) unsigned char smem[];
    stage_scene(sc, smem, n_items, staged);
    typename StatsSel<COUNT>::type stats;
    KdStackEntry stack[RSB_KD_STACK];
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        V3 o = v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        V3 d = v3(directions[3 * i], directions[3 * i + 1], directions[3 * i + 2]);
        double md = max_distance ? max_distance[i] : RSB_INF;
        HitRec rec;
        bool hit = world_hit(sc, o, d, md, stack, &rec, stats);
        if (hit) {
            Isect is;
            world_hit_geometry(sc, o, d, rec, &is);
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
    if (COUNT) {
        __syncwarp();
        flush_stats(stats, counters);
    }
}

// rays from `origin` toward target + (jx, jy, 0)*half_window, (jx, jy) uniform in [-1, 1) from Philox(seed, index)
template <bool COUNT>
__global__ void __launch_bounds__(128)
k_hit_sweep(Scene sc, int n_items, int staged, long long n, long long first_index, unsigned long long seed,
            double ox, double oy, double oz, double tx, double ty, double tz, double half_window,
            unsigned long long* out_hits, double* out_sum_t, unsigned long long* out_xor_prim, DevCounters* counters) {
    extern __shared__ __align__(16) unsigned char smem[];
    stage_scene(sc, smem, n_items, staged);
    typename StatsSel<COUNT>::type stats;
    KdStackEntry stack[RSB_KD_STACK];
    unsigned long long hits = 0, xr = 0;
    double sum_t = 0.0;
    long long stride = (long long)gridDim.x * blockDim.x;
    V3 o = v3(ox, oy, oz);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Philox4x32 px;
        px.init(seed, (unsigned long long)(first_index + i), 0u);
        double u1 = (double)(px.next_u64() >> 11) * (1.0 / 9007199254740992.0);
        double u2 = (double)(px.next_u64() >> 11) * (1.0 / 9007199254740992.0);
        V3 p = v3(tx + (2.0 * u1 - 1.0) * half_window, ty + (2.0 * u2 - 1.0) * half_window, tz);
        V3 d = normalise(v3(p.x - o.x, p.y - o.y, p.z - o.z));
        HitRec rec;
        if (world_hit(sc, o, d, RSB_INF, stack, &rec, stats)) {
            hits += 1;
            sum_t += rec.t;
            xr ^= (unsigned long long)(unsigned)rec.prim * 0x9E3779B97F4A7C15ULL + (unsigned long long)(first_index + i);
        }
    }
    __syncwarp();
    hits = warp_sum(hits);
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
        sum_t += __shfl_down_sync(RSB_FULL_MASK, sum_t, k);
        xr ^= __shfl_down_sync(RSB_FULL_MASK, xr, k);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out_hits, hits);
        atomicAdd(out_sum_t, sum_t);
        atomicXor(out_xor_prim, xr);
    }
    if (COUNT) flush_stats(stats, counters);
}

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
struct RenderArgs {
    Scene sc;
    Spectral sp;
    RayConfig cfg;
    Camera cam;
    long long n_pixels;
    const int32_t* pixels;          // [n][2] or null
    double* mean;
    double* variance;
    unsigned long long* ray_count;
    unsigned long long* work_counter;
    unsigned long long* mt_state;   // [2][312][T] word-interleaved across the T threads of the grid
    LogEntry* log_pool;             // [capacity][T]
    int32_t* overflow_flag;
    DevCounters* counters;
    unsigned long long seed;
    int32_t log_capacity;
    int32_t n_items;
    int32_t staged;
    int32_t tables_staged;
};

template <int RNGMODE, bool COUNT>
__global__ void __launch_bounds__(RSB_RENDER_THREADS)
k_render(const __grid_constant__ RenderArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    Scene sc = a.sc;
    Spectral sp = a.sp;
    stage_scene(sc, smem, a.n_items, a.staged);
    if (a.tables_staged) {
        // materials + per-slice spectral tables behind the scene data
        StageLayout l = stage_layout(a.staged ? a.sc.world.n_nodes : 0, a.staged ? a.n_items : 0, a.staged ? a.sc.n_prims : 0);
        unsigned char* base = smem + l.total;
        int mat_bytes = ((sp.n_materials * (int)sizeof(Material) + 15) / 16) * 16;
        int tab_bytes = ((sp.n_materials * sp.bins * 8 + 15) / 16) * 16;
        copy16(base, a.sp.mats, mat_bytes);
        copy16(base + mat_bytes, a.sp.tables, tab_bytes);
        __syncthreads();
        sp.mats = reinterpret_cast<const Material*>(base);
        sp.tables = reinterpret_cast<const double*>(base + mat_bytes);
    }

    typename StatsSel<COUNT>::type stats;
    KdStackEntry stack[RSB_KD_STACK];
    const size_t T = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const size_t warp_tid0 = tid - lane;
    const int bins = sp.bins;
    const int spp = a.cam.pixel_samples;

    Rng rng;        // path stream
    Rng jit;        // jitter stream (MT mode: the pixel's stream before the 2*spp jitter draws were consumed)
    rng.mode = RNGMODE;
    jit.mode = RNGMODE;
    rng.mt.mt = a.mt_state ? reinterpret_cast<uint64_t*>(a.mt_state) + tid : nullptr;
    rng.mt.stride = T;
    rng.mt.mti = RSB_MT_NN;
    jit.mt.mt = a.mt_state ? reinterpret_cast<uint64_t*>(a.mt_state) + (size_t)RSB_MT_NN * T + tid : nullptr;
    jit.mt.stride = T;
    jit.mt.mti = RSB_MT_NN;

    PathLog log;
    log.base = a.log_pool + tid;
    log.stride = T;
    log.capacity = a.log_capacity;
    log.n = 0;
    log.overflow = 0;

    PathState ps;
    ps.depth = 0; ps.rays = 0;
    ps.o = v3(0, 0, 0); ps.d = v3(0, 0, 1);
    long long frame_row = 0;     // (x*ny + y): row of the pixel in the frame arrays
    long long pixel_id = 0;      // y*nx + x: RNG stream id
    int px = 0, py = 0;
    int s = spp;                 // next sample index of the current pixel; spp => need a pixel
    bool have_path = false, exhausted = false;
    double weight = 0.0;
    unsigned long long my_rays = 0, my_paths = 0, my_hits = 0;

    for (;;) {
        // ---- regenerate: next sample of this lane's pixel, or a new pixel ---------------------------
        if (!have_path && !exhausted) {
            if (s >= spp) {
                unsigned long long w = atomicAdd(a.work_counter, 1ULL);
                if (w >= (unsigned long long)a.n_pixels) {
                    exhausted = true;
                } else {
                    if (a.pixels) { px = a.pixels[2 * w]; py = a.pixels[2 * w + 1]; }
                    else { px = (int)(w / (unsigned long long)a.cam.ny); py = (int)(w % (unsigned long long)a.cam.ny); }
                    frame_row = (long long)px * a.cam.ny + py;
                    pixel_id = (long long)py * a.cam.nx + px;
                    s = 0;
                    if (RNGMODE == RNG_MT19937_64) {
                        // seed(seed + pixel_id); the jitter cursor starts at draw 0, the path cursor after the
                        // 2*spp draws RectangleSampler3D.samples(spp) consumes up front (pinhole.pyx:183)
                        jit.mt.seed(a.seed + (unsigned long long)pixel_id);
                        for (int i = 0; i < RSB_MT_NN; ++i) rng.mt.w(i) = jit.mt.w(i);
                        rng.mt.mti = RSB_MT_NN;
                        for (int i = 0; i < 2 * spp; ++i) (void)rng.mt.next_u64();
                    }
                }
            }
            if (!exhausted) {
                if (RNGMODE == RNG_PHILOX) {
                    rng.px.init(a.seed, (unsigned long long)pixel_id, (uint32_t)s);
                }
                double u1, u2;
                if (RNGMODE == RNG_MT19937_64) { u1 = jit.uniform(); u2 = jit.uniform(); }
                else { u1 = rng.uniform(); u2 = rng.uniform(); }
                V3 o, d;
                pinhole_ray(a.cam, px, py, u1, u2, &o, &d, &weight);
                path_begin(ps, log, o, d);
                have_path = true;
                my_paths += 1;
            }
        }
        // ---- one segment --------------------------------------------------------------------------
        int result = PATH_CONTINUE;
        bool ended = false;
        if (have_path) {
            my_hits += 1;
            result = path_step(sc, sp, a.cfg, ps, rng, stack, log, stats);
            ended = (result != PATH_CONTINUE);
        }
        // ---- fold finished paths into their pixel's statistics, warp-cooperatively (lane = bin) -------
        __syncwarp();
        unsigned ended_mask = __ballot_sync(RSB_FULL_MASK, ended);
        while (ended_mask) {
            int src = __ffs(ended_mask) - 1;
            ended_mask &= ended_mask - 1;
            int e_n = __shfl_sync(RSB_FULL_MASK, log.n, src);
            int e_res = __shfl_sync(RSB_FULL_MASK, result, src);
            long long e_row = __shfl_sync(RSB_FULL_MASK, frame_row, src);
            int e_s = __shfl_sync(RSB_FULL_MASK, s, src);
            double e_w = __shfl_sync(RSB_FULL_MASK, weight, src);
            PathLog elog;
            elog.base = a.log_pool + warp_tid0 + src;
            elog.stride = T;
            elog.n = e_n;
            elog.capacity = a.log_capacity;
            elog.overflow = 0;
            double* m = a.mean + e_row * bins;
            double* v = a.variance + e_row * bins;
            for (int b = lane; b < bins; b += 32) {
                double x = 0.0;
                if (e_res == PATH_EMITTED) x = replay_bin(elog, sp, b);
                x = x * e_w;                    // spectrum.mul_scalar(projection_weight), observer.pyx:408
                x = x * a.cam.sensitivity;      // add_sample(spectrum, sensitivity), power.pyx:478-481
                welford_add(x, m + b, v + b, e_s);
            }
        }
        __syncwarp();
        if (ended) {
            my_rays += ps.rays;
            s += 1;
            have_path = false;
        }
        if (__all_sync(RSB_FULL_MASK, exhausted && !have_path)) break;
    }
    if (log.overflow) atomicExch(a.overflow_flag, 1);
    // ray counter (observer.pyx:414) + roofline counters
    unsigned long long r = warp_sum(my_rays);
    unsigned long long p = warp_sum(my_paths);
    unsigned long long h = warp_sum(my_hits);
    if (lane == 0) {
        atomicAdd(a.ray_count, r);
        if (COUNT) { atomicAdd(&a.counters->paths, p); atomicAdd(&a.counters->rays, h); }
    }
    if (COUNT) flush_stats(stats, a.counters);
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

}  // namespace rsb
