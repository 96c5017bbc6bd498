// hostsim.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the product's device headers
// (source_b200/csrc/rsb_*.h) as plain host C++ and drives them serially, so that the arithmetic
// the CUDA kernels execute can be pinned against the reference in a container without a GPU.
// It is built by tests/conftest.py into tests/_build/, is never imported by source_b200, and is
// not a fallback: the product library fails loudly when no sm_100 device is present.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../source_b200/csrc/rsb_path.h"
#include "../../source_b200/csrc/rsb_trav.h"
#include "../../source_b200/csrc/scene_pack.h"

using namespace rsb;

namespace {
struct HostScene {
    PackedScene ps;
    std::vector<Mesh> meshes;
    Scene sc;
};
std::string g_err;

void bind(HostScene* h) {
    Scene& sc = h->sc;
    memset(&sc, 0, sizeof(sc));
    sc.prims = h->ps.prims.data();
    sc.n_prims = (int32_t)h->ps.prims.size();
    sc.n_world = h->ps.n_world;
    sc.world.nodes = h->ps.world.nodes.data();
    sc.world.items = h->ps.world.items.data();
    memcpy(sc.world.bounds, h->ps.world.bounds, 48);
    sc.world.n_nodes = (int32_t)h->ps.world.nodes.size();
    sc.world.max_depth = h->ps.world.depth;
    h->meshes.assign(h->ps.meshes.size(), Mesh{});
    for (size_t i = 0; i < h->ps.meshes.size(); ++i) {
        PackedMesh& pm = h->ps.meshes[i];
        Mesh& m = h->meshes[i];
        m.tri = pm.tri.data();
        m.tri_idx = pm.tri_idx.data();
        m.vnormals = pm.vnormals.empty() ? nullptr : pm.vnormals.data();
        m.tree.nodes = pm.tree.nodes.data();
        m.tree.items = pm.tree.items.data();
        memcpy(m.tree.bounds, pm.tree.bounds, 48);
        m.tree.n_nodes = (int32_t)pm.tree.nodes.size();
        m.tree.max_depth = pm.tree.depth;
        m.n_tri = pm.n_tri;
        m.idx_stride = pm.idx_stride;
        m.smoothing = pm.smoothing;
        m.closed = pm.closed;
    }
    sc.meshes = h->meshes.data();
    sc.n_meshes = (int32_t)h->meshes.size();
    sc.n_important = (int32_t)h->ps.imp_weight.size();
    sc.imp_sphere = h->ps.imp_sphere.data();
    sc.imp_weight = h->ps.imp_weight.data();
    sc.imp_cdf = h->ps.imp_cdf.data();
    sc.imp_total = h->ps.imp_total;
}
}  // namespace

extern "C" {

const char* hs_last_error() { return g_err.c_str(); }

int hs_scene_create(const RsbSceneDesc* d, uint64_t* out) {
    HostScene* h = new HostScene();
    int rc = pack_scene(d, &h->ps, &g_err);
    if (rc) { delete h; return rc; }
    bind(h);
    *out = reinterpret_cast<uint64_t>(h);
    return 0;
}

int hs_scene_destroy(uint64_t s) {
    delete reinterpret_cast<HostScene*>(s);
    return 0;
}

// mode 0: world_hit (the nested two-level loop); 1: the split pipeline run serially (world walk suspended in front of
// every Mesh.hit, mesh_query over kd_visit, resume with the per-ray memo); 2: the single-visit walk of mesh-free scenes
int hs_hit_batch_mode(uint64_t scene, int32_t mode, int64_t n, const double* origins, const double* directions, const double* max_distance,
                      int32_t* out_prim, double* out_t, int32_t* out_sub, uint8_t* out_flags, int32_t* out_node, double* out_geom,
                      float* out_uvw, uint64_t* counters /* branches, leaves, items, prim_tests, tri_tests */);

int hs_hit_batch(uint64_t scene, int64_t n, const double* origins, const double* directions, const double* max_distance,
                 int32_t* out_prim, double* out_t, int32_t* out_sub, uint8_t* out_flags, int32_t* out_node, double* out_geom,
                 float* out_uvw, uint64_t* counters) {
    return hs_hit_batch_mode(scene, 0, n, origins, directions, max_distance, out_prim, out_t, out_sub, out_flags, out_node, out_geom,
                             out_uvw, counters);
}

int hs_hit_batch_mode(uint64_t scene, int32_t mode, int64_t n, const double* origins, const double* directions, const double* max_distance,
                      int32_t* out_prim, double* out_t, int32_t* out_sub, uint8_t* out_flags, int32_t* out_node, double* out_geom,
                      float* out_uvw, uint64_t* counters) {
    HostScene* h = reinterpret_cast<HostScene*>(scene);
    CountStats stats;
    KdStackEntry stack[RSB_KD_STACK];
    for (int64_t i = 0; i < n; ++i) {
        V3 o = v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        V3 d = v3(directions[3 * i], directions[3 * i + 1], directions[3 * i + 2]);
        double md = max_distance ? max_distance[i] : RSB_INF;
        HitRec rec;
        bool hit;
        if (mode == 1) hit = world_hit_split<RSB_FEAT_ALL>(h->sc, o, d, md, stack, &rec, stats);
        else if (mode == 2) hit = world_hit_visits<RSB_FEAT_ALL>(h->sc, o, d, md, stack, &rec, stats);
        else hit = world_hit(h->sc, o, d, md, stack, &rec, stats);
        if (hit) {
            Isect is;
            world_hit_geometry(h->sc, o, d, rec, &is);
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
            if (out_geom) for (int k = 0; k < 12; ++k) out_geom[12 * i + k] = 0.0;
            if (out_uvw) { out_uvw[3 * i] = 0; out_uvw[3 * i + 1] = 0; out_uvw[3 * i + 2] = 0; }
        }
    }
    if (counters) {
        counters[0] = stats.branches; counters[1] = stats.leaves; counters[2] = stats.items;
        counters[3] = stats.prim_tests; counters[4] = stats.tri_tests;
    }
    return 0;
}

int hs_contains_batch(uint64_t scene, int64_t n, const double* points, int32_t cap, int32_t* out_count, int32_t* out_prims) {
    HostScene* h = reinterpret_cast<HostScene*>(scene);
    NoStats stats;
    KdStackEntry stack[RSB_KD_STACK];
    for (int64_t i = 0; i < n; ++i) {
        V3 p = v3(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
        int offset, count, found = 0;
        for (int k = 0; k < cap; ++k) out_prims[i * cap + k] = -1;
        if (kd_locate(h->sc.world, p, &offset, &count, stats)) {
            for (int k = 0; k < count; ++k) {
                int id = h->sc.world.items[offset + k];
                if (prim_contains(h->sc, id, p, stack, stats)) {
                    if (found < cap) out_prims[i * cap + found] = id;
                    ++found;
                }
            }
        }
        out_count[i] = found;
    }
    return 0;
}

// the state seed(d) leaves: fast = 0 the reference's init_by_array64 restated (Mt19937_64::seed), 1 the 312-step form
// through the seed-independent table (mt_seed_fast, what k_wf_seed runs)
int hs_mt_state(uint64_t seed, int32_t fast, uint64_t* out) {
    if (fast) {
        uint64_t T[RSB_MT_NN];
        mt_seed_table(T);
        mt_seed_fast(T, seed, out);
    } else {
        Mt19937_64 g;
        g.mt = out;
        g.stride = 1;
        g.seed(seed);
    }
    return 0;
}

int hs_rng_uniform(uint64_t seed, int64_t n, double* out) {
    std::vector<uint64_t> state(RSB_MT_NN);
    Rng rng;
    rng.mode = RNG_MT19937_64;
    rng.mt.mt = state.data();
    rng.mt.stride = 1;
    rng.mt.seed(seed);
    for (int64_t i = 0; i < n; ++i) out[i] = rng.uniform();
    return 0;
}

// serial equivalent of k_render (same per-pixel stream definition, same per-bin arithmetic)
int hs_render(uint64_t scene, const RsbCamera* camera, const RsbRayConfig* config, const RsbSpectral* spectral,
              const RsbRngDesc* rngd, int64_t n_pixels, const int32_t* pixels, double* mean, double* variance,
              uint64_t* ray_count, uint64_t* counters /* optional: branches, leaves, items, prim_tests, tri_tests, paths, segments */,
              const double* xyz_curves /* optional [bins][n_channels] */, double xyz_delta, double* xyz_mean /* [n_pixels][n_channels], per task */,
              double* xyz_variance, int32_t n_channels, const int32_t* channel_mode /* RSB_PROJ_* per channel */) {
    HostScene* h = reinterpret_cast<HostScene*>(scene);
    int nm = (int)h->ps.mat_type.size();
    if (spectral->n_materials != nm) { g_err = "spectral tables do not match the scene's materials"; return RSB_ERR_ARG; }
    std::vector<Material> mats((size_t)nm);
    for (int i = 0; i < nm; ++i) {
        memset(&mats[i], 0, sizeof(Material));
        mats[i].type = h->ps.mat_type[i];
        mats[i].transmission_only = h->ps.mat_transmission_only[i];
        mats[i].table = i;
        mats[i].table2 = spectral->table2 ? spectral->table2[i] : -1;
        mats[i].scale = spectral->scale ? spectral->scale[i] : 1.0;
        mats[i].index_in = spectral->index_in ? spectral->index_in[i] : 1.0;
        mats[i].index_out = spectral->index_out ? spectral->index_out[i] : 1.0;
    }
    Spectral sp;
    sp.mats = mats.data();
    sp.tables = spectral->tables;
    sp.tables_ln = nullptr;
    sp.bins = spectral->bins;
    sp.n_materials = nm;
    sp.n_tables = spectral->n_tables > 0 ? spectral->n_tables : nm;
    RayConfig cfg;
    cfg.bins = config->bins;
    cfg.extinction_min_depth = config->extinction_min_depth;
    cfg.max_depth = config->max_depth;
    cfg.importance_sampling = config->importance_sampling;
    cfg.min_wavelength = config->min_wavelength;
    cfg.max_wavelength = config->max_wavelength;
    cfg.extinction_prob = config->extinction_prob;
    cfg.important_path_weight = config->important_path_weight;
    cfg.max_distance = config->max_distance;
    Camera cam;
    cam.nx = camera->nx; cam.ny = camera->ny; cam.pixel_samples = camera->pixel_samples; cam.kind = camera->kind;
    cam.image_delta = camera->image_delta; cam.image_start_x = camera->image_start_x; cam.image_start_y = camera->image_start_y;
    cam.sensitivity = camera->sensitivity;
    memcpy(cam.to_root, camera->to_root, 12 * sizeof(double));
    cam.to_root[12] = 1.0 / camera->to_root_w;
    cam.pixel_origins = camera->pixel_origins;
    cam.pixel_directions = camera->pixel_directions;

    int cap = 6 * (std::max(cfg.max_depth, cfg.extinction_min_depth) + 2);
    std::vector<LogEntry> logbuf((size_t)cap);
    std::vector<uint64_t> st_path(RSB_MT_NN), st_jit(RSB_MT_NN);
    CountStats stats;
    KdStackEntry stack[RSB_KD_STACK];
    uint64_t paths = 0, segments = 0;
    if (!pixels) n_pixels = (int64_t)cam.nx * cam.ny;
    const int bins = sp.bins, spp = cam.pixel_samples;
    for (int64_t w = 0; w < n_pixels; ++w) {
        int px, py;
        if (pixels) { px = pixels[2 * w]; py = pixels[2 * w + 1]; }
        else { px = (int)(w / cam.ny); py = (int)(w % cam.ny); }
        long long frame_row = (long long)px * cam.ny + py;
        long long pixel_id = (long long)py * cam.nx + px;
        Rng rng, jit;
        rng.mode = rngd->mode; jit.mode = rngd->mode;
        rng.mt.mt = st_path.data(); rng.mt.stride = 1; rng.mt.mti = RSB_MT_NN;
        jit.mt.mt = st_jit.data(); jit.mt.stride = 1; jit.mt.mti = RSB_MT_NN;
        const int pairs = camera_jitter_pairs(cam.kind);
        const bool draws = camera_pixel_draws(cam, px, py);      // (an edge pixel of a VectorCamera draws nothing up front)
        std::vector<double> pre((size_t)2 * pairs * spp, 0.0);
        if (rngd->mode == RNG_MT19937_64) {
            mt_seed_pair(rngd->seed + (uint64_t)pixel_id, draws ? 2 * pairs * spp : 0, st_jit.data(), &jit.mt.mti, st_path.data(), &rng.mt.mti);
            for (size_t k = 0; draws && k < pre.size(); ++k) pre[k] = jit.uniform();      // the task's up-front draws, in draw order
        }
        double* m = mean + frame_row * bins;
        double* v = variance + frame_row * bins;
        for (int s = 0; s < spp; ++s) {
            if (rngd->mode == RNG_PHILOX) rng.px.init(rngd->seed, (uint64_t)pixel_id, (uint32_t)s);
            double u1, u2, u3 = 0.0, u4 = 0.0;
            if (rngd->mode == RNG_MT19937_64) {
                u1 = pre[2 * s]; u2 = pre[2 * s + 1];
                if (pairs == 2) { u3 = pre[2 * spp + 2 * s]; u4 = pre[2 * spp + 2 * s + 1]; }
            } else {
                if (draws) { u1 = rng.uniform(); u2 = rng.uniform(); } else { u1 = u2 = 0.0; }
                if (pairs == 2) { u3 = rng.uniform(); u4 = rng.uniform(); }
            }
            V3 o, d;
            double weight;
            pinhole_ray(cam, px, py, u1, u2, &o, &d, &weight, u3, u4);
            PathLog log;
            log.base = logbuf.data(); log.stride = 1; log.capacity = cap; log.n = 0; log.overflow = 0;
            uint32_t rays = 0;
            int res;
            if (getenv("HS_DEBUG")) {
                PathState ps;
                path_begin(ps, log, o, d);
                do {
                    fprintf(stderr, "seg depth=%d o=(%.17g %.17g %.17g) d=(%.17g %.17g %.17g)\n", ps.depth, ps.o.x, ps.o.y, ps.o.z, ps.d.x, ps.d.y, ps.d.z);
                    res = path_step(h->sc, sp, cfg, ps, rng, stack, log, stats);
                } while (res == PATH_CONTINUE);
                rays = ps.rays;
                fprintf(stderr, "end res=%d rays=%u logn=%d\n", res, rays, log.n);
                for (int k = log.n - 1; k >= 0; --k) { LogEntry e = log.get(k); fprintf(stderr, "  log op=%d table=%d v=%.17g\n", e.op, e.table, e.v); }
            } else {
                res = trace_path(h->sc, sp, cfg, o, d, rng, stack, log, &rays, stats);
            }
            if (log.overflow) { g_err = "path log overflow"; return RSB_ERR_OVERFLOW; }
            *ray_count += rays;
            paths += 1;
            segments += rays;
            double tri[RSB_PROJ_MAX] = {0.0};
            for (int b = 0; b < bins; ++b) {
                double x = 0.0;
                if (res == PATH_EMITTED) x = replay_bin(log, sp, b);
                x = x * weight;
                if (xyz_curves)
                    for (int ch = 0; ch < n_channels; ++ch) {
                        const double cv = xyz_curves[n_channels * b + ch];
                        // the pixel processors' own expressions: colour.pyx:182-184, mono/power.pyx:777, mono/radiance.pyx:193
                        if (channel_mode[ch] == RSB_PROJ_XYZ) tri[ch] += xyz_delta * x * cv;
                        else if (channel_mode[ch] == RSB_PROJ_POWER) tri[ch] += x * cv * cam.sensitivity * xyz_delta;
                        else tri[ch] += x * cv * xyz_delta;
                    }
                x = x * cam.sensitivity;
                welford_add(x, m + b, v + b, s);
            }
            if (xyz_curves)
                for (int ch = 0; ch < n_channels; ++ch)
                    welford_add(channel_mode[ch] == RSB_PROJ_XYZ ? tri[ch] * cam.sensitivity : tri[ch], xyz_mean + n_channels * w + ch,
                                xyz_variance + n_channels * w + ch, s);
        }
    }
    if (counters) {
        counters[0] = stats.branches; counters[1] = stats.leaves; counters[2] = stats.items;
        counters[3] = stats.prim_tests; counters[4] = stats.tri_tests; counters[5] = paths; counters[6] = segments;
    }
    return 0;
}

int hs_div_count(int64_t n, const double* x, const int32_t* d, double* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = div_count(x[i], (double)d[i], 1.0 / (double)d[i]);
    return 0;
}

int hs_div_exact(int64_t n, const double* x, const double* d, double* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = div_exact(x[i], d[i], exact_recip(d[i]));
    return 0;
}

// the kd traversal's plane distance exactly as kd_descend forms it: per-ray reciprocal + unsafe mask + one correction
int hs_div_recip1(int64_t n, const double* x, const double* d, double* out) {
    for (int64_t i = 0; i < n; ++i) {
        V3 dir = v3(d[i], d[i], d[i]);
        V3 rc = ray_reciprocals(dir);
        bool unsafe = recip_unsafe_mask(dir, rc) != 0;
        out[i] = d[i] == 0.0 ? x[i] / d[i] : div_recip1(x[i], d[i], rc.x, unsafe);   // zero components never reach the division
    }
    return 0;
}

int hs_frame_combine(int64_t n_pixels_total, int32_t frame_bins, int32_t slice_offset, int32_t slice_bins, const double* mean,
                     const double* variance, int32_t samples, double* fmean, double* fvar, int32_t* fsamples) {
    for (int64_t p = 0; p < n_pixels_total; ++p)
        for (int b = 0; b < slice_bins; ++b) {
            int64_t src = p * slice_bins + b, dst = p * frame_bins + slice_offset + b;
            double mt, vt;
            int nt;
            double vb = variance[src];
            if (vb < 0) vb = 0;
            stats_combine(fmean[dst], fvar[dst], fsamples[dst], mean[src], vb, samples, &mt, &vt, &nt);
            fmean[dst] = mt; fvar[dst] = vt; fsamples[dst] = nt;
        }
    return 0;
}

}  // extern "C"
