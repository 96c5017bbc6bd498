// raysect_b200.cu -- C ABI (include/raysect_b200.h): context, scene upload, kernel launches.
#include "../../include/raysect_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <exception>
#include <string>
#include <vector>

#include "kdtree_host.h"
#include "rsb_kernels.cuh"
#include "rsb_trav.cuh"
#include "scene_pack.h"

using namespace rsb;

namespace {

thread_local std::string g_error;

int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}

#define RSB_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(RSB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));         \
    } while (0)

struct DeviceScene {
    Scene sc;
    std::vector<void*> allocs;
    int32_t n_world_items = 0;
    int32_t n_materials = 0;
    std::vector<int32_t> mat_type, mat_transmission_only;
    int32_t stage_bytes = 0;   // shared memory needed to stage world tree + prims (0 = do not stage)
    bool has_mesh = false, has_csg = false;
    int32_t n_mesh_prims = 0;  // world-level Mesh primitives (Mesh.hit rounds of the query pipeline)
};

struct Context {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t total_mem = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_in = nullptr, ev_out = nullptr;
    DevCounters* d_counters = nullptr;
    unsigned long long* d_scalars = nullptr;   // [0] work counter, [1] ray count, [2] overflow flag (as int)
    unsigned char* d_slots = nullptr;   // wavefront slot state (rsb_kernels.cuh: WfSlots)
    size_t slot_bytes = 0;
    // rsb_render_slice: the last rendered slice stays on the device (mean | variance, [n_pix_frame][slice_bins]) with its
    // task list, until rsb_slice_update_frame / rsb_slice_read consumed it; grow-only buffers
    double* d_slice = nullptr;
    size_t slice_cap = 0;               // doubles
    int32_t* d_slice_pix = nullptr;
    size_t slice_pix_cap = 0;           // int32 pairs
    double* d_frame = nullptr;          // host frames pass through here: mean | variance, [n_pix_frame][frame_bins]
    int32_t* d_frame_samples = nullptr;
    size_t frame_cap = 0;               // elements
    unsigned long long* d_slice_rays = nullptr;
    struct {
        int32_t nx = 0, ny = 0, bins = 0, samples = 0; int64_t n_pixels = 0; bool listed = false, valid = false;
        bool has_bins = false, has_xyz = false;      // which statistics the render kept (spectral frame / XYZ work items)
        int32_t n_passes = 0, n_slices = 0, pass_samples = 0, n_channels = 0, slice_bins = 0;
    } slice;
    // RGB side of rsb_render_slices_xyz: curves [n_slices][bins][3] | delta [n_slices] | mean [work items][3] | variance [same];
    // the (nx, ny, 3) host frames pass through d_xyz_frame (mean | variance) and d_xyz_samples
    double* d_xyz = nullptr;
    size_t xyz_cap = 0;                 // doubles
    double* d_campix = nullptr;         // VectorCamera: per-pixel origins | directions
    size_t campix_cap = 0;              // doubles
    double* d_xyz_frame = nullptr;
    int32_t* d_xyz_samples = nullptr;
    size_t xyz_frame_cap = 0;           // elements
    unsigned char* d_rq = nullptr;      // query pipeline arrays (rsb_trav.cuh: RqBuf)
    size_t rq_bytes = 0;
    long long rq_chunk = 0;             // RSB_RQ_CHUNK: queries per pipeline pass of rsb_hit_batch / rsb_hit_sweep; 0 = rq_pass_size()
    bool query_reorder = true;          // rsb_set_query_reorder / RSB_RQ_REORDER=0: sort every batch on its coherence key first
    double* d_pass = nullptr;           // frames of passes 1.. of a multi-pass render: [2][n_passes - 1][frame]
    size_t pass_bytes = 0;
    long long chunk_items = 16LL << 20; // RSB_CHUNK_ITEMS: pixel streams seeded up front per chunk (2.5 KB + 16 B per sample each)
    unsigned int* h_idle = nullptr;     // pinned
    RsbRenderStats render_stats{};
    int slots_per_sm = 8192;            // RSB_SLOTS_PER_SM: pixel streams in flight per SM (wavefront width)
    bool use_graphs = true;             // RSB_NO_GRAPH=1 launches the wave kernels one by one (debugging)
    std::vector<cudaEvent_t> event_pool;
    unsigned long long* d_mt_table = nullptr;   // seed-independent part of MT19937-64's seed(d) (rsb_rng.h mt_seed_table)
    Material* d_mats = nullptr;
    double* d_tables = nullptr;
    size_t mats_cap = 0, tables_cap = 0;
    float last_ms = 0.f;
    RsbCounters last_counters{};
    std::vector<DeviceScene*> scenes;
};

const size_t kMaxStageBytes = 96 * 1024;
const long long kReorderMin = 4096;      // smaller batches are traversed as they come

// Kernels are instantiated per scene feature set (rsb_geom.h RSB_FEAT_*): analytic primitives only, + meshes, or
// everything (meshes and CSG), and world-level data staged in shared memory vs read from HBM/L2.  M(COUNT, FEAT) is expanded
// with compile-time constants.
#define RSB_DISPATCH_FEAT(count, feat, staged, M)                                                      \
    do {                                                                                               \
        const int feat_ = (feat) | ((staged) ? RSB_FEAT_STAGED : 0);                                   \
        if (count) {                                                                                   \
            switch (feat_) {                                                                           \
                case 0: M(true, 0); break;                                                             \
                case 1: M(true, 1); break;                                                             \
                case 2: M(true, 2); break;                                                             \
                case 3: M(true, 3); break;                                                             \
                case 4: M(true, 4); break;                                                             \
                case 5: M(true, 5); break;                                                             \
                case 6: M(true, 6); break;                                                             \
                default: M(true, 7); break;                                                            \
            }                                                                                          \
        } else {                                                                                       \
            switch (feat_) {                                                                           \
                case 0: M(false, 0); break;                                                            \
                case 1: M(false, 1); break;                                                            \
                case 2: M(false, 2); break;                                                            \
                case 3: M(false, 3); break;                                                            \
                case 4: M(false, 4); break;                                                            \
                case 5: M(false, 5); break;                                                            \
                case 6: M(false, 6); break;                                                            \
                default: M(false, 7); break;                                                           \
            }                                                                                          \
        }                                                                                              \
    } while (0)

// the query pipeline's kernels exist for the feature sets that can reach them: k_rq_walk for scenes with meshes,
// k_rq_world for scenes without
#define RSB_DISPATCH_FEAT_MESH(count, feat, staged, M)                                                 \
    do {                                                                                               \
        const int feat_ = (feat) | ((staged) ? RSB_FEAT_STAGED : 0);                                   \
        if (count) {                                                                                   \
            if (feat_ == RSB_FEAT_MESH) M(true, RSB_FEAT_MESH);                                        \
            else if (feat_ == (RSB_FEAT_MESH | RSB_FEAT_STAGED)) M(true, (RSB_FEAT_MESH | RSB_FEAT_STAGED)); \
            else if (feat_ == RSB_FEAT_ALL) M(true, RSB_FEAT_ALL);                                     \
            else M(true, (RSB_FEAT_ALL | RSB_FEAT_STAGED));                                            \
        } else {                                                                                       \
            if (feat_ == RSB_FEAT_MESH) M(false, RSB_FEAT_MESH);                                       \
            else if (feat_ == (RSB_FEAT_MESH | RSB_FEAT_STAGED)) M(false, (RSB_FEAT_MESH | RSB_FEAT_STAGED)); \
            else if (feat_ == RSB_FEAT_ALL) M(false, RSB_FEAT_ALL);                                    \
            else M(false, (RSB_FEAT_ALL | RSB_FEAT_STAGED));                                           \
        }                                                                                              \
    } while (0)
#define RSB_DISPATCH_FEAT_NOMESH(count, feat, staged, M)                                               \
    do {                                                                                               \
        const int feat_ = ((feat) & RSB_FEAT_CSG) | ((staged) ? RSB_FEAT_STAGED : 0);                  \
        if (count) {                                                                                   \
            if (feat_ == 0) M(true, 0);                                                                \
            else if (feat_ == RSB_FEAT_STAGED) M(true, RSB_FEAT_STAGED);                               \
            else if (feat_ == RSB_FEAT_CSG) M(true, RSB_FEAT_CSG);                                     \
            else M(true, (RSB_FEAT_CSG | RSB_FEAT_STAGED));                                            \
        } else {                                                                                       \
            if (feat_ == 0) M(false, 0);                                                               \
            else if (feat_ == RSB_FEAT_STAGED) M(false, RSB_FEAT_STAGED);                              \
            else if (feat_ == RSB_FEAT_CSG) M(false, RSB_FEAT_CSG);                                    \
            else M(false, (RSB_FEAT_CSG | RSB_FEAT_STAGED));                                           \
        }                                                                                              \
    } while (0)

template <class T>
int upload(DeviceScene* ds, const T* host, size_t count, const T** dev, size_t pad_to_bytes = 16) {
    size_t bytes = count * sizeof(T);
    size_t alloc = ((bytes + pad_to_bytes - 1) / pad_to_bytes) * pad_to_bytes;
    if (alloc == 0) alloc = pad_to_bytes;
    void* p = nullptr;
    RSB_CUDA(cudaMalloc(&p, alloc));
    ds->allocs.push_back(p);
    RSB_CUDA(cudaMemset(p, 0, alloc));
    if (bytes) RSB_CUDA(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice));
    *dev = reinterpret_cast<const T*>(p);
    return RSB_OK;
}

int upload_tree(DeviceScene* ds, const HostKdTree& h, KdTree* out) {
    int rc = upload(ds, h.nodes.data(), h.nodes.size(), &out->nodes);
    if (rc) return rc;
    rc = upload(ds, h.items.data(), h.items.size(), &out->items);
    if (rc) return rc;
    memcpy(out->bounds, h.bounds, sizeof(h.bounds));
    out->n_nodes = (int32_t)h.nodes.size();
    out->max_depth = h.depth;
    return RSB_OK;
}

Context* as_ctx(uint64_t h) { return reinterpret_cast<Context*>(h); }
DeviceScene* as_scene(uint64_t h) { return reinterpret_cast<DeviceScene*>(h); }

void free_scene(DeviceScene* ds) {
    for (void* p : ds->allocs) cudaFree(p);
    delete ds;
}

int grid_for(Context* c, long long n, int threads, int blocks_per_sm) {
    long long need = (n + threads - 1) / threads;
    long long cap = (long long)c->sm_count * blocks_per_sm;
    return (int)std::max(1LL, std::min(need, cap));
}

int read_counters(Context* c, cudaStream_t st) {
    DevCounters h;
    RSB_CUDA(cudaMemcpyAsync(&h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost, st));
    RSB_CUDA(cudaStreamSynchronize(st));
    c->last_counters.rays = h.rays;
    c->last_counters.branches = h.branches;
    c->last_counters.leaves = h.leaves;
    c->last_counters.items = h.items;
    c->last_counters.prim_tests = h.prim_tests;
    c->last_counters.tri_tests = h.tri_tests;
    c->last_counters.paths = h.paths;
    c->last_counters.contains = h.contains;
    c->last_counters.table_reads = h.table_reads;
    c->last_counters.contains_nodes = h.contains_nodes;
    c->last_counters.contains_items = h.contains_items;
    c->last_counters.contains_prim_tests = h.contains_prim_tests;
    return RSB_OK;
}

}  // namespace

extern "C" {

const char* rsb_last_error(void) { return g_error.c_str(); }
int rsb_version(void) { return 100; }
void rsb_free(void* p) { free(p); }

int rsb_kdtree_build(const double* boxes, int64_t n_items, int32_t max_depth, int32_t min_items, double hit_cost,
                     double empty_bonus, uint8_t** stream, int64_t* stream_bytes) {
    // n_items == 0 is legal: the reference builds a single empty leaf inside inverted (+inf, -inf) bounds
    if ((!boxes && n_items > 0) || n_items < 0 || !stream || !stream_bytes) return fail(RSB_ERR_ARG, "rsb_kdtree_build: bad arguments");
    if (empty_bonus < 0.0 || empty_bonus > 1.0)
        return fail(RSB_ERR_ARG, "The empty_bonus cost modifier must lie in the range [0.0, 1.0].");
    // nothing may unwind through the C ABI: allocation or thread-creation failures become an error code
    try {
        HostKdTree t;
        kd_build(boxes, n_items, max_depth, min_items, hit_cost, empty_bonus, &t);
        std::vector<uint8_t> bytes;
        kd_write_stream(t, &bytes);
        uint8_t* out = (uint8_t*)malloc(bytes.size());
        if (!out) return fail(RSB_ERR_ARG, "rsb_kdtree_build: out of memory");
        memcpy(out, bytes.data(), bytes.size());
        *stream = out;
        *stream_bytes = (int64_t)bytes.size();
        return RSB_OK;
    } catch (const std::exception& e) {
        return fail(RSB_ERR_ARG, std::string("rsb_kdtree_build: ") + e.what());
    }
}

int rsb_mesh_face_normals(const float* vertices, int32_t n_vertices, const int32_t* triangles, int32_t n_triangles,
                          int32_t tri_stride, float* face_normals) {
    if (!vertices || !triangles || !face_normals || (tri_stride != 3 && tri_stride != 6))
        return fail(RSB_ERR_ARG, "rsb_mesh_face_normals: bad arguments");
    for (int32_t i = 0; i < n_triangles; ++i) {
        const int32_t* row = triangles + (size_t)i * tri_stride;
        for (int k = 0; k < 3; ++k)
            if (row[k] < 0 || row[k] >= n_vertices) return fail(RSB_ERR_ARG, "The triangle array references non-existent vertices.");
        mesh_face_normal(vertices, row, face_normals + 3 * (size_t)i);
    }
    return RSB_OK;
}

int rsb_mesh_triangle_boxes(const float* vertices, int32_t n_vertices, const int32_t* triangles, int32_t n_triangles,
                            int32_t tri_stride, double* boxes) {
    if (!vertices || !triangles || !boxes || (tri_stride != 3 && tri_stride != 6))
        return fail(RSB_ERR_ARG, "rsb_mesh_triangle_boxes: bad arguments");
    const double BOX_PADDING = 1e-6;
    for (int32_t i = 0; i < n_triangles; ++i) {
        const int32_t* row = triangles + (size_t)i * tri_stride;
        for (int k = 0; k < 3; ++k)
            if (row[k] < 0 || row[k] >= n_vertices) return fail(RSB_ERR_ARG, "The triangle array references non-existent vertices.");
        const float* a = vertices + 3 * (size_t)row[0];
        const float* b = vertices + 3 * (size_t)row[1];
        const float* c = vertices + 3 * (size_t)row[2];
        double* o = boxes + 6 * (size_t)i;
        for (int k = 0; k < 3; ++k) {
            // min()/max() over float32 values, then stored as double (mesh.pyx:484-497)
            float lo = std::min(std::min(a[k], b[k]), c[k]);
            float hi = std::max(std::max(a[k], b[k]), c[k]);
            o[k] = (double)lo;
            o[3 + k] = (double)hi;
        }
        // bbox.pad(max(BOX_PADDING, bbox.largest_extent() * BOX_PADDING)) (mesh.pyx:502)
        double ex = std::max(0.0, o[3] - o[0]), ey = std::max(0.0, o[4] - o[1]), ez = std::max(0.0, o[5] - o[2]);
        double pad = std::max(BOX_PADDING, std::max(std::max(ex, ey), ez) * BOX_PADDING);
        for (int k = 0; k < 3; ++k) {
            o[k] = o[k] - pad;
            o[3 + k] = o[3 + k] + pad;
        }
    }
    return RSB_OK;
}

int rsb_context_create(int device, uint64_t* ctx) {
    if (!ctx) return fail(RSB_ERR_ARG, "rsb_context_create: null handle pointer");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(RSB_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                      "); libraysect_b200 has no CPU fallback");
    if (device < 0 || device >= count) return fail(RSB_ERR_ARG, "rsb_context_create: device index out of range");
    RSB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RSB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(RSB_ERR_CUDA, "libraysect_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major) +
                                      std::to_string(prop.minor));
    Context* c = new Context();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->total_mem = prop.totalGlobalMem;
    RSB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    RSB_CUDA(cudaEventCreate(&c->ev0));
    RSB_CUDA(cudaEventCreate(&c->ev1));
    RSB_CUDA(cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming));
    RSB_CUDA(cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming));
    RSB_CUDA(cudaMalloc(&c->d_counters, sizeof(DevCounters)));
    RSB_CUDA(cudaMemset(c->d_counters, 0, sizeof(DevCounters)));
    RSB_CUDA(cudaMalloc(&c->d_scalars, 8 * sizeof(unsigned long long)));
    RSB_CUDA(cudaMemset(c->d_scalars, 0, 8 * sizeof(unsigned long long)));
    RSB_CUDA(cudaMallocHost(&c->h_idle, sizeof(unsigned int)));
    {
        uint64_t table[RSB_MT_NN];
        mt_seed_table(table);
        RSB_CUDA(cudaMalloc(&c->d_mt_table, sizeof(table)));
        RSB_CUDA(cudaMemcpy(c->d_mt_table, table, sizeof(table), cudaMemcpyHostToDevice));
    }
    if (const char* ng = getenv("RSB_NO_GRAPH")) c->use_graphs = !(ng[0] == '1');
    if (const char* sp = getenv("RSB_CHUNK_ITEMS")) { long long v = atoll(sp); if (v >= 1024 && v <= (64LL << 20)) c->chunk_items = v; }
    if (const char* sp = getenv("RSB_RQ_REORDER")) c->query_reorder = atoi(sp) != 0;
    if (const char* sp = getenv("RSB_RQ_CHUNK")) { long long v = atoll(sp); if (v >= 1024 && v <= (64LL << 20)) c->rq_chunk = v; }
    if (const char* sp = getenv("RSB_SLOTS_PER_SM")) { int v = atoi(sp); if (v >= 32 && v <= 65536) c->slots_per_sm = v; }
    *ctx = reinterpret_cast<uint64_t>(c);
    return RSB_OK;
}

int rsb_context_destroy(uint64_t ctx) {
    Context* c = as_ctx(ctx);
    if (!c) return fail(RSB_ERR_ARG, "null context");
    cudaSetDevice(c->device);
    for (DeviceScene* s : c->scenes) free_scene(s);
    cudaFree(c->d_counters);
    cudaFree(c->d_scalars);
    cudaFree(c->d_slots);
    cudaFree(c->d_pass);
    cudaFree(c->d_rq);
    cudaFree(c->d_slice);
    cudaFree(c->d_xyz);
    cudaFree(c->d_campix);
    cudaFree(c->d_xyz_frame);
    cudaFree(c->d_xyz_samples);
    cudaFree(c->d_slice_pix);
    cudaFree(c->d_frame);
    cudaFree(c->d_frame_samples);
    cudaFree(c->d_slice_rays);
    cudaFreeHost(c->h_idle);
    for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
    cudaFree(c->d_mats);
    cudaFree(c->d_mt_table);
    cudaFree(c->d_tables);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaEventDestroy(c->ev_in);
    cudaEventDestroy(c->ev_out);
    cudaStreamDestroy(c->stream);
    delete c;
    return RSB_OK;
}

int rsb_device_info(uint64_t ctx, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, uint64_t* total_mem) {
    Context* c = as_ctx(ctx);
    if (!c) return fail(RSB_ERR_ARG, "null context");
    if (sm_count) *sm_count = c->sm_count;
    if (cc_major) *cc_major = c->cc_major;
    if (cc_minor) *cc_minor = c->cc_minor;
    if (total_mem) *total_mem = c->total_mem;
    return RSB_OK;
}

int rsb_scene_create(uint64_t ctx, const RsbSceneDesc* d, uint64_t* scene) {
    Context* c = as_ctx(ctx);
    if (!c || !d || !scene) return fail(RSB_ERR_ARG, "rsb_scene_create: null argument");
    RSB_CUDA(cudaSetDevice(c->device));
    PackedScene ps;
    std::string err;
    int prc;
    try {
        prc = pack_scene(d, &ps, &err);
    } catch (const std::exception& e) {
        return fail(RSB_ERR_ARG, std::string("rsb_scene_create: ") + e.what());
    }
    if (prc) return fail(prc, err);

    DeviceScene* ds = new DeviceScene();
    auto bail = [&](int rc) { free_scene(ds); return rc; };
    memset(&ds->sc, 0, sizeof(ds->sc));
    int rc = upload(ds, ps.prims.data(), ps.prims.size(), &ds->sc.prims);
    if (rc) return bail(rc);
    ds->sc.n_prims = (int32_t)ps.prims.size();
    for (const Prim& pr : ps.prims) {
        if (pr.type == PRIM_MESH) ds->has_mesh = true;
        // (has_csg = "needs the full-featured kernel instantiations": CSG trees, and the torus with its quartic solver)
        if (pr.type >= PRIM_UNION || pr.type == PRIM_TORUS) ds->has_csg = true;
    }
    for (int32_t i = 0; i < ps.n_world; ++i)
        if (ps.prims[i].type == PRIM_MESH) ds->n_mesh_prims += 1;
    ds->sc.n_world = ps.n_world;
    rc = upload_tree(ds, ps.world, &ds->sc.world);
    if (rc) return bail(rc);
    ds->n_world_items = (int32_t)ps.world.items.size();
    {
        // leaf-ordered item rows with their AABBs (worlds that are not staged in shared memory read these)
        std::vector<LeafRow> rows(ps.world.items.size());
        for (size_t k = 0; k < rows.size(); ++k) {
            const Prim& pr = ps.prims[ps.world.items[k]];
            memcpy(rows[k].bbox, pr.bbox, sizeof(pr.bbox));
            rows[k].id = ps.world.items[k];
            rows[k].type = pr.type;
            rows[k].pad[0] = rows[k].pad[1] = 0;
        }
        rc = upload(ds, rows.data(), rows.size(), &ds->sc.world_rows, 64);
        if (rc) return bail(rc);
    }
    {
        StageLayout l = stage_layout((int)ps.world.nodes.size(), (int)ps.world.items.size(), (int)ps.prims.size());
        ds->stage_bytes = (size_t)l.total <= kMaxStageBytes ? l.total : 0;
    }
    std::vector<Mesh> meshes(std::max<size_t>(1, ps.meshes.size()));
    memset(meshes.data(), 0, meshes.size() * sizeof(Mesh));
    for (size_t mi = 0; mi < ps.meshes.size(); ++mi) {
        const PackedMesh& pm = ps.meshes[mi];
        Mesh& m = meshes[mi];
        rc = upload(ds, pm.tri.data(), pm.tri.size(), &m.tri);
        if (rc) return bail(rc);
        rc = upload(ds, pm.tri_idx.data(), pm.tri_idx.size(), &m.tri_idx);
        if (rc) return bail(rc);
        m.vnormals = nullptr;
        if (!pm.vnormals.empty()) {
            rc = upload(ds, pm.vnormals.data(), pm.vnormals.size(), &m.vnormals);
            if (rc) return bail(rc);
        }
        rc = upload_tree(ds, pm.tree, &m.tree);
        if (rc) return bail(rc);
        m.n_tri = pm.n_tri;
        m.idx_stride = pm.idx_stride;
        m.smoothing = pm.smoothing;
        m.closed = pm.closed;
    }
    rc = upload(ds, meshes.data(), meshes.size(), &ds->sc.meshes);
    if (rc) return bail(rc);
    ds->sc.n_meshes = (int32_t)ps.meshes.size();
    ds->n_materials = (int32_t)ps.mat_type.size();
    ds->mat_type = ps.mat_type;
    ds->mat_transmission_only = ps.mat_transmission_only;
    ds->sc.n_important = (int32_t)ps.imp_weight.size();
    ds->sc.imp_total = ps.imp_total;
    if (!ps.imp_weight.empty()) {
        rc = upload(ds, ps.imp_sphere.data(), ps.imp_sphere.size(), &ds->sc.imp_sphere);
        if (rc) return bail(rc);
        rc = upload(ds, ps.imp_weight.data(), ps.imp_weight.size(), &ds->sc.imp_weight);
        if (rc) return bail(rc);
        rc = upload(ds, ps.imp_cdf.data(), ps.imp_cdf.size(), &ds->sc.imp_cdf);
        if (rc) return bail(rc);
    }
    c->scenes.push_back(ds);
    *scene = reinterpret_cast<uint64_t>(ds);
    return RSB_OK;
}

int rsb_scene_destroy(uint64_t ctx, uint64_t scene) {
    Context* c = as_ctx(ctx);
    DeviceScene* ds = as_scene(scene);
    if (!c || !ds) return fail(RSB_ERR_ARG, "null handle");
    auto it = std::find(c->scenes.begin(), c->scenes.end(), ds);
    if (it == c->scenes.end()) return fail(RSB_ERR_ARG, "scene does not belong to this context");
    c->scenes.erase(it);
    cudaSetDevice(c->device);
    free_scene(ds);
    return RSB_OK;
}

}  // extern "C"

namespace {

// carve the arrays of the query pipeline (rsb_trav.cuh) out of one allocation
struct RqHost {
    RqBuf b;
    double* ray;     // [6][cap]
    double* md;      // [cap]
    // reordering of incoherent batches (rsb_trav.cuh, rq_reorder): permutation and sort scratch
    int32_t* perm = nullptr;
    unsigned int* key = nullptr;
    unsigned int* hist = nullptr;         // [RQ_KEY_BINS]
    unsigned int* block_sums = nullptr;   // [1024]
    unsigned int* bounds = nullptr;       // [10]
};

size_t rq_carve(unsigned char* base, size_t cap, bool park, RqHost* out, bool reorder = false) {
    size_t off = 0;
    auto take = [&](size_t bytes) -> unsigned char* {
        off = (off + 255) & ~(size_t)255;
        unsigned char* p = base ? base + off : nullptr;
        off += bytes;
        return p;
    };
    memset(&out->b, 0, sizeof(out->b));
    out->b.ctr = reinterpret_cast<unsigned int*>(take(16 * sizeof(unsigned int)));
    out->ray = reinterpret_cast<double*>(take(6 * cap * sizeof(double)));
    out->md = reinterpret_cast<double*>(take(cap * sizeof(double)));
    out->b.hit_t = reinterpret_cast<double*>(take(cap * sizeof(double)));
    out->b.hit_a = reinterpret_cast<int4*>(take(cap * sizeof(int4)));
    out->b.hit_uvw = reinterpret_cast<float4*>(take(cap * sizeof(float4)));
    out->b.hit_node = reinterpret_cast<int32_t*>(take(cap * sizeof(int32_t)));
    if (park) {
        out->b.susp = reinterpret_cast<RqSusp*>(take(cap * sizeof(RqSusp)));
        out->b.susp_stack = reinterpret_cast<KdStackEntry*>(take(cap * RQ_WORLD_STACK * sizeof(KdStackEntry)));
        out->b.queue = reinterpret_cast<int2*>(take((size_t)(RQ_MAX_ROUNDS + 1) * cap * sizeof(int2)));
    }
    if (reorder) {
        out->perm = reinterpret_cast<int32_t*>(take(cap * sizeof(int32_t)));
        out->key = reinterpret_cast<unsigned int*>(take(cap * sizeof(unsigned int)));
        out->hist = reinterpret_cast<unsigned int*>(take((size_t)RQ_KEY_BINS * sizeof(unsigned int)));
        out->block_sums = reinterpret_cast<unsigned int*>(take(1024 * sizeof(unsigned int)));
        out->bounds = reinterpret_cast<unsigned int*>(take(16 * sizeof(unsigned int)));
    }
    out->b.ray = out->ray;
    out->b.ray_stride = (long long)cap;
    out->b.md = out->md;
    out->b.md_all = RSB_INF;
    out->b.cap = (long long)cap;
    return (off + 255) & ~(size_t)255;
}

int rq_reserve(Context* c, long long cap, bool park, RqHost* out, bool reorder = false) {
    size_t need = rq_carve(nullptr, (size_t)cap, park, out, reorder);
    if (c->rq_bytes < need) {
        cudaFree(c->d_rq);
        c->d_rq = nullptr; c->rq_bytes = 0;
        RSB_CUDA(cudaMalloc(&c->d_rq, need));
        c->rq_bytes = need;
    }
    rq_carve(c->d_rq, (size_t)cap, park, out, reorder);
    return RSB_OK;
}

// Queries per pipeline pass.  A pass is sorted as a whole (rq_reorder), so the larger it is the smaller the part of the scene
// a stretch of consecutive warps works in: 1e8 independent random rays over 10,000 spheres run at 1,800 / 1,987 / 2,193 Mrays/s
// with passes of 4 Mi / 16 Mi / 64 Mi queries (47 / 52 / 57 % of the HBM roofline).  Mesh-free scenes: 64 Mi (108 B of pipeline
// state per query); scenes with meshes park a walk per query (~800 B): 16 Mi (1.3 M triangles: 922 / 990 / 1,000 Mrays/s).  Never
// more than half of the device memory that is free right now, or what the buffers already hold.
long long rq_pass_size(Context* c, DeviceScene* ds, long long n) {
    if (c->rq_chunk > 0) return std::min<long long>(n, c->rq_chunk);
    long long pass = std::min<long long>(n, ds->has_mesh ? (16LL << 20) : (64LL << 20));
    RqHost probe;
    // (buffers that already hold this pass need no look at the device's free memory: cudaMemGetInfo is a driver round trip
    // that would otherwise sit inside every call)
    if (rq_carve(nullptr, (size_t)pass, ds->has_mesh, &probe, c->query_reorder) <= c->rq_bytes) return pass;
    const size_t per_query = rq_carve(nullptr, (size_t)1 << 20, ds->has_mesh, &probe, c->query_reorder) >> 20;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return std::min<long long>(pass, 4LL << 20); }
    const size_t budget = std::max(c->rq_bytes, free_b / 2);
    const long long fit = (long long)(budget / std::max<size_t>(per_query, 1));
    return std::max<long long>(std::min(pass, fit), std::min<long long>(n, 1LL << 16));
}

// Slot order of the m queries of `src` (the caller's arrays, or the sweep's generator) by coherence key (rsb_trav.cuh):
// rq.perm[slot] = query.  The input stage fills the pipeline in that order; the output stage writes answers back through it.
template <class Source>
int rq_reorder(Context* c, RqHost& rq, long long m, const Source& src, cudaStream_t st) {
    RSB_CUDA(cudaMemsetAsync(rq.bounds, 0xFF, 5 * sizeof(unsigned int), st));
    RSB_CUDA(cudaMemsetAsync(rq.bounds + 5, 0, 5 * sizeof(unsigned int), st));
    RSB_CUDA(cudaMemsetAsync(rq.hist, 0, (size_t)RQ_KEY_BINS * sizeof(unsigned int), st));
    const int grid = grid_for(c, m, 256, 8);
    k_ro_bounds<Source><<<grid, 256, 0, st>>>(m, src, rq.bounds);
    k_ro_keys<Source><<<grid, 256, 0, st>>>(m, src, rq.bounds, rq.key, rq.hist);
    k_ro_scan_bins<<<RQ_KEY_BINS / 1024, 1024, 0, st>>>(rq.hist, rq.block_sums);
    k_ro_scan_blocks<<<1, 1024, 0, st>>>(rq.block_sums);
    k_ro_scatter<<<grid, 256, 0, st>>>(m, rq.key, rq.hist, rq.block_sums, rq.perm);
    RSB_CUDA(cudaGetLastError());
    rq.b.perm = rq.perm;
    return RSB_OK;
}

// World.hit for queries [0, n) of `b` (rays in b.ray / b.md, answers in b.hit_*): one kernel for mesh-free scenes,
// world walk / Mesh.hit / resume rounds for scenes with meshes.
int rq_trace(Context* c, DeviceScene* ds, const RqBuf& b, long long n, cudaStream_t st, bool count) {
    RSB_CUDA(cudaMemsetAsync(b.ctr, 0, 16 * sizeof(unsigned int), st));
    const int feat = ds->has_csg ? RSB_FEAT_ALL : (ds->has_mesh ? RSB_FEAT_MESH : 0);
    const int staged = ds->stage_bytes ? 1 : 0;
    RqArrayClient cl;
    cl.b = b;
    if (!ds->has_mesh) {
        const size_t smem = ds->stage_bytes + RQ_WORLD_SMEM;
        const int grid = grid_for(c, n, RQ_THREADS, (feat & RSB_FEAT_CSG) ? 3 : RQ_WORLD_BLOCKS);
#define RSB_LAUNCH_RQ_WORLD(C, F)                                                                                         \
    do {                                                                                                                  \
        if (smem > 48 * 1024) RSB_CUDA(cudaFuncSetAttribute(k_rq_world<C, F, RqArrayClient>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_rq_world<C, F, RqArrayClient><<<grid, RQ_THREADS, smem, st>>>(ds->sc, ds->n_world_items, cl, b, n, c->d_counters); \
    } while (0)
        RSB_DISPATCH_FEAT_NOMESH(count, feat, staged, RSB_LAUNCH_RQ_WORLD);
#undef RSB_LAUNCH_RQ_WORLD
        RSB_CUDA(cudaGetLastError());
        return RSB_OK;
    }
    const size_t smem_walk = ds->stage_bytes + ax_bytes(feat);
    const int grid_walk = grid_for(c, n, RQ_THREADS, 16);
    const int grid_mesh = c->sm_count * RQ_MESH_BLOCKS;
    const int rounds = std::min<int>(ds->n_mesh_prims, RQ_MAX_ROUNDS);
    RSB_CUDA(cudaFuncSetAttribute(k_rq_mesh<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RQ_MESH_SMEM));
    RSB_CUDA(cudaFuncSetAttribute(k_rq_mesh<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RQ_MESH_SMEM));
#define RSB_LAUNCH_RQ_WALK(C, F, RES, LAST, ROUND)                                                                        \
    do {                                                                                                                  \
        if (smem_walk > 48 * 1024)                                                                                        \
            RSB_CUDA(cudaFuncSetAttribute(k_rq_walk<C, F, RqArrayClient, RES, LAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_walk)); \
        k_rq_walk<C, F, RqArrayClient, RES, LAST><<<grid_walk, RQ_THREADS, smem_walk, st>>>(ds->sc, ds->n_world_items, cl, b, n, ROUND, c->d_counters); \
    } while (0)
#define RSB_LAUNCH_RQ_BEGIN(C, F) RSB_LAUNCH_RQ_WALK(C, F, false, false, 0)
#define RSB_LAUNCH_RQ_RESUME(C, F) RSB_LAUNCH_RQ_WALK(C, F, true, false, r)
#define RSB_LAUNCH_RQ_LAST(C, F) RSB_LAUNCH_RQ_WALK(C, F, true, true, r)
    RSB_DISPATCH_FEAT_MESH(count, feat, staged, RSB_LAUNCH_RQ_BEGIN);
    for (int r = 0; r < rounds; ++r) {
        if (count) k_rq_mesh<true><<<grid_mesh, RQ_THREADS, RQ_MESH_SMEM, st>>>(ds->sc, b, r, c->d_counters);
        else k_rq_mesh<false><<<grid_mesh, RQ_THREADS, RQ_MESH_SMEM, st>>>(ds->sc, b, r, c->d_counters);
        // one mesh primitive: its answer is the ray's memo from now on, no walk can stop a second time
        if (r == rounds - 1 && ds->n_mesh_prims > 1) RSB_DISPATCH_FEAT_MESH(count, feat, staged, RSB_LAUNCH_RQ_LAST);
        else RSB_DISPATCH_FEAT_MESH(count, feat, staged, RSB_LAUNCH_RQ_RESUME);
    }
#undef RSB_LAUNCH_RQ_BEGIN
#undef RSB_LAUNCH_RQ_RESUME
#undef RSB_LAUNCH_RQ_LAST
#undef RSB_LAUNCH_RQ_WALK
    RSB_CUDA(cudaGetLastError());
    return RSB_OK;
}

}  // namespace

extern "C" {

int rsb_set_query_reorder(uint64_t ctx, int32_t on) {
    Context* c = as_ctx(ctx);
    if (!c) return fail(RSB_ERR_ARG, "null context");
    c->query_reorder = on != 0;
    return RSB_OK;
}

int rsb_hit_batch_dev(uint64_t ctx, uint64_t scene, void* cuda_stream, int64_t n, const double* origins,
                      const double* directions, const double* max_distance, int32_t* out_prim, double* out_t,
                      int32_t* out_sub, uint8_t* out_flags, int32_t* out_node, double* out_geom, float* out_uvw,
                      int32_t count) {
    Context* c = as_ctx(ctx);
    DeviceScene* ds = as_scene(scene);
    if (!c || !ds) return fail(RSB_ERR_ARG, "null handle");
    if (n <= 0) return RSB_OK;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    RSB_CUDA(cudaSetDevice(c->device));
    if (count) RSB_CUDA(cudaMemsetAsync(c->d_counters, 0, sizeof(DevCounters), st));
    const long long chunk = rq_pass_size(c, ds, n);
    RqHost rq;
    int rc = rq_reserve(c, chunk, ds->has_mesh, &rq, c->query_reorder);
    if (rc) return rc;
    const int feat = ds->has_csg ? RSB_FEAT_ALL : (ds->has_mesh ? RSB_FEAT_MESH : 0);
    for (long long off = 0; off < n; off += chunk) {
        const long long m = std::min<long long>(chunk, n - off);
        const int grid = grid_for(c, m, 256, 8);
        rq.b.perm = nullptr;
        if (c->query_reorder && m >= kReorderMin) {
            rc = rq_reorder(c, rq, m, BatchSource{origins + 3 * off, directions + 3 * off}, st);
            if (rc) return rc;
        }
        k_rq_batch_in<<<grid, 256, 0, st>>>(m, origins + 3 * off, directions + 3 * off, max_distance ? max_distance + off : nullptr, rq.b.perm,
                                            rq.ray, rq.b.ray_stride, rq.md);
        rc = rq_trace(c, ds, rq.b, m, st, count != 0);
        if (rc) return rc;
        const int ogrid = grid_for(c, m, 128, 16);
#define RSB_LAUNCH_OUT(F)                                                                                                     \
    k_rq_batch_out<F><<<ogrid, 128, 0, st>>>(ds->sc, rq.b, m, out_prim + off, out_t + off, out_sub + off, out_flags + off,         \
                                             out_node ? out_node + 2 * off : nullptr, out_geom ? out_geom + 12 * off : nullptr,   \
                                             out_uvw ? out_uvw + 3 * off : nullptr)
        if (feat == 0) RSB_LAUNCH_OUT(0);
        else if (feat == RSB_FEAT_MESH) RSB_LAUNCH_OUT(RSB_FEAT_MESH);
        else RSB_LAUNCH_OUT(RSB_FEAT_ALL);
#undef RSB_LAUNCH_OUT
    }
    RSB_CUDA(cudaGetLastError());
    if (count) {
        rc = read_counters(c, st);
        if (rc) return rc;
        c->last_counters.rays = (uint64_t)n;
    }
    return RSB_OK;
}

int rsb_hit_batch(uint64_t ctx, uint64_t scene, int64_t n, const double* origins, const double* directions,
                  const double* max_distance, int32_t* out_prim, double* out_t, int32_t* out_sub, uint8_t* out_flags,
                  int32_t* out_node, double* out_geom, float* out_uvw) {
    Context* c = as_ctx(ctx);
    if (!c || !as_scene(scene)) return fail(RSB_ERR_ARG, "null handle");
    if (n < 0 || !origins || !directions || !out_prim || !out_t || !out_sub || !out_flags)
        return fail(RSB_ERR_ARG, "rsb_hit_batch: null array");
    if (n == 0) return RSB_OK;
    RSB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    double *d_o = nullptr, *d_d = nullptr, *d_m = nullptr, *d_t = nullptr, *d_g = nullptr;
    int32_t *d_p = nullptr, *d_s = nullptr, *d_n = nullptr;
    uint8_t* d_f = nullptr;
    float* d_u = nullptr;
    int rc = RSB_OK;
    auto cleanup = [&]() {
        cudaFree(d_o); cudaFree(d_d); cudaFree(d_m); cudaFree(d_t); cudaFree(d_g);
        cudaFree(d_p); cudaFree(d_s); cudaFree(d_n); cudaFree(d_f); cudaFree(d_u);
    };
#define RSB_TRY(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) { cleanup(); return fail(RSB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } \
    } while (0)
    RSB_TRY(cudaMalloc(&d_o, n * 24));
    RSB_TRY(cudaMalloc(&d_d, n * 24));
    if (max_distance) RSB_TRY(cudaMalloc(&d_m, n * 8));
    RSB_TRY(cudaMalloc(&d_t, n * 8));
    RSB_TRY(cudaMalloc(&d_p, n * 4));
    RSB_TRY(cudaMalloc(&d_s, n * 4));
    RSB_TRY(cudaMalloc(&d_f, n));
    if (out_node) RSB_TRY(cudaMalloc(&d_n, n * 8));
    if (out_geom) RSB_TRY(cudaMalloc(&d_g, n * 96));
    if (out_uvw) RSB_TRY(cudaMalloc(&d_u, n * 12));
    RSB_TRY(cudaMemcpyAsync(d_o, origins, n * 24, cudaMemcpyHostToDevice, st));
    RSB_TRY(cudaMemcpyAsync(d_d, directions, n * 24, cudaMemcpyHostToDevice, st));
    if (max_distance) RSB_TRY(cudaMemcpyAsync(d_m, max_distance, n * 8, cudaMemcpyHostToDevice, st));
    RSB_TRY(cudaEventRecord(c->ev0, st));
    rc = rsb_hit_batch_dev(ctx, scene, st, n, d_o, d_d, d_m, d_p, d_t, d_s, d_f, d_n, d_g, d_u, 1);
    if (rc) { cleanup(); return rc; }
    RSB_TRY(cudaEventRecord(c->ev1, st));
    RSB_TRY(cudaMemcpyAsync(out_prim, d_p, n * 4, cudaMemcpyDeviceToHost, st));
    RSB_TRY(cudaMemcpyAsync(out_t, d_t, n * 8, cudaMemcpyDeviceToHost, st));
    RSB_TRY(cudaMemcpyAsync(out_sub, d_s, n * 4, cudaMemcpyDeviceToHost, st));
    RSB_TRY(cudaMemcpyAsync(out_flags, d_f, n, cudaMemcpyDeviceToHost, st));
    if (out_node) RSB_TRY(cudaMemcpyAsync(out_node, d_n, n * 8, cudaMemcpyDeviceToHost, st));
    if (out_geom) RSB_TRY(cudaMemcpyAsync(out_geom, d_g, n * 96, cudaMemcpyDeviceToHost, st));
    if (out_uvw) RSB_TRY(cudaMemcpyAsync(out_uvw, d_u, n * 12, cudaMemcpyDeviceToHost, st));
    RSB_TRY(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1);
    cleanup();
    return RSB_OK;
}

int rsb_hit_sweep_dev(uint64_t ctx, uint64_t scene, void* cuda_stream, int64_t n, int64_t first_index, uint64_t seed,
                      const double* origin, const double* target, double half_window, int32_t order_log2, uint64_t* out_hits_dev,
                      double* out_sum_t_dev, uint64_t* out_xor_prim_dev, int32_t count) {
    Context* c = as_ctx(ctx);
    DeviceScene* ds = as_scene(scene);
    if (!c || !ds || !origin || !target || !out_hits_dev || !out_sum_t_dev || !out_xor_prim_dev) return fail(RSB_ERR_ARG, "null argument");
    if (order_log2 < 0 || order_log2 > 31) return fail(RSB_ERR_ARG, "rsb_hit_sweep: order_log2 must be in [0, 31]");
    if (n <= 0) return RSB_OK;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    RSB_CUDA(cudaSetDevice(c->device));
    if (count) RSB_CUDA(cudaMemsetAsync(c->d_counters, 0, sizeof(DevCounters), st));
    const long long chunk = rq_pass_size(c, ds, n);
    RqHost rq;
    int rc = rq_reserve(c, chunk, ds->has_mesh, &rq, c->query_reorder);
    if (rc) return rc;
    rq.b.md = nullptr;          // every ray of a sweep is unbounded
    rq.b.md_all = RSB_INF;
    for (long long off = 0; off < n; off += chunk) {
        const long long m = std::min<long long>(chunk, n - off);
        const int grid = grid_for(c, m, 256, 8);
        const SweepSource src{first_index + off, seed, origin[0], origin[1], origin[2], target[0], target[1], target[2], half_window, order_log2};
        rq.b.perm = nullptr;
        if (c->query_reorder && m >= kReorderMin && order_log2 == 0) {      // (order_log2 > 0: the caller asked for coherent rays)
            rc = rq_reorder(c, rq, m, src, st);
            if (rc) return rc;
        }
        k_rq_sweep_gen<<<grid, 256, 0, st>>>(m, src, rq.b.perm, rq.ray, rq.b.ray_stride);
        rc = rq_trace(c, ds, rq.b, m, st, count != 0);
        if (rc) return rc;
        k_rq_sweep_reduce<<<grid, 256, 0, st>>>(rq.b, m, first_index + off, (unsigned long long*)out_hits_dev, out_sum_t_dev,
                                                (unsigned long long*)out_xor_prim_dev);
    }
    RSB_CUDA(cudaGetLastError());
    if (count) {
        rc = read_counters(c, st);
        if (rc) return rc;
        c->last_counters.rays = (uint64_t)n;
    }
    return RSB_OK;
}

int rsb_contains_batch(uint64_t ctx, uint64_t scene, int64_t n, const double* points, int32_t cap, int32_t* out_count,
                       int32_t* out_prims) {
    Context* c = as_ctx(ctx);
    DeviceScene* ds = as_scene(scene);
    if (!c || !ds) return fail(RSB_ERR_ARG, "null handle");
    if (n < 0 || !points || !out_count || !out_prims || cap <= 0) return fail(RSB_ERR_ARG, "rsb_contains_batch: bad arguments");
    if (n == 0) return RSB_OK;
    RSB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    double* d_p = nullptr;
    int32_t *d_c = nullptr, *d_o = nullptr;
    auto cleanup = [&]() { cudaFree(d_p); cudaFree(d_c); cudaFree(d_o); };
    RSB_TRY(cudaMalloc(&d_p, n * 24));
    RSB_TRY(cudaMalloc(&d_c, n * 4));
    RSB_TRY(cudaMalloc(&d_o, n * 4 * cap));
    RSB_TRY(cudaMemsetAsync(d_o, 0xff, n * 4 * cap, st));
    RSB_TRY(cudaMemcpyAsync(d_p, points, n * 24, cudaMemcpyHostToDevice, st));
    k_contains_batch<<<grid_for(c, n, 128, 16), 128, 0, st>>>(ds->sc, n, d_p, cap, d_c, d_o);
    RSB_TRY(cudaGetLastError());
    RSB_TRY(cudaMemcpyAsync(out_count, d_c, n * 4, cudaMemcpyDeviceToHost, st));
    RSB_TRY(cudaMemcpyAsync(out_prims, d_o, n * 4 * cap, cudaMemcpyDeviceToHost, st));
    RSB_TRY(cudaStreamSynchronize(st));
    cleanup();
    c->last_counters.contains = (uint64_t)n;
    return RSB_OK;
}

int rsb_rng_uniform(uint64_t ctx, uint64_t seed, int64_t n, double* out) {
    Context* c = as_ctx(ctx);
    if (!c) return fail(RSB_ERR_ARG, "null context");
    if (seed == 0) return fail(RSB_ERR_ARG, "seed must be >= 1 (the reference treats seed(0) as 'reseed from urandom')");
    if (n <= 0 || !out) return fail(RSB_ERR_ARG, "rsb_rng_uniform: bad arguments");
    RSB_CUDA(cudaSetDevice(c->device));
    unsigned long long* d_s = nullptr;
    double* d_o = nullptr;
    auto cleanup = [&]() { cudaFree(d_s); cudaFree(d_o); };
    RSB_TRY(cudaMalloc(&d_s, RSB_MT_NN * 8));
    RSB_TRY(cudaMalloc(&d_o, n * 8));
    k_rng_uniform<<<1, 1, 0, c->stream>>>(seed, n, d_s, d_o);
    RSB_TRY(cudaGetLastError());
    RSB_TRY(cudaMemcpyAsync(out, d_o, n * 8, cudaMemcpyDeviceToHost, c->stream));
    RSB_TRY(cudaStreamSynchronize(c->stream));
    cleanup();
    return RSB_OK;
}

}  // extern "C"

namespace {

// carve the slot state out of one allocation
struct Carver {
    unsigned char* base;
    size_t off = 0;
    template <class T>
    T* take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

size_t carve_slots(unsigned char* base, size_t P, size_t cap, bool mt, size_t chunk, int spp, WfSlots* st, bool mesh, RqBuf* rq) {
    Carver c{base};
    memset(rq, 0, sizeof(*rq));
    if (mesh) {
        // scenes with meshes trace through the World.hit kernel pipeline (rsb_trav.cuh): parked walks and queues per slot
        rq->ctr = c.take<unsigned int>(16);
        rq->susp = c.take<RqSusp>(P);
        rq->susp_stack = c.take<KdStackEntry>(P * RQ_WORLD_STACK);
        rq->queue = c.take<int2>((size_t)(RQ_MAX_ROUNDS + 1) * P);
        rq->cap = (long long)P;
    }
    st->ray = c.take<double>(6 * P);
    st->weight = c.take<double>(P);
    st->norm = c.take<double>(P);
    st->hit_t = c.take<double>(P);
    st->hit_a = c.take<int4>(P);
    st->hit_uvw = c.take<float4>(P);
    st->depth = c.take<int32_t>(P);
    st->rays = c.take<uint32_t>(P);
    st->sample = c.take<int32_t>(P);
    st->px = c.take<int32_t>(P);
    st->py = c.take<int32_t>(P);
    st->status = c.take<int32_t>(P);
    st->log_n = c.take<int32_t>(P);
    st->philox_idx = c.take<uint32_t>(P);
    st->work = c.take<int32_t>(P);
    st->group = c.take<int32_t>(P);
    st->additive = c.take<int32_t>(P);
    st->ended = c.take<int32_t>(2 * P);
    st->n_ended = c.take<unsigned int>(2);
    st->hit_list = c.take<int32_t>(4 * P);
    st->n_hit = c.take<unsigned int>(4);
    st->pix_mti = mt ? c.take<int32_t>(chunk) : nullptr;
    st->pix_mt = mt ? c.take<unsigned long long>(chunk * RSB_MT_NN) : nullptr;
    st->pix_jitter = mt ? c.take<double>(chunk * 2 * (size_t)spp) : nullptr;
    st->log = c.take<LogEntry>(P * cap);
    if (mesh) {
        rq->ray = st->ray;
        rq->ray_stride = (long long)P;
        rq->md = nullptr;
    }
    return (c.off + 255) & ~(size_t)255;
}

template <int RNGMODE, bool COUNT, int FEAT>
int run_wavefront(Context* c, WfArgs& a, size_t smem_scene, size_t smem_shade, size_t smem_tables, cudaStream_t st, bool time_trace,
                  RqBuf rq, int n_mesh_prims) {
    const int threads = 128;
    const int grid = (a.n_slots + threads - 1) / threads;
    const int fin_grid = std::max(1, std::min(grid, c->sm_count * 16));
    const int regen_grid = std::max(1, std::min((grid + 3) / 4, c->sm_count * 8));
    const int shade_grid = grid;
    smem_scene += ax_bytes(FEAT);   // RayAx storage of k_wf_trace behind the staged scene
    if (RNGMODE == RNG_MT19937_64) smem_shade += RSB_MT_WIN_WORDS * 8 * threads;   // k_wf_shade: MT state window per thread
    constexpr bool PIPELINE = (FEAT & RSB_FEAT_MESH) != 0;   // scenes with meshes: World.hit = walk / Mesh.hit / resume kernels
    const int mesh_rounds = std::min<int>(std::max(n_mesh_prims, 1), RQ_MAX_ROUNDS);
    const int grid_mesh = c->sm_count * RQ_MESH_BLOCKS;
    if constexpr (PIPELINE) {
        rq.md_all = a.cfg.max_distance;
        rq.ray_stride = a.n_slots;      // the slot arrays of this chunk are [..][n_slots]
        if (smem_scene > 48 * 1024) {
            RSB_CUDA(cudaFuncSetAttribute(k_rq_walk<COUNT, FEAT, RqWfClient, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scene));
            RSB_CUDA(cudaFuncSetAttribute(k_rq_walk<COUNT, FEAT, RqWfClient, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scene));
            RSB_CUDA(cudaFuncSetAttribute(k_rq_walk<COUNT, FEAT, RqWfClient, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scene));
        }
        RSB_CUDA(cudaFuncSetAttribute(k_rq_mesh<COUNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RQ_MESH_SMEM));
    } else {
        if (smem_scene > 48 * 1024)
            RSB_CUDA(cudaFuncSetAttribute(k_wf_trace<RNGMODE, COUNT, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scene));
    }
    if (smem_shade > 48 * 1024)
        RSB_CUDA(cudaFuncSetAttribute(k_wf_shade<RNGMODE, COUNT, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_shade));
    smem_tables += (threads / 32) * 32 * sizeof(LogEntry);   // k_wf_finalize: one 32-entry log window per warp
    if (a.xyz_mean) smem_tables += (size_t)(threads / 32) * a.proj_channels * a.sp.bins * sizeof(double);   // and the projection terms of one sample per warp
    if (smem_tables > 200 * 1024) return fail(RSB_ERR_UNSUPPORTED, "rsb_render: too many bins per slice for the RGB projection (shared memory)");
    if (smem_tables > 48 * 1024)
    {
        RSB_CUDA(cudaFuncSetAttribute(k_wf_finalize<RNGMODE, COUNT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tables));
        RSB_CUDA(cudaFuncSetAttribute(k_wf_finalize<RNGMODE, COUNT, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tables));
        RSB_CUDA(cudaFuncSetAttribute(k_wf_finalize<RNGMODE, COUNT, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tables));
    }
    // the accumulate kernel that holds exactly the projection code this render needs
    int fin_kind = 0;
    if (a.xyz_mean) {
        fin_kind = -1;
        if (a.proj_channels == 3 && a.proj_mode[0] == RSB_PROJ_XYZ && a.proj_mode[1] == RSB_PROJ_XYZ && a.proj_mode[2] == RSB_PROJ_XYZ) fin_kind = 3;
    }
    a.wave = 0;
    RsbRenderStats& rs = c->render_stats;
    rs.slots = std::max<int64_t>(rs.slots, a.n_slots);
    if (RNGMODE == RNG_MT19937_64) {
        k_wf_seed<<<(unsigned int)((a.n_pixels + 32 * RSB_SEED_WARPS - 1) / (32 * RSB_SEED_WARPS)), 32 * RSB_SEED_WARPS, 0, st>>>(a);
        rs.launches += 1;
    }
    k_wf_init<RNGMODE><<<grid, threads, 0, st>>>(a);
    RSB_CUDA(cudaGetLastError());
    unsigned int* h_idle = c->h_idle;
    const int kBatch = 32;
    rs.launches += 1;
    if (time_trace) {
        while (c->event_pool.size() < 5 * kBatch) {
            cudaEvent_t e;
            RSB_CUDA(cudaEventCreate(&e));
            c->event_pool.push_back(e);
        }
    }
    // One batch = kBatch waves x 4 kernels.  The batch is captured ONCE into a CUDA graph (the kernel
    // arguments only differ in the wave parity) and replayed until every slot is idle: ~115k dependent
    // launches per 1024^2 x 256 spp frame otherwise cost ~5 us of launch gap each.
    auto enqueue_batch = [&](cudaStream_t s, bool external_events) -> int {
        // (time_trace) five events per wave bracket its four phases: trace | shade | finalize | regen
        auto mark = [&](int b, int k) -> int {
            if (!time_trace) return RSB_OK;
            if (external_events) RSB_CUDA(cudaEventRecordWithFlags(c->event_pool[5 * b + k], s, cudaEventRecordExternal));
            else RSB_CUDA(cudaEventRecord(c->event_pool[5 * b + k], s));
            return RSB_OK;
        };
        for (int b = 0; b < kBatch; ++b) {
            a.wave = b;
            if (mark(b, 0)) return RSB_ERR_CUDA;
            if constexpr (PIPELINE) {
                RqWfClient cl;
                cl.a = a;
                RSB_CUDA(cudaMemsetAsync(rq.ctr, 0, 16 * sizeof(unsigned int), s));
                k_rq_walk<COUNT, FEAT, RqWfClient, false, false><<<grid, threads, smem_scene, s>>>(a.sc, a.n_items, cl, rq, a.n_slots, 0, a.counters);
                for (int r = 0; r < mesh_rounds; ++r) {
                    k_rq_mesh<COUNT><<<grid_mesh, RQ_THREADS, RQ_MESH_SMEM, s>>>(a.sc, rq, r, a.counters);
                    if (r == mesh_rounds - 1 && n_mesh_prims > 1)
                        k_rq_walk<COUNT, FEAT, RqWfClient, true, true><<<grid, threads, smem_scene, s>>>(a.sc, a.n_items, cl, rq, a.n_slots, r, a.counters);
                    else
                        k_rq_walk<COUNT, FEAT, RqWfClient, true, false><<<grid, threads, smem_scene, s>>>(a.sc, a.n_items, cl, rq, a.n_slots, r, a.counters);
                }
            } else {
                k_wf_trace<RNGMODE, COUNT, FEAT><<<grid, threads, smem_scene, s>>>(a);
            }
            if (mark(b, 1)) return RSB_ERR_CUDA;
            k_wf_shade<RNGMODE, COUNT, FEAT><<<shade_grid, threads, smem_shade, s>>>(a);
            if (mark(b, 2)) return RSB_ERR_CUDA;
            if (fin_kind == 0) k_wf_finalize<RNGMODE, COUNT, 0><<<fin_grid, threads, smem_tables, s>>>(a);
            else if (fin_kind == 3) k_wf_finalize<RNGMODE, COUNT, 3><<<fin_grid, threads, smem_tables, s>>>(a);
            else k_wf_finalize<RNGMODE, COUNT, -1><<<fin_grid, threads, smem_tables, s>>>(a);
            if (mark(b, 3)) return RSB_ERR_CUDA;
            k_wf_regen<RNGMODE, COUNT><<<regen_grid, threads, 0, s>>>(a);
            if (mark(b, 4)) return RSB_ERR_CUDA;
        }
        return RSB_OK;
    };
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const bool use_graph = c->use_graphs && st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread;
    if (use_graph) {
        RSB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        int erc = enqueue_batch(st, true);
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (erc != RSB_OK || ce != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            return fail(RSB_ERR_CUDA, std::string("graph capture of the wave batch failed: ") + cudaGetErrorString(ce));
        }
        RSB_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    }
    int rc_loop = RSB_OK;
    for (;;) {
        if (use_graph) {
            cudaError_t ge = cudaGraphLaunch(exec, st);
            if (ge != cudaSuccess) { rc_loop = fail(RSB_ERR_CUDA, std::string("cudaGraphLaunch: ") + cudaGetErrorString(ge)); break; }
        } else {
            rc_loop = enqueue_batch(st, false);
            if (rc_loop) break;
        }
        cudaError_t e1 = cudaGetLastError();
        if (e1 == cudaSuccess) e1 = cudaMemcpyAsync(h_idle, a.n_idle, sizeof(unsigned int), cudaMemcpyDeviceToHost, st);
        if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(st);
        if (e1 != cudaSuccess) { rc_loop = fail(RSB_ERR_CUDA, std::string("wave batch: ") + cudaGetErrorString(e1)); break; }
        rs.waves += kBatch;
        rs.launches += (4 + (PIPELINE ? 2 * mesh_rounds : 0)) * kBatch;
        rs.trace_launches += kBatch;
        if (time_trace) {
            for (int b = 0; b < kBatch; ++b) {
                float ms[4] = {0.f, 0.f, 0.f, 0.f};
                for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&ms[k], c->event_pool[5 * b + k], c->event_pool[5 * b + k + 1]);
                rs.trace_ms += ms[0];
                rs.shade_ms += ms[1];
                rs.finalize_ms += ms[2];
                rs.regen_ms += ms[3];
            }
        }
        if (*h_idle >= (unsigned int)a.n_slots) break;
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    if (rc_loop) return rc_loop;
    return RSB_OK;
}

}  // namespace

extern "C" {

int rsb_render_dev(uint64_t ctx, uint64_t scene, void* cuda_stream, const RsbCamera* camera, const RsbRayConfig* config,
                   const RsbSpectral* spectral, const RsbRngDesc* rng, int64_t n_pixels, const int32_t* pixels_dev,
                   double* mean_dev, double* variance_dev, uint64_t* ray_count_dev, int32_t count) {
    return rsb_render_passes_dev(ctx, scene, cuda_stream, camera, config, spectral, rng, 1, 0, n_pixels, pixels_dev, mean_dev,
                                 variance_dev, ray_count_dev, count);
}

int rsb_render_passes_dev(uint64_t ctx, uint64_t scene, void* cuda_stream, const RsbCamera* camera, const RsbRayConfig* config,
                          const RsbSpectral* spectral, const RsbRngDesc* rng, int32_t n_passes, uint64_t seed_stride,
                          int64_t n_pixels, const int32_t* pixels_dev, double* mean_dev, double* variance_dev,
                          uint64_t* ray_count_dev, int32_t count) {
    return rsb_render_slices_dev(ctx, scene, cuda_stream, camera, config, spectral, rng, n_passes, 1, seed_stride, n_pixels, pixels_dev,
                                 mean_dev, variance_dev, ray_count_dev, count);
}

// device-resident XYZ side of a render: curves and per-work-item statistics (see WfArgs)
struct XyzDev {
    const double* tab = nullptr;
    const double* delta = nullptr;
    double* mean = nullptr;
    double* variance = nullptr;
    int32_t n_channels = 0;
    int32_t mode[RSB_PROJ_MAX] = {0};
};

static int render_slices_impl(uint64_t ctx, uint64_t scene, void* cuda_stream, const RsbCamera* camera, const RsbRayConfig* config,
                              const RsbSpectral* spectral, const RsbRngDesc* rng, int32_t n_passes, int32_t n_slices, uint64_t seed_stride,
                              int64_t n_pixels, const int32_t* pixels_dev, double* mean_dev, double* variance_dev,
                              uint64_t* ray_count_dev, int32_t count, const XyzDev& xyz);

int rsb_render_slices_dev(uint64_t ctx, uint64_t scene, void* cuda_stream, const RsbCamera* camera, const RsbRayConfig* config,
                          const RsbSpectral* spectral, const RsbRngDesc* rng, int32_t n_passes, int32_t n_slices, uint64_t seed_stride,
                          int64_t n_pixels, const int32_t* pixels_dev, double* mean_dev, double* variance_dev,
                          uint64_t* ray_count_dev, int32_t count) {
    if (!mean_dev || !variance_dev) return fail(RSB_ERR_ARG, "rsb_render: null argument");
    return render_slices_impl(ctx, scene, cuda_stream, camera, config, spectral, rng, n_passes, n_slices, seed_stride, n_pixels, pixels_dev,
                              mean_dev, variance_dev, ray_count_dev, count, XyzDev{});
}

static int render_slices_impl(uint64_t ctx, uint64_t scene, void* cuda_stream, const RsbCamera* camera, const RsbRayConfig* config,
                              const RsbSpectral* spectral, const RsbRngDesc* rng, int32_t n_passes, int32_t n_slices, uint64_t seed_stride,
                              int64_t n_pixels, const int32_t* pixels_dev, double* mean_dev, double* variance_dev,
                              uint64_t* ray_count_dev, int32_t count, const XyzDev& xyz) {
    Context* c = as_ctx(ctx);
    DeviceScene* ds = as_scene(scene);
    if (!c || !ds || !camera || !config || !spectral || !rng || !ray_count_dev || (!mean_dev != !variance_dev) || (!mean_dev && !xyz.mean))
        return fail(RSB_ERR_ARG, "rsb_render: null argument");
    if (camera->nx < 1 || camera->ny < 1 || camera->pixel_samples < 1) return fail(RSB_ERR_ARG, "rsb_render: bad camera");
    if (camera->kind == RSB_CAMERA_VECTOR && (!camera->pixel_origins || !camera->pixel_directions))
        return fail(RSB_ERR_ARG, "rsb_render: a vector camera needs pixel_origins and pixel_directions");
    if (camera->kind != RSB_CAMERA_PINHOLE && camera->kind != RSB_CAMERA_ORTHOGRAPHIC && camera->kind != RSB_CAMERA_CCD &&
        camera->kind != RSB_CAMERA_VECTOR && camera->kind != RSB_CAMERA_PIXEL)
        return fail(RSB_ERR_UNSUPPORTED, "rsb_render: unknown camera kind");
    if (config->bins < 1) return fail(RSB_ERR_ARG, "Number of bins cannot be less than 1.");
    if (config->bins != spectral->bins) return fail(RSB_ERR_ARG, "rsb_render: ray bins and spectral table bins differ");
    if (spectral->n_materials != ds->n_materials) return fail(RSB_ERR_ARG, "rsb_render: spectral tables do not match the scene's materials");
    const int n_tables = spectral->n_tables > 0 ? spectral->n_tables : spectral->n_materials;
    if (n_tables < spectral->n_materials) return fail(RSB_ERR_ARG, "rsb_render: n_tables is smaller than the number of materials");
    for (int i = 0; i < ds->n_materials; ++i)
        if ((ds->mat_type[i] == RSB_MAT_CONDUCTOR || ds->mat_type[i] == RSB_MAT_ROUGH_CONDUCTOR) &&
            (!spectral->table2 || spectral->table2[i] < 0 || spectral->table2[i] >= n_tables))
            return fail(RSB_ERR_ARG, "rsb_render: a Conductor needs its extinction table (RsbSpectral.table2)");
    for (int i = 0; i < ds->n_materials && spectral->table2; ++i)
        if (spectral->table2[i] >= n_tables) return fail(RSB_ERR_ARG, "rsb_render: RsbSpectral.table2 points past the tables");
    if (config->important_path_weight < 0 || config->important_path_weight > 1.0)
        return fail(RSB_ERR_ARG, "Important path weight must be in the range [0, 1].");
    if (rng->seed == 0) return fail(RSB_ERR_ARG, "rng seed must be >= 1");
    if (rng->mode != RSB_RNG_MT19937_64 && rng->mode != RSB_RNG_PHILOX) return fail(RSB_ERR_ARG, "unknown rng mode");
    if (n_passes < 1 || n_passes > 1024) return fail(RSB_ERR_ARG, "rsb_render_passes: the number of passes must be in [1, 1024]");
    if (n_slices < 1 || n_slices > 65536) return fail(RSB_ERR_ARG, "rsb_render_slices: the number of slices must be in [1, 65536]");
    for (int k = 1; k < n_slices; ++k)
        if (spectral[k].bins != spectral->bins || spectral[k].n_materials != spectral->n_materials ||
            (spectral[k].n_tables > 0 ? spectral[k].n_tables : spectral[k].n_materials) != (spectral->n_tables > 0 ? spectral->n_tables : spectral->n_materials))
            return fail(RSB_ERR_ARG, "rsb_render_slices: every slice must have the same number of bins, materials and tables");
    if (!pixels_dev) n_pixels = (int64_t)camera->nx * camera->ny;
    if (n_pixels <= 0) return RSB_OK;
    cudaStream_t caller = (cudaStream_t)cuda_stream;
    RSB_CUDA(cudaSetDevice(c->device));
    // The wave loop is stream-captured into a CUDA graph, which the legacy default stream does not allow: run
    // on the context's own stream, ordered after / before the caller's stream with events.
    cudaStream_t st = c->stream;
    if (caller != st) {
        RSB_CUDA(cudaEventRecord(c->ev_in, caller));
        RSB_CUDA(cudaStreamWaitEvent(st, c->ev_in, 0));
    }

    // ---- per-slice material rows and spectral tables ------------------------------------------------------
    int nm = ds->n_materials;
    std::vector<Material> mats((size_t)nm * n_slices);
    const size_t tb = (size_t)n_tables * spectral->bins;
    // per slice: the tables, followed by their natural logs (exp(length * ln T) form of the Beer-Lambert pow in the replay)
    std::vector<double> both(tb * 2 * (size_t)n_slices);
    for (int k = 0; k < n_slices; ++k) {
        const RsbSpectral& sk = spectral[k];
        if ((sk.table2 == nullptr) != (spectral->table2 == nullptr)) return fail(RSB_ERR_ARG, "rsb_render_slices: inconsistent second tables");
        for (int i = 0; i < nm; ++i) {
            Material& m = mats[(size_t)k * nm + i];
            memset(&m, 0, sizeof(m));
            m.type = ds->mat_type[i];
            m.transmission_only = ds->mat_transmission_only[i];
            m.table = i;
            m.table2 = sk.table2 ? sk.table2[i] : -1;
            m.scale = sk.scale ? sk.scale[i] : 1.0;
            m.index_in = sk.index_in ? sk.index_in[i] : 1.0;
            m.index_out = sk.index_out ? sk.index_out[i] : 1.0;
        }
        double* dst = both.data() + (size_t)k * 2 * tb;
        for (size_t i = 0; i < tb; ++i) {
            dst[i] = sk.tables[i];
            dst[tb + i] = log(sk.tables[i]);
        }
    }
    size_t mat_bytes = ((mats.size() * sizeof(Material) + 15) / 16) * 16;
    size_t tab_bytes = ((tb * 8 + 15) / 16) * 16;
    size_t tab_alloc = both.size() * 8 + 16;
    if (c->mats_cap < mat_bytes) {
        cudaFree(c->d_mats);
        c->d_mats = nullptr; c->mats_cap = 0;
        RSB_CUDA(cudaMalloc(&c->d_mats, mat_bytes));
        c->mats_cap = mat_bytes;
    }
    if (c->tables_cap < tab_alloc) {
        cudaFree(c->d_tables);
        c->d_tables = nullptr; c->tables_cap = 0;
        RSB_CUDA(cudaMalloc(&c->d_tables, tab_alloc));
        c->tables_cap = tab_alloc;
    }
    RSB_CUDA(cudaMemcpyAsync(c->d_mats, mats.data(), mats.size() * sizeof(Material), cudaMemcpyHostToDevice, st));
    RSB_CUDA(cudaMemcpyAsync(c->d_tables, both.data(), both.size() * 8, cudaMemcpyHostToDevice, st));
    // the two host staging buffers above are stack/heap temporaries: make the copies complete before returning
    RSB_CUDA(cudaStreamSynchronize(st));

    WfArgs a;
    memset(&a, 0, sizeof(a));
    a.sc = ds->sc;
    a.sp.mats = c->d_mats;
    a.sp.tables = c->d_tables;
    a.sp.tables_ln = c->d_tables + (size_t)n_tables * spectral->bins;
    a.sp.bins = spectral->bins;
    a.sp.n_materials = nm;
    a.sp.n_tables = n_tables;
    a.cfg.bins = config->bins;
    a.cfg.extinction_min_depth = config->extinction_min_depth;
    a.cfg.max_depth = config->max_depth;
    a.cfg.importance_sampling = config->importance_sampling;
    a.cfg.min_wavelength = config->min_wavelength;
    a.cfg.max_wavelength = config->max_wavelength;
    a.cfg.extinction_prob = config->extinction_prob;
    a.cfg.important_path_weight = config->important_path_weight;
    a.cfg.max_distance = config->max_distance;
    a.cam.nx = camera->nx;
    a.cam.ny = camera->ny;
    a.cam.pixel_samples = camera->pixel_samples;
    a.cam.kind = camera->kind;
    a.cam.image_delta = camera->image_delta;
    a.cam.image_start_x = camera->image_start_x;
    a.cam.image_start_y = camera->image_start_y;
    a.cam.sensitivity = camera->sensitivity;
    memcpy(a.cam.to_root, camera->to_root, 12 * sizeof(double));
    if (camera->kind == RSB_CAMERA_VECTOR) {
        // per-pixel origins and directions of the VectorCamera: [nx][ny][3] each, uploaded per render (grow-only buffer)
        const size_t n3 = (size_t)camera->nx * camera->ny * 3;
        if (c->campix_cap < 2 * n3) {
            cudaFree(c->d_campix);
            c->d_campix = nullptr; c->campix_cap = 0;
            RSB_CUDA(cudaMalloc(&c->d_campix, 2 * n3 * sizeof(double)));
            c->campix_cap = 2 * n3;
        }
        RSB_CUDA(cudaMemcpyAsync(c->d_campix, camera->pixel_origins, n3 * 8, cudaMemcpyHostToDevice, st));
        RSB_CUDA(cudaMemcpyAsync(c->d_campix + n3, camera->pixel_directions, n3 * 8, cudaMemcpyHostToDevice, st));
        RSB_CUDA(cudaStreamSynchronize(st));        // the caller's arrays may go away after the call
        a.cam.pixel_origins = c->d_campix;
        a.cam.pixel_directions = c->d_campix + n3;
    }
    if (camera->to_root_w == 0.0) return fail(RSB_ERR_ARG, "rsb_render: RsbCamera.to_root_w (m33 of the camera transform) is zero");
    a.cam.to_root[12] = 1.0 / camera->to_root_w;
    a.mean = mean_dev;
    a.variance = variance_dev;
    a.xyz_tab = xyz.tab;
    a.xyz_delta = xyz.delta;
    a.xyz_mean = xyz.mean;
    a.xyz_variance = xyz.variance;
    a.proj_channels = xyz.n_channels;
    for (int k = 0; k < RSB_PROJ_MAX; ++k) a.proj_mode[k] = xyz.mode[k];
    a.ray_count = (unsigned long long*)ray_count_dev;
    a.work_counter = c->d_scalars;
    a.n_idle = (unsigned int*)(c->d_scalars + 1);
    a.overflow_flag = (int32_t*)(c->d_scalars + 2);
    RSB_CUDA(cudaMemsetAsync(c->d_scalars + 2, 0, sizeof(unsigned long long), st));
    a.counters = c->d_counters;
    a.seed = rng->seed;
    a.seed_stride = seed_stride;
    a.n_passes = n_passes;
    a.n_slices = n_slices;
    a.frame_bins = n_slices * config->bins;
    for (int i = 0; i < ds->n_materials; ++i)
        if (ds->mat_type[i] == RSB_MAT_VOLUME_EMITTER) a.has_additive = 1;
    a.n_pix_pass = n_pixels;
    a.pixels = pixels_dev;
    a.frame_elems = (long long)camera->nx * camera->ny * a.frame_bins;
    if (n_passes > 1 && mean_dev) {
        size_t need_pass = (size_t)2 * (size_t)(n_passes - 1) * (size_t)a.frame_elems * sizeof(double);
        if (c->pass_bytes < need_pass) {
            cudaFree(c->d_pass);
            c->d_pass = nullptr; c->pass_bytes = 0;
            RSB_CUDA(cudaMalloc(&c->d_pass, need_pass));
            c->pass_bytes = need_pass;
        }
        a.pass_mean = c->d_pass;
        a.pass_variance = c->d_pass + (size_t)(n_passes - 1) * (size_t)a.frame_elems;
    }
    a.n_items = ds->n_world_items;
    a.staged = ds->stage_bytes ? 1 : 0;
    size_t smem_scene = ds->stage_bytes, smem_shade = ds->stage_bytes, smem_tables = 0;
    if (n_slices == 1 && smem_shade + mat_bytes <= kMaxStageBytes && 2 * tab_bytes <= kMaxStageBytes) {
        a.tables_staged = 1;
        smem_shade += mat_bytes;
        smem_tables = 2 * tab_bytes;
    }
    // a path of D segments logs at most 3 surface + 1 roulette entries per segment plus one per enclosing
    // dielectric; budget 6 per segment
    long long cap = 6LL * ((long long)std::max(config->max_depth, config->extinction_min_depth) + 2);
    a.log_capacity = (int32_t)std::min(cap, 1LL << 16);

    // ---- slot pool: one pixel stream per slot, two CTA waves of 1024 threads per SM ---------------------
    // Pixels are processed in chunks so that the up-front MT19937-64 state of a chunk (5 KB per pixel) stays
    // within a fixed HBM budget; a 1024 x 1024 frame is one chunk (5.2 GB).
    bool mt = rng->mode == RSB_RNG_MT19937_64;
    const long long n_work = (long long)n_pixels * n_passes * n_slices;   // work items: (pass, slice, pixel task)
    long long chunk_cap = std::min<long long>(n_work, c->chunk_items);
    long long P = std::min<long long>(chunk_cap, (long long)c->sm_count * c->slots_per_sm);
    P = std::max<long long>(P, 1);
    WfSlots probe;
    RqBuf rq;
    size_t need = carve_slots(nullptr, (size_t)P, (size_t)a.log_capacity, mt, (size_t)chunk_cap, camera->pixel_samples * camera_jitter_pairs(camera->kind), &probe, ds->has_mesh, &rq);
    if (c->slot_bytes < need) {
        // the per-slot path log grows with ray_max_depth (48 KB per slot at Raysect's default 500): narrow the wavefront
        // and the seeded chunk until the pool fits what the device has left (another context, another process, ...)
        size_t free_b = 0, total_b = 0;
        RSB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        free_b += c->slot_bytes;
        while (need > free_b - free_b / 8 && (P > (long long)c->sm_count * 128 || chunk_cap > P)) {
            if (chunk_cap > P) chunk_cap = std::max<long long>(P, chunk_cap / 2);
            else { P = std::max<long long>((long long)c->sm_count * 128, P * 3 / 4); chunk_cap = std::min(chunk_cap, P); }
            need = carve_slots(nullptr, (size_t)P, (size_t)a.log_capacity, mt, (size_t)chunk_cap, camera->pixel_samples * camera_jitter_pairs(camera->kind), &probe, ds->has_mesh, &rq);
        }
    }
    if (c->slot_bytes < need) {
        cudaFree(c->d_slots);
        c->d_slots = nullptr; c->slot_bytes = 0;
        RSB_CUDA(cudaMalloc(&c->d_slots, need));
        c->slot_bytes = need;
    }
    carve_slots(c->d_slots, (size_t)P, (size_t)a.log_capacity, mt, (size_t)chunk_cap, camera->pixel_samples * camera_jitter_pairs(camera->kind), &a.st, ds->has_mesh, &rq);
    a.st.mt_table = c->d_mt_table;
    if (count & RSB_RENDER_COUNT) RSB_CUDA(cudaMemsetAsync(c->d_counters, 0, sizeof(DevCounters), st));
    const bool time_trace = (count & RSB_RENDER_TIME_TRACE) != 0;
    count &= RSB_RENDER_COUNT;
    c->render_stats = RsbRenderStats{};
    int rc = RSB_OK;
    for (long long base = 0; base < n_work && rc == RSB_OK; base += chunk_cap) {
        a.n_pixels = std::min<long long>(chunk_cap, n_work - base);
        a.item_base = base;
        a.n_slots = (int32_t)std::min<long long>(P, a.n_pixels);
        // scalars: [0] work counter, [1] idle slots, [2] overflow flag (sticky across chunks)
        RSB_CUDA(cudaMemsetAsync(c->d_scalars, 0, 2 * sizeof(unsigned long long), st));
        RSB_CUDA(cudaMemsetAsync(a.st.n_ended, 0, 2 * sizeof(unsigned int), st));
        RSB_CUDA(cudaMemsetAsync(a.st.n_hit, 0, 4 * sizeof(unsigned int), st));
        // kernels are instantiated for "analytic primitives only" and for "everything" (meshes and CSG)
        // conductors and volume emitters are compiled into the full-featured instantiation only (RSB_FEAT_RARE_MATERIALS)
        bool rare = false;
        for (int i = 0; i < ds->n_materials; ++i) rare = rare || ds->mat_type[i] >= RSB_MAT_CONDUCTOR;
        // the MESH bit is set exactly when the scene holds a mesh: those instantiations trace through the kernel pipeline
        const int feat = ((ds->has_csg || rare) ? RSB_FEAT_CSG : 0) | (ds->has_mesh ? RSB_FEAT_MESH : 0);
#define RSB_RUN_MT(C, F) rc = run_wavefront<RNG_MT19937_64, C, F>(c, a, smem_scene, smem_shade, smem_tables, st, time_trace, rq, ds->n_mesh_prims)
#define RSB_RUN_PX(C, F) rc = run_wavefront<RNG_PHILOX, C, F>(c, a, smem_scene, smem_shade, smem_tables, st, time_trace, rq, ds->n_mesh_prims)
        if (mt) RSB_DISPATCH_FEAT(count != 0, feat, a.staged, RSB_RUN_MT);
        else RSB_DISPATCH_FEAT(count != 0, feat, a.staged, RSB_RUN_PX);
#undef RSB_RUN_MT
#undef RSB_RUN_PX
    }
    if (rc) return rc;
    if (n_passes > 1 && mean_dev) {
        long long total = (long long)n_pixels * a.frame_bins;
        k_pass_combine<<<grid_for(c, total, 256, 8), 256, 0, st>>>(n_pixels, pixels_dev, camera->ny, a.frame_bins, n_passes,
                                                                    camera->pixel_samples, a.frame_elems, a.pass_mean, a.pass_variance,
                                                                    mean_dev, variance_dev);
        RSB_CUDA(cudaGetLastError());
        c->render_stats.launches += 1;
    }
    {
        int32_t overflow = 0;
        RSB_CUDA(cudaMemcpyAsync(&overflow, a.overflow_flag, 4, cudaMemcpyDeviceToHost, st));
        RSB_CUDA(cudaStreamSynchronize(st));
        if (overflow) return fail(RSB_ERR_OVERFLOW, "rsb_render: a path exceeded the per-path log capacity");
    }
    if (count) {
        rc = read_counters(c, st);
        if (rc) return rc;
    }
    if (caller != st) {
        RSB_CUDA(cudaEventRecord(c->ev_out, st));
        RSB_CUDA(cudaStreamWaitEvent(caller, c->ev_out, 0));
    }
    return RSB_OK;
}

int rsb_render(uint64_t ctx, uint64_t scene, const RsbCamera* camera, const RsbRayConfig* config, const RsbSpectral* spectral,
               const RsbRngDesc* rng, int64_t n_pixels, const int32_t* pixels, double* mean, double* variance, uint64_t* ray_count) {
    return rsb_render_passes(ctx, scene, camera, config, spectral, rng, 1, 0, n_pixels, pixels, mean, variance, ray_count);
}

int rsb_render_slice(uint64_t ctx, uint64_t scene, const RsbCamera* camera, const RsbRayConfig* config, const RsbSpectral* spectral,
                     const RsbRngDesc* rng, int32_t n_passes, uint64_t seed_stride, int64_t n_pixels, const int32_t* pixels,
                     uint64_t* ray_count) {
    return rsb_render_slices(ctx, scene, camera, config, spectral, rng, n_passes, 1, seed_stride, n_pixels, pixels, ray_count);
}

static int render_slices_host(uint64_t ctx, uint64_t scene, const RsbCamera* camera, const RsbRayConfig* config, const RsbSpectral* spectral,
                              const RsbRngDesc* rng, int32_t n_passes, int32_t n_slices, uint64_t seed_stride, int64_t n_pixels,
                              const int32_t* pixels, int32_t n_channels, const int32_t* channel_mode, const double* resampled_xyz,
                              const double* delta_wavelength, bool keep_spectral, uint64_t* ray_count) {
    Context* c = as_ctx(ctx);
    if (!c || !as_scene(scene) || !camera || !config || !ray_count) return fail(RSB_ERR_ARG, "rsb_render_slice: null argument");
    if (n_slices < 1) return fail(RSB_ERR_ARG, "rsb_render_slices: the number of slices must be at least 1");
    if (n_passes < 1) return fail(RSB_ERR_ARG, "rsb_render_passes: the number of passes must be in [1, 1024]");
    if (config->bins < 1) return fail(RSB_ERR_ARG, "Number of bins cannot be less than 1.");
    const bool want_xyz = resampled_xyz != nullptr;
    if (want_xyz && (!delta_wavelength || !channel_mode)) return fail(RSB_ERR_ARG, "rsb_render_slices_proj: null argument");
    if (want_xyz && (n_channels < 1 || n_channels > RSB_PROJ_MAX))
        return fail(RSB_ERR_ARG, "rsb_render_slices_proj: the number of projection channels must be in [1, 8]");
    for (int k = 0; want_xyz && k < n_channels; ++k)
        if (channel_mode[k] != RSB_PROJ_XYZ && channel_mode[k] != RSB_PROJ_POWER && channel_mode[k] != RSB_PROJ_RADIANCE)
            return fail(RSB_ERR_ARG, "rsb_render_slices_proj: unknown channel mode");
    if (!want_xyz && !keep_spectral) return fail(RSB_ERR_ARG, "rsb_render_slices_xyz: nothing to render (no XYZ curves, no spectral frame)");
    RSB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    c->slice.valid = false;
    const size_t frame = (size_t)camera->nx * camera->ny * config->bins * n_slices;
    if (!pixels) n_pixels = (int64_t)camera->nx * camera->ny;
    if (n_pixels < 0) return fail(RSB_ERR_ARG, "rsb_render_slice: negative pixel count");
    if (keep_spectral && c->slice_cap < 2 * frame) {
        cudaFree(c->d_slice);
        c->d_slice = nullptr; c->slice_cap = 0;
        RSB_CUDA(cudaMalloc(&c->d_slice, 2 * frame * sizeof(double)));
        c->slice_cap = 2 * frame;
    }
    XyzDev xyz;
    if (want_xyz) {
        const size_t n_tab = (size_t)n_slices * config->bins * n_channels, n_work = (size_t)n_pixels * n_passes * n_slices;
        const size_t need = n_tab + (size_t)n_slices + 2 * (size_t)n_channels * n_work;
        if (c->xyz_cap < need) {
            cudaFree(c->d_xyz);
            c->d_xyz = nullptr; c->xyz_cap = 0;
            RSB_CUDA(cudaMalloc(&c->d_xyz, std::max<size_t>(1, need) * sizeof(double)));
            c->xyz_cap = need;
        }
        RSB_CUDA(cudaMemcpyAsync(c->d_xyz, resampled_xyz, n_tab * 8, cudaMemcpyHostToDevice, st));
        RSB_CUDA(cudaMemcpyAsync(c->d_xyz + n_tab, delta_wavelength, (size_t)n_slices * 8, cudaMemcpyHostToDevice, st));
        xyz.tab = c->d_xyz;
        xyz.delta = c->d_xyz + n_tab;
        xyz.mean = c->d_xyz + n_tab + n_slices;
        xyz.variance = xyz.mean + (size_t)n_channels * n_work;
        xyz.n_channels = n_channels;
        for (int k = 0; k < n_channels; ++k) xyz.mode[k] = channel_mode[k];
    }
    if (!c->d_slice_rays) RSB_CUDA(cudaMalloc(&c->d_slice_rays, 8));
    if (pixels && c->slice_pix_cap < (size_t)n_pixels) {
        cudaFree(c->d_slice_pix);
        c->d_slice_pix = nullptr; c->slice_pix_cap = 0;
        RSB_CUDA(cudaMalloc(&c->d_slice_pix, std::max<size_t>(1, (size_t)n_pixels) * 8));
        c->slice_pix_cap = (size_t)n_pixels;
    }
    RSB_CUDA(cudaMemsetAsync(c->d_slice_rays, 0, 8, st));
    // unlisted pixels of the slice read as zero
    if (keep_spectral) RSB_CUDA(cudaMemsetAsync(c->d_slice, 0, 2 * frame * sizeof(double), st));
    if (pixels && n_pixels > 0) RSB_CUDA(cudaMemcpyAsync(c->d_slice_pix, pixels, (size_t)n_pixels * 8, cudaMemcpyHostToDevice, st));
    RSB_CUDA(cudaEventRecord(c->ev0, st));
    if (n_pixels > 0) {
        int rc = render_slices_impl(ctx, scene, st, camera, config, spectral, rng, n_passes, n_slices, seed_stride, n_pixels,
                                    pixels ? c->d_slice_pix : nullptr, keep_spectral ? c->d_slice : nullptr,
                                    keep_spectral ? c->d_slice + frame : nullptr, (uint64_t*)c->d_slice_rays, 0, xyz);
        if (rc) return rc;
    }
    RSB_CUDA(cudaEventRecord(c->ev1, st));
    unsigned long long rays = 0;
    RSB_CUDA(cudaMemcpyAsync(&rays, c->d_slice_rays, 8, cudaMemcpyDeviceToHost, st));
    RSB_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1);
    *ray_count += rays;
    c->slice.nx = camera->nx; c->slice.ny = camera->ny; c->slice.bins = config->bins * n_slices;
    c->slice.samples = camera->pixel_samples * n_passes;
    c->slice.n_pixels = n_pixels;
    c->slice.listed = pixels != nullptr;
    c->slice.has_bins = keep_spectral;
    c->slice.has_xyz = want_xyz;
    c->slice.n_channels = want_xyz ? n_channels : 0;
    c->slice.slice_bins = config->bins;
    c->slice.n_passes = n_passes; c->slice.n_slices = n_slices; c->slice.pass_samples = camera->pixel_samples;
    c->slice.valid = true;
    return RSB_OK;
}

int rsb_render_slices(uint64_t ctx, uint64_t scene, const RsbCamera* camera, const RsbRayConfig* config, const RsbSpectral* spectral,
                      const RsbRngDesc* rng, int32_t n_passes, int32_t n_slices, uint64_t seed_stride, int64_t n_pixels,
                      const int32_t* pixels, uint64_t* ray_count) {
    return render_slices_host(ctx, scene, camera, config, spectral, rng, n_passes, n_slices, seed_stride, n_pixels, pixels, 0, nullptr, nullptr,
                              nullptr, true, ray_count);
}

int rsb_render_slices_proj(uint64_t ctx, uint64_t scene, const RsbCamera* camera, const RsbRayConfig* config, const RsbSpectral* spectral,
                           const RsbRngDesc* rng, int32_t n_passes, int32_t n_slices, uint64_t seed_stride, int64_t n_pixels,
                           const int32_t* pixels, int32_t n_channels, const int32_t* channel_mode, const double* curves,
                           const double* delta_wavelength, int32_t keep_spectral, uint64_t* ray_count) {
    if (!curves || !delta_wavelength || !channel_mode) return fail(RSB_ERR_ARG, "rsb_render_slices_proj: null argument");
    return render_slices_host(ctx, scene, camera, config, spectral, rng, n_passes, n_slices, seed_stride, n_pixels, pixels, n_channels,
                              channel_mode, curves, delta_wavelength, keep_spectral != 0, ray_count);
}

int rsb_render_slices_xyz(uint64_t ctx, uint64_t scene, const RsbCamera* camera, const RsbRayConfig* config, const RsbSpectral* spectral,
                          const RsbRngDesc* rng, int32_t n_passes, int32_t n_slices, uint64_t seed_stride, int64_t n_pixels,
                          const int32_t* pixels, const double* resampled_xyz, const double* delta_wavelength, int32_t keep_spectral,
                          uint64_t* ray_count) {
    const int32_t modes[3] = {RSB_PROJ_XYZ, RSB_PROJ_XYZ, RSB_PROJ_XYZ};
    return rsb_render_slices_proj(ctx, scene, camera, config, spectral, rng, n_passes, n_slices, seed_stride, n_pixels, pixels, 3, modes,
                                  resampled_xyz, delta_wavelength, keep_spectral, ray_count);
}

static int slice_update_proj(uint64_t ctx, int32_t channel0, int32_t n_channels, bool bayer, int32_t frame_is_empty, double* frame_mean,
                             double* frame_variance, int32_t* frame_samples);

int rsb_slice_update_proj_frame(uint64_t ctx, int32_t channel0, int32_t n_channels, int32_t frame_is_empty, double* frame_mean,
                                double* frame_variance, int32_t* frame_samples) {
    return slice_update_proj(ctx, channel0, n_channels, false, frame_is_empty, frame_mean, frame_variance, frame_samples);
}

int rsb_slice_update_bayer_frame(uint64_t ctx, int32_t channel0, int32_t frame_is_empty, double* frame_mean, double* frame_variance,
                                 int32_t* frame_samples) {
    return slice_update_proj(ctx, channel0, 3, true, frame_is_empty, frame_mean, frame_variance, frame_samples);
}

static int slice_update_proj(uint64_t ctx, int32_t channel0, int32_t n_channels, bool bayer, int32_t frame_is_empty, double* frame_mean,
                             double* frame_variance, int32_t* frame_samples) {
    Context* c = as_ctx(ctx);
    if (!c || !frame_mean || !frame_variance || !frame_samples) return fail(RSB_ERR_ARG, "rsb_slice_update_proj_frame: null argument");
    if (!c->slice.valid || !c->slice.has_xyz)
        return fail(RSB_ERR_ARG, "rsb_slice_update_proj_frame: no rendered projection statistics are held (call rsb_render_slices_proj first)");
    if (channel0 < 0 || n_channels < 1 || channel0 + n_channels > c->slice.n_channels)
        return fail(RSB_ERR_ARG, "rsb_slice_update_proj_frame: channel range outside the rendered channels");
    if (c->slice.n_pixels == 0) return RSB_OK;
    RSB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int frame_channels = bayer ? 1 : n_channels;       // values per pixel of the host frame
    const size_t elems = (size_t)c->slice.nx * c->slice.ny * frame_channels;
    if (c->xyz_frame_cap < elems) {
        cudaFree(c->d_xyz_frame); cudaFree(c->d_xyz_samples);
        c->d_xyz_frame = nullptr; c->d_xyz_samples = nullptr; c->xyz_frame_cap = 0;
        RSB_CUDA(cudaMalloc(&c->d_xyz_frame, 2 * elems * sizeof(double)));
        RSB_CUDA(cudaMalloc(&c->d_xyz_samples, elems * sizeof(int32_t)));
        c->xyz_frame_cap = elems;
    }
    double* d_fm = c->d_xyz_frame;
    double* d_fv = c->d_xyz_frame + elems;
    if (frame_is_empty) {
        RSB_CUDA(cudaMemsetAsync(d_fm, 0, 2 * elems * sizeof(double), st));
        RSB_CUDA(cudaMemsetAsync(c->d_xyz_samples, 0, elems * sizeof(int32_t), st));
    } else {
        RSB_CUDA(cudaMemcpyAsync(d_fm, frame_mean, elems * 8, cudaMemcpyHostToDevice, st));
        RSB_CUDA(cudaMemcpyAsync(d_fv, frame_variance, elems * 8, cudaMemcpyHostToDevice, st));
        RSB_CUDA(cudaMemcpyAsync(c->d_xyz_samples, frame_samples, elems * 4, cudaMemcpyHostToDevice, st));
    }
    const int nch = c->slice.n_channels;
    const size_t n_tab = (size_t)c->slice.n_slices * c->slice.slice_bins * nch;
    const size_t n_work = (size_t)c->slice.n_pixels * c->slice.n_passes * c->slice.n_slices;
    const double* wm = c->d_xyz + n_tab + c->slice.n_slices;
    const double* wv = wm + (size_t)nch * n_work;
    k_xyz_combine<<<grid_for(c, (long long)c->slice.n_pixels * frame_channels, 256, 4), 256, 0, st>>>(
        c->slice.n_pixels, c->slice.listed ? c->d_slice_pix : nullptr, c->slice.ny, c->slice.n_passes, c->slice.n_slices, c->slice.pass_samples,
        nch, channel0, frame_channels, bayer ? 1 : 0, wm, wv, d_fm, d_fv, c->d_xyz_samples);
    RSB_CUDA(cudaGetLastError());
    RSB_CUDA(cudaMemcpyAsync(frame_mean, d_fm, elems * 8, cudaMemcpyDeviceToHost, st));
    RSB_CUDA(cudaMemcpyAsync(frame_variance, d_fv, elems * 8, cudaMemcpyDeviceToHost, st));
    RSB_CUDA(cudaMemcpyAsync(frame_samples, c->d_xyz_samples, elems * 4, cudaMemcpyDeviceToHost, st));
    RSB_CUDA(cudaStreamSynchronize(st));
    return RSB_OK;
}

int rsb_slice_update_xyz_frame(uint64_t ctx, int32_t frame_is_empty, double* xyz_mean, double* xyz_variance, int32_t* xyz_samples) {
    Context* c = as_ctx(ctx);
    if (c && c->slice.valid && c->slice.has_xyz && c->slice.n_channels != 3)
        return fail(RSB_ERR_ARG, "rsb_slice_update_xyz_frame: the held statistics are not those of rsb_render_slices_xyz");
    return rsb_slice_update_proj_frame(ctx, 0, 3, frame_is_empty, xyz_mean, xyz_variance, xyz_samples);
}

int rsb_slice_read(uint64_t ctx, double* mean, double* variance) {
    Context* c = as_ctx(ctx);
    if (!c || !mean || !variance) return fail(RSB_ERR_ARG, "rsb_slice_read: null argument");
    if (!c->slice.valid || !c->slice.has_bins) return fail(RSB_ERR_ARG, "rsb_slice_read: no rendered slice is held (call rsb_render_slice first)");
    RSB_CUDA(cudaSetDevice(c->device));
    const size_t frame = (size_t)c->slice.nx * c->slice.ny * c->slice.bins;
    RSB_CUDA(cudaMemcpyAsync(mean, c->d_slice, frame * 8, cudaMemcpyDeviceToHost, c->stream));
    RSB_CUDA(cudaMemcpyAsync(variance, c->d_slice + frame, frame * 8, cudaMemcpyDeviceToHost, c->stream));
    RSB_CUDA(cudaStreamSynchronize(c->stream));
    return RSB_OK;
}

int rsb_slice_update_frame(uint64_t ctx, int32_t frame_bins, int32_t slice_offset, int32_t frame_is_empty, double* frame_mean,
                           double* frame_variance, int32_t* frame_samples) {
    Context* c = as_ctx(ctx);
    if (!c || !frame_mean || !frame_variance || !frame_samples) return fail(RSB_ERR_ARG, "rsb_slice_update_frame: null argument");
    if (!c->slice.valid || !c->slice.has_bins)
        return fail(RSB_ERR_ARG, "rsb_slice_update_frame: no rendered slice is held (call rsb_render_slice first)");
    const int nx = c->slice.nx, ny = c->slice.ny, sb = c->slice.bins;
    if (slice_offset < 0 || slice_offset + sb > frame_bins)
        return fail(RSB_ERR_ARG, "The slice offset plus the bin count extends beyond the full bin count.");
    if (c->slice.n_pixels == 0) return RSB_OK;
    RSB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t npix = (size_t)nx * ny, elems = npix * (size_t)frame_bins;
    if (c->frame_cap < elems) {
        cudaFree(c->d_frame); cudaFree(c->d_frame_samples);
        c->d_frame = nullptr; c->d_frame_samples = nullptr; c->frame_cap = 0;
        RSB_CUDA(cudaMalloc(&c->d_frame, 2 * elems * sizeof(double)));
        RSB_CUDA(cudaMalloc(&c->d_frame_samples, elems * sizeof(int32_t)));
        c->frame_cap = elems;
    }
    double* d_fm = c->d_frame;
    double* d_fv = c->d_frame + elems;
    // only the slice's bin range of the frame travels: rows of frame_bins elements, slice_bins wide at slice_offset
    const size_t pitch8 = (size_t)frame_bins * 8, width8 = (size_t)sb * 8, off8 = (size_t)slice_offset * 8;
    const size_t pitch4 = (size_t)frame_bins * 4, width4 = (size_t)sb * 4, off4 = (size_t)slice_offset * 4;
    if (frame_is_empty) {
        RSB_CUDA(cudaMemsetAsync(d_fm, 0, 2 * elems * sizeof(double), st));
        RSB_CUDA(cudaMemsetAsync(c->d_frame_samples, 0, elems * sizeof(int32_t), st));
    } else {
        RSB_CUDA(cudaMemcpy2DAsync((char*)d_fm + off8, pitch8, (char*)frame_mean + off8, pitch8, width8, npix, cudaMemcpyHostToDevice, st));
        RSB_CUDA(cudaMemcpy2DAsync((char*)d_fv + off8, pitch8, (char*)frame_variance + off8, pitch8, width8, npix, cudaMemcpyHostToDevice, st));
        RSB_CUDA(cudaMemcpy2DAsync((char*)c->d_frame_samples + off4, pitch4, (char*)frame_samples + off4, pitch4, width4, npix,
                                   cudaMemcpyHostToDevice, st));
    }
    int rc = rsb_frame_combine_dev(ctx, st, (int64_t)npix, frame_bins, slice_offset, sb, c->slice.n_pixels,
                                   c->slice.listed ? c->d_slice_pix : nullptr, ny, c->d_slice, c->d_slice + npix * sb, c->slice.samples, d_fm,
                                   d_fv, c->d_frame_samples);
    if (rc) return rc;
    RSB_CUDA(cudaMemcpy2DAsync((char*)frame_mean + off8, pitch8, (char*)d_fm + off8, pitch8, width8, npix, cudaMemcpyDeviceToHost, st));
    RSB_CUDA(cudaMemcpy2DAsync((char*)frame_variance + off8, pitch8, (char*)d_fv + off8, pitch8, width8, npix, cudaMemcpyDeviceToHost, st));
    RSB_CUDA(cudaMemcpy2DAsync((char*)frame_samples + off4, pitch4, (char*)c->d_frame_samples + off4, pitch4, width4, npix,
                               cudaMemcpyDeviceToHost, st));
    RSB_CUDA(cudaStreamSynchronize(st));
    return RSB_OK;
}

int rsb_render_passes(uint64_t ctx, uint64_t scene, const RsbCamera* camera, const RsbRayConfig* config, const RsbSpectral* spectral,
                      const RsbRngDesc* rng, int32_t n_passes, uint64_t seed_stride, int64_t n_pixels, const int32_t* pixels,
                      double* mean, double* variance, uint64_t* ray_count) {
    Context* c = as_ctx(ctx);
    if (!c || !as_scene(scene) || !camera || !config || !mean || !variance || !ray_count) return fail(RSB_ERR_ARG, "rsb_render: null argument");
    if (!pixels) n_pixels = (int64_t)camera->nx * camera->ny;
    if (n_pixels <= 0) return RSB_OK;
    int rc = rsb_render_slice(ctx, scene, camera, config, spectral, rng, n_passes, seed_stride, n_pixels, pixels, ray_count);
    if (rc) return rc;
    const size_t frame = (size_t)camera->nx * camera->ny * config->bins;
    if (!pixels) return rsb_slice_read(ctx, mean, variance);
    // unlisted pixels must come back untouched: scatter the listed rows into the caller's arrays
    std::vector<double> m(frame), v(frame);
    rc = rsb_slice_read(ctx, m.data(), v.data());
    if (rc) return rc;
    const size_t bins = (size_t)config->bins;
    for (int64_t k = 0; k < n_pixels; ++k) {
        const int x = pixels[2 * k], y = pixels[2 * k + 1];
        if (x < 0 || y < 0 || x >= camera->nx || y >= camera->ny) return fail(RSB_ERR_ARG, "rsb_render: pixel outside the frame");
        const size_t row = ((size_t)x * camera->ny + y) * bins;
        memcpy(mean + row, m.data() + row, bins * 8);
        memcpy(variance + row, v.data() + row, bins * 8);
    }
    return RSB_OK;
}

// Page-lock / release a caller-owned host buffer (cudaHostRegister): the drop-in engine pins the pipeline's frame arrays
// from a helper thread WHILE the device renders, so that rsb_slice_update_frame's copies run at PCIe speed instead of
// through the driver's staging of freshly allocated, never-touched pageable memory.
int rsb_host_pin(uint64_t ctx, void* ptr, int64_t bytes) {
    Context* c = as_ctx(ctx);
    if (!c || !ptr || bytes <= 0) return fail(RSB_ERR_ARG, "rsb_host_pin: bad arguments");
    RSB_CUDA(cudaSetDevice(c->device));
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return RSB_OK; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(RSB_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e)); }
    return RSB_OK;
}

int rsb_host_unpin(uint64_t ctx, void* ptr) {
    Context* c = as_ctx(ctx);
    if (!c || !ptr) return fail(RSB_ERR_ARG, "rsb_host_unpin: bad arguments");
    RSB_CUDA(cudaSetDevice(c->device));
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) cudaGetLastError();      // not registered (any more): nothing to release
    return RSB_OK;
}

// ---- several GPUs in one process -----------------------------------------------------------------------------------
struct Comm {
    std::vector<Context*> ctx;
};

int rsb_comm_create(int32_t n, const uint64_t* ctxs, uint64_t* comm) {
    if (n < 1 || !ctxs || !comm) return fail(RSB_ERR_ARG, "rsb_comm_create: bad arguments");
    Comm* cm = new Comm();
    for (int i = 0; i < n; ++i) {
        Context* c = as_ctx(ctxs[i]);
        if (!c) { delete cm; return fail(RSB_ERR_ARG, "rsb_comm_create: null context"); }
        cm->ctx.push_back(c);
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const int a = cm->ctx[i]->device, b = cm->ctx[j]->device;
            if (a == b) continue;
            int can = 0;
            cudaError_t e = cudaDeviceCanAccessPeer(&can, a, b);
            if (e != cudaSuccess || !can) {
                cudaGetLastError();
                delete cm;
                return fail(RSB_ERR_UNSUPPORTED, "rsb_comm_create: device " + std::to_string(a) + " cannot map the memory of device " + std::to_string(b));
            }
            e = cudaSetDevice(a);
            if (e == cudaSuccess) e = cudaDeviceEnablePeerAccess(b, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) {
                cudaGetLastError();
                delete cm;
                return fail(RSB_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            }
        }
    *comm = reinterpret_cast<uint64_t>(cm);
    return RSB_OK;
}

int rsb_comm_destroy(uint64_t comm) {
    delete reinterpret_cast<Comm*>(comm);
    return RSB_OK;
}

int rsb_comm_gather_slices(uint64_t comm, int32_t root) {
    Comm* cm = reinterpret_cast<Comm*>(comm);
    if (!cm || root < 0 || root >= (int)cm->ctx.size()) return fail(RSB_ERR_ARG, "rsb_comm_gather_slices: bad arguments");
    Context* r = cm->ctx[root];
    if (!r->slice.valid || !r->slice.has_bins) return fail(RSB_ERR_ARG, "rsb_comm_gather_slices: the root holds no rendered slice");
    int64_t total = r->slice.n_pixels;
    for (size_t i = 0; i < cm->ctx.size(); ++i) {
        if ((int)i == root) continue;
        Context* p = cm->ctx[i];
        if (!p->slice.valid || !p->slice.has_bins) return fail(RSB_ERR_ARG, "rsb_comm_gather_slices: a member holds no rendered slice");
        if (p->slice.nx != r->slice.nx || p->slice.ny != r->slice.ny || p->slice.bins != r->slice.bins || p->slice.samples != r->slice.samples)
            return fail(RSB_ERR_ARG, "rsb_comm_gather_slices: the members rendered different frames");
        if (p->slice.n_pixels > 0 && !p->slice.listed)
            return fail(RSB_ERR_ARG, "rsb_comm_gather_slices: every member but the root must have rendered a pixel list (its own tiles)");
        total += p->slice.n_pixels;
    }
    const int64_t frame_pixels = (int64_t)r->slice.nx * r->slice.ny;
    if (total > frame_pixels) return fail(RSB_ERR_ARG, "rsb_comm_gather_slices: the members' pixel lists overlap");
    RSB_CUDA(cudaSetDevice(r->device));
    cudaStream_t st = r->stream;
    const size_t frame = (size_t)frame_pixels * r->slice.bins;
    // the root's task list becomes the union of all lists (or "the whole frame")
    const bool whole = total == frame_pixels;
    int32_t* d_union = nullptr;
    if (!whole && total > r->slice.n_pixels) {
        if (!r->slice.listed && r->slice.n_pixels > 0) return fail(RSB_ERR_ARG, "rsb_comm_gather_slices: the root rendered the whole frame already");
        RSB_CUDA(cudaMalloc(&d_union, (size_t)total * 8));
        if (r->slice.n_pixels > 0) RSB_CUDA(cudaMemcpyAsync(d_union, r->d_slice_pix, (size_t)r->slice.n_pixels * 8, cudaMemcpyDeviceToDevice, st));
    }
    int64_t at = r->slice.n_pixels;
    for (size_t i = 0; i < cm->ctx.size(); ++i) {
        if ((int)i == root) continue;
        Context* p = cm->ctx[i];
        if (p->slice.n_pixels == 0) continue;
        // (every member's render call has synchronised its own stream before it returned)
        const long long n = (long long)p->slice.n_pixels * r->slice.bins;
        k_gather_peer_rows<<<grid_for(r, n, 256, 8), 256, 0, st>>>(p->slice.n_pixels, p->d_slice_pix, r->slice.ny, r->slice.bins, p->d_slice,
                                                                    p->d_slice + frame, r->d_slice, r->d_slice + frame);
        RSB_CUDA(cudaGetLastError());
        if (d_union) RSB_CUDA(cudaMemcpyPeerAsync(d_union + 2 * at, r->device, p->d_slice_pix, p->device, (size_t)p->slice.n_pixels * 8, st));
        at += p->slice.n_pixels;
    }
    RSB_CUDA(cudaStreamSynchronize(st));
    if (d_union) {
        cudaFree(r->d_slice_pix);
        r->d_slice_pix = d_union;
        r->slice_pix_cap = (size_t)total;
        r->slice.listed = true;
    } else if (whole) {
        r->slice.listed = false;
    }
    r->slice.n_pixels = total;
    r->slice.has_xyz = false;     // XYZ work items stay with their owners (rsb_slice_update_xyz_frame on every member)
    return RSB_OK;
}

int rsb_frame_combine_dev(uint64_t ctx, void* cuda_stream, int64_t n_pixels_total, int32_t frame_bins, int32_t slice_offset,
                          int32_t slice_bins, int64_t n_pixels, const int32_t* pixels_dev, int32_t ny, const double* mean_dev,
                          const double* variance_dev, int32_t samples, double* frame_mean_dev, double* frame_variance_dev,
                          int32_t* frame_samples_dev) {
    Context* c = as_ctx(ctx);
    if (!c || !mean_dev || !variance_dev || !frame_mean_dev || !frame_variance_dev || !frame_samples_dev) return fail(RSB_ERR_ARG, "null argument");
    if (samples < 1) return fail(RSB_ERR_ARG, "Number of samples must not be less than 1.");
    if (slice_offset < 0 || slice_bins < 1 || slice_offset + slice_bins > frame_bins)
        return fail(RSB_ERR_ARG, "The slice offset plus the bin count extends beyond the full bin count.");
    if (!pixels_dev) n_pixels = n_pixels_total;
    if (n_pixels <= 0) return RSB_OK;
    RSB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    long long total = (long long)n_pixels * slice_bins;
    k_frame_combine<<<grid_for(c, total, 256, 8), 256, 0, st>>>(n_pixels, pixels_dev, ny, frame_bins, slice_offset, slice_bins, mean_dev,
                                                                 variance_dev, samples, frame_mean_dev, frame_variance_dev, frame_samples_dev);
    RSB_CUDA(cudaGetLastError());
    return RSB_OK;
}

int rsb_render_stats(uint64_t ctx, RsbRenderStats* out) {
    Context* c = as_ctx(ctx);
    if (!c || !out) return fail(RSB_ERR_ARG, "null argument");
    *out = c->render_stats;
    return RSB_OK;
}

int rsb_counters(uint64_t ctx, RsbCounters* out) {
    Context* c = as_ctx(ctx);
    if (!c || !out) return fail(RSB_ERR_ARG, "null argument");
    *out = c->last_counters;
    return RSB_OK;
}

int rsb_last_kernel_ms(uint64_t ctx, float* ms) {
    Context* c = as_ctx(ctx);
    if (!c || !ms) return fail(RSB_ERR_ARG, "null argument");
    *ms = c->last_ms;
    return RSB_OK;
}

}  // extern "C"
