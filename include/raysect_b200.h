/*
 * raysect_b200.h -- C ABI of libraysect_b200.so: Raysect's ray/scene intersection and spectral
 * trace hot path on NVIDIA B200 (sm_100a).
 *
 * The reference (raysect v0.9.1) has no FFI: its extension seams are Python-subclassable Cython
 * classes.  Each entry point below names the reference interface it stands behind; the Python
 * plugin classes that bind them (CudaAccelerator, CudaRenderEngine) are in source_b200/plugin.py
 * and INTEGRATION.md shows the stub a Raysect maintainer would add.
 *
 * Conventions: every function returns 0 on success, non-zero on failure with a message available
 * from rsb_last_error() (thread-local).  Handles are opaque 64-bit integers.  Host arrays passed
 * IN are copied during the call (the caller keeps ownership); host arrays passed OUT are
 * caller-allocated.  Functions with the suffix _dev take DEVICE pointers plus a CUDA stream and
 * return after enqueueing.  One context per device; a context is not thread-safe (neither are the
 * reference's primitives, which cache next_intersection() state).  There is no CPU fallback: every
 * compute entry point fails with RSB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef RAYSECT_B200_H
#define RAYSECT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSB_OK 0
#define RSB_ERR_ARG 1
#define RSB_ERR_CUDA 2
#define RSB_ERR_UNSUPPORTED 3
#define RSB_ERR_OVERFLOW 4

/* primitive rows: raysect/primitive/{sphere,box,cylinder,cone,parabola,csg}.pyx, raysect/primitive/mesh/mesh.pyx */
#define RSB_PRIM_SPHERE 0
#define RSB_PRIM_BOX 1
#define RSB_PRIM_CYLINDER 2
#define RSB_PRIM_CONE 3
#define RSB_PRIM_MESH 4
#define RSB_PRIM_UNION 5
#define RSB_PRIM_INTERSECT 6
#define RSB_PRIM_SUBTRACT 7
#define RSB_PRIM_TORUS (-2)    /* raysect/primitive/torus.pyx; params: major radius, minor radius.  World-level only (a torus has
                                  up to four crossings; the CSG operand interface here carries two) */
#define RSB_PRIM_PARABOLA (-1) /* raysect/primitive/parabola.pyx; params: radius, height.  (Analytic primitives are the
                                  types <= RSB_PRIM_CONE, meshes 4, CSG operators >= 5: the new analytic type takes -1.) */

/* materials: raysect/optical/material/{absorber,lambert,dielectric,conductor}.pyx, emitter/{uniform,unity}.pyx */
#define RSB_MAT_ABSORBER 0
#define RSB_MAT_EMITTER 1
#define RSB_MAT_LAMBERT 2
#define RSB_MAT_DIELECTRIC 3
#define RSB_MAT_CONDUCTOR 4 /* raysect/optical/material/conductor.pyx:39-147 (specular Fresnel conductor) */
#define RSB_MAT_ROUGH_CONDUCTOR 6 /* raysect/optical/material/conductor.pyx:157-344 (GGX microfacets, Smith shadowing,
                                    conductor Fresnel; roughness in RsbSpectral.scale[]) */
#define RSB_MAT_VOLUME_EMITTER 5 /* emitter/uniform.pyx:91-133, emitter/unity.pyx:79-99 on emitter/homogeneous.pyx:40-93
                                    (HomogeneousVolumeEmitter: NullSurface + emission * path length) */

#define RSB_RNG_MT19937_64 0 /* raysect/core/math/random.pyx:99-265, one stream per pixel = seed(seed + y*nx + x) */
#define RSB_RNG_PHILOX 1     /* counter based, keyed on (seed, pixel, sample) */

/* MeshData arrays (raysect/primitive/mesh/mesh.pxd:44-60) + its kd-tree stream (kdtree3d.pyx:864-984) */
typedef struct RsbMeshDesc {
    const float* vertices;       /* [n_vertices][3] */
    const int32_t* triangles;    /* [n_triangles][tri_stride]: v1 v2 v3 [n1 n2 n3] */
    const float* vertex_normals; /* [n_vertex_normals][3] or NULL */
    const float* face_normals;   /* [n_triangles][3] or NULL: computed as mesh.pyx:428-462 does */
    const uint8_t* kdtree;       /* serialised KDTree3DCore stream */
    int64_t kdtree_bytes;
    int32_t n_vertices;
    int32_t n_triangles;
    int32_t tri_stride;          /* 3 or 6 */
    int32_t n_vertex_normals;
    int32_t smoothing;
    int32_t closed;
} RsbMeshDesc;

/*
 * Flattened scenegraph.  Rows [0, n_world) are World.primitives in list order (that order defines the
 * primitive ids stored in the world kd-tree, raysect/core/acceleration/kdtree.pyx:52-55); further rows are
 * CSG operands, whose transforms and boxes are relative to their CSGRoot (raysect/primitive/csg.pyx:70-79).
 */
typedef struct RsbSceneDesc {
    int32_t n_primitives;
    int32_t n_world;
    const int32_t* prim_type;
    const int32_t* prim_material;   /* material row, -1 for CSG operands */
    const int32_t* prim_child_a;    /* CSG operand rows, else -1 */
    const int32_t* prim_child_b;
    const int32_t* prim_mesh;       /* mesh row, else -1 */
    const int32_t* prim_parent;     /* enclosing CSG row, -1 for world-level primitives */
    const double* prim_params;      /* [n][6] sphere r | box lower,upper | cylinder/cone/parabola r,h | torus R,r */
    /* matrices travel as [13]: rows 0..2 of the AffineMatrix3D, then its m33.  The bottom row of an affine matrix is
     * (0, 0, 0, m33); Point3D.transform divides by it (raysect/core/math/point.pyx:272-281), and AffineMatrix3D.inverse()
     * of a non-rigid chain can leave m33 = 1 - 1 ulp, so it is carried instead of assumed. */
    const double* prim_to_local;    /* [n][13] Node.to_local() (parent space -> local) */
    const double* prim_to_root;     /* [n][13] Node.to_root() */
    const double* prim_root_inv;    /* [n][13] to_root().inverse() (Normal3D.transform, normal.pyx:241) */
    const double* prim_bbox;        /* [n][6] Primitive.bounding_box() lower,upper in the parent space */
    const uint8_t* world_kdtree;    /* serialised _PrimitiveKDTree stream */
    int64_t world_kdtree_bytes;
    int32_t n_meshes;
    int32_t n_materials;
    const RsbMeshDesc* meshes;
    const int32_t* mat_type;        /* [n_materials] */
    const int32_t* mat_transmission_only;
    /* ImportanceManager input (raysect/optical/scenegraph/world.pyx:88-132): primitives with material.importance > 0 */
    int32_t n_important;
    int32_t pad;
    const double* imp_sphere;       /* [n_important][4] bounding sphere centre xyz, radius */
    const double* imp_weight;       /* [n_important] */
} RsbSceneDesc;

/* PinholeCamera state after _update_image_geometry (raysect/optical/observer/imaging/pinhole.pyx:148-167), or
 * OrthographicCamera state (imaging/orthographic.pyx:132-137): image_delta = width / nx, rays leave the pixel's
 * jittered position on the z = 0 plane along +z with projection weight 1 (orthographic.pyx:139-167). */
/* CCDArray (imaging/ccd.pyx:40-151): image_delta = width / nx, every sample leaves a uniformly drawn point of its pixel in a
 * cosine-weighted direction over the hemisphere in front of the sensor, projection weight 0.5; sensitivity = pixel area x
 * 2 pi (ccd.pyx:150-151).  Its pixel task draws all its points first and all its directions after them (ccd.pyx:130-131). */
#define RSB_CAMERA_PINHOLE 0
#define RSB_CAMERA_ORTHOGRAPHIC 1
#define RSB_CAMERA_CCD 2
/* VectorCamera (imaging/vector.pyx:44-156): every pixel has its own origin and viewing direction (pixel_origins /
 * pixel_directions, [nx][ny][3] doubles, observer-local).  Pixels off the edge of the image are sub-sampled: per sample two
 * draws (point_square) interpolate the direction between the four diagonal neighbours' with Vector3D.slerp; an edge pixel
 * traces its own direction and draws nothing.  Projection weight 1. */
#define RSB_CAMERA_VECTOR 3
/* Pixel (raysect/optical/observer/nonimaging/pixel.pyx:60-173), a 0-D observer: the (nx, ny) "frame" holds one entry per TASK
 * of the observer (Observer0D._generate_tasks, base/observer.pyx:634-649: pixel_samples split into tasks of samples_per_task =
 * pixel_samples of this descriptor); every task samples the same rectangle at the observer's origin -- image_delta = x_width,
 * image_start_x = y_width -- with cosine-weighted directions, weight 0.5; sensitivity = solid angle x collection area. */
#define RSB_CAMERA_PIXEL 4
typedef struct RsbCamera {
    int32_t nx, ny;
    int32_t pixel_samples;
    int32_t kind;
    double image_delta, image_start_x, image_start_y;
    double sensitivity;
    double to_root[12];       /* rows 0..2 of the observer's to_root() */
    double to_root_w;         /* its m33 (see RsbSceneDesc) */
    const double* pixel_origins;      /* RSB_CAMERA_VECTOR only, else NULL */
    const double* pixel_directions;
} RsbCamera;

/* optical Ray template (raysect/optical/ray.pyx:85-126) for one spectral slice */
typedef struct RsbRayConfig {
    int32_t bins;
    int32_t extinction_min_depth;
    int32_t max_depth;
    int32_t importance_sampling;
    double min_wavelength, max_wavelength;
    double extinction_prob;
    double important_path_weight;
    double max_distance;
} RsbRayConfig;

/* what SpectralFunction.sample(min,max,bins) / .average(min,max) return for each material in this slice */
typedef struct RsbSpectral {
    int32_t bins;
    int32_t n_materials;
    const double* tables;     /* [n_materials][bins] reflectivity | (surface or volume) emission | transmission | conductor index n */
    const double* scale;      /* [n_materials] emitter scale | RoughConductor roughness | Checkerboard scale1 */
    const double* index_in;   /* [n_materials] dielectric index.average() | Checkerboard scale2 */
    const double* index_out;  /* [n_materials] dielectric external_index.average() | Checkerboard 1 / width */
    /* materials that sample TWO spectral functions (Conductor: index n -> row i, extinction k -> row table2[i]):
     * `tables` then holds n_tables >= n_materials rows, the extra rows after the per-material ones.
     * n_tables == 0 means n_materials rows and no second tables (table2 may be NULL). */
    int32_t n_tables;
    int32_t pad;
    const int32_t* table2;    /* [n_materials] row of the material's second table, or -1.  An RSB_MAT_EMITTER with a second
                               * table is a Checkerboard (emitter/checkerboard.pyx:38-146): emission of square two */
} RsbSpectral;

typedef struct RsbRngDesc {
    int32_t mode;
    int32_t pad;
    uint64_t seed;            /* must be >= 1 */
} RsbRngDesc;

/* traversal counters behind the algorithmic-bytes roofline model */
typedef struct RsbCounters {
    uint64_t rays;        /* World.hit queries */
    uint64_t branches;    /* kd branch nodes visited by World.hit (world + mesh trees) */
    uint64_t leaves;      /* kd leaves visited by World.hit */
    uint64_t items;       /* leaf item ids read by World.hit */
    uint64_t prim_tests;  /* BoundPrimitive.hit calls on world-level primitives */
    uint64_t tri_tests;   /* _hit_triangle calls (World.hit and mesh contains rays) */
    uint64_t paths;       /* primary rays traced by rsb_render */
    uint64_t contains;    /* World.contains queries */
    uint64_t table_reads; /* per-bin spectral table rows consumed (surface + volume + emission interactions) */
    uint64_t contains_nodes;      /* kd nodes visited by World.contains point location (+ mesh contains rays) */
    uint64_t contains_items;      /* leaf item ids read by World.contains */
    uint64_t contains_prim_tests; /* BoundPrimitive.contains calls */
} RsbCounters;

const char* rsb_last_error(void);
int rsb_version(void);
void rsb_free(void* p);

/* ---- host-only helpers (no device needed) ---- */

/* KDTree3DCore.__init__ (raysect/core/math/spatial/kdtree3d.pyx:126-459): SAH build over item boxes
 * (boxes[i] = lower xyz, upper xyz; item id = i); returns the serialised stream of save() (:864-912),
 * to be released with rsb_free(). */
int rsb_kdtree_build(const double* boxes, int64_t n_items, int32_t max_depth, int32_t min_items,
                     double hit_cost, double empty_bonus, uint8_t** stream, int64_t* stream_bytes);

/* MeshData._generate_face_normals (raysect/primitive/mesh/mesh.pyx:428-462) */
int rsb_mesh_face_normals(const float* vertices, int32_t n_vertices, const int32_t* triangles,
                          int32_t n_triangles, int32_t tri_stride, float* face_normals);

/* MeshData._generate_bounding_box for every triangle (mesh.pyx:467-504): boxes[i] = lower xyz, upper xyz */
int rsb_mesh_triangle_boxes(const float* vertices, int32_t n_vertices, const int32_t* triangles,
                            int32_t n_triangles, int32_t tri_stride, double* boxes);

/* ---- device ---- */

int rsb_context_create(int device, uint64_t* ctx);
int rsb_context_destroy(uint64_t ctx);
int rsb_device_info(uint64_t ctx, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, uint64_t* total_mem);

/* Accelerator.build (raysect/core/acceleration/accelerator.pxd:37-41, kdtree.pyx:167-168) */
int rsb_scene_create(uint64_t ctx, const RsbSceneDesc* desc, uint64_t* scene);
int rsb_scene_destroy(uint64_t ctx, uint64_t scene);

/*
 * Accelerator.hit == World.hit (raysect/core/scenegraph/world.pyx:125-146) over a batch of rays.
 * origins, directions: [n][3]; max_distance: [n] or NULL (= +inf).
 * out_prim[n]: world-level primitive id or -1 (miss); out_t[n]: Intersection.ray_distance;
 * out_sub[n]: triangle id (mesh) | face code (analytic) ; out_flags[n]: bit0 = Intersection.exiting;
 * out_node[n][2]: (world kd leaf id, mesh kd leaf id or -1) using the reference's node numbering;
 * out_geom[n][12] or NULL: hit_point, inside_point, outside_point, normal in primitive-local space;
 * out_uvw[n][3] or NULL: MeshIntersection.u,v,w.
 */
int rsb_hit_batch(uint64_t ctx, uint64_t scene, int64_t n, const double* origins, const double* directions,
                  const double* max_distance, int32_t* out_prim, double* out_t, int32_t* out_sub,
                  uint8_t* out_flags, int32_t* out_node, double* out_geom, float* out_uvw);
int rsb_hit_batch_dev(uint64_t ctx, uint64_t scene, void* cuda_stream, int64_t n, const double* origins,
                      const double* directions, const double* max_distance, int32_t* out_prim, double* out_t,
                      int32_t* out_sub, uint8_t* out_flags, int32_t* out_node, double* out_geom, float* out_uvw,
                      int32_t count);

/*
 * Ray-batch sweep (BASELINE config 5): n primary rays generated ON DEVICE from (seed, index), from
 * `origin` toward a jittered window centred on `target` (half-width `half_window` in x and y),
 * hit results reduced to (hits, sum of t, xor of primitive ids) so that 1e9 rays need no host ray arrays
 * (the rays live in a device-side chunk buffer of RSB_RQ_CHUNK queries).
 * order_log2 = 0: ray `index` aims at a point drawn uniformly over the whole window (consecutive rays are
 * unrelated: the incoherent case).  order_log2 = g > 0: the window is a 2^g x 2^g grid of cells walked along
 * the Morton curve and ray `index` jitters inside cell (index mod 4^g) -- consecutive rays are neighbouring
 * cells, the way an observer hands out the pixels of an image (the coherent, primary-ray case).
 */
int rsb_hit_sweep_dev(uint64_t ctx, uint64_t scene, void* cuda_stream, int64_t n, int64_t first_index, uint64_t seed,
                      const double* origin, const double* target, double half_window, int32_t order_log2,
                      uint64_t* out_hits_dev, double* out_sum_t_dev, uint64_t* out_xor_prim_dev, int32_t count);

/*
 * Incoherent query batches (rays after a diffuse bounce, a user's unordered batch) put 32 unrelated kd walks into every
 * warp.  With reordering on, rsb_hit_batch(_dev) and rsb_hit_sweep_dev sort each pipeline pass (up to 64 Mi queries) on a
 * 20-bit coherence key first -- origin cell and octahedral direction cell, normalised to the extent the batch covers,
 * Morton-interleaved; one counting sort on the device -- traverse the permuted copy and write every answer back at the
 * caller's index.  The answers are the same (queries are independent; the reference's World.hit has no notion of order).
 * On by default (measured on a B200: 10,000 spheres 943 -> 1,732 Mrays/s, 1.3 M triangles 631 -> 862 on independent random
 * rays; an already coherent batch pays the key + scatter passes for nothing, about 8 %: turn it off for those);
 * RSB_RQ_REORDER=0 turns it off at context creation.  A sweep generated along the Morton curve (order_log2 > 0) is never
 * sorted.
 */
int rsb_set_query_reorder(uint64_t ctx, int32_t on);

/*
 * Accelerator.contains == World.contains (world.pyx:148-168): points [n][3]; out_count[n];
 * out_prims[n][cap] primitive ids in the reference's list order (kd leaf order).
 */
int rsb_contains_batch(uint64_t ctx, uint64_t scene, int64_t n, const double* points, int32_t cap,
                       int32_t* out_count, int32_t* out_prims);

/* raysect.core.math.random: seed(seed) then n x uniform() (random.pyx:215-265), evaluated on the device */
int rsb_rng_uniform(uint64_t ctx, uint64_t seed, int64_t n, double* out);

/*
 * RenderEngine.run for a PinholeCamera + SpectralPowerPipeline2D slice
 * (raysect/core/workflow.py:78-91 ; observer.pyx:363-419 ; pipeline/spectral/power.pyx:468-486):
 * renders pixel_samples samples for each listed pixel (pixels = [n_pixels][2] (x, y), or NULL for the
 * whole nx*ny frame) and writes, for pixel (x, y) and slice bin b, frame index (x*ny + y)*bins + b:
 *   mean, variance = StatsArray1D after pixel_samples x add_sample(); unlisted pixels are untouched.
 * ray_count += the reference's ray counter (primary rays + daughters spawned).
 */
int rsb_render(uint64_t ctx, uint64_t scene, const RsbCamera* camera, const RsbRayConfig* config,
               const RsbSpectral* spectral, const RsbRngDesc* rng, int64_t n_pixels, const int32_t* pixels,
               double* mean, double* variance, uint64_t* ray_count);
int rsb_render_dev(uint64_t ctx, uint64_t scene, void* cuda_stream, const RsbCamera* camera,
                   const RsbRayConfig* config, const RsbSpectral* spectral, const RsbRngDesc* rng,
                   int64_t n_pixels, const int32_t* pixels_dev, double* mean_dev, double* variance_dev,
                   uint64_t* ray_count_dev, int32_t count);

/*
 * n_passes accumulated observe() calls of the same camera rendered CONCURRENTLY (the reference's usage pattern
 * for progressive renders, demos/cornell_box.py:166-174: observe() in a loop with pipeline.accumulate = True):
 * pass p renders camera->pixel_samples samples per listed pixel from the streams seeded
 * rng->seed + p*seed_stride + y*nx + x, and the passes are merged in order p = 0, 1, ... with
 * SpectralPowerPipeline2D.update -> StatsArray3D.combine_samples (power.pyx:424-437, statsarray.pyx:612-670,
 * 780-857), starting from an EMPTY frame.  mean/variance of the listed pixels come back holding
 * n_passes*pixel_samples samples.  A pixel's samples within one pass are sequential (they share a stream);
 * passes are independent streams, so they multiply the number of pixel streams in flight by n_passes.
 * n_passes = 1, seed_stride = 0 is rsb_render / rsb_render_dev (no merge step: the frame is the pass).
 */
int rsb_render_passes(uint64_t ctx, uint64_t scene, const RsbCamera* camera, const RsbRayConfig* config,
                      const RsbSpectral* spectral, const RsbRngDesc* rng, int32_t n_passes, uint64_t seed_stride,
                      int64_t n_pixels, const int32_t* pixels, double* mean, double* variance, uint64_t* ray_count);
int rsb_render_passes_dev(uint64_t ctx, uint64_t scene, void* cuda_stream, const RsbCamera* camera,
                          const RsbRayConfig* config, const RsbSpectral* spectral, const RsbRngDesc* rng,
                          int32_t n_passes, uint64_t seed_stride, int64_t n_pixels, const int32_t* pixels_dev,
                          double* mean_dev, double* variance_dev, uint64_t* ray_count_dev, int32_t count);

/*
 * RenderEngine.run + Pipeline.update for a whole slice, the form the drop-in engine uses (workflow.py:78-91,
 * observer.pyx:299-305, power.pyx:424-437): rsb_render_slice renders the listed pixels like rsb_render_passes but
 * KEEPS the result on the device; rsb_slice_update_frame then merges it into a pipeline's HOST frame arrays
 * ((nx, ny, frame_bins) mean / variance f64, samples i32 -- StatsArray3D's own buffers) at bin offset slice_offset with
 * StatsArray3D.combine_samples (samples = n_passes * pixel_samples), for the listed pixels only: the slice's bin range of
 * the frame goes host -> device (skipped when frame_is_empty: a freshly initialised frame), is combined on the device
 * and comes back.  Several pipelines can be updated from one rendered slice.  rsb_slice_read copies the raw slice
 * (mean, variance; (nx, ny, slice_bins); unlisted pixels zero) to the host.  All buffers are owned by the context and
 * reused from call to call.
 */
int rsb_render_slice(uint64_t ctx, uint64_t scene, const RsbCamera* camera, const RsbRayConfig* config,
                     const RsbSpectral* spectral, const RsbRngDesc* rng, int32_t n_passes, uint64_t seed_stride,
                     int64_t n_pixels, const int32_t* pixels, uint64_t* ray_count);

/*
 * Every spectral slice of an observe() at once (observer.pyx:299-340: the reference calls RenderEngine.run once per
 * slice, one after the other; with spectral_rays = 512 that is 512 renders of a frame).  Slices are independent pixel
 * streams exactly like accumulated passes, so they are rendered CONCURRENTLY: work item (pass p, slice k, pixel) draws
 * from the streams seeded rng->seed + (p*n_slices + k)*seed_stride + y*nx + x (the engine's per-slice seeds when
 * seed_stride = nx*ny), is shaded with slice k's material rows / tables (spectral[k]; all slices have config->bins bins,
 * the same materials and tables) and lands in bins [k*bins, (k+1)*bins) of a frame with n_slices*bins bins per pixel.
 * rsb_render_slices keeps that frame on the device as the held slice (slice_bins = n_slices*bins) for
 * rsb_slice_update_frame / rsb_slice_read; the _dev form writes mean_dev / variance_dev [(nx, ny, n_slices*bins)].
 */
int rsb_render_slices(uint64_t ctx, uint64_t scene, const RsbCamera* camera, const RsbRayConfig* config,
                      const RsbSpectral* spectral, const RsbRngDesc* rng, int32_t n_passes, int32_t n_slices,
                      uint64_t seed_stride, int64_t n_pixels, const int32_t* pixels, uint64_t* ray_count);
int rsb_render_slices_dev(uint64_t ctx, uint64_t scene, void* cuda_stream, const RsbCamera* camera,
                          const RsbRayConfig* config, const RsbSpectral* spectral, const RsbRngDesc* rng,
                          int32_t n_passes, int32_t n_slices, uint64_t seed_stride, int64_t n_pixels,
                          const int32_t* pixels_dev, double* mean_dev, double* variance_dev, uint64_t* ray_count_dev,
                          int32_t count);
int rsb_slice_read(uint64_t ctx, double* mean, double* variance);
int rsb_slice_update_frame(uint64_t ctx, int32_t frame_bins, int32_t slice_offset, int32_t frame_is_empty,
                           double* frame_mean, double* frame_variance, int32_t* frame_samples);
/*
 * RGBPipeline2D on the device (raysect/optical/observer/pipeline/rgb.pyx:216-290, XYZPixelProcessor rgb.pyx:534-562,
 * spectrum_to_ciexyz raysect/optical/colour.pyx:158-186).  rsb_render_slices_xyz renders like rsb_render_slices and, for
 * every sample, also projects the sample's spectrum on the CIE curves -- X = sum over the bins, in index order, of
 * delta_wavelength[k] * sample[i] * resampled_xyz[k][i][0], likewise Y and Z -- and keeps Welford statistics of
 * X, Y, Z times the camera's sensitivity per work item (pass, slice, pixel): what XYZPixelProcessor.add_sample /
 * pack_results produce per pixel task.  resampled_xyz is [n_slices][bins][3] (colour.resample_ciexyz of every slice's
 * wavelength range), delta_wavelength [n_slices] (Spectrum.delta_wavelength of the slice).  keep_spectral = 0 drops the
 * per-bin statistics altogether (an observer with RGB pipelines only: no (nx, ny, bins) frame exists anywhere);
 * otherwise the spectral frame is held exactly as after rsb_render_slices.
 * rsb_slice_update_xyz_frame is RGBPipeline2D.update + finalise for the listed pixels: per pass, the slices' means and
 * variances are summed in slice order (rgb.pyx:259-265) and merged into the HOST xyz_frame arrays ((nx, ny, 3) mean /
 * variance f64, samples i32) with StatsArray3D.combine_samples(samples = pixel_samples), pass after pass.
 */
int rsb_render_slices_xyz(uint64_t ctx, uint64_t scene, const RsbCamera* camera, const RsbRayConfig* config,
                          const RsbSpectral* spectral, const RsbRngDesc* rng, int32_t n_passes, int32_t n_slices,
                          uint64_t seed_stride, int64_t n_pixels, const int32_t* pixels, const double* resampled_xyz,
                          const double* delta_wavelength, int32_t keep_spectral, uint64_t* ray_count);
int rsb_slice_update_xyz_frame(uint64_t ctx, int32_t frame_is_empty, double* xyz_mean, double* xyz_variance,
                               int32_t* xyz_samples);
/*
 * The general form: up to RSB_PROJ_MAX projection channels per render, each summing one curve times the sample's spectrum
 * over the bins in the operation order of the pixel processor it stands for, so that one render can feed an RGBPipeline2D
 * (3 channels RSB_PROJ_XYZ: delta * sample * curve, the sum times the sensitivity), PowerPipeline2D's (one channel
 * RSB_PROJ_POWER each: sample * filter * sensitivity * delta, raysect/optical/observer/pipeline/mono/power.pyx:768-779) and
 * RadiancePipeline2D's (RSB_PROJ_RADIANCE: sample * filter * delta, mono/radiance.pyx:184-195) side by side.  curves is
 * [n_slices][bins][n_channels].  rsb_slice_update_proj_frame merges channels [channel0, channel0 + n_channels) into HOST
 * frame arrays with n_channels values per pixel ((nx, ny, 3) xyz_frame; (nx, ny) StatsArray2D of the mono pipelines) as
 * the pipelines' update + finalise do (slices of a pass summed, passes merged with combine_samples; mono/power.pyx:516-556).
 */
#define RSB_PROJ_XYZ 0
#define RSB_PROJ_POWER 1
#define RSB_PROJ_RADIANCE 2
#define RSB_PROJ_MAX 8
int rsb_render_slices_proj(uint64_t ctx, uint64_t scene, const RsbCamera* camera, const RsbRayConfig* config,
                           const RsbSpectral* spectral, const RsbRngDesc* rng, int32_t n_passes, int32_t n_slices,
                           uint64_t seed_stride, int64_t n_pixels, const int32_t* pixels, int32_t n_channels,
                           const int32_t* channel_mode, const double* curves, const double* delta_wavelength,
                           int32_t keep_spectral, uint64_t* ray_count);
int rsb_slice_update_proj_frame(uint64_t ctx, int32_t channel0, int32_t n_channels, int32_t frame_is_empty,
                                double* frame_mean, double* frame_variance, int32_t* frame_samples);
/* BayerPipeline2D (raysect/optical/observer/pipeline/bayer.pyx:307-387): channels channel0 .. channel0 + 2 are its red, green
 * and blue filters (RSB_PROJ_POWER); pixel (x, y) of the (nx, ny) frame takes the channel its mosaic position selects,
 * (0, 1, 1, 2)[(x % 2) + 2 (y % 2)] (bayer.pyx:109, 345-347). */
int rsb_slice_update_bayer_frame(uint64_t ctx, int32_t channel0, int32_t frame_is_empty, double* frame_mean,
                                 double* frame_variance, int32_t* frame_samples);
/*
 * Several GPUs driven from ONE process (the reference's user runs one Python interpreter; its MulticoreEngine forks
 * workers and pickles per-pixel results back, raysect/core/workflow.py:123-327).  A communicator joins contexts on
 * different devices and enables peer access between them.  After every member has rendered ITS pixel list of the same
 * frame (rsb_render_slice(s), one host thread per context), rsb_comm_gather_slices makes the root's held slice the whole
 * result: one kernel per peer on the root's device loads the rows of that peer's listed pixels straight out of the peer's
 * memory (NVLink / NVSwitch peer mapping) and stores them in place -- copies only, exact -- and the root's task list
 * becomes the union of all lists, ready for ONE rsb_slice_update_frame.  Pixel lists must be disjoint.  XYZ statistics
 * (rsb_render_slices_xyz) are per work item and stay with their owners: call rsb_slice_update_xyz_frame on every member
 * BEFORE the gather (the xyz_frame is 3 values per pixel).  One process per GPU (torch.distributed / NCCL,
 * source_b200/distributed.py) remains the form bench.py scales with.
 */
int rsb_comm_create(int32_t n_contexts, const uint64_t* contexts, uint64_t* comm);
int rsb_comm_destroy(uint64_t comm);
int rsb_comm_gather_slices(uint64_t comm, int32_t root);
/* Page-lock / release a caller-owned host buffer (cudaHostRegister / cudaHostUnregister).  The drop-in engine pins the
 * pipeline's StatsArray3D buffers from a helper thread while the device renders, so that rsb_slice_update_frame's
 * copies do not crawl through freshly allocated, never-touched pageable memory.  Pinning an already pinned buffer
 * and releasing an unpinned one are no-ops. */
int rsb_host_pin(uint64_t ctx, void* ptr, int64_t bytes);
int rsb_host_unpin(uint64_t ctx, void* ptr);

/*
 * SpectralPowerPipeline2D.update -> StatsArray3D.combine_samples (power.pyx:424-437, statsarray.pyx:780-857):
 * merges a freshly rendered slice (mean, variance, samples_per_pixel; [n_pixels_total][slice_bins]) into an
 * accumulating frame (frame_* [n_pixels_total][frame_bins]) at bin offset slice_offset, for the listed pixels.
 */
int rsb_frame_combine_dev(uint64_t ctx, void* cuda_stream, int64_t n_pixels_total, int32_t frame_bins,
                          int32_t slice_offset, int32_t slice_bins, int64_t n_pixels, const int32_t* pixels_dev,
                          int32_t ny, const double* mean_dev, const double* variance_dev, int32_t samples,
                          double* frame_mean_dev, double* frame_variance_dev, int32_t* frame_samples_dev);

/* shape of the last rsb_render / rsb_render_dev call */
typedef struct RsbRenderStats {
    int64_t slots;            /* pixel streams in flight (wavefront width) */
    int64_t waves;            /* trace -> shade -> finalize -> regen rounds */
    int64_t launches;         /* kernels launched */
    int64_t trace_launches;   /* trace phases run (k_wf_trace, or walk / Mesh.hit / resume kernels for scenes with meshes) */
    double trace_ms;          /* summed device time of the trace phase (CUDA events on the launch stream); 0 unless
                                 the call was made with RSB_RENDER_TIME_TRACE */
    double shade_ms;          /* the same for k_wf_shade, k_wf_finalize and k_wf_regen */
    double finalize_ms;
    double regen_ms;
} RsbRenderStats;
int rsb_render_stats(uint64_t ctx, RsbRenderStats* out);

#define RSB_RENDER_COUNT 1        /* `count` argument of rsb_render_dev: collect traversal counters */
#define RSB_RENDER_TIME_TRACE 2   /* bracket the four phases of every wave with CUDA events */

/* counters of the last *_dev / host call made with count != 0 (host calls always count) */
int rsb_counters(uint64_t ctx, RsbCounters* out);
/* device time (ms, CUDA events on the launch stream) of the dominant kernel of the last host-buffer call */
int rsb_last_kernel_ms(uint64_t ctx, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* RAYSECT_B200_H */
