// rsb_path.h -- the per-ray spectral trace loop as an iterative path.
//
// The reference recurses: Ray.trace -> material.evaluate_surface -> daughter.trace -> ... and
// multiplies the returned Spectrum (float64[bins]) on the way back up.  Every material on this
// path spawns at most one daughter, so the recursion is a chain: RNG draws and geometry on the
// way down, per-bin multiplies on the way up.  Here the way down appends the multiplicative
// factors of each segment to a per-path LOG (16-B entries in HBM); only when a path ends on an
// emitter is the log replayed backwards over the bins, applying the SAME multiplies in the SAME
// order as the reference's unwind.  Paths that end in a miss / absorber / roulette contribute an
// all-zero spectrum and skip the replay.
#pragma once
#include "rsb_geom.h"
#include "rsb_rng.h"

namespace rsb {

#ifndef RSB_PI
#define RSB_PI 3.14159265358979323846
#endif
#define RSB_1_PI 0.31830988618379067154

struct RayConfig {   // optical Ray template fields, raysect/optical/ray.pyx:85-126
    int32_t bins;
    int32_t extinction_min_depth;
    int32_t max_depth;
    int32_t importance_sampling;
    double min_wavelength, max_wavelength;
    double extinction_prob;
    double important_path_weight;
    double max_distance;
};

struct Spectral {    // per spectral slice: what SpectralFunction.sample()/average() cache on the host
    const Material* mats;
    const double* tables;   // [n_tables][bins]: one row per material, then the second tables (conductor extinction)
    const double* tables_ln;   // [n_tables][bins] natural log of `tables` (device replay only), or null
    int32_t bins;
    int32_t n_materials;
    int32_t n_tables;
};

struct Camera {      // PinholeCamera, raysect/optical/observer/imaging/pinhole.pyx:148-207 | OrthographicCamera (kind 1) | CCDArray (kind 2) | VectorCamera (kind 3) | Pixel (kind 4)
    int32_t nx, ny;
    int32_t pixel_samples;
    int32_t kind;
    double image_delta, image_start_x, image_start_y;
    double sensitivity;
    double to_root[RSB_MAT_WORDS];   // rows 0..2, then 1 / m33
    const double* pixel_origins;     // VectorCamera (kind 3): [nx][ny][3] per-pixel origin and viewing direction
    const double* pixel_directions;
};

enum LogOp : int32_t {
    LOG_MULS = 0,   // s *= v                      Spectrum.mul_scalar / div_scalar (spectrum.pyx:449-468)
    LOG_MULA = 1,   // s *= table[i]               Spectrum.mul_array (spectrum.pyx:491)
    LOG_POWA = 2,   // s *= pow(table[i], v)       Dielectric.evaluate_volume (dielectric.pyx:313-330)
    LOG_EMIT = 3,   // s  = table[i] * v           UniformSurfaceEmitter.evaluate_surface (uniform.pyx:67-81)
    LOG_ADDA = 5,   // s += (0 + table[i] * scale) * v   HomogeneousVolumeEmitter.evaluate_volume (homogeneous.pyx:78-91) with
                    //                              UniformVolumeEmitter.emission_function (uniform.pyx:129-131); v = length
    LOG_FRESNEL = 4,   // s *= fresnel(v, n[i], k[i])  Conductor.evaluate_surface (conductor.pyx:122-128); n = table,
                       //                              k = the row in bits 8.. of the op word
};

struct __attribute__((aligned(16))) LogEntry {
    int32_t op;
    int32_t table;
    double v;
};

struct PathLog {
    LogEntry* base;
    size_t stride;      // entries are interleaved across threads: entry k at base[k*stride]
    int32_t n;
    int32_t capacity;
    int32_t overflow;
    int32_t additive;   // a LOG_ADDA entry was pushed: the path's spectrum is not identically zero even if it ends dark
    RSB_HD void push(int op, int table, double v) {
        if (op == LOG_ADDA) additive = 1;
        if (n < capacity) {
            LogEntry e; e.op = op; e.table = table; e.v = v;
            base[(size_t)n * stride] = e;
            ++n;
        } else {
            overflow = 1;
        }
    }
    RSB_HD LogEntry get(int k) const { return base[(size_t)k * stride]; }
};

// ----- random direction generators (raysect/core/math/random.pyx:375-445, sampler/solidangle.pyx:223-237)
RSB_HD double max0(double x) { return x > 0 ? x : 0.0; }   // Cython max(0, x)

RSB_HD V3 vector_sphere(Rng& rng) {
    double z = 1.0 - 2.0 * rng.uniform();
    double r = sqrt(max0(1.0 - z * z));
    double phi = 2.0 * RSB_PI * rng.uniform();
    return v3(r * cos(phi), r * sin(phi), z);
}

RSB_HD V3 vector_cone_uniform(Rng& rng, double theta) {
    theta *= 0.017453292519943295;
    double phi = 2.0 * RSB_PI * rng.uniform();
    double cos_theta = cos(theta);
    double z = rng.uniform() * (1 - cos_theta) + cos_theta;
    double r = sqrt(max0(1.0 - z * z));
    return v3(r * cos(phi), r * sin(phi), z);
}

// Checkerboard._flip (emitter/checkerboard.pyx:129-146): the coordinate rounded to the nanometre, its cell parity
// (C fmod of |rwidth * p| by 2), inverted for negative coordinates
RSB_HD bool checkerboard_flip(bool v, double p, double rwidth) {
    p = round(p * 1e9) / 1e9;
    if (fmod(fabs(rwidth * p), 2.0) >= 1.0) v = !v;
    if (p < 0) v = !v;
    return v;
}

// HemisphereCosineSampler.sample (core/math/sampler/solidangle.pyx:228-233) for the two draws it takes, in draw order
RSB_HD V3 hemisphere_cosine_from(double u1, double u2) {
    double r = sqrt(u1);
    double phi = 2.0 * RSB_PI * u2;
    double x = r * cos(phi);
    double y = r * sin(phi);
    return v3(x, y, sqrt(max0(1.0 - x * x - y * y)));
}

RSB_HD V3 hemisphere_cosine_sample(Rng& rng) {
    double u1 = rng.uniform();
    double u2 = rng.uniform();
    return hemisphere_cosine_from(u1, u2);
}

// ----- RoughConductor (raysect/optical/material/conductor.pyx:157-344): GGX facet distribution, Smith shadowing
RSB_HD double ggx_d(const V3& s_half, double roughness) {          // _d, conductor.pyx:292-300
    double r2 = roughness * roughness;
    double h2 = s_half.z * s_half.z;
    double k = h2 * (r2 - 1) + 1;
    return r2 / (RSB_PI * k * k);
}

RSB_HD double ggx_g1(const V3& v, double roughness) {              // _g1, conductor.pyx:307-310
    double r2 = roughness * roughness;
    return 2 * v.z / (v.z + sqrt(r2 + (1 - r2) * (v.z * v.z)));
}

RSB_HD double ggx_pdf(const V3& s_incoming, const V3& s_outgoing, double roughness) {   // pdf, conductor.pyx:203-220
    V3 s_half = v3(s_incoming.x + s_outgoing.x, s_incoming.y + s_outgoing.y, s_incoming.z + s_outgoing.z);
    if (length(s_half) == 0.0) return 0.0;
    s_half = normalise(s_half);
    return 0.25 * ggx_d(s_half, roughness) * fabs(s_half.z / dot(s_outgoing, s_half));
}

RSB_HD V3 ggx_sample(Rng& rng, const V3& s_incoming, double roughness) {               // sample, conductor.pyx:222-247
    double e1 = rng.uniform();
    double e2 = rng.uniform();
    double theta = atan(roughness * sqrt(e1) / sqrt(1 - e1));
    double phi = 2 * RSB_PI * e2;
    V3 facet_normal = v3(cos(phi) * sin(theta), sin(phi) * sin(theta), cos(theta));
    double temp = 2 * dot(s_incoming, facet_normal);
    return v3(temp * facet_normal.x - s_incoming.x, temp * facet_normal.y - s_incoming.y, temp * facet_normal.z - s_incoming.z);
}

RSB_HD double hemisphere_cosine_pdf(const V3& s) {
    if (s.z >= 0.0) return RSB_1_PI * s.z;
    return 0.0;
}

// ----- ImportanceManager (raysect/optical/scenegraph/world.pyx:134-253)
RSB_HD V3 important_direction_sample(const Scene& sc, Rng& rng, const V3& origin) {
    int index = find_index(sc.imp_cdf, sc.n_important, rng.uniform()) + 1;
    if (index >= sc.n_important) index = sc.n_important - 1;
    const double* s = sc.imp_sphere + 4 * index;
    V3 direction = v3(s[0] - origin.x, s[1] - origin.y, s[2] - origin.z);
    double distance = length(direction);
    double radius = s[3];
    if (distance == 0 || distance < radius) return vector_sphere(rng);
    double angular_radius = asin(radius / distance);
    V3 sample = vector_cone_uniform(rng, angular_radius * 180 / RSB_PI);
    direction = normalise(direction);
    // rotate_basis(direction, direction.orthogonal()) -- the cdef version in
    // raysect/core/math/cython/transform.pyx:45-70 (no re-normalisation: right = up x forward,
    // columns right | up | forward), NOT the public raysect.core.math.transform.rotate_basis
    V3 up = orthogonal(direction);
    V3 right = cross(up, direction);
    return v3(right.x * sample.x + up.x * sample.y + direction.x * sample.z,
              right.y * sample.x + up.y * sample.y + direction.y * sample.z,
              right.z * sample.x + up.z * sample.y + direction.z * sample.z);
}

RSB_HD double important_direction_pdf(const Scene& sc, const V3& origin, const V3& direction) {
    double pdf_all = 0;
    for (int i = 0; i < sc.n_important; ++i) {
        const double* s = sc.imp_sphere + 4 * i;
        V3 cone_axis = v3(s[0] - origin.x, s[1] - origin.y, s[2] - origin.z);
        double distance = length(cone_axis);
        double radius = s[3];
        double solid_angle;
        if (distance == 0 || distance < radius) {
            solid_angle = 4 * RSB_PI;
        } else {
            double t = radius / distance;
            double angular_radius_cos = sqrt(1 - t * t);
            cone_axis = normalise(cone_axis);
            if (dot(direction, cone_axis) < angular_radius_cos) continue;
            solid_angle = 2 * RSB_PI * (1 - angular_radius_cos);
        }
        double pdf_sphere = 1 / solid_angle;
        double selection_weight = sc.imp_weight[i] / sc.imp_total;
        pdf_all += selection_weight * pdf_sphere;
    }
    return pdf_all;
}

// ----- volumes: Ray._sample_volumes (raysect/optical/ray.pyx:422-455) ---------------------------
// world.contains(ray.origin) -> the one world leaf holding the origin, items in leaf order; each
// containing primitive's material.evaluate_volume(start = world hit point, end = ray origin).
// Entries are pushed in REVERSE so that the backward replay applies them in forward order.
template <int FEAT = RSB_FEAT_ALL, class Stats>
RSB_HD void log_volumes(const Scene& sc, const Spectral& sp, const V3& origin, const V3& w_hit,
                        KdStackEntry* stack, PathLog& log, Stats& stats) {
    int offset, count;
    if (!kd_locate(sc.world, origin, &offset, &count, stats)) return;
    for (int i = count - 1; i >= 0; --i) {
        int id = sc.world.items[offset + i];
        const Material& m = sp.mats[sc.prims[id].material];
        // (conductors, volume emitters: only in the full-featured instantiation, see RSB_FEAT_RARE_MATERIALS)
        const bool emitter = (FEAT & RSB_FEAT_RARE_MATERIALS) && m.type == MAT_VOLUME_EMITTER;
        if (m.type != MAT_DIELECTRIC && !emitter) continue;   // the others: evaluate_volume is the identity
        stats.prim_test();
        if (!prim_contains<FEAT>(sc, id, origin, stack, stats)) continue;
        stats.table_read();
        if (emitter) {
            // HomogeneousVolumeEmitter.evaluate_volume (homogeneous.pyx:66-91): the integration length is measured in
            // the CONTAINING primitive's local space, end -> start; a zero length contributes nothing
            const double* w2p = sc.prims[id].to_local;
            V3 s = xform_point(w2p, w_hit), e = xform_point(w2p, origin);
            double len = length(v3(s.x - e.x, s.y - e.y, s.z - e.z));
            if (len == 0) continue;
            log.push(LOG_ADDA, m.table, len);
            continue;
        }
        // length = start_point.vector_to(end_point).get_length()
        V3 v = v3(origin.x - w_hit.x, origin.y - w_hit.y, origin.z - w_hit.z);
        log.push(LOG_POWA, m.table, length(v));
    }
}

enum PathEnd : int32_t { PATH_ZERO = 0, PATH_EMITTED = 1, PATH_CONTINUE = 2 };

struct PathState {
    V3 o, d;            // world-space ray of the current segment
    int32_t depth;
    uint32_t rays;      // the reference's ray counter for this primary ray (1 + daughters spawned)
    int32_t keep_alive; // the segment was spawned by a NullSurface: trace(world, keep_alive=True), no roulette (material.pyx:147)
};

RSB_HD void path_begin(PathState& ps, PathLog& log, const V3& o, const V3& d) {
    ps.o = o;
    ps.d = d;
    ps.depth = 0;
    ps.rays = 1;
    ps.keep_alive = 0;
    log.n = 0;
    log.additive = 0;
}

// One segment of the Ray.trace recursion (raysect/optical/ray.pyx:338-401) in two stages, which the
// wavefront kernels run as separate launches and the serial harness back to back:
//   path_trace  roulette + World.hit                         (ray.pyx:380-393)
//   path_shade  material.evaluate_surface up to the point where it would call daughter.trace(),
//               then _sample_volumes and the roulette normalisation (ray.pyx:395-401)
// PATH_CONTINUE: ps holds the daughter ray.  PATH_EMITTED: the log now ends with a LOG_EMIT entry.
// PATH_ZERO: the path's spectrum is identically zero.
// Russian roulette of a segment that is about to be traced (ray.pyx:380-388).  false: the path ends here with an
// all-zero spectrum (roulette, or ray.max_depth reached -- no draw in that case).
RSB_HD bool path_roulette(const RayConfig& cfg, int depth, Rng& rng, double* normalisation) {
    if (depth < cfg.extinction_min_depth) {
        *normalisation = 1.0;
        return true;
    }
    if (depth >= cfg.max_depth || rng.probability(cfg.extinction_prob)) return false;
    *normalisation = 1 / (1 - cfg.extinction_prob);
    return true;
}

template <int FEAT = RSB_FEAT_ALL, int S = 1, class Stats>
RSB_HD int path_trace(const Scene& sc, const RayConfig& cfg, const PathState& ps, Rng& rng, KdStackEntry* stack,
                      HitRec* rec, double* normalisation, Stats& stats, double* axbuf = nullptr) {
    if (ps.keep_alive) *normalisation = 1.0;     // ray.pyx:382: "if keep_alive or self.depth < self._extinction_min_depth"
    else if (!path_roulette(cfg, ps.depth, rng, normalisation)) return PATH_ZERO;
    // -- closest hit (ray.pyx:391-393)
    if (S == 1) {
        if (!world_hit<FEAT>(sc, ps.o, ps.d, cfg.max_distance, stack, rec, stats)) return PATH_ZERO;
    } else {
        if (!world_hit_ax<FEAT, S>(sc, ps.o, ps.d, cfg.max_distance, stack, rec, stats, axbuf)) return PATH_ZERO;
    }
    return PATH_CONTINUE;
}

// MATSEL: -1 = dispatch on the material at run time (serial harness); otherwise the caller guarantees the
// hit primitive's material type (the wavefront shade kernels run one launch per material family, so the
// other families' code is compiled out and the warp stays converged).
template <int MATSEL, int FEAT = RSB_FEAT_ALL, class Stats>
RSB_HD int path_shade(const Scene& sc, const Spectral& sp, const RayConfig& cfg, PathState& ps, const HitRec& rec,
                      double normalisation, Rng& rng, KdStackEntry* stack, PathLog& log, Stats& stats) {
    const V3 o = ps.o, d = ps.d;
    const int depth = ps.depth;
    ps.keep_alive = 0;
    {
        Isect is;
        world_hit_geometry<FEAT>(sc, o, d, rec, &is);
        const Prim& prim = sc.prims[rec.prim];
        const Material& mat = sp.mats[prim.material];
        const int mtype = MATSEL >= 0 ? MATSEL : mat.type;
        const double* w2p = prim.to_local;
        const double* p2w = prim.to_root;

        // entries of this segment, pushed in reverse application order: normalisation, volumes, surface
        if (normalisation != 1.0) log.push(LOG_MULS, 0, normalisation);
        V3 w_hit = xform_point(p2w, is.hit);
        log_volumes<FEAT>(sc, sp, o, w_hit, stack, log, stats);

        if (mtype == MAT_EMITTER) {
            if (mat.table2 >= 0) {
                // Checkerboard.evaluate_surface (emitter/checkerboard.pyx:101-127): the parity of the LOCAL hit point's cell
                // along x, y, z picks square one (row table, scale) or square two (row table2, index_in)
                bool v = false;
                v = checkerboard_flip(v, is.hit.x, mat.index_out);
                v = checkerboard_flip(v, is.hit.y, mat.index_out);
                v = checkerboard_flip(v, is.hit.z, mat.index_out);
                if (v) log.push(LOG_EMIT, mat.table, mat.scale);
                else log.push(LOG_EMIT, mat.table2, mat.index_in);
            } else {
                log.push(LOG_EMIT, mat.table, mat.scale);
            }
            stats.table_read();
            return PATH_EMITTED;
        }
        if (mtype == MAT_ABSORBER) return PATH_ZERO;

        V3 next_o, next_d;
        // RoughConductor is a ContinuousBSDF like Lambert: same surface frame and multiple importance sampling, its own
        // sample / pdf / evaluate_shading (full-featured instantiation only; it is dealt to the Lambert shade list)
        const bool rough = (FEAT & RSB_FEAT_RARE_MATERIALS) && mat.type == MAT_ROUGH_CONDUCTOR;
        if (mtype == MAT_LAMBERT || rough) {
            // ContinuousBSDF.evaluate_surface (material.pyx:291-361) + Lambert (lambert.pyx:71-105)
            V3 normal = is.normal;
            V3 w_reflection_origin;
            if (is.exiting) {
                w_reflection_origin = xform_point(p2w, is.inside);
                normal = v3(-normal.x, -normal.y, -normal.z);
            } else {
                w_reflection_origin = xform_point(p2w, is.outside);
            }
            // _generate_surface_transforms (material.pyx:393-422)
            V3 tangent = orthogonal(normal);
            V3 bitangent = cross(normal, tangent);
            double p2s[9] = {tangent.x, tangent.y, tangent.z, bitangent.x, bitangent.y, bitangent.z, normal.x, normal.y, normal.z};
            double s2p[9] = {tangent.x, bitangent.x, normal.x, tangent.y, bitangent.y, normal.y, tangent.z, bitangent.z, normal.z};
            // AffineMatrix3D.mul (affinematrix.pyx:254-273), 3x3 block; the 4th product of each sum is 0*0
            double w2s[9], s2w[9];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    w2s[3 * r + c] = p2s[3 * r + 0] * w2p[0 + c] + p2s[3 * r + 1] * w2p[4 + c] + p2s[3 * r + 2] * w2p[8 + c] + 0.0 * 0.0;
                    s2w[3 * r + c] = p2w[4 * r + 0] * s2p[0 + c] + p2w[4 * r + 1] * s2p[3 + c] + p2w[4 * r + 2] * s2p[6 + c] + p2w[4 * r + 3] * 0.0;
                }
            V3 s_outgoing, w_outgoing;
            double pdf;
            // s_incoming = ray.direction.transform(world_to_surface).neg() (material.pyx:326); Lambert never looks at it
            V3 s_incoming = v3(0, 0, 0);
            const double roughness = mat.scale;
            if (rough) {
                V3 t = xform_vector33(w2s, d);
                s_incoming = v3(-t.x, -t.y, -t.z);
            }
            bool mis = cfg.importance_sampling && sc.imp_total > 0;
            if (mis) {
                if (rng.probability(cfg.important_path_weight)) {
                    w_outgoing = important_direction_sample(sc, rng, w_hit);
                    s_outgoing = xform_vector33(w2s, w_outgoing);
                } else {
                    s_outgoing = rough ? ggx_sample(rng, s_incoming, roughness) : hemisphere_cosine_sample(rng);
                    w_outgoing = xform_vector33(s2w, s_outgoing);
                }
                double pdf_important = important_direction_pdf(sc, w_hit, w_outgoing);
                double pdf_bsdf = rough ? ggx_pdf(s_incoming, s_outgoing, roughness) : hemisphere_cosine_pdf(s_outgoing);
                pdf = cfg.important_path_weight * pdf_important + (1 - cfg.important_path_weight) * pdf_bsdf;
#ifdef RSB_DEBUG_MIS
                fprintf(stderr, "MIS w_hit=(%.17g, %.17g, %.17g) w_out=(%.17g, %.17g, %.17g) pdf_imp=%.17g pdf_bsdf=%.17g pdf=%.17g\n", w_hit.x, w_hit.y, w_hit.z, w_outgoing.x, w_outgoing.y, w_outgoing.z, pdf_important, pdf_bsdf, pdf);
#endif
            } else {
                s_outgoing = rough ? ggx_sample(rng, s_incoming, roughness) : hemisphere_cosine_sample(rng);
                pdf = rough ? ggx_pdf(s_incoming, s_outgoing, roughness) : hemisphere_cosine_pdf(s_outgoing);
            }
            if (rough) {
                // RoughConductor.evaluate_shading (conductor.pyx:249-289): no transmission, no grazing incidence
                if (s_outgoing.z <= 0) return PATH_ZERO;
                if (s_incoming.z == 0) return PATH_ZERO;
                V3 s_half = normalise(v3(s_incoming.x + s_outgoing.x, s_incoming.y + s_outgoing.y, s_incoming.z + s_outgoing.z));
                // unwind order: mul_scalar(D * G / (4 cos_i)), per-bin Fresnel (_f, conductor.pyx:312-328), div_scalar(pdf)
                log.push(LOG_MULS, 0, 1.0 / pdf);
                log.push(LOG_FRESNEL | (mat.table2 << 8), mat.table, dot(s_half, s_outgoing));
                log.push(LOG_MULS, 0, ggx_d(s_half, roughness) * (ggx_g1(s_incoming, roughness) * ggx_g1(s_outgoing, roughness)) / (4 * s_incoming.z));
                stats.table_read();
                stats.table_read();
                ps.o = w_reflection_origin;
                ps.d = xform_vector33(s2w, s_outgoing);
                ps.depth = depth + 1;
                ps.rays += 1;
                return PATH_CONTINUE;
            }
            // Lambert.evaluate_shading (lambert.pyx:77-105)
            double pdf_cos = hemisphere_cosine_pdf(s_outgoing);
            if (pdf_cos == 0.0) return PATH_ZERO;
            next_o = w_reflection_origin;
            next_d = xform_vector33(s2w, s_outgoing);
            // unwind order: mul_array(reflectivity), mul_scalar(pdf_cos), div_scalar(pdf) => pushed reversed
            log.push(LOG_MULS, 0, 1.0 / pdf);
            log.push(LOG_MULS, 0, pdf_cos);
            log.push(LOG_MULA, mat.table, 0.0);
            stats.table_read();
        } else if ((FEAT & RSB_FEAT_RARE_MATERIALS) && mat.type == MAT_VOLUME_EMITTER) {
            // NullSurface.evaluate_surface (material.pyx:126-147): the ray carries on through the surface in the same
            // direction; the transit is not counted in the depth and the daughter cannot be extinguished
            next_o = is.exiting ? xform_point(p2w, is.outside) : xform_point(p2w, is.inside);
            ps.o = next_o;
            ps.rays += 1;
            ps.keep_alive = 1;
            return PATH_CONTINUE;
        } else if ((FEAT & RSB_FEAT_RARE_MATERIALS) && mat.type == MAT_CONDUCTOR) {
            // Conductor.evaluate_surface (conductor.pyx:75-130): mirror reflection, no random draws; the per-bin
            // Fresnel reflectance of the complex index n + ik is applied on the way back up
            V3 incident = normalise(xform_vector(w2p, d));
            V3 normal = normalise(is.normal);
            double ci = dot(normal, incident);
            double temp = 2 * ci;
            V3 reflected = v3(incident.x - temp * normal.x, incident.y - temp * normal.y, incident.z - temp * normal.z);
            next_d = xform_vector(p2w, reflected);
            // "we do not use the supplied exiting parameter" (conductor.pyx:108-118): the side comes from ci
            next_o = (ci > 0.0) ? xform_point(p2w, is.inside) : xform_point(p2w, is.outside);
            log.push(LOG_FRESNEL | (mat.table2 << 8), mat.table, fabs(ci));
            stats.table_read();
            stats.table_read();
        } else {
            // Dielectric.evaluate_surface (dielectric.pyx:153-303)
            V3 incident = normalise(xform_vector(w2p, d));
            V3 normal = normalise(is.normal);
            double c1 = -dot(normal, incident);
            double n1, n2;
            if (c1 < 0.0) { n1 = mat.index_in; n2 = mat.index_out; }
            else { n1 = mat.index_out; n2 = mat.index_in; }
            double gamma = n1 / n2;
            double c2s = 1 - (gamma * gamma) * (1 - c1 * c1);
            bool reflect;
            V3 transmitted = v3(0, 0, 0);
            if (c2s <= 0) {
                if (mat.transmission_only) return PATH_ZERO;
                reflect = true;   // total internal reflection
            } else {
                double temp;
                if (c1 < 0.0) temp = gamma * c1 + sqrt(c2s);
                else temp = gamma * c1 - sqrt(c2s);
                transmitted = v3(gamma * incident.x + temp * normal.x,
                                 gamma * incident.y + temp * normal.y,
                                 gamma * incident.z + temp * normal.z);
                // _fresnel (dielectric.pyx:305-308)
                double ci = c1, ct = -dot(normal, transmitted);
                double ra = (n1 * ci - n2 * ct) / (n1 * ci + n2 * ct);
                double rb = (n1 * ct - n2 * ci) / (n1 * ct + n2 * ci);
                double reflectivity = 0.5 * (ra * ra + rb * rb);
                double transmission = 1 - reflectivity;
                reflect = !(mat.transmission_only || rng.probability(transmission));
            }
            if (reflect) {
                double temp = 2 * c1;
                V3 reflected = v3(incident.x + temp * normal.x, incident.y + temp * normal.y, incident.z + temp * normal.z);
                next_d = xform_vector(p2w, reflected);
                next_o = (c1 < 0.0) ? xform_point(p2w, is.inside) : xform_point(p2w, is.outside);
            } else {
                next_d = xform_vector(p2w, transmitted);
                next_o = (c1 < 0.0) ? xform_point(p2w, is.outside) : xform_point(p2w, is.inside);
            }
        }
        // spawn_daughter (ray.pyx:506-549)
        ps.o = next_o;
        ps.d = next_d;
        ps.depth = depth + 1;
        ps.rays += 1;
        return PATH_CONTINUE;
    }
}

template <class Stats>
RSB_HD int path_step(const Scene& sc, const Spectral& sp, const RayConfig& cfg, PathState& ps, Rng& rng,
                     KdStackEntry* stack, PathLog& log, Stats& stats) {
    HitRec rec;
    double normalisation;
    int r = path_trace(sc, cfg, ps, rng, stack, &rec, &normalisation, stats);
    if (r != PATH_CONTINUE) return r;
    return path_shade<-1>(sc, sp, cfg, ps, rec, normalisation, rng, stack, log, stats);
}

// A whole path (serial harness).
template <class Stats>
RSB_HD int trace_path(const Scene& sc, const Spectral& sp, const RayConfig& cfg, const V3& o, const V3& d, Rng& rng,
                      KdStackEntry* stack, PathLog& log, uint32_t* ray_count, Stats& stats) {
    PathState ps;
    path_begin(ps, log, o, d);
    int r;
    do { r = path_step(sc, sp, cfg, ps, rng, stack, log, stats); } while (r == PATH_CONTINUE);
    *ray_count = ps.rays;
    // a path that ends dark still carries what emitting volumes added along the way: it must be replayed
    if (r == PATH_ZERO && log.additive) r = PATH_EMITTED;
    return r;
}

// x / d for an integer-valued divisor d with r = 1.0 / d precomputed: the correctly rounded quotient via
// Markstein's residual correction (q = x*r; q += r * fma(-d, q, x), twice) -- bit-identical to the IEEE
// division the reference performs, at 5 instructions instead of ~25.  The per-sample Welford update divides
// every bin of every sample by the same two small integers, so the reciprocals are formed once per path.
// Outside a safe magnitude window (and for zeros, whose sign the shortcut does not preserve) it divides.
RSB_HD double div_count(double x, double d, double r) { return div_exact(x, d, r); }

// raysect/core/math/statsarray.pyx:743-777 (_add_sample), n is the count BEFORE this sample;
// r_nn = 1.0 / (n + 1), r_nn1 = 1.0 / n (only read when n >= 1)
RSB_HD void welford_add_r(double sample, double prev_m, double prev_v, int n, double r_nn, double r_nn1, double* m, double* v) {
    if (n == 0) {
        *m = sample;
        *v = 0;
    } else {
        int prev_n = n > 1 ? n : 2;
        int nn = n + 1;
        double mm = prev_m + div_count(sample - prev_m, (double)nn, r_nn);
        *m = mm;
        *v = div_count(prev_v * (prev_n - 1) + (sample - prev_m) * (sample - mm), (double)(nn - 1), r_nn1);
    }
}

RSB_HD void welford_add(double sample, double* m, double* v, int n) {
    welford_add_r(sample, *m, *v, n, 1.0 / (double)(n + 1), n > 0 ? 1.0 / (double)n : 0.0, m, v);
}

// One log entry applied to one bin's running value (the reference's unwind, one Spectrum op at a time)
// Conductor._fresnel (conductor.pyx:132-145)
RSB_HD double conductor_fresnel(double ci, double n, double k) {
    double ci2 = ci * ci;
    double k0 = n * n + k * k;
    double k1 = k0 * ci2 + 1;
    double k2 = 2 * n * ci;
    double k3 = k0 + ci2;
    return 0.5 * ((k1 - k2) / (k1 + k2) + (k3 - k2) / (k3 + k2));
}

RSB_HD double apply_entry(double s, int op, int table, double v, const Spectral& sp, int bin) {
    if (op == LOG_MULS) return s * v;
    double t = sp.tables[(size_t)table * sp.bins + bin];
    if ((op & 0xff) == LOG_FRESNEL) return s * conductor_fresnel(v, t, sp.tables[(size_t)(op >> 8) * sp.bins + bin]);
    if (op == LOG_MULA) return s * t;
    if (op == LOG_ADDA) return s + (0.0 + t * sp.mats[table].scale) * v;
    if (op == LOG_POWA) {
#ifdef __CUDA_ARCH__
        // Dielectric.evaluate_volume's pow(transmission, length) (dielectric.pyx:326) as exp(length * ln T) with ln T
        // tabulated per slice: ~6x fewer instructions than CUDA's fp64 pow, which is itself only accurate to 2 ulp
        // and so was never bit-comparable with glibc's; |error| <= (1 + |length ln T|) * 2.2e-16 relative.
        if (sp.tables_ln != nullptr) return s * (v == 0.0 ? 1.0 : exp(v * sp.tables_ln[(size_t)table * sp.bins + bin]));
#endif
        return s * pow(t, v);
    }
    return t * v;
}

// Backward replay of a path log for one bin (the reference's unwind for samples_mv[bin]).
RSB_HD double replay_bin(const PathLog& log, const Spectral& sp, int bin) {
    double s = 0.0;
    for (int k = log.n - 1; k >= 0; --k) {
        LogEntry e = log.get(k);
        s = apply_entry(s, e.op, e.table, e.v, sp, bin);
    }
    return s;
}

// raysect/core/math/statsarray.pyx:780-857 (_combine_samples)
RSB_HD void stats_combine(double mx, double vx, int nx, double my, double vy, int ny, double* mt, double* vt, int* nt) {
    if (nx < ny) {
        int ti = nx; nx = ny; ny = ti;
        double td = mx; mx = my; my = td;
        td = vx; vx = vy; vy = td;
    }
    if (nx > 1 && ny > 1) {
        *nt = nx + ny;
        *mt = (nx * mx + ny * my) / (double)*nt;
        vx = (nx - 1) * vx / (double)nx;
        vy = (ny - 1) * vy / (double)ny;
        *vt = (nx * (mx * mx + vx) + ny * (my * my + vy)) / (double)*nt - *mt * *mt;
        *vt = *nt * *vt / (double)(*nt - 1);
        return;
    }
    if (nx == 0 && ny == 0) { *nt = 0; *mt = 0; *vt = 0; }
    else if (nx == 1) {
        if (ny == 0) { *nt = 1; *mt = mx; *vt = 0; }
        else {
            *nt = 2;
            *mt = 0.5 * (mx + my);
            double temp = mx - *mt;
            *vt = 2 * temp * temp;
        }
    } else if (nx > 1) {
        *nt = nx; *mt = mx; *vt = vx;
        if (ny == 1) { welford_add(my, mt, vt, *nt); *nt += 1; }
    }
}

// PinholeCamera._generate_rays for one sample (pinhole.pyx:169-204): jitter (u1, u2) -> local
// direction + projection weight, then _render_pixel's camera->world transform (observer.pyx:400-403).
// uniform() draws a camera takes per pixel task BEFORE any ray is traced, per sample: 2 (the point on the pixel), or 4 for the
// CCD, which draws all its pixel points first and all its directions after them (ccd.pyx:130-131)
RSB_HD int camera_jitter_pairs(int kind) { return (kind == 2 || kind == 4) ? 2 : 1; }

// VectorCamera._generate_rays (imaging/vector.pyx:124-125): only pixels off the edge of the image are sub-sampled -- an edge
// pixel's task draws NOTHING before its rays are traced
RSB_HD bool camera_pixel_draws(const Camera& cam, int px, int py) {
    return cam.kind != 3 || (0 < px && px < cam.nx - 1 && 0 < py && py < cam.ny - 1);
}

// Vector3D.slerp (core/math/vector.pyx:501-604)
RSB_HD V3 vector_slerp(const V3& a, const V3& b, double t) {
    const V3 an = normalise(a), bn = normalise(b);
    const double a_magnitude = sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    const double b_magnitude = sqrt(b.x * b.x + b.y * b.y + b.z * b.z);
    double angle = acos(an.x * bn.x + an.y * bn.y + an.z * bn.z);
    V3 v;
    if (angle < 1e-12) {
        v = an;
    } else {
        angle *= t;
        const double d = an.x * bn.x + an.y * bn.y + an.z * bn.z;
        const V3 e = normalise(v3(bn.x - an.x * d, bn.y - an.y * d, bn.z - an.z * d));
        const double c = cos(angle), s = sin(angle);
        v = v3(an.x * c + e.x * s, an.y * c + e.y * s, an.z * c + e.z * s);
    }
    const double m = (1 - t) * a_magnitude + t * b_magnitude;
    return v3(v.x * m, v.y * m, v.z * m);
}

// (u1, u2): the sample's point draws; (u3, u4): its direction draws (CCDArray only)
RSB_HD void pinhole_ray(const Camera& cam, int px, int py, double u1, double u2, V3* o, V3* d, double* weight, double u3 = 0.0,
                        double u4 = 0.0) {
    double pixel_x = cam.image_start_x - cam.image_delta * (px + 0.5);
    double pixel_y = cam.image_start_y - cam.image_delta * (py + 0.5);
    // RectangleSampler3D.sample (surface3d.pyx:197-198): width = height = image_delta, offsets 0.5*width
    // The two uniform() calls are arguments of one C call in the Cython output,
    // new_point3d(uniform()*w - ow, uniform()*h - oh, 0), and gcc evaluates call arguments right to left:
    // the FIRST draw of a sample lands in y, the SECOND in x (verified against the compiled reference).
    if (cam.kind == 3) {
        // VectorCamera._generate_rays (imaging/vector.pyx:107-153): the pixel's own origin; off the edge of the image the
        // direction is interpolated between the four diagonal neighbours' (point_square draws: new_point2d(uniform(),
        // uniform()) -- arguments evaluated right to left, so the FIRST draw is y), at the edge it is the pixel's own
        const size_t at = ((size_t)px * cam.ny + py) * 3;
        const V3 origin = v3(cam.pixel_origins[at], cam.pixel_origins[at + 1], cam.pixel_origins[at + 2]);
        V3 direction;
        if (camera_pixel_draws(cam, px, py)) {
            const double sx = u2, sy = u1;
            const double* D = cam.pixel_directions;
            const size_t r0 = ((size_t)(px - 1) * cam.ny + py) * 3, r1 = ((size_t)(px + 1) * cam.ny + py) * 3;
            const V3 v1 = v3(D[r0 - 3], D[r0 - 2], D[r0 - 1]), v2 = v3(D[r0 + 3], D[r0 + 4], D[r0 + 5]);
            const V3 v3_ = v3(D[r1 + 3], D[r1 + 4], D[r1 + 5]), v4 = v3(D[r1 - 3], D[r1 - 2], D[r1 - 1]);
            const V3 v14 = vector_slerp(v1, v4, sx), v23 = vector_slerp(v2, v3_, sx);
            direction = normalise(vector_slerp(v14, v23, sy));
        } else {
            direction = v3(cam.pixel_directions[at], cam.pixel_directions[at + 1], cam.pixel_directions[at + 2]);
        }
        *weight = 1.0;
        *o = xform_point(cam.to_root, origin);
        *d = xform_vector(cam.to_root, direction);
        return;
    }
    double half = 0.5 * cam.image_delta;
    double jy = u1 * cam.image_delta - half;
    double jx = u2 * cam.image_delta - half;
    if (cam.kind == 1) {
        // OrthographicCamera._generate_rays (orthographic.pyx:139-167): the sample point is moved to the pixel with
        // Point3D.transform(translate(pixel_x, pixel_y, 0)) and the ray leaves it along +z; "non-physical camera
        // samples radiance directly": projection weight 1
        const double pixel_to_local[RSB_MAT_WORDS] = {1.0, 0.0, 0.0, pixel_x, 0.0, 1.0, 0.0, pixel_y, 0.0, 0.0, 1.0, 0.0, 1.0};
        V3 origin = xform_point(pixel_to_local, v3(jx, jy, 0.0));
        *weight = 1.0;
        *o = xform_point(cam.to_root, origin);
        *d = xform_vector(cam.to_root, v3(0.0, 0.0, 1.0));
        return;
    }
    if (cam.kind == 4) {
        // Pixel._generate_rays (nonimaging/pixel.pyx:152-170), a 0-D observer: every "pixel" of the frame is one TASK of the
        // same x_width x y_width rectangle at the observer's origin (image_delta = x_width, image_start_x = y_width).  A
        // RectangleSampler3D point (first draw -> y, second -> x, as above) and a cosine-weighted direction; weight 0.5
        const double w = cam.image_delta, h = cam.image_start_x;
        V3 origin = v3(u2 * w - 0.5 * w, u1 * h - 0.5 * h, 0.0);
        *weight = 0.5;
        *o = xform_point(cam.to_root, origin);
        *d = xform_vector(cam.to_root, hemisphere_cosine_from(u3, u4));
        return;
    }
    if (cam.kind == 2) {
        // CCDArray._generate_rays (imaging/ccd.pyx:114-148): point on the pixel and a cosine-weighted direction over the
        // hemisphere in front of the sensor, both moved to the pixel with .transform(translate(pixel_x, pixel_y, 0));
        // "projected area cosine is implicit in distribution": weight 0.5
        const double pixel_to_local[RSB_MAT_WORDS] = {1.0, 0.0, 0.0, pixel_x, 0.0, 1.0, 0.0, pixel_y, 0.0, 0.0, 1.0, 0.0, 1.0};
        V3 origin = xform_point(pixel_to_local, v3(jx, jy, 0.0));
        V3 direction = xform_vector(pixel_to_local, hemisphere_cosine_from(u3, u4));
        *weight = 0.5;
        *o = xform_point(cam.to_root, origin);
        *d = xform_vector(cam.to_root, direction);
        return;
    }
    V3 dir = normalise(v3(jx + pixel_x, jy + pixel_y, 0.0 + 1.0));
    *weight = dir.z;
    *o = xform_point(cam.to_root, v3(0, 0, 0));
    *d = xform_vector(cam.to_root, dir);
}

}  // namespace rsb
