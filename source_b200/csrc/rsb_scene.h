// rsb_scene.h -- flattened, read-only scene as the kernels see it (HBM layout).
//
// Layout choices (see DESIGN.md "Data layout in HBM"):
//  * kd-tree nodes are packed to 16 B (one LDG.128 per visit) instead of the reference's 32-B
//    kdnode + malloc'd item arrays (raysect/core/math/spatial/kdtree3d.pxd:38-43); leaf item ids
//    live in one flat int32 array.  Node ids are the reference's own (pre-order, lower child = id+1).
//  * primitives are 384-B rows: a first 128-B line with everything the reject tests need
//    (world AABB, shape parameters, type/material/children), then the two affine 3x4 matrices.
//  * mesh triangles are pre-gathered into 48-B rows (3 x float4: v1, v2, v3, face normal), the
//    figure SURVEY 8(d) charges per triangle test, so a test is three coalescable 16-B loads.
#pragma once
#include "rsb_math.h"

namespace rsb {

enum PrimType : int32_t {
    PRIM_SPHERE = 0,
    PRIM_BOX = 1,
    PRIM_CYLINDER = 2,
    PRIM_PARABOLA = -1,   // raysect/primitive/parabola.pyx (analytic primitives are the types <= PRIM_CONE)
    PRIM_TORUS = -2,      // raysect/primitive/torus.pyx (world-level only: not a CSG operand)
    PRIM_CONE = 3,
    PRIM_MESH = 4,
    PRIM_UNION = 5,
    PRIM_INTERSECT = 6,
    PRIM_SUBTRACT = 7,
};

enum MatType : int32_t {
    MAT_ABSORBER = 0,   // raysect/optical/material/absorber.pyx:50-55
    MAT_EMITTER = 1,    // raysect/optical/material/emitter/uniform.pyx:67-81
    MAT_LAMBERT = 2,    // raysect/optical/material/lambert.pyx:77-105
    MAT_DIELECTRIC = 3, // raysect/optical/material/dielectric.pyx:153-330
    MAT_CONDUCTOR = 4,  // raysect/optical/material/conductor.pyx:75-147 (shaded with the dielectric family)
    MAT_ROUGH_CONDUCTOR = 6,  // raysect/optical/material/conductor.pyx:157-344 (a ContinuousBSDF: shaded with the Lambert family)
    MAT_VOLUME_EMITTER = 5,   // emitter/homogeneous.pyx:40-93 with uniform.pyx:91-133 / unity.pyx:79-99 (same family)
};

// 16-byte kd-tree node.  Branch: split, upper child id, axis 0..2 (lower child is id+1).
// Leaf: axis == -1, the two halves of `split`'s storage hold (first item offset, item count).
struct __attribute__((aligned(16))) KdNode {
    union {
        double split;
        struct { int32_t item_offset; int32_t item_count; } leaf;
    };
    int32_t upper;
    int32_t axis;
};

struct KdTree {
    const KdNode* nodes;
    const int32_t* items;
    double bounds[6];   // lower xyz, upper xyz
    int32_t n_nodes;
    int32_t max_depth;
};

struct __attribute__((aligned(16))) Prim {
    // -- line 0 (128 B): reject tests + dispatch
    double bbox[6];     // AABB in the parent space (world for top-level primitives, CSG-local for CSG children)
    double params[6];   // sphere: r | box: lower xyz, upper xyz | cylinder, cone: r, h
    int32_t type;
    int32_t material;
    int32_t child_a;    // CSG operands (row indices), else -1
    int32_t child_b;
    int32_t mesh;       // mesh index, else -1
    int32_t parent;     // enclosing CSG row, -1 for world-level primitives
    int32_t pad0, pad1;
    // -- lines 1..2
    double to_local[RSB_MAT_WORDS]; // parent space -> local: rows 0..2 of the affine matrix, then 1 / m33 (rsb_math.h xform_point)
    double to_root[RSB_MAT_WORDS];  // local -> parent space
    double root_inv[RSB_MAT_WORDS]; // AffineMatrix3D.inverse() of to_root as the reference recomputes it inside
                                    // Normal3D.transform (raysect/core/math/normal.pyx:241-247); CSG children only
};

struct __attribute__((aligned(16))) F4 {
    float x, y, z, w;
};

struct Mesh {
    const F4* tri;      // [n_tri][3]: (v1.xyz, v2.x) (v2.yz, v3.xy) (v3.z, fn.xyz)
    const int32_t* tri_idx; // [n_tri][stride] original index rows (v1,v2,v3[,n1,n2,n3])
    const float* vnormals;  // [n_vn][3] or null
    KdTree tree;
    int32_t n_tri;
    int32_t idx_stride;     // 3 or 6
    int32_t smoothing;
    int32_t closed;
};

struct Material {
    int32_t type;
    int32_t transmission_only;
    int32_t table;      // row of the per-slice spectral table (reflectivity | emission | transmission | conductor n)
    int32_t table2;     // conductor: row of the extinction table k; else -1
    double scale;       // emitter scale | rough conductor: roughness
    double index_in;    // dielectric: index.average(min,max) for the slice
    double index_out;   // dielectric: external_index.average(min,max)
};

// One leaf item of the world tree with the AABB the leaf test starts with (BoundPrimitive.hit, boundprimitive.pyx:42-51),
// in leaf order: the pre-test of a leaf reads consecutive 64-B rows instead of chasing item id -> primitive row.
struct __attribute__((aligned(64))) LeafRow {
    double bbox[6];
    int32_t id;
    int32_t type;
    int32_t pad[2];
};

struct Scene {
    const Prim* prims;
    const LeafRow* world_rows;  // [world items] or null (worlds staged in shared memory keep items + primitive table there)
    KdTree world;
    const Mesh* meshes;
    int32_t n_prims;    // all rows
    int32_t n_world;    // rows [0, n_world) are world.primitives in order
    int32_t n_meshes;
    // importance manager (raysect/optical/scenegraph/world.pyx:47-260)
    int32_t n_important;
    const double* imp_sphere;   // [n][4] centre xyz, radius
    const double* imp_weight;   // [n] importance
    const double* imp_cdf;      // [n]
    double imp_total;
};

}  // namespace rsb
