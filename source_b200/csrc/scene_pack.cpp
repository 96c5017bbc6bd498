// scene_pack.cpp -- see scene_pack.h
#include "scene_pack.h"

#include <string.h>

#include <algorithm>

#include "rsb_geom.h"

namespace rsb {

void mesh_face_normal(const float* vertices, const int32_t* row, float* out) {
    const float* a = vertices + 3 * (size_t)row[0];
    const float* b = vertices + 3 * (size_t)row[1];
    const float* c = vertices + 3 * (size_t)row[2];
    V3 p1 = v3(a[0], a[1], a[2]), p2 = v3(b[0], b[1], b[2]), p3 = v3(c[0], c[1], c[2]);
    V3 e1 = v3(p2.x - p1.x, p2.y - p1.y, p2.z - p1.z);   // p1.vector_to(p2)
    V3 e2 = v3(p3.x - p1.x, p3.y - p1.y, p3.z - p1.z);
    V3 n = normalise(cross(e1, e2));
    out[0] = (float)n.x;
    out[1] = (float)n.y;
    out[2] = (float)n.z;
}

namespace {
int csg_depth_and_leaves(const RsbSceneDesc* d, int id, int depth, int* leaves, std::string* err) {
    if (id < 0 || id >= d->n_primitives) { *err = "CSG operand row out of range"; return -1; }
    int t = d->prim_type[id];
    if (t == RSB_PRIM_TORUS && depth > 0) { *err = "Torus operands inside CSG are not supported on the device path"; return -1; }
    if (t <= RSB_PRIM_CONE) { *leaves += 1; return depth; }
    if (t == RSB_PRIM_MESH) { *err = "Mesh operands inside CSG are not supported on the device path"; return -1; }
    if (depth > 16) { *err = "CSG tree too deep (cycle?)"; return -1; }
    int da = csg_depth_and_leaves(d, d->prim_child_a[id], depth + 1, leaves, err);
    if (da < 0) return -1;
    int db = csg_depth_and_leaves(d, d->prim_child_b[id], depth + 1, leaves, err);
    if (db < 0) return -1;
    return std::max(da, db);
}
}  // namespace

int pack_scene(const RsbSceneDesc* d, PackedScene* out, std::string* err) {
    // an EMPTY world is legal (the reference builds a one-leaf tree and every query misses)
    if (d->n_primitives < 0 || d->n_world < 0 || d->n_world > d->n_primitives || (d->n_primitives > 0 && d->n_world == 0)) {
        *err = "bad primitive counts";
        return RSB_ERR_ARG;
    }
    for (int i = 0; i < d->n_primitives; ++i) {
        int t = d->prim_type[i];
        if (t < RSB_PRIM_TORUS || t > RSB_PRIM_SUBTRACT) { *err = "unsupported primitive type in row " + std::to_string(i); return RSB_ERR_UNSUPPORTED; }
        if (t == RSB_PRIM_MESH && (d->prim_mesh[i] < 0 || d->prim_mesh[i] >= d->n_meshes)) { *err = "mesh row out of range"; return RSB_ERR_ARG; }
        if (i < d->n_world && (d->prim_material[i] < 0 || d->prim_material[i] >= d->n_materials)) {
            *err = "material row out of range for primitive " + std::to_string(i);
            return RSB_ERR_ARG;
        }
        if (i < d->n_world && t >= RSB_PRIM_UNION) {
            int leaves = 0;
            int depth = csg_depth_and_leaves(d, i, 0, &leaves, err);
            if (depth < 0) return RSB_ERR_UNSUPPORTED;
            if (depth > RSB_CSG_MAX_DEPTH) { *err = "CSG nesting deeper than " + std::to_string(RSB_CSG_MAX_DEPTH) + " operators"; return RSB_ERR_UNSUPPORTED; }
            if (2 * leaves > RSB_CSG_MAX_EVENTS) { *err = "CSG primitive with more than " + std::to_string(RSB_CSG_MAX_EVENTS / 2) + " leaves"; return RSB_ERR_UNSUPPORTED; }
        }
    }
    for (int i = 0; i < d->n_materials; ++i)
        if (d->mat_type[i] < RSB_MAT_ABSORBER || d->mat_type[i] > RSB_MAT_ROUGH_CONDUCTOR) {
            *err = "unsupported material type in row " + std::to_string(i);
            return RSB_ERR_UNSUPPORTED;
        }

    out->prims.assign((size_t)d->n_primitives, Prim{});
    for (int i = 0; i < d->n_primitives; ++i) {
        Prim& p = out->prims[i];
        memset(&p, 0, sizeof(p));
        memcpy(p.bbox, d->prim_bbox + 6 * (size_t)i, 48);
        memcpy(p.params, d->prim_params + 6 * (size_t)i, 48);
        p.type = d->prim_type[i];
        p.material = d->prim_material[i];
        p.child_a = d->prim_child_a[i];
        p.child_b = d->prim_child_b[i];
        p.mesh = d->prim_mesh[i];
        p.parent = d->prim_parent[i];
        // [13] per matrix: rows 0..2, then m33; the device keeps 1 / m33 (Point3D.transform's `w = 1.0 / w`, point.pyx:276)
        const double* src[3] = {d->prim_to_local + 13 * (size_t)i, d->prim_to_root + 13 * (size_t)i, d->prim_root_inv + 13 * (size_t)i};
        double* dst[3] = {p.to_local, p.to_root, p.root_inv};
        for (int k = 0; k < 3; ++k) {
            memcpy(dst[k], src[k], 96);
            if (src[k][12] == 0.0) { *err = "Bad matrix transform, 4th element of homogeneous coordinate is zero."; return RSB_ERR_ARG; }
            dst[k][12] = 1.0 / src[k][12];
        }
    }
    out->n_world = d->n_world;

    std::string kerr;
    if (kd_parse_stream(d->world_kdtree, d->world_kdtree_bytes, &out->world, &kerr) < 0) { *err = "world " + kerr; return RSB_ERR_ARG; }
    for (int32_t id : out->world.items)
        if (id < 0 || id >= d->n_world) { *err = "world kd-tree references a primitive outside World.primitives"; return RSB_ERR_ARG; }
    if (out->world.depth >= RSB_KD_STACK / 2) { *err = "world kd-tree deeper than " + std::to_string(RSB_KD_STACK / 2 - 1); return RSB_ERR_UNSUPPORTED; }

    out->meshes.assign((size_t)d->n_meshes, PackedMesh{});
    for (int mi = 0; mi < d->n_meshes; ++mi) {
        const RsbMeshDesc& md = d->meshes[mi];
        PackedMesh& m = out->meshes[mi];
        if (!md.vertices || !md.triangles || md.n_triangles <= 0 || (md.tri_stride != 3 && md.tri_stride != 6)) {
            *err = "mesh " + std::to_string(mi) + ": bad arrays";
            return RSB_ERR_ARG;
        }
        if (md.tri_stride == 6 && !md.vertex_normals) { *err = "The triangle array must have dimensions Nx3."; return RSB_ERR_ARG; }
        m.tri.resize((size_t)md.n_triangles * 3);
        for (int32_t t = 0; t < md.n_triangles; ++t) {
            const int32_t* row = md.triangles + (size_t)t * md.tri_stride;
            for (int k = 0; k < 3; ++k)
                if (row[k] < 0 || row[k] >= md.n_vertices) { *err = "The triangle array references non-existent vertices."; return RSB_ERR_ARG; }
            if (md.tri_stride == 6)
                for (int k = 3; k < 6; ++k)
                    if (row[k] < 0 || row[k] >= md.n_vertex_normals) { *err = "The triangle array references non-existent normals."; return RSB_ERR_ARG; }
            const float* a = md.vertices + 3 * (size_t)row[0];
            const float* b = md.vertices + 3 * (size_t)row[1];
            const float* c = md.vertices + 3 * (size_t)row[2];
            float fn[3];
            if (md.face_normals) memcpy(fn, md.face_normals + 3 * (size_t)t, 12);
            else mesh_face_normal(md.vertices, row, fn);
            F4* q = &m.tri[(size_t)t * 3];
            q[0].x = a[0]; q[0].y = a[1]; q[0].z = a[2]; q[0].w = b[0];
            q[1].x = b[1]; q[1].y = b[2]; q[1].z = c[0]; q[1].w = c[1];
            q[2].x = c[2]; q[2].y = fn[0]; q[2].z = fn[1]; q[2].w = fn[2];
        }
        m.tri_idx.assign(md.triangles, md.triangles + (size_t)md.n_triangles * md.tri_stride);
        if (md.vertex_normals && md.tri_stride == 6) m.vnormals.assign(md.vertex_normals, md.vertex_normals + (size_t)md.n_vertex_normals * 3);
        if (kd_parse_stream(md.kdtree, md.kdtree_bytes, &m.tree, &kerr) < 0) { *err = "mesh " + kerr; return RSB_ERR_ARG; }
        for (int32_t id : m.tree.items)
            if (id < 0 || id >= md.n_triangles) { *err = "mesh kd-tree references a non-existent triangle"; return RSB_ERR_ARG; }
        if (m.tree.depth >= RSB_KD_STACK / 2) { *err = "mesh kd-tree deeper than " + std::to_string(RSB_KD_STACK / 2 - 1); return RSB_ERR_UNSUPPORTED; }
        m.n_tri = md.n_triangles;
        m.idx_stride = md.tri_stride;
        m.smoothing = md.smoothing;
        m.closed = md.closed;
    }

    out->mat_type.assign(d->mat_type, d->mat_type + d->n_materials);
    out->mat_transmission_only.assign(d->mat_transmission_only, d->mat_transmission_only + d->n_materials);
    out->imp_total = 0;
    if (d->n_important > 0) {
        // ImportanceManager._process_primitives/_calculate_cdf (optical/scenegraph/world.pyx:88-132)
        out->imp_sphere.assign(d->imp_sphere, d->imp_sphere + (size_t)d->n_important * 4);
        out->imp_weight.assign(d->imp_weight, d->imp_weight + d->n_important);
        out->imp_cdf.resize((size_t)d->n_important);
        double total = 0;
        for (int i = 0; i < d->n_important; ++i) total += d->imp_weight[i];
        for (int i = 0; i < d->n_important; ++i) out->imp_cdf[i] = (i == 0) ? d->imp_weight[0] : out->imp_cdf[i - 1] + d->imp_weight[i];
        for (int i = 0; i < d->n_important; ++i) out->imp_cdf[i] /= total;
        out->imp_total = total;
    }
    return RSB_OK;
}

}  // namespace rsb
