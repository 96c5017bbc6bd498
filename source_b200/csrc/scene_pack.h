// scene_pack.h -- RsbSceneDesc (C ABI) -> packed host arrays in the device layout of rsb_scene.h.
// Pure host C++: validation, kd-stream parsing, triangle pre-gather, importance CDF.  The CUDA side
// (raysect_b200.cu) only uploads these vectors.
#pragma once
#include <string>
#include <vector>

#include "../../include/raysect_b200.h"
#include "kdtree_host.h"
#include "rsb_scene.h"

namespace rsb {

struct PackedMesh {
    std::vector<F4> tri;
    std::vector<int32_t> tri_idx;
    std::vector<float> vnormals;
    HostKdTree tree;
    int32_t n_tri = 0, idx_stride = 3, smoothing = 0, closed = 0;
};

struct PackedScene {
    std::vector<Prim> prims;
    int32_t n_world = 0;
    HostKdTree world;
    std::vector<PackedMesh> meshes;
    std::vector<int32_t> mat_type, mat_transmission_only;
    std::vector<double> imp_sphere, imp_weight, imp_cdf;
    double imp_total = 0;
};

// returns RSB_OK or an error code with *err filled
int pack_scene(const RsbSceneDesc* d, PackedScene* out, std::string* err);

// MeshData._generate_face_normals for one triangle (mesh.pyx:428-462)
void mesh_face_normal(const float* vertices, const int32_t* row, float* out);

}  // namespace rsb
