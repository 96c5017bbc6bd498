// kdtree_host.h -- host-side kd-tree: the reference's serialised stream <-> packed device nodes,
// and a bit-exact SAH builder.
//
// Stream format (raysect/core/math/spatial/kdtree3d.pyx:864-984), little endian:
//   i32 max_depth, i32 min_items, f64 hit_cost, f64 empty_bonus, 6 x f64 bounds (lower xyz, upper xyz),
//   i32 n_nodes, then per node:  leaf  = i32 -1, i32 count, count x i32 item ids
//                                branch = i32 axis, f64 split, i32 upper_id
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "rsb_scene.h"

namespace rsb {

struct HostKdTree {
    int32_t max_depth = 0, min_items = 1;
    double hit_cost = 0, empty_bonus = 0;
    double bounds[6] = {0, 0, 0, 0, 0, 0};
    std::vector<KdNode> nodes;
    std::vector<int32_t> items;
    int32_t depth = 0;   // deepest node actually present
};

// Parses a stream; returns bytes consumed or -1 (err filled).
int64_t kd_parse_stream(const uint8_t* data, int64_t size, HostKdTree* out, std::string* err);

// Serialises to the reference's stream format.
void kd_write_stream(const HostKdTree& tree, std::vector<uint8_t>* out);

// SAH build over item boxes (boxes[i] = lower xyz, upper xyz; item id = i), restating
// KDTree3DCore.__init__/_build/_split/_get_edges/_new_leaf/_new_branch (kdtree3d.pyx:126-459).
void kd_build(const double* boxes, int64_t n_items, int32_t max_depth, int32_t min_items, double hit_cost,
              double empty_bonus, HostKdTree* out);

}  // namespace rsb
