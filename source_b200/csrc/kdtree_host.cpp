// kdtree_host.cpp -- see kdtree_host.h
#include "kdtree_host.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <exception>
#include <thread>

namespace rsb {

namespace {
struct Reader {
    const uint8_t* p;
    int64_t left;
    bool ok = true;
    template <class T>
    T get() {
        T v{};
        if (left < (int64_t)sizeof(T)) { ok = false; return v; }
        memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        left -= sizeof(T);
        return v;
    }
};

template <class T>
void put(std::vector<uint8_t>* out, T v) {
    const uint8_t* b = reinterpret_cast<const uint8_t*>(&v);
    out->insert(out->end(), b, b + sizeof(T));
}

int32_t measure_depth(const std::vector<KdNode>& nodes) {
    // pre-order layout: iterative walk with an explicit stack of (node, depth)
    if (nodes.empty()) return 0;
    int32_t deepest = 0;
    std::vector<std::pair<int32_t, int32_t>> st;
    st.push_back({0, 0});
    while (!st.empty()) {
        auto [id, d] = st.back();
        st.pop_back();
        deepest = std::max(deepest, d);
        const KdNode& n = nodes[id];
        if (n.axis >= 0) {
            st.push_back({n.upper, d + 1});
            st.push_back({id + 1, d + 1});
        }
    }
    return deepest;
}
}  // namespace

int64_t kd_parse_stream(const uint8_t* data, int64_t size, HostKdTree* out, std::string* err) {
    Reader r{data, size};
    out->max_depth = r.get<int32_t>();
    out->min_items = r.get<int32_t>();
    out->hit_cost = r.get<double>();
    out->empty_bonus = r.get<double>();
    for (int i = 0; i < 6; ++i) out->bounds[i] = r.get<double>();
    int32_t n = r.get<int32_t>();
    if (!r.ok || n <= 0) { *err = "kd-tree stream: truncated header or no nodes"; return -1; }
    if ((int64_t)n > r.left / 8) { *err = "kd-tree stream: node count exceeds the stream"; return -1; }   // a node is >= 8 bytes
    out->nodes.resize(n);
    out->items.clear();
    for (int32_t id = 0; id < n; ++id) {
        int32_t type = r.get<int32_t>();
        KdNode node;
        memset(&node, 0, sizeof(node));
        if (type == -1) {
            int32_t count = r.get<int32_t>();
            if (!r.ok || count < 0 || (int64_t)count > r.left / 4) { *err = "kd-tree stream: bad leaf"; return -1; }
            node.axis = -1;
            node.upper = -1;
            node.leaf.item_offset = (int32_t)out->items.size();
            node.leaf.item_count = count;
            for (int32_t k = 0; k < count; ++k) out->items.push_back(r.get<int32_t>());
        } else if (type >= 0 && type <= 2) {
            node.axis = type;
            node.split = r.get<double>();
            node.upper = r.get<int32_t>();
            if (node.upper <= id || node.upper >= n) { *err = "kd-tree stream: bad upper child id"; return -1; }
        } else {
            *err = "kd-tree stream: bad node type";
            return -1;
        }
        if (!r.ok) { *err = "kd-tree stream: truncated"; return -1; }
        out->nodes[id] = node;
    }
    out->depth = measure_depth(out->nodes);
    return size - r.left;
}

void kd_write_stream(const HostKdTree& tree, std::vector<uint8_t>* out) {
    put<int32_t>(out, tree.max_depth);
    put<int32_t>(out, tree.min_items);
    put<double>(out, tree.hit_cost);
    put<double>(out, tree.empty_bonus);
    for (int i = 0; i < 6; ++i) put<double>(out, tree.bounds[i]);
    put<int32_t>(out, (int32_t)tree.nodes.size());
    for (const KdNode& n : tree.nodes) {
        if (n.axis < 0) {
            put<int32_t>(out, -1);
            put<int32_t>(out, n.leaf.item_count);
            for (int32_t k = 0; k < n.leaf.item_count; ++k) put<int32_t>(out, tree.items[n.leaf.item_offset + k]);
        } else {
            put<int32_t>(out, n.axis);
            put<double>(out, n.split);
            put<int32_t>(out, n.upper);
        }
    }
}

namespace {

struct Edge {
    double value;
    bool upper;
};

// BoundingBox3D.surface_area (boundingbox.pyx:302-315)
inline double surface_area(const double* b) {
    double dx = b[3] - b[0], dy = b[4] - b[1], dz = b[5] - b[2];
    return 2 * (dx * dy + dx * dz + dy * dz);
}

// BoundingBox3D.largest_axis / extent (boundingbox.pyx:345-387): x unless y or z is strictly larger
inline int largest_axis(const double* b) {
    double dx = std::max(0.0, b[3] - b[0]), dy = std::max(0.0, b[4] - b[1]), dz = std::max(0.0, b[5] - b[2]);
    int axis = 0;
    double largest = dx;
    if (dy > largest) { largest = dy; axis = 1; }
    if (dz > largest) { largest = dz; axis = 2; }
    return axis;
}

// A subtree under construction: node ids and item offsets are LOCAL to the subtree (its root is node 0), so that
// independent subtrees can be built by different threads and spliced into the pre-order layout afterwards.
struct Sub {
    std::vector<KdNode> nodes;
    std::vector<int32_t> items;
    int32_t depth = 0;
};

struct Builder {
    const double* boxes;
    int32_t max_depth, min_items;
    double hit_cost, empty_bonus;
    Sub* out;
    std::vector<Edge> edges;

    int32_t new_leaf(const std::vector<int32_t>& items) {
        KdNode n;
        memset(&n, 0, sizeof(n));
        n.axis = -1;
        n.upper = -1;
        n.leaf.item_offset = (int32_t)out->items.size();
        n.leaf.item_count = (int32_t)items.size();
        out->items.insert(out->items.end(), items.begin(), items.end());
        out->nodes.push_back(n);
        return (int32_t)out->nodes.size() - 1;
    }

    // appends a finished subtree behind the nodes already present, shifting its local ids / offsets
    int32_t splice(const Sub& sub) {
        const int32_t node_base = (int32_t)out->nodes.size(), item_base = (int32_t)out->items.size();
        for (KdNode n : sub.nodes) {
            if (n.axis < 0) n.leaf.item_offset += item_base;
            else n.upper += node_base;
            out->nodes.push_back(n);
        }
        out->items.insert(out->items.end(), sub.items.begin(), sub.items.end());
        out->depth = std::max(out->depth, sub.depth);
        return node_base;
    }

    // kdtree3d.pyx:166-188 (_build) + :193-308 (_split) + :422-459 (_new_branch).  `fork` > 0: the two children of
    // this node are built concurrently (the lower one on a new thread) while enough items are left to pay for it;
    // the layout -- and therefore the serialised stream -- is the same as the serial build's.
    int32_t build(std::vector<int32_t>& items, const double* bounds, int32_t depth, int fork) {
        out->depth = std::max(out->depth, depth);
        if (depth == max_depth || (int32_t)items.size() <= min_items) return new_leaf(items);

        double best_cost = (double)items.size() * hit_cost;
        double best_split = 0;
        int best_axis = -1;
        bool is_leaf = true;
        double recip_total_sa = 1.0 / surface_area(bounds);
        int longest = largest_axis(bounds);
        for (int a = 0; a < 3 && is_leaf; ++a) {
            int axis = (longest + a) % 3;
            // _get_edges (:312-352): sorted by value, lower edge before upper edge at equal value
            // (_edge_compare :82-100); equal (value, kind) pairs are indistinguishable, so the
            // libc qsort order is reproduced by any correct sort on that key
            edges.resize(items.size() * 2);
            for (size_t i = 0; i < items.size(); ++i) {
                const double* b = boxes + 6 * (size_t)items[i];
                edges[2 * i] = Edge{b[axis], false};
                edges[2 * i + 1] = Edge{b[3 + axis], true};
            }
            std::sort(edges.begin(), edges.end(), [](const Edge& x, const Edge& y) {
                if (x.value != y.value) return x.value < y.value;
                return !x.upper && y.upper;
            });
            int64_t lower_count = 0, upper_count = (int64_t)items.size();
            double lo = bounds[axis], hi = bounds[3 + axis];
            for (const Edge& e : edges) {
                if (e.upper) upper_count -= 1;
                double split = e.value;
                if (lo < split && split < hi) {
                    double lb[6], ub[6];
                    memcpy(lb, bounds, sizeof(lb));
                    memcpy(ub, bounds, sizeof(ub));
                    lb[3 + axis] = split;
                    ub[axis] = split;
                    double lower_sa = surface_area(lb);
                    double upper_sa = surface_area(ub);
                    double bonus = 1.0;
                    if (lower_count == 0 || upper_count == 0) bonus -= empty_bonus;
                    // :264  cost = 1 + bonus * (lower_sa * lower_count + upper_sa * upper_count) * recip_total_sa * hit_cost
                    double cost = 1 + bonus * (lower_sa * (double)lower_count + upper_sa * (double)upper_count) * recip_total_sa * hit_cost;
                    if (cost < best_cost) {
                        best_cost = cost;
                        best_split = split;
                        best_axis = axis;
                        is_leaf = false;
                    }
                }
                if (!e.upper) lower_count += 1;
            }
        }
        if (is_leaf) return new_leaf(items);

        std::vector<int32_t> lower_items, upper_items;
        for (int32_t id : items) {
            const double* b = boxes + 6 * (size_t)id;
            if (b[best_axis] < best_split) lower_items.push_back(id);
            if (b[3 + best_axis] > best_split) upper_items.push_back(id);
        }
        double lb[6], ub[6];
        memcpy(lb, bounds, sizeof(lb));
        memcpy(ub, bounds, sizeof(ub));
        lb[3 + best_axis] = best_split;
        ub[best_axis] = best_split;
        // release this node's list before recursing (the tree can be 30 deep over millions of items)
        std::vector<int32_t>().swap(items);

        int32_t id = (int32_t)out->nodes.size();
        out->nodes.push_back(KdNode{});
        int32_t upper_id;
        if (fork > 0 && lower_items.size() + upper_items.size() >= 20000) {
            Sub lower_sub, upper_sub;
            Builder lower_builder{boxes, max_depth, min_items, hit_cost, empty_bonus, &lower_sub, {}};
            Builder upper_builder{boxes, max_depth, min_items, hit_cost, empty_bonus, &upper_sub, {}};
            // the worker is joined on every path out of this scope, and an exception on either side is re-thrown in
            // the forking thread (a joinable std::thread destroyed during unwinding would call std::terminate)
            std::exception_ptr worker_error;
            std::thread worker([&]() {
                try { lower_builder.build(lower_items, lb, depth + 1, fork - 1); } catch (...) { worker_error = std::current_exception(); }
            });
            try {
                upper_builder.build(upper_items, ub, depth + 1, fork - 1);
            } catch (...) {
                worker.join();
                throw;
            }
            worker.join();
            if (worker_error) std::rethrow_exception(worker_error);
            splice(lower_sub);
            upper_id = splice(upper_sub);
        } else {
            build(lower_items, lb, depth + 1, 0);
            upper_id = build(upper_items, ub, depth + 1, 0);
        }
        KdNode n;
        memset(&n, 0, sizeof(n));
        n.split = best_split;
        n.upper = upper_id;
        n.axis = best_axis;
        out->nodes[id] = n;
        return id;
    }
};
}  // namespace

void kd_build(const double* boxes, int64_t n_items, int32_t max_depth, int32_t min_items, double hit_cost,
              double empty_bonus, HostKdTree* out) {
    // kdtree3d.pyx:126-153
    out->empty_bonus = empty_bonus;
    out->max_depth = std::max(0, max_depth);
    out->min_items = std::max(1, min_items);
    out->hit_cost = std::max(1.0, hit_cost);
    // no items: log(0) = -inf, and the reference's <int32_t> cast of -inf is x86's cvttsd2si "integer indefinite"
    if (out->max_depth == 0) out->max_depth = n_items > 0 ? (int32_t)ceil(8 + 1.3 * log((double)n_items)) : INT32_MIN;
    // BoundingBox3D() default is the empty box (lower=+inf... no: see note) then union of item boxes
    double b[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = 0; i < n_items; ++i) {
        const double* x = boxes + 6 * i;
        for (int k = 0; k < 3; ++k) {
            b[k] = std::min(b[k], x[k]);
            b[3 + k] = std::max(b[3 + k], x[3 + k]);
        }
    }
    memcpy(out->bounds, b, sizeof(b));
    // RSB_KD_THREADS (default: hardware threads, at most 16): the top log2(threads) levels fork; 1 = serial
    int threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char* e = getenv("RSB_KD_THREADS")) { int v = atoi(e); if (v >= 1 && v <= 256) threads = v; }
    int fork = 0;
    while ((1 << (fork + 1)) <= threads) ++fork;
    Sub root;
    Builder bd{boxes, out->max_depth, out->min_items, out->hit_cost, out->empty_bonus, &root, {}};
    std::vector<int32_t> items((size_t)n_items);
    for (int64_t i = 0; i < n_items; ++i) items[(size_t)i] = (int32_t)i;
    bd.build(items, out->bounds, 0, fork);
    out->nodes.swap(root.nodes);
    out->items.swap(root.items);
    out->depth = root.depth;
}

}  // namespace rsb
