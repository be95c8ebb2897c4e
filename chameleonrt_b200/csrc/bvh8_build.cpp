// bvh8_build.cpp — host-side builder for the compressed 8-wide BVH (format: bvh8.h).
//
// Pipeline: (1) binned-SAH BVH2 with one triangle per leaf, built in parallel over subtrees
// on a std::thread task pool; (2) optimal SAH collapse of the binary tree into 8-wide nodes
// with <= 3 triangles per leaf (dynamic programme of Ylitie et al. 2017, §4); (3) emission:
// octant-ordered child slots, 8-bit quantised child boxes that are conservative in the
// arithmetic the traversal kernel uses, inner children and leaf triangles made contiguous.
//
// The reference has no builder (Embree / OptiX do it); nothing here is
// derived from reference code.
#include "bvh8.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <limits>
#include <mutex>
#include <thread>

namespace crt {
namespace {

struct Box {
    float lo[3], hi[3];
    void reset()
    {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::numeric_limits<float>::max();
            hi[a] = -std::numeric_limits<float>::max();
        }
    }
    void grow(const Box &b)
    {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], b.lo[a]);
            hi[a] = std::max(hi[a], b.hi[a]);
        }
    }
    void grow_pt(const float *p)
    {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], p[a]);
            hi[a] = std::max(hi[a], p[a]);
        }
    }
    float half_area() const
    {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
};

struct B2Node {
    Box box;
    uint32_t left;   // inner: index of left child (right = left + 1); leaf: triangle index
    uint32_t count;  // 0 = inner, 1 = leaf
};

struct Task {
    uint32_t node, first, count;
};

constexpr int kBins = 16;
constexpr uint32_t kParallelThreshold = 1u << 15;

struct Bvh2Builder {
    const Box *prim_box;
    const float *prim_cent;  // 3 per prim
    std::vector<uint32_t> prim_ids;
    std::vector<B2Node> nodes;
    std::atomic<uint32_t> next_node{1};

    std::mutex mtx;
    std::condition_variable cv;
    std::deque<Task> shared;
    int outstanding = 0;  // tasks queued or running (guarded by mtx)

    void push_shared(const Task &t)
    {
        {
            std::lock_guard<std::mutex> lk(mtx);
            shared.push_back(t);
            ++outstanding;
        }
        cv.notify_one();
    }

    // Splits [first, first+count) of prim_ids; returns the split position.
    uint32_t split(const Task &t, const Box &cbox)
    {
        const uint32_t first = t.first, count = t.count;
        Box bin_box[3][kBins];
        uint32_t bin_cnt[3][kBins];
        float scale[3], cmin[3];
        bool valid[3];
        for (int a = 0; a < 3; ++a) {
            cmin[a] = cbox.lo[a];
            const float ext = cbox.hi[a] - cbox.lo[a];
            valid[a] = ext > 0.f;
            scale[a] = valid[a] ? kBins / ext : 0.f;
            for (int b = 0; b < kBins; ++b) {
                bin_box[a][b].reset();
                bin_cnt[a][b] = 0;
            }
        }
        for (uint32_t i = first; i < first + count; ++i) {
            const uint32_t p = prim_ids[i];
            const float *c = prim_cent + 3 * (size_t)p;
            for (int a = 0; a < 3; ++a) {
                if (!valid[a]) {
                    continue;
                }
                int b = (int)((c[a] - cmin[a]) * scale[a]);
                b = std::min(std::max(b, 0), kBins - 1);
                bin_box[a][b].grow(prim_box[p]);
                bin_cnt[a][b]++;
            }
        }
        float best_cost = std::numeric_limits<float>::max();
        int best_axis = -1, best_bin = 0;
        for (int a = 0; a < 3; ++a) {
            if (!valid[a]) {
                continue;
            }
            float right_area[kBins];
            uint32_t right_cnt[kBins];
            Box acc;
            acc.reset();
            uint32_t cnt = 0;
            for (int b = kBins - 1; b > 0; --b) {
                acc.grow(bin_box[a][b]);
                cnt += bin_cnt[a][b];
                right_area[b] = cnt ? acc.half_area() : 0.f;
                right_cnt[b] = cnt;
            }
            acc.reset();
            cnt = 0;
            for (int b = 0; b < kBins - 1; ++b) {
                acc.grow(bin_box[a][b]);
                cnt += bin_cnt[a][b];
                if (cnt == 0 || right_cnt[b + 1] == 0) {
                    continue;
                }
                const float cost = acc.half_area() * cnt + right_area[b + 1] * right_cnt[b + 1];
                if (cost < best_cost) {
                    best_cost = cost;
                    best_axis = a;
                    best_bin = b;
                }
            }
        }
        uint32_t mid = first + count / 2;
        if (best_axis >= 0) {
            const int a = best_axis;
            auto it = std::partition(prim_ids.begin() + first, prim_ids.begin() + first + count,
                                     [&](uint32_t p) {
                                         int b = (int)((prim_cent[3 * (size_t)p + a] - cmin[a]) * scale[a]);
                                         b = std::min(std::max(b, 0), kBins - 1);
                                         return b <= best_bin;
                                     });
            const uint32_t m = (uint32_t)(it - prim_ids.begin());
            if (m > first && m < first + count) {
                mid = m;
            }
        }
        return mid;
    }

    void process(Task root_task)
    {
        std::vector<Task> local;
        local.push_back(root_task);
        while (!local.empty()) {
            const Task t = local.back();
            local.pop_back();
            Box box, cbox;
            box.reset();
            cbox.reset();
            for (uint32_t i = t.first; i < t.first + t.count; ++i) {
                const uint32_t p = prim_ids[i];
                box.grow(prim_box[p]);
                cbox.grow_pt(prim_cent + 3 * (size_t)p);
            }
            B2Node &n = nodes[t.node];
            n.box = box;
            if (t.count == 1) {
                n.left = prim_ids[t.first];
                n.count = 1;
                continue;
            }
            const uint32_t mid = split(t, cbox);
            const uint32_t left = next_node.fetch_add(2);
            n.left = left;
            n.count = 0;
            const Task tl{left, t.first, mid - t.first};
            const Task tr{left + 1, mid, t.first + t.count - mid};
            if (tl.count >= kParallelThreshold && tr.count >= kParallelThreshold) {
                push_shared(tr);
                local.push_back(tl);
            } else {
                local.push_back(tr);
                local.push_back(tl);
            }
        }
    }

    void worker()
    {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(mtx);
                cv.wait(lk, [&] { return !shared.empty() || outstanding == 0; });
                if (shared.empty()) {
                    return;
                }
                t = shared.front();
                shared.pop_front();
            }
            process(t);
            bool done;
            {
                std::lock_guard<std::mutex> lk(mtx);
                --outstanding;
                done = outstanding == 0;
            }
            if (done) {
                cv.notify_all();
            }
        }
    }

    void build(uint32_t n, int threads)
    {
        prim_ids.resize(n);
        for (uint32_t i = 0; i < n; ++i) {
            prim_ids[i] = i;
        }
        nodes.resize(2 * (size_t)n - 1);
        push_shared(Task{0, 0, n});
        std::vector<std::thread> pool;
        for (int i = 1; i < threads; ++i) {
            pool.emplace_back([this] { worker(); });
        }
        worker();
        for (auto &th : pool) {
            th.join();
        }
        nodes.resize(next_node.load());
    }
};

// ---------------------------------------------------------------------------------------
// Collapse BVH2 -> BVH8 (Ylitie et al. 2017 §4.1-4.2)
// ---------------------------------------------------------------------------------------
constexpr float kPrimCost = 0.3f;
constexpr float kNodeCost = 1.0f;
enum : uint8_t { kLeaf = 0, kInternal = 1, kDistribute = 2 };

struct Decision {
    uint8_t type : 2;
    uint8_t dist_left : 3;
    uint8_t dist_right : 3;
};

struct Collapser {
    const std::vector<B2Node> &n2;
    std::vector<float> cost;         // 7 per node
    std::vector<Decision> decision;  // 7 per node
    std::vector<uint32_t> tri_count;

    explicit Collapser(const std::vector<B2Node> &nodes) : n2(nodes)
    {
        cost.resize(nodes.size() * 7);
        decision.resize(nodes.size() * 7);
        tri_count.resize(nodes.size());
    }

    // post-order over an explicit stack (SAH trees can be deep)
    void run(uint32_t root)
    {
        std::vector<std::pair<uint32_t, bool>> stack;
        stack.emplace_back(root, false);
        while (!stack.empty()) {
            auto [n, expanded] = stack.back();
            stack.pop_back();
            const B2Node &node = n2[n];
            if (node.count) {
                tri_count[n] = 1;
                const float c = node.box.half_area() * kPrimCost;
                for (int i = 0; i < 7; ++i) {
                    cost[7 * (size_t)n + i] = c;
                    decision[7 * (size_t)n + i] = Decision{kLeaf, 0, 0};
                }
                continue;
            }
            if (!expanded) {
                stack.emplace_back(n, true);
                stack.emplace_back(node.left, false);
                stack.emplace_back(node.left + 1, false);
                continue;
            }
            const uint32_t l = node.left, r = node.left + 1;
            const uint32_t cnt = tri_count[l] + tri_count[r];
            tri_count[n] = cnt;
            const float area = node.box.half_area();
            const float *cl = &cost[7 * (size_t)l], *cr = &cost[7 * (size_t)r];
            float *cn = &cost[7 * (size_t)n];
            Decision *dn = &decision[7 * (size_t)n];
            // i = 0: a single root — either a leaf (<= 3 triangles) or an internal node whose
            // 8 slots are distributed between the two subtrees
            {
                const float cost_leaf = cnt <= 3 ? area * (float)cnt * kPrimCost : std::numeric_limits<float>::infinity();
                float best = std::numeric_limits<float>::infinity();
                int bl = 0, br = 0;
                for (int k = 0; k < 7; ++k) {
                    const float c = cl[k] + cr[6 - k];
                    if (c < best) {
                        best = c;
                        bl = k;
                        br = 6 - k;
                    }
                }
                const float cost_internal = best + area * kNodeCost;
                if (cost_leaf < cost_internal) {
                    cn[0] = cost_leaf;
                    dn[0] = Decision{kLeaf, 0, 0};
                } else {
                    cn[0] = cost_internal;
                    dn[0] = Decision{kInternal, (uint8_t)bl, (uint8_t)br};
                }
            }
            // i = 1..6: a forest of up to i+1 roots
            for (int i = 1; i < 7; ++i) {
                cn[i] = cn[i - 1];
                dn[i] = dn[i - 1];
                for (int k = 0; k < i; ++k) {
                    const float c = cl[k] + cr[i - k - 1];
                    if (c < cn[i]) {
                        cn[i] = c;
                        dn[i] = Decision{kDistribute, (uint8_t)k, (uint8_t)(i - k - 1)};
                    }
                }
            }
        }
    }

    // Children (BVH2 node ids) of the forest (n, i)
    void get_children(uint32_t n, int i, uint32_t *children, int &num) const
    {
        const B2Node &node = n2[n];
        if (node.count) {
            children[num++] = n;
            return;
        }
        const Decision d = decision[7 * (size_t)n + i];
        const uint32_t l = node.left, r = node.left + 1;
        if (decision[7 * (size_t)l + d.dist_left].type == kDistribute) {
            get_children(l, d.dist_left, children, num);
        } else {
            children[num++] = l;
        }
        if (decision[7 * (size_t)r + d.dist_right].type == kDistribute) {
            get_children(r, d.dist_right, children, num);
        } else {
            children[num++] = r;
        }
    }
};

void collect_tris(const std::vector<B2Node> &n2, uint32_t n, std::vector<uint32_t> &out)
{
    std::vector<uint32_t> stack{n};
    while (!stack.empty()) {
        const uint32_t c = stack.back();
        stack.pop_back();
        if (n2[c].count) {
            out.push_back(n2[c].left);
        } else {
            stack.push_back(n2[c].left + 1);
            stack.push_back(n2[c].left);
        }
    }
}

inline float pow2_from_biased(uint8_t e)
{
    const uint32_t bits = (uint32_t)e << 23;
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}

// Writes one BVH8 node from its (<= 8) children.
void emit_node(const std::vector<B2Node> &n2, const Collapser &col, const Box &box, const uint32_t *children,
               int num_children, Bvh8Node &out, uint32_t child_base, uint32_t tri_base,
               std::vector<uint32_t> &tri_order, uint32_t *inner_children, int &num_inner)
{
    // --- slot assignment: slot s should hold the child met first by rays of octant s
    // (bit 2/1/0 set = negative x/y/z direction); greedy minimum of dot(centroid offset, d_s)
    float cx = 0.5f * (box.lo[0] + box.hi[0]), cy = 0.5f * (box.lo[1] + box.hi[1]), cz = 0.5f * (box.lo[2] + box.hi[2]);
    float cst[8][8];
    for (int c = 0; c < num_children; ++c) {
        const Box &b = n2[children[c]].box;
        const float ox = 0.5f * (b.lo[0] + b.hi[0]) - cx, oy = 0.5f * (b.lo[1] + b.hi[1]) - cy,
                    oz = 0.5f * (b.lo[2] + b.hi[2]) - cz;
        for (int s = 0; s < 8; ++s) {
            const float dx = (s & 4) ? -1.f : 1.f, dy = (s & 2) ? -1.f : 1.f, dz = (s & 1) ? -1.f : 1.f;
            cst[c][s] = ox * dx + oy * dy + oz * dz;
        }
    }
    int slot_child[8];
    for (int s = 0; s < 8; ++s) {
        slot_child[s] = -1;
    }
    bool child_done[8] = {false};
    for (int it = 0; it < num_children; ++it) {
        float best = std::numeric_limits<float>::max();
        int bc = -1, bs = -1;
        for (int c = 0; c < num_children; ++c) {
            if (child_done[c]) {
                continue;
            }
            for (int s = 0; s < 8; ++s) {
                if (slot_child[s] >= 0) {
                    continue;
                }
                if (cst[c][s] < best) {
                    best = cst[c][s];
                    bc = c;
                    bs = s;
                }
            }
        }
        child_done[bc] = true;
        slot_child[bs] = bc;
    }

    // --- quantisation frame
    std::memset(&out, 0, sizeof(out));
    float max_ext = 0.f;
    for (int a = 0; a < 3; ++a) {
        out.p[a] = box.lo[a];
        max_ext = std::max(max_ext, box.hi[a] - box.lo[a]);
    }
    float step[3];
    for (int a = 0; a < 3; ++a) {
        // flat axes get a thin but non-zero grid so that every child box has thickness
        float ext = std::max(box.hi[a] - box.lo[a], max_ext * (1.f / 65536.f));
        ext = std::max(ext, 1e-30f);
        int e = (int)std::ceil(std::log2((double)ext / 255.0));
        e = std::min(std::max(e, -100), 100);
        // make sure the far plane of the node is representable: lo + 255*2^e >= hi
        while (e < 100 && (double)box.lo[a] + 255.0 * std::ldexp(1.0, e) < (double)box.hi[a]) {
            ++e;
        }
        out.e[a] = (uint8_t)(e + 127);
        step[a] = pow2_from_biased(out.e[a]);
    }
    out.child_base = child_base;
    out.tri_base = tri_base;

    uint8_t *qlo[3] = {out.qlo_x, out.qlo_y, out.qlo_z};
    uint8_t *qhi[3] = {out.qhi_x, out.qhi_y, out.qhi_z};
    uint32_t tri_off = 0;
    num_inner = 0;
    for (int s = 0; s < 8; ++s) {
        const int c = slot_child[s];
        if (c < 0) {
            out.meta[s] = 0;
            continue;
        }
        const uint32_t cn = children[c];
        const Box &b = n2[cn].box;
        for (int a = 0; a < 3; ++a) {
            const double p = out.p[a], st = step[a];
            int lo = (int)std::floor(((double)b.lo[a] - p) / st);
            int hi = (int)std::ceil(((double)b.hi[a] - p) / st);
            lo = std::min(std::max(lo, 0), 255);
            hi = std::min(std::max(hi, 0), 255);
            // conservative in double arithmetic (p + q*step is exact in double)
            while (lo > 0 && p + lo * st > (double)b.lo[a]) {
                --lo;
            }
            while (hi < 255 && p + hi * st < (double)b.hi[a]) {
                ++hi;
            }
            if (hi <= lo) {  // give flat boxes one grid cell of thickness
                if (hi < 255) {
                    hi = lo + 1;
                } else {
                    lo = hi - 1;
                }
            }
            qlo[a][s] = (uint8_t)lo;
            qhi[a][s] = (uint8_t)hi;
        }
        const bool is_leaf = n2[cn].count || col.decision[7 * (size_t)cn].type == kLeaf;
        if (is_leaf) {
            const size_t before = tri_order.size();
            collect_tris(n2, cn, tri_order);
            const uint32_t k = (uint32_t)(tri_order.size() - before);
            const uint8_t unary = k == 1 ? 0b001 : (k == 2 ? 0b011 : 0b111);
            out.meta[s] = (uint8_t)((unary << 5) | tri_off);
            tri_off += k;
        } else {
            out.meta[s] = (uint8_t)((0b001 << 5) | (24 + s));
            out.imask |= (uint8_t)(1u << s);
            inner_children[num_inner++] = cn;
        }
    }
}

}  // namespace

void build_bvh8(const float *verts, size_t num_tris, int threads, Bvh8 &out)
{
    const auto t0 = std::chrono::steady_clock::now();
    out.nodes.clear();
    out.tri_order.clear();
    if (threads <= 0) {
        threads = std::max(1u, std::thread::hardware_concurrency());
    }
    for (int a = 0; a < 3; ++a) {
        out.scene_lo[a] = 0.f;
        out.scene_hi[a] = 0.f;
    }
    if (num_tris == 0) {
        Bvh8Node root;
        std::memset(&root, 0, sizeof(root));
        root.e[0] = root.e[1] = root.e[2] = 127;
        out.nodes.push_back(root);
        return;
    }
    std::vector<Box> boxes(num_tris);
    std::vector<float> cent(3 * num_tris);
    {
        auto fill = [&](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i) {
                Box bx;
                bx.reset();
                bx.grow_pt(verts + 9 * i);
                bx.grow_pt(verts + 9 * i + 3);
                bx.grow_pt(verts + 9 * i + 6);
                boxes[i] = bx;
                for (int a = 0; a < 3; ++a) {
                    cent[3 * i + a] = 0.5f * (bx.lo[a] + bx.hi[a]);
                }
            }
        };
        std::vector<std::thread> pool;
        const size_t chunk = (num_tris + threads - 1) / threads;
        for (int t = 0; t < threads; ++t) {
            const size_t b = std::min(num_tris, t * chunk), e = std::min(num_tris, (t + 1) * chunk);
            if (b < e) {
                pool.emplace_back(fill, b, e);
            }
        }
        for (auto &th : pool) {
            th.join();
        }
    }
    Bvh2Builder b2;
    b2.prim_box = boxes.data();
    b2.prim_cent = cent.data();
    b2.build((uint32_t)num_tris, threads);
    const std::vector<B2Node> &n2 = b2.nodes;
    for (int a = 0; a < 3; ++a) {
        out.scene_lo[a] = n2[0].box.lo[a];
        out.scene_hi[a] = n2[0].box.hi[a];
    }

    Collapser col(n2);
    col.run(0);

    out.tri_order.reserve(num_tris);
    out.nodes.reserve(num_tris / 2 + 16);
    struct Pending {
        uint32_t b2node;
        uint32_t out_index;
        uint32_t depth;
    };
    std::deque<Pending> queue;
    out.nodes.emplace_back();
    queue.push_back(Pending{0, 0, 1});
    out.sah_cost = 0.0;
    const double root_area = std::max(1e-30f, n2[0].box.half_area());
    while (!queue.empty()) {
        const Pending cur = queue.front();
        queue.pop_front();
        out.max_depth = std::max(out.max_depth, cur.depth);
        uint32_t children[8];
        int num = 0;
        const bool root_is_leaf = n2[cur.b2node].count || col.decision[7 * (size_t)cur.b2node].type == kLeaf;
        if (root_is_leaf) {
            // only possible for the scene root (<= 3 triangles): one leaf child
            children[num++] = cur.b2node;
        } else {
            col.get_children(cur.b2node, 0, children, num);
        }
        uint32_t inner[8];
        int num_inner = 0;
        const uint32_t child_base = (uint32_t)out.nodes.size();
        const uint32_t tri_base = (uint32_t)out.tri_order.size();
        Bvh8Node node;
        emit_node(n2, col, n2[cur.b2node].box, children, num, node, child_base, tri_base, out.tri_order, inner,
                  num_inner);
        out.nodes[cur.out_index] = node;
        out.sah_cost += n2[cur.b2node].box.half_area() / root_area * kNodeCost;
        for (int i = 0; i < num_inner; ++i) {
            out.nodes.emplace_back();
            queue.push_back(Pending{inner[i], child_base + (uint32_t)i, cur.depth + 1});
        }
    }
    out.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace crt
