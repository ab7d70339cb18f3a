// Host-side connectivity generation (GraphSampling) for the vc-conv hot path.
//
// From-scratch restatement of what the reference's offline C++ tool computes:
//   level graph + strided centre sampling ... GraphSampling/meshPooler.h:122-197
//   centre selection (BFS, "no centre within stride-1 hops") ... meshPooler.h:205-308
//   radius-r receptive fields in BFS discovery order ......... meshPooler.h:315-375
//   centre-to-centre graph of the next level ................. meshPooler.h:382-433
//   transposed (unpool) map .................................. meshPooler.h:436-454
//   padded [rows, 1+2N] int32 table layout ................... meshCNN.h:54-92
//   level chaining ........................................... meshCNN.h:28-49
// The tables must be BIT-IDENTICAL to the reference's, including the order of
// the neighbours inside a row (it is the order A[p,n,:] is indexed by), so the
// traversal order is reproduced exactly; what changes is the machinery: flat
// CSR adjacency, O(1) stamp arrays instead of std::set, no vector copies, a C
// ABI, and the layer list as data instead of an edited main.cpp preset.
//
// Plain C++17, no CUDA: the north_star keeps GraphSampling on the host.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vcmesh_sampling.h"

namespace {

struct Entry {
    int id;
    int hop;
};

struct Level {
    int stride = 0, r_pool = 0, r_unpool = 0;
    int n_before = 0;                            // vertices entering the level
    int n_after = 0;                             // centres leaving it
    std::vector<int> centres;                    // old ids, ascending
    std::vector<std::vector<Entry>> pool;        // [n_after] -> (old id, hop)
    std::vector<std::vector<Entry>> unpool;      // [n_before] -> (new centre id, hop)
};

struct Sampler {
    std::vector<std::vector<int>> adj;           // current level graph
    std::vector<int> must;                       // must-be-centre ids, current level numbering
    std::vector<Level> levels;
    std::string err;
    // scratch shared by all BFS calls
    std::vector<int> stamp;
    int tick = 0;
    std::vector<int> frontier, next_frontier;
};

thread_local std::string g_err;

void add_edge(std::vector<std::vector<int>> &adj, int a, int b) {
    // meshPooler.h:36-49: membership is tested on a's list only, then both
    // lists are appended (also for a == b, which appends twice).
    for (int v : adj[a])
        if (v == b) return;
    adj[a].push_back(b);
    adj[b].push_back(a);
}

bool graph_ok(const std::vector<std::vector<int>> &adj) {
    // meshPooler.h:82-117
    const int n = (int)adj.size();
    if (n < 3) return false;
    for (int p = 0; p < n; ++p)
        for (int q : adj[p]) {
            bool back = false;
            for (int r : adj[q])
                if (r == p) { back = true; break; }
            if (!back) return false;
        }
    return true;
}

inline void new_tick(Sampler &s) {
    if (++s.tick == INT32_MAX) {
        std::fill(s.stamp.begin(), s.stamp.end(), 0);
        s.tick = 1;
    }
}

// meshPooler.h:205-241. Every examined edge end is tested against the cover
// state BEFORE the visited test, exactly as the reference does.
bool can_be_centre(Sampler &s, int p, int radius, const std::vector<int> &cover) {
    if (radius <= 0) return true;
    new_tick(s);
    s.stamp[p] = s.tick;
    s.frontier.assign(1, p);
    for (int r = 0; r < radius; ++r) {
        s.next_frontier.clear();
        for (int u : s.frontier)
            for (int v : s.adj[u]) {
                if (cover[v] == 1) return false;
                if (s.stamp[v] != s.tick) {
                    s.stamp[v] = s.tick;
                    s.next_frontier.push_back(v);
                }
            }
        s.frontier.swap(s.next_frontier);
    }
    return true;
}

// meshPooler.h:315-358: ball of `radius` hops around p in discovery order.
void ball(Sampler &s, int p, int radius, std::vector<Entry> &out) {
    out.clear();
    new_tick(s);
    s.stamp[p] = s.tick;
    out.push_back({p, 0});
    s.frontier.assign(1, p);
    for (int r = 0; r < radius; ++r) {
        s.next_frontier.clear();
        for (int u : s.frontier)
            for (int v : s.adj[u])
                if (s.stamp[v] != s.tick) {
                    s.stamp[v] = s.tick;
                    s.next_frontier.push_back(v);
                    out.push_back({v, r + 1});
                }
        s.frontier.swap(s.next_frontier);
    }
}

int add_level(Sampler &s, int stride, int r_pool, int r_unpool) {
    if (stride < 1 || r_pool < 0 || r_unpool < 0) {
        s.err = "stride must be >= 1 and radii >= 0";
        return VCM_GS_EINVAL;
    }
    if (!graph_ok(s.adj)) {
        s.err = "level graph has fewer than 3 vertices or is not symmetric (meshPooler.h:82-117)";
        return VCM_GS_EGRAPH;
    }
    const int n = (int)s.adj.size();
    s.stamp.assign(n, 0);
    s.tick = 0;

    // --- centre selection, meshPooler.h:247-308 ---
    std::vector<int> cover(n, -1);               // -1 unseen, 0 covered, 1 centre
    std::vector<int> queue;
    queue.reserve(n + s.must.size() + 1);
    for (int m : s.must) {
        cover[m] = 1;
        queue.push_back(m);
    }
    if (s.must.empty()) queue.push_back(0);      // seed is NOT marked; it is decided when first reached
    for (size_t head = 0; head < queue.size(); ++head) {
        const int u = queue[head];
        for (int p : s.adj[u]) {
            if (cover[p] >= 0) continue;
            cover[p] = can_be_centre(s, p, stride - 1, cover) ? 1 : 0;
            queue.push_back(p);
        }
    }

    Level L;
    L.stride = stride;
    L.r_pool = r_pool;
    L.r_unpool = r_unpool;
    L.n_before = n;
    std::vector<int> old2new(n, -1);
    for (int i = 0; i < n; ++i)
        if (cover[i] == 1) {
            old2new[i] = (int)L.centres.size();
            L.centres.push_back(i);
        }
    L.n_after = (int)L.centres.size();

    // --- pool map ---
    L.pool.resize(L.n_after);
    for (int c = 0; c < L.n_after; ++c) ball(s, L.centres[c], r_pool, L.pool[c]);

    // --- unpool map = transpose of the radius-r_unpool balls ---
    L.unpool.assign(n, {});
    {
        std::vector<Entry> tmp;
        for (int c = 0; c < L.n_after; ++c) {
            const std::vector<Entry> *b = &L.pool[c];
            if (r_unpool != r_pool) {
                ball(s, L.centres[c], r_unpool, tmp);
                b = &tmp;
            }
            for (const Entry &e : *b) L.unpool[e.id].push_back({c, e.hop});
        }
    }

    // --- next level graph, meshPooler.h:382-433 (self is entry 0) ---
    std::vector<std::vector<int>> next(L.n_after);
    {
        const int r_cc = stride > 1 ? stride + 1 : stride;
        std::vector<Entry> tmp;
        for (int c = 0; c < L.n_after; ++c) {
            ball(s, L.centres[c], r_cc, tmp);
            for (const Entry &e : tmp)
                if (old2new[e.id] >= 0) next[c].push_back(old2new[e.id]);
        }
    }

    // --- must-include ids in the new numbering, meshPooler.h:178-193 ---
    std::vector<int> must_new;
    for (int m : s.must) {
        if (old2new[m] < 0) break;               // cannot happen: must-include vertices are centres by construction
        must_new.push_back(old2new[m]);
    }

    s.adj.swap(next);
    s.must.swap(must_new);
    s.levels.push_back(std::move(L));
    return s.levels.back().n_after;
}

const std::vector<std::vector<Entry>> *pick(const Sampler &s, int level, int unpool, int *pad_id) {
    if (level < 0 || level >= (int)s.levels.size()) return nullptr;
    const Level &L = s.levels[level];
    *pad_id = unpool ? L.n_after : L.n_before;   // meshCNN.h:107-109
    return unpool ? &L.unpool : &L.pool;
}

}  // namespace

extern "C" {

const char *vcm_gs_last_error(void) { return g_err.c_str(); }

int vcm_gs_create(const int32_t *triangles, int n_tri, int n_vert, const int32_t *must_include, int n_must,
                  vcm_gs **out) {
    if (!out || n_vert < 0 || n_tri < 0 || (n_tri > 0 && !triangles)) {
        g_err = "vcm_gs_create: bad arguments";
        return VCM_GS_EINVAL;
    }
    for (int t = 0; t < 3 * n_tri; ++t)
        if (triangles[t] < 0 || triangles[t] >= n_vert) {
            g_err = "vcm_gs_create: triangle index out of range";
            return VCM_GS_EINVAL;
        }
    for (int i = 0; i < n_must; ++i)
        if (must_include[i] < 0 || must_include[i] >= n_vert) {
            g_err = "vcm_gs_create: must-include index out of range";
            return VCM_GS_EINVAL;
        }
    Sampler *s = new Sampler;
    s->adj.resize(n_vert);
    for (int t = 0; t < n_tri; ++t) {            // meshPooler.h:51-63
        const int a = triangles[3 * t], b = triangles[3 * t + 1], c = triangles[3 * t + 2];
        add_edge(s->adj, a, b);
        add_edge(s->adj, b, c);
        add_edge(s->adj, c, a);
    }
    s->must.assign(must_include, must_include + n_must);
    *out = reinterpret_cast<vcm_gs *>(s);
    return VCM_GS_OK;
}

void vcm_gs_destroy(vcm_gs *g) { delete reinterpret_cast<Sampler *>(g); }

int vcm_gs_add_layer(vcm_gs *g, int stride, int radius_pool, int radius_unpool) {
    if (!g) {
        g_err = "vcm_gs_add_layer: null handle";
        return VCM_GS_EINVAL;
    }
    Sampler *s = reinterpret_cast<Sampler *>(g);
    const int r = add_level(*s, stride, radius_pool, radius_unpool);
    if (r < 0) g_err = s->err;
    return r;
}

int vcm_gs_num_levels(const vcm_gs *g) { return g ? (int)reinterpret_cast<const Sampler *>(g)->levels.size() : 0; }

int vcm_gs_table_shape(const vcm_gs *g, int level, int unpool, int *rows, int *cols) {
    int pad;
    const auto *m = g ? pick(*reinterpret_cast<const Sampler *>(g), level, unpool, &pad) : nullptr;
    if (!m || !rows || !cols) {
        g_err = "vcm_gs_table_shape: bad level";
        return VCM_GS_EINVAL;
    }
    size_t nmax = 0;
    for (const auto &r : *m) nmax = r.size() > nmax ? r.size() : nmax;
    *rows = (int)m->size();
    *cols = 1 + 2 * (int)nmax;
    return VCM_GS_OK;
}

int vcm_gs_table_copy(const vcm_gs *g, int level, int unpool, int32_t *dst) {
    int pad, rows, cols;
    if (vcm_gs_table_shape(g, level, unpool, &rows, &cols) != VCM_GS_OK || !dst) {
        g_err = "vcm_gs_table_copy: bad arguments";
        return VCM_GS_EINVAL;
    }
    const auto *m = pick(*reinterpret_cast<const Sampler *>(g), level, unpool, &pad);
    const int nmax = (cols - 1) / 2;
    for (int p = 0; p < rows; ++p) {             // meshCNN.h:69-89
        int32_t *row = dst + (size_t)p * cols;
        const auto &r = (*m)[p];
        row[0] = (int32_t)r.size();
        for (int j = 0; j < nmax; ++j) {
            const bool real = j < (int)r.size();
            row[1 + 2 * j] = real ? r[j].id : pad;
            row[2 + 2 * j] = real ? r[j].hop : -1;
        }
    }
    return VCM_GS_OK;
}

int vcm_gs_centres(const vcm_gs *g, int level, int32_t *dst, int capacity) {
    const Sampler *s = reinterpret_cast<const Sampler *>(g);
    if (!s || level < 0 || level >= (int)s->levels.size()) {
        g_err = "vcm_gs_centres: bad level";
        return VCM_GS_EINVAL;
    }
    const auto &c = s->levels[level].centres;
    if (dst) {
        if (capacity < (int)c.size()) {
            g_err = "vcm_gs_centres: buffer too small";
            return VCM_GS_EINVAL;
        }
        std::memcpy(dst, c.data(), c.size() * sizeof(int32_t));
    }
    return (int)c.size();
}

}  // extern "C"
