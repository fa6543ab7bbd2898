// host/rala_b200.hpp — C++11 host side of the B200 hot path: RAII over the C ABI (include/rala_b200.h).
//
// This header knows nothing about the reference's classes; host/graph_b200.cpp binds it to
// rala::Graph / rala::Overlap / rala::Pile.  Error convention = the reference's own
// (`fprintf(stderr, "[rala::X] error: ...!\n"); exit(1);`, cf. /root/reference/src/graph.cpp:418-421,
// src/overlap.cpp:55-59): a failing C-ABI call prints the library's message under the name of the
// reference function that was running and exits with status 1.  There is no CPU fallback: without an
// sm_100 device Session's constructor fails exactly like that.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../include/rala_b200.h"

namespace rala_b200 {

// Overlap records column-wise, in the layout the kernels read (rala_b200_graph_set_overlaps_columns):
// bit 31 of a_id = invalid record, bit 31 of b_id = orientation.
struct OverlapColumns {
    std::vector<uint32_t> a_id, b_id, a_begin, a_end, b_begin, b_end;
    void reserve(size_t n) {
        a_id.reserve(n); b_id.reserve(n); a_begin.reserve(n); a_end.reserve(n); b_begin.reserve(n); b_end.reserve(n);
    }
};

class Session {
public:
    // `where` names the reference function on whose behalf the calls are made (for error messages)
    explicit Session(const std::string& where, int device = 0) : where_(where) {
        const char* env = getenv("RALA_B200_DEVICE");
        if (env != nullptr) device = atoi(env);
        if (getenv("RALA_B200_DEVICES") != nullptr) {   // a multi-GPU session may follow in this process (see MultiSession)
            setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
            if (env == nullptr) device = atoi(getenv("RALA_B200_DEVICES"));   // the first device of the list
        }
        int rc = rala_b200_create(&ctx_, device);
        if (rc != RALA_B200_OK) {
            fprintf(stderr, "[%s] error: no usable B200 (sm_100) device %d (status %d); "
                "the CUDA path has no CPU fallback!\n", where_.c_str(), device, rc);
            exit(1);
        }
        check(rala_b200_graph_create(ctx_, &graph_), "graph_create");
    }
    ~Session() {
        if (graph_ != nullptr) rala_b200_graph_destroy(graph_);
        if (ctx_ != nullptr) rala_b200_destroy(ctx_);
    }
    Session(const Session&) = delete;
    Session& operator=(const Session&) = delete;

    void where(const std::string& w) { where_ = w; }

    // ---- inputs ---------------------------------------------------------------------------
    void set_overlaps(const OverlapColumns& c) {
        check(rala_b200_graph_set_overlaps_columns(graph_, c.a_id.data(), c.b_id.data(), c.a_begin.data(), c.a_end.data(),
                                                   c.b_begin.data(), c.b_end.data(), c.a_id.size()), "set_overlaps_columns");
        check(rala_b200_synchronize(ctx_), "synchronize");   // the columns may be released by the caller
    }
    void set_overlaps(const std::vector<rala_ovl_t>& records) {
        check(rala_b200_graph_set_overlaps(graph_, records.data(), records.size()), "set_overlaps");
        check(rala_b200_synchronize(ctx_), "synchronize");   // `records` may be released by the caller
    }
    void set_piles(const std::vector<rala_pile_t>& piles, const std::vector<uint8_t>& flags) {
        n_piles_ = static_cast<uint32_t>(piles.size());
        check(rala_b200_graph_set_piles(graph_, piles.data(), flags.empty() ? nullptr : flags.data(), n_piles_), "set_piles");
        check(rala_b200_synchronize(ctx_), "synchronize");
    }
    void set_hills(const std::vector<rala_hill_t>& hills) {
        n_hills_ = static_cast<uint32_t>(hills.size());
        check(rala_b200_graph_set_hills(graph_, hills.data(), n_hills_), "set_hills");
    }

    // ---- stages (reference lines in include/rala_b200.h) -------------------------------------
    void classify() { check(rala_b200_graph_classify(graph_), "classify"); }
    void retrim() { check(rala_b200_graph_retrim(graph_), "retrim"); }
    bool retrim_promote() {
        int changed = 0;
        check(rala_b200_graph_retrim_promote(graph_, &changed), "retrim_promote");
        return changed != 0;
    }
    void finalize() { check(rala_b200_graph_finalize(graph_), "finalize"); }
    void build() { check(rala_b200_graph_build(graph_), "build"); }
    void transitive() { check(rala_b200_graph_transitive(graph_), "transitive"); }

    // ---- outputs --------------------------------------------------------------------------
    rala_b200_counts_t counts() {
        rala_b200_counts_t c;
        check(rala_b200_graph_counts(graph_, &c), "counts");
        return c;
    }
    std::vector<uint32_t> hill_coverage() {
        std::vector<uint32_t> cov(n_hills_);
        check(rala_b200_graph_get_hill_coverage(graph_, cov.data()), "get_hill_coverage");
        return cov;
    }
    std::vector<rala_pile_t> piles() {
        std::vector<rala_pile_t> p(n_piles_);
        if (n_piles_) check(rala_b200_graph_get_piles(graph_, p.data()), "get_piles");
        return p;
    }
    // (a_id, b_id) of every entry of `overlaps` (graph.cpp:740-744)
    std::vector<uint32_t> connections() {
        std::vector<uint32_t> ab(2 * counts().n_overlaps);
        if (!ab.empty()) check(rala_b200_graph_get_connections(graph_, ab.data()), "get_connections");
        return ab;
    }
    std::vector<rala_ovl_t> overlaps() {
        std::vector<rala_ovl_t> o(counts().n_overlaps);
        check(rala_b200_graph_get_lists(graph_, o.empty() ? nullptr : o.data(), nullptr), "get_lists");
        return o;
    }
    void replace_overlaps(const std::vector<rala_ovl_t>& kept) {
        check(rala_b200_graph_set_kept_overlaps(graph_, kept.data(), kept.size()), "set_kept_overlaps");
        check(rala_b200_synchronize(ctx_), "synchronize");
    }
    std::vector<uint32_t> seq_to_node() {
        std::vector<uint32_t> s(n_piles_);
        if (n_piles_) check(rala_b200_graph_get_seq_to_node(graph_, s.data()), "get_seq_to_node");
        return s;
    }
    std::vector<rala_edge_t> edges() {
        std::vector<rala_edge_t> e(counts().n_edges);
        if (!e.empty()) check(rala_b200_graph_get_edges(graph_, e.data()), "get_edges");
        return e;
    }
    std::vector<uint8_t> marked() {
        std::vector<uint8_t> m(counts().n_edges);
        if (!m.empty()) check(rala_b200_graph_get_marked(graph_, m.data()), "get_marked");
        return m;
    }

    // per node the ids of its out- (which = 0) or in-edges (which = 1), ascending, without the removed edges when skip_marked
    void adjacency(int which, bool skip_marked, std::vector<uint32_t>& off, std::vector<uint32_t>& ids) {
        const auto c = counts();
        off.assign(static_cast<size_t>(c.n_nodes) + 1, 0u);
        ids.assign(static_cast<size_t>(c.n_edges) + 1, 0u);
        uint64_t n = 0;
        check(rala_b200_graph_get_adjacency(graph_, which, skip_marked ? 1 : 0, off.data(), ids.data(), &n), "get_adjacency");
        ids.resize(n);
    }

    // Graph::remove_transitive_edges on an edge list that is no longer the device-resident one
    uint64_t transitive_reduce(uint32_t n_nodes, const std::vector<rala_edge_t>& edges, std::vector<uint8_t>& marked_out) {
        uint64_t n_pairs = 0;
        marked_out.assign(edges.size(), 0);
        check(rala_b200_transitive_reduce(ctx_, n_nodes, edges.size(), edges.data(), marked_out.data(), &n_pairs), "transitive_reduce");
        return n_pairs;
    }

    uint64_t kernel_launches() const { return rala_b200_launch_count(ctx_); }

    // device time (ms) of the last run of each stage: classify, retrim, finalize, build, transitive, then single kernels
    std::vector<float> stage_ms() {
        std::vector<float> ms(RALA_B200_N_STAGES, 0.f);
        check(rala_b200_graph_stage_ms(graph_, ms.data()), "stage_ms");
        return ms;
    }

private:
    void check(int rc, const char* what) {
        if (rc == RALA_B200_OK) return;
        fprintf(stderr, "[%s] error: rala_b200 %s failed (status %d): %s!\n", where_.c_str(), what, rc,
            rala_b200_last_error(ctx_));
        exit(1);
    }

    std::string where_;
    rala_b200_ctx* ctx_ = nullptr;
    rala_b200_graph* graph_ = nullptr;
    uint32_t n_piles_ = 0, n_hills_ = 0;
};

// The same hot path on several GPUs (rala_b200_multi_*): one rank per entry of `devices`, records cut into contiguous
// file ranges, the pile table frozen (clean data: no chimeric hills or pits, so no host pile breaking between the passes).
class MultiSession {
public:
    MultiSession(const std::string& where, const std::vector<int>& devices) : where_(where), n_(static_cast<int>(devices.size())) {
        // ranks that share a GPU wait for each other inside kernels: every stream needs its own hardware queue (read by
        // the CUDA runtime when it initialises, i.e. at the first CUDA call of this process, which is the one below)
        setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
        int rc = rala_b200_multi_create(&m_, devices.data(), n_, 0, n_);
        if (rc != RALA_B200_OK) {
            fprintf(stderr, "[%s] error: no usable B200 (sm_100) devices for %d ranks (status %d); "
                "the CUDA path has no CPU fallback!\n", where_.c_str(), n_, rc);
            exit(1);
        }
    }
    ~MultiSession() { if (m_ != nullptr) rala_b200_multi_destroy(m_); }
    MultiSession(const MultiSession&) = delete;
    MultiSession& operator=(const MultiSession&) = delete;

    void where(const std::string& w) { where_ = w; }
    int ranks() const { return n_; }

    void set_piles(const std::vector<rala_pile_t>& piles, const std::vector<uint8_t>& flags) {
        n_piles_ = static_cast<uint32_t>(piles.size());
        check(rala_b200_multi_set_piles(m_, piles.data(), flags.empty() ? nullptr : flags.data(), n_piles_), "set_piles");
    }
    // records [begin, end) of the file go to rank k (contiguous ranges in rank order)
    void set_shard(int k, const OverlapColumns& c, size_t begin, size_t end) {
        check(rala_b200_multi_set_overlaps_columns(m_, k, c.a_id.data() + begin, c.b_id.data() + begin, c.a_begin.data() + begin,
                                                   c.a_end.data() + begin, c.b_begin.data() + begin, c.b_end.data() + begin,
                                                   end - begin, begin), "set_overlaps_columns");
    }
    // sizes the exchange buffers from a first step, then runs classify .. transitive on every rank
    void run() {
        check(rala_b200_multi_plan(m_), "plan");
        check(rala_b200_multi_run(m_), "run");
        check(rala_b200_multi_synchronize(m_), "synchronize");
    }
    rala_b200_multi_counts_t counts() {
        rala_b200_multi_counts_t c;
        check(rala_b200_multi_counts(m_, &c), "counts");
        return c;
    }
    std::vector<rala_pile_t> piles() {
        std::vector<rala_pile_t> p(n_piles_);
        if (n_piles_) check(rala_b200_multi_get_piles(m_, p.data()), "get_piles");
        return p;
    }
    std::vector<uint32_t> seq_to_node() {
        std::vector<uint32_t> s(n_piles_);
        if (n_piles_) check(rala_b200_multi_get_seq_to_node(m_, s.data()), "get_seq_to_node");
        return s;
    }
    // the whole edge list / removed-edge set in edge-id order: every rank holds the ids it emitted
    std::vector<rala_edge_t> edges() {
        std::vector<rala_edge_t> e(counts().n_edges);
        for (int k = 0; k < n_; ++k) {
            uint64_t first = 0, n = 0;
            check(rala_b200_multi_edge_range(m_, k, &first, &n), "edge_range");
            if (n) check(rala_b200_multi_get_edges(m_, k, e.data() + first), "get_edges");
        }
        return e;
    }
    std::vector<uint8_t> marked() {
        std::vector<uint8_t> mk(counts().n_edges);
        for (int k = 0; k < n_; ++k) {
            uint64_t first = 0, n = 0;
            check(rala_b200_multi_edge_range(m_, k, &first, &n), "edge_range");
            if (n) check(rala_b200_multi_get_marked(m_, k, mk.data() + first), "get_marked");
        }
        return mk;
    }
    uint64_t kernel_launches() const { return rala_b200_multi_launch_count(m_); }

private:
    void check(int rc, const char* what) {
        if (rc == RALA_B200_OK) return;
        fprintf(stderr, "[%s] error: rala_b200 multi %s failed (status %d): %s!\n", where_.c_str(), what, rc,
            rala_b200_multi_last_error(m_));
        exit(1);
    }

    std::string where_;
    rala_b200_multi* m_ = nullptr;
    int n_ = 0;
    uint32_t n_piles_ = 0;
};

// RALA_B200_DEVICES=0,1,2,3: the devices of the multi-GPU session (an id may repeat: several ranks on one GPU)
inline std::vector<int> devices_from_env() {
    std::vector<int> d;
    const char* env = getenv("RALA_B200_DEVICES");
    if (env == nullptr) return d;
    std::string s(env);
    size_t pos = 0;
    while (pos <= s.size()) {
        size_t comma = s.find(',', pos);
        if (comma == std::string::npos) comma = s.size();
        if (comma > pos) d.push_back(atoi(s.substr(pos, comma - pos).c_str()));
        pos = comma + 1;
    }
    return d;
}

}  // namespace rala_b200
