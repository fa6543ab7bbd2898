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

    // Graph::remove_transitive_edges on an edge list that is no longer the device-resident one
    uint64_t transitive_reduce(uint32_t n_nodes, const std::vector<rala_edge_t>& edges, std::vector<uint8_t>& marked_out) {
        uint64_t n_pairs = 0;
        marked_out.assign(edges.size(), 0);
        check(rala_b200_transitive_reduce(ctx_, n_nodes, edges.size(), edges.data(), marked_out.data(), &n_pairs), "transitive_reduce");
        return n_pairs;
    }

    uint64_t kernel_launches() const { return rala_b200_launch_count(ctx_); }

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

}  // namespace rala_b200
