// tests/host_event_code.cu — CPU check of rala_b200/csrc/common.cuh's straight-line event_code() (what the first pass
// over the records evaluates on the GPU) against the ORACLE's trim + type (oracle/rala_oracle.c: fp64, the
// reference's own arithmetic, pinned to the reference in tests/test_oracle.py), and of the kernels' trim() /
// classify() against the same oracle.  Host-only program: nvcc compiles the __host__ __device__ functions for the
// CPU; no GPU is touched.  Built and run by tests/test_exact_arith.py.
//
//   host_event_code <cases> <seed>   ->  one JSON line {"cases": .., "mismatches": .., "by_type": [...]}; exit 1 on mismatch
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../oracle/rala_oracle.h"
#include "../rala_b200/csrc/common.cuh"

static uint64_t rng_state;
static inline uint64_t rnd() {   // splitmix64
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint32_t below(uint32_t n) { return n ? (uint32_t) (rnd() % n) : 0u; }
static inline int32_t jitter() {
    static const int32_t J[] = {0, 0, 0, 0, 1, -1, 2, -2, 3, -3, 15, -15, 30, -30, 83, 84, 85, -84, 400, -400, 1000, -1000};
    return J[below(sizeof(J) / sizeof(J[0]))];
}

int main(int argc, char** argv) {
    const uint64_t cases = argc > 1 ? strtoull(argv[1], nullptr, 10) : 2000000ull;
    rng_state = argc > 2 ? strtoull(argv[2], nullptr, 10) : 1ull;
    uint64_t mismatches = 0, by_type[6] = {0, 0, 0, 0, 0, 0};
    for (uint64_t it = 0; it < cases; ++it) {
        // two piles: lengths from tiny to the 2^30 limit, valid regions trimmed at both ends or not
        uint32_t len[2], p0[2], p1[2];
        for (int k = 0; k < 2; ++k) {
            const uint32_t kind = below(8);
            len[k] = kind == 0 ? 100u + below(3000u) : kind == 1 ? (1u << 30) - 1u - below(5000u) : kind == 2 ? 200000u + below(2000000u) : 5000u + below(20000u);
            const uint32_t cut = below(4);
            p0[k] = cut == 0 ? 0u : cut == 1 ? 15u : below(len[k] / 3u + 1u);
            p1[k] = cut == 0 ? len[k] : cut == 1 ? len[k] - 15u : len[k] - below(len[k] / 3u + 1u);
            if (p1[k] <= p0[k]) { p0[k] = 0u; p1[k] = len[k]; }
            if (below(64) == 0) { p0[k] = 0u; p1[k] = 0u; }   // dead pile
        }
        const uint32_t ori = below(2);
        // a geometry in trimmed-read coordinates, then mapped back and jittered
        const uint32_t al = p1[0] > p0[0] ? p1[0] - p0[0] : len[0], bl = p1[1] > p0[1] ? p1[1] - p0[1] : len[1];
        const uint32_t shape = below(8);
        uint32_t a0, a1, b0, b1;
        const uint32_t span = 84u + below((al < bl ? al : bl));
        const uint32_t me = (al > bl ? al : bl) / 20u;
        switch (shape) {
            case 0:  // a's suffix on b's prefix (dovetail)
                a1 = al; a0 = al > span ? al - span : 0u; b0 = 0u; b1 = span < bl ? span : bl; break;
            case 1:  // b's suffix on a's prefix
                b1 = bl; b0 = bl > span ? bl - span : 0u; a0 = 0u; a1 = span < al ? span : al; break;
            case 2:  // a contained in b
                a0 = 0u; a1 = al; b0 = below(bl > al ? bl - al + 1u : 1u); b1 = b0 + al; break;
            case 3:  // b contained in a
                b0 = 0u; b1 = bl; a0 = below(al > bl ? al - bl + 1u : 1u); a1 = a0 + bl; break;
            case 4:  // near containment: begin offsets differ by about min_extension
                a0 = below(al / 2u + 1u); b0 = a0 + me + (uint32_t) jitter(); a1 = a0 + span; b1 = b0 + span + (uint32_t) jitter(); break;
            case 5:  // near containment on the end side
                a1 = al - below(al / 2u + 1u); b1 = bl - (al - a1) - me + (uint32_t) jitter(); a0 = a1 - span; b0 = b1 - span + (uint32_t) jitter(); break;
            case 6:  // internal match with overhangs around the 7/8 threshold
                a0 = below(al / 4u + 1u); b0 = below(bl / 4u + 1u); a1 = a0 + span; b1 = b0 + span;
                if (below(2)) { const uint32_t oh = (a0 < b0 ? a0 : b0); a1 = a0 + 7u * oh + (uint32_t) jitter(); b1 = b0 + (a1 - a0); }
                break;
            default:  // anything
                a0 = below(al + 50u); a1 = below(al + 50u); b0 = below(bl + 50u); b1 = below(bl + 50u);
                if (a0 > a1) { uint32_t t = a0; a0 = a1; a1 = t; }
                if (b0 > b1) { uint32_t t = b0; b0 = b1; b1 = t; }
                break;
        }
        uint32_t rec[7];
        rec[0] = 0u; rec[1] = 1u;
        rec[2] = p0[0] + a0 + (uint32_t) jitter();
        rec[3] = p0[0] + a1 + (uint32_t) jitter();
        const uint32_t fb0 = ori ? bl - b1 : b0, fb1 = ori ? bl - b0 : b1;   // forward-strand coordinates of b
        rec[4] = p0[1] + fb0 + (uint32_t) jitter();
        rec[5] = p0[1] + fb1 + (uint32_t) jitter();
        rec[6] = ori;
        if (below(1024) == 0) rec[2 + below(4)] = (uint32_t) rnd();   // garbage coordinate (wrap-around paths)
        const uint32_t piles[4] = {p0[0], p1[0], p0[1], p1[1]};

        // oracle
        uint32_t r[7];
        for (int k = 0; k < 7; ++k) r[k] = rec[k];
        const int ok = ora_trim(r, piles, 2u);
        const int type = ok ? ora_type(r, piles) : ORA_REJECT;
        const uint32_t want = !ok ? 0u : (1u | (type == ORA_KB ? 2u : 0u) | (type == ORA_KA ? 4u : 0u));
        ++by_type[ok ? type : 5];

        // straight-line form
        const uint32_t got = rb::event_code(rec[2], rec[3], rec[4], rec[5], ori, p0[0], p1[0], p0[1], p1[1]);

        // the kernels' trim() + classify()
        uint32_t got2 = 0u;
        {
            rb::Pile A{p0[0], p1[0], 0u}, B{p0[1], p1[1], 0u};
            rb::Coords c{rec[2], rec[3], rec[4], rec[5]};
            if (A.alive() && B.alive() && rb::trim(c, ori, A, B)) {
                const uint8_t t = rb::classify(c, rb::relative(c, ori, A, B));
                got2 = 1u | (t == rb::kB ? 2u : 0u) | (t == rb::kA ? 4u : 0u);
                if ((int) t != type || c.ab != r[2] || c.ae != r[3] || c.bb != r[4] || c.be != r[5]) got2 = 0xFFu;
            }
        }
        if (got != want || got2 != want) {
            if (mismatches < 10)
                fprintf(stderr, "mismatch: rec %u %u %u %u ori %u piles [%u,%u) [%u,%u): oracle %u event_code %u trim+classify %u\n",
                        rec[2], rec[3], rec[4], rec[5], ori, p0[0], p1[0], p0[1], p1[1], want, got, got2);
            ++mismatches;
        }
    }
    printf("{\"cases\": %llu, \"mismatches\": %llu, \"by_type\": [%llu, %llu, %llu, %llu, %llu, %llu]}\n",
           (unsigned long long) cases, (unsigned long long) mismatches, (unsigned long long) by_type[0],
           (unsigned long long) by_type[1], (unsigned long long) by_type[2], (unsigned long long) by_type[3],
           (unsigned long long) by_type[4], (unsigned long long) by_type[5]);
    return mismatches ? 1 : 0;
}
