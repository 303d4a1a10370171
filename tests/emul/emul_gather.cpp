// emul_gather.cpp — the peer-memory gather of the multi-GPU path (csrc/kernels/gather.inl: k_gather_claim,
// k_gather_copy, k_gather_wait, k_gather_release) on the CPU through cuda_emul.h.  The landing zone is one host
// allocation, "ranks" are plain structs whose kernels run one after the other in a shuffled order per run
// (system-scope atomics are host atomics here).  Self-checking:
//   * several runs back to back: the merged lists are exactly the multiset union of the ranks' lists, every rank's
//     block contiguous, counts and near-tau counts right, epoch parity alternates, release resets the buffer;
//   * a rank whose pass overflowed does not push (check = 1) and the root's wait reports "not yet";
//   * more pairs than the landing zone holds: nothing is written past its end, the count tells the host;
//   * a buffer that was never released makes the next-but-one claim time out (M_PUSHED = 2).
// Test infrastructure (tests/test_emul_gather.py); exit code 0 = all checks passed.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>

#define SELB_EMUL 1
#include "cuda_emul.h"

constexpr int TILE = 128;
#include "../../cuda_selection_criteria_b200/csrc/estimators.cuh"
#include "../../cuda_selection_criteria_b200/csrc/kernels/tiles.inl"       // the meta[] word indices
#include "../../cuda_selection_criteria_b200/csrc/kernels/gather.inl"

namespace {
int g_bad = 0;
#define CHECK(cond, ...) do { if (!(cond)) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); ++g_bad; } } while (0)

struct Rank {
    std::vector<unsigned long long> meta = std::vector<unsigned long long>(M_WORDS, 0);
    std::vector<uint64_t> keys, near_keys;
    std::vector<double> jac, near_j;
    GatherPush push{};
};

struct Zone {
    std::vector<uint8_t> mem;
    GatherZone z{};
    Zone(unsigned long long cap, unsigned long long near_cap) : mem(sizeof(GatherHdr) + 2 * cap * 16 + 2 * near_cap * 16, 0) {
        uint8_t* zb = mem.data();
        z.hdr = (GatherHdr*)zb;
        z.cap = cap;
        z.near_cap = near_cap;
        z.keys = (uint64_t*)(zb + sizeof(GatherHdr));                  // the layout selb200_run derives
        z.jac = (double*)(z.keys + 2 * cap);
        z.near_keys = (uint64_t*)(z.jac + 2 * cap);
        z.near_j = (double*)(z.near_keys + 2 * near_cap);
    }
};

void push(Rank& r, GatherZone z, unsigned epoch, int check, unsigned long long near_cap_local, unsigned grid) {
    emul::launch(1, 32, [&] {
        k_gather_claim(z, epoch, r.meta.data(), check, 1ull << 20, 1ull << 20, 1ull << 20, 1ull << 20, near_cap_local, &r.push);
    });
    emul::launch(grid, 256, [&] {
        k_gather_copy(z, epoch, r.meta.data(), near_cap_local, r.keys.data(), r.jac.data(), r.near_keys.data(), r.near_j.data(), &r.push);
    });
}
}  // namespace

int main() {
    std::mt19937_64 rng(2024);
    const unsigned world = 3;
    const unsigned long long near_cap_local = 64;
    // ---- several runs back to back ----
    {
        Zone zone(20000, 256);
        std::vector<Rank> ranks(world);
        unsigned long long merged[3];
        for (unsigned epoch = 0; epoch < 6; ++epoch) {
            std::vector<std::pair<uint64_t, double>> want, want_near;
            for (unsigned r = 0; r < world; ++r) {
                Rank& R = ranks[r];
                const size_t cnt = (epoch == 3 && r == 1) ? 0 : 1 + rng() % 5000, ncnt = rng() % 65;      // one empty list too; a near list never exceeds its capacity (the host grows it and redoes the pass)
                R.keys.resize(cnt + 1); R.jac.resize(cnt + 1); R.near_keys.resize(ncnt + 1); R.near_j.resize(ncnt + 1);
                for (size_t i = 0; i < cnt; ++i) { R.keys[i] = ((uint64_t)epoch << 56) | ((uint64_t)r << 48) | i; R.jac[i] = (double)(rng() % 1000) / 1000.0; want.push_back({R.keys[i], R.jac[i]}); }
                for (size_t i = 0; i < ncnt; ++i) { R.near_keys[i] = ((uint64_t)0xEE << 56) | ((uint64_t)r << 48) | i; R.near_j[i] = 0.9; if (i < near_cap_local) want_near.push_back({R.near_keys[i], R.near_j[i]}); }
                std::fill(R.meta.begin(), R.meta.end(), 0);
                R.meta[M_OUT] = cnt;
                R.meta[M_NEAR] = ncnt;
            }
            std::vector<unsigned> order(world);
            for (unsigned r = 0; r < world; ++r) order[r] = r;
            std::shuffle(order.begin(), order.end(), rng);
            for (unsigned r : order) push(ranks[r], zone.z, epoch, 1, near_cap_local, 1 + (unsigned)(rng() % 4));
            for (unsigned r = 0; r < world; ++r) CHECK(ranks[r].meta[M_PUSHED] == 1 && ranks[r].push.go == 1, "epoch %u rank %u did not push", epoch, r);
            emul::launch(1, 32, [&] { k_gather_wait(zone.z, epoch, world, &ranks[0].push, merged); });
            const unsigned b = epoch & 1u;
            CHECK(merged[2] == 0 && merged[0] == want.size() && merged[1] == want_near.size(), "epoch %u merged %llu/%llu err %llu, want %zu/%zu", epoch, merged[0], merged[1], merged[2], want.size(), want_near.size());
            std::vector<std::pair<uint64_t, double>> got, got_near;
            for (unsigned long long i = 0; i < merged[0]; ++i) got.push_back({zone.z.keys[b * zone.z.cap + i], zone.z.jac[b * zone.z.cap + i]});
            for (unsigned long long i = 0; i < merged[1]; ++i) got_near.push_back({zone.z.near_keys[b * zone.z.near_cap + i], zone.z.near_j[b * zone.z.near_cap + i]});
            // every rank's block is contiguous and in the rank's own order
            for (unsigned r = 0; r < world; ++r) {
                const unsigned long long base = ranks[r].push.base;
                for (size_t i = 0; i < ranks[r].meta[M_OUT]; ++i)
                    if (got[base + i].first != ranks[r].keys[i]) { CHECK(false, "epoch %u rank %u block broken at %zu", epoch, r, i); break; }
            }
            std::sort(got.begin(), got.end()); std::sort(want.begin(), want.end());
            std::sort(got_near.begin(), got_near.end()); std::sort(want_near.begin(), want_near.end());
            CHECK(got == want, "epoch %u merged list differs", epoch);
            CHECK(got_near == want_near, "epoch %u merged near list differs", epoch);
            emul::launch(1, 32, [&] { k_gather_release(zone.z, epoch); });
            CHECK(zone.z.hdr->count[b] == 0 && zone.z.hdr->near_count[b] == 0 && zone.z.hdr->done[b] == 0 && zone.z.hdr->consumed == epoch + 1, "epoch %u release", epoch);
        }
        printf("runs back to back: %s\n", g_bad ? "FAIL" : "ok");
    }
    // ---- a rank whose pass overflowed does not push; the root's wait says "not yet" ----
    {
        const int before = g_bad;
        Zone zone(1000, 64);
        Rank R;
        R.keys.resize(10); R.jac.resize(10); R.near_keys.resize(1); R.near_j.resize(1);
        R.meta[M_OUT] = 5;
        R.meta[M_CAND] = (1ull << 20) + 1;                  // candidate list overflowed: the host will redo the pass
        push(R, zone.z, 0, 1, near_cap_local, 2);
        unsigned long long merged[3] = {7, 7, 7};
        emul::launch(1, 32, [&] { k_gather_wait(zone.z, 0, 1, &R.push, merged); });
        CHECK(R.push.go == 0 && R.meta[M_PUSHED] == 0 && zone.z.hdr->count[0] == 0 && zone.z.hdr->done[0] == 0 && merged[2] == 2, "overflowed pass was pushed");
        R.meta[M_CAND] = 0;
        R.meta[M_NEAR] = near_cap_local + 1;                 // so did the near-tau list: same rule
        push(R, zone.z, 0, 1, near_cap_local, 2);
        CHECK(R.push.go == 0 && R.meta[M_PUSHED] == 0 && zone.z.hdr->count[0] == 0, "pass with an overflowed near list was pushed");
        R.meta[M_NEAR] = 0;                                  // the redone pass pushes without the check
        push(R, zone.z, 0, 0, near_cap_local, 2);
        emul::launch(1, 32, [&] { k_gather_wait(zone.z, 0, 1, &R.push, merged); });
        CHECK(R.meta[M_PUSHED] == 1 && merged[0] == 5 && merged[2] == 0, "redone pass");
        printf("overflowed pass: %s\n", g_bad == before ? "ok" : "FAIL");
    }
    // ---- more pairs than the landing zone holds ----
    {
        const int before = g_bad;
        Zone zone(100, 8);
        std::vector<Rank> ranks(2);
        for (unsigned r = 0; r < 2; ++r) {
            ranks[r].keys.assign(80, 0x1111111111111111ull * (r + 1)); ranks[r].jac.assign(80, 1.0 + r);
            ranks[r].near_keys.assign(8, 5); ranks[r].near_j.assign(8, 0.5);
            ranks[r].meta[M_OUT] = 80; ranks[r].meta[M_NEAR] = 6;
        }
        const uint64_t canary = 0xC0FFEE0DDF00Dull;
        zone.z.keys[1 * zone.z.cap] = canary;               // first word of the OTHER parity's buffer
        for (unsigned r = 0; r < 2; ++r) push(ranks[r], zone.z, 0, 1, near_cap_local, 3);
        unsigned long long merged[3];
        emul::launch(1, 32, [&] { k_gather_wait(zone.z, 0, 2, &ranks[0].push, merged); });
        CHECK(merged[0] == 160 && merged[1] == 12 && merged[2] == 0, "overflow counts %llu %llu", merged[0], merged[1]);
        CHECK(zone.z.keys[zone.z.cap] == canary, "wrote past the end of the key buffer");
        printf("landing zone too small: %s\n", g_bad == before ? "ok" : "FAIL");
    }
    // ---- never released: the claim of run e+2 times out ----
    {
        const int before = g_bad;
        Zone zone(1000, 64);
        Rank R;
        R.keys.resize(4); R.jac.resize(4); R.near_keys.resize(1); R.near_j.resize(1);
        for (unsigned epoch = 0; epoch < 3; ++epoch) {
            std::fill(R.meta.begin(), R.meta.end(), 0);
            R.meta[M_OUT] = 3;
            push(R, zone.z, epoch, 1, near_cap_local, 1);
            if (epoch < 2) CHECK(R.meta[M_PUSHED] == 1, "epoch %u should push", epoch);
            else CHECK(R.meta[M_PUSHED] == 2 && R.push.go == 0, "epoch 2 should time out, pushed=%llu", R.meta[M_PUSHED]);
        }
        printf("timeout: %s\n", g_bad == before ? "ok" : "FAIL");
    }
    printf(g_bad ? "FAILED (%d)\n" : "all checks passed\n", g_bad);
    return g_bad ? 1 : 0;
}
