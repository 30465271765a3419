// rattle_b200/csrc/cluster_aux_kernels.cuh — the SOURCE of the read-ingest kernels (2-bit packing, length sort) and of the
// greedy waves' bookkeeping kernels (selection, resolution, per-segment seeds, owner assignment) — on the CPU under
// tests/native/cuda_emu.h, against straightforward host code.  Prints "ok" lines; exit code = number of failures.
#define CUDA_EMU_IMPLEMENTATION
#include "cuda_emu.h"

#include <cstdio>
#include <numeric>
#include <random>
#include <string>

#include "../../rattle_b200/csrc/cluster_aux_kernels.cuh"
using namespace rtl;

static int fails = 0;
#define CHECK(cond, what)                         \
    do {                                          \
        if (!(cond)) {                            \
            printf("FAIL: %s\n", what);           \
            ++fails;                              \
        }                                         \
    } while (0)

static void test_pack(std::mt19937 &rng) {
    const char L[5] = {'A', 'C', 'G', 'T', 'U'};
    std::vector<uint64_t> off{0};
    std::string bases;
    for (int r = 0; r < 70; ++r) {
        const int len = 1 + (int)(rng() % 97);
        for (int i = 0; i < len; ++i) bases.push_back(L[rng() % 5]);
        off.push_back(bases.size());
    }
    const uint32_t n = (uint32_t)off.size() - 1;
    for (int bad = 0; bad < 2; ++bad) {
        std::string b = bases;
        if (bad) b[off[n - 3] + 1] = 'N';
        std::vector<uint32_t> pk((b.size() >> 4) + n + 2, 0xdeadbeefu);
        int err = 0;
        emu::launch(3, 64, [&]() { k_pack_bases((const uint8_t *)b.data(), off.data(), n, pk.data(), &err); });
        CHECK((err != 0) == (bad != 0), "k_pack_bases input flag");
        if (bad) continue;
        bool same = true;
        for (uint32_t r = 0; r < n; ++r) {
            const uint64_t w0 = pk_start(off.data(), r);
            if (r + 1 < n) same &= w0 + ((off[r + 1] - off[r] + 15) >> 4) <= pk_start(off.data(), r + 1);  // reads do not overlap
            for (uint64_t p = 0; p < off[r + 1] - off[r]; ++p) {
                const char c = b[off[r] + p];
                const int want = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 3 : 2;  // kmer.hpp:25-31 (T and U -> 2)
                same &= pk_code(pk.data(), w0, (int)p) == want;
            }
        }
        CHECK(same, "k_pack_bases codes");
    }
    printf("ok pack\n");
}

static void test_sort(std::mt19937 &rng, uint32_t n) {
    std::vector<uint64_t> off(n + 1, 0);
    for (uint32_t i = 0; i < n; ++i) off[i + 1] = off[i] + 7 + rng() % 40;  // many equal lengths
    uint32_t n_pad = 2048;
    while (n_pad < n) n_pad <<= 1;
    std::vector<uint64_t> keys(n_pad);
    emu::launch((n_pad + 255) / 256, 256, [&]() { k_sort_keys_init(off.data(), n, n_pad, keys.data()); });
    for (uint32_t size = 2; size <= n_pad; size <<= 1) {  // the launch sequence of cluster_engine.cu: cluster_sort_by_length
        uint32_t stride = size >> 1;
        for (; stride >= 1024; stride >>= 1)
            emu::launch((n_pad / 2 + 255) / 256, 256, [&]() { k_bitonic_step(keys.data(), n_pad, size, stride); });
        emu::launch(n_pad / 2048, 1024, [&]() { k_bitonic_local(keys.data(), size, stride); });
    }
    std::vector<uint32_t> perm(n), want(n);
    emu::launch((n + 255) / 256, 256, [&]() { k_sort_keys_perm(keys.data(), n, perm.data()); });
    std::iota(want.begin(), want.end(), 0u);
    std::stable_sort(want.begin(), want.end(), [&](uint32_t a, uint32_t b) { return off[a + 1] - off[a] > off[b + 1] - off[b]; });
    CHECK(perm == want, "length sort == stable_sort (fasta.cpp:458-464)");
    printf("ok sort %u\n", n);
}

// cluster.cpp:124-166 on a decision matrix: candidate b joins the smallest earlier SEED a with acc[a][b] set
static void host_resolve(const std::vector<uint32_t> &acc, int W, int nc, std::vector<uint8_t> &is_seed, std::vector<int> &found) {
    is_seed.assign(nc, 0);
    found.assign(nc, -1);
    for (int b = 0; b < nc; ++b) {
        for (int a = 0; a < b && found[b] < 0; ++a)
            if (is_seed[a] && acc[(size_t)a * W + b] != 0xffffffffu) found[b] = a;
        is_seed[b] = found[b] < 0;
    }
}

static void test_resolve(std::mt19937 &rng, int W, int nc, double density) {
    std::vector<uint32_t> acc((size_t)W * W, 0xffffffffu);
    for (int a = 0; a < nc; ++a)
        for (int b = a + 1; b < nc; ++b)
            if ((rng() % 10000) < density * 10000) acc[(size_t)a * W + b] = rng() & 1u;
    std::vector<int32_t> cand(W);
    for (int i = 0; i < W; ++i) cand[i] = 1000 + 3 * i;
    std::vector<uint8_t> seed_h;
    std::vector<int> found_h;
    host_resolve(acc, W, nc, seed_h, found_h);
    for (int variant = 0; variant < 2; ++variant) {
        int32_t wave[4] = {0, nc, 0, 0};
        std::vector<int32_t> seed_item(W, -1), owner(1000 + 3 * W + 8, -7);
        std::vector<uint8_t> is_seed(W, 9), owner_rev(owner.size(), 9);
        if (variant == 0)
            emu::launch(1, 32, [&]() { k_resolve(acc.data(), W, cand.data(), wave, seed_item.data(), is_seed.data(), owner.data(), owner_rev.data()); });
        else
            emu::launch(1, 1024, [&]() { k_resolve_cta(acc.data(), W, cand.data(), wave, seed_item.data(), is_seed.data(), owner.data(), owner_rev.data()); });
        bool same = true;
        int ns = 0;
        for (int b = 0; b < nc; ++b) {
            same &= (is_seed[b] != 0) == (seed_h[b] != 0);
            if (seed_h[b]) {
                same &= seed_item[ns++] == cand[b];
            } else {
                same &= owner[cand[b]] == cand[found_h[b]];
                same &= owner_rev[cand[b]] == (uint8_t)acc[(size_t)found_h[b] * W + b];
            }
        }
        same &= wave[2] == ns;
        CHECK(same, variant ? "k_resolve_cta == greedy resolution" : "k_resolve == greedy resolution");
    }
    printf("ok resolve W=%d nc=%d density=%.3f\n", W, nc, density);
}

static void test_select(std::mt19937 &rng) {
    const int M = 5000, W = 64;
    std::vector<uint8_t> taken(M);
    for (auto &t : taken) t = (rng() % 3) == 0;
    for (int cursor : {0, 777, 4990}) {
        int32_t wave[4] = {cursor, 0, 0, 0};
        std::vector<int32_t> cand(W, -1);
        std::vector<uint8_t> tk = taken;
        for (int i = 0; i < cursor; ++i) tk[i] = 1;
        emu::launch(1, 1024, [&]() { k_select(tk.data(), M, W, cand.data(), wave); });
        std::vector<int32_t> want;
        for (int i = cursor; i < M && (int)want.size() < W; ++i)
            if (!tk[i]) want.push_back(i);
        bool same = wave[1] == (int)want.size();
        for (size_t i = 0; i < want.size() && same; ++i) same &= cand[i] == want[i];
        same &= wave[0] == ((int)want.size() == W ? want.back() + 1 : M);
        CHECK(same, "k_select: first W untaken items, cursor behind the last");
        std::vector<int32_t> owner(M, -1);
        std::vector<uint8_t> orev(M, 5);
        emu::launch(2, 128, [&]() { k_mark_cand(tk.data(), cand.data(), wave, owner.data(), orev.data()); });
        bool ok = true;
        for (int c : want) ok &= tk[c] == 1 && owner[c] == c && orev[c] == 0;
        CHECK(ok, "k_mark_cand");
    }
    printf("ok select\n");
}

static void test_select_seg(std::mt19937 &rng) {
    // segments of random sizes; repeated waves must hand out every untaken item exactly once, first-untaken first
    std::vector<int32_t> seg_first{0};
    for (int s = 0; s < 300; ++s) seg_first.push_back(seg_first.back() + (int)(rng() % 9));  // empty segments included
    const int n_seg = (int)seg_first.size() - 1, M = seg_first.back();
    std::vector<uint8_t> taken(M);
    for (auto &t : taken) t = (rng() % 4) == 0;
    std::vector<uint8_t> tk = taken;
    std::vector<int32_t> cur(seg_first.begin(), seg_first.end() - 1), owner(M, -1), handed(M, 0);
    std::vector<uint8_t> orev(M, 3);
    int waves = 0;
    while (true) {
        int32_t wave[4] = {0, 0, 0, 0};
        std::vector<int32_t> seeds(n_seg, -1);
        emu::launch((n_seg + 63) / 64, 64, [&]() { k_select_seg(tk.data(), seg_first.data(), cur.data(), n_seg, seeds.data(), wave, owner.data(), orev.data()); });
        if (wave[2] == 0) break;
        ++waves;
        std::vector<int> per_seg(n_seg, 0);
        for (int k = 0; k < wave[2]; ++k) {
            const int i = seeds[k];
            const int s = (int)(std::upper_bound(seg_first.begin(), seg_first.end(), i) - seg_first.begin()) - 1;
            ++per_seg[s];
            ++handed[i];
            bool first = true;  // no untaken item of the segment before it (items handed out earlier are taken now)
            for (int j = seg_first[s]; j < i; ++j) first &= tk[j] == 1;
            CHECK(first && owner[i] == i && orev[i] == 0 && tk[i] == 1, "k_select_seg: first untaken item of its segment");
        }
        for (int s = 0; s < n_seg; ++s) CHECK(per_seg[s] <= 1, "k_select_seg: one seed per segment and wave");
    }
    bool all = true;
    for (int i = 0; i < M; ++i) all &= handed[i] == (taken[i] ? 0 : 1);
    CHECK(all, "k_select_seg: every untaken item handed out exactly once");
    printf("ok select_seg (%d waves)\n", waves);
}

static void test_apply() {
    const int M = 300;
    std::vector<uint32_t> best(M, 0xffffffffu);
    std::vector<int32_t> seed_item{11, 22, 33}, owner(M, -1);
    std::vector<uint8_t> taken(M, 0), orev(M, 0);
    best[40] = 2 * 1 + 1;
    best[41] = 2 * 2 + 0;
    best[299] = 0;
    best[5] = 2;  // below t0: untouched
    emu::launch(2, 64, [&]() { k_apply(best.data(), 10, M, seed_item.data(), taken.data(), owner.data(), orev.data()); });
    CHECK(owner[40] == 22 && orev[40] == 1 && taken[40] && owner[41] == 33 && orev[41] == 0 && owner[299] == 11 && taken[299] &&
              owner[5] == -1 && !taken[5] && best[40] == 0xffffffffu && best[5] == 2,
          "k_apply");
    printf("ok apply\n");
}

int main() {
    std::mt19937 rng(12345);
    test_pack(rng);
    for (uint32_t n : {1u, 5u, 2048u, 2049u, 5000u}) test_sort(rng, n);
    test_resolve(rng, 64, 64, 0.05);
    test_resolve(rng, 64, 37, 0.5);
    test_resolve(rng, 512, 512, 0.004);
    test_resolve(rng, 1024, 700, 0.01);
    test_select(rng);
    test_select_seg(rng);
    test_apply();
    printf("failures %d\n", fails);
    return fails;
}
