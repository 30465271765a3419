/* Deterministic synthetic long-read generator (SURVEY.md §8d shapes; the reference ships none).
 *
 *   genes x isoforms transcripts of uniform-random ACGT, length ~ N(len_mean, len_sd) clipped to [len_min, len_max];
 *   isoform i>0 of a gene = isoform 0 with one internal `exon_skip`-nt block removed (a different block per isoform);
 *   each transcript yields `reads_per_tx` reads: 5'/3' truncation U[0,trunc_max] each, then i.i.d. per-base
 *   substitution / insertion / deletion, optional reverse-complement with probability p_flip,
 *   Phred ~ round(N(14,4)) clipped to [3,40] (+33).  Reads are emitted in a seeded random order.
 *
 * Every transcript and every read draws from its own counter-based stream (splitmix64 of (seed, tx, read)), so the
 * output depends only on the parameters, not on call order.  Used by tests/, bench.py and tools/ only.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t s;
} rng_t;

static inline uint64_t splitmix(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline rng_t rng_make(uint64_t seed, uint64_t a, uint64_t b) {
    uint64_t s = seed;
    uint64_t x = splitmix(&s) ^ (a * 0xD1342543DE82EF95ull);
    uint64_t y = splitmix(&x) ^ (b * 0xA0761D6478BD642Full);
    rng_t r;
    r.s = splitmix(&y);
    return r;
}
static inline uint64_t rng_u64(rng_t *r) { return splitmix(&r->s); }
static inline double rng_unif(rng_t *r) { return (double)(rng_u64(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t rng_below(rng_t *r, uint32_t n) { return (uint32_t)(((rng_u64(r) >> 32) * (uint64_t)n) >> 32); }
static inline double rng_normal(rng_t *r) {
    double u1 = rng_unif(r), u2 = rng_unif(r);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

typedef struct {
    uint32_t n_reads;
    uint64_t total;
    char *bases, *quals;
    uint64_t *offsets; /* n_reads+1 */
    int32_t *truth_tx; /* transcript index of each read */
    uint8_t *truth_rev;
} synth_t;

static const char ALPHA[4] = {'A', 'C', 'G', 'T'};
static inline char comp(char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A'; }

synth_t *synth_create(uint64_t seed, int n_genes, int n_isoforms, int reads_per_tx, double len_mean, double len_sd,
                      int len_min, int len_max, int exon_skip, int trunc_max, double p_sub, double p_ins, double p_del,
                      double p_flip, int shuffle) {
    const int n_tx = n_genes * n_isoforms;
    const uint32_t n_reads = (uint32_t)n_tx * (uint32_t)reads_per_tx;
    synth_t *S = (synth_t *)calloc(1, sizeof(synth_t));
    S->n_reads = n_reads;
    /* transcripts */
    char **tx = (char **)calloc(n_tx, sizeof(char *));
    int *tx_len = (int *)calloc(n_tx, sizeof(int));
    for (int g = 0; g < n_genes; ++g) {
        rng_t r = rng_make(seed, 1, (uint64_t)g);
        int L = (int)lround(len_mean + len_sd * rng_normal(&r));
        if (L < len_min) L = len_min;
        if (L > len_max) L = len_max;
        char *base = (char *)malloc(L);
        for (int i = 0; i < L; ++i) base[i] = ALPHA[rng_below(&r, 4)];
        tx[g * n_isoforms] = base;
        tx_len[g * n_isoforms] = L;
        for (int iso = 1; iso < n_isoforms; ++iso) {
            int skip = exon_skip;
            if (skip > L / 3) skip = L / 3;
            int room = L - skip - 200;
            int at = 100 + (room > 0 ? (int)rng_below(&r, (uint32_t)room) : 0);
            if (at + skip > L) at = L - skip;
            char *v = (char *)malloc(L - skip);
            memcpy(v, base, at);
            memcpy(v + at, base + at + skip, L - skip - at);
            tx[g * n_isoforms + iso] = v;
            tx_len[g * n_isoforms + iso] = L - skip;
        }
    }
    /* emission order */
    uint32_t *order = (uint32_t *)malloc(sizeof(uint32_t) * n_reads);
    for (uint32_t i = 0; i < n_reads; ++i) order[i] = i;
    if (shuffle) {
        rng_t r = rng_make(seed, 2, 0);
        for (uint32_t i = n_reads; i > 1; --i) {
            uint32_t j = rng_below(&r, i);
            uint32_t t = order[i - 1];
            order[i - 1] = order[j];
            order[j] = t;
        }
    }
    /* reads: generate each into a scratch then append */
    uint64_t cap = 0;
    for (int t = 0; t < n_tx; ++t) cap += (uint64_t)reads_per_tx * (uint64_t)(tx_len[t] * 1.25 + 64);
    S->bases = (char *)malloc(cap);
    S->quals = (char *)malloc(cap);
    S->offsets = (uint64_t *)malloc(sizeof(uint64_t) * (n_reads + 1));
    S->truth_tx = (int32_t *)malloc(sizeof(int32_t) * n_reads);
    S->truth_rev = (uint8_t *)malloc(n_reads);
    uint64_t o = 0;
    for (uint32_t e = 0; e < n_reads; ++e) {
        uint32_t id = order[e];
        int t = (int)(id / (uint32_t)reads_per_tx);
        rng_t r = rng_make(seed, 3, (uint64_t)id);
        int L = tx_len[t];
        int c5 = trunc_max > 0 ? (int)rng_below(&r, (uint32_t)trunc_max + 1) : 0;
        int c3 = trunc_max > 0 ? (int)rng_below(&r, (uint32_t)trunc_max + 1) : 0;
        if (c5 + c3 > L - 50) { c5 = 0; c3 = 0; }
        char *dst = S->bases + o;
        uint64_t n = 0;
        const uint64_t lim = (uint64_t)(L * 1.25 + 60);
        for (int i = c5; i < L - c3 && n + 2 < lim; ++i) {
            double u = rng_unif(&r);
            if (u < p_del) continue;
            if (u < p_del + p_ins) dst[n++] = ALPHA[rng_below(&r, 4)];
            char b = tx[t][i];
            if (rng_unif(&r) < p_sub) {
                char nb;
                do nb = ALPHA[rng_below(&r, 4)];
                while (nb == b);
                b = nb;
            }
            dst[n++] = b;
        }
        int flip = rng_unif(&r) < p_flip;
        if (flip) {
            for (uint64_t i = 0; i < n / 2; ++i) {
                char a = dst[i], b = dst[n - 1 - i];
                dst[i] = comp(b);
                dst[n - 1 - i] = comp(a);
            }
            if (n & 1) dst[n / 2] = comp(dst[n / 2]);
        }
        for (uint64_t i = 0; i < n; ++i) {
            int q = (int)lround(14.0 + 4.0 * rng_normal(&r));
            if (q < 3) q = 3;
            if (q > 40) q = 40;
            S->quals[o + i] = (char)(33 + q);
        }
        S->offsets[e] = o;
        S->truth_tx[e] = t;
        S->truth_rev[e] = (uint8_t)flip;
        o += n;
    }
    S->offsets[n_reads] = o;
    S->total = o;
    for (int t = 0; t < n_tx; ++t) free(tx[t]);
    free(tx);
    free(tx_len);
    free(order);
    return S;
}

uint32_t synth_n_reads(const synth_t *S) { return S->n_reads; }
uint64_t synth_total_bases(const synth_t *S) { return S->total; }
void synth_copy(const synth_t *S, char *bases, char *quals, uint64_t *offsets, int32_t *truth_tx, uint8_t *truth_rev) {
    if (bases) memcpy(bases, S->bases, S->total);
    if (quals) memcpy(quals, S->quals, S->total);
    if (offsets) memcpy(offsets, S->offsets, sizeof(uint64_t) * (S->n_reads + 1));
    if (truth_tx) memcpy(truth_tx, S->truth_tx, sizeof(int32_t) * S->n_reads);
    if (truth_rev) memcpy(truth_rev, S->truth_rev, S->n_reads);
}
void synth_free(synth_t *S) {
    if (!S) return;
    free(S->bases);
    free(S->quals);
    free(S->offsets);
    free(S->truth_tx);
    free(S->truth_rev);
    free(S);
}

/* ReadSet.take: reads idx[0..n) of (src, src_off) copied back to back into dst (dst_off = their new offsets, n+1
 * entries, computed by the caller); one memcpy per read instead of one index per base */
void synth_take(const char *src, const uint64_t *src_off, const int64_t *idx, uint64_t n, const uint64_t *dst_off, char *dst) {
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t a = src_off[idx[i]], b = src_off[idx[i] + 1];
        memcpy(dst + dst_off[i], src + a, b - a);
    }
}
