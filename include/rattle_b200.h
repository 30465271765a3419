/* rattle_b200 — C ABI of the B200-native build of RATTLE's two hot paths.
 *
 * RATTLE has no FFI of its own: its boundary is two C++ free functions plus the CLI/file contract
 * (SURVEY.md §8b).  This header is what a re-written `cluster_reads` / `correct_reads` binds:
 *
 *   rtl_cluster_reads   replaces  cluster_reads()            /root/reference/cluster.hpp:44, cluster.cpp:93-259
 *   rtl_correct_reads   replaces  correct_reads()            /root/reference/correct.hpp:44, correct.cpp:311-563
 *   rtl_poa_msa         replaces  spoa align/add_alignment/  /root/reference/correct.cpp:395-405
 *                                 generate_multiple_sequence_alignment as driven by correct.cpp
 *   rtl_hps_encode /    replace   hps::to_stream/from_stream /root/reference/main.cpp:275,397 (clusters.out)
 *   rtl_hps_decode                of cluster_set_t           /root/reference/cluster.hpp:10-42
 *
 * and the function-granularity entry points the parity tests use:
 *
 *   rtl_extract_kmers   replaces  extract_kmers_from_read()  /root/reference/kmer.cpp:6-42
 *   rtl_bv_scan         replaces  the bitvector pre-filter   /root/reference/cluster.cpp:13-19,43
 *   rtl_pair_similarity replaces  get_common_kmers + calc_similarity + var + accept test
 *                                                            /root/reference/kmer.cpp:45-67, similarity.cpp:4-97,
 *                                                            utils.cpp:36-55, cluster.cpp:24-37,48-61
 *
 * Conventions: plain pointers and sizes, caller-owned HOST buffers (the library does its own H2D/D2H),
 * int status (0 = ok, <0 = error; text via rtl_last_error), no exceptions cross the ABI, one rtl_ctx per
 * GPU/process, calls on one ctx must be serialised by the caller.  Everything computes on the GPU: there is no
 * CPU fallback, and rtl_init fails when no CUDA device is usable.
 *
 * Read sets are flat: `bases` = ASCII A/C/G/T/U concatenated, `offsets` = n_reads+1 uint64 starts.
 * Cluster sets are flat: main_id/main_rev[n_clusters], cl_off[n_clusters+1], mem_id/mem_rev[cl_off[n_clusters]]
 * (member order = the order the reference stores in cluster_t::seqs, i.e. what clusters.out holds).
 */
#ifndef RATTLE_B200_H
#define RATTLE_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define RTL_OK 0
#define RTL_ERR_CUDA -1     /* CUDA runtime error (text in rtl_last_error) */
#define RTL_ERR_INPUT -2    /* input the reference does not survive either (len <= k, base outside ACGTU, k > 16) */
#define RTL_ERR_CAPACITY -3 /* caller buffer or device scratch too small */
#define RTL_ERR_STATE -4    /* call out of order (e.g. run before upload) */

#define RTL_BV_WORDS 64 /* 4096-bit 6-mer presence bitvector as 64 x uint64 (kmer.hpp:14-16) */

typedef struct rtl_ctx rtl_ctx;

/* -------------------------------------------------------------------------------- context */
int rtl_init(int device, rtl_ctx **out);
void rtl_destroy(rtl_ctx *ctx);
const char *rtl_last_error(const rtl_ctx *ctx); /* ctx may be NULL: error of the last failed rtl_init */
/* Tunables (all optional): "wave" = candidate seeds evaluated per greedy wave (default 512),
 * "task_cap" = candidate-pair buffer entries, "scratch_mb" = match scratch for oversized pairs,
 * "poa_arena_mb" = device memory for POA score rows + traceback codes, "poa_units" = concurrently running POA
 * units (default 12; both set before the first POA call), "poa_kernel" = 1 forces the int32 POA kernel,
 * "poa_gpu_sort" = 0 keeps the graphs' topological sort and row records on the host (default 1: on the GPU),
 * "poa_mirror_pct" = capacity of the device graph mirrors in percent of the default (tests: forces the host path),
 * "bv_kernel" = 2 sends bitvector scans with at most 16 seeds to the bulk-copy ring kernel (default: the register-staged
 * scan, whose seed tile is staged by bulk copies),
 * "poa_device_chain" = 0 drives every alignment step from the host (default 1: whole per-pack chains run on the GPU).
 * Returns RTL_ERR_INPUT for an unknown key. */
int rtl_set_option(rtl_ctx *ctx, const char *key, int64_t value);

/* Run all work of this ctx on an existing CUDA stream (a cudaStream_t passed as void*; NULL = the ctx's own
 * stream).  bench.py passes torch's current stream so that torch.cuda.Event brackets the library's kernels. */
int rtl_set_stream(rtl_ctx *ctx, void *cuda_stream);

/* counters of the last clustering / correction call on this ctx (for roofline arithmetic and tests) */
typedef struct {
    int64_t bv_pairs;        /* (representative, read) bitvector comparisons the scan kernel evaluated */
    int64_t bv_launches;     /* bv_scan kernel launches */
    double bv_ms;            /* device time inside bv_scan launches (CUDA events on the ctx stream) */
    int64_t full_pairs;      /* candidate (pair,strand) tasks that reached the k-mer join */
    int64_t heavy_pairs;     /* tasks that survived the join-count bound and ran LIS/chain/variance */
    double join_ms;          /* device time, join-count kernel */
    double heavy_ms;         /* device time, LIS kernel */
    double extract_ms;       /* device time, k-mer extraction + sort */
    int64_t waves;           /* greedy waves over all passes */
    int64_t rounds;          /* passes (initial + merge rounds) */
    int64_t kernel_launches; /* all kernels of this library launched by the call */
    int64_t h2d_bytes, d2h_bytes;
    int64_t poa_alignments;  /* sequence-to-graph alignments */
    int64_t poa_cells;       /* sum over alignments of query_len x graph_nodes */
    double poa_ms;           /* device time inside the POA DP+traceback kernel */
    int64_t poa_launches;
    double total_ms;         /* host wall time of the call */
    double poa_wall_ms;      /* host wall time inside the POA phases (kernels of concurrent units overlap, so
                                poa_ms, the sum of their device times, can exceed it) */
    double poa_busy_ms;      /* device time with at least one POA launch group running (union of the groups'
                                CUDA-event intervals): the denominator of the POA roofline */
    double upload_ms;        /* device time of the H2D copy of the read set (CUDA events around it, rtl_cluster_reads /
                                rtl_reads_upload): lets a caller separate the resident part of an end-to-end call */
    int64_t poa_dram_bytes;  /* bytes the POA DP kernels wrote to HBM by construction: traceback codes
                                (rows x strips x 512 B) + spilled score rows + pass hand-over words */
} rtl_stats;
int rtl_get_stats(const rtl_ctx *ctx, rtl_stats *out);

/* -------------------------------------------------------------------------------- hot path A: clustering */
/* One call = cluster_reads(): reads must already be in visitation order (main.cpp:254 sorts by length desc).
 * seq_id values are indices into the passed read set.  Output arrays must hold n_reads (cl_off: n_reads+1). */
int rtl_cluster_reads(rtl_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n_reads, int kmer_size,
                      double t_s, double t_v, double bv_threshold, double min_bv_threshold, double bv_falloff,
                      double repr_percentile, int is_rna, int32_t *main_id, uint8_t *main_rev, int64_t *cl_off,
                      int32_t *mem_id, uint8_t *mem_rev, int32_t *n_clusters);

/* Read ingest on the GPU.  rtl_reads_upload / rtl_cluster_reads pack the uploaded reads to 2 bits per
 * base on the device (kmer.hpp:25-31 codes; the k-mer extraction reads the packed copy) and flag bases outside A,C,G,T,U.
 * rtl_sort_reads_by_length replaces sort_read_set (/root/reference/fasta.cpp:458-464, called at main.cpp:254 before
 * cluster_reads): perm[i] = index of the read that comes i-th in the greedy visitation order — longest first, ties in
 * input order (a stable sort). */
int rtl_sort_reads_by_length(rtl_ctx *ctx, const uint64_t *offsets, uint32_t n_reads, uint32_t *perm);

/* Batched cluster_reads(): the read set is the concatenation of n_seg independent read sets ("segments": reads
 * seg_off[s] .. seg_off[s+1]-1, each already in visitation order), clustered as n_seg separate cluster_reads() calls with
 * the same parameters would — the per-gene loop of `rattle cluster --iso` (/root/reference/main.cpp:281-324, call at :300)
 * — but in ONE pass: one upload, one extraction, and greedy waves that span segments (pairs only exist inside a
 * segment).  Clusters come out segment after segment: those of segment s are seg_cl_off[s] .. seg_cl_off[s+1]-1
 * (seg_cl_off holds n_seg+1 entries), and their seq_ids count from the segment's first read, as the reference's would.
 * Ignores rtl_set_shard: segments are what a multi-GPU caller distributes, not pairs. */
int rtl_cluster_reads_batched(rtl_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n_reads,
                              const uint32_t *seg_off, uint32_t n_seg, int kmer_size, double t_s, double t_v,
                              double bv_threshold, double min_bv_threshold, double bv_falloff, double repr_percentile,
                              int is_rna, int32_t *main_id, uint8_t *main_rev, int64_t *cl_off, int32_t *mem_id,
                              uint8_t *mem_rev, int32_t *n_clusters, int64_t *seg_cl_off);

/* The same in two steps, so that a caller (bench.py `value`) can time the device-resident part alone:
 * upload = H2D of bases/offsets; run = extraction + all passes + D2H of the (small) result. */
int rtl_reads_upload(rtl_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n_reads);
int rtl_cluster_resident(rtl_ctx *ctx, int kmer_size, double t_s, double t_v, double bv_threshold,
                         double min_bv_threshold, double bv_falloff, double repr_percentile, int is_rna,
                         int32_t *main_id, uint8_t *main_rev, int64_t *cl_off, int32_t *mem_id, uint8_t *mem_rev,
                         int32_t *n_clusters);

/* Multi-GPU sharding hook: rank r of `world` evaluates only the (seed,target) pairs whose
 * target index t satisfies t % world == rank, and the caller supplies an exchange callback that min-reduces
 * the per-wave decision arrays across ranks (NCCL allreduce in bench.py / the CLI).  world=1 disables it. */
typedef int (*rtl_allreduce_min_fn)(void *user, void *device_ptr_u32, int64_t count);
int rtl_set_shard(rtl_ctx *ctx, int rank, int world, rtl_allreduce_min_fn fn, void *user);

/* Sharded k-mer extraction (reads block-partitioned over the GPUs, the k-mer lists and bitvectors
 * gathered so that every GPU can hold any representative).  With a broadcast callback set (and world > 1), rank r
 * extracts only reads [r*n/world, (r+1)*n/world) and the ranks exchange their slices of the k-mer lists, bitvectors and
 * popcounts: the callback copies `bytes` at `device_ptr` from rank `root` to the same address range of every other rank
 * (ncclBroadcast in rattle_b200/dist.py), ordered against the ctx stream.  Without it every rank extracts every read. */
typedef int (*rtl_broadcast_fn)(void *user, void *device_ptr, int64_t bytes, int root);
int rtl_set_broadcast(rtl_ctx *ctx, rtl_broadcast_fn fn, void *user);

/* Function-granularity entry points (same data layout as the kernels use). */
/* k-mer lists: for read i the len_i-k entries start at offsets[i]-i*k; sorted by (hash,pos). rev_* and bv_rev
 * may be NULL when both_strands=0. bv_*: n_reads x 64 uint64. */
int rtl_extract_kmers(rtl_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n_reads, int kmer_size,
                      int both_strands, uint32_t *fwd_hash, int32_t *fwd_pos, uint32_t *rev_hash, int32_t *rev_pos,
                      uint64_t *bv_fwd, uint64_t *bv_rev);
/* For every (seed s, target t) pair: common[s*n_targets+t] = popcount(bv[seed]&bv[target]) | rev_common<<16,
 * pass[..] bit0 = forward branch taken, bit1 = reverse branch taken at `bv_threshold` (cluster.cpp:19,43).
 * Operates on the reads of the last rtl_reads_upload + an internal extraction with (kmer_size, !is_rna).
 * common and pass may both be NULL: the scan runs without storing results (kernel timing, rtl_stats.bv_ms). */
int rtl_bv_scan(rtl_ctx *ctx, int kmer_size, int is_rna, const int32_t *seed_reads, int n_seeds,
                const int32_t *target_reads, int n_targets, double bv_threshold, uint32_t *common, uint8_t *pass);
/* For every task (a_read[i], b_read[i], strand[i]): n_common = |get_common_kmers|, bases and n_dist/var from
 * calc_similarity/var, accept = bases/min_len >= t_s && var < t_v.  Tasks rejected by the exact bound
 * kmer_size*n_common/min_len < t_s report bases = -1 (the reference would reject them as well). */
int rtl_pair_similarity(rtl_ctx *ctx, int kmer_size, int is_rna, const int32_t *a_read, const int32_t *b_read,
                        const uint8_t *strand, int64_t n_tasks, double t_s, double t_v, int64_t *n_common,
                        int32_t *bases, int32_t *n_dist, double *var, uint8_t *accept);

/* -------------------------------------------------------------------------------- hot path B: correction */
/* POA of n sequences in the given order with spoa (kSW, affine) scoring m/n/g/e as correct.cpp:395-405 drives it.
 * msa_out receives n rows of *msa_cols chars. Optional alignment dump: aln_off[n+1], aln_pairs[2*aln_cap]
 * as (node, pos) pairs. Returns number of rows or <0. */
int rtl_poa_msa(rtl_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n, int m, int nn, int g, int e,
                char *msa_out, int64_t cap, int *msa_cols, int64_t *aln_off, int32_t *aln_pairs, int64_t aln_cap);

/* correct_reads(): FASTQ text of corrected / uncorrected / consensi is written to the three caller buffers
 * (*_len in: capacity, out: bytes; RTL_ERR_CAPACITY leaves the needed size in *_len).  Headers are "@r<idx>"
 * unless `headers`/`header_off` are given.  Pack order = the reference's single-threaded queue order. */
int rtl_correct_reads(rtl_ctx *ctx, const char *bases, const char *quals, const uint64_t *offsets, uint32_t n_reads,
                      const char *headers, const uint64_t *header_off, const int32_t *main_id,
                      const uint8_t *main_rev, const int32_t *main_gene, const int64_t *cl_off,
                      const int32_t *mem_id, const uint8_t *mem_rev, const int32_t *mem_gene, int n_clusters,
                      double min_occ, double gap_occ, double err_ratio, int split, int min_reads, char *corrected,
                      int64_t *corrected_len, char *uncorrected, int64_t *uncorrected_len, char *consensi,
                      int64_t *consensi_len);

/* File labels of `rattle correct -l a,b,...` (main.cpp:383): the consensus headers then carry per-label read counts
 * ("labels=a:12,b:3,", correct.cpp:447-470,488-512).  Applies to the following rtl_correct_reads calls on this ctx;
 * n_labels = 0 (the default) gives the reference's label-less headers ("labels="). */
int rtl_set_labels(rtl_ctx *ctx, const char *const *labels, int n_labels);

/* Sharded correction (independent clusters, no collective): a rank that corrects a SUBSET of the cluster
 * set passes the subset to rtl_correct_reads and, here, the index each of those clusters has in the whole set — the id
 * the reference writes into every header (",gene_cluster_<cid>", "@gene_cluster_<cid>", "@transcript_cluster_<cid>";
 * correct.cpp:344-349,540-549).  Applies to the following rtl_correct_reads calls, which must pass n_clusters == n;
 * n = 0 (the default) numbers the clusters 0..n_clusters-1. */
int rtl_set_cluster_ids(rtl_ctx *ctx, const int32_t *ids, int n);

/* -------------------------------------------------------------------------------- clusters.out codec */
int64_t rtl_hps_encode(int n_clusters, const int32_t *main_id, const uint8_t *main_rev, const int32_t *main_gene,
                       const int64_t *cl_off, const int32_t *mem_id, const uint8_t *mem_rev,
                       const int32_t *mem_gene, uint8_t *out, int64_t cap);
/* First call with NULL outputs to get sizes (*n_clusters, *n_members), then with buffers. */
int rtl_hps_decode(const uint8_t *buf, int64_t len, int32_t *n_clusters, int64_t *n_members, int32_t *main_id,
                   uint8_t *main_rev, int32_t *main_gene, int64_t *cl_off, int32_t *mem_id, uint8_t *mem_rev,
                   int32_t *mem_gene);

#ifdef __cplusplus
}
#endif
#endif
