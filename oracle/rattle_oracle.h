/* TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of RATTLE's two hot paths.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * liboracle.so.  The product (rattle_b200/) never links, imports or executes anything in oracle/.
 *
 * Parity pin: every function here is checked in tests/test_oracle_*.py against
 *   (a) the reference's own functions through oracle/_ref/libref_shim.so (built from the untouched
 *       sources under /root/reference by oracle/Makefile), on the toyset and on seeded synthetic reads, and
 *   (b) committed golden vectors under tests/golden/ generated from that same reference build
 *       (tests/golden/make_golden.py), which travel to the GPU box where /root/reference does not exist.
 */
#ifndef RATTLE_ORACLE_H
#define RATTLE_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_BV_WORDS 64 /* 4096-bit 6-mer presence bitvector as 64 x uint64 (kmer.hpp:14-16) */

/* kmer.hpp:25-40 + kmer.cpp:6-42.  Lists hold len-k entries sorted by (hash,pos); returns len-k, or <0 on
 * input the reference does not survive (len<=k, base outside ACGTU). */
int orc_extract_kmers(const char *seq, int len, int k, int both_strands, uint32_t *fwd_hash, int32_t *fwd_pos,
                      uint32_t *rev_hash, int32_t *rev_pos, uint64_t *bv_fwd, uint64_t *bv_rev);

/* kmer.cpp:45-67.  Returns total number of cross pairs, writes min(total,cap) of them sorted by (first,second). */
int64_t orc_common_kmers(const uint32_t *h1, const int32_t *p1, int n1, const uint32_t *h2, const int32_t *p2,
                         int n2, int32_t *out_first, int32_t *out_second, int64_t cap);

/* similarity.cpp:4-97.  Returns number of distances (written up to dist_cap), *bases = covered bases. */
int orc_similarity(const int32_t *first, const int32_t *second, int64_t n, int k, int *bases, int32_t *distances,
                   int dist_cap);

/* utils.cpp:36-55 (n=0 -> 0, n=1 -> NaN). */
double orc_var(const int32_t *d, int n);

/* cluster.cpp:12-65 on two raw reads: -1 none, 0 forward, 1 reverse. */
int orc_pair_match(const char *s1, int l1, const char *s2, int l2, int k, double t_s, double t_v, double bv_threshold,
                   int is_rna);

/* cluster.cpp:93-259 (incl. get_main_seq :67-91).  reads must already be in visitation order.
 * Outputs: main_id/main_rev [n_clusters], cl_off [n_clusters+1], mem_id/mem_rev [n_reads]. Returns n_clusters or <0. */
int orc_cluster_reads(const char *bases, const uint64_t *offsets, uint32_t n_reads, int k, double t_s, double t_v,
                      double bv_thr, double bv_min, double bv_falloff, double repr_pct, int is_rna, int n_threads,
                      int32_t *main_id, uint8_t *main_rev, int64_t *cl_off, int32_t *mem_id, uint8_t *mem_rev);

/* Work counters of the last orc_cluster_reads call: [0]=bitvector pair tests, [1]=full k-mer comparisons,
 * [2]=accepted pairs, [3]=rounds. */
void orc_cluster_stats(int64_t out[4]);

/* hps codec of cluster_set_t (cluster.hpp:10-42; hps int/uint/vector serializers). Returns bytes written
 * (or needed when cap is too small, negated). gene may be NULL (-1 everywhere). */
int64_t orc_hps_encode(int n_clusters, const int32_t *main_id, const uint8_t *main_rev, const int32_t *main_gene,
                       const int64_t *cl_off, const int32_t *mem_id, const uint8_t *mem_rev, const int32_t *mem_gene,
                       uint8_t *out, int64_t cap);

/* ------------------------------------------------------------------ POA (poa_oracle.cpp) */
/* spoa (kSW, affine) POA of n sequences in the given order, as correct.cpp:395-405 drives it.
 * MSA rows are written back to back (each *msa_cols long). Optional alignment dump like ref_poa_msa. */
int orc_poa_msa(const char *bases, const uint64_t *offsets, uint32_t n, int m, int nn, int g, int e, char *msa_out,
                int64_t cap, int *msa_cols, int64_t *aln_off, int32_t *aln_pairs, int64_t aln_cap);

/* correct.cpp:311-563 end to end (single-threaded, pack order = queue order). FASTQ text outputs. */
int orc_correct_reads(const char *bases, const char *quals, const uint64_t *offsets, uint32_t n_reads,
                      const int32_t *main_id, const uint8_t *main_rev, const int32_t *main_gene, const int64_t *cl_off,
                      const int32_t *mem_id, const uint8_t *mem_rev, const int32_t *mem_gene, int n_clusters,
                      double min_occ, double gap_occ, double err_ratio, int split, int min_reads, char *corrected,
                      int64_t *corrected_len, char *uncorrected, int64_t *uncorrected_len, char *consensi,
                      int64_t *consensi_len);

/* total DP cells (sum over alignments of query_len x graph_nodes) of the last orc_poa_msa/orc_correct_reads call */
int64_t orc_poa_cells(void);

#ifdef __cplusplus
}
#endif
#endif
