/*
 * desamba_oracle.h -- TEST INFRASTRUCTURE ONLY (parity oracle).
 *
 * CPU restatement, in plain C, of the deSAMBA `classify` hot path (reference: /root/reference/src,
 * v1.1.12).  Nothing in the product (desamba_b200/) may include, link or call this; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg use it, as the checker.
 *
 * PINNING: this restatement is pinned against outputs of the reference itself:
 *   - oracle/_ref/deSAMBA_zero (unmodified reference sources built with -ftrivial-auto-var-init=zero, run -t 1)
 *     on the demo (SAM/SAM_FULL/DES/DES_FULL md5 identical to stock -t 4) and on synthetic read sets;
 *   - committed golden text under tests/golden/ produced by oracle/make_golden.sh.
 * The reference ships no tests of its own.
 *
 * Undefined-behaviour policy: see the header of desamba_oracle.c.
 */
#ifndef DESAMBA_ORACLE_H
#define DESAMBA_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint32_t unitig_ID, offset; } orc_sa_t;           /* bwt.h:10-13  */
typedef struct { uint32_t ref_list, length; } orc_unitig_t;        /* idx.h:19-23  */
typedef struct { char name[128]; uint64_t seq_l, seq_offset; } orc_refinfo_t; /* idx.h:13-17 */

typedef struct {
	/* FM index (bwt.c:68-104) */
	uint8_t  *bwt_occ;  uint64_t byteLen;
	uint64_t rank[6];
	uint64_t *hash_index;
	orc_sa_t *sa;       uint64_t sa_size;
	uint64_t dollar_pos;
	/* exist k-mer tables (idx.c:1105-1118, 966-982) */
	uint64_t ek_size, ek_mask; uint8_t l_ek; int single_base_max;
	uint8_t  *ek0, *ek1;
	/* unitigs, reference (idx.c:1120-1159) */
	orc_unitig_t *uni;  uint64_t n_uni;      /* n_uni entries + 1 fabricated sentinel */
	uint8_t  *ref_bin;  uint64_t ref_bin_n;
	orc_refinfo_t *ri;  uint64_t n_ri;
	uint64_t *ref_pos;  uint64_t n_rp;       /* REF_POS bitfield as raw u64: global_offset:40, ref_ID:23, direction:1 */
	/* MAPQ tables (cly_mt.c:413-437) */
	int Q_MEM[2000]; int Q_LV[20][20];
	/* options -l / -s (cly_mt.c:521-523) */
	int filter_min_length, filter_min_score, filter_min_score_LV3;
} orc_index;

typedef struct { uint32_t offset, len; uint8_t top; } orc_seed;    /* cly.h:27-32 */

typedef struct {                                                    /* cly.h:44-61 */
	uint16_t mtch_len; int16_t score; uint8_t left_len, left_ED, rigt_len, rigt_ED;
	uint8_t  direction;
	uint64_t global_offset;
	uint32_t ref_ID, ref_offset, index_in_read;
	int32_t  pre;            /* chain_anchor_pre as index into the anchor array, -1 = NULL */
	uint16_t seed_ID, chain_id;
	uint8_t  anchor_useless, duplicate;
} orc_anchor;

typedef struct {                                                    /* cly.h:69-89 */
	uint32_t ref_ID; int32_t q_t_dis; uint32_t sum_score, anchor_number;
	uint8_t  direction, with_top_anchor, primary, pri_index;
	uint32_t t_st, t_ed, q_st, q_ed, indel, chain_id;
	int32_t  cur;            /* chain_anchor_cur as index, -1 = NULL */
} orc_chain;

typedef struct {
	orc_chain  *hit;  size_t n_hit, m_hit;
	orc_anchor *anc;  size_t n_anc, m_anc;
	uint32_t fast_classify;
	/* introspection for kernel-level parity tests (filled on every call) */
	orc_seed *seeds[2]; uint32_t n_seeds[2]; uint32_t total_score[2]; /* [0]=forward strand, [1]=reverse strand */
	uint32_t m_seeds[2];
	uint32_t entered_final;  /* 1 if the read reached delete_small_score_rst with >=1 chain (updates max_read_l) */
} orc_result;

typedef struct orc_buff orc_buff;   /* per-thread scratch (Classify_buff_pool, cly.h:137-158) */

/* counters for the algorithmic-byte model */
typedef struct {
	uint64_t n_bit0, n_bit1, n_prefix, n_occ, n_locate, n_getref, n_getref_bytes, n_reads, n_bases, n_hits;
} orc_counters;
extern __thread orc_counters orc_cnt;

int  orc_index_load(orc_index *ix, const char *dir);          /* idx.c:1103-1160 + bwt.c:68-104 */
void orc_index_free(orc_index *ix);
void orc_set_opts(orc_index *ix, int l_min_match, int min_score); /* also (re)computes MAPQ tables */
orc_buff *orc_buff_new(void);
void orc_buff_free(orc_buff *b);
void orc_result_free(orc_result *r);
/* cly.c:3064-3132 */
void orc_classify_seq(const orc_index *ix, const char *seq, uint32_t read_len, orc_result *res, orc_buff *buff);

/* writers (cly_mt.c:60-344); fmt: 1 SAM, 2 SAM_FULL, 3 DES, 4 DES_FULL */
void orc_write_result(FILE *out, const orc_index *ix, const orc_result *r, const char *name,
                      const char *seq, const char *qual, uint32_t read_len, int fmt, int max_sec_N);

/* exposed primitives for unit tests */
uint64_t orc_occ(const orc_index *ix, uint64_t r, uint8_t *c);     /* bwt.c:43-65 */
uint64_t orc_hash64_1(uint64_t key);                               /* utils.c:1067-1077 */
uint64_t orc_hash64_2(uint64_t key);                               /* utils.c:1080-1091 */
int      orc_exist_kmer(const orc_index *ix, uint64_t kmer);      /* cly.c:956-972 */
int32_t  orc_lv_extd(uint8_t *ref, int32_t ref_length, uint8_t *query, int32_t query_length); /* cly.c:510-609 */
void     orc_get_ref(const uint8_t *ref_bin, uint8_t *out, int64_t off, int32_t length, int forward); /* cly.c:435-466 */
/* glibc 2.39 qsort == top-down merge sort, left taken when cmp<=0 */
void     orc_msort(void *base, size_t n, size_t size, int (*cmp)(const void *, const void *));

#ifdef __cplusplus
}
#endif
#endif
