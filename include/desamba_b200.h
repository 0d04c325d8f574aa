/*
 * desamba_b200.h -- C ABI of the B200-native deSAMBA classification hot path.
 *
 * The reference has no plugin/FFI API; the boundary it offers is the internal call
 *     classify_seq(kseq_t *read, DA_IDX *idx, cly_r *results, Classify_buff_pool *buff)      (cly.c:3064)
 * made once per read from kt_for/worker_for (cly_mt.c:369-376,389) on batches of <=5000 reads / 10 Mbases
 * (cly_mt.c:19-20,42-56), around load_idx (idx.c:1103) and the writers (cly_mt.c:60-365).
 * This library replaces exactly that: load the unchanged on-disk index into HBM, classify a BATCH of reads on the
 * GPU, hand back the per-read fields the reference's writers consume.  Plain pointers and sizes only; every entry
 * point returns 0 on success or a negative DSB_E_* code (the reference aborts instead: utils.h:111, utils.c:112-134).
 * There is no CPU fallback: without a CUDA device every compute entry point fails with DSB_E_CUDA.
 */
#ifndef DESAMBA_B200_H
#define DESAMBA_B200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSB_OK            0
#define DSB_E_ARG        -1   /* bad argument */
#define DSB_E_IO         -2   /* index file missing / short */
#define DSB_E_CUDA       -3   /* CUDA runtime error (dsb_last_error() has the text) */
#define DSB_E_NOMEM      -4
#define DSB_E_CAPACITY   -5   /* a per-read capacity (anchors, chains, matches, hits) was exceeded; see dsb_opts */

typedef struct dsb_index dsb_index;   /* index resident in one GPU's HBM  (replaces DA_IDX, idx.h:72-91) */
typedef struct dsb_ctx   dsb_ctx;     /* stream + scratch for batches      (replaces Classify_buff_pool, cly.h:137-158) */

/* classify options: MAP_opt fields that reach the hot path (cly.h:14-24, cly_mt.c:486-503,521-523) + capacities */
typedef struct {
	int32_t  l_min_match;        /* -l, default 170 */
	int32_t  min_score;          /* -s, default 64  */
	uint32_t max_anchors;        /* per-read anchor capacity (reference: unbounded realloc), default 16384 */
	uint32_t max_matches;        /* per-extension 9-mer match capacity (spd_match_set), default 16384 */
	uint32_t max_read_len;       /* longest read accepted in a batch, default 1<<20 */
	uint32_t warps_per_sm;       /* resident classify warps per SM (scratch is per warp), default 16 */
	uint32_t pool_scale_pct;     /* initial size of the per-batch pools (seed tasks, staging, anchors, chains, hits) in % of the
	                                built-in sizing from the batch's bases and reads, default 100; a pool that still overflows is
	                                doubled and the batch re-run inside dsb_batch_download (dsb_batch_retries) */
} dsb_opts;

/* one result line of cly_r.hit (chain_item, cly.h:69-89): exactly the fields the writers read
 * (cly_mt.c:85-98,266-293,303-338) */
typedef struct {
	uint32_t ref_ID;
	uint32_t t_st, t_ed;
	uint32_t q_st, q_ed;
	uint32_t sum_score;
	uint32_t indel;
	uint8_t  direction;          /* FORWARD 1 / REVERSE 0 (utils.h:66-67) */
	uint8_t  primary;            /* 1 PRIMARY, 2 SECONDARY, 3 SUPPLEMENTARY (cly.h:65-67) */
	uint8_t  pri_index;
	uint8_t  pad;
} dsb_hit;

/* per read: cly_r (cly.h:93-100) */
typedef struct {
	uint64_t hit_off;            /* index of the read's first dsb_hit in the hits array */
	uint32_t n_hit;              /* cly_r.hit.n */
	uint32_t n_anchor;           /* cly_r.anchor_v.n (printed by DES/DES_FULL, cly_mt.c:175) */
	uint8_t  fast_classify;      /* cly_r.fast_classify */
	uint8_t  entered_final;      /* read reached delete_small_score_rst with >=1 chain (it then updates max_read_l, cly.c:2958) */
	uint16_t error;              /* 0, or 1 anchors / 2 matches / 3 window / 4 hits: a capacity was hit for this read (no result) */
	uint32_t read_len;
} dsb_read_result;

/* island seeds (CLY_seed, cly.h:27-32) -- exposed for kernel-level parity tests */
typedef struct { uint32_t offset; uint16_t len; uint8_t top; uint8_t pad; } dsb_seed;

typedef struct {               /* reference REF_INFO (idx.h:13-17), for the writers */
	char     name[128];
	uint64_t seq_l, seq_offset;
} dsb_ref_info;

const char *dsb_last_error(void);
const char *dsb_version(void);

/* load_idx + load_bwt (idx.c:1103-1160, bwt.c:68-104): reads <dir>/deSAMBA.{bwt,sa,exki,exk0,exk1,unv,ref_b,ref_i,ref_p}
 * (.acg is not needed on the GPU) and uploads to `device`. */
int  dsb_index_load(const char *dir, int device, dsb_index **out);
/* the same index on another GPU of the box, copied device to device from a loaded one (peer copy over NVLink / NVSwitch):
 * a multi-GPU run reads, uploads and re-cuts the index once */
int  dsb_index_clone(const dsb_index *src, int device, dsb_index **out);
void dsb_index_free(dsb_index *ix);
uint64_t dsb_index_n_ref(const dsb_index *ix);
const dsb_ref_info *dsb_index_ref_info(const dsb_index *ix);          /* host copy of .ref_i */
uint64_t dsb_index_hbm_bytes(const dsb_index *ix);
int  dsb_index_l_ek(const dsb_index *ix);

void dsb_opts_default(dsb_opts *o);
int  dsb_ctx_create(dsb_index *ix, const dsb_opts *o, dsb_ctx **out);
void dsb_ctx_free(dsb_ctx *ctx);
/* allocate every device buffer of the context now for batches of up to max_reads reads / max_bases bases (they grow on demand
 * otherwise -- a cudaMalloc in the middle of a run stalls every stream of the GPU) */
int  dsb_ctx_reserve(dsb_ctx *ctx, uint32_t max_reads, uint64_t max_bases);

/*
 * One batch = the reads of one kt_for call (cly_mt.c:389).  seqs: concatenated read bases (ASCII, no separators),
 * offs[n_reads+1]: start of each read in seqs.  max_read_l_in: Classify_buff_pool.max_read_l carried from the
 * previous batches (cly.c:2958; 0 at start); *max_read_l_out receives the value after this batch, so batches chained
 * in input order reproduce the reference's -t 1 semantics.
 * Results: rr[n_reads]; hits[hits_cap] (hit_off indexes it); *n_hits_out = hits used.  DSB_E_CAPACITY if hits_cap is
 * too small (nothing is lost inside the context: call dsb_batch_download again with a larger array).
 *
 * dsb_classify_batch      = upload + run + download with HOST buffers (the end-to-end call).
 * dsb_batch_upload/run/download = the same three steps separately (run works on inputs resident in HBM).
 * A context holds two input sets: dsb_batch_upload of the NEXT batch may be called while the batch run before is still on the
 * GPU (after its dsb_batch_run, before its dsb_batch_download) -- its reads then travel on a copy stream of their own under the
 * running kernels.  The order per context is  upload(k+1); download(k); run(k+1)  -- one batch ahead, never two; the
 * results of batch k stay readable until run(k+1).  (kt_pipeline's step 0 reading the next batch while step 1 classifies,
 * cly_mt.c:369-398.)
 */
int dsb_classify_batch(dsb_ctx *ctx, const char *seqs, const uint64_t *offs, uint32_t n_reads,
                       int32_t max_read_l_in, int32_t *max_read_l_out,
                       dsb_read_result *rr, dsb_hit *hits, uint64_t hits_cap, uint64_t *n_hits_out);
int dsb_batch_upload(dsb_ctx *ctx, const char *seqs, const uint64_t *offs, uint32_t n_reads);
/* Second (and last) piece of cross-read state of the reference: the capacity of Classify_buff_pool.bin_read (BUFF_REALLOC,
 * utils.h:117-122, cly.c:1241): an alignment that runs 7 bases off the START of a read compares with a byte of that buffer's
 * malloc header, which depends on the capacity.  A context carries the value from batch to batch by itself (= one thread of
 * the reference); a caller that deals consecutive batches to several contexts sets the value valid before the batch
 * (it depends on the read lengths only: capacity = 2*len+20 whenever 2*len exceeds it, for reads >= 40 bp). */
int dsb_ctx_set_bin_capacity(dsb_ctx *ctx, uint32_t m_bin_read);
uint32_t dsb_ctx_bin_capacity(dsb_ctx *ctx);
int dsb_batch_run(dsb_ctx *ctx, int32_t max_read_l_in);
int dsb_batch_download(dsb_ctx *ctx, int32_t *max_read_l_out, dsb_read_result *rr, dsb_hit *hits, uint64_t hits_cap, uint64_t *n_hits_out);
int dsb_batch_sync(dsb_ctx *ctx);
/* re-runs of the last batch after a pool overflow (0 in the normal case) */
int dsb_batch_retries(dsb_ctx *ctx);

/* introspection for tests / profiling (valid after dsb_batch_run + dsb_batch_sync) */
int dsb_batch_get_seeds(dsb_ctx *ctx, uint32_t read, int strand /*0 fwd,1 rev*/, dsb_seed *out, uint32_t cap, uint32_t *n_out, uint32_t *total_score);
/* device time of each kernel of the last dsb_batch_run in ms (CUDA events on the context's stream), up to `cap` values:
 * [0] encode+probe, [1] islands, [2] seed fast, [3] chain, [4] seed slow (strand 0), [5] chain, [6] seed slow (strand 1),
 * [7] chain, [8] score (warp per read), [9] score of the deferred repeat-rich reads (CTA per read), [10] finalize */
int dsb_batch_kernel_ms(dsb_ctx *ctx, float *ms, int cap);
/* stream marks for timing several contexts (batches) in flight on one GPU: dsb_ctx_mark records mark 0 or 1 on the
 * context's stream; dsb_ctx_elapsed_ms waits for b's mark and returns b.mark_b - a.mark_a in ms */
int dsb_ctx_mark(dsb_ctx *ctx, int which);
int dsb_ctx_elapsed_ms(dsb_ctx *a, int mark_a, dsb_ctx *b, int mark_b, float *ms);
/* developer aid: where the batch last run on `ctx` sat on the device's time line, in ms from mark `ref_mark` of context `ref`:
 * t[0] the stream turns to the batch (upload starts), t[1] reads resident / first kernel starts, t[2 .. 12] the 11 kernel groups
 * of dsb_batch_kernel_ms done.  cap >= 13. */
int dsb_batch_timeline(dsb_ctx *ref, int ref_mark, dsb_ctx *ctx, float *t, int cap);
/* work-list sizes of the last run: [0] reads in slow pass 0, [1] in slow pass 1, [2] scored (warp per read), [3] scored again by a CTA
 * each (repeat-rich), [4..6] seed tasks of the fast / slow 0 / slow 1 pass, [7..9] staging chunks of the passes, [10] anchors,
 * [11] chains handed to scoring */
int dsb_batch_work(dsb_ctx *ctx, uint32_t out[12]);
/* number of kernel launches issued by the last dsb_batch_run */
int dsb_batch_launches(dsb_ctx *ctx);
/* CUDA stream handle (cudaStream_t) the context launches on */
void *dsb_ctx_stream(dsb_ctx *ctx);
/* device-side counters of the last dsb_batch_run (algorithmic-byte model):
 * [0] hit slots used, [1] reads taken, [2] first read >= 510 bp that reached the filter, [3] max_read_l of the batch,
 * [4] get_exist_kmer calls on unmasked k-mers (table-0 probes), [5] table-1 probes, [6] prefix-table lookups, [7] occ calls,
 * [8] SA/unitig/ref_pos locates, [9] get_ref calls while seeding, [10] packed reference bytes they cover, [11] reads that hit a
 * capacity, [12] get_ref calls while scoring, [13] their packed bytes */
int dsb_batch_counters(dsb_ctx *ctx, uint64_t out[16]);

/* per-read device time of the last run, out[n_reads][8] in units of 1024 SM cycles:
 * fast seeding, chaining, slow seeding, 9-mer index build, sdp middle, sdp right, sdp left, whole read */
int dsb_batch_profile(dsb_ctx *ctx, uint32_t *out);

/* pinned host memory for the caller's batch buffers (the driver batches reads into these; cly_mt.c:42-56 equivalent) */
int  dsb_host_alloc(size_t bytes, void **out);
void dsb_host_free(void *p);
/* page-lock / release memory the caller allocated itself (e.g. huge-page batch buffers of the driver) */
int  dsb_host_register(void *p, size_t bytes);
void dsb_host_unregister(void *p);
/* Host seconds dsb_classify_batch has spent on this context so far: out[0] upload (per-read layout tables, copies of the reads
 * into the stream), [1] run (kernel launches), [2] download (waiting for the batch, result copies), [3] calls, [4] re-runs
 * after a pool overflow. */
int  dsb_ctx_host_seconds(dsb_ctx *ctx, double *out, int cap);
/* How host threads wait for the GPU: 0 (default) the CUDA runtime's choice (spins while there are more cores than contexts),
 * 1 sleep until the stream is done.  With several contexts per GPU and one host thread each (the driver: up to 6 x 8 threads)
 * spinning takes the cores the FASTQ reader needs.  Takes effect for devices opened afterwards (dsb_index_load / _clone). */
int  dsb_set_sync_mode(int blocking);
/* free / total HBM of a device in bytes */
int  dsb_device_memory(int device, uint64_t *free_bytes, uint64_t *total_bytes);

/* HBM random-gather microbenchmark (roofline denominator): n_gathers random reads of `bytes_each`
 * (1..128, power of two) from a table of table_bytes; returns sector-granular GB/s in *gbs and elapsed ms. */
int dsb_gather_bench(int device, uint64_t table_bytes, uint64_t n_gathers, int bytes_each, double *gbs, double *ms);

#ifdef __cplusplus
}
#endif
#endif
