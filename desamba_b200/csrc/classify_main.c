/*
 * classify_main.c -- `deSAMBA-b200 classify`: drop-in for `deSAMBA classify` (cly_mt.c:482-562) with the per-read
 * classifier running on B200 GPUs through the C ABI of include/desamba_b200.h.
 *
 * Same options (-h -t -l -r -o -s -f), same index directory, same text output in input order, same stderr summary.
 * What replaces the reference's host (kt_pipeline + kt_for thread pool, cly_mt.c:369-398):
 *   reader thread   FASTQ(.gz) -> batches in host memory (parallel indexer for plain FASTQ) (read_reads, cly_mt.c:42-56; kseq, utils.c:939-977)
 *   one thread/GPU  dsb_classify_batch on its own context; batches are dealt in input order
 *   writer (main)   formats the result records of each batch in input order (output_results, cly_mt.c:350-365)
 * The index is replicated per GPU, reads are sharded by batch, nothing is exchanged between GPUs.
 * Extra options: -g INT GPUs to use [all visible], -c INT contexts (batches in flight) per GPU [up to 6, as HBM allows], -B INT reads per batch
 * [262144], -M INT Mbases per batch [512], -P INT helper threads of the FASTQ reader [cores - 6, at most 48; 0 = serial reader],
 * -A / -m INT per-read anchor / match capacity, -L INT longest read accepted, -p INT pool sizing in % (dsb_opts).
 * -t is accepted and ignored (the thread pool it sized no longer exists).
 *
 * Classify_buff_pool.max_read_l (cly.c:2958) is the reference's only cross-read state; with -t 1 it is the running maximum
 * in input order.  It only matters through the test `max_read_l < 510`, so a batch needs its predecessors' value only
 * when it holds a read < 510 bp while an unfinished earlier batch holds one >= 510 bp; only then does a GPU wait
 * (batch_order.h: the value handed to a batch is built from batches before it in input order only).
 */
#define _GNU_SOURCE
#include "desamba_b200.h"
#include "batch_order.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <pthread.h>
#include <sys/time.h>
#include <sys/resource.h>
#include <zlib.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <sys/mman.h>

enum { FMT_SAM = 1, FMT_SAM_FULL = 2, FMT_DES = 3, FMT_DES_FULL = 4 };

typedef struct {
	int l_min_match, n_threads, max_sec_N, fmt, min_score, n_gpus, ctx_per_gpu;
	uint32_t batch_reads; uint64_t batch_bases;
	FILE *out;
	int n_parse_threads;                    /* helper threads of the FASTQ reader (0: serial reader only) */
	uint32_t max_anchors, max_matches, max_read_len, pool_scale_pct;   /* 0: library default (dsb_opts) */
} opts_t;

/* ---------------------------------------------------------------- batches */
enum { SLOT_FREE = BO_FREE, SLOT_READY = BO_READY, SLOT_BUSY = BO_BUSY, SLOT_DONE = BO_DONE };
/* the buffers of a batch travel as one set: a slot that starts to fill takes the set that was released LAST (its pages are
 * resident and warm) -- rotating through all 2 x contexts x GPUs + 2 slots made nearly every batch of a run fault fresh memory in
 * (measured on 8 GPUs: 98 slots, 120 batches, 4.5 of the reader's 6.8 s in the copies) */
typedef struct {
	char *seqs; size_t m_seqs;              /* ordinary (huge-page) memory, or pinned with DSB_PINNED=1 */
	uint64_t *offs; size_t m_offs;          /* n_reads + 1 */
	char *quals; size_t m_quals;            /* SAM_FULL only, same offsets as seqs */
	char *names; size_t m_names;            /* NUL-terminated, concatenated */
	uint32_t *name_off; size_t m_name_off;
	dsb_read_result *rr; size_t m_rr;
	dsb_hit *hits; size_t m_hits;
} bufset_t;
typedef struct {
	int state; int rc;
	uint64_t seq_no;
	uint32_t n_reads; uint64_t n_bases;
	union {
		bufset_t B;
		struct {
			char *seqs; size_t m_seqs;
			uint64_t *offs; size_t m_offs;
			char *quals; size_t m_quals;
			char *names; size_t m_names;
			uint32_t *name_off; size_t m_name_off;
			dsb_read_result *rr; size_t m_rr;
			dsb_hit *hits; size_t m_hits;
		};
	};
	size_t n_names;
	int has_long, has_short;
	uint32_t m_bin_read_in;                 /* capacity of the reference's bin_read buffer before this batch (dsb_ctx_set_bin_capacity) */
	uint64_t n_hits;
	uint32_t *long_r, *long_l; uint32_t n_long, m_long;   /* reads longer than -L: (index in the batch, true length), ascending; they travel with no bases */
} slot_t;

typedef struct {
	opts_t *o;
	int n_slots; slot_t *slot;
	bufset_t *free_set; int n_free_set;     /* released buffer sets, last in first out (under mu) */
	pthread_mutex_t mu; pthread_cond_t cv;
	uint64_t n_filled, n_claimed, n_written; int eof;
	bo_slot *bo;                            /* per slot: what batch_order.h needs (kept under mu) */
	uint64_t done_upto; int32_t prefix_max; /* every batch below done_upto is finished; max_read_l after them */
	uint64_t n_capacity_reads;              /* reads that exceeded a per-read capacity: written as unclassified, with a warning */
	uint64_t n_long_reads;                  /* reads longer than -L: written as unclassified, with a warning */
	int error; char errmsg[600];
	int n_files; char **files;
	uint64_t total_sequences;
	double t_reader_wait, t_reader_work, t_worker_call, t_worker_wait, t_writer_wait, t_writer_fmt;   /* DSB_VERBOSE: where the host time goes */
	double t_rd_read, t_rd_index, t_rd_copy;   /* parts of t_reader_work: block input (not hidden by the read-ahead), indexing, copies into the batch */
	int max_ahead;                          /* batches the reader may be ahead of the writer */
	int started; char early_msg[4096];      /* "Processing file" lines of the time before "Start classify" */
} shared_t;

typedef struct { shared_t *sh; int gpu; dsb_index *ix; dsb_ctx *ctx; } worker_t;
static double now_s(void);

/* the x-alloc convention of the reference (utils.c:112-134): out of memory ends the run with a message */
static void *xrealloc(void *p, size_t n)
{
	void *q = realloc(p, n ? n : 1);
	if (!q) { fprintf(stderr, "[deSAMBA-b200] out of memory (%zu bytes)\n", n); exit(1); }
	return q;
}
static void *xcalloc(size_t n, size_t sz)
{
	void *q = calloc(n ? n : 1, sz);
	if (!q) { fprintf(stderr, "[deSAMBA-b200] out of memory (%zu x %zu bytes)\n", n, sz); exit(1); }
	return q;
}

static void fail(shared_t *sh, const char *what, int rc)
{
	pthread_mutex_lock(&sh->mu);
	if (!sh->error) { sh->error = rc ? rc : -1; snprintf(sh->errmsg, sizeof sh->errmsg, "%s: %s", what, dsb_last_error()); }
	pthread_cond_broadcast(&sh->cv);
	pthread_mutex_unlock(&sh->mu);
}

#include "fastq_reader.h"

/* batch buffers grow by doubling from 32 MB, and from 128 MB straight to `full` (the batch limit) */
/* The batch buffers are ordinary (huge-page) memory by default: cudaHostAlloc costs ~1 s per GB on an 8-GPU box (19 allocations
 * = 6.9 s for a 4 GB input, more than reading, classifying and writing it), and the H2D copy of a batch from pageable memory
 * goes through the library's pinned staging ring (three threads per batch, the next batch while the current one runs).
 * DSB_PINNED=1 pins them (long multi-GPU runs). */
static double g_t_pinned = 0; static int g_n_pinned = 0; static int g_pageable = 1, g_register = 0, g_host_only = 0;
static int grow_pinned(void **p, size_t *m, size_t need, size_t keep, size_t full)
{
	if (need <= *m) return 0;
	const double t0 = now_s();
	size_t nm = (size_t)32 << 20;
	while (need > nm) nm *= 2;
	if ((nm > full || nm >= ((size_t)128 << 20)) && need <= full) nm = full;       /* a batch of long reads will fill up: no more copies */
	void *np = NULL;
	if (g_pageable) {
		if (posix_memalign(&np, (size_t)2 << 20, nm)) return -1;
		madvise(np, nm, MADV_HUGEPAGE);
		if (g_register && nm >= ((size_t)8 << 20) && dsb_host_register(np, nm) != DSB_OK) g_register = 0;   /* DSB_REGISTER=1: page-lock the huge pages */
	} else if (dsb_host_alloc(nm, &np) != DSB_OK) return -1;
	if (*p) { memcpy(np, *p, keep); if (g_pageable) { if (g_register) dsb_host_unregister(*p); free(*p); } else dsb_host_free(*p); }
	*p = np; *m = nm;
	g_t_pinned += now_s() - t0; g_n_pinned++;
	return 0;
}

/* A read longer than the longest the contexts accept (-L, dsb_opts.max_read_len) does not end the run: it goes into its batch
 * with no bases (the GPU sees an empty read), is written as unclassified with its true length, and is counted for a warning. */
static uint32_t g_max_read_len = 1u << 20;
static void slot_note_long(slot_t *b, uint32_t r, uint32_t L)
{
	if (b->n_long == b->m_long) { b->m_long = b->m_long ? b->m_long * 2 : 16; b->long_r = xrealloc(b->long_r, b->m_long * 4); b->long_l = xrealloc(b->long_l, b->m_long * 4); }
	b->long_r[b->n_long] = r; b->long_l[b->n_long] = L; b->n_long++;
}
static inline uint32_t slot_true_len(const slot_t *b, uint32_t r, uint32_t L, int *is_long)
{
	*is_long = 0;
	if (!b->n_long) return L;
	uint32_t lo = 0, hi = b->n_long;
	while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (b->long_r[mid] < r) lo = mid + 1; else hi = mid; }
	if (lo < b->n_long && b->long_r[lo] == r) { *is_long = 1; return b->long_l[lo]; }
	return L;
}

/* one record into the batch (serial path) */
static int slot_add(shared_t *sh, slot_t *b, const char *name, size_t n_name, const char *seq, const char *qual, size_t n_qual, size_t L)
{
	opts_t *o = sh->o;
	if (grow_pinned((void **)&b->seqs, &b->m_seqs, b->n_bases + L + 16, b->n_bases, o->batch_bases + ((size_t)4 << 20)) ||
	    grow_pinned((void **)&b->offs, &b->m_offs, ((size_t)b->n_reads + 2) * 8, ((size_t)b->n_reads + 1) * 8, ((size_t)o->batch_reads + 2) * 8)) return -1;
	if (b->n_reads == 0) b->offs[0] = 0;
	memcpy(b->seqs + b->n_bases, seq, L);
	if (o->fmt == FMT_SAM_FULL) {
		if (b->n_bases + L + 1 > b->m_quals) { b->m_quals = (b->n_bases + L + 1) * 2; b->quals = xrealloc(b->quals, b->m_quals); }
		if (n_qual == L) memcpy(b->quals + b->n_bases, qual, L); else memset(b->quals + b->n_bases, '*', L);
	}
	if ((size_t)b->n_reads + 1 > b->m_name_off) { b->m_name_off = ((size_t)b->n_reads + 1) * 2; b->name_off = xrealloc(b->name_off, b->m_name_off * 4); }
	if (b->n_names + n_name + 1 > b->m_names) { b->m_names = (b->n_names + n_name + 1) * 2; b->names = xrealloc(b->names, b->m_names); }
	b->name_off[b->n_reads] = (uint32_t)b->n_names;
	memcpy(b->names + b->n_names, name, n_name); b->names[b->n_names + n_name] = 0; b->n_names += n_name + 1;
	b->n_bases += L; b->n_reads++; b->offs[b->n_reads] = b->n_bases;
	if (L >= 510) b->has_long = 1; else b->has_short = 1;
	return 0;
}

/* records [i0, i1) of an indexed block into the batch: offsets serially, bytes by the helper threads */
typedef struct { const char *map; const fq_rec_t *r; size_t i0, i1; slot_t *b; uint32_t first_read; int quals; } copy_job_t;
static void *copy_thread(void *a)
{
	copy_job_t *j = (copy_job_t *)a; slot_t *b = j->b;
	for (size_t i = j->i0; i < j->i1; i++) {
		const fq_rec_t *r = j->r + i;
		const uint32_t k = j->first_read + (uint32_t)(i - j->i0);
		memcpy(b->seqs + b->offs[k], j->map + r->seq, r->n_seq);
		if (j->quals) memcpy(b->quals + b->offs[k], j->map + r->qual, r->n_seq);
		char *nm = b->names + b->name_off[k];
		memcpy(nm, j->map + r->name, r->n_name); nm[r->n_name] = 0;
	}
	return NULL;
}
static int slot_add_block(shared_t *sh, slot_t *b, const char *map, const fq_rec_t *r, size_t i0, size_t i1, int n_thr)
{
	opts_t *o = sh->o;
	uint64_t bases = 0, names = 0;
	for (size_t i = i0; i < i1; i++) { bases += r[i].n_seq; names += r[i].n_name + 1; }
	const size_t n = i1 - i0;
	if (grow_pinned((void **)&b->seqs, &b->m_seqs, b->n_bases + bases + 16, b->n_bases, o->batch_bases + ((size_t)4 << 20)) ||
	    grow_pinned((void **)&b->offs, &b->m_offs, ((size_t)b->n_reads + n + 2) * 8, ((size_t)b->n_reads + 1) * 8, ((size_t)o->batch_reads + 2) * 8)) return -1;
	if (o->fmt == FMT_SAM_FULL && b->n_bases + bases + 1 > b->m_quals) { b->m_quals = (b->n_bases + bases + 1) * 2; b->quals = xrealloc(b->quals, b->m_quals); }
	if ((size_t)b->n_reads + n > b->m_name_off) { b->m_name_off = ((size_t)b->n_reads + n) * 2; b->name_off = xrealloc(b->name_off, b->m_name_off * 4); }
	if (b->n_names + names > b->m_names) { b->m_names = (b->n_names + names) * 2; b->names = xrealloc(b->names, b->m_names); }
	if (b->n_reads == 0) b->offs[0] = 0;
	const uint32_t first = b->n_reads;
	for (size_t i = i0; i < i1; i++) {
		const uint32_t L = r[i].n_seq;
		b->name_off[b->n_reads] = (uint32_t)b->n_names; b->n_names += r[i].n_name + 1;
		b->n_bases += L; b->n_reads++; b->offs[b->n_reads] = b->n_bases;
		if (L >= 510) b->has_long = 1; else b->has_short = 1;
	}
	if (n_thr > 64) n_thr = 64;
	if ((size_t)n_thr > n) n_thr = (int)n;
	if (n_thr < 1) n_thr = 1;
	copy_job_t job[64]; pthread_t th[64];
	/* shares of about the same number of bases */
	size_t at = i0; uint64_t done = 0;
	for (int t = 0; t < n_thr; t++) {
		const uint64_t upto = bases * (uint64_t)(t + 1) / (uint64_t)n_thr;
		size_t e = at;
		while (e < i1 && (done < upto || t == n_thr - 1)) { done += r[e].n_seq; e++; }
		if (t == n_thr - 1) e = i1;
		job[t].map = map; job[t].r = r; job[t].i0 = at; job[t].i1 = e; job[t].b = b; job[t].first_read = first + (uint32_t)(at - i0); job[t].quals = (o->fmt == FMT_SAM_FULL);
		at = e;
	}
	for (int t = 1; t < n_thr; t++) pthread_create(&th[t], NULL, copy_thread, &job[t]);
	copy_thread(&job[0]);
	for (int t = 1; t < n_thr; t++) pthread_join(th[t], NULL);
	return 0;
}

static uint64_t g_fq_block = (uint64_t)256 << 20;   /* bytes of a plain FASTQ file indexed at a time (DSB_FQ_BLOCK_MB: tests use small blocks) */
#define FQ_BLOCK g_fq_block
#define FQ_MARGIN ((uint64_t)16 << 20)     /* read beyond the block: the records that start in it must be complete */
/* Block input.  Default: the block is MAPPED -- the page-cache / tmpfs pages themselves, no copy into a buffer of our own -- by
 * the read-ahead thread while the records of the current block are copied into batches; the helper threads fill in its page
 * tables, a share each (one thread populating alone was as slow as the copy it replaces), and the read-ahead thread also unmaps
 * the block before.  Measured on the 16-core box, 32 GB of FASTQ in tmpfs: reader 1.78 s against 2.00 s with pread under GPU
 * load, 1.45-1.49 s against 1.53-1.56 s alone.  DSB_FQ_MMAP=0 (or a failed mmap) reads the block with pread into one of two
 * reusable buffers instead (an input file that is truncated while it is mapped ends the run with SIGBUS). */
static int g_fq_mmap = 1;
typedef struct { const char *map; void *mbase; size_t mlen; int rc; } block_t;     /* map[file offset] for the offsets of the block */
typedef struct { int fd, n_thr, active; uint64_t pos, len; char *buf; block_t blk, old; pthread_t th; } prefetch_t;   /* old: the block before, unmapped by the read-ahead thread */
typedef struct { char *p; size_t len; } populate_job_t;
static void *populate_thread(void *a)
{
	populate_job_t *j = (populate_job_t *)a;
#ifdef MADV_POPULATE_READ
	if (madvise(j->p, j->len, MADV_POPULATE_READ) == 0) return NULL;
#endif
	volatile char sink = 0;
	for (size_t i = 0; i < j->len; i += 4096) sink += j->p[i];       /* older kernels: a read fault per page */
	(void)sink;
	return NULL;
}
static void block_get(int fd, uint64_t pos, uint64_t len, char *buf, int n_thr, block_t *o)
{
	o->map = NULL; o->mbase = NULL; o->mlen = 0; o->rc = -1;
	if (g_fq_mmap) {
		const uint64_t moff = pos & ~(uint64_t)4095;
		const size_t mlen = (size_t)(pos + len - moff);
		void *m = mmap(NULL, mlen, PROT_READ, MAP_PRIVATE, fd, (off_t)moff);
		if (m != MAP_FAILED) {
			/* page tables of the block filled in by the helper threads, a share each (one thread populating alone was as slow as
			 * the copy it replaces) */
			populate_job_t job[64]; pthread_t th[64];
			int n = n_thr < 1 ? 1 : n_thr > 64 ? 64 : n_thr;
			const size_t span = ((mlen + n - 1) / n + 4095) & ~(size_t)4095;
			int k = 0;
			for (size_t at = 0; at < mlen && k < 64; at += span, k++) { job[k].p = (char *)m + at; job[k].len = at + span <= mlen ? span : mlen - at; }
			for (int t = 1; t < k; t++) pthread_create(&th[t], NULL, populate_thread, &job[t]);
			populate_thread(&job[0]);
			for (int t = 1; t < k; t++) pthread_join(th[t], NULL);
			o->mbase = m; o->mlen = mlen; o->map = (const char *)m - moff; o->rc = 0; return;
		}
	}
	if (buf && fq_read_block(fd, pos, len, buf, n_thr) == 0) { o->map = buf - pos; o->rc = 0; }
}
static void block_release(block_t *b) { if (b->mbase) munmap(b->mbase, b->mlen); b->mbase = NULL; b->map = NULL; b->mlen = 0; }
static void *prefetch_main(void *a) { prefetch_t *p = (prefetch_t *)a; block_get(p->fd, p->pos, p->len, p->buf, p->n_thr, &p->blk); block_release(&p->old); return NULL; }
static void *reader_main(void *arg)
{
	shared_t *sh = (shared_t *)arg; opts_t *o = sh->o;
	stream_t st; memset(&st, 0, sizeof st); st.buf = xrealloc(NULL, SBUF); st.own_buf = st.buf;
	/* .gz files of the command line are inflated ahead by threads of their own (fastq_reader.h): the one being parsed and the
	 * next GZ_AHEAD - 1 */
	gzq_t **gzq = xcalloc(sh->n_files, sizeof *gzq);
	int gz_next = 0;                                   /* files below gz_next have been looked at */
	/* inflating a file takes ~4 x as long as parsing it: 8 in flight hide it (measured on 16 files: 2.1 s with one stream, 1.7 s
	 * with 4 ahead, 0.5 - 0.65 s with 8); each queue holds at most 128 MB */
	const int gz_ahead = getenv("DSB_GZ_AHEAD") ? atoi(getenv("DSB_GZ_AHEAD")) : (o->n_parse_threads > 0 ? 8 : 0);
	rec_t rec; memset(&rec, 0, sizeof rec);
	int file_i = 0, stream_open = 0, pending = 0;      /* pending: rec holds a record not yet stored */
	uint32_t m_bin_read = 0;                           /* running BUFF_REALLOC capacity over all reads, in input order */
	long plen = 0;
	/* plain 4-line FASTQ: read and indexed a block at a time by the helper threads (fastq_reader.h) */
	const char *map = NULL; uint64_t map_size = 0, map_pos = 0;           /* map = blockbuf - (offset of the block): file offsets index it */
	char *blockbuf2[2] = {NULL, NULL}; int cur_buf = 0;      /* pread mode: two reusable buffers */
	block_t cur, stale; memset(&cur, 0, sizeof cur); memset(&stale, 0, sizeof stale);
	int par_file = 0;                                         /* the open file goes through the parallel indexer */
	prefetch_t pf; memset(&pf, 0, sizeof pf);
	fq_list_t lists[64]; memset(lists, 0, sizeof lists);
	fq_rec_t *recs = NULL; size_t m_recs = 0, n_recs = 0, i_rec = 0;
	struct { size_t i; uint32_t len; } *long_rec = NULL; size_t n_long_rec = 0, m_long_rec = 0;   /* over-long records of the indexed block */
	const int n_thr = o->n_parse_threads;
	for (;;) {
		const double tw0 = now_s();
		pthread_mutex_lock(&sh->mu);
		slot_t *b = &sh->slot[sh->n_filled % sh->n_slots];
		while ((b->state != SLOT_FREE || sh->n_filled - sh->n_written >= (uint64_t)sh->max_ahead) && !sh->error) pthread_cond_wait(&sh->cv, &sh->mu);
		int err = sh->error;
		pthread_mutex_unlock(&sh->mu);
		if (err) break;
		const double tw1 = now_s();
		sh->t_reader_wait += tw1 - tw0;
		pthread_mutex_lock(&sh->mu);
		if (sh->n_free_set) b->B = sh->free_set[--sh->n_free_set]; else memset(&b->B, 0, sizeof b->B);
		pthread_mutex_unlock(&sh->mu);
		b->n_reads = 0; b->n_bases = 0; b->n_names = 0; b->has_long = b->has_short = 0; b->m_bin_read_in = m_bin_read; b->n_long = 0;
		int end_of_input = 0;
		while (b->n_reads < o->batch_reads && b->n_bases < o->batch_bases) {
			if (par_file) {
				if (i_rec < n_recs) {                   /* as many indexed records as the batch takes */
					size_t e = i_rec; uint32_t nr = b->n_reads; uint64_t nb = b->n_bases;
					while (e < n_recs && nr < o->batch_reads && nb < o->batch_bases) { nb += recs[e].n_seq; nr++; e++; }
					const uint32_t first_in_batch = b->n_reads;
					const double tc0 = now_s();
					if (slot_add_block(sh, b, map, recs, i_rec, e, n_thr)) { fail(sh, "pinned alloc", -4); end_of_input = 1; break; }
					sh->t_rd_copy += now_s() - tc0;
					for (size_t i = i_rec; i < e; i++) { const size_t L = recs[i].n_seq; if (L >= 40 && 2 * L > m_bin_read) m_bin_read = (uint32_t)(2 * L + 20); }
					for (size_t k = 0; k < n_long_rec; k++)
						if (long_rec[k].i >= i_rec && long_rec[k].i < e) {
							slot_note_long(b, first_in_batch + (uint32_t)(long_rec[k].i - i_rec), long_rec[k].len);
							if (2 * (uint64_t)long_rec[k].len > m_bin_read) m_bin_read = (uint32_t)(2 * (uint64_t)long_rec[k].len + 20);
						}
					i_rec = e;
					continue;
				}
				if (cur.mbase) { block_release(&stale); stale = cur; memset(&cur, 0, sizeof cur); } else block_release(&cur);   /* a mapping is unmapped by the next read-ahead */
				map = NULL;
				if (map_pos < map_size) {               /* next block */
					uint64_t next = map_pos;
					const uint64_t len = (map_size - map_pos < FQ_BLOCK + FQ_MARGIN) ? map_size - map_pos : FQ_BLOCK + FQ_MARGIN;
					long n = -1;
					const double tr0 = now_s(); double tr1 = tr0;
					if (pf.active) {                    /* the block may have been read ahead */
						pthread_join(pf.th, NULL); pf.active = 0;
						if (pf.blk.rc == 0 && pf.pos == map_pos && pf.len == len && pf.fd == st.fd) { cur = pf.blk; if (!cur.mbase) cur_buf ^= 1; }
						else block_release(&pf.blk);
						memset(&pf.blk, 0, sizeof pf.blk);
					}
					if (!cur.map) block_get(st.fd, map_pos, len, blockbuf2[cur_buf], n_thr, &cur);
					if (cur.map) {
						map = cur.map; tr1 = now_s();
						n = fq_index_block(map, map_pos + len, map_pos + len == map_size, map_pos, map_pos + FQ_BLOCK, n_thr, lists, &recs, &m_recs, &next);
					}
					sh->t_rd_read += tr1 - tr0; sh->t_rd_index += now_s() - tr1;
					if (n >= 0) {
						n_recs = (size_t)n; i_rec = 0; map_pos = next;
						n_long_rec = 0;
						for (size_t i = 0; i < n_recs; i++)
							if (recs[i].n_seq > g_max_read_len) {
								if (n_long_rec == m_long_rec) { m_long_rec = m_long_rec ? m_long_rec * 2 : 16; long_rec = xrealloc(long_rec, m_long_rec * sizeof *long_rec); }
								long_rec[n_long_rec].i = i; long_rec[n_long_rec].len = recs[i].n_seq; n_long_rec++;
								recs[i].n_seq = 0;
							}
						if (next < map_size) {             /* read ahead while this block's records go into batches */
							pf.fd = st.fd; pf.n_thr = n_thr; pf.pos = next; pf.buf = blockbuf2[cur_buf ^ 1];
							pf.len = (map_size - next < FQ_BLOCK + FQ_MARGIN) ? map_size - next : FQ_BLOCK + FQ_MARGIN;
							memset(&pf.blk, 0, sizeof pf.blk);
							pf.old = stale; memset(&stale, 0, sizeof stale);
							if (pthread_create(&pf.th, NULL, prefetch_main, &pf) == 0) pf.active = 1; else block_release(&pf.old);
						}
						continue;
					}
					/* not strict 4-line FASTQ from here on: the serial reader takes over at the start of the block */
					block_release(&cur); map = NULL; par_file = 0;
					lseek(st.fd, (off_t)map_pos, SEEK_SET);
					st.n = st.pos = st.eof = 0; st.last_char = 0;
					n_recs = i_rec = 0;
					continue;
				}
				par_file = 0; n_recs = i_rec = 0;
				if (pf.active) { pthread_join(pf.th, NULL); pf.active = 0; block_release(&pf.blk); }
				close(st.fd); stream_open = 0; file_i++;
				continue;
			}
			if (!pending) {
				if (!stream_open) {
					if (file_i >= sh->n_files) { end_of_input = 1; break; }
					/* start inflating the .gz files among this one and the next gz_ahead - 1 */
					if (gz_next < file_i) gz_next = file_i;
					for (; gz_ahead > 0 && gz_next < sh->n_files && gz_next < file_i + gz_ahead; gz_next++) {
						struct stat sb2;
						if (stat(sh->files[gz_next], &sb2) != 0 || !S_ISREG(sb2.st_mode)) continue;   /* regular files only: opening a FIFO ahead of its turn could block */
						const int fd2 = open(sh->files[gz_next], O_RDONLY);
						if (fd2 < 0) continue;                 /* (reported when its turn comes) */
						unsigned char mg[2] = {0, 0};
						if (pread(fd2, mg, 2, 0) == 2 && mg[0] == 0x1f && mg[1] == 0x8b) gzq[gz_next] = gzq_start(fd2);
						if (!gzq[gz_next]) close(fd2);
					}
					st.fp = NULL; st.q = NULL;
					unsigned char magic[2] = {0, 0};
					int got = 0;
					if (gzq[file_i]) { st.q = gzq[file_i]; gzq[file_i] = NULL; st.fd = -1; magic[0] = 0x1f; magic[1] = 0x8b; got = 2; }
					else {
						st.fd = open(sh->files[file_i], O_RDONLY);
						if (st.fd < 0) { fprintf(stderr, "[xzopen] fail to open file '%s'\n", sh->files[file_i]); fail(sh, "open reads", -2); end_of_input = 1; break; }
						got = (int)pread(st.fd, magic, 2, 0);
						/* gzip by its magic; a pipe cannot be looked into without consuming it: zlib reads it, compressed or not */
						if ((got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) || got < 0) {
							st.fp = gzdopen(st.fd, "r");
							if (!st.fp) { fail(sh, "gzdopen", -2); end_of_input = 1; break; }
							gzbuffer(st.fp, 1 << 20);
						}
					}
					st.need_qual = (o->fmt == FMT_SAM_FULL);
					st.n = st.pos = st.eof = 0; st.last_char = 0; stream_open = 1;
					pthread_mutex_lock(&sh->mu);             /* (the reader starts while the index is still loading) */
					if (sh->started) fprintf(stderr, "Processing file: [%s].\n", sh->files[file_i]);
					else { const size_t l = strlen(sh->early_msg); snprintf(sh->early_msg + l, sizeof sh->early_msg - l, "Processing file: [%s].\n", sh->files[file_i]); }
					pthread_mutex_unlock(&sh->mu);
					struct stat sb;
					if (!st.fp && !st.q && n_thr > 0 && got >= 1 && magic[0] == '@' && fstat(st.fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0) {
						if (!g_fq_mmap && !blockbuf2[0]) { blockbuf2[0] = malloc(((uint64_t)256 << 20) + FQ_MARGIN + 16); blockbuf2[1] = malloc(((uint64_t)256 << 20) + FQ_MARGIN + 16); cur_buf = 0; }
						if (g_fq_mmap || (blockbuf2[0] && blockbuf2[1])) {
							par_file = 1; map = NULL; map_size = (uint64_t)sb.st_size; map_pos = 0; n_recs = i_rec = 0;
							continue;
						}
					}
				}
				plen = read_record(&st, &rec);
				if (plen < 0) { if (st.q) st_release_queue(&st); else if (st.fp) gzclose(st.fp); else close(st.fd); stream_open = 0; file_i++; continue; }   /* -1 end of file; -2 ends the file like the reference ends its run */
				pending = 1;
			}
			size_t L = (size_t)plen;
			if (L > g_max_read_len) { slot_note_long(b, b->n_reads, (uint32_t)(L > 0xffffffffu ? 0xffffffffu : L)); if (2 * L > m_bin_read) m_bin_read = (uint32_t)(2 * L + 20); L = 0; rec.n_qual = 0; }
			if (slot_add(sh, b, rec.name, rec.n_name, rec.seq, rec.qual, rec.n_qual, L)) { fail(sh, "pinned alloc", -4); end_of_input = 1; break; }
			if (L >= 40 && 2 * L > m_bin_read) m_bin_read = (uint32_t)(2 * L + 20);
			pending = 0;
		}
		sh->t_reader_work += now_s() - tw1;
		pthread_mutex_lock(&sh->mu);
		if (b->n_reads) {
			b->seq_no = sh->n_filled; b->state = SLOT_READY;
			bo_slot *q = &sh->bo[sh->n_filled % sh->n_slots];
			q->state = BO_READY; q->has_long = b->has_long; q->has_short = b->has_short; q->max_out = 0; q->seq_no = b->seq_no;
			sh->n_filled++; sh->total_sequences += b->n_reads;
		}
		if (end_of_input) sh->eof = 1;
		pthread_cond_broadcast(&sh->cv);
		pthread_mutex_unlock(&sh->mu);
		if (end_of_input) break;
	}
	if (pf.active) { pthread_join(pf.th, NULL); block_release(&pf.blk); }
	block_release(&cur); block_release(&stale);
	free(blockbuf2[0]); free(blockbuf2[1]);
	for (int t = 0; t < 64; t++) free(lists[t].r);
	free(recs); free(long_rec);
	st_release_queue(&st);
	for (int i = 0; i < sh->n_files; i++) gzq_close(gzq[i]);
	free(gzq);
	free(st.own_buf); free(rec.name); free(rec.seq); free(rec.qual);
	return NULL;
}

/* ---------------------------------------------------------------- GPU workers */
/* One thread per context, one batch ahead: the reads of batch k + 1 are uploaded (dsb_batch_upload: a copy stream of its own) while
 * the kernels of batch k run; then the results of k are fetched and k + 1 is started -- kt_pipeline's overlap of its read step
 * with its classify step (cly_mt.c:369-398), per context. */
static int worker_finish(worker_t *w, slot_t *b, uint64_t id, int32_t max_in)
{
	shared_t *sh = w->sh;
	size_t want = (size_t)b->n_reads * 24 + 4096;
	int32_t max_out = max_in; int rc;
	for (;;) {
		if (want > b->m_hits) { b->m_hits = want; free(b->hits); b->hits = xrealloc(NULL, b->m_hits * sizeof *b->hits); }
		if (g_host_only) { memset(b->rr, 0, (size_t)b->n_reads * sizeof *b->rr); b->n_hits = 0; rc = DSB_OK; break; }   /* DSB_HOST_ONLY: the host pipeline's own ceiling */
		rc = dsb_batch_download(w->ctx, &max_out, b->rr, b->hits, b->m_hits, &b->n_hits);
		if (rc == DSB_E_CAPACITY && b->n_hits > b->m_hits) { want = b->n_hits; continue; }    /* more hits than the array holds: again with a larger one */
		break;
	}
	uint64_t n_cap = 0;
	if (rc == DSB_E_CAPACITY) {
		/* single reads beyond a per-read capacity (-A / -m): they carry dsb_read_result.error and no hits; the run goes on and
		 * they are written as unclassified (the reference grows its vectors without bound instead) */
		for (uint32_t r = 0; r < b->n_reads; r++) if (b->rr[r].error) { n_cap++; b->rr[r].n_hit = 0; }
		rc = DSB_OK;
	}
	if (rc != DSB_OK) { fail(sh, "dsb_batch_download", rc); return -1; }
	pthread_mutex_lock(&sh->mu);
	sh->n_capacity_reads += n_cap;
	b->state = SLOT_DONE; b->rc = rc;
	sh->bo[id % sh->n_slots].max_out = max_out; sh->bo[id % sh->n_slots].state = BO_DONE;
	bo_finished(sh->bo, sh->n_slots, sh->n_claimed, &sh->done_upto, &sh->prefix_max);
	pthread_cond_broadcast(&sh->cv);
	pthread_mutex_unlock(&sh->mu);
	return 0;
}

static void *worker_main(void *arg)
{
	worker_t *w = (worker_t *)arg; shared_t *sh = w->sh;
	slot_t *cur = NULL; uint64_t cur_id = 0; int32_t cur_max_in = 0;     /* the batch on the GPU */
	for (;;) {
		/* the next batch, if there is one; wait for it only while this context has nothing on the GPU */
		slot_t *nxt = NULL; uint64_t nxt_id = 0;
		pthread_mutex_lock(&sh->mu);
		while (!sh->error && sh->n_claimed == sh->n_filled && !sh->eof && !cur) pthread_cond_wait(&sh->cv, &sh->mu);
		if (!sh->error && sh->n_claimed < sh->n_filled) {
			nxt_id = sh->n_claimed++;
			nxt = &sh->slot[nxt_id % sh->n_slots];
			nxt->state = SLOT_BUSY; sh->bo[nxt_id % sh->n_slots].state = BO_BUSY;
		}
		const int err = sh->error;
		pthread_mutex_unlock(&sh->mu);
		if (err || (!nxt && !cur)) break;
		const double tc0 = now_s();
		if (nxt) {
			if ((size_t)nxt->n_reads > nxt->m_rr) { nxt->m_rr = (size_t)nxt->n_reads * 2; nxt->rr = xrealloc(nxt->rr, nxt->m_rr * sizeof *nxt->rr); }
			if (!g_host_only) {
				dsb_ctx_set_bin_capacity(w->ctx, nxt->m_bin_read_in);
				const int rc = dsb_batch_upload(w->ctx, nxt->seqs, nxt->offs, nxt->n_reads);
				if (rc != DSB_OK) { fail(sh, "dsb_batch_upload", rc); break; }
			}
		}
		if (cur && worker_finish(w, cur, cur_id, cur_max_in)) break;
		cur = NULL;
		double t_dep = 0;
		if (nxt) {
			/* the only cross-batch dependency (batch_order.h) */
			int32_t max_in = 0;
			pthread_mutex_lock(&sh->mu);
			const double tw0 = now_s();
			while (!sh->error && !bo_may_start(sh->bo, sh->n_slots, nxt_id, sh->done_upto, sh->prefix_max, &max_in)) pthread_cond_wait(&sh->cv, &sh->mu);
			t_dep = now_s() - tw0;
			sh->t_worker_wait += t_dep;
			const int err2 = sh->error;
			pthread_mutex_unlock(&sh->mu);
			if (err2) break;
			if (!g_host_only) {
				const int rc = dsb_batch_run(w->ctx, max_in);
				if (rc != DSB_OK) { fail(sh, "dsb_batch_run", rc); break; }
			}
			cur = nxt; cur_id = nxt_id; cur_max_in = max_in;
		}
		pthread_mutex_lock(&sh->mu);
		sh->t_worker_call += now_s() - tc0 - t_dep;
		pthread_mutex_unlock(&sh->mu);
	}
	return NULL;
}

/* ---------------------------------------------------------------- writers (cly_mt.c:60-344) */
typedef struct { char *s; size_t n, m; } obuf_t;
static inline void ob_need(obuf_t *b, size_t add) { if (b->n + add > b->m) { b->m = (b->n + add) * 2 + 4096; b->s = xrealloc(b->s, b->m); } }
static const char PRI_STR[3][4] = {"PRI", "SEC", "SUP"};

/* decimal text of an int / a string, without printf: the SAM lines of 20 M short reads were 1.5 of the run's 1.7 s in sprintf */
static inline char *put_int(char *p, int v)
{
	unsigned u = (unsigned)v;
	if (v < 0) { *p++ = '-'; u = 0u - u; }
	char t[12]; int n = 0;
	do { t[n++] = (char)('0' + u % 10); u /= 10; } while (u);
	while (n) *p++ = t[--n];
	return p;
}
static inline char *put_mem(char *p, const char *s, size_t n) { memcpy(p, s, n); return p + n; }
#define PUT_LIT(p, lit) put_mem(p, lit, sizeof(lit) - 1)

static void put_hit(obuf_t *ob, const dsb_hit *c, const dsb_ref_info *ri, int rst_cnt)      /* print_hit, cly_mt.c:60-105 */
{
	ob_need(ob, 400);
	ob->n += sprintf(ob->s + ob->n, "%3d %s %s %20s ts:%-10d te:%-10d qs:%-10d qe:%-10d %-5d\t%d\t\n", rst_cnt, PRI_STR[c->primary - 1], c->direction ? "F" : "R",
	                 ri[c->ref_ID].name, (int)c->t_st, (int)c->t_ed, (int)c->q_st, (int)c->q_ed, (int)c->sum_score, (int)c->indel);
}

static void format_read(obuf_t *ob, const opts_t *o, const dsb_ref_info *ri, const dsb_read_result *r, const dsb_hit *hits,
                        const char *name, const char *seq, const char *qual, uint32_t L, int no_bases)
{
	const size_t ln = strlen(name);
	const dsb_hit *c_s = hits + r->hit_off, *c_e = c_s + r->n_hit;
	if (o->fmt == FMT_DES || o->fmt == FMT_DES_FULL) {              /* output_one_result_des / _full, cly_mt.c:158-246 */
		ob_need(ob, ln + 200);
		ob->n += sprintf(ob->s + ob->n, "%s\t%s\t%s\t%ld\tn_rst:[%ld]\tn_anc:[%ld]\t\n", name, r->n_hit ? "CLASSIFY" : "UNCLASSIFY",
		                 r->fast_classify ? "FAST" : "SLOW", (long)L, (long)r->n_hit, (long)r->n_anchor);
		int rst_cnt = 0;
		for (const dsb_hit *c = c_s; c < c_e; c++) if (c->pri_index == 0) put_hit(ob, c, ri, rst_cnt++);
		for (const dsb_hit *c = c_s; c < c_e; c++)
			if (c->pri_index > 0 && (o->fmt == FMT_DES_FULL || c->pri_index <= o->max_sec_N)) put_hit(ob, c, ri, rst_cnt++);
		ob_need(ob, 2); ob->s[ob->n++] = '\n';
		return;
	}
	const int full = o->fmt == FMT_SAM_FULL && !no_bases;           /* output_one_result_sam, cly_mt.c:248-344 (a read longer than -L has no bases here: '*') */
	const size_t lsq = full ? L : 1;
	ob_need(ob, ln + 2 * lsq + 400);
	char *p = ob->s + ob->n;
	#define PUT_SEQ_QUAL() do { if (full) { memcpy(p, seq, L); p += L; *p++ = '\t'; memcpy(p, qual, L); p += L; *p++ = '\t'; } else { memcpy(p, "*\t*\t", 4); p += 4; } } while (0)
	if (r->n_hit == 0) {
		p = put_mem(p, name, ln); p = PUT_LIT(p, "\t4\t*\t0\t0\t*\t*\t0\t0\t");
		PUT_SEQ_QUAL();
		*p++ = '\n';
		ob->n = p - ob->s;
		return;
	}
	const int flag0 = c_s->direction ? 0 : 0x10;
	int mapQ_PRI;
	if (r->n_hit == 1 || (c_s->sum_score - c_s[1].sum_score > 5)) mapQ_PRI = 30;      /* unsigned compare, as in the reference */
	else mapQ_PRI = (int)((c_s->sum_score - c_s[1].sum_score) << 2);
	{
		const char *rn = ri[c_s->ref_ID].name;
		p = put_mem(p, name, ln); *p++ = '\t'; p = put_int(p, flag0); *p++ = '\t'; p = put_mem(p, rn, strlen(rn)); *p++ = '\t';
		p = put_int(p, (int)c_s->t_st); *p++ = '\t'; p = put_int(p, mapQ_PRI); *p++ = '\t';
		p = put_int(p, (int)c_s->q_st); *p++ = 'S'; p = put_int(p, (int)(c_s->q_ed - c_s->q_st)); *p++ = 'M'; p = put_int(p, (int)(L - c_s->q_ed)); *p++ = 'S';
		p = PUT_LIT(p, "\t*\t0\t0\t");
	}
	PUT_SEQ_QUAL();
	p = PUT_LIT(p, "AS:i:"); p = put_int(p, (int)c_s->sum_score); *p++ = '\t'; *p++ = '\n';
	ob->n = p - ob->s;
	for (int loop = 0; loop <= 1; loop++)
		for (const dsb_hit *c = c_s + 1; c < c_e; c++) {
			int show = 0, fl = c->direction ? 0 : 0x10, mapQ = 0;
			if (loop == 0 && c->pri_index == 0) { show = 1; fl += 0x800; mapQ = mapQ_PRI < 30 ? mapQ_PRI : 30; }
			else if (loop == 1 && c->pri_index > 0 && c->pri_index <= o->max_sec_N) { show = 1; fl += 0x100; }
			if (!show) continue;
			const char *rn = ri[c->ref_ID].name; const size_t lrn = strlen(rn);
			const char clip = loop == 0 ? 'H' : 'S';
			ob_need(ob, ln + lrn + 200);
			char *q = ob->s + ob->n;
			q = put_mem(q, name, ln); *q++ = '\t'; q = put_int(q, fl); *q++ = '\t'; q = put_mem(q, rn, lrn); *q++ = '\t';
			q = put_int(q, (int)c->t_st); *q++ = '\t'; q = put_int(q, mapQ); *q++ = '\t';
			q = put_int(q, (int)c->q_st); *q++ = clip; q = put_int(q, (int)(c->q_ed - c->q_st)); *q++ = 'M'; q = put_int(q, (int)(L - c->q_ed)); *q++ = clip;
			q = PUT_LIT(q, "\t*\t0\t0\t*\t*\tAS:i:"); q = put_int(q, (int)c->sum_score); *q++ = '\t'; *q++ = '\n';
			ob->n = q - ob->s;
		}
	#undef PUT_SEQ_QUAL
}

/* ---------------------------------------------------------------- main */
static void usage(void)
{
	fprintf(stderr, "\nProgram:   deSAMBA-b200 (B200-native classify hot path of deSAMBA)\nVersion:   %s\n\n", dsb_version());
	fprintf(stderr, "  Usage:     deSAMBA-b200  classify  [Options] <IndexDir> [ReadFiles.fq][...]>\n");
	fprintf(stderr, "  Basic:   \n    <IndexDir>      FOLDER   the directory contains deSAMBA index\n    [ReadFiles.fq]  FILES    reads files, FASTQ(A) format, separated by space\n");
	fprintf(stderr, "  Options:\n    -h,             help\n    -t, INT         accepted for compatibility, ignored [4]\n");
	fprintf(stderr, "    -l, INT         minimum matching length, ignored for NGS reads [170]\n    -r, INT         max Output number of secondary alignments[5]\n");
	fprintf(stderr, "    -o, FILE        output results into file [stdout]\n    -s, INT         MIN score[64]\n");
	fprintf(stderr, "    -f, STR         output format, one of: SAM (default), SAM_FULL, DES, DES_FULL\n");
	fprintf(stderr, "    -g, INT         number of GPUs [all visible]\n    -c, INT         batches in flight per GPU [up to 6, as HBM allows]\n    -B, INT         reads per batch [262144]\n    -M, INT         Mbases per batch [512]\n    -P, INT         FASTQ reader threads [cores - 6; 0: serial]\n");
	fprintf(stderr, "    -A, INT         anchors kept per read [16384]\n    -m, INT         9-mer matches kept per extension [16384]\n    -L, INT         longest read accepted [1048576]\n    -p, INT         size of the per-batch device pools in %% of the built-in sizing [100]\n\n");
}

typedef struct { const opts_t *o; const dsb_ref_info *ri; slot_t *b; uint32_t r0, r1; obuf_t ob; } fmt_job_t;
static void *fmt_thread(void *a)
{
	fmt_job_t *j = (fmt_job_t *)a; slot_t *b = j->b;
	for (uint32_t r = j->r0; r < j->r1; r++) {
		int lg; const uint32_t L = slot_true_len(b, r, (uint32_t)(b->offs[r + 1] - b->offs[r]), &lg);
		format_read(&j->ob, j->o, j->ri, b->rr + r, b->hits, b->names + b->name_off[r], b->seqs + b->offs[r], b->quals ? b->quals + b->offs[r] : NULL, L, lg);
	}
	return NULL;
}

static double now_s(void) { struct timeval t; gettimeofday(&t, NULL); return t.tv_sec + t.tv_usec * 1e-6; }
static double cpu_s(void) { struct rusage r; getrusage(RUSAGE_SELF, &r); return r.ru_utime.tv_sec + r.ru_stime.tv_sec + 1e-6 * (r.ru_utime.tv_usec + r.ru_stime.tv_usec); }

static int classify_main(int argc, char **argv)
{
	opts_t o = {170, 4, 5, FMT_SAM, 64, 0, 0, 262144, 512ull << 20, stdout, -1, 0, 0, 0, 0};
	int c;
	while ((c = getopt(argc, argv, "ht:l:r:f:o:s:g:B:M:c:P:A:m:L:p:")) >= 0) {
		if (c == 'h') { usage(); return 0; }
		else if (c == 't') o.n_threads = atoi(optarg);
		else if (c == 'l') o.l_min_match = atoi(optarg);
		else if (c == 'r') o.max_sec_N = atoi(optarg);
		else if (c == 'o') { o.out = fopen(optarg, "w"); if (!o.out) { fprintf(stderr, "[xopen] fail to open file '%s'\n", optarg); return 1; } }
		else if (c == 's') o.min_score = atoi(optarg);
		else if (c == 'g') o.n_gpus = atoi(optarg);
		else if (c == 'P') o.n_parse_threads = atoi(optarg);
		else if (c == 'c') o.ctx_per_gpu = atoi(optarg);
		else if (c == 'A') o.max_anchors = (uint32_t)atol(optarg);
		else if (c == 'm') o.max_matches = (uint32_t)atol(optarg);
		else if (c == 'L') o.max_read_len = (uint32_t)atol(optarg);
		else if (c == 'p') o.pool_scale_pct = (uint32_t)atol(optarg);
		else if (c == 'B') o.batch_reads = (uint32_t)atol(optarg);
		else if (c == 'M') o.batch_bases = (uint64_t)atol(optarg) << 20;
		else if (c == 'f') {
			if (!strcmp(optarg, "SAM")) o.fmt = FMT_SAM; else if (!strcmp(optarg, "SAM_FULL")) o.fmt = FMT_SAM_FULL;
			else if (!strcmp(optarg, "DES")) o.fmt = FMT_DES; else if (!strcmp(optarg, "DES_FULL")) o.fmt = FMT_DES_FULL;
		}
	}
	if (optind + 2 > argc) { usage(); return 0; }
	if (o.batch_reads < 1) o.batch_reads = 1;
	if (o.batch_bases < 1) o.batch_bases = 1;
	const char *index_dir = argv[optind++];
	if (o.n_gpus <= 0) { const char *e = getenv("DSB_GPUS"); o.n_gpus = e ? atoi(e) : 0; }
	fprintf(stderr, "loading index\t");
	const int auto_ctx = o.ctx_per_gpu < 1;                 /* as many as HBM allows, up to 6 (decided once the index is resident) */
	if (auto_ctx) o.ctx_per_gpu = 6;
	if (o.ctx_per_gpu > 8) o.ctx_per_gpu = 8;
	if (o.n_parse_threads < 0) { long nc = sysconf(_SC_NPROCESSORS_ONLN); o.n_parse_threads = nc >= 16 ? (nc - 6 > 48 ? 48 : (int)nc - 6) : nc >= 8 ? 4 : nc >= 4 ? 2 : 1;   /* the GPU worker threads stage pageable batches with 3 threads each: leave them cores */ }
	if (o.n_parse_threads > 64) o.n_parse_threads = 64;
	const int verbose = getenv("DSB_VERBOSE") != NULL;
	g_pageable = getenv("DSB_PINNED") == NULL;
	g_register = getenv("DSB_REGISTER") != NULL;
	if (getenv("DSB_FQ_MMAP")) g_fq_mmap = atoi(getenv("DSB_FQ_MMAP")) != 0;
	g_host_only = getenv("DSB_HOST_ONLY") != NULL;     /* test / developer aid: reader + writer only -- no device is opened, every read is written as unclassified */
	if (getenv("DSB_FQ_BLOCK_MB") && atol(getenv("DSB_FQ_BLOCK_MB")) >= 1 && atol(getenv("DSB_FQ_BLOCK_MB")) <= 256) g_fq_block = (uint64_t)atol(getenv("DSB_FQ_BLOCK_MB")) << 20;
	if (getenv("DSB_FQ_BLOCK_KB") && atol(getenv("DSB_FQ_BLOCK_KB")) >= 1 && atol(getenv("DSB_FQ_BLOCK_KB")) <= (256 << 10)) g_fq_block = (uint64_t)atol(getenv("DSB_FQ_BLOCK_KB")) << 10;   /* tests */
	{ dsb_opts d0; dsb_opts_default(&d0); g_max_read_len = o.max_read_len ? o.max_read_len : d0.max_read_len; }
	const double t_start = now_s();
	#define STAMP(what) do { if (verbose) fprintf(stderr, "[deSAMBA-b200] %-34s at %7.3f s\n", what, now_s() - t_start); } while (0)
	/* the reader starts at once: the first batches are parsed while the index is loaded into HBM */
	shared_t sh; memset(&sh, 0, sizeof sh);
	sh.o = &o; sh.n_slots = 2 * ((o.n_gpus > 0 ? o.n_gpus : 8) * o.ctx_per_gpu) + 2; sh.slot = xcalloc(sh.n_slots, sizeof(slot_t)); sh.bo = xcalloc(sh.n_slots, sizeof(bo_slot)); sh.free_set = xcalloc(sh.n_slots + 1, sizeof(bufset_t));
	sh.max_ahead = getenv("DSB_READ_AHEAD") ? atoi(getenv("DSB_READ_AHEAD")) : 3;   /* batches filled while the index loads: few -- faulting fresh buffers in from a dozen threads slows the
	                                                                                 * loader's and the contexts' device allocations (same mm lock) by more than it saves later (measured: 9.0 s wall against 4.5 s) */
	if (sh.max_ahead < 1) sh.max_ahead = 1;
	if (sh.max_ahead > sh.n_slots) sh.max_ahead = sh.n_slots;
	pthread_mutex_init(&sh.mu, NULL); pthread_cond_init(&sh.cv, NULL);
	sh.n_files = argc - optind; sh.files = argv + optind;
	pthread_t rd;
	pthread_create(&rd, NULL, reader_main, &sh);
	/* load_idx runs ONCE (idx.c:1103): the index is read, uploaded and re-cut on GPU 0; the other GPUs get device-to-device copies */
	#define BAIL(...) do { fprintf(stderr, __VA_ARGS__); pthread_mutex_lock(&sh.mu); if (!sh.error) sh.error = -1; pthread_cond_broadcast(&sh.cv); pthread_mutex_unlock(&sh.mu); pthread_join(rd, NULL); return 1; } while (0)
	dsb_index *gix[64];
	int n_gpus = 0;
	/* the worker threads sleep while their batch is on the GPU: up to 6 x 8 of them would otherwise spin on the reader's cores */
	dsb_set_sync_mode(getenv("DSB_SPIN") == NULL);
	double t_load0 = 0, t_clone = 0;
	for (int g = 0; g < (o.n_gpus > 0 ? o.n_gpus : 64) && !g_host_only; g++) {
		dsb_index *ix = NULL;
		const double tl = now_s();
		int rc = (g == 0 || getenv("DSB_NO_CLONE")) ? dsb_index_load(index_dir, g, &ix) : dsb_index_clone(gix[0], g, &ix);
		if (rc != DSB_OK) {
			if (g == 0 || o.n_gpus > 0) BAIL("\n[deSAMBA-b200] cannot load index on GPU %d: %s\n", g, dsb_last_error());
			break;                                       /* ran out of visible devices */
		}
		if (g == 0) t_load0 = now_s() - tl; else t_clone += now_s() - tl;
		gix[g] = ix; n_gpus++;
	}
	STAMP("index resident in HBM");
	/* several contexts (streams) per GPU: the expensive tail reads of one batch overlap the next batch.  Every device buffer of a
	 * context is allocated at its final size now (no cudaMalloc while batches are in flight); without -c, contexts are added
	 * while a quarter of the HBM stays free for the pools that may still grow */
	dsb_opts dop; dsb_opts_default(&dop);
	dop.l_min_match = o.l_min_match; dop.min_score = o.min_score;
	if (o.max_anchors) dop.max_anchors = o.max_anchors;
	if (o.max_matches) dop.max_matches = o.max_matches;
	if (o.max_read_len) dop.max_read_len = o.max_read_len;
	if (o.pool_scale_pct) dop.pool_scale_pct = o.pool_scale_pct;
	worker_t *w = xcalloc((size_t)(n_gpus ? n_gpus : 1) * o.ctx_per_gpu, sizeof *w);
	int n_workers = 0, ctx_made = 0;
	for (int k = 0; k < o.ctx_per_gpu && !g_host_only; k++) {   /* round k: one more context on every GPU */
		int ok = 1;
		for (int g = 0; g < n_gpus && ok; g++) {
			uint64_t fr = 0, tot = 0;
			if (auto_ctx && k >= 2 && dsb_device_memory(g, &fr, &tot) == DSB_OK && fr < tot / 4) { ok = 0; break; }
			worker_t *x = &w[n_workers];
			x->gpu = g; x->ix = gix[g];
			if (dsb_ctx_create(x->ix, &dop, &x->ctx) != DSB_OK) BAIL("\n[deSAMBA-b200] %s\n", dsb_last_error());
			if (!getenv("DSB_NO_RESERVE") && dsb_ctx_reserve(x->ctx, o.batch_reads, o.batch_bases + ((uint64_t)4 << 20)) != DSB_OK) {
				if (k == 0) BAIL("\n[deSAMBA-b200] %s\n", dsb_last_error());
				dsb_ctx_free(x->ctx); x->ctx = NULL; ok = 0; break;   /* HBM is full: the contexts made so far do the work */
			}
			n_workers++;
		}
		if (!ok) { while (n_workers > (k) * n_gpus) { n_workers--; dsb_ctx_free(w[n_workers].ctx); } break; }
		ctx_made = k + 1;
	}
	o.ctx_per_gpu = ctx_made;
	if (g_host_only) n_workers = 2;                         /* no device is opened at all: two threads hand the batches on with empty results */
	STAMP("contexts created");
	const dsb_ref_info *ri = g_host_only ? NULL : dsb_index_ref_info(w[0].ix);
	const double t0 = now_s(), c0 = cpu_s();
	fprintf(stderr, "Start classify\n");
	pthread_mutex_lock(&sh.mu);
	sh.started = 1; fputs(sh.early_msg, stderr);
	sh.max_ahead = 2 * n_workers + 2;
	if (sh.max_ahead > sh.n_slots) sh.max_ahead = sh.n_slots;
	pthread_cond_broadcast(&sh.cv);
	pthread_mutex_unlock(&sh.mu);
	pthread_t *th = xcalloc(n_workers, sizeof *th);
	for (int k = 0; k < n_workers; k++) { w[k].sh = &sh; pthread_create(&th[k], NULL, worker_main, &w[k]); }

	/* the text of a batch is formatted by helper threads (contiguous shares of the batch's reads, written in order) */
	const int n_fmt = o.n_parse_threads > 1 ? (o.n_parse_threads > 16 ? 16 : o.n_parse_threads) : 1;
	fmt_job_t fj[16]; memset(fj, 0, sizeof fj);
	obuf_t ob = {0};
	for (;;) {
		const double tq0 = now_s();
		pthread_mutex_lock(&sh.mu);
		slot_t *b = &sh.slot[sh.n_written % sh.n_slots];
		while (!sh.error && !(b->state == SLOT_DONE && b->seq_no == sh.n_written) && !(sh.eof && sh.n_written == sh.n_filled)) pthread_cond_wait(&sh.cv, &sh.mu);
		const int stop = sh.error || !(b->state == SLOT_DONE && b->seq_no == sh.n_written);
		pthread_mutex_unlock(&sh.mu);
		if (stop) break;
		const double tq1 = now_s();
		sh.t_writer_wait += tq1 - tq0;
		if (n_fmt > 1 && b->n_reads >= 4096) {
			pthread_t ft[16];
			for (int t = 0; t < n_fmt; t++) {
				fj[t].o = &o; fj[t].ri = ri; fj[t].b = b; fj[t].ob.n = 0;
				fj[t].r0 = (uint32_t)((uint64_t)b->n_reads * t / n_fmt); fj[t].r1 = (uint32_t)((uint64_t)b->n_reads * (t + 1) / n_fmt);
			}
			for (int t = 1; t < n_fmt; t++) pthread_create(&ft[t], NULL, fmt_thread, &fj[t]);
			fmt_thread(&fj[0]);
			for (int t = 1; t < n_fmt; t++) pthread_join(ft[t], NULL);
			for (int t = 0; t < n_fmt; t++) fwrite(fj[t].ob.s, 1, fj[t].ob.n, o.out);
		} else {
			ob.n = 0;
			for (uint32_t r = 0; r < b->n_reads; r++) {
				int lg; const uint32_t L = slot_true_len(b, r, (uint32_t)(b->offs[r + 1] - b->offs[r]), &lg);
				format_read(&ob, &o, ri, b->rr + r, b->hits, b->names + b->name_off[r], b->seqs + b->offs[r], b->quals ? b->quals + b->offs[r] : NULL, L, lg);
				if (ob.n > (8u << 20)) { fwrite(ob.s, 1, ob.n, o.out); ob.n = 0; }
			}
			fwrite(ob.s, 1, ob.n, o.out);
		}
		sh.t_writer_fmt += now_s() - tq1;
		sh.n_long_reads += b->n_long;
		pthread_mutex_lock(&sh.mu);
		sh.free_set[sh.n_free_set++] = b->B; memset(&b->B, 0, sizeof b->B);
		b->state = SLOT_FREE; sh.n_written++;
		pthread_cond_broadcast(&sh.cv);
		pthread_mutex_unlock(&sh.mu);
	}
	pthread_join(rd, NULL);
	for (int k = 0; k < n_workers; k++) pthread_join(th[k], NULL);
	STAMP("last record written");
	if (verbose) fprintf(stderr, "[deSAMBA-b200] batch buffers: %d allocations, %.3f s (%s)\n", g_n_pinned, g_t_pinned, g_pageable ? "pageable" : "pinned");
	if (verbose) fprintf(stderr, "[deSAMBA-b200] host time: reader %.3f s work (block input %.3f, indexing %.3f, copies %.3f) + %.3f s waiting for a free batch; GPU calls %.3f s (sum over %d worker threads); writer %.3f s formatting + %.3f s waiting\n",
	                     sh.t_reader_work, sh.t_rd_read, sh.t_rd_index, sh.t_rd_copy, sh.t_reader_wait, sh.t_worker_call, n_workers, sh.t_writer_fmt, sh.t_writer_wait);
	if (verbose && !g_host_only) {
		double sum[5] = {0, 0, 0, 0, 0};
		for (int k = 0; k < n_workers; k++) { double h[5] = {0, 0, 0, 0, 0}; dsb_ctx_host_seconds(w[k].ctx, h, 5); for (int i = 0; i < 5; i++) sum[i] += h[i]; }
		fprintf(stderr, "[deSAMBA-b200] GPU calls: %.0f batches, upload %.3f s, launches %.3f s, waiting + results %.3f s (sums over the worker threads); %.0f pool-overflow re-runs\n", sum[3], sum[0], sum[1], sum[2], sum[4]);
	}
	fflush(o.out);
	if (o.out != stdout) fclose(o.out);
	if (sh.error) { fprintf(stderr, "[deSAMBA-b200] error %d: %s\n", sh.error, sh.errmsg); return 1; }
	const double sec = now_s() - t0;
	fprintf(stderr, "%ld sequences processed in %.3fs (%.1f Kseq/m).\n", (long)sh.total_sequences, sec, sh.total_sequences / 1.0e3 / (sec / 60));   /* report_stats, cly_mt.c:439-446 */
	fprintf(stderr, "Classify CPU: %.3f sec\n", cpu_s() - c0);
	fprintf(stderr, "GPUs: %d (%d contexts each); index: %.3f s load on GPU 0 + %.3f s device-to-device copies\n", n_gpus, o.ctx_per_gpu, t_load0, t_clone);
	if (sh.n_long_reads) fprintf(stderr, "[deSAMBA-b200] warning: %llu read(s) longer than %u bases were written as unclassified (raise -L)\n", (unsigned long long)sh.n_long_reads, g_max_read_len);
	if (sh.n_capacity_reads) fprintf(stderr, "[deSAMBA-b200] warning: %llu read(s) exceeded a per-read capacity and were written as unclassified (raise -A / -m)\n", (unsigned long long)sh.n_capacity_reads);
	/* Freeing tens of GB of device memory buffer by buffer took 0.2 - 3.8 s here; the process is over and the driver reclaims all
	 * of it at once.  DSB_FREE_AT_EXIT=1 keeps the orderly teardown (leak checkers). */
	if (!getenv("DSB_FREE_AT_EXIT")) { fflush(stdout); fflush(stderr); _exit(0); }
	for (int k = 0; k < n_workers; k++) if (w[k].ctx) dsb_ctx_free(w[k].ctx);
	for (int g = 0; g < n_gpus; g++) dsb_index_free(gix[g]);
	STAMP("device memory released");
	return 0;
}

int main(int argc, char **argv)
{
	if (argc >= 2 && !strcmp(argv[1], "classify")) return classify_main(argc - 1, argv + 1);
	if (argc >= 2 && (!strcmp(argv[1], "-h") || !strcmp(argv[1], "--help"))) { usage(); return 0; }
	fprintf(stderr, "deSAMBA-b200: only the `classify` command is implemented here (index construction and analysis stay with the reference)\n");
	usage();
	return 1;
}
