/*
 * fastq_reader.h -- FASTA/FASTQ(.gz) input of `deSAMBA-b200 classify`.
 *
 * (1) the serial reader: kseq_read semantics (utils.c:939-977) over read(2) / zlib, one record at a time;
 * (2) a parallel indexer for plain (uncompressed) 4-line FASTQ, the format sequencers write: helper threads read a block of
 *     the file into a buffer, find and check the records that start in their share of it, and the driver copies the bases of
 *     a whole batch into the batch buffer with the same helpers.  Anything that is not strict 4-line FASTQ (FASTA, wrapped lines, a
 *     truncated or inconsistent record, junk between records) makes the indexer give up at the start of the block and the
 *     serial reader takes over from there, so the records handed on are the same in every case (tests/test_reader.py).
 * A single-threaded parser feeds ~0.8 Gbases/s; one B200 classifies 6.
 */
#ifndef DSB_FASTQ_READER_H
#define DSB_FASTQ_READER_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <pthread.h>
#include <zlib.h>

/* ---------------------------------------------------------------- .gz input inflated ahead of the parser
 * A gzip stream cannot be inflated in parallel, but the files of a run can: each of the next few .gz files of the command line
 * gets a thread that inflates it into a bounded queue of chunks, so the parser never waits for zlib on inputs that come as many
 * files (a sequencer's output directory), and on a single file inflating at least overlaps parsing. */
#define GZ_CHUNK (4 << 20)
#define GZ_QUEUE_CAP ((size_t)128 << 20)         /* inflated bytes a file may be ahead of the parser */
typedef struct gz_chunk { struct gz_chunk *next; int n; unsigned char data[GZ_CHUNK]; } gz_chunk_t;
typedef struct {
	int fd; gzFile fp; pthread_t th; int started, done, stop;
	pthread_mutex_t mu; pthread_cond_t cv;
	gz_chunk_t *head, *tail; size_t queued;
} gzq_t;
static __attribute__((unused)) void *gzq_main(void *a)
{
	gzq_t *q = (gzq_t *)a;
	for (;;) {
		pthread_mutex_lock(&q->mu);
		while (q->queued >= GZ_QUEUE_CAP && !q->stop) pthread_cond_wait(&q->cv, &q->mu);
		const int stop = q->stop;
		pthread_mutex_unlock(&q->mu);
		if (stop) break;
		gz_chunk_t *c = (gz_chunk_t *)malloc(sizeof(gz_chunk_t));
		if (!c) break;
		c->next = NULL;
		c->n = gzread(q->fp, c->data, GZ_CHUNK);
		if (c->n <= 0) { free(c); break; }              /* end of the stream (or a damaged one: the file ends here, as with gzread in the parser) */
		pthread_mutex_lock(&q->mu);
		if (q->tail) q->tail->next = c; else q->head = c;
		q->tail = c; q->queued += (size_t)c->n;
		pthread_cond_broadcast(&q->cv);
		pthread_mutex_unlock(&q->mu);
	}
	pthread_mutex_lock(&q->mu);
	q->done = 1;
	pthread_cond_broadcast(&q->cv);
	pthread_mutex_unlock(&q->mu);
	return NULL;
}
/* takes over fd (a gzip file positioned at its start); NULL if the thread cannot be started (fd stays open) */
static __attribute__((unused)) gzq_t *gzq_start(int fd)
{
	gzq_t *q = (gzq_t *)calloc(1, sizeof(gzq_t));
	if (!q) return NULL;
	q->fd = fd; q->fp = gzdopen(fd, "r");
	if (!q->fp) { free(q); return NULL; }
	gzbuffer(q->fp, 1 << 20);
	pthread_mutex_init(&q->mu, NULL); pthread_cond_init(&q->cv, NULL);
	if (pthread_create(&q->th, NULL, gzq_main, q) != 0) { free(q); return NULL; }   /* (gzFile leaks its small state; the caller goes on with the fd) */
	q->started = 1;
	return q;
}
/* next chunk (blocking); NULL at the end of the stream */
static __attribute__((unused)) gz_chunk_t *gzq_pop(gzq_t *q)
{
	pthread_mutex_lock(&q->mu);
	while (!q->head && !q->done) pthread_cond_wait(&q->cv, &q->mu);
	gz_chunk_t *c = q->head;
	if (c) { q->head = c->next; if (!q->head) q->tail = NULL; q->queued -= (size_t)c->n; pthread_cond_broadcast(&q->cv); }
	pthread_mutex_unlock(&q->mu);
	return c;
}
static __attribute__((unused)) void gzq_close(gzq_t *q)
{
	if (!q) return;
	pthread_mutex_lock(&q->mu); q->stop = 1; pthread_cond_broadcast(&q->cv); pthread_mutex_unlock(&q->mu);
	pthread_join(q->th, NULL);
	for (gz_chunk_t *c = q->head; c;) { gz_chunk_t *nx = c->next; free(c); c = nx; }
	gzclose(q->fp);
	pthread_mutex_destroy(&q->mu); pthread_cond_destroy(&q->cv);
	free(q);
}

/* ---------------------------------------------------------------- serial reader (kseq_read semantics, utils.c:939-977) */
typedef struct {
	gzFile fp; int fd; unsigned char *buf; int n, pos, eof; int last_char; int need_qual;
	gzq_t *q; gz_chunk_t *chunk; unsigned char *own_buf;      /* q: the stream comes from an inflating thread, buf points into its current chunk */
} stream_t;
#define SBUF (4 << 20)
/* plain files are read with read(2) (zlib's transparent mode costs an extra copy of every byte); .gz through zlib */
static inline int st_fill(stream_t *s)
{
	if (s->q) {
		if (s->chunk) { free(s->chunk); s->chunk = NULL; }
		s->chunk = gzq_pop(s->q);
		s->pos = 0;
		if (!s->chunk) { s->buf = s->own_buf; s->eof = 1; s->n = 0; return -1; }
		s->buf = s->chunk->data; s->n = s->chunk->n;
		return 0;
	}
	s->n = s->fp ? gzread(s->fp, s->buf, SBUF) : (int)read(s->fd, s->buf, SBUF);
	s->pos = 0;
	if (s->n <= 0) { s->eof = 1; s->n = 0; return -1; }
	return 0;
}
/* a stream is detached from its inflating thread when its file is done */
static inline void st_release_queue(stream_t *s)
{
	if (s->chunk) { free(s->chunk); s->chunk = NULL; }
	if (s->q) { gzq_close(s->q); s->q = NULL; }
	if (s->own_buf) s->buf = s->own_buf;
}
static inline int st_getc(stream_t *s)
{
	if (s->pos >= s->n) {
		if (s->eof || st_fill(s)) return -1;
	}
	return s->buf[s->pos++];
}
/* append bytes up to (not including) the next '\n' (or any whitespace if `word`) to *dst; returns the delimiter or -1 */
static int st_getuntil(stream_t *s, int word, char **dst, size_t *n, size_t *m)
{
	for (;;) {
		if (s->pos >= s->n) {
			if (s->eof || st_fill(s)) return -1;
		}
		int i = s->pos;
		if (word) { while (i < s->n && s->buf[i] != '\n' && s->buf[i] != ' ' && s->buf[i] != '\t' && s->buf[i] != '\r' && s->buf[i] != '\v' && s->buf[i] != '\f') i++; }
		else { unsigned char *p = memchr(s->buf + i, '\n', s->n - i); i = p ? (int)(p - s->buf) : s->n; }
		size_t add = i - s->pos;
		if (dst) {
			if (*n + add + 1 > *m) { *m = (*n + add + 1) * 2; *dst = realloc(*dst, *m); }
			memcpy(*dst + *n, s->buf + s->pos, add); *n += add;
		}
		s->pos = i;
		if (i < s->n) { s->pos++; return s->buf[i]; }
	}
}

typedef struct { char *name, *seq, *qual; size_t n_name, m_name, n_seq, m_seq, n_qual, m_qual; } rec_t;
/* returns seq length >= 0, -1 at end of file, -2 on a truncated quality string (kseq's error codes) */
static long read_record(stream_t *s, rec_t *r)
{
	int c;
	if (s->last_char == 0) {
		while ((c = st_getc(s)) != -1 && c != '>' && c != '@');
		if (c == -1) return -1;
		s->last_char = c;
	}
	r->n_name = r->n_seq = r->n_qual = 0;
	c = st_getuntil(s, 1, &r->name, &r->n_name, &r->m_name);
	if (c == -1 && r->n_name == 0) return -1;
	if (c != '\n' && c != -1) st_getuntil(s, 0, NULL, NULL, NULL);         /* comment */
	while ((c = st_getc(s)) != -1 && c != '>' && c != '+' && c != '@') {
		if (c == '\n') continue;
		if (r->n_seq + 2 > r->m_seq) { r->m_seq = (r->n_seq + 2) * 2; r->seq = realloc(r->seq, r->m_seq); }
		r->seq[r->n_seq++] = (char)c;
		st_getuntil(s, 0, &r->seq, &r->n_seq, &r->m_seq);
		while (r->n_seq && r->seq[r->n_seq - 1] == '\r') r->n_seq--;
	}
	s->last_char = (c == '>' || c == '@') ? c : 0;
	if (c != '+') return (long)r->n_seq;                                     /* FASTA record */
	st_getuntil(s, 0, NULL, NULL, NULL);                                     /* rest of the '+' line */
	/* fast path (4-line FASTQ, quality not printed): the quality line is exactly as long as the sequence */
	if (!s->need_qual && r->n_seq && (size_t)(s->n - s->pos) > r->n_seq && s->buf[s->pos + r->n_seq] == '\n') {
		s->pos += (int)r->n_seq + 1; s->last_char = 0;
		return (long)r->n_seq;
	}
	while (r->n_qual < r->n_seq) {
		c = st_getuntil(s, 0, &r->qual, &r->n_qual, &r->m_qual);
		while (r->n_qual && r->qual[r->n_qual - 1] == '\r') r->n_qual--;
		if (c == -1) break;
	}
	s->last_char = 0;
	if (r->n_qual != r->n_seq) return -2;
	return (long)r->n_seq;
}


/* ---------------------------------------------------------------- parallel indexer for plain 4-line FASTQ */
typedef struct { uint64_t name, seq, qual; uint32_t n_name, n_seq; } fq_rec_t;      /* byte offsets into the mapping */
typedef struct { fq_rec_t *r; size_t n, m; int bad; uint64_t first, next; } fq_list_t;   /* first: start of the first record at or after the share, next: start of the first record at or after its end */

static inline int fq_is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

/* the record starting at p: fills *o and *next (start of the following record); 0 ok, -1 not strict 4-line FASTQ / cut off */
static inline int fq_parse_at(const char *map, uint64_t size, int at_eof, uint64_t p, fq_rec_t *o, uint64_t *next)
{
	if (p >= size || map[p] != '@') return -1;
	const char *e1 = memchr(map + p, '\n', size - p);
	if (!e1) return -1;
	uint64_t q = p + 1, l1 = (uint64_t)(e1 - map);
	while (q < l1 && !fq_is_space(map[q])) q++;
	o->name = p + 1; o->n_name = (uint32_t)(q - (p + 1));
	const uint64_t s0 = l1 + 1;
	if (s0 >= size) return -1;
	const char *e2 = memchr(map + s0, '\n', size - s0);
	if (!e2) return -1;
	uint64_t s1 = (uint64_t)(e2 - map);
	const uint64_t p3 = s1 + 1;
	while (s1 > s0 && map[s1 - 1] == '\r') s1--;
	if (s1 == s0 || s1 - s0 > 0x7fffffffu) return -1;
	if (map[s0] == '@' || map[s0] == '>' || map[s0] == '+') return -1;
	o->seq = s0; o->n_seq = (uint32_t)(s1 - s0);
	if (p3 >= size || map[p3] != '+') return -1;
	const char *e3 = memchr(map + p3, '\n', size - p3);
	if (!e3) return -1;
	const uint64_t q0 = (uint64_t)(e3 - map) + 1;
	if (q0 > size) return -1;
	const char *e4 = (q0 < size) ? memchr(map + q0, '\n', size - q0) : NULL;
	if (!e4 && !at_eof) return -1;                                   /* ran into the end of the buffered part, not of the file */
	uint64_t q1 = e4 ? (uint64_t)(e4 - map) : size;
	const uint64_t nx = e4 ? q1 + 1 : size;
	while (q1 > q0 && map[q1 - 1] == '\r') q1--;
	if (q1 - q0 != o->n_seq) return -1;
	o->qual = q0;
	if (nx < size && map[nx] != '@') return -1;
	if (nx >= size && !at_eof) return -1;
	*next = nx;
	return 0;
}

/* the records that START in [lo, hi); lo == 0 or any offset (the first record start at or after lo is looked for: a line
 * starting with '@' that parses as a record whose successor parses too).  out->next = start of the first record >= hi. */
static void fq_index_range(const char *map, uint64_t size, int at_eof, uint64_t lo, uint64_t hi, int lo_is_start, fq_list_t *out)
{
	out->n = 0; out->bad = 0; out->first = out->next = lo;
	uint64_t p = lo;
	fq_rec_t r; uint64_t nx;
	if (!lo_is_start) {
		int found = 0;
		if (p > 0) { const char *e = memchr(map + p - 1, '\n', size - (p - 1)); p = e ? (uint64_t)(e - map) + 1 : size; }
		for (int tries = 0; tries < 64 && p < size && p < hi + (1u << 20); tries++) {
			fq_rec_t r2; uint64_t nx2;
			if (map[p] == '@' && fq_parse_at(map, size, at_eof, p, &r, &nx) == 0 && (nx >= size || fq_parse_at(map, size, at_eof, nx, &r2, &nx2) == 0)) { found = 1; break; }
			const char *e = memchr(map + p, '\n', size - p);
			p = e ? (uint64_t)(e - map) + 1 : size;
		}
		if (p >= size) { if (at_eof) out->first = out->next = size; else out->bad = 1; return; }
		if (!found) { out->bad = 1; return; }
		out->first = p;
	}
	while (p < hi && p < size) {
		if (fq_parse_at(map, size, at_eof, p, &r, &nx) != 0) { out->bad = 1; return; }
		if (out->n == out->m) { out->m = out->m ? out->m * 2 : 4096; out->r = (fq_rec_t *)realloc(out->r, out->m * sizeof(fq_rec_t)); }
		out->r[out->n++] = r;
		p = nx;
	}
	out->next = p;
}

typedef struct { const char *map; uint64_t size, lo, hi; int at_eof, lo_is_start; fq_list_t *out; } fq_index_job_t;
static void *fq_index_thread(void *a) { fq_index_job_t *j = (fq_index_job_t *)a; fq_index_range(j->map, j->size, j->at_eof, j->lo, j->hi, j->lo_is_start, j->out); return NULL; }

/* index the records that start in the block [lo, hi) of the mapping with n_thr threads; lo must be a record start.
 * map[0, size) need only be readable from lo on; at_eof = size is the end of the file (else: of the buffered part, and a
 * record that runs into it makes the block fail).  Returns the number of records (written to *recs in file order) and the
 * start of the next block in *next, or -1 when the block is not strict 4-line FASTQ. */
static long fq_index_block(const char *map, uint64_t size, int at_eof, uint64_t lo, uint64_t hi, int n_thr, fq_list_t *lists, fq_rec_t **recs, size_t *m_recs, uint64_t *next)
{
	if (hi > size) hi = size;
	if (n_thr < 1) n_thr = 1;
	if (n_thr > 64) n_thr = 64;
	const uint64_t span = (hi - lo + n_thr - 1) / n_thr;
	fq_index_job_t job[64]; pthread_t th[64];
	for (int t = 0; t < n_thr; t++) {
		job[t].map = map; job[t].size = size; job[t].at_eof = at_eof; job[t].lo = lo + t * span; job[t].hi = (t == n_thr - 1) ? hi : lo + (t + 1) * span;
		if (job[t].lo > hi) job[t].lo = hi;
		if (job[t].hi > hi) job[t].hi = hi;
		job[t].lo_is_start = (t == 0); job[t].out = &lists[t];
	}
	for (int t = 1; t < n_thr; t++) pthread_create(&th[t], NULL, fq_index_thread, &job[t]);
	fq_index_thread(&job[0]);
	for (int t = 1; t < n_thr; t++) pthread_join(th[t], NULL);
	/* the shares must meet: the records of share t end where share t + 1 found its first record start */
	size_t n = 0; uint64_t expect = lo;
	for (int t = 0; t < n_thr; t++) {
		if (lists[t].bad || lists[t].first != expect) return -1;
		expect = lists[t].next;
		n += lists[t].n;
	}
	if (n > *m_recs) { *m_recs = n + n / 4 + 1024; *recs = (fq_rec_t *)realloc(*recs, *m_recs * sizeof(fq_rec_t)); }
	size_t k = 0;
	for (int t = 0; t < n_thr; t++) { if (lists[t].n) memcpy(*recs + k, lists[t].r, lists[t].n * sizeof(fq_rec_t)); k += lists[t].n; }
	*next = expect;
	return (long)n;
}

/* ---- block input: the file is read (not mapped: a page fault per 4 KB under one mm lock is slower than the serial reader)
 * into a reusable buffer, FQ_MARGIN bytes beyond the block so that the records starting in the block are complete */
typedef struct { int fd; char *dst; uint64_t off, len; long got; } fq_read_job_t;
static void *fq_read_thread(void *a)
{
	fq_read_job_t *j = (fq_read_job_t *)a; uint64_t done = 0;
	while (done < j->len) {
		const ssize_t n = pread(j->fd, j->dst + done, j->len - done, (off_t)(j->off + done));
		if (n <= 0) break;
		done += (uint64_t)n;
	}
	j->got = (long)done;
	return NULL;
}
/* bytes [pos, pos + len) of fd into buf with n_thr threads; returns 0 when all of them arrived */
static int fq_read_block(int fd, uint64_t pos, uint64_t len, char *buf, int n_thr)
{
	if (n_thr < 1) n_thr = 1;
	if (n_thr > 64) n_thr = 64;
	fq_read_job_t job[64]; pthread_t th[64];
	const uint64_t span = ((len + n_thr - 1) / n_thr + 4095) & ~4095ull;
	int n = 0;
	for (uint64_t o = 0; o < len && n < 64; o += span, n++) { job[n].fd = fd; job[n].dst = buf + o; job[n].off = pos + o; job[n].len = (o + span <= len) ? span : len - o; job[n].got = 0; }
	for (int t = 1; t < n; t++) pthread_create(&th[t], NULL, fq_read_thread, &job[t]);
	if (n) fq_read_thread(&job[0]);
	for (int t = 1; t < n; t++) pthread_join(th[t], NULL);
	for (int t = 0; t < n; t++) if ((uint64_t)job[t].got != job[t].len) return -1;
	return 0;
}
#endif
