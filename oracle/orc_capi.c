/*
 * orc_capi.c -- TEST INFRASTRUCTURE ONLY.  Flat-array entry points over the oracle restatement for ctypes (tests/,
 * __graft_entry__.smoke(), bench.py's cpu_baseline leg).  Result records use the byte layout of dsb_hit /
 * dsb_read_result / dsb_seed (include/desamba_b200.h) so that parity tests compare arrays directly.
 * Reads are classified in input order with ONE scratch buffer = the reference's `-t 1` semantics (running max_read_l,
 * cly.c:2958).
 */
#include "desamba_oracle.c"
#include <pthread.h>

typedef struct { uint32_t ref_ID, t_st, t_ed, q_st, q_ed, sum_score, indel; uint8_t direction, primary, pri_index, pad; } capi_hit;
typedef struct { uint64_t hit_off; uint32_t n_hit, n_anchor; uint8_t fast_classify, entered_final; uint16_t error; uint32_t read_len; } capi_rr;
typedef struct { uint32_t offset; uint16_t len; uint8_t top, pad; } capi_seed;
typedef struct { orc_index ix; } capi_handle;

void *orc_capi_open(const char *dir, int l_min_match, int min_score)
{
	capi_handle *h = (capi_handle *)calloc(1, sizeof *h);
	if (orc_index_load(&h->ix, dir)) { free(h); return NULL; }
	orc_set_opts(&h->ix, l_min_match, min_score);
	return h;
}
void orc_capi_close(void *h_) { capi_handle *h = (capi_handle *)h_; if (h) { orc_index_free(&h->ix); free(h); } }
int orc_capi_l_ek(void *h_) { return ((capi_handle *)h_)->ix.l_ek; }

/* m_bin_read carried between calls (policy P3 of the oracle): set before / read after orc_capi_classify */
static __thread uint32_t g_m_bin_read = 0;
void orc_capi_set_m_bin_read(uint32_t m) { g_m_bin_read = m; }
uint32_t orc_capi_get_m_bin_read(void) { return g_m_bin_read; }

int64_t orc_capi_classify(void *h_, const char *seqs, const uint64_t *offs, uint32_t n, int32_t max_read_l_in, int32_t *max_read_l_out,
                          capi_rr *rr, capi_hit *hits, uint64_t cap)
{
	capi_handle *h = (capi_handle *)h_;
	orc_buff *buff = orc_buff_new();
	buff->max_read_l = max_read_l_in;
	buff->m_bin_read = g_m_bin_read;
	orc_result res; memset(&res, 0, sizeof res);
	uint64_t used = 0; int64_t ret = 0;
	for (uint32_t r = 0; r < n; r++) {
		uint32_t L = (uint32_t)(offs[r + 1] - offs[r]);
		orc_classify_seq(&h->ix, seqs + offs[r], L, &res, buff);
		rr[r].hit_off = used; rr[r].n_hit = (uint32_t)res.n_hit; rr[r].n_anchor = (uint32_t)res.n_anc;
		rr[r].fast_classify = (uint8_t)res.fast_classify; rr[r].entered_final = (uint8_t)res.entered_final; rr[r].error = 0; rr[r].read_len = L;
		if (used + res.n_hit > cap) { ret = -1; break; }
		for (size_t i = 0; i < res.n_hit; i++) {
			const orc_chain *c = res.hit + i; capi_hit *o = hits + used + i;
			o->ref_ID = c->ref_ID; o->t_st = c->t_st; o->t_ed = c->t_ed; o->q_st = c->q_st; o->q_ed = c->q_ed; o->sum_score = c->sum_score;
			o->indel = c->indel; o->direction = c->direction; o->primary = c->primary; o->pri_index = c->pri_index; o->pad = 0;
		}
		used += res.n_hit;
	}
	if (max_read_l_out) *max_read_l_out = buff->max_read_l;
	g_m_bin_read = buff->m_bin_read;
	orc_result_free(&res); orc_buff_free(buff);
	return ret < 0 ? ret : (int64_t)used;
}

/* island seeds of one read; strand 0 forward, 1 reverse (getIsland, cly.c:1236-1268) */
int orc_capi_seeds(void *h_, const char *seq, uint32_t L, int strand, capi_seed *out, uint32_t cap, uint32_t *total_score)
{
	capi_handle *h = (capi_handle *)h_;
	if (L < MIN_READ_LEN) { if (total_score) *total_score = 0; return 0; }
	orc_buff *buff = orc_buff_new();
	orc_result res; memset(&res, 0, sizeof res);
	search_dir_t sd[2];
	get_island(&h->ix, seq, L, buff, &res, sd);
	uint32_t n = res.n_seeds[strand];
	if (total_score) *total_score = res.total_score[strand];
	for (uint32_t i = 0; i < n && i < cap; i++) { out[i].offset = res.seeds[strand][i].offset; out[i].len = (uint16_t)res.seeds[strand][i].len; out[i].top = res.seeds[strand][i].top; out[i].pad = 0; }
	orc_result_free(&res); orc_buff_free(buff);
	return (int)n;
}

void orc_capi_counters(uint64_t out[16], int reset)
{
	uint64_t v[16] = {orc_cnt.n_hits, orc_cnt.n_reads, 0, 0, orc_cnt.n_bit0, orc_cnt.n_bit1, orc_cnt.n_prefix, orc_cnt.n_occ, orc_cnt.n_locate,
	                  orc_cnt.n_getref, orc_cnt.n_getref_bytes, 0, orc_cnt.n_bases, 0, 0, 0};
	memcpy(out, v, sizeof v);
	if (reset) memset(&orc_cnt, 0, sizeof orc_cnt);
}

/* multi-threaded throughput run for bench.py's CPU legs ("port" baseline): reads are dealt to threads in contiguous chunks;
 * results are discarded except the hit count (timing only; parity uses orc_capi_classify). */
typedef struct { capi_handle *h; const char *seqs; const uint64_t *offs; uint32_t lo, hi; uint64_t n_hits; } mt_job;
static void *mt_worker(void *a_)
{
	mt_job *a = (mt_job *)a_;
	orc_buff *buff = orc_buff_new();
	orc_result res; memset(&res, 0, sizeof res);
	for (uint32_t r = a->lo; r < a->hi; r++) {
		orc_classify_seq(&a->h->ix, a->seqs + a->offs[r], (uint32_t)(a->offs[r + 1] - a->offs[r]), &res, buff);
		a->n_hits += res.n_hit;
	}
	orc_result_free(&res); orc_buff_free(buff);
	return NULL;
}
uint64_t orc_capi_classify_mt(void *h_, const char *seqs, const uint64_t *offs, uint32_t n, int n_threads)
{
	if (n_threads < 1) n_threads = 1;
	if (n_threads > 256) n_threads = 256;
	pthread_t th[256]; mt_job job[256];
	for (int t = 0; t < n_threads; t++) {
		job[t].h = (capi_handle *)h_; job[t].seqs = seqs; job[t].offs = offs; job[t].n_hits = 0;
		job[t].lo = (uint32_t)((uint64_t)n * t / n_threads); job[t].hi = (uint32_t)((uint64_t)n * (t + 1) / n_threads);
		pthread_create(&th[t], NULL, mt_worker, &job[t]);
	}
	uint64_t tot = 0;
	for (int t = 0; t < n_threads; t++) { pthread_join(th[t], NULL); tot += job[t].n_hits; }
	return tot;
}

/* anchors of ONE seeding pass of one read, for the anchor-level parity test of the seeding engine:
 * dir = index into search_dir after the swap of getIsland (cly.c:1261-1266), slow = 0 fast_classify (cly.c:1476-1546) /
 * 1 slow_classify (cly.c:1548-1611).  out[]: {ref_ID, ref_offset, index_in_read, mtch_len | score << 16, direction | useless << 8}.
 * *strand_out = 0 forward / 1 reverse strand of that search direction; counters[5] = prefix look-ups, occ, locates,
 * get_ref calls, get_ref bytes of the pass. */
typedef struct { uint32_t ref_ID, ref_offset, index_in_read, len_score, dir_useless; } capi_anchor;
int orc_capi_seed_pass(void *h_, const char *seq, uint32_t L, int dir, int slow, capi_anchor *out, uint32_t cap, int *strand_out, int *both_out, uint64_t *counters)
{
	capi_handle *h = (capi_handle *)h_;
	if (L < MIN_READ_LEN) return 0;
	orc_buff *buff = orc_buff_new();
	orc_result res; memset(&res, 0, sizeof res);
	search_dir_t sd[2];
	get_island(&h->ix, seq, L, buff, &res, sd);
	if (strand_out) *strand_out = (sd[dir].direction == FORWARD) ? 0 : 1;
	if (both_out) *both_out = ((sd[0].total_score - sd[1].total_score) <= (sd[0].total_score >> 3)) ? 1 : 0;
	orc_counters before = orc_cnt;
	res.n_anc = 0;
	if (slow) slow_classify(&h->ix, sd + dir, L, &res); else fast_classify(&h->ix, sd + dir, L, &res);
	if (counters) {
		counters[0] = orc_cnt.n_prefix - before.n_prefix; counters[1] = orc_cnt.n_occ - before.n_occ; counters[2] = orc_cnt.n_locate - before.n_locate;
		counters[3] = orc_cnt.n_getref - before.n_getref; counters[4] = orc_cnt.n_getref_bytes - before.n_getref_bytes;
	}
	int n = (int)res.n_anc;
	for (int i = 0; i < n && (uint32_t)i < cap; i++) {
		const orc_anchor *a = res.anc + i;
		out[i].ref_ID = a->ref_ID; out[i].ref_offset = a->ref_offset; out[i].index_in_read = a->index_in_read;
		out[i].len_score = (uint32_t)a->mtch_len | ((uint32_t)(uint16_t)a->score << 16);
		out[i].dir_useless = (uint32_t)a->direction | ((uint32_t)a->anchor_useless << 8);
	}
	orc_result_free(&res); orc_buff_free(buff);
	return n;
}
