/*
 * desamba_oracle.c -- TEST INFRASTRUCTURE ONLY (parity oracle); see desamba_oracle.h.
 *
 * A sequential CPU restatement of deSAMBA's per-read classifier, written from the behaviour of
 * /root/reference/src (cly.c, bwt.c, idx.c, cly_mt.c, lib/utils.c); each function cites the lines it follows.
 * Pointers of the reference are indices here, realloc'd vectors are explicit arrays, qsort is an explicit
 * emulation of glibc's merge sort.  All arithmetic mirrors the DECLARED C types of the reference (mixed
 * signed/unsigned MAX/MIN/ABS macros included) because results depend on the usual arithmetic conversions.
 *
 * Undefined-behaviour policy.  The parity target is O_def = the unmodified reference built with
 * -ftrivial-auto-var-init=zero and run with -t 1.  Where the reference reads memory it never wrote we define:
 *   P1  stack windows (ref[2000] per anchor pair, ref[1000] per extension call, LV flank arrays, LV work arrays)
 *       read 0 where never loaded, and keep bytes of earlier loads of the same call (zero-init semantics).
 *   P2  the three 13-byte flank arrays of map_seed are adjacent in the order q_pre, t_pre, t_suf and get_new_ed's
 *       q_buff/t_buff reuse the t_pre/t_suf slots (frame layout of the O_def binary, objdump of map_seed);
 *       the 5 padding bytes below q_pre read 0.
 *   P3  the two read strands are contiguous (forward, then reverse complement: cly.c:1247-1259).  Bytes after the
 *       reverse strand read ORC_OOB (0 = 'A': untouched realloc slack).  Bytes BEFORE the forward strand are the
 *       glibc chunk header of the reference's bin_read buffer (BUFF_REALLOC, utils.h:117-122): [-1..-6] = 0 (high
 *       bytes of the size field), [-7] = (chunk_size >> 8) & 0xff -- a base code 0..3 for buffers of 256..1023 bytes,
 *       i.e. for short reads -- and [-8] = low size byte | flags, never a base.  chunk_size follows from the running
 *       capacity m_bin_read of the buffer (grows to 2*len+20 whenever 2*len exceeds it), which is carried over the
 *       reads in input order (-t 1).  Pinned by tests/golden/syn_short1 (a read starting right of a 7-A run).
 */
#include "desamba_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_OOB 0
#define GUARD 128

/* utils.h:61-64 -- same textual macros so that mixed-type operands convert exactly as in the reference */
#define MAX(a,b) (((a) > (b))?(a):(b))
#define MIN(a,b) (((a) < (b))?(a):(b))
#define ABS(a) (((a) > 0)?(a): (- (a)))
#define ABS_U(a,b) (((a) > (b))?((a) - (b)): ((b) - (a)))

#define FORWARD 1
#define REVERSE 0
#define L_PRE_IDX 13
#define PRE_IDX_MASK 0x3FFFFFF
#define SA_MASK 0x7
#define SA_OFF 3
#define MIN_UNI_L 35
#define LV_L 12
#define S_A_KEMR_L 9
#define OVER_SEARCH_M2 50
#define MIN_SCORE_MEM 12
#define NO_SA 0xFFFFFFFFFFFFFFFFull

__thread orc_counters orc_cnt;   /* per thread: the multi-threaded timing run must not share a counter cache line */

/* ------------------------------------------------------------------ index loading (idx.c:1103-1160, bwt.c:68-104) */
static void *slurp(const char *dir, const char *ext, size_t skip_hdr, uint64_t *hdr, size_t elem, size_t extra_elems)
{
	char path[4096];
	snprintf(path, sizeof path, "%s/deSAMBA%s", dir, ext);
	FILE *f = fopen(path, "rb");
	if (!f) { fprintf(stderr, "[oracle] cannot open %s\n", path); return NULL; }
	if (skip_hdr && fread(hdr, 8, 1, f) != 1) { fclose(f); return NULL; }
	uint64_t n = *hdr;
	uint8_t *p = (uint8_t *)calloc(n + extra_elems, elem);
	if (!p || fread(p, elem, n, f) != n) { fprintf(stderr, "[oracle] short read %s\n", path); fclose(f); free(p); return NULL; }
	fclose(f);
	return p;
}

static void mapq_tables(orc_index *ix)     /* cly_mt.c:413-437, called with P_E = 0.15, L_REF = ref_bin.n*4 (cly_mt.c:527) */
{
	double P_E = 0.15; uint64_t L_REF = ix->ref_bin_n * 4;
	double REF_SIZE_PUNALTY = -10 * log(L_REF) / log(10);
	double MATCH_SCORE = -10 * log(0.25 / (1 - P_E)) / log(10);
	double MISMATCH_PUNALTY = -10 * log(0.75 / (P_E)) / log(10);
	for (int i = 0; i < 2000; i++)
		ix->Q_MEM[i] = REF_SIZE_PUNALTY + i * MATCH_SCORE + 0.5;
	for (int j = 0; j < 20; j++)
		for (int i = 0; i < 20; i++) {
			ix->Q_LV[i][j] = (j - i) * MATCH_SCORE + i * MISMATCH_PUNALTY + 0.5;
			if (j < 5) ix->Q_LV[i][j] += 15;
			ix->Q_LV[i][j] = MAX(ix->Q_LV[i][j], -8);
		}
}

void orc_set_opts(orc_index *ix, int l_min_match, int min_score)   /* cly_mt.c:521-527 */
{
	ix->filter_min_length = l_min_match;
	ix->filter_min_score = min_score;
	ix->filter_min_score_LV3 = min_score + 10;
	mapq_tables(ix);
}

int orc_index_load(orc_index *ix, const char *dir)
{
	memset(ix, 0, sizeof *ix);
	char path[4096];
	snprintf(path, sizeof path, "%s/deSAMBA.bwt", dir);
	FILE *f = fopen(path, "rb");
	if (!f) { fprintf(stderr, "[oracle] cannot open %s\n", path); return -1; }
	if (fread(&ix->byteLen, 8, 1, f) != 1) return -1;
	ix->bwt_occ = (uint8_t *)calloc(ix->byteLen + 168, 1);           /* one zero block of slack: occ(len_bwt) when len%256==0 */
	if (fread(ix->bwt_occ, 1, ix->byteLen, f) != ix->byteLen) return -1;
	if (fread(ix->rank, 8, 5, f) != 5) return -1;
	ix->rank[5] = ix->rank[0] - 1;                                   /* bwt.c:81 */
	uint64_t nh = (1ull << (L_PRE_IDX << 1)) + 1;
	ix->hash_index = (uint64_t *)malloc(nh * 8);
	if (fread(ix->hash_index, 8, nh, f) != nh) return -1;
	fclose(f);
	ix->sa = (orc_sa_t *)slurp(dir, ".sa", 1, &ix->sa_size, sizeof(orc_sa_t), 0);
	uint64_t eks = 0;
	snprintf(path, sizeof path, "%s/deSAMBA.exki", dir);
	f = fopen(path, "rb");
	if (!f || fread(&eks, 8, 1, f) != 1) return -1;
	fclose(f);
	ix->ek_size = eks;
	/* set_ekmer_par, idx.c:966-982 */
	ix->ek_mask = (1ull << 37) - 1; ix->l_ek = 20;
	switch (eks >> 27) {
		case 1:   ix->ek_mask = (1ull << 30) - 1; ix->l_ek = 16; break;
		case 2:   ix->ek_mask = (1ull << 31) - 1; ix->l_ek = 17; break;
		case 4:   ix->ek_mask = (1ull << 32) - 1; ix->l_ek = 17; break;
		case 8:   ix->ek_mask = (1ull << 33) - 1; ix->l_ek = 18; break;
		case 16:  ix->ek_mask = (1ull << 34) - 1; ix->l_ek = 18; break;
		case 32:  ix->ek_mask = (1ull << 35) - 1; ix->l_ek = 19; break;
		case 64:  ix->ek_mask = (1ull << 36) - 1; ix->l_ek = 19; break;
		case 128: ix->ek_mask = (1ull << 37) - 1; ix->l_ek = 20; break;
	}
	ix->single_base_max = 0.8 * ix->l_ek;
	uint64_t n = eks;
	ix->ek0 = (uint8_t *)slurp(dir, ".exk0", 0, &n, 1, 0);
	ix->ek1 = (uint8_t *)slurp(dir, ".exk1", 0, &n, 1, 0);
	ix->uni = (orc_unitig_t *)slurp(dir, ".unv", 1, &ix->n_uni, sizeof(orc_unitig_t), 1);
	if (!ix->sa || !ix->ek0 || !ix->ek1 || !ix->uni) return -1;
	ix->uni[ix->n_uni].ref_list = ix->uni[ix->n_uni - 1].ref_list + 1 + ix->uni[ix->n_uni - 1].length; /* idx.c:1127 */
	ix->uni[ix->n_uni].length = 0;
	ix->dollar_pos = ix->n_uni - 1 - 1;                              /* idx.c:1128 */
	ix->ref_bin = (uint8_t *)slurp(dir, ".ref_b", 1, &ix->ref_bin_n, 1, 1024);
	ix->ri = (orc_refinfo_t *)slurp(dir, ".ref_i", 1, &ix->n_ri, sizeof(orc_refinfo_t), 0);
	ix->ref_pos = (uint64_t *)slurp(dir, ".ref_p", 1, &ix->n_rp, 8, 0);
	if (!ix->ref_bin || !ix->ri || !ix->ref_pos) return -1;
	orc_set_opts(ix, 170, 64);
	return 0;
}

void orc_index_free(orc_index *ix)
{
	free(ix->bwt_occ); free(ix->hash_index); free(ix->sa); free(ix->ek0); free(ix->ek1);
	free(ix->uni); free(ix->ref_bin); free(ix->ri); free(ix->ref_pos);
	memset(ix, 0, sizeof *ix);
}

/* ------------------------------------------------------------------ small primitives */
uint64_t orc_hash64_1(uint64_t key)        /* utils.c:1067-1077 */
{
	key = (~key + (key << 21));
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8));
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4));
	key = key ^ key >> 28;
	key = (key + (key << 31));
	return key;
}

uint64_t orc_hash64_2(uint64_t key)        /* utils.c:1080-1091 */
{
	key += ~(key << 32);
	key ^= (key >> 22);
	key += ~(key << 13);
	key ^= (key >> 8);
	key += (key << 3);
	key ^= (key >> 15);
	key += ~(key << 27);
	key ^= (key >> 31);
	return key;
}

int orc_exist_kmer(const orc_index *ix, uint64_t kmer)   /* cly.c:956-972 */
{
	if (kmer == 0) return 0;
	orc_cnt.n_bit0++;
	uint64_t h1 = orc_hash64_1(kmer) & ix->ek_mask;
	if (((ix->ek0[h1 >> 3] >> (7 - (h1 & 7))) & 1) == 0) return 0;
	orc_cnt.n_bit1++;
	uint64_t h2 = orc_hash64_2(kmer) & ix->ek_mask;
	return (ix->ek1[h2 >> 3] >> (7 - (h2 & 7))) & 1;
}

/* count nibbles equal to c among the first n (0..255) nibbles of a 128-byte block payload */
static inline uint32_t nib_count(const uint8_t *data, uint32_t n, uint8_t c)
{
	uint32_t cnt = 0;
	uint64_t pat = 0x1111111111111111ull * c;
	for (uint32_t w = 0; w * 16 < n; w++) {
		uint64_t x;
		memcpy(&x, data + w * 8, 8);
		x ^= pat;
		x |= x >> 1; x |= x >> 2;
		uint64_t m = ~x & 0x1111111111111111ull;
		uint32_t rem = n - w * 16;
		if (rem < 16) m &= (1ull << (rem * 4)) - 1;
		cnt += (uint32_t)__builtin_popcountll(m);
	}
	return cnt;
}

uint64_t orc_occ(const orc_index *ix, uint64_t r, uint8_t *c)   /* bwt.c:43-65 (LUT replaced by nibble compare; same counts) */
{
	orc_cnt.n_occ++;
	const uint8_t *blk = ix->bwt_occ + (r >> 8) * 168;
	uint32_t in = (uint32_t)(r & 0xff);
	if (*c == 0xff) {
		*c = (blk[40 + (in >> 1)] >> ((in & 1) << 2)) & 0xf;
		if (*c == 5) return ix->dollar_pos;
	}
	uint64_t base;
	memcpy(&base, blk + ((*c) << 3), 8);
	return base + nib_count(blk + 40, in, *c);
}

void orc_get_ref(const uint8_t *ref_bin, uint8_t *out, int64_t off, int32_t length, int forward)   /* cly.c:435-466 */
{
	if (off < 0) off = 0;
	if (length < 0) length = 0;
	orc_cnt.n_getref++; orc_cnt.n_getref_bytes += (uint64_t)(length + 3) / 4;
	uint64_t o = (uint64_t)off;
	if (forward)
		for (uint32_t k = 0; k < (uint32_t)length; k++, o++)
			out[k] = (ref_bin[o >> 2] >> ((3 - (o & 3)) << 1)) & 3;
	else
		for (uint32_t k = 0; k < (uint32_t)length; k++, o--)   /* o wraps below 0 exactly like the reference's offset-- */
			out[k] = (ref_bin[o >> 2] >> ((3 - (o & 3)) << 1)) & 3;
}

/* glibc 2.39 msort_with_tmp: n1 = n/2 (left), n2 = n - n1; merge takes from the left run while cmp(l, r) <= 0 */
static void msort_rec(char *b, size_t n, size_t s, int (*cmp)(const void *, const void *), char *tmp)
{
	if (n <= 1) return;
	size_t n1 = n / 2, n2 = n - n1;
	char *b1 = b, *b2 = b + n1 * s;
	msort_rec(b1, n1, s, cmp, tmp);
	msort_rec(b2, n2, s, cmp, tmp);
	char *t = tmp;
	while (n1 > 0 && n2 > 0) {
		if (cmp(b1, b2) <= 0) { memcpy(t, b1, s); b1 += s; n1--; }
		else { memcpy(t, b2, s); b2 += s; n2--; }
		t += s;
	}
	if (n1 > 0) memcpy(t, b1, n1 * s);
	memcpy(b, tmp, (n - n2) * s);
}

void orc_msort(void *base, size_t n, size_t size, int (*cmp)(const void *, const void *))
{
	if (n <= 1) return;
	char *tmp = (char *)malloc(n * size);
	msort_rec((char *)base, n, size, cmp, tmp);
	free(tmp);
}

/* ------------------------------------------------------------------ scratch */
typedef struct { uint32_t t_pos, q_pos, len, score; } spd_match;           /* cly.h:127-133 */
typedef struct { uint32_t kmer, next, pos; } sa_hash_t;                    /* cly.h:102-107 */
typedef struct { uint16_t next; uint16_t seed_ID; uint8_t s_or_e; } sc_hash_t; /* cly.h:120-125 (15+1 bit field) */

struct orc_buff {
	uint8_t *bin_base; size_t m_bin;          /* GUARD | fwd L | rev L | GUARD */
	uint64_t *kmer; size_t m_kmer;            /* fwd L | rev L */
	sa_hash_t *sa_hash[2]; size_t m_sa_hash[2];
	sc_hash_t *sc_hash; size_t m_sc_hash;
	spd_match *sms; size_t n_sms, m_sms;
	int max_read_l;
	uint32_t m_bin_read;                      /* capacity of the reference's bin_read buffer (policy P3) */
};

orc_buff *orc_buff_new(void) { return (orc_buff *)calloc(1, sizeof(orc_buff)); }
void orc_buff_free(orc_buff *b)
{
	if (!b) return;
	free(b->bin_base); free(b->kmer); free(b->sa_hash[0]); free(b->sa_hash[1]); free(b->sc_hash); free(b->sms); free(b);
}
void orc_result_free(orc_result *r)
{
	free(r->hit); free(r->anc); free(r->seeds[0]); free(r->seeds[1]);
	memset(r, 0, sizeof *r);
}

static spd_match *sms_push(orc_buff *b)       /* kv_pushp_2: no zeroing of the new slot (kvec.h:103-109) */
{
	if (b->n_sms == b->m_sms) {
		b->m_sms = b->m_sms ? b->m_sms << 1 : 10;
		b->sms = (spd_match *)realloc(b->sms, b->m_sms * sizeof(spd_match));
	}
	return b->sms + b->n_sms++;
}

static orc_anchor *anchor_push(orc_result *r)
{
	if (r->n_anc == r->m_anc) {
		r->m_anc = r->m_anc ? r->m_anc << 1 : 10;
		r->anc = (orc_anchor *)realloc(r->anc, r->m_anc * sizeof(orc_anchor));
	}
	orc_anchor *a = r->anc + r->n_anc++;
	memset(a, 0, sizeof *a);
	a->pre = -1;
	return a;
}

static orc_chain *chain_push(orc_result *r)
{
	if (r->n_hit == r->m_hit) {
		r->m_hit = r->m_hit ? r->m_hit << 1 : 10;
		r->hit = (orc_chain *)realloc(r->hit, r->m_hit * sizeof(orc_chain));
	}
	orc_chain *c = r->hit + r->n_hit++;
	memset(c, 0, sizeof *c);
	return c;
}

typedef struct {                /* SEARCH_DIR, cly.c:946-954 */
	orc_seed *seed_v; uint32_t l_seed_v;
	uint8_t *bin_read; uint64_t *kmer;
	uint32_t direction, total_score;
} search_dir_t;

/* ------------------------------------------------------------------ islands (cly.c:360-398, 1071-1268) */
static void store_kmers(const uint8_t *bin, uint32_t n_kmer, uint8_t l_ek, int single_base_max, uint64_t *out)
{
	int cnt[4] = {0, 0, 0, 0};
	uint64_t mask = (l_ek == 32) ? ~0ull : ((1ull << (2 * l_ek)) - 1);
	uint64_t kmer = 0;
	for (uint32_t i = 0; i < l_ek; i++) { cnt[bin[i]]++; kmer = (kmer << 2) | bin[i]; }
	for (uint32_t i = 0; i < n_kmer; i++) {
		if (i) {
			cnt[bin[i - 1]]--; cnt[bin[i + l_ek - 1]]++;
			kmer = ((kmer << 2) | bin[i + l_ek - 1]) & mask;
		}
		int bad = cnt[0] >= single_base_max || cnt[1] >= single_base_max || cnt[2] >= single_base_max || cnt[3] >= single_base_max;
		out[i] = bad ? 0 : kmer;
	}
}

#define STEP_EK 3
static uint32_t scan_islands(const orc_index *ix, const uint64_t *kmer_v, uint32_t l_kmer_v, orc_seed *seed_v, uint32_t direction)
{   /* search_exist_kmer_M2, cly.c:1071-1160 */
	uint32_t l_seed_v = 0;
	if (direction == FORWARD) {
		for (uint32_t i = STEP_EK - 1; i < l_kmer_v; i += STEP_EK) {
			if (orc_exist_kmer(ix, kmer_v[i]) == 1) {
				uint32_t offset = i, len = 1;
				for (int j = 1; j < STEP_EK; ++j) {
					if (orc_exist_kmer(ix, kmer_v[i - j]) == 1) { offset--; len++; }
					else break;
				}
				for (int j = 1; i + j < l_kmer_v; ++j) {
					if (orc_exist_kmer(ix, kmer_v[i + j]) == 1) { len++; if (len > 60) break; }
					else break;
				}
				seed_v[l_seed_v].offset = offset;
				seed_v[l_seed_v].len = len;
				l_seed_v++;
				i = offset + len;
			}
		}
	} else {
		for (int i = l_kmer_v - STEP_EK; i >= 0; i -= STEP_EK) {
			if (orc_exist_kmer(ix, kmer_v[i]) == 1) {
				uint32_t offset = i, len = 1;
				for (int j = 1; j < STEP_EK; ++j) {
					if (orc_exist_kmer(ix, kmer_v[i + j]) == 1) { offset++; len++; }
					else break;
				}
				for (int j = 1; j <= i; ++j) {
					if (orc_exist_kmer(ix, kmer_v[i - j]) == 1) { len++; if (len > 60) break; }
					else break;
				}
				seed_v[l_seed_v].offset = offset - len + 1;
				seed_v[l_seed_v].len = len;
				l_seed_v++;
				i = offset - len;
			}
		}
	}
	return l_seed_v;
}

#define SEED_RANGE 100
static void seed_vector(const orc_index *ix, uint8_t *bin, uint64_t *kmer_buff, uint32_t l_kmer_buff,
                        orc_seed *seed_v, uint32_t direction, search_dir_t *sd)
{   /* get_seed_vector_M2, cly.c:1162-1234 */
	store_kmers(bin, l_kmer_buff, ix->l_ek, ix->single_base_max, kmer_buff);
	uint32_t l_seed_v = scan_islands(ix, kmer_buff, l_kmer_buff, seed_v, direction);
	uint32_t total_score = 0;
	int max_index = 0; uint32_t max_length = 0; uint32_t index_end = SEED_RANGE;
	for (uint32_t m = 0; m < l_seed_v; m++) {
		seed_v[m].top = 0;
		uint32_t wpos = (direction == FORWARD) ? seed_v[m].offset : l_kmer_buff - seed_v[m].offset - seed_v[m].len;
		if (wpos < index_end) {
			if (max_length < seed_v[m].len) { max_length = seed_v[m].len; max_index = m; }
			seed_v[max_index].top = 0;
		} else {
			seed_v[max_index].top = 1;
			index_end += SEED_RANGE;
			total_score += max_length;
			max_index = m;
			max_length = seed_v[m].len;
		}
	}
	seed_v[max_index].top = 1;     /* written even with zero seeds (slot 0 always exists here) */
	total_score += max_length;
	sd->seed_v = seed_v; sd->l_seed_v = l_seed_v; sd->bin_read = bin; sd->kmer = kmer_buff;
	sd->direction = direction; sd->total_score = total_score;
}

static const uint8_t cly_bit(char ch)     /* CLY_Bit, cly.c:17-35: A0 C1 G2 T3 (either case), everything else 1 */
{
	switch (ch) { case 'A': case 'a': return 0; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 1; }
}

static void get_island(const orc_index *ix, const char *seq, uint32_t read_len, orc_buff *buff, orc_result *res, search_dir_t *sd)
{   /* getIsland, cly.c:1236-1268 */
	size_t need = (size_t)read_len * 2 + 2 * GUARD;
	if (need > buff->m_bin) { buff->m_bin = need + 64; buff->bin_base = (uint8_t *)realloc(buff->bin_base, buff->m_bin); }
	if ((size_t)read_len * 2 > buff->m_kmer) { buff->m_kmer = (size_t)read_len * 2 + 64; buff->kmer = (uint64_t *)realloc(buff->kmer, buff->m_kmer * 8); }
	for (int s = 0; s < 2; s++)
		if ((read_len >> 2) + 2 > res->m_seeds[s]) { res->m_seeds[s] = (read_len >> 2) + 64; res->seeds[s] = (orc_seed *)realloc(res->seeds[s], res->m_seeds[s] * sizeof(orc_seed)); }
	uint32_t l_kmer_buff = read_len - ix->l_ek + 1;
	uint8_t *bin_F = buff->bin_base + GUARD, *bin_R = bin_F + read_len;
	memset(buff->bin_base, ORC_OOB, GUARD);
	memset(bin_R + read_len, ORC_OOB, GUARD);
	{	/* policy P3: the chunk header in front of the reference's buffer */
		if (2 * read_len > buff->m_bin_read) buff->m_bin_read = 2 * read_len + 20;          /* BUFF_REALLOC */
		uint32_t chunk = (buff->m_bin_read + 8 + 15) & ~15u;
		if (chunk < 32) chunk = 32;
		bin_F[-7] = (uint8_t)((chunk >> 8) & 0xff);
		bin_F[-8] = 0xff;
	}
	for (uint32_t k = 0; k < read_len; ++k) bin_F[k] = cly_bit(seq[k]);
	seed_vector(ix, bin_F, buff->kmer, l_kmer_buff, res->seeds[0], FORWARD, sd);
	for (uint32_t k = 0; k < read_len; ++k) bin_R[read_len - k - 1] = 3 - bin_F[k];
	seed_vector(ix, bin_R, buff->kmer + read_len, l_kmer_buff, res->seeds[1], REVERSE, sd + 1);
	res->n_seeds[0] = sd[0].l_seed_v; res->n_seeds[1] = sd[1].l_seed_v;
	res->total_score[0] = sd[0].total_score; res->total_score[1] = sd[1].total_score;
	if (sd[0].total_score < sd[1].total_score) { search_dir_t t = sd[0]; sd[0] = sd[1]; sd[1] = t; }
}

/* ------------------------------------------------------------------ FM-index search (cly.c:1286-1298, 1344-1447) */
typedef struct { uint64_t *set; int l, m; } sp_set_t;
static inline int sp_set_insert(uint64_t node, sp_set_t *s)
{
	if (s->l == s->m) s->l = 0;
	int i = 0;
	for (; i < s->l; i++) if (s->set[i] == node) return 0;
	s->set[i] = node;
	s->l++;
	return 1;
}

typedef struct { int match_len; uint64_t sp, sa_sp; int sa_sp_l, kmer_index, read_offset; } mem_rst_t;  /* cly.c:619-627 */

static void bwt_single_search(const orc_index *ix, uint64_t sp, const uint8_t *string, int max_match_len, sp_set_t *sp_set, mem_rst_t *out)
{   /* cly.c:1344-1383 */
	uint64_t new_sp, sa_sp = NO_SA;
	int match_len = 0, sa_sp_l = 0;
	while (1) {
		if (match_len >= max_match_len) break;
		if ((sp & SA_MASK) == 0) { sa_sp = sp; sa_sp_l = 0; }
		else sa_sp_l--;
		uint8_t c = 0xff;
		new_sp = orc_occ(ix, sp, &c) + ix->rank[c];
		if (c != *string) break;
		match_len++;
		string--;
		if (sp_set_insert(new_sp, sp_set) == 0) { out->match_len = -1000; return; }
		sp = new_sp;
	}
	out->sp = sp; out->match_len = match_len; out->sa_sp = sa_sp; out->sa_sp_l = sa_sp_l;
}

static int bwt_MEM_search(const orc_index *ix, const uint8_t *string, uint64_t pre_v, int max_rst, int l_min_mth, int l_max_mth,
                          sp_set_t *sp_set, mem_rst_t *mem_rst)
{   /* cly.c:1388-1447 */
	int n_rst = 0;
	const uint64_t *rank = ix->rank;
	orc_cnt.n_prefix++;
	uint64_t sp = ix->hash_index[pre_v], ep = ix->hash_index[pre_v + 1], new_sp, new_ep;
	string -= L_PRE_IDX;
	int match_len = L_PRE_IDX;
	uint8_t c;
	while (1) {
		c = *string;
		string--;
		new_sp = rank[c] + orc_occ(ix, sp, &c);
		new_ep = rank[c] + orc_occ(ix, ep, &c);
		if (match_len >= l_min_mth - 1) {
			if (new_sp + max_rst >= new_ep) break;
			if (match_len >= l_max_mth) return 0;
		}
		if (new_sp + 1 >= new_ep) break;
		match_len++;
		sp = new_sp; ep = new_ep;
	}
	if (new_sp >= new_ep) return 0;
	if (new_sp + 1 == new_ep) {
		if (sp_set_insert(new_sp, sp_set) == 0) return 0;
		bwt_single_search(ix, new_sp, string, MAX(0, l_max_mth - match_len), sp_set, mem_rst + n_rst);
		mem_rst[n_rst].match_len += match_len + 1;
		if (mem_rst[n_rst].match_len >= l_min_mth) n_rst++;
	} else {
		for (uint64_t c_sp = new_sp; c_sp < new_ep; c_sp++) {
			if (sp_set_insert(c_sp, sp_set) == 0) continue;
			bwt_single_search(ix, c_sp, string, MAX(0, l_max_mth - match_len), sp_set, mem_rst + n_rst);
			mem_rst[n_rst].match_len += match_len + 1;
			if (mem_rst[n_rst].match_len >= l_min_mth) n_rst++;
		}
	}
	return n_rst;
}

/* ------------------------------------------------------------------ Landau-Vishkin flank (cly.c:510-609) */
#define LV_ERROR 4
#define LV_BASE LV_ERROR
int32_t orc_lv_extd(uint8_t *ref, int32_t ref_length, uint8_t *query, int32_t query_length)
{
	if (ref_length < query_length) {
		int32_t t = ref_length; ref_length = query_length; query_length = t;
		uint8_t *p = ref; ref = query; query = p;
	}
	int32_t mn_data[99], ed_data[99];
	memset(mn_data, 0, sizeof mn_data); memset(ed_data, 0, sizeof ed_data);   /* policy P1 */
	int32_t *mn = mn_data + LV_BASE + 1, *ed = ed_data + LV_BASE + 1;
	int32_t prev_mn, cur_mn, next_mn, prev_ed, cur_ed, next_ed;
	uint8_t old_ref_end = ref[ref_length], old_query_end = query[query_length];
	ref[ref_length] = '#';
	query[query_length] = '$';
	int32_t best_score = query_length;
	for (int i = -LV_BASE - 1; i <= LV_BASE + 1; i++) { mn[i] = -1; ed[i] = (i > 0) ? (i) : (-i); }
	for (int i = 0; i <= LV_ERROR; i++) {
		prev_mn = -1; cur_mn = (i - 1); next_mn = mn[-i + 1];
		prev_ed = i + 1; cur_ed = i; next_ed = ed[-i + 1];
		for (int j = -i; j <= LV_ERROR; j++) {
			if (cur_mn + j < ref_length - 1) {
				int best = cur_mn + 1 - cur_ed;
				mn[j] = cur_mn + 1; ed[j] = cur_ed + 1;
				if (best < next_mn + 1 - next_ed) { mn[j] = next_mn + 1; ed[j] = next_ed + 1; best = next_mn - next_ed; }
				if (best < prev_mn - prev_ed) { mn[j] = prev_mn + 1; ed[j] = prev_ed + 1; }
			} else {
				int best = cur_mn - cur_ed;
				mn[j] = cur_mn; ed[j] = cur_ed + 1;
				if (best < prev_mn - prev_ed) { mn[j] = prev_mn; ed[j] = prev_ed + 1; best = prev_mn - prev_ed; }
				if (best < next_mn + 1 - next_ed) { mn[j] = next_mn + 1; ed[j] = next_ed + 1; }
			}
			int mn_j = MIN(mn[j], query_length);
			mn_j = MIN(mn_j, ref_length - j);
			for (; ref[mn_j + j] == query[mn_j]; mn_j++);
			mn[j] = mn_j;
			if (query[mn_j] == '$' || ref[mn_j + j] == '#') {
				best_score = MIN(ed[j] - 1, best_score);
				if (j <= i + 1) goto done;
			}
			prev_mn = cur_mn; cur_mn = next_mn; next_mn = mn[j + 2];
			prev_ed = cur_ed; cur_ed = next_ed; next_ed = ed[j + 2];
		}
	}
done:
	ref[ref_length] = old_ref_end;
	query[query_length] = old_query_end;
	return best_score;
}

/* ------------------------------------------------------------------ locate + anchors (cly.c:471-496, 629-694, 706-939) */
typedef struct { uint8_t pad[8]; uint8_t A[13], B[13], C[13]; uint8_t tail[9]; } lv_frame;  /* policy P2: q_pre|t_pre|t_suf adjacent */

typedef struct { uint8_t *bin_read; uint32_t read_L; uint16_t seed_ID; int direction; } seed_info_t;

static int64_t get_uni(const orc_index *ix, uint64_t bwt_pos, int search_l, uint64_t *global_offset, uint32_t *uni_offset_)
{   /* cly.c:471-496; returns the unitig index */
	orc_cnt.n_locate++;
	int64_t u = ix->sa[bwt_pos >> SA_OFF].unitig_ID;
	uint32_t uni_offset = ix->sa[bwt_pos >> SA_OFF].offset + search_l + 1;
	if (search_l > 0)
		for (; uni_offset >= ix->uni[u].length;) { uni_offset -= (ix->uni[u].length + 1); u++; }
	/* the search_l<=0 normalisation loop of the reference tests an unsigned for <0: dead code */
	uint64_t rp = ix->ref_pos[ix->uni[u].ref_list];
	*global_offset = (rp & 0xFFFFFFFFFFull) + uni_offset;
	*uni_offset_ = uni_offset;
	return u;
}

static void get_new_ed(const orc_index *ix, lv_frame *fr, uint32_t *e_d, uint32_t *len_, uint32_t *l_mem_ext,
                       int32_t q_off, uint64_t t_off, uint32_t l_read, uint8_t *q_b, int is_FWD)
{   /* cly.c:629-694 */
	const uint8_t *t_b = ix->ref_bin;
	memset(fr->B, 0, 13); memset(fr->C, 0, 13);
	uint8_t *q = fr->B, *t = fr->C;
	uint32_t len, max_len;
	if (is_FWD) {
		if (q_off < 0) q_off = 0;
		max_len = q_off;
		len = MIN(12, max_len);
		for (uint8_t k = 0; k < len; k++) q[k] = q_b[q_off - k];
	} else {
		max_len = l_read - q_off;
		len = MIN(12, max_len);
		q = q_b + q_off;
	}
	orc_get_ref(t_b, t, t_off, len, !is_FWD);
	if (len > 0 && t[0] == q[0]) {
		int mtc;
		do {
			for (mtc = 0; mtc < len; mtc++) if (t[mtc] != q[mtc]) break;
			if (mtc > 0) {
				*l_mem_ext += mtc;
				max_len -= mtc;
				len = MIN(12, max_len);
				if (is_FWD) {
					q_off -= mtc; t_off -= mtc;
					for (uint8_t k = 0; k < len; k++) q[k] = q_b[q_off - k];
				} else { t_off += mtc; q += mtc; }
				orc_get_ref(t_b, t, t_off, len, !is_FWD);
			}
		} while (mtc > 0);
	}
	*e_d = orc_lv_extd(t, len, q, len);
	*len_ = len;
}

#define MIN_S_1 12
#define MIN_S_2 20
static int32_t map_seed(const orc_index *ix, mem_rst_t *m_r, seed_info_t *s_i, orc_result *res)
{   /* cly.c:706-939 */
	uint64_t b_p = m_r->sp;
	int32_t q_off = m_r->read_offset;
	uint32_t l_m = m_r->match_len;
	uint8_t *q_b = s_i->bin_read;
	const uint8_t *t_b = ix->ref_bin;
	int64_t uni = -1;
	uint32_t u_off = 0;
	uint64_t t_off = 0;
	uint32_t l_pre, l_suf = 0, d_pre, d_suf = 0;
	int32_t s = 0, max_s = 0;
	lv_frame fr;
	memset(&fr, 0, sizeof fr);
	do {
		uint8_t *q_pre = fr.A, *t_pre = fr.B, *t_suf = fr.C, *q_suf;
		l_pre = MIN(q_off + 1, LV_L);
		for (uint8_t k = 0; k < l_pre; k++) q_pre[k] = q_b[q_off - k];
		int s_l = 0;
		if (m_r->sa_sp != NO_SA)
			uni = get_uni(ix, m_r->sa_sp, m_r->sa_sp_l, &t_off, &u_off);
		else {
			uint8_t c; uint64_t new_sp;
			while (1) {
				if ((b_p & SA_MASK) == 0) break;
				c = 0xff;
				new_sp = orc_occ(ix, b_p, &c) + ix->rank[c];
				if (c == 4) break;
				t_pre[s_l++] = c;
				b_p = new_sp;
				if (s_l >= l_pre) break;
			}
			if ((b_p & SA_MASK) == 0) uni = get_uni(ix, b_p, s_l, &t_off, &u_off);
			else l_pre = s_l;
		}
		if (uni >= 0) {
			if (ix->uni[uni].length < MIN_UNI_L) break;
			l_pre = MIN(l_pre, u_off);
			orc_get_ref(t_b, t_pre, t_off - 1, l_pre, 0);
		}
		d_pre = orc_lv_extd(t_pre, l_pre, q_pre, l_pre);
		s = ix->Q_MEM[l_m] + ix->Q_LV[d_pre][l_pre];
		if (s < MIN_S_1 && l_pre == LV_L && uni < 0) { s = 0; break; }
		if (uni < 0) {
			while (b_p & SA_MASK) {
				uint8_t c = 0xff;
				b_p = orc_occ(ix, b_p, &c) + ix->rank[c];
				s_l++;
			}
			uni = get_uni(ix, b_p, s_l, &t_off, &u_off);
			if (ix->uni[uni].length < MIN_UNI_L) { s = 0; break; }
		}
		int32_t q_off_r = q_off + l_m + 1;
		uint32_t l_max_suf = MIN(ix->uni[uni].length - u_off - l_m, s_i->read_L - q_off_r);
		if (l_max_suf != 0) {
			l_suf = MIN(l_max_suf, LV_L);
			q_suf = q_b + q_off_r;
			orc_get_ref(t_b, t_suf, t_off + l_m, l_suf, 1);
			if (t_suf[0] == q_suf[0]) {
				int mtc;
				do {
					for (mtc = 0; mtc < l_suf; mtc++) if (t_suf[mtc] != q_suf[mtc]) break;
					if (mtc > 0) {
						l_m += mtc;
						s = ix->Q_MEM[l_m] + ix->Q_LV[d_pre][l_pre];
						l_max_suf -= mtc;
						l_suf = MIN(l_max_suf, LV_L);
						q_suf += mtc;
						orc_get_ref(t_b, t_suf, t_off + l_m, l_suf, 1);
					}
				} while (mtc > 0);
			}
			d_suf = orc_lv_extd(t_suf, l_suf, q_suf, l_suf);
			s += ix->Q_LV[d_suf][l_suf];
		} else
			l_suf = d_suf = 0;
		if (s <= MIN_S_2 && l_suf == LV_L) { s = 0; break; }
	} while (0);

	if (s > 0) {
		uint16_t am_mtch_len = l_m; int16_t am_score = s;
		uint8_t am_left_len = l_pre, am_left_ED = d_pre, am_rigt_len = l_suf, am_rigt_ED = d_suf;
		uint32_t r_p_s = ix->uni[uni].ref_list, r_p_e = ix->uni[uni + 1].ref_list;
		int ref_search_l = (l_pre < LV_L || d_pre == 0) ? 1 : 0;
		int ref_search_r = (l_suf < LV_L || d_suf == 0) ? 1 : 0;
		if ((int64_t)r_p_e - (int64_t)r_p_s > 50)
			if (!((int64_t)r_p_e - (int64_t)r_p_s < 1000)) return 50;
		for (uint32_t c_r_p = r_p_s; c_r_p < r_p_e; c_r_p++) {
			uint64_t rp = ix->ref_pos[c_r_p];
			uint64_t rp_global = rp & 0xFFFFFFFFFFull; uint32_t rp_ref = (uint32_t)((rp >> 40) & 0x7FFFFF);
			uint32_t ed_l, ed_r, len_l, len_r;
			uint32_t l_m_ext_l = 0, l_m_ext_r;
			if (ref_search_l || ref_search_r) {
				if (ref_search_l) {
					get_new_ed(ix, &fr, &ed_l, &len_l, &l_m_ext_l, q_off, rp_global + u_off - 1, s_i->read_L, q_b, 1);
					am_left_len = len_l; am_left_ED = ed_l;
				}
				am_mtch_len = l_m + l_m_ext_l;
				if (ref_search_r) {
					l_m_ext_r = 0;
					get_new_ed(ix, &fr, &ed_r, &len_r, &l_m_ext_r, q_off + l_m + 1, rp_global + u_off + l_m, s_i->read_L, q_b, 0);
					am_rigt_len = len_r; am_rigt_ED = ed_r;
					am_mtch_len += l_m_ext_r;
				}
				am_score = ix->Q_MEM[am_mtch_len] + ix->Q_LV[am_left_ED][am_left_len] + ix->Q_LV[am_rigt_ED][am_rigt_len];
				if (am_score < MIN_S_2) continue;
			}
			max_s = MAX(max_s, am_score);
			orc_anchor *a = anchor_push(res);
			a->direction = s_i->direction;
			a->index_in_read = q_off + 1 - l_m_ext_l;
			a->global_offset = rp_global + u_off - l_m_ext_l;
			a->ref_ID = rp_ref;
			a->ref_offset = a->global_offset - ix->ri[a->ref_ID].seq_offset;
			a->mtch_len = am_mtch_len; a->score = am_score;
			a->left_len = am_left_len; a->left_ED = am_left_ED; a->rigt_len = am_rigt_len; a->rigt_ED = am_rigt_ED;
			a->seed_ID = s_i->seed_ID;
			a->duplicate = 0;
		}
	}
	return max_s;
}

/* ------------------------------------------------------------------ seed scheduling (cly.c:1476-1611) */
#define MEM_search_FAST 2
#define MIN_MEM_LEN_FAST 21
static int fast_classify(const orc_index *ix, search_dir_t *s_d, uint32_t read_len, orc_result *res)
{
	uint8_t l_ek = ix->l_ek;
	int min_index = MIN_MEM_LEN_FAST - l_ek;
	uint64_t *kmer = s_d->kmer;
	uint8_t *bin_read = s_d->bin_read;
	uint64_t sp_set_BUFF[500];
	sp_set_t sp_set = {sp_set_BUFF, 0, 500};
	mem_rst_t m_r[MEM_search_FAST];
	memset(m_r, 0, sizeof m_r);
	orc_seed *sv_b = s_d->seed_v, *sv_e = sv_b + s_d->l_seed_v;
	seed_info_t s_i = {bin_read, read_len, 0, (int)s_d->direction};
	for (orc_seed *c_sv = sv_b; c_sv < sv_e; c_sv++) {
		if (c_sv->top == 0) continue;
		sp_set.l = 0;
		s_i.seed_ID = c_sv - sv_b;
		uint32_t a_b_idx = res->n_anc;
		for (int j = c_sv->len - 1; j >= min_index;) {
			int kmer_index = c_sv->offset + j;
			uint64_t prefixValue = kmer[kmer_index] & PRE_IDX_MASK;
			int string_index = kmer_index + l_ek - 1;
			int n = bwt_MEM_search(ix, bin_read + string_index, prefixValue, MEM_search_FAST, MIN_MEM_LEN_FAST - 1, string_index, &sp_set, m_r);
			if (n == 0) { j -= 2; continue; }
			j -= 3;
			int max_score = 0;
			for (mem_rst_t *c_mr = m_r; c_mr < m_r + n; ++c_mr) {
				c_mr->read_offset = string_index - c_mr->match_len;
				int c_score = map_seed(ix, c_mr, &s_i, res);
				max_score = MAX(c_score, max_score);
			}
			if (max_score > 35) j -= 7;
			if (max_score > 256) {
				if (max_score > 512) c_sv++;
				break;
			}
		}
		int top_score = 35;
		for (size_t k = a_b_idx; k < res->n_anc; k++) top_score = MAX(top_score, res->anc[k].score);
		for (size_t k = a_b_idx; k < res->n_anc; k++) res->anc[k].anchor_useless = (res->anc[k].score < top_score) ? 1 : 0;
	}
	return 0;   /* super_repeat bookkeeping is commented out in the reference (cly.c:849-888,1545) */
}

static int mem_rst_cmp_by_match_len(const void *a_, const void *b_)     /* cly.c:1328-1331 */
{
	return ((const mem_rst_t *)b_)->match_len - ((const mem_rst_t *)a_)->match_len;
}

#define MEM_search_SLOW 8
#define MIN_MEM_LEN_SLOW 20
static void slow_classify(const orc_index *ix, search_dir_t *sd, uint32_t read_len, orc_result *res)
{
	int l_ek = ix->l_ek;
	uint8_t *bin_read = sd->bin_read;
	uint64_t *kmer = sd->kmer;
	orc_seed *sv_f = sd->seed_v;
	uint64_t sp_set_BUFF[500];
	sp_set_t sp_set = {sp_set_BUFF, 0, 500};
	static __thread mem_rst_t mem_rst[MEM_search_SLOW * 800 + 1];
	int mem_rst_num;
	seed_info_t seed_info = {bin_read, read_len, 0, (int)sd->direction};
	for (uint32_t i = 0; i < sd->l_seed_v; i++) {
		if ((int)(sv_f[i].len) < 3 && sv_f->top == 0) continue;       /* sv_f->top (seed 0), as written (cly.c:1564) */
		int min_match_len = MIN(MIN_MEM_LEN_SLOW - 1, l_ek + 1);
		sp_set.l = 0;
		mem_rst_num = 0;
		for (int j = sv_f[i].len - 1; j >= 1; j -= 2) {
			int k_idx = sv_f[i].offset + j;
			uint64_t pre_v = kmer[k_idx] & PRE_IDX_MASK;
			int s_idx = k_idx + l_ek - 1;
			int n = bwt_MEM_search(ix, bin_read + s_idx, pre_v, MEM_search_SLOW, min_match_len, s_idx, &sp_set, mem_rst + mem_rst_num);
			for (int k = mem_rst_num; k < mem_rst_num + n; k++)
				mem_rst[k].read_offset = k_idx + l_ek - 1 - mem_rst[k].match_len;
			mem_rst_num += n;
		}
		if (mem_rst_num == 0) continue;
		if (mem_rst_num > 1) orc_msort(mem_rst, mem_rst_num, sizeof(mem_rst_t), mem_rst_cmp_by_match_len);
		seed_info.seed_ID = i;
		uint32_t a_b_idx = res->n_anc;
		int max_search = MIN(mem_rst_num, MEM_search_SLOW);
		for (mem_rst_t *c = mem_rst; c < mem_rst + max_search; ++c) map_seed(ix, c, &seed_info, res);
		int top_score = 35;
		for (size_t k = a_b_idx; k < res->n_anc; k++) top_score = MAX(top_score, res->anc[k].score);
		for (size_t k = a_b_idx; k < res->n_anc; k++) res->anc[k].anchor_useless = (res->anc[k].score < top_score) ? 1 : 0;
	}
	res->fast_classify = 0;
}

/* ------------------------------------------------------------------ chaining (cly.c:38-52, 72-112, 201-349) */
static int chain_cmp_by_score(const void *a_, const void *b_)
{
	const orc_chain *a = (const orc_chain *)a_, *b = (const orc_chain *)b_;
	if (a->with_top_anchor != b->with_top_anchor) return (a->with_top_anchor) ? (-1) : (1);
	int score_a = a->sum_score + ((a->q_ed - a->q_st) << 1);
	score_a -= (a->indel << 2);
	int score_b = b->sum_score + ((b->q_ed - b->q_st) << 1);
	score_b -= (b->indel << 2);
	if (score_a < score_b) return 1;
	if (score_a > score_b) return -1;
	return 0;
}

static void chain_insert_meta(orc_result *res, int32_t ai, orc_chain *c, int new_chain, int dis_minus)
{
	orc_anchor *anchor = res->anc + ai;
	uint32_t ref_l = anchor->ref_offset, ref_r = ref_l + anchor->mtch_len;
	uint32_t read_l = anchor->index_in_read, read_r = read_l + anchor->mtch_len;
	if (new_chain) {
		anchor->chain_id = c->chain_id;
		anchor->pre = -1;
		c->ref_ID = anchor->ref_ID;
		c->direction = anchor->direction;
		c->q_t_dis = anchor->ref_offset - anchor->index_in_read;
		c->t_st = ref_l; c->t_ed = ref_r; c->q_st = read_l; c->q_ed = read_r;
		c->with_top_anchor = !anchor->anchor_useless;
		c->anchor_number = 1;
		c->sum_score = (anchor->duplicate) ? 1 : anchor->score;
		c->indel = 0;
		c->cur = ai;
	} else {
		anchor->chain_id = c->chain_id;
		c->with_top_anchor |= (!anchor->anchor_useless);
		if (c->q_ed >= read_r) return;
		c->t_ed = MAX(ref_r, c->t_ed);
		c->q_ed = read_r;
		anchor->pre = c->cur;
		c->cur = ai;
		c->q_t_dis = anchor->ref_offset - anchor->index_in_read;
		c->indel += dis_minus;
		c->anchor_number++;
		c->sum_score += (anchor->duplicate) ? 1 : anchor->score;
	}
}

#define MAX_dis_MINUS 30
#define MAX_waiting_len 400
static void chain_insert_M2(orc_result *res, int32_t ai)
{
	orc_anchor *anchor = res->anc + ai;
	uint8_t direction = anchor->direction;
	uint32_t ref_ID = anchor->ref_ID;
	int32_t dis = anchor->ref_offset - anchor->index_in_read;
	int dis_minus = 0;
	for (size_t k = 0; k < res->n_hit; k++) {
		orc_chain *c_s = res->hit + k;
		if (c_s->direction == direction && c_s->ref_ID == ref_ID &&
		    (dis_minus = ABS(dis - c_s->q_t_dis)) < MAX_dis_MINUS &&
		    ABS_U(c_s->t_ed, anchor->ref_offset) < MAX_waiting_len) {
			chain_insert_meta(res, ai, c_s, 0, dis_minus);
			return;
		}
	}
	orc_chain *new_c = chain_push(res);
	new_c->chain_id = res->n_hit - 1;
	chain_insert_meta(res, ai, new_c, 1, dis_minus);
}

static int anchor_cmp_by_chr_ID_and_pos(const void *a_, const void *b_)     /* cly.c:226-235: 0/1 comparator */
{
	const orc_anchor *a = (const orc_anchor *)a_, *b = (const orc_anchor *)b_;
	if (a->ref_ID != b->ref_ID) return a->ref_ID > b->ref_ID;
	if (a->direction != b->direction) return a->direction > b->direction;
	return a->ref_offset > b->ref_offset;
}

#define MAX_ANCHOR_OVERLAP 3
static void chain_insert_M3(orc_result *res)
{
	int score_v[1024];
	orc_anchor *A = res->anc;
	int32_t n = (int32_t)res->n_anc;
	orc_msort(A, res->n_anc, sizeof(orc_anchor), anchor_cmp_by_chr_ID_and_pos);
	for (int32_t chr_st = 0; chr_st < n;) {
		int32_t chr_ed = chr_st + 1, c_a;
		uint32_t ref_ID = A[chr_st].ref_ID;
		uint32_t direction = A[chr_st].direction;
		for (; chr_ed < n && A[chr_ed].ref_ID == ref_ID && A[chr_ed].direction == direction &&
		       A[chr_ed].ref_offset - A[chr_ed - 1].ref_offset < 2000; chr_ed++);
		if (chr_ed - chr_st > 1024) chr_ed = chr_st + 1024;
		int32_t max_anchor = -1; int max_score = 0, anchor_max_score;
		for (c_a = chr_st; c_a < chr_ed; c_a++) {
			A[c_a].pre = -1;
			anchor_max_score = A[c_a].score;
			uint32_t max_t = A[c_a].ref_offset + MAX_ANCHOR_OVERLAP;
			uint32_t max_q = A[c_a].index_in_read + MAX_ANCHOR_OVERLAP;
			for (int32_t pre = c_a - 1; pre >= chr_st; pre--) {
				if (A[pre].index_in_read + A[pre].mtch_len > max_q) continue;
				if (A[pre].ref_offset + A[pre].mtch_len > max_t) continue;
				if (A[pre].index_in_read + 1000 < max_q) break;
				if (A[pre].ref_offset + 1000 < max_t) break;
				int indel = A[pre].index_in_read - A[pre].ref_offset - (max_q - max_t);
				int ABS_indel = ABS(indel);
				if (ABS_indel > 200) continue;
				int new_score = score_v[pre - chr_st] + A[c_a].mtch_len - (ABS_indel >> 4) - ((max_q - A[pre].index_in_read) >> 8);
				if (new_score > anchor_max_score) { anchor_max_score = new_score; A[c_a].pre = pre; }
			}
			score_v[c_a - chr_st] = anchor_max_score;
			if (max_score < anchor_max_score) { max_score = anchor_max_score; max_anchor = c_a; }
		}
		int sum_INDEL = 0, anchor_number = 1; int32_t pre = max_anchor;
		int sum_score = (A[max_anchor].duplicate) ? 1 : A[max_anchor].score;
		int with_top = !A[max_anchor].anchor_useless;
		for (; A[pre].pre != -1; anchor_number++) {
			int32_t pre_ = A[pre].pre;
			sum_INDEL += (A[pre].index_in_read - A[pre_].index_in_read) - (A[pre].ref_offset - A[pre_].ref_offset);
			with_top |= (!A[pre].anchor_useless);
			sum_score += (A[pre].duplicate) ? 1 : A[pre].score;
			pre = pre_;
		}
		orc_chain *new_c = chain_push(res);
		A = res->anc;
		new_c->chain_id = res->n_hit - 1;
		new_c->ref_ID = ref_ID;
		new_c->direction = direction;
		new_c->q_t_dis = A[max_anchor].ref_offset - A[max_anchor].index_in_read;
		new_c->t_st = A[pre].ref_offset;
		new_c->t_ed = A[max_anchor].ref_offset + A[max_anchor].mtch_len;
		new_c->q_st = A[pre].index_in_read;
		new_c->q_ed = A[max_anchor].index_in_read + A[max_anchor].mtch_len;
		new_c->with_top_anchor = with_top;
		new_c->anchor_number = anchor_number;
		new_c->sum_score = sum_score;
		new_c->indel = sum_INDEL;
		new_c->cur = max_anchor;
		chr_st = chr_ed;
	}
}

static void resolve_tree(orc_result *res)       /* cly.c:326-349 */
{
	res->n_hit = 0;
	if (res->n_anc < 50)
		for (int32_t a = 0; a < (int32_t)res->n_anc; a++) chain_insert_M2(res, a);
	else
		chain_insert_M3(res);
	if (res->n_hit > 1) orc_msort(res->hit, res->n_hit, sizeof(orc_chain), chain_cmp_by_score);
	size_t rst_num = MIN(5, res->n_hit);
	while (rst_num < res->n_hit && res->hit[rst_num].with_top_anchor == 1) rst_num++;
	res->n_hit = rst_num;
}

/* ------------------------------------------------------------------ 9-mer sparse DP scoring (cly.c:1691-1818, 2173-2224, 2335-2849) */
static void sc_hash_idx(sc_hash_t *sc_hash, orc_result *res)     /* cly.c:1691-1710 */
{
	memset(sc_hash, 0, 256 * sizeof(sc_hash_t));
	int sc_con_index = 256;
	for (size_t h = 0; h < res->n_hit; h++) {
		orc_chain *c_h = res->hit + h;
		for (int i = 1; i >= 0; i--) {
			uint16_t c_key = ((i == 1) ? (c_h->t_st - c_h->q_st) : (c_h->t_ed - c_h->q_ed)) & 0xff;
			while (sc_hash[c_key].next != 0) c_key = sc_hash[c_key].next;
			sc_hash[c_key].seed_ID = (uint16_t)((h + 1) & 0x7fff);
			sc_hash[c_key].s_or_e = i;
			sc_hash[c_key].next = sc_con_index;
			sc_hash[sc_con_index++].next = 0;
		}
	}
}

static int combine_chain(orc_chain *c_st, int chain_ID, sc_hash_t *sc_hash, int dis, int isleft, int c_q_pos, orc_chain **combined)
{   /* cly.c:1763-1808 */
	uint16_t key = (dis) & 0xff;
	orc_chain *c, *c_h = c_st + chain_ID;
	while (sc_hash[key].next != 0) {
		uint16_t seed_ID = sc_hash[key].seed_ID;
		c = c_st + seed_ID - 1;
		int dis_con = (isleft) ? (c->t_ed - c->q_ed) : (c->t_st - c->q_st);
		int q_pos_con = (!isleft) ? (c->q_st) : (c->q_ed - S_A_KEMR_L);
		if (dis == dis_con && c_h != c && isleft != sc_hash[key].s_or_e && ABS_U(c_q_pos, q_pos_con) < 8 &&
		    c_h->ref_ID == c->ref_ID && c_h->direction == c->direction && c->sum_score != 0 && seed_ID - 1 > chain_ID) {
			c_h->sum_score += c->sum_score;
			c_h->anchor_number += c->anchor_number;
			c_h->indel += c->indel;
			c_h->q_st = MIN(c_h->q_st, c->q_st);
			c_h->t_st = MIN(c_h->t_st, c->t_st);
			c_h->q_ed = MAX(c_h->q_ed, c->q_ed);
			c_h->t_ed = MAX(c_h->t_ed, c->t_ed);
			c->sum_score = 0;
			c->t_st = c->t_ed = c->q_st = c->q_ed = 0;
			*combined = c;
			return 1;
		}
		key = sc_hash[key].next;
	}
	return 0;
}

static inline int MEM_search(const uint8_t *q, const uint8_t *t, int forward, int max)     /* cly.c:1810-1818 */
{
	int len = 0;
	if (forward) for (; len < max && *q++ == *t++; len++);
	else for (; len < max && *q-- == *t--; len++);
	return len;
}

static const uint32_t hash_size[20] = {
	0x00001, 0x00002, 0x00004, 0x00008, 0x00010, 0x00020, 0x00040, 0x00080, 0x00100, 0x00200,
	0x00400, 0x00800, 0x01000, 0x02000, 0x04000, 0x08000, 0x10000, 0x20000, 0x40000, 0x80000};

static int build_hash_table_M2(search_dir_t *search_dir, orc_result *res, int q_len, orc_buff *buff)
{   /* cly.c:2173-2224 */
	int both_dir = 0;
	for (size_t i = 0; i < res->n_hit; i++) {
		both_dir |= (res->hit[i].direction == FORWARD) ? 0x2 : 0x1;
		if (both_dir == 3) break;
	}
	int key_len = 10;
	for (; key_len < 18; key_len++) if (hash_size[key_len] >= (uint32_t)q_len) break;
	uint64_t MASK = (1ull << (2 * S_A_KEMR_L)) - 1;
	uint64_t KEY_MASK = (1ull << key_len) - 1;
	for (int c_dir = 2; c_dir >= 1; c_dir--) {
		if ((c_dir & both_dir) == 0) continue;
		uint32_t direction = (c_dir == 1) ? REVERSE : FORWARD;
		search_dir_t *c_sd = ((search_dir->direction == direction) ? 0 : 1) + search_dir;
		int slot = (c_dir == 2) ? 0 : 1;
		size_t need = (size_t)hash_size[key_len] + q_len + 8;
		if (need > buff->m_sa_hash[slot]) { buff->m_sa_hash[slot] = need; buff->sa_hash[slot] = (sa_hash_t *)realloc(buff->sa_hash[slot], need * sizeof(sa_hash_t)); }
		sa_hash_t *h = buff->sa_hash[slot];
		int kmer_con_index = hash_size[key_len];
		for (int index = 0; index < kmer_con_index; index++) h[index].next = 0;
		const uint8_t *q = c_sd->bin_read;
		uint64_t kmer = 0;
		for (int k = 0; k < S_A_KEMR_L - 1; k++) kmer = (kmer << 2) | q[k];
		for (uint32_t c_pos = 0; c_pos < (uint32_t)(q_len - S_A_KEMR_L + 1); c_pos++) {
			kmer = ((kmer << 2) | q[c_pos + S_A_KEMR_L - 1]) & MASK;
			uint32_t next = kmer & KEY_MASK;
			while (h[next].next != 0) next = h[next].next;
			uint32_t id = kmer_con_index++;
			h[id].kmer = kmer; h[id].next = 0; h[id].pos = c_pos;
			h[next].next = id;
		}
	}
	return key_len;
}

static void sdp_match(orc_buff *buff, uint32_t q_bg, uint32_t q_ed, const uint8_t *q_str, const uint8_t *t_str, uint32_t t_len, int key_len,
                      const sa_hash_t *sa_hash, uint32_t t_st, int isForward)
{   /* cly.c:2335-2440 */
	uint64_t KEY_MASK = (1ull << key_len) - 1;
	uint32_t t_kmer_num = t_len - S_A_KEMR_L + 1;
	uint64_t MASK = (1ull << (2 * S_A_KEMR_L)) - 1;
	if (isForward) {
		const uint8_t *c_t_str = t_str + 4;
		uint64_t kmer = 0;
		for (int k = 0; k < S_A_KEMR_L - 1; k++) kmer = (kmer << 2) | c_t_str[k];
		for (int i = 4; i < t_kmer_num; i++, c_t_str++) {
			kmer = ((kmer << 2) | c_t_str[S_A_KEMR_L - 1]) & MASK;
			if ((i & 0x03) != 0) continue;
			uint32_t next = sa_hash[kmer & KEY_MASK].next;
			while (next != 0) {
				if (sa_hash[next].kmer == kmer) {
					uint32_t q_pos = sa_hash[next].pos;
					if (q_pos >= q_bg && q_pos <= q_ed) {
						int back_len = MEM_search(q_str + q_pos - 1, c_t_str - 1, 0, 4);
						if (back_len < 4 || i == 4) {
							uint32_t max_search = q_ed - q_pos - 1;
							max_search = MIN(max_search, t_len - i - 1) + OVER_SEARCH_M2;
							int forward_len = MEM_search(q_str + q_pos + S_A_KEMR_L, c_t_str + S_A_KEMR_L, 1, max_search);
							int total_len = back_len + forward_len + 1;
							if (total_len >= 4) {
								spd_match *p = sms_push(buff);
								p->len = total_len;
								p->q_pos = q_pos - back_len;
								p->t_pos = i - back_len + t_st;
							}
						}
					}
				}
				next = sa_hash[next].next;
			}
		}
	} else {
		const uint8_t *c_t_str = t_str + t_len - S_A_KEMR_L - 4;
		uint64_t kmer = 0;
		for (int k = 0; k < S_A_KEMR_L; k++) kmer = (kmer << 2) | c_t_str[k];
		kmer <<= 2;                                                       /* bit2_preKmer_init */
		for (int i = 4; i < t_kmer_num; i++, c_t_str--) {
			kmer = (kmer >> 2) | ((uint64_t)c_t_str[0] << ((S_A_KEMR_L << 1) - 2));
			if ((i & 0x03) != 0) continue;
			uint32_t next = sa_hash[kmer & KEY_MASK].next;
			while (next != 0) {
				if (sa_hash[next].kmer == kmer) {
					uint32_t q_pos = sa_hash[next].pos;
					if (q_pos >= q_bg && q_pos <= q_ed) {
						int forward_len = MEM_search(q_str + q_pos + S_A_KEMR_L, c_t_str + S_A_KEMR_L, 1, 4);
						if (forward_len < 4 || i == 4) {
							uint32_t max_search = q_pos;
							max_search = MIN(max_search, c_t_str - t_str) + OVER_SEARCH_M2;
							int back_len = MEM_search(q_str + q_pos - 1, c_t_str - 1, 0, max_search);
							int total_len = back_len + forward_len + 1;
							if (total_len >= 4) {
								spd_match *p = sms_push(buff);
								p->len = total_len;
								p->q_pos = q_pos - back_len;
								p->t_pos = c_t_str - t_str - back_len + t_st;
							}
						}
					}
				}
				next = sa_hash[next].next;
			}
		}
	}
}

#define MAX_sms_overlap (6)
#define MAX_sms_overlap_middle (6)
static int sdp_middle_M2(const orc_index *ix, orc_result *res, int32_t c_a, orc_buff *buff, const uint8_t *q_str, const sa_hash_t *sa_hash, int key_len)
{   /* cly.c:2444-2530 */
	int score = 10000;
	const orc_anchor *A = res->anc;
	uint64_t t_offset = ix->ri[A[c_a].ref_ID].seq_offset;
	int32_t pre_a = -1;
	while (c_a != -1) {
		pre_a = A[c_a].pre;
		if (pre_a != -1) {
			int pre_mch = A[pre_a].mtch_len;
			int pre_refoffset = A[pre_a].ref_offset - 3;
			int total_ref_len = A[c_a].ref_offset - (pre_refoffset + pre_mch) + 3;
			buff->n_sms = 0;
			spd_match *p = sms_push(buff);
			p->score = score;
			p->q_pos = A[pre_a].index_in_read;
			p->t_pos = A[pre_a].ref_offset;
			p->len = A[pre_a].mtch_len - S_A_KEMR_L + 1;
			if (total_ref_len > 12) {
				uint8_t ref[2000 + 128];
				memset(ref, 0, sizeof ref);                                      /* policy P1 */
				if (!(total_ref_len < 2000)) { fprintf(stderr, "[oracle] xassert total_ref_len<2000\n"); abort(); }
				uint64_t ref_offset = pre_refoffset + t_offset + pre_mch;
				orc_get_ref(ix->ref_bin, ref, ref_offset, total_ref_len, 1);
				sdp_match(buff, A[pre_a].index_in_read + pre_mch - 8, A[c_a].index_in_read - 1, q_str, ref, total_ref_len, key_len, sa_hash,
				          pre_refoffset + pre_mch, 1);
			}
			p = sms_push(buff);
			p->q_pos = A[c_a].index_in_read;
			p->t_pos = A[c_a].ref_offset;
			p->len = A[c_a].mtch_len - S_A_KEMR_L + 1;
			if (buff->n_sms > 1) {
				spd_match *base = buff->sms, *spd_ed = base + buff->n_sms;
				for (spd_match *c_spd = base + 1; c_spd < spd_ed; c_spd++) {
					int max_score = c_spd->len;
					uint32_t max_q = c_spd->q_pos + MAX_sms_overlap_middle;
					uint32_t max_t = c_spd->t_pos + MAX_sms_overlap_middle;
					for (spd_match *c_pre = c_spd - 1; c_pre >= base; c_pre--) {
						int pre_q_ed = c_pre->q_pos + c_pre->len + S_A_KEMR_L - 1;
						int pre_t_ed = c_pre->t_pos + c_pre->len + S_A_KEMR_L - 1;
						if (pre_q_ed > max_q) continue;
						if (pre_t_ed > max_t) continue;
						int indel = c_pre->q_pos - c_pre->t_pos - (max_q - max_t);
						int ABS_indel = ABS(indel);
						if (ABS_indel > 200) continue;
						int new_score = c_pre->score + c_spd->len - (ABS_indel >> 3);
						if (pre_q_ed > c_spd->q_pos || pre_t_ed > c_spd->t_pos) {
							int overlap_q = pre_q_ed - c_spd->q_pos;
							int overlap_t = pre_t_ed - c_spd->t_pos;
							new_score -= MAX(overlap_q, overlap_t);
						}
						max_score = MAX(max_score, new_score);
					}
					score = MAX(max_score, score);
					c_spd->score = max_score;
				}
			}
		} else
			score += A[c_a].mtch_len - S_A_KEMR_L + 1;
		c_a = pre_a;
	}
	return score - 10000;
}

static int sdp_right_M2(const orc_index *ix, orc_result *res, orc_buff *buff, const uint8_t *q_str, const sa_hash_t *sa_hash, int key_len,
                        int chain_ID, uint32_t l_read, sc_hash_t *sc_hash, int score_ori)
{   /* cly.c:2532-2677 */
	orc_chain *c_st = res->hit;
	score_ori += 10000;
	int total_max_score = score_ori;
	int max_sms_id = 0;
	orc_chain *c_h = c_st + chain_ID, *combined;
	buff->n_sms = 0;
	uint8_t ref[1000 + 64];
	memset(ref, 0, sizeof ref);                                                  /* policy P1 */
	spd_match *p = sms_push(buff);
	p->score = score_ori; p->q_pos = c_h->q_ed; p->t_pos = c_h->t_ed; p->len = 1 - S_A_KEMR_L;
	uint32_t current_sms = 1;
	uint64_t t_offset_global = ix->ri[c_h->ref_ID].seq_offset;
	uint64_t t_length = ix->ri[c_h->ref_ID].seq_l;
	uint32_t c_t_offset = c_h->t_ed - 3;
	int last_search = 0;
	while (1) {
		if (buff->n_sms == current_sms) {
			uint32_t next_step = t_length - c_t_offset;
			if (next_step < MIN_SCORE_MEM) break;
			uint32_t max_search_ref;
			if (l_read - c_h->q_ed < 600) {
				if (last_search == 1) break;
				last_search = 1;
				max_search_ref = l_read - c_h->q_ed + 60;
			} else
				max_search_ref = t_length - c_t_offset;
			max_search_ref = MIN(600, max_search_ref);
			orc_get_ref(ix->ref_bin, ref, c_t_offset + t_offset_global, max_search_ref + OVER_SEARCH_M2, 1);
			int search_q_ed = (int)buff->sms[max_sms_id].q_pos + 1000;
			search_q_ed = MIN(search_q_ed, l_read);
			int search_q_st = MAX(search_q_ed - 2000, c_h->q_st - 8);
			sdp_match(buff, search_q_st, search_q_ed, q_str, ref, max_search_ref, key_len, sa_hash, c_t_offset, 1);
			c_t_offset += max_search_ref - S_A_KEMR_L - 3;
			if (buff->n_sms == current_sms) break;
			if (buff->sms[current_sms].t_pos > buff->sms[max_sms_id].t_pos + 1000) break;
		}
		spd_match *c_sms = buff->sms + current_sms++;
		int max_score = c_sms->len;
		uint32_t max_pre_q = c_sms->q_pos + MAX_sms_overlap;
		uint32_t max_pre_t = c_sms->t_pos + MAX_sms_overlap;
		spd_match *c_sms_ed = buff->sms, *c_pre = buff->sms + current_sms - 2;
		for (; c_pre >= c_sms_ed; c_pre--) {
			int pre_q_ed = c_pre->q_pos + c_pre->len + S_A_KEMR_L - 1;
			int pre_t_ed = c_pre->t_pos + c_pre->len + S_A_KEMR_L - 1;
			if (pre_q_ed > max_pre_q) continue;
			if (pre_t_ed > max_pre_t) continue;
			if (c_pre->t_pos + 600 < max_pre_t) break;
			int indel = c_pre->q_pos - c_pre->t_pos - (max_pre_q - max_pre_t);
			int ABS_indel = ABS(indel);
			if (ABS_indel > 200) continue;
			int new_score = c_pre->score + c_sms->len - (ABS_indel >> 3);
			if (pre_q_ed > c_sms->q_pos || pre_t_ed > c_sms->t_pos) {
				int overlap_q = pre_q_ed - c_sms->q_pos;
				int overlap_t = pre_t_ed - c_sms->t_pos;
				new_score -= MAX(overlap_q, overlap_t);
			}
			max_score = MAX(max_score, new_score);
		}
		c_sms->score = max_score;
		if (c_sms->len >= 8 &&
		    combine_chain(c_st, chain_ID, sc_hash, c_sms->t_pos - c_sms->q_pos, 0, c_sms->q_pos, &combined) == 1) {
			uint32_t c_len = c_sms->len;     /* read before sdp_middle_M2 reuses the sms buffer */
			total_max_score = MAX(score_ori, max_score) - c_len + sdp_middle_M2(ix, res, combined->cur, buff, q_str, sa_hash, key_len);
			score_ori = total_max_score;
			max_sms_id = 0;
			buff->n_sms = 0;
			p = sms_push(buff);
			p->score = total_max_score; p->q_pos = c_h->q_ed; p->t_pos = c_h->t_ed; p->len = -S_A_KEMR_L;
			current_sms = 1;
			c_t_offset = c_h->t_ed;
			continue;
		}
		if (total_max_score < max_score) { total_max_score = max_score; max_sms_id = current_sms - 1; }
		if (c_sms->t_pos > buff->sms[max_sms_id].t_pos + 1000) break;
	}
	c_h->q_ed = buff->sms[max_sms_id].q_pos + buff->sms[max_sms_id].len + S_A_KEMR_L;
	c_h->t_ed = buff->sms[max_sms_id].t_pos + buff->sms[max_sms_id].len + S_A_KEMR_L;
	return total_max_score - 10000;
}

static int sdp_left_M2(const orc_index *ix, orc_result *res, orc_buff *buff, const uint8_t *q_str, const sa_hash_t *sa_hash, int key_len,
                       int chain_ID, uint32_t l_read, sc_hash_t *sc_hash, int score_ori)
{   /* cly.c:2679-2819 */
	(void)l_read;
	orc_chain *c_st = res->hit;
	score_ori += 10000;
	int total_max_score = score_ori;
	int max_sms_id = 0;
	orc_chain *c_h = c_st + chain_ID, *combined;
	buff->n_sms = 0;
	uint8_t ref[1000 + 64];
	memset(ref, 0, sizeof ref);                                                  /* policy P1 */
	spd_match *p = sms_push(buff);
	p->score = score_ori; p->q_pos = c_h->q_st; p->t_pos = c_h->t_st;          /* len keeps its previous content, unused */
	uint32_t current_sms = 1;
	uint64_t t_offset_global = ix->ri[c_h->ref_ID].seq_offset;
	uint32_t c_t_offset = c_h->t_st + 3;
	int last_search = 0;
	while (1) {
		if (buff->n_sms == current_sms) {
			uint32_t next_step = c_t_offset;
			if (next_step < MIN_SCORE_MEM) break;
			uint32_t max_search_ref;
			if (c_h->q_st < 600) {
				if (last_search == 1) break;
				last_search = 1;
				max_search_ref = c_h->q_st + 60;
			} else
				max_search_ref = c_t_offset;
			max_search_ref = MIN(600, max_search_ref);
			if (t_offset_global == 0 && c_t_offset < OVER_SEARCH_M2 + max_search_ref)
				orc_get_ref(ix->ref_bin, ref, c_t_offset + t_offset_global - max_search_ref, max_search_ref, 1);
			else
				orc_get_ref(ix->ref_bin, ref, c_t_offset + t_offset_global - max_search_ref - OVER_SEARCH_M2, max_search_ref + OVER_SEARCH_M2, 1);
			int search_q_st = (int)buff->sms[max_sms_id].q_pos - 1000;
			search_q_st = MAX(search_q_st, 0);
			int search_q_ed = MIN(search_q_st + 2000, c_h->q_st - 1);
			sdp_match(buff, search_q_st, search_q_ed, q_str, ref + OVER_SEARCH_M2, max_search_ref, key_len, sa_hash, c_t_offset - max_search_ref, 0);
			c_t_offset = c_t_offset - max_search_ref + S_A_KEMR_L + 3;
			if (buff->n_sms == current_sms) break;
			if (buff->sms[current_sms].t_pos + 1000 < buff->sms[max_sms_id].t_pos) break;
		}
		spd_match *c_sms = buff->sms + current_sms++;
		int max_score = c_sms->len;
		uint32_t min_pre_q = c_sms->q_pos + c_sms->len - MAX_sms_overlap + S_A_KEMR_L - 1;
		uint32_t min_pre_t = c_sms->t_pos + c_sms->len - MAX_sms_overlap + S_A_KEMR_L - 1;
		spd_match *c_sms_ed = buff->sms, *c_pre = buff->sms + current_sms - 2;
		for (; c_pre >= c_sms_ed; c_pre--) {
			if (c_pre->q_pos < min_pre_q) continue;
			if (c_pre->t_pos < min_pre_t) continue;
			if (min_pre_t + 600 < c_pre->t_pos) break;
			int indel = c_pre->q_pos - c_pre->t_pos - (min_pre_q - min_pre_t);
			int ABS_indel = ABS(indel);
			if (ABS_indel > 200) continue;
			int new_score = c_pre->score + c_sms->len - (ABS_indel >> 3);
			if (min_pre_q + MAX_sms_overlap > c_pre->q_pos || min_pre_t + MAX_sms_overlap > c_pre->t_pos) {
				int overlap_q = min_pre_q + MAX_sms_overlap - c_pre->q_pos;
				int overlap_t = min_pre_t + MAX_sms_overlap - c_pre->t_pos;
				new_score -= MAX(overlap_q, overlap_t);
			}
			max_score = MAX(max_score, new_score);
		}
		c_sms->score = max_score;
		if (c_sms->len >= 8 &&
		    combine_chain(c_st, chain_ID, sc_hash, c_sms->t_pos - c_sms->q_pos, 1, c_sms->q_pos + c_sms->len, &combined) == 1) {
			uint32_t c_len = c_sms->len;
			total_max_score = MAX(score_ori, max_score) - c_len + sdp_middle_M2(ix, res, combined->cur, buff, q_str, sa_hash, key_len);
			score_ori = total_max_score;
			max_sms_id = 0;
			buff->n_sms = 0;
			p = sms_push(buff);
			p->score = total_max_score; p->q_pos = c_h->q_st; p->t_pos = c_h->t_st;
			current_sms = 1;
			c_t_offset = c_h->t_st;
			continue;
		}
		if (total_max_score < max_score) { total_max_score = max_score; max_sms_id = current_sms - 1; }
		if (c_sms->t_pos + 1000 < buff->sms[max_sms_id].t_pos) break;
	}
	c_h->q_st = buff->sms[max_sms_id].q_pos;
	c_h->t_st = buff->sms[max_sms_id].t_pos;
	return total_max_score - 10000;
}

static void get_score_M2(const orc_index *ix, search_dir_t *search_dir, orc_buff *buff, uint32_t l_read, orc_result *res, sc_hash_t *sc_hash)
{   /* cly.c:2821-2849 */
	int key_len = build_hash_table_M2(search_dir, res, l_read, buff);
	for (size_t i = 0; i < res->n_hit; i++) {
		orc_chain *st_hit = res->hit;
		if (st_hit[i].sum_score == 0) continue;
		search_dir_t *c_sd = ((search_dir->direction == st_hit[i].direction) ? 0 : 1) + search_dir;
		const sa_hash_t *sa_hash = (st_hit[i].direction == FORWARD) ? buff->sa_hash[0] : buff->sa_hash[1];
		int score = sdp_middle_M2(ix, res, st_hit[i].cur, buff, c_sd->bin_read, sa_hash, key_len);
		score = sdp_right_M2(ix, res, buff, c_sd->bin_read, sa_hash, key_len, (int)i, l_read, sc_hash, score);
		score = sdp_left_M2(ix, res, buff, c_sd->bin_read, sa_hash, key_len, (int)i, l_read, sc_hash, score);
		st_hit[i].sum_score = score;
	}
}

/* ------------------------------------------------------------------ final filter / sort / primary (cly.c:54-64, 2853-3058) */
static int chain_cmp_by_pos(const void *a_, const void *b_)
{
	const orc_chain *a = (const orc_chain *)a_, *b = (const orc_chain *)b_;
	if (a->ref_ID > b->ref_ID) return 1;
	if (a->ref_ID < b->ref_ID) return -1;
	if (a->t_st > b->t_st) return 1;
	if (a->t_st < b->t_st) return -1;
	if (a->sum_score < b->sum_score) return 1;
	if (a->sum_score > b->sum_score) return -1;
	return 0;
}

static int chain_cmp_by_MEM_score(const void *a_, const void *b_)      /* asymmetric on ties, as written */
{
	const orc_chain *a = (const orc_chain *)a_, *b = (const orc_chain *)b_;
	int score_a = (a->sum_score << 5);
	int score_b = (b->sum_score << 5);
	if (score_a < score_b) return 1;
	if (score_a > score_b) return -1;
	return (a->sum_score % 2);
}

#define FILTER_MIN_SCORE_SHROT_3G_READ 30
#define FILTER_MIN_SCORE_2G_READ 26
static void delete_small_score_rst(const orc_index *ix, orc_result *res, search_dir_t *search_dir, orc_buff *buff, uint32_t l_read)
{
	res->entered_final = 0;
	if (res->n_hit == 0) return;
	res->entered_final = 1;
	if (res->n_hit > 200) {
		size_t rst_num = 200;
		for (; rst_num < res->n_hit && res->hit[rst_num].sum_score > 50; rst_num++);
		res->n_hit = rst_num;
	}
	res->n_hit = MIN(400, res->n_hit);
	uint32_t n_sc_hash = 256 + (res->n_hit << 1);
	if (n_sc_hash + 2 > buff->m_sc_hash) { buff->m_sc_hash = n_sc_hash + 22; buff->sc_hash = (sc_hash_t *)realloc(buff->sc_hash, buff->m_sc_hash * sizeof(sc_hash_t)); }
	sc_hash_t *sc_hash = buff->sc_hash;
	sc_hash_idx(sc_hash, res);
	get_score_M2(ix, search_dir, buff, l_read, res, sc_hash);

	orc_chain *st_c = res->hit, *ed_c = st_c + res->n_hit, *c_c;
	if (res->n_hit > 1) orc_msort(res->hit, res->n_hit, sizeof(orc_chain), chain_cmp_by_pos);
	for (c_c = st_c; c_c < ed_c - 1; c_c++) {
		if (c_c->sum_score == 0) continue;
		for (orc_chain *next_c = c_c + 1; next_c < ed_c; next_c++) {
			if (c_c->ref_ID == next_c->ref_ID) {
				if (c_c->direction != next_c->direction) continue;
				if (next_c->sum_score == 0) continue;
				if (next_c->t_st < c_c->t_st + 5 && next_c->q_st < c_c->q_st + 5 && next_c->sum_score < c_c->sum_score + 5) {
					next_c->sum_score = 0;
					next_c->q_ed = next_c->q_st;
					next_c->t_ed = next_c->t_st;
					continue;
				}
				int dis_t = next_c->t_st - c_c->t_ed;
				int dis_q = next_c->q_st - c_c->q_ed;
				int dis_t_q = ABS(dis_t - dis_q);
				if ((dis_t > -20 && dis_t < 1000 && dis_q > -20 && dis_q < 1000) && dis_t_q < 200) {
					c_c->t_ed = MAX(c_c->t_ed, next_c->t_ed);
					c_c->q_ed = MAX(c_c->q_ed, next_c->q_ed);
					c_c->sum_score += next_c->sum_score;
					next_c->sum_score = 0;
					next_c->q_ed = next_c->q_st;
					next_c->t_ed = next_c->t_st;
				}
			} else
				break;
		}
	}
	buff->max_read_l = MAX(buff->max_read_l, l_read);
	if (buff->max_read_l < 510) {
		for (c_c = st_c; c_c < ed_c; c_c++) {
			int score = c_c->sum_score + ((c_c->q_ed - c_c->q_st) >> 5);
			if (score < FILTER_MIN_SCORE_2G_READ) c_c->sum_score = 0;
		}
	} else if (l_read < 310) {
		for (c_c = st_c; c_c < ed_c; c_c++) {
			int score = c_c->sum_score + ((c_c->q_ed - c_c->q_st) >> 5);
			if (score < FILTER_MIN_SCORE_SHROT_3G_READ) c_c->sum_score = 0;
		}
	} else {
		for (c_c = st_c; c_c < ed_c; c_c++) {
			int score = c_c->sum_score + ((c_c->q_ed - c_c->q_st) >> 5);
			if (score < (ix->filter_min_score_LV3) && (c_c->q_ed - c_c->q_st < ix->filter_min_length || score < ix->filter_min_score))
				c_c->sum_score = 0;
		}
	}
	if (res->n_hit > 1) orc_msort(res->hit, res->n_hit, sizeof(orc_chain), chain_cmp_by_MEM_score);
	for (c_c = st_c; c_c < ed_c; c_c++) if (c_c->sum_score == 0) break;
	res->n_hit = c_c - st_c;
}

#define PRIMARY 1
#define SECONDARY 2
#define SUPPLYMENTARY 3
static void detect_primary(orc_chain *hit, uint32_t n_hit, uint32_t read_len)     /* cly.c:2995-3058 */
{
	if (n_hit == 0) return;
	int primary_v[800];
	uint8_t primary_v_idx[800];
	int n_primary_v = 1;
	hit->pri_index = primary_v_idx[0] = 0;
	primary_v[0] = 0;
	hit->primary = PRIMARY;
	orc_chain *ed_hit = hit + n_hit;
	for (orc_chain *c_hit = hit; c_hit < ed_hit; c_hit++) if (c_hit->q_st > 4294960000) c_hit->q_st = 0;
	for (orc_chain *c_hit = hit + 1; c_hit < ed_hit; c_hit++) {
		int overlap = 0;
		for (int i = 0; i < n_primary_v; i++) {
			int primary_st, primary_ed;
			if (hit[primary_v[i]].direction == c_hit->direction) {
				primary_st = hit[primary_v[i]].q_st;
				primary_ed = hit[primary_v[i]].q_ed;
			} else {
				primary_st = read_len - hit[primary_v[i]].q_ed;
				primary_ed = read_len - hit[primary_v[i]].q_st;
			}
			uint32_t overlap_st = MAX(c_hit->q_st, primary_st);
			uint32_t overlap_ed = MIN(c_hit->q_ed, primary_ed);
			if ((overlap_st < overlap_ed) && (((overlap_ed - overlap_st) << 1) >= (c_hit->q_ed - c_hit->q_st))) overlap = 1;
			if (overlap) {
				c_hit->primary = SECONDARY;
				c_hit->pri_index = ++primary_v_idx[i];
				int max_gap = MAX((hit[primary_v[i]].sum_score >> 6), 5);
				if (c_hit->sum_score + max_gap > hit[primary_v[i]].sum_score) c_hit->pri_index = 1;
				if (primary_v_idx[i] == 255) primary_v_idx[i] = 254;
				break;
			}
		}
		if (overlap == 0) {
			c_hit->primary = SUPPLYMENTARY;
			c_hit->pri_index = primary_v_idx[n_primary_v] = 0;
			primary_v[n_primary_v++] = c_hit - hit;
			if (n_primary_v > 750) n_primary_v = 750;
		}
	}
}

/* ------------------------------------------------------------------ classify_seq (cly.c:3064-3132) */
#define MIN_READ_LEN 40
void orc_classify_seq(const orc_index *ix, const char *seq, uint32_t read_len, orc_result *res, orc_buff *buff)
{
	search_dir_t sd[2];
	res->n_anc = 0;
	res->fast_classify = 1;
	res->n_hit = 0;
	res->n_seeds[0] = res->n_seeds[1] = 0;
	res->entered_final = 0;
	orc_cnt.n_reads++; orc_cnt.n_bases += read_len;
	if (read_len < MIN_READ_LEN) return;
	get_island(ix, seq, read_len, buff, res, sd);
	int both_direction = ((sd[0].total_score - sd[1].total_score) <= (sd[0].total_score >> 3)) ? 1 : 0;
	int super_repeat = fast_classify(ix, sd, read_len, res);
	if (both_direction) super_repeat += fast_classify(ix, sd + 1, read_len, res);
	resolve_tree(res);
	int run_slow_mode = 0;
	if (res->n_hit <= 0) run_slow_mode = 1;
	else if (res->hit[0].anchor_number < 5 && super_repeat < 3) {
		run_slow_mode = 1;
		if (read_len <= 300 && res->hit[0].sum_score > 200) run_slow_mode = 0;
	}
	if (run_slow_mode) {
		res->n_anc = 0;
		slow_classify(ix, sd, read_len, res);
		resolve_tree(res);
		if (both_direction || res->n_hit <= 0 || (res->hit[0].anchor_number < 5 && super_repeat < 3)) {
			slow_classify(ix, sd + 1, read_len, res);
			resolve_tree(res);
		}
	}
	delete_small_score_rst(ix, res, sd, buff, read_len);
	detect_primary(res->hit, (uint32_t)res->n_hit, read_len);
	orc_cnt.n_hits += res->n_hit;
}

/* ------------------------------------------------------------------ writers (cly_mt.c:60-344) */
static const char primary_string[3][4] = {"PRI", "SEC", "SUP"};
static void print_hit(FILE *out, const orc_chain *c, const orc_refinfo_t *r_i, int rst_cnt)     /* cly_mt.c:60-105 */
{
	fprintf(out, "%3d %s %s %20s ts:%-10d te:%-10d qs:%-10d qe:%-10d %-5d\t%d\t\n",
	        rst_cnt, primary_string[c->primary - 1], (c->direction) ? "F" : "R", r_i[c->ref_ID].name,
	        c->t_st, c->t_ed, c->q_st, c->q_ed, c->sum_score, c->indel);
}

void orc_write_result(FILE *out, const orc_index *ix, const orc_result *r, const char *name, const char *seq, const char *qual,
                      uint32_t read_len, int fmt, int max_sec_N)
{
	const orc_refinfo_t *r_i = ix->ri;
	if (fmt == 3 || fmt == 4) {           /* output_one_result_des / _full, cly_mt.c:158-246 */
		fprintf(out, "%s\t%s\t%s\t%ld\tn_rst:[%ld]\tn_anc:[%ld]\t\n", name, (r->n_hit) ? "CLASSIFY" : "UNCLASSIFY",
		        (r->fast_classify) ? "FAST" : "SLOW", (long)read_len, (long)r->n_hit, (long)r->n_anc);
		int rst_cnt = 0;
		for (size_t i = 0; i < r->n_hit; i++) if (r->hit[i].pri_index == 0) print_hit(out, r->hit + i, r_i, rst_cnt++);
		for (size_t i = 0; i < r->n_hit; i++)
			if (r->hit[i].pri_index > 0 && (fmt == 4 || r->hit[i].pri_index <= max_sec_N)) print_hit(out, r->hit + i, r_i, rst_cnt++);
		fprintf(out, "\n");
		return;
	}
	const char *seq_s = (fmt == 2) ? seq : "*", *qual_s = (fmt == 2) ? qual : "*";     /* output_one_result_sam, cly_mt.c:248-344 */
	if (r->n_hit == 0) {
		fprintf(out, "%s\t4\t*\t0\t0\t*\t*\t0\t0\t%s\t%s\t\n", name, seq_s, qual_s);
		return;
	}
	const orc_chain *c_s = r->hit, *c_e = c_s + r->n_hit;
	int flag = c_s->direction ? 0 : 0x10;
	int mapQ_PRI = 0;
	if (r->n_hit == 1 || (c_s->sum_score - c_s[1].sum_score > 5)) mapQ_PRI = 30;
	else mapQ_PRI = (c_s->sum_score - c_s[1].sum_score) << 2;
	fprintf(out, "%s\t%d\t%s\t%d\t%d\t%dS%dM%dS\t*\t0\t0\t%s\t%s\tAS:i:%d\t\n", name, flag, r_i[c_s->ref_ID].name, c_s->t_st, mapQ_PRI,
	        c_s->q_st, c_s->q_ed - c_s->q_st, read_len - c_s->q_ed, seq_s, qual_s, c_s->sum_score);
	for (int loop = 0; loop <= 1; loop++)
		for (const orc_chain *c = c_s + 1; c < c_e; c++) {
			int show = 0, fl = c->direction ? 0 : 0x10, mapQ = 0;
			if (loop == 0 && c->pri_index == 0) { show = 1; fl += 0x800; mapQ = MIN(30, mapQ_PRI); }
			else if (loop == 1 && c->pri_index > 0 && c->pri_index <= max_sec_N) { show = 1; fl += 0x100; }
			if (show)
				fprintf(out, "%s\t%d\t%s\t%d\t%d\t%d%c%dM%d%c\t*\t0\t0\t*\t*\tAS:i:%d\t\n", name, fl, r_i[c->ref_ID].name, c->t_st, mapQ,
				        c->q_st, (loop == 0) ? 'H' : 'S', c->q_ed - c->q_st, read_len - c->q_ed, (loop == 0) ? 'H' : 'S', c->sum_score);
		}
}
