// dsb_classify.cuh -- the warp-per-read phases of the classifier (anchor chaining, 9-mer sparse-DP scoring) and the per-read
// state between the phase kernels.  Behaviour follows the reference function by function (citations inline); structure is
// this project's own: anchors/chains are index-linked arrays in HBM pools, the reference window and the hash table of its
// scanned 9-mers live in shared memory, the read range is streamed through that table.  Seeding is in dsb_seedcore.h /
// dsb_seed.cuh (lane per island seed over a flat task list).
#pragma once
#include "dsb_device.cuh"
#include <climits>
#include "../../include/desamba_b200.h"

#define S_A_KEMR_L 9
#define OVER_SEARCH_M2 50
#define MIN_SCORE_MEM 12

struct DevAnchor {              // Anchor (cly.h:44-61), only the fields read after map_seed
	uint32_t ref_ID, ref_offset, index_in_read;
	int32_t  pre;               // chain_anchor_pre as index, -1 = NULL
	uint16_t mtch_len; int16_t score;
	uint8_t  direction, useless, duplicate, pad;
};
struct DevChain {               // chain_item (cly.h:69-89)
	uint32_t ref_ID; int32_t q_t_dis; uint32_t sum_score, anchor_number;
	uint32_t t_st, t_ed, q_st, q_ed, indel;
	int32_t  cur;               // chain_anchor_cur as index
	uint8_t  direction, with_top_anchor, primary, pri_index;
};
struct DevSms { uint32_t t_pos, q_pos, len, score; };   // spd_match (cly.h:127-133)
struct ScHash { uint16_t next; uint16_t seed_ID; uint16_t s_or_e; uint16_t pad; };   // seed_con_hash (cly.h:120-125)

struct WarpSmem {               // per-warp shared memory
	uint8_t  refwin[2176];      // ref[2000] of sdp_middle_M2 / ref[1000] of sdp_right/left_M2 (+ over-read slack)
};
// per-warp shared memory of the scoring kernels: hash table of the 9-mers of the scanned TARGET positions of one sdp_match
#define TT_SLOTS 1024           // >= 2 x the scanned positions of a window (t_len < 2000 -> at most 497)
#define TT_EMPTY 0xffffffffu
struct MatchSmem {
	uint32_t tslot[TT_SLOTS];   // open addressing: (9-mer << 9) | scanned position index, or TT_EMPTY
	uint32_t bloom[512];        // one hash bit per scanned 9-mer (16 bits per table slot): the streaming loop tests this first
	uint32_t n_cand, n_tmp, pad[2];
};

struct WarpScratch {            // per-warp HBM scratch of the chaining / scoring kernels
	DevAnchor *anc, *anc_tmp;   // anc: the read's anchors (in the anchor pool); anc_tmp: merge-sort scratch of chain_insert_M3
	DevChain  *chain, *chain_tmp;
	DevSms    *sms;
	int       *score_v;         // 1024
	ScHash    *sc_hash;         // 256 + 2*400 + 8
	DevSms    *sms_tmp;         // sdp_match: matches of the current window before they are ordered (max_matches)
	uint2     *cand;            // sdp_match: (scanned target index, read position) pairs with equal 9-mers (CAND_CAP)
	uint64_t  *sort_key[2];     // sdp_match: order keys of sms_tmp + merge-sort ping-pong (max_matches each)
	uint32_t  *sort_idx[2];     // merge-sort permutation ping-pong (max_matches each)
};

// per-read state that lives in HBM between the phase kernels
struct ReadWork {
	uint32_t anc_off, n_anc;      // the read's anchors in the anchor pool (cly_r.anchor_v)
	uint32_t chain_off, n_chain;  // chains kept by resolve_tree, in the chain pool
	uint16_t error; uint8_t fast_classify, pad;
};
enum { LIST_SLOW0 = 0, LIST_SLOW1 = 1, LIST_SCORE = 2, LIST_SCORE_HEAVY = 3, N_LISTS = 4 };
// HEAVY: a read whose sparse DP collects more than DEFER_SMS matches in one extension (repeats) gives up in k_score and is
// re-scored from its (untouched) pool chains by k_score_heavy, a whole CTA per read
#define DEFER_SMS 1024
#define ERR_DEFER 7
enum { PASS_FAST = 0, PASS_SLOW0 = 1, PASS_SLOW1 = 2 };
// control block (u32): [0..3] list lengths, [5] anchor pool cursor, [6] chain pool cursor, [7] pools that overflowed (OVF_*),
// [8..31] work cursors of the launches, [32..34] seed tasks of the three seeding passes, [36..38] their fetch cursors,
// [40..42] staging-chunk cursors of the passes
enum { CTL_LIST_N = 0, CTL_ANC_CURSOR = 5, CTL_CHAIN_CURSOR = 6, CTL_OVERFLOW = 7, CTL_CURSOR = 8, CTL_TASK_N = 32, CTL_TASK_CURSOR = 36, CTL_CHUNK_CURSOR = 40, CTL_WORDS = 48 };
enum { OVF_TASKS = 1, OVF_CHUNKS = 2, OVF_ANCHORS = 4, OVF_CHAINS = 8, OVF_HITS = 16, OVF_STUCK = 1u << 30 };

struct ClassifyParams {
	DevIndex ix;
	uint32_t n_reads;
	const uint64_t *read_off;   // n_reads+1, offsets into the concatenated ASCII reads
	const uint64_t *bin_off;    // per read: offset of [GUARD | fwd | rev | GUARD] in bin
	const uint8_t  *bin;
	const uint32_t *seed_off;   // per read: first seed slot (per strand array)
	const dsb_seed *seeds[2];   // [0] forward strand, [1] reverse strand
	const uint32_t *n_seeds[2];
	const uint32_t *total_score[2];
	const uint32_t *order;      // read ids, longest first (work order of the chaining pass over all reads)
	uint32_t *prof;             // per read: 8 x u32 phase times in units of 1024 cycles (-, chain, -, sdp_match [part of the next three], middle, right, left, total)
	// seeding: task lists / per-seed records of the passes (ping-pong: fast and slow 1 use [0], slow 0 uses [1]), staging chunks,
	// the tasks of a (strand, read) in the current pass
	SeedTaskRef *tasks[2]; SeedRec *recs[2]; uint32_t task_cap;
	const uint4 *chunks;
	uint32_t *task_first[2], *task_cnt[2];
	// state between the phase kernels
	ReadWork *work;
	DevAnchor *anc_pool; uint32_t anc_pool_cap;
	DevChain *chain_pool; uint32_t chain_pool_cap;
	uint32_t *list[N_LISTS];
	uint32_t *ctl;
	// per-warp scratch
	uint8_t  *scratch; uint64_t scratch_stride;
	uint32_t max_anchors, max_matches;
	// outputs
	dsb_read_result *rr;
	dsb_hit *hits; uint64_t hits_cap; unsigned long long *hits_cursor;
	unsigned long long *counters;
};

struct SearchDir {              // SEARCH_DIR (cly.c:946-954)
	const dsb_seed *seed_v; uint32_t l_seed_v;
	const uint8_t *bin_read;
	uint32_t direction, total_score;
};

struct DpTeam;
struct ReadState {
	const DevIndex *ix;
	WarpSmem *sm;
	MatchSmem *mt;               // scoring kernels only
	uint32_t l_read;             // scoring: length of the read (positions that have a 9-mer: 0 .. l_read - 9)
	DpTeam *team;                // helper warps of the CTA for the sparse DP of heavy reads (k_score_heavy), else nullptr
	WarpScratch ws;
	uint32_t n_anc, n_hit, n_sms;
	uint32_t max_anchors, max_matches;
	uint32_t fast_classify;
	int error;
	int sp_l;                    // SP_SET.l
	uint32_t c_prefix, c_occ, c_locate, c_getref, c_getref_bytes;   // algorithmic counters
	long long t_phase[8];       // clock64 per phase
};
#define PH_BEGIN() const long long ph_t0_ = clock64()
#define PH_END(S, k) (S).t_phase[k] += clock64() - ph_t0_
#define CNT_GETREF(S, len) do { (S).c_getref++; (S).c_getref_bytes += ((uint32_t)DSB_MAX((int)(len), 0) + 3) >> 2; } while (0)

#include "dsb_seed.cuh"

// ---------------------------------------------------------------- chaining (cly.c:38-52, 72-112, 201-349)
struct ChainCmpByScore {
	__device__ int operator()(const DevChain &a, const DevChain &b) const {
		if (a.with_top_anchor != b.with_top_anchor) return (a.with_top_anchor) ? (-1) : (1);
		int score_a = a.sum_score + ((a.q_ed - a.q_st) << 1);
		score_a -= (a.indel << 2);
		int score_b = b.sum_score + ((b.q_ed - b.q_st) << 1);
		score_b -= (b.indel << 2);
		if (score_a < score_b) return 1;
		if (score_a > score_b) return -1;
		return 0;
	}
};

#define MAX_dis_MINUS 30
#define MAX_waiting_len 400
__device__ __forceinline__ void chain_insert_M2(ReadState &S, int32_t ai)
{
	DevAnchor *A = S.ws.anc; DevChain *C = S.ws.chain;
	const DevAnchor an = A[ai];
	const int32_t dis = an.ref_offset - an.index_in_read;
	const uint32_t ref_l = an.ref_offset, ref_r = ref_l + an.mtch_len;
	const uint32_t read_l = an.index_in_read, read_r = read_l + an.mtch_len;
	int dis_minus = 0;
	for (uint32_t k = 0; k < S.n_hit; k++) {
		DevChain c = C[k];
		if (c.direction == an.direction && c.ref_ID == an.ref_ID &&
		    (dis_minus = DSB_ABS(dis - c.q_t_dis)) < MAX_dis_MINUS &&
		    DSB_ABS_U(c.t_ed, an.ref_offset) < MAX_waiting_len) {
			// chain_insert_meta, new_chain == false (cly.c:95-111)
			c.with_top_anchor |= (!an.useless);
			if (c.q_ed >= read_r) { C[k] = c; return; }
			c.t_ed = DSB_MAX(ref_r, c.t_ed);
			c.q_ed = read_r;
			A[ai].pre = c.cur;
			c.cur = ai;
			c.q_t_dis = an.ref_offset - an.index_in_read;
			c.indel += dis_minus;
			c.anchor_number++;
			c.sum_score += (an.duplicate) ? 1 : an.score;
			C[k] = c;
			return;
		}
	}
	DevChain c;                                   // chain_insert_meta, new_chain == true (cly.c:78-94)
	A[ai].pre = -1;
	c.ref_ID = an.ref_ID; c.direction = an.direction;
	c.q_t_dis = an.ref_offset - an.index_in_read;
	c.t_st = ref_l; c.t_ed = ref_r; c.q_st = read_l; c.q_ed = read_r;
	c.with_top_anchor = !an.useless;
	c.anchor_number = 1;
	c.sum_score = (an.duplicate) ? 1 : an.score;
	c.indel = 0; c.cur = ai; c.primary = 0; c.pri_index = 0;
	C[S.n_hit++] = c;
}

struct AnchorCmp {                                // Anchor_cmp_by_chr_ID_and_pos (cly.c:226-235): a 0/1 comparator
	__device__ int operator()(const DevAnchor &a, const DevAnchor &b) const {
		if (a.ref_ID != b.ref_ID) return a.ref_ID > b.ref_ID;
		if (a.direction != b.direction) return a.direction > b.direction;
		return a.ref_offset > b.ref_offset;
	}
};

// The reference sorts with qsort (= glibc merge sort, taking the left run while cmp <= 0).  For a comparator that is a
// consistent ordering, that is THE stable sort by the comparator's key, so it can be computed any way: here each element's
// final position is counted directly (elements with a smaller key, or an equal key and a smaller index), 32 elements at a
// time, and the array is permuted through scratch.  (chain_cmp_by_MEM_score is NOT consistent: k_finalize keeps the
// explicit merge emulation for it.)
template <typename T>
__device__ __noinline__ void warp_stable_sort_by_key(T *a, T *tmp, uint32_t n, uint64_t *key)
{
	const int lane = lane_id();
	__syncwarp();
	for (uint32_t i0 = 0; i0 < n; i0 += 32) {
		const uint32_t i = i0 + lane;
		const uint64_t ki = (i < n) ? key[i] : 0;
		uint32_t r = 0;
		#pragma unroll 4
		for (uint32_t j = 0; j < n; j++) { const uint64_t kj = key[j]; r += (kj < ki || (kj == ki && j < i)) ? 1u : 0u; }
		if (i < n) tmp[r] = a[i];
	}
	__syncwarp();
	for (uint32_t i = lane; i < n; i += 32) a[i] = tmp[i];
	__syncwarp();
}

// Sorted permutation of n 64-bit keys by a warp, for large n (counting ranks is O(n^2)): chunks of CHUNK keys are rank-sorted, then merged pairwise (every
// element finds its place in the other run by binary search; the left run wins ties, so the order is stable).  key0 holds the
// keys; key0/key1 and idx0/idx1 are ping-pong buffers of n entries.  Returns the index array: result[i] = index of the i-th
// smallest key.
// Warp `w_` of NW warps of a CTA work together when NW > 1 (k_score_heavy's team, all of them called with the same arguments):
// chunks and elements are dealt round-robin, the passes are separated by CTA barriers.
template <uint32_t CHUNK, int NW = 1>
__device__ __noinline__ const uint32_t *warp_sort_perm(uint32_t n, uint64_t *key0, uint64_t *key1, uint32_t *idx0, uint32_t *idx1, int w_ = 0)
{
	const int lane = lane_id(), w = NW > 1 ? w_ : 0, nw = NW;
	if (nw > 1) __syncthreads(); else __syncwarp();
	for (uint32_t c0 = (uint32_t)w * CHUNK; c0 < n; c0 += (uint32_t)nw * CHUNK) {     // chunk sort: (key0, identity) -> (key1, idx1)
		const uint32_t cn = DSB_MIN((uint32_t)CHUNK, n - c0);
		for (uint32_t i0 = 0; i0 < cn; i0 += 32) {
			const uint32_t i = i0 + lane;
			const uint64_t ki = (i < cn) ? key0[c0 + i] : 0;
			uint32_t r = 0;
			for (uint32_t j = 0; j < cn; j++) { const uint64_t kj = key0[c0 + j]; r += (kj < ki || (kj == ki && j < i)) ? 1u : 0u; }
			if (i < cn) { key1[c0 + r] = ki; idx1[c0 + r] = c0 + i; }
		}
	}
	if (nw > 1) __syncthreads(); else __syncwarp();
	uint64_t *ks = key1, *kd = key0; uint32_t *is = idx1, *id = idx0;
	for (uint32_t wd = CHUNK; wd < n; wd <<= 1) {                        // merge runs of width wd
		for (uint32_t p = (uint32_t)w * 32 + lane; p < n; p += (uint32_t)nw * 32) {
			const uint32_t pair = p / (2 * wd) * (2 * wd), l0 = pair, l1 = DSB_MIN(pair + wd, n), r1 = DSB_MIN(pair + 2 * wd, n);
			const uint64_t k = ks[p];
			uint32_t dst;
			if (p < l1) {                                                // left run: elements of the right run that are < k come first
				uint32_t lo = l1, hi = r1;
				while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (ks[mid] < k) lo = mid + 1; else hi = mid; }
				dst = p + (lo - l1);
			} else {                                                     // right run: elements of the left run that are <= k come first
				uint32_t lo = l0, hi = l1;
				while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (ks[mid] <= k) lo = mid + 1; else hi = mid; }
				dst = (p - l1) + lo;
			}
			kd[dst] = k; id[dst] = is[p];
		}
		if (nw > 1) __syncthreads(); else __syncwarp();
		uint64_t *tk = ks; ks = kd; kd = tk;
		uint32_t *ti = is; is = id; id = ti;
	}
	return is;
}

#define SORT_CHUNK 1024
template <typename T>
__device__ __noinline__ void warp_stable_sort_large(T *a, T *tmp, uint32_t n, uint64_t *key0, uint64_t *key1, uint32_t *idx0, uint32_t *idx1)
{
	const uint32_t *is = warp_sort_perm<SORT_CHUNK>(n, key0, key1, idx0, idx1);
	for (uint32_t i = lane_id(); i < n; i += 32) tmp[i] = a[is[i]];
	__syncwarp();
	for (uint32_t i = lane_id(); i < n; i += 32) a[i] = tmp[i];
	__syncwarp();
}

#define MAX_ANCHOR_OVERLAP 3
__device__ __noinline__ void chain_insert_M3(ReadState &S)
{
	int *score_v = S.ws.score_v;
	DevAnchor *A = S.ws.anc; DevChain *C = S.ws.chain;
	const int32_t n = (int32_t)S.n_anc;
	{	// qsort by (ref_ID, direction, ref_offset), cly.c:226-243
		uint64_t *key = (uint64_t *)S.ws.sms;                  // the match buffer is idle while chaining
		if ((uint64_t)n * 8 <= (uint64_t)S.max_matches * sizeof(DevSms)) {
			for (int32_t i = lane_id(); i < n; i += 32) { const DevAnchor a = A[i]; key[i] = ((uint64_t)a.ref_ID << 33) | ((uint64_t)(a.direction ? 1 : 0) << 32) | a.ref_offset; }
			if (n <= 2 * SORT_CHUNK) warp_stable_sort_by_key(A, S.ws.anc_tmp, (uint32_t)n, key);
			else if ((uint64_t)n * 16 <= (uint64_t)S.max_matches * sizeof(DevSms) && (uint64_t)n * 8 <= (uint64_t)S.max_anchors * sizeof(DevChain))
				warp_stable_sort_large(A, S.ws.anc_tmp, (uint32_t)n, key, key + n, (uint32_t *)S.ws.chain_tmp, (uint32_t *)S.ws.chain_tmp + n);   // chain_tmp is idle here
			else warp_stable_sort_by_key(A, S.ws.anc_tmp, (uint32_t)n, key);
		} else
			glibc_msort(A, S.ws.anc_tmp, n, AnchorCmp());
	}
	for (int32_t chr_st = 0; chr_st < n;) {
		int32_t chr_ed = chr_st + 1, c_a;
		const uint32_t ref_ID = A[chr_st].ref_ID;
		const uint32_t direction = A[chr_st].direction;
		for (; chr_ed < n && A[chr_ed].ref_ID == ref_ID && A[chr_ed].direction == direction &&
		       A[chr_ed].ref_offset - A[chr_ed - 1].ref_offset < 2000; chr_ed++);
		if (chr_ed - chr_st > 1024) chr_ed = chr_st + 1024;
		int32_t max_anchor = -1; int max_score = 0, anchor_max_score;
		for (c_a = chr_st; c_a < chr_ed; c_a++) {
			const DevAnchor ca = A[c_a];
			int32_t ca_pre = -1;
			anchor_max_score = ca.score;
			const uint32_t max_t = ca.ref_offset + MAX_ANCHOR_OVERLAP;
			const uint32_t max_q = ca.index_in_read + MAX_ANCHOR_OVERLAP;
			// look-back over earlier anchors of the group, 32 per step (lane 0 = nearest); the reference's sequential loop takes
			// the first maximum it meets walking backwards and stops at the first `break` (cly.c:283-303)
			int best_score = INT_MIN, best_pre = -1;
			for (int32_t base = c_a - 1; base >= chr_st; base -= 32) {
				const int32_t pre = base - lane_id();
				const bool act = pre >= chr_st;
				DevAnchor pa; pa.index_in_read = 0; pa.ref_offset = 0; pa.mtch_len = 0;
				if (act) pa = A[pre];
				const bool pass = act && !(pa.index_in_read + pa.mtch_len > max_q) && !(pa.ref_offset + pa.mtch_len > max_t);
				const bool brk = pass && ((pa.index_in_read + 1000 < max_q) || (pa.ref_offset + 1000 < max_t));
				const uint32_t bm = __ballot_sync(DSB_FULL, brk);
				const int first = bm ? (__ffs(bm) - 1) : 32;
				if (pass && lane_id() < first) {
					const int indel = pa.index_in_read - pa.ref_offset - (max_q - max_t);
					const int ABS_indel = DSB_ABS(indel);
					if (!(ABS_indel > 200)) {
						const int new_score = score_v[pre - chr_st] + ca.mtch_len - (ABS_indel >> 4) - ((max_q - pa.index_in_read) >> 8);
						if (new_score > best_score) { best_score = new_score; best_pre = pre; }      // lanes walk downwards: first maximum = highest pre
					}
				}
				if (bm) break;
			}
			{	// arg-max over lanes: highest score, ties -> highest pre
				#pragma unroll
				for (int d = 16; d; d >>= 1) {
					const int os = __shfl_xor_sync(DSB_FULL, best_score, d), op = __shfl_xor_sync(DSB_FULL, best_pre, d);
					if (os > best_score || (os == best_score && op > best_pre)) { best_score = os; best_pre = op; }
				}
				if (best_pre >= 0 && best_score > anchor_max_score) { anchor_max_score = best_score; ca_pre = best_pre; }
			}
			if (lane_id() == 0) { A[c_a].pre = ca_pre; score_v[c_a - chr_st] = anchor_max_score; }
			__syncwarp();
			if (max_score < anchor_max_score) { max_score = anchor_max_score; max_anchor = c_a; }
		}
		int sum_INDEL = 0, anchor_number = 1; int32_t pre = max_anchor;
		int sum_score = (A[max_anchor].duplicate) ? 1 : A[max_anchor].score;
		int with_top = !A[max_anchor].useless;
		for (; A[pre].pre != -1; anchor_number++) {
			const int32_t pre_ = A[pre].pre;
			sum_INDEL += (A[pre].index_in_read - A[pre_].index_in_read) - (A[pre].ref_offset - A[pre_].ref_offset);
			with_top |= (!A[pre].useless);
			sum_score += (A[pre].duplicate) ? 1 : A[pre].score;
			pre = pre_;
		}
		DevChain c;
		c.ref_ID = ref_ID; c.direction = (uint8_t)direction;
		c.q_t_dis = A[max_anchor].ref_offset - A[max_anchor].index_in_read;
		c.t_st = A[pre].ref_offset;
		c.t_ed = A[max_anchor].ref_offset + A[max_anchor].mtch_len;
		c.q_st = A[pre].index_in_read;
		c.q_ed = A[max_anchor].index_in_read + A[max_anchor].mtch_len;
		c.with_top_anchor = (uint8_t)with_top;
		c.anchor_number = anchor_number;
		c.sum_score = sum_score;
		c.indel = sum_INDEL;
		c.cur = max_anchor; c.primary = 0; c.pri_index = 0;
		C[S.n_hit++] = c;
		chr_st = chr_ed;
	}
}

__device__ __noinline__ void resolve_tree(ReadState &S)     // cly.c:326-349
{
	S.n_hit = 0;
	if (S.n_anc < 50)
		for (int32_t a = 0; a < (int32_t)S.n_anc; a++) chain_insert_M2(S, a);
	else
		chain_insert_M3(S);
	if (S.n_hit > 32 && (uint64_t)S.n_hit * 8 <= (uint64_t)S.max_matches * sizeof(DevSms)) {       // chain_cmp_by_score (cly.c:38-52) as a key
		uint64_t *key = (uint64_t *)S.ws.sms;
		__syncwarp();
		for (uint32_t i = lane_id(); i < S.n_hit; i += 32) {
			const DevChain c = S.ws.chain[i];
			int score = c.sum_score + ((c.q_ed - c.q_st) << 1);
			score -= (c.indel << 2);
			key[i] = ((uint64_t)(c.with_top_anchor ? 0 : 1) << 32) | (uint32_t)~((uint32_t)score ^ 0x80000000u);   // with_top first, then descending (signed) score
		}
		warp_stable_sort_by_key(S.ws.chain, S.ws.chain_tmp, S.n_hit, key);
	} else if (S.n_hit > 1) glibc_msort(S.ws.chain, S.ws.chain_tmp, (int)S.n_hit, ChainCmpByScore());
	uint32_t rst_num = DSB_MIN(5u, S.n_hit);
	while (rst_num < S.n_hit && S.ws.chain[rst_num].with_top_anchor == 1) rst_num++;
	S.n_hit = rst_num;
}

// ---------------------------------------------------------------- 9-mer sparse DP scoring (cly.c:1691-1818, 2335-2849)
__device__ __forceinline__ void sc_hash_idx(ReadState &S)      // cly.c:1691-1710
{
	ScHash *sc_hash = S.ws.sc_hash;
	for (int i = lane_id(); i < 256; i += 32) { ScHash z; z.next = 0; z.seed_ID = 0; z.s_or_e = 0; z.pad = 0; sc_hash[i] = z; }
	__syncwarp();
	int sc_con_index = 256;
	for (uint32_t h = 0; h < S.n_hit; h++) {
		const DevChain c_h = S.ws.chain[h];
		for (int i = 1; i >= 0; i--) {
			uint16_t c_key = ((i == 1) ? (c_h.t_st - c_h.q_st) : (c_h.t_ed - c_h.q_ed)) & 0xff;
			while (sc_hash[c_key].next != 0) c_key = sc_hash[c_key].next;
			ScHash e; e.seed_ID = (uint16_t)((h + 1) & 0x7fff); e.s_or_e = (uint16_t)i; e.next = (uint16_t)sc_con_index; e.pad = 0;
			sc_hash[c_key] = e;
			ScHash z; z.next = 0; z.seed_ID = 0; z.s_or_e = 0; z.pad = 0;
			sc_hash[sc_con_index++] = z;
		}
	}
}

__device__ __noinline__ int combine_chain(ReadState &S, int chain_ID, int dis, int isleft, int c_q_pos, int *combined)
{   // cly.c:1763-1808
	const ScHash *sc_hash = S.ws.sc_hash;
	DevChain *c_st = S.ws.chain;
	uint16_t key = (dis) & 0xff;
	while (sc_hash[key].next != 0) {
		const uint16_t seed_ID = sc_hash[key].seed_ID;
		const int ci = seed_ID - 1;
		const DevChain c = c_st[ci];
		const int dis_con = (isleft) ? (c.t_ed - c.q_ed) : (c.t_st - c.q_st);
		const int q_pos_con = (!isleft) ? (c.q_st) : (c.q_ed - S_A_KEMR_L);
		if (dis == dis_con && chain_ID != ci && isleft != (int)sc_hash[key].s_or_e && DSB_ABS_U(c_q_pos, q_pos_con) < 8 &&
		    c_st[chain_ID].ref_ID == c.ref_ID && c_st[chain_ID].direction == c.direction && c.sum_score != 0 && seed_ID - 1 > chain_ID) {
			DevChain h = c_st[chain_ID];
			h.sum_score += c.sum_score;
			h.anchor_number += c.anchor_number;
			h.indel += c.indel;
			h.q_st = DSB_MIN(h.q_st, c.q_st);
			h.t_st = DSB_MIN(h.t_st, c.t_st);
			h.q_ed = DSB_MAX(h.q_ed, c.q_ed);
			h.t_ed = DSB_MAX(h.t_ed, c.t_ed);
			c_st[chain_ID] = h;
			DevChain z = c;
			z.sum_score = 0; z.t_st = z.t_ed = z.q_st = z.q_ed = 0;
			c_st[ci] = z;
			*combined = ci;
			return 1;
		}
		key = sc_hash[key].next;
	}
	return 0;
}

// MEM_search (cly.c:1810-1818); q in HBM (read strands incl. guards), t in the shared reference window
__device__ __forceinline__ int MEM_search_fwd(const uint8_t *q, const uint8_t *t, int max)
{
	int len = 0;
	for (; len < max && __ldg(q) == *t; len++, q++, t++);
	return len;
}
__device__ __forceinline__ int MEM_search_bwd(const uint8_t *q, const uint8_t *t, int max)
{
	int len = 0;
	for (; len < max && __ldg(q) == *t; len++, q--, t--);
	return len;
}

__device__ __forceinline__ bool sms_push(ReadState &S, uint32_t t_pos, uint32_t q_pos, uint32_t len)
{
	if (S.n_sms >= S.max_matches) { S.error = 2; return false; }
	DevSms *p = S.ws.sms + S.n_sms++;
	p->t_pos = t_pos; p->q_pos = q_pos; p->len = len;
	return true;
}

struct SdpArgs { uint32_t q_bg, q_ed; const uint8_t *q_str, *t_str; uint32_t t_len, t_st; bool fwd; };
struct FlushArgs { SdpArgs A; MatchSmem *M; const uint2 *cand; DevSms *sms_tmp; uint64_t *key; uint32_t n_sms, max_matches; };
// Heavy reads (repeats: thousands of matches per window) are scored by k_score_heavy with a whole CTA: warp 0 runs the
// read, the other warps only help with phase (A) below -- each takes a contiguous range of the earlier matches.
#ifndef TEAM_WARPS
#define TEAM_WARPS 32
#endif
#define TEAM_MIN_FIRST 128
struct DpTeam {
	int cmd;                          // 1: DP look-back posted, 2: sort posted, 3: candidate extension posted, < 0: exit
	int kind; uint32_t first, nb;
	uint32_t sort_n; uint64_t *sort_key[2]; uint32_t *sort_idx[2];
	FlushArgs flush; uint32_t flush_n;
	const DevSms *sms;
	DevSms item[32]; int stopped0[32];
	int best[2][TEAM_WARPS][32]; uint8_t brk[2][TEAM_WARPS][32];   // per round (ping-pong), tile and lane: best candidate, walk hit its break
	uint4 tile[TEAM_WARPS][32];       // predecessor tile, one per warp
};

// ---------------------------------------------------------------- sdp_match (cly.c:2335-2440) without a per-read index
// The reference hashes all 9-mers of the READ once (build_hash_table_M2, cly.c:2173-2224) and looks up every 4th 9-mer of a
// reference window; only read positions inside [q_bg, q_ed] (<= 2000 wide) count.  A per-read index is ~100 KB of scattered
// per-warp scratch (it was 100 B/base of DRAM traffic and the top stall of k_score, profiles/r1f).  Here the roles are
// swapped: the <= 497 scanned TARGET 9-mers of the window go into a small hash table in shared memory, and the warp streams
// the read range [q_bg, q_ed] through it -- 16 consecutive positions per lane from one coalesced 16-byte load, the 9-mers cut
// out of a 48-bit register.  Equal 9-mers give (target index, read position) candidates; each candidate then runs the
// reference's own test + extension on a lane of its own (sdp_extend), and the surviving matches are put into the
// reference's push order: target scan order, ascending read position (the order the hash chains are walked in).
#define CAND_CAP (32 * 512 + 4096)
// The per-warp shared structures are reached through pointers kept in ReadState, which the compiler can only treat as
// generic addresses (LD.E / generic atomics); these wrappers keep the accesses in the shared window (LDS / STS / ATOMS).
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_u32_ro(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }   // table is read-only while streaming
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t atoms_cas(uint32_t a, uint32_t cmp, uint32_t v) { uint32_t o; asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(o) : "r"(a), "r"(cmp), "r"(v) : "memory"); return o; }
__device__ __forceinline__ uint32_t atoms_add(uint32_t a, uint32_t v) { uint32_t o; asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory"); return o; }
__device__ __forceinline__ void reds_or(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }

__device__ __forceinline__ uint32_t pack4(uint32_t w) { return ((w & 0x03030303u) * 0x40100401u) >> 24; }   // 4 base codes (first in the low byte) -> 8 bits, first base in the high bits
__device__ __forceinline__ uint32_t tt_hash(uint32_t kmer, uint32_t tbits) { return (kmer * 0x9E3779B1u) >> (32 - tbits); }

// the per-(scanned target position, read position) body of sdp_match (cly.c:2353-2384 / 2402-2433)
__device__ __forceinline__ bool sdp_extend(const SdpArgs &A, uint32_t k, uint32_t cur, DevSms &m)
{
	const int i = 4 + 4 * (int)k;
	const uint8_t *c_t_str = A.fwd ? (A.t_str + i) : (A.t_str + A.t_len - S_A_KEMR_L - i);
	if (A.fwd) {
		const int back_len = MEM_search_bwd(A.q_str + cur - 1, c_t_str - 1, 4);
		if (back_len < 4 || i == 4) {
			uint32_t max_search = A.q_ed - cur - 1;
			max_search = DSB_MIN(max_search, A.t_len - i - 1) + OVER_SEARCH_M2;
			const int forward_len = MEM_search_fwd(A.q_str + cur + S_A_KEMR_L, c_t_str + S_A_KEMR_L, max_search);
			const int total_len = back_len + forward_len + 1;
			if (total_len >= 4) { m.t_pos = i - back_len + A.t_st; m.q_pos = cur - back_len; m.len = total_len; return true; }
		}
	} else {
		const int forward_len = MEM_search_fwd(A.q_str + cur + S_A_KEMR_L, c_t_str + S_A_KEMR_L, 4);
		if (forward_len < 4 || i == 4) {
			uint32_t max_search = cur;
			max_search = DSB_MIN((long)max_search, (long)(c_t_str - A.t_str)) + OVER_SEARCH_M2;
			const int back_len = MEM_search_bwd(A.q_str + cur - 1, c_t_str - 1, max_search);
			const int total_len = back_len + forward_len + 1;
			if (total_len >= 4) { m.t_pos = (uint32_t)((long)(c_t_str - A.t_str) - back_len + A.t_st); m.q_pos = cur - back_len; m.len = total_len; return true; }
		}
	}
	return false;
}

// the collected candidates, 32 at a time: test + extension; a surviving match goes (unordered) into sms_tmp with its order
// key (target index, read position of the 9-mer).  Warp `w` of `nw` (k_score_heavy's team) takes every nw-th group of 32.
__device__ __noinline__ void flush_range(const FlushArgs &F, uint32_t nc, int w, int nw)
{
	const uint32_t a_ntmp = smem_addr(&F.M->n_tmp);
	for (uint32_t c0 = (uint32_t)w * 32; c0 < nc; c0 += (uint32_t)nw * 32) {
		const uint32_t c = c0 + lane_id();
		if (c < nc) {
			const uint2 cd = F.cand[c];
			DevSms m; m.score = 0;
			if (sdp_extend(F.A, cd.x, cd.y, m)) {
				const uint32_t slot = atoms_add(a_ntmp, 1u);
				if (F.n_sms + slot < F.max_matches) { F.sms_tmp[slot] = m; F.key[slot] = ((uint64_t)cd.x << 32) | cd.y; }
			}
		}
	}
}

template <bool HEAVY>
__device__ __noinline__ void sdp_flush_cand(ReadState &S, const SdpArgs &A)
{
	MatchSmem *M = S.mt;
	__syncwarp();
	const uint32_t a_ncand = smem_addr(&M->n_cand);
	const uint32_t nc = DSB_MIN(lds_u32(a_ncand), (uint32_t)CAND_CAP);
	if (HEAVY && nc >= 1024) {                             // k_score_heavy: the whole CTA extends the candidates
		DpTeam *T = S.team;
		if (lane_id() == 0) {
			FlushArgs F; F.A = A; F.M = M; F.cand = S.ws.cand; F.sms_tmp = S.ws.sms_tmp; F.key = S.ws.sort_key[0]; F.n_sms = S.n_sms; F.max_matches = S.max_matches;
			T->flush = F; T->flush_n = nc; T->cmd = 3;
		}
		__syncthreads();                                   // job posted: the helper warps wake up
		flush_range(T->flush, nc, 0, TEAM_WARPS);
		__syncthreads();                                   // every candidate has been looked at
	} else {
		const uint32_t a_ntmp = smem_addr(&M->n_tmp);
		for (uint32_t c0 = 0; c0 < nc; c0 += 32) {
			const uint32_t c = c0 + lane_id();
			if (c < nc) {
				const uint2 cd = S.ws.cand[c];
				DevSms m; m.score = 0;
				if (sdp_extend(A, cd.x, cd.y, m)) {
					const uint32_t slot = atoms_add(a_ntmp, 1u);
					if (S.n_sms + slot < S.max_matches) { S.ws.sms_tmp[slot] = m; S.ws.sort_key[0][slot] = ((uint64_t)cd.x << 32) | cd.y; }
				}
			}
		}
		__syncwarp();
	}
	if (lane_id() == 0) sts_u32(a_ncand, 0);
	__syncwarp();
}

// Sorted permutation of n 64-bit keys by a warp, for large n: see warp_sort_perm above.

template <bool HEAVY>
__device__ __noinline__ void sdp_match(ReadState &S, uint32_t q_bg, uint32_t q_ed, const uint8_t *q_str, const uint8_t *t_str, uint32_t t_len,
                                       uint32_t t_st, bool isForward)
{
	SdpArgs A; A.q_bg = q_bg; A.q_ed = q_ed; A.q_str = q_str; A.t_str = t_str; A.t_len = t_len; A.t_st = t_st; A.fwd = isForward;
	PH_BEGIN();
	__syncwarp();
	if (A.t_len < S_A_KEMR_L + 4) return;
	const uint32_t t_kmer_num = A.t_len - S_A_KEMR_L + 1;
	const uint32_t n_pos = (t_kmer_num > 4) ? (t_kmer_num - 4 + 3) / 4 : 0;       // i = 4, 8, ... < t_kmer_num
	if (n_pos == 0 || S.l_read < S_A_KEMR_L) return;
	if (n_pos > 511) { S.error = 3; return; }                                      // (windows are < 2000 bases everywhere)
	const uint32_t lo = A.q_bg, hi = DSB_MIN(A.q_ed, S.l_read - S_A_KEMR_L);       // read positions that have a 9-mer and lie in [q_bg, q_ed]
	if (lo > hi) return;
	MatchSmem *M = S.mt;
	const int lane = lane_id();
#ifdef DSB_PROF_SDP
	long long pt_ = clock64();
	#define SDP_PT(k) do { const long long n_ = clock64(); S.t_phase[k] += n_ - pt_; pt_ = n_; } while (0)
#else
	#define SDP_PT(k) do { } while (0)
#endif
	// (1) the scanned target 9-mers -> shared hash table (open addressing, equal 9-mers take slots of their own) + filter bits
	uint32_t tbits = 6;
	while ((1u << tbits) < 2 * n_pos) tbits++;
	const uint32_t tmask = (1u << tbits) - 1;
	const uint32_t fbits = tbits + 4;                                              // filter bits = 16 x table slots (<= 16384)
	const uint32_t a_tslot = smem_addr(M->tslot), a_bloom = smem_addr(M->bloom), a_ncand = smem_addr(&M->n_cand);
	for (uint32_t s = lane; s <= tmask; s += 32) sts_u32(a_tslot + 4 * s, TT_EMPTY);
	for (uint32_t s = lane; s < (1u << (fbits - 5)); s += 32) sts_u32(a_bloom + 4 * s, 0);
	if (lane == 0) { sts_u32(a_ncand, 0); sts_u32(smem_addr(&M->n_tmp), 0); }
	__syncwarp();
	for (uint32_t k = lane; k < n_pos; k += 32) {
		const int i = 4 + 4 * (int)k;
		const uint8_t *c_t_str = A.fwd ? (A.t_str + i) : (A.t_str + A.t_len - S_A_KEMR_L - i);
		uint32_t kmer = 0;
		#pragma unroll
		for (int b = 0; b < S_A_KEMR_L; b++) kmer = (kmer << 2) | c_t_str[b];
		const uint32_t f = tt_hash(kmer, fbits);
		reds_or(a_bloom + 4 * (f >> 5), 1u << (f & 31));
		uint32_t h = tt_hash(kmer, tbits);
		while (atoms_cas(a_tslot + 4 * h, TT_EMPTY, (kmer << 9) | k) != TT_EMPTY) h = (h + 1) & tmask;
	}
	__syncwarp();
	SDP_PT(0);
	// (2) stream the read range, 512 positions per round (the loads of the next round are in flight while this one is looked
	// up): filter bits of 16 consecutive 9-mers per lane, then the table for the few that pass.  A candidate is a (scanned
	// target index, read position) pair with equal 9-mers.
	const uintptr_t a_lo = (uintptr_t)(A.q_str + lo) & ~(uintptr_t)15;
	const int64_t p_first = (int64_t)(a_lo - (uintptr_t)A.q_str);                  // read position of byte 0 of the first 16-byte chunk (<= lo)
	// a short range (the gap between two anchors of a chain: a few hundred positions) would leave most lanes without a chunk
	// while the others look up 16 positions one after the other: there `share` = 2 or 4 lanes split a chunk (they load the same
	// 24 bytes -- one request) and a round covers 256 or 128 positions
	const uint32_t n_span = (uint32_t)((int64_t)hi - p_first + 1);
	const int share = n_span <= 128 ? 4 : n_span <= 256 ? 2 : 1, per_lane = 16 / share, j_first = (lane & (share - 1)) * per_lane;
	const int64_t round = 512 / share;
	const uint32_t j_mine = ((1u << per_lane) - 1) << j_first;
	uint4 a = make_uint4(0, 0, 0, 0); uint2 b = make_uint2(0, 0);
	{ const int64_t p0 = p_first + 16 * (lane / share); if (p0 <= (int64_t)hi) { a = __ldg((const uint4 *)(A.q_str + p0)); b = __ldg((const uint2 *)(A.q_str + p0 + 16)); } }
	for (int64_t pb = p_first; pb <= (int64_t)hi; pb += round) {
		const int64_t p0 = pb + 16 * (lane / share);
		const uint4 ca = a; const uint2 cb = b;
		{ const int64_t pn = p0 + round; if (pn <= (int64_t)hi) { a = __ldg((const uint4 *)(A.q_str + pn)); b = __ldg((const uint2 *)(A.q_str + pn + 16)); } }
		uint32_t maybe = 0;                                                        // bit j: the 9-mer at p0 + j passes the filter
		uint64_t bits = 0;
		if (p0 <= (int64_t)hi && p0 + 15 >= (int64_t)lo) {
			bits = ((uint64_t)pack4(ca.x) << 40) | ((uint64_t)pack4(ca.y) << 32) | ((uint64_t)pack4(ca.z) << 24) |
			       ((uint64_t)pack4(ca.w) << 16) | ((uint64_t)pack4(cb.x) << 8) | (uint64_t)pack4(cb.y);   // base t of the chunk at bits 47-2t, 46-2t
			#pragma unroll 4
			for (int j = j_first; j < j_first + per_lane; j++) {
				const uint32_t kmer = (uint32_t)(bits >> (30 - 2 * j)) & 0x3ffffu;
				const uint32_t f = tt_hash(kmer, fbits);
				maybe |= ((lds_u32_ro(a_bloom + 4 * (f >> 5)) >> (f & 31)) & 1u) << j;
			}
			const int j_min = (int)DSB_MAX((int64_t)lo - p0, (int64_t)0), j_max = (int)DSB_MIN((int64_t)hi - p0, (int64_t)15);
			maybe &= (0xffffu >> (15 - j_max)) & (0xffffu << j_min) & j_mine;
		}
		while (__any_sync(DSB_FULL, maybe != 0)) {                                  // one position per lane and turn
			if (lds_u32(a_ncand) + 32 * n_pos > CAND_CAP) sdp_flush_cand<HEAVY>(S, A);     // room for every scanned position on every lane
			if (maybe) {
				const int j = __ffs(maybe) - 1;
				maybe &= maybe - 1;
				const uint32_t p = (uint32_t)(p0 + j);
				const uint32_t kmer = (uint32_t)(bits >> (30 - 2 * j)) & 0x3ffffu;
				uint32_t h = tt_hash(kmer, tbits), slot;
				for (; (slot = lds_u32_ro(a_tslot + 4 * h)) != TT_EMPTY; h = (h + 1) & tmask)     // every scanned target position with this 9-mer
					if ((slot >> 9) == kmer) S.ws.cand[atoms_add(a_ncand, 1u)] = make_uint2(slot & 0x1ff, p);
			}
		}
	}
	SDP_PT(1);
	sdp_flush_cand<HEAVY>(S, A);
	SDP_PT(2);
	// (3) order: target scan order, then ascending read position (unique keys), appended to the match array
	const uint32_t total = lds_u32(smem_addr(&M->n_tmp));
	if (total) {
		if (S.n_sms + total > S.max_matches) { S.error = 2; return; }
		DevSms *dst = S.ws.sms + S.n_sms;
		const DevSms *src = S.ws.sms_tmp;
		const uint64_t *key = S.ws.sort_key[0];
		if (total <= 256) {
			for (uint32_t i0 = 0; i0 < total; i0 += 32) {
				const uint32_t i = i0 + lane;
				const uint64_t ki = (i < total) ? key[i] : 0;
				uint32_t r = 0;
				for (uint32_t j = 0; j < total; j++) r += (key[j] < ki) ? 1u : 0u;
				if (i < total) { const DevSms m = src[i]; dst[r].t_pos = m.t_pos; dst[r].q_pos = m.q_pos; dst[r].len = m.len; }
			}
		} else {
			const uint32_t *perm;
			if (HEAVY && total >= 1024) {                                              // k_score_heavy: the whole CTA orders the matches
				DpTeam *T = S.team;
				__syncwarp();
				if (lane == 0) { T->sort_n = total; T->sort_key[0] = S.ws.sort_key[0]; T->sort_key[1] = S.ws.sort_key[1]; T->sort_idx[0] = S.ws.sort_idx[0]; T->sort_idx[1] = S.ws.sort_idx[1]; T->cmd = 2; }
				__syncthreads();                                                       // job posted: the helper warps wake up
				perm = warp_sort_perm<128, TEAM_WARPS>(total, S.ws.sort_key[0], S.ws.sort_key[1], S.ws.sort_idx[0], S.ws.sort_idx[1], 0);
			} else
				perm = warp_sort_perm<128>(total, S.ws.sort_key[0], S.ws.sort_key[1], S.ws.sort_idx[0], S.ws.sort_idx[1]);
			for (uint32_t i = lane; i < total; i += 32) { const DevSms m = src[perm[i]]; dst[i].t_pos = m.t_pos; dst[i].q_pos = m.q_pos; dst[i].len = m.len; }
		}
		S.n_sms += total;
	}
	__syncwarp();
	PH_END(S, 3);
}

__device__ __forceinline__ void refwin_zero(ReadState &S, int nbytes)     // zero-initialised stack window (policy P1)
{
	__syncwarp();
	uint32_t *w = (uint32_t *)S.sm->refwin;
	for (int i = lane_id(); i < (nbytes + 3) / 4; i += 32) w[i] = 0;
	__syncwarp();
}

__device__ __forceinline__ int warp_max(int v)
{
	return __reduce_max_sync(DSB_FULL, v);
}
__device__ __forceinline__ DevSms load_sms(const DevSms *p)
{
	const uint4 v = *(const uint4 *)p;
	DevSms r; r.t_pos = v.x; r.q_pos = v.y; r.len = v.z; r.score = v.w;
	return r;
}

#define MAX_sms_overlap (6)
#define MAX_sms_overlap_middle (6)
// ---------------------------------------------------------------- sparse DP over 9-mer matches, one LANE per match
// The reference scores the matches of a window one after the other; each looks back over ALL earlier matches (walking
// backwards, stopping at its first `break`) and takes the best predecessor (cly.c:2480-2528, 2621-2652, 2765-2797).  Here
// a block of <= 32 consecutive matches is scored at once: lane j owns match first+j.  Earlier blocks are final, so every
// lane walks them in the same order (one broadcast load per predecessor serves 32 matches); the dependencies inside the
// block are resolved by a short sequential pass.  pass / break / candidate score are the reference's own expressions.
enum { DP_MIDDLE = 0, DP_RIGHT = 1, DP_LEFT = 2 };

// (KIND is a run-time value: one copy of the DP code keeps the scoring kernel inside the instruction cache -- with three
// template instances of every routine `no instruction` was the top stall, profiles/r1j)
__device__ __forceinline__ void dp_eval(const int KIND, const DevSms &c, const DevSms &p, bool &pass, bool &brk, bool &has, int &cand)
{
	has = false; brk = false; cand = 0;
	if (KIND == DP_LEFT) {
		const uint32_t min_pre_q = c.q_pos + c.len - MAX_sms_overlap + S_A_KEMR_L - 1;
		const uint32_t min_pre_t = c.t_pos + c.len - MAX_sms_overlap + S_A_KEMR_L - 1;
		pass = !(p.q_pos < min_pre_q) && !(p.t_pos < min_pre_t);
		if (!pass) return;
		if (min_pre_t + 600 < p.t_pos) { brk = true; return; }
		const int indel = p.q_pos - p.t_pos - (min_pre_q - min_pre_t);
		const int ABS_indel = DSB_ABS(indel);
		if (ABS_indel > 200) return;
		int new_score = p.score + c.len - (ABS_indel >> 3);
		if (min_pre_q + MAX_sms_overlap > p.q_pos || min_pre_t + MAX_sms_overlap > p.t_pos) {
			const int overlap_q = min_pre_q + MAX_sms_overlap - p.q_pos;
			const int overlap_t = min_pre_t + MAX_sms_overlap - p.t_pos;
			new_score -= DSB_MAX(overlap_q, overlap_t);
		}
		has = true; cand = new_score;
	} else {
		const uint32_t max_q = c.q_pos + MAX_sms_overlap;
		const uint32_t max_t = c.t_pos + MAX_sms_overlap;
		const int pre_q_ed = p.q_pos + p.len + S_A_KEMR_L - 1;
		const int pre_t_ed = p.t_pos + p.len + S_A_KEMR_L - 1;
		pass = !(pre_q_ed > max_q) && !(pre_t_ed > max_t);
		if (!pass) return;
		if (KIND == DP_RIGHT && (p.t_pos + 600 < max_t)) { brk = true; return; }
		const int indel = p.q_pos - p.t_pos - (max_q - max_t);
		const int ABS_indel = DSB_ABS(indel);
		if (ABS_indel > 200) return;
		int new_score = p.score + c.len - (ABS_indel >> 3);
		if (pre_q_ed > c.q_pos || pre_t_ed > c.t_pos) {
			const int overlap_q = pre_q_ed - c.q_pos;
			const int overlap_t = pre_t_ed - c.t_pos;
			new_score -= DSB_MAX(overlap_q, overlap_t);
		}
		has = true; cand = new_score;
	}
}

// phase (A) over the predecessors [lo, hi), walked downwards: best candidate per lane and whether the lane's walk hit its break.
// The predecessors are fetched 32 at a time (one coalesced 512-byte request, the next tile already in flight), parked in a
// shared-memory tile and read back by all lanes as broadcasts, highest index first.
__device__ __noinline__ void dp_range(const int KIND, const DevSms *sms, int lo, int hi, const DevSms &my, bool stopped, int &best, bool &brk_out, uint32_t a_tile)
{
	const int lane = lane_id();
	int top = hi - 1;                                    // tile = entries top, top-1, ..., top-31 (lane l holds entry top - l)
	uint4 cur = make_uint4(0, 0, 0, 0);
	if (top - lane >= lo) cur = *(const uint4 *)(sms + top - lane);
	while (top >= lo) {
		if (__all_sync(DSB_FULL, stopped)) break;                                  // (also: every lane is done with the previous tile)
		const uint32_t a_buf = a_tile;
		asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(a_buf + 16 * lane), "r"(cur.x), "r"(cur.y), "r"(cur.z), "r"(cur.w) : "memory");
		uint4 nxt = make_uint4(0, 0, 0, 0);
		if (top - 32 - lane >= lo) nxt = *(const uint4 *)(sms + top - 32 - lane);
		__syncwarp();
		const int n_here = min(32, top - lo + 1);
		if (!stopped) {
			#pragma unroll 2
			for (int k = 0; k < n_here; k++) {
				DevSms p;
				asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(p.t_pos), "=r"(p.q_pos), "=r"(p.len), "=r"(p.score) : "r"(a_buf + 16 * k) : "memory");
				bool pass, brk, has; int cand;
				dp_eval(KIND, my, p, pass, brk, has, cand);
				if (brk) { stopped = true; brk_out = true; break; }
				if (has) best = DSB_MAX(best, cand);
			}
		}
		cur = nxt; top -= 32;
	}
	__syncwarp();
}

// Phase (A) of a posted job by the whole CTA.  The predecessors first-1 .. 0 are cut into tiles of 32, nearest first; in round
// r warp w walks tile 32 r + w, then all warps combine the 32 tiles of the round IN ORDER (a lane's walk ends at its first
// break, later tiles do not count) -- every warp computes the same, so all leave the loop together once every lane has
// stopped.  (Round 1 gave each warp one contiguous 1/32 of the predecessors: the walks of a repeat-rich window end after a few
// thousand of them, so the two or three nearest warps did all the work -- 82 % of the kernel's stall samples were the other
// warps at the barrier.)  Returns the best candidate of the lane's match over the earlier blocks (INT_MIN: none).
__device__ __forceinline__ int dp_team_range(DpTeam *T, int w)
{
	const int lane = lane_id();
	const int first = (int)T->first, n_tiles = (first + 31) >> 5, kind = T->kind;
	const DevSms *sms = T->sms;
	const DevSms my = T->item[lane];
	bool stopped = T->stopped0[lane] != 0;
	int best = INT_MIN;
	const uint32_t a_buf = smem_addr(T->tile[w]);
	int t = w;
	uint4 cur = make_uint4(0, 0, 0, 0);
	if (t < n_tiles && first - 1 - 32 * t - lane >= 0) cur = *(const uint4 *)(sms + first - 1 - 32 * t - lane);
	for (int r = 0; 32 * r < n_tiles; r++, t += TEAM_WARPS) {
		uint4 nxt = make_uint4(0, 0, 0, 0);                  // my tile of the next round is in flight while this one is walked
		if (t + TEAM_WARPS < n_tiles && first - 1 - 32 * (t + TEAM_WARPS) - lane >= 0) nxt = *(const uint4 *)(sms + first - 1 - 32 * (t + TEAM_WARPS) - lane);
		int tb = INT_MIN; bool tbrk = false;
		if (t < n_tiles) {
			asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(a_buf + 16 * lane), "r"(cur.x), "r"(cur.y), "r"(cur.z), "r"(cur.w) : "memory");
			__syncwarp();
			const int n_here = min(32, first - 32 * t);
			if (!stopped) {
				#pragma unroll 2
				for (int k = 0; k < n_here; k++) {
					DevSms p;
					asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(p.t_pos), "=r"(p.q_pos), "=r"(p.len), "=r"(p.score) : "r"(a_buf + 16 * k) : "memory");
					bool pass, brk, has; int cand;
					dp_eval(kind, my, p, pass, brk, has, cand);
					if (brk) { tbrk = true; break; }
					if (has) tb = DSB_MAX(tb, cand);
				}
			}
			__syncwarp();
		}
		const int buf = r & 1;
		T->best[buf][w][lane] = tb; T->brk[buf][w][lane] = tbrk ? 1 : 0;
		__syncthreads();                                 // the tiles of the round are done
		if (!stopped)
			for (int ww = 0; ww < TEAM_WARPS; ww++) { best = DSB_MAX(best, T->best[buf][ww][lane]); if (T->brk[buf][ww][lane]) { stopped = true; break; } }
		cur = nxt;
		if (__all_sync(DSB_FULL, stopped)) break;        // (the same in every warp)
	}
	return best;
}

__device__ __forceinline__ void dp_team_helper_loop(DpTeam *T, int w)
{
	for (;;) {
		__syncthreads();                                 // job posted
		if (T->cmd < 0) break;
		if (T->cmd == 3) { flush_range(T->flush, T->flush_n, w, TEAM_WARPS); __syncthreads(); }
		else if (T->cmd == 2) (void)warp_sort_perm<128, TEAM_WARPS>(T->sort_n, T->sort_key[0], T->sort_key[1], T->sort_idx[0], T->sort_idx[1], w);
		else (void)dp_team_range(T, w);
	}
}

// scores of the matches [first, first + nb), nb <= 32; lane j holds match first+j in `my` and receives its score.
template <bool HEAVY>
__device__ __noinline__ int dp_block(const int KIND, const DevSms *sms, uint32_t first, uint32_t nb, DevSms my, DpTeam *team, uint32_t a_tile)
{
	const int lane = lane_id();
	const bool mine = (uint32_t)lane < nb;
	// the block's matches, parked in shared memory behind the predecessor tile: read back as broadcasts in (0) and (B)
	const uint32_t a_blk = a_tile + 512;
	__syncwarp();
	asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(a_blk + 16 * lane), "r"(my.t_pos), "r"(my.q_pos), "r"(my.len), "r"(0u) : "memory");
	__syncwarp();
	// (0) inside the block, positions only: does my walk stop (break) before it leaves the block, and at which lane?
	int stop_at = -1;                                    // highest k < lane whose predecessor test says `break`
	if (KIND != DP_MIDDLE) {
		for (int k = (int)nb - 2; k >= 0; k--) {
			DevSms p;
			asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(p.t_pos), "=r"(p.q_pos), "=r"(p.len), "=r"(p.score) : "r"(a_blk + 16 * k) : "memory");
			if (mine && k < lane && stop_at < 0) { bool pass, brk, has; int cand; dp_eval(KIND, my, p, pass, brk, has, cand); if (brk) stop_at = k; }
		}
	}
	// (A) predecessors in earlier blocks: all lanes walk first-1 .. 0 together, each with its own stop
	int best = (int)my.len;
	const bool stopped = !mine || stop_at >= 0;
	if (HEAVY && first >= TEAM_MIN_FIRST) {
		team->item[lane] = my; team->stopped0[lane] = stopped ? 1 : 0;
		if (lane == 0) { team->kind = KIND; team->first = first; team->nb = nb; team->sms = sms; team->cmd = 1; }
		__syncthreads();                                 // job posted: the helper warps wake up
		const int tb = dp_team_range(team, 0);
		if (!stopped) best = DSB_MAX(best, tb);
		__syncwarp();
	} else {
		bool brk = false;
		dp_range(KIND, sms, 0, (int)first, my, stopped, best, brk, a_tile);
	}
	// (B) predecessors inside the block, in order: match j needs the final scores of the matches before it
	int my_score = best;
	for (uint32_t j = 1; j < nb; j++) {
		DevSms c;
		asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c.t_pos), "=r"(c.q_pos), "=r"(c.len), "=r"(c.score) : "r"(a_blk + 16 * j) : "memory");
		const int c_stop = __shfl_sync(DSB_FULL, stop_at, j);
		int cand_max = INT_MIN;
		if ((uint32_t)lane < j && lane > c_stop) {
			DevSms p = my; p.score = (uint32_t)my_score;
			bool pass, brk, has; int cand;
			dp_eval(KIND, c, p, pass, brk, has, cand);
			if (has) cand_max = cand;
		}
		cand_max = __reduce_max_sync(DSB_FULL, cand_max);
		if ((uint32_t)lane == j && cand_max > my_score) my_score = cand_max;
	}
	return my_score;
}
template <bool HEAVY>
__device__ __noinline__ int sdp_middle_M2(ReadState &S, int32_t c_a, const uint8_t *q_str)
{   // cly.c:2444-2530
	const DevIndex &ix = *S.ix;
	int score = 10000;
	const DevAnchor *A = S.ws.anc;
	const uint64_t t_offset = __ldg(ix.ref_info + A[c_a].ref_ID).y;
	int32_t pre_a = -1;
	while (c_a != -1) {
		const DevAnchor ca = A[c_a];
		pre_a = ca.pre;
		if (pre_a != -1) {
			const DevAnchor pa = A[pre_a];
			const int pre_mch = pa.mtch_len;
			const int pre_refoffset = pa.ref_offset - 3;
			const int total_ref_len = ca.ref_offset - (pre_refoffset + pre_mch) + 3;
			S.n_sms = 0;
			DevSms *base = S.ws.sms;
			base[0].score = score; base[0].q_pos = pa.index_in_read; base[0].t_pos = pa.ref_offset; base[0].len = pa.mtch_len - S_A_KEMR_L + 1;
			S.n_sms = 1;
			if (total_ref_len > 12) {
				if (!(total_ref_len < 2000)) { S.error = 3; return 0; }         // xassert(total_ref_len < 2000) aborts the reference (cly.c:2473)
				refwin_zero(S, 2128);
				const uint64_t ref_offset = pre_refoffset + t_offset + pre_mch;
				CNT_GETREF(S, total_ref_len); get_ref_coop(ix, S.sm->refwin, ref_offset, total_ref_len);
				sdp_match<HEAVY>(S, pa.index_in_read + pre_mch - 8, ca.index_in_read - 1, q_str, S.sm->refwin, total_ref_len, pre_refoffset + pre_mch, true);
				if (S.error) return 0;
				if (!HEAVY && S.n_sms > DEFER_SMS) { S.error = ERR_DEFER; return 0; }
			}
			if (!sms_push(S, ca.ref_offset, ca.index_in_read, ca.mtch_len - S_A_KEMR_L + 1)) return 0;
			if (S.n_sms > 1) {
				__syncwarp();
				for (uint32_t first = 1; first < S.n_sms; first += 32) {
					const uint32_t nb = DSB_MIN(32u, S.n_sms - first);
					DevSms my; my.t_pos = my.q_pos = my.len = my.score = 0;
					if ((uint32_t)lane_id() < nb) my = load_sms(base + first + lane_id());
					const int sc = dp_block<HEAVY>(DP_MIDDLE, base, first, nb, my, S.team, smem_addr(S.mt->tslot));
					if ((uint32_t)lane_id() < nb) base[first + lane_id()].score = sc;
					__syncwarp();
					score = DSB_MAX(warp_max(((uint32_t)lane_id() < nb) ? sc : INT_MIN), score);
				}
			}
		} else
			score += ca.mtch_len - S_A_KEMR_L + 1;
		c_a = pre_a;
	}
	return score - 10000;
}

template <bool HEAVY>
__device__ __noinline__ int sdp_right_M2(ReadState &S, const uint8_t *q_str, int chain_ID, uint32_t l_read, int score_ori)
{   // cly.c:2532-2677
	const DevIndex &ix = *S.ix;
	DevChain *c_st = S.ws.chain;
	DevSms *sms = S.ws.sms;
	score_ori += 10000;
	int total_max_score = score_ori;
	int max_sms_id = 0;
	int combined = 0;
	refwin_zero(S, 1064);
	{ const DevChain ch = c_st[chain_ID]; sms[0].score = score_ori; sms[0].q_pos = ch.q_ed; sms[0].t_pos = ch.t_ed; sms[0].len = 1 - S_A_KEMR_L; }
	S.n_sms = 1;
	uint32_t current_sms = 1;
	const ulonglong2 ri = __ldg(ix.ref_info + c_st[chain_ID].ref_ID);
	const uint64_t t_offset_global = ri.y, t_length = ri.x;
	uint32_t c_t_offset = c_st[chain_ID].t_ed - 3;
	uint32_t best_t = c_st[chain_ID].t_ed;               // sms[max_sms_id].t_pos
	int last_search = 0;
	while (1) {
		if (S.n_sms == current_sms) {
			const DevChain ch = c_st[chain_ID];
			const uint32_t next_step = (uint32_t)(t_length - c_t_offset);
			if (next_step < MIN_SCORE_MEM) break;
			uint32_t max_search_ref;
			if (l_read - ch.q_ed < 600) {
				if (last_search == 1) break;
				last_search = 1;
				max_search_ref = l_read - ch.q_ed + 60;
			} else
				max_search_ref = (uint32_t)(t_length - c_t_offset);
			max_search_ref = DSB_MIN(600, max_search_ref);
			__syncwarp();
			CNT_GETREF(S, max_search_ref + OVER_SEARCH_M2); get_ref_coop(ix, S.sm->refwin, c_t_offset + t_offset_global, max_search_ref + OVER_SEARCH_M2);
			int search_q_ed = (int)sms[max_sms_id].q_pos + 1000;
			search_q_ed = DSB_MIN(search_q_ed, l_read);
			const int search_q_st = DSB_MAX(search_q_ed - 2000, ch.q_st - 8);
			sdp_match<HEAVY>(S, search_q_st, search_q_ed, q_str, S.sm->refwin, max_search_ref, c_t_offset, true);
			if (S.error) return 0;
			if (!HEAVY && S.n_sms > DEFER_SMS) { S.error = ERR_DEFER; return 0; }
			c_t_offset += max_search_ref - S_A_KEMR_L - 3;
			if (S.n_sms == current_sms) break;
			if (sms[current_sms].t_pos > sms[max_sms_id].t_pos + 1000) break;
		}
		// score the pending matches in blocks of 32 (dp_block), then replay the reference's per-match control flow
		const uint32_t first = current_sms, nb = DSB_MIN(32u, S.n_sms - current_sms);
		DevSms my; my.t_pos = my.q_pos = my.len = my.score = 0;
		if ((uint32_t)lane_id() < nb) my = load_sms(sms + first + lane_id());
		const int my_score = dp_block<HEAVY>(DP_RIGHT, sms, first, nb, my, S.team, smem_addr(S.mt->tslot));
		if ((uint32_t)lane_id() < nb) sms[first + lane_id()].score = my_score;
		__syncwarp();
		// combine_chain returns 0 at once when the bucket of the match's diagonal is empty (cly.c:1771)
		const bool need = (uint32_t)lane_id() < nb && my.len >= 8 && S.ws.sc_hash[(my.t_pos - my.q_pos) & 0xff].next != 0;
		const uint32_t cm = __ballot_sync(DSB_FULL, need);
		bool restarted = false, done = false;
		for (uint32_t j = 0; j < nb; j++) {
			DevSms c_sms;
			c_sms.t_pos = __shfl_sync(DSB_FULL, my.t_pos, j); c_sms.q_pos = __shfl_sync(DSB_FULL, my.q_pos, j); c_sms.len = __shfl_sync(DSB_FULL, my.len, j);
			const int max_score = __shfl_sync(DSB_FULL, my_score, j);
			current_sms = first + j + 1;
			if (((cm >> j) & 1) && combine_chain(S, chain_ID, c_sms.t_pos - c_sms.q_pos, 0, c_sms.q_pos, &combined) == 1) {
				total_max_score = DSB_MAX(score_ori, max_score) - c_sms.len + sdp_middle_M2<HEAVY>(S, c_st[combined].cur, q_str);
				if (S.error) return 0;
				score_ori = total_max_score;
				max_sms_id = 0;
				const DevChain ch = c_st[chain_ID];
				__syncwarp();
				if (lane_id() == 0) { sms[0].score = total_max_score; sms[0].q_pos = ch.q_ed; sms[0].t_pos = ch.t_ed; sms[0].len = -S_A_KEMR_L; }
				__syncwarp();
				best_t = ch.t_ed;
				S.n_sms = 1;
				current_sms = 1;
				c_t_offset = ch.t_ed;
				restarted = true;
				break;
			}
			if (total_max_score < max_score) { total_max_score = max_score; max_sms_id = first + j; best_t = c_sms.t_pos; }
			if (c_sms.t_pos > best_t + 1000) { done = true; break; }
		}
		if (restarted) continue;
		if (done) break;
	}
	__syncwarp();
	c_st[chain_ID].q_ed = sms[max_sms_id].q_pos + sms[max_sms_id].len + S_A_KEMR_L;
	c_st[chain_ID].t_ed = sms[max_sms_id].t_pos + sms[max_sms_id].len + S_A_KEMR_L;
	return total_max_score - 10000;
}

template <bool HEAVY>
__device__ __noinline__ int sdp_left_M2(ReadState &S, const uint8_t *q_str, int chain_ID, int score_ori)
{   // cly.c:2679-2819
	const DevIndex &ix = *S.ix;
	DevChain *c_st = S.ws.chain;
	DevSms *sms = S.ws.sms;
	score_ori += 10000;
	int total_max_score = score_ori;
	int max_sms_id = 0;
	int combined = 0;
	refwin_zero(S, 1064);
	{ const DevChain ch = c_st[chain_ID]; sms[0].score = score_ori; sms[0].q_pos = ch.q_st; sms[0].t_pos = ch.t_st; }
	S.n_sms = 1;
	uint32_t current_sms = 1;
	const uint64_t t_offset_global = __ldg(ix.ref_info + c_st[chain_ID].ref_ID).y;
	uint32_t c_t_offset = c_st[chain_ID].t_st + 3;
	uint32_t best_t = c_st[chain_ID].t_st;               // sms[max_sms_id].t_pos
	int last_search = 0;
	while (1) {
		if (S.n_sms == current_sms) {
			const DevChain ch = c_st[chain_ID];
			const uint32_t next_step = c_t_offset;
			if (next_step < MIN_SCORE_MEM) break;
			uint32_t max_search_ref;
			if (ch.q_st < 600) {
				if (last_search == 1) break;
				last_search = 1;
				max_search_ref = ch.q_st + 60;
			} else
				max_search_ref = c_t_offset;
			max_search_ref = DSB_MIN(600, max_search_ref);
			__syncwarp();
			if (t_offset_global == 0 && c_t_offset < OVER_SEARCH_M2 + max_search_ref)
				{ CNT_GETREF(S, max_search_ref); get_ref_coop(ix, S.sm->refwin, (int64_t)(c_t_offset + t_offset_global - max_search_ref), max_search_ref); }
			else
				{ CNT_GETREF(S, max_search_ref + OVER_SEARCH_M2); get_ref_coop(ix, S.sm->refwin, (int64_t)(c_t_offset + t_offset_global - max_search_ref - OVER_SEARCH_M2), max_search_ref + OVER_SEARCH_M2); }
			int search_q_st = (int)sms[max_sms_id].q_pos - 1000;
			search_q_st = DSB_MAX(search_q_st, 0);
			const int search_q_ed = DSB_MIN(search_q_st + 2000, ch.q_st - 1);
			sdp_match<HEAVY>(S, search_q_st, search_q_ed, q_str, S.sm->refwin + OVER_SEARCH_M2, max_search_ref, c_t_offset - max_search_ref, false);
			if (S.error) return 0;
			if (!HEAVY && S.n_sms > DEFER_SMS) { S.error = ERR_DEFER; return 0; }
			c_t_offset = c_t_offset - max_search_ref + S_A_KEMR_L + 3;
			if (S.n_sms == current_sms) break;
			if (sms[current_sms].t_pos + 1000 < sms[max_sms_id].t_pos) break;
		}
		// score the pending matches in blocks of 32 (dp_block), then replay the reference's per-match control flow
		const uint32_t first = current_sms, nb = DSB_MIN(32u, S.n_sms - current_sms);
		DevSms my; my.t_pos = my.q_pos = my.len = my.score = 0;
		if ((uint32_t)lane_id() < nb) my = load_sms(sms + first + lane_id());
		const int my_score = dp_block<HEAVY>(DP_LEFT, sms, first, nb, my, S.team, smem_addr(S.mt->tslot));
		if ((uint32_t)lane_id() < nb) sms[first + lane_id()].score = my_score;
		__syncwarp();
		const bool need = (uint32_t)lane_id() < nb && my.len >= 8 && S.ws.sc_hash[(my.t_pos - my.q_pos) & 0xff].next != 0;
		const uint32_t cm = __ballot_sync(DSB_FULL, need);
		bool restarted = false, done = false;
		for (uint32_t j = 0; j < nb; j++) {
			DevSms c_sms;
			c_sms.t_pos = __shfl_sync(DSB_FULL, my.t_pos, j); c_sms.q_pos = __shfl_sync(DSB_FULL, my.q_pos, j); c_sms.len = __shfl_sync(DSB_FULL, my.len, j);
			const int max_score = __shfl_sync(DSB_FULL, my_score, j);
			current_sms = first + j + 1;
			if (((cm >> j) & 1) && combine_chain(S, chain_ID, c_sms.t_pos - c_sms.q_pos, 1, c_sms.q_pos + c_sms.len, &combined) == 1) {
				total_max_score = DSB_MAX(score_ori, max_score) - c_sms.len + sdp_middle_M2<HEAVY>(S, c_st[combined].cur, q_str);
				if (S.error) return 0;
				score_ori = total_max_score;
				max_sms_id = 0;
				const DevChain ch = c_st[chain_ID];
				__syncwarp();
				if (lane_id() == 0) { sms[0].score = total_max_score; sms[0].q_pos = ch.q_st; sms[0].t_pos = ch.t_st; }
				__syncwarp();
				best_t = ch.t_st;
				S.n_sms = 1;
				current_sms = 1;
				c_t_offset = ch.t_st;
				restarted = true;
				break;
			}
			if (total_max_score < max_score) { total_max_score = max_score; max_sms_id = first + j; best_t = c_sms.t_pos; }
			if (c_sms.t_pos + 1000 < best_t) { done = true; break; }
		}
		if (restarted) continue;
		if (done) break;
	}
	__syncwarp();
	c_st[chain_ID].q_st = sms[max_sms_id].q_pos;
	c_st[chain_ID].t_st = sms[max_sms_id].t_pos;
	return total_max_score - 10000;
}

// delete_small_score_rst up to (not including) the max_read_l-dependent filter (cly.c:2883-2957)
template <bool HEAVY>
__device__ __noinline__ void score_and_merge(ReadState &S, const SearchDir *search_dir, uint32_t l_read)
{
	if (S.n_hit == 0) return;
	DevChain *C = S.ws.chain;
	if (S.n_hit > 200) {
		uint32_t rst_num = 200;
		for (; rst_num < S.n_hit && C[rst_num].sum_score > 50; rst_num++);
		S.n_hit = rst_num;
	}
	S.n_hit = DSB_MIN(400u, S.n_hit);
	sc_hash_idx(S);
	// get_score_M2 (cly.c:2821-2849)
	S.l_read = l_read;                                             // (build_hash_table_M2, cly.c:2173-2224, is not needed: see sdp_match)
	for (uint32_t i = 0; i < S.n_hit; i++) {
		if (C[i].sum_score == 0) continue;
		const uint32_t dir = C[i].direction;
		const SearchDir *c_sd = ((search_dir->direction == dir) ? 0 : 1) + search_dir;
		int score; { PH_BEGIN(); score = sdp_middle_M2<HEAVY>(S, C[i].cur, c_sd->bin_read); PH_END(S, 4); }
		if (S.error) return;
		{ PH_BEGIN(); score = sdp_right_M2<HEAVY>(S, c_sd->bin_read, (int)i, l_read, score); PH_END(S, 5); }
		if (S.error) return;
		{ PH_BEGIN(); score = sdp_left_M2<HEAVY>(S, c_sd->bin_read, (int)i, score); PH_END(S, 6); }
		if (S.error) return;
		C[i].sum_score = score;
	}
	if (S.n_hit > 1) {
		// qsort by chain_cmp_by_pos (cly.c:2853-2870): a consistent order (ref_ID, t_st ascending, sum_score descending), so the
		// reference's merge sort is THE stable sort by that key -- two stable passes, least significant field first
		uint64_t *key = S.ws.sort_key[0];
		__syncwarp();
		for (uint32_t i = lane_id(); i < S.n_hit; i += 32) key[i] = (uint32_t)~C[i].sum_score;
		warp_stable_sort_by_key(C, S.ws.chain_tmp, S.n_hit, key);
		for (uint32_t i = lane_id(); i < S.n_hit; i += 32) key[i] = ((uint64_t)C[i].ref_ID << 32) | C[i].t_st;
		warp_stable_sort_by_key(C, S.ws.chain_tmp, S.n_hit, key);
	}
	const int n = (int)S.n_hit;
	for (int ci = 0; ci < n - 1; ci++) {
		if (C[ci].sum_score == 0) continue;
		DevChain c_c = C[ci];
		for (int ni = ci + 1; ni < n; ni++) {
			DevChain next_c = C[ni];
			if (c_c.ref_ID == next_c.ref_ID) {
				if (c_c.direction != next_c.direction) continue;
				if (next_c.sum_score == 0) continue;
				if (next_c.t_st < c_c.t_st + 5 && next_c.q_st < c_c.q_st + 5 && next_c.sum_score < c_c.sum_score + 5) {
					next_c.sum_score = 0; next_c.q_ed = next_c.q_st; next_c.t_ed = next_c.t_st;
					C[ni] = next_c;
					continue;
				}
				const int dis_t = next_c.t_st - c_c.t_ed;
				const int dis_q = next_c.q_st - c_c.q_ed;
				const int dis_t_q = DSB_ABS(dis_t - dis_q);
				if ((dis_t > -20 && dis_t < 1000 && dis_q > -20 && dis_q < 1000) && dis_t_q < 200) {
					c_c.t_ed = DSB_MAX(c_c.t_ed, next_c.t_ed);
					c_c.q_ed = DSB_MAX(c_c.q_ed, next_c.q_ed);
					c_c.sum_score += next_c.sum_score;
					next_c.sum_score = 0; next_c.q_ed = next_c.q_st; next_c.t_ed = next_c.t_st;
					C[ni] = next_c;
				}
			} else
				break;
		}
		C[ci] = c_c;
	}
}

// ---------------------------------------------------------------- classify_seq (cly.c:3064-3132) up to the class filter
#define MIN_READ_LEN 40
// ================================================================ the phases of classify_seq (cly.c:3064-3132)
// classify_seq is cut at its decision points into three kinds of work, each run by its own kernel over a list of reads
// (so that all warps of the GPU execute the same code at the same time):
//   seed   fast_classify / slow_classify of the strand(s)            -> anchors appended to the read's anchor vector
//   chain  resolve_tree + the run_slow_mode decisions (cly.c:3101-3127) -> next list: slow pass 0, slow pass 1 or scoring
//   score  delete_small_score_rst up to the class filter (cly.c:2883-2957) -> pre-filter hits for the finalize kernel
__device__ __forceinline__ bool setup_dirs(const ClassifyParams &P, uint32_t r, uint32_t read_len, SearchDir sd[2])
{
	const uint8_t *bin_F = P.bin + P.bin_off[r] + DSB_GUARD;
	for (int s = 0; s < 2; s++) {
		sd[s].seed_v = P.seeds[s] + P.seed_off[r];
		sd[s].l_seed_v = P.n_seeds[s][r];
		sd[s].bin_read = s ? bin_F + read_len : bin_F;
		sd[s].direction = s ? DSB_REVERSE : DSB_FORWARD;
		sd[s].total_score = P.total_score[s][r];
	}
	if (sd[0].total_score < sd[1].total_score) { SearchDir t = sd[0]; sd[0] = sd[1]; sd[1] = t; }   // cly.c:1261-1266
	return ((sd[0].total_score - sd[1].total_score) <= (sd[0].total_score >> 3));                   // both_direction, cly.c:3095
}

__device__ __forceinline__ void read_begin(ReadState &S)
{
	S.n_anc = 0; S.n_hit = 0; S.n_sms = 0; S.fast_classify = 1; S.error = 0; S.sp_l = 0;
	S.c_prefix = S.c_occ = S.c_locate = S.c_getref = S.c_getref_bytes = 0;
	for (int k = 0; k < 8; k++) S.t_phase[k] = 0;
}

__device__ __forceinline__ void read_end(const ClassifyParams &P, ReadState &S, uint32_t r, long long t0)
{
	S.t_phase[7] = clock64() - t0;
	if (P.prof && lane_id() < 8) P.prof[(uint64_t)r * 8 + lane_id()] += (uint32_t)(S.t_phase[lane_id()] >> 10);
	__syncwarp();
}

__device__ __forceinline__ void list_push(const ClassifyParams &P, int list, uint32_t r)
{
	if (lane_id() == 0) { const uint32_t i = atomicAdd(P.ctl + CTL_LIST_N + list, 1u); P.list[list][i] = r; }
}

// a read that ends without hits (or with an error): its result record is final
__device__ __forceinline__ void write_empty_result(const ClassifyParams &P, uint32_t r, uint32_t read_len, uint32_t n_anchor, uint32_t fast, int error)
{
	if (lane_id() == 0) {
		dsb_read_result out;
		out.hit_off = 0; out.n_hit = 0; out.n_anchor = n_anchor; out.fast_classify = (uint8_t)fast; out.entered_final = 0; out.error = (uint16_t)error; out.read_len = read_len;
		P.rr[r] = out;
		if (error) atomicAdd(P.counters + DSB_CNT_N_ERRORS, 1ull);
	}
}

// Chaining phase of one read after a seeding pass: first the anchors of the pass are gathered from the per-seed staging lists
// into the read's anchor vector (gather_strand, dsb_seed.cuh: the order in which fast_classify / slow_classify push them),
// then resolve_tree and the decisions of classify_seq (cly.c:3098-3127); a read that goes on to a slow pass gets its seed
// tasks appended to that pass's list.
__device__ void phase_chain(const ClassifyParams &P, ReadState &S, uint32_t r, int pass)
{
	const long long t0 = clock64();
	const int lane = lane_id();
	const uint32_t read_len = (uint32_t)(P.read_off[r + 1] - P.read_off[r]);
	ReadWork w;
	w.anc_off = w.n_anc = w.chain_off = w.n_chain = 0; w.error = 0; w.fast_classify = 1; w.pad = 0;
	if (read_len < MIN_READ_LEN) {                                     // cly.c:3089: untouched (unmapped) result
		if (pass == PASS_FAST) { if (lane == 0) P.work[r] = w; write_empty_result(P, r, read_len, 0, 1, 0); }
		return;
	}
	read_begin(S);
	if (pass != PASS_FAST) { w = P.work[r]; w.fast_classify = 0; }     // slow_classify: results->fast_classify = false (cly.c:1610)
	if (w.error) { write_empty_result(P, r, read_len, w.n_anc, w.fast_classify, w.error); return; }
	SearchDir sd[2];
	const bool both_direction = setup_dirs(P, r, read_len, sd);
	const int super_repeat = 0;                                        // always 0 in the reference (cly.c:849-888,1545)
	{	// gather: fast = strand pass 0 (+ 1 if both_direction), slow 0 = strand pass 0 from scratch (cly.c:3116), slow 1 = strand
		// pass 1 appended to the (re-ordered) anchors of slow 0 (cly.c:3121-3125)
		const int d0 = (pass == PASS_SLOW1) ? 1 : 0, d1 = (pass == PASS_FAST && both_direction) ? 1 : d0;
		const uint32_t n_old = (pass == PASS_SLOW1) ? w.n_anc : 0u;
		GatherSum G; G.n_anchor = n_old; G.c_prefix = G.c_occ = G.c_locate = G.c_getref = G.c_getref_bytes = 0; G.error = 0;
		for (int d = d0; d <= d1; d++) gather_strand(P, r, sd[d].direction == DSB_FORWARD ? 0u : 1u, pass, nullptr, G, false);
		const uint32_t total = G.n_anchor;
		int err = __reduce_max_sync(DSB_FULL, G.error);
		uint32_t off = 0;
		if (total && !err) {
			if (lane == 0) off = atomicAdd(P.ctl + CTL_ANC_CURSOR, total);
			off = __shfl_sync(DSB_FULL, off, 0);
			if ((uint64_t)off + total > P.anc_pool_cap) { err = 1; if (lane == 0) atomicOr(P.ctl + CTL_OVERFLOW, OVF_ANCHORS); }
			else if (total > S.max_anchors) err = 1;
		}
		if (err) {
			w.n_anc = total; w.error = (uint16_t)err;
			if (lane == 0) P.work[r] = w;
			write_empty_result(P, r, read_len, total, w.fast_classify, err);
			return;
		}
		DevAnchor *dst = P.anc_pool + off;
		if (n_old) {
			const uint64_t *src = (const uint64_t *)(P.anc_pool + w.anc_off); uint64_t *d64 = (uint64_t *)dst;
			for (uint32_t i = lane; i < n_old * 3; i += 32) d64[i] = src[i];
		}
		GatherSum G2 = G; G2.n_anchor = n_old;
		for (int d = d0; d <= d1; d++) gather_strand(P, r, sd[d].direction == DSB_FORWARD ? 0u : 1u, pass, dst, G2, true);
		__syncwarp();
		w.anc_off = off; w.n_anc = total;
		const uint32_t c_prefix = __reduce_add_sync(DSB_FULL, G.c_prefix), c_occ = __reduce_add_sync(DSB_FULL, G.c_occ), c_locate = __reduce_add_sync(DSB_FULL, G.c_locate);
		const uint32_t c_getref = __reduce_add_sync(DSB_FULL, G.c_getref), c_getref_bytes = __reduce_add_sync(DSB_FULL, G.c_getref_bytes);
		if (lane == 0) {
			unsigned long long *C = P.counters;
			atomicAdd(C + DSB_CNT_N_PREFIX, (unsigned long long)c_prefix);
			atomicAdd(C + DSB_CNT_N_OCC, (unsigned long long)c_occ);
			atomicAdd(C + DSB_CNT_N_LOCATE, (unsigned long long)c_locate);
			atomicAdd(C + DSB_CNT_N_GETREF, (unsigned long long)c_getref);
			atomicAdd(C + DSB_CNT_N_GETREF_BYTES, (unsigned long long)c_getref_bytes);
		}
	}
	S.ws.anc = P.anc_pool + w.anc_off; S.n_anc = w.n_anc;
	{ PH_BEGIN(); resolve_tree(S); PH_END(S, 1); }
	int next;                                                           // -1: finished without hits
	if (pass == PASS_FAST) {                                            // cly.c:3101-3112
		bool run_slow_mode = false;
		if (S.n_hit <= 0) run_slow_mode = true;
		else if (S.ws.chain[0].anchor_number < 5 && super_repeat < 3) {
			run_slow_mode = true;
			if (read_len <= 300 && S.ws.chain[0].sum_score > 200) run_slow_mode = false;
		}
		next = run_slow_mode ? LIST_SLOW0 : LIST_SCORE;
	} else if (pass == PASS_SLOW0)                                      // cly.c:3119-3121
		next = (both_direction || S.n_hit <= 0 || (S.ws.chain[0].anchor_number < 5 && super_repeat < 3)) ? LIST_SLOW1 : LIST_SCORE;
	else
		next = S.n_hit ? LIST_SCORE : -1;
	if (next == LIST_SCORE) {
		uint32_t off = 0;
		if (lane == 0) off = atomicAdd(P.ctl + CTL_CHAIN_CURSOR, S.n_hit);
		off = __shfl_sync(DSB_FULL, off, 0);
		if ((uint64_t)off + S.n_hit > P.chain_pool_cap) {
			if (lane == 0) atomicOr(P.ctl + CTL_OVERFLOW, OVF_CHAINS);
			write_empty_result(P, r, read_len, w.n_anc, w.fast_classify, 1); read_end(P, S, r, t0); return;
		}
		const uint32_t *src = (const uint32_t *)S.ws.chain; uint32_t *dst = (uint32_t *)(P.chain_pool + off);
		for (uint32_t i = lane; i < S.n_hit * (uint32_t)(sizeof(DevChain) / 4); i += 32) dst[i] = src[i];
		w.chain_off = off; w.n_chain = S.n_hit;
	} else if (next == LIST_SLOW0 || next == LIST_SLOW1) {
		const int d = (next == LIST_SLOW0) ? 0 : 1;
		if (!slow_tasks_append(P, r, sd[d].direction == DSB_FORWARD ? 0u : 1u, next == LIST_SLOW0 ? PASS_SLOW0 : PASS_SLOW1)) {
			write_empty_result(P, r, read_len, w.n_anc, 0, 1); read_end(P, S, r, t0); return;
		}
	}
	if (lane == 0) P.work[r] = w;
	if (next >= 0) list_push(P, next, r);
	else write_empty_result(P, r, read_len, w.n_anc, w.fast_classify, 0);
	read_end(P, S, r, t0);
}

template <bool HEAVY>
__device__ void phase_score(const ClassifyParams &P, ReadState &S, uint32_t r)
{
	const long long t0 = clock64();
	const uint32_t read_len = (uint32_t)(P.read_off[r + 1] - P.read_off[r]);
	read_begin(S);
	const ReadWork w = P.work[r];
	SearchDir sd[2];
	setup_dirs(P, r, read_len, sd);
	S.ws.anc = P.anc_pool + w.anc_off; S.n_anc = w.n_anc;
	{
		const uint32_t *src = (const uint32_t *)(P.chain_pool + w.chain_off); uint32_t *dst = (uint32_t *)S.ws.chain;
		for (uint32_t i = lane_id(); i < w.n_chain * (uint32_t)(sizeof(DevChain) / 4); i += 32) dst[i] = src[i];
		__syncwarp();
		S.n_hit = w.n_chain;
	}
	dsb_read_result out;
	out.hit_off = 0; out.n_hit = 0; out.n_anchor = w.n_anc; out.fast_classify = w.fast_classify; out.entered_final = 1; out.error = 0; out.read_len = read_len;
	score_and_merge<HEAVY>(S, sd, read_len);
	if (S.error == ERR_DEFER) {                          // too heavy for one warp: k_score_heavy starts over from the pool chains
		list_push(P, LIST_SCORE_HEAVY, r);
		read_end(P, S, r, t0);
		return;
	}
	if (S.error) { out.error = (uint16_t)S.error; out.entered_final = 0; S.n_hit = 0; }
	// hand the pre-filter chains to the finalize kernel: reserve 2*n slots (second half = merge-sort scratch)
	unsigned long long off = 0;
	if (S.n_hit) {
		if (lane_id() == 0) off = atomicAdd(P.hits_cursor, (unsigned long long)(2 * S.n_hit));
		off = __shfl_sync(DSB_FULL, off, 0);
		if (off + 2ull * S.n_hit > P.hits_cap) { out.error = 4; out.entered_final = 0; S.n_hit = 0; if (lane_id() == 0) atomicOr(P.ctl + CTL_OVERFLOW, OVF_HITS); }
	}
	out.hit_off = off; out.n_hit = S.n_hit;
	for (uint32_t i = lane_id(); i < S.n_hit; i += 32) {
		const DevChain c = S.ws.chain[i];
		dsb_hit h;
		h.ref_ID = c.ref_ID; h.t_st = c.t_st; h.t_ed = c.t_ed; h.q_st = c.q_st; h.q_ed = c.q_ed;
		h.sum_score = c.sum_score; h.indel = c.indel; h.direction = c.direction; h.primary = 0; h.pri_index = 0; h.pad = 0;
		P.hits[off + i] = h;
	}
	if (lane_id() == 0) {
		P.rr[r] = out;
		unsigned long long *C = P.counters;
		if (out.entered_final) {                          // Classify_buff_pool.max_read_l bookkeeping (cly.c:2958), resolved in k_finalize
			atomicMax(C + DSB_CNT_MAX_READ_L, (unsigned long long)read_len);
			if (read_len >= 510) atomicMin(C + DSB_CNT_FIRST_LONG, (unsigned long long)r);
		}
		if (out.error) atomicAdd(C + DSB_CNT_N_ERRORS, 1ull);
		atomicAdd(C + DSB_CNT_N_GETREF_SCORE, (unsigned long long)S.c_getref);
		atomicAdd(C + DSB_CNT_N_GETREF_BYTES_SCORE, (unsigned long long)S.c_getref_bytes);
	}
	read_end(P, S, r, t0);
}
