// dsb_device.cuh -- device-side index view and warp-level primitives of the deSAMBA hot path (sm_100a).
//
// Execution model of the classify kernel: ONE WARP PER READ.  All 32 lanes run the per-read control flow in
// lock-step with identical register values ("warp-uniform" code); the lanes split the work only inside the
// cooperative primitives below (FM-index occ blocks, visited-row set, reference-window unpacking, the read's
// 9-mer index), which return identical values to every lane again.  Uniform stores of identical values to one
// address are benign; lane-partitioned stores are followed by __syncwarp().
//
// Arithmetic mirrors the DECLARED C types of the reference (mixed signed/unsigned MAX/MIN/ABS macros, utils.h:61-64)
// because the results depend on the usual arithmetic conversions.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define DSB_MAX(a,b) (((a) > (b))?(a):(b))
#define DSB_MIN(a,b) (((a) < (b))?(a):(b))
#define DSB_ABS(a) (((a) > 0)?(a): (- (a)))
#define DSB_ABS_U(a,b) (((a) > (b))?((a) - (b)): ((b) - (a)))

#define DSB_FORWARD 1
#define DSB_REVERSE 0
#define DSB_GUARD 64            // zero bytes before the forward strand and after the reverse strand (out-of-buffer policy P3)
#define DSB_FULL 0xffffffffu

// ---- HBM layout of the index -------------------------------------------------------------------------------------
// FM index: the reference's 168-byte blocks (5 x u64 counts + 256 nibbles, bwt.c:32-41) are re-cut at load time into
// 128-byte lines of 128 symbols:  u64 cnt[5] (A,C,G,T,# before the line) at byte 0 | pad | bit-plane 0 at byte 48 |
// bit-plane 1 at byte 64 | bit-plane 2 at byte 80 | pad.  Plane k holds bit k of the symbol code (A0 C1 G2 T3 #4 $5,
// padding 7); symbol i of the line is bit i of the 128-bit little-endian plane.  occ() = 1 count word + 3 x 16 B =
// 3 sectors of one aligned line, counted with and/xor/popc.
struct DevIndex {
	const uint8_t  *occ;        // n_lines * 128 B
	uint64_t        n_lines;
	uint64_t        rank[6];    // bwt.c:80-81
	uint64_t        dollar_pos; // idx.c:1128
	const uint64_t *prefix;     // hash_index[4^13+1], bwt.c:82-85
	const uint2    *sa;         // {unitig_ID, offset} per 8 rows, bwt.h:10-13
	const uint2    *uni;        // {ref_list, length}, n_uni + sentinel, idx.h:19-23
	uint64_t        n_uni;      // entries before the sentinel
	const uint64_t *ref_pos;    // REF_POS bitfield global_offset:40 ref_ID:23 direction:1, idx.h:33-39
	const uint8_t  *ref_bin;    // 2-bit packed reference, first base in bits 7-6, idx.c:594-603
	uint64_t        ref_bin_n;  // bytes in ref_bin; 1 KiB of zero slack follows, anything further reads as base 0
	const ulonglong2 *ref_info; // {seq_l, seq_offset}, idx.h:13-17
	const uint8_t  *ek0, *ek1;  // exist-k-mer bit tables, MSB-first, idx.c:1018-1021
	const uint32_t *ek0_sum;    // one bit per BYTE of ek0 (byte != 0), or null: 1/8 of the table, L2-resident for small indexes
	uint64_t        ek_mask;
	int             l_ek;
	int             single_base_max;
	const int      *q_mem;      // Q_MEM[2000], cly_mt.c:413-437 (host-computed, double -> int exactly as the reference)
	const int      *q_lv;       // Q_LV[20][20] row-major [d][l]
};

__device__ __forceinline__ int Q_MEM_at(const DevIndex &ix, uint32_t l) { return __ldg(ix.q_mem + l); }
__device__ __forceinline__ int Q_LV_at(const DevIndex &ix, uint32_t d, uint32_t l) { return __ldg(ix.q_lv + d * 20 + l); }

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// utils.c:1067-1091
__device__ __forceinline__ uint64_t dsb_hash64_1(uint64_t key)
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
__device__ __forceinline__ uint64_t dsb_hash64_2(uint64_t key)
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

__device__ __forceinline__ uint32_t ref_base_at(const DevIndex &ix, uint64_t o)
{
	const uint64_t byte = o >> 2;
	if (byte >= ix.ref_bin_n + 1024) return 0;     // the reference faults or reads foreign heap here
	return (__ldg(ix.ref_bin + byte) >> ((3 - (o & 3)) << 1)) & 3;
}

// get_ref (cly.c:435-466): unpack `length` bases starting at `off` forward, or walking backwards from `off`.
// Cooperative: lane k writes out[k], out[k+32], ...; out may be shared or global scratch.
__device__ __forceinline__ void get_ref_coop(const DevIndex &ix, uint8_t *out, int64_t off, int32_t length, bool forward)
{
	if (off < 0) off = 0;
	if (length < 0) length = 0;
	const uint64_t o = (uint64_t)off;
	for (uint32_t k = lane_id(); k < (uint32_t)length; k += 32)
		out[k] = (uint8_t)ref_base_at(ix, forward ? o + k : o - k);
	__syncwarp();
}

// Landau-Vishkin flank edit distance, <= 4 errors (cly.c:510-609), both strings of the same length here.
// `ref`/`query` may point into the shared flank frame or into the read; the reference writes '#'/'$' sentinels at
// [ref_length]/[query_length] and restores them -- indices never pass the sentinels, so they are virtual here.
static __device__ __noinline__ int32_t lv_extd_dev(const uint8_t *ref, int32_t ref_length, const uint8_t *query, int32_t query_length)
{
	if (ref_length < query_length) {
		int32_t t = ref_length; ref_length = query_length; query_length = t;
		const uint8_t *p = ref; ref = query; query = p;
	}
	int32_t mn_d[12], ed_d[12];
	int32_t *mn = mn_d + 5, *ed = ed_d + 5;
	#pragma unroll
	for (int i = -5; i <= 5; i++) { mn[i] = -1; ed[i] = (i > 0) ? (i) : (-i); }
	mn[6] = 0; ed[6] = 0;
	int32_t best_score = query_length;
	#define DSB_R(idx) (((idx) == ref_length) ? (uint32_t)'#' : (uint32_t)ref[(idx)])
	#define DSB_Q(idx) (((idx) == query_length) ? (uint32_t)'$' : (uint32_t)query[(idx)])
	#pragma unroll 1
	for (int i = 0; i <= 4; i++) {
		int32_t prev_mn = -1, cur_mn = (i - 1), next_mn = mn[-i + 1];
		int32_t prev_ed = i + 1, cur_ed = i, next_ed = ed[-i + 1];
		#pragma unroll 1
		for (int j = -i; j <= 4; j++) {
			int32_t m, e;
			if (cur_mn + j < ref_length - 1) {
				int best = cur_mn + 1 - cur_ed;
				m = cur_mn + 1; e = cur_ed + 1;
				if (best < next_mn + 1 - next_ed) { m = next_mn + 1; e = next_ed + 1; best = next_mn - next_ed; }
				if (best < prev_mn - prev_ed) { m = prev_mn + 1; e = prev_ed + 1; }
			} else {
				int best = cur_mn - cur_ed;
				m = cur_mn; e = cur_ed + 1;
				if (best < prev_mn - prev_ed) { m = prev_mn; e = prev_ed + 1; best = prev_mn - prev_ed; }
				if (best < next_mn + 1 - next_ed) { m = next_mn + 1; e = next_ed + 1; }
			}
			ed[j] = e;
			int mn_j = DSB_MIN(m, query_length);
			mn_j = DSB_MIN(mn_j, ref_length - j);
			for (; DSB_R(mn_j + j) == DSB_Q(mn_j); mn_j++);
			mn[j] = mn_j;
			if (DSB_Q(mn_j) == '$' || DSB_R(mn_j + j) == '#') {
				best_score = DSB_MIN(e - 1, best_score);
				if (j <= i + 1) return best_score;
			}
			prev_mn = cur_mn; cur_mn = next_mn; next_mn = mn[j + 2];
			prev_ed = cur_ed; cur_ed = next_ed; next_ed = ed[j + 2];
		}
	}
	#undef DSB_R
	#undef DSB_Q
	return best_score;
}

// Emulation of glibc 2.39 qsort (msort_with_tmp): top-down, n1 = n/2, left element taken while cmp(l, r) <= 0
// (SURVEY.md 5.9-H).  Warp-uniform, iterative (explicit stack), elements are moved with operator=.
template <typename T, typename Cmp>
__device__ void glibc_msort(T *b, T *tmp, int n, Cmp cmp)
{
	if (n <= 1) return;
	// frames: (lo, n, stage) ; stage 0 = descend left, 1 = descend right, 2 = merge
	int st_lo[40], st_n[40]; char st_stage[40];
	int sp = 0;
	st_lo[0] = 0; st_n[0] = n; st_stage[0] = 0;
	while (sp >= 0) {
		const int lo = st_lo[sp], cn = st_n[sp];
		if (cn <= 1) { sp--; continue; }
		const int n1 = cn / 2, n2 = cn - n1;
		if (st_stage[sp] == 0) { st_stage[sp] = 1; sp++; st_lo[sp] = lo; st_n[sp] = n1; st_stage[sp] = 0; continue; }
		if (st_stage[sp] == 1) { st_stage[sp] = 2; sp++; st_lo[sp] = lo + n1; st_n[sp] = n2; st_stage[sp] = 0; continue; }
		int i1 = lo, e1 = lo + n1, i2 = lo + n1, e2 = lo + cn, t = 0;
		while (i1 < e1 && i2 < e2) {
			if (cmp(b[i1], b[i2]) <= 0) tmp[t++] = b[i1++];
			else tmp[t++] = b[i2++];
		}
		while (i1 < e1) tmp[t++] = b[i1++];
		for (int k = 0; k < t; k++) b[lo + k] = tmp[k];      // the right run's tail is already in place
		sp--;
	}
}
