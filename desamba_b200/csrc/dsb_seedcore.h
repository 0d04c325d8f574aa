// dsb_seedcore.h -- the seeding engine: FM-index backward search, locate, Landau-Vishkin flank scoring -> anchors
// (reference behaviour: bwt_MEM_search / bwt_single_search cly.c:1344-1447, map_seed cly.c:706-939, get_new_ed
// cly.c:629-694, lv_extd cly.c:510-609, the per-seed schedules of fast_classify / slow_classify cly.c:1476-1611).
//
// Execution model: ONE LANE PER ISLAND SEED, WARP-UNIFORM MICRO-STEPS.  The unit of sequential work of the reference is
// one island seed (its k-mer loop has data-dependent strides and one visited-row set); seeds are independent of each
// other.  Every seed of a pass is a task in a flat list; a lane takes a task and carries it through a small state
// machine whose states are the memory-latency points of the work:
//     FETCH  take the next task            CTRL    start the next k-mer search / the next map_seed / finish the seed
//     OCC    one FM-index step (one backward-extension step with 2 occ, or one LF step of a single-row search or of a
//            locate walk)                  LOCATE  SA sample -> unitig -> first reference position
//     FLANK  fetch reference + read windows, exact-match extension, Landau-Vishkin on 2-bit packed registers
//     LV     one row of the Landau-Vishkin      RP     next reference position of the unitig (anchor push / per-reference re-extension)
// Each turn the warp votes for one state and only the lanes in that state run its handler, all executing the same
// instructions, all their loads in flight together.  Nothing in a handler is warp-cooperative: the handlers are plain
// per-lane functions, compiled for the device (k_seed, dsb_seed.cuh) and -- unchanged -- for the host, where
// tests/emul runs 32 emulated lanes against the oracle.  The two collective steps (task fetch, staging-chunk
// allocation) are outside, in the warp loop.
//
// All strings are 2-bit packed, MSB first: a "window" is a 64-bit word whose top two bits are the first base.  The three
// 13-byte flank arrays of map_seed's stack frame (q_pre | t_pre | t_suf, oracle policy P2) are three 32-bit registers.
#pragma once
#include "dsb_index_view.h"
#include "../../include/desamba_b200.h"

#if defined(__CUDACC__)
#define SC_HD __host__ __device__ __forceinline__
#define SC_HDN static __host__ __device__ __noinline__
#else
#define SC_HD static inline
#define SC_HDN static
#endif
#if defined(__CUDA_ARCH__)
#define SC_DEVICE 1
#else
#define SC_DEVICE 0
#endif

#define SC_MAX(a,b) (((a) > (b))?(a):(b))
#define SC_MIN(a,b) (((a) < (b))?(a):(b))

// ---------------------------------------------------------------- portable primitives
SC_HD uint64_t sc_ld64(const uint64_t *p)
{
#if SC_DEVICE
	return __ldg(p);
#else
	return *p;
#endif
}
SC_HD uint32_t sc_ld32(const uint32_t *p)
{
#if SC_DEVICE
	return __ldg(p);
#else
	return *p;
#endif
}
SC_HD int sc_ldi(const int *p)
{
#if SC_DEVICE
	return __ldg(p);
#else
	return *p;
#endif
}
SC_HD uint2 sc_ld2(const uint2 *p)
{
#if SC_DEVICE
	return __ldg(p);
#else
	return *p;
#endif
}
SC_HD uint4 sc_ld128(const uint4 *p)
{
#if SC_DEVICE
	return __ldg(p);
#else
	return *p;
#endif
}
SC_HD int sc_popc64(uint64_t x)
{
#if SC_DEVICE
	return __popcll(x);
#else
	return __builtin_popcountll(x);
#endif
}
SC_HD int sc_clz64(uint64_t x)               // 64 for x == 0
{
#if SC_DEVICE
	return __clzll((long long)x);
#else
	return x ? __builtin_clzll(x) : 64;
#endif
}
SC_HD uint64_t sc_brev64(uint64_t x)
{
#if SC_DEVICE
	return __brevll(x);
#else
	x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
	x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
	x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
	return __builtin_bswap64(x);
#endif
}
SC_HD uint64_t sc_bswap64(uint64_t x)
{
#if SC_DEVICE
	const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
	return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
#else
	return __builtin_bswap64(x);
#endif
}
// reverse the order of the 32 two-bit groups of a window
SC_HD uint64_t sc_rev2(uint64_t x)
{
	x = sc_brev64(x);
	return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}
// (hi:lo) << sh, upper 64 bits; sh in 0..63
SC_HD uint64_t sc_funnel(uint64_t hi, uint64_t lo, uint32_t sh) { return sh ? ((hi << sh) | (lo >> (64 - sh))) : hi; }
// length of the common prefix (in bases) of two windows
SC_HD int sc_common(uint64_t a, uint64_t b) { return sc_clz64(a ^ b) >> 1; }

// ---------------------------------------------------------------- data of a seeding pass
struct SeedTaskRef { uint32_t read; uint32_t sk; };        // sk: bit 31 = strand (0 forward, 1 reverse complement), low bits = index of the island seed in the strand's seed list
// what a finished seed leaves behind for the ordered gather (phase_chain): its anchors as a chunk list in the staging pool
struct SeedRec { uint32_t first_chunk, count; int32_t top_score; uint32_t flag512; uint32_t c_occ, c_getref, c_getref_bytes, c_pl; };   // c_pl = prefix look-ups | locates << 16
#define STAGE_PER_CHUNK 3                                  // a chunk = 3 staged anchors of 16 B + {next chunk, -, -, -} = 64 B
#define SC_NO_CHUNK 0xffffffffu
#define VIS1_SLOTS 64                                      // tier 1 of the visited-row set (shared memory): open addressing, at most VIS1_MAX rows
#define VIS1_MAX 48
#define VIS2_SLOTS 1024                                    // tier 2 (HBM, generation tagged): the rows beyond the first VIS1_MAX of a seed
#define SEED_MEM_SLOTS 8                                   // MEM results a lane holds: 2 of one search (fast) / the 8 longest of the seed (slow)

struct SeedEnv {
	DevIndex ix;
	const uint64_t *pk;            // 2-bit packed FORWARD strands, 32 bases per word, first base in the top bits; word 0 and one word behind every read are padding
	const uint64_t *read_off;      // start of each read in the concatenated input (read length)
	const uint64_t *bits_off;      // per read: 5 * (its first packed word - 1)
	const uint32_t *seed_off;      // per read: first seed slot of the strand arrays
	const dsb_seed *seeds[2];
	const SeedTaskRef *tasks;
	SeedRec *recs;                 // one per task
	uint4 *chunks; uint32_t n_chunks;
	int slow;                      // 0: fast_classify schedule, 1: slow_classify schedule
	int big_rows;                  // the BWT has >= 2^32 rows: tier-1 tags are confirmed against the full rows
};

// per-lane memory outside the registers
struct LaneMem {
#if defined(__CUDACC__)
	uint32_t vis1;                 // shared-memory address of tier-1 slot 0 of this lane (slot i at + 128 i bytes: conflict free)
#else
	uint32_t *vis1;
#endif
	uint64_t *vis1_full;           // full rows of the tier-1 slots (only touched when big_rows)
	uint64_t *vis2;
	MemRst *mem;
#if defined(__CUDACC__)
	uint32_t lvs;                  // shared-memory address of the lane's Landau-Vishkin state (4 x u64, word k at + 256 k bytes)
#else
	uint64_t *lvs;
#endif
};

enum { ST_FETCH = 0, ST_CTRL, ST_OCC, ST_LOCATE, ST_FLANK, ST_LV, ST_RP, ST_DEAD, SC_N_STATES };
enum { CK_KMER = 0, CK_MAP, CK_FIN, CK_ROWS, CK_MAPDONE }; // kinds of CTRL step
enum { OK_EXT = 0, OK_SINGLE, OK_WALK1, OK_WALK2 };        // kinds of OCC step
enum { FK_PRE = 0, FK_SUF, FK_NEWL, FK_NEWR };             // kinds of FLANK step
enum { AL_PRE = 0, AL_SUF };                               // what follows a locate

struct SeedLane {
	uint32_t st, kind;
	// the task
	uint32_t task, pkw, read_len, strand, s_off;
	int32_t  j;
	// what the seed has produced
	uint32_t n_out, first_chunk, cur_chunk, flag512;
	int32_t  top_score;
	uint32_t c_occ, c_getref, c_getref_bytes, c_pl;
	// visited rows (sp_set, cly.c:1286-1298)
	uint64_t vis_mask; int32_t sp_l; uint32_t vis_gen;
	// the running k-mer search
	int32_t  si, pos, ml, ext_ml, pos_ext, sa_l;
	uint64_t sp, ep;               // search: row interval [sp, ep) / single-row search: row, last sampled row; map_seed: b_p, t_off
	uint64_t row_next; uint32_t rows_left;
	uint32_t n_found, i_mem; int32_t max_score;
	// the running map_seed
	int32_t  q_off; uint32_t l_m; int32_t s_l, uni; uint32_t uni_len, u_off;
	uint32_t l_pre, d_pre, l_suf, d_suf; int32_t s, max_s;
	uint32_t A, B, C;              // q_pre | t_pre | t_suf (13 entries each, entry e at bits 31-2e, 30-2e); B bits 5..0: entry holding '$'
	uint64_t loc_row; int32_t loc_l; uint32_t after_loc;
	int32_t  fl_q; uint64_t fl_t; uint32_t fl_max, fl_ext;
	uint32_t c_r_p, r_p_e, rp_ref, ext_l; uint64_t rp_global;
	uint32_t am_mtch_len; int32_t am_score; uint32_t am_ll, am_le, am_rl, am_re, rsl, rsr;
	uint32_t push, row_res, lv_st;
	int32_t  ret, error;
};

// ---------------------------------------------------------------- windows on the read strands and on the reference
SC_HD uint64_t fwd_win(const uint64_t *rd, int64_t p)              // 32 bases of the forward strand from position p (>= -32)
{
	const int64_t w = p >> 5;
	return sc_funnel(sc_ld64(rd + w), sc_ld64(rd + w + 1), 2u * (uint32_t)(p & 31));
}
// 32 bases of the lane's strand: positions p, p+1, ...   (reverse strand: base p = 3 - forward[len-1-p], cly.c:1258-1259)
SC_HD uint64_t strand_win(const SeedEnv &E, const SeedLane &L, int64_t p)
{
	const uint64_t *rd = E.pk + L.pkw;
	if (L.strand == 0) return fwd_win(rd, p);
	return ~sc_rev2(fwd_win(rd, (int64_t)L.read_len - 1 - p - 31));
}
// positions p, p-1, p-2, ...
SC_HD uint64_t strand_win_left(const SeedEnv &E, const SeedLane &L, int64_t p)
{
	const uint64_t *rd = E.pk + L.pkw;
	if (L.strand == 0) return sc_rev2(fwd_win(rd, p - 31));
	return ~fwd_win(rd, (int64_t)L.read_len - 1 - p);
}
// one base of the lane's strand.  Position -1 is read by bwt_MEM_search before its length check (cly.c:1402-1411): in front
// of the forward strand sits the malloc header (0 at [-1..-6], oracle policy P3), in front of the reverse strand the forward one
SC_HD uint32_t strand_base(const SeedEnv &E, const SeedLane &L, int32_t p)
{
	const uint64_t *rd = E.pk + L.pkw;
	if (L.strand == 0) {
		if (p < 0) return 0;
		return (uint32_t)(sc_ld64(rd + (p >> 5)) >> (62 - 2 * (p & 31))) & 3;
	}
	if (p < 0) { const int32_t f = (int32_t)L.read_len + p; return (uint32_t)(sc_ld64(rd + (f >> 5)) >> (62 - 2 * (f & 31))) & 3; }
	const int32_t f = (int32_t)L.read_len - 1 - p;
	return 3 - ((uint32_t)(sc_ld64(rd + (f >> 5)) >> (62 - 2 * (f & 31))) & 3);
}

SC_HD uint32_t ref_base_at(const DevIndex &ix, uint64_t o)
{
	const uint64_t byte = o >> 2;
	if (byte >= ix.ref_bin_n + 1024) return 0;     // the reference faults or reads foreign heap here
#if SC_DEVICE
	return (__ldg(ix.ref_bin + byte) >> ((3 - (o & 3)) << 1)) & 3;
#else
	return (ix.ref_bin[byte] >> ((3 - (o & 3)) << 1)) & 3;
#endif
}
// get_ref forward (cly.c:435-466): 32 bases of the packed reference from global position o
SC_HD uint64_t ref_win(const DevIndex &ix, uint64_t o)
{
	const uint64_t w = o >> 5;
	if ((w + 2) * 8 <= ix.ref_bin_n + 1024) {
		const uint64_t *p = (const uint64_t *)ix.ref_bin + w;
		return sc_funnel(sc_bswap64(sc_ld64(p)), sc_bswap64(sc_ld64(p + 1)), 2u * (uint32_t)(o & 31));
	}
	uint64_t v = 0;
	for (int k = 0; k < 32; k++) v = (v << 2) | ref_base_at(ix, o + k);
	return v;
}
// get_ref backward: positions o, o-1, ...; positions below 0 read as base 0 (the reference wraps around there)
SC_HD uint64_t ref_win_left(const DevIndex &ix, uint64_t o)
{
	if (o >= 31) return sc_rev2(ref_win(ix, o - 31));
	return sc_rev2(ref_win(ix, 0)) << (2 * (31 - (uint32_t)o));
}

SC_HD int Q_MEM_at(const DevIndex &ix, uint32_t l) { return sc_ldi(ix.q_mem + l); }
SC_HD int Q_LV_at(const DevIndex &ix, uint32_t d, uint32_t l) { return sc_ldi(ix.q_lv + d * 20 + l); }

// ---------------------------------------------------------------- occ (bwt.c:43-65) on the bit-plane lines
SC_HD uint32_t plane_count(const uint4 &p0, const uint4 &p1, const uint4 &p2, int in, uint32_t c)
{   // number of symbols equal to c among the first `in` (0..127) symbols of the line
	const uint64_t a0 = (uint64_t)p0.x | ((uint64_t)p0.y << 32), a1 = (uint64_t)p0.z | ((uint64_t)p0.w << 32);
	const uint64_t b0 = (uint64_t)p1.x | ((uint64_t)p1.y << 32), b1 = (uint64_t)p1.z | ((uint64_t)p1.w << 32);
	const uint64_t c0 = (uint64_t)p2.x | ((uint64_t)p2.y << 32), c1 = (uint64_t)p2.z | ((uint64_t)p2.w << 32);
	const uint64_t x0 = (c & 1) ? 0ull : ~0ull, x1 = (c & 2) ? 0ull : ~0ull, x2 = (c & 4) ? 0ull : ~0ull;
	const uint64_t m_lo = (in >= 64) ? ~0ull : ((1ull << in) - 1);
	const uint64_t m_hi = (in > 64) ? ((1ull << (in - 64)) - 1) : 0ull;
	return (uint32_t)(sc_popc64((a0 ^ x0) & (b0 ^ x1) & (c0 ^ x2) & m_lo) + sc_popc64((a1 ^ x0) & (b1 ^ x1) & (c1 ^ x2) & m_hi));
}
SC_HD uint32_t plane_symbol(const uint4 &p0, const uint4 &p1, const uint4 &p2, int in)
{
	const uint32_t w = (uint32_t)in >> 5, sh = (uint32_t)in & 31;
	const uint32_t q0 = (w == 0) ? p0.x : (w == 1) ? p0.y : (w == 2) ? p0.z : p0.w;
	const uint32_t q1 = (w == 0) ? p1.x : (w == 1) ? p1.y : (w == 2) ? p1.z : p1.w;
	const uint32_t q2 = (w == 0) ? p2.x : (w == 1) ? p2.y : (w == 2) ? p2.z : p2.w;
	return ((q0 >> sh) & 1) | (((q1 >> sh) & 1) << 1) | (((q2 >> sh) & 1) << 2);
}

// ---------------------------------------------------------------- visited-row set (sp_set_insert, cly.c:1286-1298)
// The reference keeps <= 500 rows in an array, scans it on every insert, empties it when it is full and at the start of
// every seed.  Same semantics with two hash tiers: the first VIS1_MAX rows of a seed go to a 64-slot table of 32-bit tags in
// shared memory (occupancy = a 64-bit register; rows are < 2^32 unless the index is beyond 4 G symbols -- then a tag match
// is confirmed against the full row kept in HBM), later rows to a generation-tagged table in HBM.
SC_HD uint32_t vis1_ld(const LaneMem &M, uint32_t i)
{
#if SC_DEVICE
	uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(M.vis1 + 128u * i) : "memory"); return v;
#elif defined(__CUDACC__)
	return 0 * M.vis1 * i;                                 // (host pass of nvcc: never called)
#else
	return M.vis1[i];
#endif
}
SC_HD void vis1_st(const LaneMem &M, uint32_t i, uint32_t v)
{
#if SC_DEVICE
	asm volatile("st.shared.u32 [%0], %1;" :: "r"(M.vis1 + 128u * i), "r"(v) : "memory");
#elif defined(__CUDACC__)
	(void)M; (void)i; (void)v;
#else
	M.vis1[i] = v;
#endif
}
SC_HDN void vis2_zero(uint64_t *vis2)                     // (cold: once per 16 M seeds of a lane)
{
	#pragma unroll 1
	for (int i = 0; i < VIS2_SLOTS; i++) vis2[i] = 0;
}
SC_HD void vis_clear(SeedLane &L, const LaneMem &M)
{
	L.vis_mask = 0; L.sp_l = 0;
	if (++L.vis_gen >= (1u << 24)) { vis2_zero(M.vis2); L.vis_gen = 1; }
}
SC_HD int vis_insert(const SeedEnv &E, SeedLane &L, const LaneMem &M, uint64_t row)
{
	if (L.sp_l == SP_SET_CAP) vis_clear(L, M);
	const uint64_t hh = row * 0x9E3779B97F4A7C15ull;
	const uint32_t tag = (uint32_t)row;
	uint32_t i = (uint32_t)(hh >> 58);
	const uint64_t mask = L.vis_mask;
	while ((mask >> i) & 1) {
		if (vis1_ld(M, i) == tag && (!E.big_rows || M.vis1_full[i] == row)) return 0;
		i = (i + 1) & (VIS1_SLOTS - 1);
	}
	if (sc_popc64(mask) < VIS1_MAX) {
		vis1_st(M, i, tag);
		if (E.big_rows) M.vis1_full[i] = row;
		L.vis_mask = mask | (1ull << i);
		L.sp_l++;
		return 1;
	}
	const uint64_t key = ((uint64_t)L.vis_gen << 40) | (row & 0xFFFFFFFFFFull);
	uint32_t h = (uint32_t)(hh >> 54);
	for (;;) {
		const uint64_t v = M.vis2[h];
		if (v == key) return 0;
		if ((uint32_t)(v >> 40) != L.vis_gen) { M.vis2[h] = key; L.sp_l++; return 1; }
		h = (h + 1) & (VIS2_SLOTS - 1);
	}
}

// ---------------------------------------------------------------- Landau-Vishkin flank edit distance (lv_extd, cly.c:510-609)
// Both strings have the same length l <= 12 everywhere on this path, so the reference's swap never happens.  The reference
// indexes up to 5 bytes BEFORE either string when the flank is short: those bytes are the neighbouring
// array of the frame or the preceding read bases, so a string comes as an "extended" window: index i (-5 <= i <= 12) at
// bits 63-2(i+5), 62-2(i+5).  Index l is the sentinel ('#' / '$'): it matches nothing, which is a length limit here.
// The match-number / edit-distance rows mn[-5..6], ed[-5..6] are twelve 4-bit fields of a 64-bit register each (mn stored +1);
// mn[6] = ed[6] = 0 is the zero-initialised word behind the initialised part (oracle policy P1).
SC_HD int lv_get(uint64_t v, int j) { return (int)((v >> (4 * (j + 5))) & 15); }
SC_HD uint64_t lv_set(uint64_t v, int j, int x) { const int sh = 4 * (j + 5); return (v & ~(15ull << sh)) | ((uint64_t)x << sh); }
SC_HDN int32_t lv_packed(uint64_t er, uint64_t eq, int32_t len)
{
	uint64_t MN = 1ull << 44, ED = 0x054321012345ull;
	int32_t best_score = len;
	#pragma unroll 1
	for (int i = 0; i <= 4; i++) {
		int prev_mn = -1, cur_mn = i - 1, next_mn = lv_get(MN, -i + 1) - 1;
		int prev_ed = i + 1, cur_ed = i, next_ed = lv_get(ED, -i + 1);
		#pragma unroll 1
		for (int j = -i; j <= 4; j++) {
			int m, e;
			if (cur_mn + j < len - 1) {
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
			ED = lv_set(ED, j, e);
			int mn_j = SC_MIN(m, len);
			mn_j = SC_MIN(mn_j, len - j);
			{	// in-line match along diagonal j: reference index mn_j + j against query index mn_j, up to the sentinels
				const int a = mn_j + j;
				int run = sc_common(er << (2 * (a + 5)), eq << (2 * (mn_j + 5)));
				run = SC_MIN(run, SC_MIN(len - a, len - mn_j));
				mn_j += run;
			}
			MN = lv_set(MN, j, mn_j + 1);
			if (mn_j == len || mn_j + j == len) {
				best_score = SC_MIN(e - 1, best_score);
				if (j <= i + 1) return best_score;
			}
			prev_mn = cur_mn; cur_mn = next_mn; next_mn = lv_get(MN, j + 2) - 1;
			prev_ed = cur_ed; cur_ed = next_ed; next_ed = lv_get(ED, j + 2);
		}
	}
	return best_score;
}
// the same on bytes: index i of a string at [i + 5]; used when the t_pre array holds the '$' symbol (a walk that ran off the
// start of unitig 0, cly.c:749-756) and by the tests as the plain statement of the packed version
SC_HDN int32_t lv_bytes(const uint8_t *ref, const uint8_t *query, int32_t len)
{
	int32_t mn_d[12], ed_d[12];
	int32_t *mn = mn_d + 5, *ed = ed_d + 5;
	#pragma unroll 1
	for (int i = -5; i <= 5; i++) { mn[i] = -1; ed[i] = (i > 0) ? (i) : (-i); }
	mn[6] = 0; ed[6] = 0;
	int32_t best_score = len;
	#define SC_R(idx) (((idx) == len) ? (uint32_t)'#' : (uint32_t)ref[(idx) + 5])
	#define SC_Q(idx) (((idx) == len) ? (uint32_t)'$' : (uint32_t)query[(idx) + 5])
	#pragma unroll 1
	for (int i = 0; i <= 4; i++) {
		int32_t prev_mn = -1, cur_mn = (i - 1), next_mn = mn[-i + 1];
		int32_t prev_ed = i + 1, cur_ed = i, next_ed = ed[-i + 1];
		#pragma unroll 1
		for (int j = -i; j <= 4; j++) {
			int32_t m, e;
			if (cur_mn + j < len - 1) {
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
			int mn_j = SC_MIN(m, len);
			mn_j = SC_MIN(mn_j, len - j);
			for (; SC_R(mn_j + j) == SC_Q(mn_j); mn_j++);
			mn[j] = mn_j;
			if (mn_j == len || mn_j + j == len) {
				best_score = SC_MIN(e - 1, best_score);
				if (j <= i + 1) return best_score;
			}
			prev_mn = cur_mn; cur_mn = next_mn; next_mn = mn[j + 2];
			prev_ed = cur_ed; cur_ed = next_ed; next_ed = ed[j + 2];
		}
	}
	#undef SC_R
	#undef SC_Q
	return best_score;
}

// ---------------------------------------------------------------- the flank frame in registers
#define FR_POISON 0x20u                                    // B bit 5: entry (B & 15) holds the '$' symbol
// write the first l entries of a frame array from the top of a window
SC_HD uint32_t frame_merge(uint32_t x, uint64_t w, uint32_t l)
{
	if (l == 0) return x;
	const uint32_t m = 0xffffffffu << (32 - 2 * l);        // l <= 13
	uint32_t y = ((uint32_t)(w >> 32) & m) | (x & ~m);
	if ((x & FR_POISON) && (x & 15) < l) y &= ~0x3Fu;      // the '$' entry is overwritten
	return y;
}
SC_HD uint32_t frame_tail(uint32_t x) { return (x >> 6) & 0x3FF; }      // entries 8..12: what index -5..-1 of the NEXT array of the frame reads
SC_HD uint64_t frame_ext(uint32_t pre10, uint32_t x) { return ((uint64_t)pre10 << 54) | ((uint64_t)(x & ~0x3Fu) << 22); }
SC_HD void frame_bytes(uint8_t *out, uint32_t pre10, uint32_t x, bool poison_here, uint32_t pre_poison_entry)
{   // out[0..4] = index -5..-1, out[5..17] = entries 0..12
	for (int k = 0; k < 5; k++) out[k] = (uint8_t)((pre10 >> (8 - 2 * k)) & 3);
	for (int e = 0; e < 13; e++) out[5 + e] = (uint8_t)((x >> (30 - 2 * e)) & 3);
	if (poison_here) out[5 + (x & 15)] = 5;
	if (pre_poison_entry >= 8 && pre_poison_entry <= 12) out[pre_poison_entry - 8] = 5;
}
// Landau-Vishkin of two frame arrays / of a frame array and a read window; B may hold '$'
SC_HD int32_t lv_frames(uint32_t ref_pre_arr, uint32_t ref_arr, uint64_t eq, int32_t len, bool ref_is_B, uint32_t B)
{
	const bool poisoned = (B & FR_POISON) != 0;
	if (!poisoned) return lv_packed(frame_ext(frame_tail(ref_pre_arr), ref_arr), eq, len);
	uint8_t r[18], q[18];
	frame_bytes(r, frame_tail(ref_pre_arr), ref_arr, ref_is_B, ref_is_B ? 99u : (B & 15));
	for (int k = 0; k < 18; k++) q[k] = (uint8_t)((eq >> (62 - 2 * k)) & 3);
	return lv_bytes(r, q, len);
}

// ---------------------------------------------------------------- small helpers of the state machine
SC_HD void count_getref(SeedLane &L, uint32_t len) { L.c_getref++; L.c_getref_bytes += (len + 3) >> 2; }
SC_HD void go(SeedLane &L, uint32_t st, uint32_t kind) { L.st = st; L.kind = kind; }
// map_seed returns `score` (cly.c:938): the seed's schedule goes on in CTRL
SC_HD void map_done(SeedLane &L, int32_t score) { L.ret = score; go(L, ST_CTRL, CK_MAPDONE); }

// one window call site each for the read and for the reference: 32 bases of the lane's strand from position p, rightwards
// (p, p+1, ...) or leftwards (p, p-1, ...).  Reverse strand: base p = 3 - forward[len-1-p] (cly.c:1258-1259).
SC_HD uint64_t strand_window(const SeedEnv &E, const SeedLane &L, int64_t p, uint32_t left)
{
	const uint64_t *rd = E.pk + L.pkw;
	const uint32_t rev = L.strand ^ left;
	const int64_t base = L.strand ? (int64_t)L.read_len - 1 - p : p;
	uint64_t w = fwd_win(rd, rev ? base - 31 : base);
	if (rev) w = sc_rev2(w);
	return L.strand ? ~w : w;
}
SC_HDN uint64_t ref_win_slow(const DevIndex &ix, uint64_t o)
{
	uint64_t v = 0;
	for (int k = 0; k < 32; k++) v = (v << 2) | ref_base_at(ix, o + k);
	return v;
}
// get_ref (cly.c:435-466) forward from o, or backward from o (positions below 0 read as base 0: the reference wraps around there)
SC_HD uint64_t ref_window(const DevIndex &ix, uint64_t o, uint32_t left)
{
	const uint64_t s = left ? (o >= 31 ? o - 31 : 0) : o;
	const uint64_t wi = s >> 5;
	uint64_t w;
	if ((wi + 2) * 8 <= ix.ref_bin_n + 1024) {
		const uint64_t *p = (const uint64_t *)ix.ref_bin + wi;
		w = sc_funnel(sc_bswap64(sc_ld64(p)), sc_bswap64(sc_ld64(p + 1)), 2u * (uint32_t)(s & 31));
	} else
		w = ref_win_slow(ix, s);
	if (left) { w = sc_rev2(w); if (o < 31) w <<= 2 * (31 - (uint32_t)o); }
	return w;
}

// a MEM result of the running search (cly.c:1429-1431 / 1441-1443): fast mode keeps the <= 2 of the search in arrival order,
// slow mode the 8 longest of the whole seed, longest first, earlier ones first among equals -- the head of the reference's
// stable sort by match length (qsort with MEM_rst_cmp_by_match_len, cly.c:1595: a consistent order, so glibc's merge sort is
// THE stable sort) of which only the first 8 are mapped (cly.c:1598)
SC_HD void mem_keep(const SeedEnv &E, SeedLane &L, const LaneMem &M, const MemRst &r)
{
	if (!E.slow) { M.mem[L.n_found++] = r; return; }
	uint32_t n = L.n_found;
	if (n == SEED_MEM_SLOTS && M.mem[n - 1].match_len >= r.match_len) return;
	uint32_t p = (n == SEED_MEM_SLOTS) ? n - 1 : n;        // slot that becomes free
	while (p > 0 && M.mem[p - 1].match_len < r.match_len) { M.mem[p] = M.mem[p - 1]; p--; }
	M.mem[p] = r;
	if (n < SEED_MEM_SLOTS) L.n_found = n + 1;
}

// ---------------------------------------------------------------- FETCH: a lane takes task t
SC_HD void task_begin(const SeedEnv &E, SeedLane &L, const LaneMem &M, uint32_t t)
{
	const SeedTaskRef tr = E.tasks[t];
	const uint32_t r = tr.read, strand = tr.sk >> 31, k = tr.sk & 0x7fffffffu;
	const dsb_seed sv = E.seeds[strand][E.seed_off[r] + k];
	L.task = t;
	L.read_len = (uint32_t)(E.read_off[r + 1] - E.read_off[r]);
	L.pkw = (uint32_t)(E.bits_off[r] / 5) + 1;
	L.strand = strand;
	L.s_off = sv.offset;
	L.j = (int32_t)sv.len - 1;
	L.n_out = 0; L.first_chunk = SC_NO_CHUNK; L.cur_chunk = SC_NO_CHUNK; L.flag512 = 0; L.top_score = 35;
	L.c_occ = L.c_getref = L.c_getref_bytes = L.c_pl = 0;
	L.n_found = 0; L.i_mem = 0; L.max_score = 0; L.error = 0; L.push = 0;
	vis_clear(L, M);
	go(L, ST_CTRL, CK_KMER);
}

// ---------------------------------------------------------------- CTRL: everything between the memory-bound steps of a seed
// (the kinds run top to bottom, so one turn can pass through several of them)
SC_HD void h_ctrl(const SeedEnv &E, SeedLane &L, const LaneMem &M)
{
	const DevIndex &ix = E.ix;
	const int l_min = E.slow ? SC_MIN(19, ix.l_ek + 1) : 20;
	if (L.kind == CK_MAPDONE) {
		// map_seed returned L.ret: cly.c:1519-1533 (fast) / 1599-1600 (slow)
		L.i_mem++;
		if (L.error) L.kind = CK_FIN;
		else if (E.slow) L.kind = (L.i_mem < L.n_found) ? CK_MAP : CK_FIN;
		else {
			L.max_score = SC_MAX(L.ret, L.max_score);
			if (L.i_mem < L.n_found) L.kind = CK_MAP;
			else {
				if (L.max_score > 35) L.j -= 7;
				if (L.max_score > 256) {
					if (L.max_score > 512) L.flag512 = 1;          // "skip next and break": the gather drops the next seed (cly.c:1530-1531)
					L.kind = CK_FIN;
				} else L.kind = CK_KMER;
			}
		}
	}
	if (L.kind == CK_ROWS) {
		// the rows of the interval the backward extension ended with (cly.c:1425-1446): every row not seen before starts a
		// single-row search from the same read position; L.row_res: a single-row search has just ended (1) / was aborted (0)
		uint32_t pending = L.row_res;
		L.row_res = 0;
		for (;;) {
			if (pending) {
				const int32_t total = L.ml + L.ext_ml + 1;
				if (total >= l_min) {
					MemRst r; r.match_len = total; r.sa_sp_l = L.sa_l; r.sp = L.sp; r.sa_sp = L.ep; r.read_offset = L.si - total; r.pad = 0;
					mem_keep(E, L, M, r);
				}
				pending = 0;
			}
			if (!L.rows_left) break;
			const uint64_t c_sp = L.row_next++;
			L.rows_left--;
			if (vis_insert(E, L, M, c_sp) == 0) continue;
			L.sp = c_sp; L.ep = NO_SA; L.sa_l = 0; L.ml = 0; L.pos = L.pos_ext;
			if (L.si - L.ext_ml <= 0) { pending = 1; continue; }            // bwt_single_search leaves at once (cly.c:1356)
			go(L, ST_OCC, OK_SINGLE);
			return;
		}
		// bwt_MEM_search returned: cly.c:1510-1517 / 1582-1591
		if (E.slow || L.n_found == 0) { L.j -= 2; L.kind = CK_KMER; }
		else { L.j -= 3; L.i_mem = 0; L.max_score = 0; L.kind = CK_MAP; }
	}
	if (L.kind == CK_KMER) {
		const bool more = E.slow ? (L.j >= 1) : (L.j >= 21 - ix.l_ek);     // cly.c:1500 / 1570
		if (!more) {
			if (E.slow && L.n_found > 0) { L.i_mem = 0; L.kind = CK_MAP; }  // cly.c:1592-1600
			else L.kind = CK_FIN;
		}
	}
	if (L.kind == CK_FIN) {
		SeedRec rec;
		rec.first_chunk = L.first_chunk; rec.count = L.n_out; rec.top_score = L.top_score; rec.flag512 = L.flag512 | ((uint32_t)L.error << 8);
		rec.c_occ = L.c_occ; rec.c_getref = L.c_getref; rec.c_getref_bytes = L.c_getref_bytes; rec.c_pl = L.c_pl;
		E.recs[L.task] = rec;
		go(L, ST_FETCH, 0);
		return;
	}
	// CK_KMER: one k-mer of the island -- prefix-table interval of its last 13 bases, then backward extension (cly.c:1502-1509,
	// 1399-1401).  CK_MAP: map_seed of MEM result i_mem (cly.c:706-733).  Both start with a window of the read.
	const bool is_map = L.kind == CK_MAP;
	MemRst m; m.sp = 0; m.sa_sp = NO_SA; m.match_len = 0; m.sa_sp_l = 0; m.read_offset = 0;
	if (is_map) m = M.mem[L.i_mem];
	const int32_t si = (int32_t)L.s_off + L.j + ix.l_ek - 1;
	const uint64_t w = strand_window(E, L, is_map ? (int64_t)m.read_offset : (int64_t)si - 12, is_map ? 1u : 0u);
	if (!is_map) {
		const uint64_t pre_v = w >> 38;
		L.sp = sc_ld64(ix.prefix + pre_v); L.ep = sc_ld64(ix.prefix + pre_v + 1);
		L.c_pl++;
		L.si = si; L.pos = si - L_PRE_IDX; L.ml = L_PRE_IDX;
		if (!E.slow) L.n_found = 0;
		go(L, ST_OCC, OK_EXT);
		return;
	}
	L.sp = m.sp; L.q_off = m.read_offset; L.l_m = (uint32_t)m.match_len;      // (b_p lives in L.sp)
	L.uni = -1; L.s_l = 0; L.s = 0; L.max_s = 0;
	L.A = L.B = L.C = 0;                                       // the frame starts zeroed (trivial-auto-var-init, oracle policy P1)
	L.l_pre = (uint32_t)SC_MIN(L.q_off + 1, LV_L);
	L.l_suf = L.d_suf = L.d_pre = 0; L.u_off = 0; L.uni_len = 0;
	L.A = frame_merge(0, w, L.l_pre);
	if (m.sa_sp != NO_SA) { L.loc_row = m.sa_sp; L.loc_l = m.sa_sp_l; L.after_loc = AL_PRE; go(L, ST_LOCATE, 0); }
	else if ((L.sp & SA_MASK) == 0) { L.loc_row = L.sp; L.loc_l = 0; L.after_loc = AL_PRE; go(L, ST_LOCATE, 0); }
	else go(L, ST_OCC, OK_WALK1);
}

// ---------------------------------------------------------------- OCC: one FM-index step
SC_HD void h_occ(const SeedEnv &E, SeedLane &L, const LaneMem &M)
{
	const DevIndex &ix = E.ix;
	const uint32_t kind = L.kind;
	const uint64_t rowA = L.sp;
	const uint4 *lineA = (const uint4 *)(ix.occ + (rowA >> 7) * 128);
	const uint4 pA0 = sc_ld128(lineA + 3), pA1 = sc_ld128(lineA + 4), pA2 = sc_ld128(lineA + 5);
	uint32_t c = 0;
	if (kind <= OK_SINGLE) c = strand_base(E, L, L.pos);
	uint4 pB0 = pA0, pB1 = pA1, pB2 = pA2; uint64_t cntB = 0;
	if (kind == OK_EXT) {
		const uint4 *lineB = (const uint4 *)(ix.occ + (L.ep >> 7) * 128);
		cntB = sc_ld64((const uint64_t *)lineB + c);
		pB0 = sc_ld128(lineB + 3); pB1 = sc_ld128(lineB + 4); pB2 = sc_ld128(lineB + 5);
	}
	const int inA = (int)(rowA & 127);
	uint32_t cA = c;
	if (kind != OK_EXT) cA = plane_symbol(pA0, pA1, pA2, inA);
	// LF(rowA) for symbol cA; '$' (5) goes to DOLLOR_POS (bwt.c:54-55).  The count word of the symbol is a second access to the line just fetched.
	uint64_t lfA = ix.dollar_pos + ix.rank[5];
	if (cA != 5) lfA = ix.rank[cA] + sc_ld64((const uint64_t *)lineA + cA) + plane_count(pA0, pA1, pA2, inA, cA);
	if (kind <= OK_SINGLE) {
		uint64_t ins = lfA; bool do_ins = false, first_row = false;
		if (kind == OK_EXT) {
			// one step of the backward extension of bwt_MEM_search (cly.c:1403-1422)
			const uint64_t new_sp = lfA;
			const uint64_t new_ep = ix.rank[c] + cntB + plane_count(pB0, pB1, pB2, (int)(L.ep & 127), c);
			L.c_occ += 2;
			L.pos--;
			const int l_min = E.slow ? SC_MIN(19, ix.l_ek + 1) : 20, max_rst = E.slow ? 8 : 2;
			bool brk = false, none = false;
			if (L.ml >= l_min - 1) {
				if (new_sp + max_rst >= new_ep) brk = true;
				else if (L.ml >= L.si) none = true;                // longer than what is left of the read: no result
			}
			if (!brk && !none && new_sp + 1 >= new_ep) brk = true;
			if (!brk && !none) { L.ml++; L.sp = new_sp; L.ep = new_ep; return; }
			// the rows [new_sp, new_ep) go through single-row searches (cly.c:1425-1446): the first one starts right here
			L.ext_ml = L.ml; L.pos_ext = L.pos; L.row_res = 0;
			if (none || new_sp >= new_ep) { L.rows_left = 0; go(L, ST_CTRL, CK_ROWS); return; }
			L.row_next = new_sp + 1; L.rows_left = (uint32_t)(new_ep - new_sp) - 1;
			do_ins = true; first_row = true;
		} else {
			// one step of bwt_single_search (cly.c:1354-1382); L.ep = last sampled row, L.sa_l = steps since
			L.c_occ++;
			if ((rowA & SA_MASK) == 0) { L.ep = rowA; L.sa_l = 0; } else L.sa_l--;
			if (cA != c) { L.row_res = 1; go(L, ST_CTRL, CK_ROWS); return; }
			L.ml++; L.pos--;
			do_ins = true;
		}
		if (do_ins) {
			if (vis_insert(E, L, M, ins) == 0) { L.row_res = 0; go(L, ST_CTRL, CK_ROWS); return; }   // row seen before: this single-row search is dropped
			L.sp = ins;
			if (first_row) { L.ep = NO_SA; L.sa_l = 0; L.ml = 0; L.pos = L.pos_ext; L.kind = OK_SINGLE; }
			if (L.ml >= L.si - L.ext_ml) { L.row_res = 1; go(L, ST_CTRL, CK_ROWS); }            // cly.c:1356: the read is used up
		}
		return;
	}
	L.c_occ++;
	if (kind == OK_WALK1) {
		// locate walk of map_seed while the left flank is collected (cly.c:741-757)
		bool end = false;
		if (cA == 4) end = true;                                   // the "begin" of a unitig
		else {
			if (cA == 5) L.B = (L.B & ~0x3Fu) | FR_POISON | (uint32_t)L.s_l;
			else L.B |= cA << (30 - 2 * L.s_l);
			L.s_l++;
			L.sp = lfA;
			if ((uint32_t)L.s_l >= L.l_pre || (L.sp & SA_MASK) == 0) end = true;
		}
		if (!end) return;
		if ((L.sp & SA_MASK) == 0) { L.loc_row = L.sp; L.loc_l = L.s_l; L.after_loc = AL_PRE; go(L, ST_LOCATE, 0); }
		else { L.l_pre = (uint32_t)L.s_l; go(L, ST_FLANK, FK_PRE); }
		return;
	}
	// OK_WALK2: on to the next sampled row once the left flank has passed (cly.c:782-788)
	L.sp = lfA; L.s_l++;
	if (L.s_l > (1 << 20)) { L.error = 5; map_done(L, 0); return; }   // cannot happen on a well-formed index; never hang the GPU
	if ((L.sp & SA_MASK) == 0) { L.loc_row = L.sp; L.loc_l = L.s_l; L.after_loc = AL_SUF; go(L, ST_LOCATE, 0); }
}

// ---------------------------------------------------------------- right flank of map_seed (cly.c:796-835)
SC_HD void suf_begin(SeedLane &L)
{
	const int32_t q_off_r = L.q_off + (int32_t)L.l_m + 1;
	const uint32_t l_max_suf = SC_MIN(L.uni_len - L.u_off - L.l_m, L.read_len - (uint32_t)q_off_r);
	if (l_max_suf != 0) {
		L.fl_q = q_off_r; L.fl_max = l_max_suf; L.fl_ext = 0;
		count_getref(L, SC_MIN(l_max_suf, (uint32_t)LV_L));
		go(L, ST_FLANK, FK_SUF);
		return;
	}
	L.l_suf = L.d_suf = 0;                                     // cly.c:828-829; the test of cly.c:831 needs l_suf == 12
	if (!(L.s > 0)) map_done(L, 0); else go(L, ST_RP, 1);
}

// ---------------------------------------------------------------- LOCATE (get_uni, cly.c:471-496)
SC_HD void h_locate(const SeedEnv &E, SeedLane &L)
{
	const DevIndex &ix = E.ix;
	L.c_pl += 1u << 16;
	const uint2 sa = sc_ld2(ix.sa + (L.loc_row >> SA_OFF));
	int64_t u = sa.x;
	uint32_t uni_offset = sa.y + (uint32_t)L.loc_l + 1;
	uint2 ul = sc_ld2(ix.uni + u);
	if (L.loc_l > 0)
		for (;;) {
			if (!(uni_offset >= ul.y) || u >= (int64_t)ix.n_uni) break;            // (bound: the reference walks off its table here)
			uni_offset -= (ul.y + 1); u++;
			ul = sc_ld2(ix.uni + u);
		}
	const uint64_t rp = sc_ld64(ix.ref_pos + ul.x);
	L.ep = (rp & 0xFFFFFFFFFFull) + uni_offset;                  // t_off lives in L.ep
	L.u_off = uni_offset; L.uni = (int32_t)u; L.uni_len = ul.y;
	if (ul.y < MIN_UNI_L) { map_done(L, 0); return; }          // cly.c:767 / 791
	if (L.after_loc == AL_PRE) { L.l_pre = SC_MIN(L.l_pre, L.u_off); go(L, ST_FLANK, FK_PRE); }
	else suf_begin(L);
}

// ---------------------------------------------------------------- RP: the reference positions of the unitig (cly.c:840-937)
SC_HD void rp_next(SeedLane &L)
{
	L.c_r_p++;
	if (L.c_r_p >= L.r_p_e) map_done(L, L.max_s); else go(L, ST_RP, 0);
}
SC_HD void rp_score(const SeedEnv &E, SeedLane &L)
{   // cly.c:914-920
	L.am_score = (int16_t)(Q_MEM_at(E.ix, L.am_mtch_len) + Q_LV_at(E.ix, L.am_le, L.am_ll) + Q_LV_at(E.ix, L.am_re, L.am_rl));
	if (L.am_score < 20) rp_next(L); else L.push = 1;
}
SC_HD void new_ed_begin(SeedLane &L, uint32_t right)
{   // get_new_ed (cly.c:629-662): q_buff / t_buff take the t_pre / t_suf slots of the frame and start zeroed (oracle policy P2)
	L.B = 0; L.C = 0; L.fl_ext = 0;
	if (right) {
		const int32_t q2 = L.q_off + (int32_t)L.l_m + 1;
		L.fl_q = q2; L.fl_t = L.rp_global + L.u_off + L.l_m; L.fl_max = L.read_len - (uint32_t)q2;
	} else {
		const int32_t qo = SC_MAX(L.q_off, 0);
		L.fl_q = qo; L.fl_t = L.rp_global + L.u_off - 1; L.fl_max = (uint32_t)qo;
	}
	count_getref(L, SC_MIN(12u, L.fl_max));
	go(L, ST_FLANK, right ? FK_NEWR : FK_NEWL);
}
SC_HD void h_rp(const SeedEnv &E, SeedLane &L)
{
	const DevIndex &ix = E.ix;
	if (L.kind == 1) {
		// first turn of a map_seed that passed both flanks (cly.c:840-888)
		L.am_mtch_len = L.l_m & 0xffff; L.am_score = (int16_t)L.s;
		L.am_ll = L.l_pre & 0xff; L.am_le = L.d_pre & 0xff; L.am_rl = L.l_suf & 0xff; L.am_re = L.d_suf & 0xff;
		const uint32_t r_p_s = sc_ld2(ix.uni + L.uni).x, r_p_e = sc_ld2(ix.uni + L.uni + 1).x;
		L.rsl = (L.l_pre < LV_L || L.d_pre == 0) ? 1 : 0;      // an edit distance of 0: the extension is not over
		L.rsr = (L.l_suf < LV_L || L.d_suf == 0) ? 1 : 0;
		const int64_t n = (int64_t)r_p_e - (int64_t)r_p_s;
		if (n > 50 && !(n < 1000)) { map_done(L, 50); return; }
		L.c_r_p = r_p_s; L.r_p_e = r_p_e;
		if (r_p_s >= r_p_e) { map_done(L, L.max_s); return; }
		L.kind = 0;
	}
	const uint64_t rp = sc_ld64(ix.ref_pos + L.c_r_p);
	L.rp_global = rp & 0xFFFFFFFFFFull; L.rp_ref = (uint32_t)((rp >> 40) & 0x7FFFFF);
	L.ext_l = 0;
	if (L.rsl | L.rsr) {
		if (!L.rsl) L.am_mtch_len = L.l_m & 0xffff;
		new_ed_begin(L, L.rsl ? 0u : 1u);
		return;
	}
	L.push = 1;
}
// the pending anchor of a lane, after the collective part of the push gave it a staging slot
SC_HD uint4 push_make(const SeedEnv &E, SeedLane &L)
{
	L.max_s = SC_MAX(L.max_s, L.am_score);
	L.top_score = SC_MAX(L.top_score, L.am_score);
	const uint64_t g = L.rp_global + L.u_off - L.ext_l;
#if SC_DEVICE
	const uint64_t seq_offset = __ldg(&E.ix.ref_info[L.rp_ref].y);
#else
	const uint64_t seq_offset = E.ix.ref_info[L.rp_ref].y;
#endif
	uint4 a;
	a.x = L.rp_ref;
	a.y = (uint32_t)(g - seq_offset);
	a.z = (uint32_t)(L.q_off + 1 - (int32_t)L.ext_l);
	a.w = (L.am_mtch_len & 0xffff) | ((uint32_t)(uint16_t)(int16_t)L.am_score << 16);
	return a;
}

// ---------------------------------------------------------------- what follows the Landau-Vishkin of a flank
SC_HD void flank_done(const SeedEnv &E, SeedLane &L, uint32_t len, int32_t ed)
{
	const DevIndex &ix = E.ix;
	const uint32_t fk = L.kind;
	if (fk == FK_PRE) {
		// cly.c:773-795
		L.d_pre = (uint32_t)ed;
		L.s = Q_MEM_at(ix, L.l_m) + Q_LV_at(ix, L.d_pre, L.l_pre);
		if (L.s < 12 && L.l_pre == LV_L && L.uni < 0) { map_done(L, 0); return; }
		if (L.uni < 0) {
			if ((L.sp & SA_MASK) == 0) { L.loc_row = L.sp; L.loc_l = L.s_l; L.after_loc = AL_SUF; go(L, ST_LOCATE, 0); }
			else go(L, ST_OCC, OK_WALK2);
			return;
		}
		suf_begin(L);
		return;
	}
	if (fk == FK_SUF) {
		// cly.c:826-842
		L.l_suf = len; L.d_suf = (uint32_t)ed;
		L.s = Q_MEM_at(ix, L.l_m) + Q_LV_at(ix, L.d_pre, L.l_pre) + Q_LV_at(ix, L.d_suf, L.l_suf);
		if ((L.s <= 20 && L.l_suf == LV_L) || !(L.s > 0)) { map_done(L, 0); return; }
		go(L, ST_RP, 1);
		return;
	}
	if (fk == FK_NEWL) {
		L.am_ll = len & 0xff; L.am_le = (uint32_t)ed & 0xff; L.ext_l = L.fl_ext;
		L.am_mtch_len = (L.l_m + L.ext_l) & 0xffff;
		if (L.rsr) { new_ed_begin(L, 1); return; }
		rp_score(E, L);
		return;
	}
	L.am_rl = len & 0xff; L.am_re = (uint32_t)ed & 0xff;
	L.am_mtch_len = (L.am_mtch_len + L.fl_ext) & 0xffff;
	rp_score(E, L);
}

// ---------------------------------------------------------------- FLANK: windows, exact-match extension, start of the Landau-Vishkin
// Landau-Vishkin state of a lane between its row turns: the two extended strings and the mn / ed rows, in shared memory
SC_HD uint64_t lvs_ld(const LaneMem &M, uint32_t k)
{
#if SC_DEVICE
	uint64_t v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(M.lvs + 256u * k) : "memory"); return v;
#elif defined(__CUDACC__)
	return 0 * M.lvs * k;
#else
	return M.lvs[k];
#endif
}
SC_HD void lvs_st(const LaneMem &M, uint32_t k, uint64_t v)
{
#if SC_DEVICE
	asm volatile("st.shared.u64 [%0], %1;" :: "r"(M.lvs + 256u * k), "l"(v) : "memory");
#elif defined(__CUDACC__)
	(void)M; (void)k; (void)v;
#else
	M.lvs[k] = v;
#endif
}

SC_HD void h_flank(const SeedEnv &E, SeedLane &L, const LaneMem &M)
{
	const DevIndex &ix = E.ix;
	const uint32_t fk = L.kind;
	const bool pre = fk == FK_PRE;
	const uint32_t left = (fk == FK_PRE || fk == FK_NEWL) ? 1u : 0u;
	// the reference window (get_ref, cly.c:435-466: a negative offset is clamped to 0) and the read window
	uint64_t t;
	if (pre) { const int64_t o = (int64_t)L.ep - 1; t = o < 0 ? 0 : (uint64_t)o; }
	else if (fk == FK_SUF) t = L.ep + L.l_m;
	else { const int64_t o = (int64_t)L.fl_t; t = o < 0 ? 0 : (uint64_t)o; }
	uint64_t rw = 0, qraw = 0;
	if (!pre || L.uni >= 0) rw = ref_window(ix, t, left);
	if (!pre) qraw = strand_window(E, L, left ? (int64_t)L.fl_q : (int64_t)L.fl_q - 5, left);   // rightwards: 5 bases in front of the cursor, what negative indices of the query read
	uint64_t er, eq; uint32_t len;
	if (pre) {
		// left flank of map_seed (cly.c:764-772): the reference bases left of the match when the unitig is known, else what the walk collected
		if (L.uni >= 0) { L.B = frame_merge(L.B, rw, L.l_pre); count_getref(L, L.l_pre); }
		len = L.l_pre;
		er = frame_ext(frame_tail(L.A), L.B); eq = frame_ext(0, L.A);
	} else {
		// exact-match extension in strides of <= 12 (cly.c:806-823 / 663-689), then Landau-Vishkin on what follows
		const uint64_t qw = left ? qraw : (qraw << 10);
		const int cap = left ? 32 : 27;
		int d = 0;
		for (;;) {
			len = SC_MIN(L.fl_max, 12u);
			if (fk == FK_NEWL) L.B = frame_merge(L.B, qw << (2 * d), len);
			if (d + 12 > cap) {                                        // the windows are used up: fetch again at the new cursors
				if (left) { L.fl_q -= d; L.fl_t -= d; } else { L.fl_q += d; L.fl_t += d; }
				return;
			}
			const uint32_t mtc = SC_MIN((uint32_t)sc_common(rw << (2 * d), qw << (2 * d)), len);
			if (mtc == 0) break;
			if (fk == FK_SUF) L.l_m += mtc; else L.fl_ext += mtc;
			L.fl_max -= mtc;
			d += (int)mtc;
			count_getref(L, SC_MIN(L.fl_max, 12u));
		}
		L.C = frame_merge(L.C, rw << (2 * d), len);
		er = frame_ext(frame_tail(L.B), L.C);
		eq = left ? frame_ext(frame_tail(L.A), L.B) : (qraw << (2 * d));
	}
	if (L.B & FR_POISON) {                                         // t_pre holds '$' (a walk off the start of unitig 0): the byte statement
		uint8_t r[18], q[18];
		for (int k = 0; k < 18; k++) { r[k] = (uint8_t)((er >> (62 - 2 * k)) & 3); q[k] = (uint8_t)((eq >> (62 - 2 * k)) & 3); }
		const uint32_t pe = L.B & 15;
		if (pre) r[5 + pe] = 5; else if (pe >= 8 && pe <= 12) r[pe - 8] = 5;
		flank_done(E, L, len, lv_bytes(r, q, (int32_t)len));
		return;
	}
	lvs_st(M, 0, er); lvs_st(M, 1, eq); lvs_st(M, 2, 1ull << 44); lvs_st(M, 3, 0x054321012345ull);
	L.lv_st = (len << 8) | (len << 16);                            // row 0, best_score = len
	L.st = ST_LV;
}

// ---------------------------------------------------------------- LV: one row of the Landau-Vishkin (lv_extd, cly.c:544-607)
// Lanes in different rows run together (row i visits the diagonals j = -i .. 4).
SC_HD void h_lv(const SeedEnv &E, SeedLane &L, const LaneMem &M)
{
	const uint64_t er = lvs_ld(M, 0), eq = lvs_ld(M, 1);
	uint64_t MN = lvs_ld(M, 2), ED = lvs_ld(M, 3);
	int i = (int)(L.lv_st & 15), best_score = (int)((L.lv_st >> 8) & 0xff);
	const int len = (int)((L.lv_st >> 16) & 0xff);
	int prev_mn = -1, cur_mn = i - 1, next_mn = lv_get(MN, -i + 1) - 1;
	int prev_ed = i + 1, cur_ed = i, next_ed = lv_get(ED, -i + 1);
	bool fin = false;
	#pragma unroll 1
	for (int j = -i; j <= 4 && !fin; j++) {
		int m, e;
		if (cur_mn + j < len - 1) {
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
		ED = lv_set(ED, j, e);
		int mn_j = SC_MIN(m, len);
		mn_j = SC_MIN(mn_j, len - j);
		{	// in-line match along diagonal j: reference index mn_j + j against query index mn_j, up to the sentinels
			const int a = mn_j + j;
			int run = sc_common(er << (2 * (a + 5)), eq << (2 * (mn_j + 5)));
			run = SC_MIN(run, SC_MIN(len - a, len - mn_j));
			mn_j += run;
		}
		MN = lv_set(MN, j, mn_j + 1);
		if (mn_j == len || mn_j + j == len) {
			best_score = SC_MIN(e - 1, best_score);
			if (j <= i + 1) fin = true;
		}
		prev_mn = cur_mn; cur_mn = next_mn; next_mn = lv_get(MN, j + 2) - 1;
		prev_ed = cur_ed; cur_ed = next_ed; next_ed = lv_get(ED, j + 2);
	}
	if (!fin && ++i > 4) fin = true;
	if (fin) { L.st = ST_FLANK; flank_done(E, L, (uint32_t)len, best_score); return; }
	lvs_st(M, 2, MN); lvs_st(M, 3, ED);
	L.lv_st = (uint32_t)i | ((uint32_t)best_score << 8) | ((uint32_t)len << 16);
}

// ---------------------------------------------------------------- which state runs this turn
// Every lane computes a key from its state and the number of lanes k sharing it; the state of the largest key runs.  The
// fullest state wins, so a state with few lanes waits while a fuller one exists and parked lanes pile up until their state is
// worth a turn.  Free lanes (FETCH) are refilled as soon as fetch_min of them wait, or when nothing else is left to do.
// policy 3: the light states (a few dozen instructions: CTRL, LOCATE, RP) also go first once fetch_min lanes wait in them.
#ifndef SC_FETCH_MIN
#define SC_FETCH_MIN 4
#endif
#ifndef SC_POLICY
#define SC_POLICY 2
#endif
SC_HD uint32_t vote_key(uint32_t st, int k, int policy, int fetch_min)
{
	if (st == ST_DEAD) return 0;
	uint32_t prio = (uint32_t)k;                                   // 1..32
	if (st == ST_FETCH) prio = (k >= fetch_min) ? 100u : 0u;
	else if (policy >= 3 && (st == ST_CTRL || st == ST_LOCATE || st == ST_RP) && k >= fetch_min) prio = 64u + (uint32_t)k;
	return (prio << 3) | (7u - st) | 0x800u;
}
SC_HD uint32_t vote_state(uint32_t best_key) { return best_key ? 7u - (best_key & 7u) : (uint32_t)ST_DEAD; }
