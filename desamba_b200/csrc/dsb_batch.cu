// dsb_batch.cu -- the batch pipeline behind dsb_classify_batch():
//   K0 k_encode_probe : ASCII -> 2-bit codes of both strands (bytes for the scoring kernels, packed words for seeding); every
//                       l_ek-mer of the read (both strands) is hashed and probed in the two exist-k-mer bit tables
//                       (get_exist_kmer, cly.c:956-972) -> membership bit-vectors
//   K1 k_islands      : replays the reference's island scan + top labelling on the bit-vectors (cly.c:1071-1234) -> seeds,
//                       and the seed-task list of the fast pass
//   K2 k_seed x3      : lane per island seed over a flat task list (dsb_seedcore.h): FM-index search, locate, flanks -> anchors
//      k_chain x3     : warp per read: ordered gather of the anchors, resolve_tree, which pass comes next (cly.c:3098-3127)
//      k_score, k_score_heavy : warp / CTA per read: 9-mer sparse-DP scoring (dsb_classify.cuh)
//   K3 k_finalize     : class filter (needs the running max_read_l of the input order), final sort, primary detection
// One CUDA stream per context, no host synchronisation between the kernels.
#include "dsb_internal.h"
#include "dsb_classify.cuh"
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <mutex>
#include <condition_variable>
#include <thread>

#define PROBE_TILE 1024
#define PROBE_THREADS 256
#define N_BITVEC 5                 // E_fwd, E_rev, T0_fwd, T0_rev, Z (k-mer not masked as low-complexity)
#ifndef CLASSIFY_WARPS_PER_BLOCK
#define CLASSIFY_WARPS_PER_BLOCK 2      // small blocks release their SM share early in the tail of a launch (batches in flight overlap better)
#endif
#ifndef HEAVY_BLOCKS
#define HEAVY_BLOCKS 24               // CTAs of k_score_heavy (one heavy read at a time each; a handful of reads per batch)
#endif

static inline uint32_t bits_words(uint32_t len) { return (len + 31) / 32 + 1; }
static inline uint32_t seed_slots(uint32_t len) { return len / 2 + 2; }

// ------------------------------------------------------------------------------------------------ K0
struct ProbeParams {
	DevIndex ix;
	const char *seqs; const uint64_t *read_off, *bin_off, *bits_off;
	const uint2 *tiles;
	const uint8_t *hdr7;       // per read: the byte 7 before the forward strand (oracle policy P3: the reference's malloc chunk header)
	uint8_t *bin; uint32_t *bits;
	uint64_t *pk;              // 2-bit packed forward strands for the seeding engine (layout: SeedEnv.pk, dsb_seedcore.h)
};

__device__ __forceinline__ uint32_t cly_bit(char ch)      // CLY_Bit, cly.c:17-35: A0 C1 G2 T3 (either case), everything else 1
{
	switch (ch) { case 'A': case 'a': return 0; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 1; }
}

__device__ __forceinline__ uint64_t revcomp_kmer(uint64_t km, int l)
{
	uint64_t x = __brevll(~km);
	x = ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);
	return x >> (64 - 2 * l);
}

__global__ void __launch_bounds__(PROBE_THREADS) k_encode_probe(const __grid_constant__ ProbeParams P)
{
	__shared__ __align__(16) uint8_t s_code[PROBE_TILE + 32];
	const uint2 tile = P.tiles[blockIdx.x];
	const uint32_t r = tile.x, start = tile.y;
	const uint64_t off = P.read_off[r];
	const uint32_t len = (uint32_t)(P.read_off[r + 1] - off);
	const int l_ek = P.ix.l_ek;
	const uint32_t n_kmer = len - l_ek + 1;
	uint8_t *bin_F = P.bin + P.bin_off[r] + DSB_GUARD, *bin_R = bin_F + len;
	const uint32_t nb = min((uint32_t)(PROBE_TILE + l_ek - 1), len - start);
	const int tid = threadIdx.x;
	for (uint32_t k = tid; k < nb; k += PROBE_THREADS) {
		const uint32_t code = cly_bit(P.seqs[off + start + k]);
		s_code[k] = (uint8_t)code;
		if (k < PROBE_TILE) { bin_F[start + k] = (uint8_t)code; bin_R[len - 1 - (start + k)] = (uint8_t)(3 - code); }   // cly.c:1250-1259
	}
	if (start == 0 && tid < DSB_GUARD) {                          // out-of-buffer policy P3
		bin_F[tid - DSB_GUARD] = (tid == DSB_GUARD - 7) ? P.hdr7[r] : (tid == DSB_GUARD - 8) ? (uint8_t)0xff : (uint8_t)0;
		bin_R[len + tid] = 0;
	}
	__syncthreads();
	if (tid < 32 && start + 32 * tid < len) {                     // 32 packed words of 32 bases, first base in the top bits
		const uint32_t *w4 = (const uint32_t *)(s_code + 32 * tid);
		uint64_t v = 0;
		#pragma unroll
		for (int k = 0; k < 8; k++) v = (v << 8) | (((w4[k] & 0x03030303u) * 0x40100401u) >> 24);
		const uint32_t valid = min(32u, len - (start + 32 * tid));
		if (valid < 32) v &= ~0ull << (64 - 2 * valid);
		P.pk[P.bits_off[r] / N_BITVEC + 1 + (start >> 5) + tid] = v;
	}
	const uint32_t W = (len + 31) / 32 + 1;
	uint32_t *bits = P.bits + P.bits_off[r];
	const int sbm = P.ix.single_base_max;
	const uint64_t mask = P.ix.ek_mask;
	#pragma unroll
	for (int it = 0; it < PROBE_TILE / PROBE_THREADS; it++) {
		const uint32_t p = it * PROBE_THREADS + tid, i = start + p;
		const bool valid = i < n_kmer;
		uint64_t km = 0; uint32_t cnt = 0;
		if (valid)
			for (int b = 0; b < l_ek; b++) { const uint32_t c = s_code[p + b]; km = (km << 2) | c; cnt += 1u << (c * 8); }
		// store_kmers (cly.c:360-398): a k-mer with any base count >= single_base_max is stored as 0 and never exists
		const bool ok = valid && (int)(cnt & 0xff) < sbm && (int)((cnt >> 8) & 0xff) < sbm && (int)((cnt >> 16) & 0xff) < sbm && (int)(cnt >> 24) < sbm;
		const uint64_t kr = revcomp_kmer(km, l_ek);
		uint32_t t0f = 0, t0r = 0, ef = 0, er = 0;
		if (ok) {
			const uint64_t hf = dsb_hash64_1(km) & mask, hr = dsb_hash64_1(kr) & mask;
			if (P.ix.ek0_sum) {                            // summary bit (L2) first; the table byte only where it is set
				const uint64_t yf = hf >> 3, yr = hr >> 3;
				const uint32_t sf = (__ldg(P.ix.ek0_sum + (yf >> 5)) >> (yf & 31)) & 1, sr = (__ldg(P.ix.ek0_sum + (yr >> 5)) >> (yr & 31)) & 1;
				if (sf) t0f = (__ldg(P.ix.ek0 + yf) >> (7 - (hf & 7))) & 1;
				if (sr) t0r = (__ldg(P.ix.ek0 + yr) >> (7 - (hr & 7))) & 1;
			} else {
				const uint32_t bf = __ldg(P.ix.ek0 + (hf >> 3)), br = __ldg(P.ix.ek0 + (hr >> 3));
				t0f = (bf >> (7 - (hf & 7))) & 1; t0r = (br >> (7 - (hr & 7))) & 1;
			}
			if (t0f) { const uint64_t h2 = dsb_hash64_2(km) & mask; ef = (__ldg(P.ix.ek1 + (h2 >> 3)) >> (7 - (h2 & 7))) & 1; }
			if (t0r) { const uint64_t h2 = dsb_hash64_2(kr) & mask; er = (__ldg(P.ix.ek1 + (h2 >> 3)) >> (7 - (h2 & 7))) & 1; }
		}
		const uint32_t m_ef = __ballot_sync(DSB_FULL, ef), m_er = __ballot_sync(DSB_FULL, er);
		const uint32_t m_t0f = __ballot_sync(DSB_FULL, t0f), m_t0r = __ballot_sync(DSB_FULL, t0r), m_z = __ballot_sync(DSB_FULL, ok);
		const uint32_t w = i >> 5;                     // start and p are multiples of 32 per warp
		if ((tid & 31) == 0 && w < W) {
			bits[w] = m_ef; bits[W + w] = m_er; bits[2 * W + w] = m_t0f; bits[3 * W + w] = m_t0r; bits[4 * W + w] = m_z;
		}
	}
}

// ------------------------------------------------------------------------------------------------ K1
struct IslandParams {
	int l_ek; uint32_t n_reads;
	const uint64_t *read_off, *bits_off; const uint32_t *seed_off;
	const uint32_t *bits;
	dsb_seed *seeds[2]; uint32_t *n_seeds[2]; uint32_t *total_score[2];
	unsigned long long *counters;
	SeedTaskRef *tasks; uint32_t task_cap; uint32_t *ctl; uint32_t *task_first[2], *task_cnt[2];
};

__device__ __forceinline__ uint32_t bit_at(const uint32_t *v, uint32_t i) { return (__ldg(v + (i >> 5)) >> (i & 31)) & 1; }

// search_exist_kmer_M2 + get_seed_vector_M2 (cly.c:1071-1234).  Both strands run the SAME scan on a bit-vector indexed
// by forward k-mer position: the reference's right-to-left scan of the reverse strand is the mirror image of its
// left-to-right scan of the forward strand; only the reported offset is mirrored back.
__global__ void __launch_bounds__(128) k_islands(const __grid_constant__ IslandParams P)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t r = t >> 1, s = t & 1;
	uint32_t calls_nz = 0, calls_t0 = 0;
	if (r < P.n_reads) {
		const uint32_t len = (uint32_t)(P.read_off[r + 1] - P.read_off[r]);
		uint32_t n_seed = 0, total = 0;
		if (len >= 40) {
			const uint32_t n = len - P.l_ek + 1, W = (len + 31) / 32 + 1;
			const uint32_t *E = P.bits + P.bits_off[r] + s * W, *T0 = E + 2 * W, *Z = P.bits + P.bits_off[r] + 4 * W;
			dsb_seed *out = P.seeds[s] + P.seed_off[r];
			uint32_t max_index = 0, max_length = 0, index_end = 100;
			#define PROBE(idx) (calls_nz += bit_at(Z, (idx)), calls_t0 += bit_at(T0, (idx)), bit_at(E, (idx)))
			for (uint32_t i = 2; i < n; i += 3) {
				{   // the probes i, i+3, ... that fall into the current 32-bit word and miss are taken in one step
					const uint32_t w = i >> 5, b = i & 31, left = n - (w << 5);
					uint32_t pm = ((b % 3 == 0) ? 0x49249249u : (b % 3 == 1) ? 0x92492492u : 0x24924924u) & (0xffffffffu << b);
					if (left < 32) pm &= (1u << left) - 1;
					const uint32_t e = __ldg(E + w) & pm;
					const uint32_t miss = e ? (pm & ((1u << (__ffs(e) - 1)) - 1)) : pm;
					if (miss) { calls_nz += __popc(__ldg(Z + w) & miss); calls_t0 += __popc(__ldg(T0 + w) & miss); }
					if (!e) { i = (w << 5) + (31 - __clz(pm)); continue; }
					i = (w << 5) + __ffs(e) - 1;
				}
				if (!PROBE(i)) continue;
				uint32_t offset = i, l = 1;
				for (int j = 1; j < 3; ++j) { if (PROBE(i - j)) { offset--; l++; } else break; }
				{   // right extension (cly.c:1118-1130) a word at a time: the run of existing k-mers behind i, cut at l = 61 and at n;
					// every probed position counts for the probe counters -- the hits, and the miss that ends the run if there is one
					const uint32_t want = 61 - l;                          // hits after which `l > 60` stops the loop
					uint32_t run = 0;
					for (uint32_t p = i + 1; p < n && run < want;) {
						const uint32_t b = p & 31, avail = min(32 - b, n - p);
						const uint32_t x = ~(__ldg(E + (p >> 5)) >> b);
						const uint32_t ones = x ? min((uint32_t)(__ffs(x) - 1), avail) : avail;
						run += ones; p += ones;
						if (ones < avail) break;
					}
					const uint32_t hits = min(run, want);
					const uint32_t probes = hits + ((hits == run && i + run + 1 < n && run < want) ? 1u : 0u);
					for (uint32_t p = i + 1, left = probes; left;) {        // probe counters over [i + 1, i + probes]
						const uint32_t b = p & 31, take = min(32 - b, left);
						const uint32_t m = ((take == 32) ? 0xffffffffu : ((1u << take) - 1)) << b;
						calls_nz += __popc(__ldg(Z + (p >> 5)) & m); calls_t0 += __popc(__ldg(T0 + (p >> 5)) & m);
						p += take; left -= take;
					}
					l += hits;
				}
				dsb_seed sd; sd.offset = s ? (n - offset - l) : offset; sd.len = (uint16_t)l; sd.top = 0; sd.pad = 0;
				out[n_seed] = sd;
				// top labelling (cly.c:1174-1226); the window position is the mirrored offset for the reverse strand
				if (offset < index_end) { if (max_length < l) { max_length = l; max_index = n_seed; } out[max_index].top = 0; }   // (un-marks seed 0 when it was marked as its own predecessor)
				else { out[max_index].top = 1; index_end += 100; total += max_length; max_index = n_seed; max_length = l; }
				n_seed++;
				i = offset + l;
			}
			#undef PROBE
			if (n_seed) out[max_index].top = 1;
			total += max_length;
		}
		P.n_seeds[s][r] = n_seed; P.total_score[s][r] = total;
	}
	{	// seed tasks of the fast pass: the top seeds of the strand with the larger total score, of both strands when the scores are
		// close (cly.c:1261-1266, 3095-3099, 1494-1496).  Threads 2r and 2r + 1 are neighbours in the warp.
		uint32_t n_seed = 0, total = 0;
		if (r < P.n_reads) { n_seed = P.n_seeds[s][r]; total = P.total_score[s][r]; }
		const uint32_t other = __shfl_xor_sync(DSB_FULL, total, 1);
		if (r < P.n_reads) {
			const bool first = s ? (total > other) : (total >= other);                    // search_dir[0] after the swap
			const uint32_t t0 = first ? total : other, t1 = first ? other : total;
			const bool both = (t0 - t1) <= (t0 >> 3);
			const dsb_seed *sv = P.seeds[s] + P.seed_off[r];
			uint32_t n_top = 0;
			if (first || both) for (uint32_t k = 0; k < n_seed; k++) n_top += sv[k].top ? 1u : 0u;
			uint32_t base = 0;
			if (n_top) {
				base = atomicAdd(P.ctl + CTL_TASK_N + PASS_FAST, n_top);
				if ((uint64_t)base + n_top > P.task_cap) { atomicOr(P.ctl + CTL_OVERFLOW, OVF_TASKS); n_top = 0; }
			}
			P.task_first[s][r] = base; P.task_cnt[s][r] = n_top;
			if (n_top) {
				SeedTaskRef *out = P.tasks + base;
				for (uint32_t k = 0, at = 0; k < n_seed; k++) if (sv[k].top) { SeedTaskRef t; t.read = r; t.sk = (s << 31) | k; out[at++] = t; }
			}
		}
	}
	for (int d = 16; d; d >>= 1) { calls_nz += __shfl_xor_sync(DSB_FULL, calls_nz, d); calls_t0 += __shfl_xor_sync(DSB_FULL, calls_t0, d); }
	if ((threadIdx.x & 31) == 0 && (calls_nz | calls_t0)) {
		atomicAdd(P.counters + DSB_CNT_N_BIT0, (unsigned long long)calls_nz);
		atomicAdd(P.counters + DSB_CNT_N_BIT1, (unsigned long long)calls_t0);
	}
}

// ------------------------------------------------------------------------------------------------ K2
struct ScratchLayout { uint64_t anc_tmp, chain, chain_tmp, sms, sms_tmp, cand, sort_key[2], sort_idx[2], score_v, sc_hash, total; };
static ScratchLayout scratch_layout(uint32_t max_anchors, uint32_t max_matches)
{
	ScratchLayout L; uint64_t o = 0;
	auto take = [&](uint64_t bytes) { uint64_t at = o; o += (bytes + 127) & ~127ull; return at; };
	L.anc_tmp = take((uint64_t)max_anchors * sizeof(DevAnchor));
	L.chain = take((uint64_t)max_anchors * sizeof(DevChain)); L.chain_tmp = take((uint64_t)max_anchors * sizeof(DevChain));
	L.sms = take((uint64_t)max_matches * sizeof(DevSms));
	L.score_v = take(1024 * sizeof(int));
	L.sc_hash = take((256 + 2 * 400 + 8) * sizeof(ScHash));
	L.sms_tmp = take((uint64_t)max_matches * sizeof(DevSms)); L.cand = take((uint64_t)CAND_CAP * sizeof(uint2));
	for (int s = 0; s < 2; s++) { L.sort_key[s] = take((uint64_t)max_matches * 8); L.sort_idx[s] = take((uint64_t)max_matches * 4); }
	L.total = o;
	return L;
}

struct ClassifyLaunch { ClassifyParams P; ScratchLayout L; };

// Persistent phase kernels: every warp pulls read ids from a work list (chaining after the fast pass: all reads, longest first;
// later phases: the lists the chain kernel fills) and runs one phase of classify_seq for that read in its own scratch.
__device__ __forceinline__ void scratch_setup(const ClassifyLaunch &A, ReadState &S, uint32_t slot)
{
	uint8_t *base = A.P.scratch + (uint64_t)slot * A.P.scratch_stride;
	S.ix = &A.P.ix;
	S.ws.anc = nullptr; S.ws.anc_tmp = (DevAnchor *)(base + A.L.anc_tmp);
	S.ws.chain = (DevChain *)(base + A.L.chain); S.ws.chain_tmp = (DevChain *)(base + A.L.chain_tmp);
	S.ws.sms = (DevSms *)(base + A.L.sms); S.ws.score_v = (int *)(base + A.L.score_v);
	S.ws.sc_hash = (ScHash *)(base + A.L.sc_hash);
	S.ws.sms_tmp = (DevSms *)(base + A.L.sms_tmp); S.ws.cand = (uint2 *)(base + A.L.cand);
	for (int s = 0; s < 2; s++) { S.ws.sort_key[s] = (uint64_t *)(base + A.L.sort_key[s]); S.ws.sort_idx[s] = (uint32_t *)(base + A.L.sort_idx[s]); }
	S.max_anchors = A.P.max_anchors; S.max_matches = A.P.max_matches;
	S.team = nullptr; S.mt = nullptr; S.l_read = 0;
}
__device__ __forceinline__ void warp_setup(const ClassifyLaunch &A, ReadState &S, uint8_t *smem_raw)
{
	const int warp = threadIdx.x >> 5;
	S.sm = (WarpSmem *)smem_raw + warp;
	scratch_setup(A, S, blockIdx.x * CLASSIFY_WARPS_PER_BLOCK + warp);
}
// next read of the launch's work list, or 0xffffffff; list < 0: all reads in `order`
__device__ __forceinline__ uint32_t next_read(const ClassifyParams &P, int list, int cursor)
{
	uint32_t i = 0;
	if (lane_id() == 0) i = atomicAdd(P.ctl + CTL_CURSOR + cursor, 1u);
	i = __shfl_sync(DSB_FULL, i, 0);
	const uint32_t n = (list < 0) ? P.n_reads : P.ctl[CTL_LIST_N + list];
	if (i >= n) return 0xffffffffu;
	return (list < 0) ? P.order[i] : P.list[list][i];
}

#ifndef CLASSIFY_WARPS_PER_SM
#define CLASSIFY_WARPS_PER_SM 16          // resident classify warps per SM the phase kernels are compiled for (register budget 65536 / (32 * this))
#endif
#ifndef SEED_WARPS_PER_SM
#define SEED_WARPS_PER_SM 16              // resident warps per SM of k_seed (register budget; 8 KB of shared memory per warp for the visited-row sets)
#endif
#define SCORE_MIN_BLOCKS (CLASSIFY_WARPS_PER_SM / CLASSIFY_WARPS_PER_BLOCK)
// lane per island seed over the flat task list of the pass (dsb_seed.cuh)
__global__ void __launch_bounds__(SEED_WARPS_PER_BLOCK * 32, SEED_WARPS_PER_SM / SEED_WARPS_PER_BLOCK) k_seed(const __grid_constant__ SeedPassParams P)
{
	__shared__ uint32_t s_vis1[SEED_WARPS_PER_BLOCK][VIS1_SLOTS][32];
	__shared__ uint64_t s_lvs[SEED_WARPS_PER_BLOCK][4][32];
	if (P.ctl[CTL_OVERFLOW]) return;                      // a pool of an earlier kernel overflowed: the batch is run again with larger pools (dsb_batch_download)
	seed_warp_loop(P, s_vis1[threadIdx.x >> 5], s_lvs[threadIdx.x >> 5]);
}

__global__ void __launch_bounds__(CLASSIFY_WARPS_PER_BLOCK * 32, SCORE_MIN_BLOCKS) k_chain(const __grid_constant__ ClassifyLaunch A, int pass, int list, int cursor)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	ReadState S;
	if (A.P.ctl[CTL_OVERFLOW]) return;
	warp_setup(A, S, smem_raw);
	for (uint32_t r; (r = next_read(A.P, list, cursor)) != 0xffffffffu;) phase_chain(A.P, S, r, pass);
}

__global__ void __launch_bounds__(CLASSIFY_WARPS_PER_BLOCK * 32, SCORE_MIN_BLOCKS) k_score(const __grid_constant__ ClassifyLaunch A, int list, int cursor)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	ReadState S;
	if (A.P.ctl[CTL_OVERFLOW]) return;
	warp_setup(A, S, smem_raw);
	S.mt = (MatchSmem *)(smem_raw + CLASSIFY_WARPS_PER_BLOCK * sizeof(WarpSmem)) + (threadIdx.x >> 5);
	// the list is in no particular order: reads with many anchors or many bases (the expensive ones) are scored first
	for (int sweep = 0; sweep < 2; sweep++)
		for (uint32_t r; (r = next_read(A.P, list, cursor + 12 * sweep)) != 0xffffffffu;) {
			const bool is_big = A.P.work[r].n_anc > 200 || (A.P.read_off[r + 1] - A.P.read_off[r]) > 16000;
			if (is_big == (sweep == 0)) phase_score<false>(A.P, S, r);
		}
}

// Reads with many anchors (repeats) carry the long tail of a batch: their sparse DP looks back over thousands of matches.
// They get a whole CTA each: warp 0 runs phase_score, the other warps serve its DP look-backs (dp_team_helper_loop).
__global__ void __launch_bounds__(TEAM_WARPS * 32) k_score_heavy(const __grid_constant__ ClassifyLaunch A, int list, int cursor, uint32_t slot0)
{
	__shared__ __align__(16) WarpSmem wsm;
	__shared__ __align__(16) MatchSmem msm;
	__shared__ DpTeam team;
	const int warp = threadIdx.x >> 5;
	if (A.P.ctl[CTL_OVERFLOW]) return;                    // (uniform over the CTA)
	if (warp == 0) {
		ReadState S;
		scratch_setup(A, S, slot0 + blockIdx.x);             // scratch slots of their own
		S.sm = &wsm; S.team = &team; S.mt = &msm;
		for (uint32_t r; (r = next_read(A.P, list, cursor)) != 0xffffffffu;) phase_score<true>(A.P, S, r);
		if (lane_id() == 0) team.cmd = -1;
		__syncthreads();                                 // releases the helpers
	} else
		dp_team_helper_loop(&team, warp);
}

// ------------------------------------------------------------------------------------------------ K3
struct FinalizeParams {
	uint32_t n_reads; int32_t max_read_l_in;
	int filter_min_length, filter_min_score, filter_min_score_LV3;
	dsb_read_result *rr; dsb_hit *hits;
	const unsigned long long *counters;
	const uint32_t *ctl;
};

struct HitCmpByMEMScore {           // chain_cmp_by_MEM_score (cly.c:54-64): asymmetric on ties, as written
	__device__ int operator()(const dsb_hit &a, const dsb_hit &b) const {
		const int score_a = (a.sum_score << 5), score_b = (b.sum_score << 5);
		if (score_a < score_b) return 1;
		if (score_a > score_b) return -1;
		return (a.sum_score % 2);
	}
};

#define FILTER_MIN_SCORE_SHROT_3G_READ 30
#define FILTER_MIN_SCORE_2G_READ 26
// tail of delete_small_score_rst (cly.c:2958-2993) + detect_primary (cly.c:2995-3058); one thread per read
__global__ void __launch_bounds__(128) k_finalize(const __grid_constant__ FinalizeParams P)
{
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= P.n_reads || P.ctl[CTL_OVERFLOW]) return;
	dsb_read_result rr = P.rr[r];
	uint32_t n = rr.n_hit;
	if (n == 0) return;
	dsb_hit *hit = P.hits + rr.hit_off, *tmp = hit + n;
	const uint32_t l_read = rr.read_len;
	// buff->max_read_l after this read (input order, -t 1 semantics): < 510 iff nothing >= 510 entered before or at r
	const bool max_small = P.max_read_l_in < 510 && (unsigned long long)r < P.counters[DSB_CNT_FIRST_LONG];
	for (uint32_t i = 0; i < n; i++) {
		dsb_hit c = hit[i];
		const int score = c.sum_score + ((c.q_ed - c.q_st) >> 5);
		bool kill;
		if (max_small) kill = score < FILTER_MIN_SCORE_2G_READ;
		else if (l_read < 310) kill = score < FILTER_MIN_SCORE_SHROT_3G_READ;
		else kill = score < (P.filter_min_score_LV3) && (c.q_ed - c.q_st < P.filter_min_length || score < P.filter_min_score);
		if (kill) { c.sum_score = 0; hit[i] = c; }
	}
	if (n > 1) glibc_msort(hit, tmp, (int)n, HitCmpByMEMScore());
	uint32_t k = 0;
	for (; k < n; k++) if (hit[k].sum_score == 0) break;
	n = k;
	rr.n_hit = n;
	P.rr[r] = rr;
	if (n == 0) return;
	// detect_primary
	int primary_v[400]; uint8_t primary_v_idx[400];      // n <= 400 (cly.c:2897), so the reference's 750 cap is never reached
	int n_primary_v = 1;
	primary_v[0] = 0; primary_v_idx[0] = 0;
	for (uint32_t i = 0; i < n; i++) if (hit[i].q_st > 4294960000u) hit[i].q_st = 0;
	{ dsb_hit h = hit[0]; h.pri_index = 0; h.primary = 1; hit[0] = h; }
	for (uint32_t ci = 1; ci < n; ci++) {
		dsb_hit c_hit = hit[ci];
		int overlap = 0;
		for (int i = 0; i < n_primary_v; i++) {
			const dsb_hit ph = hit[primary_v[i]];
			int primary_st, primary_ed;
			if (ph.direction == c_hit.direction) { primary_st = ph.q_st; primary_ed = ph.q_ed; }
			else { primary_st = l_read - ph.q_ed; primary_ed = l_read - ph.q_st; }
			const uint32_t overlap_st = DSB_MAX(c_hit.q_st, primary_st);
			const uint32_t overlap_ed = DSB_MIN(c_hit.q_ed, primary_ed);
			if ((overlap_st < overlap_ed) && (((overlap_ed - overlap_st) << 1) >= (c_hit.q_ed - c_hit.q_st))) overlap = 1;
			if (overlap) {
				c_hit.primary = 2;
				c_hit.pri_index = ++primary_v_idx[i];
				const int max_gap = DSB_MAX((ph.sum_score >> 6), 5);
				if (c_hit.sum_score + max_gap > ph.sum_score) c_hit.pri_index = 1;
				if (primary_v_idx[i] == 255) primary_v_idx[i] = 254;
				break;
			}
		}
		if (overlap == 0) {
			c_hit.primary = 3;
			c_hit.pri_index = primary_v_idx[n_primary_v] = 0;
			primary_v[n_primary_v++] = ci;
		}
		hit[ci] = c_hit;
	}
}

// ------------------------------------------------------------------------------------------------ host side
static int ensure(DevBuf &b, size_t bytes)
{
	if (bytes <= b.cap) return DSB_OK;
	if (b.p) cudaFree(b.p);
	b.p = nullptr; b.cap = 0;
	const size_t want = bytes + bytes / 4 + 256;
	DSB_CUDA(cudaMalloc(&b.p, want));
	b.cap = want;
	return DSB_OK;
}

extern "C" void dsb_opts_default(dsb_opts *o)
{
	if (!o) return;
	o->l_min_match = 170; o->min_score = 64;               // cly_mt.c:486
	o->max_anchors = 16384; o->max_matches = 16384; o->max_read_len = 1u << 20; o->warps_per_sm = CLASSIFY_WARPS_PER_SM; o->pool_scale_pct = 100;
}

extern "C" int dsb_ctx_create(dsb_index *ix, const dsb_opts *o, dsb_ctx **out)
{
	if (!ix || !out) { dsb_set_error("dsb_ctx_create: null argument"); return DSB_E_ARG; }
	*out = nullptr;
	DSB_CUDA(cudaSetDevice(ix->device));
	dsb_ctx *c = new dsb_ctx();
	c->ix = ix;
	if (o) c->opts = *o; else dsb_opts_default(&c->opts);
	if (c->opts.max_anchors < 1024) c->opts.max_anchors = 1024;
	if (c->opts.max_matches < 1024) c->opts.max_matches = 1024;
	if (c->opts.warps_per_sm < CLASSIFY_WARPS_PER_BLOCK) c->opts.warps_per_sm = CLASSIFY_WARPS_PER_BLOCK;
	if (c->opts.warps_per_sm > 32) c->opts.warps_per_sm = 32;
	c->stream = nullptr; c->copy_stream = nullptr; c->ran = false; c->launches = 0; c->h_stage = nullptr; for (int k = 0; k < 2 * DSB_STAGE_THREADS; k++) c->ev_stage[k] = nullptr;
	c->n_reads = 0; c->m_bin_read = 0; c->scratch_stride = 0; c->hits_cap = 0; c->task_cap = 0; c->n_chunks = 0; c->max_read_l_in = 0; c->retries = 0;
	for (int k = 0; k < 5; k++) c->grow[k] = 0;
	cudaDeviceProp prop;
	DSB_CUDA(cudaGetDeviceProperties(&prop, ix->device));
	c->n_sm = prop.multiProcessorCount;
	c->n_warps = c->n_sm * (int)(c->opts.warps_per_sm / CLASSIFY_WARPS_PER_BLOCK) * CLASSIFY_WARPS_PER_BLOCK;
	c->seed_blocks = c->n_sm * (SEED_WARPS_PER_SM / SEED_WARPS_PER_BLOCK);
	c->heavy_blocks = HEAVY_BLOCKS;
	if (const char *e = getenv("DSB_HEAVY_BLOCKS")) { const int v = atoi(e); if (v >= 1 && v <= 1024) c->heavy_blocks = v; }   // developer knob
	DSB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	DSB_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
	for (int k = 0; k < 2; k++) DSB_CUDA(cudaEventCreateWithFlags(&c->in[k].ev_up, cudaEventDisableTiming));
	for (int i = 0; i < DSB_N_EV; i++) DSB_CUDA(cudaEventCreateWithFlags(&c->ev[i], dsb_blocking_sync() ? cudaEventBlockingSync : cudaEventDefault));
	int rc = ensure(c->counters, DSB_CNT_COUNT * 8);
	if (rc != DSB_OK) { dsb_ctx_free(c); return rc; }
	*out = c;
	return DSB_OK;
}

extern "C" void dsb_ctx_free(dsb_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->ix->device);
	if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
	if (c->stream) cudaStreamSynchronize(c->stream);
	DevBuf *bufs[] = {&c->in[0].seqs, &c->in[1].seqs, &c->read_off, &c->bin_off, &c->bits_off, &c->seed_off, &c->tiles, &c->bin, &c->bits, &c->seeds[0], &c->seeds[1],
	                  &c->n_seeds[0], &c->n_seeds[1], &c->total_score[0], &c->total_score[1], &c->scratch, &c->rr, &c->hits, &c->counters, &c->prof, &c->work, &c->anc_pool, &c->chain_pool,
	                  &c->lists[0], &c->lists[1], &c->lists[2], &c->lists[3], &c->ctl, &c->order, &c->hdr7,
	                  &c->pk, &c->tasks[0], &c->tasks[1], &c->recs, &c->chunks, &c->lane_mem, &c->vis2, &c->vis1_full, &c->vis_gen,
	                  &c->task_first[0], &c->task_first[1], &c->task_cnt[0], &c->task_cnt[1]};
	for (DevBuf *b : bufs) if (b->p) cudaFree(b->p);
	for (int k = 0; k < 2; k++) { if (c->in[k].h_pin) cudaFreeHost(c->in[k].h_pin); if (c->in[k].ev_up) cudaEventDestroy(c->in[k].ev_up); }
	if (c->h_stage) { cudaFreeHost(c->h_stage); for (int k = 0; k < 2 * DSB_STAGE_THREADS; k++) if (c->ev_stage[k]) cudaEventDestroy(c->ev_stage[k]); }
	for (int i = 0; i < DSB_N_EV; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	if (c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

extern "C" int dsb_ctx_set_bin_capacity(dsb_ctx *c, uint32_t m_bin_read) { if (!c) return DSB_E_ARG; c->m_bin_read = m_bin_read; return DSB_OK; }
extern "C" uint32_t dsb_ctx_bin_capacity(dsb_ctx *c) { return c ? c->m_bin_read : 0; }

extern "C" void *dsb_ctx_stream(dsb_ctx *c) { return c ? (void *)c->stream : nullptr; }

extern "C" int dsb_host_alloc(size_t bytes, void **out)
{
	if (!out) return DSB_E_ARG;
	DSB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
	return DSB_OK;
}
extern "C" void dsb_host_free(void *p) { if (p) cudaFreeHost(p); }
// page-lock memory the caller allocated itself (huge pages: far fewer pages to pin than cudaHostAlloc's 4 KB pages)
extern "C" int dsb_host_register(void *p, size_t bytes)
{
	if (!p || !bytes) return DSB_E_ARG;
	DSB_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
	return DSB_OK;
}
extern "C" void dsb_host_unregister(void *p) { if (p) cudaHostUnregister(p); }
extern "C" int dsb_device_memory(int device, uint64_t *free_bytes, uint64_t *total_bytes)
{
	DSB_CUDA(cudaSetDevice(device));
	size_t f = 0, t = 0;
	DSB_CUDA(cudaMemGetInfo(&f, &t));
	if (free_bytes) *free_bytes = f;
	if (total_bytes) *total_bytes = t;
	return DSB_OK;
}

// device buffers of the upload step for a batch of n_reads reads / n_bases bases (grow-only)
static int reserve_upload(dsb_ctx *c, BatchIn &I, uint64_t n_reads, uint64_t n_bases, uint64_t n_tiles, uint64_t bo, uint64_t wo, uint64_t so)
{
	const size_t tbl_bytes = (size_t)(n_reads + 1) * 8;
	int rc;
	if ((rc = ensure(I.seqs, n_bases + 16)) || (rc = ensure(c->read_off, tbl_bytes)) || (rc = ensure(c->bin_off, tbl_bytes)) ||
	    (rc = ensure(c->bits_off, tbl_bytes)) || (rc = ensure(c->seed_off, (size_t)(n_reads + 1) * 4)) || (rc = ensure(c->tiles, n_tiles * 8 + 8)) ||
	    (rc = ensure(c->bin, bo + 64)) || (rc = ensure(c->bits, wo * 4 + 64)) ||
	    (rc = ensure(c->seeds[0], so * sizeof(dsb_seed) + 64)) || (rc = ensure(c->seeds[1], so * sizeof(dsb_seed) + 64)) ||
	    (rc = ensure(c->n_seeds[0], (size_t)n_reads * 4)) || (rc = ensure(c->n_seeds[1], (size_t)n_reads * 4)) ||
	    (rc = ensure(c->total_score[0], (size_t)n_reads * 4)) || (rc = ensure(c->total_score[1], (size_t)n_reads * 4)) ||
	    (rc = ensure(c->rr, (size_t)n_reads * sizeof(dsb_read_result))) || (rc = ensure(c->order, (size_t)n_reads * 4)) || (rc = ensure(c->hdr7, (size_t)n_reads + 16)))
		return rc;
	return DSB_OK;
}
static int ensure_zeroed(DevBuf &b, size_t bytes, cudaStream_t st);
static ScratchLayout scratch_layout(uint32_t max_anchors, uint32_t max_matches);
// device buffers of the run step: scratch of the warp-per-read kernels, pools between the phase kernels, seeding memory
static int reserve_run(dsb_ctx *c, uint64_t n, uint64_t n_bases, uint64_t bits_words)
{
	cudaStream_t st = c->stream;
	const ScratchLayout L = scratch_layout(c->opts.max_anchors, c->opts.max_matches);
	c->scratch_stride = L.total;
	int rc;
	if ((rc = ensure(c->scratch, (size_t)L.total * (c->n_warps + c->heavy_blocks))) != DSB_OK) return rc;
	// Pools between the phase kernels, sized from the batch; a kernel that runs out of one sets its bit in ctl[CTL_OVERFLOW] and
	// dsb_batch_download doubles that pool and runs the batch again (grow[] keeps the factor for the batches that follow).
	const uint64_t sc = c->opts.pool_scale_pct ? c->opts.pool_scale_pct : 100;
	auto pool = [&](uint64_t base, int k) { return std::min<uint64_t>(((base * sc / 100) << c->grow[k]) + 1024, 0xfffffff0ull); };
	const uint64_t task_cap = pool(n_bases / 32 + 4ull * n, 0), n_chunks = pool(n_bases / 16 + 16ull * n, 1);
	const uint64_t anc_cap = pool(n_bases / 8 + 64ull * n + (1u << 16), 2), chain_cap = pool(n_bases / 32 + 16ull * n + (1u << 14), 3);
	// hits: the pre-filter chains of a read use 2 slots each (second half = merge-sort scratch)
	const uint64_t hits_cap = pool(std::max<uint64_t>(4096, (uint64_t)n * 24), 4);
	if ((rc = ensure(c->hits, hits_cap * sizeof(dsb_hit))) != DSB_OK) return rc;
	if ((rc = ensure(c->prof, (size_t)n * 32)) != DSB_OK) return rc;
	if ((rc = ensure(c->work, (size_t)n * sizeof(ReadWork))) || (rc = ensure(c->anc_pool, anc_cap * sizeof(DevAnchor))) ||
	    (rc = ensure(c->chain_pool, chain_cap * sizeof(DevChain))) || (rc = ensure(c->ctl, CTL_WORDS * 4)))
		return rc;
	for (int l = 0; l < N_LISTS; l++) if ((rc = ensure(c->lists[l], (size_t)n * 4)) != DSB_OK) return rc;
	// seeding: packed strands, task lists, per-task records, staging chunks, per-lane memory of the persistent k_seed grid
	const uint64_t seed_lanes = (uint64_t)c->seed_blocks * SEED_WARPS_PER_BLOCK * 32;
	const bool big_rows = c->ix->dev.n_lines * 128 >= (1ull << 32);
	if ((rc = ensure(c->pk, (bits_words / N_BITVEC + 4) * 8)) || (rc = ensure(c->tasks[0], task_cap * sizeof(SeedTaskRef))) ||
	    (rc = ensure(c->tasks[1], task_cap * sizeof(SeedTaskRef))) || (rc = ensure(c->recs, task_cap * sizeof(SeedRec))) ||
	    (rc = ensure(c->chunks, n_chunks * 64)) || (rc = ensure(c->lane_mem, seed_lanes * SEED_MEM_SLOTS * sizeof(MemRst))) ||
	    (rc = ensure_zeroed(c->vis2, seed_lanes * VIS2_SLOTS * 8, st)) || (rc = ensure_zeroed(c->vis_gen, seed_lanes * 4, st)) ||
	    (rc = ensure(c->vis1_full, big_rows ? seed_lanes * VIS1_SLOTS * 8 : 64)))
		return rc;
	for (int s = 0; s < 2; s++) if ((rc = ensure(c->task_first[s], (size_t)n * 4)) || (rc = ensure(c->task_cnt[s], (size_t)n * 4))) return rc;
	c->task_cap = (uint32_t)(std::min<uint64_t>(c->tasks[0].cap, c->tasks[1].cap) / sizeof(SeedTaskRef));
	c->task_cap = (uint32_t)std::min<uint64_t>(c->task_cap, c->recs.cap / sizeof(SeedRec));
	c->n_chunks = (uint32_t)std::min<uint64_t>(c->chunks.cap / 64, 0xfffffff0u);
	c->hits_cap = c->hits.cap / sizeof(dsb_hit);
	return DSB_OK;
}

// pinned host memory of a batch's per-read tables (grow-only)
static int pin_tables(BatchIn &I, size_t need)
{
	if (need <= I.h_pin_cap) return DSB_OK;
	if (I.h_pin) cudaFreeHost(I.h_pin);
	I.h_pin = nullptr; I.h_pin_cap = 0;
	DSB_CUDA(cudaHostAlloc(&I.h_pin, need, cudaHostAllocDefault));
	I.h_pin_cap = need;
	return DSB_OK;
}

// the pinned ring of the staged upload: DSB_STAGE_THREADS x 2 pieces, an event per piece ("the copy out of it has landed")
static int stage_ring(dsb_ctx *c)
{
	if (c->h_stage) return DSB_OK;
	DSB_CUDA(cudaHostAlloc(&c->h_stage, 2 * DSB_STAGE_THREADS * (size_t)DSB_STAGE_BYTES, cudaHostAllocDefault));
	for (int k = 0; k < 2 * DSB_STAGE_THREADS; k++) DSB_CUDA(cudaEventCreateWithFlags(&c->ev_stage[k], cudaEventDisableTiming | (dsb_blocking_sync() ? cudaEventBlockingSync : 0)));
	return DSB_OK;
}

// All device buffers of a context for batches of up to max_reads reads and max_bases bases, allocated now (the driver calls this
// before its timed interval: no cudaMalloc while batches are in flight); also sizes the pinned staging of the offset tables.
extern "C" int dsb_ctx_reserve(dsb_ctx *c, uint32_t max_reads, uint64_t max_bases)
{
	if (!c || max_reads == 0) { dsb_set_error("dsb_ctx_reserve: bad argument"); return DSB_E_ARG; }
	DSB_CUDA(cudaSetDevice(c->ix->device));
	const uint64_t n = max_reads, nb = max_bases;
	const uint64_t n_tiles = nb / PROBE_TILE + n, bo = 2 * nb + (2 * DSB_GUARD + 16) * n, wo = (uint64_t)N_BITVEC * (nb / 32 + 2 * n), so = nb / 2 + 2 * n;
	if (so >= 0xffffffffull) { dsb_set_error("dsb_ctx_reserve: batch limit too large (seed slots overflow 32 bits)"); return DSB_E_ARG; }
	int rc;
	for (int k = 0; k < 2; k++) if ((rc = reserve_upload(c, c->in[k], n, nb, n_tiles, bo, wo, so)) != DSB_OK) return rc;   // both input sets: the next batch is uploaded while one runs
	if ((rc = reserve_run(c, n, nb, wo)) != DSB_OK) return rc;
	const size_t pin_need = (size_t)(n + 1) * 8 * 4 + n_tiles * 8 + n + 64;
	for (int k = 0; k < 2; k++) if ((rc = pin_tables(c->in[k], pin_need)) != DSB_OK) return rc;
	if ((rc = stage_ring(c)) != DSB_OK) return rc;
	DSB_CUDA(cudaStreamSynchronize(c->stream));
	return DSB_OK;
}

static double host_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// counting gate per device for the staged (pageable) upload path
struct StageGate {
	static constexpr int MAX_DEV = 64;
	static std::mutex mu; static std::condition_variable cv; static int busy[MAX_DEV]; static int limit;
	int dev;
	explicit StageGate(int device) : dev(device < 0 || device >= MAX_DEV ? 0 : device)
	{
		std::unique_lock<std::mutex> lk(mu);
		if (limit < 0) { const char *e = getenv("DSB_STAGE_CONCURRENCY"); limit = e ? atoi(e) : 3; if (limit < 1) limit = 1 << 20; }
		cv.wait(lk, [&] { return busy[dev] < limit; });
		busy[dev]++;
	}
	~StageGate() { { std::lock_guard<std::mutex> lk(mu); busy[dev]--; } cv.notify_all(); }
};
std::mutex StageGate::mu; std::condition_variable StageGate::cv; int StageGate::busy[StageGate::MAX_DEV] = {0}; int StageGate::limit = -1;

static int upload_impl(dsb_ctx *c, const char *seqs, const uint64_t *offs, uint32_t n_reads);
extern "C" int dsb_batch_upload(dsb_ctx *c, const char *seqs, const uint64_t *offs, uint32_t n_reads)
{
	const double t0 = host_now();
	const int rc = upload_impl(c, seqs, offs, n_reads);
	if (c) c->host_s[0] += host_now() - t0;
	return rc;
}
static int upload_impl(dsb_ctx *c, const char *seqs, const uint64_t *offs, uint32_t n_reads)
{
	if (!c || (n_reads && (!seqs || !offs))) { dsb_set_error("dsb_batch_upload: null argument"); return DSB_E_ARG; }
	DSB_CUDA(cudaSetDevice(c->ix->device));
	if (c->has_pending) { dsb_set_error("dsb_batch_upload: the batch uploaded before has not been run yet"); return DSB_E_ARG; }
	// the input set the running batch does not use (or the same one again when nothing is in flight)
	const int u = c->run_pending ? (c->cur ^ 1) : c->cur;
	BatchIn &I = c->in[u];
	if (!c->run_pending) c->ran = false;
	I.n_reads = n_reads;
	if (n_reads == 0) { I.n_tiles = 0; I.n_bases = 0; c->pend = u; c->has_pending = true; return DSB_OK; }
	const uint64_t n_bases = offs[n_reads] - offs[0];
	if (offs[0] != 0) { dsb_set_error("dsb_batch_upload: offs[0] must be 0"); return DSB_E_ARG; }
	// per-read layout tables (host): bin / bit-vector / seed-slot offsets and the probe tile list
	const size_t tbl_bytes = (size_t)(n_reads + 1) * 8;
	DSB_CUDA(cudaEventRecord(c->ev[DSB_N_KERNELS + 3], c->copy_stream));        // the copy stream turns to this batch (dsb_batch_timeline)
	uint64_t n_tiles = 0; uint32_t max_len = 0;
	for (uint32_t r = 0; r < n_reads; r++) {
		if (offs[r + 1] < offs[r] || offs[r + 1] - offs[r] > c->opts.max_read_len) { dsb_set_error("read %u: bad offsets or longer than max_read_len %u", r, c->opts.max_read_len); return DSB_E_ARG; }
		const uint32_t len = (uint32_t)(offs[r + 1] - offs[r]);
		if (len >= 40) n_tiles += (len + PROBE_TILE - 1) / PROBE_TILE;
		max_len = std::max(max_len, len);
	}
	int rc;
	{
		const size_t pin_need = tbl_bytes * 3 + (size_t)(n_reads + 1) * 8 + n_tiles * 8 + n_reads + 64;
		if (pin_need > I.h_pin_cap && (rc = pin_tables(I, pin_need + pin_need / 4)) != DSB_OK) return rc;
	}
	uint64_t *h_bin_off = (uint64_t *)I.h_pin, *h_bits_off = h_bin_off + n_reads + 1, *h_tiles64 = h_bits_off + n_reads + 1;
	uint2 *h_tiles = (uint2 *)h_tiles64;
	uint32_t *h_seed_off = (uint32_t *)(h_tiles64 + n_tiles);
	uint64_t bo = 0, wo = 0, so = 0, ti = 0;
	for (uint32_t r = 0; r < n_reads; r++) {
		const uint32_t len = (uint32_t)(offs[r + 1] - offs[r]);
		h_bin_off[r] = bo; h_bits_off[r] = wo; h_seed_off[r] = (uint32_t)so;
		if (len >= 40) {
			bo += ((uint64_t)2 * len + 2 * DSB_GUARD + 15) & ~15ull;
			wo += (uint64_t)N_BITVEC * bits_words(len);
			so += seed_slots(len);
			for (uint32_t st = 0; st < len; st += PROBE_TILE) { h_tiles[ti].x = r; h_tiles[ti].y = st; ti++; }
		}
	}
	h_bin_off[n_reads] = bo; h_bits_off[n_reads] = wo; h_seed_off[n_reads] = (uint32_t)so;
	// work order of the first seeding pass: longest reads first (the expensive tail starts early)
	uint32_t *h_order = h_seed_off + n_reads + 1;
	// (a stable counting sort by length: a comparison sort of a million 150-bp reads cost more host time than their kernels)
	{
		std::vector<uint32_t> &first = c->h_len_first;
		first.assign((size_t)max_len + 2, 0);
		for (uint32_t r = 0; r < n_reads; r++) first[(uint32_t)(offs[r + 1] - offs[r])]++;
		uint32_t run = 0;
		for (uint32_t l = max_len + 1; l-- > 0;) { const uint32_t n_l = first[l]; first[l] = run; run += n_l; }   // first[l] = reads longer than l
		for (uint32_t r = 0; r < n_reads; r++) h_order[first[(uint32_t)(offs[r + 1] - offs[r])]++] = r;
	}
	// policy P3: the capacity of the reference's bin_read buffer (BUFF_REALLOC, utils.h:117-122) in input order decides the
	// chunk-header byte an alignment that runs 7 bases off the start of a read compares with
	uint8_t *h_hdr7 = (uint8_t *)(h_order + n_reads);
	for (uint32_t r = 0; r < n_reads; r++) {
		const uint32_t len = (uint32_t)(offs[r + 1] - offs[r]);
		if (len >= 40 && 2 * len > c->m_bin_read) c->m_bin_read = 2 * len + 20;
		uint32_t chunk = (c->m_bin_read + 8 + 15) & ~15u;
		if (chunk < 32) chunk = 32;
		h_hdr7[r] = (uint8_t)((chunk >> 8) & 0xff);
	}
	if (so >= 0xffffffffull) { dsb_set_error("batch too large (seed slots overflow 32 bits): split the batch"); return DSB_E_ARG; }
	I.n_tiles = (uint32_t)n_tiles; I.n_bases = n_bases; I.bits_words = wo; I.seed_slots = so; I.bin_bytes = bo; I.max_len = max_len;
	I.h_bits_off.assign(h_bits_off, h_bits_off + n_reads + 1);
	I.h_seed_off.assign(h_seed_off, h_seed_off + n_reads + 1);
	I.h_off.assign(offs, offs + n_reads + 1);
	// only the reads' own buffer here: the tables and work buffers of the context may still hold the batch in flight and its
	// results -- they are sized when this batch is run (run_impl)
	if ((rc = ensure(I.seqs, n_bases + 16)) != DSB_OK) return rc;
	cudaStream_t st = c->copy_stream;
	{	// the reads: straight from the caller's buffer when it is page-locked, else through the context's pinned staging ring (the
		// runtime's own path for pageable memory serialises the copies of all contexts of a process)
		cudaPointerAttributes pa;
		const bool pinned = cudaPointerGetAttributes(&pa, seqs) == cudaSuccess && (pa.type == cudaMemoryTypeHost || pa.type == cudaMemoryTypeManaged);
		(void)cudaGetLastError();
		if (pinned || n_bases < (1u << 20)) DSB_CUDA(cudaMemcpyAsync(I.seqs.p, seqs, n_bases, cudaMemcpyHostToDevice, st));
		else {
			if ((rc = stage_ring(c)) != DSB_OK) return rc;
			// At most DSB_STAGE_CONCURRENCY (3) contexts of a device stage a batch at a time, each with DSB_STAGE_THREADS host threads
			// (one thread copies ~6.7 GB/s: 75 ms per 0.5-GB batch during which this context has no kernel to run).  Contexts that
			// share a GPU evenly finish together; without the limit they then all copy together while the device idles (measured
			// with one staging thread: 6 contexts in lock step, 74 of 405 5-ms bins of the device's time line empty, 67.5 ms
			// per batch against 57.5 with page-locked inputs).
			StageGate gate(c->ix->device);
			const int dev = c->ix->device;
			const uint64_t n_pieces = (n_bases + DSB_STAGE_BYTES - 1) / DSB_STAGE_BYTES;
			int trc[DSB_STAGE_THREADS];
			auto work = [&](int t) {
				trc[t] = DSB_OK;
				if (cudaSetDevice(dev) != cudaSuccess) { trc[t] = DSB_E_CUDA; return; }
				for (uint64_t i = (uint64_t)t; i < n_pieces; i += DSB_STAGE_THREADS) {
					const int k = 2 * t + (int)((i / DSB_STAGE_THREADS) & 1);
					const uint64_t off = i * DSB_STAGE_BYTES;
					const size_t nb = (size_t)std::min<uint64_t>(DSB_STAGE_BYTES, n_bases - off);
					char *stg = (char *)c->h_stage + (size_t)k * DSB_STAGE_BYTES;
					if (cudaEventSynchronize(c->ev_stage[k]) != cudaSuccess) { trc[t] = DSB_E_CUDA; return; }   // the copy that last used this piece has landed
					memcpy(stg, seqs + off, nb);
					if (cudaMemcpyAsync((char *)I.seqs.p + off, stg, nb, cudaMemcpyHostToDevice, st) != cudaSuccess || cudaEventRecord(c->ev_stage[k], st) != cudaSuccess) { trc[t] = DSB_E_CUDA; return; }
				}
			};
			const int n_thr = (int)std::min<uint64_t>(DSB_STAGE_THREADS, n_pieces);
			std::thread helper[DSB_STAGE_THREADS];
			for (int t = 1; t < n_thr; t++) helper[t] = std::thread(work, t);
			work(0);
			for (int t = 1; t < n_thr; t++) helper[t].join();
			for (int t = 0; t < n_thr; t++) if (trc[t] != DSB_OK) { dsb_set_error("staged upload: %s", cudaGetErrorString(cudaGetLastError())); return trc[t]; }
		}
	}
	DSB_CUDA(cudaEventRecord(I.ev_up, st));
	c->pend = u; c->has_pending = true;
	return DSB_OK;
}

// the per-read tables of in[u] -> device, on the compute stream (they are single-buffered: the batch before has finished there)
static int upload_tables(dsb_ctx *c, BatchIn &I, cudaStream_t st)
{
	const uint32_t n_reads = I.n_reads;
	const size_t tbl_bytes = (size_t)(n_reads + 1) * 8;
	const uint64_t *h_bin_off = (const uint64_t *)I.h_pin, *h_bits_off = h_bin_off + n_reads + 1, *h_tiles64 = h_bits_off + n_reads + 1;
	const uint32_t *h_seed_off = (const uint32_t *)(h_tiles64 + I.n_tiles);
	const uint32_t *h_order = h_seed_off + n_reads + 1;
	const uint8_t *h_hdr7 = (const uint8_t *)(h_order + n_reads);
	DSB_CUDA(cudaMemcpyAsync(c->read_off.p, I.h_off.data(), tbl_bytes, cudaMemcpyHostToDevice, st));
	DSB_CUDA(cudaMemcpyAsync(c->bin_off.p, h_bin_off, tbl_bytes, cudaMemcpyHostToDevice, st));
	DSB_CUDA(cudaMemcpyAsync(c->bits_off.p, h_bits_off, tbl_bytes, cudaMemcpyHostToDevice, st));
	DSB_CUDA(cudaMemcpyAsync(c->seed_off.p, h_seed_off, (size_t)(n_reads + 1) * 4, cudaMemcpyHostToDevice, st));
	if (I.n_tiles) DSB_CUDA(cudaMemcpyAsync(c->tiles.p, h_tiles64, (size_t)I.n_tiles * 8, cudaMemcpyHostToDevice, st));
	DSB_CUDA(cudaMemcpyAsync(c->order.p, h_order, (size_t)n_reads * 4, cudaMemcpyHostToDevice, st));
	DSB_CUDA(cudaMemcpyAsync(c->hdr7.p, h_hdr7, (size_t)n_reads, cudaMemcpyHostToDevice, st));
	return DSB_OK;
}

// zero-initialised grow-only buffer (the visited-row tables and their generation counters start at zero)
static int ensure_zeroed(DevBuf &b, size_t bytes, cudaStream_t st)
{
	if (bytes <= b.cap) return DSB_OK;
	int rc = ensure(b, bytes);
	if (rc != DSB_OK) return rc;
	DSB_CUDA(cudaMemsetAsync(b.p, 0, b.cap, st));
	return DSB_OK;
}

static int run_impl(dsb_ctx *c, int32_t max_read_l_in, bool rerun);
extern "C" int dsb_batch_run(dsb_ctx *c, int32_t max_read_l_in)
{
	const double t0 = host_now();
	const int rc = run_impl(c, max_read_l_in, false);
	if (c) c->host_s[1] += host_now() - t0;
	return rc;
}
static int run_impl(dsb_ctx *c, int32_t max_read_l_in, bool rerun)
{
	if (!c) return DSB_E_ARG;
	DSB_CUDA(cudaSetDevice(c->ix->device));
	cudaStream_t st = c->stream;
	c->launches = 0;
	c->max_read_l_in = max_read_l_in;
	// the batch uploaded last; without one, the current batch again (re-run after a pool overflow, repeated runs of a resident batch)
	if (c->has_pending && !rerun) { c->cur = c->pend; c->has_pending = false; }   // (results of the batch before that nobody fetched are dropped)
	BatchIn &I = c->in[c->cur];
	c->n_reads = I.n_reads; c->n_tiles = I.n_tiles; c->n_bases = I.n_bases; c->bits_words = I.bits_words; c->seed_slots = I.seed_slots; c->bin_bytes = I.bin_bytes; c->max_len = I.max_len;
	const uint32_t n = c->n_reads;
	if (n == 0) { c->ran = true; c->run_pending = true; return DSB_OK; }
	const ScratchLayout L = scratch_layout(c->opts.max_anchors, c->opts.max_matches);
	int rc;
	if ((rc = reserve_upload(c, I, n, I.n_bases, I.n_tiles, I.bin_bytes, I.bits_words, I.seed_slots)) != DSB_OK) return rc;
	if ((rc = reserve_run(c, n, c->n_bases, c->bits_words)) != DSB_OK) return rc;
	DSB_CUDA(cudaStreamWaitEvent(st, I.ev_up, 0));                 // the reads are resident
	if ((rc = upload_tables(c, I, st)) != DSB_OK) return rc;
	const bool big_rows = c->ix->dev.n_lines * 128 >= (1ull << 32);
	DSB_CUDA(cudaMemsetAsync(c->ctl.p, 0, CTL_WORDS * 4, st));
	DSB_CUDA(cudaMemsetAsync(c->prof.p, 0, (size_t)n * 32, st));
	unsigned long long *cnt = (unsigned long long *)c->counters.p;
	DSB_CUDA(cudaMemsetAsync(cnt, 0, DSB_CNT_COUNT * 8, st));
	DSB_CUDA(cudaMemsetAsync(cnt + DSB_CNT_FIRST_LONG, 0xff, 8, st));
	DSB_CUDA(cudaEventRecord(c->ev[0], st));
	if (c->n_tiles) {
		ProbeParams P;
		P.ix = c->ix->dev; P.seqs = (const char *)I.seqs.p; P.read_off = (const uint64_t *)c->read_off.p; P.bin_off = (const uint64_t *)c->bin_off.p;
		P.bits_off = (const uint64_t *)c->bits_off.p; P.tiles = (const uint2 *)c->tiles.p; P.hdr7 = (const uint8_t *)c->hdr7.p; P.bin = (uint8_t *)c->bin.p; P.bits = (uint32_t *)c->bits.p;
		P.pk = (uint64_t *)c->pk.p;
		k_encode_probe<<<c->n_tiles, PROBE_THREADS, 0, st>>>(P);
		c->launches++;
	}
	DSB_CUDA(cudaEventRecord(c->ev[1], st));
	{
		IslandParams P;
		P.l_ek = c->ix->dev.l_ek; P.n_reads = n; P.read_off = (const uint64_t *)c->read_off.p; P.bits_off = (const uint64_t *)c->bits_off.p;
		P.seed_off = (const uint32_t *)c->seed_off.p; P.bits = (const uint32_t *)c->bits.p;
		for (int s = 0; s < 2; s++) { P.seeds[s] = (dsb_seed *)c->seeds[s].p; P.n_seeds[s] = (uint32_t *)c->n_seeds[s].p; P.total_score[s] = (uint32_t *)c->total_score[s].p; }
		P.counters = cnt;
		P.tasks = (SeedTaskRef *)c->tasks[PASS_FAST & 1].p; P.task_cap = c->task_cap; P.ctl = (uint32_t *)c->ctl.p;
		for (int s = 0; s < 2; s++) { P.task_first[s] = (uint32_t *)c->task_first[s].p; P.task_cnt[s] = (uint32_t *)c->task_cnt[s].p; }
		k_islands<<<(2 * n + 127) / 128, 128, 0, st>>>(P);
		c->launches++;
	}
	DSB_CUDA(cudaEventRecord(c->ev[2], st));
	{
		ClassifyLaunch A;
		ClassifyParams &P = A.P;
		P.ix = c->ix->dev; P.n_reads = n; P.read_off = (const uint64_t *)c->read_off.p; P.bin_off = (const uint64_t *)c->bin_off.p; P.bin = (const uint8_t *)c->bin.p;
		P.seed_off = (const uint32_t *)c->seed_off.p;
		for (int s = 0; s < 2; s++) { P.seeds[s] = (const dsb_seed *)c->seeds[s].p; P.n_seeds[s] = (const uint32_t *)c->n_seeds[s].p; P.total_score[s] = (const uint32_t *)c->total_score[s].p; }
		P.order = (const uint32_t *)c->order.p; P.prof = (uint32_t *)c->prof.p;
		for (int s = 0; s < 2; s++) { P.tasks[s] = (SeedTaskRef *)c->tasks[s].p; P.recs[s] = (SeedRec *)c->recs.p; P.task_first[s] = (uint32_t *)c->task_first[s].p; P.task_cnt[s] = (uint32_t *)c->task_cnt[s].p; }
		P.task_cap = c->task_cap; P.chunks = (const uint4 *)c->chunks.p;
		P.work = (ReadWork *)c->work.p;
		P.anc_pool = (DevAnchor *)c->anc_pool.p; P.anc_pool_cap = (uint32_t)std::min<uint64_t>(c->anc_pool.cap / sizeof(DevAnchor), 0xfffffff0u);
		P.chain_pool = (DevChain *)c->chain_pool.p; P.chain_pool_cap = (uint32_t)std::min<uint64_t>(c->chain_pool.cap / sizeof(DevChain), 0xfffffff0u);
		for (int l = 0; l < N_LISTS; l++) P.list[l] = (uint32_t *)c->lists[l].p;
		P.ctl = (uint32_t *)c->ctl.p;
		P.scratch = (uint8_t *)c->scratch.p; P.scratch_stride = L.total;
		P.max_anchors = c->opts.max_anchors; P.max_matches = c->opts.max_matches;
		P.rr = (dsb_read_result *)c->rr.p; P.hits = (dsb_hit *)c->hits.p; P.hits_cap = c->hits_cap; P.hits_cursor = cnt + DSB_CNT_HITS_CURSOR;
		P.counters = cnt;
		A.L = L;
		SeedPassParams SP;
		SP.E.ix = c->ix->dev; SP.E.pk = (const uint64_t *)c->pk.p; SP.E.read_off = P.read_off; SP.E.bits_off = (const uint64_t *)c->bits_off.p; SP.E.seed_off = P.seed_off;
		SP.E.seeds[0] = P.seeds[0]; SP.E.seeds[1] = P.seeds[1]; SP.E.recs = (SeedRec *)c->recs.p; SP.E.chunks = (uint4 *)c->chunks.p; SP.E.n_chunks = c->n_chunks;
		SP.E.big_rows = big_rows ? 1 : 0;
		SP.ctl = P.ctl; SP.task_cap = c->task_cap;
		SP.policy = SC_POLICY; SP.fetch_min = SC_FETCH_MIN;
		if (const char *e = getenv("DSB_SEED_POLICY")) { int a = 0, b = 0; if (sscanf(e, "%d,%d", &a, &b) == 2 && a >= 0 && b >= 1) { SP.policy = a; SP.fetch_min = b; } }   // developer knob (profiles/)
		SP.lane_mem = (MemRst *)c->lane_mem.p; SP.vis2 = (uint64_t *)c->vis2.p; SP.vis1_full = (uint64_t *)c->vis1_full.p; SP.vis_gen = (uint32_t *)c->vis_gen.p;
		auto seed = [&](int pass) { SP.pass = pass; SP.E.slow = pass != PASS_FAST; SP.E.tasks = (const SeedTaskRef *)c->tasks[pass & 1].p; k_seed<<<c->seed_blocks, SEED_WARPS_PER_BLOCK * 32, 0, st>>>(SP); };
		const int blocks = c->n_warps / CLASSIFY_WARPS_PER_BLOCK, threads = CLASSIFY_WARPS_PER_BLOCK * 32;
		const size_t smem = CLASSIFY_WARPS_PER_BLOCK * sizeof(WarpSmem);
		// classify_seq's control flow (cly.c:3098-3131) as a sequence of phase kernels over work lists
		seed(PASS_FAST);                                                        DSB_CUDA(cudaEventRecord(c->ev[3], st));
		k_chain<<<blocks, threads, smem, st>>>(A, PASS_FAST, -1, 1);            DSB_CUDA(cudaEventRecord(c->ev[4], st));
		seed(PASS_SLOW0);                                                       DSB_CUDA(cudaEventRecord(c->ev[5], st));
		k_chain<<<blocks, threads, smem, st>>>(A, PASS_SLOW0, LIST_SLOW0, 3);   DSB_CUDA(cudaEventRecord(c->ev[6], st));
		seed(PASS_SLOW1);                                                       DSB_CUDA(cudaEventRecord(c->ev[7], st));
		k_chain<<<blocks, threads, smem, st>>>(A, PASS_SLOW1, LIST_SLOW1, 5);   DSB_CUDA(cudaEventRecord(c->ev[8], st));
		// scoring: a warp per read; the few reads that give up there (ERR_DEFER) then get a CTA each
		k_score<<<blocks, threads, smem + CLASSIFY_WARPS_PER_BLOCK * sizeof(MatchSmem), st>>>(A, LIST_SCORE, 6);
		DSB_CUDA(cudaEventRecord(c->ev[9], st));
		k_score_heavy<<<c->heavy_blocks, TEAM_WARPS * 32, 0, st>>>(A, LIST_SCORE_HEAVY, 7, (uint32_t)c->n_warps);
		DSB_CUDA(cudaEventRecord(c->ev[10], st));
		c->launches += 8;
	}
	{
		FinalizeParams P;
		P.n_reads = n; P.max_read_l_in = max_read_l_in;
		P.filter_min_length = c->opts.l_min_match; P.filter_min_score = c->opts.min_score; P.filter_min_score_LV3 = c->opts.min_score + 10;   // cly_mt.c:521-523
		P.rr = (dsb_read_result *)c->rr.p; P.hits = (dsb_hit *)c->hits.p; P.counters = cnt; P.ctl = (const uint32_t *)c->ctl.p;
		k_finalize<<<(n + 127) / 128, 128, 0, st>>>(P);
		c->launches++;
	}
	DSB_CUDA(cudaEventRecord(c->ev[11], st));
	DSB_CUDA(cudaGetLastError());
	c->ran = true; c->run_pending = true;
	return DSB_OK;
}

extern "C" int dsb_batch_sync(dsb_ctx *c)
{
	if (!c) return DSB_E_ARG;
	DSB_CUDA(cudaSetDevice(c->ix->device));
	DSB_CUDA(cudaStreamSynchronize(c->copy_stream));
	DSB_CUDA(cudaStreamSynchronize(c->stream));
	return DSB_OK;
}

#define DSB_MAX_RETRIES 8
static int download_impl(dsb_ctx *c, int32_t *max_read_l_out, dsb_read_result *rr, dsb_hit *hits, uint64_t hits_cap, uint64_t *n_hits_out);
extern "C" int dsb_batch_download(dsb_ctx *c, int32_t *max_read_l_out, dsb_read_result *rr, dsb_hit *hits, uint64_t hits_cap, uint64_t *n_hits_out)
{
	const double t0 = host_now();
	const int rc = download_impl(c, max_read_l_out, rr, hits, hits_cap, n_hits_out);
	if (c) { c->run_pending = false; c->host_s[2] += host_now() - t0; c->host_s[3] += 1; c->host_s[4] += c->retries; }   // (the results stay on the device until the next run)
	return rc;
}
static int download_impl(dsb_ctx *c, int32_t *max_read_l_out, dsb_read_result *rr, dsb_hit *hits, uint64_t hits_cap, uint64_t *n_hits_out)
{
	if (!c || !c->ran) { dsb_set_error("dsb_batch_download: no batch has been run"); return DSB_E_ARG; }
	DSB_CUDA(cudaSetDevice(c->ix->device));
	if (n_hits_out) *n_hits_out = 0;
	if (max_read_l_out) *max_read_l_out = c->max_read_l_in;
	if (c->n_reads == 0) return DSB_OK;
	cudaStream_t st = c->stream;
	unsigned long long h_cnt[DSB_CNT_COUNT];
	c->retries = 0;
	for (;;) {
		uint32_t ovf = 0;
		DSB_CUDA(cudaMemcpyAsync(h_cnt, c->counters.p, sizeof h_cnt, cudaMemcpyDeviceToHost, st));
		DSB_CUDA(cudaMemcpyAsync(&ovf, (uint32_t *)c->ctl.p + CTL_OVERFLOW, 4, cudaMemcpyDeviceToHost, st));
		DSB_CUDA(cudaStreamSynchronize(st));
		if (ovf & OVF_STUCK) { dsb_set_error("seeding did not terminate (malformed index?)"); return DSB_E_CUDA; }
		if (!ovf || c->retries >= DSB_MAX_RETRIES) break;
		// a pool between the phase kernels overflowed (a batch far richer in seeds / anchors / chains than the sizing assumes):
		// double the pools concerned and run the batch again -- its inputs are still resident
		for (int k = 0; k < 5; k++) if (ovf & (1u << k)) c->grow[k]++;
		c->retries++;
		int rc = run_impl(c, c->max_read_l_in, true);          // (the batch uploaded ahead, if any, keeps waiting)
		if (rc != DSB_OK) return rc;
	}
	{
		uint32_t ovf = 0;
		DSB_CUDA(cudaMemcpy(&ovf, (uint32_t *)c->ctl.p + CTL_OVERFLOW, 4, cudaMemcpyDeviceToHost));
		if (ovf) { dsb_set_error("device pools still too small after %d doublings (overflow mask 0x%x): split the batch", c->retries, ovf); return DSB_E_NOMEM; }
	}
	if (rr) DSB_CUDA(cudaMemcpyAsync(rr, c->rr.p, (size_t)c->n_reads * sizeof(dsb_read_result), cudaMemcpyDeviceToHost, st));
	DSB_CUDA(cudaStreamSynchronize(st));
	const uint64_t used = std::min<uint64_t>(h_cnt[DSB_CNT_HITS_CURSOR], c->hits_cap);
	if (n_hits_out) *n_hits_out = used;
	if (max_read_l_out) *max_read_l_out = std::max<int32_t>((int32_t)h_cnt[DSB_CNT_MAX_READ_L], c->max_read_l_in);
	if (hits && used) {
		if (used > hits_cap) { dsb_set_error("hits array too small: %llu needed, %llu given", (unsigned long long)used, (unsigned long long)hits_cap); return DSB_E_CAPACITY; }
		DSB_CUDA(cudaMemcpyAsync(hits, c->hits.p, used * sizeof(dsb_hit), cudaMemcpyDeviceToHost, st));
		DSB_CUDA(cudaStreamSynchronize(st));
	}
	if (h_cnt[DSB_CNT_N_ERRORS]) {
		dsb_set_error("%llu read(s) exceeded a per-read capacity (dsb_read_result.error): raise dsb_opts.max_anchors / max_matches", h_cnt[DSB_CNT_N_ERRORS]);
		return DSB_E_CAPACITY;
	}
	return DSB_OK;
}

extern "C" int dsb_batch_retries(dsb_ctx *c) { return c ? c->retries : 0; }

extern "C" int dsb_classify_batch(dsb_ctx *c, const char *seqs, const uint64_t *offs, uint32_t n_reads, int32_t max_read_l_in, int32_t *max_read_l_out,
                                  dsb_read_result *rr, dsb_hit *hits, uint64_t hits_cap, uint64_t *n_hits_out)
{
	int rc;
	if (max_read_l_out) *max_read_l_out = max_read_l_in;
	if ((rc = dsb_batch_upload(c, seqs, offs, n_reads)) != DSB_OK) return rc;
	if ((rc = dsb_batch_run(c, max_read_l_in)) != DSB_OK) return rc;
	return dsb_batch_download(c, max_read_l_out, rr, hits, hits_cap, n_hits_out);
}

// where the host time of the batch calls went on this context so far: seconds in upload (layout tables + copies of the reads
// into the stream), run (kernel launches), download (waiting for the batch + result copies); calls; pool-overflow re-runs
extern "C" int dsb_ctx_host_seconds(dsb_ctx *c, double *out, int cap)
{
	if (!c || !out) return DSB_E_ARG;
	for (int i = 0; i < 5 && i < cap; i++) out[i] = c->host_s[i];
	return DSB_OK;
}

extern "C" int dsb_batch_get_seeds(dsb_ctx *c, uint32_t read, int strand, dsb_seed *out, uint32_t cap, uint32_t *n_out, uint32_t *total_score)
{
	if (!c || !c->ran || read >= c->n_reads || strand < 0 || strand > 1) return DSB_E_ARG;
	DSB_CUDA(cudaSetDevice(c->ix->device));
	DSB_CUDA(cudaStreamSynchronize(c->stream));
	uint32_t n = 0, ts = 0;
	DSB_CUDA(cudaMemcpy(&n, (uint32_t *)c->n_seeds[strand].p + read, 4, cudaMemcpyDeviceToHost));
	DSB_CUDA(cudaMemcpy(&ts, (uint32_t *)c->total_score[strand].p + read, 4, cudaMemcpyDeviceToHost));
	if (n_out) *n_out = n;
	if (total_score) *total_score = ts;
	if (out && n) {
		if (n > cap) return DSB_E_CAPACITY;
		DSB_CUDA(cudaMemcpy(out, (dsb_seed *)c->seeds[strand].p + c->in[c->cur].h_seed_off[read], (size_t)n * sizeof(dsb_seed), cudaMemcpyDeviceToHost));
	}
	return DSB_OK;
}

extern "C" int dsb_batch_kernel_ms(dsb_ctx *c, float *ms, int cap)
{
	if (!c || !c->ran || !ms) return DSB_E_ARG;
	DSB_CUDA(cudaSetDevice(c->ix->device));
	DSB_CUDA(cudaStreamSynchronize(c->stream));
	for (int i = 0; i < DSB_N_KERNELS && i < cap; i++) { ms[i] = 0; if (c->n_reads) DSB_CUDA(cudaEventElapsedTime(&ms[i], c->ev[i], c->ev[i + 1])); }
	return DSB_OK;
}

// stream marks for timing several batches in flight: mark 0/1 = events on this context's stream; elapsed = b.mark_b - a.mark_a
extern "C" int dsb_ctx_mark(dsb_ctx *c, int which)
{
	if (!c || which < 0 || which > 1) return DSB_E_ARG;
	DSB_CUDA(cudaSetDevice(c->ix->device));
	DSB_CUDA(cudaEventRecord(c->ev[DSB_N_KERNELS + 1 + which], c->stream));
	return DSB_OK;
}
extern "C" int dsb_ctx_elapsed_ms(dsb_ctx *a, int mark_a, dsb_ctx *b, int mark_b, float *ms)
{
	if (!a || !b || !ms || (mark_a | mark_b) < 0 || mark_a > 1 || mark_b > 1) return DSB_E_ARG;
	DSB_CUDA(cudaSetDevice(a->ix->device));
	DSB_CUDA(cudaEventSynchronize(b->ev[DSB_N_KERNELS + 1 + mark_b]));
	DSB_CUDA(cudaEventElapsedTime(ms, a->ev[DSB_N_KERNELS + 1 + mark_a], b->ev[DSB_N_KERNELS + 1 + mark_b]));
	return DSB_OK;
}

// Developer aid: where the batch last run on `c` sat on the device's time line -- ms from mark `ref_mark` of context `ref` to
// t[0] the stream turning to the batch (start of the upload) and t[1 .. DSB_N_KERNELS + 1] the kernel boundaries (t[1] = reads
// resident, first kernel starts; t[i + 1] = kernel group i done).
extern "C" int dsb_batch_timeline(dsb_ctx *ref, int ref_mark, dsb_ctx *c, float *t, int cap)
{
	if (!ref || !c || !c->ran || !t || ref_mark < 0 || ref_mark > 1 || cap < DSB_N_KERNELS + 2) return DSB_E_ARG;
	DSB_CUDA(cudaSetDevice(c->ix->device));
	DSB_CUDA(cudaStreamSynchronize(c->stream));
	const cudaEvent_t r = ref->ev[DSB_N_KERNELS + 1 + ref_mark];
	DSB_CUDA(cudaEventElapsedTime(&t[0], r, c->ev[DSB_N_KERNELS + 3]));
	for (int i = 0; i <= DSB_N_KERNELS; i++) DSB_CUDA(cudaEventElapsedTime(&t[1 + i], r, c->ev[i]));
	return DSB_OK;
}

extern "C" int dsb_batch_profile(dsb_ctx *c, uint32_t *out)
{
	if (!c || !c->ran || !out) return DSB_E_ARG;
	DSB_CUDA(cudaSetDevice(c->ix->device));
	DSB_CUDA(cudaStreamSynchronize(c->stream));
	if (c->n_reads) DSB_CUDA(cudaMemcpy(out, c->prof.p, (size_t)c->n_reads * 32, cudaMemcpyDeviceToHost));
	return DSB_OK;
}

extern "C" int dsb_batch_launches(dsb_ctx *c) { return c ? c->launches : 0; }

// work-list sizes of the last run: [0] slow pass 0, [1] slow pass 1, [2] scored (warp), [3] scored again (CTA per read), [4..6] seed
// tasks of the fast / slow 0 / slow 1 pass, [7..9] staging chunks of the passes, [10] anchors, [11] chains kept
extern "C" int dsb_batch_work(dsb_ctx *c, uint32_t out[12])
{
	if (!c || !c->ran || !out) return DSB_E_ARG;
	DSB_CUDA(cudaSetDevice(c->ix->device));
	DSB_CUDA(cudaStreamSynchronize(c->stream));
	uint32_t h[CTL_WORDS];
	DSB_CUDA(cudaMemcpy(h, c->ctl.p, sizeof h, cudaMemcpyDeviceToHost));
	for (int i = 0; i < 4; i++) out[i] = h[CTL_LIST_N + i];
	for (int i = 0; i < 3; i++) { out[4 + i] = h[CTL_TASK_N + i]; out[7 + i] = h[CTL_CHUNK_CURSOR + i]; }
	out[10] = h[CTL_ANC_CURSOR]; out[11] = h[CTL_CHAIN_CURSOR];
	return DSB_OK;
}

extern "C" int dsb_batch_counters(dsb_ctx *c, uint64_t out[16])
{
	if (!c || !c->ran || !out) return DSB_E_ARG;
	DSB_CUDA(cudaSetDevice(c->ix->device));
	DSB_CUDA(cudaStreamSynchronize(c->stream));
	DSB_CUDA(cudaMemcpy(out, c->counters.p, DSB_CNT_COUNT * 8, cudaMemcpyDeviceToHost));
	return DSB_OK;
}

// ------------------------------------------------------------------------------------------------ roofline denominator
// Random-gather microbenchmark: every thread reads `bytes_each` bytes at pseudo-random aligned positions
// of a table far larger than L2; the figure reported is sector-granular traffic (max(bytes_each,32) per gather) per second.
template <int BYTES>
__global__ void __launch_bounds__(256) k_gather(const uint8_t *table, uint64_t n_slots, uint64_t n_gathers, uint64_t seed, unsigned long long *sink)
{
	const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, nthr = gridDim.x * (uint64_t)blockDim.x;
	uint64_t acc = 0;
	constexpr int U = (BYTES <= 16) ? 8 : 4;             // independent gathers in flight per thread
	for (uint64_t g = tid; g < n_gathers; g += nthr * U) {
		const uint8_t *p[U];
		#pragma unroll
		for (int u = 0; u < U; u++) p[u] = table + (dsb_hash64_1((g + u * nthr) ^ seed) % n_slots) * BYTES;
		#pragma unroll
		for (int u = 0; u < U; u++) {
			if (g + u * nthr >= n_gathers) break;
			if (BYTES == 1) acc += __ldg(p[u]);
			else if (BYTES == 2) acc += __ldg((const uint16_t *)p[u]);
			else if (BYTES == 4) acc += __ldg((const uint32_t *)p[u]);
			else if (BYTES == 8) acc += __ldg((const uint64_t *)p[u]);
			else { for (int k = 0; k < BYTES / 16; k++) { const uint4 v = __ldg((const uint4 *)p[u] + k); acc += v.x + v.y + v.z + v.w; } }
		}
	}
	if (acc == 0x123456789abcdefull) atomicAdd(sink, 1ull);
}

extern "C" int dsb_gather_bench(int device, uint64_t table_bytes, uint64_t n_gathers, int bytes_each, double *gbs, double *ms_out)
{
	if (!gbs || table_bytes < (1u << 20) || n_gathers == 0) return DSB_E_ARG;
	DSB_CUDA(cudaSetDevice(device));
	uint8_t *table = nullptr; unsigned long long *sink = nullptr;
	DSB_CUDA(cudaMalloc(&table, table_bytes));
	DSB_CUDA(cudaMalloc(&sink, 8));
	DSB_CUDA(cudaMemset(table, 1, table_bytes));
	DSB_CUDA(cudaMemset(sink, 0, 8));
	cudaEvent_t e0, e1;
	DSB_CUDA(cudaEventCreate(&e0)); DSB_CUDA(cudaEventCreate(&e1));
	cudaDeviceProp prop; DSB_CUDA(cudaGetDeviceProperties(&prop, device));
	const int blocks = prop.multiProcessorCount * 8, threads = 256;      // 2048 threads per SM x 4-8 gathers in flight each
	float best = 1e30f;
	for (int rep = 0; rep < 4; rep++) {
		const uint64_t n_slots = table_bytes / bytes_each, seed = 0x9e3779b97f4a7c15ull * (rep + 1);
		DSB_CUDA(cudaEventRecord(e0));
		switch (bytes_each) {
			case 1: k_gather<1><<<blocks, threads>>>(table, n_slots, n_gathers, seed, sink); break;
			case 2: k_gather<2><<<blocks, threads>>>(table, n_slots, n_gathers, seed, sink); break;
			case 4: k_gather<4><<<blocks, threads>>>(table, n_slots, n_gathers, seed, sink); break;
			case 8: k_gather<8><<<blocks, threads>>>(table, n_slots, n_gathers, seed, sink); break;
			case 16: k_gather<16><<<blocks, threads>>>(table, n_slots, n_gathers, seed, sink); break;
			case 32: k_gather<32><<<blocks, threads>>>(table, n_slots, n_gathers, seed, sink); break;
			case 64: k_gather<64><<<blocks, threads>>>(table, n_slots, n_gathers, seed, sink); break;
			case 128: k_gather<128><<<blocks, threads>>>(table, n_slots, n_gathers, seed, sink); break;
			default: cudaFree(table); cudaFree(sink); dsb_set_error("bytes_each must be a power of two in 1..128"); return DSB_E_ARG;
		}
		DSB_CUDA(cudaEventRecord(e1));
		DSB_CUDA(cudaEventSynchronize(e1));
		float ms = 0; DSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
		if (rep > 0 && ms < best) best = ms;
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	cudaFree(table); cudaFree(sink);
	const double sector_bytes = (double)std::max(bytes_each, 32) * (double)n_gathers;
	*gbs = sector_bytes / (best * 1e-3) / 1e9;
	if (ms_out) *ms_out = best;
	return DSB_OK;
}
