// dsb_seed.cuh -- device side of the seeding engine (dsb_seedcore.h): the warp loop of k_seed, the ordered gather of the
// per-seed anchor lists (start of phase_chain) and the task lists of the slow passes.
//
// k_seed is a persistent kernel over a FLAT list of seed tasks (island seeds of all reads of the pass: fast_classify takes
// the top seeds of the chosen strand(s), cly.c:1494-1496; slow_classify all seeds but the short ones, cly.c:1563-1565).
// A lane owns one task at a time; each turn the warp votes for a state (pick_state) and the lanes in that state run its
// handler together.  Anchors go to per-seed chunk lists in a staging pool; phase_chain then copies the lists of a read to
// its anchor vector in the order the reference pushes them (seed order within a strand pass), dropping the seed behind one
// that scored > 512 (cly.c:1530-1531) and setting anchor_useless per seed (cly.c:1536-1542, 1601-1607).
#pragma once

struct SeedPassParams {
	SeedEnv E;
	uint32_t *ctl;                 // control block of the batch (CTL_*)
	int pass;                      // PASS_FAST / PASS_SLOW0 / PASS_SLOW1
	uint32_t task_cap;
	int policy, fetch_min;         // of the state vote (pick_state)
	// per resident lane of the kernel (grid * block threads)
	MemRst *lane_mem; uint64_t *vis2, *vis1_full; uint32_t *vis_gen;
};

#ifndef SEED_WARPS_PER_BLOCK
#define SEED_WARPS_PER_BLOCK 4
#endif

__device__ __forceinline__ void seed_warp_loop(const SeedPassParams &P, uint32_t (*s_vis1)[32], uint64_t (*s_lvs)[32])
{
	const int lane = threadIdx.x & 31;
	const uint32_t glane = (blockIdx.x * blockDim.x) + threadIdx.x;
	const SeedEnv &E = P.E;
	LaneMem M;
	M.vis1 = (uint32_t)__cvta_generic_to_shared(&s_vis1[0][lane]);
	M.lvs = (uint32_t)__cvta_generic_to_shared(&s_lvs[0][lane]);
	M.vis1_full = P.vis1_full + (E.big_rows ? (uint64_t)glane * VIS1_SLOTS : 0);
	M.vis2 = P.vis2 + (uint64_t)glane * VIS2_SLOTS;
	M.mem = P.lane_mem + (uint64_t)glane * SEED_MEM_SLOTS;
	SeedLane L;
	memset(&L, 0, sizeof L);
	L.st = ST_FETCH;
	L.vis_gen = P.vis_gen[glane];
	const uint32_t n_tasks = min(P.ctl[CTL_TASK_N + P.pass], P.task_cap);
	uint32_t *task_cursor = P.ctl + CTL_TASK_CURSOR + P.pass, *chunk_cursor = P.ctl + CTL_CHUNK_CURSOR + P.pass;
	const uint32_t lt = (1u << lane) - 1;
	for (uint32_t turn = 0;; turn++) {
		if (turn > (1u << 26)) { if (lane == 0) atomicOr(P.ctl + CTL_OVERFLOW, OVF_STUCK); break; }     // never hang the GPU on a malformed index
		const int k = __popc(__match_any_sync(DSB_FULL, L.st));
		const int sel = (int)vote_state(__reduce_max_sync(DSB_FULL, vote_key(L.st, k, P.policy, P.fetch_min)));
		if (sel == ST_DEAD) break;
		if (sel == ST_FETCH) {
			const uint32_t fm = __ballot_sync(DSB_FULL, L.st == ST_FETCH);
			uint32_t base = 0;
			if (lane == __ffs(fm) - 1) base = atomicAdd(task_cursor, (uint32_t)__popc(fm));
			base = __shfl_sync(DSB_FULL, base, __ffs(fm) - 1);
			if (L.st == ST_FETCH) {
				const uint32_t t = base + __popc(fm & lt);
				if (t < n_tasks) task_begin(E, L, M, t); else L.st = ST_DEAD;
			}
		} else if (L.st == (uint32_t)sel) {
			switch (sel) {
				case ST_CTRL: h_ctrl(E, L, M); break;
				case ST_OCC: h_occ(E, L, M); break;
				case ST_LOCATE: h_locate(E, L); break;
				case ST_FLANK: h_flank(E, L, M); break;
				case ST_LV: h_lv(E, L, M); break;
				default: h_rp(E, L); break;
			}
		}
		__syncwarp();
		// pending anchor pushes of this turn: a staging chunk holds 3 anchors, new chunks are taken with one atomic per warp
		const uint32_t pm = __ballot_sync(DSB_FULL, L.push != 0);
		if (pm) {
			const bool want = L.push != 0;
			const bool need = want && (L.n_out % STAGE_PER_CHUNK) == 0;
			const uint32_t nm = __ballot_sync(DSB_FULL, need);
			uint32_t base = 0;
			if (nm) {
				if (lane == __ffs(nm) - 1) base = atomicAdd(chunk_cursor, (uint32_t)__popc(nm));
				base = __shfl_sync(DSB_FULL, base, __ffs(nm) - 1);
			}
			if (want) {
				L.push = 0;
				bool ok = true;
				if (need) {
					const uint32_t c = base + __popc(nm & lt);
					if (c >= E.n_chunks) { ok = false; atomicOr(P.ctl + CTL_OVERFLOW, OVF_CHUNKS); L.error = 1; map_done(L, 0); }
					else {
						E.chunks[(uint64_t)c * 4 + 3] = make_uint4(SC_NO_CHUNK, 0, 0, 0);
						if (L.n_out == 0) L.first_chunk = c; else E.chunks[(uint64_t)L.cur_chunk * 4 + 3].x = c;
						L.cur_chunk = c;
					}
				}
				if (ok) {
					E.chunks[(uint64_t)L.cur_chunk * 4 + (L.n_out % STAGE_PER_CHUNK)] = push_make(E, L);
					L.n_out++;
					rp_next(L);
				}
			}
		}
	}
	P.vis_gen[glane] = L.vis_gen;
}

// ---------------------------------------------------------------- task list of a slow pass (phase_chain decides it): the seeds
// of one strand of read r that slow_classify visits (cly.c:1561-1565); returns false when the list is full
__device__ __forceinline__ bool slow_tasks_append(const ClassifyParams &P, uint32_t r, uint32_t strand, int pass)
{
	const int lane = lane_id();
	const dsb_seed *sv = P.seeds[strand] + P.seed_off[r];
	const uint32_t n = P.n_seeds[strand][r];
	const bool top0 = n ? (sv[0].top != 0) : false;                // sv_f->top: seed 0's flag, as written (cly.c:1564)
	uint32_t total = 0;
	for (uint32_t k0 = 0; k0 < n; k0 += 32) {
		const uint32_t k = k0 + lane;
		const bool ok = k < n && !((int)sv[k].len < 3 && !top0);
		total += __popc(__ballot_sync(DSB_FULL, ok));
	}
	uint32_t base = 0;
	if (lane == 0 && total) base = atomicAdd(P.ctl + CTL_TASK_N + pass, total);
	base = __shfl_sync(DSB_FULL, base, 0);
	if (lane == 0) { P.task_first[strand][r] = base; P.task_cnt[strand][r] = total; }
	if ((uint64_t)base + total > P.task_cap) {
		if (lane == 0) { atomicOr(P.ctl + CTL_OVERFLOW, OVF_TASKS); P.task_cnt[strand][r] = 0; }
		return false;
	}
	SeedTaskRef *out = P.tasks[pass & 1] + base;
	uint32_t at = 0;
	for (uint32_t k0 = 0; k0 < n; k0 += 32) {
		const uint32_t k = k0 + lane;
		const bool ok = k < n && !((int)sv[k].len < 3 && !top0);
		const uint32_t m = __ballot_sync(DSB_FULL, ok);
		if (ok) { SeedTaskRef t; t.read = r; t.sk = (strand << 31) | k; out[at + __popc(m & ((1u << lane) - 1))] = t; }
		at += __popc(m);
	}
	return true;
}

// ---------------------------------------------------------------- ordered gather: the anchors of one strand pass, appended to dst
// Two sweeps over the records of the strand's tasks: count (what survives the "> 512 skips the next seed" rule), then copy.
struct GatherSum { uint32_t n_anchor; uint32_t c_prefix, c_occ, c_locate, c_getref, c_getref_bytes; int error; };
__device__ __forceinline__ uint32_t gather_dropped(uint32_t F, uint32_t ADJ, uint32_t &carry)
{   // F: task scored > 512; ADJ: task's seed directly follows the seed of the task before it; carry: the task before this group skips its successor
	uint32_t dropped = 0;
	#pragma unroll
	for (int b = 0; b < 32; b++) {
		const uint32_t d = ((ADJ >> b) & 1) ? carry : 0u;
		dropped |= d << b;
		carry = (!d && ((F >> b) & 1)) ? 1u : 0u;
	}
	return dropped;
}
__device__ __noinline__ void gather_strand(const ClassifyParams &P, uint32_t r, uint32_t strand, int pass, DevAnchor *dst, GatherSum &G, bool copy)
{
	const int lane = lane_id();
	const uint32_t t0 = P.task_first[strand][r], n = P.task_cnt[strand][r];
	const SeedTaskRef *tasks = P.tasks[pass & 1];
	const SeedRec *recs = P.recs[pass & 1];
	uint32_t carry = 0, prev_k = 0xfffffff0u;
	uint32_t n_out = G.n_anchor;
	for (uint32_t i0 = 0; i0 < n; i0 += 32) {
		const uint32_t i = i0 + lane;
		SeedRec rec; rec.first_chunk = SC_NO_CHUNK; rec.count = 0; rec.top_score = 35; rec.flag512 = 0; rec.c_occ = rec.c_getref = rec.c_getref_bytes = rec.c_pl = 0;
		uint32_t k = 0xfffffff0u;
		if (i < n) { rec = recs[t0 + i]; k = tasks[t0 + i].sk & 0x7fffffffu; }
		uint32_t kp = __shfl_up_sync(DSB_FULL, k, 1);
		if (lane == 0) kp = prev_k;
		prev_k = __shfl_sync(DSB_FULL, k, 31);
		const uint32_t F = __ballot_sync(DSB_FULL, (rec.flag512 & 1) != 0), ADJ = __ballot_sync(DSB_FULL, i < n && kp + 1 == k);
		const uint32_t dropped = gather_dropped(F, ADJ, carry);
		const bool keep = i < n && !((dropped >> lane) & 1);
		const uint32_t cnt = keep ? rec.count : 0u;
		uint32_t x = cnt;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(DSB_FULL, x, d); if (lane >= d) x += y; }
		const uint32_t total = __shfl_sync(DSB_FULL, x, 31);
		if (!copy) {
			if (keep) {
				G.c_prefix += rec.c_pl & 0xffff; G.c_locate += rec.c_pl >> 16; G.c_occ += rec.c_occ; G.c_getref += rec.c_getref; G.c_getref_bytes += rec.c_getref_bytes;
				if (rec.flag512 >> 8) G.error = (int)(rec.flag512 >> 8);
			}
		} else if (cnt) {
			uint32_t at = n_out + x - cnt, ch = rec.first_chunk;
			for (uint32_t a = 0; a < cnt; a++) {
				if (a && (a % STAGE_PER_CHUNK) == 0) ch = P.chunks[(uint64_t)ch * 4 + 3].x;
				const uint4 s = P.chunks[(uint64_t)ch * 4 + (a % STAGE_PER_CHUNK)];
				DevAnchor A;
				A.ref_ID = s.x; A.ref_offset = s.y; A.index_in_read = s.z; A.pre = -1;
				A.mtch_len = (uint16_t)(s.w & 0xffff); A.score = (int16_t)(s.w >> 16);
				A.direction = strand ? DSB_REVERSE : DSB_FORWARD;
				A.useless = ((int)A.score < rec.top_score) ? 1 : 0;        // cly.c:1536-1542
				A.duplicate = 0; A.pad = 0;
				dst[at + a] = A;
			}
		}
		n_out += total;
	}
	G.n_anchor = n_out;
}
