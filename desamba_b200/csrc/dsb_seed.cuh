// dsb_seed.cuh -- seeding: FM-index backward search, locate, Landau-Vishkin flank scoring -> anchors.
//
// Execution model: ONE LANE PER ISLAND SEED.  A seed is the reference's unit of sequential work (its k-mer loop has
// data-dependent strides and shares one visited-row set, cly.c:1497-1542 / 1563-1608); different seeds of a read are
// independent except for (a) the order in which their anchors are appended and (b) fast mode's "a seed that scored > 512
// makes the next seed be skipped" rule (cly.c:1530-1531).  So the 32 lanes of the read's warp each pull seeds from a
// shared counter and run the whole per-seed search privately (32 dependent-load chains in flight per warp instead of 1);
// anchors go to lane-private chunk lists in a staging pool, and a warp-cooperative pass then drops the seeds the skip rule
// removes and copies the lists to the anchor array in seed order -- the same array the reference builds serially.
// Everything in the first half of this file is lane-private code: no warp collectives, lanes diverge freely.
#pragma once

struct SeedRec { uint32_t first_chunk, count; int32_t top_score; uint32_t flag512; uint32_t c_prefix, c_occ, c_locate, c_getref, c_getref_bytes, pad[3]; };
#define ANCHOR_CHUNK 8

struct LaneCtx {
	const DevIndex *ix;
	uint64_t *sp_set;              // visited-row set: open-addressing table of SP_TAB slots, interleaved across lanes (slot i at sp_set[i * 32])
	int sp_l;                      // SP_SET.l: rows in the set
	uint32_t sp_gen;               // generation tag of the live entries (clearing the set = a new generation)
	MemRst *mem;                   // 256 results + 256 merge-sort scratch (slow mode)
	DevAnchor *pool; uint32_t *chunk_next; uint32_t *chunk_cursor; uint32_t n_chunks;
	uint32_t first_chunk, cur_chunk, n_out; int top_score;
	DevAnchor *lin; uint32_t lin_cap;   // lane-per-READ mode (short reads): anchors go to a lane-private linear buffer instead
	int error;
	uint32_t c_prefix, c_occ, c_locate, c_getref, c_getref_bytes;
	uint8_t fr[64];                // pad[8] | q_pre[13] | t_pre[13] | t_suf[13]  (frame layout policy P2 of the oracle)
};

// ---------------------------------------------------------------- occ (bwt.c:43-65) by one lane
// One 128-byte line per 128 BWT symbols: u64 cnt[5] | pad | three 128-bit bit-planes of the symbols at bytes 48/64/80
// (dsb_device.cuh).  The loads of a call are independent of each other; counting is ~20 integer instructions.
__device__ __forceinline__ uint32_t plane_count(const uint4 &p0, const uint4 &p1, const uint4 &p2, int in, uint32_t c)
{   // number of symbols equal to c among the first `in` (0..127) symbols of the line
	const uint64_t a0 = (uint64_t)p0.x | ((uint64_t)p0.y << 32), a1 = (uint64_t)p0.z | ((uint64_t)p0.w << 32);
	const uint64_t b0 = (uint64_t)p1.x | ((uint64_t)p1.y << 32), b1 = (uint64_t)p1.z | ((uint64_t)p1.w << 32);
	const uint64_t c0 = (uint64_t)p2.x | ((uint64_t)p2.y << 32), c1 = (uint64_t)p2.z | ((uint64_t)p2.w << 32);
	const uint64_t x0 = (c & 1) ? 0ull : ~0ull, x1 = (c & 2) ? 0ull : ~0ull, x2 = (c & 4) ? 0ull : ~0ull;
	const uint64_t m_lo = (in >= 64) ? ~0ull : ((1ull << in) - 1);
	const uint64_t m_hi = (in > 64) ? ((1ull << (in - 64)) - 1) : 0ull;
	return __popcll((a0 ^ x0) & (b0 ^ x1) & (c0 ^ x2) & m_lo) + __popcll((a1 ^ x0) & (b1 ^ x1) & (c1 ^ x2) & m_hi);
}
// occ(r, c) for a known symbol c (0..4)
__device__ __noinline__ uint64_t occ_t(const DevIndex &ix, uint64_t r, uint32_t c)
{
	const uint4 *line = (const uint4 *)(ix.occ + (r >> 7) * 128);
	const uint64_t base = __ldg((const uint64_t *)line + c);
	const uint4 p0 = __ldg(line + 3), p1 = __ldg(line + 4), p2 = __ldg(line + 5);
	return base + plane_count(p0, p1, p2, (int)(r & 127), c);
}
// occ(r, *c) with *c == 0xff: takes c = BWT[r] first; '$' (5) returns DOLLOR_POS (bwt.c:50-56)
__device__ __noinline__ uint64_t occ_char_t(const DevIndex &ix, uint64_t r, uint32_t &c)
{
	const uint4 *line = (const uint4 *)(ix.occ + (r >> 7) * 128);
	const uint4 h0 = __ldg(line), h1 = __ldg(line + 1);
	const uint64_t h4 = __ldg((const uint64_t *)line + 4);
	const uint4 p0 = __ldg(line + 3), p1 = __ldg(line + 4), p2 = __ldg(line + 5);
	const int in = (int)(r & 127);
	const uint32_t w = in >> 5, sh = in & 31;
	const uint32_t q0 = (w == 0) ? p0.x : (w == 1) ? p0.y : (w == 2) ? p0.z : p0.w;
	const uint32_t q1 = (w == 0) ? p1.x : (w == 1) ? p1.y : (w == 2) ? p1.z : p1.w;
	const uint32_t q2 = (w == 0) ? p2.x : (w == 1) ? p2.y : (w == 2) ? p2.z : p2.w;
	c = ((q0 >> sh) & 1) | (((q1 >> sh) & 1) << 1) | (((q2 >> sh) & 1) << 2);
	if (c == 5) return ix.dollar_pos;
	uint64_t base;
	switch (c) {
		case 0: base = (uint64_t)h0.x | ((uint64_t)h0.y << 32); break;
		case 1: base = (uint64_t)h0.z | ((uint64_t)h0.w << 32); break;
		case 2: base = (uint64_t)h1.x | ((uint64_t)h1.y << 32); break;
		case 3: base = (uint64_t)h1.z | ((uint64_t)h1.w << 32); break;
		default: base = h4; break;
	}
	return base + plane_count(p0, p1, p2, in, c);
}

// ---------------------------------------------------------------- visited-row set (sp_set_insert, cly.c:1286-1298)
// The reference keeps <= 500 rows in an array and scans it linearly on every insert; the set is emptied when it is full
// and at the start of every seed.  Same semantics here with a hash table: slot = (generation << 40) | row (rows of a
// BWT are < 2^40, like REF_POS.global_offset), 0 = never used; emptying the set = next generation, stale slots read as
// free.  (The linear scan was 69 % of the instructions of the seeding kernel on short reads, profiles/r1d_*.)
#define SP_TAB 1024
#define SP_SMALL 128            // the first SP_SMALL / 2 rows of a set live in a table of this size (1 KB per lane, cache
                                // resident; most seeds never need more) -- only later rows go to the SP_TAB-slot table behind it
__device__ __forceinline__ void sp_set_clear_t(LaneCtx &L)
{
	L.sp_l = 0;
	if (++L.sp_gen >= (1u << 24)) {
		for (int i = 0; i < SP_SMALL + SP_TAB; i++) L.sp_set[i * 32] = 0;
		L.sp_gen = 1;
	}
}
__device__ __noinline__ int sp_set_insert_t(LaneCtx &L, uint64_t node)
{
	if (L.sp_l == SP_SET_CAP) sp_set_clear_t(L);
	const uint64_t key = ((uint64_t)L.sp_gen << 40) | (node & 0xFFFFFFFFFFull);
	const uint64_t hh = node * 0x9E3779B97F4A7C15ull;
	uint32_t h = (uint32_t)(hh >> 57);                                       // top 7 bits
	for (;;) {                                                               // (at most SP_SMALL / 2 live slots: a free one exists)
		const uint64_t v = L.sp_set[h * 32];
		if (v == key) return 0;
		if ((uint32_t)(v >> 40) != L.sp_gen) break;
		h = (h + 1) & (SP_SMALL - 1);
	}
	if (L.sp_l < SP_SMALL / 2) { L.sp_set[h * 32] = key; L.sp_l++; return 1; }
	uint64_t *big = L.sp_set + SP_SMALL * 32;
	h = (uint32_t)(hh >> 54);                                                // top 10 bits
	for (;;) {
		const uint64_t v = big[h * 32];
		if (v == key) return 0;
		if ((uint32_t)(v >> 40) != L.sp_gen) { big[h * 32] = key; L.sp_l++; return 1; }
		h = (h + 1) & (SP_TAB - 1);
	}
}

// ---------------------------------------------------------------- FM-index search (cly.c:1344-1447)
__device__ __noinline__ void bwt_single_search_t(LaneCtx &L, uint64_t sp, const uint8_t *string, int max_match_len, MemRst *out)
{
	const DevIndex &ix = *L.ix;
	uint64_t new_sp, sa_sp = NO_SA;
	int match_len = 0, sa_sp_l = 0;
	while (1) {
		if (match_len >= max_match_len) break;
		if ((sp & SA_MASK) == 0) { sa_sp = sp; sa_sp_l = 0; }
		else sa_sp_l--;
		uint32_t c;
		new_sp = occ_char_t(ix, sp, c);
		new_sp += ix.rank[c];
		L.c_occ++;
		if (c != (uint32_t)__ldg(string)) break;
		match_len++;
		string--;
		asm volatile("prefetch.global.L1 [%0];" :: "l"(ix.occ + (new_sp >> 7) * 128));   // the next step's FM line, while the set is probed
		if (sp_set_insert_t(L, new_sp) == 0) { out->match_len = -1000; return; }
		sp = new_sp;
	}
	out->sp = sp; out->match_len = match_len; out->sa_sp = sa_sp; out->sa_sp_l = sa_sp_l;
}

__device__ __noinline__ int bwt_MEM_search_t(LaneCtx &L, const uint8_t *string, uint64_t pre_v, int max_rst, int l_min_mth, int l_max_mth, MemRst *mem_rst)
{
	const DevIndex &ix = *L.ix;
	int n_rst = 0;
	const ulonglong2 pe = make_ulonglong2(__ldg(ix.prefix + pre_v), __ldg(ix.prefix + pre_v + 1));
	uint64_t sp = pe.x, ep = pe.y, new_sp, new_ep;
	L.c_prefix++;
	string -= L_PRE_IDX;
	int match_len = L_PRE_IDX;
	while (1) {
		const uint32_t c = __ldg(string);
		string--;
		new_sp = ix.rank[c] + occ_t(ix, sp, c);
		new_ep = ix.rank[c] + occ_t(ix, ep, c);
		L.c_occ += 2;
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
		if (sp_set_insert_t(L, new_sp) == 0) return 0;
		bwt_single_search_t(L, new_sp, string, DSB_MAX(0, l_max_mth - match_len), mem_rst + n_rst);
		mem_rst[n_rst].match_len += match_len + 1;
		if (mem_rst[n_rst].match_len >= l_min_mth) n_rst++;
	} else {
		for (uint64_t c_sp = new_sp; c_sp < new_ep; c_sp++) {
			if (sp_set_insert_t(L, c_sp) == 0) continue;
			bwt_single_search_t(L, c_sp, string, DSB_MAX(0, l_max_mth - match_len), mem_rst + n_rst);
			mem_rst[n_rst].match_len += match_len + 1;
			if (mem_rst[n_rst].match_len >= l_min_mth) n_rst++;
		}
	}
	return n_rst;
}

// ---------------------------------------------------------------- locate + anchors (cly.c:435-496, 629-694, 706-939)
__device__ __noinline__ void get_ref_t(LaneCtx &L, uint8_t *out, int64_t off, int32_t length, bool forward)
{   // get_ref, cly.c:435-466
	if (off < 0) off = 0;
	if (length < 0) length = 0;
	L.c_getref++; L.c_getref_bytes += ((uint32_t)length + 3) >> 2;
	const uint64_t o = (uint64_t)off;
	const DevIndex &ix = *L.ix;
	if (length == 0) return;
	// <= 16 bases well inside the packed reference (every call of the seeding path): two aligned words instead of a load per base
	const uint64_t x_hi = forward ? o + (uint32_t)(length - 1) : o;
	if (length <= 16 && (forward || o >= (uint64_t)(length - 1)) && (x_hi >> 2) + 8 < ix.ref_bin_n + 1024) {
		const uint64_t x_lo = forward ? o : o - (uint32_t)(length - 1);
		const uint64_t w = (x_lo >> 2) & ~3ull;                          // byte offset of the first word
		const uint32_t w0 = __ldg((const uint32_t *)(ix.ref_bin + w)), w1 = __ldg((const uint32_t *)(ix.ref_bin + w + 4));
		const uint64_t v = ((uint64_t)__byte_perm(w0, 0, 0x0123) << 32) | __byte_perm(w1, 0, 0x0123);   // base 4w + d at bits 63-2d, 62-2d
		const uint32_t d0 = (uint32_t)(o - 4 * w);
		for (int k = 0; k < length; k++) { const uint32_t d = forward ? d0 + k : d0 - k; out[k] = (uint8_t)((v >> (62 - 2 * d)) & 3); }
		return;
	}
	for (uint32_t k = 0; k < (uint32_t)length; k++) out[k] = (uint8_t)ref_base_at(ix, forward ? o + k : o - k);
}

__device__ __noinline__ int64_t get_uni_t(LaneCtx &L, uint64_t bwt_pos, int search_l, uint64_t *global_offset, uint32_t *uni_offset_)
{   // get_uni, cly.c:471-496
	const DevIndex &ix = *L.ix;
	L.c_locate++;
	const uint2 sa = __ldg(ix.sa + (bwt_pos >> SA_OFF));
	int64_t u = sa.x;
	uint32_t uni_offset = sa.y + search_l + 1;
	if (search_l > 0)
		for (;;) { const uint32_t len = __ldg(ix.uni + u).y; if (!(uni_offset >= len) || u >= (int64_t)ix.n_uni) break; uni_offset -= (len + 1); u++; }   // (bound: the reference walks off its table here)
	const uint64_t rp = __ldg(ix.ref_pos + __ldg(ix.uni + u).x);
	*global_offset = (rp & 0xFFFFFFFFFFull) + uni_offset;
	*uni_offset_ = uni_offset;
	return u;
}

__device__ __noinline__ void get_new_ed_t(LaneCtx &L, uint32_t *e_d, uint32_t *len_, uint32_t *l_mem_ext,
                                          int32_t q_off, uint64_t t_off, uint32_t l_read, const uint8_t *q_b, bool is_FWD)
{   // get_new_ed, cly.c:629-694; q_buff / t_buff reuse the t_pre / t_suf slots of the frame and start zeroed
	uint8_t *fr = L.fr;
	for (int k = 0; k < 13; k++) { fr[FR_B + k] = 0; fr[FR_C + k] = 0; }
	const uint8_t *q = fr + FR_B; uint8_t *t = fr + FR_C;
	uint32_t len, max_len;
	if (is_FWD) {
		if (q_off < 0) q_off = 0;
		max_len = q_off;
		len = DSB_MIN(12, max_len);
		for (uint32_t k = 0; k < len; k++) fr[FR_B + k] = __ldg(q_b + q_off - k);
	} else {
		max_len = l_read - q_off;
		len = DSB_MIN(12, max_len);
		q = q_b + q_off;
	}
	get_ref_t(L, t, t_off, len, !is_FWD);
	if (len > 0 && t[0] == q[0]) {
		int mtc;
		do {
			for (mtc = 0; mtc < len; mtc++) if (t[mtc] != q[mtc]) break;
			if (mtc > 0) {
				*l_mem_ext += mtc;
				max_len -= mtc;
				len = DSB_MIN(12, max_len);
				if (is_FWD) {
					q_off -= mtc; t_off -= mtc;
					for (uint32_t k = 0; k < len; k++) fr[FR_B + k] = __ldg(q_b + q_off - k);
				} else { t_off += mtc; q += mtc; }
				get_ref_t(L, t, t_off, len, !is_FWD);
			}
		} while (mtc > 0);
	}
	*e_d = lv_extd_dev(t, len, q, len);
	*len_ = len;
}

__device__ __forceinline__ bool anchor_push_t(LaneCtx &L, const DevAnchor &a)
{
	if (L.lin) {
		if (L.n_out >= L.lin_cap) { L.error = 6; return false; }      // the read is redone by the warp-per-read path
		L.lin[L.n_out++] = a;
		L.top_score = DSB_MAX(L.top_score, (int)a.score);
		return true;
	}
	if ((L.n_out & (ANCHOR_CHUNK - 1)) == 0) {
		const uint32_t c = atomicAdd(L.chunk_cursor, 1u);
		if (c >= L.n_chunks) { L.error = 1; return false; }
		L.chunk_next[c] = 0xffffffffu;
		if (L.n_out == 0) L.first_chunk = c; else L.chunk_next[L.cur_chunk] = c;
		L.cur_chunk = c;
	}
	L.pool[L.cur_chunk * ANCHOR_CHUNK + (L.n_out & (ANCHOR_CHUNK - 1))] = a;
	L.n_out++;
	L.top_score = DSB_MAX(L.top_score, (int)a.score);
	return true;
}

struct SeedInfo { const uint8_t *bin_read; uint32_t read_L; uint32_t direction; };

#define MIN_S_1 12
#define MIN_S_2 20
__device__ __noinline__ int32_t map_seed_t(LaneCtx &L, const MemRst *m_r, const SeedInfo &s_i)
{   // map_seed, cly.c:706-939
	const DevIndex &ix = *L.ix;
	uint64_t b_p = m_r->sp;
	const int32_t q_off = m_r->read_offset;
	uint32_t l_m = m_r->match_len;
	const uint8_t *q_b = s_i.bin_read;
	int64_t uni = -1;
	uint32_t u_off = 0;
	uint64_t t_off = 0;
	uint32_t l_pre, l_suf = 0, d_pre, d_suf = 0;
	int32_t s = 0, max_s = 0;
	uint8_t *fr = L.fr;
	for (int k = 0; k < 16; k++) ((uint32_t *)fr)[k] = 0;                 // the frame starts zeroed (trivial-auto-var-init)
	do {
		uint8_t *q_pre = fr + FR_A, *t_pre = fr + FR_B, *t_suf = fr + FR_C;
		const uint8_t *q_suf;
		l_pre = DSB_MIN(q_off + 1, LV_L);
		for (uint32_t k = 0; k < l_pre; k++) q_pre[k] = __ldg(q_b + q_off - k);
		int s_l = 0;
		if (m_r->sa_sp != NO_SA)
			uni = get_uni_t(L, m_r->sa_sp, m_r->sa_sp_l, &t_off, &u_off);
		else {
			uint32_t c; uint64_t new_sp;
			while (1) {
				if ((b_p & SA_MASK) == 0) break;
				new_sp = occ_char_t(ix, b_p, c);
				new_sp += ix.rank[c];
				L.c_occ++;
				if (c == 4) break;
				t_pre[s_l++] = (uint8_t)c;
				b_p = new_sp;
				if (s_l >= l_pre) break;
			}
			if ((b_p & SA_MASK) == 0) uni = get_uni_t(L, b_p, s_l, &t_off, &u_off);
			else l_pre = s_l;
		}
		if (uni >= 0) {
			if (__ldg(ix.uni + uni).y < MIN_UNI_L) break;
			l_pre = DSB_MIN(l_pre, u_off);
			get_ref_t(L, t_pre, t_off - 1, l_pre, false);
		}
		d_pre = lv_extd_dev(t_pre, l_pre, q_pre, l_pre);
		s = Q_MEM_at(ix, l_m) + Q_LV_at(ix, d_pre, l_pre);
		if (s < MIN_S_1 && l_pre == LV_L && uni < 0) { s = 0; break; }
		if (uni < 0) {
			for (int guard = 0; b_p & SA_MASK; guard++) {
				if (guard > (1 << 20)) { L.error = 5; return 0; }          // cannot happen on a well-formed index; never hang the GPU
				uint32_t c;
				b_p = occ_char_t(ix, b_p, c);
				b_p += ix.rank[c];
				L.c_occ++;
				s_l++;
			}
			uni = get_uni_t(L, b_p, s_l, &t_off, &u_off);
			if (__ldg(ix.uni + uni).y < MIN_UNI_L) { s = 0; break; }
		}
		const int32_t q_off_r = q_off + l_m + 1;
		uint32_t l_max_suf = DSB_MIN(__ldg(ix.uni + uni).y - u_off - l_m, s_i.read_L - q_off_r);
		if (l_max_suf != 0) {
			l_suf = DSB_MIN(l_max_suf, LV_L);
			q_suf = q_b + q_off_r;
			get_ref_t(L, t_suf, t_off + l_m, l_suf, true);
			if (t_suf[0] == __ldg(q_suf)) {
				int mtc;
				do {
					for (mtc = 0; mtc < l_suf; mtc++) if (t_suf[mtc] != __ldg(q_suf + mtc)) break;
					if (mtc > 0) {
						l_m += mtc;
						s = Q_MEM_at(ix, l_m) + Q_LV_at(ix, d_pre, l_pre);
						l_max_suf -= mtc;
						l_suf = DSB_MIN(l_max_suf, LV_L);
						q_suf += mtc;
						get_ref_t(L, t_suf, t_off + l_m, l_suf, true);
					}
				} while (mtc > 0);
			}
			d_suf = lv_extd_dev(t_suf, l_suf, q_suf, l_suf);
			s += Q_LV_at(ix, d_suf, l_suf);
		} else
			l_suf = d_suf = 0;
		if (s <= MIN_S_2 && l_suf == LV_L) { s = 0; break; }
	} while (0);

	if (s > 0) {
		uint16_t am_mtch_len = (uint16_t)l_m; int16_t am_score = (int16_t)s;
		uint8_t am_left_len = (uint8_t)l_pre, am_left_ED = (uint8_t)d_pre, am_rigt_len = (uint8_t)l_suf, am_rigt_ED = (uint8_t)d_suf;
		const uint32_t r_p_s = __ldg(ix.uni + uni).x, r_p_e = __ldg(ix.uni + uni + 1).x;
		const bool ref_search_l = (l_pre < LV_L || d_pre == 0);
		const bool ref_search_r = (l_suf < LV_L || d_suf == 0);
		if ((int64_t)r_p_e - (int64_t)r_p_s > 50)
			if (!((int64_t)r_p_e - (int64_t)r_p_s < 1000)) return 50;
		for (uint32_t c_r_p = r_p_s; c_r_p < r_p_e; c_r_p++) {
			const uint64_t rp = __ldg(ix.ref_pos + c_r_p);
			const uint64_t rp_global = rp & 0xFFFFFFFFFFull; const uint32_t rp_ref = (uint32_t)((rp >> 40) & 0x7FFFFF);
			uint32_t ed_l, ed_r, len_l, len_r;
			uint32_t l_m_ext_l = 0, l_m_ext_r;
			if (ref_search_l || ref_search_r) {
				if (ref_search_l) {
					get_new_ed_t(L, &ed_l, &len_l, &l_m_ext_l, q_off, rp_global + u_off - 1, s_i.read_L, q_b, true);
					am_left_len = (uint8_t)len_l; am_left_ED = (uint8_t)ed_l;
				}
				am_mtch_len = (uint16_t)(l_m + l_m_ext_l);
				if (ref_search_r) {
					l_m_ext_r = 0;
					get_new_ed_t(L, &ed_r, &len_r, &l_m_ext_r, q_off + l_m + 1, rp_global + u_off + l_m, s_i.read_L, q_b, false);
					am_rigt_len = (uint8_t)len_r; am_rigt_ED = (uint8_t)ed_r;
					am_mtch_len = (uint16_t)(am_mtch_len + l_m_ext_r);
				}
				am_score = (int16_t)(Q_MEM_at(ix, am_mtch_len) + Q_LV_at(ix, am_left_ED, am_left_len) + Q_LV_at(ix, am_rigt_ED, am_rigt_len));
				if (am_score < MIN_S_2) continue;
			}
			max_s = DSB_MAX(max_s, am_score);
			DevAnchor a;
			a.direction = (uint8_t)s_i.direction;
			a.index_in_read = q_off + 1 - l_m_ext_l;
			const uint64_t g = rp_global + u_off - l_m_ext_l;
			a.ref_ID = rp_ref;
			a.ref_offset = (uint32_t)(g - __ldg(ix.ref_info + rp_ref).y);
			a.mtch_len = am_mtch_len; a.score = am_score;
			a.pre = -1; a.useless = 0; a.duplicate = 0; a.pad = 0;
			if (!anchor_push_t(L, a)) return max_s;
		}
	}
	return max_s;
}

// ---------------------------------------------------------------- per-seed search schedules (cly.c:1476-1611)
__device__ __forceinline__ uint64_t prefix13(const uint8_t *bin_read, int string_index)
{   // low 26 bits of the l_ek-mer ending at string_index (= kmer[kmer_index] & PRE_IDX_MASK, cly.c:1504; seeds hold only non-zero k-mers)
	uint64_t v = 0;
	#pragma unroll
	for (int k = 12; k >= 0; k--) v = (v << 2) | __ldg(bin_read + string_index - k);
	return v;
}

#define MEM_search_FAST 2
#define MIN_MEM_LEN_FAST 21
#define MEM_search_SLOW 8
#define MIN_MEM_LEN_SLOW 20
struct MemRstCmp { __device__ int operator()(const MemRst &a, const MemRst &b) const { return b.match_len - a.match_len; } };

// The per-seed schedules of fast_classify (cly.c:1494-1543) and slow_classify (cly.c:1563-1608) as step functions: the
// lanes of a warp work on different seeds, but a warp-uniform outer loop makes every busy lane do ONE step at a time
// (one k-mer search, or one map_seed), so that lanes enter bwt_MEM_search_t / map_seed_t together and their dependent
// loads are in flight at the same time, instead of 32 private instruction streams.
struct SeedTask {
	dsb_seed sv;
	int j;                   // k-mer index inside the island (counts down)
	int stage;               // 0: searching, 1: slow mode -- mapping the sorted MEM results, 2: done
	int n_mem, i_mem;        // slow mode: results collected / next one to map
	uint32_t flag512;
};

__device__ __forceinline__ void seed_task_begin(SeedTask &T, const dsb_seed sv, bool slow, int l_ek)
{
	T.sv = sv; T.j = (int)sv.len - 1; T.stage = 0; T.n_mem = 0; T.i_mem = 0; T.flag512 = 0;
	if (!slow && T.j < MIN_MEM_LEN_FAST - l_ek) T.stage = 2;
	if (slow && T.j < 1) T.stage = 2;
}

// one step of a top seed in fast mode = one iteration of the k-mer loop (cly.c:1500-1534)
__device__ __forceinline__ void fast_seed_step(LaneCtx &L, SeedTask &T, const SeedInfo &s_i)
{
	const int l_ek = L.ix->l_ek;
	const int min_index = MIN_MEM_LEN_FAST - l_ek;
	const uint8_t *bin_read = s_i.bin_read;
	MemRst m_r[MEM_search_FAST];
	const int kmer_index = T.sv.offset + T.j;
	const int string_index = kmer_index + l_ek - 1;
	const uint64_t prefixValue = prefix13(bin_read, string_index);
	const int n = bwt_MEM_search_t(L, bin_read + string_index, prefixValue, MEM_search_FAST, MIN_MEM_LEN_FAST - 1, string_index, m_r);
	if (n == 0) T.j -= 2;
	else {
		T.j -= 3;
		int max_score = 0;
		for (int k = 0; k < n; k++) {
			m_r[k].read_offset = string_index - m_r[k].match_len;
			const int c_score = map_seed_t(L, m_r + k, s_i);
			max_score = DSB_MAX(c_score, max_score);
			if (L.error) { T.stage = 2; return; }
		}
		if (max_score > 35) T.j -= 7;
		if (max_score > 256) {
			if (max_score > 512) T.flag512 = 1;
			T.stage = 2;
			return;
		}
	}
	if (T.j < min_index) T.stage = 2;
}

// one step of a seed in slow mode: one k-mer search (cly.c:1570-1591), then -- after the sort by match length -- one
// map_seed of the at most 8 longest results (cly.c:1595-1600)
__device__ __forceinline__ void slow_seed_step(LaneCtx &L, SeedTask &T, const SeedInfo &s_i)
{
	const int l_ek = L.ix->l_ek;
	MemRst *mem_rst = L.mem;                                    // <= 30 searches * 8 results per seed (len <= 61)
	if (T.stage == 0) {
		const int min_match_len = DSB_MIN(MIN_MEM_LEN_SLOW - 1, l_ek + 1);
		const int k_idx = T.sv.offset + T.j;
		const int s_idx = k_idx + l_ek - 1;
		const uint64_t pre_v = prefix13(s_i.bin_read, s_idx);
		const int n = bwt_MEM_search_t(L, s_i.bin_read + s_idx, pre_v, MEM_search_SLOW, min_match_len, s_idx, mem_rst + T.n_mem);
		for (int k = T.n_mem; k < T.n_mem + n; k++) mem_rst[k].read_offset = k_idx + l_ek - 1 - mem_rst[k].match_len;
		T.n_mem += n;
		T.j -= 2;
		if (T.j < 1) {
			if (T.n_mem == 0) { T.stage = 2; return; }
			if (T.n_mem > 1) glibc_msort(mem_rst, mem_rst + 256, T.n_mem, MemRstCmp());
			T.n_mem = DSB_MIN(T.n_mem, MEM_search_SLOW);
			T.stage = 1;
		}
		return;
	}
	map_seed_t(L, mem_rst + T.i_mem, s_i);
	T.i_mem++;
	if (L.error || T.i_mem >= T.n_mem) T.stage = 2;
}

// ================================================================ warp level
// The seeding jobs in S.sm->job[0 .. n_jobs) (strand passes of one or several reads: fast_classify, cly.c:1476-1546, or
// slow_classify, cly.c:1548-1611): the lanes pull seeds from the combined numbering of all jobs, then the anchors are appended
// to S.ws.anc job by job in seed order with anchor_useless set per seed (cly.c:1536-1542, 1601-1607); job[j].anc_end
// receives S.n_anc after job j.
__device__ __forceinline__ int seed_job_of(const WarpSmem *sm, int n_jobs, uint32_t k)
{
	int j = 0;
	while (j + 1 < n_jobs && k >= sm->job[j + 1].base) j++;
	return j;
}

__device__ __noinline__ void seed_pass(ReadState &S, int n_jobs, bool slow)
{
	WarpSmem *sm = S.sm;
	const int lane = lane_id();
	__syncwarp();
	const uint32_t n_seed = sm->job[n_jobs - 1].base + sm->job[n_jobs - 1].l_seed_v;
	if (slow) S.fast_classify = 0;
	if (lane < n_jobs) sm->job[lane].anc_end = S.n_anc;              // (jobs without seeds keep the count of their predecessor, fixed below)
	__syncwarp();
	if (n_seed == 0) return;
	LaneCtx L;
	L.ix = S.ix; L.sp_set = S.ws.sp_set + lane; L.sp_l = 0; L.sp_gen = S.ws.sp_gen[lane]; L.mem = S.ws.lane_mem + lane * 512;
	L.pool = S.ws.anc_tmp; L.chunk_next = S.ws.chunk_next; L.chunk_cursor = &sm->chunk_cursor; L.n_chunks = S.max_anchors / ANCHOR_CHUNK;
	L.error = 0; L.c_prefix = L.c_occ = L.c_locate = L.c_getref = L.c_getref_bytes = 0; L.lin = nullptr; L.lin_cap = 0;
	SeedInfo s_i = {nullptr, 0, 0};
	SeedRec *rec = S.ws.seed_rec;
	SeedTask T; T.stage = 2;
	uint32_t carry = 0;                                          // 1: the next seed in array order is skipped
	// Every seed that yields an anchor holds at least one chunk of the staging pool, so the seeds are taken in windows small
	// enough for the pool (one window for all but reads of several 100 kb): search the window's seeds, append, reuse the pool.
	const uint32_t win = DSB_MAX(32u, (L.n_chunks / 2) & ~31u);
	for (uint32_t w0 = 0; w0 < n_seed; w0 += win) {
	const uint32_t w1 = DSB_MIN(n_seed, w0 + win);
	__syncwarp();
	if (lane == 0) { sm->next_seed = w0; sm->chunk_cursor = 0; }
	__syncwarp();
	int my_k = -1; bool exhausted = false;
	for (;;) {
		// lanes without a seed pull the next eligible one (ineligible seeds get an empty record on the way)
		while (my_k < 0 && !exhausted) {
			const uint32_t k = atomicAdd(&sm->next_seed, 1u);
			if (k >= w1) { exhausted = true; break; }
			const SeedJob &J = sm->job[seed_job_of(sm, n_jobs, k)];
			const dsb_seed sv = J.seed_v[k - J.base];
			const bool eligible = slow ? !((int)(sv.len) < 3 && J.seed_v[0].top == 0)   // sv_f->top: seed 0's flag, as written (cly.c:1564)
			                           : (sv.top != 0);
			if (eligible && !L.error) {
				my_k = (int)k;
				s_i.bin_read = J.bin_read; s_i.read_L = J.read_len; s_i.direction = J.direction;
				sp_set_clear_t(L); L.n_out = 0; L.first_chunk = 0xffffffffu; L.cur_chunk = 0; L.top_score = 35;
				L.c_prefix = L.c_occ = L.c_locate = L.c_getref = L.c_getref_bytes = 0;
				seed_task_begin(T, sv, slow, L.ix->l_ek);
			} else {
				SeedRec r; r.first_chunk = 0xffffffffu; r.count = 0; r.top_score = 35; r.flag512 = 0;
				r.c_prefix = r.c_occ = r.c_locate = r.c_getref = r.c_getref_bytes = 0; r.pad[0] = r.pad[1] = r.pad[2] = 0;
				rec[k] = r;
			}
		}
		if (__all_sync(DSB_FULL, my_k < 0)) break;
		if (my_k >= 0) {
			if (T.stage != 2) { if (slow) slow_seed_step(L, T, s_i); else fast_seed_step(L, T, s_i); }
			if (T.stage == 2) {
				SeedRec r; r.first_chunk = L.first_chunk; r.count = L.n_out; r.top_score = L.top_score; r.flag512 = T.flag512;
				r.c_prefix = L.c_prefix; r.c_occ = L.c_occ; r.c_locate = L.c_locate; r.c_getref = L.c_getref; r.c_getref_bytes = L.c_getref_bytes;
				r.pad[0] = r.pad[1] = r.pad[2] = 0;
				rec[my_k] = r;
				my_k = -1;
			}
		}
		__syncwarp();
	}
	S.ws.sp_gen[lane] = L.sp_gen;
	__syncwarp();
	L.error = __reduce_max_sync(DSB_FULL, L.error);
	if (L.error) { S.error = L.error; return; }
	// ordered append: drop the seeds removed by the "> 512 skips the next seed" rule (cly.c:1530-1531; it does not reach into
	// the next strand pass), prefix-sum, copy
	uint32_t n_anc = S.n_anc;
	uint32_t c_prefix = 0, c_occ = 0, c_locate = 0, c_getref = 0, c_getref_bytes = 0;      // counters of the seeds the reference would have run
	for (uint32_t base = w0; base < w1; base += 32) {
		const uint32_t k = base + lane;
		SeedRec r; r.first_chunk = 0xffffffffu; r.count = 0; r.top_score = 35; r.flag512 = 0;
		r.c_prefix = r.c_occ = r.c_locate = r.c_getref = r.c_getref_bytes = 0;
		int jk = 0; bool job_first = false, job_last = false;
		if (k < w1) {
			r = rec[k];
			jk = seed_job_of(sm, n_jobs, k);
			job_first = (k == sm->job[jk].base); job_last = (k + 1 == sm->job[jk].base + sm->job[jk].l_seed_v);
		}
		const uint32_t F = __ballot_sync(DSB_FULL, r.flag512 != 0), J1 = __ballot_sync(DSB_FULL, job_first);
		uint32_t skipped = 0;
		#pragma unroll
		for (int b = 0; b < 32; b++) {
			const uint32_t sk = ((J1 >> b) & 1) ? 0u : carry;
			skipped |= sk << b;
			carry = (!sk && ((F >> b) & 1)) ? 1u : 0u;
		}
		const bool drop = (skipped >> lane) & 1;
		const uint32_t cnt = drop ? 0u : r.count;
		if (!drop) { c_prefix += r.c_prefix; c_occ += r.c_occ; c_locate += r.c_locate; c_getref += r.c_getref; c_getref_bytes += r.c_getref_bytes; }
		uint32_t x = cnt;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(DSB_FULL, x, d); if (lane >= d) x += y; }
		const uint32_t total = __shfl_sync(DSB_FULL, x, 31);
		if (n_anc + total > S.max_anchors) { S.error = 1; return; }
		if (job_last) sm->job[jk].anc_end = n_anc + x;
		uint32_t dst = n_anc + x - cnt;
		uint32_t ch = r.first_chunk;
		for (uint32_t i = 0; i < cnt; i++) {
			if (i && (i & (ANCHOR_CHUNK - 1)) == 0) ch = S.ws.chunk_next[ch];
			DevAnchor a = S.ws.anc_tmp[ch * ANCHOR_CHUNK + (i & (ANCHOR_CHUNK - 1))];
			a.useless = (a.score < r.top_score) ? 1 : 0;
			S.ws.anc[dst + i] = a;
		}
		n_anc += total;
	}
	__syncwarp();
	S.n_anc = n_anc;
	S.c_prefix += __reduce_add_sync(DSB_FULL, c_prefix); S.c_occ += __reduce_add_sync(DSB_FULL, c_occ); S.c_locate += __reduce_add_sync(DSB_FULL, c_locate);
	S.c_getref += __reduce_add_sync(DSB_FULL, c_getref); S.c_getref_bytes += __reduce_add_sync(DSB_FULL, c_getref_bytes);
	}
	__syncwarp();
	if (lane == 0)
		for (int j = 0; j < n_jobs; j++) if (sm->job[j].l_seed_v == 0 && j > 0) sm->job[j].anc_end = sm->job[j - 1].anc_end;
	__syncwarp();
}
