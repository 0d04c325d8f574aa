// seed_emul.cpp -- TEST INFRASTRUCTURE ONLY.  Host-side emulation of the warp loop of k_seed: the per-lane handlers of
// desamba_b200/csrc/dsb_seedcore.h are compiled unchanged for the CPU and driven by 32 emulated lanes per emulated warp
// with the same state vote as the kernel.  tests/test_seed_engine.py compares the anchors of every seed with the oracle
// (fast_classify / slow_classify), checks the packed Landau-Vishkin against its plain byte statement, and reads the
// lanes-per-turn statistics the scheduling policy is tuned with.  Nothing of the product links or loads this.
#include "../../desamba_b200/csrc/dsb_seedcore.h"
#include <vector_functions.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>

namespace {

int sc_policy = SC_POLICY, sc_fetch_min = SC_FETCH_MIN;

struct HostIndex {
	DevIndex dev;
	std::vector<uint8_t> occ, ref_bin;
	std::vector<uint64_t> prefix, ref_pos, ref_info;
	std::vector<uint2> sa, uni;
	std::vector<int> q;
	uint64_t n_rows;
};

bool slurp(const std::string &path, size_t skip, std::vector<uint8_t> &out, uint64_t *hdr)
{
	FILE *f = fopen(path.c_str(), "rb");
	if (!f) return false;
	if (skip && fread(hdr, 8, 1, f) != 1) { fclose(f); return false; }
	fseek(f, 0, SEEK_END); const long end = ftell(f); fseek(f, (long)skip, SEEK_SET);
	out.resize((size_t)(end - (long)skip));
	const bool ok = fread(out.data(), 1, out.size(), f) == out.size();
	fclose(f);
	return ok;
}

// the re-cut of the FM blocks (dsb_index.cu k_recut_fm), on the host
void recut(const uint8_t *blocks, uint64_t nb, std::vector<uint8_t> &lines)
{
	lines.assign((2 * nb + 2) * 128, 0);
	uint64_t cnt[5];
	for (uint64_t b = 0; b < nb; b++) {
		const uint8_t *src = blocks + b * 168;
		memcpy(cnt, src, 40);
		for (int half = 0; half < 2; half++) {
			uint64_t *l = (uint64_t *)(lines.data() + (2 * b + half) * 128);
			const uint8_t *nib = src + 40 + 64 * half;
			for (int k = 0; k < 5; k++) l[k] = cnt[k];
			uint64_t pl[3][2] = {{0, 0}, {0, 0}, {0, 0}};
			for (int i = 0; i < 128; i++) {
				uint32_t v = (nib[i >> 1] >> ((i & 1) << 2)) & 0xf;
				if (v < 5) cnt[v]++;
				if (v > 5) v = 7;
				for (int k = 0; k < 3; k++) if ((v >> k) & 1) pl[k][i >> 6] |= 1ull << (i & 63);
			}
			l[6] = pl[0][0]; l[7] = pl[0][1]; l[8] = pl[1][0]; l[9] = pl[1][1]; l[10] = pl[2][0]; l[11] = pl[2][1];
		}
	}
	uint64_t *l = (uint64_t *)(lines.data() + 2 * nb * 128);
	for (int k = 0; k < 5; k++) l[k] = cnt[k];
	for (int k = 6; k < 12; k++) l[k] = ~0ull;
}

struct Stats { uint64_t turns[SC_N_STATES], lanes[SC_N_STATES]; };

struct Warp {
	SeedLane L[32];
	uint32_t vis1[32][VIS1_SLOTS];
	uint64_t lvs[32][4];
	std::vector<uint64_t> vis1_full, vis2;
	std::vector<MemRst> mem;
	bool dead;
};

} // namespace

extern "C" {

void *emul_open(const char *dir)
{
	HostIndex *h = new HostIndex();
	memset(&h->dev, 0, sizeof h->dev);
	const std::string d = std::string(dir) + "/deSAMBA";
	uint64_t hdr = 0;
	{
		std::vector<uint8_t> bwt;
		if (!slurp(d + ".bwt", 8, bwt, &hdr)) { delete h; return nullptr; }
		const uint64_t byteLen = hdr, nb = byteLen / 168;
		recut(bwt.data(), nb, h->occ);
		h->dev.n_lines = 2 * nb + 1;
		memcpy(h->dev.rank, bwt.data() + byteLen, 40);
		h->dev.rank[5] = h->dev.rank[0] - 1;
		const uint64_t nh = (1ull << 26) + 1;
		h->prefix.resize(nh + 2);
		memcpy(h->prefix.data(), bwt.data() + byteLen + 40, nh * 8);
		h->n_rows = nb * 256;
	}
	std::vector<uint8_t> buf;
	if (!slurp(d + ".sa", 8, buf, &hdr)) { delete h; return nullptr; }
	h->sa.resize(hdr + 2); memcpy(h->sa.data(), buf.data(), hdr * 8);
	if (!slurp(d + ".unv", 8, buf, &hdr)) { delete h; return nullptr; }
	{
		const uint64_t n = hdr;
		h->uni.resize(n + 16); memcpy(h->uni.data(), buf.data(), n * 8);
		h->uni[n].x = h->uni[n - 1].x + 1 + h->uni[n - 1].y; h->uni[n].y = 0;       // idx.c:1127 (as the product's loader)
		h->dev.n_uni = n; h->dev.dollar_pos = n - 2;
	}
	if (!slurp(d + ".ref_b", 8, buf, &hdr)) { delete h; return nullptr; }
	h->dev.ref_bin_n = hdr;
	h->ref_bin.assign(hdr + 1024 + 32, 0); memcpy(h->ref_bin.data(), buf.data(), hdr);
	if (!slurp(d + ".ref_i", 8, buf, &hdr)) { delete h; return nullptr; }
	h->ref_info.resize(hdr * 2);
	for (uint64_t i = 0; i < hdr; i++) { memcpy(&h->ref_info[2 * i], buf.data() + i * 144 + 128, 16); }
	if (!slurp(d + ".ref_p", 8, buf, &hdr)) { delete h; return nullptr; }
	h->ref_pos.resize(hdr + 2); memcpy(h->ref_pos.data(), buf.data(), hdr * 8);
	{	// exist k-mer parameters (idx.c:966-982) -- only l_ek reaches the seeding pass
		FILE *f = fopen((d + ".exki").c_str(), "rb"); uint64_t ek = 0;
		if (!f || fread(&ek, 8, 1, f) != 1) { if (f) fclose(f); delete h; return nullptr; }
		fclose(f);
		int l_ek = 20;
		switch (ek >> 27) { case 1: l_ek = 16; break; case 2: case 4: l_ek = 17; break; case 8: case 16: l_ek = 18; break; case 32: case 64: l_ek = 19; break; }
		h->dev.l_ek = l_ek; h->dev.single_base_max = (int)(0.8 * l_ek);
	}
	{	// MAPQ tables, as dsb_index.cu (cly_mt.c:413-437)
		const double P_E = 0.15; const uint64_t L_REF = h->dev.ref_bin_n * 4;
		const double REF_SIZE_PUNALTY = -10 * log(L_REF) / log(10);
		const double MATCH_SCORE = -10 * log(0.25 / (1 - P_E)) / log(10);
		const double MISMATCH_PUNALTY = -10 * log(0.75 / (P_E)) / log(10);
		h->q.resize(65536 + 400);
		for (int i = 0; i < 65536; i++) h->q[i] = REF_SIZE_PUNALTY + i * MATCH_SCORE + 0.5;
		int *lv = h->q.data() + 65536;
		for (int j = 0; j < 20; j++)
			for (int i = 0; i < 20; i++) {
				int v = (j - i) * MATCH_SCORE + i * MISMATCH_PUNALTY + 0.5;
				if (j < 5) v += 15;
				lv[i * 20 + j] = v > -8 ? v : -8;
			}
	}
	h->dev.occ = h->occ.data(); h->dev.prefix = h->prefix.data(); h->dev.sa = h->sa.data(); h->dev.uni = h->uni.data();
	h->dev.ref_pos = h->ref_pos.data(); h->dev.ref_bin = h->ref_bin.data(); h->dev.ref_info = (const ulonglong2 *)h->ref_info.data();
	h->dev.q_mem = h->q.data(); h->dev.q_lv = h->q.data() + 65536;
	return h;
}
void emul_close(void *h) { delete (HostIndex *)h; }
int emul_l_ek(void *h) { return ((HostIndex *)h)->dev.l_ek; }
void emul_set_policy(int policy, int fetch_min) { sc_policy = policy; sc_fetch_min = fetch_min; }
int emul_n_states(void) { return SC_N_STATES; }

// One seeding pass over a task list.  seqs/offs: ASCII reads; seeds0/seeds1 + seed_off: the island seeds of the two strands
// (from the oracle); tasks: the seeds to run.  Outputs: recs[n_tasks]; the staged anchors of task t at anc[anc_off[t] ..
// anc_off[t] + recs[t].count) as {ref_ID, ref_offset, index_in_read, mtch_len | score << 16}; stats[2 * SC_N_STATES] =
// turns and busy lanes per state.  Returns the number of anchors or -1 when anc_cap is too small.
int64_t emul_seed_pass(void *h_, const char *seqs, const uint64_t *offs, uint32_t n_reads,
                       const dsb_seed *seeds0, const dsb_seed *seeds1, const uint32_t *seed_off,
                       const SeedTaskRef *tasks, uint32_t n_tasks, int slow, int n_warps, int big_rows,
                       SeedRec *recs, uint64_t *anc_off, uint4 *anc, uint64_t anc_cap, uint64_t *stats)
{
	HostIndex *h = (HostIndex *)h_;
	// packed forward strands in the layout of k_encode_probe
	std::vector<uint64_t> bits_off(n_reads + 1);
	uint64_t wo = 0;
	for (uint32_t r = 0; r < n_reads; r++) {
		const uint32_t len = (uint32_t)(offs[r + 1] - offs[r]);
		bits_off[r] = wo;
		if (len >= 40) wo += 5ull * ((len + 31) / 32 + 1);
	}
	bits_off[n_reads] = wo;
	std::vector<uint64_t> pk(wo / 5 + 4, 0x5a5a5a5a5a5a5a5aull);           // (padding words hold junk on purpose)
	for (uint32_t r = 0; r < n_reads; r++) {
		const uint32_t len = (uint32_t)(offs[r + 1] - offs[r]);
		if (len < 40) continue;
		uint64_t *w = pk.data() + bits_off[r] / 5 + 1;
		for (uint32_t k = 0; k < (len + 31) / 32; k++) w[k] = 0;
		for (uint32_t i = 0; i < len; i++) {
			uint32_t c;
			switch (seqs[offs[r] + i]) { case 'A': case 'a': c = 0; break; case 'G': case 'g': c = 2; break; case 'T': case 't': c = 3; break; default: c = 1; }
			w[i >> 5] |= (uint64_t)c << (62 - 2 * (i & 31));
		}
	}
	SeedEnv E;
	E.ix = h->dev; E.pk = pk.data(); E.read_off = offs; E.bits_off = bits_off.data(); E.seed_off = seed_off;
	E.seeds[0] = seeds0; E.seeds[1] = seeds1; E.tasks = tasks; E.recs = recs;
	std::vector<uint4> chunks((size_t)(anc_cap / 2 + 1024) * 4);
	E.chunks = chunks.data(); E.n_chunks = (uint32_t)(chunks.size() / 4);
	E.slow = slow; E.big_rows = big_rows;
	uint32_t chunk_cursor = 0, task_cursor = 0;
	std::vector<Warp> W(n_warps);
	for (auto &w : W) {
		w.vis1_full.assign(32 * VIS1_SLOTS, 0); w.vis2.assign(32 * VIS2_SLOTS, 0); w.mem.resize(32 * SEED_MEM_SLOTS);
		memset(w.L, 0, sizeof w.L);
		for (int l = 0; l < 32; l++) { w.L[l].st = ST_FETCH; w.L[l].vis_gen = 0; }
		w.dead = false;
	}
	Stats S; memset(&S, 0, sizeof S);
	int n_dead = 0;
	bool overflow = false;
	while (n_dead < n_warps) {
		for (auto &w : W) {                                    // the warps take turns, one state turn each
			if (w.dead) continue;
			int cnt[SC_N_STATES] = {0};
			for (int l = 0; l < 32; l++) cnt[w.L[l].st]++;
			uint32_t best = 0;
			for (int l = 0; l < 32; l++) { const uint32_t key = vote_key(w.L[l].st, cnt[w.L[l].st], sc_policy, sc_fetch_min); if (key > best) best = key; }
			const int sel = (int)vote_state(best);
			if (sel == ST_DEAD) { w.dead = true; n_dead++; continue; }
			S.turns[sel]++; S.lanes[sel] += cnt[sel];
			for (int l = 0; l < 32; l++) {
				SeedLane &L = w.L[l];
				if ((int)L.st != sel) continue;
				LaneMem M; M.vis1 = w.vis1[l]; M.lvs = w.lvs[l]; M.vis1_full = w.vis1_full.data() + l * VIS1_SLOTS; M.vis2 = w.vis2.data() + l * VIS2_SLOTS; M.mem = w.mem.data() + l * SEED_MEM_SLOTS;
				switch (sel) {
					case ST_FETCH: { const uint32_t t = task_cursor++; if (t < n_tasks) task_begin(E, L, M, t); else L.st = ST_DEAD; break; }
					case ST_CTRL: h_ctrl(E, L, M); break;
					case ST_OCC: h_occ(E, L, M); break;
					case ST_LOCATE: h_locate(E, L); break;
					case ST_FLANK: h_flank(E, L, M); break;
					case ST_LV: h_lv(E, L, M); break;
					case ST_RP: h_rp(E, L); break;
				}
			}
			// the pending anchor pushes of this turn (collective on the device: the chunk allocation)
			for (int l = 0; l < 32; l++) {
				SeedLane &L = w.L[l];
				if (!L.push) continue;
				L.push = 0;
				const uint32_t slot = L.n_out % STAGE_PER_CHUNK;
				if (slot == 0) {
					const uint32_t c = chunk_cursor++;
					if (c >= E.n_chunks) { overflow = true; L.error = 1; map_done(L, 0); continue; }
					E.chunks[c * 4 + 3] = make_uint4(SC_NO_CHUNK, 0, 0, 0);
					if (L.n_out == 0) L.first_chunk = c; else E.chunks[L.cur_chunk * 4 + 3].x = c;
					L.cur_chunk = c;
				}
				E.chunks[L.cur_chunk * 4 + slot] = push_make(E, L);
				L.n_out++;
				rp_next(L);
			}
		}
	}
	if (stats) for (int s = 0; s < SC_N_STATES; s++) { stats[2 * s] = S.turns[s]; stats[2 * s + 1] = S.lanes[s]; }
	if (overflow) return -1;
	uint64_t total = 0;
	for (uint32_t t = 0; t < n_tasks; t++) {
		anc_off[t] = total;
		uint32_t c = recs[t].first_chunk;
		for (uint32_t i = 0; i < recs[t].count; i++) {
			if (i && i % STAGE_PER_CHUNK == 0) c = E.chunks[c * 4 + 3].x;
			if (total >= anc_cap) return -1;
			anc[total++] = E.chunks[c * 4 + i % STAGE_PER_CHUNK];
		}
	}
	return (int64_t)total;
}

// packed Landau-Vishkin against its byte statement: strings with `pre` bytes in front, as the engine passes them
int emul_lv_packed(const uint8_t *ref18, const uint8_t *query18, int len)
{
	uint64_t er = 0, eq = 0;
	for (int k = 0; k < 18; k++) { er |= (uint64_t)(ref18[k] & 3) << (62 - 2 * k); eq |= (uint64_t)(query18[k] & 3) << (62 - 2 * k); }
	return lv_packed(er, eq, len);
}
int emul_lv_bytes(const uint8_t *ref18, const uint8_t *query18, int len) { return lv_bytes(ref18, query18, len); }

}
