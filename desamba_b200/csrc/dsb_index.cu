// dsb_index.cu -- index loader: reads the reference's unchanged on-disk index (load_idx idx.c:1103-1160, load_bwt
// bwt.c:68-104, set_ekmer_par idx.c:966-982, calculate_MAPQ_TABLE cly_mt.c:413-437) and makes it resident in one
// GPU's HBM.  The FM-index blocks are re-cut at load time (see dsb_device.cuh); every other array is uploaded as is.
#include "dsb_internal.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <cmath>
#include <chrono>
#include <new>
#include <exception>

static thread_local char g_err[512] = "";
void dsb_set_error(const char *fmt, ...)
{
	va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
extern "C" const char *dsb_last_error(void) { return g_err; }
extern "C" const char *dsb_version(void) { return "desamba_b200 0.1 (sm_100a)"; }

// FM-index re-cut (see dsb_index_load): thread b turns block b (u64 cnt[5] + 256 nibble symbols) into lines 2b and 2b+1
__global__ void __launch_bounds__(128) k_recut_fm(const uint8_t *blocks, uint64_t nb, uint8_t *lines)
{
	const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nb) return;
	const uint8_t *src = blocks + b * 168;
	uint64_t cnt[5];
	for (int k = 0; k < 5; k++) cnt[k] = ((const uint64_t *)src)[k];
	for (int half = 0; half < 2; half++) {
		uint64_t *l = (uint64_t *)(lines + (2 * b + half) * 128);
		const uint8_t *nib = src + 40 + 64 * half;
		for (int k = 0; k < 5; k++) l[k] = cnt[k];
		uint64_t pl[3][2] = {{0, 0}, {0, 0}, {0, 0}};
		for (int i = 0; i < 128; i++) {
			uint32_t v = (nib[i >> 1] >> ((i & 1) << 2)) & 0xf;
			if (v < 5) cnt[v]++;
			if (v > 5) v = 7;                                          // padding never equals a countable symbol
			for (int k = 0; k < 3; k++) if ((v >> k) & 1) pl[k][i >> 6] |= 1ull << (i & 63);
		}
		l[6] = pl[0][0]; l[7] = pl[0][1]; l[8] = pl[1][0]; l[9] = pl[1][1]; l[10] = pl[2][0]; l[11] = pl[2][1];   // bytes 48 / 64 / 80
	}
	if (b == nb - 1) {                                                 // trailing line: the totals, every symbol = padding
		uint64_t *l = (uint64_t *)(lines + 2 * nb * 128);
		for (int k = 0; k < 5; k++) l[k] = cnt[k];
		for (int k = 6; k < 12; k++) l[k] = ~0ull;
	}
}

// summary of a bit table: bit j of word w = (byte 32 w + j of the table != 0)
__global__ void __launch_bounds__(256) k_byte_summary(const uint8_t *table, uint64_t n_words, uint32_t *sum)
{
	const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= n_words) return;
	const uint4 a = ((const uint4 *)table)[2 * w], b = ((const uint4 *)table)[2 * w + 1];
	const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
	uint32_t m = 0;
	for (int i = 0; i < 8; i++)
		for (int k = 0; k < 4; k++) if ((v[i] >> (8 * k)) & 0xff) m |= 1u << (4 * i + k);
	sum[w] = m;
}

namespace {

struct HostFile {
	FILE *f; std::string path; uint64_t size;
	HostFile(const char *dir, const char *ext) : size(0)
	{
		path = std::string(dir) + "/deSAMBA" + ext; f = fopen(path.c_str(), "rb");
		if (f) { fseeko(f, 0, SEEK_END); size = (uint64_t)ftello(f); fseeko(f, 0, SEEK_SET); }
	}
	~HostFile() { if (f) fclose(f); }
	bool rd(void *p, size_t n) { return f && fread(p, 1, n, f) == n; }
};
struct DevTmp { void *p; DevTmp() : p(nullptr) {} ~DevTmp() { if (p) cudaFree(p); } };      // device scratch of the loader, released on every path

// upload a host array, remember the allocation; `extra` zero bytes are appended
int upload(dsb_index *ix, const void *h, size_t bytes, size_t extra, void **d_out)
{
	void *d = nullptr;
	DSB_CUDA(cudaMalloc(&d, bytes + extra + 16));
	ix->allocs.push_back({d, bytes + extra + 16});
	ix->hbm_bytes += bytes + extra + 16;
	DSB_CUDA(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
	DSB_CUDA(cudaMemset((char *)d + bytes, 0, extra + 16));
	*d_out = d;
	return DSB_OK;
}

// read `<u64 n> n*elem bytes` (or, with have_n, n*elem bytes without a header) and upload
int load_array(dsb_index *ix, const char *dir, const char *ext, size_t elem, bool header, uint64_t *n_io, size_t extra,
               void **d_out, std::vector<uint8_t> *keep = nullptr)
{
	HostFile hf(dir, ext);
	if (!hf.f) { dsb_set_error("cannot open %s", hf.path.c_str()); return DSB_E_IO; }
	if (header && !hf.rd(n_io, 8)) { dsb_set_error("short read %s", hf.path.c_str()); return DSB_E_IO; }
	if (*n_io > hf.size / elem) { dsb_set_error("%s: header says %llu elements, the file has %llu bytes", hf.path.c_str(), (unsigned long long)*n_io, (unsigned long long)hf.size); return DSB_E_IO; }
	const size_t bytes = (size_t)(*n_io) * elem;
	std::vector<uint8_t> local;
	std::vector<uint8_t> &buf = keep ? *keep : local;
	buf.resize(bytes);
	if (!hf.rd(buf.data(), bytes)) { dsb_set_error("short read %s", hf.path.c_str()); return DSB_E_IO; }
	return upload(ix, buf.data(), bytes, extra, d_out);
}

} // namespace

static int load_impl(dsb_index *ix, const char *dir, bool verbose)
{
	auto t_prev = std::chrono::steady_clock::now();
	auto lap = [&](const char *what) { if (!verbose) return; auto t = std::chrono::steady_clock::now(); fprintf(stderr, "[dsb_index_load] %-28s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count()); t_prev = t; };
	int rc = DSB_OK;
	void *d = nullptr;
	// ---- .bwt: u64 byteLen | occ blocks | u64 rank[5] | u64 hash_index[4^13+1]   (bwt.c:75-85)
	{
		HostFile hf(dir, ".bwt");
		uint64_t byteLen = 0;
		const uint64_t nh = (1ull << 26) + 1;
		if (!hf.f || !hf.rd(&byteLen, 8) || byteLen % 168 != 0 || byteLen == 0 || hf.size < 8 + 40 + nh * 8 || byteLen != hf.size - 8 - 40 - nh * 8) { dsb_set_error("cannot read %s (not a deSAMBA FM index)", hf.path.c_str()); return DSB_E_IO; }
		std::vector<uint8_t> blocks(byteLen);
		if (!hf.rd(blocks.data(), byteLen)) { dsb_set_error("short read %s", hf.path.c_str()); return DSB_E_IO; }
		lap("read FM blocks");
		uint64_t rank[6];
		if (!hf.rd(rank, 40)) { dsb_set_error("short read %s", hf.path.c_str()); return DSB_E_IO; }
		rank[5] = rank[0] - 1;                                     // bwt.c:81
		memcpy(ix->dev.rank, rank, sizeof rank);
		const uint64_t nb = byteLen / 168;
		ix->bwt_len_blocks = nb;
		// re-cut ON THE GPU: 168-B blocks of 256 nibble symbols -> 128-B lines of 128 symbols as three bit-planes
		// (dsb_index_view.h), +1 trailing line holding the totals so that occ(len_bwt, c) is defined when len_bwt % 256 == 0
		// (the reference reads past its array there).  One thread per block; a host loop took ~10 ns per symbol, which is
		// half a minute for a bacteria-scale index.
		const uint64_t n_lines = nb * 2 + 1;
		DevTmp d_blocks;
		DSB_CUDA(cudaMalloc(&d_blocks.p, byteLen + 16));
		DSB_CUDA(cudaMemcpy(d_blocks.p, blocks.data(), byteLen, cudaMemcpyHostToDevice));
		blocks.clear(); blocks.shrink_to_fit();
		DSB_CUDA(cudaMalloc(&d, n_lines * 128 + 128 + 16));
		ix->allocs.push_back({d, n_lines * 128 + 128 + 16});
		ix->hbm_bytes += n_lines * 128 + 128 + 16;
		DSB_CUDA(cudaMemset(d, 0, n_lines * 128 + 128 + 16));
		k_recut_fm<<<(unsigned)((nb + 127) / 128), 128>>>((const uint8_t *)d_blocks.p, nb, (uint8_t *)d);
		DSB_CUDA(cudaGetLastError());
		DSB_CUDA(cudaDeviceSynchronize());
		ix->dev.occ = (const uint8_t *)d; ix->dev.n_lines = n_lines;
		lap("upload + re-cut FM blocks (GPU)");
		std::vector<uint64_t> hidx(nh);
		if (!hf.rd(hidx.data(), nh * 8)) { dsb_set_error("short read %s (prefix table)", hf.path.c_str()); return DSB_E_IO; }
		lap("read prefix table");
		if ((rc = upload(ix, hidx.data(), nh * 8, 0, &d)) != DSB_OK) return rc;
		ix->dev.prefix = (const uint64_t *)d;
		lap("upload prefix table");
	}
	// ---- .sa (bwt.c:95-98)
	if ((rc = load_array(ix, dir, ".sa", 8, true, &ix->sa_size, 0, &d)) != DSB_OK) return rc;
	ix->dev.sa = (const uint2 *)d;
	lap("SA");
	// ---- exist k-mer tables (idx.c:1105-1118) + set_ekmer_par (idx.c:966-982: exact size classes, anything else is the largest)
	{
		HostFile hf(dir, ".exki");
		if (!hf.f || !hf.rd(&ix->ek_size, 8) || ix->ek_size == 0) { dsb_set_error("cannot read %s", hf.path.c_str()); return DSB_E_IO; }
		uint64_t mask = (1ull << 37) - 1; int l_ek = 20;
		switch (ix->ek_size) {
			case 1ull << 27: mask = (1ull << 30) - 1; l_ek = 16; break;
			case 1ull << 28: mask = (1ull << 31) - 1; l_ek = 17; break;
			case 1ull << 29: mask = (1ull << 32) - 1; l_ek = 17; break;
			case 1ull << 30: mask = (1ull << 33) - 1; l_ek = 18; break;
			case 1ull << 31: mask = (1ull << 34) - 1; l_ek = 18; break;
			case 1ull << 32: mask = (1ull << 35) - 1; l_ek = 19; break;
			case 1ull << 33: mask = (1ull << 36) - 1; l_ek = 19; break;
			case 1ull << 34: mask = (1ull << 37) - 1; l_ek = 20; break;
		}
		if (mask / 8 >= ix->ek_size) { dsb_set_error("%s: table size %llu does not fit a deSAMBA size class", hf.path.c_str(), (unsigned long long)ix->ek_size); return DSB_E_IO; }
		ix->dev.ek_mask = mask; ix->dev.l_ek = l_ek;
		ix->dev.single_base_max = (int)(0.8 * l_ek);
	}
	{ uint64_t n = ix->ek_size; if ((rc = load_array(ix, dir, ".exk0", 1, false, &n, 0, &d)) != DSB_OK) return rc; ix->dev.ek0 = (const uint8_t *)d; }
	{ uint64_t n = ix->ek_size; if ((rc = load_array(ix, dir, ".exk1", 1, false, &n, 0, &d)) != DSB_OK) return rc; ix->dev.ek1 = (const uint8_t *)d; }
	// A sparse table 0 (viral-scale index: 1 % of the bits, 8 % of the bytes set) gets a summary with one bit per byte: it
	// stays in L2 and answers most probes without a DRAM access (k_encode_probe is DRAM-bound on these probes).
	ix->dev.ek0_sum = nullptr;
	if (ix->ek_size <= (1ull << 28) && ix->ek_size % 32 == 0) {
		void *ds = nullptr;
		DSB_CUDA(cudaMalloc(&ds, ix->ek_size / 8 + 16));
		ix->allocs.push_back({ds, ix->ek_size / 8 + 16});
		ix->hbm_bytes += ix->ek_size / 8 + 16;
		k_byte_summary<<<(unsigned)((ix->ek_size / 32 + 255) / 256), 256>>>(ix->dev.ek0, ix->ek_size / 32, (uint32_t *)ds);
		DSB_CUDA(cudaGetLastError());
		DSB_CUDA(cudaDeviceSynchronize());
		ix->dev.ek0_sum = (const uint32_t *)ds;
	}
	lap("exist k-mer tables");
	// ---- .unv + fabricated sentinel (idx.c:1123-1129)
	{
		HostFile hf(dir, ".unv");
		if (!hf.f || !hf.rd(&ix->n_uni, 8) || ix->n_uni < 2 || ix->n_uni > hf.size / 8) { dsb_set_error("cannot read %s", hf.path.c_str()); return DSB_E_IO; }
		std::vector<uint32_t> u((ix->n_uni + 1) * 2);
		if (!hf.rd(u.data(), ix->n_uni * 8)) { dsb_set_error("short read %s", hf.path.c_str()); return DSB_E_IO; }
		u[ix->n_uni * 2] = u[(ix->n_uni - 1) * 2] + 1 + u[(ix->n_uni - 1) * 2 + 1];
		u[ix->n_uni * 2 + 1] = 0;
		ix->dev.dollar_pos = ix->n_uni - 1 - 1;                    // idx.c:1128
		if ((rc = upload(ix, u.data(), u.size() * 4, 64, &d)) != DSB_OK) return rc;
		ix->dev.uni = (const uint2 *)d; ix->dev.n_uni = ix->n_uni;
	}
	// ---- .ref_b (idx.c:1141-1145); 1 KiB of zero slack behind it for windows that run past the last base
	if ((rc = load_array(ix, dir, ".ref_b", 1, true, &ix->ref_bin_n, 1024, &d)) != DSB_OK) return rc;
	ix->dev.ref_bin = (const uint8_t *)d; ix->dev.ref_bin_n = ix->ref_bin_n;
	// ---- .ref_i (idx.c:1148-1152): host copy for the writers, {seq_l, seq_offset} on the device
	{
		HostFile hf(dir, ".ref_i");
		uint64_t n = 0;
		if (!hf.f || !hf.rd(&n, 8) || n > hf.size / sizeof(dsb_ref_info)) { dsb_set_error("cannot read %s", hf.path.c_str()); return DSB_E_IO; }
		ix->ref_info.resize(n);
		if (!hf.rd(ix->ref_info.data(), n * sizeof(dsb_ref_info))) { dsb_set_error("short read %s", hf.path.c_str()); return DSB_E_IO; }
		std::vector<uint64_t> ri(n * 2);
		for (uint64_t i = 0; i < n; i++) { ix->ref_info[i].name[127] = 0; ri[2 * i] = ix->ref_info[i].seq_l; ri[2 * i + 1] = ix->ref_info[i].seq_offset; }
		if ((rc = upload(ix, ri.data(), n * 16, 0, &d)) != DSB_OK) return rc;
		ix->dev.ref_info = (const ulonglong2 *)d;
	}
	// ---- .ref_p (idx.c:1155-1159)
	if ((rc = load_array(ix, dir, ".ref_p", 8, true, &ix->n_rp, 0, &d)) != DSB_OK) return rc;
	ix->dev.ref_pos = (const uint64_t *)d;
	// ---- MAPQ tables: the reference's expressions evaluated in double and truncated to int (cly_mt.c:413-437),
	//      P_E = 0.15, L_REF = ref_bin.n * 4 (cly_mt.c:527).  Q_MEM is continued past 2000 entries so that the
	//      device never indexes outside the table (the reference leaves l_m >= 2000 unchecked).
	{
		const double P_E = 0.15; const uint64_t L_REF = ix->ref_bin_n * 4;
		const double REF_SIZE_PUNALTY = -10 * log(L_REF) / log(10);
		const double MATCH_SCORE = -10 * log(0.25 / (1 - P_E)) / log(10);
		const double MISMATCH_PUNALTY = -10 * log(0.75 / (P_E)) / log(10);
		std::vector<int> q(65536 + 400);
		for (int i = 0; i < 65536; i++) q[i] = REF_SIZE_PUNALTY + i * MATCH_SCORE + 0.5;
		int *lv = q.data() + 65536;
		for (int j = 0; j < 20; j++)
			for (int i = 0; i < 20; i++) {
				int v = (j - i) * MATCH_SCORE + i * MISMATCH_PUNALTY + 0.5;
				if (j < 5) v += 15;
				lv[i * 20 + j] = v > -8 ? v : -8;
			}
		if ((rc = upload(ix, q.data(), q.size() * 4, 0, &d)) != DSB_OK) return rc;
		ix->dev.q_mem = (const int *)d; ix->dev.q_lv = (const int *)d + 65536;
	}
	lap("unitigs, ref, tables");
	return DSB_OK;
}

static int g_blocking_sync = 0;       // dsb_set_sync_mode
static int open_device(int device)
{

	int n_dev = 0;
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { dsb_set_error("no CUDA device (there is no CPU fallback)"); return DSB_E_CUDA; }
	if (device < 0 || device >= n_dev) { dsb_set_error("device %d out of range (%d devices)", device, n_dev); return DSB_E_ARG; }
	DSB_CUDA(cudaSetDevice(device));
	// host threads that wait for a batch sleep instead of spinning (dsb_set_sync_mode); CUDA 12 applies the flags to the current
	// device whether or not its primary context exists already
	if (g_blocking_sync && cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync) != cudaSuccess) (void)cudaGetLastError();
	DSB_CUDA(cudaFree(0));
	return DSB_OK;
}

int dsb_blocking_sync() { return g_blocking_sync; }
extern "C" int dsb_set_sync_mode(int blocking)
{
	g_blocking_sync = blocking ? 1 : 0;
	return DSB_OK;
}

extern "C" int dsb_index_load(const char *dir, int device, dsb_index **out)
{
	if (!dir || !out) { dsb_set_error("dsb_index_load: null argument"); return DSB_E_ARG; }
	*out = nullptr;
	int rc = open_device(device);
	if (rc != DSB_OK) return rc;
	dsb_index *ix = nullptr;
	try {
		ix = new dsb_index();
		ix->device = device; ix->hbm_bytes = 0;
		memset(&ix->dev, 0, sizeof ix->dev);
		rc = load_impl(ix, dir, getenv("DSB_VERBOSE") != nullptr);
	} catch (const std::bad_alloc &) { dsb_set_error("out of host memory while loading the index"); rc = DSB_E_NOMEM; }
	catch (const std::exception &e) { dsb_set_error("index loader: %s", e.what()); rc = DSB_E_IO; }
	if (rc != DSB_OK) { dsb_index_free(ix); return rc; }
	*out = ix;
	return DSB_OK;
}

// The same index on another GPU of the box: every array is copied device to device (NVLink / NVSwitch peer copy; through the
// host where peer access is not available) instead of being read, uploaded and re-cut again (load_idx runs once, idx.c:1103).
extern "C" int dsb_index_clone(const dsb_index *src, int device, dsb_index **out)
{
	if (!src || !out) { dsb_set_error("dsb_index_clone: null argument"); return DSB_E_ARG; }
	*out = nullptr;
	int rc = open_device(device);
	if (rc != DSB_OK) return rc;
	dsb_index *ix = nullptr;
	try {
		ix = new dsb_index(*src);                                  // scalars, host copy of .ref_i
		ix->device = device; ix->allocs.clear();
	} catch (const std::bad_alloc &) { dsb_set_error("out of host memory"); return DSB_E_NOMEM; }
	{ int can = 0; if (cudaDeviceCanAccessPeer(&can, device, src->device) == cudaSuccess && can) { cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError(); else if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError(); } }
	rc = DSB_OK;
	for (const auto &a : src->allocs) {
		void *d = nullptr;
		if (cudaMalloc(&d, a.second) != cudaSuccess) { dsb_set_error("dsb_index_clone: cudaMalloc of %zu bytes on device %d failed", a.second, device); rc = DSB_E_CUDA; break; }
		ix->allocs.push_back({d, a.second});
		if (cudaMemcpyPeerAsync(d, device, a.first, src->device, a.second, 0) != cudaSuccess) { dsb_set_error("dsb_index_clone: peer copy failed: %s", cudaGetErrorString(cudaGetLastError())); rc = DSB_E_CUDA; break; }
	}
	if (rc == DSB_OK && cudaDeviceSynchronize() != cudaSuccess) { dsb_set_error("dsb_index_clone: %s", cudaGetErrorString(cudaGetLastError())); rc = DSB_E_CUDA; }
	if (rc != DSB_OK) { dsb_index_free(ix); return rc; }
	// the device pointers of the view, translated allocation by allocation
	auto tr = [&](const void *p) -> const void * {
		if (!p) return nullptr;
		for (size_t i = 0; i < src->allocs.size(); i++) {
			const char *b = (const char *)src->allocs[i].first;
			if ((const char *)p >= b && (const char *)p < b + src->allocs[i].second) return (const char *)ix->allocs[i].first + ((const char *)p - b);
		}
		return nullptr;
	};
	DevIndex &v = ix->dev; const DevIndex &s = src->dev;
	v.occ = (const uint8_t *)tr(s.occ); v.prefix = (const uint64_t *)tr(s.prefix); v.sa = (const uint2 *)tr(s.sa); v.uni = (const uint2 *)tr(s.uni);
	v.ref_pos = (const uint64_t *)tr(s.ref_pos); v.ref_bin = (const uint8_t *)tr(s.ref_bin); v.ref_info = (const ulonglong2 *)tr(s.ref_info);
	v.ek0 = (const uint8_t *)tr(s.ek0); v.ek1 = (const uint8_t *)tr(s.ek1); v.ek0_sum = (const uint32_t *)tr(s.ek0_sum);
	v.q_mem = (const int *)tr(s.q_mem); v.q_lv = (const int *)tr(s.q_lv);
	*out = ix;
	return DSB_OK;
}

extern "C" void dsb_index_free(dsb_index *ix)
{
	if (!ix) return;
	cudaSetDevice(ix->device);
	for (auto &a : ix->allocs) cudaFree(a.first);
	delete ix;
}
extern "C" uint64_t dsb_index_n_ref(const dsb_index *ix) { return ix ? ix->ref_info.size() : 0; }
extern "C" const dsb_ref_info *dsb_index_ref_info(const dsb_index *ix) { return ix ? ix->ref_info.data() : nullptr; }
extern "C" uint64_t dsb_index_hbm_bytes(const dsb_index *ix) { return ix ? ix->hbm_bytes : 0; }
extern "C" int dsb_index_l_ek(const dsb_index *ix) { return ix ? ix->dev.l_ek : 0; }
