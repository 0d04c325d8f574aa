// dsb_internal.h -- host-side objects behind the C ABI (include/desamba_b200.h).
#pragma once
#include "../../include/desamba_b200.h"
#include "dsb_device.cuh"
#include <cuda_runtime.h>
#include <vector>
#include <string>
#include <utility>

void dsb_set_error(const char *fmt, ...);
int dsb_blocking_sync();                 // dsb_set_sync_mode (dsb_index.cu)
#define DSB_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	dsb_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return DSB_E_CUDA; } } while (0)

struct dsb_index {
	int device;
	DevIndex dev;                       // device pointers + scalars, passed to kernels by value
	std::vector<std::pair<void *, size_t>> allocs;   // every cudaMalloc of this index (pointer, bytes)
	uint64_t hbm_bytes;
	std::vector<dsb_ref_info> ref_info; // host copy of .ref_i for the writers
	uint64_t ref_bin_n;                 // bytes of .ref_b (L_REF = 4 * this, cly_mt.c:527)
	uint64_t bwt_len_blocks;            // reference occ blocks (256 symbols each)
	uint64_t n_uni, n_rp, sa_size, ek_size;
};

struct DevBuf {                         // grow-only device buffer
	void *p; size_t cap;
	DevBuf() : p(nullptr), cap(0) {}
};

#define DSB_STAGE_BYTES (8u << 20)     // one piece of the staged (pageable) upload
#define DSB_STAGE_THREADS 3            // host threads that stage a batch, two pieces of the ring each
#define DSB_N_KERNELS 11               // timed kernel groups of one dsb_batch_run
#define DSB_N_EV (DSB_N_KERNELS + 4)   // kernel boundaries + 2 user marks + start of the upload

// One uploaded batch: its reads in HBM, the per-read layout tables (pinned host memory, copied to the device when the batch is
// run) and its sizes.  A context has two, so that dsb_batch_upload of the NEXT batch may run -- on a stream of its own -- while
// the kernels of the current one are still busy (include/desamba_b200.h).
struct BatchIn {
	DevBuf seqs;
	void *h_pin = nullptr; size_t h_pin_cap = 0;
	uint32_t n_reads = 0, n_tiles = 0, max_len = 0; uint64_t n_bases = 0, bits_words = 0, seed_slots = 0, bin_bytes = 0;
	std::vector<uint64_t> h_off;        // host copies of the per-read offset tables
	std::vector<uint64_t> h_bits_off;
	std::vector<uint32_t> h_seed_off;
	cudaEvent_t ev_up = nullptr;        // the reads have landed in `seqs`
};

struct dsb_ctx {
	dsb_index *ix;
	dsb_opts opts;
	cudaStream_t stream;
	cudaStream_t copy_stream;           // uploads of the reads
	BatchIn in[2]; int cur = 0, pend = 0;
	bool has_pending = false;           // in[pend] is uploaded and waits for dsb_batch_run
	bool run_pending = false;           // in[cur] has been run and not yet downloaded
	int n_sm, n_warps;                  // resident classify warps = n_sm * warps_per_sm
	int seed_blocks;                    // grid of k_seed (persistent, SEED_WARPS_PER_SM warps per SM)
	int heavy_blocks;                   // CTAs of k_score_heavy
	// batch inputs (device)
	DevBuf read_off, bin_off, bits_off, seed_off, tiles, bin, bits, seeds[2], n_seeds[2], total_score[2];
	// classify scratch + outputs
	DevBuf scratch, rr, hits, counters, prof, work, anc_pool, chain_pool, lists[4], ctl, order, hdr7;
	// seeding: packed forward strands, task lists (ping-pong), per-task records, staging chunks, per-lane memory of k_seed
	DevBuf pk, tasks[2], recs, chunks, lane_mem, vis2, vis1_full, vis_gen, task_first[2], task_cnt[2];
	uint32_t task_cap, n_chunks;
	uint32_t grow[5];                   // pool growth (x 2^k) after overflows: tasks, chunks, anchors, chains, hits
	int32_t max_read_l_in;              // of the last dsb_batch_run (re-runs after a pool overflow, dsb_batch_download)
	int retries;                        // re-runs of the last batch
	double host_s[5] = {0, 0, 0, 0, 0};  // dsb_ctx_host_seconds
	uint64_t scratch_stride;
	uint64_t hits_cap;
	// pinned ring for reads that arrive in pageable memory
	void *h_stage; cudaEvent_t ev_stage[2 * DSB_STAGE_THREADS];
	// batch state
	uint32_t m_bin_read;                // capacity of the reference's bin_read buffer after the batches seen so far (policy P3)
	uint32_t n_reads, n_tiles; uint64_t n_bases, bits_words, seed_slots, bin_bytes; uint32_t max_len;   // of the batch last run (in[cur])
	std::vector<uint32_t> h_len_first;  // counting sort of the reads by length (work order)
	cudaEvent_t ev[DSB_N_EV];
	int launches;
	bool ran;
};

// device-side batch counters (one block of u64 in ctx->counters)
enum {
	DSB_CNT_HITS_CURSOR = 0,   // next free slot in hits
	DSB_CNT_WORK,              // classify work counter
	DSB_CNT_FIRST_LONG,        // min read index with entered_final && read_len >= 510 (init: ~0)
	DSB_CNT_MAX_READ_L,        // max read_len over entered_final reads
	DSB_CNT_N_BIT0,            // algorithmic counters: get_exist_kmer calls with k-mer != 0
	DSB_CNT_N_BIT1,            //   table-1 probes (table-0 hit)
	DSB_CNT_N_PREFIX,          //   prefix-table lookups (bwt_MEM_search calls)
	DSB_CNT_N_OCC,             //   occ calls
	DSB_CNT_N_LOCATE,          //   get_uni calls
	DSB_CNT_N_GETREF,          //   get_ref calls of the seeding phase (map_seed / get_new_ed flanks)
	DSB_CNT_N_GETREF_BYTES,    //   packed reference bytes those calls cover
	DSB_CNT_N_ERRORS,          // reads that hit a capacity
	DSB_CNT_N_GETREF_SCORE,    //   get_ref calls of the scoring phase (sdp windows)
	DSB_CNT_N_GETREF_BYTES_SCORE,
	DSB_CNT_COUNT = 16
};
