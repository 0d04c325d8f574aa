// dsb_device.cuh -- device-side helpers shared by the kernels of the deSAMBA hot path (sm_100a): warp-level primitives of the
// warp-per-read phases (chaining, scoring), hashes of the exist-k-mer tables, the glibc merge-sort emulation.  The index view
// (DevIndex) is in dsb_index_view.h, the seeding engine in dsb_seedcore.h.
//
// Arithmetic mirrors the DECLARED C types of the reference (mixed signed/unsigned MAX/MIN/ABS macros, utils.h:61-64)
// because the results depend on the usual arithmetic conversions.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "dsb_seedcore.h"

#define DSB_MAX(a,b) (((a) > (b))?(a):(b))
#define DSB_MIN(a,b) (((a) < (b))?(a):(b))
#define DSB_ABS(a) (((a) > 0)?(a): (- (a)))
#define DSB_ABS_U(a,b) (((a) > (b))?((a) - (b)): ((b) - (a)))

#define DSB_FORWARD 1
#define DSB_REVERSE 0
#define DSB_GUARD 64            // zero bytes before the forward strand and after the reverse strand (out-of-buffer policy P3)
#define DSB_FULL 0xffffffffu

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

// get_ref (cly.c:435-466), forward: unpack `length` bases of the packed reference from position `off` into out[0 .. length)
// (one byte per base).  Cooperative: lane k writes out[k], out[k+32], ...; bytes behind `length` are never written: the window
// keeps what it held there (zero-initialised, then earlier loads of the same call: oracle policy P1).
// (A variant that unpacked 16 bases per lane from two aligned words executed 5x fewer instructions and made k_score 10 % SLOWER,
// twice: the byte loop's 64 independent loads per lane are all in flight at once; profiles/README.md.)
__device__ __forceinline__ void get_ref_coop(const DevIndex &ix, uint8_t *out, int64_t off, int32_t length)
{
	if (off < 0) off = 0;
	if (length < 0) length = 0;
	const uint64_t o = (uint64_t)off;
	for (uint32_t k = lane_id(); k < (uint32_t)length; k += 32) out[k] = (uint8_t)ref_base_at(ix, o + k);
	__syncwarp();
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
