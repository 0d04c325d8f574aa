// dsb_index_view.h -- the index as the kernels see it (pointers into HBM + scalars), and the constants of the path.
// Plain C++: included by the CUDA sources and by the host-side emulation of the seeding engine in tests/emul.
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#else
#include <vector_types.h>
#endif

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
	const int      *q_mem;      // Q_MEM[65536], cly_mt.c:413-437 (host-computed, double -> int exactly as the reference)
	const int      *q_lv;       // Q_LV[20][20] row-major [d][l]
};

#define L_PRE_IDX 13
#define PRE_IDX_MASK 0x3FFFFFFu
#define SA_MASK 0x7
#define SA_OFF 3
#define MIN_UNI_L 35
#define LV_L 12
#define NO_SA 0xFFFFFFFFFFFFFFFFull
#define SP_SET_CAP 500

struct MemRst { int match_len; int sa_sp_l; uint64_t sp, sa_sp; int read_offset; int pad; };   // MEM_rst (cly.c:619-627)
