/*
 * batch_order.h -- the one cross-batch dependency of the driver, as plain C so that it can be tested without a GPU
 * (tests/test_batch_order.py).
 *
 * Classify_buff_pool.max_read_l (cly.c:2958) is the reference's running maximum of the read lengths that reached the class
 * filter; with -t 1 it runs in input order.  It is read only as `max_read_l < 510` (cly.c:2960), and a read >= 510 bp
 * raises it above the threshold before its own test.  So a batch needs the value of its predecessors only when it holds a
 * read < 510 bp, and only the predecessors that hold a read >= 510 bp can change the answer.
 *
 * The driver finishes batches out of order (several contexts per GPU, several GPUs).  State kept under its mutex:
 *   done_upto   every batch with seq_no < done_upto is finished (contiguous prefix)
 *   prefix_max  max_read_l after those batches
 * and per batch in flight (ring of slots): state, has_long, has_short, max_out (max_read_l after the batch, given its max_in).
 * A value handed to batch `my` is built ONLY from batches with seq_no < my; never from a global that later batches update.
 */
#ifndef DSB_BATCH_ORDER_H
#define DSB_BATCH_ORDER_H
#include <stdint.h>

enum { BO_FREE = 0, BO_READY, BO_BUSY, BO_DONE };
typedef struct { int state, has_long, has_short; int32_t max_out; uint64_t seq_no; } bo_slot;

/* May batch `my` start now?  1: yes, *max_in = the value to hand to dsb_classify_batch; 0: an earlier batch that holds a read
 * >= 510 bp is still running and this batch holds a read < 510 bp -- wait for a completion and ask again. */
static inline int bo_may_start(const bo_slot *slot, int n_slots, uint64_t my, uint64_t done_upto, int32_t prefix_max, int32_t *max_in)
{
	int32_t m = prefix_max;
	const bo_slot *me = &slot[my % (uint64_t)n_slots];
	if (m >= 510 || !me->has_short) { *max_in = m; return 1; }       /* the threshold is passed for good / no read of this batch asks */
	for (uint64_t k = done_upto; k < my; k++) {
		const bo_slot *e = &slot[k % (uint64_t)n_slots];
		if (e->state == BO_DONE) { if (e->max_out > m) m = e->max_out; }
		else if (e->has_long) return 0;                              /* could still pass the threshold */
	}
	*max_in = m;                                                     /* (unfinished predecessors hold reads < 510 bp only: they cannot change the answer) */
	return 1;
}

/* batch `seq` has finished with max_out: advance the contiguous prefix */
static inline void bo_finished(const bo_slot *slot, int n_slots, uint64_t n_claimed, uint64_t *done_upto, int32_t *prefix_max)
{
	while (*done_upto < n_claimed) {
		const bo_slot *e = &slot[*done_upto % (uint64_t)n_slots];
		if (e->state != BO_DONE || e->seq_no != *done_upto) break;
		if (e->max_out > *prefix_max) *prefix_max = e->max_out;
		(*done_upto)++;
	}
}
#endif
