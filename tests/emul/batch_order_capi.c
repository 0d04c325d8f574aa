/* TEST INFRASTRUCTURE: the driver's cross-batch ordering logic (desamba_b200/csrc/batch_order.h) exported for ctypes */
#include "../../desamba_b200/csrc/batch_order.h"
int capi_bo_may_start(const bo_slot *slot, int n_slots, uint64_t my, uint64_t done_upto, int32_t prefix_max, int32_t *max_in) { return bo_may_start(slot, n_slots, my, done_upto, prefix_max, max_in); }
void capi_bo_finished(const bo_slot *slot, int n_slots, uint64_t n_claimed, uint64_t *done_upto, int32_t *prefix_max) { bo_finished(slot, n_slots, n_claimed, done_upto, prefix_max); }
