/* stubs in a separate TU without prototypes: never called with dev == NULL */
#include <stdlib.h>
#define STUB(name) int name() { abort(); return 0; }
STUB(asam_reserve) STUB(asam_upload_ipool) STUB(asam_upload_desc) STUB(asam_upload_node2q) STUB(asam_upload_q2node)
STUB(asam_upload_fslot) STUB(asam_set_full_tasks) STUB(asam_set_leaf_tasks) STUB(asam_set_bs_leaf_count)
STUB(asam_set_shard_schedule) STUB(asam_btasks_prepend) STUB(asam_hessian_clear_range) STUB(asam_device_info)
const char *asam_last_error(void) { return ""; }
