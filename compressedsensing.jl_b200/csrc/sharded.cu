// Column-sharded single-dictionary OMP (one process per GPU, NCCL over NVLink) -- placeholder
// entry points; see include/csb200.h.  Filled in after the single-GPU path is parity-green.
#include "../../include/csb200.h"
extern "C" {
int csb200_comm_unique_id(void*) { return CSB200_ERR_UNSUPPORTED; }
int csb200_comm_create(const void*, int, int, int, csb200_comm**) { return CSB200_ERR_UNSUPPORTED; }
int csb200_comm_destroy(csb200_comm*) { return CSB200_OK; }
int csb200_omp_sharded(csb200_dict*, csb200_comm*, const void*, int64_t, double, int64_t*, double*, int64_t*, double*,
                       int64_t*, double*) { return CSB200_ERR_UNSUPPORTED; }
}
