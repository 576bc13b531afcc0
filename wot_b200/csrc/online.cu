// Online-kernel unbalanced Sinkhorn (placeholder until the tile kernels land).
#include "solver_state.cuh"

namespace wotb {

int sinkhorn_online(wotb_ctx *, const double *, int64_t, const double *, int64_t, int, double, const double *,
                    const wotb_params *, double *, double *, double *, wotb_info *) {
    set_error("online kernel not built in this revision");
    return WOTB_ERR_INVALID;
}

int coupling_online(wotb_ctx *, const double *, int64_t, const double *, int64_t, int, double, const double *,
                    const double *, double, double, void *, int64_t, int, double *, cudaStream_t) {
    set_error("online kernel not built in this revision");
    return WOTB_ERR_INVALID;
}

}  // namespace wotb
