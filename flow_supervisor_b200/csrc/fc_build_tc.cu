// tcgen05 / TMEM build path (placeholder until the tensor-core kernel lands).
#include "fc_common.cuh"

namespace fc {

size_t tc_build_workspace_bytes(int, int, int, int, int, int) { return 0; }

int tc_build(const float*, const float*, void*, const Pyramid&, int, int, int, int, int, void*, size_t,
             cudaStream_t) {
    set_error("fc_build: tensor-core math modes are not available in this build");
    return FC_EINVAL;
}

}  // namespace fc
