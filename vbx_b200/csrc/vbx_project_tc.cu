// tcgen05 (5th-gen tensor core) projection rho = X . V with 3xTF32 split precision.
// Placeholder until the UMMA/TMA kernel lands: reports "unavailable" so the C ABI uses the FFMA tiles.
#include "vbx_internal.cuh"

namespace vbx {
int launch_project_tcgen05(const Plan &, const float *, int, const float *, float *, cudaStream_t, std::string *err) {
    if (err) *err = "not built";
    return -1;
}
}  // namespace vbx
