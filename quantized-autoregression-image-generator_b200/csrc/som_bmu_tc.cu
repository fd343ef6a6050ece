// placeholder until the tcgen05 kernel lands
#include "som_common.cuh"
namespace som {
bool tc_supported(int64_t, int, int) { return false; }
size_t tc_workspace_bytes(int64_t, int, int) { return 0; }
int launch_bmu_tc(const float*, const Geom&, const float*, const float*, int, int64_t, int64_t*, float*,
                  void*, size_t, cudaStream_t) {
    return fail(SOM_E_UNSUPPORTED, "bmu: tensor-core variant not built");
}
}
