// K1 (tensor-core variant): dispatch between the tcgen05 BMU kernels.  sm_100a only.
//
// Replaces patchify + torch.cdist + torch.argmin of Codebook.get_patches_bmu
// (/root/reference/models/Codebook.py:77-99).  The reduced distance
//     rd[p][j] = ||c_j||^2 - 2 x_p . c_j          (d^2 = rd + ||x_p||^2, row-constant dropped)
// is computed as an error-compensated 3xTF32 GEMM (hi/lo split of both operands, three products, norms through
// one extra k-step) with the argmin fused into the TMEM epilogue.  Static rule on the shape, no autotuner:
//   D <= 16  -> som_bmu_tc_s.cu  (config S: R = 4 resident patch tiles, chunk-tracking epilogue + exact refine)
//   D  > 16  -> som_bmu_tc_l.cu  (resident-A for D <= 64, TMA-streamed A for larger D, split-K when there are
//                                 fewer patch tiles than SMs; CTA pairs for large problems)
#include "som_common.cuh"
#include "som_tc_ptx.cuh"

namespace som {

// som_bmu_tc_s.cu
bool tc_s_applicable(int D);
size_t tc_s_workspace_bytes(int64_t n_patches, int D, int K);
int launch_bmu_tc_s(const float* x, const Geom& g, const float* W, const float* cn, int K, int64_t unit_offset,
                    int64_t* out_idx, float* out_rd, void* ws, size_t ws_bytes, cudaStream_t st);
// som_bmu_tc_l.cu
size_t tc_l_workspace_bytes(int64_t n_patches, int D, int K);
int launch_bmu_tc_l(const float* x, const Geom& g, const float* W, const float* cn, int K, int64_t unit_offset,
                    int64_t* out_idx, float* out_rd, void* ws, size_t ws_bytes, cudaStream_t st);

bool tc_supported(int64_t n_patches, int D, int K) {
    if (n_patches <= 0 || D <= 0 || K <= 0) return false;
    if (n_patches / tc::TM >= (1ll << 30)) return false;
    if ((int64_t)D > (1 << 20)) return false;
    if ((int64_t)K + tc::TN >= (1ll << 31)) return false;
    static int cc_major = -1;
    if (cc_major < 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
            return false;
        cc_major = v;
    }
    return cc_major == 10 && tc::get_encode_fn() != nullptr;
}

size_t tc_workspace_bytes(int64_t n_patches, int D, int K) {
    if (!tc_supported(n_patches, D, K)) return 0;
    if (tc_s_applicable(D)) return tc_s_workspace_bytes(n_patches, D, K);
    return tc_l_workspace_bytes(n_patches, D, K);
}

int launch_bmu_tc(const float* x, const Geom& g, const float* W, const float* cn, int K,
                  int64_t unit_offset, int64_t* out_idx, float* out_rd, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
    if (g.n_patches == 0) return SOM_OK;
    if (tc_s_applicable(g.D)) return launch_bmu_tc_s(x, g, W, cn, K, unit_offset, out_idx, out_rd, ws, ws_bytes, st);
    return launch_bmu_tc_l(x, g, W, cn, K, unit_offset, out_idx, out_rd, ws, ws_bytes, st);
}

}  // namespace som
