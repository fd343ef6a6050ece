// K1 (FFMA variant): Best-Matching-Unit search, fp32 FFMA register-tiled, fused argmin.
//
// Replaces patchify + torch.cdist + torch.argmin of Codebook.get_patches_bmu
// (/root/reference/models/Codebook.py:77-99, models/layers.py:8-34).  For each patch p
//     score_j = x_p . W_j - 0.5 ||W_j||^2          (d^2 = ||x_p||^2 - 2 score_j, ||x_p||^2 row-constant)
// and the kernel keeps the running (max score, lowest index) per patch, so the (N*Seq) x K distance
// matrix never exists.  Dropping the row-constant term makes near-tie comparisons more accurate
// than the reference's own expansion.
//
// Tiling: CTA = 128 patches x 128 units, 256 threads, 8x8 register tile per thread (split 4+4 in
// both directions so every shared-memory read is a conflict-free LDS.128), K-slab of 16 features
// double-buffered in shared memory, next slab prefetched global->registers during the FFMAs.
// patchify is address arithmetic in the loader (VEC-wide loads along a patch row).  Units are
// walked in ascending order inside a thread, so "strictly greater" keeps the lowest index; the
// 16 threads sharing a patch row merge with (score desc, index asc).
// grid.y splits the unit range when there are too few patch tiles to fill 148 SMs; the per-split
// candidates are merged by split_merge_kernel with the same rule.
// Bound: FP32 FFMA pipe -- 2*K*D flop per patch against 4*D+8 bytes.
#include "som_common.cuh"

namespace som {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;
constexpr int FOFF_TAB = 2048;      // feature-offset table entries kept in shared memory

template <int V> struct VT;
template <> struct VT<1> { using T = float; };
template <> struct VT<2> { using T = float2; };
template <> struct VT<4> { using T = float4; };

struct BmuArgs {
    const float* x;
    Geom g;
    const float* W;
    const float* cn;
    int K;
    int n_unit_tiles;        // ceil(K / BN)
    int tiles_per_split;
    int64_t unit_offset;
    int64_t* out_idx;
    float* out_rd;
    float* cand_score;       // [splits][n]  (splits > 1)
    int* cand_idx;
};

template <int XV, int WV>
__global__ void __launch_bounds__(NT, 2) bmu_ffma_kernel(BmuArgs a) {
    __shared__ __align__(16) float xs[2][BK][BM];
    __shared__ __align__(16) float cs[2][BK][BN];
    __shared__ int64_t pb[BM];
    __shared__ int foff[FOFF_TAB];

    const Geom& g = a.g;
    const int D = g.D;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int64_t n = g.n_patches;
    const int t0 = blockIdx.y * a.tiles_per_split;
    int t1 = t0 + a.tiles_per_split;
    if (t1 > a.n_unit_tiles) t1 = a.n_unit_tiles;
    const int nKT = (D + BK - 1) / BK;
    const bool x_resident = (nKT == 1);
    const bool use_tab = (D <= FOFF_TAB);

    if (tid < BM) pb[tid] = (m0 + tid < n) ? patch_base(g, m0 + tid) : (int64_t)-1;
    if (use_tab)
        for (int d = tid; d < D; d += NT) foff[d] = feat_off(g, d);
    __syncthreads();

    // loader roles: row r = tid % 128 (a patch for x, a unit for W), 8 consecutive features
    const int lr = tid & 127;
    const int lk = (tid >> 7) * 8;
    const int64_t my_pb = pb[lr];

    float xr[8], wr[8];

    auto load_x = [&](int kt) {
#pragma unroll
        for (int q = 0; q < 8 / XV; ++q) {
            int d = kt * BK + lk + q * XV;
            float tmp[XV];
#pragma unroll
            for (int e = 0; e < XV; ++e) tmp[e] = 0.f;
            if (my_pb >= 0 && d < D) {
                int off = use_tab ? foff[d] : feat_off(g, d);
                typename VT<XV>::T v = __ldg(reinterpret_cast<const typename VT<XV>::T*>(a.x + my_pb + off));
                const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
                for (int e = 0; e < XV; ++e) tmp[e] = f[e];
            }
#pragma unroll
            for (int e = 0; e < XV; ++e) xr[q * XV + e] = tmp[e];
        }
    };
    auto load_w = [&](int ut, int kt) {
        int unit = ut * BN + lr;
#pragma unroll
        for (int q = 0; q < 8 / WV; ++q) {
            int d = kt * BK + lk + q * WV;
            float tmp[WV];
#pragma unroll
            for (int e = 0; e < WV; ++e) tmp[e] = 0.f;
            if (unit < a.K && d < D) {
                typename VT<WV>::T v =
                    __ldg(reinterpret_cast<const typename VT<WV>::T*>(a.W + (int64_t)unit * D + d));
                const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
                for (int e = 0; e < WV; ++e) tmp[e] = f[e];
            }
#pragma unroll
            for (int e = 0; e < WV; ++e) wr[q * WV + e] = tmp[e];
        }
    };
    auto store_x = [&](int buf) {
#pragma unroll
        for (int e = 0; e < 8; ++e) xs[buf][lk + e][lr] = xr[e];
    };
    auto store_w = [&](int buf) {
#pragma unroll
        for (int e = 0; e < 8; ++e) cs[buf][lk + e][lr] = wr[e];
    };

    float best[8];
    int bidx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { best[i] = -INFINITY; bidx[i] = 0; }

    const int total = (t1 - t0) * nKT;
    if (total > 0) {
        load_x(0);
        load_w(t0, 0);
        store_x(0);
        store_w(0);
    }
    __syncthreads();

    float acc[8][8];
    int ut = t0, kt = 0;
    for (int it = 0; it < total; ++it) {
        const bool has_next = (it + 1 < total);
        int nut = ut, nkt = kt + 1;
        if (nkt == nKT) { nkt = 0; nut = ut + 1; }
        if (has_next) {
            if (!x_resident) load_x(nkt);
            load_w(nut, nkt);
        }
        if (kt == 0) {
            // acc starts at -0.5 ||c||^2 (padding units: -inf, never selected)
            float cv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int unit = ut * BN + ((j < 4) ? (tx * 4 + j) : (64 + tx * 4 + (j - 4)));
                cv[j] = (unit < a.K) ? -0.5f * __ldg(a.cn + unit) : -INFINITY;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = cv[j];
        }
        const int cb = it & 1;
        const int xb = x_resident ? 0 : cb;
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&xs[xb][k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&xs[xb][k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&cs[cb][k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&cs[cb][k][64 + tx * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt == nKT - 1) {
            const int ubase = ut * BN + tx * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float mx = acc[i][0];
#pragma unroll
                for (int j = 1; j < 8; ++j) mx = fmaxf(mx, acc[i][j]);
                if (mx > best[i]) {
                    best[i] = mx;
                    int q = 7;
#pragma unroll
                    for (int j = 6; j >= 0; --j) q = (acc[i][j] == mx) ? j : q;
                    bidx[i] = ubase + ((q < 4) ? q : (60 + q));
                }
            }
        }
        if (has_next) {
            if (!x_resident) store_x((it + 1) & 1);
            store_w((it + 1) & 1);
        }
        __syncthreads();
        ut = nut; kt = nkt;
    }

    // merge the 16 threads (tx) that share each patch row: score desc, index asc
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float s = best[i];
        int bi = bidx[i];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            float os = __shfl_xor_sync(0xffffffffu, s, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (os > s || (os == s && oi < bi)) { s = os; bi = oi; }
        }
        best[i] = s; bidx[i] = bi;
    }
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int64_t p = m0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4)));
            if (p < n) {
                if (gridDim.y == 1) {
                    a.out_idx[p] = (int64_t)bidx[i] + a.unit_offset;
                    if (a.out_rd) a.out_rd[p] = -2.0f * best[i];
                } else {
                    a.cand_score[(int64_t)blockIdx.y * n + p] = best[i];
                    a.cand_idx[(int64_t)blockIdx.y * n + p] = bidx[i];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) split_merge_kernel(const float* __restrict__ score,
                                                          const int* __restrict__ idx, int splits,
                                                          int64_t n, int64_t unit_offset,
                                                          int64_t* __restrict__ out_idx,
                                                          float* __restrict__ out_rd) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    float s = score[p];
    int bi = idx[p];
    for (int r = 1; r < splits; ++r) {
        float os = score[(int64_t)r * n + p];
        int oi = idx[(int64_t)r * n + p];
        if (os > s || (os == s && oi < bi)) { s = os; bi = oi; }
    }
    out_idx[p] = (int64_t)bi + unit_offset;
    if (out_rd) out_rd[p] = -2.0f * s;
}

// how many unit splits: enough CTAs for two waves, but every split keeps >= 2 unit tiles
int ffma_splits(int64_t n_patches, int K) {
    int64_t patch_tiles = ceil_div64(n_patches, BM);
    int unit_tiles = (int)ceil_div64(K, BN);
    int64_t want = ceil_div64(2LL * sm_count(), patch_tiles > 0 ? patch_tiles : 1);
    int splits = (int)(want < 1 ? 1 : want);
    int max_splits = unit_tiles / 2 > 0 ? unit_tiles / 2 : 1;
    if (splits > max_splits) splits = max_splits;
    if (splits > 64) splits = 64;
    return splits;
}

size_t ffma_workspace_bytes(int64_t n_patches, int K) {
    int splits = ffma_splits(n_patches, K);
    if (splits <= 1) return 0;
    return align_up((size_t)splits * n_patches * 4, 256) * 2;
}

int launch_bmu_ffma(const float* x, const Geom& g, const float* W, const float* cn, int K,
                    int64_t unit_offset, int64_t* out_idx, float* out_rd, void* ws, size_t ws_bytes,
                    cudaStream_t st) {
    const int64_t n = g.n_patches;
    if (n == 0) return SOM_OK;
    BmuArgs a;
    a.x = x; a.g = g; a.W = W; a.cn = cn; a.K = K;
    a.n_unit_tiles = (int)ceil_div64(K, BN);
    int splits = ffma_splits(n, K);
    a.tiles_per_split = (int)ceil_div64(a.n_unit_tiles, splits);
    splits = (int)ceil_div64(a.n_unit_tiles, a.tiles_per_split);
    a.unit_offset = unit_offset; a.out_idx = out_idx; a.out_rd = out_rd;
    a.cand_score = nullptr; a.cand_idx = nullptr;
    if (splits > 1) {
        size_t half = align_up((size_t)splits * n * 4, 256);
        SOM_REQUIRE(ws != nullptr && ws_bytes >= 2 * half, SOM_E_WORKSPACE,
                    "bmu(ffma): workspace %zu < required %zu", ws_bytes, 2 * half);
        a.cand_score = (float*)ws;
        a.cand_idx = (int*)((char*)ws + half);
    }
    int64_t patch_tiles = ceil_div64(n, BM);
    SOM_REQUIRE(patch_tiles < (1LL << 31), SOM_E_SHAPE, "bmu(ffma): too many patches");
    dim3 grid((unsigned)patch_tiles, (unsigned)splits);
    int xv = g.vec;
    int wv = (g.D % 4 == 0 && ((uintptr_t)W & 15) == 0) ? 4 : 1;
    if (wv == 1) xv = 1;
    if (xv == 4) bmu_ffma_kernel<4, 4><<<grid, NT, 0, st>>>(a);
    else if (xv == 2) bmu_ffma_kernel<2, 4><<<grid, NT, 0, st>>>(a);
    else if (wv == 4) bmu_ffma_kernel<1, 4><<<grid, NT, 0, st>>>(a);
    else bmu_ffma_kernel<1, 1><<<grid, NT, 0, st>>>(a);
    int rc = check_launch("bmu_ffma_kernel");
    if (rc) return rc;
    if (splits > 1) {
        split_merge_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(a.cand_score, a.cand_idx, splits,
                                                                        n, unit_offset, out_idx, out_rd);
        return check_launch("split_merge_kernel");
    }
    return SOM_OK;
}

}  // namespace som
