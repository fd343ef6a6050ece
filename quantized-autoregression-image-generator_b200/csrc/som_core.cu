// libsomcb core: error plumbing, geometry, and the small HBM-bound kernels of the SOM path
// (codebook norms, candidate merge, hit histogram, quantise/gather, Adam, row compaction).
#include "som_common.cuh"

#include <atomic>
#include <mutex>
#include <string.h>

namespace som {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return SOM_OK;
}

unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
    static int cached[64];
    static std::once_flag once;
    std::call_once(once, [] { memset(cached, 0, sizeof(cached)); });
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int make_geom(Geom* g, const void* x, int64_t n_img, int C, int H, int W, int pH, int pW) {
    SOM_REQUIRE(n_img >= 0 && C > 0 && H > 0 && W > 0 && pH > 0 && pW > 0, SOM_E_BADARG,
                "geometry: non-positive dimension (n_img=%lld C=%d H=%d W=%d pH=%d pW=%d)",
                (long long)n_img, C, H, W, pH, pW);
    SOM_REQUIRE(H % pH == 0 && W % pW == 0, SOM_E_SHAPE,
                "geometry: image %dx%d is not divisible by patch %dx%d", H, W, pH, pW);
    g->n_img = n_img; g->C = C; g->H = H; g->W = W; g->pH = pH; g->pW = pW;
    g->gH = H / pH; g->gW = W / pW; g->seq = g->gH * g->gW;
    g->D = C * pH * pW;
    g->n_patches = n_img * g->seq;
    g->img_stride = (int64_t)C * H * W;
    uintptr_t a = (uintptr_t)x;
    SOM_REQUIRE((a & 3) == 0, SOM_E_BADARG, "geometry: fp32 buffer is not 4-byte aligned");
    g->vec = 1;
    if (pW % 4 == 0 && (a & 15) == 0) g->vec = 4;
    else if (pW % 2 == 0 && (a & 7) == 0) g->vec = 2;
    return SOM_OK;
}

// ---------------------------------------------------------------------------------------------
// K0: ||W_j||^2, one warp per unit, fixed summation order (lane-strided, then butterfly)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) norm2_kernel(const float* __restrict__ W, int K, int D,
                                                    float* __restrict__ out) {
    int warp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    int lane = threadIdx.x & 31;
    if (warp >= K) return;
    const float* row = W + (int64_t)warp * D;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) { float v = row[d]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) out[warp] = s;
}

// ---------------------------------------------------------------------------------------------
// K1b: candidate merge
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) merge_kernel(const float* __restrict__ rd,
                                                    const int64_t* __restrict__ idx, int R,
                                                    int64_t n, int64_t* __restrict__ out_idx,
                                                    float* __restrict__ out_rd) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    float best = rd[p];
    int64_t bi = idx[p];
    for (int r = 1; r < R; ++r) {
        float v = rd[(int64_t)r * n + p];
        int64_t i = idx[(int64_t)r * n + p];
        if (v < best || (v == best && i < bi)) { best = v; bi = i; }
    }
    out_idx[p] = bi;
    if (out_rd) out_rd[p] = best;
}

// ---------------------------------------------------------------------------------------------
// K5: hit histogram.  Small K: shared-memory private counts per CTA, one global merge per CTA.
// Large K: direct 64-bit global atomics (integer, so the result is order-independent).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) hist_smem_kernel(const int64_t* __restrict__ idx, int64_t n,
                                                        int K, unsigned long long* __restrict__ counts) {
    extern __shared__ unsigned int sh[];
    for (int i = threadIdx.x; i < K; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += stride) {
        int64_t j = idx[p];
        if (j >= 0 && j < K) atomicAdd(&sh[j], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        unsigned int c = sh[i];
        if (c) atomicAdd(&counts[i], (unsigned long long)c);
    }
}

__global__ void __launch_bounds__(256) hist_global_kernel(const int64_t* __restrict__ idx, int64_t n,
                                                          int K, unsigned long long* __restrict__ counts) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += stride) {
        int64_t j = idx[p];
        if (j >= 0 && j < K) atomicAdd(&counts[j], 1ull);
    }
}

// ---------------------------------------------------------------------------------------------
// token assembly for the Transformer trainer (train_quantized_transformer.py:411-455): one pass writes
// hr_input and hr_target from the two BMU index tensors -- replaces add / cat / repeat / cat
// one thread per output element; a row of hr_input followed by the row of hr_target
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) assemble_tokens_kernel(const int64_t* __restrict__ lr_idx,
                                                              const int64_t* __restrict__ hr_idx, int64_t n,
                                                              int lr_seq, int hr_seq, int64_t lr_K, int64_t hr_K,
                                                              int base_model, int64_t* __restrict__ hr_input,
                                                              int64_t* __restrict__ hr_target) {
    const int in_w = base_model ? lr_seq + hr_seq : 1 + hr_seq;
    const int row_w = in_w + hr_seq + 1;
    const int64_t total = n * (int64_t)row_w;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t i = t / row_w;
        const int c = (int)(t - i * row_w);
        if (c < in_w) {
            int64_t v;
            if (base_model) v = (c < lr_seq) ? lr_idx[i * lr_seq + c] : hr_idx[i * hr_seq + (c - lr_seq)] + lr_K;
            else v = (c == 0) ? hr_K : hr_idx[i * hr_seq + (c - 1)];
            hr_input[i * in_w + c] = v;
        } else {
            const int k = c - in_w;
            hr_target[i * (int64_t)(hr_seq + 1) + k] = (k < hr_seq) ? hr_idx[i * hr_seq + k] : hr_K;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// quantise: out[offset(p, d)] = table[idx[p]][d]   (gather + fused unpatchify)
// one thread per VEC-wide run; consecutive threads walk d fastest inside a patch
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) quantize_kernel(const int64_t* __restrict__ idx,
                                                       const float* __restrict__ table, int K,
                                                       Geom g, float* __restrict__ out) {
    const int dv = g.D / VEC;
    int64_t total = g.n_patches * dv;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += stride) {
        int64_t p = t / dv;
        int d = (int)(t - p * dv) * VEC;
        int64_t u = idx[p];
        if (u < 0) u = 0;
        if (u >= K) u = K - 1;
        const float* src = table + u * (int64_t)g.D + d;
        float* dst = out + patch_base(g, p) + feat_off(g, d);
        if (VEC == 4) *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(src);
        else if (VEC == 2) *reinterpret_cast<float2*>(dst) = *reinterpret_cast<const float2*>(src);
        else *dst = *src;
    }
}

// ---------------------------------------------------------------------------------------------
// K4: Adam, torch.optim.Adam single-tensor rule (lerp form of the first moment)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ W, float* __restrict__ m,
                                                   float* __restrict__ v, const float* __restrict__ g,
                                                   int64_t n, float w1, float b2, float one_m_b2,
                                                   float step_size, float bc2_sqrt, float eps) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        float gi = g[i], mi = m[i], vi = v[i];
        float diff = gi - mi;
        // torch lerp: weight < 0.5 ? start + w*diff : end - diff*(1-w)
        mi = (w1 < 0.5f) ? __fmaf_rn(w1, diff, mi) : __fmaf_rn(-diff, 1.0f - w1, gi);
        vi = __fmaf_rn(one_m_b2 * gi, gi, vi * b2);
        float denom = __fdiv_rn(__fsqrt_rn(vi), bc2_sqrt) + eps;
        W[i] = __fmaf_rn(-step_size, __fdiv_rn(mi, denom), W[i]);
        m[i] = mi;
        v[i] = vi;
    }
}

// Same rule with the step count read from device memory, so a captured CUDA graph of the training step can be
// replayed: t = *steps_done + 1; bias corrections in double as on the host path.  step_bump_kernel advances it.
__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ W, float* __restrict__ m,
                                                       float* __restrict__ v, const float* __restrict__ g,
                                                       int64_t n, double lr, double b1, double b2, float eps,
                                                       const int64_t* __restrict__ steps_done) {
    const double t = (double)(*steps_done + 1);
    const float w1 = (float)(1.0 - b1), b2f = (float)b2, one_m_b2 = (float)(1.0 - b2);
    const float step_size = (float)(lr / (1.0 - pow(b1, t)));
    const float bc2_sqrt = (float)sqrt(1.0 - pow(b2, t));
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        float gi = g[i], mi = m[i], vi = v[i];
        float diff = gi - mi;
        mi = (w1 < 0.5f) ? __fmaf_rn(w1, diff, mi) : __fmaf_rn(-diff, 1.0f - w1, gi);
        vi = __fmaf_rn(one_m_b2 * gi, gi, vi * b2f);
        float denom = __fdiv_rn(__fsqrt_rn(vi), bc2_sqrt) + eps;
        W[i] = __fmaf_rn(-step_size, __fdiv_rn(mi, denom), W[i]);
        m[i] = mi;
        v[i] = vi;
    }
}
__global__ void step_bump_kernel(int64_t* steps_done) { *steps_done += 1; }

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ W, int D,
                                                          const int64_t* __restrict__ keep,
                                                          int64_t n_keep, float* __restrict__ out) {
    int64_t total = n_keep * D;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += stride) {
        int64_t r = t / D;
        int d = (int)(t - r * D);
        out[t] = W[keep[r] * (int64_t)D + d];
    }
}

static inline int grid_for(int64_t work_items, int threads, int per_sm) {
    int64_t want = ceil_div64(work_items, threads);
    int64_t cap = (int64_t)sm_count() * per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

}  // namespace som

using namespace som;

extern "C" {

int som_version(void) { return SOM_ABI_VERSION; }

const char* som_last_error(void) { return som::g_err; }

uint64_t som_launch_count(void) { return som::launch_count(); }

int som_device_info(int* sm, int* major, int* minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return (int)e; }
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) { set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return (int)e; }
    if (sm) *sm = p.multiProcessorCount;
    if (major) *major = p.major;
    if (minor) *minor = p.minor;
    return SOM_OK;
}

int som_prepare_codebook_f32(const float* W, int K, int D, float* c_norm2, void* stream) {
    SOM_REQUIRE(W && c_norm2, SOM_E_BADARG, "prepare_codebook: null pointer");
    SOM_REQUIRE(K > 0 && D > 0, SOM_E_BADARG, "prepare_codebook: K=%d D=%d", K, D);
    int threads = 256;
    int blocks = (int)ceil_div64((int64_t)K * 32, threads);
    norm2_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(W, K, D, c_norm2);
    return check_launch("norm2_kernel");
}

int som_merge_candidates(const float* rd, const int64_t* idx, int R, int64_t n,
                         int64_t* out_idx, float* out_rd, void* stream) {
    SOM_REQUIRE(rd && idx && out_idx, SOM_E_BADARG, "merge_candidates: null pointer");
    SOM_REQUIRE(R > 0 && n >= 0, SOM_E_BADARG, "merge_candidates: R=%d n=%lld", R, (long long)n);
    if (n == 0) return SOM_OK;
    int blocks = (int)ceil_div64(n, 256);
    merge_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(rd, idx, R, n, out_idx, out_rd);
    return check_launch("merge_kernel");
}

int som_histogram_i64(const int64_t* idx, int64_t n, int K, int64_t* counts, void* stream) {
    SOM_REQUIRE(idx && counts, SOM_E_BADARG, "histogram: null pointer");
    SOM_REQUIRE(K > 0 && n >= 0, SOM_E_BADARG, "histogram: K=%d n=%lld", K, (long long)n);
    if (n == 0) return SOM_OK;
    auto* c = reinterpret_cast<unsigned long long*>(counts);
    if (K <= 12288) {   // 48 KB of private counters: no opt-in needed, 4 CTAs/SM
        int blocks = grid_for(n, 512 * 8, 4);
        hist_smem_kernel<<<blocks, 512, (size_t)K * sizeof(unsigned int), (cudaStream_t)stream>>>(idx, n, K, c);
        return check_launch("hist_smem_kernel");
    }
    int blocks = grid_for(n, 256 * 4, 8);
    hist_global_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(idx, n, K, c);
    return check_launch("hist_global_kernel");
}

int som_quantize_nchw_f32(const int64_t* idx, const float* table, int K,
                          int64_t n_img, int C, int H, int Wd, int pH, int pW,
                          float* out, void* stream) {
    SOM_REQUIRE(idx && table && out, SOM_E_BADARG, "quantize: null pointer");
    SOM_REQUIRE(K > 0, SOM_E_BADARG, "quantize: K=%d", K);
    Geom g;
    int rc = make_geom(&g, out, n_img, C, H, Wd, pH, pW);
    if (rc) return rc;
    if (g.n_patches == 0) return SOM_OK;
    int vec = g.vec;
    if (((uintptr_t)table & 15) != 0 || g.D % 4 != 0) vec = (vec == 4) ? 1 : vec;
    if (vec == 2 && (((uintptr_t)table & 7) != 0 || g.D % 2 != 0)) vec = 1;
    int64_t items = g.n_patches * (g.D / vec);
    int blocks = grid_for(items, 256, 16);
    cudaStream_t s = (cudaStream_t)stream;
    if (vec == 4) quantize_kernel<4><<<blocks, 256, 0, s>>>(idx, table, K, g, out);
    else if (vec == 2) quantize_kernel<2><<<blocks, 256, 0, s>>>(idx, table, K, g, out);
    else quantize_kernel<1><<<blocks, 256, 0, s>>>(idx, table, K, g, out);
    return check_launch("quantize_kernel");
}

int som_assemble_tokens_i64(const int64_t* lr_idx, const int64_t* hr_idx, int64_t n, int lr_seq, int hr_seq,
                            int64_t lr_K, int64_t hr_K, int base_model, int64_t* hr_input, int64_t* hr_target,
                            void* stream) {
    SOM_REQUIRE(hr_idx && hr_input && hr_target, SOM_E_BADARG, "assemble_tokens: null pointer");
    SOM_REQUIRE(!base_model || lr_idx, SOM_E_BADARG, "assemble_tokens: base model needs lr_idx");
    SOM_REQUIRE(n >= 0 && hr_seq > 0 && lr_seq >= 0, SOM_E_BADARG, "assemble_tokens: n=%lld lr_seq=%d hr_seq=%d",
                (long long)n, lr_seq, hr_seq);
    if (n == 0) return SOM_OK;
    const int in_w = base_model ? lr_seq + hr_seq : 1 + hr_seq;
    const int64_t items = n * (int64_t)(in_w + hr_seq + 1);
    int blocks = grid_for(items, 256, 16);
    assemble_tokens_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(lr_idx, hr_idx, n, lr_seq, hr_seq, lr_K, hr_K,
                                                                     base_model, hr_input, hr_target);
    return check_launch("assemble_tokens_kernel");
}

int som_adam_f32(float* W, float* m, float* v, const float* g, int64_t n,
                 double lr, double b1, double b2, double eps, int64_t step, void* stream) {
    SOM_REQUIRE(W && m && v && g, SOM_E_BADARG, "adam: null pointer");
    SOM_REQUIRE(n >= 0 && step >= 1, SOM_E_BADARG, "adam: n=%lld step=%lld", (long long)n, (long long)step);
    if (n == 0) return SOM_OK;
    // scalars in double on the host exactly as torch/optim/adam.py does, then narrowed
    double bc1 = 1.0 - pow(b1, (double)step);
    double bc2 = 1.0 - pow(b2, (double)step);
    double step_size = lr / bc1;
    double bc2_sqrt = sqrt(bc2);
    int blocks = grid_for(n, 256, 16);
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        W, m, v, g, n, (float)(1.0 - b1), (float)b2, (float)(1.0 - b2),
        (float)step_size, (float)bc2_sqrt, (float)eps);
    return check_launch("adam_kernel");
}

int som_adam_devstep_f32(float* W, float* m, float* v, const float* g, int64_t n,
                         double lr, double b1, double b2, double eps, int64_t* steps_done, void* stream) {
    SOM_REQUIRE(W && m && v && g && steps_done, SOM_E_BADARG, "adam(devstep): null pointer");
    SOM_REQUIRE(n >= 0, SOM_E_BADARG, "adam(devstep): n=%lld", (long long)n);
    if (n > 0) {
        int blocks = grid_for(n, 256, 16);
        adam_dev_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, m, v, g, n, lr, b1, b2, (float)eps, steps_done);
        int rc = check_launch("adam_dev_kernel");
        if (rc) return rc;
    }
    step_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(steps_done);
    return check_launch("step_bump_kernel");
}

int som_gather_rows_f32(const float* W, int D, const int64_t* keep, int64_t n_keep,
                        float* out, void* stream) {
    SOM_REQUIRE(W && keep && out, SOM_E_BADARG, "gather_rows: null pointer");
    SOM_REQUIRE(D > 0 && n_keep >= 0, SOM_E_BADARG, "gather_rows: D=%d n_keep=%lld", D, (long long)n_keep);
    if (n_keep == 0) return SOM_OK;
    int blocks = grid_for(n_keep * D, 256, 16);
    gather_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, D, keep, n_keep, out);
    return check_launch("gather_rows_kernel");
}

}  // extern "C"
