// PTX helpers shared by the tcgen05 BMU kernels (sm_100a): mbarrier, TMA, tcgen05 MMA / TMEM loads.
#pragma once
#include <cuda.h>
#include <stdint.h>

#include "som_common.cuh"

namespace som {
namespace tc {

constexpr int TM = 128;              // patches per MMA tile (UMMA M)
constexpr int TN = 256;              // units per MMA tile (UMMA N)
constexpr int KBLK = 32;             // floats per k-block: one 128-byte swizzle row
constexpr int A_BLK_BYTES = TM * KBLK * 4;     // 16 KB
constexpr int B_BLK_BYTES = TN * KBLK * 4;     // 32 KB
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) |
                           ((uint32_t)(TM >> 4) << 24);
constexpr float PAD_NORM = 1.0e30f;            // ||c||^2 of padding units: never the minimum

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU.  The non-suspending
// test_wait is the fast path (the phase has usually completed already); try_wait sleeps otherwise.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_test(bar, parity)) return;
    for (uint32_t spin = 0; !mbar_try(bar, parity); ++spin)
        if (spin > (1u << 20)) __trap();
}
// warp-level wait: ONE lane polls (with optional back-off for long waits), the rest of the warp parks at
// __syncwarp.  Hundreds of threads spinning on try_wait saturate the barrier unit and slow down the single
// MMA-issuing thread's own barrier checks (measured: ~100+ cycles per check under mass polling).
template <bool BACKOFF>
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
    if (lane == 0) {
        if (!mbar_test(bar, parity)) {
            for (uint32_t spin = 0; !mbar_try(bar, parity); ++spin) {
                if (BACKOFF) __nanosleep(spin < 64 ? 64 : 256);
                if (spin > (1u << 20)) __trap();
            }
        }
    }
    __syncwarp();
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// ---- CTA-pair (cta_group::2) helpers ------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_count_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(cta)
        : "memory");
}
// The same arrival WITHOUT release semantics: for signals that publish no memory writes (an epilogue warp handing a
// TMEM accumulator back after tcgen05.ld + tcgen05.fence::before_thread_sync).  The release form compiles to
// MEMBAR.ALL.GPU + ERRBAR in front of the arrive (ncu: 16 % of all stall samples of the FP16-split kernel, on the
// critical path of every tile).
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(cta)
        : "memory");
}
// bounded wait with cluster-scope acquire (barriers that peer CTAs arrive on)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (spin > (1u << 20)) __trap();
    }
}
// TMA load issued by either CTA of a pair; the transaction bytes land on the LEADER's barrier (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
// M = 128 * CG rows (one 128-row half per CTA), N = 256, K = 8, kind::tf32, K-major operands, fp32 accumulate
template <int CG>
__device__ __forceinline__ void tc_mma_tf32_cg(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accum) {
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) |
                               ((uint32_t)((TM * CG) >> 4) << 24);
    if (CG == 1) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    }
}
// commit: arrive (once the issued MMAs completed) on the barrier at this offset in every CTA of the pair
template <int CG>
__device__ __forceinline__ void tc_commit_cg(uint64_t* bar) {
    if (CG == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                     : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .b16 m;\n"
            "mov.b16 m, 3;\n"
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n"
            "}\n" ::"r"(smem_u32(bar))
            : "memory");
    }
}

// one lane of a converged warp (elect.sync): the MMA warp keeps warp-uniform control flow, so descriptors
// live in uniform registers, and only the tcgen05 instructions themselves are predicated on the elected lane
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, %1;\n"
        "@px mov.s32 %0, 1;\n"
        "}\n"
        : "+r"(pred)
        : "r"(0xFFFFFFFFu));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accum)
        : "memory");
}
// kind::f16 with FP16 operands (a/b format 0), fp32 accumulate: M128 x N256 x K16 at the cycle cost of the K8 TF32 MMA
constexpr uint32_t IDESC_F16 = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TN >> 3) << 17) |
                               ((uint32_t)(TM >> 4) << 24);
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC_F16), "r"(accum)
        : "memory");
}
// kind::f16 (FP16 operands, fp32 accumulate), M = 128 * CG rows (one 128-row half per CTA), N = 256, K = 16
template <int CG>
__device__ __forceinline__ void tc_mma_f16_cg(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accum) {
    constexpr uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TN >> 3) << 17) |
                               ((uint32_t)((TM * CG) >> 4) << 24);
    if (CG == 1) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    }
}
// K-major, SWIZZLE_128B operand descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO=1 | SBO=1024>>4
// | version=1 | layout_type=2.  Advancing one K=8 step inside the 128-byte row adds 32 B (2 units).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

#define SOM_R32(a) "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), \
    "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]),           \
    "=r"(a[16]), "=r"(a[17]), "=r"(a[18]), "=r"(a[19]), "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]),         \
    "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]), "=r"(a[30]), "=r"(a[31])
#define SOM_RW32(a) "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), \
    "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]),           \
    "+r"(a[16]), "+r"(a[17]), "+r"(a[18]), "+r"(a[19]), "+r"(a[20]), "+r"(a[21]), "+r"(a[22]), "+r"(a[23]),         \
    "+r"(a[24]), "+r"(a[25]), "+r"(a[26]), "+r"(a[27]), "+r"(a[28]), "+r"(a[29]), "+r"(a[30]), "+r"(a[31])

// asynchronous TMEM load of 32 consecutive columns of this warp's 32 lanes (no wait)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : SOM_R32(r)
        : "r"(taddr)
        : "memory");
}
// wait for all outstanding TMEM loads; the "+r" operands pin every use of r[] after the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" : SOM_RW32(r)::"memory");
}
__device__ __forceinline__ float min32(const uint32_t (&r)[32]) {
    float m = __uint_as_float(r[0]);
#pragma unroll
    for (int i = 1; i < 32; ++i) m = fminf(m, __uint_as_float(r[i]));
    return m;
}
__device__ __forceinline__ int first_eq32(const uint32_t (&r)[32], float m) {
    int q = 31;
#pragma unroll
    for (int i = 30; i >= 0; --i) q = (__uint_as_float(r[i]) == m) ? i : q;
    return q;
}
__device__ __forceinline__ float tf32_rna(float v) {
    uint32_t o;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(o) : "f"(v));
    return __uint_as_float(o);
}


// builder-side loader: D features of patch row `src` into registers (static indexing, DCAP >= D)
template <int DCAP>
__device__ __forceinline__ void load_row(float (&xr)[DCAP], const float* src, bool ok, int D, int vec,
                                         const int* foff) {
#pragma unroll
    for (int d = 0; d < DCAP; d += 4) {
        float tmp[4] = {0.f, 0.f, 0.f, 0.f};
        if (ok && d < D) {
            if (vec == 4) {
                float4 v = __ldg(reinterpret_cast<const float4*>(src + foff[d]));
                tmp[0] = v.x; tmp[1] = v.y; tmp[2] = v.z; tmp[3] = v.w;
            } else if (vec == 2) {
                float2 v0 = __ldg(reinterpret_cast<const float2*>(src + foff[d]));
                tmp[0] = v0.x; tmp[1] = v0.y;
                if (d + 2 < D) {
                    float2 v1 = __ldg(reinterpret_cast<const float2*>(src + foff[d + 2]));
                    tmp[2] = v1.x; tmp[3] = v1.y;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (d + e < D) tmp[e] = __ldg(src + foff[d + e]);
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) xr[d + e] = tmp[e];
    }
}


// ---- host side: tensor maps ------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (EncodeTiledFn)p;
    return fn;
}

// general 2-D fp32 tensor map: `cols` floats per row (row pitch `pitch_bytes`), box = box_cols x box_rows
static inline int make_map2d(CUtensorMap* map, void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                             uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = get_encode_fn();
    SOM_REQUIRE(fn != nullptr, SOM_E_UNSUPPORTED, "bmu(tc): cuTensorMapEncodeTiled is not available");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SOM_REQUIRE(r == CUDA_SUCCESS, SOM_E_UNSUPPORTED, "bmu(tc): cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SOM_OK;
}

}  // namespace tc
}  // namespace som
