// K1 (tensor-core variant): BMU search as an error-compensated 3xTF32 GEMM on tcgen05 with the
// argmin fused into the TMEM epilogue.  sm_100a only.
//
// Replaces patchify + torch.cdist + torch.argmin of Codebook.get_patches_bmu
// (/root/reference/models/Codebook.py:77-99).  The reduced distance
//     rd[p][j] = ||c_j||^2 - 2 x_p . c_j          (d^2 = rd + ||x_p||^2, row-constant dropped)
// is ONE GEMM over an augmented inner dimension K' = 3D + 3:
//     A'[p] = [ x_hi | x_lo | x_hi | 1 1 1 | 0.. ]            (patches, 128-row tiles -> UMMA M)
//     B'[j] = [-2c_hi|-2c_hi|-2c_lo| n1 n2 n3 | 0.. ]         (units, 256-row tiles  -> UMMA N)
// hi = RNA-rounded TF32 part, lo = TF32-rounded remainder, n1+n2+n3 = ||c_j||^2.  Every TF32
// product is exact in the fp32 accumulator; the dropped lo*lo term is 2^-24 relative, so the result
// is fp32-faithful and BMU indices stay stable.
//
// Pipeline per CTA (persistent, 1 CTA/SM, 320 threads):
//   warp 0   TMA producer : cp.async.bulk.tensor (SWIZZLE_128B) of A'/B' k-blocks (32 floats = 128 B
//                           rows) into shared memory, mbarrier full/empty ring
//   warp 1   MMA issuer   : one elected thread, tcgen05.mma.cta_group::1.kind::tf32, M=128 N=256 K=8,
//                           fp32 accumulators in TMEM, 2 accumulator stages x 256 columns
//   warps2-9 epilogue     : tcgen05.ld 32x32b of their TMEM lane quarter (warp%4) and column half,
//                           running (min, first index) per patch row in registers; the two column
//                           halves merge through shared memory once per patch tile
// Operand reuse: for small K' (<= 64 floats) R patch tiles stay resident in shared memory and each
// streamed unit tile feeds R MMAs; for K' <= 224 one patch tile stays resident and unit k-blocks
// stream; otherwise both operands stream through a 4-stage ring.
// The operand split (with fused patchify) is a pre-pass into an L2-sized workspace chunk.
// Bound: tensor pipe at TF32 rate / 3 -- algorithmic 2*K*D flop per patch.
#include "som_common.cuh"

#include <cuda.h>

namespace som {

namespace tc {

constexpr int TM = 128;              // patches per MMA tile (UMMA M)
constexpr int TN = 256;              // units per MMA tile (UMMA N)
constexpr int KBLK = 32;             // floats per k-block: one 128-byte swizzle row
constexpr int A_BLK_BYTES = TM * KBLK * 4;     // 16 KB
constexpr int B_BLK_BYTES = TN * KBLK * 4;     // 32 KB
constexpr int NUM_THREADS = 320;
constexpr int EPI_THREADS = 256;
constexpr int MAX_STAGES = 4;
constexpr int MAX_R = 4;
constexpr uint32_t SMEM_LIMIT = 232448;        // 227 KB
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) |
                           ((uint32_t)(TM >> 4) << 24);
constexpr float PAD_NORM = 1.0e30f;            // ||c||^2 of padding units: never the minimum

struct Params {
    int KB;                 // k-blocks per row (KP / 32)
    int ksteps;             // MMA k-steps carrying data: ceil((3D+3)/8)
    int NT;                 // unit tiles (K_pad / 256)
    int R;                  // resident patch tiles per super-tile
    int a_resident;         // A' tiles stay in shared memory across unit tiles
    int stage_kb;           // k-blocks of B' per ring stage (KB: whole unit tile, or 1)
    int n_stages;
    int n_mtiles;           // patch tiles in this chunk
    int64_t rows;           // valid patches in this chunk
    int64_t unit_offset;
    int64_t* out_idx;       // already offset to the chunk's first patch
    float* out_rd;          // idem or nullptr
    uint32_t a_bytes;       // shared-memory bytes of the resident A region
    uint32_t stage_bytes;   // bytes per ring stage
};

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
// bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try(bar, parity); ++spin)
        if (spin > (1u << 20)) __trap();
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
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
// K-major, SWIZZLE_128B operand descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO=1 | SBO=1024>>4
// | version=1 | layout_type=2.  Advancing one K=8 step inside the 128-byte row adds 32 B (2 units).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tf32_rna(float v) {
    uint32_t o;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(o) : "f"(v));
    return __uint_as_float(o);
}

struct __align__(8) Barriers {
    uint64_t full[MAX_STAGES];
    uint64_t empty[MAX_STAGES];
    uint64_t a_full[MAX_R];
    uint64_t a_empty[MAX_R];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

// ---- the GEMM + argmin kernel ----------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(NUM_THREADS, 1)
bmu_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
              const Params P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Barriers bars;
    __shared__ float mrg_val[TM];
    __shared__ int mrg_idx[TM];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_res = tiles;                       // [R][KB] blocks of 16 KB   (a_resident)
    uint8_t* ring = tiles + P.a_bytes;            // [n_stages] stages

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); }
        for (int r = 0; r < MAX_R; ++r) { mbar_init(&bars.a_full[r], 1); mbar_init(&bars.a_empty[r], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&bars.acc_full[a], 1); mbar_init(&bars.acc_empty[a], EPI_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_base)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars.tmem_base;

    const int n_super = (P.n_mtiles + R - 1) / R;
    const int groups = P.KB / P.stage_kb;          // ring stages consumed per unit tile

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, a_epar = 1;
            for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
                const int r_eff = min(R, P.n_mtiles - st * R);
                const int m_base = st * R * TM;
                if (P.a_resident) {
                    for (int r = 0; r < r_eff; ++r) {
                        mbar_wait(&bars.a_empty[r], a_epar);
                        mbar_expect_tx(&bars.a_full[r], (uint32_t)P.KB * A_BLK_BYTES);
                        for (int kb = 0; kb < P.KB; ++kb)
                            tma_load_2d(&map_a, &bars.a_full[r], a_res + (size_t)(r * P.KB + kb) * A_BLK_BYTES,
                                        kb * KBLK, m_base + r * TM);
                    }
                    a_epar ^= 1;
                }
                for (int n = 0; n < P.NT; ++n) {
                    for (int g = 0; g < groups; ++g) {
                        mbar_wait(&bars.empty[stage], phase ^ 1);
                        uint8_t* sbase = ring + (size_t)stage * P.stage_bytes;
                        mbar_expect_tx(&bars.full[stage], P.stage_bytes);
                        for (int j = 0; j < P.stage_kb; ++j)
                            tma_load_2d(&map_b, &bars.full[stage], sbase + (size_t)j * B_BLK_BYTES,
                                        (g * P.stage_kb + j) * KBLK, n * TN);
                        if (!P.a_resident)
                            tma_load_2d(&map_a, &bars.full[stage], sbase + B_BLK_BYTES, g * KBLK, m_base);
                        if (++stage == P.n_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0, a_fpar = 0;
            for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
                const int r_eff = min(R, P.n_mtiles - st * R);
                for (int n = 0; n < P.NT; ++n) {
                    if (P.stage_kb == P.KB) {
                        // whole unit tile in one stage; feeds r_eff resident patch tiles
                        mbar_wait(&bars.full[stage], phase);
                        const uint32_t b_addr = smem_u32(ring + (size_t)stage * P.stage_bytes);
                        for (int r = 0; r < r_eff; ++r) {
                            uint32_t a_addr;
                            if (P.a_resident) {
                                if (n == 0) mbar_wait(&bars.a_full[r], a_fpar);
                                a_addr = smem_u32(a_res + (size_t)r * P.KB * A_BLK_BYTES);
                            } else {
                                a_addr = b_addr + B_BLK_BYTES;       // (KB == 1, R == 1)
                            }
                            mbar_wait(&bars.acc_empty[acc], acc_phase ^ 1);
                            tc_fence_after();
                            const uint32_t d_addr = tmem_base + (uint32_t)acc * TN;
                            for (int ks = 0; ks < P.ksteps; ++ks) {
                                const int kb = ks >> 2, k = ks & 3;
                                tc_mma_tf32(d_addr, umma_desc(a_addr + kb * A_BLK_BYTES + k * 32),
                                            umma_desc(b_addr + kb * B_BLK_BYTES + k * 32), ks > 0);
                            }
                            tc_commit(&bars.acc_full[acc]);
                            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                        }
                        tc_commit(&bars.empty[stage]);
                        if (++stage == P.n_stages) { stage = 0; phase ^= 1; }
                    } else {
                        // one k-block per stage, one patch tile per super-tile
                        if (P.a_resident && n == 0) mbar_wait(&bars.a_full[0], a_fpar);
                        mbar_wait(&bars.acc_empty[acc], acc_phase ^ 1);
                        tc_fence_after();
                        const uint32_t d_addr = tmem_base + (uint32_t)acc * TN;
                        for (int kb = 0; kb < P.KB; ++kb) {
                            mbar_wait(&bars.full[stage], phase);
                            tc_fence_after();
                            const uint32_t b_addr = smem_u32(ring + (size_t)stage * P.stage_bytes);
                            const uint32_t a_addr = P.a_resident ? smem_u32(a_res + (size_t)kb * A_BLK_BYTES)
                                                                 : b_addr + B_BLK_BYTES;
                            const int nk = min(4, P.ksteps - kb * 4);
                            for (int k = 0; k < nk; ++k)
                                tc_mma_tf32(d_addr, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32),
                                            (kb | k) != 0);
                            tc_commit(&bars.empty[stage]);
                            if (++stage == P.n_stages) { stage = 0; phase ^= 1; }
                        }
                        tc_commit(&bars.acc_full[acc]);
                        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                    }
                }
                if (P.a_resident) {
                    for (int r = 0; r < r_eff; ++r) tc_commit(&bars.a_empty[r]);
                    a_fpar ^= 1;
                }
            }
        }
    } else {
        // ================================ epilogue ====================================
        const int ew = warp - 2;                    // 0..7
        const int half = ew >> 2;                   // column half of the accumulator
        const int lg = warp & 3;                    // TMEM lane quarter this warp may access
        const int row = lg * 32 + lane;             // patch row inside the tile
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
            const int r_eff = min(R, P.n_mtiles - st * R);
            float best[R];
            int bidx[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { best[r] = INFINITY; bidx[r] = 0; }
            for (int n = 0; n < P.NT; ++n) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (r < r_eff) {
                        mbar_wait(&bars.acc_full[acc], acc_phase);
                        tc_fence_after();
                        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * TN + half * 128);
                        const int col0 = n * TN + half * 128;
#pragma unroll 1
                        for (int c = 0; c < 4; ++c) {
                            float v[32];
                            tmem_ld32(taddr + c * 32, v);
                            float m = v[0];
#pragma unroll
                            for (int i = 1; i < 32; ++i) m = fminf(m, v[i]);
                            if (m < best[r]) {
                                best[r] = m;
                                int q = 31;
#pragma unroll
                                for (int i = 30; i >= 0; --i) q = (v[i] == m) ? i : q;
                                bidx[r] = col0 + c * 32 + q;
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars.acc_empty[acc]);
                        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                    }
                }
            }
            // merge the two column halves (score asc, index asc) and store, one patch tile at a time
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r < r_eff) {
                    if (half == 1) { mrg_val[row] = best[r]; mrg_idx[row] = bidx[r]; }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (half == 0) {
                        float ov = mrg_val[row];
                        int oi = mrg_idx[row];
                        float bv = best[r];
                        int bi = bidx[r];
                        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                        const int64_t p = (int64_t)(st * R + r) * TM + row;
                        if (p < P.rows) {
                            P.out_idx[p] = (int64_t)bi + P.unit_offset;
                            if (P.out_rd) P.out_rd[p] = bv;
                        }
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---- operand split pre-passes ----------------------------------------------------------------------
// B'[j] = [-2 hi(c) | -2 hi(c) | -2 lo(c) | n1 n2 n3 | 0..], rows >= K are padding units
__global__ void __launch_bounds__(256) split_w_kernel(const float* __restrict__ W, const float* __restrict__ cn,
                                                      int K, int D, int K_pad, int KP, float* __restrict__ Bp) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)K_pad * KP) return;
    int row = (int)(t / KP);
    int kp = (int)(t - (int64_t)row * KP);
    float out = 0.f;
    if (row < K) {
        if (kp < 3 * D) {
            int seg = kp / D;
            int d = kp - seg * D;
            float w = W[(int64_t)row * D + d];
            float hi = tf32_rna(w);
            out = -2.0f * (seg < 2 ? hi : tf32_rna(w - hi));
        } else if (kp < 3 * D + 3) {
            float nrm = cn[row];
            float n1 = tf32_rna(nrm);
            float n2 = tf32_rna(nrm - n1);
            float n3 = tf32_rna(nrm - n1 - n2);
            out = (kp == 3 * D) ? n1 : (kp == 3 * D + 1 ? n2 : n3);
        }
    } else if (kp == 3 * D) {
        out = PAD_NORM;
    }
    Bp[t] = out;
}

// A'[p] = [ hi(x) | lo(x) | hi(x) | 1 1 1 | 0.. ] with patchify fused as address arithmetic.
// One thread per (patch, feature); consecutive threads walk the features of one patch.
__global__ void __launch_bounds__(256) split_x_kernel(const float* __restrict__ x, Geom g, int64_t p0,
                                                      int64_t rows, int KP, float* __restrict__ Ap) {
    const int D = g.D;
    const int per = D + 1;                         // the extra slot writes the constant tail
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= rows * per) return;
    int64_t pr = t / per;
    int d = (int)(t - pr * per);
    float* dst = Ap + pr * KP;
    if (d < D) {
        float v = __ldg(x + patch_base(g, p0 + pr) + feat_off(g, d));
        float hi = tf32_rna(v);
        float lo = tf32_rna(v - hi);
        dst[d] = hi;
        dst[D + d] = lo;
        dst[2 * D + d] = hi;
    } else {
        for (int k = 3 * D; k < KP; ++k) dst[k] = (k < 3 * D + 3) ? 1.0f : 0.f;
    }
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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

static int make_map(CUtensorMap* map, void* base, uint64_t rows, uint64_t kp, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    SOM_REQUIRE(fn != nullptr, SOM_E_UNSUPPORTED, "bmu(tc): cuTensorMapEncodeTiled is not available");
    cuuint64_t dims[2] = {kp, rows};
    cuuint64_t strides[1] = {kp * 4};
    cuuint32_t box[2] = {KBLK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SOM_REQUIRE(r == CUDA_SUCCESS, SOM_E_UNSUPPORTED, "bmu(tc): cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SOM_OK;
}

struct Plan {
    int D, K, KP, KB, ksteps, K_pad, NT;
    int R, a_resident, stage_kb, n_stages;
    uint32_t a_bytes, stage_bytes, smem_bytes;
    int64_t chunk_rows;            // patches per workspace chunk (multiple of R * 128)
    size_t off_b, off_a, total;
};

static void make_plan(Plan* pl, int64_t n, int D, int K) {
    pl->D = D; pl->K = K;
    const int kprime = 3 * D + 3;
    pl->KP = (kprime + KBLK - 1) / KBLK * KBLK;
    pl->KB = pl->KP / KBLK;
    pl->ksteps = (kprime + 7) / 8;
    pl->K_pad = (K + TN - 1) / TN * TN;
    pl->NT = pl->K_pad / TN;
    const uint32_t budget = SMEM_LIMIT - 1024 /*alignment slack*/ - 2048 /*static: barriers + merge*/;
    if (pl->KB <= 2) {
        pl->a_resident = 1; pl->stage_kb = pl->KB; pl->n_stages = 2;
        pl->stage_bytes = (uint32_t)pl->KB * B_BLK_BYTES;
        int R = MAX_R;
        while (R > 1 && (uint32_t)R * pl->KB * A_BLK_BYTES + 2u * pl->stage_bytes > budget) --R;
        pl->R = R;
        pl->a_bytes = (uint32_t)R * pl->KB * A_BLK_BYTES;
    } else if (pl->KB <= 7) {
        pl->a_resident = 1; pl->stage_kb = 1; pl->R = 1;
        pl->stage_bytes = B_BLK_BYTES;
        pl->a_bytes = (uint32_t)pl->KB * A_BLK_BYTES;
        pl->n_stages = (int)((budget - pl->a_bytes) / pl->stage_bytes);
        if (pl->n_stages > MAX_STAGES) pl->n_stages = MAX_STAGES;
    } else {
        pl->a_resident = 0; pl->stage_kb = 1; pl->R = 1;
        pl->stage_bytes = B_BLK_BYTES + A_BLK_BYTES;
        pl->a_bytes = 0;
        pl->n_stages = MAX_STAGES;
    }
    pl->smem_bytes = pl->a_bytes + (uint32_t)pl->n_stages * pl->stage_bytes + 1024;
    // workspace chunk: keep A' within ~48 MB so it stays L2 resident between the two kernels
    const int64_t super_rows = (int64_t)pl->R * TM;
    const int64_t wave_rows = super_rows * sm_count();
    int64_t max_rows = (48ll << 20) / ((int64_t)pl->KP * 4);
    int64_t chunk = max_rows / wave_rows * wave_rows;
    if (chunk < wave_rows) chunk = (max_rows / super_rows > 0 ? max_rows / super_rows : 1) * super_rows;
    int64_t need = ceil_div64(n, super_rows) * super_rows;
    if (chunk > need) chunk = need;
    pl->chunk_rows = chunk;
    size_t o = 0;
    pl->off_b = o; o = align_up(o + (size_t)pl->K_pad * pl->KP * 4, 1024);
    pl->off_a = o; o = align_up(o + (size_t)chunk * pl->KP * 4, 1024);
    pl->total = o;
}

template <int R>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const Params& P, uint32_t smem, int grid,
                       cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(bmu_tc_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(SMEM_LIMIT - 2048));
        if (e != cudaSuccess) { set_error("bmu(tc): smem opt-in: %s", cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    bmu_tc_kernel<R><<<grid, NUM_THREADS, smem, st>>>(ma, mb, P);
    return check_launch("bmu_tc_kernel");
}

}  // namespace tc

bool tc_supported(int64_t n_patches, int D, int K) {
    if (n_patches <= 0 || D <= 0 || K <= 0) return false;
    if ((int64_t)3 * D + 3 > (1 << 20)) return false;
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
    tc::Plan pl;
    tc::make_plan(&pl, n_patches, D, K);
    return pl.total;
}

int launch_bmu_tc(const float* x, const Geom& g, const float* W, const float* cn, int K,
                  int64_t unit_offset, int64_t* out_idx, float* out_rd, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
    using namespace tc;
    const int64_t n = g.n_patches;
    if (n == 0) return SOM_OK;
    Plan pl;
    make_plan(&pl, n, g.D, K);
    SOM_REQUIRE(ws != nullptr && ws_bytes >= pl.total, SOM_E_WORKSPACE,
                "bmu(tc): workspace %zu < required %zu", ws_bytes, pl.total);
    SOM_REQUIRE(((uintptr_t)ws & 1023) == 0 || ((uintptr_t)ws & 255) == 0, SOM_E_BADARG,
                "bmu(tc): workspace must be 256-byte aligned");
    SOM_REQUIRE(pl.smem_bytes + 2048 <= SMEM_LIMIT, SOM_E_SHAPE, "bmu(tc): shared memory plan too large");
    float* Bp = (float*)((char*)ws + pl.off_b);
    float* Ap = (float*)((char*)ws + pl.off_a);

    {
        int64_t items = (int64_t)pl.K_pad * pl.KP;
        split_w_kernel<<<(unsigned)ceil_div64(items, 256), 256, 0, st>>>(W, cn, K, g.D, pl.K_pad, pl.KP, Bp);
        int rc = check_launch("split_w_kernel");
        if (rc) return rc;
    }
    CUtensorMap map_a, map_b;
    int rc = make_map(&map_b, Bp, (uint64_t)pl.K_pad, (uint64_t)pl.KP, TN);
    if (rc) return rc;
    rc = make_map(&map_a, Ap, (uint64_t)pl.chunk_rows, (uint64_t)pl.KP, TM);
    if (rc) return rc;

    Params P;
    P.KB = pl.KB; P.ksteps = pl.ksteps; P.NT = pl.NT; P.R = pl.R; P.a_resident = pl.a_resident;
    P.stage_kb = pl.stage_kb; P.n_stages = pl.n_stages; P.unit_offset = unit_offset;
    P.a_bytes = pl.a_bytes; P.stage_bytes = pl.stage_bytes;

    for (int64_t p0 = 0; p0 < n; p0 += pl.chunk_rows) {
        const int64_t rows = (n - p0 < pl.chunk_rows) ? n - p0 : pl.chunk_rows;
        {
            int64_t items = rows * (g.D + 1);
            split_x_kernel<<<(unsigned)ceil_div64(items, 256), 256, 0, st>>>(x, g, p0, rows, pl.KP, Ap);
            rc = check_launch("split_x_kernel");
            if (rc) return rc;
        }
        P.rows = rows;
        P.n_mtiles = (int)ceil_div64(rows, TM);
        P.out_idx = out_idx + p0;
        P.out_rd = out_rd ? out_rd + p0 : nullptr;
        const int n_super = (P.n_mtiles + pl.R - 1) / pl.R;
        const int grid = n_super < sm_count() ? n_super : sm_count();
        switch (pl.R) {
            case 4: rc = launch_gemm<4>(map_a, map_b, P, pl.smem_bytes, grid, st); break;
            case 3: rc = launch_gemm<3>(map_a, map_b, P, pl.smem_bytes, grid, st); break;
            case 2: rc = launch_gemm<2>(map_a, map_b, P, pl.smem_bytes, grid, st); break;
            default: rc = launch_gemm<1>(map_a, map_b, P, pl.smem_bytes, grid, st); break;
        }
        if (rc) return rc;
    }
    return SOM_OK;
}

}  // namespace som
