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
// One persistent CTA per SM, 448 threads, warp-specialised:
//   warp 0     TMA producer : cp.async.bulk.tensor (SWIZZLE_128B) of B' (and, for large D, A') k-blocks
//                             of 32 floats = 128-byte rows into a shared-memory mbarrier ring
//   warp 1     MMA issuer   : one elected thread, tcgen05.mma.cta_group::1.kind::tf32, M=128 N=256 K=8,
//                             fp32 accumulators in TMEM, 2 accumulator stages x 256 columns
//   warps 2-5  A' builders  : (D <= 73) read patches straight from NCHW (patchify = address arithmetic,
//                             rows prefetched into registers one super-tile ahead), split hi/lo and
//                             write the SWIZZLE_128B operand tile into shared memory themselves
//                             (generic stores + fence.proxy.async) -- no intermediate in HBM
//   warps 6-13 epilogue     : tcgen05.ld 32x32b of their TMEM lane quarter (warp%4) and column half,
//                             double-buffered, running minimum per patch row in registers
// Three static configurations (no autotuner):
//   S  K' <= 64  (D <= 20): R=3 patch tiles resident, each streamed 256-unit tile feeds 3 MMAs; the
//      epilogue tracks (min, 8-unit chunk) only -- under one ALU op per distance -- and the builder
//      warps resolve the index inside the winning chunk with exact fp32 FFMA scores while the next
//      super-tile computes
//   M  K' <= 224 (D <= 73): one patch tile resident, unit k-blocks stream; exact in-register index scan
//   L  larger D: both operands stream through a 4-stage ring; A' comes from a pre-pass split kernel
// Bound: tensor pipe at TF32 rate / 3 -- algorithmic 2*K*D flop per patch.
#include "som_common.cuh"
#include "som_tc_ptx.cuh"

#include <cuda.h>
#include <stdlib.h>

namespace som {

namespace tc {

constexpr int NUM_THREADS = 448;
constexpr int BUILD_WARP0 = 2, BUILD_WARPS = 4;
constexpr int EPI_WARP0 = 6, EPI_THREADS = 256;
constexpr int MAX_STAGES = 4;
constexpr int MAX_R = 3;
constexpr int CHUNK = 8;             // units per refine chunk (config S)
constexpr int DCAP_S = 20, DCAP_M = 76;
constexpr uint32_t SMEM_LIMIT = 232448;        // 227 KB
constexpr uint32_t STATIC_SMEM = 2048;         // barriers + merge buffers + feature offsets (upper bound)

struct Params {
    int KB;                 // k-blocks per row (KP / 32)
    int KP;                 // padded K' in floats
    int ksteps;             // MMA k-steps carrying data: ceil((3D+3)/8)
    int NT;                 // unit tiles (K_pad / 256)
    int a_resident;         // A' tiles stay in shared memory across unit tiles
    int stage_kb;           // k-blocks of B' per ring stage (KB: whole unit tile, or 1)
    int n_stages;
    int n_mtiles;           // patch tiles in this launch
    int64_t rows;           // valid patches in this launch
    int64_t unit_offset;
    int64_t* out_idx;
    float* out_rd;
    uint32_t a_bytes;       // shared-memory bytes of the resident A region
    uint32_t stage_bytes;   // bytes per ring stage
    const float* x;         // fused builders: NCHW source + geometry
    Geom g;
    const float* W;         // config S refine: original codebook rows, ||c||^2, unit count
    const float* cn;
    int K;
    int dbg;                // SOM_TC_DEBUG bit mask (timing experiments only; results are wrong)
};

struct __align__(8) Barriers {
    uint64_t full[MAX_STAGES];
    uint64_t empty[MAX_STAGES];
    uint64_t a_full[MAX_R];
    uint64_t a_empty[MAX_R];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint64_t ref_full[2];       // config S: chunk bases of a super-tile are in out_idx (alternating)
    uint32_t tmem_base;
    uint32_t pad;
};

// one 128 x 256 x 8 TF32 MMA per k-step; descriptors advance by constants, so a step costs two 32-bit
// adds and the tcgen05.mma itself (the single issuing thread must stay under 128 cycles per step)
template <int KS_MAX>
__device__ __forceinline__ void issue_tile(uint32_t d_addr, uint64_t adesc, uint64_t bdesc, int ksteps) {
#pragma unroll
    for (int ks = 0; ks < KS_MAX; ++ks) {
        if (ks < ksteps) {
            const uint32_t oa = (uint32_t)(((ks >> 2) * A_BLK_BYTES + (ks & 3) * 32) >> 4);
            const uint32_t ob = (uint32_t)(((ks >> 2) * B_BLK_BYTES + (ks & 3) * 32) >> 4);
            tc_mma_tf32(d_addr, adesc + oa, bdesc + ob, ks > 0 ? 1u : 0u);
        }
    }
}

__device__ long long g_tl[5][64];

// ---- the GEMM + argmin kernel ----------------------------------------------------------------------
// R      patch tiles per super-tile (resident A' slots)
// EXACT  epilogue resolves the index in registers (else: winning CHUNK-unit chunk, refined by builders)
// FUSED  builder warps create A' in shared memory from NCHW (else A' arrives by TMA from a pre-pass)
// DCAP   per-row register capacity of the builders (>= D)
template <int R, bool EXACT, bool FUSED, int DCAP, bool PROF = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
bmu_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
              const Params P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Barriers bars;
    __shared__ float mrg_val[TM];
    __shared__ int mrg_idx[TM];
    __shared__ int foff_s[FUSED ? DCAP : 1];
    long long (*tl)[64] = g_tl;                             // PROF: timeline of 64 consecutive tiles (CTA 0)
    (void)tl;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_res = tiles;                       // [R][KB] blocks of 16 KB   (a_resident)
    uint8_t* ring = tiles + P.a_bytes;            // [n_stages] stages

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); }
        for (int r = 0; r < MAX_R; ++r) { mbar_init(&bars.a_full[r], BUILD_WARPS); mbar_init(&bars.a_empty[r], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&bars.acc_full[a], 1); mbar_init(&bars.acc_empty[a], EPI_THREADS / 32); }
        mbar_init(&bars.ref_full[0], EPI_THREADS / 64);
        mbar_init(&bars.ref_full[1], EPI_THREADS / 64);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (FUSED) {
        for (int d = threadIdx.x; d < P.g.D; d += NUM_THREADS) foff_s[d] = feat_off(P.g, d);
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
            uint32_t phase = 0;
            for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
                const int m_base = st * R * TM;
                for (int n = 0; n < P.NT; ++n) {
                    for (int g = 0; g < groups; ++g) {
                        mbar_wait(&bars.empty[stage], phase ^ 1);
                        uint8_t* sbase = ring + (size_t)stage * P.stage_bytes;
                        if ((P.dbg & 4) && (st != (int)blockIdx.x || n * groups + g >= P.n_stages)) {
                            mbar_arrive(&bars.full[stage]);
                            if (++stage == P.n_stages) { stage = 0; phase ^= 1; }
                            continue;
                        }
                        mbar_expect_tx(&bars.full[stage], P.stage_bytes);
                        for (int j = 0; j < P.stage_kb; ++j)
                            tma_load_2d(&map_b, &bars.full[stage], sbase + (size_t)j * B_BLK_BYTES,
                                        (g * P.stage_kb + j) * KBLK, n * TN);
                        if (!FUSED)
                            tma_load_2d(&map_a, &bars.full[stage], sbase + B_BLK_BYTES, g * KBLK, m_base);
                        if (++stage == P.n_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0 && P.stage_kb == P.KB) {
            // ---- config S: whole unit tile per stage, R resident patch tiles, <= 8 k-steps per tile
            uint64_t adesc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) adesc[r] = umma_desc(smem_u32(a_res + (size_t)r * P.KB * A_BLK_BYTES));
            const uint64_t bdesc0 = umma_desc(smem_u32(ring));
            const uint32_t stage_units = P.stage_bytes >> 4;
            long long pt[3] = {0, 0, 0};
            const long long pt_begin = PROF ? clock64() : 0;
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0, a_fpar = 0;
            for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
                const int r_eff = min(R, P.n_mtiles - st * R);
                for (int n = 0; n < P.NT; ++n) {
                    long long c0 = PROF ? clock64() : 0;
                    mbar_wait(&bars.full[stage], phase);
                    if (PROF) pt[0] += clock64() - c0;
                    const uint64_t bd = bdesc0 + (uint32_t)stage * stage_units;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (r < r_eff) {
                            if (PROF) { c0 = clock64(); const long long ti = pt[2] - 480; if (blockIdx.x == 0 && ti >= 0 && ti < 64) tl[0][ti] = c0; }
                            if (n == 0 && !(P.dbg & 16)) mbar_wait(&bars.a_full[r], a_fpar);
                            mbar_wait(&bars.acc_empty[acc], acc_phase ^ 1);
                            tc_fence_after();
                            if (PROF) { pt[0] += clock64() - c0; c0 = clock64(); }
                            issue_tile<8>(tmem_base + (uint32_t)acc * TN, adesc[r], bd, P.ksteps);
                            tc_commit(&bars.acc_full[acc]);
                            if (n == P.NT - 1) tc_commit(&bars.a_empty[r]);       // slot free early
                            if (PROF) { pt[1] += clock64() - c0; ++pt[2];
                                        const long long ti = pt[2] - 1 - 480; if (blockIdx.x == 0 && ti >= 0 && ti < 64) { tl[1][ti] = c0; tl[2][ti] = clock64(); } }
                            acc ^= 1;
                            acc_phase ^= (acc == 0);
                        }
                    }
                    tc_commit(&bars.empty[stage]);
                    if (++stage == P.n_stages) { stage = 0; phase ^= 1; }
                }
                a_fpar ^= 1;
            }
            if (PROF && blockIdx.x == 0) {
                __threadfence();
                for (int i = 0; i < 64; ++i)
                    printf("[tl] tile %d mma_wait_begin %lld issue_begin %lld issue_end %lld | epi_full_seen %lld epi_arrive %lld\n", i,
                           tl[0][i] - tl[0][0], tl[1][i] - tl[0][0], tl[2][i] - tl[0][0], tl[3][i] - tl[0][0], tl[4][i] - tl[0][0]);
            }
            if (PROF && blockIdx.x % 21 == 0 && pt[2] > 0)
                printf("[bmu_tc mma] tiles=%lld cyc/tile total=%lld waits=%lld issue+commit=%lld\n", pt[2],
                       (clock64() - pt_begin) / pt[2], pt[0] / pt[2], pt[1] / pt[2]);
        } else if (P.stage_kb != P.KB) {
            // configs M / L.  The whole warp runs the loop (warp-uniform control flow keeps the operand
            // descriptors in uniform registers -- a lane-0-only branch costs ~3x more per MMA issue); one
            // elected lane issues the tcgen05 instructions.
            const bool leader = elect_one();
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0, a_fpar = 0;
            const uint64_t adesc0 = umma_desc(smem_u32(a_res));
            const uint64_t bdesc0 = umma_desc(smem_u32(ring));
            const uint32_t stage_units = P.stage_bytes >> 4;
            const uint32_t a_in_stage = (uint32_t)B_BLK_BYTES >> 4;      // config L: A' block after B' block
            for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
                for (int n = 0; n < P.NT; ++n) {
                    // one k-block per stage, one patch tile per super-tile
                    if (FUSED && n == 0) mbar_wait(&bars.a_full[0], a_fpar);
                    mbar_wait(&bars.acc_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_addr = tmem_base + (uint32_t)acc * TN;
                    for (int kb = 0; kb < P.KB; ++kb) {
                        mbar_wait(&bars.full[stage], phase);
                        tc_fence_after();
                        const uint64_t bdesc = bdesc0 + (uint32_t)stage * stage_units;
                        const uint64_t ad = FUSED ? adesc0 + (uint32_t)kb * ((uint32_t)A_BLK_BYTES >> 4)
                                                  : bdesc + a_in_stage;
                        const int nk = min(4, P.ksteps - kb * 4);
                        if (leader) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (k < nk) tc_mma_tf32(d_addr, ad + 2u * k, bdesc + 2u * k, (kb | k) != 0);
                            tc_commit(&bars.empty[stage]);
                        }
                        if (++stage == P.n_stages) { stage = 0; phase ^= 1; }
                    }
                    if (leader) {
                        tc_commit(&bars.acc_full[acc]);
                        if (FUSED && n == P.NT - 1) tc_commit(&bars.a_empty[0]);
                    }
                    acc ^= 1;
                    acc_phase ^= (acc == 0);
                }
                a_fpar ^= 1;
            }
        }
    } else if (warp >= BUILD_WARP0 && warp < BUILD_WARP0 + BUILD_WARPS) {
        // ================================ A' builders (+ config S refine) ==============
        if (FUSED && !(P.dbg & 16)) {
            const int t = threadIdx.x - BUILD_WARP0 * 32;       // patch row inside the tile
            const int D = P.g.D;
            const int vec = P.g.vec;
            const uint32_t row_off = (uint32_t)t * 128u;
            const uint32_t sw = (uint32_t)(t & 7);
            uint32_t a_epar = 1;
            int it_ref = 0;                          // super-tiles refined so far by this CTA
            auto st_elem = [&](uint8_t* slot, int kp, float val) {
                uint32_t off = (uint32_t)(kp >> 5) * A_BLK_BYTES + row_off +
                               (((((uint32_t)kp & 31u) >> 2) ^ sw) << 4) + (((uint32_t)kp & 3u) << 2);
                *reinterpret_cast<float*>(slot + off) = val;
            };
            auto st_chunk = [&](uint8_t* slot, int q, float4 val) {
                uint32_t off = (uint32_t)(q >> 3) * A_BLK_BYTES + row_off + ((((uint32_t)q & 7u) ^ sw) << 4);
                *reinterpret_cast<float4*>(slot + off) = val;
            };
            // exact fp32 scores of the CHUNK candidates of one row; strict '>' in ascending unit order
            // keeps the lowest index on ties (same arithmetic as the FFMA variant)
            auto refine_row = [&](int64_t p) {
                float xr[DCAP];
                load_row<DCAP>(xr, P.x + patch_base(P.g, p), true, D, vec, foff_s);
                int u0 = (int)P.out_idx[p];
                if (u0 < 0 || u0 >= P.K) u0 = 0;    // defensive: the epilogue only writes bases in [0, K_pad)
                float best = -INFINITY;
                int bu = u0;
                const bool v4 = ((D & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.W) & 15) == 0);
#pragma unroll
                for (int j = 0; j < CHUNK; ++j) {
                    const int u = u0 + j;
                    const bool ok = u < P.K;
                    const float* wr = P.W + (int64_t)(ok ? u : 0) * D;
                    float accv = ok ? -0.5f * __ldg(P.cn + u) : -INFINITY;
#pragma unroll
                    for (int d = 0; d < DCAP; d += 4) {
                        if (d < D) {
                            if (v4) {
                                const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + d));
                                accv = fmaf(xr[d], w4.x, accv);
                                accv = fmaf(xr[d + 1], w4.y, accv);
                                accv = fmaf(xr[d + 2], w4.z, accv);
                                accv = fmaf(xr[d + 3], w4.w, accv);
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if (d + e < D) accv = fmaf(xr[d + e], __ldg(wr + d + e), accv);
                            }
                        }
                    }
                    if (accv > best) { best = accv; bu = u; }
                }
                P.out_idx[p] = (int64_t)bu + P.unit_offset;
                if (P.out_rd) P.out_rd[p] = -2.0f * best;
            };
            auto refine_super = [&](int st_done) {
                // two alternating barriers: the epilogue can be at most one super-tile ahead of this
                // wait on the same barrier, so a parity can never alias
                mbar_wait(&bars.ref_full[it_ref & 1], (uint32_t)(it_ref >> 1) & 1u);
                ++it_ref;
                const int r_done = min(R, P.n_mtiles - st_done * R);
                for (int r = 0; r < r_done; ++r) {
                    const int64_t p = (int64_t)(st_done * R + r) * TM + t;
                    if (p < P.rows) refine_row(p);
                }
            };

            long long bt[4] = {0, 0, 0, 0};
            float xv[R][DCAP];
            auto prefetch = [&](int st_next) {
                const int r_nxt = (st_next < n_super) ? min(R, P.n_mtiles - st_next * R) : 0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int64_t p = (int64_t)(st_next * R + r) * TM + t;
                    const bool ok = (r < r_nxt) && (p < P.rows);
                    load_row<DCAP>(xv[r], P.x + (ok ? patch_base(P.g, p) : 0), ok, D, vec, foff_s);
                }
            };
            prefetch(blockIdx.x);
            int st_prev = -1;
            for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
                const int r_eff = min(R, P.n_mtiles - st * R);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (r < r_eff) {
                        long long b0 = PROF ? clock64() : 0;
                        mbar_wait(&bars.a_empty[r], a_epar);
                        if (PROF) { bt[0] += clock64() - b0; b0 = clock64(); }
                        uint8_t* slot = a_res + (size_t)r * P.KB * A_BLK_BYTES;
                        if ((P.dbg & 2) && st != (int)blockIdx.x) {
                        } else if ((D & 3) == 0) {
                            // 16-byte chunks: chunk q of the row lands at ((q & 7) ^ (t & 7)) inside its
                            // 128-byte swizzle row -- 8 consecutive rows fill one conflict-free wavefront
                            const int dq = D >> 2;
#pragma unroll
                            for (int d4 = 0; d4 < DCAP / 4; ++d4) {
                                if (d4 < dq) {
                                    float4 hi, lo;
                                    hi.x = tf32_rna(xv[r][4 * d4]);     lo.x = tf32_rna(xv[r][4 * d4] - hi.x);
                                    hi.y = tf32_rna(xv[r][4 * d4 + 1]); lo.y = tf32_rna(xv[r][4 * d4 + 1] - hi.y);
                                    hi.z = tf32_rna(xv[r][4 * d4 + 2]); lo.z = tf32_rna(xv[r][4 * d4 + 2] - hi.z);
                                    hi.w = tf32_rna(xv[r][4 * d4 + 3]); lo.w = tf32_rna(xv[r][4 * d4 + 3] - hi.w);
                                    st_chunk(slot, d4, hi);
                                    st_chunk(slot, dq + d4, lo);
                                    st_chunk(slot, 2 * dq + d4, hi);
                                }
                            }
                            st_chunk(slot, 3 * dq, make_float4(1.f, 1.f, 1.f, 0.f));
                            for (int q = 3 * dq + 1; q < (P.KP >> 2); ++q) st_chunk(slot, q, make_float4(0.f, 0.f, 0.f, 0.f));
                        } else {
#pragma unroll
                            for (int d = 0; d < DCAP; ++d) {
                                if (d < D) {
                                    const float v = xv[r][d];
                                    const float hi = tf32_rna(v);
                                    const float lo = tf32_rna(v - hi);
                                    st_elem(slot, d, hi);
                                    st_elem(slot, D + d, lo);
                                    st_elem(slot, 2 * D + d, hi);
                                }
                            }
                            for (int kp = 3 * D; kp < P.KP; ++kp) st_elem(slot, kp, kp < 3 * D + 3 ? 1.0f : 0.f);
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars.a_full[r]);
                        if (PROF) { bt[1] += clock64() - b0; ++bt[3]; }
                    }
                }
                a_epar ^= 1;
                if (!(P.dbg & 8)) prefetch(st + gridDim.x);            // rows of the next super-tile, a full tile time ahead
                long long b1 = PROF ? clock64() : 0;
                if (!EXACT && st_prev >= 0 && !(P.dbg & 1)) refine_super(st_prev);
                if (PROF) bt[2] += clock64() - b1;
                st_prev = st;
            }
            if (!EXACT && st_prev >= 0 && !(P.dbg & 1)) refine_super(st_prev);
            if (PROF && blockIdx.x == 0 && t == 0 && bt[3] > 0)
                printf("[bmu_tc bld ] slots=%lld cyc/slot wait_a_empty=%lld write=%lld refine(per super-tile incl wait)=%lld\n",
                       bt[3], bt[0] / bt[3], bt[1] / bt[3], bt[2] * R / bt[3]);
        }
    } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + EPI_THREADS / 32) {
        // ================================ epilogue ====================================
        const int ew = warp - EPI_WARP0;            // 0..7
        const int half = ew >> 2;                   // column half of the accumulator
        const int lg = warp & 3;                    // TMEM lane quarter this warp may access
        const int row = lg * 32 + lane;             // patch row inside the tile
        int acc = 0, it_epi = 0;
        uint32_t acc_phase = 0;
        long long et[3] = {0, 0, 0};
        for (int st = blockIdx.x; st < n_super; st += gridDim.x, ++it_epi) {
            const int r_eff = min(R, P.n_mtiles - st * R);
            float best[R];
            int bidx[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { best[r] = INFINITY; bidx[r] = 0; }
            for (int n = 0; n < P.NT; ++n) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (r < r_eff) {
                        long long e0 = PROF ? clock64() : 0;
                        mbar_wait(&bars.acc_full[acc], acc_phase);
                        tc_fence_after();
                        if (PROF) { et[0] += clock64() - e0; e0 = clock64();
                                    const long long ti = et[2] - 480; if (blockIdx.x == 0 && threadIdx.x == EPI_WARP0 * 32 && ti >= 0 && ti < 64) tl[3][ti] = e0; }
                        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * TN + half * 128);
                        const int col0 = n * TN + half * 128;
                        uint32_t va[32], vb[32];
                        auto consume = [&](const uint32_t (&v)[32], int c) {
                            if (EXACT) {
                                const float m = min32(v);
                                if (m < best[r]) { best[r] = m; bidx[r] = col0 + c * 32 + first_eq32(v, m); }
                            } else {
                                // minimum per CHUNK(8)-column group; the group index is all we keep
                                float q[4];
#pragma unroll
                                for (int g = 0; g < 4; ++g) {
                                    float m8 = __uint_as_float(v[g * 8]);
#pragma unroll
                                    for (int i = 1; i < 8; ++i) m8 = fminf(m8, __uint_as_float(v[g * 8 + i]));
                                    q[g] = m8;
                                }
                                const float m = fminf(fminf(q[0], q[1]), fminf(q[2], q[3]));
                                if (m < best[r]) {
                                    best[r] = m;
                                    const int sub = (q[0] == m) ? 0 : (q[1] == m) ? 1 : (q[2] == m) ? 2 : 3;
                                    bidx[r] = col0 + c * 32 + sub * CHUNK;
                                }
                            }
                        };
                        tmem_ld32_issue(taddr, va);
                        tmem_ld_wait(va);
                        tmem_ld32_issue(taddr + 32, vb);
                        consume(va, 0);
                        tmem_ld_wait(vb);
                        tmem_ld32_issue(taddr + 64, va);
                        consume(vb, 1);
                        tmem_ld_wait(va);
                        tmem_ld32_issue(taddr + 96, vb);
                        consume(va, 2);
                        tmem_ld_wait(vb);
                        // accumulator fully read: hand it back before the last reduction
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars.acc_empty[acc]);
                        if (PROF) { const long long ti = et[2] - 480; if (blockIdx.x == 0 && threadIdx.x == EPI_WARP0 * 32 && ti >= 0 && ti < 64) tl[4][ti] = clock64(); }
                        consume(vb, 3);
                        if (PROF) { et[1] += clock64() - e0; ++et[2]; }
                        acc ^= 1;
                        acc_phase ^= (acc == 0);
                    }
                }
            }
            // merge the two column halves (value asc, index asc) and store, one patch tile at a time
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r < r_eff && !(P.dbg & 32)) {
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
                            // config S leaves the chunk base here; the builders finish the index
                            P.out_idx[p] = (int64_t)bi + (EXACT ? P.unit_offset : 0);
                            if (EXACT && P.out_rd) P.out_rd[p] = bv;
                        }
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
            }
            if (!EXACT && half == 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.ref_full[it_epi & 1]);   // release: chunk bases visible
            }
        }
        if (PROF && blockIdx.x == 0 && threadIdx.x == EPI_WARP0 * 32 && et[2] > 0)
            printf("[bmu_tc epi ] tiles=%lld cyc/tile wait_acc_full=%lld process=%lld\n", et[2], et[0] / et[2],
                   et[1] / et[2]);
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

// config L: A'[p] = [ hi(x) | lo(x) | hi(x) | 1 1 1 | 0.. ] with patchify fused as address arithmetic.
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
enum Config { CFG_S = 0, CFG_M = 1, CFG_L = 2 };

struct Plan {
    int cfg;
    int D, K, KP, KB, ksteps, K_pad, NT;
    int R, a_resident, stage_kb, n_stages;
    uint32_t a_bytes, stage_bytes, smem_bytes;
    int64_t chunk_rows;            // config L: patches per workspace chunk (multiple of 128)
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
    const uint32_t budget = SMEM_LIMIT - 1024 /*alignment slack*/ - STATIC_SMEM;
    pl->chunk_rows = 0;
    if (pl->KB <= 2 && D <= DCAP_S) {
        pl->cfg = CFG_S;
        pl->a_resident = 1; pl->stage_kb = pl->KB; pl->n_stages = 2; pl->R = MAX_R;
        { const char* e = getenv("SOM_TC_R1"); if (e && e[0] == '1') pl->R = 1; }
        pl->stage_bytes = (uint32_t)pl->KB * B_BLK_BYTES;
        pl->a_bytes = (uint32_t)pl->R * pl->KB * A_BLK_BYTES;
    } else if (pl->KB <= 7 && D <= DCAP_M) {
        pl->cfg = CFG_M;
        pl->a_resident = 1; pl->stage_kb = 1; pl->R = 1;
        pl->stage_bytes = B_BLK_BYTES;
        pl->a_bytes = (uint32_t)pl->KB * A_BLK_BYTES;
        pl->n_stages = (int)((budget - pl->a_bytes) / pl->stage_bytes);
        if (pl->n_stages > MAX_STAGES) pl->n_stages = MAX_STAGES;
    } else {
        pl->cfg = CFG_L;
        pl->a_resident = 0; pl->stage_kb = 1; pl->R = 1;
        pl->stage_bytes = B_BLK_BYTES + A_BLK_BYTES;
        pl->a_bytes = 0;
        pl->n_stages = MAX_STAGES;
        // workspace chunk: keep A' within ~48 MB so it stays L2 resident between the two kernels
        const int64_t wave_rows = (int64_t)TM * sm_count();
        int64_t max_rows = (48ll << 20) / ((int64_t)pl->KP * 4);
        int64_t chunk = max_rows / wave_rows * wave_rows;
        if (chunk < TM) chunk = (max_rows / TM > 0 ? max_rows / TM : 1) * TM;
        int64_t need = ceil_div64(n, TM) * TM;
        if (chunk > need) chunk = need;
        pl->chunk_rows = chunk;
    }
    pl->smem_bytes = pl->a_bytes + (uint32_t)pl->n_stages * pl->stage_bytes + 1024;
    size_t o = 0;
    pl->off_b = o; o = align_up(o + (size_t)pl->K_pad * pl->KP * 4, 1024);
    pl->off_a = o; o = align_up(o + (size_t)pl->chunk_rows * pl->KP * 4, 1024);
    pl->total = o;
}

template <int R, bool EXACT, bool FUSED, int DCAP, bool PROF = false>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const Params& P, uint32_t smem, int grid,
                       cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(bmu_tc_kernel<R, EXACT, FUSED, DCAP, PROF>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(SMEM_LIMIT - STATIC_SMEM));
        if (e != cudaSuccess) { set_error("bmu(tc): smem opt-in: %s", cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    bmu_tc_kernel<R, EXACT, FUSED, DCAP, PROF><<<grid, NUM_THREADS, smem, st>>>(ma, mb, P);
    return check_launch("bmu_tc_kernel");
}

}  // namespace tc

// som_bmu_tc_s.cu: config S (D <= 16)
bool tc_s_applicable(int D);
size_t tc_s_workspace_bytes(int64_t n_patches, int D, int K);
int launch_bmu_tc_s(const float* x, const Geom& g, const float* W, const float* cn, int K, int64_t unit_offset,
                    int64_t* out_idx, float* out_rd, void* ws, size_t ws_bytes, cudaStream_t st);
// som_bmu_tc_l.cu: config L (streamed operands, fused builders, optional split-K)
size_t tc_l_workspace_bytes(int64_t n_patches, int D, int K);
int launch_bmu_tc_l(const float* x, const Geom& g, const float* W, const float* cn, int K, int64_t unit_offset,
                    int64_t* out_idx, float* out_rd, void* ws, size_t ws_bytes, cudaStream_t st);
int tc_l_splits(int64_t n_patches, int D, int K);
static bool use_l2(int64_t n_patches, int D, int K) {
    static int mode = -1;           // SOM_TC_L=0: never, 2: always (A/B comparisons only); default: split-K shapes
    if (mode < 0) { const char* e = getenv("SOM_TC_L"); mode = e ? atoi(e) : 1; }
    if (mode == 0) return false;
    if (D <= 64) return true;       // resident-A mode of the fused-builder kernel (config M, D in 17..64)
    const int kb = (3 * D + 3 + tc::KBLK - 1) / tc::KBLK;
    if (kb <= 7 && D <= tc::DCAP_M) return false;
    return mode == 2 || tc_l_splits(n_patches, D, K) > 1;
}
static bool use_s4(int D) {
    static int old = -1;            // SOM_TC_OLD_S=1: previous config-S kernel (A/B comparisons only)
    if (old < 0) { const char* e = getenv("SOM_TC_OLD_S"); old = (e && e[0] == '1') ? 1 : 0; }
    return !old && tc_s_applicable(D);
}

bool tc_supported(int64_t n_patches, int D, int K) {
    if (n_patches <= 0 || D <= 0 || K <= 0) return false;
    if (n_patches / tc::TM >= (1ll << 30)) return false;
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
    if (use_s4(D)) return tc_s_workspace_bytes(n_patches, D, K);
    if (use_l2(n_patches, D, K)) return tc_l_workspace_bytes(n_patches, D, K);
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
    if (use_s4(g.D)) return launch_bmu_tc_s(x, g, W, cn, K, unit_offset, out_idx, out_rd, ws, ws_bytes, st);
    if (use_l2(n, g.D, K)) return launch_bmu_tc_l(x, g, W, cn, K, unit_offset, out_idx, out_rd, ws, ws_bytes, st);
    Plan pl;
    make_plan(&pl, n, g.D, K);
    SOM_REQUIRE(ws != nullptr && ws_bytes >= pl.total, SOM_E_WORKSPACE,
                "bmu(tc): workspace %zu < required %zu", ws_bytes, pl.total);
    SOM_REQUIRE(((uintptr_t)ws & 255) == 0, SOM_E_BADARG, "bmu(tc): workspace must be 256-byte aligned");
    SOM_REQUIRE(pl.smem_bytes + STATIC_SMEM <= SMEM_LIMIT, SOM_E_SHAPE, "bmu(tc): shared memory plan too large");
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

    Params P;
    P.KB = pl.KB; P.KP = pl.KP; P.ksteps = pl.ksteps; P.NT = pl.NT; P.a_resident = pl.a_resident;
    P.stage_kb = pl.stage_kb; P.n_stages = pl.n_stages; P.unit_offset = unit_offset;
    P.a_bytes = pl.a_bytes; P.stage_bytes = pl.stage_bytes;
    P.x = x; P.g = g; P.W = W; P.cn = cn; P.K = K;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("SOM_TC_DEBUG"); dbg = e ? atoi(e) : 0; } P.dbg = dbg; }

    if (pl.cfg != CFG_L) {
        // fused builders: one launch over all patches, no operand round trip through memory
        map_a = map_b;                               // unused by the kernel
        P.rows = n;
        P.n_mtiles = (int)ceil_div64(n, TM);
        P.out_idx = out_idx;
        P.out_rd = out_rd;
        const int n_super = (P.n_mtiles + pl.R - 1) / pl.R;
        const int grid = n_super < sm_count() ? n_super : sm_count();
        if (pl.cfg == CFG_S) {
            static int prof = -1;       // SOM_TC_PROFILE=1: per-role cycle breakdown printed by CTA 0
            if (prof < 0) { const char* e = getenv("SOM_TC_PROFILE"); prof = (e && e[0] == '1') ? 1 : 0; }
            if (pl.R == 1) return launch_gemm<1, false, true, DCAP_S, true>(map_a, map_b, P, pl.smem_bytes, grid, st);
            if (prof) return launch_gemm<MAX_R, false, true, DCAP_S, true>(map_a, map_b, P, pl.smem_bytes, grid, st);
            return launch_gemm<MAX_R, false, true, DCAP_S>(map_a, map_b, P, pl.smem_bytes, grid, st);
        }
        return launch_gemm<1, true, true, DCAP_M>(map_a, map_b, P, pl.smem_bytes, grid, st);
    }

    rc = make_map(&map_a, Ap, (uint64_t)pl.chunk_rows, (uint64_t)pl.KP, TM);
    if (rc) return rc;
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
        const int grid = P.n_mtiles < sm_count() ? P.n_mtiles : sm_count();
        rc = launch_gemm<1, true, false, 4>(map_a, map_b, P, pl.smem_bytes, grid, st);
        if (rc) return rc;
    }
    return SOM_OK;
}

}  // namespace som
