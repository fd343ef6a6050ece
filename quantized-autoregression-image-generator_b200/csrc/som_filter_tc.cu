// K3 (tensor-core variant): neighbourhood filter  out = scale * T @ in  as a banded-Toeplitz GEMM on tcgen05.
// sm_100a only.
//
//   out[a][d] = scale * sum_{t=-h..h} w(|t|) * in[a+t][d],   w(t) = expf(-( float(t*t) / two_var ))
//
// T is the K x K Gaussian neighbourhood matrix the reference builds row by row (models/Codebook.py:112-125) and
// multiplies densely (`:128-130` forward, autograd backward); h is where expf underflows to exactly 0.
// For a tile of 128 units [a0, a0+128) the product is a dense GEMM  M = 128 units, N = features,
// inner = the 128 + 2h input rows [a0-h, a0+127+h]:  A[ul][jl] = w(jl - h - ul) is the SAME Toeplitz strip for every
// tile, B[d][jl] = in[a0-h+jl][d].  fp32-faithful 3xTF32 (hi.hi + lo.hi + hi.lo, fp32 accumulation in TMEM), like the
// BMU kernels.  The FFMA kernel (som_filter.cu) reaches 26-29 TFLOP/s of 2*K*D*band useful flop (C4 96 us, C3 36-43 us).
//
//   pre-pass  split_in_t_kernel : in (K x D) -> transposed, zero-padded, hi | lo split operand  Bt[d][h + j]
//                                 (K-major for the MMA: the reduction index must be the contiguous one)
//   main      filter_tc_kernel  : one CTA per (unit tile, feature tile of TNF), 320 threads, warp-specialised:
//     warp 0     TMA producer: B_hi / B_lo blocks (TNF x 32 floats each) of k-block kb into a 3-stage ring
//     warp 1     MMA issuer  : 4 k-steps x 3 products of M128 x N(TNF) x K8 kind::tf32 per k-block
//     warps 2-9  builders    : write the Toeplitz A_hi / A_lo block of k-block kb (128 x 32, SWIZZLE_128B rows) from
//                              a shared weight table -- the strip never travels through L2 (every CTA would read the
//                              same bytes at the same time: measured 14.5 B/clk/SM for such hot tiles);
//                              afterwards warps 2-5 are the epilogue: tcgen05.ld of their TMEM lane quarter,
//                              scale, 128-byte row stores.
// Bound: tensor pipe / shared-memory operand bandwidth (A 4 KB + B TNF*32 B per MMA).
// Algorithmic work: 2*K*D*(2h+1) flop; executed: 3 * 2*128*TNF*L per tile with L = 32*ceil((128+2h)/32).
#include "som_common.cuh"
#include "som_peer.cuh"
#include "som_tc_ptx.cuh"

namespace som {
SOM_TRACE_TU(trace_set_filter_tc)
namespace ftc {
using namespace tc;

constexpr int TMU = 128;                     // units per tile (UMMA M)
constexpr int NSTAGE = 3;
constexpr int NACC = 4;                      // TMEM accumulators, k-block kb goes to kb % NACC (see the MMA loop)
constexpr int BUILD_WARPS = 8;                // two threads per unit row, 16 inner positions each
constexpr int NUM_THREADS = (2 + BUILD_WARPS) * 32;      // 320
constexpr int TAB_PAD = 160;                 // |jl - h - ul| <= h + 158
constexpr int MAX_H = 1400;                  // table of 2 * (2h + 2*TAB_PAD + 1) floats <= 25 KB

struct Params {
    int K, D, h, nkb;
    float two_var, scale;
    float* out;
    int ks;                 // k-split: CTA blockIdx.z takes the k-blocks kb = z (mod ks); ks is 1 or NACC
    float* partial;         // ks > 1: [ks][K][D] unscaled partial sums, added by filter_reduce_kernel
};

struct __align__(8) Barriers {
    uint64_t full[NSTAGE], empty[NSTAGE], acc_full;
    uint32_t tmem_base, pad;
};

template <int TNF> constexpr int stage_bytes() { return 2 * A_BLK_BYTES + 2 * TNF * KBLK * 4; }
static inline size_t smem_bytes(int tnf, int h) {
    return 1024 + (size_t)NSTAGE * (2 * A_BLK_BYTES + 2 * tnf * KBLK * 4) + sizeof(Barriers) +
           2 * (size_t)(2 * (h + TAB_PAD) + 1) * sizeof(float) + 16;
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int TNF>
__device__ __forceinline__ void mma_tf32_n(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accum) {
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TNF >> 3) << 17) |
                               ((uint32_t)(TMU >> 4) << 24);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}

// Bt_hi / Bt_lo [D][Kp]: column h + j holds in[j][d] (hi / lo part), zero elsewhere.  32 x 32 tiles through shared
// memory: reads coalesced along d, writes coalesced along j.
// MC: `in` is the NVSwitch multicast address of the ranks' accumulator rows (som_peer.cu) and the load is the in-switch
// reduction multimem.ld_reduce.add -- the reduce-scatter half of the data-parallel tail happens in this pre-pass's read,
// with no reduced copy of the rows in between; block (0, 0) also reduces the packed buffer's 4-float tail exactly.
struct PeerIn {
    peer::Pads bufs;        // every rank's packed buffer (peer addresses), for the exact tail
    int64_t q_tail;         // float4 index of the tail in a packed buffer
    int world;
    float4* tail_out;
};

template <bool MC>
__global__ void __launch_bounds__(256) split_in_t_kernel(const float* __restrict__ in, int K, int D, int h, int Kp,
                                                         float* __restrict__ Bhi, float* __restrict__ Blo, PeerIn pin) {
    pdl_begin();
    trace_stamp(s_trace_buf, 1);
    __shared__ float tile[32][33];
    const int jp0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    if (MC) {
        const int rr = threadIdx.x >> 3, c4 = (threadIdx.x & 7) * 4;         // 32 rows x 8 quads (D % 4 == 0)
        const int j = jp0 + rr - h, d = d0 + c4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j >= 0 && j < K && d < D) v = peer::mm_ld_reduce(in + (int64_t)j * D + d);
        tile[rr][c4] = v.x; tile[rr][c4 + 1] = v.y; tile[rr][c4 + 2] = v.z; tile[rr][c4 + 3] = v.w;
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
            *pin.tail_out = peer::exact_tail(pin.bufs, pin.q_tail, pin.world);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = jp0 + ty + 8 * i - h, d = d0 + tx;
            tile[ty + 8 * i][tx] = (j >= 0 && j < K && d < D) ? __ldg(in + (int64_t)j * D + d) : 0.f;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int d = d0 + ty + 8 * i, jp = jp0 + tx;
        if (d < D && jp < Kp) {
            const float v = tile[tx][ty + 8 * i];
            const float hi = tf32_rna(v);
            Bhi[(int64_t)d * Kp + jp] = hi;
            Blo[(int64_t)d * Kp + jp] = tf32_rna(v - hi);
        }
    }
}

template <int TNF>
__global__ void __launch_bounds__(NUM_THREADS, 1)
filter_tc_kernel(const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                 const Params P) {
    pdl_launch_dependents();            // (the wait comes after the barrier / weight-table / TMEM prologue)
    constexpr int STAGE = stage_bytes<TNF>();
    constexpr int B_BYTES = TNF * KBLK * 4;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment as an OFFSET into the shared array: rounding the pointer through uintptr_t made the compiler
    // forget the address space, and every access below compiled to generic LD.E / ST.E (cuobjdump, round 2)
    uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    Barriers& bars = *reinterpret_cast<Barriers*>(ring + NSTAGE * STAGE);
    float* tab_hi = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(&bars) + sizeof(Barriers));
    const int off = P.h + TAB_PAD;
    const int tab_n = 2 * off + 1;
    float* tab_lo = tab_hi + tab_n;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ut = blockIdx.x, ct = blockIdx.y;
    // k-split (few tiles, e.g. a data-parallel rank's slice of units): the NACC round-robin accumulators of the unsplit
    // kernel become NACC CTAs -- CTA z runs exactly the accumulation chain of accumulator z, and filter_reduce_kernel
    // adds the partial tiles in the same fixed order, so the result is bit-identical to the unsplit kernel's
    const int ksp = (int)blockIdx.z, KS = P.ks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&bars.full[s], 1 + BUILD_WARPS); mbar_init(&bars.empty[s], 1); }
        mbar_init(&bars.acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // weight table, generated exactly as the FFMA kernel / the reference do: (j-b)^2 exact, fp32 divide, expf
    for (int i = threadIdx.x; i < tab_n; i += NUM_THREADS) {
        const int t = i - off;
        const int at = t < 0 ? -t : t;
        float w = 0.f;
        if (at <= P.h) {
            const float sq = (float)((long long)at * (long long)at);
            w = expf(-(__fdiv_rn(sq, P.two_var)));
        }
        const float hi = tf32_rna(w);
        tab_hi[i] = hi;
        tab_lo[i] = tf32_rna(w - hi);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_base)),
                     "r"(NACC * TNF));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars.tmem_base;
    const int nkb = P.nkb;
    pdl_wait();
    trace_stamp(s_trace_buf, 2);

    if (warp == 0) {
        // ================================ TMA producer (B blocks) ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = ksp; kb < nkb; kb += KS) {
                mbar_wait(&bars.empty[stage], phase ^ 1);
                uint8_t* sb = ring + (size_t)stage * STAGE + 2 * A_BLK_BYTES;
                mbar_expect_tx(&bars.full[stage], 2 * B_BYTES);
                // padded inner coordinate of unit tile ut starts at ut * 128 (= a0 - h + h)
                tma_load_2d(&map_bhi, &bars.full[stage], sb, ut * TMU + kb * KBLK, ct * TNF);
                tma_load_2d(&map_blo, &bars.full[stage], sb + B_BYTES, ut * TMU + kb * KBLK, ct * TNF);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        const bool leader = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = ksp; kb < nkb; kb += KS) {
            mbar_wait(&bars.full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(ring + (size_t)stage * STAGE);
            const uint64_t ahi = umma_desc(sa), alo = umma_desc(sa + A_BLK_BYTES);
            const uint64_t bhi = umma_desc(sa + 2 * A_BLK_BYTES), blo = umma_desc(sa + 2 * A_BLK_BYTES + B_BYTES);
            // The tensor core adds into its fp32 accumulator with truncation, a bias that grows with the length of
            // the accumulation chain (measured 1.7e-6 relative after 372 MMAs into one accumulator, band 851): the
            // k-blocks are dealt round-robin to NACC accumulators and the epilogue adds those in fp32 (4e-7).
            const uint32_t d_addr = tmem_base + (KS > 1 ? 0u : (uint32_t)(kb % NACC) * TNF);
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < KBLK / 8; ++ks) {
                    mma_tf32_n<TNF>(d_addr, ahi + 2u * ks, bhi + 2u * ks, (kb >= NACC || ks != 0) ? 1u : 0u);
                    mma_tf32_n<TNF>(d_addr, alo + 2u * ks, bhi + 2u * ks, 1u);
                    mma_tf32_n<TNF>(d_addr, ahi + 2u * ks, blo + 2u * ks, 1u);
                }
                tc_commit(&bars.empty[stage]);
                if (kb + KS >= nkb) tc_commit(&bars.acc_full);
            }
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
    } else {
        // ================================ A builders, then epilogue =====================
        {
            const int bt = threadIdx.x - 64;
            const int ul = bt & (TMU - 1);                      // unit row of the tile this thread builds
            const int qh = (bt >> 7) * (KBLK / 8);              // its half of the row's eight 16-byte chunks
            const uint32_t row_off = (uint32_t)ul * 128u;
            const uint32_t sw = (uint32_t)(ul & 7);
            int stage = 0;
            uint32_t phase = 0;
            // explicit shared-space accesses through 32-bit addresses: written with pointers, the table reads and the tile
            // stores compiled to GENERIC LD.E / ST.E.128 with 64-bit address arithmetic (ncu source page, round 2)
            const uint32_t ring_u = smem_u32(ring), tab_hi_u = smem_u32(tab_hi), tab_lo_u = smem_u32(tab_lo);
            for (int kb = ksp; kb < nkb; kb += KS) {
                mbar_wait_warp<false>(&bars.empty[stage], phase ^ 1, lane);
                const uint32_t sa = ring_u + (uint32_t)stage * (uint32_t)STAGE + row_off;
                // A[ul][jl] = w(jl - h - ul), jl = 32 kb + c: consecutive rows read consecutive table entries
                const uint32_t t_off = (uint32_t)(kb * KBLK - P.h - ul + off) * 4u;
                const uint32_t th = tab_hi_u + t_off, tl = tab_lo_u + t_off;
#pragma unroll
                for (int qq = 0; qq < KBLK / 8; ++qq) {
                    const int q = qh + qq;
                    float4 hi, lo;
                    hi.x = lds_f32(th + 16u * q); hi.y = lds_f32(th + 16u * q + 4u);
                    hi.z = lds_f32(th + 16u * q + 8u); hi.w = lds_f32(th + 16u * q + 12u);
                    lo.x = lds_f32(tl + 16u * q); lo.y = lds_f32(tl + 16u * q + 4u);
                    lo.z = lds_f32(tl + 16u * q + 8u); lo.w = lds_f32(tl + 16u * q + 12u);
                    sts_v4(sa + ((((uint32_t)q) ^ sw) << 4), hi);
                    sts_v4(sa + (uint32_t)A_BLK_BYTES + ((((uint32_t)q) ^ sw) << 4), lo);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.full[stage]);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
        if (warp >= 6) goto done;                               // warps 2-5 own the four TMEM lane quarters
        // epilogue: TMEM lane quarter (warp & 3), 32 columns at a time
        mbar_wait_warp<true>(&bars.acc_full, 0, lane);
        tc_fence_after();
        const int lg = warp & 3;
        const int u = ut * TMU + lg * 32 + lane;
        const bool row_ok = u < P.K;
        const bool v4 = ((P.D & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.out) & 15) == 0);
        float* orow = (KS > 1 ? P.partial + (int64_t)ksp * P.K * P.D : P.out) + (int64_t)(row_ok ? u : 0) * P.D;
        const float oscale = KS > 1 ? 1.0f : P.scale;            // split: raw partial sums, scaled after the reduction
#pragma unroll 1
        for (int c = 0; c < TNF / 32; ++c) {
            uint32_t v[32], w[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(c * 32);
            tmem_ld32_issue(taddr, v);
            tmem_ld_wait(v);
            if (KS == 1) {
#pragma unroll
                for (int a = 1; a < NACC; ++a) {                 // fixed order: ((a0 + a1) + a2) + a3
                    tmem_ld32_issue(taddr + (uint32_t)(a * TNF), w);
                    tmem_ld_wait(w);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
                }
            }
            const int d0 = ct * TNF + c * 32;
            if (row_ok && d0 < P.D) {
                if (v4 && d0 + 32 <= P.D) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<float4*>(orow + d0 + 4 * q) =
                            make_float4(oscale * __uint_as_float(v[4 * q]), oscale * __uint_as_float(v[4 * q + 1]),
                                        oscale * __uint_as_float(v[4 * q + 2]), oscale * __uint_as_float(v[4 * q + 3]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (d0 + i < P.D) orow[d0 + i] = oscale * __uint_as_float(v[i]);
                }
            }
        }
    }

done:
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NACC * TNF));
    }
}

// out = scale * (((p0 + p1) + p2) + p3): the unsplit epilogue's order
__global__ void __launch_bounds__(256) filter_reduce_kernel(const float* __restrict__ partial, int64_t n, float scale,
                                                            float* __restrict__ out) {
    pdl_begin();
    trace_stamp(s_trace_buf, 3);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        float v = partial[i];
#pragma unroll
        for (int a = 1; a < NACC; ++a) v += partial[(int64_t)a * n + i];
        out[i] = scale * v;
    }
}

struct Plan {
    int h, L, nkb, Kp, tnf, ks;
    size_t off_hi, off_lo, off_part, total;
};

static void make_plan(Plan* pl, int K, int D, int h) {
    pl->h = h;
    pl->L = (TMU + 2 * h + KBLK - 1) / KBLK * KBLK;
    pl->nkb = pl->L / KBLK;                             // >= 5 >= NACC: every accumulator is written
    const int n_ut = (K + TMU - 1) / TMU;
    pl->Kp = (n_ut - 1) * TMU + pl->L;                  // multiple of 32: 128-byte row pitch granularity for TMA
    pl->tnf = D > 64 ? 128 : 64;
    const size_t one = align_up((size_t)D * pl->Kp * sizeof(float), 1024);
    pl->off_hi = 0;
    pl->off_lo = one;
    // static rule: with fewer tiles than half the SMs, the NACC accumulation chains of a tile go to NACC CTAs
    const int64_t tiles = (int64_t)n_ut * ((D + pl->tnf - 1) / pl->tnf);
    pl->ks = (tiles * 2 <= sm_count()) ? NACC : 1;
    pl->off_part = 2 * one;
    pl->total = 2 * one + (pl->ks > 1 ? align_up((size_t)pl->ks * K * D * sizeof(float), 1024) : 0);
}

}  // namespace ftc

// largest t with expf(-(float(t*t)/two_var)) > 0 (som_filter.cu)
int filter_band_half_width(float two_var, int K);

// Static rule on the shape: the tensor-core filter pays a transposing pre-pass and 128-unit tiles.
bool filter_tc_applicable(int K, int D, int h) {
    // ... and one CTA per tile: with fewer than 32 tiles most SMs idle and the FFMA kernel (32-unit tiles) is faster
    // (measured at C1, K = 1024, D = 64: 8 tiles, 28 us against 21 us)
    // (fewer than 74 tiles are k-split over 4 CTAs each; below ~12 tiles the FFMA kernel wins again)
    const int64_t tiles = ceil_div64(K, ftc::TMU) * ceil_div64(D, D > 64 ? 128 : 64);
    if (K < 256 || D < 48 || h > ftc::MAX_H || tiles < 12) return false;
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

size_t filter_tc_workspace_bytes(int K, int D, int h) {
    ftc::Plan pl;
    ftc::make_plan(&pl, K, D, h);
    return pl.total;
}

// pin != nullptr: `in` is a multicast address, read with the in-switch reduction (see split_in_t_kernel)
int launch_filter_tc_in(const float* in, float* out, int K, int D, float two_var, int h, float scale, void* ws,
                        size_t ws_bytes, cudaStream_t st, const ftc::PeerIn* pin) {
    using namespace ftc;
    Plan pl;
    make_plan(&pl, K, D, h);
    SOM_REQUIRE(ws != nullptr && ws_bytes >= pl.total, SOM_E_WORKSPACE, "filter(tc): workspace %zu < required %zu",
                ws_bytes, pl.total);
    SOM_REQUIRE(((uintptr_t)ws & 255) == 0, SOM_E_BADARG, "filter(tc): workspace must be 256-byte aligned");
    float* Bhi = (float*)((char*)ws + pl.off_hi);
    float* Blo = (float*)((char*)ws + pl.off_lo);
    {
        dim3 grid((unsigned)(pl.Kp / 32), (unsigned)ceil_div64(D, 32));
        if (pin != nullptr) launch_pdl(split_in_t_kernel<true>, grid, 256, 0, st, in, K, D, h, pl.Kp, Bhi, Blo, *pin);
        else launch_pdl(split_in_t_kernel<false>, grid, 256, 0, st, in, K, D, h, pl.Kp, Bhi, Blo, PeerIn{});
        int rc = check_launch("split_in_t_kernel");
        if (rc) return rc;
    }
    CUtensorMap map_bhi, map_blo;
    int rc = make_map2d(&map_bhi, Bhi, (uint64_t)D, (uint64_t)pl.Kp, (uint64_t)pl.Kp * 4, KBLK, (uint32_t)pl.tnf,
                        CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map2d(&map_blo, Blo, (uint64_t)D, (uint64_t)pl.Kp, (uint64_t)pl.Kp * 4, KBLK, (uint32_t)pl.tnf,
                    CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    Params P;
    P.K = K; P.D = D; P.h = h; P.nkb = pl.nkb; P.two_var = two_var; P.scale = scale; P.out = out;
    P.ks = pl.ks; P.partial = (float*)((char*)ws + pl.off_part);
    const size_t smem = smem_bytes(pl.tnf, h);
    static PerDeviceFlag attr_done;
    if (attr_done.pending()) {
        cudaError_t e = cudaFuncSetAttribute(filter_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(filter_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("filter(tc): smem opt-in: %s", cudaGetErrorString(e)); return (int)e; }
        attr_done.set();
    }
    dim3 grid((unsigned)ceil_div64(K, TMU), (unsigned)ceil_div64(D, pl.tnf), (unsigned)pl.ks);
    if (pl.tnf == 128) launch_pdl(filter_tc_kernel<128>, grid, NUM_THREADS, smem, st, map_bhi, map_blo, P);
    else launch_pdl(filter_tc_kernel<64>, grid, NUM_THREADS, smem, st, map_bhi, map_blo, P);
    rc = check_launch("filter_tc_kernel");
    if (rc || pl.ks == 1) return rc;
    const int64_t n = (int64_t)K * D;
    int blocks = (int)ceil_div64(n, 256 * 4);
    if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
    launch_pdl(filter_reduce_kernel, blocks, 256, 0, st, P.partial, n, scale, out);
    return check_launch("filter_reduce_kernel");
}

int launch_filter_tc(const float* in, float* out, int K, int D, float two_var, int h, float scale, void* ws,
                     size_t ws_bytes, cudaStream_t st) {
    return launch_filter_tc_in(in, out, K, D, two_var, h, scale, ws, ws_bytes, st, nullptr);
}

// som_peer.cu: out = scale * T @ (sum over the ranks of the rows at the multicast address mc_in), tail reduced exactly
int launch_filter_tc_peer(const float* mc_in, float* out, int K, int D, float two_var, int h, float scale, void* ws,
                          size_t ws_bytes, cudaStream_t st, const peer::Pads& bufs, int64_t q_tail, int world,
                          float* tail_out) {
    ftc::PeerIn pin;
    pin.bufs = bufs;
    pin.q_tail = q_tail;
    pin.world = world;
    pin.tail_out = reinterpret_cast<float4*>(tail_out);
    return launch_filter_tc_in(mc_in, out, K, D, two_var, h, scale, ws, ws_bytes, st, &pin);
}

}  // namespace som
