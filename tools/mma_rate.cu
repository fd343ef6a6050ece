// Microbenchmark: issue rate of tcgen05.mma on B200 for the operand modes the BMU kernels can use.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mma_rate.bin tools/mma_rate.cu
// Prints cycles per MMA instruction (one CTA / CTA pair per SM, all SMs busy) for:
//   kind (tf32 / f16), cta_group 1 / 2, N, A from shared memory (SS) or TMEM (TS), and with
//   concurrent shared-memory store / TMEM-load traffic from other warps (what builders / epilogue do).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t a) {
    return (uint64_t)((a >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// MODE bits: 1 = TS (A in TMEM), 2 = f16 kind (else tf32), 4 = smem store traffic, 8 = tmem ld traffic
template <int CG, int N, int MODE>
__global__ void __launch_bounds__(384, 1) rate_kernel(int iters, int kblocks, long long* out, const uint8_t* src, long long* fill_bytes) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int stop;
    uint8_t* tiles = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool TS = MODE & 1, F16 = MODE & 2, STS = MODE & 4, LDT = MODE & 8, BULK = MODE & 16;
    __shared__ uint64_t fbar[2];
    const int a_blk = 128 * 128, b_blk = (N / CG) * 128;          // bytes per k-block (128-byte rows)
    uint8_t* A = tiles;
    uint8_t* B = tiles + kblocks * a_blk;
    uint8_t* scratch = B + kblocks * b_blk;                       // 32 KB of store-traffic target
    for (int i = threadIdx.x; i < (kblocks * (a_blk + b_blk) + 32768) / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(tiles)[i] = make_uint4(0, 0, 0, 0);
    if (MODE & 32) {   // random finite operands: sign + exponent near 1.0 + random mantissa
        for (int i = threadIdx.x; i < kblocks * (a_blk + b_blk) / 4; i += blockDim.x) {
            uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            reinterpret_cast<uint32_t*>(tiles)[i] = (h & 0x807FFFFFu) | ((124u + (h >> 29)) << 23);
        }
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&fbar[0], 1); mbar_init(&fbar[1], 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const bool leader = (CG == 1) || cluster_rank() == 0;
    constexpr uint32_t M = 128 * CG;
    constexpr uint32_t fmt = F16 ? 1u : 2u;                      // bf16 : tf32
    constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((M >> 4) << 24);
    if (warp == 0 && lane == 0 && leader) {
        const uint64_t ad = umma_desc(smem_u32(A)), bd = umma_desc(smem_u32(B));
        const uint32_t a_tm = tmem + 256;                        // TS: A operand columns (garbage data is fine)
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int kb = 0; kb < kblocks; ++kb) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t a = ad + (uint32_t)((kb * a_blk + k * 32) >> 4);
                    const uint64_t b = bd + (uint32_t)((kb * b_blk + k * 32) >> 4);
                    const uint32_t acc = (it | kb | k) != 0;
                    if (TS) {
                        if (F16)
                            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                                         ::"r"(tmem), "r"(a_tm + k * 8), "l"(b), "r"(idesc), "r"(acc) : "memory");
                        else
                            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                                         ::"r"(tmem), "r"(a_tm + k * 8), "l"(b), "r"(idesc), "r"(acc) : "memory");
                    } else if (CG == 1) {
                        if (F16)
                            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                         ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
                        else
                            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                                         ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
                    } else {
                        if (F16)
                            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                         ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
                        else
                            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                                         ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
                    }
                }
            }
        }
        if (CG == 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        else
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        stop = 1;
        if (CG == 2) {   // release the peer CTA's traffic warps through distributed shared memory
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32((const void*)&stop)), "r"(1u));
            asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(1u) : "memory");
        }
        out[blockIdx.x] = t1 - t0;
    } else if (warp == 1 && lane == 0 && BULK) {
        // TMA-like fill traffic: 16 KB bulk copies global(L2) -> shared, two in flight, back to back
        long long bytes = 0;
        uint32_t ph[2] = {0, 0};
        const uint8_t* g = src + (size_t)(blockIdx.x % 8) * 131072;
        for (int i = 0; i < 2; ++i) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&fbar[i])), "r"(16384) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(scratch + i * 16384)), "l"(g + i * 16384), "r"(16384), "r"(smem_u32(&fbar[i])) : "memory");
        }
        int i = 0, blk = 2;
        while (!stop) {
            mbar_wait(&fbar[i], ph[i]); ph[i] ^= 1; bytes += 16384;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&fbar[i])), "r"(16384) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(scratch + i * 16384)), "l"(g + (blk & 7) * 16384), "r"(16384), "r"(smem_u32(&fbar[i])) : "memory");
            i ^= 1; ++blk;
        }
        mbar_wait(&fbar[0], ph[0]); mbar_wait(&fbar[1], ph[1]);
        fill_bytes[blockIdx.x] = bytes;
    } else if (warp >= 4 && warp < 8 && STS) {
        // builder-like traffic: 16-byte shared stores, conflict-free, as fast as the warp can issue
        uint4* dst = reinterpret_cast<uint4*>(scratch) + (warp - 4) * 512 + lane;
        uint4 v = make_uint4(lane, warp, 0, 0);
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dst[i * 32] = v;
        }
    } else if (warp >= 8 && LDT) {
        // epilogue-like traffic: tcgen05.ld 32x32b.x32 of this warp's lane quarter, continuously
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 384;
        uint32_t sink = 0;
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
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
                    : "r"(taddr + (i & 1) * 32) : "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) sink ^= r[j];
            }
        }
        if (sink == 0x12345678u) out[0] = 0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}


// ---- pipeline skeleton of the BMU kernel's config S: MMA issuer + 8 epilogue warps, operands static ----
// NT = columns per accumulator stage, ST = stages (NT*ST <= 512), KS = k-steps per tile, EPI = epilogue flavour
// (0: chunk-min like the kernel, 1: loads only, 2: 3-input min)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait2(uint64_t* bar, uint32_t parity) {
    if (mbar_test(bar, parity)) return;
    mbar_wait(bar, parity);
}
#define R32OUT(a) "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), \
    "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]),           \
    "=r"(a[16]), "=r"(a[17]), "=r"(a[18]), "=r"(a[19]), "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]),         \
    "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]), "=r"(a[30]), "=r"(a[31])
#define R32IO(a) "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), \
    "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]),           \
    "+r"(a[16]), "+r"(a[17]), "+r"(a[18]), "+r"(a[19]), "+r"(a[20]), "+r"(a[21]), "+r"(a[22]), "+r"(a[23]),         \
    "+r"(a[24]), "+r"(a[25]), "+r"(a[26]), "+r"(a[27]), "+r"(a[28]), "+r"(a[29]), "+r"(a[30]), "+r"(a[31])
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : R32OUT(r) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ldwait(uint32_t (&r)[32]) { asm volatile("tcgen05.wait::ld.sync.aligned;\n" : R32IO(r)::"memory"); }
__device__ __forceinline__ float min3(float a, float b, float c) {
    float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}

template <int NT, int ST, int KS, int EPI, int VAR = 0>
__global__ void __launch_bounds__(384, 1) pipe_kernel(int tiles_n, long long* out, float* sink_out) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t acc_full[ST], acc_empty[ST];
    __shared__ uint32_t tmem_base_s;
    uint8_t* tiles = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a_blk = 128 * 128, b_blk = NT * 128;
    uint8_t* A = tiles;
    uint8_t* B = tiles + (VAR ? 6 : 2) * a_blk;
    for (int i = threadIdx.x; i < (VAR ? (6 * a_blk + 4 * b_blk) : 2 * (a_blk + b_blk)) / 4; i += blockDim.x)
        reinterpret_cast<float*>(tiles)[i] = (float)((i * 37) % 101) * 0.01f;
    if (threadIdx.x == 0) {
        for (int s = 0; s < ST; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
    if (warp == 0 && lane == 0) {
        const uint64_t ad0 = umma_desc(smem_u32(A)), bd0 = umma_desc(smem_u32(B));
        int s = 0; uint32_t ph = 0;
        long long t0 = clock64();
        for (int t = 0; t < tiles_n; ++t) {
            mbar_wait2(&acc_empty[s], ph ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t ad = VAR ? ad0 + (uint32_t)(((t % 3) * 2 * a_blk) >> 4) : ad0;
            const uint64_t bd = VAR ? bd0 + (uint32_t)((((t / 3) & 1) * 2 * b_blk) >> 4) : bd0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const uint64_t a = ad + (uint32_t)((((ks >> 2) & 1) * a_blk + (ks & 3) * 32) >> 4);
                const uint64_t b = bd + (uint32_t)((((ks >> 2) & 1) * b_blk + (ks & 3) * 32) >> 4);
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tmem + s * NT), "l"(a), "l"(b), "r"(idesc), "r"(ks > 0 ? 1u : 0u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&acc_full[s])) : "memory");
            if (++s == ST) { s = 0; ph ^= 1; }
        }
        // drain: wait for the epilogue to release every stage once more
        for (int i = 0; i < ST; ++i) { mbar_wait2(&acc_empty[s], ph ^ 1); if (++s == ST) { s = 0; ph ^= 1; } }
        out[blockIdx.x] = clock64() - t0;
    } else if (warp >= 4) {
        const int ew = warp - 4, half = ew >> 2, lg = warp & 3;
        constexpr int CW = NT / 2;                      // columns per warp per tile
        int s = 0; uint32_t ph = 0;
        float best = INFINITY; int bidx = 0;
        for (int t = 0; t < tiles_n; ++t) {
            mbar_wait2(&acc_full[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(s * NT + half * CW);
            uint32_t va[32], vb[32];
            auto consume = [&](const uint32_t (&v)[32], int c) {
                if (EPI == 1) { best = fminf(best, __uint_as_float(v[c])); return; }
                float q[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (EPI == 2) {
                        float m8 = min3(__uint_as_float(v[g * 8]), __uint_as_float(v[g * 8 + 1]), __uint_as_float(v[g * 8 + 2]));
                        m8 = min3(m8, __uint_as_float(v[g * 8 + 3]), __uint_as_float(v[g * 8 + 4]));
                        m8 = min3(m8, __uint_as_float(v[g * 8 + 5]), __uint_as_float(v[g * 8 + 6]));
                        q[g] = fminf(m8, __uint_as_float(v[g * 8 + 7]));
                    } else {
                        float m8 = __uint_as_float(v[g * 8]);
#pragma unroll
                        for (int i = 1; i < 8; ++i) m8 = fminf(m8, __uint_as_float(v[g * 8 + i]));
                        q[g] = m8;
                    }
                }
                const float m = fminf(fminf(q[0], q[1]), fminf(q[2], q[3]));
                if (m < best) {
                    best = m;
                    const int sub = (q[0] == m) ? 0 : (q[1] == m) ? 1 : (q[2] == m) ? 2 : 3;
                    bidx = t * NT + c * 32 + sub * 8;
                }
            };
            ld32(taddr, va);
#pragma unroll
            for (int c = 0; c < CW / 32; c += 2) {
                ldwait(va);
                if (c + 1 < CW / 32) ld32(taddr + (c + 1) * 32, vb);
                consume(va, c);
                if (c + 1 < CW / 32) {
                    ldwait(vb);
                    if (c + 2 < CW / 32) ld32(taddr + (c + 2) * 32, va);
                    else {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[s])) : "memory");
                    }
                    consume(vb, c + 1);
                }
            }
            if (++s == ST) { s = 0; ph ^= 1; }
        }
        if (best == 123.456f) sink_out[threadIdx.x] = best + bidx;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

template <int NT, int ST, int KS, int EPI, int VAR = 0>
static void run_pipe(const char* name) {
    const int grid = 148, tiles_n = 2048 * 256 / NT;
    long long* out; float* sink;
    cudaMalloc(&out, grid * sizeof(long long));
    cudaMalloc(&sink, 4096);
    const int smem = 1024 + (VAR ? 6 * 128 * 128 + 4 * NT * 128 : 2 * (128 * 128 + NT * 128));
    cudaFuncSetAttribute(pipe_kernel<NT, ST, KS, EPI, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; ++rep) {
        pipe_kernel<NT, ST, KS, EPI, VAR><<<grid, 384, smem>>>(tiles_n, out, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-44s ERROR %s\n", name, cudaGetErrorString(e)); exit(1); }
    }
    long long h[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double sum = 0;
    for (int i = 0; i < grid; ++i) sum += h[i];
    const double per256 = sum / grid / tiles_n * (256.0 / NT);
    printf("%-52s %.0f cyc per 128x256 tile (MMA floor %d)\n", name, per256, KS * 128);
    cudaFree(out); cudaFree(sink);
}

// issue-queue probe: timestamps after each of 24 back-to-back MMA issues (one CTA)
__global__ void __launch_bounds__(128, 1) queue_probe(long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* tiles = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(tiles)[i] = 1.0f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((128u >> 4) << 24);
    if (warp == 0 && lane == 0) {
        const uint64_t ad = umma_desc(smem_u32(tiles)), bd = umma_desc(smem_u32(tiles + 16384));
        long long ts[26];
        ts[0] = clock64();
#pragma unroll
        for (int i = 0; i < 24; ++i) {
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tmem + (i & 1) * 256), "l"(ad + (uint32_t)((i & 3) * 2)), "l"(bd + (uint32_t)((i & 3) * 2)), "r"(idesc), "r"(1u) : "memory");
            ts[i + 1] = clock64();
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, 0);
        ts[25] = clock64();
        for (int i = 0; i < 26; ++i) out[i] = ts[i] - ts[0];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

// issue-queue probe 2: [7 dependent MMAs + commit] x 4 tiles alternating accumulators; timestamps after every op
__global__ void __launch_bounds__(128, 1) queue_probe2(long long* out, int same_acc) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tmem_base_s;
    uint8_t* tiles = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(tiles)[i] = 1.0f;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((128u >> 4) << 24);
    if (warp == 0 && lane == 0) {
        const uint64_t ad = umma_desc(smem_u32(tiles)), bd = umma_desc(smem_u32(tiles + 16384));
        long long ts[40];
        int k = 0;
        ts[k++] = clock64();
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tmem + (same_acc ? 0 : (t & 1) * 256)), "l"(ad + (uint32_t)((i & 3) * 2)), "l"(bd + (uint32_t)((i & 3) * 2)), "r"(idesc), "r"(i > 0 ? 1u : 0u) : "memory");
                ts[k++] = clock64();
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[t])) : "memory");
            ts[k++] = clock64();
        }
        mbar_wait(&bar[3], 0);
        ts[k++] = clock64();
        for (int i = 0; i < k; ++i) out[i] = ts[i] - ts[0];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}
static void run_probe2(int same_acc) {
    long long* out; cudaMalloc(&out, 40 * 8);
    cudaFuncSetAttribute(queue_probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    for (int rep = 0; rep < 2; ++rep) { queue_probe2<<<1, 128, 60000>>>(out, same_acc); cudaDeviceSynchronize(); }
    long long h[40]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("probe2 (same_acc=%d): per tile [7 MMA issue stamps | commit stamp]; last = completion\n", same_acc);
    for (int t = 0; t < 4; ++t) { for (int i = 0; i < 8; ++i) printf(" %lld%s", h[1 + t * 8 + i], i == 6 ? " |" : ""); printf("\n"); }
    printf(" done %lld\n", h[33]);
}

static void run_probe() {
    long long* out; cudaMalloc(&out, 26 * 8);
    cudaFuncSetAttribute(queue_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    for (int rep = 0; rep < 2; ++rep) { queue_probe<<<1, 128, 60000>>>(out); cudaDeviceSynchronize(); }
    long long h[26]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("issue timestamps (cycles after first issue) for 24 MMAs, then completion:\n");
    for (int i = 0; i < 26; ++i) printf(" %lld", h[i]);
    printf("\n");
}

template <int CG, int N, int MODE>
static void run(const char* name, int kblocks) {
    const int iters = 40000 / kblocks, grid = 148;
    long long* out; long long* fb; uint8_t* src;
    cudaMalloc(&out, grid * sizeof(long long));
    cudaMalloc(&fb, grid * sizeof(long long));
    cudaMalloc(&src, 8 * 131072);
    cudaMemset(src, 0, 8 * 131072);
    cudaMemset(out, 0, grid * sizeof(long long));
    cudaMemset(fb, 0, grid * sizeof(long long));
    const int smem = 1024 + kblocks * (128 * 128 + (N / CG) * 128) + 32768;
    cudaFuncSetAttribute(rate_kernel<CG, N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<CG, N, MODE>, iters, kblocks, out, (const uint8_t*)src, fb);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-44s ERROR %s\n", name, cudaGetErrorString(e)); exit(1); }
    }
    long long h[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double sum = 0; int cnt = 0; long long mx = 0;
    for (int i = 0; i < grid; ++i) if (h[i] > 0) { sum += h[i]; ++cnt; if (h[i] > mx) mx = h[i]; }
    const double n_mma = (double)iters * kblocks * 4;
    long long hb[148];
    cudaMemcpy(hb, fb, sizeof(hb), cudaMemcpyDeviceToHost);
    double fbs = 0; for (int i = 0; i < grid; ++i) fbs += hb[i];
    printf("%-44s kblocks=%d  avg %.1f cyc/MMA  (max %.1f)  -> %.0f MAC/clk/SM  fill %.1f B/clk/SM\n", name, kblocks, sum / cnt / n_mma,
           mx / n_mma, 128.0 * N * ((MODE & 2) ? 16 : 8) / (sum / cnt / n_mma), fbs / grid / (sum / cnt) * (CG == 2 ? 1.0 : 1.0));
    cudaFree(out); cudaFree(fb); cudaFree(src);
}


// ---- sustained peaks: a long tcgen05 TF32 run (power-capped clock) and an FP32 FFMA run, timed with events ----
__global__ void __launch_bounds__(1024) ffma_peak_kernel(float* out, int iters) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-6f + i;
    const float b = 1.000001f, c = 1e-7f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456f) out[threadIdx.x] = s;
}

template <int MODE>
static void sustained_tc(const char* name, int kdepth) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148, kblocks = 2, iters = 600000;     // 4.8 M MMAs per CTA
    long long* out; long long* fb; uint8_t* src;
    cudaMalloc(&out, grid * sizeof(long long)); cudaMalloc(&fb, grid * sizeof(long long)); cudaMalloc(&src, 8 * 131072);
    const int smem = 1024 + kblocks * (128 * 128 + 256 * 128) + 32768;
    cudaFuncSetAttribute(rate_kernel<1, 256, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, rate_kernel<1, 256, MODE>, 1000, kblocks, out, (const uint8_t*)src, fb);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, rate_kernel<1, 256, MODE>, iters, kblocks, out, (const uint8_t*)src, fb);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    const double n_mma = (double)iters * kblocks * 4;
    const double tf = 148.0 * n_mma * 128 * 256 * kdepth * 2 / (ms * 1e-3) / 1e12;
    printf("SUSTAINED %s tcgen05 (random operands): %.1f ms, %.1f TFLOP/s dense -> %.1f TFLOP/s fp32-faithful (3 products), "
           "%.1f cycles/MMA, average SM clock %.3f GHz  [%s]\n", name, ms, tf, tf / 3.0, (double)h[0] / n_mma,
           (double)h[0] / (ms * 1e-3) / 1e9, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
    cudaFree(out); cudaFree(fb); cudaFree(src);
}

static void run_sustained() {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    // tensor pipe: ~0.4-0.6 s of back-to-back M128 N256 MMAs on every SM, kind::tf32 (K = 8) then kind::f16 with bf16
    // operands (K = 16): the second line is what a two-way 16-bit split (3 products at twice the rate) would run at
    sustained_tc<32>("tf32", 8);
    sustained_tc<34>("f16 kind (bf16 operands)", 16);
    // FP32 FFMA: 148 x 2 CTAs x 1024 threads, 64 independent FMAs per iteration
    {
        float* out; cudaMalloc(&out, 4096);
        const int iters = 200000;
        ffma_peak_kernel<<<148 * 2, 1024>>>(out, 1000);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        ffma_peak_kernel<<<148 * 2, 1024>>>(out, iters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 148.0 * 2 * 1024 * (double)iters * 64 * 2;
        printf("SUSTAINED fp32 FFMA: %.1f ms, %.1f TFLOP/s\n", ms, fl / (ms * 1e-3) / 1e12);
        cudaFree(out);
    }
}

int main(int argc, char** argv) {
    if (argc > 1 && argv[1][0] == 's') { run_sustained(); return 0; }
    run_probe();
    run_probe2(0);
    run_probe2(1);
    run<1, 256, 0>("tf32 cg1 SS N=256", 2);
    run<1, 256, 0>("tf32 cg1 SS N=256", 4);
    run<1, 128, 0>("tf32 cg1 SS N=128", 2);
    run<1, 64, 0>("tf32 cg1 SS N=64", 2);
    run<1, 256, 1>("tf32 cg1 TS N=256 (A in TMEM)", 2);
    run<1, 128, 1>("tf32 cg1 TS N=128 (A in TMEM)", 2);
    run<2, 256, 0>("tf32 cg2 SS M=256 N=256", 2);
    run<2, 256, 0>("tf32 cg2 SS M=256 N=256", 4);
    run<2, 128, 0>("tf32 cg2 SS M=256 N=128", 2);
    run<1, 256, 2>("bf16 cg1 SS N=256", 2);
    run<2, 256, 2>("bf16 cg2 SS M=256 N=256", 2);
    run<1, 256, 4>("tf32 cg1 SS N=256 + smem stores", 2);
    run<1, 256, 8>("tf32 cg1 SS N=256 + tmem loads", 2);
    run<1, 256, 12>("tf32 cg1 SS N=256 + stores + tmem loads", 2);
    run<2, 256, 12>("tf32 cg2 SS M=256 N=256 + stores + tmem lds", 2);
    run<1, 256, 9>("tf32 cg1 TS N=256 + tmem loads", 2);
    run<1, 256, 32>("tf32 cg1 SS N=256 RANDOM data", 2);
    run<1, 256, 32>("tf32 cg1 SS N=256 RANDOM data", 4);
    run<2, 256, 32>("tf32 cg2 SS M=256 N=256 RANDOM data", 2);
    run<1, 256, 34>("bf16 cg1 SS N=256 RANDOM data", 2);
    run<1, 256, 33>("tf32 cg1 TS N=256 RANDOM data", 2);
    run<1, 256, 16>("tf32 cg1 SS N=256 + bulk fills", 2);
    run<1, 256, 28>("tf32 cg1 SS N=256 + fills + stores + tmem lds", 2);
    run<2, 256, 16>("tf32 cg2 SS M=256 N=256 + bulk fills", 2);
    run<1, 128, 16>("tf32 cg1 SS N=128 + bulk fills", 2);
    run<1, 256, 17>("tf32 cg1 TS N=256 + bulk fills", 2);
    run_pipe<256, 2, 7, 0>("pipe N=256 x2 stages, 7 ksteps, chunk-min epi");
    run_pipe<256, 2, 7, 0, 1>("pipe N=256 x2, 7 ksteps, chunk-min, 3 A slots + 2 B stages");
    run_pipe<256, 2, 7, 1>("pipe N=256 x2 stages, 7 ksteps, loads-only epi");
    run_pipe<256, 2, 7, 2>("pipe N=256 x2 stages, 7 ksteps, min3 epi");
    run_pipe<128, 4, 7, 0>("pipe N=128 x4 stages, 7 ksteps, chunk-min epi");
    run_pipe<128, 4, 7, 2>("pipe N=128 x4 stages, 7 ksteps, min3 epi");
    run_pipe<128, 4, 7, 1>("pipe N=128 x4 stages, 7 ksteps, loads-only epi");
    run_pipe<256, 2, 25, 0>("pipe N=256 x2 stages, 25 ksteps(2 kblk reuse), chunk-min");
    run_pipe<256, 2, 6, 2>("pipe N=256 x2 stages, 6 ksteps, min3 epi");
    run_pipe<128, 4, 6, 2>("pipe N=128 x4 stages, 6 ksteps, min3 epi");
    return 0;
}
