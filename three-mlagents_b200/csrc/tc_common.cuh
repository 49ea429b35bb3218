// tc_common.cuh — PTX wrappers for tcgen05 / TMEM / mbarrier shared by the tensor-core translation units
// (mlp_tc.cu, mlp_train.cu).  sm_100a only.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

static constexpr int H = 256;

// ------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *holder) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {   // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(NCOLS) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *r) {   // 32 lanes x 16 consecutive columns
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | base_offset=0 | lbo_mode=0 | layout_type=SWIZZLE_NONE(0) [61,64)
// bulk-TMA global -> shared load completing on an mbarrier (no tensor map: plain byte ranges)
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_raw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 (1<<4) | a=bf16 (1<<7) | b=bf16 (1<<10) |
// K-major A and B (bits 15,16 = 0) | N>>3 at [17,23) | M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// ------------------------------------------------------------------------ shared-memory operand layouts
// resident weight tile: 256 rows, K-adjacent core matrices contiguous
static constexpr uint32_t kLBO = 128;
static constexpr uint32_t kSBO = (H / 8) * kLBO;         // 4096 B per 8-row group (K = 256)
static constexpr uint32_t kWBytes = (H / 8) * kSBO;      // 131072
// activation tile: 128 rows; core matrices along K are 144 B apart (16 B pad) so that a warp whose lanes
// hold the 32 consecutive 16-byte chunks of ONE row (a fully coalesced 512-byte global load) stores them
// without shared-memory bank conflicts
static constexpr uint32_t kaLBO = 144;
static constexpr uint32_t kaSBO = (H / 8) * kaLBO;       // 4608
static constexpr uint32_t kABytes = (128 / 8) * kaSBO;   // 73728 (also reused as the 64 KB epilogue stage)

// stage `nrows` (<= R) rows of a row-major bf16 [*,256] matrix into the K-major core-matrix layout.
// chunk q (16 bytes): r = q%8 + 8*(q/(8*32)), kb = (q/8)%32  -> a quarter-warp writes 128 contiguous bytes.
template <int R>
__device__ __forceinline__ void stage_rows(uint8_t *dst, const __nv_bfloat16 *__restrict__ src, int64_t row0, int64_t nvalid) {
    constexpr int KB = H / 8;
    for (int q = threadIdx.x; q < R * KB; q += blockDim.x) {
        const int rin = q & 7, kb = (q >> 3) & (KB - 1), rg = q >> 8;
        const int r = rg * 8 + rin;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < nvalid) v = __ldg(reinterpret_cast<const uint4 *>(src + (row0 + r) * H + kb * 8));
        *reinterpret_cast<uint4 *>(dst + rg * kSBO + kb * kLBO + rin * 16) = v;
    }
}
