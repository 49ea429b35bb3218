// mlp_fwd_pipe.cu — the rollout's policy/value forward with the tensor core OFF the critical path.
//
// Same function as tc_tower_forward_dual_kernel (mlp_tc.cu): ActorCriticPolicy.forward of both towers for the n observations
// of one rollout step (SB3 OnPolicyAlgorithm.collect_rollouts, entered from backend/mlagents/training.py:166), inference only
// (no activations kept).  The classic kernel runs layer 1, the 16 MMAs of layer 2 and the epilogue of a 128-row tile strictly one
// after the other (11.7 K cycles per tile, 2.3 K of them tensor time).  Here:
//   * the layer-2 accumulator is double-buffered in TMEM (2 x 256 columns): the MMAs of tile k+1 run while the CUDA cores drain
//     tile k (bias + tanh + head dot products) — tcgen05.ld overlaps MMAs queued on other columns (profiles/r2_train_variants.txt);
//   * a DRIVER warp issues the MMAs (the issuing lane is blocked ~70 cycles per MMA while the queue is full), the 16 worker
//     warps never wait for the tensor core: by the time they have computed layer 1 of tile k+1 in registers, M1(k) — issued one
//     epilogue earlier — has long released the operand tile;
//   * layer 1 is the training kernel's (weights broadcast from shared memory, FFMA2, 4 rows x 16 columns per thread step), which
//     fits the 96-register budget of a 17-warp CTA.
// Arithmetic is the classic kernel's operation for operation (FMA order over k, tanh.approx, bf16 packing, K-step order of the
// MMAs, summation tree of the head), so outputs are BIT-IDENTICAL to it — tests/test_tc_gpu.py compares the two.
#include <algorithm>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"

static constexpr int kPipeWorkers = 512;                   // 16 worker warps + 1 driver warp

template <int D, int NOUT>
struct PipeSmem {
    static constexpr uint32_t w = 0;                                   // W2 bf16 operand image, K-major (kLBO/kSBO)
    static constexpr uint32_t tile = kWBytes;                          // H1 tile [128][256] bf16, same layout
    static constexpr uint32_t w1t = tile + 128 * 512;                  // float [D][256] (W1 transposed)
    static constexpr uint32_t b1 = w1t + D * H * 4;
    static constexpr uint32_t b2 = b1 + H * 4;
    static constexpr uint32_t wh = b2 + H * 4;                         // float [NOUT][256]
    static constexpr uint32_t xs = wh + NOUT * H * 4;                  // float [128][D]
    static constexpr uint32_t part = xs + 128 * D * 4;                 // float [3][128][NOUT]
    static constexpr uint32_t bar = (part + 3 * 128 * NOUT * 4 + 15) & ~15u;
    static constexpr uint32_t total = bar + 64;
};

struct PipeTowerArgs {
    const float *W1, *B1; const __nv_bfloat16 *W2; const float *B2, *Wh, *Bh; float *out;
};

__device__ __forceinline__ bool pipe_elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void pipe_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int D, int NOUT>
__device__ __forceinline__ void
tower_forward_pipe_body(const PipeTowerArgs &p, const float *__restrict__ x, const int32_t *__restrict__ index, int64_t M,
                        const int32_t *rows_dev, const int64_t bid, const int64_t nblk) {
    using L = PipeSmem<D, NOUT>;
    constexpr int NT = kPipeWorkers;
    constexpr int XPT = (128 * D + NT - 1) / NT;           // gathered obs elements per worker thread per tile
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *Ws = smem + L::w, *Ts = smem + L::tile;
    float *w1t = reinterpret_cast<float *>(smem + L::w1t), *b1s = reinterpret_cast<float *>(smem + L::b1);
    float *b2s = reinterpret_cast<float *>(smem + L::b2), *whs = reinterpret_cast<float *>(smem + L::wh);
    float *xs = reinterpret_cast<float *>(smem + L::xs), *part = reinterpret_cast<float *>(smem + L::part);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L::bar);      // full[2]: layer-2 MMAs into accumulator b done
    uint64_t *rdy = full + 2, *barw = full + 3;                        // workers -> driver "tile staged"; W2 image landed
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(smem + L::bar + 32);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const bool is_driver = warp_u == NT / 32;
    if (rows_dev) M = min(M, (int64_t)*rows_dev);
    const int64_t ntiles = (M + 127) / 128;
    if (bid >= ntiles) return;
    auto worker_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(kPipeWorkers) : "memory"); };

    float xpre[XPT];
    auto prefetch_x = [&](int64_t tile) {                  // element e = tid + NT*i of the [128][D] obs tile (workers only)
        const int64_t row0 = tile * 128;
#pragma unroll
        for (int i = 0; i < XPT; ++i) {
            const int e = tid + NT * i, r = e / D, k = e - r * D;
            float v = 0.0f;
            if (e < 128 * D && row0 + r < M) {
                const int64_t src = index ? (int64_t)__ldg(index + row0 + r) : row0 + r;
                v = __ldg(x + src * D + k);
            }
            xpre[i] = v;
        }
    };
    if (!is_driver) prefetch_x(bid);

    if (warp == 0) tmem_alloc<512>(tmem_holder);
    if (tid == 32) {
        mbar_init(full, 1); mbar_init(full + 1, 1); mbar_init(rdy, 1); mbar_init(barw, 1); fence_barrier_init();
        mbar_expect_tx(barw, kWBytes); bulk_load(smem_u32(Ws), p.W2, kWBytes, barw);     // one bulk-TMA load of the W2 image
    }
    if (!is_driver) {
        // all of a thread's parameter loads are issued before the first store (one global latency instead of one per element)
        constexpr int NW1 = (H * D + NT - 1) / NT, NWH = (NOUT * H + NT - 1) / NT;
        float w1v[NW1], whv[NWH], bv1 = 0.0f, bv2 = 0.0f;
#pragma unroll
        for (int q = 0; q < NW1; ++q) { const int e = tid + q * NT; if (e < H * D) { const int k = e / H, j = e - k * H; w1v[q] = __ldg(p.W1 + j * D + k); } }
#pragma unroll
        for (int q = 0; q < NWH; ++q) { const int e = tid + q * NT; if (e < NOUT * H) whv[q] = __ldg(p.Wh + e); }
        if (tid < H) { bv1 = __ldg(p.B1 + tid); bv2 = __ldg(p.B2 + tid); }
#pragma unroll
        for (int q = 0; q < NW1; ++q) { const int e = tid + q * NT; if (e < H * D) w1t[e] = w1v[q]; }
#pragma unroll
        for (int q = 0; q < NWH; ++q) { const int e = tid + q * NT; if (e < NOUT * H) whs[e] = whv[q]; }
        if (tid < H) { b1s[tid] = bv1; b2s[tid] = bv2; }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_holder, 0);
    const uint32_t t_addr = smem_u32(Ts), w_addr = smem_u32(Ws);

    if (is_driver) {
        // ============================== driver warp: layer-2 MMAs of tile k into accumulator k & 1 ==============================
        if (pipe_elect_one()) {
            mbar_wait(barw, 0);                            // W2 has landed
            uint32_t ph = 0, k = 0;
            for (int64_t tile = bid; tile < ntiles; tile += nblk, ++k) {
                mbar_wait(rdy, ph);
                ph ^= 1u;
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < H / 16; ++kk)
                    umma_bf16(tmem_base + (k & 1u) * 256u, make_desc_raw(t_addr + kk * 2 * kLBO, kLBO, kSBO), make_desc_raw(w_addr + kk * 2 * kLBO, kLBO, kSBO),
                              make_idesc(128, 256), kk > 0 ? 1u : 0u);
                umma_commit(full + (k & 1u));
            }
        }
        __syncwarp();
    } else {
        // ======================================= worker warps =======================================
        const int rt = (warp & 3) * 32 + lane, cq = warp >> 2;               // epilogue: accumulator row, 64-column quarter
        const int l1row = (warp & 3) * 32 + (lane & 7), l1col = cq * 64 + (lane >> 3) * 16;   // layer 1: rows l1row + 8i, 16 columns
        uint32_t h1p[32];
        auto layer1 = [&]() {                              // = mlp_train.cu layer1: H1 = tanh(x W1^T + b1), 4 rows x 16 columns
            float xr[4][D];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < D; ++k) xr[i][k] = xs[(l1row + 8 * i) * D + k];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const int col = l1col + g * 2;
                const float2 bb = *reinterpret_cast<const float2 *>(b1s + col);
                float2 ww[D];
#pragma unroll
                for (int k = 0; k < D; ++k) ww[k] = *reinterpret_cast<const float2 *>(w1t + k * H + col);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float2 v = bb;
#pragma unroll
                    for (int k = 0; k < D; ++k) v = __ffma2_rn(make_float2(xr[i][k], xr[i][k]), ww[k], v);
                    h1p[i * 8 + g] = pack_bf16(tanh_fast(v.x), tanh_fast(v.y));
                }
            }
        };
        auto store_h1 = [&]() {                            // a quarter-warp writes 8 rows of one chunk column: conflict-free
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint8_t *dst = Ts + ((l1row + 8 * i) >> 3) * kSBO + (l1col >> 3) * kLBO + (lane & 7) * 16;
                *reinterpret_cast<uint4 *>(dst) = make_uint4(h1p[i * 8], h1p[i * 8 + 1], h1p[i * 8 + 2], h1p[i * 8 + 3]);
                *reinterpret_cast<uint4 *>(dst + kLBO) = make_uint4(h1p[i * 8 + 4], h1p[i * 8 + 5], h1p[i * 8 + 6], h1p[i * 8 + 7]);
            }
        };
        auto stage_x = [&]() {
#pragma unroll
            for (int i = 0; i < XPT; ++i) { const int e = tid + NT * i; if (e < 128 * D) xs[e] = xpre[i]; }
        };
        uint32_t phf[2] = {0u, 0u};
        // tile 0: stage, layer 1, hand to the driver
        stage_x();
        worker_sync();
        if (bid + nblk < ntiles) prefetch_x(bid + nblk);
        layer1();
        store_h1();
        fence_proxy_async();
        tc_fence_before();
        worker_sync();
        if (tid == 0) pipe_mbar_arrive(rdy);
        uint32_t k = 0;
        for (int64_t tile = bid; tile < ntiles; tile += nblk, ++k) {
            const int64_t row0 = tile * 128;
            const bool has_next = tile + nblk < ntiles;
            if (has_next) {
                // ---- layer 1 of tile k+1 (registers) while the tensor core works on tile k
                stage_x();                                 // xs is free: layer 1 of tile k read it before the last barrier
                worker_sync();
                if (tile + 2 * nblk < ntiles) prefetch_x(tile + 2 * nblk);
                layer1();
            }
            // ---- M1(k) done: the accumulator is ready and the operand tile is free
            mbar_wait(full + (k & 1u), phf[k & 1u]);
            phf[k & 1u] ^= 1u;
            tc_fence_after();
            if (has_next) {
                store_h1();
                fence_proxy_async();
                tc_fence_before();
                worker_sync();                             // also: every worker's epilogue of tile k-1 (accumulator (k+1)&1) is complete
                if (tid == 0) pipe_mbar_arrive(rdy);       // -> driver: M1(k+1) into the other accumulator
            }
            // ---- epilogue of tile k: bias + tanh, head partial dots (identical to mlp_tc.cu tower_forward_body)
            float hsum[NOUT];
#pragma unroll
            for (int a = 0; a < NOUT; ++a) hsum[a] = 0.0f;
            {
                const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (k & 1u) * 256u + (uint32_t)(cq * 64);
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t acc[16];
                    tmem_ld16(taddr + c * 16, acc);
                    const int col = cq * 64 + c * 16;
                    float hv[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 bb = *reinterpret_cast<const float4 *>(b2s + col + 4 * q);
                        hv[4 * q + 0] = tanh_fast(__uint_as_float(acc[4 * q + 0]) + bb.x);
                        hv[4 * q + 1] = tanh_fast(__uint_as_float(acc[4 * q + 1]) + bb.y);
                        hv[4 * q + 2] = tanh_fast(__uint_as_float(acc[4 * q + 2]) + bb.z);
                        hv[4 * q + 3] = tanh_fast(__uint_as_float(acc[4 * q + 3]) + bb.w);
                    }
#pragma unroll
                    for (int a = 0; a < NOUT; ++a)
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 ww = *reinterpret_cast<const float4 *>(whs + a * H + col + 4 * q);
                            hsum[a] = fmaf(hv[4 * q + 0], ww.x, hsum[a]); hsum[a] = fmaf(hv[4 * q + 1], ww.y, hsum[a]);
                            hsum[a] = fmaf(hv[4 * q + 2], ww.z, hsum[a]); hsum[a] = fmaf(hv[4 * q + 3], ww.w, hsum[a]);
                        }
                }
            }
            if (cq > 0) {
#pragma unroll
                for (int a = 0; a < NOUT; ++a) part[((cq - 1) * 128 + rt) * NOUT + a] = hsum[a];
            }
            tc_fence_before();
            worker_sync();                                 // partial sums complete, accumulator k & 1 drained
            if (cq == 0 && row0 + rt < M) {
#pragma unroll
                for (int a = 0; a < NOUT; ++a)
                    p.out[(row0 + rt) * NOUT + a] = ((hsum[a] + part[rt * NOUT + a]) + (part[(128 + rt) * NOUT + a] + part[(256 + rt) * NOUT + a])) + __ldg(p.Bh + a);
            }
            worker_sync();                                 // `part` may be rewritten by the next epilogue
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

template <int D, int A>
__global__ void __launch_bounds__(kPipeWorkers + 32, 1)
tc_tower_forward_pipe_dual_kernel(const __grid_constant__ PipeTowerArgs pi, const __grid_constant__ PipeTowerArgs vf, const float *__restrict__ x,
                                  const int32_t *__restrict__ index, int64_t M, const int32_t *rows_dev) {
    const int64_t bid = blockIdx.x >> 1, nblk = gridDim.x >> 1;
    if (blockIdx.x & 1) tower_forward_pipe_body<D, 1>(vf, x, index, M, rows_dev, bid, nblk);
    else tower_forward_pipe_body<D, A>(pi, x, index, M, rows_dev, bid, nblk);
}

template <int D, int A>
static int pipe_dual_launch_t(const PipeTowerArgs &pi, const PipeTowerArgs &vf, const float *x, const int32_t *index, int64_t M,
                              const int32_t *rows_dev, cudaStream_t st) {
    static int attr_done = 0;
    constexpr uint32_t smem = PipeSmem<D, A>::total > PipeSmem<D, 1>::total ? PipeSmem<D, A>::total : PipeSmem<D, 1>::total;
    static_assert(smem <= 232448, "pipelined tower kernel exceeds the 227 KB shared-memory limit");
    if (!attr_done) {
        TMLA_CUDA(cudaFuncSetAttribute(tc_tower_forward_pipe_dual_kernel<D, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = 1;
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    const unsigned grid = (unsigned)std::min<int64_t>(2 * ((M + 127) / 128), sms & ~1);
    tc_tower_forward_pipe_dual_kernel<D, A><<<grid, kPipeWorkers + 32, smem, st>>>(pi, vf, x, index, M, rows_dev);
    TMLA_LAUNCH_CHECK();
    return TMLA_OK;
}

// both towers of a policy step, inference only; TMLA_EINVAL (no message) for shapes this kernel does not cover
int tc_tower_forward_pipe_dual_launch(int D, int n_actions, const float *const *W1, const float *const *B1, const void *const *W2,
                                      const float *const *B2, const float *const *Wh, const float *const *Bh, const float *x,
                                      const int32_t *index, int64_t M, const int32_t *rows_dev, float *const *out, cudaStream_t st) {
    PipeTowerArgs a[2];
    for (int t = 0; t < 2; ++t) a[t] = PipeTowerArgs{W1[t], B1[t], (const __nv_bfloat16 *)W2[t], B2[t], Wh[t], Bh[t], out[t]};
#define TFP(DD, AA) if (D == DD && n_actions == AA) return pipe_dual_launch_t<DD, AA>(a[0], a[1], x, index, M, rows_dev, st)
    TFP(6, 5); TFP(4, 5); TFP(4, 4);
#undef TFP
    return TMLA_EINVAL;
}
