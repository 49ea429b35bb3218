// mlp_train.cu — one PPO minibatch of one tower (policy or value), forward + loss + backward in ONE kernel.
//
// Replaces, for the minibatch loop of SB3's PPO.train (reached from backend/mlagents/training.py:166 with the
// hyper-parameters of training.py:379-389), the chain
//     evaluate_actions (tower forward) -> advantage normalisation -> clipped surrogate / value MSE / entropy
//     -> autograd backward through head, layer 2 and layer 1
// that the unfused path runs as 9 kernels with every [rows,256] activation making a round trip through HBM
// (10 KB per sample).  Here one persistent CTA per SM walks 128-row tiles and keeps everything on chip:
//
//   P1  gather obs rows, layer 1 on CUDA cores            -> H1 (bf16) in the shared-memory operand tile
//   P2  tcgen05.mma  Z2 = H1 . W2^T  (W2 resident, K-major)   -> TMEM;  H1 rows leave for HBM meanwhile
//   P3  TMEM -> bias + tanh -> head dot products           -> H2 (bf16) overwrites the tile (MMA done)
//   P4  loss per row (softmax / ratio / clip / entropy, or value MSE), analytic d(loss)/d(head output)
//   P5  dZ2 = (dout . Wh) * (1 - H2^2) in place; dWh, db2 accumulated in registers
//   P6  tcgen05.mma  dH1 = dZ2 . W2  — the SAME resident W2 bytes read through an MN-major descriptor
//       (no transposed copy);  dZ2 rows leave for HBM meanwhile
//   P7  TMEM -> dH1 (bf16) -> tile;  dZ1 = dH1 * (1 - H1^2);  dW1, db1 accumulated in registers
//
// Only H1 and dZ2 (2 x 512 B per sample) are written out, for the split-K weight-gradient GEMM
// dW2 = dZ2^T . H1 whose 256x256 fp32 accumulator needs all 512 TMEM columns (tc_wgrad_mn_kernel below: both
// operands are read through MN-major descriptors straight from row-major staging, no transposes).
//
// MN-major trick: a [rows][256] bf16 tile staged in the canonical K-major SWIZZLE_NONE layout
// (chunk (r, cb) of 16 bytes at (r/8)*G + cb*C + (r%8)*16) is, byte for byte, also the canonical MN-major
// layout of its transpose with SBO = C (stride between 8-element groups along MN) and LBO = G (stride between
// 8-row groups along K) — cute::UMMA "((1,n),(8,k)):((X,SBO),(1,LBO))" in uint128 units.
#include <algorithm>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "mlp_common.cuh"

static constexpr int kTrainThreads = 512;                  // 16 warps

__host__ __device__ constexpr uint32_t make_idesc_major(int M, int N, int a_mn, int b_mn) {
    return make_idesc(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct TowerTrainArgs {
    // tower parameters (fp32 views into the flat vector; W2 from the bf16 pack)
    const float *W1, *B1;
    const __nv_bfloat16 *W2;
    const float *B2, *Wh, *Bh;
    // minibatch
    const float *x;                 // obs [total][D]
    const int32_t *index;           // int32[M] row offsets into x / the [T*N] buffers (NULL = identity)
    int64_t M;
    const int32_t *actions;         // policy tower
    const float *adv, *old_logp;
    const float *returns;           // value tower
    const double *adv_sums;         // (sum, sum of squares, count) or NULL
    int normalize;
    float clip, ent_coef, vf_coef, inv_rows;
    // outputs
    __nv_bfloat16 *h1_out, *dz2_out;    // [M][256] each, for the dW2 GEMM
    float *out;                          // optional [M][NOUT] head outputs (logits / values)
    float *gW1, *gB1, *gB2, *gWh, *gBh;  // gradient slices (accumulated with atomics; caller zeroes)
    float *stats;                        // float[8] (tmla_ppo_loss layout), accumulated
};

template <int D, int NOUT>
struct TrainSmem {
    static constexpr uint32_t w = 0;                                   // W2 bf16, K-major (kLBO/kSBO)
    static constexpr uint32_t a = kWBytes;                             // the tile: H1 -> H2 -> dZ2 -> dH1 (kaLBO/kaSBO)
    static constexpr uint32_t w1 = a + kABytes;                        // float [256][D]
    static constexpr uint32_t b1 = w1 + H * D * 4;                     // float [256]
    static constexpr uint32_t b2 = b1 + H * 4;                         // float [256]
    static constexpr uint32_t wh = b2 + H * 4;                         // float [NOUT][256]
    static constexpr uint32_t xs = wh + NOUT * H * 4;                  // float [128][D]
    static constexpr uint32_t part = xs + 128 * D * 4;                 // float [3][128][NOUT]
    static constexpr uint32_t dout = part + 3 * 128 * NOUT * 4;        // float [128][NOUT]
    static constexpr uint32_t scal = dout + 128 * NOUT * 4;            // float [16]: stats[0..5), dbh[NOUT]
    static constexpr uint32_t bar = (scal + 64 + 15) & ~15u;
    static constexpr uint32_t total = bar + 64;
};

template <int D, int NOUT>
__global__ void __launch_bounds__(kTrainThreads, 1)
tc_tower_train_kernel(const __grid_constant__ TowerTrainArgs p) {
    using L = TrainSmem<D, NOUT>;
    constexpr int NT = kTrainThreads;
    constexpr bool PI = NOUT > 1;
    constexpr int XPT = (128 * D + NT - 1) / NT;
    constexpr int NV = NOUT + 1 + D + 1;                   // per-column gradient values: dWh[NOUT], db2, dW1[D], db1
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *Ws = smem + L::w, *As = smem + L::a;
    float *w1s = reinterpret_cast<float *>(smem + L::w1), *b1s = reinterpret_cast<float *>(smem + L::b1);
    float *b2s = reinterpret_cast<float *>(smem + L::b2), *whs = reinterpret_cast<float *>(smem + L::wh);
    float *xs = reinterpret_cast<float *>(smem + L::xs), *part = reinterpret_cast<float *>(smem + L::part);
    float *douts = reinterpret_cast<float *>(smem + L::dout), *scal = reinterpret_cast<float *>(smem + L::scal);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L::bar);
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(smem + L::bar + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t M = p.M;
    const int64_t ntiles = (M + 127) / 128;
    if ((int64_t)blockIdx.x >= ntiles) return;

    // column-owner role (P1, P5, P7b): 4 hidden units c0..c0+3, rows rgrp + 8*i
    const int half = warp & 1, rgrp = warp >> 1;
    const int c0 = half * 128 + lane * 4;
    const uint32_t slot0 = (uint32_t)(c0 >> 3) * kaLBO + (uint32_t)rgrp * 16 + (uint32_t)(lane & 1) * 8;   // + i*kaSBO
    // row-owner role (P3, P7a): row rt of the tile, 64-column quarter cq
    const int rt = (warp & 3) * 32 + lane, cq = warp >> 2;

    float xpre[XPT];
    int32_t a_pre = 0;
    float f0_pre = 0.0f, f1_pre = 0.0f;                    // policy: advantage, old log-prob;  value: return
    auto prefetch = [&](int64_t tile) {
        const int64_t row0 = tile * 128;
#pragma unroll
        for (int i = 0; i < XPT; ++i) {
            const int e = tid + NT * i, r = e / D, k = e - r * D;
            float v = 0.0f;
            if (e < 128 * D && row0 + r < M) {
                const int64_t src = p.index ? (int64_t)__ldg(p.index + row0 + r) : row0 + r;
                v = __ldg(p.x + src * D + k);
            }
            xpre[i] = v;
        }
        if (tid < 128 && row0 + tid < M) {
            const int64_t src = p.index ? (int64_t)__ldg(p.index + row0 + tid) : row0 + tid;
            if (PI) { a_pre = __ldg(p.actions + src); f0_pre = __ldg(p.adv + src); f1_pre = __ldg(p.old_logp + src); }
            else f0_pre = __ldg(p.returns + src);
        }
    };
    prefetch(blockIdx.x);

    if (warp == 0) tmem_alloc<256>(tmem_holder);
    if (tid == 32) { mbar_init(bar, 1); fence_barrier_init(); }
    stage_rows<H>(Ws, p.W2, 0, H);
    for (int e = tid; e < H * D; e += NT) w1s[e] = p.W1[e];
    for (int e = tid; e < NOUT * H; e += NT) whs[e] = p.Wh[e];
    if (tid < H) { b1s[tid] = p.B1[tid]; b2s[tid] = p.B2[tid]; }
    if (tid < 16) scal[tid] = 0.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t idesc_fwd = make_idesc_major(128, 256, 0, 0);   // A K-major, B K-major
    const uint32_t idesc_dgr = make_idesc_major(128, 256, 0, 1);   // A K-major, B MN-major (W2 read transposed)
    const uint32_t a_addr = smem_u32(As), w_addr = smem_u32(Ws);
    uint32_t phase = 0;

    // advantage normalisation constants (PPO.train: (adv - mean) / (std + 1e-8), std unbiased)
    float adv_mean = 0.0f, adv_inv_std = 1.0f;
    if (PI && p.normalize) {
        const double cnt = p.adv_sums[2], mu = p.adv_sums[0] / cnt;
        double var = (p.adv_sums[1] - p.adv_sums[0] * mu) / (cnt - 1.0);
        var = var > 0.0 ? var : 0.0;
        adv_mean = (float)mu;
        adv_inv_std = 1.0f / ((float)sqrt(var) + 1e-8f);
        if (blockIdx.x == 0 && tid == 0) { p.stats[6] = adv_mean; p.stats[7] = (float)sqrt(var); }
    }

    // per-column gradient accumulators, persistent across the CTA's tiles (pairs for the packed fp32 pipe)
    float2 gWh01[NOUT], gWh23[NOUT], gW101[D], gW123[D];
    float2 gb2_01 = make_float2(0.f, 0.f), gb2_23 = gb2_01, gb1_01 = gb2_01, gb1_23 = gb2_01;
#pragma unroll
    for (int a = 0; a < NOUT; ++a) gWh01[a] = gWh23[a] = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < D; ++k) gW101[k] = gW123[k] = make_float2(0.f, 0.f);

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * 128;
        const int32_t a_cur = a_pre;
        const float f0_cur = f0_pre, f1_cur = f1_pre;
#pragma unroll
        for (int i = 0; i < XPT; ++i) { const int e = tid + NT * i; if (e < 128 * D) xs[e] = xpre[i]; }
        __syncthreads();
        // ---- P1: layer 1, H1 = tanh(x W1^T + b1) -> K-major tile
        {
            float2 w01[D], w23[D];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                w01[k] = make_float2(w1s[(c0 + 0) * D + k], w1s[(c0 + 1) * D + k]);
                w23[k] = make_float2(w1s[(c0 + 2) * D + k], w1s[(c0 + 3) * D + k]);
            }
            const float4 bb = *reinterpret_cast<const float4 *>(b1s + c0);
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const float *xr = xs + (rgrp + 8 * i) * D;
                float2 v01 = make_float2(bb.x, bb.y), v23 = make_float2(bb.z, bb.w);
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const float2 xx = make_float2(xr[k], xr[k]);
                    v01 = __ffma2_rn(xx, w01[k], v01);
                    v23 = __ffma2_rn(xx, w23[k], v23);
                }
                *reinterpret_cast<uint2 *>(As + i * kaSBO + slot0) =
                    make_uint2(pack_bf16(tanh_fast(v01.x), tanh_fast(v01.y)), pack_bf16(tanh_fast(v23.x), tanh_fast(v23.y)));
            }
        }
        fence_proxy_async();
        __syncthreads();
        // ---- P2: Z2 = H1 . W2^T on the tensor core; H1 rows -> HBM and next-tile prefetch meanwhile
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < H / 16; ++kk)
                umma_bf16(tmem_base, make_desc_raw(a_addr + kk * 2 * kaLBO, kaLBO, kaSBO),
                          make_desc_raw(w_addr + kk * 2 * kLBO, kLBO, kSBO), idesc_fwd, kk > 0 ? 1u : 0u);
            umma_commit(bar);
        }
#pragma unroll 2
        for (int i = 0; i < 8; ++i) {
            const int r = warp + 16 * i;
            const uint4 v = *reinterpret_cast<const uint4 *>(As + (r >> 3) * kaSBO + lane * kaLBO + (r & 7) * 16);
            if (row0 + r < M) reinterpret_cast<uint4 *>(p.h1_out + (row0 + r) * H)[lane] = v;
        }
        if (tile + gridDim.x < ntiles) prefetch(tile + gridDim.x);
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        __syncthreads();                                   // every warp has copied its H1 rows out: the tile may be overwritten
        // ---- P3: H2 = tanh(Z2 + b2) -> tile (K-major), head partial dot products
        {
            float2 hs[NOUT];
#pragma unroll
            for (int a = 0; a < NOUT; ++a) hs[a] = make_float2(0.f, 0.f);
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cq * 64);
            uint8_t *trow = As + (rt >> 3) * kaSBO + (rt & 7) * 16;
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
                        hs[a] = __ffma2_rn(make_float2(hv[4 * q + 0], hv[4 * q + 1]), make_float2(ww.x, ww.y), hs[a]);
                        hs[a] = __ffma2_rn(make_float2(hv[4 * q + 2], hv[4 * q + 3]), make_float2(ww.z, ww.w), hs[a]);
                    }
                uint32_t o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = pack_bf16(hv[2 * j], hv[2 * j + 1]);
                const int kb = col >> 3;
                *reinterpret_cast<uint4 *>(trow + kb * kaLBO) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4 *>(trow + (kb + 1) * kaLBO) = make_uint4(o[4], o[5], o[6], o[7]);
            }
            if (cq > 0) {
#pragma unroll
                for (int a = 0; a < NOUT; ++a) part[((cq - 1) * 128 + rt) * NOUT + a] = hs[a].x + hs[a].y;
            }
            tc_fence_before();
            __syncthreads();                               // TMEM drained, H2 tile and partial sums complete
            // ---- P4: loss of row rt (threads of column quarter 0 = tid 0..127), d(loss)/d(head output)
            if (cq == 0) {
                const bool valid = row0 + rt < M;
                float z[NOUT], dz[NOUT];
#pragma unroll
                for (int a = 0; a < NOUT; ++a)
                    z[a] = (((hs[a].x + hs[a].y) + part[rt * NOUT + a]) + (part[(128 + rt) * NOUT + a] + part[(256 + rt) * NOUT + a])) + __ldg(p.Bh + a);
                if (p.out && valid) {
#pragma unroll
                    for (int a = 0; a < NOUT; ++a) p.out[(row0 + rt) * NOUT + a] = z[a];
                }
                float st[4] = {0.f, 0.f, 0.f, 0.f};        // policy: pg, entropy-loss, kl, clipfrac;  value: st[0] = squared error
                if (PI) {
                    float m = z[0];
#pragma unroll
                    for (int j = 1; j < NOUT; ++j) m = fmaxf(m, z[j]);
                    float pr[NOUT], lp[NOUT], S = 0.0f;
#pragma unroll
                    for (int j = 0; j < NOUT; ++j) { pr[j] = expf(z[j] - m); S += pr[j]; }
                    const float logS = logf(S), invS = 1.0f / S;
                    float ent = 0.0f;
#pragma unroll
                    for (int j = 0; j < NOUT; ++j) { lp[j] = (z[j] - m) - logS; pr[j] *= invS; ent -= pr[j] * lp[j]; }
                    float logp = lp[0];
#pragma unroll
                    for (int j = 1; j < NOUT; ++j) logp = (a_cur == j) ? lp[j] : logp;
                    float adv = f0_cur;
                    if (p.normalize) adv = (adv - adv_mean) * adv_inv_std;
                    const float lr = logp - f1_cur;
                    const float ratio = expf(lr);
                    const float lo = 1.0f - p.clip, hi = 1.0f + p.clip;
                    const float s1 = adv * ratio, s2 = adv * fminf(fmaxf(ratio, lo), hi);
                    const bool inside = (ratio >= lo) && (ratio <= hi);
                    const bool active = inside || (s1 < s2);
                    const float dlogp = active ? (-adv * ratio * p.inv_rows) : 0.0f;
#pragma unroll
                    for (int j = 0; j < NOUT; ++j)
                        dz[j] = valid ? dlogp * ((a_cur == j ? 1.0f : 0.0f) - pr[j]) + p.ent_coef * p.inv_rows * pr[j] * (lp[j] + ent) : 0.0f;
                    if (valid) {
                        st[0] = -fminf(s1, s2); st[1] = -ent; st[2] = (ratio - 1.0f) - lr;
                        st[3] = (fabsf(ratio - 1.0f) > p.clip) ? 1.0f : 0.0f;
                    }
                } else {
                    const float dv = z[0] - f0_cur;
                    dz[0] = valid ? p.vf_coef * 2.0f * dv * p.inv_rows : 0.0f;
                    if (valid) st[0] = dv * dv;
                }
#pragma unroll
                for (int a = 0; a < NOUT; ++a) douts[rt * NOUT + a] = dz[a];
                // per-tile scalar reductions: stats and the head-bias gradient
#pragma unroll
                for (int q = 0; q < (PI ? 4 : 1); ++q) {
                    const float s = warp_sum_f(st[q]);
                    if (lane == 0) atomicAdd(scal + q, s);
                }
#pragma unroll
                for (int a = 0; a < NOUT; ++a) {
                    const float s = warp_sum_f(dz[a]);
                    if (lane == 0) atomicAdd(scal + 8 + a, s);
                }
            }
        }
        __syncthreads();
        // ---- P5: dZ2 = (dout . Wh) * (1 - H2^2) in place;  dWh += dout^T H2,  db2 += sum dZ2
        {
            float2 wh01[NOUT], wh23[NOUT];
#pragma unroll
            for (int a = 0; a < NOUT; ++a) {
                const float4 ww = *reinterpret_cast<const float4 *>(whs + a * H + c0);
                wh01[a] = make_float2(ww.x, ww.y); wh23[a] = make_float2(ww.z, ww.w);
            }
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const int r = rgrp + 8 * i;
                uint2 *slot = reinterpret_cast<uint2 *>(As + i * kaSBO + slot0);
                const uint2 hr = *slot;
                const float2 h01 = make_float2(bf16_lo(hr.x), bf16_hi(hr.x)), h23 = make_float2(bf16_lo(hr.y), bf16_hi(hr.y));
                float2 d01 = make_float2(0.f, 0.f), d23 = d01;
#pragma unroll
                for (int a = 0; a < NOUT; ++a) {
                    const float da = douts[r * NOUT + a];
                    const float2 dd = make_float2(da, da);
                    d01 = __ffma2_rn(dd, wh01[a], d01);
                    d23 = __ffma2_rn(dd, wh23[a], d23);
                    gWh01[a] = __ffma2_rn(dd, h01, gWh01[a]);
                    gWh23[a] = __ffma2_rn(dd, h23, gWh23[a]);
                }
                const float2 one = make_float2(1.f, 1.f);
                const float2 s01 = __ffma2_rn(make_float2(-h01.x, -h01.y), h01, one), s23 = __ffma2_rn(make_float2(-h23.x, -h23.y), h23, one);
                d01 = __fmul2_rn(d01, s01); d23 = __fmul2_rn(d23, s23);
                gb2_01 = __fadd2_rn(gb2_01, d01); gb2_23 = __fadd2_rn(gb2_23, d23);
                *slot = make_uint2(pack_bf16(d01.x, d01.y), pack_bf16(d23.x, d23.y));
            }
        }
        fence_proxy_async();
        __syncthreads();
        // ---- P6: dH1 = dZ2 . W2 on the tensor core (W2 tile read MN-major); dZ2 rows -> HBM, H1 rows come back
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < H / 16; ++kk)            // K = layer-2 output index j: 16 rows of the W2 tile per step
                umma_bf16(tmem_base, make_desc_raw(a_addr + kk * 2 * kaLBO, kaLBO, kaSBO),
                          make_desc_raw(w_addr + kk * 2 * kSBO, /*LBO (K groups)*/ kSBO, /*SBO (N groups)*/ kLBO), idesc_dgr, kk > 0 ? 1u : 0u);
            umma_commit(bar);
        }
#pragma unroll 2
        for (int i = 0; i < 8; ++i) {
            const int r = warp + 16 * i;
            const uint4 v = *reinterpret_cast<const uint4 *>(As + (r >> 3) * kaSBO + lane * kaLBO + (r & 7) * 16);
            if (row0 + r < M) reinterpret_cast<uint4 *>(p.dz2_out + (row0 + r) * H)[lane] = v;
        }
        uint2 h1pre[16];                                   // this thread's H1 values (written in P2, L2-resident)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int64_t r = row0 + rgrp + 8 * i;
            h1pre[i] = (r < M) ? *reinterpret_cast<const uint2 *>(p.h1_out + r * H + c0) : make_uint2(0u, 0u);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        __syncthreads();                                   // dZ2 rows copied out by every warp
        // ---- P7a: dH1 (fp32, TMEM) -> bf16 -> tile
        {
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cq * 64);
            uint8_t *trow = As + (rt >> 3) * kaSBO + (rt & 7) * 16;
#pragma unroll 2
            for (int c = 0; c < 4; ++c) {
                uint32_t acc[16];
                tmem_ld16(taddr + c * 16, acc);
                uint32_t o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = pack_bf16(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
                const int kb = (cq * 64 + c * 16) >> 3;
                *reinterpret_cast<uint4 *>(trow + kb * kaLBO) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4 *>(trow + (kb + 1) * kaLBO) = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
        tc_fence_before();
        __syncthreads();
        // ---- P7b: dZ1 = dH1 * (1 - H1^2);  dW1 += dZ1^T x,  db1 += sum dZ1  (fully unrolled: h1pre stays in registers)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint2 dr = *reinterpret_cast<const uint2 *>(As + i * kaSBO + slot0);
            const uint2 hr = h1pre[i];
            const float2 h01 = make_float2(bf16_lo(hr.x), bf16_hi(hr.x)), h23 = make_float2(bf16_lo(hr.y), bf16_hi(hr.y));
            const float2 one = make_float2(1.f, 1.f);
            const float2 s01 = __ffma2_rn(make_float2(-h01.x, -h01.y), h01, one), s23 = __ffma2_rn(make_float2(-h23.x, -h23.y), h23, one);
            const float2 d01 = __fmul2_rn(make_float2(bf16_lo(dr.x), bf16_hi(dr.x)), s01);
            const float2 d23 = __fmul2_rn(make_float2(bf16_lo(dr.y), bf16_hi(dr.y)), s23);
            gb1_01 = __fadd2_rn(gb1_01, d01); gb1_23 = __fadd2_rn(gb1_23, d23);
            const float *xr = xs + (rgrp + 8 * i) * D;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float2 xx = make_float2(xr[k], xr[k]);
                gW101[k] = __ffma2_rn(d01, xx, gW101[k]);
                gW123[k] = __ffma2_rn(d23, xx, gW123[k]);
            }
        }
        __syncthreads();                                   // tile and xs are free for the next iteration
    }

    // ---- flush: reduce the 8 row-group partials per column through shared memory, then one atomic per value
    float *red = reinterpret_cast<float *>(smem);          // [8 rgrp][NV][256] floats <= 128 KB (the W2 region; all MMAs are done)
    {
        auto put = [&](int v, float2 lo, float2 hi) {
            *reinterpret_cast<float4 *>(red + ((rgrp * NV + v) * H + c0)) = make_float4(lo.x, lo.y, hi.x, hi.y);
        };
#pragma unroll
        for (int a = 0; a < NOUT; ++a) put(a, gWh01[a], gWh23[a]);
        put(NOUT, gb2_01, gb2_23);
#pragma unroll
        for (int k = 0; k < D; ++k) put(NOUT + 1 + k, gW101[k], gW123[k]);
        put(NOUT + 1 + D, gb1_01, gb1_23);
    }
    __syncthreads();
    for (int e = tid; e < NV * H; e += NT) {
        const int v = e / H, col = e - v * H;
        float s = 0.0f;
#pragma unroll
        for (int g = 0; g < 8; ++g) s += red[(g * NV + v) * H + col];
        float *dst;
        if (v < NOUT) dst = p.gWh + v * H + col;
        else if (v == NOUT) dst = p.gB2 + col;
        else if (v <= NOUT + D) dst = p.gW1 + col * D + (v - NOUT - 1);
        else dst = p.gB1 + col;
        atomicAdd(dst, s);
    }
    if (tid < 16) {
        const float s = scal[tid];
        if (tid >= 8) { if (tid - 8 < NOUT) atomicAdd(p.gBh + (tid - 8), s); }
        else if (PI) {
            // stats: pg_loss, value_loss, entropy_loss, approx_kl, clip_fraction, loss
            if (tid == 0) { atomicAdd(p.stats + 0, s * p.inv_rows); atomicAdd(p.stats + 5, s * p.inv_rows); }
            if (tid == 1) { atomicAdd(p.stats + 2, s * p.inv_rows); atomicAdd(p.stats + 5, p.ent_coef * s * p.inv_rows); }
            if (tid == 2) atomicAdd(p.stats + 3, s * p.inv_rows);
            if (tid == 3) atomicAdd(p.stats + 4, s * p.inv_rows);
        } else if (tid == 0) {
            atomicAdd(p.stats + 1, s * p.inv_rows); atomicAdd(p.stats + 5, p.vf_coef * s * p.inv_rows);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem_base);
}

// --------------------------------------------------------- G[256,256] += X[rows,256]^T . Y[rows,256]
// MN-major operands: 64-row chunks of X and Y are copied row-major -> padded K-major-by-row tiles with cp.async
// (16 bytes per thread, no registers, no transposes) through a 3-stage ring; the tensor core reads both tiles
// through MN-major descriptors.  256x256 fp32 accumulators in all 512 TMEM columns; split-K over CTAs finished
// with red.global.add.v4.f32.
static constexpr int kmRows = 64;
static constexpr uint32_t kmTile = (kmRows / 8) * kaSBO;       // 36864 B: 64 rows x 256 columns (padded)
static constexpr uint32_t kmStage = 2 * kmTile;                // X and Y
static constexpr int kmStages = 3;
static constexpr uint32_t kWgradMnSmem = kmStages * kmStage + 128;   // 221312 (>= the 128 KB epilogue stage)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(256, 1)
tc_wgrad_mn_kernel(const __nv_bfloat16 *__restrict__ X, const __nv_bfloat16 *__restrict__ Y, float *__restrict__ G, int64_t rows) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + kmStages * kmStage);      // bar[0..2]
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(smem + kmStages * kmStage + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t nchunks = (rows + kmRows - 1) / kmRows;
    if ((int64_t)blockIdx.x >= nchunks) return;

    if (warp == 0) tmem_alloc<512>(tmem_holder);
    if (tid == 32) { for (int s = 0; s < kmStages; ++s) mbar_init(bar + s, 1); fence_barrier_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t idesc = make_idesc_major(128, 256, 1, 1);
    const uint32_t s_addr = smem_u32(smem);
    const int64_t stride = gridDim.x;
    const int64_t my_chunks = (nchunks - blockIdx.x + stride - 1) / stride;

    auto issue = [&](int64_t j) {                          // async copy of this CTA's j-th chunk into stage j % 3
        const int64_t row0 = (blockIdx.x + j * stride) * kmRows;
        const uint32_t st = s_addr + (uint32_t)(j % kmStages) * kmStage;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = warp + 8 * i;
            const bool ok = row0 + r < rows;
            const int64_t g = (ok ? row0 + r : 0) * H + lane * 8;
            const uint32_t d = st + (r >> 3) * kaSBO + lane * kaLBO + (r & 7) * 16;
            cp_async16(d, X + g, ok ? 16u : 0u);
            cp_async16(d + kmTile, Y + g, ok ? 16u : 0u);
        }
    };
    uint32_t ph[kmStages] = {0u, 0u, 0u};
    issue(0);
    cp_async_commit();
    if (my_chunks > 1) issue(1);
    cp_async_commit();
    for (int64_t j = 0; j < my_chunks; ++j) {
        const int s = (int)(j % kmStages);
        cp_async_wait<1>();                                // this thread's copies of chunk j have landed
        fence_proxy_async();
        __syncthreads();                                   // ... and everybody else's
        if (tid == 0) {
            tc_fence_after();
            const uint32_t x_addr = s_addr + s * kmStage, y_addr = x_addr + kmTile;
#pragma unroll
            for (int mh = 0; mh < 2; ++mh)                 // output rows (X columns) 0..127 / 128..255 -> TMEM columns 0..255 / 256..511
#pragma unroll
                for (int kk = 0; kk < kmRows / 16; ++kk)   // K = 16 rows of the chunk = two 8-row groups
                    umma_bf16(tmem_base + mh * 256, make_desc_raw(x_addr + mh * 16 * kaLBO + kk * 2 * kaSBO, kaSBO, kaLBO),
                              make_desc_raw(y_addr + kk * 2 * kaSBO, kaSBO, kaLBO), idesc, (j == 0 && kk == 0) ? 0u : 1u);
            umma_commit(bar + s);
        }
        // refill the stage that chunk j-1 used (its MMAs were issued one iteration ago) with chunk j+2
        if (j + 2 < my_chunks) {
            if (j >= 1) { const int sp = (int)((j + 2) % kmStages); mbar_wait(bar + sp, ph[sp]); ph[sp] ^= 1u; }
            issue(j + 2);
        }
        cp_async_commit();
    }
    // drain: wait for the last commit of every stage that still has one pending
    for (int64_t j = (my_chunks > 3 ? my_chunks - 3 : 0); j < my_chunks; ++j) {
        const bool waited = (j + 3 < my_chunks);           // already consumed by a refill wait
        if (!waited) { const int s = (int)(j % kmStages); mbar_wait(bar + s, ph[s]); ph[s] ^= 1u; }
    }
    tc_fence_after();
    // epilogue: TMEM -> registers -> smem stage (fp32 [128][256], 16-byte chunks XOR-swizzled by row) ->
    // coalesced red.global.add.v4.f32
#pragma unroll 1
    for (int mh = 0; mh < 2; ++mh) {
        __syncthreads();
        {
            const int rt = (warp & 3) * 32 + lane;
            const int colbase = (warp >> 2) * 128;
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mh * 256 + colbase);
            uint8_t *srow = smem + rt * 1024;
#pragma unroll 2
            for (int c0 = 0; c0 < 128; c0 += 16) {
                uint32_t acc[16];
                tmem_ld16(taddr + c0, acc);
                const int ch = (colbase + c0) >> 2;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4 *>(srow + (((ch + q) ^ (rt & 7)) << 4)) = make_uint4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const int r = warp + 8 * i;
            float *grow = G + (int64_t)(mh * 128 + r) * H;
#pragma unroll
            for (int hseg = 0; hseg < 2; ++hseg) {
                const int ch = hseg * 32 + lane;
                const float4 v = *reinterpret_cast<const float4 *>(smem + r * 1024 + ((ch ^ (r & 7)) << 4));
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(grow + ch * 4), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------ descriptor probe (tests only)
// One CTA, A bf16 [128][256], B bf16 [256][256] staged exactly like the production tiles:
//   mode 0: out[128][256] = A . B^T   (both K-major)
//   mode 1: out[128][256] = A . B     (B tile read through the MN-major descriptor)
//   mode 2: out[256][256] = A^T . B[0:128]   (both MN-major; the wgrad shape)
__global__ void __launch_bounds__(256, 1)
tc_probe_kernel(const __nv_bfloat16 *__restrict__ A, const __nv_bfloat16 *__restrict__ B, float *__restrict__ out, int mode) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *Ws = smem, *As = smem + kWBytes;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + kWBytes + kABytes);
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(smem + kWBytes + kABytes + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc<512>(tmem_holder);
    if (tid == 32) { mbar_init(bar, 1); fence_barrier_init(); }
    stage_rows<H>(Ws, B, 0, H);
    for (int i = 0; i < 16; ++i) {
        const int r = warp + 8 * i;
        *reinterpret_cast<uint4 *>(As + (r >> 3) * kaSBO + lane * kaLBO + (r & 7) * 16) = __ldg(reinterpret_cast<const uint4 *>(A + r * H) + lane);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t a_addr = smem_u32(As), w_addr = smem_u32(Ws);
    if (tid == 0) {
        if (mode == 0) {
            for (int kk = 0; kk < 16; ++kk)
                umma_bf16(tmem_base, make_desc_raw(a_addr + kk * 2 * kaLBO, kaLBO, kaSBO), make_desc_raw(w_addr + kk * 2 * kLBO, kLBO, kSBO),
                          make_idesc_major(128, 256, 0, 0), kk > 0);
        } else if (mode == 1) {
            for (int kk = 0; kk < 16; ++kk)
                umma_bf16(tmem_base, make_desc_raw(a_addr + kk * 2 * kaLBO, kaLBO, kaSBO), make_desc_raw(w_addr + kk * 2 * kSBO, kSBO, kLBO),
                          make_idesc_major(128, 256, 0, 1), kk > 0);
        } else {
            for (int mh = 0; mh < 2; ++mh)
                for (int kk = 0; kk < 8; ++kk)
                    umma_bf16(tmem_base + mh * 256, make_desc_raw(a_addr + mh * 16 * kaLBO + kk * 2 * kaSBO, kaSBO, kaLBO),
                              make_desc_raw(w_addr + kk * 2 * kSBO, kSBO, kLBO), make_idesc_major(128, 256, 1, 1), kk > 0);
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    const int rt = (warp & 3) * 32 + lane, colbase = (warp >> 2) * 128;
    for (int mh = 0; mh < (mode == 2 ? 2 : 1); ++mh)
        for (int c = 0; c < 8; ++c) {
            uint32_t acc[16];
            tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mh * 256 + colbase + c * 16), acc);
            for (int j = 0; j < 16; ++j) out[(int64_t)(mh * 128 + rt) * H + colbase + c * 16 + j] = __uint_as_float(acc[j]);
        }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------ host side
int tc_wgrad_launch(const void *X, const void *Y, float *G, int64_t rows, cudaStream_t st);   // mlp_tc.cu (PRMT-transposing variant)

static int sm_count_train() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}

static int g_wgrad_impl = 1;                               // 1 = MN-major cp.async kernel, 0 = PRMT-transposing kernel (mlp_tc.cu)

int tc_wgrad_mn_launch(const void *X, const void *Y, float *G, int64_t rows, cudaStream_t st) {
    static int attr_done = 0;
    if (!attr_done) {
        TMLA_CUDA(cudaFuncSetAttribute(tc_wgrad_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgradMnSmem));
        attr_done = 1;
    }
    const unsigned grid = (unsigned)std::min<int64_t>((rows + kmRows - 1) / kmRows, sm_count_train());
    tc_wgrad_mn_kernel<<<grid, 256, kWgradMnSmem, st>>>((const __nv_bfloat16 *)X, (const __nv_bfloat16 *)Y, G, rows);
    TMLA_LAUNCH_CHECK();
    return TMLA_OK;
}

template <int D, int NOUT>
static int tower_train_launch_t(const TowerTrainArgs &a, cudaStream_t st) {
    static int attr_done = 0;
    constexpr uint32_t smem = TrainSmem<D, NOUT>::total;
    static_assert(smem <= 232448, "fused tower training kernel exceeds the 227 KB shared-memory limit");
    static_assert((NOUT + 1 + D + 1) * 8 * H * 4 <= (int)kWBytes, "gradient flush stage must fit in the W2 region");
    if (!attr_done) {
        TMLA_CUDA(cudaFuncSetAttribute(tc_tower_train_kernel<D, NOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = 1;
    }
    const unsigned grid = (unsigned)std::min<int64_t>((a.M + 127) / 128, sm_count_train());
    tc_tower_train_kernel<D, NOUT><<<grid, kTrainThreads, smem, st>>>(a);
    TMLA_LAUNCH_CHECK();
    return TMLA_OK;
}

extern "C" {

int tmla_ppo_minibatch_supported(int obs_dim, int hidden, int n_actions) {
    return (hidden == H && (obs_dim == 4 || obs_dim == 6) && n_actions == 5) ? 1 : 0;
}

int64_t tmla_ppo_minibatch_scratch(int hidden, int64_t rows) { return 2 * rows * (int64_t)hidden; }

int tmla_ppo_minibatch_bf16(const float *params, const void *wpack, int obs_dim, int hidden, int n_actions, const float *obs,
                            const int32_t *index, int64_t rows, int64_t global_rows, const int32_t *actions,
                            const float *advantages, const float *old_logp, const float *returns, const double *adv_sums,
                            int normalize_advantage, float clip_range, float ent_coef, float vf_coef, float *grads,
                            void *scratch, float *stats_out, float *logits_out, float *values_out, void *stream) {
    TMLA_REQUIRE(params && wpack && obs && actions && advantages && old_logp && returns && grads && scratch && stats_out, "NULL buffer");
    TMLA_REQUIRE(rows > 0 && global_rows >= rows, "bad row counts");
    TMLA_REQUIRE(!normalize_advantage || adv_sums, "adv_sums required when normalising");
    if (!tmla_ppo_minibatch_supported(obs_dim, hidden, n_actions)) {
        tmla_set_error("tmla_ppo_minibatch_bf16: fused path covers hidden=256, obs_dim 4 or 6, 5 actions (got %d/%d/%d)", obs_dim, hidden, n_actions);
        return TMLA_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const MlpOffsets o = mlp_offsets(obs_dim, n_actions);
    TMLA_CUDA(cudaMemsetAsync(grads, 0, sizeof(float) * o.total, st));
    TMLA_CUDA(cudaMemsetAsync(stats_out, 0, 8 * sizeof(float), st));
    __nv_bfloat16 *h1 = reinterpret_cast<__nv_bfloat16 *>(scratch), *dz2 = h1 + rows * H;
    for (int t = 0; t < 2; ++t) {
        TowerTrainArgs a;
        a.W1 = params + o.w1[t]; a.B1 = params + o.b1[t];
        a.W2 = reinterpret_cast<const __nv_bfloat16 *>(wpack) + (int64_t)(2 * t) * H * H;
        a.B2 = params + o.b2[t]; a.Wh = params + o.wh[t]; a.Bh = params + o.bh[t];
        a.x = obs; a.index = index; a.M = rows;
        a.actions = actions; a.adv = advantages; a.old_logp = old_logp; a.returns = returns;
        a.adv_sums = adv_sums; a.normalize = normalize_advantage;
        a.clip = clip_range; a.ent_coef = ent_coef; a.vf_coef = vf_coef; a.inv_rows = (float)(1.0 / (double)global_rows);
        a.h1_out = h1; a.dz2_out = dz2; a.out = t == 0 ? logits_out : values_out;
        a.gW1 = grads + o.w1[t]; a.gB1 = grads + o.b1[t]; a.gB2 = grads + o.b2[t]; a.gWh = grads + o.wh[t]; a.gBh = grads + o.bh[t];
        a.stats = stats_out;
        int rc;
        if (obs_dim == 6) rc = t == 0 ? tower_train_launch_t<6, 5>(a, st) : tower_train_launch_t<6, 1>(a, st);
        else rc = t == 0 ? tower_train_launch_t<4, 5>(a, st) : tower_train_launch_t<4, 1>(a, st);
        if (rc) return rc;
        rc = g_wgrad_impl ? tc_wgrad_mn_launch(dz2, h1, grads + o.w2[t], rows, st) : tc_wgrad_launch(dz2, h1, grads + o.w2[t], rows, st);
        if (rc) return rc;
    }
    return TMLA_OK;
}

int tmla_tc_wgrad_mn(const void *X, const void *Y, float *G, int64_t rows, void *stream) {
    TMLA_REQUIRE(X && Y && G && rows > 0, "bad arguments");
    return tc_wgrad_mn_launch(X, Y, G, rows, (cudaStream_t)stream);
}

int tmla_tc_wgrad_select(int impl) {
    g_wgrad_impl = impl ? 1 : 0;
    return TMLA_OK;
}

int tmla_tc_probe(const void *A, const void *B, float *out, int mode, void *stream) {
    TMLA_REQUIRE(A && B && out && mode >= 0 && mode <= 2, "bad arguments");
    static int attr_done = 0;
    const int smem = (int)(kWBytes + kABytes + 64);
    if (!attr_done) {
        TMLA_CUDA(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_done = 1;
    }
    tc_probe_kernel<<<1, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16 *)A, (const __nv_bfloat16 *)B, out, mode);
    TMLA_LAUNCH_CHECK();
    return TMLA_OK;
}

}  // extern "C"
