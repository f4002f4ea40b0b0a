// lfm_b200 - tcgen05 attention for the DiT token grid: T = 256 tokens, head_dim = 64, non-causal, no mask.
//
// Replaces timm Attention's softmax(q k^T / sqrt(dh)) v  (reference models/DiT.py:120 via timm).
// Input: the qkv GEMM's output [B*T, 3*D] bf16 (feature index = which*D + head*64 + d), read in place by TMA.
// Output: [B*T, D] bf16 (feature index = head*64 + d), i.e. the ".transpose(1,2).reshape(B,N,C)" layout.
//
// One CTA = one (sample, head, 128-query tile).  Whole K and V of the head (256 x 64) sit in shared memory.
//   S[128 x 256] = Q K^T        4 x tcgen05.mma 128x256x16, SS, both K-major          -> TMEM cols [0,256)
//   softmax                     1 thread per query row (TMEM lane), fp32, exp2         -> P (bf16)
//   O[128 x 64]  = P V          16 x tcgen05.mma 128x64x16; V is the MN-major B operand -> TMEM cols [128,192)
// P is either written back to TMEM over the dead S columns (A-from-TMEM "TS" MMA, P_TMEM = true: 80 KB smem,
// 256 TMEM columns => 2 CTAs per SM so one CTA's softmax overlaps the other's loads/MMAs), or staged through
// 128B-swizzled shared memory (P_TMEM = false: +64 KB smem).
//
// 160 threads: warps 0..3 = softmax/epilogue (TMEM lane quadrant = warp), warp 4 = TMA + MMA issue + TMEM alloc.
#pragma once
#include "common.cuh"

namespace lfm {

constexpr int kAttnT = 256;
constexpr int kAttnDh = 64;
constexpr int kAttnThreads = 160;
constexpr int kAttnQBytes = 128 * 64 * 2;
constexpr int kAttnKVBytes = 256 * 64 * 2;
constexpr int kAttnPBytes = 128 * 256 * 2;

template <bool P_TMEM>
constexpr int attn_smem_bytes() {
    return kAttnQBytes + 2 * kAttnKVBytes + (P_TMEM ? 0 : kAttnPBytes) + 1024 + 128;
}

LFM_DEVICE float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool P_TMEM>
__global__ void __launch_bounds__(kAttnThreads)
attention_t256_d64(const __grid_constant__ CUtensorMap tmap_q,   // qkv [M, 3D], box {64, 128}
                   const __grid_constant__ CUtensorMap tmap_kv,  // qkv [M, 3D], box {64, 256}
                   __nv_bfloat16* __restrict__ out, int D, float scale_log2e, float* __restrict__ dbg_s) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = smem + kAttnQBytes;
    uint8_t* sV = sK + kAttnKVBytes;
    uint8_t* sP = sV + kAttnKVBytes;  // only when !P_TMEM
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kAttnKVBytes + (P_TMEM ? 0 : kAttnPBytes));
    uint64_t* bar_qk = bars + 0;
    uint64_t* bar_v = bars + 1;
    uint64_t* bar_s = bars + 2;
    uint64_t* bar_p = bars + 3;
    uint64_t* bar_o = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int H = gridDim.y;
    const int tok0 = b * kAttnT;

    if (warp == 4) {
        if (lane == 0) {
            prefetch_tmap(&tmap_q);
            prefetch_tmap(&tmap_kv);
            mbar_init(bar_qk, 1);
            mbar_init(bar_v, 1);
            mbar_init(bar_s, 1);
            mbar_init(bar_p, 128);
            mbar_init(bar_o, 1);
            fence_barrier_init();
            fence_proxy_async();
        }
        __syncwarp();
        tmem_alloc<256>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t kColO = 128;

    if (warp == 4) {
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_qk, kAttnQBytes + kAttnKVBytes);
            tma_load_2d(sQ, &tmap_q, bar_qk, h * kAttnDh, tok0 + qt * 128);
            tma_load_2d(sK, &tmap_kv, bar_qk, D + h * kAttnDh, tok0);
            mbar_arrive_expect_tx(bar_v, kAttnKVBytes);
            tma_load_2d(sV, &tmap_kv, bar_v, 2 * D + h * kAttnDh, tok0);

            // S = Q K^T
            mbar_wait(bar_qk, 0);
            tc_fence_after();
            {
                constexpr uint32_t idesc = make_idesc_bf16(128, 256, 0, 0);
                const uint64_t dq = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
                const uint64_t dk = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ss(tmem, dq + 2 * k, dk + 2 * k, idesc, k != 0);
                umma_commit(bar_s);
            }
            // O = P V
            mbar_wait(bar_p, 0);
            tc_fence_after();
            mbar_wait(bar_v, 0);
            tc_fence_after();
            {
                constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    // V[k*16 .. k*16+15][0..63]: two 8-token swizzle atoms of 1024 B
                    const uint64_t dv = make_smem_desc_sw128(smem_u32(sV + k * 2048), 1024, 1024);
                    if (P_TMEM) {
                        umma_ts(tmem + kColO, tmem + k * 8, dv, idesc, k != 0);
                    } else {
                        const uint64_t dp = make_smem_desc_sw128(smem_u32(sP + (k >> 2) * 16384), 16, 1024) + 2 * (k & 3);
                        umma_ss(tmem + kColO, dp, dv, idesc, k != 0);
                    }
                }
                umma_commit(bar_o);
            }
        }
    } else {
        const int r = warp * 32 + lane;  // query row inside the tile == TMEM lane
        const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        mbar_wait(bar_s, 0);
        tc_fence_after();
        // pass 1: row max
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(taddr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
            if (dbg_s != nullptr) {
                float* dp = dbg_s + (static_cast<size_t>((b * H + h) * kAttnT + qt * 128 + r)) * kAttnT + c * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j) dp[j] = __uint_as_float(v[j]);
            }
        }
        const float mxs = mx * scale_log2e;
        // pass 2: p = exp2(s*scale*log2e - max), row sum, P -> TMEM / smem as bf16
        float sum = 0.f;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(taddr + c * 32, v);
            tmem_ld_wait();
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), scale_log2e, -mxs));
                const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), scale_log2e, -mxs));
                sum += p0 + p1;
                pk[j] = pack_bf16x2(p0, p1);
            }
            if (P_TMEM) {
                tmem_st_32x32b_x16(taddr + c * 16, pk);
            } else {
                // K-major, 128B swizzle: k-block (64 keys) = c/2, 16-byte chunk = (c%2)*4 + jj, XOR (r % 8)
                uint8_t* rowp = sP + (c >> 1) * 16384 + r * 128;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int chunk = ((c & 1) * 4 + jj) ^ (r & 7);
                    *reinterpret_cast<uint4*>(rowp + chunk * 16) =
                        make_uint4(pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
                }
            }
        }
        if (P_TMEM) {
            tmem_st_wait();
            tc_fence_before();
        } else {
            fence_proxy_async();
        }
        mbar_arrive(bar_p);

        // epilogue: O / sum -> bf16 -> out[(tok0 + qt*128 + r), h*64 .. +63]
        mbar_wait(bar_o, 0);
        tc_fence_after();
        const float inv = 1.0f / sum;
        __nv_bfloat16* orow = out + static_cast<size_t>(tok0 + qt * 128 + r) * D + h * kAttnDh;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(taddr + kColO + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]) * inv, __uint_as_float(v[8 * j + 1]) * inv);
                o.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv);
                o.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv);
                o.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv);
                reinterpret_cast<uint4*>(orow + c * 32)[j] = o;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc<256>(tmem);
    }
}

}  // namespace lfm
