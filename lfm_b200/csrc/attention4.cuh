// lfm_b200 - persistent tcgen05 attention for the DiT token grid (T = 256, head_dim = 64), version 4 (EXPERIMENTAL:
// selected only by LFM_ATTN_VARIANT=4 / lfm_dbg_attention(variant = 4); the default path is attention3.cuh).
//
// Motivation (profiles/r1f_attention3_full.md): version 3 is latency-bound - two softmax warps per SM sub-partition,
// issue slots 29 % busy, MUFU 37 %, warps stalled on the TMEM -> register path.  Version 4 doubles the softmax
// warps: every query tile g gets TWO warpgroups, one per key half h (keys [128 h, 128 h + 128)), so four warps share a
// sub-partition and the TMEM loads of one hide behind the exponentials of the others.  The two halves were already
// independent in version 3 (own row max m_h, own unnormalised P_h, own accumulator O_h = P_h V_h, merged in the
// epilogue), so the split needs no exchange during the softmax; only the per-row (m_h, l_h) pairs cross over through
// shared memory for the merge.  To stay inside the register file (20 warps) a half is walked in 32-column pieces,
// twice: a max pass and an exp pass (the S values are re-read from TMEM instead of being held in 128 registers).
//
// Roles (640 threads): warps 0-3 (g=0,h=0), 4-7 (g=0,h=1), 8-11 (g=1,h=0), 12-15 (g=1,h=1) softmax / epilogue
// (TMEM lane quadrant = warp % 4); warp 16 TMA loader; warp 17 MMA issuer + TMEM allocation; warps 18-19 idle.
// TMEM columns of tile g (base 256 g): S = Q_g K^T in [0,256); then P_0 -> [0,64), O_0 -> [64,128), P_1 -> [128,192),
// O_1 -> [192,256): warpgroup (g, h) only ever touches [128 h, 128 h + 128) until both accumulators are complete.
#pragma once
#include "attention.cuh"
#include "attention2.cuh"
#include "common.cuh"

namespace lfm {

constexpr int kA4Threads = 640;

__global__ void __launch_bounds__(kA4Threads, 1)
attention4_t256_d64(const __grid_constant__ CUtensorMap tmap_kv,   // qkv [M, 3D] bf16, box {64, 256}
                    const __grid_constant__ CUtensorMap tmap_out,  // out [M, D]  bf16, box {64, 128}
                    int D, int H, int num_items, float scale_log2e, int reverse) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ float2 s_stats[2][2][128];  // [tile g][half h][row] = {m_h * scale_log2e, l_h}
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kA2StageBytes);
    uint64_t* full = bars;          // [2] loader -> MMA
    uint64_t* empty = bars + 2;     // [2] MMA commit + one arrival per query tile (its TMA store has drained the slot)
    uint64_t* s_full = bars + 4;    // [2] MMA -> the two softmax warpgroups of tile g
    uint64_t* p_full = bars + 6;    // [2][2] warpgroup (g, h) -> MMA (128 arrivals): P_h written
    uint64_t* o_full = bars + 10;   // [2] MMA -> warpgroups of tile g: O_0 and O_1 complete
    uint64_t* s_empty = bars + 12;  // [2] warpgroups of tile g -> MMA (256 arrivals): TMEM region g is free again
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 17) {
        if (lane == 0) {
            prefetch_tmap(&tmap_kv);
            prefetch_tmap(&tmap_out);
            for (int i = 0; i < 2; ++i) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], 3);
                mbar_init(&s_full[i], 1);
                mbar_init(&p_full[2 * i], 128);
                mbar_init(&p_full[2 * i + 1], 128);
                mbar_init(&o_full[i], 1);
                mbar_init(&s_empty[i], 256);
            }
            fence_barrier_init();
            fence_proxy_async();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    pdl_trigger();
    const uint32_t tmem = *tmem_slot;

    if (warp == 16) {
        // ===================== TMA loader =====================
        if (lane == 0) {
            int i = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++i) {
                const int stage = i & 1;
                const int it = reverse ? num_items - 1 - item : item;
                const int b = it / H, h = it % H;
                uint8_t* st = smem + stage * kA2StageBytes;
                mbar_wait(&empty[stage], ((i >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&full[stage], kA2StageBytes);
                tma_load_2d(st, &tmap_kv, &full[stage], h * kAttnDh, b * kAttnT);                          // Q
                tma_load_2d(st + kAttnKVBytes, &tmap_kv, &full[stage], D + h * kAttnDh, b * kAttnT);       // K
                tma_load_2d(st + 2 * kAttnKVBytes, &tmap_kv, &full[stage], 2 * D + h * kAttnDh, b * kAttnT);  // V
            }
        }
    } else if (warp == 17) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(128, 256, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);  // B (= V) MN-major
            int i = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++i) {
                const int stage = i & 1;
                const uint32_t hp = i & 1;
                uint8_t* st = smem + stage * kA2StageBytes;
                mbar_wait(&full[stage], (i >> 1) & 1);
                tc_fence_after();
                const uint64_t dk = make_smem_desc_sw128(smem_u32(st + kAttnKVBytes), 16, 1024);
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    mbar_wait(&s_empty[g], hp ^ 1);
                    tc_fence_after();
                    const uint64_t dq = make_smem_desc_sw128(smem_u32(st + g * kAttnQBytes), 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_ss(tmem + g * 256, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                    umma_commit(&s_full[g]);
                }
                // O_h = P_h V[128 h : 128 h + 128]; P_h at columns 128 h + [0,64), O_h at 128 h + [64,128)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        mbar_wait(&p_full[2 * g + hh], hp);
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint64_t dv =
                                make_smem_desc_sw128(smem_u32(st + 2 * kAttnKVBytes + (8 * hh + k) * 2048), 1024, 1024);
                            umma_ts(tmem + g * 256 + 128 * hh + 64, tmem + g * 256 + 128 * hh + k * 8, dv, idesc_o, k != 0);
                        }
                        if (hh == 1) umma_commit(&o_full[g]);
                    }
                }
                umma_commit(&empty[stage]);  // K, V (and Q) of this stage are no longer read by the tensor core
            }
        }
    } else if (warp < 16) {
        // ===================== softmax + epilogue warpgroups =====================
        const int g = warp >> 3;         // query tile
        const int hh = (warp >> 2) & 1;  // key half
        const int r = (warp & 3) * 32 + lane;
        const uint32_t taddr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + g * 256;
        const uint32_t thalf = taddr + 128 * hh;
        int i = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++i) {
            const int stage = i & 1;
            const uint32_t hp = i & 1;
            const int it = reverse ? num_items - 1 - item : item;
            const int b = it / H, h = it % H;
            uint8_t* sO = smem + stage * kA2StageBytes + g * kAttnQBytes;  // Q_g's slot, reused for the output tile
            mbar_wait(&s_full[g], hp);
            tc_fence_after();
            // pass 1: row max over this half's 128 keys, 32 columns at a time (the next piece loads while this one is reduced)
            float mx = -INFINITY;
            {
                uint32_t va[32], vb[32];
                tmem_ld_32x32b_x32(thalf, va);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(thalf + 32, vb);
#pragma unroll
                for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(va[j]));
                tmem_ld_wait();
                tmem_ld_32x32b_x32(thalf + 64, va);
#pragma unroll
                for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(vb[j]));
                tmem_ld_wait();
                tmem_ld_32x32b_x32(thalf + 96, vb);
#pragma unroll
                for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(va[j]));
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(vb[j]));
            }
            const float ms = mx * scale_log2e;
            // pass 2: P = exp2(S * scale - m) -> bf16 pairs -> columns [16 c, 16 c + 16) of this half (already consumed S)
            float sum = 0.f;
            {
                uint32_t v[32];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    tmem_ld_32x32b_x32(thalf + c * 32, v);
                    tmem_ld_wait();
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), scale_log2e, -ms));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), scale_log2e, -ms));
                        sum += p0 + p1;
                        pk[j] = pack_bf16x2(p0, p1);
                    }
                    tmem_st_32x32b_x16(thalf + c * 16, pk);
                }
            }
            s_stats[g][hh][r] = make_float2(ms, sum);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[2 * g + hh]);

            // epilogue: this warpgroup merges output columns [32 hh, 32 hh + 32) of both accumulators
            mbar_wait(&o_full[g], hp);
            tc_fence_after();
            asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");  // the other half's (m, l) are in shared memory
            const float2 so = s_stats[g][hh ^ 1][r];
            const float mA = hh == 0 ? ms : so.x, mB = hh == 0 ? so.x : ms;
            const float lA = hh == 0 ? sum : so.y, lB = hh == 0 ? so.y : sum;
            const float mm = fmaxf(mA, mB);
            const float aA = ex2_approx(mA - mm), aB = ex2_approx(mB - mm);
            const float inv = 1.0f / fmaf(aA, lA, aB * lB);
            const float wA = aA * inv, wB = aB * inv;
            {
                uint8_t* rowp = sO + r * 128;
                uint32_t va[32], vb[32];
                tmem_ld_32x32b_x32(taddr + 64 + hh * 32, va);    // O_0 columns [32 hh, 32 hh + 32)
                tmem_ld_32x32b_x32(taddr + 192 + hh * 32, vb);   // O_1
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        o[e] = fmaf(wA, __uint_as_float(va[8 * j + e]), wB * __uint_as_float(vb[8 * j + e]));
                    uint4 u;
                    u.x = pack_bf16x2(o[0], o[1]);
                    u.y = pack_bf16x2(o[2], o[3]);
                    u.z = pack_bf16x2(o[4], o[5]);
                    u.w = pack_bf16x2(o[6], o[7]);
                    *reinterpret_cast<uint4*>(rowp + (((hh * 4 + j) ^ (r & 7)) << 4)) = u;
                }
            }
            tc_fence_before();
            mbar_arrive(&s_empty[g]);  // TMEM region g may be overwritten by the next head's S (256 arrivals)
            fence_proxy_async();
            // tile-wide barrier (both warpgroups of tile g), then one thread issues the TMA store
            asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
            if ((warp & 7) == 0 && lane == 0) {
                tma_store_2d(&tmap_out, sO, h * kAttnDh, b * kAttnT + g * 128);
                tma_store_commit();
                tma_store_wait_read<0>();   // the slot can be refilled by the loader
                mbar_arrive(&empty[stage]);
            }
        }
        if ((warp & 7) == 0 && lane == 0) tma_store_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 17) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

}  // namespace lfm
