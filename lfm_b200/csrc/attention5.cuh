// lfm_b200 - persistent tcgen05 attention for the DiT token grid (T = 256, head_dim = 64), version 5 ("ping-pong").
//
// Same arithmetic, TMEM layout and softmax code as attention3.cuh (every S value is read from TMEM once; two key
// halves with their own row max, merged in the epilogue).  What changes is the SCHEDULE.  In version 3 one MMA
// thread walks both 128-query tiles of a head in a fixed program order, so the two softmax warpgroups run in
// lock-step: both wait for S at the same time, both hit the MUFU pipe at the same time, and the chain
//   S MMA -> ld -> exp(A) -> ld -> exp(B) -> P V -> epilogue -> (TMEM free) -> next S MMA
// is fully serialised - MUFU is busy 37 % of the time, the tensor pipe 18 % (profiles/r1f_attention3_full.md).
// Version 5 gives every query tile its OWN MMA issuer thread (tcgen05.mma from different threads execute in issue
// order on the one tensor core, the two tiles touch disjoint TMEM columns) and starts tile 1 half a period late
// (its first S MMA waits for tile 0's first P_B): the two (MMA thread, warpgroup) pipelines then run out of phase,
// the exponentials of one tile overlap the TMEM loads, P V MMAs, epilogue and next S MMA of the other.
//
// Roles (384 threads): warps 0-3 / 4-7 softmax + epilogue warpgroups of tile 0 / 1; warp 8 TMA loader (2-stage
// ring of Q, K, V); warp 9 MMA issuer of tile 0 + TMEM allocation; warp 10 MMA issuer of tile 1; warp 11 idle
// (setmaxnreg moves registers between whole warpgroups).
// A shared-memory stage is released by four arrivals: the commit of each tile's last MMA and each tile's TMA store.
#pragma once
#include "attention.cuh"
#include "attention2.cuh"
#include "attention3.cuh"
#include "common.cuh"

namespace lfm {


constexpr int kA5Threads = 384;  // 12 warps: registers are re-balanced between the warpgroups with setmaxnreg

template <int N>
LFM_DEVICE void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
LFM_DEVICE void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__global__ void __launch_bounds__(kA5Threads, 1)
attention5_t256_d64(const __grid_constant__ CUtensorMap tmap_kv,   // qkv [M, 3D] bf16, box {64, 256}
                    const __grid_constant__ CUtensorMap tmap_out,  // out [M, D]  bf16, box {64, 128}
                    int D, int H, int num_items, float scale_log2e, int reverse, int stagger) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kA2StageBytes);
    uint64_t* full = bars;          // [2] loader -> MMA
    uint64_t* empty = bars + 2;     // [2] MMA commit + one arrival per warpgroup (its TMA store has drained the slot)
    uint64_t* s_full = bars + 4;    // [2] MMA -> softmax WG g
    uint64_t* p_full = bars + 6;    // [2] softmax WG g -> MMA   (128 arrivals): P_A written AND S_B held in registers
    uint64_t* o_full = bars + 8;    // [2] MMA -> softmax WG g
    uint64_t* s_empty = bars + 10;  // [2] softmax WG g -> MMA   (128 arrivals): TMEM region g is free again
    uint64_t* pb_full = bars + 12;  // [2] softmax WG g -> MMA   (128 arrivals): P_B written
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 9) {
        if (lane == 0) {
            prefetch_tmap(&tmap_kv);
            prefetch_tmap(&tmap_out);
            for (int i = 0; i < 2; ++i) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], 4);
                mbar_init(&s_full[i], 1);
                mbar_init(&p_full[i], 128);
                mbar_init(&o_full[i], 1);
                mbar_init(&s_empty[i], 128);
                mbar_init(&pb_full[i], 128);
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

    // warps 8-11 (loader, two MMA issuers, one idle warp) need few registers; the softmax warps hold 128 S values each
    if (warp == 8) {
        setmaxnreg_dec<40>();
        // ===================== TMA loader =====================
        if (lane == 0) {
            int i = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++i) {
                const int stage = i & 1;
                const int it = reverse ? num_items - 1 - item : item;  // last samples first: see GemmEpi::reverse_m
                const int b = it / H, h = it % H;
                uint8_t* st = smem + stage * kA2StageBytes;
                mbar_wait(&empty[stage], ((i >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&full[stage], kA2StageBytes);
                tma_load_2d(st, &tmap_kv, &full[stage], h * kAttnDh, b * kAttnT);                          // Q
                tma_load_2d(st + kAttnKVBytes, &tmap_kv, &full[stage], D + h * kAttnDh, b * kAttnT);       // K
                tma_load_2d(st + 2 * kAttnKVBytes, &tmap_kv, &full[stage], 2 * D + h * kAttnDh, b * kAttnT);  // V
            }
        }
    } else if (warp == 9 || warp == 10) {
        // ===================== MMA issuer of query tile g = warp - 9 =====================
        setmaxnreg_dec<40>();
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(128, 256, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);  // B (= V) MN-major
            const int g = warp - 9;
            // stagger: tile 1 starts once tile 0 has finished its first softmax (phase 0 of pb_full[0] complete)
            if (g == 1 && stagger) mbar_wait(&pb_full[0], 0);
            int i = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++i) {
                const int stage = i & 1;
                const uint32_t hp = i & 1;
                uint8_t* st = smem + stage * kA2StageBytes;
                mbar_wait(&full[stage], (i >> 1) & 1);
                tc_fence_after();
                const uint64_t dk = make_smem_desc_sw128(smem_u32(st + kAttnKVBytes), 16, 1024);
                mbar_wait(&s_empty[g], hp ^ 1);
                tc_fence_after();
                const uint64_t dq = make_smem_desc_sw128(smem_u32(st + g * kAttnQBytes), 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ss(tmem + g * 256, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                umma_commit(&s_full[g]);
                // O_A = P_A V[0:128]  -> TMEM cols [128,192);  O_B = P_B V[128:256] -> cols [192,256)
                mbar_wait(&p_full[g], hp);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t dv = make_smem_desc_sw128(smem_u32(st + 2 * kAttnKVBytes + k * 2048), 1024, 1024);
                    umma_ts(tmem + g * 256 + 128, tmem + g * 256 + k * 8, dv, idesc_o, k != 0);
                }
                mbar_wait(&pb_full[g], hp);
                tc_fence_after();
#pragma unroll
                for (int k = 8; k < 16; ++k) {
                    const uint64_t dv = make_smem_desc_sw128(smem_u32(st + 2 * kAttnKVBytes + k * 2048), 1024, 1024);
                    umma_ts(tmem + g * 256 + 192, tmem + g * 256 + k * 8, dv, idesc_o, k != 8);
                }
                umma_commit(&o_full[g]);
                umma_commit(&empty[stage]);  // this tile's MMAs no longer read Q_g, K, V of this stage
            }
        }
    } else if (warp == 11) {
        setmaxnreg_dec<40>();
    } else {
        // ===================== softmax + epilogue warpgroups =====================
        setmaxnreg_inc<216>();
        const int g = warp >> 2;
        const int r = (warp & 3) * 32 + lane;
        const uint32_t taddr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + g * 256;
        int i = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++i) {
            const int stage = i & 1;
            const uint32_t hp = i & 1;
            const int it = reverse ? num_items - 1 - item : item;
            const int b = it / H, h = it % H;
            uint8_t* sO = smem + stage * kA2StageBytes + g * kAttnQBytes;  // Q_g's slot, reused for the output tile
            mbar_wait(&s_full[g], hp);
            tc_fence_after();
            // one half (128 keys) at a time: the half's S values live in registers, are read from TMEM once
            float mAs, mBs, sumA = 0.f, sumB = 0.f;
            {
                uint32_t v[4][32];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(taddr + c * 32, v[c]);
                tmem_ld_wait();
                const float mx = row_max128(v);
                mAs = mx * scale_log2e;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(v[c][2 * j]), scale_log2e, -mAs));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(v[c][2 * j + 1]), scale_log2e, -mAs));
                        sumA += p0 + p1;
                        pk[j] = pack_bf16x2(p0, p1);
                    }
                    tmem_st_32x32b_x16(taddr + c * 16, pk);  // P_A -> cols [0,64) (S_A is already in registers)
                }
            }
            {
                uint32_t v[4][32];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(taddr + 128 + c * 32, v[c]);
                tmem_ld_wait();   // S_B in registers: columns [128,256) may now be overwritten by O_A / O_B
                tmem_st_wait();   // P_A visible
                tc_fence_before();
                mbar_arrive(&p_full[g]);
                const float mx = row_max128(v);
                mBs = mx * scale_log2e;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(v[c][2 * j]), scale_log2e, -mBs));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(v[c][2 * j + 1]), scale_log2e, -mBs));
                        sumB += p0 + p1;
                        pk[j] = pack_bf16x2(p0, p1);
                    }
                    tmem_st_32x32b_x16(taddr + 64 + c * 16, pk);  // P_B -> cols [64,128)
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&pb_full[g]);

            // epilogue: O / sum -> bf16 -> swizzled smem tile (Q_g's slot) -> one TMA store
            mbar_wait(&o_full[g], hp);
            tc_fence_after();
            const float ms = fmaxf(mAs, mBs);
            const float aA = ex2_approx(mAs - ms), aB = ex2_approx(mBs - ms);
            const float inv = 1.0f / fmaf(aA, sumA, aB * sumB);
            const float wA = aA * inv, wB = aB * inv;
            {
                uint8_t* rowp = sO + r * 128;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t va[32], vb[32];
                    tmem_ld_32x32b_x32(taddr + 128 + c * 32, va);
                    tmem_ld_32x32b_x32(taddr + 192 + c * 32, vb);
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
                        *reinterpret_cast<uint4*>(rowp + (((c * 4 + j) ^ (r & 7)) << 4)) = u;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&s_empty[g]);  // TMEM region g may be overwritten by the next head's S
            fence_proxy_async();
            // warpgroup-wide barrier (named barrier 1 + g, 128 threads), then one thread issues the TMA store
            asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
            if ((warp & 3) == 0 && lane == 0) {
                tma_store_2d(&tmap_out, sO, h * kAttnDh, b * kAttnT + g * 128);
                tma_store_commit();
                tma_store_wait_read<0>();   // the slot can be refilled by the loader
                mbar_arrive(&empty[stage]);
            }
        }
        if ((warp & 3) == 0 && lane == 0) tma_store_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

}  // namespace lfm
