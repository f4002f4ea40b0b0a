// lfm_b200 - persistent tcgen05 attention for the DiT token grid (T = 256, head_dim = 64), version 3:
// same structure as attention2.cuh, but every S value is read from TMEM exactly ONCE (the kernel is bound by the
// TMEM -> register path: profiles/r1c_launches.md).  The 256 keys are treated as two halves with their own row max
// (m_A, m_B) and unnormalised probabilities P_A = exp(S_A - m_A), P_B = exp(S_B - m_B); O_A = P_A V_A and
// O_B = P_B V_B accumulate in separate TMEM columns and are merged in the epilogue:
//     O = (a_A O_A + a_B O_B) / (a_A l_A + a_B l_B),  a_X = exp(m_X - max(m_A, m_B)),  l_X = rowsum(P_X)
// which equals softmax(S) V exactly.  TMEM reads per query row: 256 (S) + 128 (O_A, O_B) columns instead of
// 2 x 256 + 64; and O_A's MMAs overlap the exponentials of the second half.
//
// One persistent CTA per SM loops over (sample, head) work items; both 128-query tiles of a head are processed
// together so K and V are staged once.  Roles (320 threads):
//   warps 0-3  softmax / epilogue warpgroup for query tile 0   (TMEM lane quadrant = warp % 4)
//   warps 4-7  softmax / epilogue warpgroup for query tile 1
//   warp  8    TMA loader: Q (256 x 64), K, V of the NEXT head stream into a 2-stage shared-memory ring while the
//              current head is being processed
//   warp  9    MMA issuer (one thread) + TMEM allocation (all 512 columns)
// TMEM: S_g = Q_g K^T in columns [256 g, 256 g + 256); the bf16 probabilities P_g overwrite the first 128 of those
// columns (A operand of the TS-form P V MMA) and O_g lands in columns [256 g + 128, 256 g + 192).
// The output tile is staged in the (dead) shared-memory slot of Q_g and written with one TMA store per tile.
// Replaces timm Attention's softmax(q k^T / sqrt(dh)) v (reference models/DiT.py:120).
#pragma once
#include "attention.cuh"
#include "attention2.cuh"
#include "common.cuh"

namespace lfm {

// max of three (one FMNMX3 on sm_100)
LFM_DEVICE float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// Row maximum of 128 fp32 values held as 4 x 32 registers: eight independent 3-input chains (a serial fmaxf chain
// costs 128 dependent instructions of a warp that has nothing else to issue until the maximum is known).
LFM_DEVICE float row_max128(const uint32_t (*v)[32]) {
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = fmaxf(__uint_as_float(v[k >> 1][(k & 1) * 16]), __uint_as_float(v[k >> 1][(k & 1) * 16 + 1]));
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int j = 2; j < 16; j += 2)
            m[k] = fmax3(m[k], __uint_as_float(v[k >> 1][(k & 1) * 16 + j]), __uint_as_float(v[k >> 1][(k & 1) * 16 + j + 1]));
    return fmaxf(fmax3(m[0], m[1], m[2]), fmax3(m[3], m[4], fmax3(m[5], m[6], m[7])));
}

// X2 (LFM_ATTN_X2): the softmax / merge arithmetic in packed f32x2 instructions - per PAIR of scores one fma (scale, minus the row
// maximum), two MUFU.EX2, one add into a packed partial sum and one bf16x2 conversion (5 issue slots instead of 7); same values, the
// row sum is accumulated in a different (equally valid) order.
template <bool X2>
__global__ void __launch_bounds__(kA2Threads, 1)
attention3_t256_d64(const __grid_constant__ CUtensorMap tmap_kv,   // qkv [M, 3D] bf16, box {64, 256}
                    const __grid_constant__ CUtensorMap tmap_out,  // out [M, D]  bf16, box {64, 128}
                    int D, int H, int num_items, float scale_log2e, int reverse) {
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
                mbar_init(&empty[i], 3);
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

    if (warp == 8) {
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
    } else if (warp == 9) {
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
                // O_A = P_A V[0:128]  -> TMEM cols [128,192);  O_B = P_B V[128:256] -> cols [192,256)
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    mbar_wait(&p_full[g], hp);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint64_t dv = make_smem_desc_sw128(smem_u32(st + 2 * kAttnKVBytes + k * 2048), 1024, 1024);
                        umma_ts(tmem + g * 256 + 128, tmem + g * 256 + k * 8, dv, idesc_o, k != 0);
                    }
                }
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    mbar_wait(&pb_full[g], hp);
                    tc_fence_after();
#pragma unroll
                    for (int k = 8; k < 16; ++k) {
                        const uint64_t dv = make_smem_desc_sw128(smem_u32(st + 2 * kAttnKVBytes + k * 2048), 1024, 1024);
                        umma_ts(tmem + g * 256 + 192, tmem + g * 256 + k * 8, dv, idesc_o, k != 8);
                    }
                    umma_commit(&o_full[g]);
                }
                umma_commit(&empty[stage]);  // K, V (and Q) of this stage are no longer read by the tensor core
            }
        }
    } else {
        // ===================== softmax + epilogue warpgroups =====================
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
            float mAs, mBs, sumA, sumB;
            {
                uint32_t v[4][32];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(taddr + c * 32, v[c]);
                tmem_ld_wait();
                const float mx = row_max128(v);
                mAs = mx * scale_log2e;
                float sacc[4] = {0.f, 0.f, 0.f, 0.f};  // independent partial sums: no serial FADD chain behind the MUFU pipe
                uint64_t sacc2[4] = {0ull, 0ull, 0ull, 0ull};
                const uint64_t sc2 = pk2(scale_log2e, scale_log2e), nm2 = pk2(-mAs, -mAs);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float p0, p1;
                        if constexpr (X2) {
                            float a0, a1;
                            upk2(fma2(pk2(__uint_as_float(v[c][2 * j]), __uint_as_float(v[c][2 * j + 1])), sc2, nm2), a0, a1);
                            p0 = ex2_approx(a0);
                            p1 = ex2_approx(a1);
                            sacc2[j & 3] = add2(sacc2[j & 3], pk2(p0, p1));
                        } else {
                            p0 = ex2_approx(fmaf(__uint_as_float(v[c][2 * j]), scale_log2e, -mAs));
                            p1 = ex2_approx(fmaf(__uint_as_float(v[c][2 * j + 1]), scale_log2e, -mAs));
                            sacc[j & 3] += p0 + p1;
                        }
                        pk[j] = pack_bf16x2(p0, p1);
                    }
                    tmem_st_32x32b_x16(taddr + c * 16, pk);  // P_A -> cols [0,64) (S_A is already in registers)
                }
                if constexpr (X2) {
                    float lo, hi;
                    upk2(add2(add2(sacc2[0], sacc2[1]), add2(sacc2[2], sacc2[3])), lo, hi);
                    sumA = lo + hi;
                } else {
                    sumA = (sacc[0] + sacc[1]) + (sacc[2] + sacc[3]);
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
                float sacc[4] = {0.f, 0.f, 0.f, 0.f};
                uint64_t sacc2[4] = {0ull, 0ull, 0ull, 0ull};
                const uint64_t sc2 = pk2(scale_log2e, scale_log2e), nm2 = pk2(-mBs, -mBs);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float p0, p1;
                        if constexpr (X2) {
                            float a0, a1;
                            upk2(fma2(pk2(__uint_as_float(v[c][2 * j]), __uint_as_float(v[c][2 * j + 1])), sc2, nm2), a0, a1);
                            p0 = ex2_approx(a0);
                            p1 = ex2_approx(a1);
                            sacc2[j & 3] = add2(sacc2[j & 3], pk2(p0, p1));
                        } else {
                            p0 = ex2_approx(fmaf(__uint_as_float(v[c][2 * j]), scale_log2e, -mBs));
                            p1 = ex2_approx(fmaf(__uint_as_float(v[c][2 * j + 1]), scale_log2e, -mBs));
                            sacc[j & 3] += p0 + p1;
                        }
                        pk[j] = pack_bf16x2(p0, p1);
                    }
                    tmem_st_32x32b_x16(taddr + 64 + c * 16, pk);  // P_B -> cols [64,128)
                }
                if constexpr (X2) {
                    float lo, hi;
                    upk2(add2(add2(sacc2[0], sacc2[1]), add2(sacc2[2], sacc2[3])), lo, hi);
                    sumB = lo + hi;
                } else {
                    sumB = (sacc[0] + sacc[1]) + (sacc[2] + sacc[3]);
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
                        if constexpr (X2) {
                            const uint64_t wA2 = pk2(wA, wA), wB2 = pk2(wB, wB);
#pragma unroll
                            for (int e = 0; e < 8; e += 2)
                                upk2(fma2(wA2, pk2(__uint_as_float(va[8 * j + e]), __uint_as_float(va[8 * j + e + 1])),
                                          mul2(wB2, pk2(__uint_as_float(vb[8 * j + e]), __uint_as_float(vb[8 * j + e + 1])))),
                                     o[e], o[e + 1]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                o[e] = fmaf(wA, __uint_as_float(va[8 * j + e]), wB * __uint_as_float(vb[8 * j + e]));
                        }
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
