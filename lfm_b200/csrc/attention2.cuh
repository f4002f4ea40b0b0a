// lfm_b200 - persistent tcgen05 attention for the DiT token grid (T = 256, head_dim = 64), version 2.
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
#include "common.cuh"

namespace lfm {

constexpr int kA2Threads = 320;
constexpr int kA2StageBytes = 3 * kAttnKVBytes;  // Q (both tiles) + K + V = 96 KB
constexpr int kA2SmemBytes = 2 * kA2StageBytes + 1024 + 256;

__global__ void __launch_bounds__(kA2Threads, 1)
attention2_t256_d64(const __grid_constant__ CUtensorMap tmap_kv,   // qkv [M, 3D] bf16, box {64, 256}
                    const __grid_constant__ CUtensorMap tmap_out,  // out [M, D]  bf16, box {64, 128}
                    int D, int H, int num_items, float scale_log2e) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kA2StageBytes);
    uint64_t* full = bars;          // [2] loader -> MMA
    uint64_t* empty = bars + 2;     // [2] MMA commit + one arrival per warpgroup (its TMA store has drained the slot)
    uint64_t* s_full = bars + 4;    // [2] MMA -> softmax WG g
    uint64_t* p_full = bars + 6;    // [2] softmax WG g -> MMA   (128 arrivals)
    uint64_t* o_full = bars + 8;    // [2] MMA -> softmax WG g
    uint64_t* s_empty = bars + 10;  // [2] softmax WG g -> MMA   (128 arrivals): TMEM region g is free again
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

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
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        // ===================== TMA loader =====================
        if (lane == 0) {
            int i = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++i) {
                const int stage = i & 1;
                const int b = item / H, h = item % H;
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
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    mbar_wait(&p_full[g], hp);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const uint64_t dv = make_smem_desc_sw128(smem_u32(st + 2 * kAttnKVBytes + k * 2048), 1024, 1024);
                        umma_ts(tmem + g * 256 + 128, tmem + g * 256 + k * 8, dv, idesc_o, k != 0);
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
            const int b = item / H, h = item % H;
            uint8_t* sO = smem + stage * kA2StageBytes + g * kAttnQBytes;  // Q_g's slot, reused for the output tile
            mbar_wait(&s_full[g], hp);
            tc_fence_after();
            // pass 1: row max (tcgen05.ld double-buffered)
            float mx = -INFINITY;
            {
                uint32_t va[32], vb[32];
                tmem_ld_32x32b_x32(taddr, va);
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    tmem_ld_wait();
                    tmem_ld_32x32b_x32(taddr + (c + 1) * 32, vb);
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(va[j]));
                    tmem_ld_wait();
                    if (c + 2 < 8) tmem_ld_32x32b_x32(taddr + (c + 2) * 32, va);
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(vb[j]));
                }
            }
            const float mxs = mx * scale_log2e;
            // pass 2: p = exp2(s * scale * log2e - max'), row sum, P (bf16) back into TMEM over S
            float sum = 0.f;
            {
                uint32_t va[32], vb[32], pk[16];
                tmem_ld_32x32b_x32(taddr, va);
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    tmem_ld_wait();
                    tmem_ld_32x32b_x32(taddr + (c + 1) * 32, vb);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(va[2 * j]), scale_log2e, -mxs));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(va[2 * j + 1]), scale_log2e, -mxs));
                        sum += p0 + p1;
                        pk[j] = pack_bf16x2(p0, p1);
                    }
                    tmem_ld_wait();  // vb landed; also orders the store below after every earlier load of these columns
                    tmem_st_32x32b_x16(taddr + c * 16, pk);
                    if (c + 2 < 8) tmem_ld_32x32b_x32(taddr + (c + 2) * 32, va);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(vb[2 * j]), scale_log2e, -mxs));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(vb[2 * j + 1]), scale_log2e, -mxs));
                        sum += p0 + p1;
                        pk[j] = pack_bf16x2(p0, p1);
                    }
                    // P columns [16(c+1), 16(c+1)+16) alias S columns < 32(c+1): all read already (see above), but a
                    // load of chunk c+2 may be in flight - it touches columns >= 32(c+2) only.
                    tmem_st_32x32b_x16(taddr + (c + 1) * 16, pk);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[g]);

            // epilogue: O / sum -> bf16 -> swizzled smem tile (Q_g's slot) -> one TMA store
            mbar_wait(&o_full[g], hp);
            tc_fence_after();
            const float inv = 1.0f / sum;
            {
                uint32_t va[32], vb[32];
                tmem_ld_32x32b_x32(taddr + 128, va);
                tmem_ld_32x32b_x32(taddr + 160, vb);
                tmem_ld_wait();
                uint8_t* rowp = sO + r * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 o;
                    o.x = pack_bf16x2(__uint_as_float(va[8 * j + 0]) * inv, __uint_as_float(va[8 * j + 1]) * inv);
                    o.y = pack_bf16x2(__uint_as_float(va[8 * j + 2]) * inv, __uint_as_float(va[8 * j + 3]) * inv);
                    o.z = pack_bf16x2(__uint_as_float(va[8 * j + 4]) * inv, __uint_as_float(va[8 * j + 5]) * inv);
                    o.w = pack_bf16x2(__uint_as_float(va[8 * j + 6]) * inv, __uint_as_float(va[8 * j + 7]) * inv);
                    *reinterpret_cast<uint4*>(rowp + ((j ^ (r & 7)) << 4)) = o;
                    o.x = pack_bf16x2(__uint_as_float(vb[8 * j + 0]) * inv, __uint_as_float(vb[8 * j + 1]) * inv);
                    o.y = pack_bf16x2(__uint_as_float(vb[8 * j + 2]) * inv, __uint_as_float(vb[8 * j + 3]) * inv);
                    o.z = pack_bf16x2(__uint_as_float(vb[8 * j + 4]) * inv, __uint_as_float(vb[8 * j + 5]) * inv);
                    o.w = pack_bf16x2(__uint_as_float(vb[8 * j + 6]) * inv, __uint_as_float(vb[8 * j + 7]) * inv);
                    *reinterpret_cast<uint4*>(rowp + (((4 + j) ^ (r & 7)) << 4)) = o;
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
