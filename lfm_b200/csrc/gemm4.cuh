// lfm_b200 - 4-CTA-cluster tcgen05 GEMM: two cta_group::2 pairs that share their A rows by TMA multicast.
//
// The pair kernel (gemm2.cuh) is bound by L2 -> SM operand traffic (32 KB per 64-wide K slab per SM, ~70 % tensor
// pipe active, the same place cuBLAS lands: profiles/r1b_gemm2_full.md).  Here a cluster of four CTAs computes a
// 256 x 512 output block: pair p = cluster_rank / 2 owns the 256 x 256 tile of n-block 2 nq + p; both pairs need
// the same 256 rows of A, so each CTA loads HALF of its 128-row A slab (64 rows, 8 KB) and multicasts it to the CTA
// with the same in-pair rank in the other pair.  L2 traffic per SM per slab: 8 KB (A) + 16 KB (own half of W) =
// 24 KB instead of 32 KB.
//
// Synchronisation differences from gemm2: a shared-memory slot is refilled by the own producer AND by the other pair's
// multicast, so every MMA issuer's commit releases the slot in all four CTAs (multicast mask 0b1111) and the empty
// barriers expect two arrivals (one per pair).  Full barriers live in each pair's leader and expect 64 KB per stage.
#pragma once
#include "gemm2.cuh"

namespace lfm {

constexpr int kG4Threads = kG2Threads;
constexpr int kG4SmemBytes = kG2SmemBytes;

LFM_DEVICE void tma_load_2d_2sm_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1, uint16_t mask) {
    const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%4, %5}], [%2], %3;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar), "h"(mask), "r"(c0), "r"(c1)
        : "memory");
}
LFM_DEVICE void umma_commit_mask(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(kG4Threads, 1)
gemm4_bf16_tcgen05(const __grid_constant__ CUtensorMap tmap_a64,  // A [M, K], box {64, 64}  (half of a CTA's A slab)
                   const __grid_constant__ CUtensorMap tmap_b,    // W [N, K], box {64, 128}
                   const __grid_constant__ CUtensorMap tmap_out,  // out [M, ldo]: box {128 bytes, 32 rows}, 128B swizzle
                   int M, int N, int K, GemmEpi ep) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kG2Stages * kG2ABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kG2Stages * kG2StageBytes);
    uint64_t* full_bar = bars;                        // [stages] (pair leaders)
    uint64_t* empty_bar = bars + kG2Stages;           // [stages] per CTA: one commit arrival from EACH pair's issuer
    uint64_t* tmem_full = bars + 2 * kG2Stages;       // [2] per CTA, commit multicast inside the pair
    uint64_t* tmem_empty = bars + 2 * kG2Stages + 2;  // [2] pair leaders: 16 epilogue warps of the pair arrive
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kG2Stages + 4);
    uint8_t* smem_stage = smem + kG2Stages * kG2StageBytes + 1024;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();  // 0..3
    const uint32_t pair = crank >> 1;          // which 256-column half of the cluster's 256 x 512 block
    const uint32_t prank = crank & 1;          // rank inside the cta_group::2 pair (0 = leader)
    const int cluster_id = blockIdx.x >> 2;
    const int num_clusters = gridDim.x >> 2;

    const int n_blocks = (N + kG2BlockN - 1) / kG2BlockN;
    const int nq_blocks = (n_blocks + 1) / 2;
    const int m_blocks = (M + 255) / 256;
    const int num_items = m_blocks * nq_blocks;
    const int num_kb = K / 64;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a64);
        prefetch_tmap(&tmap_b);
        prefetch_tmap(&tmap_out);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kG2Stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 2);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 16);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 2) tmem_alloc_2cta<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (all four CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint16_t mc_mask = static_cast<uint16_t>((1u << prank) | (1u << (2 + prank)));  // same in-pair rank, both pairs
            for (int item = cluster_id; item < num_items; item += num_clusters) {
                const int m_blk = item / nq_blocks, n_blk = 2 * (item % nq_blocks) + static_cast<int>(pair);
                // this CTA's quarter of the A block: rows of in-pair rank `prank`, half `pair`
                const int row_a = m_blk * 256 + static_cast<int>(prank) * 128 + static_cast<int>(pair) * 64;
                const int row_b = n_blk * kG2BlockN + static_cast<int>(prank) * 128;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);  // slot free in THIS CTA (both pairs have consumed it)
                    if (prank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kG2StageBytes);
                    tma_load_2d_2sm_mc(smem_a + stage * kG2ABytes + pair * (kG2ABytes / 2), &tmap_a64, &full_bar[stage],
                                       kb * 64, row_a, mc_mask);
                    tma_load_2d_2sm(smem_b + stage * kG2BBytes, &tmap_b, &full_bar[stage], kb * 64, row_b);
                    if (++stage == kG2Stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (each pair's leader CTA, one thread) =====================
        if (prank == 0 && lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(256, kG2BlockN, 0, 0);
            const uint16_t pair_mask = static_cast<uint16_t>(3u << (2 * pair));
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int item = cluster_id; item < num_items; item += num_clusters) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * kG2BlockN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + stage * kG2ABytes), 16, 1024);
                    const uint64_t db = make_smem_desc_sw128(smem_u32(smem_b + stage * kG2BBytes), 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_ss_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    umma_commit_mask(&empty_bar[stage], 0xF);  // slot released in all four CTAs (A is shared)
                    if (kb == num_kb - 1) umma_commit_mask(&tmem_full[acc], pair_mask);
                    if (++stage == kG2Stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 8 warps, 128 rows x (2 x 128 columns) =====================
        // TMEM -> registers -> (bias / GELU / gate) -> swizzled smem staging tile -> TMA store, or TMA reduce-add
        // for the gated residual (x += g * (acc + b) is applied at L2: the SM never reads x).
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        uint8_t* stg0 = smem_stage + (warp - 4) * 4096;  // one staging tile per warp (measured: a second one at the
        constexpr int sbuf = 0;                           // cost of a pipeline stage does not pay)
        constexpr bool kBf16Out = (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = cluster_id; item < num_items; item += num_clusters) {
            const int m_blk = item / nq_blocks, n_blk = 2 * (item % nq_blocks) + static_cast<int>(pair);
            constexpr int nch = 4, width = kG2BlockN;
            const int row0 = m_blk * 256 + static_cast<int>(prank) * 128 + q * 32;  // first row of this warp
            const int row = row0 + lane;
            const int nbase = n_blk * kG2BlockN + half * (width / 2);
            const float* gate_row = nullptr;
            if (EPI == EPI_GATE_RESID_F32 && ep.gate != nullptr)  // gate == nullptr: plain residual add (gate 1)
                gate_row = ep.gate + static_cast<size_t>((row < M ? row : M - 1) / ep.rows_per_sample) * ep.gate_stride;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kG2BlockN + half * (width / 2);
            uint32_t va[32], vb[32];
            float f[32];
            tmem_ld_32x32b_x32(taddr, va);
#pragma unroll 1
            for (int c = 0; c < nch; c += 2) {
                tmem_ld_wait();
                tmem_ld_32x32b_x32(taddr + (c + 1) * 32, vb);
                epilogue_math<EPI>(va, f, ep, nbase + c * 32, N, gate_row);
                // this staging tile is free once all but the most recent TMA op of this warp have READ their tile
                uint8_t* stg = stg0 + sbuf * 4096;
                if (lane == 0) tma_store_wait_read<0>();
                __syncwarp();
                if (kBf16Out) {
                    stage_row_bf16_half(stg, lane, f, 0);
                } else {
                    stage_row_f32(stg, lane, f);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && nbase + c * 32 < N) {
                        if (EPI == EPI_GATE_RESID_F32)
                            tma_reduce_add_2d(&tmap_out, stg, nbase + c * 32, row0);
                        else
                            tma_store_2d(&tmap_out, stg, nbase + c * 32, row0);
                    }
                    if (lane == 0) tma_store_commit();
                }
                tmem_ld_wait();
                if (c + 2 < nch) tmem_ld_32x32b_x32(taddr + (c + 2) * 32, va);
                epilogue_math<EPI>(vb, f, ep, nbase + (c + 1) * 32, N, gate_row);
                if (kBf16Out) {
                    stage_row_bf16_half(stg, lane, f, 1);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && nbase + c * 32 < N) tma_store_2d(&tmap_out, stg, nbase + c * 32, row0);  // 64 bf16 cols
                    if (lane == 0) tma_store_commit();
                } else {
                    stg = stg0 + sbuf * 4096;
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
                    stage_row_f32(stg, lane, f);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && nbase + (c + 1) * 32 < N) {
                        if (EPI == EPI_GATE_RESID_F32)
                            tma_reduce_add_2d(&tmap_out, stg, nbase + (c + 1) * 32, row0);
                        else
                            tma_store_2d(&tmap_out, stg, nbase + (c + 1) * 32, row0);
                    }
                    if (lane == 0) tma_store_commit();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 2 * pair);  // this pair's leader
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (lane == 0) tma_store_wait<0>();  // all global writes of this warp are complete before the CTA exits
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2cta<512>(tmem_base);
    }
}

}  // namespace lfm
