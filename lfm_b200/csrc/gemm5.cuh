// lfm_b200 - CTA-pair tcgen05 GEMM with a 512 x 256 cluster tile ("gemm5"): the pair kernel of gemm2.cuh fed with 25 % less
// L2 traffic per FLOP.
//
//   C[M, N] = A[M, K] * W[N, K]^T
//
// Why: with one 256 x 256 tile per CTA pair every SM pulls 32 KB from L2 per 64-wide K slab, i.e. 64 B/clk at full
// tensor rate - 9.5 KB/clk for the chip, while the L2 delivers ~6.3 KB/clk (B300_MICROARCH.md "LTS throughput cap"): the
// mainloop cannot pass ~66 % of the tensor peak, which is what gemm2 measures (63-66 % tensor-pipe active, 1.07 GB
// L2 -> SM per fc1 launch; profiles/r2b_gemm_vs_cublas.md).  cuBLAS' kernel for the same shapes (nvjet 256x256 per CTA,
// 2-CTA) moves 0.81 GB: it computes 512 x 256 per pair.  Same here: the pair issues TWO M = 256 cta_group::2 MMAs per K
// step - rows [0, 256) and [256, 512) of the tile - against ONE 256-row W slab; per slab each CTA stages 2 x 128 rows
// of A and its 128-row half of W: 48 KB per 2 x (256 x 256 x 64) MACs = 48 B/clk/SM.
//
// TMEM: the two fp32 accumulators fill all 512 columns, so a tile's epilogue can no longer hide behind the NEXT tile's
// whole mainloop (gemm2's double buffer).  Instead the two halves of a tile run SKEWED: half 0 leads half 1 by kSkew K
// slabs.  Half 0 therefore finishes (and starts draining) kSkew slabs before half 1, and on the next tile half 0
// restarts as soon as ITS accumulator is drained, while half 1's is still being read.  The 4-stage TMA ring holds the
// slabs between the two cursors.
//
// Tail: what does not fill a whole wave of 512 x 256 tiles is cut into 256 x 256 halves or 256 x 128 quarters (one
// accumulator each, alternating, i.e. double-buffered exactly as in gemm2), whichever finishes first.
//
// Roles per CTA (384 threads) as in gemm2.cuh: warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warp 2 TMEM
// allocator, warps 4..11 epilogue (TMEM lane quadrant = warp % 4, column half = (warp - 4) / 4).  The epilogue code is
// gemm2's: one 256-row x <= 256-column sub-tile at a time from one accumulator.
#pragma once
#include "common.cuh"
#include "gemm.cuh"
#include "gemm2.cuh"

namespace lfm {

constexpr int kG5Threads = 384;
constexpr int kG5Stages = 4;
constexpr int kG5Skew = 2;                      // K slabs by which half 0 leads half 1 (< kG5Stages)
constexpr int kG5ABytes = 128 * 64 * 2;        // one 128-row A block
constexpr int kG5BBytes = 128 * 64 * 2;        // this CTA's half of the W slab
constexpr int kG5StageBytes = 2 * kG5ABytes + kG5BBytes;  // 48 KB
constexpr int kG5SmemBytes = kG5Stages * kG5StageBytes + 1024 /*barriers*/ + kG2StagingBytes + 1024 /*align slack*/;

// Work list of one launch (identical in every thread): `full_count` 512 x 256 tiles, then the remaining `rem` tiles
// cut into `split` pieces each (1: whole, 2: 256 x 256 halves, 4: 256 x 128 quarters).
struct G5Sched {
    int n_blocks, m512, full_count, split, num_items;
};
LFM_DEVICE G5Sched g5_schedule(int M, int N, int clusters) {
    G5Sched s;
    s.n_blocks = (N + 255) / 256;
    s.m512 = (M + 511) / 512;
    const int tiles = s.m512 * s.n_blocks;
    s.full_count = (tiles / clusters) * clusters;
    const int rem = tiles - s.full_count;
    s.split = 1;
    if (rem > 0) {
        // rounds(f) / f = time of the tail in units of one full tile
        int best_num = (rem + clusters - 1) / clusters * 4, best = 1;  // cost * 4
        for (int f = 2; f <= 4; f *= 2) {
            const int c = ((rem * f + clusters - 1) / clusters) * (4 / f);
            if (c < best_num) {
                best_num = c;
                best = f;
            }
        }
        s.split = best;
    }
    s.num_items = s.full_count + rem * s.split;
    return s;
}
// One 256-row sub-tile of a work item.
struct G5Sub {
    int m_blk;   // 256-row block index
    int n_blk;   // 256-column block index
    int nh;      // which 128-column half (width == 128)
    int width;   // 256 or 128
};
// item -> number of sub-tiles (2 for a full 512 x 256 tile) and sub-tile j
LFM_DEVICE int g5_decode(const G5Sched& s, int item, int reverse_m, G5Sub* sub) {
    int tile, nsub = 1, h = 0, nh = 0, width = 256;
    if (item < s.full_count) {
        tile = item;
        nsub = 2;
    } else {
        const int w = item - s.full_count;
        tile = s.full_count + w / s.split;
        const int p = w % s.split;
        if (s.split == 1) {
            nsub = 2;
        } else if (s.split == 2) {
            h = p;
        } else {
            h = p >> 1;
            nh = p & 1;
            width = 128;
        }
    }
    int mb = tile / s.n_blocks;
    const int nb = tile % s.n_blocks;
    if (reverse_m) mb = s.m512 - 1 - mb;
    for (int j = 0; j < nsub; ++j) {
        sub[j].m_blk = 2 * mb + (nsub == 2 ? j : h);
        sub[j].n_blk = nb;
        sub[j].nh = nh;
        sub[j].width = width;
    }
    return nsub;
}

// Read one staged 32-row x 32-column fp32 chunk back (swizzled 128-byte rows) and write it with coalesced 16-byte
// accesses: plain stores, or red.global.add for the gated residual (x += ...; one add per element: deterministic).
template <int EPI>
LFM_DEVICE void g5_flush_f32(const uint8_t* stg, const GemmEpi& ep, int row_l, int col0, int M, int N, int sr, int sc) {
    float* outp = static_cast<float*>(ep.out);
    const int col = col0 + sc * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + sr;
        const float4 v = *reinterpret_cast<const float4*>(stg + r * 128 + ((sc ^ (r & 7)) << 4));
        if (row_l + r < M && col < N) {
            float* p = outp + static_cast<size_t>(row_l + r) * ep.ldo + col;
            if (EPI == EPI_GATE_RESID_F32)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            else
                *reinterpret_cast<float4*>(p) = v;
        }
    }
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kG5Threads, 1)
gemm5_bf16_tcgen05(const __grid_constant__ CUtensorMap tmap_a,   // A [M, K], box {64, 128} (or the 4-D im2col map)
                   const __grid_constant__ CUtensorMap tmap_b,   // W [N, K], box {64, 128}
                   const __grid_constant__ CUtensorMap tmap_out, // out [M, ldo]: box {128 bytes, 32 rows}, 128B swizzle
                   const __grid_constant__ CUtensorMap tmap_bh,  // W [N, K], box {64, 64}: 128-column quarter tiles
                   int M, int N, int K, GemmEpi ep, ConvGeom cg, int sched_clusters, int skew /* 0..kG5Skew */, int dbg_flags) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kG5Stages * kG5StageBytes);
    uint64_t* full_bar = bars;                        // [stages] (used in the leader CTA only)
    uint64_t* empty_bar = bars + kG5Stages;           // [stages] per CTA, signalled by the leader's multicast commit
    uint64_t* tmem_full = bars + 2 * kG5Stages;       // [2] per CTA, multicast commit
    uint64_t* tmem_empty = bars + 2 * kG5Stages + 2;  // [2] leader only: 16 epilogue warps (8 per CTA) arrive
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kG5Stages + 4);
    uint8_t* smem_stage = smem + kG5Stages * kG5StageBytes + 1024;  // 8 x 4 KB epilogue staging tiles (1024-aligned)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int num_kb = K / 64;
    const G5Sched sched = g5_schedule(M, N, sched_clusters);  // the host sized the grid from the same schedule

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_bh);
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        prefetch_tmap(&tmap_out);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kG5Stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
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
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = cluster_id; item < sched.num_items; item += num_clusters) {
                G5Sub sub[2];
                const int nsub = g5_decode(sched, item, ep.reverse_m, sub);
                const bool quarter = sub[0].width != 256;
                const int row_b = (cg.batch_m > 0 ? (sub[0].m_blk * 256 / cg.batch_m) * cg.b_batch_rows : 0) + sub[0].n_blk * 256 +
                                  sub[0].nh * 128 + static_cast<int>(rank) * (sub[0].width / 2);
                const uint32_t tx_bytes = 2 * (nsub * kG5ABytes + (quarter ? kG5BBytes / 2 : kG5BBytes));
                int row_a[2], img0[2] = {0, 0}, h0[2] = {0, 0}, w0[2] = {0, 0};
                for (int j = 0; j < nsub; ++j) {
                    const int row_v = sub[j].m_blk * 256 + static_cast<int>(rank) * 128;
                    row_a[j] = cg.a_mod > 0 ? row_v % cg.a_mod : row_v;
                    if (cg.taps != 0) {
                        img0[j] = row_a[j] / cg.HW;
                        h0[j] = (row_a[j] % cg.HW) / cg.W;
                        w0[j] = row_a[j] % cg.W;
                    }
                }
                int tap = 0, cb = 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                    uint8_t* st = smem + stage * kG5StageBytes;
                    for (int j = 0; j < nsub; ++j) {
                        if (cg.taps == 0) {
                            tma_load_2d_2sm(st + j * kG5ABytes, &tmap_a, &full_bar[stage], kb * 64, row_a[j]);
                        } else {
                            const int r = tap / 3, sx = tap - 3 * r;
                            tma_load_4d_2sm(st + j * kG5ABytes, &tmap_a, &full_bar[stage], cb * 64, cg.stride * w0[j] + sx - 1,
                                            cg.stride * h0[j] + r - 1, img0[j]);
                        }
                    }
                    if (cg.taps != 0 && ++cb == cg.cblocks) {
                        cb = 0;
                        ++tap;
                    }
                    tma_load_2d_2sm(st + 2 * kG5ABytes, quarter ? &tmap_bh : &tmap_b, &full_bar[stage], kb * 64, row_b);
                    if (++stage == kG5Stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA, one thread) =====================
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc_full = make_idesc_bf16(256, 256, 0, 0);
            constexpr uint32_t idesc_half = make_idesc_bf16(256, 128, 0, 0);
            int front = 0, back = 0;  // stage cursors: `front` = next slab to WAIT for, `back` = next slab to RELEASE
            uint32_t front_phase = 0;
            int acc = 0;              // next accumulator to use
            uint32_t acc_phase[2] = {0, 0};
            auto issue = [&](uint32_t tmem_d, int stage, int a_block, uint32_t idesc, bool first) {
                const uint8_t* st = smem + stage * kG5StageBytes;
                const uint64_t da = make_smem_desc_sw128(smem_u32(st + a_block * kG5ABytes), 16, 1024);
                const uint64_t db = make_smem_desc_sw128(smem_u32(st + 2 * kG5ABytes), 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ss_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc, !(first && k == 0));
            };
            auto wait_front = [&]() {  // the slab at the front cursor has landed; returns its stage and advances
                mbar_wait(&full_bar[front], front_phase);
                tc_fence_after();
                const int s = front;
                if (++front == kG5Stages) {
                    front = 0;
                    front_phase ^= 1;
                }
                return s;
            };
            auto release_back = [&]() {  // all MMAs reading the slab at the back cursor have been issued
                umma_commit_2cta(&empty_bar[back]);
                if (++back == kG5Stages) back = 0;
            };
            auto take_acc = [&]() {  // next accumulator, once the epilogue has drained it
                const int a = acc;
                mbar_wait(&tmem_empty[a], acc_phase[a] ^ 1);
                tc_fence_after();
                acc_phase[a] ^= 1;
                acc ^= 1;
                return a;
            };
            for (int item = cluster_id; item < sched.num_items; item += num_clusters) {
                G5Sub sub[2];
                const int nsub = g5_decode(sched, item, ep.reverse_m, sub);
                const uint32_t idesc = sub[0].width == 256 ? idesc_full : idesc_half;
                if (nsub == 1) {
                    const int a = take_acc();
                    const uint32_t d = tmem_base + a * 256;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        const int s = wait_front();
                        issue(d, s, 0, idesc, kb == 0);
                        release_back();
                        if (kb == num_kb - 1) umma_commit_2cta(&tmem_full[a]);
                    }
                } else {
                    // skewed halves: half 0 (rows 0..255, A block 0) leads half 1 (A block 1) by up to kG5Skew slabs
                    const int a0 = take_acc();
                    const uint32_t d0 = tmem_base + a0 * 256;
                    int stages_held[kG5Skew + 1];  // stages of the slabs half 0 has consumed and half 1 not yet
                    int lead = 0;
                    const int pre = num_kb < skew ? num_kb : skew;
                    for (int kb = 0; kb < pre; ++kb) {
                        stages_held[lead++] = wait_front();
                        issue(d0, stages_held[lead - 1], 0, idesc, kb == 0);
                        if (kb == num_kb - 1) umma_commit_2cta(&tmem_full[a0]);
                    }
                    const int a1 = take_acc();
                    const uint32_t d1 = tmem_base + a1 * 256;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        // half 1 on slab kb (the oldest held stage), then half 0 on slab kb + skew
                        if (lead == 0) {  // skew 0: both halves take the slab together
                            stages_held[lead++] = wait_front();
                            issue(d0, stages_held[0], 0, idesc, kb == 0);
                            if (kb == num_kb - 1) umma_commit_2cta(&tmem_full[a0]);
                        }
                        const int s1 = stages_held[0];
                        issue(d1, s1, 1, idesc, kb == 0);
                        release_back();
                        for (int i = 1; i < lead; ++i) stages_held[i - 1] = stages_held[i];
                        --lead;
                        if (kb == num_kb - 1) umma_commit_2cta(&tmem_full[a1]);
                        const int k0 = kb + pre;
                        if (pre > 0 && k0 < num_kb) {
                            const int s0 = wait_front();
                            stages_held[lead++] = s0;
                            issue(d0, s0, 0, idesc, false);
                            if (k0 == num_kb - 1) umma_commit_2cta(&tmem_full[a0]);
                        }
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 8 warps, one 128 rows x (2 x <= 128 columns) sub-tile at a time =====================
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const uint64_t keep_policy = ep.l2_keep ? l2_policy_evict_last() : 0;
        uint8_t* stg = smem_stage + (warp - 4) * 4096;
        constexpr bool kBf16Out = (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16);
        int acc = 0;
        uint32_t acc_phase[2] = {0, 0};
        for (int item = cluster_id; item < sched.num_items; item += num_clusters) {
            G5Sub sub[2];
            const int nsub = g5_decode(sched, item, ep.reverse_m, sub);
            for (int j = 0; j < nsub; ++j) {
                const int width = sub[j].width;
                const int nch = width / 64;  // 32-column chunks per warp: 4 or 2
                const int row_l = sub[j].m_blk * 256 + static_cast<int>(rank) * 128 + q * 32;
                const int row = row_l + lane;
                const int nbase = sub[j].n_blk * 256 + sub[j].nh * 128 + half * (width / 2);
                const float* gate_row = nullptr;
                if (EPI == EPI_GATE_RESID_F32 && ep.gate != nullptr)
                    gate_row = ep.gate + static_cast<size_t>((row < M ? row : M - 1) / ep.rows_per_sample) * ep.gate_stride;
                const float* add_row = nullptr;
                if (EPI == EPI_BIAS_F32 && ep.addend != nullptr) add_row = ep.addend + static_cast<size_t>(row < M ? row : M - 1) * ep.ldo;
                const int a = acc;
                acc ^= 1;
                mbar_wait(&tmem_full[a], acc_phase[a]);
                acc_phase[a] ^= 1;
                tc_fence_after();
                if (dbg_flags & 1) {  // measurement aid: skip the drain (results are not written)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(&tmem_empty[a], 0);
                    continue;
                }
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * 256 + half * (width / 2);
                uint32_t va[32], vb[32];
                float f[32];
                tmem_ld_32x32b_x32(taddr, va);
                if (!(dbg_flags & 2)) {
                    // Default: registers -> swizzled staging tile -> COALESCED global stores by the same warp (every store
                    // instruction writes 4 whole 128-byte row segments).  No TMA store: its read-completion wait sits on
                    // the critical path of an epilogue that is no longer hidden behind a mainloop (measured: 7200 clk per
                    // 256-column sub-tile with TMA stores queued behind the producer's loads).
                    const int sr = lane >> 3, sc = lane & 7;  // this lane's row-in-group / 16-byte chunk when reading back
#pragma unroll 1
                    for (int c = 0; c < nch; c += 2) {
                        tmem_ld_wait();
                        tmem_ld_32x32b_x32(taddr + (c + 1) * 32, vb);
                        epilogue_math<EPI>(va, f, ep, nbase + c * 32, N, gate_row, add_row);
                        if ((EPI == EPI_BIAS_F32 || EPI == EPI_BIAS_BF16) && ep.gn_bins != nullptr) gn_accumulate(f, ep, row, M, nbase + c * 32, N, lane);
                        __syncwarp();  // the previous read-back of the staging tile is complete
                        if (kBf16Out) {
                            stage_row_bf16_half(stg, lane, f, 0);
                        } else {
                            stage_row_f32(stg, lane, f);
                            __syncwarp();
                            g5_flush_f32<EPI>(stg, ep, row_l, nbase + c * 32, M, N, sr, sc);
                            __syncwarp();
                        }
                        tmem_ld_wait();
                        if (c + 2 < nch) tmem_ld_32x32b_x32(taddr + (c + 2) * 32, va);
                        epilogue_math<EPI>(vb, f, ep, nbase + (c + 1) * 32, N, gate_row, add_row);
                        if ((EPI == EPI_BIAS_F32 || EPI == EPI_BIAS_BF16) && ep.gn_bins != nullptr) gn_accumulate(f, ep, row, M, nbase + (c + 1) * 32, N, lane);
                        if (kBf16Out) {
                            stage_row_bf16_half(stg, lane, f, 1);
                            __syncwarp();
                            __nv_bfloat16* outp = static_cast<__nv_bfloat16*>(ep.out);
                            const int col = nbase + c * 32 + sc * 8;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int r = 4 * i + sr;
                                const uint4 u = *reinterpret_cast<const uint4*>(stg + r * 128 + ((sc ^ (r & 7)) << 4));
                                if (row_l + r < M && col < N)
                                    *reinterpret_cast<uint4*>(outp + static_cast<size_t>(row_l + r) * ep.ldo + col) = u;
                            }
                        } else {
                            stage_row_f32(stg, lane, f);
                            __syncwarp();
                            g5_flush_f32<EPI>(stg, ep, row_l, nbase + (c + 1) * 32, M, N, sr, sc);
                        }
                    }
                } else {
#pragma unroll 1
                for (int c = 0; c < nch; c += 2) {
                    tmem_ld_wait();
                    tmem_ld_32x32b_x32(taddr + (c + 1) * 32, vb);
                    epilogue_math<EPI>(va, f, ep, nbase + c * 32, N, gate_row, add_row);
                    if ((EPI == EPI_BIAS_F32 || EPI == EPI_BIAS_BF16) && ep.gn_bins != nullptr) gn_accumulate(f, ep, row, M, nbase + c * 32, N, lane);
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
                    if (kBf16Out) {
                        stage_row_bf16_half(stg, lane, f, 0);
                    } else {
                        stage_row_f32(stg, lane, f);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0 && nbase + c * 32 < N) {
                            if (EPI == EPI_GATE_RESID_F32 && ep.l2_keep)
                                tma_reduce_add_2d_hint(&tmap_out, stg, nbase + c * 32, row_l, keep_policy);
                            else if (EPI == EPI_GATE_RESID_F32)
                                tma_reduce_add_2d(&tmap_out, stg, nbase + c * 32, row_l);
                            else
                                tma_store_2d(&tmap_out, stg, nbase + c * 32, row_l);
                        }
                        if (lane == 0) tma_store_commit();
                    }
                    tmem_ld_wait();
                    if (c + 2 < nch) tmem_ld_32x32b_x32(taddr + (c + 2) * 32, va);
                    epilogue_math<EPI>(vb, f, ep, nbase + (c + 1) * 32, N, gate_row, add_row);
                    if ((EPI == EPI_BIAS_F32 || EPI == EPI_BIAS_BF16) && ep.gn_bins != nullptr) gn_accumulate(f, ep, row, M, nbase + (c + 1) * 32, N, lane);
                    if (kBf16Out) {
                        stage_row_bf16_half(stg, lane, f, 1);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0 && nbase + c * 32 < N) tma_store_2d(&tmap_out, stg, nbase + c * 32, row_l);  // 64 bf16 cols
                        if (lane == 0) tma_store_commit();
                    } else {
                        if (lane == 0) tma_store_wait_read<0>();
                        __syncwarp();
                        stage_row_f32(stg, lane, f);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0 && nbase + (c + 1) * 32 < N) {
                            if (EPI == EPI_GATE_RESID_F32 && ep.l2_keep)
                                tma_reduce_add_2d_hint(&tmap_out, stg, nbase + (c + 1) * 32, row_l, keep_policy);
                            else if (EPI == EPI_GATE_RESID_F32)
                                tma_reduce_add_2d(&tmap_out, stg, nbase + (c + 1) * 32, row_l);
                            else
                                tma_store_2d(&tmap_out, stg, nbase + (c + 1) * 32, row_l);
                        }
                        if (lane == 0) tma_store_commit();
                    }
                }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&tmem_empty[a], 0);  // leader's barrier
            }
        }
        if (lane == 0) tma_store_wait<0>();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2cta<512>(tmem_base);
    }
}

}  // namespace lfm
