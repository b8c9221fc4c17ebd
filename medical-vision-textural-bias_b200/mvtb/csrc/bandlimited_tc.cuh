// bandlimited_tc.cuh — the H-axis passes of the band-limited path on the 5th-generation tensor cores.
//
// The pruned DFT along H is a skinny GEMM per column tile:  Y[col, (f, re|im)] = sum_h x[h, col] * T[(f, re|im), h].
// On CUDA cores (bandlimited_quad.cuh) it costs ~16 instructions per voxel and the kernels sit at ~50 % issue
// utilisation, not on HBM.  Here the multiply-adds go to tcgen05.mma (kind::tf32, fp32 accumulators in TMEM) with the
// 3xTF32 split  x*T ~ xh*Th + xl*Th + xh*Tl  (xh = tf32(x), xl = tf32(x - xh); rel-L2 ~1e-6, measured by
// tools/tc_probe.cu), and the SM's threads only convert and move data:
//
//   producer   (1 warp)     2-D TMA boxes (cp.async.bulk.tensor, UTMALDG) of 16 `box_stages` rows x 128 columns of x into a
//                           ring in shared memory, completion counted in bytes on an mbarrier
//   converters (<= 16 warps) thread = column: LDS, split into (hi, lo), tcgen05.st into a ring of A-operand slots in
//                           TMEM (lane = column, one 32-bit TMEM column per h); one group of 4 warps per slot
//   MMA issuers (6 warps)   (tile parity, term): per 8 rows D_term += A_hi*B_hi | A_lo*B_hi | A_hi*B_lo, A from TMEM, B (the
//                           cos / -sin table, split on the host) from shared memory; tcgen05.commit frees the A slot
//   epilogue   (4 warps)    tcgen05.ld of the finished 128 x N accumulator, coalesced 8-byte stores into Y[f][col]
//
// All hand-offs are mbarriers; every wait is bounded (a protocol bug sets *status instead of hanging the GPU).
// Y has exactly the layout and meaning k_bl_fwd_h4a produces, so the W/D stage and the inverse pass are unchanged.
// GPU only (no emulator build); included by bandlimited.cu inside namespace mvtb.
#pragma once
#ifndef MVTB_EMU
// tc_common.cuh is included by bandlimited.cu before namespace mvtb opens

static const int kTcRows = 16;          // rows of x per stage (two MMA K-steps of 8)
static const int kTcMaxBoxStages = 4;   // stages per TMA box (= per A-operand slot), at most
static const int kTcMaxRing = 8;        // boxes in the raw shared-memory ring, at most
static const int kTcMaxSlots = 4;       // A-operand slots in TMEM, at most
static const int kTcConvGroups = 4;     // converter groups of 4 warps (one per TMEM lane quarter), at most; a.a_slots of them work
static const int kTcIssuers = 6;        // MMA-issuing warps: (tile parity, 3xTF32 term), each with its own accumulator
static const int kTcTerms = 3;          // issuers (and accumulators) per tile
static const int kTcWarpIss0 = 1, kTcWarpEpi0 = 1 + kTcIssuers, kTcWarpConv0 = kTcWarpEpi0 + 4;
static const int kTcFwdThreads = 32 * (kTcWarpConv0 + 4 * kTcConvGroups);

struct TcFwdArgs {
    float2* Y;               // [nvol][NF][NC]
    const float* tab;        // [2][H * N]: B operand hi, lo in tc::op_offset layout; row 2f = cos, 2f+1 = -sin, k = h
    int H, NC, NF, N;        // N = 2 NF rounded up to a multiple of 16
    int n_tiles, tiles_per_vol;
    int box_stages;          // stages (of 16 rows) per TMA box and per A-operand slot
    int ring_boxes;          // boxes in the raw shared-memory ring
    int a_slots;             // A-operand slots in TMEM (each 32 * box_stages columns)
    int acc_sets;            // 2: the accumulators of even / odd tiles are separate (6 N TMEM columns); 1: one set of three (3 N),
                             // for N = 64 (NF = 26, 32), where six would leave room for single-stage slots only
    const float* abt;        // null, or the intensity prologue map of every volume of the launch (a, b, t): y = x != 0 ? a x + b : t,
                             // applied by the converters as they read x (mvtb_kspace_chain_ex_f32, pre_abt)
    int* status;
    long long* prof;         // null, or wait cycles / event timeline of CTA 0 (MVTB_TC_PROF)
};

struct TcBars {
    unsigned long long raw_full[kTcMaxRing], raw_empty[kTcMaxRing];
    unsigned long long a_full[kTcMaxSlots], a_empty[kTcMaxSlots];
    unsigned long long d_full[2], d_empty[2];    // per tile parity: the accumulators are double-buffered
};

// bounded wait that also gives up when another role has failed
__device__ __forceinline__ bool tc_wait_raw(unsigned long long* bar, uint32_t parity, volatile int* abort_flag, int* status, int code) {
    const uint32_t b = tc::smem_u32(bar);
    for (int round = 0; round < (1 << 12); ++round) {
        MVTB_UNROLL_N(1)
        for (int it = 0; it < 256; ++it) {
            uint32_t ok;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(b), "r"(parity), "r"(20000u) : "memory");
            if (ok) return true;
        }
        if (*abort_flag) return false;
    }
    *abort_flag = 1;
    atomicCAS(status, 0, code);
    return false;
}
// MVTB_TC_PROF: cycles each role of CTA 0 spends waiting, per barrier kind (code), into a.prof[warp * 8 + code]
__device__ __forceinline__ bool tc_wait(unsigned long long* bar, uint32_t parity, volatile int* abort_flag, int* status, int code,
                                        long long* prof = nullptr) {
    if (prof == nullptr || blockIdx.x != 0) return tc_wait_raw(bar, parity, abort_flag, status, code);
    const long long t0 = clock64();
    const bool ok = tc_wait_raw(bar, parity, abort_flag, status, code);
    if ((threadIdx.x & 31) == 0) prof[(threadIdx.x >> 5) * 8 + code] += clock64() - t0;
    return ok;
}

// What the measurements on a B200 dictated (tools/tc_rate.cu, tc_rate2.cu, MVTB_TC_PROF; profiles/r02_tc_*):
//  * one issuing warp pays ~105-125 cycles per tcgen05.mma whatever the tile (M = 64/128, N = 16..128, A from shared
//    memory or TMEM), while the tensor pipe takes ~25 for 128 x 32 x 8: several warps issue, one per 3xTF32 term, each
//    into its own accumulator; the epilogue adds them.  A lone `if (lane == 0)` issuer costs ~200 (ptxas wraps
//    every UTCMMA in an ELECT / BRA.U.ANY loop): the whole warp runs the loop and one elected lane issues.
//  * with ONE set of accumulators the issuers sat 40 % of the kernel waiting for the epilogue to drain the previous
//    tile (MVTB_TC_PROF: 3.1 k of 7.6 k cycles per tile): two sets of three issuers take even / odd tiles, so the
//    accumulators are double-buffered within the same 6 N columns of TMEM.
//  * a copy costs its issuing thread ~90 cycles (1-D bulk) to ~550 (2-D tensor map) and lands ~2.9 k cycles later: one
//    TMA box of several stages (24 KB for H = 240), five boxes deep.
//  * every hand-off costs a wake-up (~100-400 cycles): the unit of work between roles is a box of `box_stages` stages
//    (48 rows for H = 240), not a stage -- per stage the roles spent most of their time waking up (826 cycles per
//    stage against a budget of 356).
//  * cvt.rna.tf32.f32 is ~8 integer instructions: the split uses a mask and one subtraction (2 instructions per voxel).
__global__ void __launch_bounds__(kTcFwdThreads, 1)
k_bl_fwd_tc(const __grid_constant__ CUtensorMap tmap, TcFwdArgs a) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ TcBars bars;
    __shared__ uint32_t s_tmem;
    __shared__ int s_abort;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = a.H, NC = a.NC, N = a.N;
    const size_t tab_bytes = (size_t)H * N * sizeof(float);
    float* tab_hi = (float*)tc_smem;
    float* tab_lo = (float*)(tc_smem + tab_bytes);
    float* raw = (float*)(tc_smem + ((2 * tab_bytes + 1023) & ~(size_t)1023));   // [ring][box_stages * 16][128]
    const int bs = a.box_stages, n_box = H / (kTcRows * bs);
    const int box_floats = bs * kTcRows * 128;
    // TMEM: [A ring: a_slots x (32 bs)][six accumulators of N columns]
    const uint32_t slot_cols = 32u * (uint32_t)bs;
    const uint32_t a_col0 = 0, d_col0 = (uint32_t)a.a_slots * slot_cols;

    for (size_t i = tid; i < 2 * tab_bytes / 16; i += blockDim.x) ((float4*)tc_smem)[i] = __ldg((const float4*)a.tab + i);
    if (tid == 0) {
        s_abort = 0;
        for (int i = 0; i < kTcMaxRing; ++i) { tc::mbar_init(tc::smem_u32(&bars.raw_full[i]), 1); tc::mbar_init(tc::smem_u32(&bars.raw_empty[i]), 4); }   // 4 = warps of the one group that reads the box
        for (int i = 0; i < kTcMaxSlots; ++i) { tc::mbar_init(tc::smem_u32(&bars.a_full[i]), 4); tc::mbar_init(tc::smem_u32(&bars.a_empty[i]), kTcIssuers); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(tc::smem_u32(&bars.d_full[i]), kTcTerms); tc::mbar_init(tc::smem_u32(&bars.d_empty[i]), 4); }
        tc::mbar_init_fence();
    }
    const uint32_t tmem_cols = 512;
    if (warp == 1) tc::tmem_alloc(tc::smem_u32(&s_tmem), tmem_cols);
    tc::fence_async_smem();                                           // the table was written by threads, read by the MMA unit
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    volatile int* abortp = &s_abort;
    const long long t_start = a.prof ? clock64() : 0;

    if (warp == 0) {
        // ------------------------------------------------------------ producer: one TMA box per slot of the raw ring
        unsigned ib = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles && !*abortp; tile += gridDim.x) {
            const int vol = tile / a.tiles_per_vol, c0 = (tile - vol * a.tiles_per_vol) * 128;
            bool ok = true;
            for (int b = 0; b < n_box; ++b, ++ib) {
                const int rb = (int)(ib % (unsigned)a.ring_boxes);
                if (!tc_wait(&bars.raw_empty[rb], ((ib / (unsigned)a.ring_boxes) & 1u) ^ 1u, abortp, a.status, 1, a.prof)) { ok = false; break; }
                if (tc::elect_one()) {
                    const uint32_t full = tc::smem_u32(&bars.raw_full[rb]);
                    tc::mbar_arrive_expect_tx(full, (uint32_t)box_floats * 4u);           // the whole box, columns past NC arrive as zeros
                    tc::tma_load_2d(tc::smem_u32(raw + (size_t)rb * box_floats), &tmap, c0, vol * H + b * bs * kTcRows, full);
                }
                __syncwarp();
            }
            if (!ok) break;
        }
    } else if (warp < kTcWarpEpi0) {
        // ------------------------------------------------------------ MMA issuers: 2 `bs` MMAs and one commit per box
        const int iss = warp - kTcWarpIss0, set = iss / kTcTerms, term = iss - set * kTcTerms;   // term 0 = hi*hi, 1 = lo*hi, 2 = hi*lo
        const uint32_t idesc = tc::idesc_tf32(128, N);
        const uint32_t chunk_stride = (uint32_t)(N / 8) * 128u, group_stride = 128u;
        const uint32_t b0 = tc::smem_u32(term == 2 ? tab_lo : tab_hi);
        const uint32_t a_off = term == 1 ? 16u : 0u;                         // within a stage's 32 columns: [hi 16 | lo 16]
        const uint32_t d = tmem + d_col0 + (uint32_t)(a.acc_sets == 2 ? iss : term) * (uint32_t)N;
        const bool two = a.acc_sets == 2;
        unsigned ib = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles && !*abortp; tile += gridDim.x, ++tcount) {
            if (two ? (int)(tcount & 1u) != set : set != 0) {
                // The other set's tile.  Still observe every fill and sign the slot off: a parity wait is only sound
                // when the waiter is never a phase ahead of or behind its barrier, so all six issuers see all boxes.
                bool ok = true;
                for (int b = 0; b < n_box; ++b, ++ib) {
                    const int sl = (int)(ib % (unsigned)a.a_slots);
                    if (!tc_wait(&bars.a_full[sl], (ib / (unsigned)a.a_slots) & 1u, abortp, a.status, 7, a.prof)) { ok = false; break; }
                    if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bars.a_empty[sl]));
                    __syncwarp();
                }
                if (!ok) break;
                continue;
            }
            if (!tc_wait(&bars.d_empty[set], (((two ? tcount >> 1 : tcount)) & 1u) ^ 1u, abortp, a.status, 2, a.prof)) break;
            tc::fence_after_sync();
            uint32_t acc = 0;
            bool ok = true;
            for (int b = 0; b < n_box; ++b, ++ib) {
                const int sl = (int)(ib % (unsigned)a.a_slots);
                if (!tc_wait(&bars.a_full[sl], (ib / (unsigned)a.a_slots) & 1u, abortp, a.status, 3, a.prof)) { ok = false; break; }
                tc::fence_after_sync();
                if (tc::elect_one()) {
                    for (int k = 0; k < 2 * bs; ++k) {                       // K-steps of 8 rows
                        const uint32_t koff = (uint32_t)(b * bs * (kTcRows / 8) + k) * 2u * chunk_stride;
                        tc::mma_ts(d, tmem + a_col0 + (uint32_t)sl * slot_cols + 32u * (uint32_t)(k >> 1) + a_off + 8u * (uint32_t)(k & 1),
                                   tc::smem_desc(b0 + koff, chunk_stride, group_stride), idesc, acc);
                        acc = 1;
                    }
                    tc::mma_commit(tc::smem_u32(&bars.a_empty[sl]));
                }
                acc = 1;
                __syncwarp();
            }
            if (!ok) break;
            if (tc::elect_one()) tc::mma_commit(tc::smem_u32(&bars.d_full[set]));
            __syncwarp();
        }
    } else if (warp < kTcWarpConv0) {
        // ------------------------------------------------------------ epilogue: sum of the tile's three accumulators -> Y[f][col]
        const int q = warp & 3;                                       // TMEM lanes 32 q .. 32 q + 31
        const int m = 32 * q + lane;
        unsigned tcount = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles && !*abortp; tile += gridDim.x, ++tcount) {
            const int vol = tile / a.tiles_per_vol, c0 = (tile - vol * a.tiles_per_vol) * 128;
            const unsigned set = a.acc_sets == 2 ? tcount & 1u : 0u;
            if (!tc_wait(&bars.d_full[set], (a.acc_sets == 2 ? tcount >> 1 : tcount) & 1u, abortp, a.status, 4, a.prof)) break;
            tc::fence_after_sync();
            const bool okc = c0 + m < NC;
            float2* yv = a.Y + ((size_t)vol * a.NF) * NC + c0 + m;
            const uint32_t t0 = tmem + ((uint32_t)(32 * q) << 16) + d_col0 + set * (uint32_t)(kTcTerms * N);
            for (int c = 0; c < N / 8; ++c) {
                uint32_t v[kTcTerms][8];
                MVTB_UNROLL
                for (int i = 0; i < kTcTerms; ++i) tc::tmem_ld8(t0 + (uint32_t)i * (uint32_t)N + 8u * c, v[i]);
                tc::tmem_ld_wait();
                if (c == N / 8 - 1) {                                // everything is in registers: release the accumulators
                    tc::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bars.d_empty[set]));
                }
                MVTB_UNROLL
                for (int j = 0; j < 4; ++j) {
                    const int f = 4 * c + j;
                    float r[2];
                    MVTB_UNROLL
                    for (int e = 0; e < 2; ++e)                       // small terms first, then hi*hi
                        r[e] = (__uint_as_float(v[1][2 * j + e]) + __uint_as_float(v[2][2 * j + e])) + __uint_as_float(v[0][2 * j + e]);
                    if (okc && f < a.NF) yv[(size_t)f * NC] = make_float2(r[0], r[1]);
                }
            }
        }
    } else {
        // ------------------------------------------------------------ converters: a box of raw rows -> (hi, lo) in a TMEM slot
        // Group g takes boxes g, g + G, ... with G = a_slots groups, and the ring holds a multiple of G boxes: every TMEM
        // slot and every ring slot is then filled by ONE group, in order, so that no waiter can be two phases ahead of its
        // barrier (a parity wait cannot tell "two completions ago" from "not yet"; with 3 slots and 4 groups that hung).
        const int grp = (warp - kTcWarpConv0) >> 2;
        const int G = a.a_slots;
        const int q = warp & 3;                                       // TMEM lane quarter
        const int m = 32 * q + lane;
        const unsigned n_ib = (unsigned)n_box * (unsigned)((a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
        for (unsigned ib = (unsigned)grp; grp < G && ib < n_ib && !*abortp; ib += (unsigned)G) {
            const int rb = (int)(ib % (unsigned)a.ring_boxes), sl = (int)(ib % (unsigned)a.a_slots);
            if (!tc_wait(&bars.raw_full[rb], (ib / (unsigned)a.ring_boxes) & 1u, abortp, a.status, 5, a.prof)) break;
            if (!tc_wait(&bars.a_empty[sl], ((ib / (unsigned)a.a_slots) & 1u) ^ 1u, abortp, a.status, 6, a.prof)) break;
            tc::fence_after_sync();
            const float* rp = raw + (size_t)rb * box_floats + m;
            const uint32_t t0 = tmem + ((uint32_t)(32 * q) << 16) + a_col0 + (uint32_t)sl * slot_cols;
            float pa = 1.f, pb = 0.f, pt = 0.f;
            if (a.abt != nullptr) {                                       // the box's volume: tile = blockIdx.x + (ib / n_box) gridDim.x
                const int vol = (int)((blockIdx.x + (ib / (unsigned)n_box) * gridDim.x) / (unsigned)a.tiles_per_vol);
                pa = __ldg(a.abt + 3 * vol); pb = __ldg(a.abt + 3 * vol + 1); pt = __ldg(a.abt + 3 * vol + 2);
            }
            for (int k = 0; k < bs; ++k) {
                float v[kTcRows];
                MVTB_UNROLL
                for (int j = 0; j < kTcRows; ++j) v[j] = rp[(k * kTcRows + j) * 128];
                if (a.abt != nullptr) {
                    MVTB_UNROLL
                    for (int j = 0; j < kTcRows; ++j) v[j] = v[j] != 0.f ? fmaf(pa, v[j], pb) : pt;
                }
                uint32_t hi[kTcRows], lo[kTcRows];
                MVTB_UNROLL
                for (int j = 0; j < kTcRows; ++j) {
                    hi[j] = __float_as_uint(v[j]) & 0xffffe000u;                         // see the note on the split above
                    lo[j] = __float_as_uint(v[j] - __uint_as_float(hi[j]));
                }
                tc::tmem_st16(t0 + 32u * (uint32_t)k, hi);
                tc::tmem_st16(t0 + 32u * (uint32_t)k + 16u, lo);
            }
            tc::tmem_st_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                tc::mbar_arrive(tc::smem_u32(&bars.a_full[sl]));
                tc::mbar_arrive(tc::smem_u32(&bars.raw_empty[rb]));
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (a.prof && blockIdx.x == 0 && tid == 0) a.prof[0] = clock64() - t_start;
    if (warp == 1) tc::tmem_dealloc(tmem, tmem_cols);
}
#endif  // MVTB_EMU
