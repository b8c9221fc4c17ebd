// bandlimited_tc.cuh — the H-axis passes of the band-limited path on the 5th-generation tensor cores.
//
// The pruned DFT along H is a skinny GEMM per column tile:  Y[col, (f, re|im)] = sum_h x[h, col] * T[(f, re|im), h].
// On CUDA cores (bandlimited_quad.cuh) it costs ~16 instructions per voxel and the kernels sit at ~50 % issue
// utilisation, not on HBM.  Here the multiply-adds go to tcgen05.mma (kind::tf32, fp32 accumulators in TMEM) with the
// 3xTF32 split  x*T ~ xh*Th + xl*Th + xh*Tl  (xh = tf32(x), xl = tf32(x - xh); rel-L2 ~1e-6, measured by
// tools/tc_probe.cu), and the SM's threads only convert and move data:
//
//   producer   (1 thread)   1-D bulk copies (cp.async.bulk, UBLKCP) of 16 rows x 128 columns of x into a ring of raw
//                           shared-memory stages, completion counted in bytes on an mbarrier
//   converters (8 warps)    thread = (column, half of the 16 rows): LDS, split into (hi, lo), tcgen05.st into a ring
//                           of A-operand slots in TMEM (lane = column, one 32-bit TMEM column per h)
//   MMA issuer (1 thread)   per 8 rows: D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo, A from TMEM, B (the cos / -sin table,
//                           split on the host) from shared memory; tcgen05.commit frees the A slot
//   epilogue   (4 warps)    tcgen05.ld of the finished 128 x N accumulator, coalesced 8-byte stores into Y[f][col]
//
// All hand-offs are mbarriers; every wait is bounded (a protocol bug sets *status instead of hanging the GPU).
// Y has exactly the layout and meaning k_bl_fwd_h4a produces, so the W/D stage and the inverse pass are unchanged.
// GPU only (no emulator build); included by bandlimited.cu inside namespace mvtb.
#pragma once
#ifndef MVTB_EMU
// tc_common.cuh is included by bandlimited.cu before namespace mvtb opens

static const int kTcRows = 16;          // rows of x per stage (two MMA K-steps of 8)
static const int kTcRawStages = 16;     // raw shared-memory stages (8 KB each), at most; filled box by box (k stages per TMA copy)
static const int kTcASlotsMax = 8;      // A-operand slots in TMEM (32 columns each: 16 hi + 16 lo)
static const int kTcConvGroups = 4;     // converter groups of 4 warps (one per TMEM lane quarter); group g takes stages g, g+4, ...
static const int kTcIssuers = 6;        // MMA-issuing warps: (3xTF32 term, K-step of the stage), each with its own accumulator
static const int kTcWarpIss0 = 1, kTcWarpEpi0 = 1 + kTcIssuers, kTcWarpConv0 = kTcWarpEpi0 + 4;
static const int kTcFwdThreads = 32 * (kTcWarpConv0 + 4 * kTcConvGroups);

struct TcFwdArgs {
    const float* x;          // [nvol][H][NC] (the LDG path; the TMA path reads through the tensor map)
    float2* Y;               // [nvol][NF][NC]
    const float* tab;        // [2][H * N]: B operand hi, lo in tc::op_offset layout; row 2f = cos, 2f+1 = -sin, k = h
    int H, NC, NF, N;        // N = 2 NF rounded up to a multiple of 16
    int n_tiles, tiles_per_vol;
    int box_stages;          // TMA path: stages (of 16 rows) per tensor-map box
    int ring_boxes;          // TMA path: boxes in the raw shared-memory ring
    int* status;
    long long* prof;         // null, or [32 warps][8] wait cycles of CTA 0 (+ [0][0] = total cycles)
};

struct TcBars {
    unsigned long long raw_full[kTcRawStages], raw_empty[kTcRawStages];
    unsigned long long a_full[kTcASlotsMax], a_empty[kTcASlotsMax];
    unsigned long long d_full, d_empty;
};

// bounded wait that also gives up when another role has failed
__device__ __forceinline__ bool tc_wait_raw(unsigned long long* bar, uint32_t parity, volatile int* abort_flag, int* status, int code) {
    const uint32_t b = tc::smem_u32(bar);
    // try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or ~the hint) instead
    // of polling -- with two dozen waiting warps per SM, polling starved the few that had work of issue slots
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

// MVTB_TC_PROF timeline: (event, stage counter, cycle) of CTA 0's first events per warp into a.prof[256 + warp * 64 ...]
#define TC_TRACE(ev, itv)                                                                                        \
    do {                                                                                                         \
        if (a.prof != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && tr_n < 64)                        \
            a.prof[256 + (threadIdx.x >> 5) * 64 + tr_n++] = ((long long)(ev) << 56) | ((long long)(itv) << 40) | ((clock64() - t_start) & 0xffffffffffLL); \
    } while (0)

// The split of x: hi = x with the 13 low mantissa bits cleared (exactly a tf32 value), lo = x - hi (exact in fp32, at
// most 2^-10 |x|; the tensor core reads its top 11 bits).  Two instructions per voxel.  cvt.rna.tf32.f32 for both parts
// is slightly more accurate but ptxas expands it to ~8 integer instructions each: ~19 instructions per voxel, more than
// the FFMA kernels spend on the whole DFT, and the converter warps became the bottleneck (MVTB_TC_PROF).
// What the measurements on a B200 dictated (tools/tc_rate.cu, MVTB_TC_PROF):
//  * one issuing warp pays ~105 cycles per tcgen05.mma whatever the tile, while the tensor pipe takes ~25 for
//    128 x 32 x 8: so six warps issue, one per (3xTF32 term, K-step), each into its own accumulator; the epilogue adds
//    the six.  One issuer made the 90 MMAs of a tile cost 9.5 k cycles against 5.3 k of HBM time.
//  * a 1-D bulk copy costs its issuing thread ~90 cycles: 16 row copies per stage starve the pipeline (1.4 k cycles
//    per stage).  One 2-D tensor-map copy (TMA) per stage brings the whole 16 x 128 box.
//  * a converter warp spends ~1 k cycles per stage (TMEM store + wait + barrier hand-offs): four groups of warps work
//    on four stages at a time.
// TMA_LOAD = true: x arrives by tensor-map copies into a raw shared-memory ring (producer warp); false: every
// converter thread loads its own column of its group's next stage with coalesced LDG, one stage ahead, in registers.
template <bool TMA_LOAD>
__global__ void __launch_bounds__(kTcFwdThreads, 1)
k_bl_fwd_tc(const __grid_constant__ CUtensorMap tmap, TcFwdArgs a) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ TcBars bars;
    __shared__ uint32_t s_tmem;
    __shared__ int s_abort;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = a.H, NC = a.NC, N = a.N;
    const long long t_start = clock64();
    int tr_n = 0;
    const size_t tab_bytes = (size_t)H * N * sizeof(float);
    float* tab_hi = (float*)tc_smem;
    float* tab_lo = (float*)(tc_smem + tab_bytes);
    float* raw = (float*)(tc_smem + ((2 * tab_bytes + 1023) & ~(size_t)1023));   // [kTcRawStages][16][128]
    const int n_stage = H / kTcRows;
    // TMEM: [A ring: slots x 32][six accumulators of N columns]
    const int a_slots = 6 * N <= 512 - 32 * kTcASlotsMax ? kTcASlotsMax : 4;                    // a power of two
    const int a_shift = a_slots == 8 ? 3 : 2;
    const uint32_t a_col0 = 0, d_col0 = (uint32_t)a_slots * 32u;

    for (size_t i = tid; i < 2 * tab_bytes / 16; i += blockDim.x) ((float4*)tc_smem)[i] = __ldg((const float4*)a.tab + i);
    if (tid == 0) {
        s_abort = 0;
        for (int i = 0; i < kTcRawStages; ++i) { tc::mbar_init(tc::smem_u32(&bars.raw_full[i]), 1); tc::mbar_init(tc::smem_u32(&bars.raw_empty[i]), 4); }
        for (int i = 0; i < kTcASlotsMax; ++i) { tc::mbar_init(tc::smem_u32(&bars.a_full[i]), 4); tc::mbar_init(tc::smem_u32(&bars.a_empty[i]), kTcIssuers); }
        tc::mbar_init(tc::smem_u32(&bars.d_full), kTcIssuers);
        tc::mbar_init(tc::smem_u32(&bars.d_empty), 4);
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

    if (warp == 0) {
        // ------------------------------------------------------------ producer: one TMA box (16 rows x 128 columns) per stage
        unsigned it = 0;
        if (!TMA_LOAD) it = 0xffffffffu;
        // One copy brings `box_stages` stages (measured: a copy costs the issuing thread ~550 cycles and lands ~2.9 k
        // cycles later, so 8 KB copies eight deep delivered 22 B/clk per SM, just under the HBM share of an SM;
        // 24 KB boxes, five deep, keep 120 KB in flight).  Stage s of the ring is `full` when its box has landed; the
        // box is refilled when all its stages have been released.
        const int bs = a.box_stages, ring = a.ring_boxes * bs;            // stages in the ring
        const int n_box = n_stage / bs;
        unsigned ib = 0;                                                  // box counter
        for (int tile = blockIdx.x; TMA_LOAD && tile < a.n_tiles && !*abortp; tile += gridDim.x) {
            const int vol = tile / a.tiles_per_vol, c0 = (tile - vol * a.tiles_per_vol) * 128;
            bool ok = true;
            for (int b = 0; b < n_box; ++b, ++ib) {
                const int rb = (int)(ib % (unsigned)a.ring_boxes);        // ring slot of this box
                const uint32_t ph = ((ib / (unsigned)a.ring_boxes) & 1u) ^ 1u;
                for (int k = 0; k < bs && ok; ++k)                        // every stage of the slot must have been released
                    if (!tc_wait(&bars.raw_empty[rb * bs + k], ph, abortp, a.status, 1, a.prof)) ok = false;
                if (!ok) break;
                if (tc::elect_one()) {
                    const uint32_t full = tc::smem_u32(&bars.raw_full[rb]);
                    tc::mbar_arrive_expect_tx(full, (uint32_t)bs * kTcRows * 128 * 4);  // the whole box, columns past NC arrive as zeros
                    tc::tma_load_2d(tc::smem_u32(raw + (size_t)rb * bs * kTcRows * 128), &tmap, c0, vol * H + b * bs * kTcRows, full);
                }
                __syncwarp();
                TC_TRACE(6, ib);
            }
            if (!ok) break;
        }
        (void)ring;
    } else if (warp < kTcWarpEpi0) {
        // ------------------------------------------------------------ MMA issuers
        const int iss = warp - kTcWarpIss0, term = iss >> 1, ks = iss & 1;   // term 0 = hi*hi, 1 = lo*hi, 2 = hi*lo
        const uint32_t idesc = tc::idesc_tf32(128, N);
        const uint32_t chunk_stride = (uint32_t)(N / 8) * 128u, group_stride = 128u;
        const uint32_t b0 = tc::smem_u32(term == 2 ? tab_lo : tab_hi);
        const uint32_t a_off = (term == 1 ? 16u : 0u) + (uint32_t)ks * 8u;   // lo half of the slot for term 1
        const uint32_t d = tmem + d_col0 + (uint32_t)iss * (uint32_t)N;
        unsigned it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles && !*abortp; tile += gridDim.x, ++tcount) {
            if (!tc_wait(&bars.d_empty, (tcount & 1) ^ 1, abortp, a.status, 2, a.prof)) break;
            tc::fence_after_sync();
            uint32_t acc = 0;
            bool ok = true;
            for (int st = 0; st < n_stage; ++st, ++it) {
                const int sl = it & (a_slots - 1);
                if (!tc_wait(&bars.a_full[sl], (it >> a_shift) & 1, abortp, a.status, 3, a.prof)) { ok = false; break; }
                TC_TRACE(1, it);
                tc::fence_after_sync();
                if (tc::elect_one()) {
                    const uint32_t koff = (uint32_t)(st * (kTcRows / 8) + ks) * 2u * chunk_stride;
                    tc::mma_ts(d, tmem + a_col0 + (uint32_t)sl * 32u + a_off, tc::smem_desc(b0 + koff, chunk_stride, group_stride), idesc, acc);
                    tc::mma_commit(tc::smem_u32(&bars.a_empty[sl]));
                }
                acc = 1;
                __syncwarp();
                TC_TRACE(2, it);
            }
            if (!ok) break;
            if (tc::elect_one()) tc::mma_commit(tc::smem_u32(&bars.d_full));
            __syncwarp();
        }
    } else if (warp < kTcWarpConv0) {
        // ------------------------------------------------------------ epilogue: sum of the six accumulators -> Y[f][col]
        const int q = warp & 3;                                       // TMEM lanes 32 q .. 32 q + 31
        const int m = 32 * q + lane;
        unsigned tcount = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles && !*abortp; tile += gridDim.x, ++tcount) {
            const int vol = tile / a.tiles_per_vol, c0 = (tile - vol * a.tiles_per_vol) * 128;
            if (!tc_wait(&bars.d_full, tcount & 1, abortp, a.status, 4, a.prof)) break;
            TC_TRACE(7, tcount);
            tc::fence_after_sync();
            const bool okc = c0 + m < NC;
            float2* yv = a.Y + ((size_t)vol * a.NF) * NC + c0 + m;
            const uint32_t t0 = tmem + ((uint32_t)(32 * q) << 16) + d_col0;
            for (int c = 0; c < N / 8; ++c) {
                uint32_t v[6][8];
                MVTB_UNROLL
                for (int i = 0; i < 6; ++i) tc::tmem_ld8(t0 + (uint32_t)i * (uint32_t)N + 8u * c, v[i]);
                tc::tmem_ld_wait();
                if (c == N / 8 - 1) {                                // everything is in registers: release the accumulators
                    tc::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bars.d_empty));
                    TC_TRACE(8, tcount);
                }
                MVTB_UNROLL
                for (int j = 0; j < 4; ++j) {
                    const int f = 4 * c + j;
                    float r[2];
                    MVTB_UNROLL
                    for (int e = 0; e < 2; ++e) {                     // small terms first, then the two hi*hi halves
                        const float sm = (__uint_as_float(v[2][2 * j + e]) + __uint_as_float(v[3][2 * j + e])) +
                                         (__uint_as_float(v[4][2 * j + e]) + __uint_as_float(v[5][2 * j + e]));
                        r[e] = sm + (__uint_as_float(v[0][2 * j + e]) + __uint_as_float(v[1][2 * j + e]));
                    }
                    if (okc && f < a.NF) yv[(size_t)f * NC] = make_float2(r[0], r[1]);
                }
            }
        }
    } else {
        // ------------------------------------------------------------ converters: rows of x -> (hi, lo) in TMEM
        const int grp = (warp - kTcWarpConv0) >> 2;                  // this group takes stages grp, grp + 4, ...
        const int q = warp & 3;                                       // TMEM lane quarter
        const int m = 32 * q + lane;
        if (TMA_LOAD) {
            unsigned it = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles && !*abortp; tile += gridDim.x) {
                bool ok = true;
                for (int st = 0; st < n_stage; ++st, ++it) {
                    if ((int)(it % kTcConvGroups) != grp) continue;
                    const int sl = it & (a_slots - 1);
                    const unsigned ibx = it / (unsigned)a.box_stages;                      // box counter (n_stage % box_stages == 0)
                    const int kb = (int)(it - ibx * (unsigned)a.box_stages);               // stage within the box
                    const int rb = (int)(ibx % (unsigned)a.ring_boxes);
                    const int s = rb * a.box_stages + kb;
                    if (!tc_wait(&bars.raw_full[rb], (ibx / (unsigned)a.ring_boxes) & 1u, abortp, a.status, 5, a.prof)) { ok = false; break; }
                    TC_TRACE(3, it);
                    const float* rp = raw + (size_t)s * kTcRows * 128 + m;
                    float v[kTcRows];
                    MVTB_UNROLL
                    for (int j = 0; j < kTcRows; ++j) v[j] = rp[j * 128];
                    uint32_t hi[kTcRows], lo[kTcRows];
                    MVTB_UNROLL
                    for (int j = 0; j < kTcRows; ++j) {
                        hi[j] = __float_as_uint(v[j]) & 0xffffe000u;             // tc_split below
                        lo[j] = __float_as_uint(v[j] - __uint_as_float(hi[j]));
                    }
                    if (!tc_wait(&bars.a_empty[sl], ((it >> a_shift) & 1) ^ 1, abortp, a.status, 6, a.prof)) { ok = false; break; }
                    TC_TRACE(4, it);
                    tc::fence_after_sync();
                    const uint32_t t0 = tmem + ((uint32_t)(32 * q) << 16) + a_col0 + (uint32_t)sl * 32u;
                    tc::tmem_st16(t0, hi);
                    tc::tmem_st16(t0 + 16u, lo);
                    tc::tmem_st_wait();
                    tc::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) {
                        tc::mbar_arrive(tc::smem_u32(&bars.a_full[sl]));
                        tc::mbar_arrive(tc::smem_u32(&bars.raw_empty[s]));
                    }
                    TC_TRACE(5, it);
                }
                if (!ok) break;
            }
        } else {
            // own stages: it = grp + 4 k; stage `it` is rows (it % n_stage) * 16 .. of tile number it / n_stage of this CTA
            const unsigned n_it = (unsigned)n_stage * (unsigned)((a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
            auto load_stage = [&](unsigned it, float* v) {
                const int ord = (int)(it / (unsigned)n_stage), st = (int)(it - (unsigned)ord * (unsigned)n_stage);
                const int tile = (int)blockIdx.x + ord * (int)gridDim.x;
                const int vol = tile / a.tiles_per_vol, c0 = (tile - vol * a.tiles_per_vol) * 128;
                const bool okc = c0 + m < NC;
                const float* src = a.x + ((size_t)vol * H + (size_t)st * kTcRows) * NC + c0 + (okc ? m : 0);
                MVTB_UNROLL
                for (int j = 0; j < kTcRows; ++j) v[j] = okc ? __ldcs(src + (size_t)j * NC) : 0.f;
            };
            float vn[kTcRows];
            unsigned it = (unsigned)grp;
            if (it < n_it) load_stage(it, vn);
            for (; it < n_it && !*abortp; it += kTcConvGroups) {
                float v[kTcRows];
                MVTB_UNROLL
                for (int j = 0; j < kTcRows; ++j) v[j] = vn[j];
                if (it + kTcConvGroups < n_it) load_stage(it + kTcConvGroups, vn);     // the next own stage is in flight while this one converts
                const int sl = it & (a_slots - 1);
                uint32_t hi[kTcRows], lo[kTcRows];
                MVTB_UNROLL
                for (int j = 0; j < kTcRows; ++j) {
                    hi[j] = __float_as_uint(v[j]) & 0xffffe000u;                 // tc_split below
                    lo[j] = __float_as_uint(v[j] - __uint_as_float(hi[j]));
                }
                if (!tc_wait(&bars.a_empty[sl], ((it >> a_shift) & 1) ^ 1, abortp, a.status, 6, a.prof)) break;
                TC_TRACE(4, it);
                tc::fence_after_sync();
                const uint32_t t0 = tmem + ((uint32_t)(32 * q) << 16) + a_col0 + (uint32_t)sl * 32u;
                tc::tmem_st16(t0, hi);
                tc::tmem_st16(t0 + 16u, lo);
                tc::tmem_st_wait();
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bars.a_full[sl]));
                TC_TRACE(5, it);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, tmem_cols);
    if (a.prof && blockIdx.x == 0 && tid == 0) a.prof[0] = clock64() - t_start;
}
#endif  // MVTB_EMU
