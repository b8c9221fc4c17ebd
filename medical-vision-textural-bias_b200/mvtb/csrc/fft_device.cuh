// fft_device.cuh — batched in-place mixed-radix FFT stages on shared memory.
//
// Forward = decimation in frequency: natural order in, mixed-radix digit-reversed order out
// (position p holds bin pos2k[p]).  Inverse = decimation in time over the same positions:
// digit-reversed in, natural out.  Nothing is ever permuted: the k-space pointwise stage
// looks frequencies up through pos2k, so forward -> pointwise -> inverse needs no transpose,
// no ping-pong buffer and no fftshift (the reference's fftshift/ifftshift copies,
// F:270-279, are index arithmetic here).
//
// A stage with radix R over blocks of length L (sub-length m = L/R):
//   fwd:  v_q = sum_p x[j + p m] w_R^{pq};  x[j + q m] = v_q * w_L^{jq}
//   inv:  v_q = x[j + q m] * conj(w_L^{jq});  x[j + p m] = sum_q v_q conj(w_R^{pq})
// w_L^{jq} = tw[j q (n/L)] with j q < L, so the table index never wraps.
#pragma once
#include "mvtb_common.cuh"
#include "prime_tables.cuh"

namespace mvtb {

template <int R, bool INV>
struct Butterfly;

template <bool INV>
struct Butterfly<2, INV> {
    static __device__ __forceinline__ void run(cf* v, const cf*, int) {
        cf a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <bool INV>
struct Butterfly<4, INV> {
    static __device__ __forceinline__ void run(cf* v, const cf*, int) {
        cf s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]);
        cf s13 = cadd(v[1], v[3]), d13 = csub(v[1], v[3]);
        cf r = INV ? cmuli(d13) : cmulni(d13);   // (+-i)(x1 - x3)
        v[0] = cadd(s02, s13);
        v[1] = cadd(d02, r);
        v[2] = csub(s02, s13);
        v[3] = csub(d02, r);
    }
};

template <bool INV>
struct Butterfly<3, INV> {
    static __device__ __forceinline__ void run(cf* v, const cf*, int) {
        const float S = 0.86602540378443864676f;
        cf bc = cadd(v[1], v[2]);
        cf t = cmk(v[0].x - 0.5f * bc.x, v[0].y - 0.5f * bc.y);
        cf s = cscale(csub(v[1], v[2]), S);
        cf r = INV ? cmuli(s) : cmulni(s);
        v[0] = cadd(v[0], bc);
        v[1] = cadd(t, r);
        v[2] = csub(t, r);
    }
};

template <bool INV>
struct Butterfly<5, INV> {
    static __device__ __forceinline__ void run(cf* v, const cf*, int) {
        const float C1 = 0.30901699437494742410f, C2 = -0.80901699437494742410f;
        const float S1 = 0.95105651629515357212f, S2 = 0.58778525229247312917f;
        cf t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
        cf t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
        cf a = v[0];
        cf m1 = cmk(a.x + C1 * t1.x + C2 * t2.x, a.y + C1 * t1.y + C2 * t2.y);
        cf m2 = cmk(a.x + C2 * t1.x + C1 * t2.x, a.y + C2 * t1.y + C1 * t2.y);
        cf n1 = cmk(S1 * t3.x + S2 * t4.x, S1 * t3.y + S2 * t4.y);
        cf n2 = cmk(S2 * t3.x - S1 * t4.x, S2 * t3.y - S1 * t4.y);
        cf r1 = INV ? cmuli(n1) : cmulni(n1);
        cf r2 = INV ? cmuli(n2) : cmulni(n2);
        v[0] = cadd(a, cadd(t1, t2));
        v[1] = cadd(m1, r1);
        v[4] = csub(m1, r1);
        v[2] = cadd(m2, r2);
        v[3] = csub(m2, r2);
    }
};

// ---- exact division of small non-negative ints by a loop-invariant divisor: x / d = umulhi(x, M) when x d < 2^32,
// M = floor(2^32 / d) + 1 (2^32 / d itself when d is a power of two), computed on the host (PassDev)
struct FastDiv { int d; unsigned M; };
__device__ __forceinline__ FastDiv fastdiv_from(int d, unsigned M) { FastDiv f; f.d = d; f.M = M; return f; }
__device__ __forceinline__ int fastdiv(int x, const FastDiv& f) { return f.d > 1 ? (int)__umulhi((unsigned)x, f.M) : x; }

// Which (sequence, butterfly) tasks a thread runs.  SEQ_FAST kernels (sequences fastest across threads,
// nthr % count == 0) give every thread one fixed sequence `sq` and butterflies b0, b0 + bstep, ...;
// the others enumerate task = tid, tid + nthr, ... with butterflies fastest.
struct TaskMap { int sq, b0, bstep; };

// One butterfly (radix R <= 5) with its stage twiddles, in place at p[q * es], q = 0..R-1; j = index inside the block.
template <int R, bool INV, bool PRIME = (R >= 7)>
struct StageTask {
    static __device__ __forceinline__ void run(cf* p, int es, int j, int tstep, const cf* __restrict__ tw) {
        cf v[R];
        MVTB_UNROLL
        for (int q = 0; q < R; ++q) v[q] = p[q * es];
        // w^q, q = 1..R-1, from one table load and R-2 complex products: the table loads go through the same LSU
        // pipe as the tile traffic, which is the busiest unit of these kernels; the FMA pipe has room
        if (INV) {
            if (j != 0) {
                const cf w1 = __ldg(tw + j * tstep);
                cf wq = w1;
                MVTB_UNROLL
                for (int q = 1; q < R; ++q) {
                    v[q] = cmulc(v[q], wq);
                    if (q + 1 < R) wq = cmul(wq, w1);
                }
            }
            Butterfly<R, true>::run(v, tw, 0);
        } else {
            Butterfly<R, false>::run(v, tw, 0);
            if (j != 0) {
                const cf w1 = __ldg(tw + j * tstep);
                cf wq = w1;
                MVTB_UNROLL
                for (int q = 1; q < R; ++q) {
                    v[q] = cmul(v[q], wq);
                    if (q + 1 < R) wq = cmul(wq, w1);
                }
            }
        }
        MVTB_UNROLL
        for (int q = 0; q < R; ++q) p[q * es] = v[q];
    }
};

// Odd prime P >= 7 by the symmetric direct DFT, streamed:
//   y_q, y_{P-q} = A_q -+ i B_q,  A_q = x0 + sum_k (x_k + x_{P-k}) cos(2 pi kq/P),
//                                 B_q = sum_k (x_k - x_{P-k}) sin(2 pi kq/P),  k = 1..(P-1)/2
// Only the P-1 accumulators live in registers: the inputs are read from shared memory pair by pair, and with
// both loops unrolled the roots are compile-time literals (prime_tables.cuh) that end up as FFMA immediates.
// (Holding all P inputs as well needed 255 registers for P = 31 and one CTA per SM.)
template <int P, bool INV>
struct StageTask<P, INV, true> {
    static __device__ __forceinline__ void run(cf* p, int es, int j, int tstep, const cf* __restrict__ tw) {
        constexpr int H = (P - 1) / 2;
        const cf x0 = p[0];
        cf tot = x0;
        cf A[H], B[H];
        MVTB_UNROLL
        for (int q = 0; q < H; ++q) { A[q] = x0; B[q] = cmk(0.f, 0.f); }
        MVTB_UNROLL
        for (int k = 1; k <= H; ++k) {
            cf a = p[k * es], b = p[(P - k) * es];
            if (INV && j != 0) {
                a = cmulc(a, __ldg(tw + j * k * tstep));
                b = cmulc(b, __ldg(tw + j * (P - k) * tstep));
            }
            const cf sm = cadd(a, b), df = csub(a, b);
            tot = cadd(tot, sm);
            MVTB_UNROLL
            for (int q = 1; q <= H; ++q) {
                const float c = PrimeTab<P>::c((k * q) % P), sn = PrimeTab<P>::s((k * q) % P);
                A[q - 1].x = fmaf(sm.x, c, A[q - 1].x);
                A[q - 1].y = fmaf(sm.y, c, A[q - 1].y);
                B[q - 1].x = fmaf(df.x, sn, B[q - 1].x);
                B[q - 1].y = fmaf(df.y, sn, B[q - 1].y);
            }
        }
        p[0] = tot;
        MVTB_UNROLL
        for (int q = 1; q <= H; ++q) {
            const cf r = INV ? cmuli(B[q - 1]) : cmulni(B[q - 1]);
            cf yp = cadd(A[q - 1], r), ym = csub(A[q - 1], r);
            if (!INV && j != 0) {
                yp = cmul(yp, __ldg(tw + j * q * tstep));
                ym = cmul(ym, __ldg(tw + j * (P - q) * tstep));
            }
            p[q * es] = yp;
            p[(P - q) * es] = ym;
        }
    }
};

// One radix-R stage over `count` sequences living in shared memory.
// element (seq s, index j) is at  base[s * seq_stride + j * elem_stride].
template <int R, bool INV, bool SEQ_FAST>
__device__ __forceinline__ void fft_stage(cf* base, int seq_stride, int elem_stride, int count, const PassDev& ps,
                                          const cf* __restrict__ tw, const TaskMap& tm, int tid, int nthr) {
    const int m = ps.m, L = ps.L, tstep = ps.ts1;
    const int es = m * elem_stride;
    const FastDiv dm = fastdiv_from(m, ps.magic_m);
    if (SEQ_FAST) {
        cf* sbase = base + (size_t)tm.sq * seq_stride;
        for (int b = tm.b0; b < ps.per_seq; b += tm.bstep) {
            const int blk = fastdiv(b, dm), j = b - blk * m;
            StageTask<R, INV>::run(sbase + (size_t)(blk * L + j) * elem_stride, es, j, tstep, tw);
        }
    } else {
        const FastDiv dp = fastdiv_from(ps.per_seq, ps.magic_ps);
        const int total = ps.per_seq * count;
        for (int task = tid; task < total; task += nthr) {
            const int sq = fastdiv(task, dp), b = task - sq * ps.per_seq;
            const int blk = fastdiv(b, dm), j = b - blk * m;
            StageTask<R, INV>::run(base + (size_t)sq * seq_stride + (size_t)(blk * L + j) * elem_stride, es, j, tstep, tw);
        }
    }
}

// Two consecutive stages (radices R1 then R2, both <= 5) fused in registers: the same arithmetic, tables and
// positions as running fft_stage<R1> and fft_stage<R2> back to back, but one pass over shared memory and one
// barrier instead of two.  A task owns the R1*R2 elements  blk*L + p1*(L/R1) + p2*(L/(R1 R2)) + j2.
template <int R1, int R2, bool INV>
__device__ __forceinline__ void fft_task2(cf* p, int e1, int e2, int m2, int j2, int ts1, int ts2, const cf* __restrict__ tw) {
    {
        cf v[R1][R2];
        MVTB_UNROLL
        for (int a = 0; a < R1; ++a) {
            MVTB_UNROLL
            for (int c = 0; c < R2; ++c) v[a][c] = p[a * e1 + c * e2];
        }
        if (!INV) {
            MVTB_UNROLL
            for (int c = 0; c < R2; ++c) {                  // stage 1: over p1, twiddle w_L^(j1 q1), j1 = c*m2 + j2
                cf t[R1];
                MVTB_UNROLL
                for (int a = 0; a < R1; ++a) t[a] = v[a][c];
                Butterfly<R1, false>::run(t, tw, 0);
                const cf w1 = __ldg(tw + (c * m2 + j2) * ts1);
                cf wa = w1;
                MVTB_UNROLL
                for (int a = 1; a < R1; ++a) {
                    t[a] = cmul(t[a], wa);
                    if (a + 1 < R1) wa = cmul(wa, w1);
                }
                MVTB_UNROLL
                for (int a = 0; a < R1; ++a) v[a][c] = t[a];
            }
            cf w2[R2];                                      // stage 2 twiddles w_m1^(j2 c): the same for every a
            w2[0] = cmk(1.f, 0.f);
            w2[1] = __ldg(tw + j2 * ts2);
            MVTB_UNROLL
            for (int c = 2; c < R2; ++c) w2[c] = cmul(w2[c - 1], w2[1]);
            MVTB_UNROLL
            for (int a = 0; a < R1; ++a) {                  // stage 2: over p2
                Butterfly<R2, false>::run(v[a], tw, 0);
                if (j2 != 0) {
                    MVTB_UNROLL
                    for (int c = 1; c < R2; ++c) v[a][c] = cmul(v[a][c], w2[c]);
                }
            }
        } else {
            cf w2[R2];
            w2[0] = cmk(1.f, 0.f);
            w2[1] = __ldg(tw + j2 * ts2);
            MVTB_UNROLL
            for (int c = 2; c < R2; ++c) w2[c] = cmul(w2[c - 1], w2[1]);
            MVTB_UNROLL
            for (int a = 0; a < R1; ++a) {
                if (j2 != 0) {
                    MVTB_UNROLL
                    for (int c = 1; c < R2; ++c) v[a][c] = cmulc(v[a][c], w2[c]);
                }
                Butterfly<R2, true>::run(v[a], tw, 0);
            }
            MVTB_UNROLL
            for (int c = 0; c < R2; ++c) {
                cf t[R1];
                const cf w1 = __ldg(tw + (c * m2 + j2) * ts1);
                cf wa = w1;
                t[0] = v[0][c];
                MVTB_UNROLL
                for (int a = 1; a < R1; ++a) {
                    t[a] = cmulc(v[a][c], wa);
                    if (a + 1 < R1) wa = cmul(wa, w1);
                }
                Butterfly<R1, true>::run(t, tw, 0);
                MVTB_UNROLL
                for (int a = 0; a < R1; ++a) v[a][c] = t[a];
            }
        }
        MVTB_UNROLL
        for (int a = 0; a < R1; ++a) {
            MVTB_UNROLL
            for (int c = 0; c < R2; ++c) p[a * e1 + c * e2] = v[a][c];
        }
    }
}

template <int R1, int R2, bool INV, bool SEQ_FAST>
__device__ __forceinline__ void fft_stage2(cf* base, int seq_stride, int elem_stride, int count, const PassDev& ps,
                                           const cf* __restrict__ tw, const TaskMap& tm, int tid, int nthr) {
    const int m2 = ps.m, m1 = m2 * R2, L = ps.L;
    const int e1 = m1 * elem_stride, e2 = m2 * elem_stride;
    const FastDiv dm = fastdiv_from(m2, ps.magic_m);
    if (SEQ_FAST) {
        cf* sbase = base + (size_t)tm.sq * seq_stride;
        for (int b = tm.b0; b < ps.per_seq; b += tm.bstep) {
            const int blk = fastdiv(b, dm), j2 = b - blk * m2;
            fft_task2<R1, R2, INV>(sbase + (size_t)(blk * L + j2) * elem_stride, e1, e2, m2, j2, ps.ts1, ps.ts2, tw);
        }
    } else {
        const FastDiv dp = fastdiv_from(ps.per_seq, ps.magic_ps);
        const int total = ps.per_seq * count;
        for (int task = tid; task < total; task += nthr) {
            const int sq = fastdiv(task, dp), b = task - sq * ps.per_seq;
            const int blk = fastdiv(b, dm), j2 = b - blk * m2;
            fft_task2<R1, R2, INV>(base + (size_t)sq * seq_stride + (size_t)(blk * L + j2) * elem_stride, e1, e2, m2, j2, ps.ts1, ps.ts2, tw);
        }
    }
}

template <bool INV, bool SEQ_FAST>
__device__ __forceinline__ void fft_stage2_dispatch(cf* base, int seq_stride, int elem_stride, int count, const PassDev& ps,
                                                    const cf* __restrict__ tw, const TaskMap& tm, int tid, int nthr) {
    switch (ps.r1 * 8 + ps.r2) {
#define MVTB_CASE2(A, B) case A * 8 + B: fft_stage2<A, B, INV, SEQ_FAST>(base, seq_stride, elem_stride, count, ps, tw, tm, tid, nthr); break;
        MVTB_CASE2(2, 3) MVTB_CASE2(2, 4) MVTB_CASE2(2, 5) MVTB_CASE2(3, 3) MVTB_CASE2(3, 4) MVTB_CASE2(3, 5)
        MVTB_CASE2(4, 4) MVTB_CASE2(4, 5)
#undef MVTB_CASE2
        default: break;
    }
}

// Any other prime radix (> 31): out-of-place direct DFT through a scratch copy of the tile, one task per output
// element, O(R) each.  Slow but it makes every axis length work (181 = MNI, 37, 41, ...).
template <bool INV, bool SEQ_FAST>
__device__ __forceinline__ void fft_stage_generic(int R, cf* base, cf* scratch, int seq_stride, int elem_stride, int count,
                                                  int n, int L, const cf* __restrict__ tw, int tid, int nthr) {
    const int m = L / R;
    const int total = n * count;
    const int tstep = n / L, rstep = n / R;
    for (int task = tid; task < total; task += nthr) {
        int sq, e;
        if (SEQ_FAST) { sq = task % count; e = task / count; }
        else          { e = task % n; sq = task / n; }
        const int blk = e / L, r = e - blk * L;
        const int q = r / m, j = r - q * m;                       // output index q of butterfly (blk, j)
        const cf* src = base + (size_t)sq * seq_stride + (size_t)(blk * L + j) * elem_stride;
        cf acc = cmk(0.f, 0.f);
        int idx = 0;                                              // (p q) mod R
        for (int p = 0; p < R; ++p) {
            cf x = src[(size_t)p * m * elem_stride];
            if (INV && j != 0 && p != 0) x = cmulc(x, __ldg(tw + j * p * tstep));
            const cf w = __ldg(tw + idx * rstep);                 // w_R^(pq) = (cos, -sin)
            acc = cadd(acc, INV ? cmulc(x, w) : cmul(x, w));
            idx += q;
            if (idx >= R) idx -= R;
        }
        if (!INV && j != 0 && q != 0) acc = cmul(acc, __ldg(tw + j * q * tstep));
        scratch[(size_t)sq * seq_stride + (size_t)e * elem_stride] = acc;
    }
    __syncthreads();
    for (int task = tid; task < total; task += nthr) {
        int sq, e;
        if (SEQ_FAST) { sq = task % count; e = task / count; }
        else          { e = task % n; sq = task / n; }
        const size_t o = (size_t)sq * seq_stride + (size_t)e * elem_stride;
        base[o] = scratch[o];
    }
}

// MAXR bounds the radices compiled into a kernel (5: 2/3/4/5, 13: + 7/11/13, 31: all): each unrolled prime
// stage is P^2 FMAs of code, so kernels for axes without big primes are instantiated without them.
template <bool INV, bool SEQ_FAST, int MAXR>
__device__ __forceinline__ void fft_stage_dispatch(cf* base, cf* scratch, int seq_stride, int elem_stride, int count, int n,
                                                   const PassDev& ps, const cf* __restrict__ tw, const TaskMap& tm, int tid, int nthr) {
    const int R = ps.r1;
    switch (R) {
#define MVTB_CASE(RR) case RR: fft_stage<RR, INV, SEQ_FAST>(base, seq_stride, elem_stride, count, ps, tw, tm, tid, nthr); break;
        MVTB_CASE(2) MVTB_CASE(3) MVTB_CASE(4) MVTB_CASE(5)
        default:
            if (MAXR > 5) {
                switch (R) {
                    MVTB_CASE(7) MVTB_CASE(11) MVTB_CASE(13)
                    default:
                        if (MAXR > 13) {
                            switch (R) {
                                MVTB_CASE(17) MVTB_CASE(19) MVTB_CASE(23) MVTB_CASE(29) MVTB_CASE(31)
                                default:
                                    if (scratch) fft_stage_generic<INV, SEQ_FAST>(R, base, scratch, seq_stride, elem_stride, count, n, ps.L, tw, tid, nthr);
                                    break;
                            }
                        }
                        break;
                }
            }
            break;
#undef MVTB_CASE
    }
}

template <bool SEQ_FAST>
__device__ __forceinline__ TaskMap fft_task_map(int count, int tid, int nthr) {
    TaskMap tm;
    tm.sq = 0; tm.b0 = 0; tm.bstep = 1;
    if (SEQ_FAST) { tm.sq = tid % count; tm.b0 = tid / count; tm.bstep = nthr / count; }
    return tm;
}

// Whole transform; the caller has synchronised before, and a __syncthreads() follows every pass.
template <bool SEQ_FAST, int MAXR>
__device__ __forceinline__ void fft_forward(const AxisDev& ax, cf* base, int seq_stride, int elem_stride, int count,
                                            int tid, int nthr, cf* scratch = nullptr) {
    const TaskMap tm = fft_task_map<SEQ_FAST>(count, tid, nthr);
    for (int s = 0; s < ax.npass; ++s) {
        const PassDev& ps = ax.pass[s];
        if (ps.r2) fft_stage2_dispatch<false, SEQ_FAST>(base, seq_stride, elem_stride, count, ps, ax.tw, tm, tid, nthr);
        else fft_stage_dispatch<false, SEQ_FAST, MAXR>(base, scratch, seq_stride, elem_stride, count, ax.n, ps, ax.tw, tm, tid, nthr);
        __syncthreads();
    }
}

template <bool SEQ_FAST, int MAXR>
__device__ __forceinline__ void fft_inverse(const AxisDev& ax, cf* base, int seq_stride, int elem_stride, int count,
                                            int tid, int nthr, cf* scratch = nullptr) {
    const TaskMap tm = fft_task_map<SEQ_FAST>(count, tid, nthr);
    for (int s = ax.npass - 1; s >= 0; --s) {
        const PassDev& ps = ax.pass[s];
        if (ps.r2) fft_stage2_dispatch<true, SEQ_FAST>(base, seq_stride, elem_stride, count, ps, ax.tw, tm, tid, nthr);
        else fft_stage_dispatch<true, SEQ_FAST, MAXR>(base, scratch, seq_stride, elem_stride, count, ax.n, ps, ax.tw, tm, tid, nthr);
        __syncthreads();
    }
}

}  // namespace mvtb
