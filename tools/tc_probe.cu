// tc_probe.cu — building-block check for the tensor-core pruned-DFT kernels (bandlimited_tc.cuh):
// one CTA computes D[128 x N] = A[128 x K] * B[N x K]^T with tcgen05.mma kind::tf32 from shared-memory operands in
// the no-swizzle K-major canonical layout, accumulator in TMEM, read back with tcgen05.ld, and compares with the
// host.  It answers, on a B200, the questions the PTX manual would (it is not in this image):
//   * which of the two byte offsets of the shared-memory descriptor strides the 8-row groups and which the
//     16-byte K chunks (hypothesis 0 / 1);
//   * exact results on small-integer data (layout proof), N = 32 and N = 240, K advanced by descriptor start address;
//   * accuracy of the 3xTF32 split (hi*hi + lo*hi + hi*lo) on DFT-like data against float64.
// Every wait is bounded; a failed wait prints a code instead of hanging the box.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tc_probe tc_probe.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version: Blackwell
    return d;                                     // base offset 0, layout type 0 = no swizzle
}

__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (int it = 0; it < (1 << 22); ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

// byte offset of element (mn, k) of an operand with MN rows, K columns: [k/4][mn/8][mn%8][4 floats]
__host__ __device__ inline size_t op_off(int mn, int k, int MN) {
    return ((size_t)(k / 4) * (MN / 8) + (mn / 8)) * 128 + (size_t)(mn % 8) * 16 + (size_t)(k % 4) * 4;
}

// A: [npass][128 x K], B: [npass][N x K] already in the canonical layout (host prepared); D = sum over passes
__global__ void __launch_bounds__(128, 1)
k_probe(const float* __restrict__ Ag, const float* __restrict__ Bg, float* __restrict__ D, int N, int K, int npass, int hyp,
        int* __restrict__ status) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t a_bytes = (size_t)128 * K * 4, b_bytes = (size_t)N * K * 4;
    float* sA = (float*)smem;
    float* sB = (float*)(smem + a_bytes * npass);
    for (size_t i = tid; i < a_bytes * npass / 4; i += blockDim.x) sA[i] = Ag[i];
    for (size_t i = tid; i < b_bytes * npass / 4; i += blockDim.x) sB[i] = Bg[i];
    const uint32_t ncols = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // operand tiles written by threads -> async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(128, N);
        // strides: between 8-row groups 128 B; between 16-byte K chunks (MN/8)*128 B
        const uint32_t a_grp = 128, a_chk = 16 * 128, b_grp = 128, b_chk = (uint32_t)(N / 8) * 128;
        uint32_t acc = 0;
        for (int p = 0; p < npass; ++p)
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint32_t a_addr = smem_u32(sA) + (uint32_t)(p * a_bytes) + (uint32_t)ks * 2 * a_chk;
                const uint32_t b_addr = smem_u32(sB) + (uint32_t)(p * b_bytes) + (uint32_t)ks * 2 * b_chk;
                const uint64_t da = hyp == 0 ? make_desc(a_addr, a_chk, a_grp) : make_desc(a_addr, a_grp, a_chk);
                const uint64_t db = hyp == 0 ? make_desc(b_addr, b_chk, b_grp) : make_desc(b_addr, b_grp, b_chk);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                acc = 1;
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
    }
    const bool ok = mbar_wait(smem_u32(&s_bar), 0);
    if (!ok && tid == 0) status[0] = 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok) {
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols) : "memory");
}

static float tf32_rn(float x) {           // round to nearest (ties away), 10 explicit mantissa bits: cvt.rna.tf32.f32
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x1000u;
    u &= 0xffffe000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

static void run_case(const char* name, int N, int K, int npass, int hyp, const std::vector<float>& A, const std::vector<float>& B,
                     const std::vector<double>& ref) {
    // A, B: npass row-major logical operands [128][K], [N][K]
    std::vector<float> Al((size_t)npass * 128 * K), Bl((size_t)npass * N * K);
    for (int p = 0; p < npass; ++p) {
        for (int m = 0; m < 128; ++m)
            for (int k = 0; k < K; ++k) Al[(size_t)p * 128 * K + op_off(m, k, 128) / 4] = A[((size_t)p * 128 + m) * K + k];
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) Bl[(size_t)p * N * K + op_off(n, k, N) / 4] = B[((size_t)p * N + n) * K + k];
    }
    float *dA, *dB, *dD;
    int* dS;
    CK(cudaMalloc(&dA, Al.size() * 4)); CK(cudaMalloc(&dB, Bl.size() * 4)); CK(cudaMalloc(&dD, (size_t)128 * N * 4)); CK(cudaMalloc(&dS, 4));
    CK(cudaMemcpy(dA, Al.data(), Al.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bl.data(), Bl.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, (size_t)128 * N * 4));
    CK(cudaMemset(dS, 0, 4));
    const size_t smem = (Al.size() + Bl.size()) * 4;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_probe<<<1, 128, smem>>>(dA, dB, dD, N, K, npass, hyp, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s hyp=%d  CUDA error: %s\n", name, hyp, cudaGetErrorString(e)); exit(2); }
    std::vector<float> D((size_t)128 * N);
    int st = 0;
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    double num = 0, den = 0, maxabs = 0;
    for (size_t i = 0; i < D.size(); ++i) {
        const double d = (double)D[i] - ref[i];
        num += d * d; den += ref[i] * ref[i];
        if (fabs(d) > maxabs) maxabs = fabs(d);
    }
    printf("%-28s N=%3d K=%3d passes=%d hyp=%d  wait_failed=%d  rel_l2=%.3e  max_abs=%.3e  D[0][0..3]=%g %g %g %g  ref=%g %g %g %g\n", name, N, K, npass,
           hyp, st, sqrt(num / (den > 0 ? den : 1)), maxabs, D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
}

int main() {
    // ---- layout proof: small integers, exact in tf32 and in the fp32 accumulator
    for (int hyp = 0; hyp < 1; ++hyp)
        for (int N : {32, 240}) {
            const int K = 16;
            std::vector<float> A((size_t)128 * K), B((size_t)N * K);
            std::vector<double> ref((size_t)128 * N, 0.0);
            for (int m = 0; m < 128; ++m) for (int k = 0; k < K; ++k) A[(size_t)m * K + k] = (float)((m * 7 + k * 3) % 11 - 5);
            for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) B[(size_t)n * K + k] = (float)((n * 5 + k * 2) % 13 - 6);
            for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
                double s = 0;
                for (int k = 0; k < K; ++k) s += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
                ref[(size_t)m * N + n] = s;
            }
            run_case("integers", N, K, 1, hyp, A, B, ref);
        }
    // ---- 3xTF32 accuracy on DFT-like data: A = x (N(0,1)-ish), B = cos/sin table, K = 96, both hypotheses again
    for (int hyp = 0; hyp < 1; ++hyp)
        for (int N : {32, 240}) {
            const int K = 96;
            std::vector<double> x((size_t)128 * K), t((size_t)N * K);
            unsigned s = 12345u;
            auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) / 16777216.0) * 2.0 - 1.0; };
            for (auto& v : x) v = rnd() * 3.0;
            for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) t[(size_t)n * K + k] = (n & 1) ? sin(2 * M_PI * (n / 2) * k / 240.0) : cos(2 * M_PI * (n / 2) * k / 240.0);
            std::vector<double> ref((size_t)128 * N, 0.0);
            for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
                double acc = 0;
                for (int k = 0; k < K; ++k) acc += (double)(float)x[(size_t)m * K + k] * (double)(float)t[(size_t)n * K + k];
                ref[(size_t)m * N + n] = acc;
            }
            // passes: (hi, hi), (lo, hi), (hi, lo)
            std::vector<float> A((size_t)3 * 128 * K), B((size_t)3 * N * K);
            for (size_t i = 0; i < (size_t)128 * K; ++i) {
                const float v = (float)x[i], hi = tf32_rn(v), lo = tf32_rn(v - hi);
                A[i] = hi; A[(size_t)128 * K + i] = lo; A[(size_t)2 * 128 * K + i] = hi;
            }
            for (size_t i = 0; i < (size_t)N * K; ++i) {
                const float v = (float)t[i], hi = tf32_rn(v), lo = tf32_rn(v - hi);
                B[i] = hi; B[(size_t)N * K + i] = hi; B[(size_t)2 * N * K + i] = lo;
            }
            run_case("3xTF32 dft-like", N, K, 3, hyp, A, B, ref);
            // 1xTF32 for scale
            std::vector<float> A1(A.begin(), A.begin() + (size_t)128 * K), B1(B.begin(), B.begin() + (size_t)N * K);
            run_case("1xTF32 dft-like", N, K, 1, hyp, A1, B1, ref);
            // raw fp32 bits fed as tf32 (does the tensor core truncate or round?)
            std::vector<float> Ar((size_t)128 * K), Br((size_t)N * K);
            for (size_t i = 0; i < Ar.size(); ++i) Ar[i] = (float)x[i];
            for (size_t i = 0; i < Br.size(); ++i) Br[i] = (float)t[i];
            run_case("raw fp32 as tf32", N, K, 1, hyp, Ar, Br, ref);
        }
    return 0;
}
