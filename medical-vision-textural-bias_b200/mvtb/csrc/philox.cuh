// philox.cuh — counter-based Philox4x32-10, shared by the voxel kernels and the fused inverse + select kernel.
#pragma once
#include "mvtb_common.cuh"

namespace mvtb {

// ------------------------------------------------------------------ Philox4x32-10 (Salmon et al., SC'11)
struct Philox {
    static __device__ __forceinline__ uint4 run(uint4 c, uint2 k) {
        const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
        MVTB_UNROLL
        for (int r = 0; r < 10; ++r) {
            const unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
            const unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
            c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
            k.x += W0;
            k.y += W1;
        }
        return c;
    }
    // 24-bit uniform in [0, 1): the grid torch.rand's CPU generator also lands on
    static __device__ __forceinline__ float to_unit(unsigned r) { return (float)(r >> 8) * 5.9604644775390625e-8f; }
};

}  // namespace mvtb
