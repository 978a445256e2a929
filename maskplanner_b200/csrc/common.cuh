// Shared helpers for the libmaskplanner_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/maskplanner_b200.h"

namespace mpb {

void set_error(const char *fmt, ...);

inline int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return MPB_ERR_CUDA;
    }
    return MPB_OK;
}

#define MPB_REQUIRE(cond, msg)                                          \
    do {                                                                \
        if (!(cond)) {                                                  \
            ::mpb::set_error("%s: %s", __func__, msg);                  \
            return MPB_ERR_INVALID_ARGUMENT;                            \
        }                                                               \
    } while (0)

#define MPB_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e_ = (call);                                                    \
        if (e_ != cudaSuccess) {                                                    \
            ::mpb::set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e_)); \
            return MPB_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

int sm_count();

// Opt a kernel into `bytes` of dynamic shared memory once (cached per call site): after the first call
// no cudaFuncSetAttribute is issued any more, which keeps later launches legal inside CUDA-graph capture.
#define MPB_ENSURE_DYN_SMEM(kernel, bytes)                                                               \
    do {                                                                                                 \
        static int configured_ = 0;                                                                      \
        if ((int)(bytes) > configured_) {                                                                \
            MPB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            configured_ = (int)(bytes);                                                                  \
        }                                                                                                \
    } while (0)

// ---- warp-level integer reductions (REDUX.*; one instruction on sm_80+) ----------------------
__device__ __forceinline__ int redux_max_s32(int v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ int redux_min_s32(int v) { return __reduce_min_sync(0xffffffffu, v); }
__device__ __forceinline__ unsigned redux_min_u32(unsigned v) { return __reduce_min_sync(0xffffffffu, v); }

// ---- packed fp32x2 arithmetic and contraction ----------------------------------------------------
// ptxas 12.9 fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even though both carry an explicit .rn
// (seen in SASS with the intrinsics, with inline PTX, with fma(1, a, b) as the sum and with
// -fmad=false alike).  Kernels with a bit-exact "every product and every sum individually rounded"
// contract (FPS) therefore keep packed ops for differences and products only and add the products
// with SCALAR __fadd_rn, which ptxas never contracts.
__device__ __forceinline__ float2 add2_products_rn(float2 a, float2 b)
{
    return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
}

// ---- division of 32-bit indices by a launch-time constant --------------------------------------
// Granlund-Montgomery round-up method: exact for every 32-bit n and d >= 1, four ALU instructions instead of the
// ~20 (through the XU pipe) of a runtime udiv.  Built on the host, passed by value.
struct FastDiv {
    uint32_t d, m, s;
    __device__ __forceinline__ uint32_t div(uint32_t n) const
    {
        if (d == 1) return n;
        const uint32_t t = __umulhi(m, n);
        return (t + ((n - t) >> 1)) >> (s - 1);
    }
};
inline FastDiv make_fastdiv(uint32_t d)
{
    FastDiv f;
    f.d = d, f.m = 0, f.s = 0;
    if (d > 1) {
        uint32_t s = 0;
        while ((1ull << s) < d) ++s;
        f.s = s;
        f.m = (uint32_t)(((1ull << 32) * ((1ull << s) - d)) / d + 1);
    }
    return f;
}

// ---- thread-block-cluster primitives (raw PTX, no cooperative_groups dependency) --------------
__device__ __forceinline__ unsigned cluster_ctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_nctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive_release()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire()
{
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Address of `smem_ptr` (a shared-memory address of this CTA) as seen in CTA `rank` of the cluster.
__device__ __forceinline__ unsigned map_shared_rank(const void *smem_ptr, unsigned rank)
{
    unsigned local = (unsigned)__cvta_generic_to_shared(smem_ptr), remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    return remote;
}
__device__ __forceinline__ void st_shared_cluster_v4(unsigned addr, int a, int b, int c, int d)
{
    asm volatile("st.shared::cluster.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}
__device__ __forceinline__ void st_shared_cluster_s32(unsigned addr, int a)
{
    asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}

}  // namespace mpb
