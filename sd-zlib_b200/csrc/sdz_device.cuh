// sdz_device.cuh - small sm_100a device helpers shared by the kernels:
// mbarrier + cp.async.bulk (TMA bulk copy) wrappers and sub-warp group utilities.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sdz {

__device__ __forceinline__ uint32_t smem_addr(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// arrive (count 1) and announce `bytes` of async transactions for the current phase
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}

// TMA bulk copy global -> shared (1-D, 16-byte aligned, size multiple of 16); completion is
// signalled as `bytes` transaction bytes on `bar`.  SASS: UBLKCP.
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// 4-byte asynchronous copy global -> shared (SASS LDGSTS); completion is tracked with
// commit_group / wait_group, not with a register scoreboard, so the copy can stay in flight
// across loop iterations.
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src_gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(dst_smem)), "l"(src_gmem) : "memory");
}
// predicated forms (no branch, no divergence): the instruction is issued for the whole warp
__device__ __forceinline__ void cp_async4_if(void* dst_smem, const void* src_gmem, bool pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q cp.async.ca.shared.global [%0], [%1], 4;\n\t}" ::"r"(
                     smem_addr(dst_smem)),
                 "l"(src_gmem), "r"((uint32_t)pred)
                 : "memory");
}
__device__ __forceinline__ void st_u8_if(uint8_t* p, uint32_t v, bool pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.u8 [%0], %1;\n\t}" ::"l"(
                     __cvta_generic_to_global(p)),
                 "r"(v), "r"((uint32_t)pred)
                 : "memory");
}
__device__ __forceinline__ void st_u16_if(uint16_t* p, uint32_t v, bool pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.u16 [%0], %1;\n\t}" ::"l"(
                     __cvta_generic_to_global(p)),
                 "h"((uint16_t)v), "r"((uint32_t)pred)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// all groups except the most recently committed one are complete
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

__device__ __forceinline__ uint32_t lane_id()
{
    uint32_t l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

}  // namespace sdz
