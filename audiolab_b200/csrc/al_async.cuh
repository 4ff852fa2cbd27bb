// cp.async (LDGSTS) helpers and the dynamic shared memory declaration used by the staged kernels.
#pragma once
#include <cuda_runtime.h>

#ifndef AL_DYN_SMEM
#define AL_DYN_SMEM(T, name) extern __shared__ __align__(16) unsigned char name##_raw_[]; T* name = reinterpret_cast<T*>(name##_raw_)
#endif

namespace al {

__device__ __forceinline__ void al_cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void al_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void al_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

}  // namespace al
