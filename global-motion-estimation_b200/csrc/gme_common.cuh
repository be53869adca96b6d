// gme_common.cuh -- shared device/host helpers for the sm_100a GME kernels.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gme_b200.h"

namespace gme {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

// ---------------------------------------------------------------------------------------
// launch bookkeeping (host)
// ---------------------------------------------------------------------------------------
void note_launch();                 // counts kernel launches (gme_launch_count)
int check_launch(const char *what); // cudaGetLastError -> GME_OK / GME_ERR_CUDA

// Builds a 3-D (W, H, n) uint8 tensor map with a (box_w, box_h, 1) box.  Returns false when
// the planes cannot travel by TMA (alignment) -- callers then use the cooperative loader.
bool make_plane_tensor_map(CUtensorMap *map, const uint8_t *base, int n, int H, int W, size_t pitch,
                           size_t plane_stride, int box_w, int box_h);

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init()
{
    // make the init visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// TMA: 3-D tiled load global -> shared, completion on an mbarrier (SASS: UTMALDG).
// Out-of-bounds elements (negative or past-the-end coordinates) are zero-filled.
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y,
                                            int z)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// Packed byte arithmetic.  SASS on sm_100a: VABSDIFF4 (.ACC when the add operand is live), IDP.4A.
__device__ __forceinline__ uint32_t sad4_acc(uint32_t a, uint32_t b, uint32_t acc)
{
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}

__device__ __forceinline__ uint32_t absdiff4(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(0u));
    return d;
}

__device__ __forceinline__ uint32_t ssd4_acc(uint32_t a, uint32_t b, uint32_t acc)
{
    uint32_t d = absdiff4(a, b);
    return __dp4a(d, d, acc);   // unsigned 4-way dot product: acc += sum d_i * d_i
}

template <int PNORM>
__device__ __forceinline__ uint32_t cost4_acc(uint32_t a, uint32_t b, uint32_t acc)
{
    if constexpr (PNORM == GME_PNORM_MAE)
        return sad4_acc(a, b, acc);
    else
        return ssd4_acc(a, b, acc);
}

// mask keeping the first nvalid (1..4) bytes of a little-endian word
__device__ __host__ __forceinline__ uint32_t byte_mask(int nvalid)
{
    return nvalid >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nvalid)) - 1u);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

}  // namespace gme
