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
// cudaFuncAttributeMaxDynamicSharedMemorySize, set once per (kernel, device) and raised only when a launch needs more
void ensure_dynamic_smem(const void *kernel, size_t bytes);

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

// The float32 value np.sum gives for the bs*bs squared differences of one block (compute_dfd + mse, bbme.py:61-64,94)
// when block_size > 16, i.e. when the exact sum can exceed 2^24 and the reference's float32 arithmetic ROUNDS
// (SURVEY A.2).  NumPy reduces a contiguous float32 buffer pairwise (numpy/core/src/umath/loops_utils.h.src,
// pairwise_sum_FLOAT; numpy pinned by requirements.txt:2): runs of at most 128 values are summed with eight strided
// accumulators -- every partial sum there is an integer below 2^24, hence exact -- and longer runs are split at
// n/2 rounded down to a multiple of 8, the two halves added in float32.  `leaf(start, len)` returns the exact integer
// sum of the row-major values [start, start + len); the tree above the leaves is followed with round-to-nearest adds.
// The result is an integer-valued float below 2^32 (bs <= 255), returned as the integer.  Oracle: pairwise_sum_f32 in
// oracle/gme_oracle.c; pinned against the reference by tests/golden/bbme_f32_rounding.npz.
template <class Leaf>
__device__ __forceinline__ uint32_t pairwise_sum_f32_tree(int n, Leaf leaf)
{
    if (n <= 128) return leaf(0, n);
    int st_start[12], st_len[12], st_state[12];       // depth <= 9 for n <= 255 * 255
    float st_left[12];
    int sp = 0;
    st_start[0] = 0; st_len[0] = n; st_state[0] = 0;
    float ret = 0.f;
    while (sp >= 0) {
        const int s = st_start[sp], l = st_len[sp];
        if (l <= 128) { ret = (float)leaf(s, l); --sp; continue; }
        int n2 = l / 2;
        n2 -= n2 % 8;
        if (st_state[sp] == 0) {
            st_state[sp] = 1;
            ++sp; st_start[sp] = s; st_len[sp] = n2; st_state[sp] = 0;
        } else if (st_state[sp] == 1) {
            st_left[sp] = ret;
            st_state[sp] = 2;
            ++sp; st_start[sp] = s + n2; st_len[sp] = l - n2; st_state[sp] = 0;
        } else {
            ret = __fadd_rn(st_left[sp], ret);
            --sp;
        }
    }
    return (uint32_t)ret;
}

}  // namespace gme
