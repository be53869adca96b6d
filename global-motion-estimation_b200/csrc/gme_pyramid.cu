// gme_pyramid.cu -- one Gaussian pyramid level (K3), HBM-bound.
//
// Replaces cv2.pyrDown as called by utils.get_pyramids (utils.py:34-51): separable
// [1 4 6 4 1] x [1 4 6 4 1], BORDER_REFLECT_101, out = (sum + 128) >> 8, output
// ((H+1)/2, (W+1)/2).  Integer arithmetic, bit-exact against OpenCV.
//
// A thread produces 8 horizontally adjacent output pixels for RY consecutive output rows.
// Per input row it issues one 128-bit load (16 pixels) plus the two neighbouring words,
// filters horizontally on packed 16-bit pairs (two pixels per 32-bit lane op; the row sums
// fit 12 bits and the 5x5 sum + 128 fits 16 bits, so the packed lanes never carry), and
// keeps the five row sums of the vertical tap in a rolling register window, so every
// input byte is read from DRAM once.
#include "gme_common.cuh"

namespace gme {

__device__ __forceinline__ int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

struct PyrArgs {
    const uint8_t *src;
    size_t sp, sstride;
    uint8_t *dst;
    size_t dp, dstride;
    int H, W, Ho, Wo;
    int vec_ok;   // src rows 16-byte aligned, dst rows 8-byte aligned
};

// horizontal pass for the 8 outputs whose centres are input columns ix0, ix0+2, ..., ix0+14 of (virtual) row r;
// result: 4 registers of packed 16-bit pairs (h0,h1) (h2,h3) (h4,h5) (h6,h7)
__device__ __forceinline__ void hrow(const PyrArgs &a, const uint8_t *plane, int r, int ix0, bool fast, uint32_t (&h)[4])
{
    const uint8_t *row = plane + (size_t)reflect101(r, a.H) * a.sp;
    uint32_t w[6];   // words covering columns ix0-4 .. ix0+19
    if (fast) {
        const uint4 m = *reinterpret_cast<const uint4 *>(row + ix0);
        w[0] = *reinterpret_cast<const uint32_t *>(row + ix0 - 4);
        w[1] = m.x; w[2] = m.y; w[3] = m.z; w[4] = m.w;
        w[5] = *reinterpret_cast<const uint32_t *>(row + ix0 + 16);
    } else {
        // borders (reflect-101) and unaligned planes: only columns ix0-2 .. ix0+16 matter
#pragma unroll
        for (int i = 0; i < 6; i++) {
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int col = ix0 - 4 + 4 * i + k;
                if (col >= ix0 - 2 && col <= ix0 + 16) v |= (uint32_t)row[reflect101(col, a.W)] << (8 * k);
            }
            w[i] = v;
        }
    }
    uint32_t e[6], o[6];   // even / odd columns of each word as 16-bit pairs
#pragma unroll
    for (int i = 0; i < 6; i++) {
        e[i] = w[i] & 0x00FF00FFu;
        o[i] = (w[i] >> 8) & 0x00FF00FFu;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t em = __funnelshift_r(e[j], e[j + 1], 16);       // (E[2j-1], E[2j])
        const uint32_t ep = __funnelshift_r(e[j + 1], e[j + 2], 16);   // (E[2j+1], E[2j+2])
        const uint32_t om = __funnelshift_r(o[j], o[j + 1], 16);       // (O[2j-1], O[2j])
        h[j] = em + ep + 6u * e[j + 1] + 4u * (om + o[j + 1]);
    }
}

template <int RY>
__global__ void __launch_bounds__(256) pyr_down_kernel(PyrArgs a)
{
    const int tx = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy0 = (blockIdx.y * blockDim.y + threadIdx.y) * RY;
    const int ox0 = tx * 8;
    if (ox0 >= a.Wo || oy0 >= a.Ho) return;
    const uint8_t *splane = a.src + (size_t)blockIdx.z * a.sstride;
    uint8_t *dplane = a.dst + (size_t)blockIdx.z * a.dstride;
    const int ix0 = ox0 * 2;
    const bool fast = a.vec_ok && ix0 >= 4 && ix0 + 20 <= a.W;
    const bool full = ox0 + 8 <= a.Wo;

    uint32_t hw[5][4];
    const int r0 = 2 * oy0 - 2;
#pragma unroll
    for (int i = 0; i < 3; i++) hrow(a, splane, r0 + i, ix0, fast, hw[i + 2]);
#pragma unroll
    for (int t = 0; t < RY; t++) {
        const int oy = oy0 + t;
        if (oy >= a.Ho) break;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) hw[i][j] = hw[i + 2][j];
        hrow(a, splane, 2 * oy + 1, ix0, fast, hw[3]);
        hrow(a, splane, 2 * oy + 2, ix0, fast, hw[4]);
        uint32_t t2[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t s = hw[0][j] + hw[4][j] + 4u * (hw[1][j] + hw[3][j]) + 6u * hw[2][j] + 0x00800080u;
            t2[j] = (s >> 8) & 0x00FF00FFu;
        }
        const uint32_t lo = __byte_perm(t2[0], t2[1], 0x6420), hi = __byte_perm(t2[2], t2[3], 0x6420);
        uint8_t *out = dplane + (size_t)oy * a.dp + ox0;
        if (full && a.vec_ok) {
            *reinterpret_cast<uint2 *>(out) = make_uint2(lo, hi);
        } else {
            const int nvalid = min(8, a.Wo - ox0);
            for (int k = 0; k < nvalid; k++) out[k] = (uint8_t)(((k < 4 ? lo : hi) >> (8 * (k & 3))) & 0xFF);
        }
    }
}

int launch_pyr_down(const uint8_t *src, size_t sp, size_t sstride, uint8_t *dst, size_t dp, size_t dstride, int n,
                    int H, int W, cudaStream_t stream)
{
    PyrArgs a;
    a.src = src; a.sp = sp; a.sstride = sstride;
    a.dst = dst; a.dp = dp; a.dstride = dstride;
    a.H = H; a.W = W; a.Ho = (H + 1) / 2; a.Wo = (W + 1) / 2;
    a.vec_ok = ((reinterpret_cast<uintptr_t>(src) | sp | sstride) % 16 == 0 &&
                (reinterpret_cast<uintptr_t>(dst) | dp | dstride) % 8 == 0) ? 1 : 0;
    constexpr int RY = 8;
    dim3 block(64, 4);
    const int tx = (a.Wo + 7) / 8, ty = (a.Ho + RY - 1) / RY;
    dim3 grid((tx + block.x - 1) / block.x, (ty + block.y - 1) / block.y, n);
    pyr_down_kernel<RY><<<grid, block, 0, stream>>>(a);
    note_launch();
    return check_launch("pyr_down_kernel");
}

}  // namespace gme
