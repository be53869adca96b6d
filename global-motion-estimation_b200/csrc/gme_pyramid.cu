// gme_pyramid.cu -- one Gaussian pyramid level (K3), HBM-bound.
//
// Replaces cv2.pyrDown as called by utils.get_pyramids (utils.py:34-51): separable
// [1 4 6 4 1] x [1 4 6 4 1], BORDER_REFLECT_101, out = (sum + 128) >> 8, output
// ((H+1)/2, (W+1)/2).  Integer arithmetic, bit-exact against OpenCV.
//
// One CTA produces a 128 x 64 tile of the output.  Its input footprint (131 rows x 288 bytes:
// the tile, the 2-pixel filter halo and padding to 16-byte chunks) is staged in shared memory
// with 16-byte cp.async copies -- every input byte leaves DRAM once, in coalesced 128-bit
// requests, without passing through registers.  Chunks that touch the frame border are filled
// byte by byte through the reflect-101 index map instead, so the filter code never sees a
// border.  A thread then produces 8 adjacent output pixels for 4 consecutive output rows:
// per input row one 128-bit and two 32-bit shared loads, the horizontal tap on packed
// 16-bit pairs (two pixels per 32-bit lane op; row sums fit 12 bits and the 5x5 sum + 128
// fits 16 bits, so the packed lanes never carry), and a rolling window of five row sums for
// the vertical tap.  Output rows leave as 8-byte stores, 128 bytes per 16 threads.
// (Round 2 tried persistent CTAs with two staging buffers, the next tile's cp.async in flight while the current one is
// filtered: 72 us instead of 59 at 1080p -- two 37 KB buffers leave three CTAs per SM instead of six, and the lost
// warps cost more than the hidden staging latency gains.  Fusing the second level into the first does not pay either:
// the kernel is issue-bound, and the 2-pixel halo of the coarser level means recomputing 27 % of the finer one.
// The two 32-bit loads of hrow are 4-way bank conflicts (every lane reads the same word of its 16-byte chunk; 44 % of
// the kernel's shared wavefronts, ncu r02); taking those words from the neighbouring lanes with two shuffles instead
// removed the conflicts and cost more than it saved -- 48 registers instead of 40, 63.6 us instead of 59.3.)
#include "gme_common.cuh"

namespace gme {

constexpr int kPyrTileX = 128;                    // output pixels per tile row
constexpr int kPyrRY = 4;                         // output rows per thread
// output rows per tile: 64 (256 threads, 131 input rows) for large levels, 32 (128 threads) for small ones, where
// the coarser tile grid would waste a quarter of the CTAs on rows past the frame
constexpr int kPyrInPitch = 2 * kPyrTileX + 32;   // 288 bytes: 16 bytes of left padding, tile, halo, padding
constexpr int kPyrChunks = kPyrInPitch / 16;      // 18

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
    int vec_in;    // src rows 16-byte aligned: interior chunks travel by cp.async
    int vec_out;   // dst rows 8-byte aligned: 8-byte stores
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

// horizontal tap for the 8 outputs whose centres are tile bytes c0, c0+2, ..., c0+14 of one staged row:
// 4 registers of packed 16-bit pairs (h0,h1) (h2,h3) (h4,h5) (h6,h7).  Output j covers bytes
// c0+2j-2 .. c0+2j+2: the first four taps are one IDP.4A against (1,4,6,4), the fifth is byte 0 of the next
// word (a second IDP.4A against (1,0,0,0)); for odd j the four bytes are an aligned word, for even j they
// straddle two words (one funnel shift).  The dot products run on the FMA pipe, the shifts on the ALU pipe.
__device__ __forceinline__ void hrow(const uint8_t *row, int c0, uint32_t (&h)[4])
{
    uint32_t w[6];   // words covering tile bytes c0-4 .. c0+19
    const uint4 m = *reinterpret_cast<const uint4 *>(row + c0);
    w[0] = *reinterpret_cast<const uint32_t *>(row + c0 - 4);
    w[1] = m.x; w[2] = m.y; w[3] = m.z; w[4] = m.w;
    w[5] = *reinterpret_cast<const uint32_t *>(row + c0 + 16);
    constexpr uint32_t kTaps = 0x04060401u, kLast = 0x00000001u;
    uint32_t s[5];   // s[i] = bytes c0+4i-2 .. c0+4i+1
#pragma unroll
    for (int i = 0; i < 5; i++) s[i] = __funnelshift_r(w[i], w[i + 1], 16);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t even = __dp4a(s[q], kTaps, __dp4a(s[q + 1], kLast, 0u));        // output 2q: bytes c0+4q-2 ..
        const uint32_t odd = __dp4a(w[q + 1], kTaps, __dp4a(w[q + 2], kLast, 0u));     // output 2q+1: bytes c0+4q ..
        h[q] = __byte_perm(even, odd, 0x5410);                                         // (even, odd) as 16-bit halves
    }
}

template <int kPyrTileY>
__global__ void __launch_bounds__((kPyrTileX / 8) * (kPyrTileY / kPyrRY)) pyr_down_kernel(PyrArgs a)
{
    constexpr int kPyrThreads = (kPyrTileX / 8) * (kPyrTileY / kPyrRY);
    constexpr int kPyrInRows = 2 * kPyrTileY + 3;
    extern __shared__ __align__(16) uint8_t tile[];   // [kPyrInRows][kPyrInPitch]
    const int ox_t = blockIdx.x * kPyrTileX, oy_t = blockIdx.y * kPyrTileY;
    const uint8_t *splane = a.src + (size_t)blockIdx.z * a.sstride;
    uint8_t *dplane = a.dst + (size_t)blockIdx.z * a.dstride;
    const int col_t = 2 * ox_t - 16;                  // image column of tile byte 0 (multiple of 16)
    const int row_t = 2 * oy_t - 2;                   // image row of tile row 0
    const int rows_needed = min(kPyrInRows, 2 * (a.Ho - oy_t) + 3);
    // chunks that hold a needed byte: tile bytes 14 .. 2*valid_out + 16
    const int last_chunk = min(kPyrChunks - 1, (2 * min(kPyrTileX, a.Wo - ox_t) + 16) / 16);

    // ---- stage the input footprint -----------------------------------------------------------
    // warp w copies tile rows w, w + 8, ...; lane c copies chunk c.  Chunks that overlap the frame travel whole
    // (the row pitch is padded to 16 bytes, so a chunk that straddles column W reads padding, not foreign
    // memory); the at most two reflect-101 columns on either side of the frame are patched in afterwards.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool generic = !a.vec_in || a.W < 32;
    if (!generic) {
        const int col = col_t + 16 * lane;
        if (lane <= last_chunk && col >= 0 && col < a.W) {
            uint32_t dst = smem_u32(tile) + (uint32_t)(warp * kPyrInPitch + 16 * lane);
            if (row_t >= 0 && row_t + rows_needed <= a.H) {        // no row is reflected: walk two pointers
                const uint8_t *src = splane + (size_t)(row_t + warp) * a.sp + col;
                const size_t step = (size_t)(kPyrThreads / 32) * a.sp;
#pragma unroll 4
                for (int r = warp; r < rows_needed; r += kPyrThreads / 32) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                    dst += (kPyrThreads / 32) * kPyrInPitch;
                    src += step;
                }
            } else {
                for (int r = warp; r < rows_needed; r += kPyrThreads / 32) {
                    const uint8_t *src = splane + (size_t)reflect101(row_t + r, a.H) * a.sp + col;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                    dst += (kPyrThreads / 32) * kPyrInPitch;
                }
            }
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (col_t < 0) {                                 // left frame border: columns -2, -1 mirror columns 2, 1
            for (int r = threadIdx.x; r < rows_needed; r += kPyrThreads) {
                uint8_t *t = tile + r * kPyrInPitch + 16;
                t[-2] = t[2];
                t[-1] = t[1];
            }
        }
        if (col_t + kPyrInPitch > a.W) {                 // right frame border: columns W, W+1 mirror W-2, W-3
            const int b = a.W - col_t;                   // tile byte of image column W
            for (int r = threadIdx.x; r < rows_needed; r += kPyrThreads) {
                uint8_t *t = tile + r * kPyrInPitch + b;
                if (b < kPyrInPitch) t[0] = t[-2];
                if (b + 1 < kPyrInPitch) t[1] = t[-3];
            }
        }
    } else {
        for (int i = threadIdx.x; i < rows_needed * kPyrChunks; i += kPyrThreads) {
            const int r = i / kPyrChunks, c = i - r * kPyrChunks;
            if (c > last_chunk) continue;
            const int col = col_t + 16 * c;
            const uint8_t *srow = splane + (size_t)reflect101(row_t + r, a.H) * a.sp;
            uint32_t v[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                v[q] = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) v[q] |= (uint32_t)srow[reflect101(col + 4 * q + k, a.W)] << (8 * k);
            }
            *reinterpret_cast<uint4 *>(tile + r * kPyrInPitch + 16 * c) = make_uint4(v[0], v[1], v[2], v[3]);
        }
    }
    __syncthreads();

    // ---- filter -------------------------------------------------------------------------------
    const int tx = threadIdx.x % (kPyrTileX / 8), ty = threadIdx.x / (kPyrTileX / 8);
    const int ox0 = ox_t + tx * 8, oy0 = oy_t + ty * kPyrRY;
    if (ox0 >= a.Wo || oy0 >= a.Ho) return;
    const int c0 = 16 + tx * 16;                      // tile byte of the first output's centre
    const uint8_t *trow = tile + (size_t)(2 * ty * kPyrRY) * kPyrInPitch;   // tile row of image row 2*oy0 - 2
    uint8_t *out = dplane + (size_t)oy0 * a.dp + ox0;
    // the thread's 8 x kPyrRY outputs all exist and the rows can be stored as 8-byte words: no per-row checks
    const bool whole = ox0 + 8 <= a.Wo && oy0 + kPyrRY <= a.Ho && a.vec_out;

    uint32_t hw[5][4];
#pragma unroll
    for (int i = 0; i < 3; i++) hrow(trow + i * kPyrInPitch, c0, hw[i + 2]);
#pragma unroll
    for (int t = 0; t < kPyrRY; t++) {
        if (!whole && oy0 + t >= a.Ho) break;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) hw[i][j] = hw[i + 2][j];
        hrow(trow + (2 * t + 3) * kPyrInPitch, c0, hw[3]);
        hrow(trow + (2 * t + 4) * kPyrInPitch, c0, hw[4]);
        uint32_t s[4];   // packed pairs of (5x5 sum + 128): the result pixel is the HIGH byte of each 16-bit half
#pragma unroll
        for (int j = 0; j < 4; j++)
            s[j] = (hw[0][j] + hw[4][j] + 0x00800080u) + 4u * (hw[1][j] + hw[3][j]) + 6u * hw[2][j];
        const uint32_t lo = __byte_perm(s[0], s[1], 0x7531), hi = __byte_perm(s[2], s[3], 0x7531);
        if (whole) {
            *reinterpret_cast<uint2 *>(out + (size_t)t * a.dp) = make_uint2(lo, hi);
        } else {
            const int nvalid = min(8, a.Wo - ox0);
            for (int k = 0; k < nvalid; k++) out[(size_t)t * a.dp + k] = (uint8_t)(((k < 4 ? lo : hi) >> (8 * (k & 3))) & 0xFF);
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
    a.vec_in = ((reinterpret_cast<uintptr_t>(src) | sp | sstride) % 16 == 0) ? 1 : 0;
    a.vec_out = ((reinterpret_cast<uintptr_t>(dst) | dp | dstride) % 8 == 0) ? 1 : 0;
    if (a.Ho >= 512) {
        constexpr int TY = 64, smem = (2 * TY + 3) * kPyrInPitch;
        dim3 grid((a.Wo + kPyrTileX - 1) / kPyrTileX, (a.Ho + TY - 1) / TY, n);
        pyr_down_kernel<TY><<<grid, (kPyrTileX / 8) * (TY / kPyrRY), smem, stream>>>(a);
    } else {
        constexpr int TY = 32, smem = (2 * TY + 3) * kPyrInPitch;
        dim3 grid((a.Wo + kPyrTileX - 1) / kPyrTileX, (a.Ho + TY - 1) / TY, n);
        pyr_down_kernel<TY><<<grid, (kPyrTileX / 8) * (TY / kPyrRY), smem, stream>>>(a);
    }
    note_launch();
    return check_launch("pyr_down_kernel");
}

}  // namespace gme
