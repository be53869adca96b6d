// gme_compensate.cu -- block-wise motion compensation fused with the PSNR error sum (K5+K6).
//
// Replaces motion.compensate_frame (motion.py:289-321) and the squared-error sum of
// utils.PSNR (utils.py:100-116).  The reference's "affine warp" is a per-block integer
// translation gather: comp[a, b] = frame[a - d1, b - d0] when the source pixel is inside the
// frame, else frame[a, b]; rows/columns past the last whole block are copied.  HBM-bound:
// algorithmic traffic = read frame + read cur + write comp = 3*H*W bytes per pair.
//
// A thread owns 16 consecutive pixels of one row: when they share one motion vector and the
// source run is inside the frame it gathers them with (at most 5) aligned 32-bit loads and
// funnel shifts, reads the 16 pixels of `cur` with one 128-bit load, accumulates the squared
// error with VABSDIFF4 + IDP.4A and writes one 128-bit store.  Anything else (frame borders,
// block sizes that are not multiples of 16, unaligned planes) takes a per-pixel path.
#include "gme_common.cuh"

namespace gme {

struct CompArgs {
    const uint8_t *frame; size_t fp, fstride;
    const void *field; int field_is_i16; int R, C, bs;
    const uint8_t *cur; size_t cp, cstride;
    uint8_t *comp; size_t op, ostride;
    int H, W;
    int vec_ok;
    unsigned long long *sse; size_t sse_stride;   // squared-error sum of plane k at sse[k * sse_stride]
    uint8_t *dprev, *dcomp; size_t dp, dstride;   // optional |cur - frame| and |cur - comp| planes (results.py:78-83)
};

__device__ __forceinline__ void load_vector(const CompArgs &a, int plane, int i, int j, int &d0, int &d1)
{
    const size_t idx = (((size_t)plane * a.R + i) * a.C + j) * 2;
    if (a.field_is_i16) {
        const short2 v = *reinterpret_cast<const short2 *>(static_cast<const int16_t *>(a.field) + idx);
        d0 = v.x; d1 = v.y;
    } else {
        const int2 v = *reinterpret_cast<const int2 *>(static_cast<const int32_t *>(a.field) + idx);
        d0 = v.x; d1 = v.y;
    }
}

// per-warp partial sums -> one 64-bit atomic per CTA
__device__ __forceinline__ void block_sum_u32_to_u64(unsigned int v, unsigned long long *dst)
{
    __shared__ unsigned int warp_sums[8];
    v = __reduce_add_sync(0xFFFFFFFFu, v);
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if ((tid & 31) == 0) warp_sums[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) t += warp_sums[w];
        if (t) atomicAdd(dst, t);
    }
}

// A thread owns 16 consecutive pixels of kCompRows consecutive rows (all inside one macroblock row when the block
// size is a multiple of kCompRows, so they share one motion vector): the loads of the rows are independent, which
// gives the memory system four requests per thread to overlap.  Block (32, 8): a CTA covers 512 x 32 pixels.
constexpr int kCompRows = 4;

template <bool HAS_CUR>
__global__ void __launch_bounds__(256) compensate_kernel(CompArgs a)
{
    const int plane = blockIdx.z;
    const int row0 = (blockIdx.y * 8 + threadIdx.y) * kCompRows;
    const int b0 = (blockIdx.x * 32 + threadIdx.x) * 16;
    unsigned int err = 0;
    if (row0 < a.H && b0 < a.W) {
        const uint8_t *fplane = a.frame + (size_t)plane * a.fstride;
        const uint8_t *cplane = HAS_CUR ? a.cur + (size_t)plane * a.cstride : nullptr;
        uint8_t *oplane = a.comp + (size_t)plane * a.ostride;
        const int npx = min(16, a.W - b0);
        const int j0 = b0 / a.bs;
        const bool fast = a.vec_ok && npx == 16 && j0 == (b0 + 15) / a.bs && a.bs % kCompRows == 0;
        if (fast) {
            // branch-light gather of 16-pixel runs: aligned 32-bit loads of the source run (words that lie outside
            // the row are not touched), funnel shift to the destination alignment, then -- only for runs that
            // straddle the frame edge -- a per-byte blend with the unmoved pixels (motion.py:311-318)
            const int i = row0 / a.bs;
            int lo = 16, hi = 0, d1 = 0, s0 = 0;                             // bytes [lo, hi) of a run are moved
            if (i < a.R && j0 < a.C) {
                int d0;
                load_vector(a, plane, i, j0, d0, d1);
                d0 = clampi(d0, -(1 << 24), 1 << 24);                        // any |d| >= frame size: source outside
                d1 = clampi(d1, -(1 << 24), 1 << 24);
                s0 = b0 - d0;                                                // motion.py:313
                lo = max(0, -s0);
                hi = min(16, a.W - s0);
            }
            const int sa = s0 & ~3;                                          // floor to a word boundary (also for s0 < 0)
            const int sh = (s0 - sa) * 8;
            const bool inside = sa >= 0 && sa + 20 <= (int)a.fp;
            uint32_t px[kCompRows][4];
            uint4 cur4[kCompRows];
            bool moved[kCompRows];
#pragma unroll
            for (int r = 0; r < kCompRows; r++) {
                const int a_row = row0 + r;
                const int na = a_row - d1;                                   // motion.py:312
                moved[r] = lo < hi && a_row < a.H && na >= 0 && na < a.H;
                if (a_row < a.H) {
                    if (moved[r]) {
                        const uint8_t *srow = fplane + (size_t)na * a.fp;
                        uint32_t w[5];
                        if (inside) {
#pragma unroll
                            for (int k = 0; k < 5; k++) w[k] = __ldg(reinterpret_cast<const uint32_t *>(srow + sa) + k);
                        } else {
#pragma unroll
                            for (int k = 0; k < 5; k++) {
                                const int c = sa + 4 * k;
                                w[k] = (c >= 0 && c + 4 <= (int)a.fp) ? __ldg(reinterpret_cast<const uint32_t *>(srow + c)) : 0u;
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 4; k++) px[r][k] = __funnelshift_r(w[k], w[k + 1], sh);
                    }
                    if (!moved[r] || lo > 0 || hi < 16) {                    // some pixels stay where they are
                        const uint4 v = *reinterpret_cast<const uint4 *>(fplane + (size_t)a_row * a.fp + b0);
                        const uint32_t orig[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const uint32_t m = moved[r] ? (byte_mask(clampi(hi - 4 * k, 0, 4)) & ~byte_mask(clampi(lo - 4 * k, 0, 4))) : 0u;
                            px[r][k] = (moved[r] ? (px[r][k] & m) : 0u) | (orig[k] & ~m);
                        }
                    }
                    if (HAS_CUR) cur4[r] = *reinterpret_cast<const uint4 *>(cplane + (size_t)a_row * a.cp + b0);
                }
            }
#pragma unroll
            for (int r = 0; r < kCompRows; r++) {
                const int a_row = row0 + r;
                if (a_row < a.H) {
                    *reinterpret_cast<uint4 *>(oplane + (size_t)a_row * a.op + b0) = make_uint4(px[r][0], px[r][1], px[r][2], px[r][3]);
                    if (HAS_CUR) {
                        err = ssd4_acc(px[r][0], cur4[r].x, err);
                        err = ssd4_acc(px[r][1], cur4[r].y, err);
                        err = ssd4_acc(px[r][2], cur4[r].z, err);
                        err = ssd4_acc(px[r][3], cur4[r].w, err);
                    }
                }
            }
        } else {
            for (int a_row = row0; a_row < min(row0 + kCompRows, a.H); a_row++) {
                const uint8_t *frow = fplane + (size_t)a_row * a.fp;
                uint8_t *orow = oplane + (size_t)a_row * a.op;
                const int i = a_row / a.bs;
                for (int k = 0; k < npx; k++) {
                    const int b = b0 + k, j = b / a.bs;
                    uint8_t v = frow[b];
                    if (i < a.R && j < a.C) {
                        int d0, d1;
                        load_vector(a, plane, i, j, d0, d1);
                        const long na = (long)a_row - d1, nb = (long)b - d0;
                        if (na >= 0 && na < a.H && nb >= 0 && nb < a.W) v = fplane[(size_t)na * a.fp + nb];
                    }
                    orow[b] = v;
                    if (HAS_CUR) {
                        const int d = (int)v - (int)cplane[(size_t)a_row * a.cp + b];
                        err += (unsigned int)(d * d);
                    }
                }
            }
        }
    }
    if (HAS_CUR) block_sum_u32_to_u64(err, a.sse + (size_t)plane * a.sse_stride);
}

// The case the pipeline runs (results.py:52-59): 16 x 16 blocks (BBME_BLOCK_SIZE), 16-byte aligned planes.  Same
// arithmetic as compensate_kernel, stripped to what this case needs: the 16-pixel run of a thread is exactly one
// block column, the four rows of a thread share the block row, all index math is shifts, and the row pointers are
// walked instead of recomputed.  Runs whose source leaves the frame go through the per-byte blend.
template <bool HAS_CUR, bool DIFFS>
__global__ void __launch_bounds__(256) compensate16_kernel(CompArgs a)
{
    static_assert(!DIFFS || HAS_CUR, "the difference images need the current frame");
    const int plane = blockIdx.z;
    const int row0 = (blockIdx.y * 8 + threadIdx.y) * kCompRows;
    const int b0 = (blockIdx.x * 32 + threadIdx.x) * 16;
    unsigned int err = 0;
    if (row0 < a.H && b0 + 16 <= a.W) {
        const uint8_t *fplane = a.frame + (size_t)plane * a.fstride;
        const int i = row0 >> 4, j0 = b0 >> 4;
        int lo = 16, hi = 0, d1 = 0, s0 = 0;                                 // bytes [lo, hi) of a run are moved
        if (i < a.R && j0 < a.C) {
            int d0;
            load_vector(a, plane, i, j0, d0, d1);
            d0 = clampi(d0, -(1 << 24), 1 << 24);
            d1 = clampi(d1, -(1 << 24), 1 << 24);
            s0 = b0 - d0;                                                    // motion.py:313
            lo = max(0, -s0);
            hi = min(16, a.W - s0);
        }
        const bool whole = lo == 0 && hi == 16;                              // the common case: the run moves as a whole
        const int sa = s0 & ~3;
        const int sh = (s0 - sa) * 8;
        const bool inside = sa >= 0 && sa + 20 <= (int)a.fp;
        const uint8_t *orig = fplane + (size_t)row0 * a.fp + b0;
        const uint8_t *crow = HAS_CUR ? a.cur + (size_t)plane * a.cstride + (size_t)row0 * a.cp + b0 : nullptr;
        uint8_t *orow = a.comp + (size_t)plane * a.ostride + (size_t)row0 * a.op + b0;
        const int nrows = min(kCompRows, a.H - row0);
        const int na0 = row0 - d1;                                           // motion.py:312
        uint32_t px[kCompRows][4];
        uint4 cur4[kCompRows];
        if (whole && inside && na0 >= 0 && na0 + kCompRows <= a.H && nrows == kCompRows) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(fplane + (size_t)na0 * a.fp + sa);
            const size_t fpw = a.fp >> 2;
#pragma unroll
            for (int r = 0; r < kCompRows; r++) {
                uint32_t w[5];
#pragma unroll
                for (int k = 0; k < 5; k++) w[k] = __ldg(src + r * fpw + k);
#pragma unroll
                for (int k = 0; k < 4; k++) px[r][k] = __funnelshift_r(w[k], w[k + 1], sh);
                if (HAS_CUR) cur4[r] = *reinterpret_cast<const uint4 *>(crow + (size_t)r * a.cp);
            }
        } else {
#pragma unroll
            for (int r = 0; r < kCompRows; r++) {
                if (r < nrows) {
                    const int na = na0 + r;
                    const bool moved = lo < hi && na >= 0 && na < a.H;
                    if (moved) {
                        const uint8_t *srow = fplane + (size_t)na * a.fp;
                        uint32_t w[5];
#pragma unroll
                        for (int k = 0; k < 5; k++) {
                            const int c = sa + 4 * k;
                            w[k] = (c >= 0 && c + 4 <= (int)a.fp) ? __ldg(reinterpret_cast<const uint32_t *>(srow + c)) : 0u;
                        }
#pragma unroll
                        for (int k = 0; k < 4; k++) px[r][k] = __funnelshift_r(w[k], w[k + 1], sh);
                    }
                    if (!moved || !whole) {
                        const uint4 v = *reinterpret_cast<const uint4 *>(orig + (size_t)r * a.fp);
                        const uint32_t o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const uint32_t m = moved ? (byte_mask(clampi(hi - 4 * k, 0, 4)) & ~byte_mask(clampi(lo - 4 * k, 0, 4))) : 0u;
                            px[r][k] = (moved ? (px[r][k] & m) : 0u) | (o[k] & ~m);
                        }
                    }
                    if (HAS_CUR) cur4[r] = *reinterpret_cast<const uint4 *>(crow + (size_t)r * a.cp);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kCompRows; r++) {
            if (r < nrows) {
                *reinterpret_cast<uint4 *>(orow + (size_t)r * a.op) = make_uint4(px[r][0], px[r][1], px[r][2], px[r][3]);
                if (HAS_CUR) {
                    err = ssd4_acc(px[r][0], cur4[r].x, err);
                    err = ssd4_acc(px[r][1], cur4[r].y, err);
                    err = ssd4_acc(px[r][2], cur4[r].z, err);
                    err = ssd4_acc(px[r][3], cur4[r].w, err);
                }
                if (DIFFS) {                                 // results.py:78-83: |current - compensated|, |current - previous|
                    const size_t off = (size_t)plane * a.dstride + (size_t)(row0 + r) * a.dp + b0;
                    *reinterpret_cast<uint4 *>(a.dcomp + off) = make_uint4(absdiff4(px[r][0], cur4[r].x), absdiff4(px[r][1], cur4[r].y),
                                                                           absdiff4(px[r][2], cur4[r].z), absdiff4(px[r][3], cur4[r].w));
                    const uint4 o = *reinterpret_cast<const uint4 *>(orig + (size_t)r * a.fp);
                    *reinterpret_cast<uint4 *>(a.dprev + off) = make_uint4(absdiff4(o.x, cur4[r].x), absdiff4(o.y, cur4[r].y),
                                                                           absdiff4(o.z, cur4[r].z), absdiff4(o.w, cur4[r].w));
                }
            }
        }
    }
    if (HAS_CUR) block_sum_u32_to_u64(err, a.sse + (size_t)plane * a.sse_stride);
}

// |x - y| per pixel (the difference images of results.py:78-83 when the fused kernel does not apply)
__global__ void __launch_bounds__(256) absdiff_kernel(const uint8_t *x, size_t xp, size_t xstride, const uint8_t *y,
                                                      size_t yp, size_t ystride, uint8_t *out, size_t op, size_t ostride,
                                                      int H, int W)
{
    const int plane = blockIdx.z;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (r >= H || c0 >= W) return;
    const uint8_t *xr = x + (size_t)plane * xstride + (size_t)r * xp + c0;
    const uint8_t *yr = y + (size_t)plane * ystride + (size_t)r * yp + c0;
    uint8_t *o = out + (size_t)plane * ostride + (size_t)r * op + c0;
    for (int k = 0; k < min(16, W - c0); k++) o[k] = (uint8_t)abs((int)xr[k] - (int)yr[k]);
}

__global__ void __launch_bounds__(256) sse_kernel(const uint8_t *x, size_t xp, size_t xstride, const uint8_t *y,
                                                  size_t yp, size_t ystride, int H, int W, int vec_ok,
                                                  unsigned long long *sse)
{
    const int plane = blockIdx.z;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    unsigned int err = 0;
    if (r < H && c0 < W) {
        const uint8_t *xr = x + (size_t)plane * xstride + (size_t)r * xp + c0;
        const uint8_t *yr = y + (size_t)plane * ystride + (size_t)r * yp + c0;
        if (vec_ok && c0 + 16 <= W) {
            const uint4 u = *reinterpret_cast<const uint4 *>(xr), v = *reinterpret_cast<const uint4 *>(yr);
            err = ssd4_acc(u.x, v.x, err);
            err = ssd4_acc(u.y, v.y, err);
            err = ssd4_acc(u.z, v.z, err);
            err = ssd4_acc(u.w, v.w, err);
        } else {
            for (int k = 0; k < min(16, W - c0); k++) {
                const int d = (int)xr[k] - (int)yr[k];
                err += (unsigned int)(d * d);
            }
        }
    }
    block_sum_u32_to_u64(err, sse + plane);
}

static inline bool aligned16(const void *p, size_t pitch, size_t stride)
{
    return ((reinterpret_cast<uintptr_t>(p) | pitch | stride) % 16) == 0;
}

int launch_compensate(const uint8_t *frame, size_t fp, size_t fstride, const void *field, int field_is_i16, int R,
                      int C, const uint8_t *cur, size_t cp, size_t cstride, uint8_t *comp, size_t op, size_t ostride,
                      int n, int H, int W, uint64_t *sse, cudaStream_t stream, uint8_t *dprev = nullptr,
                      uint8_t *dcomp = nullptr, size_t dp = 0, size_t dstride = 0, size_t sse_stride = 1)
{
    CompArgs a;
    a.sse_stride = sse_stride;
    a.dprev = dprev; a.dcomp = dcomp; a.dp = dp; a.dstride = dstride;
    a.frame = frame; a.fp = fp; a.fstride = fstride;
    a.field = field; a.field_is_i16 = field_is_i16;
    a.bs = (R > 0) ? H / R : 0;                                   // motion.py:303: rows only
    if (a.bs <= 0) { a.bs = 1; a.R = 0; a.C = 0; } else { a.R = R; a.C = C; }   // bs == 0: the reference's loops are empty
    a.cur = cur; a.cp = cp; a.cstride = cstride;
    a.comp = comp; a.op = op; a.ostride = ostride;
    a.H = H; a.W = W;
    a.vec_ok = (aligned16(frame, fp, fstride) && aligned16(comp, op, ostride) && (!cur || aligned16(cur, cp, cstride))) ? 1 : 0;
    a.sse = reinterpret_cast<unsigned long long *>(sse);
    dim3 block(32, 8);
    dim3 grid(((W + 15) / 16 + block.x - 1) / block.x, (H + block.y * kCompRows - 1) / (block.y * kCompRows), n);
    const bool lean = a.vec_ok && a.bs == 16 && W % 16 == 0;      // the pipeline's case
    const bool diffs = dprev && dcomp && cur && sse;
    const bool fused_diffs = diffs && lean && aligned16(dprev, dp, dstride) && aligned16(dcomp, dp, dstride);
    if (cur && sse) {
        if (sse_stride == 1) cudaMemsetAsync(sse, 0, sizeof(uint64_t) * n, stream);
        else cudaMemset2DAsync(sse, sse_stride * sizeof(uint64_t), 0, sizeof(uint64_t), n, stream);
        if (fused_diffs) compensate16_kernel<true, true><<<grid, block, 0, stream>>>(a);
        else if (lean) compensate16_kernel<true, false><<<grid, block, 0, stream>>>(a);
        else compensate_kernel<true><<<grid, block, 0, stream>>>(a);
    } else {
        if (lean) compensate16_kernel<false, false><<<grid, block, 0, stream>>>(a);
        else compensate_kernel<false><<<grid, block, 0, stream>>>(a);
    }
    note_launch();
    if (diffs && !fused_diffs) {
        dim3 g2(((W + 15) / 16 + 31) / 32, (H + 7) / 8, n);
        absdiff_kernel<<<g2, block, 0, stream>>>(cur, cp, cstride, frame, fp, fstride, dprev, dp, dstride, H, W);
        absdiff_kernel<<<g2, block, 0, stream>>>(cur, cp, cstride, comp, op, ostride, dcomp, dp, dstride, H, W);
        note_launch();
        note_launch();
    }
    return check_launch("compensate_kernel");
}

int launch_sse(const uint8_t *x, size_t xp, size_t xstride, const uint8_t *y, size_t yp, size_t ystride, int n, int H,
               int W, uint64_t *sse, cudaStream_t stream)
{
    const int vec_ok = (aligned16(x, xp, xstride) && aligned16(y, yp, ystride)) ? 1 : 0;
    dim3 block(32, 8);
    dim3 grid(((W + 15) / 16 + block.x - 1) / block.x, (H + block.y - 1) / block.y, n);
    cudaMemsetAsync(sse, 0, sizeof(uint64_t) * n, stream);
    sse_kernel<<<grid, block, 0, stream>>>(x, xp, xstride, y, yp, ystride, H, W, vec_ok,
                                           reinterpret_cast<unsigned long long *>(sse));
    note_launch();
    return check_launch("sse_kernel");
}

}  // namespace gme
