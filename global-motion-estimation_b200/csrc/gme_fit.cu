// gme_fit.cu -- the affine global-motion fit (K4): first estimate, model field, outlier
// mask and masked least squares.  KB-scale data, latency-bound: one CTA per frame pair.
//
// Replaces motion.compute_first_parameters (motion.py:176-188), parameter_projection
// (motion.py:191-207), affine_model / get_motion_field_affine (motion.py:91-105,139-157)
// and the fit of best_affine_parameters[_robust] (motion.py:33-88, 210-286).
//
// Everything that must be bit-exact is integer: the L1 differences, the order-statistic
// threshold (exact radix select instead of the reference's sort) and the twelve masked sums
// of the normal equations (int64, so the sums are exact; the reference accumulates the same
// terms in float64 with one rounding per block).  Only the final 3x3 solve is floating
// point: LU with partial pivoting + unit-lower/upper solves of the identity, the sequence
// LAPACK's dgesv uses for np.linalg.inv, then the 3x3 matrix-vector product.  Stated
// tolerance on the parameters: 1e-9 absolute and relative (tests/).
#include "gme_common.cuh"

namespace gme {

constexpr int kFitThreads = 256;

__device__ __forceinline__ long long block_sum_ll(long long v, long long *scratch /* >= 8 */)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    __syncthreads();   // scratch reuse
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    long long t = 0;
#pragma unroll
    for (int w = 0; w < kFitThreads / 32; w++) t += scratch[w];
    return t;
}

// motion.compute_first_parameters: mean of each channel in float64 (exact integer sum / count),
// cast to float32 (motion.py:186-188); carried as float64 (exact promotion).
__global__ void __launch_bounds__(kFitThreads) first_params_kernel(const int32_t *dense, long N, double *params)
{
    __shared__ long long scratch[8];
    const int32_t *f = dense + (size_t)blockIdx.x * N * 2;
    long long s0 = 0, s1 = 0;
    for (long i = threadIdx.x; i < N; i += kFitThreads) {
        const int2 v = *reinterpret_cast<const int2 *>(f + 2 * i);
        s0 += v.x;
        s1 += v.y;
    }
    s0 = block_sum_ll(s0, scratch);
    s1 = block_sum_ll(s1, scratch);
    if (threadIdx.x == 0) {
        double *p = params + (size_t)blockIdx.x * 6;
        p[0] = (double)(float)((double)s0 / (double)N);
        p[1] = 0.0; p[2] = 0.0;
        p[3] = (double)(float)((double)s1 / (double)N);
        p[4] = 0.0; p[5] = 0.0;
    }
}

// affine_model (motion.py:102-104): [[1,i,j,0,0,0],[0,0,0,1,i,j]] @ p in float64, then Python round() =
// round-half-to-even.  np.matmul gives this product to BLAS dgemv, so the summation order is the
// BLAS kernel's: d0 = (a0 + j*a2) + i*a1 without contraction, d1 = b0 + fma(i, b1, j*b2) -- pinned against
// live NumPy (OpenBLAS 0.3.30) on 50 000 random cases, see oracle/gme_oracle.c.  Only matters within 1 ulp
// of a .5 tie.
__device__ __forceinline__ void model_vector(const double *p, int i, int j, int &m0, int &m1)
{
    const double di = (double)i, dj = (double)j;
    const double d0 = __dadd_rn(__dadd_rn(p[0], __dmul_rn(dj, p[2])), __dmul_rn(di, p[1]));
    const double d1 = __dadd_rn(p[3], __fma_rn(di, p[4], __dmul_rn(dj, p[5])));
    m0 = (int)(int16_t)__double2int_rn(d0);   // stored as int16 (motion.py:150)
    m1 = (int)(int16_t)__double2int_rn(d1);
}

__global__ void affine_field_kernel(const double *params, int R, int C, int16_t *field)
{
    const long N = (long)R * C;
    const double *p = params + (size_t)blockIdx.y * 6;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N) return;
    int m0, m1;
    model_vector(p, (int)(idx / C), (int)(idx % C), m0, m1);
    *reinterpret_cast<short2 *>(field + ((size_t)blockIdx.y * N + idx) * 2) = make_short2((short)m0, (short)m1);
}

struct FitLevel {
    const int32_t *gt;       // BBME field of this level, int32[n][R][C][2]
    int R, C;
    double w;                // 1 / (level_h * level_w)   (motion.py:250)
    uint8_t *outlier;        // optional outputs of this level
    int32_t *threshold;
    int16_t *model_field;
};

struct FitArgs {
    FitLevel lv[2];          // the pipeline fits L1 and L2 in ONE launch (the second starts from the first's result)
    int nlevels;
    double pct;              // MOTION_VECTOR_ERROR_THRESHOLD_PERCENTAGE
    int robust, project;
    double *params;
    size_t params_stride;    // doubles between the parameter rows of consecutive pairs (>= 6)
    int32_t *status;
    int status_or;           // OR into status instead of overwriting
    const long long *first_sums;   // pipeline: per-pair channel sums of the dense field (written by the dense
    long first_count;              // BBME kernel); the first estimate (motion.py:186-188) is formed here
    size_t cache_bytes;            // dynamic shared memory available for the distances of one level
    int16_t *final_field;          // optional: model field of the FINAL parameters on the last level's grid
};

// 3x3 inverse the way np.linalg.inv gets it (dgesv on the identity): LU with partial pivoting,
// multipliers by reciprocal (dgetf2), forward / backward substitution per identity column.
__device__ bool inverse3(const double (&A)[3][3], double (&inv)[3][3])
{
    double lu[3][3];
    int piv[3] = {0, 1, 2};
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) lu[r][c] = A[r][c];
    for (int k = 0; k < 3; k++) {
        int p = k;
        double big = fabs(lu[k][k]);
        for (int r = k + 1; r < 3; r++)
            if (fabs(lu[r][k]) > big) { big = fabs(lu[r][k]); p = r; }
        if (big == 0.0 || big != big) return false;              // exactly singular -> LinAlgError in the reference
        if (p != k) {
            for (int c = 0; c < 3; c++) { const double t = lu[k][c]; lu[k][c] = lu[p][c]; lu[p][c] = t; }
            const int t = piv[k]; piv[k] = piv[p]; piv[p] = t;
        }
        const double rcp = __ddiv_rn(1.0, lu[k][k]);
        for (int r = k + 1; r < 3; r++) {
            lu[r][k] = __dmul_rn(lu[r][k], rcp);
            for (int c = k + 1; c < 3; c++) lu[r][c] = __dsub_rn(lu[r][c], __dmul_rn(lu[r][k], lu[k][c]));
        }
    }
    for (int col = 0; col < 3; col++) {
        double y[3];
        for (int r = 0; r < 3; r++) {                            // L y = P e_col
            double v = (piv[r] == col) ? 1.0 : 0.0;
            for (int c = 0; c < r; c++) v = __dsub_rn(v, __dmul_rn(lu[r][c], y[c]));
            y[r] = v;
        }
        for (int r = 2; r >= 0; r--) {                           // U x = y
            double v = y[r];
            for (int c = r + 1; c < 3; c++) v = __dsub_rn(v, __dmul_rn(lu[r][c], inv[c][col]));
            inv[r][col] = __ddiv_rn(v, lu[r][r]);
        }
    }
    return true;
}

// Exact warp-wide sum of an int64 per lane with three REDUX.SUM instead of five 64-bit shuffle steps: the two's-
// complement bit pattern is cut into limbs of 22, 22 and 20 bits, each limb summed over the 32 lanes (< 2^27), and
// the limb sums recombined modulo 2^64 -- exact whenever the true sum fits an int64, which the callers guarantee.
__device__ __forceinline__ long long warp_sum_ll(long long v)
{
    const unsigned long long u = (unsigned long long)v;
    const unsigned long long s0 = __reduce_add_sync(0xFFFFFFFFu, (unsigned)(u & 0x3FFFFFu));
    const unsigned long long s1 = __reduce_add_sync(0xFFFFFFFFu, (unsigned)((u >> 22) & 0x3FFFFFu));
    const unsigned long long s2 = __reduce_add_sync(0xFFFFFFFFu, (unsigned)(u >> 44));
    return (long long)(s0 + (s1 << 22) + (s2 << 44));
}

// (row, column) of the field elements tid, tid + step, tid + 2 step, ... without a division per element
struct RowCol {
    int row, col, qs, rs, C;
    __device__ __forceinline__ RowCol(int first, int step, int C_) : row(first / C_), col(first % C_), qs(step / C_), rs(step % C_), C(C_) {}
    __device__ __forceinline__ void next()
    {
        row += qs;
        col += rs;
        if (col >= C) { col -= C; ++row; }
    }
};

// One CTA of kFitBig threads per frame pair.  The field is KB-scale, so the kernel is latency-bound: the
// distances are computed once (float64 model vector per block) and kept in shared memory for the radix
// select and the masked sums; every block-wide step is a warp-shuffle reduction or scan.
constexpr int kFitBig = 1024;
constexpr int kFitWarps = kFitBig / 32;
constexpr int kRadixBits = 11, kRadixBins = 1 << kRadixBits;

__global__ void __launch_bounds__(kFitBig) affine_fit_kernel(FitArgs a)
{
    extern __shared__ __align__(16) unsigned int diff_cache[];   // [N] of the current level when it fits
    __shared__ long long red[kFitWarps][12];
    __shared__ unsigned int hist[kRadixBins];
    __shared__ unsigned int warp_tot[kFitWarps];
    __shared__ double p[6];
    __shared__ unsigned int sel_prefix, sel_rank, sh_dmax;

    const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *params = a.params + (size_t)pair * a.params_stride;

    if (tid < 6) {
        if (a.first_sums)   // motion.compute_first_parameters (motion.py:186-188): float64 mean, stored as float32
            p[tid] = (tid == 0 || tid == 3)
                         ? (double)(float)((double)a.first_sums[2 * pair + (tid == 3)] / (double)a.first_count) : 0.0;
        else
            p[tid] = params[tid];
    }
    int st_all = 0;

    for (int level = 0; level < a.nlevels; level++) {
        const FitLevel L = a.lv[level];
        const int N = L.R * L.C;
        const int32_t *gt = L.gt + (size_t)pair * N * 2;
        const bool cached = a.robust && (size_t)N * sizeof(unsigned int) <= a.cache_bytes;
        __syncthreads();
        if (a.project && (tid == 0 || tid == 3)) p[tid] = p[tid] * 2.0;   // motion.parameter_projection (motion.py:204-207)
        if (tid == 0) sh_dmax = 0;
        __syncthreads();

        // ---- L1 distance between the BBME field and the model field (motion.py:232-239) --------
        auto diff_at = [&](int idx) -> unsigned int {
            int m0, m1;
            model_vector(p, idx / L.C, idx % L.C, m0, m1);
            const int2 g = *reinterpret_cast<const int2 *>(gt + 2 * idx);
            return (unsigned int)(abs(g.x - m0) + abs(g.y - m1));
        };
        auto diff_of = [&](int idx) -> unsigned int { return cached ? diff_cache[idx] : diff_at(idx); };

        unsigned int thr = 0xFFFFFFFFu;
        if (a.robust) {
            // threshold = sorted(diff)[N - int(pct*N)]  (element 0 when int(pct*N) == 0: Python's [-0])
            const int t = (int)(a.pct * (double)N);
            unsigned int rank = (unsigned int)(t == 0 ? 0 : N - t);   // 0-based rank of the threshold
            unsigned int prefix = 0;                                   // bits of the answer found so far
            unsigned int dmax = 0;
            // the field is read with four independent loads in flight per thread: the kernel is latency-bound
            RowCol rc(tid, kFitBig, L.C);
            for (int base = tid; base < N; base += 4 * kFitBig) {
                int2 g[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = base + u * kFitBig;
                    g[u] = i < N ? __ldg(reinterpret_cast<const int2 *>(gt + 2 * i)) : make_int2(0, 0);
                }
#pragma unroll
                for (int u = 0; u < 4; u++, rc.next()) {
                    const int i = base + u * kFitBig;
                    if (i < N) {
                        int m0, m1;
                        model_vector(p, rc.row, rc.col, m0, m1);
                        const unsigned int d = (unsigned int)(abs(g[u].x - m0) + abs(g[u].y - m1));
                        if (cached) diff_cache[i] = d;
                        dmax = max(dmax, d);
                        if (L.model_field)
                            *reinterpret_cast<short2 *>(L.model_field + ((size_t)pair * N + i) * 2) = make_short2((short)m0, (short)m1);
                    }
                }
            }
            dmax = __reduce_max_sync(0xFFFFFFFFu, dmax);
            if (lane == 0 && dmax) atomicMax(&sh_dmax, dmax);
            __syncthreads();
            dmax = sh_dmax;
            // exact radix select, 11 bits per pass, starting at the highest digit that is populated
            int shift = 0;
            while (shift + kRadixBits < 32 && (dmax >> (shift + kRadixBits)) != 0) shift += kRadixBits;
            for (; shift >= 0; shift -= kRadixBits) {
                for (int i = tid; i < kRadixBins; i += kFitBig) hist[i] = 0;
                __syncthreads();
                const unsigned int hi_mask = (shift + kRadixBits >= 32) ? 0u : (0xFFFFFFFFu << (shift + kRadixBits));
                for (int i = tid; i < N; i += kFitBig) {
                    const unsigned int d = diff_of(i);
                    if ((d & hi_mask) == prefix) atomicAdd(&hist[(d >> shift) & (kRadixBins - 1)], 1u);
                }
                __syncthreads();
                // each thread owns 2 consecutive bins; block-wide exclusive scan of the pair sums by warp shuffles
                const unsigned int b0 = hist[2 * tid], b1 = hist[2 * tid + 1];
                unsigned int incl = b0 + b1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned int up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= o) incl += up;
                }
                if (lane == 31) warp_tot[warp] = incl;
                __syncthreads();
                if (warp == 0) {
                    unsigned int v = warp_tot[lane], sc = v;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned int up = __shfl_up_sync(0xFFFFFFFFu, sc, o);
                        if (lane >= o) sc += up;
                    }
                    warp_tot[lane] = sc - v;                           // exclusive
                }
                __syncthreads();
                const unsigned int before = warp_tot[warp] + incl - (b0 + b1);
                if (rank >= before && rank < before + b0 + b1) {       // exactly one thread
                    const bool second = rank >= before + b0;
                    sel_prefix = prefix | ((unsigned int)(2 * tid + (second ? 1 : 0)) << shift);
                    sel_rank = rank - before - (second ? b0 : 0u);
                }
                __syncthreads();
                prefix = sel_prefix;
                rank = sel_rank;
                __syncthreads();
            }
            thr = prefix;
        } else if (L.model_field) {
            RowCol rc(tid, kFitBig, L.C);
            for (int i = tid; i < N; i += kFitBig, rc.next()) {
                int m0, m1;
                model_vector(p, rc.row, rc.col, m0, m1);
                *reinterpret_cast<short2 *>(L.model_field + ((size_t)pair * N + i) * 2) = make_short2((short)m0, (short)m1);
            }
        }

        // ---- masked normal equations (motion.py:246-282): twelve exact integer sums -----------
        long long S[12];
#pragma unroll
        for (int k = 0; k < 12; k++) S[k] = 0;
        RowCol rcs(tid, kFitBig, L.C);
        for (int base = tid; base < N; base += 4 * kFitBig) {
            int2 g[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + u * kFitBig;
                g[u] = i < N ? __ldg(reinterpret_cast<const int2 *>(gt + 2 * i)) : make_int2(0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; u++, rcs.next()) {
                const int i = base + u * kFitBig;
                if (i < N) {
                    const bool out = a.robust ? (diff_of(i) > thr) : false;   // strict '>' (motion.py:244)
                    if (L.outlier) L.outlier[(size_t)pair * N + i] = out ? 1 : 0;
                    if (!out) {
                        const long long x = 4 * rcs.row, y = 4 * rcs.col;      // motion.py:254-255: x = 4 i, y = 4 j
                        S[0] += 1;         S[1] += x;             S[2] += y;
                        S[3] += x * x;     S[4] += x * y;         S[5] += y * y;
                        S[6] += g[u].x;    S[7] += x * g[u].x;    S[8] += y * g[u].x;
                        S[9] += g[u].y;    S[10] += x * g[u].y;   S[11] += y * g[u].y;
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 12; k++) {
            S[k] = warp_sum_ll(S[k]);
            if (lane == 0) red[warp][k] = S[k];
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int k = 0; k < 12; k++) S[k] = warp_sum_ll(red[lane][k]);
        }

        if (tid == 0) {
            const double w = L.w;
            double M[3][3], inv[3][3];
            M[0][0] = (double)S[0] * w; M[0][1] = (double)S[1] * w; M[0][2] = (double)S[2] * w;
            M[1][0] = M[0][1];          M[1][1] = (double)S[3] * w; M[1][2] = (double)S[4] * w;
            M[2][0] = M[0][2];          M[2][1] = M[1][2];          M[2][2] = (double)S[5] * w;
            const double r0[3] = {(double)S[6] * w, (double)S[7] * w, (double)S[8] * w};
            const double r1[3] = {(double)S[9] * w, (double)S[10] * w, (double)S[11] * w};
            if (inverse3(M, inv)) {
                for (int r = 0; r < 3; r++) {
                    p[r] = __dadd_rn(__dadd_rn(__dmul_rn(inv[r][0], r0[0]), __dmul_rn(inv[r][1], r0[1])),
                                     __dmul_rn(inv[r][2], r0[2]));
                    p[3 + r] = __dadd_rn(__dadd_rn(__dmul_rn(inv[r][0], r1[0]), __dmul_rn(inv[r][1], r1[1])),
                                         __dmul_rn(inv[r][2], r1[2]));
                }
            } else {
                st_all |= 1;
                const double qnan = __longlong_as_double(0x7FF8000000000000LL);
                for (int r = 0; r < 6; r++) p[r] = qnan;
            }
            if (L.threshold) L.threshold[pair] = a.robust ? (int32_t)thr : 0;
        }
    }
    __syncthreads();
    if (tid < 6) params[tid] = p[tid];
    if (tid == 0 && a.status) a.status[pair] = a.status_or ? (a.status[pair] | st_all) : st_all;
    if (a.final_field) {                       // motion.get_motion_field_affine with the final parameters (results.py:52-54)
        const FitLevel L = a.lv[a.nlevels - 1];
        const int N = L.R * L.C;
        RowCol rc(tid, kFitBig, L.C);
        for (int i = tid; i < N; i += kFitBig, rc.next()) {
            int m0, m1;
            model_vector(p, rc.row, rc.col, m0, m1);
            *reinterpret_cast<short2 *>(a.final_field + ((size_t)pair * N + i) * 2) = make_short2((short)m0, (short)m1);
        }
    }
}

// bbme.hierarchical_wrapper's merge step (bbme.py:587-604): the coarser field is upsampled by 2 in both axes (nearest),
// its vectors truncated to int32 and doubled (bbme.rescale_motion_field, bbme.py:537-546), padded with one zero row or
// column where the finer field has one more, and averaged with the finer field: out = (2*trunc(coarse) + fine) / 2.
template <typename T>
__global__ void hier_merge_kernel(const T *coarse, int Rc, int Cc, const int32_t *fine, int R, int C, double *out)
{
    const long N = (long)R * C;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N) return;
    const int i = (int)(idx / C), j = (int)(idx % C);
    const T *cp = coarse + (size_t)blockIdx.y * Rc * Cc * 2;
    int u0 = 0, u1 = 0;
    if (i / 2 < Rc && j / 2 < Cc) {
        u0 = 2 * (int)cp[((size_t)(i / 2) * Cc + j / 2) * 2];          // (int): truncation toward zero, like ndarray assignment
        u1 = 2 * (int)cp[((size_t)(i / 2) * Cc + j / 2) * 2 + 1];
    }
    const int2 f = *reinterpret_cast<const int2 *>(fine + ((size_t)blockIdx.y * N + idx) * 2);
    double *o = out + ((size_t)blockIdx.y * N + idx) * 2;
    o[0] = (double)(u0 + f.x) / 2.0;
    o[1] = (double)(u1 + f.y) / 2.0;
}

int launch_hier_merge(const void *coarse, int coarse_is_f64, int Rc, int Cc, const int32_t *fine, int R, int C, int n,
                      double *out, cudaStream_t stream)
{
    dim3 grid((unsigned)(((long)R * C + 255) / 256), n);
    if (coarse_is_f64)
        hier_merge_kernel<double><<<grid, 256, 0, stream>>>(static_cast<const double *>(coarse), Rc, Cc, fine, R, C, out);
    else
        hier_merge_kernel<int32_t><<<grid, 256, 0, stream>>>(static_cast<const int32_t *>(coarse), Rc, Cc, fine, R, C, out);
    note_launch();
    return check_launch("hier_merge_kernel");
}

int launch_first_params(const int32_t *dense, int n, int R, int C, double *params, cudaStream_t stream)
{
    first_params_kernel<<<n, kFitThreads, 0, stream>>>(dense, (long)R * C, params);
    note_launch();
    return check_launch("first_params_kernel");
}

static int launch_fit(FitArgs &a, int n, cudaStream_t stream)
{
    size_t want = 0;
    for (int l = 0; l < a.nlevels; l++) want = max(want, (size_t)a.lv[l].R * a.lv[l].C * sizeof(unsigned int));
    a.cache_bytes = a.robust ? min(want, (size_t)160 * 1024) : 0;
    // per launch, not once per process: the attribute belongs to the current device's copy of the kernel
    ensure_dynamic_smem(reinterpret_cast<const void *>(affine_fit_kernel), 160 * 1024);
    affine_fit_kernel<<<n, kFitBig, a.cache_bytes, stream>>>(a);
    note_launch();
    return check_launch("affine_fit_kernel");
}

int launch_affine_fit(const int32_t *gt, int n, int R, int C, int level_h, int level_w, double pct, int robust,
                      int project, double *params, uint8_t *outlier, int32_t *threshold, int16_t *model_field,
                      int32_t *status, int status_or, const long long *first_sums, long first_count,
                      cudaStream_t stream)
{
    FitArgs a{};
    a.lv[0] = FitLevel{gt, R, C, 1.0 / (double)((long long)level_h * level_w), outlier, threshold, model_field};
    a.nlevels = 1;
    a.pct = pct; a.robust = robust; a.project = project;
    a.params = params; a.params_stride = 6; a.status = status; a.status_or = status_or;
    a.first_sums = first_sums; a.first_count = first_count;
    a.final_field = nullptr;
    return launch_fit(a, n, stream);
}

// The pipeline's sequential tail in one launch: first estimate from the dense channel sums, projection + robust fit
// on L1, projection + robust fit on L2 (motion.py:128-134), then the model field of the result (results.py:52-54).
int launch_pipeline_fits(const int32_t *f1, int R1, int C1, int h1, int w1, uint8_t *out1, const int32_t *f2, int R2,
                         int C2, int h2, int w2, uint8_t *out2, int n, double pct, double *params, int32_t *status,
                         const long long *first_sums, long first_count, int16_t *final_field, cudaStream_t stream,
                         size_t params_stride)
{
    FitArgs a{};
    a.lv[0] = FitLevel{f1, R1, C1, 1.0 / (double)((long long)h1 * w1), out1, nullptr, nullptr};
    a.lv[1] = FitLevel{f2, R2, C2, 1.0 / (double)((long long)h2 * w2), out2, nullptr, nullptr};
    a.nlevels = 2;
    a.pct = pct; a.robust = 1; a.project = 1;
    a.params = params; a.params_stride = params_stride; a.status = status; a.status_or = 1;
    a.first_sums = first_sums; a.first_count = first_count;
    a.final_field = final_field;
    return launch_fit(a, n, stream);
}

int launch_affine_field(const double *params, int n, int R, int C, int16_t *field, cudaStream_t stream)
{
    const long N = (long)R * C;
    dim3 grid((unsigned)((N + 255) / 256), n);
    affine_field_kernel<<<grid, 256, 0, stream>>>(params, R, C, field);
    note_launch();
    return check_launch("affine_field_kernel");
}

}  // namespace gme
