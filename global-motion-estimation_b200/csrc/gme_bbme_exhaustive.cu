// gme_bbme_exhaustive.cu -- exhaustive block matching (K1), the integer-pipe-bound kernel.
//
// Replaces bbme.exhaustive_search (bbme.py:105-179).  One CTA owns NB horizontally adjacent
// macroblocks of one block row.  The search window of the CURRENT frame
// (2*sw + 2*bs - 1 rows) is staged in shared memory by one TMA box load; candidates that
// fall outside the frame are skipped by predicate exactly like bbme.py:157-162 (TMA's
// zero fill is never used as data).  A thread owns one COLUMN offset of one macroblock and
// walks down the window: each window row is loaded once (WPR+1 words, funnel-shifted to
// the thread's byte alignment) and scored against all BS anchor rows held in registers,
// feeding BS rotating accumulators -- one candidate (row offset) completes per window row.
// Inner loop: 1 LDS per ~13 VABSDIFF4.ACC (SAD) or VABSDIFF4+IDP.4A pairs (SSD).
//
// Tie-breaking: the reference scans column offset outer, row offset inner and keeps the first
// strict minimum (bbme.py:146-174).  A thread sees its row offsets in ascending order (strict
// '<'), threads are merged by a 64-bit key (cost, column index, row index) with atomicMin.
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "gme_common.cuh"

namespace gme {

struct ExhaustiveArgs {
    const uint8_t *prev;
    size_t prev_stride;
    const uint8_t *cur;
    size_t cur_stride;
    int H, W;
    size_t pitch;
    int R, C;
    int sw;
    int32_t *field;
    int nb;             // macroblocks per CTA
    int tpb;            // threads per macroblock (<= ncand)
    int win_w, win_h;   // staged window (bytes per row, rows)
    int use_tma;
    int strips_x;       // second-generation kernel: strips per block row, and strips in all (planes x block rows x strips_x)
    int strips;
};

template <int BS, int PNORM, int NT>
__global__ void __launch_bounds__(NT) bbme_exhaustive_kernel(const __grid_constant__ CUtensorMap cur_map,
                                                             ExhaustiveArgs a)
{
    constexpr int WPR = (BS + 3) / 4;
    constexpr uint32_t LAST_MASK = (BS % 4 == 0) ? 0xFFFFFFFFu : ((1u << (8 * (BS % 4))) - 1u);
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;

    const int plane = blockIdx.z, bi = blockIdx.y, bj0 = blockIdx.x * a.nb;
    const int br = bi * BS, bc0 = bj0 * BS;
    const int ncand = 2 * a.sw + BS;                 // offsets per axis: [-sw, sw+bs-1]  (bbme.py:146-149)
    // image coordinates of window (0, 0); the first column is rounded down to a 16-byte boundary because
    // TMA traps (cudaErrorIllegalInstruction) on a box whose innermost coordinate is not 16-byte aligned
    const int wr0 = br - a.sw, wc0 = (bc0 - a.sw) & ~15;
    const int xoff = (bc0 - a.sw) - wc0;
    const uint8_t *prev_plane = a.prev + (size_t)plane * a.prev_stride;
    const uint8_t *cur_plane = a.cur + (size_t)plane * a.cur_stride;

    uint32_t *win = reinterpret_cast<uint32_t *>(smem);
    const int win_pw = a.win_w / 4;
    const size_t win_bytes = ((size_t)a.win_w * a.win_h + 32 + 127) / 128 * 128;
    uint32_t *anchors = reinterpret_cast<uint32_t *>(smem + win_bytes);                    // [nb][BS][WPR]
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(anchors + a.nb * BS * WPR + (a.nb * BS * WPR & 1));

    // ---- stage the window (TMA) and the anchor blocks ------------------------------------
    if (a.use_tma) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(&bar, (uint32_t)(a.win_w * a.win_h));
            tma_load_3d(smem, &cur_map, &bar, wc0, wr0, plane);
        }
    } else {
        for (int i = threadIdx.x; i < win_pw * a.win_h; i += NT) {
            const int rr = wr0 + i / win_pw, cc = wc0 + (i % win_pw) * 4;
            uint32_t v = 0;
            if (rr >= 0 && rr < a.H) {
                const uint8_t *p = cur_plane + (size_t)rr * a.pitch;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (cc + k >= 0 && cc + k < a.W) v |= (uint32_t)p[cc + k] << (8 * k);
            }
            win[i] = v;
        }
    }
    for (int i = threadIdx.x; i < a.nb * BS * WPR; i += NT) {
        const int b = i / (BS * WPR), r = (i / WPR) % BS, w = i % WPR;
        uint32_t v = 0;
        if (bj0 + b < a.C) {
            const uint8_t *p = prev_plane + (size_t)(br + r) * a.pitch + (bc0 + b * BS) + 4 * w;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (4 * w + k < BS) v |= (uint32_t)p[k] << (8 * k);
        }
        anchors[i] = v;
    }
    for (int i = threadIdx.x; i < a.nb; i += NT) keys[i] = ~0ull;
    __syncthreads();
    if (a.use_tma) mbar_wait(&bar, 0);

    // ---- per-thread search ------------------------------------------------------------------
    const int b = threadIdx.x / a.tpb, t_in = threadIdx.x % a.tpb;
    if (b < a.nb && bj0 + b < a.C) {
        uint32_t anc[BS][WPR];
#pragma unroll
        for (int r = 0; r < BS; r++)
#pragma unroll
            for (int w = 0; w < WPR; w++) anc[r][w] = anchors[(b * BS + r) * WPR + w];

        const int bc = bc0 + b * BS;
        // row offsets j (wr = j - sw) whose candidate is inside the frame: 0 <= br + wr <= H - BS
        const int jlo = max(0, a.sw - br), jhi = min(ncand - 1, a.H - BS - br + a.sw);
        unsigned long long best_key = ~0ull;
        for (int ci = t_in; ci < ncand; ci += a.tpb) {
            const int left = bc + ci - a.sw;
            if (left < 0 || left > a.W - BS) continue;          // bbme.py:157-162, column part
            const int x0 = xoff + b * BS + ci;                   // byte column inside the window
            const int sh = (x0 & 3) * 8;
            const uint32_t *wp = win + (x0 >> 2);
            uint32_t acc[BS];
#pragma unroll
            for (int s = 0; s < BS; s++) acc[s] = 0;
            uint32_t best = 0xFFFFFFFFu;
            int bestj = 0;
            for (int y0 = 0; y0 < a.win_h; y0 += BS) {
#pragma unroll
                for (int yy = 0; yy < BS; yy++) {
                    const int y = y0 + yy;
                    if (y < a.win_h) {
                        const uint32_t *row = wp + y * win_pw;
                        uint32_t raw[WPR + 1], w[WPR];
#pragma unroll
                        for (int i = 0; i <= WPR; i++) raw[i] = row[i];
#pragma unroll
                        for (int i = 0; i < WPR; i++) w[i] = __funnelshift_r(raw[i], raw[i + 1], sh);
                        w[WPR - 1] &= LAST_MASK;
#pragma unroll
                        for (int k = 0; k < BS; k++) {           // window row y is row k of candidate j = y - k
                            const int s = (yy - k + BS) % BS;
#pragma unroll
                            for (int i = 0; i < WPR; i++) acc[s] = cost4_acc<PNORM>(w[i], anc[k][i], acc[s]);
                        }
                        const int sdone = (yy + 1) % BS;         // candidate j = y - BS + 1 is complete
                        const int j = y - BS + 1;
                        if (j >= jlo && j <= jhi && acc[sdone] < best) { best = acc[sdone]; bestj = j; }
                        acc[sdone] = 0;
                    }
                }
            }
            if (best != 0xFFFFFFFFu) {
                const unsigned long long key =
                    ((unsigned long long)best << 32) | ((unsigned long long)ci << 16) | (unsigned long long)bestj;
                best_key = key < best_key ? key : best_key;
            }
        }
        if (best_key != ~0ull) atomicMin(&keys[b], best_key);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.nb; i += NT) {
        if (bj0 + i < a.C) {
            const unsigned long long key = keys[i];
            const int ci = (int)((key >> 16) & 0xFFFF), j = (int)(key & 0xFFFF);
            int32_t *f = a.field + (((size_t)plane * a.R + bi) * a.C + bj0 + i) * 2;
            f[0] = ci - a.sw;      // column offset -> channel 0 (bbme.py:176)
            f[1] = j - a.sw;       // row offset    -> channel 1 (bbme.py:177)
        }
    }
}

// ---------------------------------------------------------------------------------------
// Second-generation tiled kernel (block sizes 8, 12, 16).  What changed against the kernel above, and why
// (ncu, profiles/r01e_*): the 16 x 16 x 4 fully unrolled update did not fit the instruction cache
// (no_instruction was the top stall), 116-130 registers allowed one CTA per SM, every window row cost
// BS-1 wasted fill/drain rows, and the funnel shifts and min-updates competed with VABSDIFF4 for the ALU pipe.
//   * TWO threads share a column offset: thread h scores anchor rows h*BS/2 .. and walks window rows
//     BS/2*h lower, so both finish candidate j in the same iteration and one SHFL adds the halves.
//     Half the accumulators and anchor registers (2-3 CTAs per SM), a quarter of the unrolled code,
//     and only BS/2-1 fill rows.
//   * the window is expanded once per CTA into four byte-shifted copies, so the inner loop loads
//     aligned words and never shifts.  The copies use a row pitch = 2 (mod 4) words and a copy stride
//     = 4 (mod 32) words: 16 consecutive byte columns x 2 halves hit 32 different banks.
//   * the running first-minimum is one VIMNMX on the key (cost << 8 | row index).
//   * rows that can only belong to out-of-frame candidates are not visited at all.
// ---------------------------------------------------------------------------------------
struct Exhaustive2Geom {
    int tpb;          // column offsets handled per macroblock per pass (thread pairs)
    int rawpw;        // raw window pitch, words
    int cpitch;       // shifted-copy row pitch, words (= 2 mod 4)
    int cstride;      // words between the four copies (= 4 mod 32)
    unsigned cpitch_rcp;   // ceil(2^32 / cpitch): i / cpitch == umulhi(i, cpitch_rcp) for i < 2^32 / cpitch
};

template <int BS, int PNORM, int SPLIT, bool PERSIST>
__global__ void __launch_bounds__(384, 2) bbme_exhaustive2_kernel(const __grid_constant__ CUtensorMap cur_map,
                                                                  ExhaustiveArgs a, Exhaustive2Geom g)
{
    static_assert(BS % 4 == 0 && BS >= 8 && (SPLIT == 1 || SPLIT == 2), "block sizes 8, 12, 16");
    constexpr int WPR = BS / 4, HB = BS / SPLIT;     // HB anchor rows per thread
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[2];

    const int NT = blockDim.x;
    const int ncand = 2 * a.sw + BS;
    const size_t raw_bytes = ((size_t)a.win_w * a.win_h + 127) / 128 * 128;
    uint32_t *copies = reinterpret_cast<uint32_t *>(smem + 2 * raw_bytes);                           // [4][cstride]
    const int nanc = a.nb * BS * WPR + (a.nb * BS * WPR & 1);                                        // words per anchor buffer
    uint32_t *anchors2 = copies + 4 * g.cstride;                                                     // [2][nb][BS][WPR]
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(anchors2 + 2 * nanc);
    constexpr int KA = (WPR + SPLIT - 1) / SPLIT;       // anchor words a thread fetches per strip: nb*BS*WPR <= KA * blockDim.x

    // Persistent CTA: it walks the strips t = blockIdx.x, blockIdx.x + gridDim.x, ... (strip = nb macroblocks of one block
    // row of one plane) and keeps TWO raw windows: while strip t is searched, the window of the next strip is already
    // travelling by TMA into the other buffer (its own mbarrier, phase = use count mod 2).  One strip per CTA spent half
    // of a CTA's lifetime in launch, TMA latency and drain with only two CTAs per SM to hide it (profiles/r01k).
    auto strip_origin = [&](int t, int &plane, int &bi, int &bj0) {
        const int row = t / a.strips_x;
        bj0 = (t - row * a.strips_x) * a.nb;
        plane = row / a.R;
        bi = row - plane * a.R;
    };
    auto issue_window = [&](int t, int buf) {                    // one thread: TMA of strip t's window into raw[buf]
        int plane, bi, bj0;
        strip_origin(t, plane, bi, bj0);
        fence_proxy_async();                                     // the buffer was read through the generic proxy before
        mbar_arrive_expect_tx(&bar[buf], (uint32_t)(a.win_w * a.win_h));
        tma_load_3d(smem + buf * raw_bytes, &cur_map, &bar[buf], (bj0 * BS - a.sw) & ~15, bi * BS - a.sw, plane);
    };
    // the anchor blocks of strip t, KA words per thread: one aligned 32-bit load per word (block columns are multiples
    // of BS, BS % 4 == 0, planes and pitch are 4-byte aligned)
    auto fetch_anchors = [&](int t, uint32_t (&v)[KA]) {
        int plane, bi, bj0;
        strip_origin(t, plane, bi, bj0);
        const uint8_t *pp = a.prev + (size_t)plane * a.prev_stride + (size_t)(bi * BS) * a.pitch + bj0 * BS;
#pragma unroll
        for (int k = 0; k < KA; k++) {
            const int i = threadIdx.x + k * NT;
            v[k] = 0;
            if (i < a.nb * BS * WPR) {
                const int b = i / (BS * WPR), r = (i / WPR) % BS, w = i % WPR;
                if (bj0 + b < a.C) v[k] = __ldg(reinterpret_cast<const uint32_t *>(pp + (size_t)r * a.pitch + b * BS + 4 * w));
            }
        }
    };
    auto store_anchors = [&](uint32_t *dst, const uint32_t (&v)[KA]) {
#pragma unroll
        for (int k = 0; k < KA; k++) {
            const int i = threadIdx.x + k * NT;
            if (i < a.nb * BS * WPR) dst[i] = v[k];
        }
    };
    if (a.use_tma) {
        if (threadIdx.x == 0) {
            mbar_init(&bar[0], 1);
            mbar_init(&bar[1], 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0 && (int)blockIdx.x < a.strips) issue_window(blockIdx.x, 0);
    }
    if ((int)blockIdx.x < a.strips) {
        uint32_t v[KA];
        fetch_anchors(blockIdx.x, v);
        store_anchors(anchors2, v);
    }

    // (PERSIST = false: one strip per CTA, the loop body runs once and all the prefetch code folds away)
    for (int it = 0, t = blockIdx.x; PERSIST ? t < a.strips : it == 0; ++it, t += gridDim.x) {
    const int buf = PERSIST ? (it & 1) : 0;
    int plane, bi, bj0;
    strip_origin(t, plane, bi, bj0);
    const int br = bi * BS, bc0 = bj0 * BS;
    const int wr0 = br - a.sw, wc0 = (bc0 - a.sw) & ~15;         // TMA: 16-byte aligned first column
    const int xoff = (bc0 - a.sw) - wc0;
    const uint8_t *prev_plane = a.prev + (size_t)plane * a.prev_stride;
    const uint8_t *cur_plane = a.cur + (size_t)plane * a.cur_stride;
    uint32_t *raw = reinterpret_cast<uint32_t *>(smem + buf * raw_bytes);

    // ---- stage: raw window (TMA, already in flight; the next strip's is issued now), anchor blocks, result keys ----
    if (a.use_tma) {
        // raw[buf ^ 1] was last read by the copy expansion of the previous strip, which every thread has left
        if (PERSIST && threadIdx.x == 0 && t + (int)gridDim.x < a.strips) issue_window(t + gridDim.x, buf ^ 1);
    } else {
        for (int i = threadIdx.x; i < g.rawpw * a.win_h; i += NT) {
            const int rr = wr0 + i / g.rawpw, cc = wc0 + (i % g.rawpw) * 4;
            uint32_t v = 0;
            if (rr >= 0 && rr < a.H) {
                const uint8_t *p = cur_plane + (size_t)rr * a.pitch;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (cc + k >= 0 && cc + k < a.W) v |= (uint32_t)p[cc + k] << (8 * k);
            }
            raw[i] = v;
        }
    }
    // anchor blocks: this strip's are in anchors2[buf] already (fetched while the previous strip was searched); the next
    // strip's are requested now and stored after the search, so their global-memory latency never shows
    uint32_t *anchors = anchors2 + buf * nanc;
    const bool has_next = PERSIST && t + (int)gridDim.x < a.strips;
    uint32_t next_anc[KA];
    if (has_next) fetch_anchors(t + gridDim.x, next_anc);
    for (int i = threadIdx.x; i < a.nb; i += NT) keys[i] = ~0ull;
    __syncthreads();
    if (a.use_tma) mbar_wait(&bar[buf], PERSIST ? (uint32_t)((it >> 1) & 1) : 0u);

    // ---- four byte-shifted copies of the window ----------------------------------------------------------------
    // (flat index over rows x words; the row comes from a multiply-high with the host's reciprocal of cpitch, exact
    // for every index of a window: a 32-bit division per element made this loop a quarter of a CTA's lifetime)
    for (int i = threadIdx.x; i < a.win_h * g.cpitch; i += NT) {
        const int r = (int)__umulhi((unsigned)i, g.cpitch_rcp), w = i - r * g.cpitch;
        const uint32_t lo = w < g.rawpw ? raw[r * g.rawpw + w] : 0u;
        const uint32_t hi = w + 1 < g.rawpw ? raw[r * g.rawpw + w + 1] : 0u;
        copies[i] = lo;
        copies[g.cstride + i] = __funnelshift_r(lo, hi, 8);
        copies[2 * g.cstride + i] = __funnelshift_r(lo, hi, 16);
        copies[3 * g.cstride + i] = __funnelshift_r(lo, hi, 24);
    }
    __syncthreads();

    // ---- search ---------------------------------------------------------------------------------------------
    const int q = threadIdx.x / SPLIT, h = threadIdx.x % SPLIT;
    const int b = q / g.tpb, t_in = q - b * g.tpb;
    const bool block_ok = b < a.nb && bj0 + b < a.C;
    const int bb = block_ok ? b : 0;
    uint32_t anc[HB][WPR];
#pragma unroll
    for (int k = 0; k < HB; k++)
#pragma unroll
        for (int w = 0; w < WPR; w++) anc[k][w] = anchors[(bb * BS + h * HB + k) * WPR + w];

    // row offsets j (wr = j - sw) whose candidate is inside the frame: 0 <= br + wr <= H - BS  (bbme.py:157-162)
    const int jlo = max(0, a.sw - br), jhi = min(ncand - 1, a.H - BS - br + a.sw);
    const int nj = jhi - jlo + 1, nrows = nj + HB - 1;
    const int bc = bc0 + bb * BS;
    unsigned long long best_key = ~0ull;
    for (int pass = 0; pass * g.tpb < ncand; pass++) {
        const int ci = t_in + pass * g.tpb;
        const int left = bc + ci - a.sw;
        const bool active = block_ok && ci < ncand && left >= 0 && left <= a.W - BS && nj > 0;   // column part of bbme.py:157-162
        // both threads of a pair share `active`, so the pair either walks the window or skips it together
        const unsigned pair_mask = SPLIT == 2 ? __ballot_sync(0xFFFFFFFFu, active) : 0u;
        if (!active) continue;
        const int x0 = xoff + bb * BS + ci;                                             // byte column inside the window
        const uint32_t *row = copies + (x0 & 3) * g.cstride + (jlo + h * HB) * g.cpitch + (x0 >> 2);
        uint32_t acc[HB];
#pragma unroll
        for (int s = 0; s < HB; s++) acc[s] = 0;
        uint32_t key = 0xFFFFFFFFu;
        // one window row: it is row h*HB + k of candidate t - k for every k it can still belong to
        auto step = [&](auto tt_c, auto first_c, int t0) {
            constexpr int tt = decltype(tt_c)::value;
            constexpr bool first = decltype(first_c)::value;
            uint32_t w[WPR];
#pragma unroll
            for (int i = 0; i < WPR; i++) w[i] = row[i];
            row += g.cpitch;
#pragma unroll
            for (int k = 0; k < HB; k++) {
                if (first && k > tt) continue;               // fill: candidate tt - k < 0 does not exist
                const int s = (tt - k + HB) % HB;
#pragma unroll
                for (int i = 0; i < WPR; i++) acc[s] = cost4_acc<PNORM>(w[i], anc[k][i], acc[s]);
            }
            if (!first || tt == HB - 1) {                    // candidate t - HB + 1 is complete in both halves
                constexpr int sdone = (tt + 1) % HB;
                uint32_t total = acc[sdone];
                if (SPLIT == 2) total += __shfl_xor_sync(pair_mask, total, 1);
                acc[sdone] = 0;
                key = min(key, total * 256u + (uint32_t)(t0 + tt - (HB - 1)));   // first minimum: smaller row index wins ties
            }
        };
#define GME_STEP(tt, first) step(std::integral_constant<int, tt>{}, std::integral_constant<bool, first>{}, t0)
#define GME_ROWS(first, guard)                                                                                      \
        do {                                                                                                        \
            if (!(guard) || t0 + 0 < nrows) GME_STEP(0, first);                                                     \
            if (HB > 1 && (!(guard) || t0 + 1 < nrows)) GME_STEP((1 % HB), first);                                  \
            if (HB > 2 && (!(guard) || t0 + 2 < nrows)) GME_STEP((2 % HB), first);                                  \
            if (HB > 3 && (!(guard) || t0 + 3 < nrows)) GME_STEP((3 % HB), first);                                  \
            if (HB > 4 && (!(guard) || t0 + 4 < nrows)) GME_STEP((4 % HB), first);                                  \
            if (HB > 5 && (!(guard) || t0 + 5 < nrows)) GME_STEP((5 % HB), first);                                  \
            if (HB > 6 && (!(guard) || t0 + 6 < nrows)) GME_STEP((6 % HB), first);                                  \
            if (HB > 7 && (!(guard) || t0 + 7 < nrows)) GME_STEP((7 % HB), first);                                  \
            if (HB > 8 && (!(guard) || t0 + 8 < nrows)) GME_STEP((8 % HB), first);                                  \
            if (HB > 9 && (!(guard) || t0 + 9 < nrows)) GME_STEP((9 % HB), first);                                  \
            if (HB > 10 && (!(guard) || t0 + 10 < nrows)) GME_STEP((10 % HB), first);                               \
            if (HB > 11 && (!(guard) || t0 + 11 < nrows)) GME_STEP((11 % HB), first);                               \
        } while (0)
        int t0 = 0;
        // nrows = nj + HB - 1 >= HB: the first block of HB rows always exists in full and is the (triangular) fill
        GME_ROWS(true, false);
        for (t0 = HB; t0 + HB <= nrows; t0 += HB) GME_ROWS(false, false);
        if (t0 < nrows) GME_ROWS(false, true);
#undef GME_ROWS
#undef GME_STEP
        if (h == 0 && key != 0xFFFFFFFFu) {
            const unsigned long long k64 = ((unsigned long long)(key >> 8) << 32) | ((unsigned long long)ci << 16) |
                                           (unsigned long long)(jlo + (int)(key & 255u));
            best_key = k64 < best_key ? k64 : best_key;
        }
    }
    // one 64-bit shared atomic per warp and macroblock instead of one per thread: when the whole warp works on one
    // macroblock (the usual case) the keys are first reduced with shuffles
    {
        const int first_bb = __shfl_sync(0xFFFFFFFFu, bb, 0);
        if (__all_sync(0xFFFFFFFFu, bb == first_bb)) {
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, best_key, o);
                best_key = other < best_key ? other : best_key;
            }
            if ((threadIdx.x & 31) == 0 && best_key != ~0ull) atomicMin(&keys[bb], best_key);
        } else if (best_key != ~0ull) {
            atomicMin(&keys[bb], best_key);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.nb; i += NT) {
        if (bj0 + i < a.C) {
            const unsigned long long key = keys[i];
            const int ci = (int)((key >> 16) & 0xFFFF), j = (int)(key & 0xFFFF);
            int32_t *f = a.field + (((size_t)plane * a.R + bi) * a.C + bj0 + i) * 2;
            f[0] = ci - a.sw;      // column offset -> channel 0 (bbme.py:176)
            f[1] = j - a.sw;       // row offset    -> channel 1 (bbme.py:177)
        }
    }
    if (has_next) store_anchors(anchors2 + (buf ^ 1) * nanc, next_anc);
    if (PERSIST) __syncthreads();  // keys and copies are rewritten by the next strip, which also reads the new anchors
    }   // strips
}

template <int BS, int PNORM>
static int launch_fast2(ExhaustiveArgs a, int n, cudaStream_t stream, bool *handled)
{
    // two threads per column offset pay off when a thread's share of the anchor would not fit the registers /
    // the unrolled update would not fit the instruction cache (block size 16); smaller blocks keep one thread
    constexpr int WPR = BS / 4, SPLIT = BS >= 16 ? 2 : 1;
    const int ncand = 2 * a.sw + BS;
    *handled = false;
    if (ncand > 256) return GME_OK;                              // key packs the row index in 8 bits
    Exhaustive2Geom g;
    g.tpb = min(ncand, 384 / SPLIT);
    // work of one thread for one column offset (packed updates); strips below ~3000 are "short": a CTA lives ~2 us and
    // half of that is launch, TMA latency and drain
    const long work = (long)(ncand + BS / SPLIT - 1) * (BS / SPLIT) * WPR * (PNORM == GME_PNORM_MSE ? 2 : 1);
    // (tried for short strips: two column offsets per thread so that a CTA holds more macroblocks -- 0.50 vs 0.53 of
    // the pipe ceiling on config 1; and evenly filled strips, nb = 9 instead of 10 there -- 0.49)
    const int nb = max(1, min(384 / SPLIT / g.tpb, a.C));
    const int threads = (nb * g.tpb * SPLIT + 31) / 32 * 32;
    const int win_h = 2 * a.sw + 2 * BS - 1;
    int win_w = 2 * a.sw + 2 * BS - 1 + (nb - 1) * BS + 4 + 15;  // +15: the first column is rounded down to 16 bytes
    win_w = (win_w + 15) / 16 * 16;
    g.rawpw = win_w / 4;
    g.cpitch = g.rawpw + 2;                                      // rawpw = 0 (mod 4)  ->  2 (mod 4)
    g.cstride = (win_h * g.cpitch + 31) / 32 * 32 + (SPLIT == 2 ? 4 : 8);   // 4 (mod 32); one thread per column: 8
    g.cpitch_rcp = (unsigned)((0x100000000ull + g.cpitch - 1) / g.cpitch);
    const size_t raw_bytes = ((size_t)win_w * win_h + 127) / 128 * 128;
    const size_t smem = 2 * raw_bytes + (size_t)4 * g.cstride * 4 + (size_t)2 * (nb * BS * WPR + 1) * 4 + (size_t)nb * 8 + 16;
    if (smem > 110 * 1024) return GME_OK;                        // window too large: previous kernel / generic path
    a.nb = nb; a.tpb = g.tpb; a.win_w = win_w; a.win_h = win_h;
    CUtensorMap map;
    a.use_tma = (win_w <= 256 && win_h <= 256 &&
                 make_plane_tensor_map(&map, a.cur, n, a.H, a.W, a.pitch, a.cur_stride, win_w, win_h)) ? 1 : 0;
    if (!a.use_tma) memset(&map, 0, sizeof(map));

    a.strips_x = (a.C + nb - 1) / nb;
    if ((long long)a.strips_x * a.R * n > 0x7FFFFFFFLL) return GME_OK;
    a.strips = a.strips_x * a.R * n;
    // Long strips (config 5: 87 window rows x 64 packed updates per thread): two persistent CTAs per SM, each walking its
    // share of the strips with the next window and anchors prefetched.  Short strips (config 1: 47 rows x 36 updates,
    // ~2 us): one strip per CTA -- measured 8 % faster there; the hardware's CTA scheduler balances the uneven border
    // strips, which a static walk does not.
    if (work >= 3000 && a.strips > 2 * kNumSMs) {
        auto kern = bbme_exhaustive2_kernel<BS, PNORM, SPLIT, true>;
        ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem);
        kern<<<2 * kNumSMs, threads, smem, stream>>>(map, a, g);
    } else {
        auto kern = bbme_exhaustive2_kernel<BS, PNORM, SPLIT, false>;
        ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem);
        kern<<<a.strips, threads, smem, stream>>>(map, a, g);
    }
    note_launch();
    *handled = true;
    return check_launch("bbme_exhaustive2_kernel");
}

// ---------------------------------------------------------------------------------------
// Generic path: any block size / window, a warp per macroblock, lanes over candidates.
// ---------------------------------------------------------------------------------------
template <int PNORM>
__global__ void __launch_bounds__(256) bbme_exhaustive_generic_kernel(ExhaustiveArgs a, int bs)
{
    const int plane = blockIdx.z;
    const long nblocks = (long)a.R * a.C;
    const long b = (long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (b >= nblocks) return;
    const int lane = threadIdx.x & 31;
    const int bi = (int)(b / a.C), bj = (int)(b % a.C), br = bi * bs, bc = bj * bs;
    const uint8_t *anchor = a.prev + (size_t)plane * a.prev_stride + (size_t)br * a.pitch + bc;
    const uint8_t *cur_plane = a.cur + (size_t)plane * a.cur_stride;
    const int ncand = 2 * a.sw + bs;
    unsigned long long best_key = ~0ull;
    for (long idx = lane; idx < (long)ncand * ncand; idx += 32) {      // idx = ci * ncand + j: the scan order
        const int ci = (int)(idx / ncand), j = (int)(idx % ncand);
        const int top = br + j - a.sw, left = bc + ci - a.sw;
        if (top < 0 || left < 0 || top + bs > a.H || left + bs > a.W) continue;
        const uint8_t *cand = cur_plane + (size_t)top * a.pitch + left;
        // exact sum over the row-major pixels [start, start + len) of the block
        auto range_sum = [&](int start, int len) -> uint32_t {
            uint32_t sum = 0;
            int r = start / bs, c = start - r * bs;
            for (int p = 0; p < len; p++) {
                const int d = (int)anchor[(size_t)r * a.pitch + c] - (int)cand[(size_t)r * a.pitch + c];
                sum += (PNORM == GME_PNORM_MAE) ? (uint32_t)abs(d) : (uint32_t)(d * d);
                if (++c == bs) { c = 0; ++r; }
            }
            return sum;
        };
        // MSE beyond block size 16: the reference's float32 pairwise sum rounds (gme_common.cuh)
        const uint32_t acc = (PNORM == GME_PNORM_MSE && bs > 16) ? pairwise_sum_f32_tree(bs * bs, range_sum)
                                                                  : range_sum(0, bs * bs);
        const unsigned long long key = ((unsigned long long)acc << 32) | (unsigned long long)idx;
        best_key = key < best_key ? key : best_key;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, best_key, o);
        best_key = other < best_key ? other : best_key;
    }
    if (lane == 0) {
        const long idx = (long)(best_key & 0xFFFFFFFFull);
        int32_t *f = a.field + ((size_t)plane * nblocks + b) * 2;
        f[0] = (int)(idx / ncand) - a.sw;
        f[1] = (int)(idx % ncand) - a.sw;
    }
}

// ---------------------------------------------------------------------------------------
// Host launcher
// ---------------------------------------------------------------------------------------
template <int BS, int PNORM>
static int launch_fast(ExhaustiveArgs a, int n, cudaStream_t stream, bool *handled)
{
    constexpr int NT = 384;
    constexpr int WPR = (BS + 3) / 4;
    const int ncand = 2 * a.sw + BS;
    *handled = false;
    if (ncand > 0xFFFF) return GME_OK;
    int tpb = min(ncand, NT);
    int nb = max(1, min(NT / tpb, a.C));
    const int win_h = 2 * a.sw + 2 * BS - 1;
    int win_w = 2 * a.sw + 2 * BS - 1 + (nb - 1) * BS + 4 + 15;  // +4: trailing word of the funnel shift; +15: alignment
    win_w = (win_w + 15) / 16 * 16;
    const size_t win_bytes = ((size_t)win_w * win_h + 32 + 127) / 128 * 128;
    const size_t smem = win_bytes + (size_t)(nb * BS * WPR + 1) * 4 + (size_t)nb * 8 + 16;
    if (smem > 200 * 1024) return GME_OK;                        // window too large for shared memory: generic path
    a.nb = nb; a.tpb = tpb; a.win_w = win_w; a.win_h = win_h;
    CUtensorMap map;
    a.use_tma = (win_w <= 256 && win_h <= 256 &&
                 make_plane_tensor_map(&map, a.cur, n, a.H, a.W, a.pitch, a.cur_stride, win_w, win_h)) ? 1 : 0;
    if (!a.use_tma) memset(&map, 0, sizeof(map));
    auto kern = bbme_exhaustive_kernel<BS, PNORM, NT>;
    ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem);
    dim3 grid((a.C + nb - 1) / nb, a.R, n);
    kern<<<grid, NT, smem, stream>>>(map, a);
    note_launch();
    *handled = true;
    return check_launch("bbme_exhaustive_kernel");
}

template <int PNORM>
static int launch_exhaustive_pn(ExhaustiveArgs a, int n, int bs, cudaStream_t stream)
{
    bool handled = false;
    int rc = GME_OK;
    switch (bs) {
    case 2: rc = launch_fast<2, PNORM>(a, n, stream, &handled); break;
    case 4: rc = launch_fast<4, PNORM>(a, n, stream, &handled); break;
    case 8: rc = launch_fast2<8, PNORM>(a, n, stream, &handled);
            if (!handled && rc == GME_OK) rc = launch_fast<8, PNORM>(a, n, stream, &handled);
            break;
    case 12: rc = launch_fast2<12, PNORM>(a, n, stream, &handled);
             if (!handled && rc == GME_OK) rc = launch_fast<12, PNORM>(a, n, stream, &handled);
             break;
    case 16: rc = launch_fast2<16, PNORM>(a, n, stream, &handled);
             if (!handled && rc == GME_OK) rc = launch_fast<16, PNORM>(a, n, stream, &handled);
             break;
    default: break;
    }
    if (handled || rc != GME_OK) return rc;
    const long nblocks = (long)a.R * a.C;
    const int warps = 8;
    dim3 grid((unsigned)((nblocks + warps - 1) / warps), 1, n);
    bbme_exhaustive_generic_kernel<PNORM><<<grid, warps * 32, 0, stream>>>(a, bs);
    note_launch();
    return check_launch("bbme_exhaustive_generic_kernel");
}

// ---------------------------------------------------------------------------------------
// Integer-pipe probe: the roofline denominator of the exhaustive search.  Every thread runs `iters` rounds of
// 4 x 8 independent packed cost updates (VABSDIFF4.ACC for SAD, VABSDIFF4 + IDP.4A for SSD) on registers -- no memory
// traffic and NO other instruction in the loop body (round 1's probe refreshed an operand with SHF + LOP3 every eight
// updates and so measured 8/10 of the pipe).  The second operand of every update is the accumulator of the neighbouring
// chain, so no two updates see the same inputs: with loop-invariant operands ptxas merges the absolute differences of
// the unrolled rounds and folds the dot products into a multiplication.
// Elapsed time gives the sustained pixel-pair rate of the instruction mix itself; the ceiling of the pipe is
// 148 SMs x 64 lanes x 4 pixels x f_SM (74.4 T pixel-pairs/s at 1965 MHz), bench.py reports both.
// ---------------------------------------------------------------------------------------
template <int PNORM>
__device__ __forceinline__ uint32_t probe_update(uint32_t a, uint32_t b, uint32_t acc)
{
    uint32_t r;
    if constexpr (PNORM == GME_PNORM_MAE) {
        asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(acc));
    } else {
        uint32_t d;
        asm volatile("vabsdiff4.u32.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(0u));
        asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(d), "r"(d), "r"(acc));
    }
    return r;
}

template <int PNORM>
__global__ void __launch_bounds__(256) sad_probe_kernel(uint32_t seed, int iters, uint32_t *out)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a[8], acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = (tid + 1u) * 2654435761u + k * 0x9E3779B9u + seed; acc[k] = seed ^ (tid * 0x85EBCA6Bu) ^ k; }
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = probe_update<PNORM>(a[(k + r) & 7], acc[(k + 1) & 7], acc[k]);
    }
    uint32_t t = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) t ^= acc[k];
    if (t == 0x12345678u) out[tid & 1023] = t;          // keeps the chains alive; practically never taken
}

int launch_sad_probe(int pnorm, int ctas, int iters, uint32_t *out, cudaStream_t stream)
{
    if (pnorm == GME_PNORM_MAE)
        sad_probe_kernel<GME_PNORM_MAE><<<ctas, 256, 0, stream>>>(12345u, iters, out);
    else
        sad_probe_kernel<GME_PNORM_MSE><<<ctas, 256, 0, stream>>>(12345u, iters, out);
    note_launch();
    return check_launch("sad_probe_kernel");
}

int launch_bbme_exhaustive(const uint8_t *prev, size_t prev_stride, const uint8_t *cur, size_t cur_stride, int n,
                           int H, int W, size_t pitch, int bs, int sw, int pnorm, int32_t *field, cudaStream_t stream)
{
    ExhaustiveArgs a{};
    a.prev = prev; a.prev_stride = prev_stride;
    a.cur = cur; a.cur_stride = cur_stride;
    a.H = H; a.W = W; a.pitch = pitch;
    a.R = H / bs; a.C = W / bs;
    a.sw = sw;
    a.field = field;
    if (a.R == 0 || a.C == 0 || n == 0) return GME_OK;
    return pnorm == GME_PNORM_MAE ? launch_exhaustive_pn<GME_PNORM_MAE>(a, n, bs, stream)
                                  : launch_exhaustive_pn<GME_PNORM_MSE>(a, n, bs, stream);
}

}  // namespace gme
