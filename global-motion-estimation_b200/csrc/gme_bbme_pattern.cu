// gme_bbme_pattern.cu -- three-step, 2D-log and diamond block matching (K2).
//
// Replaces bbme.threestep_search (bbme.py:182-341), bbme.twodlog_search (bbme.py:344-433)
// and bbme.diamond_search (bbme.py:436-534).  One CTA owns a tile of macroblocks; the part
// of the CURRENT frame the tile can plausibly reach (tile + margin) is staged in shared
// memory by one TMA box load (cooperative loads when the planes are not TMA-aligned).  A
// group of G lanes owns one macroblock: every lane keeps its share of the anchor block in
// registers, evaluates its share of each candidate with packed byte ops (VABSDIFF4 /
// IDP.4A) and the group sums with warp shuffles.  The walk of the search is data
// dependent and unbounded in the reference, so a candidate that leaves the staged window
// is evaluated straight from global memory -- same integers, only slower.
//
// Scan orders, strict '<' first-minimum tie-breaking, the diamond clamp to H-bs-1, the
// swapped SDSP offsets, the double-counted three-step offset and the unbounded 2D-log walk
// are reproduced exactly; the oracle is oracle/gme_oracle.c.
#include <cstring>

#include "gme_common.cuh"

namespace gme {

struct PatternArgs {
    const uint8_t *prev;
    size_t prev_stride;
    const uint8_t *cur;
    size_t cur_stride;
    int H, W;
    size_t pitch;
    int R, C;          // macroblocks per frame
    int sw, procedure;
    int32_t *field;    // [n][R][C][2]
    int tbx, tby;      // macroblocks per tile
    int margin;        // staged margin around the tile, pixels
    int win_w, win_h;  // staged window: win_w bytes per row (multiple of 16), win_h rows
    int use_tma;
    int edge_tiles;    // dense diamond kernel: the clamped block columns (first, last two) are tiles of their own
    uint32_t key_scale;   // 16, as a run-time value (dense diamond kernel: keeps the key arithmetic off the ALU pipe)
    unsigned long long *sums;   // optional, [n][2]: per-plane sums of the two field channels (pipeline: the dense
                                // first estimate, motion.py:186-188, is a mean -- no second pass over the field)
};

constexpr uint32_t kInfCost = 0xFFFFFFFFu;   // every real cost is < 2^32 - 1 (bs <= 255 checked on the host)

// ---------------------------------------------------------------------------------------
// Fast path: compile-time block size, G lanes per macroblock, window in shared memory.
// ---------------------------------------------------------------------------------------
template <int BS, int G, int PNORM>
struct BlockEval {
    static constexpr int WPR = (BS + 3) / 4;          // 32-bit words per block row
    static constexpr int UNITS = BS * WPR;            // words per block
    static constexpr int UPL = (UNITS + G - 1) / G;   // words per lane

    uint32_t anchor[UPL];
    const uint8_t *cur_plane;
    const uint32_t *win;      // shared window
    size_t pitch;
    int win_pw;               // window pitch in words
    int wr0, wc0, wr1, wc1;   // window covers rows [wr0, wr1) and columns [wc0, wc1)
    int lane_g;               // lane index inside the group
    uint32_t gmask;           // shuffle mask of the group

    __device__ __forceinline__ void load_anchor(const uint8_t *prev_plane, int br, int bc)
    {
#pragma unroll
        for (int t = 0; t < UPL; t++) {
            const int u = lane_g + t * G;
            uint32_t v = 0;
            if (u < UNITS) {
                const int ur = u / WPR, uw = u % WPR;
                const uint8_t *p = prev_plane + (size_t)(br + ur) * pitch + bc + 4 * uw;
                if constexpr (BS % 4 == 0) {
                    v = __ldg(reinterpret_cast<const uint32_t *>(p));   // planes and pitch are 4-byte aligned (checked on the host)
                } else {
                    const int nv = min(4, BS - 4 * uw);
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (k < nv) v |= (uint32_t)p[k] << (8 * k);
                }
            }
            anchor[t] = v;
        }
    }

    __device__ __forceinline__ bool inside(int r, int c) const
    {
        return r >= wr0 && r + BS <= wr1 && c >= wc0 && c + BS <= wc1;
    }

    // this lane's share of the cost of the candidate whose top-left pixel is (r, c)
    __device__ __forceinline__ uint32_t partial_smem(int r, int c) const
    {
        uint32_t acc = 0;
        const int x0 = c - wc0;
        const int sh = (x0 & 3) * 8;
        const uint32_t *base = win + (r - wr0) * win_pw + (x0 >> 2);
#pragma unroll
        for (int t = 0; t < UPL; t++) {
            const int u = lane_g + t * G;
            if (UNITS % G == 0 || u < UNITS) {
                const int ur = u / WPR, uw = u % WPR;
                const uint32_t *p = base + ur * win_pw + uw;
                uint32_t v = __funnelshift_r(p[0], p[1], sh);
                if (BS % 4 != 0 && uw == WPR - 1) v &= byte_mask(BS - 4 * (WPR - 1));
                acc = cost4_acc<PNORM>(v, anchor[t], acc);
            }
        }
        return acc;
    }

    __device__ __forceinline__ uint32_t partial_gmem(int r, int c) const
    {
        uint32_t acc = 0;
#pragma unroll
        for (int t = 0; t < UPL; t++) {
            const int u = lane_g + t * G;
            if (UNITS % G == 0 || u < UNITS) {
                const int ur = u / WPR, uw = u % WPR;
                const int nv = min(4, BS - 4 * uw);
                const size_t off = (size_t)(r + ur) * pitch + (size_t)(c + 4 * uw);
                const uint32_t *p = reinterpret_cast<const uint32_t *>(cur_plane + (off & ~(size_t)3));
                const int mis = (int)(off & 3);
                const uint32_t lo = __ldg(p);
                const uint32_t hi = (mis + nv > 4) ? __ldg(p + 1) : 0u;   // never touch a word with no valid byte
                uint32_t v = __funnelshift_r(lo, hi, mis * 8);
                v &= byte_mask(nv);
                acc = cost4_acc<PNORM>(v, anchor[t], acc);
            }
        }
        return acc;
    }

    __device__ __forceinline__ uint32_t reduce(uint32_t v) const
    {
        if constexpr (G == 32) {
            return __reduce_add_sync(0xFFFFFFFFu, v);        // one REDUX.SUM instead of five shuffle steps
        } else {
#pragma unroll
            for (int o = G / 2; o >= 1; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
            return v;
        }
    }

    // costs of N candidates; all loads are issued before the reductions (ILP)
    template <int N>
    __device__ __forceinline__ void eval(const int (&r)[N], const int (&c)[N], uint32_t (&cost)[N]) const
    {
        bool all_in = true;
#pragma unroll
        for (int k = 0; k < N; k++) all_in &= inside(r[k], c[k]);
        uint32_t part[N];
        if (all_in) {
#pragma unroll
            for (int k = 0; k < N; k++) part[k] = partial_smem(r[k], c[k]);
        } else {
#pragma unroll
            for (int k = 0; k < N; k++) part[k] = partial_gmem(r[k], c[k]);
        }
#pragma unroll
        for (int k = 0; k < N; k++) cost[k] = reduce(part[k]);
    }
};

// ---------------------------------------------------------------------------------------
// The three searches, written once over any evaluator E (fast or generic).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool cand_in_frame(int top, int left, int bs, int H, int W)
{
    // bbme.py:239-244, 405-411
    return !(top < 0 || left < 0 || top + bs - 1 > H - 1 || left + bs - 1 > W - 1);
}

// One LDSP step (bbme.py:494-513) evaluated candidate by candidate at CLAMPED positions -- the form that is always
// valid.  Moves (mr, mc) to the first strict minimum in the reference's order; returns true when that is the centre.
template <class E>
__device__ __forceinline__ bool ldsp_step_clamped(const E &e, int rmax, int cmax, int &mr, int &mc)
{
    constexpr int LR[9] = {0, 2, 1, 0, -1, -2, -1, 0, 1};      // LDSP offsets (row, col), bbme.py:463-472
    constexpr int LC[9] = {0, 0, 1, 2, 1, 0, -1, -2, -1};
    int r[9], c[9];
    uint32_t cost[9];
#pragma unroll
    for (int k = 0; k < 9; k++) {
        r[k] = clampi(mr + LR[k], 0, rmax);                    // bbme.py:503-504
        c[k] = clampi(mc + LC[k], 0, cmax);
    }
    e.template eval<9>(r, c, cost);
    uint32_t best = kInfCost;
    int best_r = mr, best_c = mc;
#pragma unroll
    for (int k = 0; k < 9; k++)
        if (cost[k] < best) { best = cost[k]; best_r = r[k]; best_c = c[k]; }
    const bool stop = (best_r == mr) && (best_c == mc);        // tuple equality, bbme.py:512
    mr = best_r;
    mc = best_c;
    return stop;
}

// The final SDSP (bbme.py:515-529) at clamped positions; the offset list is applied swapped (bbme.py:518-521).
template <class E>
__device__ __forceinline__ void sdsp_clamped(const E &e, int rmax, int cmax, int mr, int mc, int &out_r, int &out_c)
{
    constexpr int SR[5] = {0, 0, 1, 0, -1};
    constexpr int SC[5] = {0, 1, 0, -1, 0};
    int r[5], c[5];
    uint32_t cost[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        r[k] = clampi(mr + SR[k], 0, rmax);
        c[k] = clampi(mc + SC[k], 0, cmax);
    }
    e.template eval<5>(r, c, cost);
    uint32_t best = kInfCost;
    out_r = mr;
    out_c = mc;
#pragma unroll
    for (int k = 0; k < 5; k++)
        if (cost[k] < best) { best = cost[k]; out_r = r[k]; out_c = c[k]; }
}

// Candidate-at-a-time evaluation for centres near the frame border, where the clamp of bbme.py:503-504 acts but every
// (clamped) candidate still lies in the staged window: shared memory only, no inside test, no global-memory variant.
// The general evaluator (any position, global memory when needed) is kept OUT of line: it is two orders of magnitude
// rarer, and inlined into the walk it made every warp that held one border block run ~900 extra instructions.
template <class E>
struct WindowOnly {
    const E &e;
    template <int N>
    __device__ __forceinline__ void eval(const int (&r)[N], const int (&c)[N], uint32_t (&cost)[N]) const
    {
        uint32_t part[N];
#pragma unroll
        for (int k = 0; k < N; k++) part[k] = e.partial_smem(r[k], c[k]);
#pragma unroll
        for (int k = 0; k < N; k++) cost[k] = e.reduce(part[k]);
    }
};

// (by value, results packed in the return value: taking the address of the evaluator or of the walk's position would
// move them from registers to the stack in the hot loop as well)
template <class E>
__device__ __noinline__ int2 ldsp_step_cold(const E e, int rmax, int cmax, int mr, int mc)
{
    ldsp_step_clamped(e, rmax, cmax, mr, mc);      // the step stops exactly when the position does not move (bbme.py:512)
    return make_int2(mr, mc);
}

template <class E>
__device__ __noinline__ int2 sdsp_cold(const E e, int rmax, int cmax, int mr, int mc)
{
    int out_r, out_c;
    sdsp_clamped(e, rmax, cmax, mr, mc, out_r, out_c);
    return make_int2(out_r, out_c);
}

template <class E>
__device__ __forceinline__ void diamond_walk(const E &e, int bs, int H, int W, int br, int bc, int &out0, int &out1)
{
    const int rmax = H - bs - 1, cmax = W - bs - 1;   // bbme.py:503-504 (off by one, kept)
    int mr = br, mc = bc;
    while (!ldsp_step_clamped(e, rmax, cmax, mr, mc)) {}
    int best_r, best_c;
    sdsp_clamped(e, rmax, cmax, mr, mc, best_r, best_c);
    out1 = best_r - br;                               // bbme.py:531-532
    out0 = best_c - bc;
}

template <class E>
__device__ __forceinline__ void threestep_walk(const E &e, int bs, int sw, int H, int W, int br, int bc, int &out0,
                                               int &out1)
{
    const int span = 2 * sw + bs;
    const int steps[3] = {span / 3, span / 5, span / 10};     // bbme.py:211-213
    int dx = 0, dy = 0, tmp_dx = 0, tmp_dy = 0;
    int orow = br, ocol = bc;
#pragma unroll
    for (int s = 0; s < 3; s++) {
        const int st = steps[s];
        int r[9], c[9];
        bool ok[9];
        uint32_t cost[9];
#pragma unroll
        for (int k = 0; k < 9; k++) {                          // k = 3*ic + ir: column offset outer (bbme.py:229-231)
            const int offr = (k % 3 - 1) * st, offc = (k / 3 - 1) * st;
            const int top = orow + offr, left = ocol + offc;
            ok[k] = cand_in_frame(top, left, bs, H, W);
            r[k] = ok[k] ? top : br;                           // out-of-frame candidates are skipped, never loaded
            c[k] = ok[k] ? left : bc;
        }
        e.template eval<9>(r, c, cost);
        uint32_t best = kInfCost;
        int bx = (s == 0) ? dx : tmp_dx, by = (s == 0) ? dy : tmp_dy;
#pragma unroll
        for (int k = 0; k < 9; k++)
            if (ok[k] && cost[k] < best) { best = cost[k]; bx = (k % 3 - 1) * st; by = (k / 3 - 1) * st; }
        if (s == 0) {
            dx = bx; dy = by;
            orow = br + dx; ocol = bc + dy;                    // bbme.py:260-261
        } else if (s == 1) {
            tmp_dx = bx; tmp_dy = by;
            dx += tmp_dx; dy += tmp_dy;                        // bbme.py:296-297
            orow += dx; ocol += dy;                            // bbme.py:300-301: step-1 offset counted twice
        } else {
            tmp_dx = bx; tmp_dy = by;                          // stale step-2 value survives when nothing was in frame
            dx += tmp_dx; dy += tmp_dy;                        // bbme.py:335-336
        }
    }
    out0 = dy;                                                 // bbme.py:338-339
    out1 = dx;
}

template <class E>
__device__ __forceinline__ void twodlog_walk(const E &e, int bs, int sw, int H, int W, int br, int bc, int &out0,
                                             int &out1)
{
    int dx = 0, dy = 0;                                        // bbme.py:371
    int x = br, y = bc;
    int step = sw;
    // The centre of every iteration after the first is the winner of the previous one, so its cost is already
    // known (the reference evaluates it again and gets the same number): only the other positions are scored.
    bool have_centre = false;
    uint32_t centre_cost = 0;
    while (step > 1) {                                         // bbme.py:381
        uint32_t best = kInfCost;
        if (step > 2) {                                        // cross, bbme.py:387-393
            int r[5] = {x, x + step, x - step, x, x};
            int c[5] = {y, y, y, y + step, y - step};
            bool ok[5];
            uint32_t cost[5];
#pragma unroll
            for (int k = 0; k < 5; k++) ok[k] = cand_in_frame(r[k], c[k], bs, H, W);
            if (have_centre) {
                int pr[4], pc[4];
                uint32_t c4[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    pr[k] = ok[k + 1] ? r[k + 1] : br;
                    pc[k] = ok[k + 1] ? c[k + 1] : bc;
                }
                e.template eval<4>(pr, pc, c4);
                cost[0] = centre_cost;
#pragma unroll
                for (int k = 0; k < 4; k++) cost[k + 1] = c4[k];
            } else {
                int pr[5], pc[5];
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    pr[k] = ok[k] ? r[k] : br;
                    pc[k] = ok[k] ? c[k] : bc;
                }
                e.template eval<5>(pr, pc, cost);
            }
#pragma unroll
            for (int k = 0; k < 5; k++)
                if (ok[k] && cost[k] < best) { best = cost[k]; dx = r[k]; dy = c[k]; }
        } else {                                               // step == 2: 3x3, row outer (bbme.py:394-398)
            int r[9], c[9];
            bool ok[9];
            uint32_t cost[9];
#pragma unroll
            for (int k = 0; k < 9; k++) {
                r[k] = x + (k / 3 - 1) * 2;
                c[k] = y + (k % 3 - 1) * 2;
                ok[k] = cand_in_frame(r[k], c[k], bs, H, W);
            }
            if (have_centre) {
                int pr[8], pc[8];
                uint32_t c8[8];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int q = k < 4 ? k : k + 1;           // every position but the centre (index 4)
                    pr[k] = ok[q] ? r[q] : br;
                    pc[k] = ok[q] ? c[q] : bc;
                }
                e.template eval<8>(pr, pc, c8);
#pragma unroll
                for (int k = 0; k < 8; k++) cost[k < 4 ? k : k + 1] = c8[k];
                cost[4] = centre_cost;
            } else {
                int pr[9], pc[9];
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    pr[k] = ok[k] ? r[k] : br;
                    pc[k] = ok[k] ? c[k] : bc;
                }
                e.template eval<9>(pr, pc, cost);
            }
#pragma unroll
            for (int k = 0; k < 9; k++)
                if (ok[k] && cost[k] < best) { best = cost[k]; dx = r[k]; dy = c[k]; }
        }
        if ((dx == x && dy == y) || step == 2) step /= 2;      // bbme.py:423-425
        x = dx;
        y = dy;
        have_centre = best != kInfCost;                        // (x, y) is now the position that scored `best`
        centre_cost = best;
    }
    out1 = dx - br;                                            // bbme.py:430-431 (-position when the loop never ran)
    out0 = dy - bc;
}

template <class E>
__device__ __forceinline__ void run_search(const E &e, int procedure, int bs, int sw, int H, int W, int br, int bc,
                                           int &o0, int &o1)
{
    if (procedure == GME_SEARCH_DIAMOND)
        diamond_walk(e, bs, H, W, br, bc, o0, o1);
    else if (procedure == GME_SEARCH_THREESTEP)
        threestep_walk(e, bs, sw, H, W, br, bc, o0, o1);
    else
        twodlog_walk(e, bs, sw, H, W, br, bc, o0, o1);
}

// ---------------------------------------------------------------------------------------
// Window staging shared by the pattern and exhaustive kernels.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_window(uint8_t *smem_win, uint64_t *bar, const CUtensorMap *map, int use_tma,
                                             const uint8_t *plane, int plane_idx, int H, int W, size_t pitch, int wr0,
                                             int wc0, int win_w, int win_h)
{
    if (use_tma) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(bar, (uint32_t)(win_w * win_h));
            tma_load_3d(smem_win, map, bar, wc0, wr0, plane_idx);
        }
        mbar_wait(bar, 0);
    } else {
        // cooperative loader for planes that are not TMA-aligned: zero-fill outside the frame like TMA does
        uint32_t *w = reinterpret_cast<uint32_t *>(smem_win);
        const int pw = win_w / 4;
        for (int i = threadIdx.x; i < pw * win_h; i += blockDim.x) {
            const int rr = wr0 + i / pw, cc = wc0 + (i % pw) * 4;
            uint32_t v = 0;
            if (rr >= 0 && rr < H) {
                const uint8_t *p = plane + (size_t)rr * pitch;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (cc + k >= 0 && cc + k < W) v |= (uint32_t)p[cc + k] << (8 * k);
            }
            w[i] = v;
        }
        __syncthreads();
    }
}

template <int BS, int G, int PNORM, int NT>
__global__ void __launch_bounds__(NT, 3) bbme_pattern_kernel(const __grid_constant__ CUtensorMap cur_map, PatternArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;

    const int plane = blockIdx.z;
    const int tile_r = blockIdx.y * a.tby, tile_c = blockIdx.x * a.tbx;   // first macroblock of the tile
    // TMA needs the box to start on a 16-byte boundary of the row (a misaligned innermost coordinate
    // traps with cudaErrorIllegalInstruction on sm_100a): round the window's first column down to 16
    const int wr0 = tile_r * BS - a.margin, wc0 = (tile_c * BS - a.margin) & ~15;
    const uint8_t *prev_plane = a.prev + (size_t)plane * a.prev_stride;
    const uint8_t *cur_plane = a.cur + (size_t)plane * a.cur_stride;

    stage_window(smem, &bar, &cur_map, a.use_tma, cur_plane, plane, a.H, a.W, a.pitch, wr0, wc0, a.win_w, a.win_h);

    BlockEval<BS, G, PNORM> e;
    e.cur_plane = cur_plane;
    e.win = reinterpret_cast<const uint32_t *>(smem);
    e.pitch = a.pitch;
    e.win_pw = a.win_w / 4;
    e.wr0 = wr0;
    e.wc0 = wc0;
    e.wr1 = wr0 + a.win_h;
    e.wc1 = wc0 + a.win_w - 4;      // the funnel shift reads one word past the block: keep it inside the row
    const int lane = threadIdx.x & 31;
    e.lane_g = lane % G;
    e.gmask = (G == 32) ? 0xFFFFFFFFu : (((1u << G) - 1u) << (lane - e.lane_g));

    const int group = threadIdx.x / G, ngroups = NT / G;
    int32_t *field = a.field + (size_t)plane * a.R * a.C * 2;
    // 16, but not a constant the compiler can turn back into a shift: the kernel is bound by the ALU pipe (shifts, byte
    // permutes, VABSDIFF4, min), and `cost * k16 + j` is a multiply-add on the other one
    const uint32_t k16 = a.key_scale;
    int sum0 = 0, sum1 = 0;
    for (int b = group; b < a.tbx * a.tby; b += ngroups) {
        const int bi = tile_r + b / a.tbx, bj = tile_c + b % a.tbx;
        if (bi >= a.R || bj >= a.C) continue;
        const int br = bi * BS, bc = bj * BS;
        e.load_anchor(prev_plane, br, bc);
        int o0, o1;
        run_search(e, a.procedure, BS, a.sw, a.H, a.W, br, bc, o0, o1);
        if (e.lane_g == 0) {
            int2 v = make_int2(o0, o1);
            *reinterpret_cast<int2 *>(field + ((size_t)bi * a.C + bj) * 2) = v;
            sum0 += o0;
            sum1 += o1;
        }
    }
    if (a.sums) {
        __syncwarp();
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            sum0 += __shfl_xor_sync(0xFFFFFFFFu, sum0, o);
            sum1 += __shfl_xor_sync(0xFFFFFFFFu, sum1, o);
        }
        if (lane == 0) {
            if (sum0) atomicAdd(a.sums + 2 * plane, (unsigned long long)(long long)sum0);
            if (sum1) atomicAdd(a.sums + 2 * plane + 1, (unsigned long long)(long long)sum1);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Diamond search on 16 x 16 macroblocks -- the search the GME pipeline runs on L1 and L2
// (motion.py:224-229), i.e. the dominant kernel of the whole path.  One warp per macroblock;
// every candidate is scored straight from the staged window by all 32 lanes.  (The first
// generation kept the 20 x 20 neighbourhood of the centre in registers, one row per lane:
// 20 of 32 lanes busy, 16 of them useful per candidate, ALU-pipe bound; round 1, profiles/r01k.)
//
// Lane l owns two 4-pixel units of the macroblock: word (l & 3) of rows (l >> 2) and
// (l >> 2) + 8.  Its two anchor words stay in registers for the whole walk; a candidate is
// 2 x (two aligned LDS.32 + one funnel shift + VABSDIFF4 + IDP.4A) and one REDUX.SUM -- no
// masks, no idle lanes, no neighbourhood registers.  The window pitch is 240 bytes
// (60 words = 28 mod 32), so the eight rows x four words one warp-wide load touches fall
// into 32 different banks.  The step body is specialised on the direction of the previous
// move (nine classes of static code, 9 KB in all: a first version that was also specialised
// on the byte alignment of the centre column, 58 KB, ran out of the 32 KB instruction cache
// and stalled on fetches, profiles/r02b): the row offsets are immediates, the word address
// and shift of a column offset are computed once per step, and a step scores only the
// positions the previous step has not scored (the
// cost of a position does not depend on the step that asks for it, so reuse is exact; the
// strict-'<' scan over the nine costs in the reference's order keeps the tie-breaking).
// Centres within 2 pixels of the clamp bounds of bbme.py:503-504, or whose candidates
// leave the staged window, take the candidate-at-a-time evaluator (BlockEval<16, 32>, the
// same lane mapping and anchor registers), which clamps and reads global memory.
// ---------------------------------------------------------------------------------------
constexpr int kD16Pitch = 240;                         // window row pitch, bytes (TMA box width)
constexpr int kD16Rows = 128;                          // window rows: 4 macroblock rows + 2 x 32 margin
constexpr int kD16Margin = 32;

__host__ __device__ constexpr int ldsp_r(int k)
{
    constexpr int t[9] = {0, 2, 1, 0, -1, -2, -1, 0, 1};          // bbme.py:463-472 (row offsets)
    return t[k];
}
__host__ __device__ constexpr int ldsp_c(int k)
{
    constexpr int t[9] = {0, 0, 1, 2, 1, 0, -1, -2, -1};
    return t[k];
}
// The centre moved by LDSP offset kb: candidate j of the new step is the position candidate ldsp_reuse(kb, j) of the
// previous step had (-1: not scored yet).  kb = 0: no previous step.
__host__ __device__ constexpr int ldsp_reuse(int kb, int j)
{
    if (kb == 0) return -1;
    const int r = ldsp_r(kb) + ldsp_r(j), c = ldsp_c(kb) + ldsp_c(j);
    for (int i = 0; i < 9; i++)
        if (ldsp_r(i) == r && ldsp_c(i) == c) return i;
    return -1;
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// 16 x cost of the candidate (DR, DC) pixels away from the centre.  `xb`: shared BYTE address of this lane's first unit
// for the centre candidate (any alignment; warp-uniform modulo 4): the unit is read as two aligned words and
// funnel-shifted into place -- the word address and the shift depend on DC only, so candidates of one step share them.
// The scale (cost << 4) leaves room for the candidate index of the first-minimum key.
template <int PNORM, int DR, int DC>
__device__ __forceinline__ uint32_t d16_cand(uint32_t xb, uint32_t a0, uint32_t a1)
{
    const uint32_t A = xb + (uint32_t)DC;
    const uint32_t addr = (A & ~3u) + (uint32_t)(DR * kD16Pitch);
    const uint32_t sh = A << 3;                                    // the funnel shift uses the low five bits: (A % 4) * 8
    const uint32_t u0 = lds_u32(addr), u1 = lds_u32(addr + 4);
    const uint32_t u2 = lds_u32(addr + 8 * kD16Pitch), u3 = lds_u32(addr + 8 * kD16Pitch + 4);
    uint32_t acc = cost4_acc<PNORM>(__funnelshift_r(u0, u1, sh), a0, 0u);
    acc = cost4_acc<PNORM>(__funnelshift_r(u2, u3, sh), a1, acc);
    return __reduce_add_sync(0xFFFFFFFFu, acc) << 4;
}

// The same for a centre whose column is word-aligned in the window (every fresh walk: macroblock columns and the
// window origin are multiples of 16): the word offset and the shift of a column offset are compile-time constants,
// and the three candidates in the centre's own column need no second word and no shift.  (Measured: 351 -> 336 us on
// the full-resolution level.  Extending it to the vertical moves, which keep the alignment, and to the SDSP gave
// nothing more, 335.9 us, and was dropped.  The predicate must come from warp-uniform values such as mc - wc0: derived
// from xb, which holds the lane offset, every warp reduction downstream got a divergent-warp fallback -- 2.2 x the code.)
template <int PNORM, int DR, int DC>
__device__ __forceinline__ uint32_t d16_cand_aligned(uint32_t xb, uint32_t a0, uint32_t a1)
{
    constexpr int wofs = DC < 0 ? -4 : 0;
    constexpr int sh = ((DC + 4) & 3) * 8;
    const uint32_t addr = xb + (uint32_t)(DR * kD16Pitch + wofs);
    uint32_t v0, v1;
    if constexpr (sh == 0) {
        v0 = lds_u32(addr);
        v1 = lds_u32(addr + 8 * kD16Pitch);
    } else {
        v0 = __funnelshift_r(lds_u32(addr), lds_u32(addr + 4), sh);
        v1 = __funnelshift_r(lds_u32(addr + 8 * kD16Pitch), lds_u32(addr + 8 * kD16Pitch + 4), sh);
    }
    uint32_t acc = cost4_acc<PNORM>(v0, a0, 0u);
    acc = cost4_acc<PNORM>(v1, a1, acc);
    return __reduce_add_sync(0xFFFFFFFFu, acc) << 4;
}

template <int PNORM>
__device__ __forceinline__ void d16_first_aligned(uint32_t (&c)[9], uint32_t xb, uint32_t a0, uint32_t a1)
{
    c[0] = d16_cand_aligned<PNORM, ldsp_r(0), ldsp_c(0)>(xb, a0, a1);
    c[1] = d16_cand_aligned<PNORM, ldsp_r(1), ldsp_c(1)>(xb, a0, a1);
    c[2] = d16_cand_aligned<PNORM, ldsp_r(2), ldsp_c(2)>(xb, a0, a1);
    c[3] = d16_cand_aligned<PNORM, ldsp_r(3), ldsp_c(3)>(xb, a0, a1);
    c[4] = d16_cand_aligned<PNORM, ldsp_r(4), ldsp_c(4)>(xb, a0, a1);
    c[5] = d16_cand_aligned<PNORM, ldsp_r(5), ldsp_c(5)>(xb, a0, a1);
    c[6] = d16_cand_aligned<PNORM, ldsp_r(6), ldsp_c(6)>(xb, a0, a1);
    c[7] = d16_cand_aligned<PNORM, ldsp_r(7), ldsp_c(7)>(xb, a0, a1);
    c[8] = d16_cand_aligned<PNORM, ldsp_r(8), ldsp_c(8)>(xb, a0, a1);
}

template <int PNORM, int KB, int J>
__device__ __forceinline__ uint32_t d16_next(const uint32_t (&c)[9], uint32_t xb, uint32_t a0, uint32_t a1)
{
    constexpr int i = ldsp_reuse(KB, J);
    if constexpr (i >= 0)
        return c[i];
    else
        return d16_cand<PNORM, ldsp_r(J), ldsp_c(J)>(xb, a0, a1);
}

struct D16Fast {          // centres for which the immediate-offset path applies (see the kernel)
    int r_lo, r_hi, c_lo, c_hi;
    __device__ __forceinline__ bool ok(int r, int c) const { return r >= r_lo && r <= r_hi && c >= c_lo && c <= c_hi; }
};

// One LDSP step (bbme.py:494-513) after the move KB (1..8): the centre, its shared address and the costs already
// known move with static offsets; false (nothing evaluated, c[] stale) when the new centre is not a fast-path one.
template <int PNORM, int KB>
__device__ __forceinline__ bool d16_update(uint32_t (&c)[9], int &mr, int &mc, uint32_t &xb, const D16Fast &f,
                                           uint32_t a0, uint32_t a1)
{
    if constexpr (KB != 0) {
        // the previous centre was a fast-path one, so only the bounds the move runs towards can be crossed
        mr += ldsp_r(KB);
        mc += ldsp_c(KB);
        if (ldsp_r(KB) > 0 && mr > f.r_hi) return false;
        if (ldsp_r(KB) < 0 && mr < f.r_lo) return false;
        if (ldsp_c(KB) > 0 && mc > f.c_hi) return false;
        if (ldsp_c(KB) < 0 && mc < f.c_lo) return false;
        xb += (uint32_t)(ldsp_r(KB) * kD16Pitch + ldsp_c(KB));
    }
    const uint32_t n0 = d16_next<PNORM, KB, 0>(c, xb, a0, a1), n1 = d16_next<PNORM, KB, 1>(c, xb, a0, a1),
                   n2 = d16_next<PNORM, KB, 2>(c, xb, a0, a1), n3 = d16_next<PNORM, KB, 3>(c, xb, a0, a1),
                   n4 = d16_next<PNORM, KB, 4>(c, xb, a0, a1), n5 = d16_next<PNORM, KB, 5>(c, xb, a0, a1),
                   n6 = d16_next<PNORM, KB, 6>(c, xb, a0, a1), n7 = d16_next<PNORM, KB, 7>(c, xb, a0, a1),
                   n8 = d16_next<PNORM, KB, 8>(c, xb, a0, a1);
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3; c[4] = n4; c[5] = n5; c[6] = n6; c[7] = n7; c[8] = n8;
    return true;
}

template <int PNORM>
__device__ __forceinline__ bool d16_ldsp(uint32_t (&c)[9], int kb, int &mr, int &mc, uint32_t &xb, const D16Fast &f,
                                         uint32_t a0, uint32_t a1)
{
    switch (kb) {
    case 1: return d16_update<PNORM, 1>(c, mr, mc, xb, f, a0, a1);
    case 2: return d16_update<PNORM, 2>(c, mr, mc, xb, f, a0, a1);
    case 3: return d16_update<PNORM, 3>(c, mr, mc, xb, f, a0, a1);
    case 4: return d16_update<PNORM, 4>(c, mr, mc, xb, f, a0, a1);
    case 5: return d16_update<PNORM, 5>(c, mr, mc, xb, f, a0, a1);
    case 6: return d16_update<PNORM, 6>(c, mr, mc, xb, f, a0, a1);
    case 7: return d16_update<PNORM, 7>(c, mr, mc, xb, f, a0, a1);
    default: return d16_update<PNORM, 8>(c, mr, mc, xb, f, a0, a1);
    }
}

// the final SDSP (bbme.py:515-529) around the same centre: key of the first strict minimum, 16 * cost + index, over
// the effective (row, col) order (0,0) (0,1) (1,0) (0,-1) (-1,0) -- the reference applies its offset list swapped
template <int PNORM>
__device__ __forceinline__ uint32_t d16_sdsp(uint32_t centre_cost16, uint32_t xb, uint32_t a0, uint32_t a1)
{
    uint32_t key = centre_cost16;
    key = min(key, d16_cand<PNORM, 0, 1>(xb, a0, a1) + 1u);
    key = min(key, d16_cand<PNORM, 1, 0>(xb, a0, a1) + 2u);
    key = min(key, d16_cand<PNORM, 0, -1>(xb, a0, a1) + 3u);
    key = min(key, d16_cand<PNORM, -1, 0>(xb, a0, a1) + 4u);
    return key;
}

template <int PNORM, int NT>
__global__ void __launch_bounds__(NT, 4) bbme_diamond16_kernel(const __grid_constant__ CUtensorMap cur_map, PatternArgs a)
{
    constexpr int BS = 16;
    constexpr int TBX = 8, TBY = 4;                   // macroblocks per tile (fixed: index math by shifts)
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ int next_block;                        // walks differ in length: warps take macroblocks from a queue
    if (threadIdx.x == 0) next_block = NT / 32;       // (the first NT/32 are handed out statically)

    const int plane = blockIdx.z;
    const int tile_r = blockIdx.y * TBY, tile_c = blockIdx.x * TBX;
    const int wr0 = tile_r * BS - kD16Margin, wc0 = (tile_c * BS - kD16Margin) & ~15;
    const uint8_t *prev_plane = a.prev + (size_t)plane * a.prev_stride;
    const uint8_t *cur_plane = a.cur + (size_t)plane * a.cur_stride;

    stage_window(smem, &bar, &cur_map, a.use_tma, cur_plane, plane, a.H, a.W, a.pitch, wr0, wc0, kD16Pitch, kD16Rows);

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0);   // tells the compiler it is warp-uniform
    BlockEval<BS, 32, PNORM> e;                       // border / far-away evaluator; owns the anchor registers
    e.cur_plane = cur_plane;
    e.win = reinterpret_cast<const uint32_t *>(smem);
    e.pitch = a.pitch;
    e.win_pw = kD16Pitch / 4;
    e.wr0 = wr0;
    e.wc0 = wc0;
    e.wr1 = wr0 + kD16Rows;
    e.wc1 = wc0 + kD16Pitch - 4;
    e.lane_g = lane;
    e.gmask = 0xFFFFFFFFu;

    constexpr unsigned SRP = 2u | (2u << 3) | (3u << 6) | (2u << 9) | (1u << 12);
    constexpr unsigned SCP = 2u | (3u << 3) | (2u << 6) | (1u << 9) | (2u << 12);
    const int rmax = a.H - BS - 1, cmax = a.W - BS - 1;        // bbme.py:503-504 (off by one, kept)
    // centres for which the immediate-offset path applies: no clamp can act on the 5 x 5 neighbourhood of offsets, and
    // every word a candidate reads (16 + 2 rows, 16 + 2 columns + the word the funnel shift completes) is staged
    D16Fast fast, inwin;
    fast.r_lo = max(2, wr0 + 2); fast.r_hi = min(rmax - 2, wr0 + kD16Rows - BS - 2);
    fast.c_lo = max(2, wc0 + 2); fast.c_hi = min(cmax - 2, wc0 + kD16Pitch - BS - 2 - 7);
    // centres whose clamped candidates all lie in the staged window (a clamped position lies between the centre and
    // the unclamped one): the frame-border blocks
    inwin.r_lo = wr0 + 2; inwin.r_hi = wr0 + kD16Rows - BS - 2;
    inwin.c_lo = wc0 + 2; inwin.c_hi = wc0 + kD16Pitch - BS - 2 - 7;
    // this lane's first unit: row lane / 4, word lane % 4 (the second is eight rows below)
    const uint32_t lane_base = smem_u32(smem) + (uint32_t)((lane >> 2) * kD16Pitch + (lane & 3) * 4);
    int32_t *field = a.field + (size_t)plane * a.R * a.C * 2;

    for (int b = warp; b < TBX * TBY; b = __shfl_sync(0xFFFFFFFFu, lane == 0 ? atomicAdd(&next_block, 1) : 0, 0)) {
        const int bi = tile_r + b / TBX, bj = tile_c + b % TBX;
        if (bi >= a.R || bj >= a.C) continue;
        const int br = bi * BS, bc = bj * BS;
        e.load_anchor(prev_plane, br, bc);
        const uint32_t a0 = e.anchor[0], a1 = e.anchor[1];

        int mr = br, mc = bc;
        uint32_t c[9];                                             // 16 x cost of the nine LDSP positions around (mr, mc)
        int kb = 0;                                                // the move that leads to the next centre; 0: c[] holds nothing
        bool last_fast = false;
        uint32_t xb = 0;
        for (;;) {                                                 // LDSP, bbme.py:494-513
            if (kb == 0) {                                         // (re)start at (mr, mc)
                if (!fast.ok(mr, mc)) {
                    bool stop;
                    if (inwin.ok(mr, mc)) {
                        stop = ldsp_step_clamped(WindowOnly<decltype(e)>{e}, rmax, cmax, mr, mc);
                    } else {
                        const int2 to = ldsp_step_cold(e, rmax, cmax, mr, mc);
                        stop = to.x == mr && to.y == mc;
                        mr = to.x;
                        mc = to.y;
                    }
                    if (stop) { last_fast = false; break; }
                    continue;
                }
                xb = lane_base + (uint32_t)((mr - wr0) * kD16Pitch + (mc - wc0));
                if (((mc - wc0) & 3) == 0)                         // (warp-uniform) every fresh walk starts here
                    d16_first_aligned<PNORM>(c, xb, a0, a1);
                else                                               // a restart after the border evaluator moved the centre
                    d16_update<PNORM, 0>(c, mr, mc, xb, fast, a0, a1);
            } else if (!d16_ldsp<PNORM>(c, kb, mr, mc, xb, fast, a0, a1)) {
                kb = 0;                                            // the move took the centre off the fast path
                continue;
            }
            // first strict minimum in candidate order (bbme.py:506-510): costs < 2^24, so 16 * cost + index fits
            uint32_t key = c[0];
#pragma unroll
            for (int j = 1; j < 9; j++) key = min(key, c[j] + (uint32_t)j);
            kb = (int)(key & 15u);
            if (kb == 0) { last_fast = true; break; }              // the centre wins: positions are distinct here
        }

        int out_r, out_c;
        if (last_fast) {                                           // SDSP around the centre of the last step
            const int ks = (int)(d16_sdsp<PNORM>(c[0], xb, a0, a1) & 15u);
            out_r = mr + (int)((SRP >> (3 * ks)) & 7u) - 2;
            out_c = mc + (int)((SCP >> (3 * ks)) & 7u) - 2;
        } else if (inwin.ok(mr, mc)) {
            sdsp_clamped(WindowOnly<decltype(e)>{e}, rmax, cmax, mr, mc, out_r, out_c);
        } else {
            const int2 o = sdsp_cold(e, rmax, cmax, mr, mc);
            out_r = o.x;
            out_c = o.y;
        }
        if (lane == 0)
            *reinterpret_cast<int2 *>(field + ((size_t)bi * a.C + bj) * 2) = make_int2(out_c - bc, out_r - br);   // bbme.py:531-532
    }
}

// ---------------------------------------------------------------------------------------
// Diamond search on 2 x 2 blocks -- the dense first estimate of the GME pipeline (motion.py:27-29).
//
// One THREAD per block (the walks of neighbouring blocks diverge, so nothing is shared between
// lanes).  The nine LDSP candidates lie in the 6 x 6 pixel neighbourhood of the centre: the thread
// loads it once per step (6 rows x 3 aligned words, byte-aligned with two funnel shifts per row),
// then each candidate's four pixels are gathered with ONE byte permute from two row registers and
// scored with VABSDIFF4 (+ IDP.4A) against the anchor word.  The first minimum comes from a
// min over (cost << 4 | index) keys.  The SDSP reuses the registers of the last LDSP step.  Steps
// whose neighbourhood leaves the staged window or on which the clamp of bbme.py:503-504 could act
// use the candidate-at-a-time evaluator (BlockEval<2, 1>).
// Round 1 measured 15.5 of 32 lanes active here and blamed the different walk lengths; the instruction-level profile
// (profiles/r02i) showed something else: the blocks on which the clamp acts -- one or two lanes of a quarter of all
// warps -- dragged their warps through the ~900-instruction general evaluator.  A per-lane work queue (tried, r02i)
// made it worse: its hand-over and finish code ran in every iteration with a few lanes.  What is built: the clamped
// block columns (first, last two) are tiles of their own and the clamped block rows are whole warps (a warp is 32
// consecutive blocks of a tile row), so border blocks only meet border blocks; they use a window-only evaluator, and
// the general one (walks that leave the window) is out of line.  The per-plane channel sums that the first estimate
// needs are accumulated here (see PatternArgs::sums).
// What remains is the walk-length divergence (17 of 32 lanes active on average).  Compacting the walks between steps
// was built and measured (round 2): centre + anchor of every block in shared memory, rounds of 1..k LDSP steps, the
// blocks still walking appended to the next round's list with one ballot and one shared atomic per warp, register-path
// and clamped centres in separate lists.  Bit-identical, and slower: 81.7 us (k = 2) .. 88 us (k = 1) against 73 us --
// the list/state indirection alone costs 12 us (k = 100, no re-queueing: 85 us), compaction wins back 4.
// ---------------------------------------------------------------------------------------
template <int PNORM, int NT>
__global__ void __launch_bounds__(NT, 3) bbme_diamond2_kernel(const __grid_constant__ CUtensorMap cur_map, PatternArgs a)
{
    constexpr int BS = 2;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;

    const int plane = blockIdx.z;
    // Tile columns: when the frame is wide enough (a.edge_tiles, set by the launcher), the block columns on which the clamp
    // of bbme.py:503-504 acts -- the first one and the last two -- are tiles of their own.  Those blocks take the
    // candidate-at-a-time evaluator; inside an ordinary tile they would be one or two lanes of every warp and make
    // each of those warps run that (long) code with the other lanes idle.
    constexpr int TBX = 64;                           // a.tbx of a full tile (index math by shifts)
    const int tile_r = blockIdx.y * a.tby;
    int tile_c, tw, shift = 6;                        // first block column, width, log2 of the queue's row stride
    if (!a.edge_tiles) { tile_c = blockIdx.x * TBX; tw = min(TBX, a.C - tile_c); }
    else if (blockIdx.x == 0) { tile_c = 0; tw = 1; shift = 0; }
    else if (blockIdx.x == gridDim.x - 1) { tile_c = a.C - 2; tw = 2; shift = 1; }
    else { tile_c = 1 + (blockIdx.x - 1) * TBX; tw = min(TBX, a.C - 2 - tile_c); }
    const int wr0 = tile_r * BS - a.margin, wc0 = (tile_c * BS - a.margin) & ~15;
    const uint8_t *prev_plane = a.prev + (size_t)plane * a.prev_stride;
    const uint8_t *cur_plane = a.cur + (size_t)plane * a.cur_stride;

    stage_window(smem, &bar, &cur_map, a.use_tma, cur_plane, plane, a.H, a.W, a.pitch, wr0, wc0, a.win_w, a.win_h);

    BlockEval<BS, 1, PNORM> e;                        // fallback evaluator
    e.cur_plane = cur_plane;
    e.win = reinterpret_cast<const uint32_t *>(smem);
    e.pitch = a.pitch;
    e.win_pw = a.win_w / 4;
    e.wr0 = wr0;
    e.wc0 = wc0;
    e.wr1 = wr0 + a.win_h;
    e.wc1 = wc0 + a.win_w - 4;
    e.lane_g = 0;
    e.gmask = 1u << (threadIdx.x & 31);

    constexpr unsigned long long LRP = 0x321012342ull, LCP = 0x101234322ull;   // the LDSP / SDSP tables as nibbles (value + 2)
    constexpr unsigned SRP = 0x12322u, SCP = 0x21232u;
    const int rmax = a.H - BS - 1, cmax = a.W - BS - 1;        // bbme.py:503-504 (off by one, kept)
    // centres for which the register path applies: no clamp can act, and rows mr-2 .. mr+3 / the three aligned
    // words that hold columns mc-2 .. mc+3 lie inside the staged window
    const int fr_lo = max(2, wr0 + 2), fr_hi = min(rmax - 2, wr0 + a.win_h - 4);
    const int fc_lo = max(2, wc0 + 2), fc_hi = min(cmax - 2, wc0 + 2 + a.win_w - 12);
    const uint32_t *win = reinterpret_cast<const uint32_t *>(smem);
    const int win_pw = a.win_w / 4;
    const int lane = threadIdx.x & 31;
    const int total = a.tby << shift;                            // tby rows of 2^shift slots (slots >= tw are skipped)
    const int rows_left = a.R - tile_r;                          // block rows of the frame this tile can hold

    int32_t *field = a.field + (size_t)plane * a.R * a.C * 2;
    // 16, but not a constant the compiler can turn back into a shift: the kernel is bound by the ALU pipe (shifts, byte
    // permutes, VABSDIFF4, min), and `cost * k16 + j` is a multiply-add on the other one
    const uint32_t k16 = a.key_scale;
    int sum0 = 0, sum1 = 0;
    const WindowOnly<decltype(e)> w{e};
    for (int b = threadIdx.x; b < total; b += NT) {              // a warp: 32 consecutive blocks of one tile row
        const int lr = b >> shift, lc = b & ((1 << shift) - 1);
        if (lr >= rows_left || lc >= tw) continue;
        const int bi = tile_r + lr, bj = tile_c + lc;
        const int br = bi * BS, bc = bj * BS;
        const uint8_t *pa = prev_plane + (unsigned)(br * (int)a.pitch + bc);      // 2-byte aligned: even column, pitch % 4 == 0
        const uint32_t anchor = (uint32_t)__ldg(reinterpret_cast<const unsigned short *>(pa)) |
                                ((uint32_t)__ldg(reinterpret_cast<const unsigned short *>(pa + a.pitch)) << 16);
        e.anchor[0] = anchor & 0xFFFFu;
        e.anchor[1] = anchor >> 16;

        uint32_t z0[6], z1[6];                                   // row i: z0 = bytes 0..3, z1 = bytes 4..7 (byte 0 = column mc - 2)
        auto cost_of = [&](uint32_t px) -> uint32_t { return cost4_acc<PNORM>(px, anchor, 0u); };
        int mr = br, mc = bc;
        bool last_fast = false;
        uint32_t centre_cost = 0;
        for (;;) {                                               // LDSP, bbme.py:494-513
            if (mr >= fr_lo && mr <= fr_hi && mc >= fc_lo && mc <= fc_hi) {
                const int x = mc - 2 - wc0;
                const uint32_t *p = win + (mr - 2 - wr0) * win_pw + (x >> 2);
                const int sh = (x & 3) * 8;
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    const uint32_t w0 = p[i * win_pw], w1 = p[i * win_pw + 1], w2 = p[i * win_pw + 2];
                    z0[i] = __funnelshift_r(w0, w1, sh);
                    z1[i] = __funnelshift_r(w1, w2, sh);
                }
                // candidate (dr, dc): rows dr+2, dr+3; bytes dc+2, dc+3 of each
                uint32_t c[9];
                c[0] = cost_of(__byte_perm(z0[2], z0[3], 0x7632));                                            // ( 0,  0)
                c[1] = cost_of(__byte_perm(z0[4], z0[5], 0x7632));                                            // ( 2,  0)
                c[2] = cost_of(__byte_perm(__funnelshift_r(z0[3], z1[3], 24), __funnelshift_r(z0[4], z1[4], 24), 0x5410));   // ( 1,  1)
                c[3] = cost_of(__byte_perm(z1[2], z1[3], 0x5410));                                            // ( 0,  2)
                c[4] = cost_of(__byte_perm(__funnelshift_r(z0[1], z1[1], 24), __funnelshift_r(z0[2], z1[2], 24), 0x5410));   // (-1,  1)
                c[5] = cost_of(__byte_perm(z0[0], z0[1], 0x7632));                                            // (-2,  0)
                c[6] = cost_of(__byte_perm(z0[1], z0[2], 0x6521));                                            // (-1, -1)
                c[7] = cost_of(__byte_perm(z0[2], z0[3], 0x5410));                                            // ( 0, -2)
                c[8] = cost_of(__byte_perm(z0[3], z0[4], 0x6521));                                            // ( 1, -1)
                uint32_t key = c[0] * k16;                       // first strict minimum in candidate order; costs < 2^18
#pragma unroll
                for (int j = 1; j < 9; j++) key = min(key, c[j] * k16 + (uint32_t)j);
                const int kb = (int)(key & 15u);
                if (kb == 0) { last_fast = true; centre_cost = c[0]; break; }
                mr += (int)((LRP >> (4 * kb)) & 15) - 2;
                mc += (int)((LCP >> (4 * kb)) & 15) - 2;
            } else if (mr >= wr0 + 2 && mr <= wr0 + a.win_h - 4 && mc >= wc0 + 2 && mc <= wc0 + 2 + a.win_w - 12) {
                // frame-border block: the clamp acts, but every clamped candidate is in the staged window
                if (ldsp_step_clamped(w, rmax, cmax, mr, mc)) { last_fast = false; break; }
            } else {
                const int2 to = ldsp_step_cold(e, rmax, cmax, mr, mc);
                const bool stop = to.x == mr && to.y == mc;
                mr = to.x;
                mc = to.y;
                if (stop) { last_fast = false; break; }
            }
        }
        int out_r, out_c;
        if (last_fast) {                                         // SDSP on the registers of the last step, bbme.py:515-529
            uint32_t key = centre_cost * k16;
            key = min(key, cost_of(__byte_perm(__funnelshift_r(z0[2], z1[2], 24), __funnelshift_r(z0[3], z1[3], 24), 0x5410)) * k16 + 1u);   // (0, +1)
            key = min(key, cost_of(__byte_perm(z0[3], z0[4], 0x7632)) * k16 + 2u);                           // (+1, 0)
            key = min(key, cost_of(__byte_perm(z0[2], z0[3], 0x6521)) * k16 + 3u);                           // (0, -1)
            key = min(key, cost_of(__byte_perm(z0[1], z0[2], 0x7632)) * k16 + 4u);                           // (-1, 0)
            const int ks = (int)(key & 15u);
            out_r = mr + (int)((SRP >> (4 * ks)) & 15) - 2;
            out_c = mc + (int)((SCP >> (4 * ks)) & 15) - 2;
        } else if (mr >= wr0 + 2 && mr <= wr0 + a.win_h - 4 && mc >= wc0 + 2 && mc <= wc0 + 2 + a.win_w - 12) {
            sdsp_clamped(w, rmax, cmax, mr, mc, out_r, out_c);
        } else {
            const int2 o = sdsp_cold(e, rmax, cmax, mr, mc);
            out_r = o.x;
            out_c = o.y;
        }
        const int o0 = out_c - bc, o1 = out_r - br;              // bbme.py:531-532
        *reinterpret_cast<int2 *>(field + ((size_t)bi * a.C + bj) * 2) = make_int2(o0, o1);
        sum0 += o0;
        sum1 += o1;
    }
    __syncwarp();
    if (a.sums) {
        sum0 = __reduce_add_sync(0xFFFFFFFFu, sum0);
        sum1 = __reduce_add_sync(0xFFFFFFFFu, sum1);
        if (lane == 0) {
            if (sum0) atomicAdd(a.sums + 2 * plane, (unsigned long long)(long long)sum0);
            if (sum1) atomicAdd(a.sums + 2 * plane + 1, (unsigned long long)(long long)sum1);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Generic path: any block size 1..255, a full warp per macroblock, global memory only.
// ---------------------------------------------------------------------------------------
template <int PNORM>
struct GenericEval {
    const uint8_t *prev_block;   // anchor, top-left pixel
    const uint8_t *cur_plane;
    size_t pitch;
    int bs, lane;

    template <int N>
    __device__ __forceinline__ void eval(const int (&r)[N], const int (&c)[N], uint32_t (&cost)[N]) const
    {
        const int npix = bs * bs;
#pragma unroll 1
        for (int k = 0; k < N; k++) {
            const uint8_t *cand = cur_plane + (size_t)r[k] * pitch + c[k];
            // exact warp-wide sum over the row-major pixels [start, start + len) of the block
            auto range_sum = [&](int start, int len) -> uint32_t {
                uint32_t acc = 0;
                for (int p = start + lane; p < start + len; p += 32) {
                    const int pr = p / bs, pc = p - pr * bs;
                    const int d = (int)prev_block[(size_t)pr * pitch + pc] - (int)cand[(size_t)pr * pitch + pc];
                    acc += (PNORM == GME_PNORM_MAE) ? (uint32_t)abs(d) : (uint32_t)(d * d);
                }
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
                return acc;
            };
            if (PNORM == GME_PNORM_MSE && bs > 16)
                cost[k] = pairwise_sum_f32_tree(npix, range_sum);     // the reference's float32 sum rounds here
            else
                cost[k] = range_sum(0, npix);                         // exact in the reference's float32 too
        }
    }
};

template <int PNORM>
__global__ void __launch_bounds__(256) bbme_pattern_generic_kernel(PatternArgs a, int bs)
{
    const int plane = blockIdx.z;
    const int warps_per_cta = blockDim.x / 32;
    const long nblocks = (long)a.R * a.C;
    const long b = (long)blockIdx.x * warps_per_cta + threadIdx.x / 32;
    if (b >= nblocks) return;
    const int bi = (int)(b / a.C), bj = (int)(b % a.C);
    const int br = bi * bs, bc = bj * bs;
    GenericEval<PNORM> e;
    e.prev_block = a.prev + (size_t)plane * a.prev_stride + (size_t)br * a.pitch + bc;
    e.cur_plane = a.cur + (size_t)plane * a.cur_stride;
    e.pitch = a.pitch;
    e.bs = bs;
    e.lane = threadIdx.x & 31;
    int o0, o1;
    run_search(e, a.procedure, bs, a.sw, a.H, a.W, br, bc, o0, o1);
    if (e.lane == 0) {
        int32_t *f = a.field + ((size_t)plane * nblocks + b) * 2;
        f[0] = o0;
        f[1] = o1;
    }
}

// ---------------------------------------------------------------------------------------
// Host launcher
// ---------------------------------------------------------------------------------------
template <int BS, int G, int PNORM>
static int launch_fast(PatternArgs a, int n, cudaStream_t stream)
{
    constexpr int NT = 256;
    // tile: about 128 x 64 pixels of macroblocks, margin sized to typical motion; the window row
    // length is = 16 (mod 32) bytes so that consecutive rows start 4 banks apart (conflict-free
    // for 8 rows x 4 words, the footprint of one warp-wide unit load)
    int tbx = max(1, 128 / BS), tby = max(1, 64 / BS);
    if (BS <= 4) { tbx = 64 / BS; tby = 32 / BS; }
    tbx = min(tbx, a.C);
    tby = min(tby, a.R);
    // three-step and 2D-log take big first steps (up to 2 * int(span/3) + ...): a wider margin keeps their
    // candidates inside the staged window; win_w stays within the 256-byte TMA box
    int margin = (BS <= 4) ? 12 : 32;
    if (BS >= 8 && a.procedure != GME_SEARCH_DIAMOND) margin = 40;
    int win_w = tbx * BS + 2 * margin + 4 + 15;          // +4: the funnel shift reads one word beyond the block;
    win_w = (win_w + 15) / 16 * 16;                      // +15: the first column is rounded down to 16 bytes (TMA)
    if (win_w % 32 == 0) win_w += 16;
    const int win_h = tby * BS + 2 * margin;
    if (win_w > 256 || win_h > 256) return GME_ERR_UNSUPPORTED;   // TMA box limit; not reachable with the tiles above
    a.tbx = tbx; a.tby = tby; a.margin = margin; a.win_w = win_w; a.win_h = win_h;
    CUtensorMap map;
    a.use_tma = make_plane_tensor_map(&map, a.cur, n, a.H, a.W, a.pitch, a.cur_stride, win_w, win_h) ? 1 : 0;
    if (!a.use_tma) memset(&map, 0, sizeof(map));
    const size_t smem = (size_t)win_w * win_h + 32;      // slack for the trailing word of the last row
    auto kern = bbme_pattern_kernel<BS, G, PNORM, NT>;
    ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem);
    dim3 grid((a.C + tbx - 1) / tbx, (a.R + tby - 1) / tby, n);
    kern<<<grid, NT, smem, stream>>>(map, a);
    note_launch();
    return check_launch("bbme_pattern_kernel");
}

template <int PNORM>
static int launch_diamond16(PatternArgs a, int n, cudaStream_t stream)
{
    constexpr int NT = 256, BS = 16;
    const int tbx = 8, tby = 4;                          // fixed in the kernel (TBX, TBY)
    a.tbx = tbx; a.tby = tby; a.margin = kD16Margin; a.win_w = kD16Pitch; a.win_h = kD16Rows;
    CUtensorMap map;
    a.use_tma = make_plane_tensor_map(&map, a.cur, n, a.H, a.W, a.pitch, a.cur_stride, kD16Pitch, kD16Rows) ? 1 : 0;
    if (!a.use_tma) memset(&map, 0, sizeof(map));
    const size_t smem = (size_t)kD16Pitch * kD16Rows + 32;   // slack: the border evaluator's funnel shift reads one word past a row
    auto kern = bbme_diamond16_kernel<PNORM, NT>;
    ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem);
    dim3 grid((a.C + tbx - 1) / tbx, (a.R + tby - 1) / tby, n);
    kern<<<grid, NT, smem, stream>>>(map, a);
    note_launch();
    return check_launch("bbme_diamond16_kernel");
}

template <int PNORM>
static int launch_diamond2(PatternArgs a, int n, cudaStream_t stream)
{
    constexpr int NT = 256, BS = 2;
    // 1024 blocks per CTA, four per thread
    const int tbx = 64, tby = min(16, a.R);              // (tbx is fixed in the kernel; 16 rows measured best of 4..32)
    const int margin = 12;
    const int edge_tiles = a.C >= 8 ? 1 : 0;             // the clamped block columns (first, last two) as tiles of their own
    a.edge_tiles = edge_tiles;
    int win_w = tbx * BS + 2 * margin + 4 + 15;
    win_w = (win_w + 15) / 16 * 16;
    if (win_w % 32 == 0) win_w += 16;
    const int win_h = tby * BS + 2 * margin;
    a.tbx = tbx; a.tby = tby; a.margin = margin; a.win_w = win_w; a.win_h = win_h;
    CUtensorMap map;
    a.use_tma = make_plane_tensor_map(&map, a.cur, n, a.H, a.W, a.pitch, a.cur_stride, win_w, win_h) ? 1 : 0;
    if (!a.use_tma) memset(&map, 0, sizeof(map));
    const size_t smem = (size_t)win_w * win_h + 32;
    auto kern = bbme_diamond2_kernel<PNORM, NT>;
    const int tiles_x = edge_tiles ? 2 + (a.C - 3 + tbx - 1) / tbx : (a.C + tbx - 1) / tbx;
    dim3 grid(tiles_x, (a.R + tby - 1) / tby, n);
    kern<<<grid, NT, smem, stream>>>(map, a);
    note_launch();
    return check_launch("bbme_diamond2_kernel");
}

template <int PNORM>
static int launch_pattern_pn(PatternArgs a, int n, int bs, cudaStream_t stream)
{
    if (bs == 2 && a.procedure == GME_SEARCH_DIAMOND) return launch_diamond2<PNORM>(a, n, stream);
    if (bs == 16 && a.procedure == GME_SEARCH_DIAMOND && !a.sums) return launch_diamond16<PNORM>(a, n, stream);
    switch (bs) {
    case 2: return launch_fast<2, 1, PNORM>(a, n, stream);
    case 4: return launch_fast<4, 1, PNORM>(a, n, stream);
    case 8: return launch_fast<8, 4, PNORM>(a, n, stream);
    case 12: return launch_fast<12, 4, PNORM>(a, n, stream);
    // (three-step / 2D-log with a warp per macroblock: G = 32 in this kernel measured 2-10 % slower in round 1; a
    //  dedicated kernel on the diamond kernel's skeleton -- queue, 240-byte pitch, window-only evaluator, general one
    //  out of line -- 24 % slower in round 2, 731 vs 588 us at 1080p: these walks are dominated by per-candidate
    //  steering that is uniform across the lanes, and two macroblocks per warp halve it per macroblock)
    case 16: return launch_fast<16, 16, PNORM>(a, n, stream);
    default: break;
    }
    if (a.sums) return GME_ERR_UNSUPPORTED;              // channel sums are only produced by the tiled kernel
    const long nblocks = (long)a.R * a.C;
    const int warps = 8;
    dim3 grid((unsigned)((nblocks + warps - 1) / warps), 1, n);
    bbme_pattern_generic_kernel<PNORM><<<grid, warps * 32, 0, stream>>>(a, bs);
    note_launch();
    return check_launch("bbme_pattern_generic_kernel");
}

int launch_bbme_pattern(const uint8_t *prev, size_t prev_stride, const uint8_t *cur, size_t cur_stride, int n, int H,
                        int W, size_t pitch, int bs, int sw, int procedure, int pnorm, int32_t *field,
                        unsigned long long *sums, cudaStream_t stream)
{
    PatternArgs a{};
    a.prev = prev; a.prev_stride = prev_stride;
    a.cur = cur; a.cur_stride = cur_stride;
    a.H = H; a.W = W; a.pitch = pitch;
    a.R = H / bs; a.C = W / bs;
    a.sw = sw; a.procedure = procedure;
    a.field = field;
    a.sums = sums;
    a.key_scale = 16;
    if (a.R == 0 || a.C == 0 || n == 0) return GME_OK;
    return pnorm == GME_PNORM_MAE ? launch_pattern_pn<GME_PNORM_MAE>(a, n, bs, stream)
                                  : launch_pattern_pn<GME_PNORM_MSE>(a, n, bs, stream);
}

}  // namespace gme
