// gme_capi.cu -- the extern "C" boundary (include/gme_b200.h): validation, dispatch, the
// whole-pipeline entry point and the TMA tensor-map helper.  No torch types cross this file.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "gme_common.cuh"

namespace gme {

// kernels' host launchers (one per .cu)
int launch_bbme_pattern(const uint8_t *, size_t, const uint8_t *, size_t, int, int, int, size_t, int, int, int, int,
                        int32_t *, unsigned long long *, cudaStream_t);
int launch_bbme_exhaustive(const uint8_t *, size_t, const uint8_t *, size_t, int, int, int, size_t, int, int, int,
                           int32_t *, cudaStream_t);
int launch_sad_probe(int, int, int, uint32_t *, cudaStream_t);
int launch_pyr_down(const uint8_t *, size_t, size_t, uint8_t *, size_t, size_t, int, int, int, cudaStream_t);
int launch_first_params(const int32_t *, int, int, int, double *, cudaStream_t);
int launch_affine_fit(const int32_t *, int, int, int, int, int, double, int, int, double *, uint8_t *, int32_t *,
                      int16_t *, int32_t *, int, const long long *, long, cudaStream_t);
int launch_affine_field(const double *, int, int, int, int16_t *, cudaStream_t);
int launch_pipeline_fits(const int32_t *, int, int, int, int, uint8_t *, const int32_t *, int, int, int, int, uint8_t *,
                         int, double, double *, int32_t *, const long long *, long, int16_t *, cudaStream_t, size_t);
int launch_compensate(const uint8_t *, size_t, size_t, const void *, int, int, int, const uint8_t *, size_t, size_t,
                      uint8_t *, size_t, size_t, int, int, int, uint64_t *, cudaStream_t, uint8_t * = nullptr,
                      uint8_t * = nullptr, size_t = 0, size_t = 0, size_t = 1);
int launch_hier_merge(const void *, int, int, int, const int32_t *, int, int, int, double *, cudaStream_t);
int launch_sse(const uint8_t *, size_t, size_t, const uint8_t *, size_t, size_t, int, int, int, uint64_t *,
               cudaStream_t);

static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_last_cuda_error{0};

// ---- per-stage timing (bench only) ----------------------------------------------------
static std::mutex g_timing_mu;
static bool g_timing_on = false;
static std::vector<cudaEvent_t> g_timing_events;   // (GME_PIPELINE_STAGES + 1) events per recorded call

struct StageTimer {
    cudaStream_t st;
    bool on;
    std::vector<cudaEvent_t> ev;
    explicit StageTimer(cudaStream_t s) : st(s)
    {
        std::lock_guard<std::mutex> lk(g_timing_mu);
        on = g_timing_on;
    }
    void mark()
    {
        if (!on) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) { on = false; return; }
        cudaEventRecord(e, st);
        ev.push_back(e);
    }
    ~StageTimer()
    {
        if (ev.empty()) return;
        std::lock_guard<std::mutex> lk(g_timing_mu);
        if (ev.size() == GME_PIPELINE_STAGES + 1 && g_timing_on)
            g_timing_events.insert(g_timing_events.end(), ev.begin(), ev.end());
        else
            for (cudaEvent_t e : ev) cudaEventDestroy(e);
    }
};

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void ensure_dynamic_smem(const void *kernel, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> granted;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    size_t &have = granted[{kernel, dev}];
    if (bytes <= have) return;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess) have = bytes;
}

int check_launch(const char *what)
{
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return GME_OK;
    g_last_cuda_error.store((int)e);
    fprintf(stderr, "[gme_b200] %s: %s\n", what, cudaGetErrorString(e));
    return GME_ERR_CUDA;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time libcuda dependency.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

bool make_plane_tensor_map(CUtensorMap *map, const uint8_t *base, int n, int H, int W, size_t pitch,
                           size_t plane_stride, int box_w, int box_h)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    if ((reinterpret_cast<uintptr_t>(base) % 16) || (pitch % 16) || (n > 1 && plane_stride % 16)) return false;
    if (box_w > 256 || box_h > 256 || box_w % 16) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)(n > 1 ? plane_stride : pitch * (size_t)H)};
    if (strides[1] % 16) return false;
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// ---------------------------------------------------------------------------------------
// pipeline workspace layout
// ---------------------------------------------------------------------------------------
struct Level { int H, W; size_t pitch, plane; };

struct Layout {
    Level l1, l0;                 // half and quarter resolution
    size_t off_prev1, off_cur1, off_prev0, off_cur0;
    size_t off_dense, off_f1, off_f2, off_out1, off_out2, off_model, off_sums, off_total;
    int R0, C0, R1, C1, R2, C2;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static Layout make_layout(int n, int H, int W)
{
    Layout L;
    L.l1.H = (H + 1) / 2; L.l1.W = (W + 1) / 2;
    L.l0.H = (L.l1.H + 1) / 2; L.l0.W = (L.l1.W + 1) / 2;
    L.l1.pitch = align_up(L.l1.W, 16); L.l1.plane = align_up(L.l1.pitch * L.l1.H, 256);
    L.l0.pitch = align_up(L.l0.W, 16); L.l0.plane = align_up(L.l0.pitch * L.l0.H, 256);
    L.R0 = L.l0.H / 2; L.C0 = L.l0.W / 2;        // dense estimate: block_size 2 (motion.py:27-29)
    L.R1 = L.l1.H / 16; L.C1 = L.l1.W / 16;      // BBME_BLOCK_SIZE = 16 (motion.py:9)
    L.R2 = H / 16; L.C2 = W / 16;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o = align_up(o + bytes, 256); return at; };
    L.off_prev1 = take(L.l1.plane * n * 2);              // 2n planes: [prev | cur], or one run of n + d frames
    L.off_cur1 = L.off_prev1 + L.l1.plane * n;
    L.off_prev0 = take(L.l0.plane * n * 2);
    L.off_cur0 = L.off_prev0 + L.l0.plane * n;
    L.off_dense = take((size_t)n * L.R0 * L.C0 * 2 * sizeof(int32_t));
    L.off_f1 = take((size_t)n * L.R1 * L.C1 * 2 * sizeof(int32_t));
    L.off_f2 = take((size_t)n * L.R2 * L.C2 * 2 * sizeof(int32_t));
    L.off_out1 = take((size_t)n * L.R1 * L.C1);
    L.off_out2 = take((size_t)n * L.R2 * L.C2);
    L.off_model = take((size_t)n * L.R2 * L.C2 * 2 * sizeof(int16_t));
    L.off_sums = take((size_t)n * 2 * sizeof(unsigned long long));
    L.off_total = o;
    return L;
}

static int bbme_dispatch(const uint8_t *prev, size_t ps, const uint8_t *cur, size_t cs, int n, int H, int W,
                         size_t pitch, int bs, int sw, int procedure, int pnorm, int32_t *field, cudaStream_t st,
                         unsigned long long *sums = nullptr)
{
    if (!prev || !cur || !field) return GME_ERR_INVALID_ARGUMENT;
    if (n < 0 || H <= 0 || W <= 0 || bs <= 0) return GME_ERR_INVALID_ARGUMENT;
    if (procedure < 0 || procedure > 3 || pnorm < 0 || pnorm > 1) return GME_ERR_INVALID_ARGUMENT;
    if (pitch % 4 || ps % 4 || cs % 4 || pitch < (size_t)W) return GME_ERR_ALIGNMENT;
    if ((reinterpret_cast<uintptr_t>(prev) | reinterpret_cast<uintptr_t>(cur)) % 4) return GME_ERR_ALIGNMENT;
    if (bs > 255) return GME_ERR_UNSUPPORTED;
    if (procedure != GME_SEARCH_DIAMOND && sw < 0) return GME_ERR_INVALID_ARGUMENT;
    if (procedure == GME_SEARCH_DIAMOND && (H <= bs || W <= bs)) return GME_ERR_UNSUPPORTED;
    if (procedure == GME_SEARCH_EXHAUSTIVE)
        return launch_bbme_exhaustive(prev, ps, cur, cs, n, H, W, pitch, bs, sw, pnorm, field, st);
    return launch_bbme_pattern(prev, ps, cur, cs, n, H, W, pitch, bs, sw, procedure, pnorm, field, sums, st);
}

}  // namespace gme

using namespace gme;

extern "C" {

int gme_version(void) { return GME_ABI_VERSION; }

const char *gme_error_string(int code)
{
    switch (code) {
    case GME_OK: return "ok";
    case GME_ERR_INVALID_ARGUMENT: return "invalid argument";
    case GME_ERR_UNSUPPORTED: return "unsupported geometry (undefined in the reference)";
    case GME_ERR_ALIGNMENT: return "pitch / base pointer not 4-byte aligned";
    case GME_ERR_CUDA: return "CUDA error";
    case GME_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown error";
    }
}

int gme_last_cuda_error(void) { return g_last_cuda_error.load(); }

uint64_t gme_launch_count(void) { return g_launches.load(); }

int gme_sad_peak_probe(int pnorm, int ctas, int iters, uint32_t *scratch, uint64_t *pixel_pairs, void *stream)
{
    if (!scratch || !pixel_pairs || ctas <= 0 || iters <= 0 || pnorm < 0 || pnorm > 1) return GME_ERR_INVALID_ARGUMENT;
    *pixel_pairs = (uint64_t)ctas * 256u * (uint64_t)iters * 4u * 8u * 4u;   // threads x rounds x 8 updates x 4 pixels
    return launch_sad_probe(pnorm, ctas, iters, scratch, static_cast<cudaStream_t>(stream));
}

int gme_stage_timing_enable(int enable)
{
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_timing_on = enable != 0;
    for (cudaEvent_t e : g_timing_events) cudaEventDestroy(e);
    g_timing_events.clear();
    return GME_OK;
}

int gme_stage_timing_read(double *ms_sum, int *calls)
{
    if (!ms_sum || !calls) return GME_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    for (int s = 0; s < GME_PIPELINE_STAGES; s++) ms_sum[s] = 0.0;
    const size_t per = GME_PIPELINE_STAGES + 1;
    *calls = (int)(g_timing_events.size() / per);
    int rc = GME_OK;
    if (!g_timing_events.empty() && cudaEventSynchronize(g_timing_events.back()) != cudaSuccess) rc = check_launch("stage timing");
    for (size_t c = 0; rc == GME_OK && c < (size_t)*calls; c++)
        for (int s = 0; s < GME_PIPELINE_STAGES; s++) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, g_timing_events[c * per + s], g_timing_events[c * per + s + 1]) != cudaSuccess) {
                rc = check_launch("stage timing");
                break;
            }
            ms_sum[s] += ms;
        }
    for (cudaEvent_t e : g_timing_events) cudaEventDestroy(e);
    g_timing_events.clear();
    return rc;
}

int gme_bbme_motion_field(const uint8_t *prev, size_t prev_plane_stride, const uint8_t *cur, size_t cur_plane_stride,
                          int n, int H, int W, size_t pitch, int block_size, int search_window, int procedure,
                          int pnorm, int32_t *field, void *stream)
{
    return bbme_dispatch(prev, prev_plane_stride, cur, cur_plane_stride, n, H, W, pitch, block_size, search_window,
                         procedure, pnorm, field, static_cast<cudaStream_t>(stream));
}

int gme_pyr_down(const uint8_t *src, size_t src_pitch, size_t src_plane_stride, uint8_t *dst, size_t dst_pitch,
                 size_t dst_plane_stride, int n, int H, int W, void *stream)
{
    if (!src || !dst || n < 0 || H <= 0 || W <= 0) return GME_ERR_INVALID_ARGUMENT;
    if (src_pitch < (size_t)W || dst_pitch < (size_t)((W + 1) / 2)) return GME_ERR_INVALID_ARGUMENT;
    if (n == 0) return GME_OK;
    return launch_pyr_down(src, src_pitch, src_plane_stride, dst, dst_pitch, dst_plane_stride, n, H, W,
                           static_cast<cudaStream_t>(stream));
}

int gme_first_parameters(const int32_t *dense_field, int n, int R, int C, double *params, void *stream)
{
    if (!dense_field || !params || n < 0 || R <= 0 || C <= 0) return GME_ERR_INVALID_ARGUMENT;
    if (n == 0) return GME_OK;
    return launch_first_params(dense_field, n, R, C, params, static_cast<cudaStream_t>(stream));
}

int gme_affine_fit(const int32_t *gt_field, int n, int R, int C, int level_h, int level_w, double pct, int robust,
                   int project, double *params, uint8_t *outlier, int32_t *threshold, int16_t *model_field,
                   int32_t *status, void *stream)
{
    if (!gt_field || !params || n < 0 || level_h <= 0 || level_w <= 0) return GME_ERR_INVALID_ARGUMENT;
    if (!(pct >= 0.0 && pct <= 1.0)) return GME_ERR_INVALID_ARGUMENT;   // also rejects NaN: int(pct * N) must lie in [0, N]
    if (R <= 0 || C <= 0) return GME_ERR_UNSUPPORTED;   // empty field: the reference indexes an empty array
    if ((long long)R * C > 0x7FFFFFFFLL) return GME_ERR_UNSUPPORTED;
    if (n == 0) return GME_OK;
    return launch_affine_fit(gt_field, n, R, C, level_h, level_w, pct, robust, project, params, outlier, threshold,
                             model_field, status, 0, nullptr, 0, static_cast<cudaStream_t>(stream));
}

int gme_affine_field(const double *params, int n, int R, int C, int16_t *field, void *stream)
{
    if (!params || !field || n < 0 || R < 0 || C < 0) return GME_ERR_INVALID_ARGUMENT;
    if (n == 0 || R == 0 || C == 0) return GME_OK;
    return launch_affine_field(params, n, R, C, field, static_cast<cudaStream_t>(stream));
}

int gme_compensate(const uint8_t *frame, size_t frame_pitch, size_t frame_plane_stride, const void *field,
                   int field_is_i16, int R, int C, const uint8_t *cur, size_t cur_pitch, size_t cur_plane_stride,
                   uint8_t *comp, size_t comp_pitch, size_t comp_plane_stride, int n, int H, int W, uint64_t *sse,
                   void *stream)
{
    if (!frame || !comp || n < 0 || H <= 0 || W <= 0 || R < 0 || C < 0) return GME_ERR_INVALID_ARGUMENT;
    if (!field && R * C > 0) return GME_ERR_INVALID_ARGUMENT;
    if ((cur == nullptr) != (sse == nullptr)) return GME_ERR_INVALID_ARGUMENT;
    if (frame_pitch % 4 || (reinterpret_cast<uintptr_t>(frame) % 4) || frame_plane_stride % 4) return GME_ERR_ALIGNMENT;
    if (n == 0) return GME_OK;
    return launch_compensate(frame, frame_pitch, frame_plane_stride, field, field_is_i16, R, C, cur, cur_pitch,
                             cur_plane_stride, comp, comp_pitch, comp_plane_stride, n, H, W, sse,
                             static_cast<cudaStream_t>(stream));
}

int gme_compensate_diffs(const uint8_t *frame, size_t frame_pitch, size_t frame_plane_stride, const void *field,
                         int field_is_i16, int R, int C, const uint8_t *cur, size_t cur_pitch, size_t cur_plane_stride,
                         uint8_t *comp, size_t comp_pitch, size_t comp_plane_stride, uint8_t *diff_prev, uint8_t *diff_comp,
                         size_t diff_pitch, size_t diff_plane_stride, int n, int H, int W, uint64_t *sse, void *stream)
{
    if (!frame || !comp || !cur || !sse || !diff_prev || !diff_comp || n < 0 || H <= 0 || W <= 0 || R < 0 || C < 0)
        return GME_ERR_INVALID_ARGUMENT;
    if (!field && R * C > 0) return GME_ERR_INVALID_ARGUMENT;
    if (diff_pitch < (size_t)W) return GME_ERR_INVALID_ARGUMENT;
    if (frame_pitch % 4 || (reinterpret_cast<uintptr_t>(frame) % 4) || frame_plane_stride % 4) return GME_ERR_ALIGNMENT;
    if (n == 0) return GME_OK;
    return launch_compensate(frame, frame_pitch, frame_plane_stride, field, field_is_i16, R, C, cur, cur_pitch,
                             cur_plane_stride, comp, comp_pitch, comp_plane_stride, n, H, W, sse,
                             static_cast<cudaStream_t>(stream), diff_prev, diff_comp, diff_pitch, diff_plane_stride);
}

int gme_hier_merge(const void *coarse, int coarse_is_f64, int Rc, int Cc, const int32_t *fine, int R, int C, int n,
                   double *out, void *stream)
{
    if (!coarse || !fine || !out || n < 0 || Rc <= 0 || Cc <= 0 || R <= 0 || C <= 0) return GME_ERR_INVALID_ARGUMENT;
    // bbme.py:596-602: the upsampled field gets ONE zero row, or else ONE zero column, when the shapes differ; anything
    // else does not broadcast in the reference (it raises), so it is not a geometry this entry point accepts
    const bool same = (2 * Rc == R && 2 * Cc == C), row = (2 * Rc + 1 == R && 2 * Cc == C), col = (2 * Rc == R && 2 * Cc + 1 == C);
    if (!(same || row || col)) return GME_ERR_UNSUPPORTED;
    if (n == 0) return GME_OK;
    return launch_hier_merge(coarse, coarse_is_f64, Rc, Cc, fine, R, C, n, out, static_cast<cudaStream_t>(stream));
}

int gme_sse(const uint8_t *a, size_t a_pitch, size_t a_plane_stride, const uint8_t *b, size_t b_pitch,
            size_t b_plane_stride, int n, int H, int W, uint64_t *sse, void *stream)
{
    if (!a || !b || !sse || n < 0 || H <= 0 || W <= 0) return GME_ERR_INVALID_ARGUMENT;
    if (n == 0) return GME_OK;
    return launch_sse(a, a_pitch, a_plane_stride, b, b_pitch, b_plane_stride, n, H, W, sse,
                      static_cast<cudaStream_t>(stream));
}

size_t gme_pipeline_workspace_bytes(int n, int H, int W)
{
    if (n <= 0 || H <= 0 || W <= 0) return 0;
    return make_layout(n, H, W).off_total;
}

void *gme_pipeline_workspace_ptr(void *workspace, int n, int H, int W, int which)
{
    if (!workspace || n <= 0 || H <= 0 || W <= 0) return nullptr;
    const Layout L = make_layout(n, H, W);
    uint8_t *ws = static_cast<uint8_t *>(workspace);
    switch (which) {
    case 0: return ws + L.off_dense;
    case 1: return ws + L.off_f1;
    case 2: return ws + L.off_f2;
    case 3: return ws + L.off_out1;
    case 4: return ws + L.off_out2;
    case 5: return ws + L.off_model;
    default: return nullptr;
    }
}

int gme_pipeline(const uint8_t *prev, size_t prev_plane_stride, const uint8_t *cur, size_t cur_plane_stride, int n,
                 int H, int W, size_t pitch, int procedure, int search_window, double outlier_fraction, double *params,
                 size_t params_stride, uint8_t *comp, size_t comp_pitch, size_t comp_plane_stride, uint64_t *sse,
                 size_t sse_stride, int32_t *status, void *workspace, size_t workspace_bytes, void *stream)
{
    if (params_stride < 6 || (sse && sse_stride < 1)) return GME_ERR_INVALID_ARGUMENT;
    if (!prev || !cur || !params || !workspace || n < 0 || H <= 0 || W <= 0) return GME_ERR_INVALID_ARGUMENT;
    if (!(outlier_fraction >= 0.0 && outlier_fraction <= 1.0)) return GME_ERR_INVALID_ARGUMENT;
    if (sse && !comp) return GME_ERR_INVALID_ARGUMENT;
    if (n == 0) return GME_OK;
    const Layout L = make_layout(n, H, W);
    if (workspace_bytes < L.off_total) return GME_ERR_WORKSPACE;
    // every level needs at least one 16x16 block and a diamond-searchable quarter-resolution frame
    if (L.R1 < 1 || L.C1 < 1 || L.l1.H <= 16 || L.l1.W <= 16 || L.l0.H <= 2 || L.l0.W <= 2) return GME_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t *ws = static_cast<uint8_t *>(workspace);
    uint8_t *prev1 = ws + L.off_prev1, *cur1 = ws + L.off_cur1, *prev0 = ws + L.off_prev0, *cur0 = ws + L.off_cur0;
    int32_t *dense = reinterpret_cast<int32_t *>(ws + L.off_dense);
    int32_t *f1 = reinterpret_cast<int32_t *>(ws + L.off_f1), *f2 = reinterpret_cast<int32_t *>(ws + L.off_f2);
    uint8_t *out1 = ws + L.off_out1, *out2 = ws + L.off_out2;
    int16_t *model = reinterpret_cast<int16_t *>(ws + L.off_model);
    int rc;
    StageTimer timer(st);
    timer.mark();
#define GME_TRY(x) do { rc = (x); if (rc != GME_OK) return rc; } while (0)
    // utils.get_pyramids for both frames (motion.py:123-124).  When prev and cur are two views of one
    // sequence buffer (cur = prev + d planes, d <= n) every frame's pyramid is built once for all the pairs
    // it belongs to, instead of once as "previous" and once as "current" -- same bytes, half the traffic.
    const ptrdiff_t gap = cur - prev;
    const bool sequence = cur_plane_stride == prev_plane_stride && prev_plane_stride > 0 && gap > 0 &&
                          gap % (ptrdiff_t)prev_plane_stride == 0 && gap / (ptrdiff_t)prev_plane_stride <= n;
    if (sequence) {
        const int d = (int)(gap / (ptrdiff_t)prev_plane_stride);
        cur1 = prev1 + (size_t)d * L.l1.plane;
        cur0 = prev0 + (size_t)d * L.l0.plane;
        GME_TRY(launch_pyr_down(prev, pitch, prev_plane_stride, prev1, L.l1.pitch, L.l1.plane, n + d, H, W, st));
        GME_TRY(launch_pyr_down(prev1, L.l1.pitch, L.l1.plane, prev0, L.l0.pitch, L.l0.plane, n + d, L.l1.H, L.l1.W, st));
    } else {
        GME_TRY(launch_pyr_down(prev, pitch, prev_plane_stride, prev1, L.l1.pitch, L.l1.plane, n, H, W, st));
        GME_TRY(launch_pyr_down(cur, pitch, cur_plane_stride, cur1, L.l1.pitch, L.l1.plane, n, H, W, st));
        GME_TRY(launch_pyr_down(prev1, L.l1.pitch, L.l1.plane, prev0, L.l0.pitch, L.l0.plane, n, L.l1.H, L.l1.W, st));
        GME_TRY(launch_pyr_down(cur1, L.l1.pitch, L.l1.plane, cur0, L.l0.pitch, L.l0.plane, n, L.l1.H, L.l1.W, st));
    }
    timer.mark();
    // the three block-matching passes are independent of the parameters: dense L0 (motion.py:27-29),
    // then the block_size-16 fields of L1 and L2 (motion.py:224-229)
    unsigned long long *sums = reinterpret_cast<unsigned long long *>(ws + L.off_sums);
    if (cudaMemsetAsync(sums, 0, sizeof(unsigned long long) * 2 * n, st) != cudaSuccess) return check_launch("memset");
    GME_TRY(bbme_dispatch(prev0, L.l0.plane, cur0, L.l0.plane, n, L.l0.H, L.l0.W, L.l0.pitch, 2, 2, GME_SEARCH_DIAMOND,
                          GME_PNORM_MSE, dense, st, sums));
    timer.mark();
    GME_TRY(bbme_dispatch(prev1, L.l1.plane, cur1, L.l1.plane, n, L.l1.H, L.l1.W, L.l1.pitch, 16, search_window,
                          procedure, GME_PNORM_MSE, f1, st));
    timer.mark();
    GME_TRY(bbme_dispatch(prev, prev_plane_stride, cur, cur_plane_stride, n, H, W, pitch, 16, search_window, procedure,
                          GME_PNORM_MSE, f2, st));
    timer.mark();
    if (status && cudaMemsetAsync(status, 0, sizeof(int32_t) * n, st) != cudaSuccess) return check_launch("memset");
    // the sequential part: first estimate, then project + robust fit per level (motion.py:128-134)
    // one launch: first estimate (mean of the dense field, from the channel sums) -> project + robust fit on L1 ->
    // project + robust fit on L2 -> model field of the final parameters at block_size 16 (results.py:52-54)
    GME_TRY(launch_pipeline_fits(f1, L.R1, L.C1, L.l1.H, L.l1.W, out1, f2, L.R2, L.C2, H, W, out2, n, outlier_fraction, params, status,
                                 reinterpret_cast<const long long *>(sums), (long)L.R0 * L.C0, comp ? model : nullptr, st,
                                 params_stride));
    timer.mark();
    if (comp) {
        // motion.compensate_frame of previous + the squared error against current (results.py:59,109)
        GME_TRY(launch_compensate(prev, pitch, prev_plane_stride, model, 1, L.R2, L.C2, sse ? cur : nullptr, pitch,
                                  cur_plane_stride, comp, comp_pitch, comp_plane_stride, n, H, W, sse, st, nullptr, nullptr, 0,
                                  0, sse ? sse_stride : 1));
    }
    timer.mark();
#undef GME_TRY
    return GME_OK;
}

}  // extern "C"
