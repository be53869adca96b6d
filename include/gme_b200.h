/*
 * gme_b200.h -- C ABI of the B200-native global-motion-estimation hot path.
 *
 * The reference (Samaretas/global-motion-estimation) is pure Python and has no FFI of its
 * own: its boundary is the module surface of bbme.py / motion.py / utils.py.  These entry
 * points are what a binding for that surface calls (ctypes stub: INTEGRATION.md); each one
 * names the reference function it replaces (paths relative to
 * /root/reference/global_motion_estimation/).
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer owned by the caller (no ownership transfer, no
 *    allocation inside the library).  `stream` is a cudaStream_t passed as void*; all work
 *    is asynchronous on it.  Thread-safe for distinct streams/buffers: no entry point keeps
 *    state between calls (the launch counter and the bench-only stage timer are the only
 *    globals, both lock-protected).
 *  - A "plane set" is n grayscale uint8 planes of H rows x W columns: plane k starts at
 *    base + k*plane_stride, row r at + r*pitch (bytes).  pitch % 4 == 0 is required;
 *    pitch % 16 == 0 with 16-byte aligned bases lets the search windows travel by TMA
 *    (otherwise a cooperative-load path is used; results are identical).
 *    prev and cur may alias one sequence buffer (cur = prev + d*plane_stride).
 *    Every plane must be backed by H*pitch readable bytes: the kernels read whole aligned
 *    words, so the padding columns of the LAST row (bytes W..pitch-1) are touched too.
 *  - Motion fields are int32[n][R][C][2] with R = H/bs, C = W/bs; channel 0 = column
 *    (horizontal) displacement, channel 1 = row (vertical) displacement (bbme.py:176-177).
 *  - Return value: GME_OK, or a negative GME_ERR_* (never throws, never exits).
 *    Kernel launch errors are returned as GME_ERR_CUDA; gme_last_cuda_error() has the code.
 */
#ifndef GME_B200_H
#define GME_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GME_ABI_VERSION 2   /* 2: gme_pipeline takes the outlier fraction and output strides; gme_pipeline_set_outlier_fraction is gone */

#if defined(__GNUC__)
#define GME_API __attribute__((visibility("default")))
#else
#define GME_API
#endif

enum {
    GME_OK = 0,
    GME_ERR_INVALID_ARGUMENT = -1, /* null pointer, non-positive size, bad enum            */
    GME_ERR_UNSUPPORTED = -2,      /* geometry the reference leaves undefined (see each fn) */
    GME_ERR_ALIGNMENT = -3,        /* pitch % 4 != 0                                        */
    GME_ERR_CUDA = -4,             /* a CUDA runtime/driver call failed                     */
    GME_ERR_WORKSPACE = -5         /* workspace too small                                   */
};

/* bbme.searching_procedures (bbme.py:609-614) and bbme.pnorm_distances (bbme.py:608) */
enum { GME_SEARCH_EXHAUSTIVE = 0, GME_SEARCH_THREESTEP = 1, GME_SEARCH_TWODLOG = 2, GME_SEARCH_DIAMOND = 3 };
enum { GME_PNORM_MAE = 0, GME_PNORM_MSE = 1 };

GME_API int gme_version(void);
GME_API const char *gme_error_string(int code);
GME_API int gme_last_cuda_error(void);

/* bbme.get_motion_field (bbme.py:12-38) + the four *_search procedures (bbme.py:105-534),
 * batched over n frame pairs.  Anchor blocks come from prev, candidates from cur.
 * Costs are exact integers (SAD / SSD), which is what the reference's float32 sums are for
 * MAE with block_size <= 255 and MSE with block_size <= 16; MSE with a larger
 * block rounds in the reference (NumPy's pairwise float32 sum) and the kernels reproduce that
 * rounding node by node, so the fields are bit-identical there too.
 * GME_ERR_UNSUPPORTED: block_size > 255; diamond with H <= bs or W <= bs (bbme.py:503-504
 * clamps to a negative bound there). */
GME_API int gme_bbme_motion_field(const uint8_t *prev, size_t prev_plane_stride,
                          const uint8_t *cur, size_t cur_plane_stride,
                          int n, int H, int W, size_t pitch,
                          int block_size, int search_window, int procedure, int pnorm,
                          int32_t *field, void *stream);

/* cv2.pyrDown as used by utils.get_pyramids (utils.py:34-51): one level, n planes.
 * dst planes are ((H+1)/2) x ((W+1)/2). */
GME_API int gme_pyr_down(const uint8_t *src, size_t src_pitch, size_t src_plane_stride,
                 uint8_t *dst, size_t dst_pitch, size_t dst_plane_stride,
                 int n, int H, int W, void *stream);

/* motion.compute_first_parameters (motion.py:176-188): params[k] = [mean ch0, 0, 0, mean ch1, 0, 0]
 * rounded through float32, stored as float64[n][6]. */
GME_API int gme_first_parameters(const int32_t *dense_field, int n, int R, int C, double *params, void *stream);

/* motion.best_affine_parameters_robust (motion.py:210-286) minus its BBME call, plus
 * motion.parameter_projection (motion.py:191-207) when project != 0:
 *   params (float64[n][6], in/out): old parameters in, new parameters out;
 *   pct = motion.MOTION_VECTOR_ERROR_THRESHOLD_PERCENTAGE, in [0, 1] (GME_ERR_INVALID_ARGUMENT otherwise);
 *   robust = 0 gives motion.best_affine_parameters (motion.py:33-88, no outlier mask);
 *   level_h/level_w: shape of the frame the field was estimated on (w = 1/(h*w));
 *   outlier (uint8[n][R][C]), threshold (int32[n]), model_field (int16[n][R][C][2]) are
 *   optional outputs (may be NULL); status (int32[n], may be NULL): 0 ok, 1 = singular
 *   normal matrix (the reference raises numpy.linalg.LinAlgError there). */
GME_API int gme_affine_fit(const int32_t *gt_field, int n, int R, int C, int level_h, int level_w,
                   double pct, int robust, int project, double *params,
                   uint8_t *outlier, int32_t *threshold, int16_t *model_field, int32_t *status,
                   void *stream);

/* motion.get_motion_field_affine (motion.py:139-157): int16[n][R][C][2]. */
GME_API int gme_affine_field(const double *params, int n, int R, int C, int16_t *field, void *stream);

/* motion.compensate_frame (motion.py:289-321), fused with the squared-error sum of
 * utils.PSNR (utils.py:100-116) against `cur` when cur != NULL:
 *   field is int16 or int32 [n][R][C][2] (field_is_i16 selects);
 *   comp receives the compensated planes; sse (uint64[n]) the sum of (cur-comp)^2. */
GME_API int gme_compensate(const uint8_t *frame, size_t frame_pitch, size_t frame_plane_stride,
                   const void *field, int field_is_i16, int R, int C,
                   const uint8_t *cur, size_t cur_pitch, size_t cur_plane_stride,
                   uint8_t *comp, size_t comp_pitch, size_t comp_plane_stride,
                   int n, int H, int W, uint64_t *sse, void *stream);

/* gme_compensate plus the two difference images results.py writes per pair (results.py:78-83):
 * diff_prev = |cur - frame|, diff_comp = |cur - comp| (uint8 planes sharing diff_pitch / diff_plane_stride).
 * cur and sse are required here. */
GME_API int gme_compensate_diffs(const uint8_t *frame, size_t frame_pitch, size_t frame_plane_stride,
                         const void *field, int field_is_i16, int R, int C,
                         const uint8_t *cur, size_t cur_pitch, size_t cur_plane_stride,
                         uint8_t *comp, size_t comp_pitch, size_t comp_plane_stride,
                         uint8_t *diff_prev, uint8_t *diff_comp, size_t diff_pitch, size_t diff_plane_stride,
                         int n, int H, int W, uint64_t *sse, void *stream);

/* The merge step of bbme.hierarchical_wrapper (bbme.py:587-604, with rescale_motion_field, bbme.py:537-546):
 * out[n][R][C][2] (float64) = (2 * trunc(coarse[i/2][j/2]) + fine[i][j]) / 2, the upsampled coarse field padded with one
 * zero row or one zero column where the fine field has one more.  coarse is int32 or float64 [n][Rc][Cc][2].
 * GME_ERR_UNSUPPORTED for shape pairs the reference itself cannot merge (it raises there). */
GME_API int gme_hier_merge(const void *coarse, int coarse_is_f64, int Rc, int Cc,
                   const int32_t *fine, int R, int C, int n, double *out, void *stream);

/* utils.PSNR (utils.py:100-116): sse[k] = sum (a-b)^2 over plane k; the caller finishes
 * mse = sse/(H*W), 20*log10(255/sqrt(mse)). */
GME_API int gme_sse(const uint8_t *a, size_t a_pitch, size_t a_plane_stride,
            const uint8_t *b, size_t b_pitch, size_t b_plane_stride,
            int n, int H, int W, uint64_t *sse, void *stream);

/* motion.global_motion_estimation + get_motion_field_affine + compensate_frame + PSNR
 * (motion.py:109-136, results.py:50-59,109) for n frame pairs in one call.
 * procedure/search_window apply to the block_size-16 levels only; the dense first
 * estimate is always diamond with block_size 2 and the cost is always MSE, as in the
 * reference (motion.py:27-29, 224-229).  Reference behaviour = (GME_SEARCH_DIAMOND, 2).
 * outlier_fraction = motion.MOTION_VECTOR_ERROR_THRESHOLD_PERCENTAGE (motion.py:10; reference: 0.3), the share of
 * blocks the robust fits may flag; the reference keeps it as a module constant that users edit (README:137-141),
 * the drop-in passes its current value with every call.  GME_ERR_INVALID_ARGUMENT outside [0, 1].
 * Outputs: params float64[n][6], pair k at params + k*params_stride doubles (params_stride >= 6); comp (optional)
 * + sse (optional, needs comp), pair k at sse + k*sse_stride uint64 words; status int32[n].  The strides let the
 * caller interleave both into ONE row buffer ([n][7] 8-byte words: six parameters + the squared error), which is what
 * the multi-GPU path gathers -- no packing kernel between the pipeline and the collective.  Workspace layout is private; size it with gme_pipeline_workspace_bytes. */
GME_API size_t gme_pipeline_workspace_bytes(int n, int H, int W);
GME_API int gme_pipeline(const uint8_t *prev, size_t prev_plane_stride,
                 const uint8_t *cur, size_t cur_plane_stride,
                 int n, int H, int W, size_t pitch,
                 int procedure, int search_window, double outlier_fraction,
                 double *params, size_t params_stride,
                 uint8_t *comp, size_t comp_pitch, size_t comp_plane_stride,
                 uint64_t *sse, size_t sse_stride, int32_t *status,
                 void *workspace, size_t workspace_bytes, void *stream);

/* Introspection for tests and the bench: pointers into a gme_pipeline workspace.
 * which: 0 dense L0 field, 1 L1 field, 2 L2 field (int32), 3 L1 outlier mask, 4 L2 outlier
 * mask (uint8), 5 model field at full resolution (int16). */
GME_API void *gme_pipeline_workspace_ptr(void *workspace, int n, int H, int W, int which);

/* Per-stage device timing of gme_pipeline for the bench (roofline of the dominant kernel is measured
 * live, inside the timed region).  While enabled, every gme_pipeline call records cudaEvents on its
 * stream around its stages: 0 pyramids (2 launches for a sequence, 4 otherwise), 1 dense L0 BBME, 2 L1 BBME,
 * 3 L2 BBME, 4 first estimate + both robust fits + model field (1 launch), 5 compensation + squared error.
 * gme_stage_timing_read synchronises on the last recorded event, adds up the elapsed milliseconds per
 * stage over all calls since the last read into ms_sum[GME_PIPELINE_STAGES], stores the number of calls
 * and clears the record.  Not usable while the stream is being captured into a CUDA graph. */
#define GME_PIPELINE_STAGES 6
GME_API int gme_stage_timing_enable(int enable);
GME_API int gme_stage_timing_read(double *ms_sum, int *calls);

/* Roofline denominator of the exhaustive search (bench only): runs the search's packed cost instruction mix
 * (pnorm 0: VABSDIFF4.ACC, 1: VABSDIFF4 + IDP.4A) on registers, ctas x 256 threads x iters rounds, and stores
 * the number of pixel-pair updates issued in *pixel_pairs (host).  scratch: >= 1024 uint32 on the device.
 * Time it with events: pixel_pairs / elapsed = sustained integer-pipe peak for that norm. */
GME_API int gme_sad_peak_probe(int pnorm, int ctas, int iters, uint32_t *scratch, uint64_t *pixel_pairs, void *stream);

/* Number of kernel launches issued through this library since load (for bench accounting). */
GME_API uint64_t gme_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GME_B200_H */
