/*
 * fastbarnes_b200.h -- C ABI of the B200-native optimized-convolution Barnes interpolation.
 *
 * Drop-in boundary for ONE path of MeteoSwiss/fast-barnes-py (v2.0.0):
 *   fastbarnes.interpolation.barnes(..., method='optimized_convolution' | 'convolution')
 *   fastbarnes.interpolationS2.barnes_S2(..., method='optimized_convolution_S2')
 * The reference has no FFI of its own (it is Python + Numba @njit); its boundary is the
 * Python call surface.  Each entry point below cites the reference function it replaces
 * (paths relative to the reference root).  The Python side that binds these symbols with
 * ctypes lives in fast-barnes-py_b200/fastbarnes/ (see INTEGRATION.md).
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in the signatures (a CUDA stream is passed
 *     as void*; NULL = the legacy default stream).
 *   - every function returns FB_OK (0) or a negative FB_E* code; fb_last_error() returns a
 *     thread-local message for the last failure.
 *   - `*_host` entry points take HOST pointers and are synchronous (H2D, kernels, D2H inside);
 *     `*_dev` entry points take DEVICE pointers plus a caller-provided workspace and only
 *     enqueue work on the stream.
 *   - grids are returned with reversed dimensions like the reference: [x], [y][x], [z][y][x],
 *     float32, NaN where the convolved weight is below the max_dist threshold.
 *   - fields: a call may carry `nfields` independent fields (time steps / ensemble members)
 *     on the same grid.  Samples of all fields are concatenated; field b owns the samples
 *     [sample_offsets[b], sample_offsets[b+1]).  nfields == 1 is the reference call.
 *   - there is no CPU fallback: without a CUDA device every compute entry returns FB_ECUDA.
 */
#ifndef FASTBARNES_B200_H
#define FASTBARNES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB_OK        0
#define FB_EINVAL   -1   /* invalid argument */
#define FB_ECUDA    -2   /* CUDA runtime error / no device */
#define FB_ENOMEM   -3   /* workspace too small / allocation failed */
#define FB_EKERNEL  -4   /* rectangular kernel does not fit (grid or on-chip ring storage) */

/* fb_problem.flags */
#define FB_FLAG_SEGMENTED_1D 1   /* dim 1, nfields 1: cut the single grid line into overlapping segments that are
                                    swept in parallel (exact in exact arithmetic, rounding-level difference to the
                                    reference's single accumulator chain); default is the bit-exact sequential walk */

#define FB_FLAG_FP32 2            /* dim 2 or 3: fp32 working precision for the sweeps (grid as interleaved float2 (value,
                                    weight) nodes, packed f32x2 arithmetic with Kahan-compensated accumulators, half the
                                    HBM traffic).  Injection stays fp64.  Not bit-identical to the reference: |field -
                                    reference| stays within a few float32 ulps of the field values except next to the
                                    max_dist boundary (tolerance stated in tests/test_gpu_parity.py); needs
                                    3 <= T and the rings of all passes on chip (FB_EKERNEL otherwise); no out64 */

#define FB_METHOD_OPTIMIZED_CONVOLUTION 0   /* interpolation.py:169-176 */
#define FB_METHOD_CONVOLUTION           1   /* interpolation.py:178-185 */
/* exact Gaussian sums (fb_barnes_exact_*), the reference's accuracy yardsticks */
#define FB_METHOD_NAIVE                 2   /* interpolation.py:195-196, :862-938 */
#define FB_METHOD_RADIUS                3   /* interpolation.py:187-193, :809-855 */
#define FB_METHOD_NAIVE_S2              4   /* interpolationS2.py:134-135, :260-301 */

/* Problem description shared by the interpolation entry points
 * (the arguments of _interpolate_opt_convol, interpolation.py:329). */
typedef struct fb_problem {
    int32_t dim;             /* 1, 2 or 3 */
    int32_t method;          /* FB_METHOD_* */
    int32_t num_iter;        /* number of self-convolutions n */
    int32_t flags;           /* FB_FLAG_* */
    int64_t nfields;         /* independent fields on the same grid (>= 1) */
    int64_t size[3];         /* grid extension (x, y, z) */
    double  sigma[3];        /* Gaussian width per axis */
    double  x0[3];           /* grid start point */
    double  step[3];         /* grid step per axis */
    double  max_dist_weight; /* exp(-max_dist^2/2), interpolation.py:167 */
} fb_problem;

/* ---- library / device -------------------------------------------------------------------- */
const char *fb_last_error(void);
int  fb_version(void);
/* number of visible CUDA devices (0 if none / driver missing) */
int  fb_device_count(void);
/* selects the CUDA device used by the calling thread's subsequent calls */
int  fb_set_device(int device);

/* ---- kernel parameters (host arithmetic, bit-identical to the Numba expressions) --------- */
/* interpolation.py:549-552  _get_half_kernel_size_opt */
int32_t fb_half_kernel_size_opt(double sigma, double step, int num_iter);
/* interpolation.py:783-785  _get_half_kernel_size */
int32_t fb_half_kernel_size(double sigma, double step, int num_iter);
/* interpolation.py:561-569  _get_tail_value */
double  fb_tail_value(double sigma, double step, int num_iter);
/* interpolation.py:424-425 (== :389-390, :472-473, and :634-635 with tail 0) conv_scale_factor */
double  fb_conv_scale_factor(int dim, const int32_t *kernel_size, const double *tail_value,
                             const double *sigma, const double *step, int num_iter,
                             double max_dist_weight);

/* ---- whole path, HOST buffers ------------------------------------------------------------- */
/* interpolation.py:329-367 _interpolate_opt_convol (method 0) / :575-612 _interpolate_convol
 * (method 1) for prob->nfields fields.
 *   pts  [nsamples][dim] float64, val [nsamples] float64 (not modified),
 *   sample_offsets [nfields+1] or NULL (NULL: equal split, nsamples % nfields == 0),
 *   out  [nfields][prod(size)] float32, index order z,y,x,
 *   out64 optional (may be NULL): the pre-cast fp64 quotient vg/wg+offset, same layout.  */
int fb_barnes_host(const fb_problem *prob, int64_t nsamples, const int64_t *sample_offsets,
                   const double *pts, const double *val, float *out, double *out64);

/* ---- whole path, DEVICE buffers ----------------------------------------------------------- */
/* bytes of scratch fb_barnes_dev needs for this problem and sample count */
int64_t fb_workspace_bytes(const fb_problem *prob, int64_t nsamples);
/* Same computation as fb_barnes_host with every pointer in device memory; sample_offsets is a
 * HOST pointer (or NULL).  Enqueues on `stream` and returns without synchronising.          */
int fb_barnes_dev(const fb_problem *prob, int64_t nsamples, const int64_t *sample_offsets,
                  const double *d_pts, const double *d_val, float *d_out, double *d_out64,
                  void *d_workspace, int64_t workspace_bytes, void *stream);

/* ---- 3D z-slab decomposition (multi-GPU; no counterpart in the single-threaded reference) ---- */
/* Halo planes a slab needs on each interior side: num_iter * (T_z + 1). */
int64_t fb_slab_halo_planes(const fb_problem *prob);
/* 1: the slab calls keep the extended B buffer as ONE array of interleaved (value, weight) double2 nodes
 * [z_ext][H][W][2] starting at offset_vB (the second-generation sweep kernels); 0: two arrays of planes at offset_vB and
 * offset_wB.  The halo exchange moves whole planes of whichever form is in use. */
int fb_slab_interleaved(const fb_problem *prob);
/* Workspace size of a slab of z_count own planes with halo_lo / halo_hi halo planes, and the byte
 * offsets of its extended B buffers (values, weights; [z_ext][y][x] float64) inside the workspace:
 * the caller writes the neighbours' planes into the halo parts between phase 1 and phase 2. */
int fb_slab_layout(const fb_problem *prob, int64_t nsamples, int64_t z_count, int64_t halo_lo, int64_t halo_hi,
                   int want_out64, int64_t *workspace_bytes, int64_t *offset_vB, int64_t *offset_wB);
/* Phase 1: centre (over ALL samples) + inject + x sweep + y sweep for the own planes
 * [z_begin, z_begin + z_count) of the volume described by prob (dim 3, nfields 1). */
int fb_slab_phase1_dev(const fb_problem *prob, int64_t z_begin, int64_t z_count, int64_t halo_lo, int64_t halo_hi,
                       int64_t nsamples, const double *d_pts, const double *d_val, int want_out64,
                       void *d_workspace, int64_t workspace_bytes, void *stream);
/* Phase 1 in two steps, for callers that overlap the halo exchange with the sweeps: fb_slab_inject_dev does the centring
 * and the injection of all own planes; fb_slab_sweeps_dev then runs the x and y sweeps of the own planes
 * [plane_begin, plane_begin + plane_count) (relative to z_begin; planes are independent in these sweeps), so that the
 * planes the neighbours need can be finished -- and sent -- first. */
int fb_slab_inject_dev(const fb_problem *prob, int64_t z_begin, int64_t z_count, int64_t halo_lo, int64_t halo_hi,
                       int64_t nsamples, const double *d_pts, const double *d_val, int want_out64,
                       void *d_workspace, int64_t workspace_bytes, void *stream);
int fb_slab_sweeps_dev(const fb_problem *prob, int64_t z_begin, int64_t z_count, int64_t halo_lo, int64_t halo_hi,
                       int64_t nsamples, int want_out64, int64_t plane_begin, int64_t plane_count,
                       void *d_workspace, int64_t workspace_bytes, void *stream);
/* Phase 2 (after the halo planes have been filled): z sweep + mask + divide + cast;
 * d_out [z_count][y][x] float32 receives the own planes. */
int fb_slab_phase2_dev(const fb_problem *prob, int64_t z_begin, int64_t z_count, int64_t halo_lo, int64_t halo_hi,
                       int64_t nsamples, float *d_out, double *d_out64, void *d_workspace,
                       int64_t workspace_bytes, void *stream);

/* The same, leaving the result where the z sweep writes it: the extended float32 volume [z_ext][y][x] at byte offset
 * *offset_out32 of the workspace (and the fp64 quotient at *offset_out64, -1 if not requested); the own planes are planes
 * halo_lo .. halo_lo + z_count - 1 of it.  Saves one device copy of the result per call. */
int fb_slab_result_offsets(const fb_problem *prob, int64_t nsamples, int64_t z_count, int64_t halo_lo, int64_t halo_hi,
                           int want_out64, int64_t *offset_out32, int64_t *offset_out64);
int fb_slab_phase2_inplace_dev(const fb_problem *prob, int64_t z_begin, int64_t z_count, int64_t halo_lo, int64_t halo_hi,
                               int64_t nsamples, int want_out64, void *d_workspace, int64_t workspace_bytes, void *stream);

/* Interprocess events for the peer-mapped halo exchange (one process per GPU on one node): an event created here can be
 * opened in another process through its 64-byte handle; fb_event_record / fb_stream_wait_event are cudaEventRecord /
 * cudaStreamWaitEvent on the given streams (0 = the default stream).  The order of a record and the wait that is meant to see
 * it is the order of the two host calls: synchronise the processes between them. */
int fb_ipc_event_create(void **event, void *handle64);
int fb_ipc_event_open(const void *handle64, void **event);
int fb_event_record(void *event, void *stream);
int fb_stream_wait_event(void *stream, void *event);
int fb_event_destroy(void *event);

/* ---- stages (private-but-tested functions of the reference) -------------------------------- */
/* interpolation.py:485-533 _accumulate_tail_array (alpha) / :729-772 _accumulate_array
 * (alpha = 0) applied in place to n_outer*n_inner independent lines of length len stored as
 * lines[outer][k][inner] (inner contiguous), HOST memory.  rect_len = 2T+1.               */
int fb_accumulate_lines_host(double *lines, int64_t n_outer, int64_t len, int64_t n_inner,
                             int64_t rect_len, int num_iter, double alpha);
/* interpolation.py:373-479 _convolve_tail_{1,2,3}d / :617-724 _convolve_{1,2,3}d: in place on
 * HOST grids vg, wg [z][y][x]; weights below conv_scale_factor become NaN.                 */
int fb_convolve_host(int dim, double *vg, double *wg, const int64_t *size,
                     const int32_t *kernel_size, int num_iter, const double *tail_value,
                     double conv_scale_factor);
/* interpolation.py:205-212 + :219-322: centre the values (offset returned per field) and inject
 * them; vg, wg [nfields][z][y][x] HOST grids receive the injected fields.                   */
int fb_inject_host(const fb_problem *prob, int64_t nsamples, const int64_t *sample_offsets,
                   const double *pts, const double *val, double *vg, double *wg, double *offsets);

/* ---- exact Gaussian sums ("next" row N3) --------------------------------------------------- */
/* prob->method = FB_METHOD_NAIVE (_interpolate_naive, interpolation.py:862-938; dim 1-3),
 * FB_METHOD_RADIUS (_interpolate_radius, :809-855; dim 2, sigma[0] == sigma[1]; the kd-tree search
 * of util/kdtree.py is an exhaustive scan with the same inclusion rule; uses prob->max_dist_weight
 * and min_weight) or FB_METHOD_NAIVE_S2 (_interpolate_naive_S2, interpolationS2.py:260-301; pts
 * are (lon, lat) in degrees).  nfields must be 1; num_iter and flags are ignored.
 * out64 [z][y][x] float64 like the reference (no float32 cast).  Every grid point sums the
 * samples in sample order; the reference sums with np.dot / np.sum or in kd-tree order, so
 * results agree to rounding (~1e-13 relative), not bit for bit.                               */
int fb_barnes_exact_host(const fb_problem *prob, int64_t nsamples, const double *pts, const double *val,
                         double min_weight, double *out64);
/* same on DEVICE buffers; d_scratch: >= 256 bytes of device memory; enqueues on `stream`. */
int fb_barnes_exact_dev(const fb_problem *prob, int64_t nsamples, const double *d_pts, const double *d_val,
                        double min_weight, double *d_out64, void *d_scratch, void *stream);

/* ---- S2 path ------------------------------------------------------------------------------ */
/* util/lambert_conformal.py:50-94 create_proj -> proj[5] = (center_lon, n, n_inv, F, rho0) */
int fb_lambert_create_proj(double center_lon, double center_lat, double lat1, double lat2,
                           double *proj);
/* util/lambert_conformal.py:113-123 to_map: geoc, mapc [n][2] HOST */
int fb_lambert_to_map_host(const double *geoc, double *mapc, int64_t n, const double *proj);
/* interpolationS2.py:180-196 interpolate_opt_convol_S2_part1: project the samples, run the
 * optimized convolution on the fixed Lambert grid lam_x0 = (-32,-2), lam_size =
 * (int(64/step0), int(44/step1)); lam_field [lam_size1][lam_size0] float32 HOST.           */
int fb_s2_part1_host(int64_t nsamples, const double *pts, const double *val, const double *sigma,
                     const double *step, int num_iter, double max_dist_weight,
                     const double *proj, float *lam_field);
/* interpolationS2.py:212-254 _resample (part2): bilinear resampling of the Lambert field to
 * the lon/lat grid; res [size1][size0] float32 HOST.                                        */
int fb_s2_resample_host(const float *lam_field, int64_t lam_w, int64_t lam_h, const double *lam_x0,
                        const double *x0, const double *step, const int64_t *size,
                        const double *proj, float *res);
/* interpolationS2.py:144-177 _interpolate_opt_convol_S2 with resample=True: part1 + part2 with
 * the Lambert field kept on the device; res [size1][size0] float32 HOST.                    */
int fb_barnes_s2_host(int64_t nsamples, const double *pts, const double *val, const double *sigma,
                      const double *x0, const double *step, const int64_t *size, int num_iter,
                      double max_dist_weight, const double *proj, float *res);

/* Generalised S2 ("next" row N4): the reference hard-codes its Lambert map (interpolationS2.py:187-188
 * lam_x0 = (-32, -2), extent 64 x 44 map degrees; :208 projection centre (11.5, 34.5), standard
 * parallels 42.5 / 65.5).  fb_s2_map makes projection and map window arguments; the *_map entry
 * points run the same three steps on it.  Output pixels whose bilinear stencil leaves the map
 * window become NaN.                                                                          */
typedef struct fb_s2_map {
    double proj[5];         /* fb_lambert_create_proj: (center_lon, n, n_inv, F, rho0) */
    double lam_x0[2];       /* start of the grid in map coordinates (degrees) */
    double lam_extent[2];   /* extent of the grid in map degrees; lam_size = (int)(extent / step) */
} fb_s2_map;
/* the reference's fixed map */
int fb_s2_default_map(fb_s2_map *map);
int fb_s2_part1_map_host(int64_t nsamples, const double *pts, const double *val, const double *sigma,
                         const double *step, int num_iter, double max_dist_weight, const fb_s2_map *map,
                         float *lam_field);
int fb_barnes_s2_map_host(int64_t nsamples, const double *pts, const double *val, const double *sigma,
                          const double *x0, const double *step, const int64_t *size, int num_iter,
                          double max_dist_weight, const fb_s2_map *map, float *res);

/* ---- tuning ------------------------------------------------------------------------------- */
/* process-wide tuning switches that never change results (bit-identical either way):
 *   "sweepq" (default 1): 2D / 3D fp64 grids run on interleaved (value, weight) nodes with the q kernels (TMA row
 *       staging, rings in tensor + shared memory) wherever 2T+2 >= 8 and the rings fit on chip; 0: first-generation kernel
 *   "sweepq_stages" (default 3), "sweepq_prefetch" (default 0), "sweepq_warps" (default 8): staging slots per warp, extra
 *       L2 prefetch lead in chunks, warps per CTA of the q kernels
 *   "sweepq_deep_staging" (default 1): the transposing / in-place q sweeps take a fourth staging slot, with 7 warps per CTA
 *       where 8 do not fit;  "sweepq_cap_warps_xy" / "sweepq_cap_warps_final" (default 8): most warps per CTA
 *   "sweepq_reserve_sms" (default 0): SMs the persistent q kernels leave free (room for kernels of other streams)
 *   "sweepp" (default 1): pass-parallel kernels for small batches: 0 off, 1 per axis when the q kernel would have at most
 *       1.5 x SMs units of work (or cannot run), 2 always
 *   "line1d" (default 1): 1D grids of >= 1024 points run the two-warp line kernel; 0: lane-pair walk
 *   "sparse_inject" (default 0): the x sweep synthesises its rows from the samples binned by cell instead of reading a
 *       dense injection grid (large device-resident batches only; measured slower than the dense path)
 *   "inject_lists" (default 1): with interleaved nodes, link the records of a node into a list (two passes over
 *       the samples) instead of count / allocate / place (three)
 *   "host_chunk_fields" (default 16): fields per chunk of the pipelined fb_barnes_host path     */
int  fb_set_option(const char *name, int value);

/* ---- introspection for benchmarks --------------------------------------------------------- */
/* number of kernels launched by this library in the calling process so far */
int64_t fb_kernel_launch_count(void);
/* With profiling enabled (fb_set_profiling(1)) the interpolation entry points bracket their
 * stages with CUDA events on the launching stream.  fb_last_profile waits for the last call of
 * the calling thread and returns the duration (ms) of up to 5 segments:
 *   [0] zero-fill + init  [1] min/max + injection  [2] x sweep  [3] y sweep  [4] z sweep
 * (absent sweeps report 0) and the number of kernels that call launched.                     */
int  fb_set_profiling(int enabled);
int  fb_last_profile(double *ms_segments, int nsegments, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* FASTBARNES_B200_H */
