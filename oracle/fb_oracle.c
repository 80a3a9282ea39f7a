/*
 * fb_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference's optimized-convolution Barnes
 * interpolation hot path (MeteoSwiss/fast-barnes-py v2.0.0).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library, and only as the checker or the CPU baseline -- never as the
 * thing shipped.  The product path (fast-barnes-py_b200/) never links it.
 *
 * Parity pin: this restatement is checked bit-for-bit against the reference itself
 * (imported under Numba in the build container, see oracle/gen_golden.py and
 * tests/test_oracle_vs_reference.py) and against the reference's own known-answer
 * vectors (tests/AccumulationTest.py).  Committed fixtures: tests/golden/.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference root).  Compile with -ffp-contract=off: the Numba code has no FMA.
 *
 * The line loops of the sweeps are OpenMP-parallel over grid lines.  Lines are
 * independent, so the result is bit-identical for any thread count; the reference
 * itself is single-threaded.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FBO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* fastbarnes/interpolation.py:205-212  _normalize_values                     */
FBO_API double fbo_normalize_values(double *val, int64_t n)
{
    double mn = val[0], mx = val[0];
    int has_nan = 0;
    for (int64_t i = 0; i < n; i++) {
        if (val[i] != val[i]) has_nan = 1;
        if (val[i] < mn) mn = val[i];
        if (val[i] > mx) mx = val[i];
    }
    double offset = has_nan ? NAN : (mn + mx) / 2.0;
    for (int64_t i = 0; i < n; i++) val[i] -= offset;
    return offset;
}

/* ------------------------------------------------------------------------- */
/* fastbarnes/interpolation.py:219-237 / 241-270 / 274-322  _inject_data_{1,2,3}d
 * vg, wg have reversed dims: [x], [y][x], [z][y][x] (row-major).            */
FBO_API void fbo_inject_data(int dim, double *vg, double *wg, const double *pts,
                             const double *val, int64_t n, const double *x0,
                             const double *step, const int64_t *size)
{
    const int64_t W = size[0];
    const int64_t H = dim > 1 ? size[1] : 1;
    for (int64_t k = 0; k < n; k++) {
        if (dim == 1) {
            double xc = (pts[k] - x0[0]) / step[0];
            if (xc < 0.0 || xc >= (double)(size[0] - 1)) continue;
            int64_t xi = (int64_t)xc;
            double xw = xc - (double)xi;
            double w = (1.0 - xw);
            vg[xi] += w * val[k];
            wg[xi] += w;
            w = xw;
            vg[xi + 1] += w * val[k];
            wg[xi + 1] += w;
        } else if (dim == 2) {
            double xc = (pts[2 * k] - x0[0]) / step[0];
            double yc = (pts[2 * k + 1] - x0[1]) / step[1];
            if (xc < 0.0 || yc < 0.0 || xc >= (double)(size[0] - 1) || yc >= (double)(size[1] - 1)) continue;
            int64_t xi = (int64_t)xc, yi = (int64_t)yc;
            double xw = xc - (double)xi, yw = yc - (double)yi;
            double w;
            w = (1.0 - xw) * (1.0 - yw);
            vg[yi * W + xi] += w * val[k];
            wg[yi * W + xi] += w;
            w = xw * (1.0 - yw);
            vg[yi * W + xi + 1] += w * val[k];
            wg[yi * W + xi + 1] += w;
            w = xw * yw;
            vg[(yi + 1) * W + xi + 1] += w * val[k];
            wg[(yi + 1) * W + xi + 1] += w;
            w = (1.0 - xw) * yw;
            vg[(yi + 1) * W + xi] += w * val[k];
            wg[(yi + 1) * W + xi] += w;
        } else {
            double xc = (pts[3 * k] - x0[0]) / step[0];
            double yc = (pts[3 * k + 1] - x0[1]) / step[1];
            double zc = (pts[3 * k + 2] - x0[2]) / step[2];
            if (xc < 0.0 || yc < 0.0 || zc < 0.0 || xc >= (double)(size[0] - 1) ||
                yc >= (double)(size[1] - 1) || zc >= (double)(size[2] - 1)) continue;
            int64_t xi = (int64_t)xc, yi = (int64_t)yc, zi = (int64_t)zc;
            double xw = xc - (double)xi, yw = yc - (double)yi, zw = zc - (double)zi;
            double w;
#define NODE(z, y, x) (((z) * H + (y)) * W + (x))
            w = (1.0 - xw) * (1.0 - yw) * (1.0 - zw);
            vg[NODE(zi, yi, xi)] += w * val[k];
            wg[NODE(zi, yi, xi)] += w;
            w = xw * (1.0 - yw) * (1.0 - zw);
            vg[NODE(zi, yi, xi + 1)] += w * val[k];
            wg[NODE(zi, yi, xi + 1)] += w;
            w = xw * yw * (1.0 - zw);
            vg[NODE(zi, yi + 1, xi + 1)] += w * val[k];
            wg[NODE(zi, yi + 1, xi + 1)] += w;
            w = (1.0 - xw) * yw * (1.0 - zw);
            vg[NODE(zi, yi + 1, xi)] += w * val[k];
            wg[NODE(zi, yi + 1, xi)] += w;
            w = (1.0 - xw) * (1.0 - yw) * zw;
            vg[NODE(zi + 1, yi, xi)] += w * val[k];
            wg[NODE(zi + 1, yi, xi)] += w;
            w = xw * (1.0 - yw) * zw;
            vg[NODE(zi + 1, yi, xi + 1)] += w * val[k];
            wg[NODE(zi + 1, yi, xi + 1)] += w;
            w = xw * yw * zw;
            vg[NODE(zi + 1, yi + 1, xi + 1)] += w * val[k];
            wg[NODE(zi + 1, yi + 1, xi + 1)] += w;
            w = (1.0 - xw) * yw * zw;
            vg[NODE(zi + 1, yi + 1, xi)] += w * val[k];
            wg[NODE(zi + 1, yi + 1, xi)] += w;
#undef NODE
        }
    }
}

/* ------------------------------------------------------------------------- */
/* fastbarnes/interpolation.py:549-552  _get_half_kernel_size_opt             */
FBO_API int32_t fbo_half_kernel_size_opt(double sigma, double step, int num_iter)
{
    double s = sigma / step;
    return (int32_t)((sqrt(1.0 + 12 * s * s / num_iter) - 1.0) / 2.0);
}

/* fastbarnes/interpolation.py:783-785  _get_half_kernel_size ('convolution') */
FBO_API int32_t fbo_half_kernel_size(double sigma, double step, int num_iter)
{
    return (int32_t)(sqrt(3.0 / num_iter) * sigma / step + 0.5);
}

/* fastbarnes/interpolation.py:561-569  _get_tail_value                       */
FBO_API double fbo_tail_value(double sigma, double step, int num_iter)
{
    int64_t hks = fbo_half_kernel_size_opt(sigma, step, num_iter);
    int64_t ks = 2 * hks + 1;
    double sigma_rect_sqr = (double)((hks + 1) * hks) / 3.0 * (step * step);
    double hs = (double)(hks + 1) * step;
    return 0.5 * (double)ks * (sigma * sigma / num_iter - sigma_rect_sqr) /
           (hs * hs - sigma * sigma / num_iter);
}

/* `float64_array ** int` as Numba lowers it (numba/cpython/numbers.py int_power_impl:
 * square-and-multiply, NOT libm pow -- measured: libm pow differs in the last bit for
 * about half of all inputs at n=4).                                                  */
static double int_power(double a, int64_t b)
{
    double r = 1.0;
    int invert = 0;
    int64_t e = b;
    if (b < 0) { invert = 1; e = -b; }
    if (e > 0x10000) return pow(a, (double)b);
    while (e != 0) {
        if (e & 1) r *= a;
        e >>= 1;
        a *= a;
    }
    return invert ? 1.0 / r : r;
}

/* fastbarnes/interpolation.py:424-425 (and :389-390, :472-473)  conv_scale_factor.
 * kernel_size[m], tail_value[m] per axis; np.prod multiplies left to right from 1. */
FBO_API double fbo_conv_scale_factor(int dim, const int32_t *kernel_size, const double *tail_value,
                                     const double *sigma, const double *step, int num_iter,
                                     double max_dist_weight)
{
    double prod = 1.0;
    for (int m = 0; m < dim; m++) {
        double f = int_power((double)kernel_size[m] + 2 * tail_value[m], num_iter) /
                   sqrt(2 * M_PI) / (sigma[m] / step[m]);
        prod *= f;
    }
    return prod * max_dist_weight;
}

/* ------------------------------------------------------------------------- */
/* fastbarnes/interpolation.py:485-533  _accumulate_tail_array
 * Returns 0 if the result ends up in in_arr, 1 if it ends up in h_arr.      */
FBO_API int fbo_accumulate_tail_array(double *in_arr, double *h_arr, int64_t arr_len,
                                      int64_t rect_len, int num_iter, double alpha)
{
    int64_t h0 = (rect_len - 1) / 2;
    int64_t h0_1 = h0 + 1;
    int64_t h1 = rect_len - h0;
    int which = 0;
    for (int i = 0; i < num_iter; i++) {
        double accu = 0.0;
        int64_t k;
        for (k = -h0; k < 0; k++) accu += in_arr[k + h0];
        for (k = 0; k < h1; k++) {
            accu += in_arr[k + h0];
            h_arr[k] = accu + alpha * in_arr[k + h0_1];
        }
        for (k = h1; k < arr_len - h0_1; k++) {
            accu += (in_arr[k + h0] - in_arr[k - h1]);
            h_arr[k] = accu + alpha * (in_arr[k - h1] + in_arr[k + h0_1]);
        }
        k = arr_len - h0_1;
        accu += (in_arr[k + h0] - in_arr[k - h1]);
        h_arr[k] = accu + alpha * in_arr[k - h1];
        for (k = arr_len - h0; k < arr_len; k++) {
            accu -= in_arr[k - h1];
            h_arr[k] = accu + alpha * in_arr[k - h1];
        }
        double *h = in_arr;
        in_arr = h_arr;
        h_arr = h;
        which ^= 1;
    }
    return which;
}

/* fastbarnes/interpolation.py:729-772  _accumulate_array ('convolution', row N1) */
FBO_API int fbo_accumulate_array(double *in_arr, double *h_arr, int64_t arr_len,
                                 int64_t rect_len, int num_iter)
{
    int64_t h0 = (rect_len - 1) / 2;
    int64_t h1 = rect_len - h0;
    int which = 0;
    for (int i = 0; i < num_iter; i++) {
        double accu = 0.0;
        int64_t k;
        for (k = -h0; k < 0; k++) accu += in_arr[k + h0];
        for (k = 0; k < h1; k++) {
            accu += in_arr[k + h0];
            h_arr[k] = accu;
        }
        for (k = h1; k < arr_len - h0; k++) {
            accu += (in_arr[k + h0] - in_arr[k - h1]);
            h_arr[k] = accu;
        }
        for (k = arr_len - h0; k < arr_len; k++) {
            accu -= in_arr[k - h1];
            h_arr[k] = accu;
        }
        double *h = in_arr;
        in_arr = h_arr;
        h_arr = h;
        which ^= 1;
    }
    return which;
}

/* One grid line (start pointer + element stride) through the n-fold filter,
 * as the `arr[...] = _accumulate_tail_array(arr[...].copy(), h_arr, ...)` idiom of
 * fastbarnes/interpolation.py:380-383, 404-418, 440-466.  tail < 0 selects the
 * plain `_accumulate_array` of the 'convolution' method (:625-711).            */
static void sweep_line(double *base, int64_t stride, int64_t len, int64_t rect_len, int num_iter,
                       double alpha, int plain, double *buf_a, double *buf_b)
{
    for (int64_t i = 0; i < len; i++) buf_a[i] = base[i * stride];
    int which = plain ? fbo_accumulate_array(buf_a, buf_b, len, rect_len, num_iter)
                      : fbo_accumulate_tail_array(buf_a, buf_b, len, rect_len, num_iter, alpha);
    const double *res = which ? buf_b : buf_a;
    for (int64_t i = 0; i < len; i++) base[i * stride] = res[i];
}

/* ------------------------------------------------------------------------- */
/* fastbarnes/interpolation.py:373-394 / 398-430 / 434-479  _convolve_tail_{1,2,3}d
 * (plain != 0: _convolve_{1,2,3}d, :617-724).  In place on vg, wg.
 * The NaN-masking threshold is passed in (fbo_conv_scale_factor).            */
FBO_API void fbo_convolve(int dim, double *vg, double *wg, const int64_t *size,
                          const int32_t *kernel_size, int num_iter, const double *tail_value,
                          int plain, double conv_scale_factor, int nthreads)
{
    const int64_t W = size[0];
    const int64_t H = dim > 1 ? size[1] : 1;
    const int64_t D = dim > 2 ? size[2] : 1;
    int64_t maxlen = W > H ? W : H;
    if (D > maxlen) maxlen = D;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        double *buf_a = (double *)malloc(sizeof(double) * (size_t)maxlen);
        double *buf_b = (double *)malloc(sizeof(double) * (size_t)maxlen);
        /* x direction: lines (k, j) */
#pragma omp for schedule(static)
        for (int64_t l = 0; l < D * H; l++) {
            sweep_line(vg + l * W, 1, W, kernel_size[0], num_iter, tail_value[0], plain, buf_a, buf_b);
            sweep_line(wg + l * W, 1, W, kernel_size[0], num_iter, tail_value[0], plain, buf_a, buf_b);
        }
        if (dim > 1) {
            /* y direction: lines (k, i) */
#pragma omp for schedule(static)
            for (int64_t l = 0; l < D * W; l++) {
                int64_t k = l / W, i = l % W;
                sweep_line(vg + k * H * W + i, W, H, kernel_size[1], num_iter, tail_value[1], plain, buf_a, buf_b);
                sweep_line(wg + k * H * W + i, W, H, kernel_size[1], num_iter, tail_value[1], plain, buf_a, buf_b);
            }
        }
        if (dim > 2) {
            /* z direction: lines (j, i) */
#pragma omp for schedule(static)
            for (int64_t l = 0; l < H * W; l++) {
                sweep_line(vg + l, H * W, D, kernel_size[2], num_iter, tail_value[2], plain, buf_a, buf_b);
                sweep_line(wg + l, H * W, D, kernel_size[2], num_iter, tail_value[2], plain, buf_a, buf_b);
            }
        }
        /* set smaller weights to NaN (:392-394, :427-430, :475-479) */
#pragma omp for schedule(static)
        for (int64_t i = 0; i < D * H * W; i++)
            if (wg[i] < conv_scale_factor) wg[i] = NAN;
        free(buf_a);
        free(buf_b);
    }
}

/* ------------------------------------------------------------------------- */
/* fastbarnes/interpolation.py:329-367 _interpolate_opt_convol  (plain != 0:
 * :575-612 _interpolate_convol).  val is modified in place (centred), as in the
 * reference.  out32 receives (vg/wg+offset).astype(float32); if out64 != NULL it
 * receives the pre-cast fp64 quotient; if vg_out/wg_out != NULL they receive the
 * post-sweep fields (wg NaN-masked).  Returns offset.                           */
FBO_API double fbo_interpolate(int dim, const double *pts, double *val, int64_t n,
                               const double *sigma, const double *x0, const double *step,
                               const int64_t *size, int num_iter, double max_dist_weight,
                               int plain, int nthreads, float *out32, double *out64,
                               double *vg_out, double *wg_out, double *vin_out, double *win_out)
{
    double offset = fbo_normalize_values(val, n);
    int64_t total = 1;
    for (int m = 0; m < dim; m++) total *= size[m];
    double *vg = (double *)calloc((size_t)total, sizeof(double));
    double *wg = (double *)calloc((size_t)total, sizeof(double));
    fbo_inject_data(dim, vg, wg, pts, val, n, x0, step, size);
    if (vin_out) memcpy(vin_out, vg, sizeof(double) * (size_t)total);
    if (win_out) memcpy(win_out, wg, sizeof(double) * (size_t)total);

    int32_t ks[3];
    double tv[3];
    for (int m = 0; m < dim; m++) {
        int32_t hk = plain ? fbo_half_kernel_size(sigma[m], step[m], num_iter)
                           : fbo_half_kernel_size_opt(sigma[m], step[m], num_iter);
        ks[m] = 2 * hk + 1;
        tv[m] = plain ? 0.0 : fbo_tail_value(sigma[m], step[m], num_iter);
    }
    double csf = fbo_conv_scale_factor(dim, ks, tv, sigma, step, num_iter, max_dist_weight);
    fbo_convolve(dim, vg, wg, size, ks, num_iter, tv, plain, csf, nthreads);

#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t i = 0; i < total; i++) {
        double q = vg[i] / wg[i] + offset;
        if (out64) out64[i] = q;
        out32[i] = (float)q;
    }
    if (vg_out) memcpy(vg_out, vg, sizeof(double) * (size_t)total);
    if (wg_out) memcpy(wg_out, wg, sizeof(double) * (size_t)total);
    free(vg);
    free(wg);
    return offset;
}

/* ------------------------------------------------------------------------- */
/* fastbarnes/util/lambert_conformal.py:45-46, 50-94  create_proj
 * proj = (center_lon, n, n_inv, F, rho0)                                      */
#define RAD_PER_DEGREE (M_PI / 180.0)
#define HALF_RAD_PER_DEGREE (RAD_PER_DEGREE / 2.0)

FBO_API void fbo_lambert_create_proj(double center_lon, double center_lat, double lat1,
                                     double lat2, double *proj)
{
    double n;
    if (lat1 != lat2)
        n = log(cos(lat1 * RAD_PER_DEGREE) / cos(lat2 * RAD_PER_DEGREE)) /
            log(tan((90.0 + lat2) * HALF_RAD_PER_DEGREE) / tan((90.0 + lat1) * HALF_RAD_PER_DEGREE));
    else
        n = sin(lat1 * RAD_PER_DEGREE);
    double n_inv = 1.0 / n;
    double F = cos(lat1 * RAD_PER_DEGREE) * pow(tan((90.0 + lat1) * HALF_RAD_PER_DEGREE), n) / n;
    double rho0 = F / pow(tan((90.0 + center_lat) * HALF_RAD_PER_DEGREE), n);
    proj[0] = center_lon;
    proj[1] = n;
    proj[2] = n_inv;
    proj[3] = F;
    proj[4] = rho0;
}

/* fastbarnes/util/lambert_conformal.py:113-123  to_map  (geoc, mapc: [N][2]) */
FBO_API void fbo_lambert_to_map(const double *geoc, double *mapc, int64_t n, const double *proj)
{
    const double center_lon = proj[0], nn = proj[1], F = proj[3], rho0 = proj[4];
    for (int64_t i = 0; i < n; i++) {
        double rho = F / pow(tan((90.0 + geoc[2 * i + 1]) * HALF_RAD_PER_DEGREE), nn);
        double arg = nn * (geoc[2 * i] - center_lon) * RAD_PER_DEGREE;
        mapc[2 * i] = rho * sin(arg) / RAD_PER_DEGREE;
        mapc[2 * i + 1] = (rho0 - rho * cos(arg)) / RAD_PER_DEGREE;
    }
}

/* fastbarnes/interpolationS2.py:212-254  _resample  (incl. to_map2,
 * fastbarnes/util/lambert_conformal.py:127-136).  lam_field f32 [lamH][lamW];
 * res f32 [size[1]][size[0]].                                                 */
FBO_API void fbo_resample(const float *lam_field, int64_t lam_w, const double *lam_x0,
                          const double *x0, const double *step, const int64_t *size,
                          const double *proj, float *res, int nthreads)
{
    const double center_lon = proj[0], nn = proj[1], F = proj[3], rho0 = proj[4];
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t j = 0; j < size[1]; j++) {
        double geoy = (double)j * step[1] + x0[1];
        double rho = F / pow(tan((90.0 + geoy) * HALF_RAD_PER_DEGREE), nn);
        for (int64_t i = 0; i < size[0]; i++) {
            double geox = x0[0] + (double)i * step[0];
            double arg = nn * (geox - center_lon) * RAD_PER_DEGREE;
            double mapx = rho * sin(arg) / RAD_PER_DEGREE;
            double mapy = (rho0 - rho * cos(arg)) / RAD_PER_DEGREE;
            mapx -= lam_x0[0];
            mapx /= step[0];
            mapy -= lam_x0[1];
            mapy /= step[1];
            int32_t ix = (int32_t)mapx, iy = (int32_t)mapy;
            mapx -= (double)ix;
            mapy -= (double)iy;
            double r = (1.0 - mapy) * (1.0 - mapx) * (double)lam_field[(int64_t)iy * lam_w + ix] +
                       mapy * (1.0 - mapx) * (double)lam_field[(int64_t)(iy + 1) * lam_w + ix] +
                       mapy * mapx * (double)lam_field[(int64_t)(iy + 1) * lam_w + ix + 1] +
                       (1.0 - mapy) * mapx * (double)lam_field[(int64_t)iy * lam_w + ix + 1];
            res[j * size[0] + i] = (float)r;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Exact Gaussian sums ("next" row N3: the accuracy yardsticks).
 *   kind 0: fastbarnes/interpolation.py:862-938  _interpolate_naive (1D/2D/3D)
 *   kind 1: fastbarnes/interpolationS2.py:260-301 _interpolate_naive_S2 with _dist_S2
 *   kind 2: fastbarnes/interpolation.py:809-855  _interpolate_radius (2D); the kd-tree radius
 *           search (util/kdtree.py:297-329, inclusion rule `sqr_dist <= radius**2`) is replaced by
 *           an exhaustive scan with the same rule.
 * The reference sums with np.dot / np.sum (kinds 0, 1) or in kd-tree traversal order (kind 2);
 * this restatement sums in sample order, so it matches the reference to rounding (~1e-13
 * relative), not bit for bit -- the tests state the tolerance.
 * val is centred in place; out [z][y][x] float64; returns the offset.                          */
FBO_API double fbo_interpolate_exact(int kind, int dim, const double *pts, double *val, int64_t n,
                                     const double *sigma, const double *x0, const double *step,
                                     const int64_t *size, double max_dist_weight, double min_weight,
                                     double *out, int nthreads)
{
    const double offset = fbo_normalize_values(val, n);
    const int64_t W = size[0], H = dim > 1 ? size[1] : 1, Dz = dim > 2 ? size[2] : 1;
    double scale[3] = {1.0, 1.0, 1.0};
    for (int m = 0; m < dim; m++) scale[m] = 2 * (sigma[m] * sigma[m]);
    const double rad_per_degree = M_PI / 180.0;
    const double search_radius = kind == 2 ? sqrt(-2.0 * log(min_weight)) * sigma[0] : 0.0;
    const double radius_sqr = search_radius * search_radius;
    (void)nthreads;
#pragma omp parallel for collapse(2) schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t k = 0; k < Dz; k++) {
        for (int64_t j = 0; j < H; j++) {
            const double zc = dim > 2 ? x0[2] + k * step[2] : 0.0;
            const double yc = dim > 1 ? x0[1] + j * step[1] : 0.0;
            for (int64_t i = 0; i < W; i++) {
                const double xc = x0[0] + i * step[0];
                double weighted_sum = 0.0, weight_total = 0.0;
                for (int64_t s = 0; s < n; s++) {
                    double weight;
                    if (kind == 1) {
                        const double lon1 = pts[2 * s], lat1 = pts[2 * s + 1];
                        const double lat0_rad = yc * rad_per_degree, lat1_rad = lat1 * rad_per_degree;
                        double arg = sin(lat0_rad) * sin(lat1_rad) +
                                     cos(lat0_rad) * cos(lat1_rad) * cos((lon1 - xc) * rad_per_degree);
                        if (arg > 1.0) arg = 1.0;
                        const double dist = acos(arg) / rad_per_degree;
                        weight = exp(-dist * dist / scale[0]);
                    } else if (kind == 2) {
                        const double dx = xc - pts[2 * s], dy = yc - pts[2 * s + 1];
                        const double sqr_dist = dx * dx + dy * dy;
                        if (!(sqr_dist <= radius_sqr)) continue;
                        weight = exp(-sqr_dist / scale[0]);
                    } else {
                        const double dx = pts[dim * s] - xc;
                        double sqr_dist = dx * dx / scale[0];
                        if (dim > 1) { const double dy = pts[dim * s + 1] - yc; sqr_dist = sqr_dist + dy * dy / scale[1]; }
                        if (dim > 2) { const double dz = pts[dim * s + 2] - zc; sqr_dist = sqr_dist + dz * dz / scale[2]; }
                        weight = exp(-sqr_dist);
                    }
                    weighted_sum += weight * val[s];
                    weight_total += weight;
                }
                const int keep = kind == 2 ? (weight_total >= max_dist_weight) : (weight_total > 0.0);
                out[(k * H + j) * W + i] = keep ? weighted_sum / weight_total + offset : NAN;
            }
        }
    }
    return offset;
}

FBO_API int fbo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
