// fb_exact.cuh -- the exact Gaussian sums of the reference (its accuracy yardsticks) for sm_100a.
//
//   KIND 0  'naive'     fastbarnes/interpolation.py:862-938   _interpolate_naive (1D / 2D / 3D)
//   KIND 1  'naive_S2'  fastbarnes/interpolationS2.py:260-301 _interpolate_naive_S2 + _dist_S2
//   KIND 2  'radius'    fastbarnes/interpolation.py:809-855   _interpolate_radius (2D, scalar sigma);
//                       the kd-tree radius search (util/kdtree.py:297-329) is an exhaustive scan with
//                       the same inclusion rule `sqr_dist <= radius**2`
//
// One thread owns one grid point and walks all samples in sample order, so the two sums are
// deterministic (the reference sums with np.dot / np.sum or in kd-tree order; results agree to
// rounding, not bit for bit).  Compute-bound fp64 work (exp, and sin/cos/acos on the sphere): the
// samples stream through shared memory in tiles that every thread of the block reads as
// broadcasts, so global traffic is one pass over the samples per block plus 8 B per grid point.
#pragma once

#define FB_EXACT_TILE 512
#define FB_EXACT_THREADS 256

struct FbExact {
    const double *pts;              // [n][dim]
    const double *val;              // [n]
    const unsigned long long *mm;   // min/max record of the values (fb_minmax_kernel)
    long long n;
    long long W, H, Dz;
    double x0[3], step[3], scale[3];    // scale = 2 sigma^2
    double radius_sqr;              // KIND 2
    double max_dist_weight;         // KIND 2
    double *out;                    // [z][y][x] float64
};

template <int KIND, int DIM>
__global__ void __launch_bounds__(FB_EXACT_THREADS)
fb_exact_kernel(const FbExact p)
{
    __shared__ double sa[FB_EXACT_TILE], sb[FB_EXACT_TILE], sc[FB_EXACT_TILE], sv[FB_EXACT_TILE];
    const double rad_per_degree = 3.141592653589793 / 180.0;
    const long long total = p.W * p.H * p.Dz;
    const long long idx = (long long)blockIdx.x * FB_EXACT_THREADS + threadIdx.x;
    const bool active = idx < total;
    const long long i = idx % p.W, j = (idx / p.W) % p.H, k = idx / (p.W * p.H);
    // grid point coordinates: x0 + i*step                                    (interpolation.py:878 ff.)
    const double xc = __dadd_rn(p.x0[0], __dmul_rn((double)i, p.step[0]));
    const double yc = DIM > 1 ? __dadd_rn(p.x0[1], __dmul_rn((double)j, p.step[1])) : 0.0;
    const double zc = DIM > 2 ? __dadd_rn(p.x0[2], __dmul_rn((double)k, p.step[2])) : 0.0;
    double sin0 = 0.0, cos0 = 0.0;
    if (KIND == 1) {
        const double lat0_rad = __dmul_rn(yc, rad_per_degree);
        sin0 = sin(lat0_rad);
        cos0 = cos(lat0_rad);
    }
    const double offset = fb_field_offset(p.mm, 0);
    double weighted_sum = 0.0, weight_total = 0.0;

    for (long long s0 = 0; s0 < p.n; s0 += FB_EXACT_TILE) {
        const int cnt = (int)((p.n - s0) < FB_EXACT_TILE ? (p.n - s0) : FB_EXACT_TILE);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += FB_EXACT_THREADS) {
            const double *q = p.pts + (s0 + t) * DIM;
            if (KIND == 1) {
                const double lat1_rad = __dmul_rn(q[1], rad_per_degree);
                sa[t] = q[0];
                sb[t] = sin(lat1_rad);
                sc[t] = cos(lat1_rad);
            } else {
                sa[t] = q[0];
                if (DIM > 1) sb[t] = q[1];
                if (DIM > 2) sc[t] = q[2];
            }
            sv[t] = __dsub_rn(p.val[s0 + t], offset);        // val -= offset  (:209-211)
        }
        __syncthreads();
        if (!active) continue;
#pragma unroll 2
        for (int t = 0; t < cnt; ++t) {
            double weight;
            if (KIND == 1) {
                // _dist_S2 (interpolationS2.py:295-301), angles in degrees
                double arg = __dadd_rn(__dmul_rn(sin0, sb[t]),
                                       __dmul_rn(__dmul_rn(cos0, sc[t]), cos(__dmul_rn(__dsub_rn(sa[t], xc), rad_per_degree))));
                if (arg > 1.0) arg = 1.0;
                const double dist = __ddiv_rn(acos(arg), rad_per_degree);
                weight = exp(__ddiv_rn(__dmul_rn(-dist, dist), p.scale[0]));
            } else if (KIND == 2) {
                const double dx = __dsub_rn(xc, sa[t]), dy = __dsub_rn(yc, sb[t]);
                const double sqr_dist = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                if (!(sqr_dist <= p.radius_sqr)) continue;
                weight = exp(__ddiv_rn(-sqr_dist, p.scale[0]));
            } else {
                const double dx = __dsub_rn(sa[t], xc);
                double sqr_dist = __ddiv_rn(__dmul_rn(dx, dx), p.scale[0]);
                if (DIM > 1) {
                    const double dy = __dsub_rn(sb[t], yc);
                    sqr_dist = __dadd_rn(sqr_dist, __ddiv_rn(__dmul_rn(dy, dy), p.scale[1]));
                }
                if (DIM > 2) {
                    const double dz = __dsub_rn(sc[t], zc);
                    sqr_dist = __dadd_rn(sqr_dist, __ddiv_rn(__dmul_rn(dz, dz), p.scale[2]));
                }
                weight = exp(-sqr_dist);
            }
            weighted_sum = __dadd_rn(weighted_sum, __dmul_rn(weight, sv[t]));
            weight_total = __dadd_rn(weight_total, weight);
        }
    }
    if (active) {
        const bool keep = KIND == 2 ? (weight_total >= p.max_dist_weight) : (weight_total > 0.0);
        p.out[idx] = keep ? __dadd_rn(__ddiv_rn(weighted_sum, weight_total), offset)
                          : __longlong_as_double(0x7ff8000000000000ll);
    }
}
