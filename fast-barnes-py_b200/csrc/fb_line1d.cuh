// fb_line1d.cuh -- the bit-exact walk of ONE long grid line (1D grids).
//
// _convolve_tail_1d (fastbarnes/interpolation.py:373-394) runs _accumulate_tail_array (:485-533) over a single line: the
// sliding accumulator of every pass is one serial chain of fp64 additions, so a bit-identical result cannot be cut into
// independent pieces.  What CAN run side by side are the 2 x n chains (value / weight field x pass): pass q+1 only needs
// the output of pass q some elements back.  Here they are the lanes of ONE warp:
//   * lane l < 2 n works for field l & 1, pass l >> 1.  All lanes execute the same instructions on their own chain (SIMT):
//     per iteration each takes a chunk of 32 elements of its input stream from a ring in shared memory (two 128-bit
//     loads per 2 elements: the newest element and the one D = 2T+2 back), runs the reference's update
//         accu += in[k+T] - in[k-T-1];   out[k] = accu + alpha * (in[k-T-1] + in[k+T+1])
//     (explicit *_rn operations in the reference's order; zero extension replaces the five loop phases) and writes its 32
//     results into the ring that is the next pass's input stream.  Pass q runs DL chunks behind pass q-1, so that what it
//     reads was written in an earlier iteration (one named barrier per iteration).
//   * a second warp feeds and drains: it copies the line into the first ring with cp.async (LDGSTS: completion by
//     cp.async groups, several chunks ahead, no registers) and takes the last pass's chunk of the previous iteration
//     out of its ring -- MODE 2: `wg[wg < csf] = nan; (vg / wg + offset).astype(float32)` (:392-394, :367), one element
//     per lane, so that the divisions never sit in the chains' warp.
// The chain costs one dependent DADD (~9 cycles) per element and pass; with everything else off that critical path the
// line advances at about one element per 12-15 cycles, against ~270 cycles of the lane-pair walk it replaces.
#pragma once
#include "fb_kernels.cuh"

#define FBL_U 32                 // elements per chunk
#define FBL_PD 4                 // chunks the loader runs ahead

struct FbLine1D {
    const double *in_v, *in_w;   // the line of every field: in[field * L + k]
    double *out_v, *out_w;       // MODE 0: result lines (may be in_v / in_w)
    float *out32;                // MODE 2: out32[field * L + k]
    double *out64;               // MODE 2, optional
    const unsigned long long *mm;
    long long L;
    int T, D, DL, RL;            // DL: chunks between consecutive passes; RL: ring length (power of two)
    double alpha, csf;
};

// the non-aligned form: the lanes of the feeding warp arrive from lane-dependent branches (loads past the end of the line,
// rows outside it) and need not have reconverged (compute-sanitizer synccheck flags `bar.sync` there)
__device__ __forceinline__ void fbl_bar() { asm volatile("barrier.sync 1, 64;" ::: "memory"); }

template <int NPASS, int MODE>
__global__ void __launch_bounds__(64, 1)
fb_line1d_kernel(const FbLine1D p)
{
    constexpr int U = FBL_U;
    extern __shared__ __align__(16) double fbl_smem[];   // rings: stream s (0 = the line, s = output of pass s), field f: ring (2 s + f)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long field = blockIdx.x;
    const long long L = p.L;
    const int T1 = p.T + 1, D = p.D, DL = p.DL;
    const unsigned mask = (unsigned)p.RL - 1u;
    const int nrings = 2 * (NPASS + 1);
    // consecutive rings start 2 doubles (4 banks) further: the lanes of the chains' warp read / write the same position of
    // different rings, which would otherwise all hit the same banks (8-way conflicts: 22 -> 13 ns per point)
    const size_t rs = (size_t)p.RL + 2;
    for (int i = threadIdx.x; i < nrings * (int)rs; i += 64) fbl_smem[i] = 0.0;
    __syncthreads();
    // iterations: pass q handles stream positions 16 (it - q DL) .. + 15; the last pass must reach position L - 1 + T1
    const long long nit = (L + T1 + U - 1) / U + (long long)(NPASS - 1) * DL + 1;

    if (wid == 0) {
        // ------------------------------ the chains ------------------------------
        const int f = lane & 1, q = lane >> 1;
        const bool mine = q < NPASS;
        const double *xin = fbl_smem + (size_t)(2 * (mine ? q : 0) + f) * rs;            // input stream of this pass
        double *xout = fbl_smem + (size_t)(2 * ((mine ? q : 0) + 1) + f) * rs;           // its output stream
        const double alpha = p.alpha;
        double accu = 0.0, new0 = 0.0;
        for (long long it = 0; it < nit; ++it) {
            fbl_bar();                                    // barrier #it: chunk `it` of the line has landed, iteration it-1 is complete
            const long long s0 = (it - (long long)q * DL) * U;                            // first stream position of this lane's chunk
            if (mine && s0 >= 0) {
                const long long k0 = s0 - T1;                                             // output position of the chunk's first step
                const bool interior = k0 >= 0 && k0 + U <= L;
                const unsigned bn = (unsigned)s0 & mask, bo = (unsigned)(s0 - D) & mask, bw = (unsigned)k0 & mask;
                double x[U], o[U], r[U];
#pragma unroll
                for (int j = 0; j < U; j += 2) {
                    // s0 and D are even: pairs are 16-byte aligned; s0 is a multiple of 32 and so is the ring length: the new chunk never wraps
                    const double2 a = *reinterpret_cast<const double2 *>(xin + bn + j);
                    const double2 b = *reinterpret_cast<const double2 *>(xin + ((bo + (unsigned)j) & mask));     // may wrap inside the chunk
                    x[j] = a.x; x[j + 1] = a.y;
                    o[j] = b.x; o[j + 1] = b.y;
                }
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const double d = __dsub_rn(new0, o[j]);
                    accu = __dadd_rn(accu, d);
                    r[j] = __dadd_rn(accu, __dmul_rn(alpha, __dadd_rn(o[j], x[j])));
                    new0 = x[j];
                }
                if (!interior) {
#pragma unroll
                    for (int j = 0; j < U; ++j) r[j] = (k0 + j >= 0 && k0 + j < L) ? r[j] : 0.0;
                }
                if ((T1 & 1) == 0) {
                    // k0 = s0 - T1 is even: 128-bit stores of pairs (the ring is circular in k: a pair never straddles its end)
#pragma unroll
                    for (int j = 0; j < U; j += 2)
                        *reinterpret_cast<double2 *>(xout + ((bw + (unsigned)j) & mask)) = make_double2(r[j], r[j + 1]);
                } else {
#pragma unroll
                    for (int j = 0; j < U; ++j) xout[(bw + (unsigned)j) & mask] = r[j];
                }
            }
        }
        fbl_bar();                                        // barrier #nit: the last iteration is complete
    } else {
        // ------------------------------ feed and drain ------------------------------
        const int f = lane >> 4, j = lane & 15;           // loader: elements j and j + 16 of a chunk of field f; drain: element `lane`
        const double *gin = (f ? p.in_w : p.in_v) + field * L;
        double *x0 = fbl_smem + (size_t)f * rs;           // ring of the line itself (stream 0)
        const double *yv = fbl_smem + (size_t)(2 * NPASS) * rs, *yw = yv + rs;           // output stream of the last pass
        double offset = 0.0;
        if (MODE == 2) offset = fb_field_offset(p.mm, field);
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        auto load_chunk = [&](long long c) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const long long k = c * U + j + 16 * h;
                if (k < L) {
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(x0 + ((unsigned)k & mask));
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gin + k) : "memory");
                } else {
                    x0[(unsigned)k & mask] = 0.0;
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto drain = [&](long long it_done) {             // the chunk the last pass wrote in iteration it_done
            const long long k = (it_done - (long long)(NPASS - 1) * DL) * U - T1 + lane;
            if (k >= 0 && k < L) {
                const double v = yv[(unsigned)k & mask], w = yw[(unsigned)k & mask];
                if (MODE == 2) {
                    // `if wg < csf: wg = nan` (interpolation.py:394); (vg / wg + offset) -> float32 (:367)
                    const double qv = (w < p.csf) ? qnan : __dadd_rn(__ddiv_rn(v, w), offset);
                    p.out32[field * L + k] = __double2float_rn(qv);
                    if (p.out64) p.out64[field * L + k] = qv;
                } else {
                    p.out_v[field * L + k] = v;
                    if (p.out_w) p.out_w[field * L + k] = w;
                }
            }
        };
        for (int c = 0; c < FBL_PD; ++c) load_chunk(c);
        for (long long it = 0; it < nit; ++it) {
            asm volatile("cp.async.wait_group %0;" ::"n"(FBL_PD - 1) : "memory");         // chunk `it` of the line has landed
            fbl_bar();                                    // barrier #it
            load_chunk(it + FBL_PD);                      // (its slots were last read D + 16 elements ago)
            if (it > 0) drain(it - 1);
        }
        fbl_bar();                                        // barrier #nit
        drain(nit - 1);
    }
}
