// fb_sweepp.cuh -- pass-parallel sweep for SMALL batches (a single field, a few fields, one small volume).
//
// The q kernels (fb_sweepq.cuh) give one warp 16 lines and all passes of them; a single 2400 x 1200 field is then 75
// (x sweep) / 150 (y sweep) units of work for 1184 warps, each a serial walk of 2400 / 1200 rows at ~250 cycles per row:
// the literal paper case (one field, N = 3490) spends 135 + 106 us in its two sweeps, launch and latency bound.
// What is parallel inside one line are the 2 x n chains (value / weight field x pass) -- pass q+1 only needs what
// pass q produced some rows back (_accumulate_tail_array, interpolation.py:485-533, one serial chain of fp64 additions
// per line, field and pass).  Here, as in the 1D kernel (fb_line1d.cuh), they are the lanes of a warp:
//   * a CTA of two warps works on NL = 32 / LPL adjacent lines (LPL = 2 n rounded up to a power of two: 4 lines at n = 3, 4).
//     Lane l of the chain warp: line l / LPL, pass (l % LPL) >> 1, field l & 1.  Per iteration every lane takes a chunk of 16
//     elements of ITS input stream from a ring in shared memory (the newest element and the one D = 2T+2 back), runs
//         accu += in[k+T] - in[k-T-1];   out[k] = accu + alpha * (in[k-T-1] + in[k+T+1])
//     (explicit *_rn operations in the reference's order; zero extension replaces the five loop phases) and writes its 16
//     results into the ring that is the next pass's input stream.  Pass q runs DL chunks behind pass q-1, so that what it
//     reads was written in an earlier iteration (one named barrier per iteration).
//   * the second warp feeds and drains: it copies the rows of the NL lines into the first rings with cp.async (LDGSTS,
//     completion by cp.async groups, chunks ahead, no registers) and takes the last pass's chunk of the previous iteration
//     out of its rings: MODE 0 same layout (in place allowed: reads lead the writes of the same lines), MODE 1 transposed
//     (the x sweep), MODE 2 `wg[wg < csf] = nan; (vg / wg + offset).astype(float32)` (:427-430, :367).
// A field of 1200 lines is 300 CTAs of 4 lines: every SM works, and a row costs ~25 cycles instead of ~250.
// Grids are arrays of interleaved (value, weight) nodes, as for the q kernels:
//   in[((outer * L + k) * n_inner + line) * 2 + field], k the sweep axis.
#pragma once
#include "fb_kernels.cuh"

#ifndef FBP_U
#define FBP_U 16                 // elements per chunk
#endif
#ifndef FBP_PD
#define FBP_PD 3                 // chunks the loader runs ahead
#endif

struct FbSweepP {
    const double *in;
    double *out;                 // MODE 0: layout of `in`; MODE 1: out[((outer * n_inner + line) * L + k) * 2 + field]
    float *out32;                // MODE 2: out32[(outer * L + k) * n_inner + line]
    double *out64;               // MODE 2, optional
    const unsigned long long *mm;
    long long n_outer, L, n_inner, n_groups;
    int T, D, DL, RL;            // DL: chunks between consecutive passes; RL: ring length (multiple of FBP_U)
    double alpha, csf;
};

__device__ __forceinline__ void fbp_bar() { asm volatile("barrier.sync 1, 64;" ::: "memory"); }
__device__ __forceinline__ int fbp_mod(int pos, int RL)
{
    const int m = pos % RL;
    return m < 0 ? m + RL : m;
}
__device__ __forceinline__ int fbp_next(int slot, int RL)      // the slot one chunk further
{
    slot += FBP_U;
    return slot >= RL ? slot - RL : slot;
}

template <int NPASS>
struct FbSweepPGeom {
    static constexpr int LPL = NPASS <= 1 ? 2 : (NPASS <= 2 ? 4 : (NPASS <= 4 ? 8 : 16));   // lanes per line
    static constexpr int NL = 32 / LPL;                                                      // lines per CTA
};

template <int NPASS, int MODE>
__global__ void __launch_bounds__(64)
fb_sweepp_kernel(const FbSweepP p)
{
    constexpr int U = FBP_U;
    constexpr int LPL = FbSweepPGeom<NPASS>::LPL, NL = FbSweepPGeom<NPASS>::NL;
    constexpr int NS = NPASS + 1;                        // streams per line: the rows themselves, the output of every pass
    extern __shared__ __align__(16) double fbp_smem[];   // ring of (line l, stream s, field f): ((l * NS + s) * 2 + f) * rs
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long outer = blockIdx.x / p.n_groups;
    const long long line0 = (blockIdx.x - outer * p.n_groups) * NL;
    const int L = (int)p.L;
    const int T1 = p.T + 1, D = p.D, DL = p.DL, RL = p.RL;
    // consecutive rings start 2 doubles (4 banks) further: the 32 lanes of the chain warp read / write the same position of
    // 32 different rings (RL is a multiple of 16 doubles, i.e. of all 32 banks)
    const int rs = RL + 2;
    for (int i = threadIdx.x; i < NL * NS * 2 * rs; i += 64) fbp_smem[i] = 0.0;
    // Which of the two warps runs the chains: several CTAs share an SM, and with fixed roles their chain warps -- the ones that
    // need the fp64 pipe and the issue slots -- would sit on the same two of the four schedulers (warp slots are handed out in
    // pairs, scheduler = slot % 4) while the other two only see the light feeding warps.  The role follows the hardware
    // warp slot of warp 0 instead, so that the chain warps of up to four resident CTAs land on four different schedulers.
    // (%warpid is only a placement hint here; either assignment is correct.)
    __shared__ unsigned s_slot0;
    if (threadIdx.x == 0) {
        unsigned ws;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(ws));
        s_slot0 = ws;
    }
    __syncthreads();
    const bool chain_on_warp0 = ((s_slot0 >> 2) & 1u) == (s_slot0 & 1u);
    const bool chains = (wid == 0) == chain_on_warp0;
    // iterations: pass q handles stream positions U (it - q DL) .. + U-1; the last pass must reach position L - 1 + T1
    const int nit = (L + T1 + U - 1) / U + (NPASS - 1) * DL + 1;

    if (chains) {
        // ------------------------------ the chains ------------------------------
        const int f = lane & 1, q = (lane % LPL) >> 1, l = lane / LPL;
        const bool mine = q < NPASS;
        const int qq = mine ? q : 0;
        const double *xin = fbp_smem + (size_t)((l * NS + qq) * 2 + f) * rs;          // input stream of this pass
        double *xout = fbp_smem + (size_t)((l * NS + qq + 1) * 2 + f) * rs;           // its output stream
        const double alpha = p.alpha;
        double accu = 0.0, new0 = 0.0;
        // ring slots of the chunk's first new element (position s0), first old element (s0 - D), first result (s0 - T1)
        int bn = 0, bo = fbp_mod(-D, RL), bw = fbp_mod(-T1, RL);
        for (int it = 0; it < nit; ++it) {
            fbp_bar();                                    // barrier #it: chunk `it` of the rows has landed, iteration it-1 is complete
            const int s0 = (it - q * DL) * U;             // first stream position of this lane's chunk
            if (mine && s0 >= 0) {
                const int k0 = s0 - T1;                   // output position of the chunk's first step
                const bool interior = k0 >= 0 && k0 + U <= L;
                double x[U], o[U], r[U];
#pragma unroll
                for (int j = 0; j < U; j += 2) {
                    // s0, D and RL are even: pairs are 16-byte aligned and never straddle the end of the ring; the new
                    // chunk starts at a multiple of U and does not wrap, the old one may
                    const double2 a = *reinterpret_cast<const double2 *>(xin + bn + j);
                    int io = bo + j;
                    io = io >= RL ? io - RL : io;
                    const double2 b = *reinterpret_cast<const double2 *>(xin + io);
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
#pragma unroll
                    for (int j = 0; j < U; j += 2) {
                        int iw = bw + j;
                        iw = iw >= RL ? iw - RL : iw;
                        *reinterpret_cast<double2 *>(xout + iw) = make_double2(r[j], r[j + 1]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < U; ++j) {
                        int iw = bw + j;
                        iw = iw >= RL ? iw - RL : iw;
                        xout[iw] = r[j];
                    }
                }
                bn = fbp_next(bn, RL);
                bo = fbp_next(bo, RL);
                bw = fbp_next(bw, RL);
            }
        }
        fbp_bar();                                        // barrier #nit: the last iteration is complete
    } else {
        // ------------------------------ feed and drain ------------------------------
        const double *gin = p.in + (size_t)outer * p.L * p.n_inner * 2;
        double offset = 0.0;
        if (MODE == 2) offset = fb_field_offset(p.mm, outer);
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        // one chunk of rows: U x NL x 2 doubles, 8 bytes per copy; (line, field) fastest: 16 NL contiguous bytes per row
        // (the chunks are loaded and drained in order: their ring slots are running indices, no division in the loop)
        int ld_slot = 0;                                  // slot of the first row of the next chunk to load
        int dr_slot = fbp_mod(-(NPASS - 1) * DL * U - T1, RL);      // slot of the first row the last pass writes in iteration 0
        auto load_chunk = [&](int c) {
            const int slot = ld_slot;
            ld_slot = fbp_next(ld_slot, RL);
#pragma unroll
            for (int h = 0; h < U * NL * 2 / 32; ++h) {
                const int e = lane + 32 * h;
                const int f = e & 1, l = (e >> 1) % NL, j = e / (2 * NL);
                const int k = c * U + j;
                double *dst = fbp_smem + (size_t)((l * NS) * 2 + f) * rs + slot + j;
                if (k < L && line0 + l < p.n_inner) {
                    const unsigned d32 = (unsigned)__cvta_generic_to_shared(dst);
                    const double *src = gin + ((size_t)k * p.n_inner + line0 + l) * 2 + f;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d32), "l"(src) : "memory");
                } else {
                    *dst = 0.0;
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // the chunk the last pass wrote in iteration it_done: rows kb .. kb + U-1 of NL lines
        auto drain = [&](int it_done) {
            const int kb = (it_done - (NPASS - 1) * DL) * U - T1;
            const int slot = dr_slot;
            dr_slot = fbp_next(dr_slot, RL);
            if (kb + U <= 0 || kb >= L) return;
            constexpr int NE = U * NL / 32;               // points per lane
            double v[NE], w[NE];
            long long o[NE];                              // element offset of the point, -1: outside the grid
#pragma unroll
            for (int h = 0; h < NE; ++h) {
                const int e = lane + 32 * h;
                // MODE 1 writes rows of one line next to each other: row fastest; MODE 0 / 2 lines next to each other
                const int j = MODE == 1 ? e % U : e / NL;
                const int l = MODE == 1 ? e / U : e % NL;
                const int k = kb + j;
                int ir = slot + j;
                ir = ir >= RL ? ir - RL : ir;
                const double *yv = fbp_smem + (size_t)((l * NS + NPASS) * 2) * rs;
                v[h] = yv[ir];
                w[h] = yv[rs + ir];
                const bool inside = k >= 0 && k < L && line0 + l < p.n_inner;
                if (MODE == 1) o[h] = inside ? (long long)(((size_t)outer * p.n_inner + line0 + l) * p.L + k) : -1;
                else o[h] = inside ? (long long)(((size_t)outer * p.L + k) * p.n_inner + line0 + l) : -1;
            }
            if (MODE == 2) {
                // `wg[wg < csf] = nan` (interpolation.py:427-430); (vg / wg + offset) -> float32 (:367).  The divisions of
                // the lane's points run interleaved (fb_div_n); masked points are overwritten afterwards instead of being
                // divided by NaN
                double q[NE];
                bool masked[NE];
#pragma unroll
                for (int h = 0; h < NE; ++h) masked[h] = (w[h] < p.csf) || o[h] < 0;
                fb_div_n<NE>(v, w, q, masked);
#pragma unroll
                for (int h = 0; h < NE; ++h) {
                    if (o[h] < 0) continue;
                    const double qv = (w[h] < p.csf) ? qnan : __dadd_rn(q[h], offset);
                    p.out32[o[h]] = __double2float_rn(qv);
                    if (p.out64) p.out64[o[h]] = qv;
                }
            } else {
#pragma unroll
                for (int h = 0; h < NE; ++h)
                    if (o[h] >= 0) *reinterpret_cast<double2 *>(p.out + 2 * o[h]) = make_double2(v[h], w[h]);
            }
        };
        for (int c = 0; c < FBP_PD; ++c) load_chunk(c);
        for (int it = 0; it < nit; ++it) {
            asm volatile("cp.async.wait_group %0;" ::"n"(FBP_PD - 1) : "memory");         // chunk `it` of the rows has landed
            fbp_bar();                                    // barrier #it
            load_chunk(it + FBP_PD);
            if (it > 0) drain(it - 1);
        }
        fbp_bar();                                        // barrier #nit
        drain(nit - 1);
    }
}
