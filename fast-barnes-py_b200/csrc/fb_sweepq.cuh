// fb_sweepq.cuh -- fp64 axis sweep, second generation ("q" kernels): one warp owns 16 lines x 2 fields
// and runs ALL passes of the n-fold filter; rows travel with the bulk asynchronous copy engine (TMA).
//
// Replaces, like fb_sweep*_kernel, the line loops of _convolve_tail_{2,3}d (fastbarnes/interpolation.py:
// 398-479) with _accumulate_tail_array (:485-533) inlined, MODE 2 also the NaN mask (:427-430 / :475-479)
// and `(vg / wg + offset).astype(np.float32)` (:367).  The arithmetic (operation order, explicit *_rn
// intrinsics, zero extension instead of the five loop phases) is the one of fb_sweep_chunk in
// fb_kernels.cuh, so results are bit-identical to the reference and to the first-generation kernels.
//
// What is different from round 1's tensor-memory kernel (two warps per 16 lines in lock step, register-staged loads; removed):
//   * Grids are arrays of interleaved (value, weight) fp64 nodes in every stage.  Lane l of a warp owns
//     line l >> 1, field l & 1: a grid row of 16 lines is ONE contiguous 256-byte segment.
//   * Input rows are fetched by TMA (`cp.async.bulk.tensor.3d...mbarrier::complete_tx::bytes`, SASS UTMALDG)
//     through a tensor map over in[outer][k][2 * n_inner]: per chunk of 8 steps one elected lane requests
//     a box of the 8 newest rows and a box of the 8 rows D = 2T+2 steps older (pass 1 re-reads them from
//     L2 instead of keeping a ring), NST chunks ahead, into a per-warp staging ring in shared memory; an
//     mbarrier per stage counts the bytes.  No address arithmetic, no load registers and no scoreboard
//     stalls in the arithmetic loop, and the depth of the prefetch is a launch parameter instead of a
//     register budget.  Rows beyond the line ends (and lines beyond the grid) arrive as zeros (the TMA
//     unit fills out-of-bounds elements): the reference's zero extension needs no load masks.
//   * One warp runs all NPASS passes: no hand-over ring, no pair barrier, no lock step.  The rings of
//     passes 2..NPASS (the element each pass subtracts D steps later) live in tensor memory (NT rings:
//     one tcgen05.ld/st.32x32b.x16 = 8 slots per ring and chunk) and, when tensor memory is full, in
//     shared memory (NS rings).  A ring has R = roundup8(D) slots plus, when D is no multiple of 8, an
//     8-slot mirror of its first slots, so that the 8 slots a chunk reads never wrap.
//   * MODE 1 (transposing x sweep) stores its 16-line x 8-step output tile with one TMA tensor store
//     (`cp.async.bulk.tensor.3d.global.shared::cta`, SASS UTMASTG) from a 128-byte-swizzled tile: 16 row
//     segments of 128 bytes; the unit clips at the grid borders.  MODE 0 stores 8 rows of 256 bytes the
//     same way, MODE 2 divides in lane pairs (l, l ^ 1) and stores float32.
//   * A CTA is 8 (or 4) independent warps and owns the SM (launch bounds 256 x 1, up to 255 registers per
//     thread); it allocates all 512 tensor-memory columns once.  Warps claim 16-line groups from a
//     global counter.
#pragma once
#include "fb_kernels.cuh"
#include <cuda.h>   // CUtensorMap (types only: the encoder is looked up at run time)
#include "fb_sparse.cuh"

#define FBQ_U 8                 // steps per chunk
#define FBQ_STAGE_BYTES 4096    // one staging slot: 8 new rows + 8 old rows of 256 bytes (16 lines x 2 fields x 8 B)
#define FBQ_TILE_BYTES 2048     // output tile of a warp (MODE 0: 8 rows x 256 B; MODE 1: 16 lines x 128 B, 128B-swizzled)
#define FBQ_MAX_STAGES 8
#ifndef FBQ_ST_X2
#define FBQ_ST_X2 1            // tensor-memory ring stores: 0 = one x16 store per ring and chunk (16 packing moves), 1 = eight x2 stores (measured: x 1.17 -> 1.14 ms, y 1.26 -> 1.20 ms)
#endif

struct FbSweepQ {
    const double *in;            // interleaved nodes: in[((outer * L + k) * n_inner + inner) * 2 + field]
    double *out;                 // MODE 0: same index space (may be `in`); MODE 1: out[((outer * n_inner + inner) * L + k) * 2 + field]
    float *out32;                // MODE 2: out32[(outer * L + k) * n_inner + inner]
    double *out64;               // MODE 2, optional: the fp64 quotient, same index space
    const unsigned long long *mm;
    long long n_outer, L, n_inner, n_groups;
    int T, D, R, RP;             // RP: physical ring slots (R, or R + 8 with the mirror)
    double alpha, csf;
    unsigned long long *work_counter;
    int nst;                     // staging slots per warp (chunks in flight), 2..FBQ_MAX_STAGES; a slot = 8 new + 8 old rows
    int pf;                      // extra chunks of lead of an L2 prefetch of the new rows (0: none)
    int tmem_cols_per_warp;      // tensor-memory columns of one warp (two warps share a lane quarter when 8 warps run)
    int tmem_alloc_cols;         // columns the CTA allocates (power of two >= 32; 512 when more than 4 warps run)
    int smem_per_warp;           // bytes of a warp's block of dynamic shared memory: [mbarriers 128 B][stages][rings]
    int off_ring;                // byte offset of the rings inside that block (stages start at 128)
    // fb_sweepqs_kernel (rows from the binned samples, fb_sparse.cuh)
    int ncw;                     // pass warps per CTA (the producer warps follow them)
    unsigned zero;               // always 0 (a value the compiler cannot fold: orders the early TMA request behind the loads)
    int nb;                      // buckets (chunks of 8 rows) per line group
    const unsigned int *bin_start;   // [n_outer * n_groups * nb + 1] first entry of every bucket
    int prod_bytes;                  // bytes of a pass warp's share of the producer area (mbarriers, bucket table, prefetch slots)
    const FbRec *nodes;              // entries: key = (row << 4 | line in group) or FB_BIN_HOLE, w / wv = the node's sums
};

__device__ __forceinline__ unsigned fbq_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void fbq_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fbq_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fbq_mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FBQ_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra FBQ_DONE;\n"
        "bra FBQ_WAIT;\n"
        "FBQ_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// wait of a warp that has nothing else to do (a producer that is ahead of its pass warp): try_wait with a suspend-time
// hint, so that the warp sleeps in hardware instead of re-issuing the test (the plain try_wait loop of eight producer
// warps took a quarter of all issue slots of the kernel)
__device__ __forceinline__ void fbq_mbar_wait_idle(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FBQ_IWAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra FBQ_IDONE;\n"
        "bra FBQ_IWAIT;\n"
        "FBQ_IDONE:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(100000u) : "memory");
}
// The requests of one chunk, by one elected lane of the (converged) warp: arm the stage's mbarrier with the bytes of
// both boxes, then two TMA tensor loads of a box of 32 doubles x 8 rows: rows c1n .. c1n+7 to dst, rows c1o .. c1o+7
// to dst + 2 KB.  The election sits inside the asm so that the compiler sees one convergent statement (a C++ `if
// (lane == 0)` around it costs an ELECT / BRA.U.ANY loop per instruction).
__device__ __forceinline__ void fbq_issue_chunk(unsigned dst, const CUtensorMap *tm, int c0, int c1n, int c1o, int c2, unsigned bar,
                                                int enable)
{
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .b32 d2;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "setp.ne.b32 q, %7, 0;\n"
        "and.pred p, p, q;\n"
        "add.u32 d2, %0, 2048;\n"
        "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%6], 4096;\n"
        "@p cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %5}], [%6];\n"
        "@p cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [d2], [%1, {%2, %4, %5}], [%6];\n"
        "}\n" ::"r"(dst), "l"(tm), "r"(c0), "r"(c1n), "r"(c1o), "r"(c2), "r"(bar), "r"(enable) : "memory");
}
__device__ __forceinline__ void fbq_tma_prefetch(const CUtensorMap *tm, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// TMA tensor store of one box from shared memory by one elected lane + commit of its bulk async-group (every lane
// commits: an empty group for the others); out-of-bounds parts of the box are clipped
__device__ __forceinline__ void fbq_tma_store(const CUtensorMap *tm, int c0, int c1, int c2, unsigned src)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "@p cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];\n"
        "cp.async.bulk.commit_group;\n"
        "}\n" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(src) : "memory");
}
__device__ __forceinline__ void fbq_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void fbq_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fbq_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fbq_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void fbq_tmem_st2(unsigned taddr, unsigned r0, unsigned r1)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(r0), "r"(r1) : "memory");
}

__device__ __forceinline__ void fbq_mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fbq_sts4(unsigned addr, int a, int b, int c, int d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void fbq_lds4(unsigned addr, int &a, int &b, int &c, int &d)
{
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}

__device__ __forceinline__ double fbq_lds(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void fbq_sts(unsigned addr, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// ------------------------------------------------------------------------------------------
// The 8 steps of one chunk.  bn / bo: newest / D-steps-older inputs of pass 1.  Ring q (q = 0 .. NR-1)
// delays the input stream of pass q + 2; rings 0 .. NS-1 live in shared memory (sr / sw: this lane's
// byte address of slot rslot / wslot of ring 0, ring pitch rp_bytes), rings NS .. NR-1 in tensor memory
// (tr / tw: address of slot rslot / wslot of the first of them, ring pitch 2 * RP columns).
// mirror: wslot == 0 and the ring has a mirror -> the chunk is also written behind slot R.
template <int NPASS, int NS, bool MASKED>
__device__ __forceinline__ void fbq_chunk(const double (&bn)[FBQ_U], const double (&bo)[FBQ_U], double (&accu)[NPASS],
                                          double (&new0)[NPASS], double (&xs)[FBQ_U], unsigned sr, unsigned sw, unsigned rp_bytes,
                                          unsigned tr, unsigned tw, unsigned rp_cols, bool mirror, unsigned mirror_slots,
                                          int t, int T1, int L, double alpha)
{
    constexpr int U = FBQ_U;
    constexpr int NR = NPASS - 1;
    constexpr int NT = NR - NS;
    unsigned oldt[NT > 0 ? NT : 1][16];
    double olds[NS > 0 ? NS : 1][U];
    if constexpr (NT > 0) {
#pragma unroll
        for (int q = 0; q < NT; ++q) fb_tmem_ld16(tr + (unsigned)q * rp_cols, oldt[q]);
    }
    if constexpr (NS > 0) {
#pragma unroll
        for (int q = 0; q < NS; ++q)
#pragma unroll
            for (int j = 0; j < U; ++j) olds[q][j] = fbq_lds(sr + (unsigned)q * rp_bytes + (unsigned)j * 256u);
    }
    if constexpr (NT > 0) {
#pragma unroll
        for (int q = 0; q < NT; ++q) fb_tmem_wait_ld16(oldt[q]);
    }
    unsigned newt[NT > 0 ? NT : 1][16];
    double news[NS > 0 ? NS : 1][U];
    if constexpr (!MASKED) {
#pragma unroll
        for (int j = 0; j < U; ++j) {
            double x = bn[j];
#pragma unroll
            for (int q = 0; q < NPASS; ++q) {
                double o;
                if (q == 0) {
                    o = bo[j];
                } else if (q - 1 < NS) {
                    o = olds[q - 1 < NS ? q - 1 : 0][j];
                    news[q - 1 < NS ? q - 1 : 0][j] = x;
                } else {
                    const int r = q - 1 - NS;
                    o = __hiloint2double((int)oldt[r >= 0 ? r : 0][2 * j + 1], (int)oldt[r >= 0 ? r : 0][2 * j]);
                    newt[r >= 0 ? r : 0][2 * j] = (unsigned)__double2loint(x);
                    newt[r >= 0 ? r : 0][2 * j + 1] = (unsigned)__double2hiint(x);
                }
                // interpolation.py:512-514 with zero extension: accu += in[k+T] - in[k-T-1];
                // out[k] = accu + alpha * (in[k-T-1] + in[k+T+1])
                const double d = __dsub_rn(new0[q], o);
                accu[q] = __dadd_rn(accu[q], d);
                const double r = __dadd_rn(accu[q], __dmul_rn(alpha, __dadd_rn(o, x)));
                new0[q] = x;
                x = r;
            }
            xs[j] = x;
        }
    } else {
        // line ends, pass by pass: the same operations; results at positions outside [0, L) are replaced by zeros
        // (zero extension of every pass's output), and a pass whose input has not begun (start of the line) or
        // whose output positions all lie beyond the line (end of the line) is skipped -- its result is zeros.
        double xin[U];
#pragma unroll
        for (int j = 0; j < U; ++j) xin[j] = bn[j];
#pragma unroll
        for (int q = 0; q < NPASS; ++q) {
            if (q >= 1) {
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    if (q - 1 < NS) {
                        news[q - 1 < NS ? q - 1 : 0][j] = xin[j];
                    } else {
                        const int r = q - 1 - NS;
                        newt[r >= 0 ? r : 0][2 * j] = (unsigned)__double2loint(xin[j]);
                        newt[r >= 0 ? r : 0][2 * j + 1] = (unsigned)__double2hiint(xin[j]);
                    }
                }
            }
            const int k0 = t - (q + 1) * T1;             // output position of step 0 of this pass
            const bool active = (t + U - 1 >= q * T1) && (k0 < L);
            if (active) {
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    double o;
                    if (q == 0) {
                        o = bo[j];
                    } else if (q - 1 < NS) {
                        o = olds[q - 1 < NS ? q - 1 : 0][j];
                    } else {
                        const int r = q - 1 - NS;
                        o = __hiloint2double((int)oldt[r >= 0 ? r : 0][2 * j + 1], (int)oldt[r >= 0 ? r : 0][2 * j]);
                    }
                    const double x = xin[j];
                    const double d = __dsub_rn(new0[q], o);
                    accu[q] = __dadd_rn(accu[q], d);
                    const double r = __dadd_rn(accu[q], __dmul_rn(alpha, __dadd_rn(o, x)));
                    new0[q] = x;
                    const int k = k0 + j;
                    xin[j] = (k >= 0 && k < L) ? r : 0.0;
                }
            } else {
#pragma unroll
                for (int j = 0; j < U; ++j) xin[j] = 0.0;
            }
        }
#pragma unroll
        for (int j = 0; j < U; ++j) xs[j] = xin[j];
    }
    if constexpr (NS > 0) {
#pragma unroll
        for (int q = 0; q < NS; ++q)
#pragma unroll
            for (int j = 0; j < U; ++j) fbq_sts(sw + (unsigned)q * rp_bytes + (unsigned)j * 256u, news[q][j]);
        if (mirror) {
#pragma unroll
            for (int q = 0; q < NS; ++q)
#pragma unroll
                for (int j = 0; j < U; ++j)
                    fbq_sts(sw + (unsigned)q * rp_bytes + mirror_slots * 256u + (unsigned)j * 256u, news[q][j]);
        }
    }
    if constexpr (NT > 0) {
#if FBQ_ST_X2
        // one 64-bit store per slot: the operands are the register pairs the passes produced (no packing moves)
#pragma unroll
        for (int q = 0; q < NT; ++q)
#pragma unroll
            for (int j = 0; j < U; ++j) fbq_tmem_st2(tw + (unsigned)q * rp_cols + 2u * (unsigned)j, newt[q][2 * j], newt[q][2 * j + 1]);
        if (mirror) {
#pragma unroll
            for (int q = 0; q < NT; ++q)
#pragma unroll
                for (int j = 0; j < U; ++j)
                    fbq_tmem_st2(tw + (unsigned)q * rp_cols + 2u * mirror_slots + 2u * (unsigned)j, newt[q][2 * j], newt[q][2 * j + 1]);
        }
#else
#pragma unroll
        for (int q = 0; q < NT; ++q) fb_tmem_st16(tw + (unsigned)q * rp_cols, newt[q]);
        if (mirror) {
#pragma unroll
            for (int q = 0; q < NT; ++q) fb_tmem_st16(tw + (unsigned)q * rp_cols + 2u * mirror_slots, newt[q]);
        }
#endif
    }
}

// Finalisation of 8 consecutive rows held by lane pairs (even lane: vg, odd lane: wg of the same line):
// `wg[wg < csf] = nan; (vg / wg + offset).astype(float32)` (interpolation.py:427-430, :367).  The lanes of a
// pair trade one operand per row pair: the even lane divides rows kb, kb+2, .., the odd lane rows kb+1, kb+3, ..
// o32 / o64 point at this lane's first row, its m-th row lies 2 * m * row_elems elements further; bit m of
// `rows` = that row exists (and the lane's line lies inside the grid).
__device__ __forceinline__ void fbq_finalize_chunk(const double (&xs)[FBQ_U], int fld, double csf, double offset, float *o32,
                                                   double *o64, unsigned row_elems, unsigned rows)
{
    constexpr int U = FBQ_U;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    double va[U / 2], wa[U / 2], qa[U / 2];
    bool masked[U / 2];
#pragma unroll
    for (int j = 0; j < U; j += 2) {
        const double send = fld ? xs[j] : xs[j + 1];
        const double recv = __shfl_xor_sync(0xffffffffu, send, 1);
        va[j / 2] = fld ? recv : xs[j];
        const double ww = fld ? xs[j + 1] : recv;
        masked[j / 2] = ww < csf;
        wa[j / 2] = ww;
    }
    fb_div_n<U / 2>(va, wa, qa, masked);
    double q[U / 2];
#pragma unroll
    for (int j = 0; j < U / 2; ++j) q[j] = masked[j] ? qnan : __dadd_rn(qa[j], offset);
    if (rows == 15u) {
#pragma unroll
        for (int j = 0; j < U / 2; ++j) *fb_row(o32, 2u * j, row_elems * 4u) = __double2float_rn(q[j]);
    } else {
#pragma unroll
        for (int j = 0; j < U / 2; ++j)
            if (rows & (1u << j)) *fb_row(o32, 2u * j, row_elems * 4u) = __double2float_rn(q[j]);
    }
    if (o64) {                                           // uniform: the fp64 quotient is an optional output
#pragma unroll
        for (int j = 0; j < U / 2; ++j)
            if (rows & (1u << j)) *fb_row(o64, 2u * j, row_elems * 8u) = q[j];
    }
}

template <int NPASS, int NS, int MODE>
__global__ void __launch_bounds__(256, 1)
fb_sweepq_kernel(const FbSweepQ p, const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out)
{
    constexpr int U = FBQ_U;
    constexpr int NR = NPASS - 1;
    constexpr int NT = NR - NS;
#ifdef FBQ_NO_EARLY
    constexpr bool EARLY_ISSUE = false;
#else
    constexpr bool EARLY_ISSUE = MODE != 2;
#endif
    static_assert(NS >= 0 && NT >= 0, "ring split");
    extern __shared__ __align__(1024) unsigned char fbq_smem[];
    __shared__ unsigned s_tmem_base;

    const int lane = threadIdx.x & 31;
    // warp index through redux.sync: the result lives in a uniform register, so everything derived from it (shared-memory
    // addresses of the TMA operations) is uniform for the compiler and UTMALDG / UTMASTG need no per-lane loop
    const int wid = __reduce_max_sync(0xffffffffu, (int)(threadIdx.x >> 5));
    const int nwarps = blockDim.x >> 5;

    if constexpr (NT > 0) {
        if (wid == 0) {
            const unsigned dst = fbq_smem_addr(&s_tmem_base);
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"((unsigned)p.tmem_alloc_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    // dynamic shared memory: [output tiles, 2 KB per warp, 1 KB aligned (MODE 0 / 1)][per warp: mbarriers 128 B, stages, rings]
    const unsigned sm0 = fbq_smem_addr(fbq_smem);       // 1 KB aligned by declaration (the swizzled tiles need it)
    const unsigned tile = sm0 + (unsigned)wid * FBQ_TILE_BYTES;
    const unsigned wsm = sm0 + (MODE == 2 ? 0u : (unsigned)nwarps * FBQ_TILE_BYTES) + (unsigned)wid * (unsigned)p.smem_per_warp;
    const unsigned bars = wsm;
    const int nst = p.nst;
    const unsigned stages = wsm + 128u;                                    // nst slots of 4 KB
    const unsigned ring_s = wsm + (unsigned)p.off_ring + (unsigned)lane * 8u;
    if (lane == 0) {
        // one mbarrier per chunk in flight; a phase = the two boxes of a chunk (new rows, old rows), 2 KB each
        for (int s = 0; s < nst; ++s) fbq_mbar_init(bars + 8u * (unsigned)s, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fbq_fence_async();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned tring = 0;
    if constexpr (NT > 0) {
        // lane quarter wid & 3; warps wid and wid + 4 split the columns of their quarter
        tring = s_tmem_base + ((unsigned)((wid & 3) * 32) << 16) + (unsigned)((wid >> 2) * p.tmem_cols_per_warp);
    }

    const int L = (int)p.L, T1 = p.T + 1, D = p.D, R = p.R, RP = p.RP;
    const unsigned rp_bytes = (unsigned)RP * 256u, rp_cols = 2u * (unsigned)RP;
    const bool has_mirror = RP > R;
    const double alpha = p.alpha;
    const int fld = lane & 1;
    const int lag = NPASS * T1;
    const int t_begin = -((U - lag % U) % U);            // (t - lag) % U == 0 at chunk starts
    const int t_end = L + lag;
    const int nchunks = (t_end - t_begin + U - 1) / U;
    const long long n_items = p.n_outer * p.n_groups;
    unsigned phases = 0;                                 // bit s: parity the next wait on stage s expects
    // this lane's addresses inside the output tile
    //   MODE 0: row j of 256 bytes, lane-th double.   MODE 1: line (lane >> 1) = 128-byte row, step j = 16-byte chunk
    //   j ^ (line & 7) (the 128-byte swizzle of the tensor map), field = 8-byte half
    const unsigned tile_lane = MODE == 1 ? tile + (unsigned)(lane >> 1) * 128u + (unsigned)fld * 8u : tile + (unsigned)lane * 8u;
    const unsigned tile_xor = MODE == 1 ? (unsigned)((lane >> 1) & 7) << 4 : 0u;

#pragma unroll 1
    for (;;) {
        // claim a 16-line group; item, outer and group go through redux.sync as well (uniform registers)
        int it = 0, ou = 0, gr = 0;
        if (lane == 0) {
            const unsigned long long claimed = atomicAdd(p.work_counter, 1ull);
            it = claimed < (unsigned long long)n_items ? (int)claimed : 0x7fffffff;
            ou = (int)((long long)it / p.n_groups);
            gr = (int)((long long)it - (long long)ou * p.n_groups);
        }
        if (__reduce_max_sync(0xffffffffu, it) == 0x7fffffff) break;
        const int outer = __reduce_max_sync(0xffffffffu, ou);
        const int group = __reduce_max_sync(0xffffffffu, gr);
        const long long inner = (long long)group * 16 + (lane >> 1);

        // rings start as zeros (zero extension to the left of the line)
        if constexpr (NT > 0) {
            unsigned z[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = 0u;
            for (int i = 0; i < NT * RP; i += 8) fb_tmem_st16(tring + 2u * (unsigned)i, z);
            fb_tmem_wait_st();
        }
        if constexpr (NS > 0) {
            for (int i = 0; i < NS * RP; ++i) fbq_sts(ring_s + (unsigned)i * 256u, 0.0);
        }
        __syncwarp();

        // requests of the chunk at stream position tt: the 8 newest rows tt .. tt+7 and the 8 rows tt-D .. tt-D+7
        // (rows outside [0, L) arrive as zeros) into staging slot sl
        auto issue = [&](int tt, int sl, int enable) {
            fbq_issue_chunk(stages + (unsigned)sl * FBQ_STAGE_BYTES, &tm_in, group * 32, tt, tt - D, outer, bars + 8u * (unsigned)sl,
                            enable);
            if (p.pf > 0 && lane == 0 && enable && tt + p.pf * U < L) fbq_tma_prefetch(&tm_in, group * 32, tt + p.pf * U, outer);
        };

        for (int c = 0; c < nst; ++c) issue(t_begin + c * U, c, c < nchunks);

        double accu[NPASS], new0[NPASS];
#pragma unroll
        for (int q = 0; q < NPASS; ++q) { accu[q] = 0.0; new0[q] = 0.0; }
        double offset = 0.0;
        if (MODE == 2) offset = fb_field_offset(p.mm, outer);

        // output of the 8 rows kb .. kb+7 that left the last pass one chunk ago (kb is a multiple of 8, 0 <= kb < L;
        // all_rows: kb + 7 < L)
        auto emit = [&](const double (&x)[U], int kb, bool all_rows) {
            if (MODE == 0 || MODE == 1) {
                fbq_bulk_wait_read0();                               // the previous tile has been read (whichever lane stored it)
                __syncwarp();
#pragma unroll
                for (int j = 0; j < U; ++j)
                    fbq_sts(MODE == 1 ? tile_lane + (((unsigned)j << 4) ^ tile_xor) : tile_lane + (unsigned)j * 256u, x[j]);
                fbq_fence_async();
                __syncwarp();
                if (MODE == 1) fbq_tma_store(&tm_out, 2 * kb, group * 16, outer, tile);
                else fbq_tma_store(&tm_out, group * 32, kb, outer, tile);
            } else {
                unsigned rows = 0;
                if (inner < p.n_inner) {
                    if (all_rows) {
                        rows = 15u;
                    } else {
#pragma unroll
                        for (int m = 0; m < U / 2; ++m)
                            if (kb + 2 * m + fld < L) rows |= 1u << m;
                    }
                }
                const long long o = ((long long)outer * p.L + kb + fld) * p.n_inner + inner;
                fbq_finalize_chunk(x, fld, p.csf, offset, p.out32 + o, p.out64 ? p.out64 + o : nullptr, (unsigned)p.n_inner, rows);
            }
        };

        int wslot = 0;                                   // ring write slot of this chunk (multiple of 8)
        int rslot = (R - D % R) % R;                     // ring read slot: (wslot - D) mod R
        int slot = 0;                                    // staging slot of this chunk
        int t = t_begin;
        // The loop is rotated: iteration c runs the passes of chunk c AND writes out the rows chunk c-1 produced, in one
        // basic block on the fast path, so that the divisions / tile stores of the output (long dependency chains, few
        // instructions) overlap with the dense fp64 stream of the passes.
        double xsp[U];
#pragma unroll
        for (int j = 0; j < U; ++j) xsp[j] = 0.0;
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c, t += U) {
            const unsigned stn = stages + (unsigned)slot * FBQ_STAGE_BYTES + (unsigned)lane * 8u;
            const unsigned sto = stn + 2048u;
            fbq_mbar_wait(bars + 8u * (unsigned)slot, (phases >> slot) & 1u);
            phases ^= 1u << slot;
            double bn[U], bo[U], xs[U];
#pragma unroll
            for (int j = 0; j < U; ++j) bn[j] = fbq_lds(stn + (unsigned)j * 256u);
#pragma unroll
            for (int j = 0; j < U; ++j) bo[j] = fbq_lds(sto + (unsigned)j * 256u);
            if constexpr (EARLY_ISSUE) {
                // the staging slot is free as soon as every lane holds its rows in registers: request the chunk nst chunks
                // ahead now, one chunk time earlier than at the end of the iteration (the memory-bound sweeps gain 4.5 %,
                // the finalising sweep loses 2.7 % and keeps the late request).  `dep` is zero -- p.zero is a launch
                // parameter that is always 0, which the compiler cannot know -- but only known once EVERY load of every lane
                // has returned: the request is data dependent on all of them.  (A first version derived the zero from the
                // last two loads alone, with arithmetic the assembler could fold: the request then overtook loads still
                // in flight, and small 3D volumes came out wrong in a third of the runs -- found by tools/fuzz_parity.py.)
                unsigned acc = 0u;
#pragma unroll
                for (int j = 0; j < U; ++j) acc ^= (unsigned)__double2hiint(bn[j]) ^ (unsigned)__double2hiint(bo[j]);
                const unsigned dep = __reduce_or_sync(0xffffffffu, acc & p.zero);
                issue(t + nst * U, slot, (int)(c + nst < nchunks) + (int)dep);
            }
            const unsigned sr = ring_s + (unsigned)rslot * 256u, sw = ring_s + (unsigned)wslot * 256u;
            const unsigned tr = tring + 2u * (unsigned)rslot, tw = tring + 2u * (unsigned)wslot;
            const bool mirror = has_mirror && wslot == 0;
            // fast path: every pass position of this chunk and every row of the previous chunk's output inside the line
            const bool fast = (t - U >= lag) && (t + U - 1 - T1 < L);
            const int kbp = t - U - lag;                 // rows kbp .. kbp+7 left the last pass in the previous chunk
            if (fast) {
                fbq_chunk<NPASS, NS, false>(bn, bo, accu, new0, xs, sr, sw, rp_bytes, tr, tw, rp_cols, mirror, (unsigned)R, t, T1, L, alpha);
                emit(xsp, kbp, true);
            } else {
                fbq_chunk<NPASS, NS, true>(bn, bo, accu, new0, xs, sr, sw, rp_bytes, tr, tw, rp_cols, mirror, (unsigned)R, t, T1, L, alpha);
                if (kbp >= 0 && kbp < L) emit(xsp, kbp, false);
            }
            // ---- the staging slot is free: request the chunk nst chunks ahead
            if constexpr (!EARLY_ISSUE) {
                __syncwarp();
                if constexpr (NPASS == 1) {
                    // with rings, the ring stores above carry the values of every staged row and cannot issue before the
                    // loads have returned; a single pass has no ring, so the request is tied to the loads explicitly
                    unsigned acc = 0u;
#pragma unroll
                    for (int j = 0; j < U; ++j) acc ^= (unsigned)__double2hiint(bn[j]) ^ (unsigned)__double2hiint(bo[j]);
                    const unsigned dep = __reduce_or_sync(0xffffffffu, acc & p.zero);
                    issue(t + nst * U, slot, (int)(c + nst < nchunks) + (int)dep);
                } else {
                    issue(t + nst * U, slot, c + nst < nchunks);
                }
            }
            if constexpr (NT > 0) fb_tmem_wait_st();
#pragma unroll
            for (int j = 0; j < U; ++j) xsp[j] = xs[j];
            wslot += U; wslot = (wslot >= R) ? 0 : wslot;
            rslot += U; rslot = (rslot >= R) ? rslot - R : rslot;
            ++slot; slot = (slot == nst) ? 0 : slot;
        }
        {
            const int kbp = t - U - lag;                 // the rows of the last chunk
            if (kbp >= 0 && kbp < L) emit(xsp, kbp, false);
        }
    }   // persistent loop

    if (MODE == 0 || MODE == 1) fbq_bulk_wait0();
    if constexpr (NT > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (wid == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem_base), "r"((unsigned)p.tmem_alloc_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// Transposing x sweep fed by the binned samples (fb_sparse.cuh) instead of a dense injection grid.
//
// The pass warps are those of fb_sweepq_kernel<.., 1> (same arithmetic, rings, transposing TMA store); what differs is
// where the rows come from.  One producer warp per pass warp of node entries (sorted
// by row) with two cursors -- the rows t .. t+7 and the rows D = 2T+2 further back that pass 1 subtracts -- and writes
// them as 2 KB tiles of zeros with the nodes scattered in (st.shared), NST chunks ahead.  Per staging slot an mbarrier
// `full` (16 arrivals: the lanes of the producer's half warp) and `empty` (32 arrivals: the pass warp's lanes, right after it has loaded
// the slot into registers) do the hand-over; a 16-byte header per slot carries {outer, group, stream position, flags}, so
// that the pass warps are driven by what arrives (the producers claim the work items).  93 % of the bench grid is zeros:
// the 16 B per grid point the dense sweep reads, and the 16 B per grid point the zero-fill writes, are never moved.
//   a pass warp's block of shared memory: +0 full[8]   +64 empty[8]   +128 headers[8 x 16 B]   +256 stages   rings
#define FBQS_FULL 0u
#define FBQS_EMPTY 64u
#define FBQS_HEAD 128u
#define FBQS_STAGES 256u
#define FBQS_FLAG_STOP 1
#define FB_BIN_END 0xfffffffeu

// A producer warp serves two pass warps AT THE SAME TIME: lanes 0-15 work for the first, lanes 16-31 for the second
// (all state is per lane and uniform inside a half warp).  Per pass warp it keeps, in shared memory,
//   * the line group's row of bin_start (first entry of every bucket; one bulk copy per line group), and
//   * FBQS_PF prefetch slots of 2 x 16 entries (FbRec, 32 bytes): the entries of the rows a .. a+7 are those of one bucket
//     (or of two neighbouring ones when the window straddles a bucket boundary), i.e. ONE contiguous range of the record
//     array, fetched with `cp.async.bulk` one chunk ahead and counted in bytes on the slot's mbarrier.
// So nothing is searched and no load result waits in a register across a loop iteration (register prefetches of this
// kind stalled on the scoreboard they share with the loads issued a moment ago: the wait for an old load is a wait for
// all of them).  The scatter of a window is one predicated 16-byte shared-memory store per lane; windows with more than
// 16 entries (clustered observations) take the rest with plain loads.
#define FBQS_PF 2                 // prefetch slots per pass warp
#define FBQS_PF_BYTES 1024u       // one slot: 16 entries of the new rows + 16 entries of the old rows

__device__ __forceinline__ void fbq_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar) : "memory");
}

// entry range of the rows a .. a+7 from the line group's bucket table in shared memory
__device__ __forceinline__ void fbq_window_range(unsigned table, int nb, int a, int t_begin, unsigned &s0, unsigned &s1)
{
    const int rel = a - t_begin;                         // multiple of 8 for the new rows; any value for the old rows
    int m0 = rel >= 0 ? rel >> 3 : -((-rel + 7) >> 3);   // floor(rel / 8)
    int m1 = m0 + ((rel & 7) ? 2 : 1);                   // one bucket behind the last one touched
    m0 = m0 < 0 ? 0 : m0;
    m1 = m1 > nb ? nb : m1;
    s0 = s1 = 0u;
    if (m0 < m1) {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(s0) : "r"(table + 4u * (unsigned)m0) : "memory");
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(s1) : "r"(table + 4u * (unsigned)m1) : "memory");
    }
}

// scatter the nodes of rows a .. a+7 into the (zeroed) tile: row r at 256 * (r - a), line y at 16 * y, (vg, wg) as a pair.
// The first 16 entries of [s0, s1) sit in the prefetch slot `pf`, further ones are read from global memory.
__device__ __forceinline__ void fbq_window_scatter(const FbSweepQ &p, unsigned s0, unsigned s1, unsigned pf, int a, unsigned tile, int hl)
{
    for (unsigned base = s0; base < s1; base += 16u) {   // uniform inside the half warp
        const unsigned idx = base + (unsigned)hl;
        unsigned pos = FB_BIN_END, pad;
        double v = 0.0, w = 0.0;
        if (idx < s1) {
            if (base == s0) {
                unsigned wlo, whi;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(pos), "=r"(pad), "=r"(wlo), "=r"(whi) : "r"(pf + 32u * (unsigned)hl) : "memory");
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(pf + 32u * (unsigned)hl + 16u) : "memory");
                w = __hiloint2double((int)whi, (int)wlo);
            } else {
                const FbRec *r = p.nodes + idx;
                pos = r->key;
                w = r->w;
                v = r->wv;
            }
        }
        const int row = (int)(pos >> 4);
        if (pos < FB_BIN_END && row >= a && row < a + FBQ_U) {       // not a hole, not beyond the range, inside the window
            const unsigned addr = tile + (unsigned)(row - a) * 256u + (pos & 15u) * 16u;
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v), "d"(w) : "memory");
        }
    }
}

template <int NPASS, int NS>
__global__ void __launch_bounds__(512, 1)
fb_sweepqs_kernel(const FbSweepQ p, const __grid_constant__ CUtensorMap tm_out)
{
    constexpr int U = FBQ_U;
    constexpr int MODE = 1;
    constexpr int NR = NPASS - 1;
    constexpr int NT = NR - NS;
    extern __shared__ __align__(1024) unsigned char fbq_smem[];
    __shared__ unsigned s_tmem_base;

    const int lane = threadIdx.x & 31;
    const int wid = __reduce_max_sync(0xffffffffu, (int)(threadIdx.x >> 5));
    const int ncw = p.ncw;                               // pass warps (8: two warpgroups); as many producer warps follow

    if constexpr (NT > 0) {
        if (wid == 0) {
            const unsigned dst = fbq_smem_addr(&s_tmem_base);
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"((unsigned)p.tmem_alloc_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    const unsigned sm0 = fbq_smem_addr(fbq_smem);
    const unsigned sm_blocks = sm0 + (unsigned)ncw * FBQ_TILE_BYTES;
    const int nst = p.nst;
    if (lane == 0 && wid < ncw) {
        const unsigned blk = sm_blocks + (unsigned)wid * (unsigned)p.smem_per_warp;
        for (int s = 0; s < nst; ++s) {
            fbq_mbar_init(blk + FBQS_FULL + 8u * (unsigned)s, 16u);      // the 16 lanes of the producer's half warp
            fbq_mbar_init(blk + FBQS_EMPTY + 8u * (unsigned)s, 32u);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fbq_fence_async();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const int L = (int)p.L, T1 = p.T + 1, D = p.D, R = p.R, RP = p.RP;
    const int lag = NPASS * T1;
    const int t_begin = -((U - lag % U) % U);
    const int t_end = L + lag;
    const int nchunks = (t_end - t_begin + U - 1) / U;
    const long long n_items = p.n_outer * p.n_groups;

    if (wid >= ncw) {
        // ================================ producer warp ================================
        // producer warp ncw + i feeds pass warp i; its two half warps take alternate chunks (half h writes staging slot
        // h: the pass warp consumes the slots 0, 1, 0, 1, .. across line groups), so two chunks are in production at a time
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;" ::: "memory");
        const int h = lane >> 4, hl = lane & 15;
        const unsigned hmask = 0xffffu << (16 * h);
        const int client = wid - ncw;
        const unsigned blk = sm_blocks + (unsigned)client * (unsigned)p.smem_per_warp;
        // producer area behind the pass warps' blocks, per pass warp: [mbarriers: table, prefetch slots][bucket table][slots]
        const unsigned parea = sm_blocks + (unsigned)ncw * (unsigned)p.smem_per_warp + (unsigned)client * (unsigned)p.prod_bytes;
        const unsigned tbar = parea, pbar = parea + 8u + 8u * (unsigned)h;
        const unsigned table = parea + 64u;
        const unsigned pslot = table + (((unsigned)p.nb + 1u) * 4u + 127u) / 128u * 128u + (unsigned)h * FBQS_PF_BYTES;
        const unsigned ebar = blk + FBQS_EMPTY + 8u * (unsigned)h, fbar = blk + FBQS_FULL + 8u * (unsigned)h;
        const unsigned head = blk + FBQS_HEAD + 16u * (unsigned)h;
        const unsigned st = blk + FBQS_STAGES + (unsigned)h * FBQ_STAGE_BYTES;
        if (lane == 0) {
            fbq_mbar_init(tbar, 1u);
            fbq_mbar_init(parea + 8u, 1u);
            fbq_mbar_init(parea + 16u, 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            fbq_fence_async();
        }
        __syncwarp();
        unsigned epar = 1u;                              // the staging slot starts free: parity 1 passes on a fresh barrier
        unsigned tpar = 0u, ppar = 0u;                   // parities of the table barrier / this half's prefetch barrier
        unsigned gbase = 0u;                             // chunks this pass warp has been fed so far (slot = chunk & 1)
        // request the first 16 entries of both windows of the chunk at stream position tt into this half's prefetch slot
        auto prefetch = [&](int tt) {
            unsigned n0, n1, o0, o1;
            fbq_window_range(table, p.nb, tt, t_begin, n0, n1);
            fbq_window_range(table, p.nb, tt - D, t_begin, o0, o1);
            if (hl == 0) {
                const unsigned cn = (n1 - n0 < 16u ? n1 - n0 : 16u) * 32u, co = (o1 - o0 < 16u ? o1 - o0 : 16u) * 32u;
                fbq_mbar_expect_tx(pbar, cn + co);
                if (cn) fbq_bulk_g2s(pslot, p.nodes + n0, cn, pbar);
                if (co) fbq_bulk_g2s(pslot + 512u, p.nodes + o0, co, pbar);
            }
        };
#pragma unroll 1
        for (;;) {
            // claim the next 16-line group for this pass warp
            unsigned long long claimed = 0;
            if (lane == 0) claimed = atomicAdd(p.work_counter, 1ull);
            claimed = __shfl_sync(0xffffffffu, claimed, 0);
            if ((long long)claimed >= n_items) {
                if (h == (int)(gbase & 1u)) {            // the slot the pass warp looks at next
                    fbq_mbar_wait(ebar, epar);
                    if (hl == 0) fbq_sts4(head, 0, 0, 0, FBQS_FLAG_STOP);
                    fbq_mbar_arrive(fbar);
                }
                break;
            }
            const int outer = (int)((long long)claimed / p.n_groups);
            const int group = (int)((long long)claimed - (long long)outer * p.n_groups);
            // the line group's bucket table (nb + 1 entries; the copy is a multiple of 16 bytes: the array is padded)
            if (lane == 0) {
                const unsigned bytes = (((unsigned)p.nb + 1u) * 4u + 15u) & ~15u;
                fbq_mbar_expect_tx(tbar, bytes);
                fbq_bulk_g2s(table, p.bin_start + ((long long)outer * p.n_groups + group) * p.nb, bytes, tbar);
            }
            fbq_mbar_wait(tbar, tpar);
            tpar ^= 1u;
            int c = h ^ (int)(gbase & 1u);               // this half's first chunk of the line group
            if (c < nchunks) prefetch(t_begin + c * U);
#pragma unroll 1
            for (; c < nchunks; c += 2) {
                const int t = t_begin + c * U;
                // rows t .. t+7 (new) and t-D .. t-D+7 (old) into this half's staging slot
                fbq_mbar_wait_idle(ebar, epar);
                epar ^= 1u;
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    asm volatile("st.shared.v2.f64 [%0], {%1, %1};" ::"r"(st + (unsigned)(k * 16 + hl) * 16u), "d"(0.0) : "memory");
                unsigned n0, n1, o0, o1;
                fbq_window_range(table, p.nb, t, t_begin, n0, n1);
                fbq_window_range(table, p.nb, t - D, t_begin, o0, o1);
                fbq_mbar_wait(pbar, ppar);
                ppar ^= 1u;
                __syncwarp(hmask);                       // the zeros of all 16 lanes are in place
                fbq_window_scatter(p, n0, n1, pslot, t, st, hl);
                fbq_window_scatter(p, o0, o1, pslot + 512u, t - D, st + 2048u, hl);
                __syncwarp(hmask);                       // the prefetch slot has been read
                if (c + 2 < nchunks) prefetch(t + 2 * U);
                if (hl == 0) fbq_sts4(head, outer, group, t, 0);
                fbq_mbar_arrive(fbar);
            }
            __syncwarp();
            gbase += (unsigned)nchunks;
        }
        if constexpr (NT > 0) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;" ::: "memory");

    // ================================== pass warp ==================================
    const unsigned tile = sm0 + (unsigned)wid * FBQ_TILE_BYTES;
    const unsigned wsm = sm_blocks + (unsigned)wid * (unsigned)p.smem_per_warp;
    const unsigned stages = wsm + FBQS_STAGES;
    const unsigned ring_s = wsm + (unsigned)p.off_ring + (unsigned)lane * 8u;
    unsigned tring = 0;
    if constexpr (NT > 0)
        tring = s_tmem_base + ((unsigned)((wid & 3) * 32) << 16) + (unsigned)((wid >> 2) * p.tmem_cols_per_warp);
    const unsigned rp_bytes = (unsigned)RP * 256u, rp_cols = 2u * (unsigned)RP;
    const bool has_mirror = RP > R;
    const double alpha = p.alpha;
    const int fld = lane & 1;
    const unsigned tile_lane = tile + (unsigned)(lane >> 1) * 128u + (unsigned)fld * 8u;
    const unsigned tile_xor = (unsigned)((lane >> 1) & 7) << 4;

    double accu[NPASS], new0[NPASS], xsp[U];
#pragma unroll
    for (int q = 0; q < NPASS; ++q) { accu[q] = 0.0; new0[q] = 0.0; }
#pragma unroll
    for (int j = 0; j < U; ++j) xsp[j] = 0.0;
    unsigned fpar = 0;
    int slot = 0, wslot = 0, rslot = 0;
    bool pend = false;                                   // xsp holds rows pend_kb .. of (pend_outer, pend_group) not yet stored
    int pend_outer = 0, pend_group = 0, pend_kb = 0;

    auto emit = [&](const double (&x)[U], int outer, int group, int kb) {
        fbq_bulk_wait_read0();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < U; ++j) fbq_sts(tile_lane + (((unsigned)j << 4) ^ tile_xor), x[j]);
        fbq_fence_async();
        __syncwarp();
        fbq_tma_store(&tm_out, 2 * kb, group * 16, outer, tile);
    };

#pragma unroll 1
    for (;;) {
        fbq_mbar_wait(wsm + FBQS_FULL + 8u * (unsigned)slot, (fpar >> slot) & 1u);
        fpar ^= 1u << slot;
        int h_outer, h_group, t, h_flags;
        fbq_lds4(wsm + FBQS_HEAD + 16u * (unsigned)slot, h_outer, h_group, t, h_flags);
        // the header is the same for all lanes: through redux.sync into uniform registers (TMA operands, branches)
        const int outer = __reduce_max_sync(0xffffffffu, h_outer);
        const int group = __reduce_max_sync(0xffffffffu, h_group);
        t = __reduce_max_sync(0xffffffffu, t + 0x40000000) - 0x40000000;      // t may be negative
        if (__reduce_max_sync(0xffffffffu, h_flags) & FBQS_FLAG_STOP) break;
        if (t == t_begin) {
            // first chunk of a line group: rings and pass state start as zeros
            if constexpr (NT > 0) {
                unsigned z[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) z[i] = 0u;
                for (int i = 0; i < NT * RP; i += 8) fb_tmem_st16(tring + 2u * (unsigned)i, z);
                fb_tmem_wait_st();
            }
            if constexpr (NS > 0) {
                for (int i = 0; i < NS * RP; ++i) fbq_sts(ring_s + (unsigned)i * 256u, 0.0);
            }
#pragma unroll
            for (int q = 0; q < NPASS; ++q) { accu[q] = 0.0; new0[q] = 0.0; }
            wslot = 0;
            rslot = (R - D % R) % R;
            __syncwarp();
        }
        const unsigned stn = stages + (unsigned)slot * FBQ_STAGE_BYTES + (unsigned)lane * 8u;
        const unsigned sto = stn + 2048u;
        double bn[U], bo[U], xs[U];
#pragma unroll
        for (int j = 0; j < U; ++j) bn[j] = fbq_lds(stn + (unsigned)j * 256u);
#pragma unroll
        for (int j = 0; j < U; ++j) bo[j] = fbq_lds(sto + (unsigned)j * 256u);
        fbq_mbar_arrive(wsm + FBQS_EMPTY + 8u * (unsigned)slot);       // the slot is in registers: the producer may refill it
        const unsigned sr = ring_s + (unsigned)rslot * 256u, sw = ring_s + (unsigned)wslot * 256u;
        const unsigned tr = tring + 2u * (unsigned)rslot, tw = tring + 2u * (unsigned)wslot;
        const bool mirror = has_mirror && wslot == 0;
        const bool fast = (t - U >= lag) && (t + U - 1 - T1 < L);
        if (fast) {
            fbq_chunk<NPASS, NS, false>(bn, bo, accu, new0, xs, sr, sw, rp_bytes, tr, tw, rp_cols, mirror, (unsigned)R, t, T1, L, alpha);
            emit(xsp, pend_outer, pend_group, pend_kb);  // on the fast path the previous chunk always left rows
        } else {
            fbq_chunk<NPASS, NS, true>(bn, bo, accu, new0, xs, sr, sw, rp_bytes, tr, tw, rp_cols, mirror, (unsigned)R, t, T1, L, alpha);
            if (pend) emit(xsp, pend_outer, pend_group, pend_kb);
        }
        if constexpr (NT > 0) fb_tmem_wait_st();
        const int kb = t - lag;
        pend = kb >= 0 && kb < L;
        pend_outer = outer; pend_group = group; pend_kb = kb;
#pragma unroll
        for (int j = 0; j < U; ++j) xsp[j] = xs[j];
        wslot += U; wslot = (wslot >= R) ? 0 : wslot;
        rslot += U; rslot = (rslot >= R) ? rslot - R : rslot;
        ++slot; slot = (slot == nst) ? 0 : slot;
    }
    if (pend) emit(xsp, pend_outer, pend_group, pend_kb);
    fbq_bulk_wait0();
    if constexpr (NT > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (wid == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem_base), "r"((unsigned)p.tmem_alloc_cols) : "memory");
    }
}
