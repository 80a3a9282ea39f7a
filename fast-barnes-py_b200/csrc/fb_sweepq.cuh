// fb_sweepq.cuh -- fp64 axis sweep, second generation ("q" kernels): one warp owns 16 lines x 2 fields
// and runs ALL passes of the n-fold filter; rows travel with the bulk asynchronous copy engine (TMA).
//
// Replaces, like fb_sweep*_kernel, the line loops of _convolve_tail_{2,3}d (fastbarnes/interpolation.py:
// 398-479) with _accumulate_tail_array (:485-533) inlined, MODE 2 also the NaN mask (:427-430 / :475-479)
// and `(vg / wg + offset).astype(np.float32)` (:367).  The arithmetic (operation order, explicit *_rn
// intrinsics, zero extension instead of the five loop phases) is the one of fb_sweep_chunk in
// fb_kernels.cuh, so results are bit-identical to the reference and to the first-generation kernels.
//
// What is different from fb_sweeph_kernel (two warps per 16 lines, lock step, register-staged loads):
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

#define FBQ_U 8                 // steps per chunk
#define FBQ_STAGE_BYTES 4096    // one staging slot: 8 new rows + 8 old rows of 256 bytes
#define FBQ_TILE_BYTES 2048     // output tile of a warp (MODE 0: 8 rows x 256 B; MODE 1: 16 lines x 128 B, 128B-swizzled)
#define FBQ_MAX_STAGES 8

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
    int nst;                     // staging slots per warp (chunks in flight), 2..FBQ_MAX_STAGES
    int pf;                      // extra chunks of lead of an L2 prefetch of the new rows (0: none)
    int tmem_cols_per_warp;      // tensor-memory columns of one warp (two warps share a lane quarter when 8 warps run)
    int tmem_alloc_cols;         // columns the CTA allocates (power of two >= 32; 512 when more than 4 warps run)
    int smem_per_warp;           // bytes of a warp's block of dynamic shared memory: [mbarriers 128 B][stages][rings]
    int off_ring;                // byte offset of the rings inside that block (stages start at 128)
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
// TMA tensor load of one box (32 doubles x 8 rows x 1) at (c0, c1, c2); completion counted in bytes on the mbarrier
__device__ __forceinline__ void fbq_tma_load(unsigned dst, const CUtensorMap *tm, int c0, int c1, int c2, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void fbq_tma_prefetch(const CUtensorMap *tm, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// TMA tensor store of one box from shared memory (bulk async-group completion); out-of-bounds parts are clipped
__device__ __forceinline__ void fbq_tma_store(const CUtensorMap *tm, int c0, int c1, int c2, unsigned src)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(c0), "r"(c1),
                 "r"(c2), "r"(src) : "memory");
}
__device__ __forceinline__ void fbq_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void fbq_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fbq_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fbq_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
            double r = __dadd_rn(accu[q], __dmul_rn(alpha, __dadd_rn(o, x)));
            new0[q] = x;
            if (MASKED) {
                const int k = t + j - (q + 1) * T1;
                r = (k >= 0 && k < L) ? r : 0.0;
            }
            x = r;
        }
        xs[j] = x;
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
#pragma unroll
        for (int q = 0; q < NT; ++q) fb_tmem_st16(tw + (unsigned)q * rp_cols, newt[q]);
        if (mirror) {
#pragma unroll
            for (int q = 0; q < NT; ++q) fb_tmem_st16(tw + (unsigned)q * rp_cols + 2u * mirror_slots, newt[q]);
        }
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
    static_assert(NS >= 0 && NT >= 0, "ring split");
    extern __shared__ __align__(1024) unsigned char fbq_smem[];
    __shared__ unsigned s_tmem_base;

    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
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
    const unsigned stages = wsm + 128u;
    const unsigned ring_s = wsm + (unsigned)p.off_ring + (unsigned)lane * 8u;
    const int nst = p.nst;
    if (lane == 0) {
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
        unsigned long long claimed = 0;
        if (lane == 0) claimed = atomicAdd(p.work_counter, 1ull);
        claimed = __shfl_sync(0xffffffffu, claimed, 0);
        if ((long long)claimed >= n_items) break;
        const int outer = (int)((long long)claimed / p.n_groups);
        const int group = (int)((long long)claimed - (long long)outer * p.n_groups);
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

        // request the rows of the chunk at stream position tt into staging slot `slot`: the 8 newest rows
        // tt .. tt+7 and the 8 rows tt-D .. tt-D+7 (rows outside [0, L) arrive as zeros)
        auto issue = [&](int tt, int slot) {
            if (lane == 0) {
                const unsigned st = stages + (unsigned)slot * FBQ_STAGE_BYTES;
                const unsigned bar = bars + 8u * (unsigned)slot;
                fbq_mbar_expect_tx(bar, FBQ_STAGE_BYTES);
                fbq_tma_load(st, &tm_in, group * 32, tt, outer, bar);
                fbq_tma_load(st + 2048u, &tm_in, group * 32, tt - D, outer, bar);
                if (p.pf > 0 && tt + p.pf * U < L) fbq_tma_prefetch(&tm_in, group * 32, tt + p.pf * U, outer);
            }
        };

        {
            const int npro = nst < nchunks ? nst : nchunks;
            for (int c = 0; c < npro; ++c) issue(t_begin + c * U, c);
        }

        double accu[NPASS], new0[NPASS];
#pragma unroll
        for (int q = 0; q < NPASS; ++q) { accu[q] = 0.0; new0[q] = 0.0; }
        double offset = 0.0;
        if (MODE == 2) offset = fb_field_offset(p.mm, outer);

        int wslot = 0;                                   // ring write slot of this chunk (multiple of 8)
        int rslot = (R - D % R) % R;                     // ring read slot: (wslot - D) mod R
        int slot = 0;                                    // staging slot of this chunk
        int t = t_begin;
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c, t += U) {
            const unsigned st = stages + (unsigned)slot * FBQ_STAGE_BYTES + (unsigned)lane * 8u;
            fbq_mbar_wait(bars + 8u * (unsigned)slot, (phases >> slot) & 1u);
            phases ^= 1u << slot;
            double bn[U], bo[U], xs[U];
#pragma unroll
            for (int j = 0; j < U; ++j) bn[j] = fbq_lds(st + (unsigned)j * 256u);
#pragma unroll
            for (int j = 0; j < U; ++j) bo[j] = fbq_lds(st + 2048u + (unsigned)j * 256u);
            const unsigned sr = ring_s + (unsigned)rslot * 256u, sw = ring_s + (unsigned)wslot * 256u;
            const unsigned tr = tring + 2u * (unsigned)rslot, tw = tring + 2u * (unsigned)wslot;
            const bool mirror = has_mirror && wslot == 0;
            const bool interior = (t >= lag) && (t + U - 1 - T1 < L);
            const int kb = t - lag;                      // rows kb .. kb+7 leave the last pass (kb is a multiple of 8)
            if (interior) {
                fbq_chunk<NPASS, NS, false>(bn, bo, accu, new0, xs, sr, sw, rp_bytes, tr, tw, rp_cols, mirror, (unsigned)R, t, T1, L, alpha);
            } else {
                fbq_chunk<NPASS, NS, true>(bn, bo, accu, new0, xs, sr, sw, rp_bytes, tr, tw, rp_cols, mirror, (unsigned)R, t, T1, L, alpha);
            }
            // ---- output of rows kb .. kb+7 (those inside the line)
            if (kb >= 0 && kb < L) {
                if (MODE == 0 || MODE == 1) {
                    if (lane == 0) fbq_bulk_wait_read0();            // the previous tile has been read
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < U; ++j)
                        fbq_sts(MODE == 1 ? tile_lane + (((unsigned)j << 4) ^ tile_xor) : tile_lane + (unsigned)j * 256u, xs[j]);
                    fbq_fence_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (MODE == 1) fbq_tma_store(&tm_out, 2 * kb, group * 16, outer, tile);
                        else fbq_tma_store(&tm_out, group * 32, kb, outer, tile);
                        fbq_bulk_commit();
                    }
                } else {
                    unsigned rows = 0;
                    if (inner < p.n_inner) {
                        if (interior) {
                            rows = 15u;
                        } else {
#pragma unroll
                            for (int m = 0; m < U / 2; ++m)
                                if (kb + 2 * m + fld < L) rows |= 1u << m;
                        }
                    }
                    const long long o = ((long long)outer * p.L + kb + fld) * p.n_inner + inner;
                    fbq_finalize_chunk(xs, fld, p.csf, offset, p.out32 + o, p.out64 ? p.out64 + o : nullptr, (unsigned)p.n_inner, rows);
                }
            }
            // ---- the staging slot is free: request the chunk nst chunks ahead
            __syncwarp();
            if (c + nst < nchunks) issue(t + nst * U, slot);
            if constexpr (NT > 0) fb_tmem_wait_st();
            wslot += U; wslot = (wslot >= R) ? 0 : wslot;
            rslot += U; rslot = (rslot >= R) ? rslot - R : rslot;
            ++slot; slot = (slot == nst) ? 0 : slot;
        }
    }   // persistent loop

    if (MODE == 0 || MODE == 1) {
        if (lane == 0) fbq_bulk_wait0();
    }
    if constexpr (NT > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (wid == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem_base), "r"((unsigned)p.tmem_alloc_cols) : "memory");
    }
}
