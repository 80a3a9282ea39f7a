// fb_sweep32.cuh -- fp32 working-precision sweeps (fb_problem.flags & FB_FLAG_FP32) for sm_100a.
//
// Same algorithm as the fp64 sweeps in fb_kernels.cuh (reference _convolve_tail_{2,3}d +
// _accumulate_tail_array, interpolation.py:373-533, with the mask / divide / cast of :427-430, :367
// fused into the last sweep), but the grid lives in HBM as interleaved float2 (value, weight) nodes
// and one thread owns BOTH fields of a line: every arithmetic step is one packed f32x2 instruction
// (FADD2 / FFMA2) for the two fields, HBM traffic per sweep halves (16 -> 8 B per node and
// direction) and a warp covers 32 lines.  The result is NOT bit-identical to the reference: it is
// the fp32 path of BASELINE.json's north_star ("within a stated tolerance and matching RMSE against
// the naive method"); tests/test_gpu_parity.py states the tolerance.
//
// Error control: a plain fp32 sliding accumulator random-walks over the whole line and keeps the
// absolute rounding error of a heavy region after leaving it (0.3 hPa worst case on the paper
// field).  Each pass therefore carries a Kahan compensation term next to its accumulator (3 more
// packed instructions per pass step); the kernels stay HBM-bound.
//
// Storage of the pass-to-pass rings (the element to subtract again D = 2T+2 steps later): pass 1
// re-reads global memory (L2), pass 2 keeps a ring in shared memory, passes 3..n keep theirs in
// tensor memory (tcgen05.ld/st, 2 columns per (value, weight) slot), 4 warps per CTA.
//
//   MODE 1: in2 [outer][k][inner] -> out2 [outer][inner][k]    (first sweep: reads the injected float2
//           nodes, A layout [..][x][y], and transposes through a padded smem tile)
//   MODE 0: in2 [outer][k][inner] -> out2 [outer][k][inner]    (3D middle sweep)
//   MODE 2: in2 [outer][k][inner] -> out32 [outer][k][inner]   float32 field (last sweep)
// The injection kernels write the float2 nodes directly (FbNodeWords<true> in fb_kernels.cuh: ordered
// fp64 sums rounded to float once).
#pragma once

typedef unsigned long long fb_f2;      // f32x2: low word = value field, high word = weight field

__device__ __forceinline__ fb_f2 fb2_pack(float lo, float hi)
{
    fb_f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void fb2_unpack(fb_f2 a, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
}
__device__ __forceinline__ fb_f2 fb2_add(fb_f2 a, fb_f2 b)
{
    fb_f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ fb_f2 fb2_sub(fb_f2 a, fb_f2 b)
{
    fb_f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ fb_f2 fb2_fma(fb_f2 a, fb_f2 b, fb_f2 c)
{
    fb_f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// the address arithmetic above hides the address space from the compiler: state it in the access
__device__ __forceinline__ fb_f2 fb_ldg2(const fb_f2 *q)
{
    fb_f2 r;
    asm("ld.global.nc.b64 %0, [%1];" : "=l"(r) : "l"(q));
    return r;
}
__device__ __forceinline__ void fb_stg2(fb_f2 *q, fb_f2 v) { asm volatile("st.global.b64 [%0], %1;" ::"l"(q), "l"(v) : "memory"); }
__device__ __forceinline__ void fb_stg1(float *q, float v) { asm volatile("st.global.f32 [%0], %1;" ::"l"(q), "f"(v) : "memory"); }

#ifndef FB32_L2_PREFETCH_CHUNKS
#define FB32_L2_PREFETCH_CHUNKS 5
#endif

struct FbSweep32 {
    const fb_f2 *in2;               // [outer][k][inner]
    fb_f2 *out2;                    // MODE 0 / 1
    float *out32;                   // MODE 2
    const unsigned long long *mm;   // MODE 2: min/max records (offset per field = outer)
    long long n_outer, L, n_inner, n_groups;   // n_groups: 32-line groups per outer index
    int T, D, R;
    float alpha, csf;
    unsigned long long *work_counter;
    int tmem_cols;                  // tensor-memory columns per CTA (0: none needed)
};

#ifndef FB32_TILE_K
#define FB32_TILE_K 32      // k extent of the transposing tile: a line receives runs of 256 bytes
#endif
#define FB32_TILE_PITCH 33

// volatile: keeps the load where it is written (the interior loop reloads a register right after its use)
__device__ __forceinline__ fb_f2 fb_ldg2_here(const fb_f2 *q)
{
    fb_f2 r;
    asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(r) : "l"(q));
    return r;
}

// tensor-memory ring reads of the chunk whose first read slot is rslot (issue only; fb_tmem_wait_ld16 later)
template <int NT, int U>
__device__ __forceinline__ void fb_sweep32_tmem_issue(unsigned (&oldr)[NT > 0 ? NT : 1][16], unsigned tring, int rslot, int R)
{
    if (NT > 0) {
        if (rslot + U <= R) {
#pragma unroll
            for (int q = 0; q < NT; ++q) fb_tmem_ld16(tring + 2u * (unsigned)(q * R + rslot), oldr[q]);
        } else {
#pragma unroll
            for (int j = 0; j < U; ++j) {
                int rj = rslot + j;
                rj = (rj >= R) ? rj - R : rj;
#pragma unroll
                for (int q = 0; q < NT; ++q) fb_tmem_ld2(tring + 2u * (unsigned)(q * R + rj), oldr[q][2 * j], oldr[q][2 * j + 1]);
            }
        }
    }
}

// The U = 8 steps of one chunk for all passes.  ring: this lane's shared-memory ring of pass 2
// (slot s at ring[s * 32]); tring: tensor-memory address of the ring of pass 3, slot 0.
// oldr holds the tensor-memory ring reads of THIS chunk, issued by the caller / the previous chunk.
// PIPE (interior chunks): every input register is reloaded for the next chunk (row pointers pn, po)
// right after its use, and the tensor-memory reads of the next chunk are issued at the end, so both
// latencies overlap a whole chunk of arithmetic without a second set of buffers.
template <int NPASS, int U, bool MASKED, bool PIPE>
__device__ __forceinline__ void fb_sweep32_chunk(
    fb_f2 (&bn)[U], fb_f2 (&bo)[U], fb_f2 (&accu)[NPASS], fb_f2 (&comp)[NPASS], fb_f2 (&new0)[NPASS],
    fb_f2 (&xs)[U], unsigned (&oldr)[(NPASS > 2 ? NPASS - 2 : 0) > 0 ? (NPASS > 2 ? NPASS - 2 : 0) : 1][16],
    fb_f2 *ring, unsigned tring, int rslot, int wslot, int R, int t, int T1, int L, fb_f2 alpha2,
    const fb_f2 *pn, const fb_f2 *po, unsigned sk8, int next_rslot)
{
    static_assert(U == 8, "one x16 tensor-memory access per ring and chunk");
    constexpr int NT = NPASS > 2 ? NPASS - 2 : 0;       // rings in tensor memory
    fb_f2 old0[U];
    if (NPASS > 1) {
        if (rslot + U <= R) {
            const fb_f2 *a = ring + rslot * 32;
#pragma unroll
            for (int j = 0; j < U; ++j) old0[j] = a[j * 32];
        } else {
#pragma unroll
            for (int j = 0; j < U; ++j) {
                int rj = rslot + j;
                rj = (rj >= R) ? rj - R : rj;
                old0[j] = ring[rj * 32];
            }
        }
        if (NT > 0) {
#pragma unroll
            for (int q = 0; q < NT; ++q) fb_tmem_wait_ld16(oldr[q]);
        }
    }
    unsigned newr[NT > 0 ? NT : 1][16];
    fb_f2 *wbase = ring + wslot * 32;
#pragma unroll
    for (int j = 0; j < U; ++j) {
        fb_f2 x = bn[j];
#pragma unroll
        for (int q = 0; q < NPASS; ++q) {
            fb_f2 o;
            if (q == 0) {
                o = bo[j];
            } else if (q == 1) {
                o = old0[j];
                wbase[j * 32] = x;
            } else {
                o = ((fb_f2)oldr[q - 2][2 * j + 1] << 32) | (fb_f2)oldr[q - 2][2 * j];
                newr[q - 2][2 * j] = (unsigned)x;
                newr[q - 2][2 * j + 1] = (unsigned)(x >> 32);
            }
            // accu += in[k+T] - in[k-T-1]  (interpolation.py:515-519), Kahan-compensated
            const fb_f2 y = fb2_sub(fb2_sub(new0[q], o), comp[q]);
            const fb_f2 s = fb2_add(accu[q], y);
            comp[q] = fb2_sub(fb2_sub(s, accu[q]), y);
            accu[q] = s;
            // out[k] = accu + alpha * (in[k-T-1] + in[k+T+1])   (:520)
            fb_f2 r = fb2_fma(alpha2, fb2_add(o, x), s);
            new0[q] = x;
            if (MASKED) {
                const int k = t + j - (q + 1) * T1;
                r = (k >= 0 && k < L) ? r : 0ull;
            }
            if (PIPE && q == 0) {
                bn[j] = fb_ldg2_here(fb_row(pn, (unsigned)j, sk8));
                bo[j] = fb_ldg2_here(fb_row(po, (unsigned)j, sk8));
            }
            x = r;
        }
        xs[j] = x;
    }
    if (NT > 0) {
#pragma unroll
        for (int q = 0; q < NT; ++q) fb_tmem_st16(tring + 2u * (unsigned)(q * R + wslot), newr[q]);
        fb_tmem_wait_st();
        if (PIPE) fb_sweep32_tmem_issue<NT, U>(oldr, tring, next_rslot, R);
    }
}

template <int NPASS, int MODE>
__global__ void __launch_bounds__(128, 2)
fb_sweep32_kernel(const FbSweep32 p)
{
    constexpr int U = 8;
    constexpr int NT = NPASS > 2 ? NPASS - 2 : 0;
    constexpr int TK = FB32_TILE_K, TP = FB32_TILE_PITCH;
    extern __shared__ __align__(16) unsigned long long fb_smem32[];
    __shared__ unsigned s_tmem_base;

    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    unsigned tmem_base = 0;
    if (NT > 0) {
        if (wid == 0) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_tmem_base);
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst), "r"((unsigned)p.tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_base = s_tmem_base;
    }
    const unsigned tring = tmem_base + ((unsigned)(wid * 32) << 16);   // this warp's lane quarter

    const int L = (int)p.L, T1 = p.T + 1, D = p.D, R = p.R;
    const long long sk = p.n_inner;
    const fb_f2 alpha2 = fb2_pack(p.alpha, p.alpha);
    // per warp: [ring of pass 2: R x 32][tile (MODE 1): TK x TP]
    const size_t warp_words = (size_t)(NPASS > 1 ? R * 32 : 0) + (MODE == 1 ? TK * TP : 0);
    fb_f2 *ring = fb_smem32 + (size_t)wid * warp_words + lane;
    fb_f2 *tile = fb_smem32 + (size_t)wid * warp_words + (NPASS > 1 ? R * 32 : 0);

    const long long n_items = p.n_outer * p.n_groups;
#pragma unroll 1
    for (;;) {
    unsigned long long claimed = 0;
    if (lane == 0) claimed = atomicAdd(p.work_counter, 1ull);
    claimed = __shfl_sync(0xffffffffu, claimed, 0);
    if ((long long)claimed >= n_items) break;
    const long long outer = (long long)claimed / p.n_groups;
    const long long group = (long long)claimed - outer * p.n_groups;
    const long long inner = group * 32 + lane;
    const bool active = inner < p.n_inner;
    const bool full_group = group * 32 + 32 <= p.n_inner;

    if (NPASS > 1) {
        for (int i = 0; i < R; ++i) ring[i * 32] = 0ull;
    }
    if (NT > 0) {
        unsigned z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
        for (int i = 0; i < NT * R; i += 8) fb_tmem_st16(tring + 2u * (unsigned)i, z);
        fb_tmem_wait_st();
    }

    const long long base = (outer * p.L) * p.n_inner + inner;      // element (outer, k = 0, inner)
    float offset = 0.0f;
    double offset64 = 0.0;
    if (MODE == 2) { offset64 = fb_field_offset(p.mm, outer); offset = (float)offset64; }
    (void)offset;

    fb_f2 accu[NPASS], comp[NPASS], new0[NPASS];
#pragma unroll
    for (int q = 0; q < NPASS; ++q) { accu[q] = 0ull; comp[q] = 0ull; new0[q] = 0ull; }
    const int lag = NPASS * T1;

    const fb_f2 *in = p.in2 + base;
    auto load_ranged = [&](fb_f2 (&r)[U], int t0) {
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int tt = t0 + j;
            r[j] = (active && tt >= 0 && tt < L) ? in[(long long)tt * sk] : 0ull;
        }
    };
    // byte offsets inside a line fit 32 bits (checked by the launcher): one IMAD.WIDE per address
    const unsigned sk8 = (unsigned)p.n_inner * 8u;
    // Straight-line loads (no predicate, no branch: a conditional load makes the compiler merge the loaded
    // registers with MOVs right behind the loads, which stalls the warp for the whole load latency).
    // Lanes beyond n_inner read the last valid line instead; nothing of theirs is ever stored.
    const fb_f2 *in_safe = p.in2 + (outer * p.L) * p.n_inner + (active ? inner : p.n_inner - 1);
    auto load_inside = [&](fb_f2 (&r)[U], int t0) {
        const fb_f2 *p0 = fb_row(in_safe, (unsigned)t0, sk8);
#pragma unroll
        for (int j = 0; j < U; ++j) r[j] = fb_ldg2(fb_row(p0, (unsigned)j, sk8));
    };
    // L2 prefetch of the rows [t0, t0 + nrows), nrows <= 16, with ONE instruction per warp: lane l takes the
    // 128-byte half (l & 1) of the warp's 256-byte segment of row t0 + (l >> 1)
    const fb_f2 *in_seg = p.in2 + (outer * p.L) * p.n_inner + group * 32 + (lane & 1) * 16;
    const bool pf_lane = group * 32 + (lane & 1) * 16 < p.n_inner;
    auto prefetch_l2 = [&](int t0, int nrows) {
        const int row = t0 + (lane >> 1);
        if (pf_lane && (lane >> 1) < nrows && row >= 0 && row < L)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(fb_row(in_seg, (unsigned)row, sk8)));
    };

    // MODE 1: rows k0 .. k0+cnt-1 of the 32 line columns; per store instruction each half warp writes 16
    // consecutive k of one line (128 B)
    auto flush_tile = [&](int k0, int cnt) {
        __syncwarp();
        constexpr int CPI = 32 / TK;                   // line columns per store instruction
        const int kk = lane % TK, sub = lane / TK;
        if (full_group && cnt == TK) {
            const fb_f2 *tp = tile + kk * TP + sub;
            fb_f2 *o = p.out2 + (outer * p.n_inner + group * 32 + sub) * p.L + k0 + kk;
            const unsigned rs8 = (unsigned)p.L * (8u * CPI);          // CPI lines further, in bytes
#pragma unroll
            for (int it = 0; it < TK; ++it) fb_stg2(fb_row(o, (unsigned)it, rs8), tp[it * CPI]);
        } else {
#pragma unroll 4
            for (int it = 0; it < TK; ++it) {
                const int col = it * CPI + sub;
                const long long inner_j = group * 32 + col;
                if (kk < cnt && inner_j < p.n_inner)
                    p.out2[(outer * p.n_inner + inner_j) * p.L + k0 + kk] = tile[kk * TP + col];
            }
        }
        __syncwarp();
    };
    // `if wg < csf: wg = nan` (interpolation.py:430); (vg / wg + offset).astype(float32) (:367): the quotient is
    // taken in fp32, the offset added in fp64 so that the large constant does not cost output bits
    auto finalize = [&](fb_f2 x) -> float {
        float v, w;
        fb2_unpack(x, v, w);
        const float q = __fdividef(v, w);      // 2 ulp of the centred quotient: far below one ulp of the field value
        const float r = __double2float_rn(__dadd_rn((double)q, offset64));
        return (w < p.csf) ? __int_as_float(0x7fc00000) : r;
    };
    auto emit = [&](int k, fb_f2 x) {
        if (MODE == 0) {
            if (active) p.out2[base + (long long)k * sk] = x;
        } else if (MODE == 1) {
            tile[(k & (TK - 1)) * TP + lane] = x;         // flushed by the caller
        } else {
            if (active) p.out32[base + (long long)k * sk] = finalize(x);
        }
    };
    auto emit_chunk = [&](const fb_f2 (&xs)[U], int kb) {
        if (MODE == 0) {
            if (active) {
                fb_f2 *o = fb_row(p.out2 + base, (unsigned)kb, sk8);
#pragma unroll
                for (int j = 0; j < U; ++j) fb_stg2(fb_row(o, (unsigned)j, sk8), xs[j]);
            }
        } else if (MODE == 1) {
            const int row0 = kb & (TK - 1);
            fb_f2 *tp = tile + row0 * TP + lane;
#pragma unroll
            for (int j = 0; j < U; ++j) tp[j * TP] = xs[j];
            if (row0 + U == TK) flush_tile(kb - row0, TK);
            else if (kb + U == L) flush_tile(kb - row0, row0 + U);
        } else {
            if (active) {
                float *o = fb_row(p.out32 + base, (unsigned)kb, sk8 >> 1);
#pragma unroll
                for (int j = 0; j < U; ++j) fb_stg1(fb_row(o, (unsigned)j, sk8 >> 1), finalize(xs[j]));
            }
        }
    };

    // stream positions as in fb_sweep_kernel: masked line start, unmasked interior, masked line end
    const int steady_lo = lag > D ? lag : D;
    const int t_begin = -((U - lag % U) % U);
    const int t_end = L + lag;
    int t_lo = steady_lo + (U - (steady_lo - t_begin) % U) % U;
    int t_hi = t_lo + ((L - t_lo) > 0 ? (L - t_lo) / U * U : 0);
    if (t_hi < t_lo) t_hi = t_lo;
    if (t_lo > t_end) { t_lo = t_hi = t_begin + (t_end - t_begin + U - 1) / U * U; }

    int wslot = 0;
    int rslot = (R - D % R) % R;
    auto advance = [&]() {
        wslot += U;
        wslot = (wslot == R) ? 0 : wslot;
        rslot += U;
        rslot = (rslot >= R) ? rslot - R : rslot;
    };

    fb_f2 xs[U];
    unsigned oldr[NT > 0 ? NT : 1][16];
    int t = t_begin;
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
        const int stop = phase == 0 ? t_lo : t_end;
        if (t < stop) {
            fb_f2 bn[U], bo[U];
#pragma unroll 1
            for (; t < stop; t += U) {
                prefetch_l2(t + FB32_L2_PREFETCH_CHUNKS * U, U);
                load_ranged(bn, t);
                load_ranged(bo, t - D);
                fb_sweep32_tmem_issue<NT, U>(oldr, tring, rslot, R);
                fb_sweep32_chunk<NPASS, U, true, false>(bn, bo, accu, comp, new0, xs, oldr, ring, tring, rslot, wslot, R, t, T1, L, alpha2,
                                                       nullptr, nullptr, 0u, 0);
                const int kb = t - lag;
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const int k = kb + j;
                    if (k >= 0 && k < L) emit(k, xs[j]);
                }
                if (MODE == 1) {
                    const int kend = (kb + U < L) ? kb + U : L;
                    if (kend > 0 && kend > kb && ((kend & (TK - 1)) == 0 || kend == L)) {
                        const int k0 = (kend - 1) & ~(TK - 1);
                        flush_tile(k0, kend - k0);
                    }
                }
                advance();
            }
        }
        // interior: software pipeline inside fb_sweep32_chunk (inputs and tensor-memory reads of chunk t+U are
        // requested while chunk t is computed)
        if (phase == 0 && t < t_hi) {
            fb_f2 an[U], ao[U];
            load_inside(an, t);
            load_inside(ao, t - D);
            fb_sweep32_tmem_issue<NT, U>(oldr, tring, rslot, R);
#pragma unroll 1
            for (; t < t_hi; t += U) {
                prefetch_l2(t + FB32_L2_PREFETCH_CHUNKS * U, U);
                // the chunk after the last interior one is loaded by the masked code: reload this one instead
                const int tn = (t + U < t_hi) ? t + U : t;
                const fb_f2 *pn = fb_row(in_safe, (unsigned)tn, sk8);
                const fb_f2 *po = fb_row(in_safe, (unsigned)(tn - D), sk8);
                int nr = rslot + U;
                nr = (nr >= R) ? nr - R : nr;
                fb_sweep32_chunk<NPASS, U, false, true>(an, ao, accu, comp, new0, xs, oldr, ring, tring, rslot, wslot, R, t, T1, L,
                                                       alpha2, pn, po, sk8, nr);
                emit_chunk(xs, t - lag);
                advance();
            }
            if (NT > 0) {
#pragma unroll
                for (int q = 0; q < NT; ++q) fb_tmem_wait_ld16(oldr[q]);     // drain the reads issued by the last chunk
            }
        }
    }
    }   // persistent loop
    if (NT > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (wid == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((unsigned)p.tmem_cols) : "memory");
    }
}
