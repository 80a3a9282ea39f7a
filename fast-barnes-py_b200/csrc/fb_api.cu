// fb_api.cu -- C ABI (include/fastbarnes_b200.h) over the sm_100a kernels in fb_kernels.cuh.
//
// Host-side responsibilities: kernel parameters (T, alpha, conv_scale_factor) in the exact
// arithmetic of the reference, workspace carving, the launch plan of the sweeps (how many of
// the n passes are fused per launch given the on-chip ring storage), and the HOST-buffer
// convenience entry points (device arena + H2D/D2H).  There is no CPU compute path: every
// entry point that computes fails with FB_ECUDA when no device is present.
#include "../../include/fastbarnes_b200.h"
#include "fb_kernels.cuh"
#include "fb_exact.cuh"
#include "fb_sweep32.cuh"
#include "fb_sweepq.cuh"
#include "fb_line1d.cuh"
#include "fb_sweepp.cuh"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#define FB_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
std::atomic<int> g_profiling{0};
// interleaved form: records linked into per-node lists (two passes over the samples) instead of count / allocate / place
std::atomic<int> g_inject_lists{1};
// second-generation sweep kernels (fb_sweepq.cuh): 0 off, non-zero (default) on for 2D / 3D fp64 grids they cover
std::atomic<int> g_sweepq{1};
std::atomic<int> g_q_reserve{0};    // SMs the q kernels leave free (z-slab runs: room for the exchange kernels)
std::atomic<int> g_q_deep{1};       // transposing / in-place q sweeps: a fourth staging slot, with 7 warps where 8 do not fit
std::atomic<int> g_q_nst{3};        // staging slots (chunks of rows in flight) per warp
std::atomic<int> g_q_pf{0};         // extra chunks of lead of the L2 prefetch (0: none)
std::atomic<int> g_sweepp{1};       // small batches: pass-parallel sweeps (fb_sweepp.cuh): 0 off, 1 when the batch is small, 2 always
std::atomic<int> g_line1d{1};       // 1D grids: the two-warp line kernel (fb_line1d.cuh) for the exact walk
std::atomic<int> g_q_warps{8};      // warps per CTA the plan starts with (8 or 4)
std::atomic<int> g_q_cap_xy{8};     // most warps per CTA of the transposing / in-place sweeps (shared memory and registers an
std::atomic<int> g_q_cap_fin{8};    // ... of the finalising sweep            SM keeps free for kernels of other streams)
// q path: injection as sort by cell + segmented reduce feeding the x sweep (fb_sparse.cuh) instead of dense grids
std::atomic<int> g_sparse{0};       // opt-in: measured slower than the dense path (x sweep 1.55 vs 1.15 ms on the bench batch), see DESIGN.md

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(FB_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define LAUNCH_CHECK()                                                                         \
    do {                                                                                       \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                   \
        cudaError_t e_ = cudaGetLastError();                                                   \
        if (e_ != cudaSuccess)                                                                 \
            return fail(FB_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// ---- profiling (CUDA events on the launching stream) -----------------------------------------
// segment i lies between event i and event i+1:
//   0 zero-fill + init   1 min/max + injection   2..4 axis sweeps (x, y, z as far as present)
constexpr int kProfSegments = 5;
struct Profile {
    cudaEvent_t ev[kProfSegments + 1] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int marked = 0;          // events recorded in the last call
    bool armed = false;
    long long launches_begin = 0, launches_end = 0;
};
thread_local Profile g_prof;

int prof_mark(int i, cudaStream_t st)
{
    if (!g_profiling.load()) return FB_OK;
    if (!g_prof.ev[i]) CUDA_TRY(cudaEventCreate(&g_prof.ev[i]));
    CUDA_TRY(cudaEventRecord(g_prof.ev[i], st));
    if (i + 1 > g_prof.marked) g_prof.marked = i + 1;
    return FB_OK;
}

// ---- kernel parameters -------------------------------------------------------------------------
// `float64 ** int` as Numba lowers it (numba/cpython/numbers.py, int_power_impl):
// square-and-multiply, not libm pow.
double int_power(double a, long long b)
{
    double r = 1.0;
    bool invert = false;
    long long e = b;
    if (b < 0) { invert = true; e = -b; }
    if (e > 0x10000) return std::pow(a, (double)b);
    while (e != 0) {
        if (e & 1) r *= a;
        e >>= 1;
        a *= a;
    }
    return invert ? 1.0 / r : r;
}

struct AxisParams { int T; double alpha; };

// ---- sweeps ------------------------------------------------------------------------------------
constexpr size_t kSmemLimit = 227 * 1024 - 1024;   // opt-in dynamic smem per CTA, minus slack
constexpr size_t kSmemPerSM = 228 * 1024;          // shared memory of one SM (each CTA also reserves 1 KB)

int sm_count(int dev)
{
    static int cached[16] = {0};
    if (cached[dev & 15] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
        cached[dev & 15] = n;
    }
    return cached[dev & 15];
}

int sm_count_current()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return sm_count(dev);
}

// Work counters of the persistent sweep launches: one slot per launch of a call, zeroed on the
// stream right before the launch.
struct SweepCounters {
    unsigned long long *base;   // device array of kSweepCounterSlots entries
    int next;
};
constexpr int kSweepCounterSlots = 12;

// chunk length U and ring depth R of a sweep with delay D = 2T+2 (see fb_sweep_kernel)
inline int sweep_chunk(int D) { return D >= FB_SWEEP_U ? FB_SWEEP_U : FB_SWEEP_U_SMALL; }
inline int sweep_ring_depth(int D)
{
    const int U = sweep_chunk(D);
    return (D + U + U - 1) / U * U;
}

size_t sweep_smem_bytes(int npass, int mode, int D)
{
    const int nr = npass - 1;
    size_t b = (size_t)nr * sweep_ring_depth(D) * 32 * sizeof(double);
    if (mode == 1) b += (size_t)FB_TILE_K * FB_TILE_PITCH * sizeof(double);
    return b;
}

// ---- second-generation sweeps: fb_sweepq_kernel ------------------------------------------------------------------
// Launch plan of one axis: warps per CTA (one CTA per SM), rings in shared memory (ns; the other
// npass - 1 - ns rings sit in tensor memory), staging depth, per-warp shared-memory layout.
struct SweepQPlan {
    bool ok;
    int warps, ns, nst, R, RP, cols_per_warp, smem_per_warp, off_ring;
    int launch_warps;            // warps a CTA is launched with (<= warps: the tensor-memory layout is that of `warps`)
};
constexpr size_t kQSmemLimit = 227 * 1024 - 1024;        // opt-in limit per CTA minus the kernel's static shared memory (1 KB with the alignment)

// dynamic shared memory of a q-sweep CTA: output tiles + the warps' blocks
inline size_t sweepq_smem_bytes(int mode, int warps, int smem_per_warp)
{
    return (mode == 2 ? 0 : (size_t)warps * FBQ_TILE_BYTES) + (size_t)warps * smem_per_warp;
}

SweepQPlan sweepq_plan(int npass, int mode, int D, int header = 128, int extra_per_warp = 0, int nst_force = 0)
{
    SweepQPlan q{};
    q.ok = false;
    if (npass < 1 || npass > FB_MAX_FUSED_PASSES || D < FBQ_U) return q;
    q.R = (D + FBQ_U - 1) / FBQ_U * FBQ_U;
    q.RP = q.R + (D % FBQ_U ? FBQ_U : 0);
    const int nr = npass - 1;
    int nst_want = g_q_nst.load();
    if (nst_want < 2) nst_want = 2;
    if (nst_want > FBQ_MAX_STAGES) nst_want = FBQ_MAX_STAGES;
    if (nst_force > 0) nst_want = nst_force;
    for (int warps = ((g_q_warps.load() >= 8 || nst_force > 0) ? 8 : 4); warps >= (nst_force > 0 ? 8 : 4); warps -= 4) {
        const int cols = warps == 8 ? 256 : 512;
        int nt = cols / (2 * q.RP);
        if (nt > nr) nt = nr;
        const int ns = nr - nt;
        if (ns > 1) continue;
        // Candidates (staging slots, warps per CTA): the finalising sweep and the sparse-fed sweep take the configured depth
        // with all warps.  The transposing / in-place sweeps are memory bound and gain more from a fourth staging slot than
        // they lose with one warp less (T = 27: 8 warps x 3 slots 1.138 ms, 7 x 3: 1.095, 7 x 4: 1.078, 6 x 5: 1.105), so
        // they try 8 x 4, 7 x 4, then the configured depth downwards.
        const int cap = nst_force > 0 ? warps : (mode == 2 ? g_q_cap_fin.load() : g_q_cap_xy.load());
        const int wmax = (cap >= 1 && cap < warps) ? cap : warps;
        int cand[16][2], nc = 0;
        if (mode != 2 && nst_force == 0 && warps == 8 && nst_want < 4 && g_q_deep.load() != 0) {
            cand[nc][0] = 4; cand[nc++][1] = wmax;
            if (wmax == 8) { cand[nc][0] = 4; cand[nc++][1] = 7; }
        }
        for (int nst = nst_want; nst >= 2 && nc < 16; --nst) { cand[nc][0] = nst; cand[nc++][1] = wmax; }
        for (int i = 0; i < nc; ++i) {
            const int nst = cand[i][0], wfit = cand[i][1];
            const size_t off_ring = (size_t)header + (size_t)nst * FBQ_STAGE_BYTES;
            const size_t per = off_ring + (size_t)ns * q.RP * 256;
            if (sweepq_smem_bytes(mode, wfit, (int)per) + (size_t)wfit * extra_per_warp > kQSmemLimit) continue;
            q.ok = true;
            q.warps = warps; q.launch_warps = wfit; q.ns = ns; q.nst = nst; q.cols_per_warp = cols;
            q.smem_per_warp = (int)per; q.off_ring = (int)off_ring;
            return q;
        }
    }
    return q;
}

// ---- tensor maps (TMA descriptors) of the q sweeps; the driver's encoder is looked up at run time, so the
// library keeps loading (and exporting its symbols) on machines without a CUDA driver
typedef CUresult (*FbTensorMapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                           const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
FbTensorMapEncodeTiled tensor_map_encoder()
{
    static std::atomic<void *> cached{nullptr};
    void *fn = cached.load();
    if (!fn) {
        cudaDriverEntryPointQueryResult qres = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            fn = nullptr;
        cudaGetLastError();
        cached.store(fn);
    }
    return (FbTensorMapEncodeTiled)fn;
}

// fp64 tensor [d2][d1][d0] (d0 contiguous) with row pitches s1, s2 in bytes, box (b0, b1, 1)
int make_tensor_map(CUtensorMap &m, const void *base, unsigned long long d0, unsigned long long d1, unsigned long long d2,
                    unsigned long long s1, unsigned long long s2, unsigned b0, unsigned b1, bool swizzle128, bool promote)
{
    FbTensorMapEncodeTiled enc = tensor_map_encoder();
    if (!enc) return fail(FB_ECUDA, "cuTensorMapEncodeTiled is not available from this CUDA driver");
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {s1, s2};
    const cuuint32_t box[3] = {b0, b1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                           promote ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(FB_ECUDA, "cuTensorMapEncodeTiled failed (%d): dims %llu x %llu x %llu, pitches %llu / %llu", (int)r, d0, d1, d2, s1, s2);
    return FB_OK;
}

template <int NPASS, int NS, int MODE>
int launch_sweepq_t(FbSweepQ p, const SweepQPlan &q, cudaStream_t st)
{
    static thread_local bool configured[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 15]) {
        CUDA_TRY(cudaFuncSetAttribute(fb_sweepq_kernel<NPASS, NS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kQSmemLimit));
        CUDA_TRY(cudaFuncSetAttribute(fb_sweepq_kernel<NPASS, NS, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
        configured[dev & 15] = true;
    }
    const long long nitems = p.n_outer * p.n_groups;
    if (nitems <= 0) return FB_OK;
    // input: in[outer][k][2 * n_inner], box = 8 rows of one 16-line group
    CUtensorMap tm_in, tm_out;
    const unsigned long long row = (unsigned long long)p.n_inner * 16ull;
    int rc = make_tensor_map(tm_in, p.in, 2ull * p.n_inner, (unsigned long long)p.L, (unsigned long long)p.n_outer, row, row * p.L, 32, FBQ_U,
                             false, true);
    if (rc != FB_OK) return rc;
    if (MODE == 0) {
        rc = make_tensor_map(tm_out, p.out, 2ull * p.n_inner, (unsigned long long)p.L, (unsigned long long)p.n_outer, row, row * p.L, 32,
                             FBQ_U, false, false);
    } else if (MODE == 1) {
        // transposed: out[outer][inner][2 * L], box = 8 steps (128 bytes, swizzled in shared memory) of 16 lines
        const unsigned long long orow = (unsigned long long)p.L * 16ull;
        rc = make_tensor_map(tm_out, p.out, 2ull * p.L, (unsigned long long)p.n_inner, (unsigned long long)p.n_outer, orow,
                             orow * p.n_inner, 16, 16, true, false);
    } else {
        tm_out = tm_in;
    }
    if (rc != FB_OK) return rc;
    const int sms = sm_count(dev);
    // few items: spread them over the SMs with fewer warps per CTA
    int warps = q.launch_warps > 0 ? q.launch_warps : q.warps;
    if (nitems < (long long)sms * warps) {
        warps = (int)((nitems + sms - 1) / sms);
        if (warps < 1) warps = 1;
    }
    long long grid = (nitems + warps - 1) / warps;
    // one CTA per SM, minus the SMs left free for kernels of other streams (the NCCL kernels of a halo exchange cannot start
    // next to a CTA that holds all of an SM's shared memory)
    int avail = sms - g_q_reserve.load();
    if (avail < 1) avail = 1;
    if (grid > avail) grid = avail;
    const size_t smem = sweepq_smem_bytes(MODE, warps, q.smem_per_warp);
    // tensor memory: the whole SM's 512 columns when two warps share a lane quarter, else what one warp's rings need
    // (small launches may place several CTAs on one SM)
    p.tmem_alloc_cols = 32;
    if (warps > 4) p.tmem_alloc_cols = 512;
    else
        while (p.tmem_alloc_cols < (NPASS - 1 - NS) * 2 * q.RP) p.tmem_alloc_cols *= 2;
    if (getenv("FB_DEBUG"))
        fprintf(stderr, "[fb] sweepq<%d,%d,%d> warps %d grid %lld smem %zu nst %d R %d RP %d D %d items %lld\n", NPASS, NS, MODE, warps,
                grid, smem, q.nst, q.R, q.RP, p.D, nitems);
    CUDA_TRY(cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned long long), st));
    fb_sweepq_kernel<NPASS, NS, MODE><<<(unsigned)grid, warps * 32, smem, st>>>(p, tm_in, tm_out);
    LAUNCH_CHECK();
    return FB_OK;
}

template <int MODE>
int launch_sweepq_m(int npass, const FbSweepQ &p, const SweepQPlan &q, cudaStream_t st)
{
#ifdef FBQ_LAB
    // lab builds (make EXTRA=-DFBQ_LAB): only the q kernels of the bench configuration, for quick A/B variants
    if (npass == 4 && q.ns == 1) return launch_sweepq_t<4, 1, MODE>(p, q, st);
    return fail(FB_EKERNEL, "lab build: kernel family compiled out");
#else
    switch (npass * 10 + q.ns) {
    case 10: return launch_sweepq_t<1, 0, MODE>(p, q, st);
    case 20: return launch_sweepq_t<2, 0, MODE>(p, q, st);
    case 21: return launch_sweepq_t<2, 1, MODE>(p, q, st);
    case 30: return launch_sweepq_t<3, 0, MODE>(p, q, st);
    case 31: return launch_sweepq_t<3, 1, MODE>(p, q, st);
    case 40: return launch_sweepq_t<4, 0, MODE>(p, q, st);
    case 41: return launch_sweepq_t<4, 1, MODE>(p, q, st);
    case 50: return launch_sweepq_t<5, 0, MODE>(p, q, st);
    case 51: return launch_sweepq_t<5, 1, MODE>(p, q, st);
    case 60: return launch_sweepq_t<6, 0, MODE>(p, q, st);
    case 61: return launch_sweepq_t<6, 1, MODE>(p, q, st);
    }
    return fail(FB_EINVAL, "unsupported q-sweep configuration: %d passes, %d shared-memory rings", npass, q.ns);
#endif
}

// ---- the x sweep fed by the binned samples (fb_sweepqs_kernel) -----------------------------------------------------
// geometry of the binning: it must agree with the stream of the x sweep (chunks of 8 rows starting at t_begin)
FbBins bins_geometry(int num_iter, int T, long long W, long long H, long long n_outer)
{
    FbBins bn{};
    const int lag = num_iter * (T + 1);
    bn.t_begin = -((FBQ_U - lag % FBQ_U) % FBQ_U);
    bn.G = (int)((H + 15) / 16);
    bn.NB = ((int)((W - 1 - bn.t_begin) / FBQ_U) + 1 + 3) / 4 * 4;       // multiple of 4: the rows of bin_start stay 16-byte aligned

    bn.nbuckets = n_outer * bn.G * bn.NB;
    return bn;
}

// a pass warp's share of the producer area: 64 B of mbarriers, the bucket table of a line group, the prefetch slots
inline int sweepqs_prod_bytes(int nb) { return 64 + (int)(((size_t)(nb + 1) * 4 + 127) / 128 * 128) + FBQS_PF * (int)FBQS_PF_BYTES; }

template <int NPASS, int NS>
int launch_sweepqs_t(FbSweepQ p, const SweepQPlan &q, cudaStream_t st)
{
    static thread_local bool configured[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 15]) {
        CUDA_TRY(cudaFuncSetAttribute(fb_sweepqs_kernel<NPASS, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kQSmemLimit));
        CUDA_TRY(cudaFuncSetAttribute(fb_sweepqs_kernel<NPASS, NS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
        configured[dev & 15] = true;
    }
    const long long nitems = p.n_outer * p.n_groups;
    if (nitems <= 0) return FB_OK;
    CUtensorMap tm_out;
    const unsigned long long orow = (unsigned long long)p.L * 16ull;
    int rc = make_tensor_map(tm_out, p.out, 2ull * p.L, (unsigned long long)p.n_inner, (unsigned long long)p.n_outer, orow,
                             orow * p.n_inner, 16, 16, true, false);
    if (rc != FB_OK) return rc;
    const int sms = sm_count(dev);
    const int warps = 8;                                 // two warpgroups of pass warps + two of producer warps
    if (q.warps != 8 || q.nst != 2) return fail(FB_EKERNEL, "internal: plan of the sparse x sweep");
    long long grid = (nitems + warps - 1) / warps;
    if (grid > sms) grid = sms;
    // producer area per pass warp: 64 B of mbarriers, the bucket table of a line group, the prefetch slots
    p.prod_bytes = sweepqs_prod_bytes(p.nb);
    const size_t smem = sweepq_smem_bytes(1, warps, q.smem_per_warp) + (size_t)warps * p.prod_bytes;
    if (smem > kQSmemLimit) return fail(FB_EKERNEL, "internal: shared memory of the sparse x sweep (%zu bytes)", smem);
    p.tmem_alloc_cols = 32;
    if (warps > 4) p.tmem_alloc_cols = 512;
    else
        while (p.tmem_alloc_cols < (NPASS - 1 - NS) * 2 * q.RP) p.tmem_alloc_cols *= 2;
    p.ncw = warps;
    const int threads = 2 * warps * 32;
    if (getenv("FB_DEBUG"))
        fprintf(stderr, "[fb] sweepqs<%d,%d> warps %d+%d grid %lld smem %zu nst %d R %d RP %d D %d items %lld buckets/group %d\n", NPASS,
                NS, warps, warps, grid, smem, q.nst, q.R, q.RP, p.D, nitems, p.nb);
    CUDA_TRY(cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned long long), st));
    fb_sweepqs_kernel<NPASS, NS><<<(unsigned)grid, threads, smem, st>>>(p, tm_out);
    LAUNCH_CHECK();
    return FB_OK;
}

int launch_sweepqs(int npass, const FbSweepQ &p, const SweepQPlan &q, cudaStream_t st)
{
#ifdef FBQ_LAB
    if (npass == 4 && q.ns == 1) return launch_sweepqs_t<4, 1>(p, q, st);
    return fail(FB_EKERNEL, "lab build: kernel family compiled out");
#else
    switch (npass * 10 + q.ns) {
    case 10: return launch_sweepqs_t<1, 0>(p, q, st);
    case 20: return launch_sweepqs_t<2, 0>(p, q, st);
    case 21: return launch_sweepqs_t<2, 1>(p, q, st);
    case 30: return launch_sweepqs_t<3, 0>(p, q, st);
    case 31: return launch_sweepqs_t<3, 1>(p, q, st);
    case 40: return launch_sweepqs_t<4, 0>(p, q, st);
    case 41: return launch_sweepqs_t<4, 1>(p, q, st);
    case 50: return launch_sweepqs_t<5, 0>(p, q, st);
    case 51: return launch_sweepqs_t<5, 1>(p, q, st);
    case 60: return launch_sweepqs_t<6, 0>(p, q, st);
    case 61: return launch_sweepqs_t<6, 1>(p, q, st);
    }
    return fail(FB_EINVAL, "unsupported q-sweep configuration: %d passes, %d shared-memory rings", npass, q.ns);
#endif
}

// one axis of the q path: src / dst are grids of interleaved (value, weight) nodes
int run_sweepq(int mode, int num_iter, const AxisParams &ax, const double *src, double *dst, float *out32, double *out64,
               const unsigned long long *mm, double csf, long long n_outer, long long L, long long n_inner, cudaStream_t st,
               SweepCounters &ctr)
{
    const int D = 2 * ax.T + 2;
    const SweepQPlan q = sweepq_plan(num_iter, mode, D);
    if (!q.ok) return fail(FB_EKERNEL, "internal: the q sweep does not cover T=%d, num_iter=%d", ax.T, num_iter);
    if (L > 2147483647LL - 8 * (long long)(ax.T + 1) - 64) return fail(FB_EINVAL, "line too long: %lld", L);
    FbSweepQ p{};
    p.in = src; p.out = dst; p.out32 = out32; p.out64 = out64; p.mm = mm;
    p.n_outer = n_outer; p.L = L; p.n_inner = n_inner; p.n_groups = (n_inner + 15) / 16;
    p.T = ax.T; p.D = D; p.R = q.R; p.RP = q.RP;
    p.alpha = ax.alpha; p.csf = csf;
    p.work_counter = ctr.base + (ctr.next++ % kSweepCounterSlots);
    p.nst = q.nst; p.pf = g_q_pf.load();
    p.tmem_cols_per_warp = q.cols_per_warp;
    p.smem_per_warp = q.smem_per_warp; p.off_ring = q.off_ring;
    if (mode == 0) return launch_sweepq_m<0>(num_iter, p, q, st);
    if (mode == 1) return launch_sweepq_m<1>(num_iter, p, q, st);
    return launch_sweepq_m<2>(num_iter, p, q, st);
}

// ---- fp32 working precision (FB_FLAG_FP32): fb_sweep32_kernel ------------------------------------------------
inline int sweep32_ring_depth(int D) { return (D + FB_SWEEP_U + FB_SWEEP_U - 1) / FB_SWEEP_U * FB_SWEEP_U; }
inline int sweep32_tmem_cols(int npass, int D)
{
    const int nt = npass > 2 ? npass - 2 : 0;
    if (nt == 0) return 0;
    int need = nt * sweep32_ring_depth(D) * 2, cols = 32;
    while (cols < need) cols *= 2;
    return cols;
}
inline size_t sweep32_smem_bytes(int npass, int mode, int D)
{
    size_t words = npass > 1 ? (size_t)sweep32_ring_depth(D) * 32 : 0;
    if (mode == 1) words += (size_t)FB32_TILE_K * FB32_TILE_PITCH;
    return 4 * words * sizeof(unsigned long long);
}
// all passes of an axis run in one launch: batched ring reads need D >= U, the rings must fit on chip
inline bool sweep32_fits(int npass, int D)
{
    return npass >= 1 && npass <= FB_MAX_FUSED_PASSES && D >= FB_SWEEP_U && sweep32_tmem_cols(npass, D) <= 512 &&
           sweep32_smem_bytes(npass, 1, D) + 1024 <= kSmemLimit;
}

template <int NPASS, int MODE>
int launch_sweep32_t(FbSweep32 p, cudaStream_t st)
{
    static thread_local size_t configured[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = sweep32_smem_bytes(NPASS, MODE, p.D);
    if (smem > 40 * 1024 && configured[dev & 15] < smem) {
        CUDA_TRY(cudaFuncSetAttribute(fb_sweep32_kernel<NPASS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
        CUDA_TRY(cudaFuncSetAttribute(fb_sweep32_kernel<NPASS, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
        configured[dev & 15] = kSmemLimit;
    }
    const long long nitems = p.n_outer * p.n_groups;
    if (nitems <= 0) return FB_OK;
    p.tmem_cols = sweep32_tmem_cols(NPASS, p.D);
    int per_sm = 2;                                          // launch bounds: 2 CTAs of 128 threads
    if ((int)(kSmemPerSM / (smem + 1024)) < per_sm) per_sm = (int)(kSmemPerSM / (smem + 1024));
    if (p.tmem_cols > 0 && per_sm * p.tmem_cols > 512) per_sm = 512 / p.tmem_cols;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)per_sm * sm_count(dev);
    if (grid > (nitems + 3) / 4) grid = (nitems + 3) / 4;
    if (getenv("FB_DEBUG")) fprintf(stderr, "[fb] sweep32<%d,%d> smem %zu tmem_cols %d per_sm %d grid %lld items %lld\n", NPASS, MODE, smem, p.tmem_cols, per_sm, grid, nitems);
    CUDA_TRY(cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned long long), st));
    fb_sweep32_kernel<NPASS, MODE><<<(unsigned)grid, 128, smem, st>>>(p);
    LAUNCH_CHECK();
    return FB_OK;
}

template <int MODE>
int launch_sweep32_m(int npass, const FbSweep32 &p, cudaStream_t st)
{
#ifdef FBQ_LAB
    (void)npass;
    return fail(FB_EKERNEL, "lab build: kernel family compiled out");
#else
    switch (npass) {
    case 1: return launch_sweep32_t<1, MODE>(p, st);
    case 2: return launch_sweep32_t<2, MODE>(p, st);
    case 3: return launch_sweep32_t<3, MODE>(p, st);
    case 4: return launch_sweep32_t<4, MODE>(p, st);
    case 5: return launch_sweep32_t<5, MODE>(p, st);
    case 6: return launch_sweep32_t<6, MODE>(p, st);
    }
    return fail(FB_EINVAL, "unsupported number of fused passes: %d", npass);
#endif
}

// one axis of the fp32 path
int run_sweep32(int mode, int num_iter, const AxisParams &ax, const fb_f2 *src2, fb_f2 *dst2, float *out32, const unsigned long long *mm, double csf, long long n_outer, long long L,
                long long n_inner, cudaStream_t st, SweepCounters &ctr)
{
    const int D = 2 * ax.T + 2;
    if (!sweep32_fits(num_iter, D))
        return fail(FB_EKERNEL, "the fp32 path does not cover this kernel (T=%d, num_iter=%d: needs 3 <= T and on-chip rings); "
                                "use the fp64 path", ax.T, num_iter);
    if ((L + (FB32_L2_PREFETCH_CHUNKS + 2) * FB_SWEEP_U) * n_inner * 8 > 4294967295LL)
        return fail(FB_EINVAL, "the fp32 path addresses a line with 32-bit element offsets: L * n_inner too large");
    FbSweep32 p{};
    p.in2 = src2; p.out2 = dst2; p.out32 = out32; p.mm = mm;
    p.n_outer = n_outer; p.L = L; p.n_inner = n_inner; p.n_groups = (n_inner + 31) / 32;
    p.T = ax.T; p.D = D; p.R = sweep32_ring_depth(D);
    p.alpha = (float)ax.alpha; p.csf = (float)csf;
    p.work_counter = ctr.base + (ctr.next++ % kSweepCounterSlots);
    if (mode == 0) return launch_sweep32_m<0>(num_iter, p, st);
    if (mode == 1) return launch_sweep32_m<1>(num_iter, p, st);
    return launch_sweep32_m<2>(num_iter, p, st);
}

template <int NPASS, int MODE, int U>
int launch_sweep_t(const FbSweep &p, size_t smem, cudaStream_t st)
{
    static thread_local size_t configured[16] = {0};   // the attribute is sticky per function and device
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && configured[dev & 15] < smem) {
        CUDA_TRY(cudaFuncSetAttribute(fb_sweep_kernel<NPASS, MODE, U>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kSmemLimit));
        CUDA_TRY(cudaFuncSetAttribute(fb_sweep_kernel<NPASS, MODE, U>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
        configured[dev & 15] = kSmemLimit;
    }
    const long long nitems = p.n_outer * p.n_groups;
    if (nitems <= 0) return FB_OK;
    static thread_local int occ[16] = {0};
    static thread_local size_t occ_smem[16] = {0};
    if (occ[dev & 15] == 0 || occ_smem[dev & 15] != smem) {
        int nb = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fb_sweep_kernel<NPASS, MODE, U>, 32, smem));
        occ[dev & 15] = nb > 0 ? nb : 1;
        occ_smem[dev & 15] = smem;
    }
    long long grid = (long long)occ[dev & 15] * sm_count(dev);
    if (grid > nitems) grid = nitems;
    CUDA_TRY(cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned long long), st));
    fb_sweep_kernel<NPASS, MODE, U><<<(unsigned)grid, 32, smem, st>>>(p);
    LAUNCH_CHECK();
    return FB_OK;
}

template <int MODE, int U>
int launch_sweep_u(int npass, const FbSweep &p, size_t smem, cudaStream_t st)
{
#ifdef FBQ_LAB
    (void)npass;
    return fail(FB_EKERNEL, "lab build: kernel family compiled out");
#else
    switch (npass) {
    case 1: return launch_sweep_t<1, MODE, U>(p, smem, st);
    case 2: return launch_sweep_t<2, MODE, U>(p, smem, st);
    case 3: return launch_sweep_t<3, MODE, U>(p, smem, st);
    case 4: return launch_sweep_t<4, MODE, U>(p, smem, st);
    case 5: return launch_sweep_t<5, MODE, U>(p, smem, st);
    case 6: return launch_sweep_t<6, MODE, U>(p, smem, st);
    }
    return fail(FB_EINVAL, "unsupported number of fused passes: %d", npass);
#endif
}

template <int MODE>
int launch_sweep_m(int npass, const FbSweep &p, size_t smem, cudaStream_t st)
{
    if (sweep_chunk(p.D) == FB_SWEEP_U) return launch_sweep_u<MODE, FB_SWEEP_U>(npass, p, smem, st);
    return launch_sweep_u<MODE, FB_SWEEP_U_SMALL>(npass, p, smem, st);
}

// A pair of fp64 grids (value field, weight field) in device memory.
struct Pair { double *v, *w; };

// One axis sweep of `num_iter` passes over the grids in `cur`.
//   mode 0: same layout.    mode 1: transposed output.    mode 2: finalised float32 output.
// The passes are fused into as few launches as the per-warp ring storage allows (a launch fusing
// f passes keeps f-1 rings of 2T+2 elements per line in shared memory; f = 1 needs none, so very
// wide kernels degrade to one launch per pass).  Launches with f >= 2 that keep the layout run in
// place; single-pass launches and the transposing launch write to `spare`, after which the roles
// of the two pairs swap.  On return `cur` holds the result (modes 0 and 1).
// largest number of passes one launch of the first-generation kernel can fuse
inline int sweep_fmax(int mode, int D)
{
    int fmax = 1;
    for (int f = 2; f <= FB_MAX_FUSED_PASSES; ++f)
        if (sweep_smem_bytes(f, mode, D) <= kSmemLimit) fmax = f;
    return fmax;
}
// first-generation sweep of one axis on planes of values and of weights (fb_sweep_kernel): 1D lines, kernels of fewer
// than 8 elements, kernels whose rings do not fit on chip (the launch is then split into groups of passes), and the
// cross-check of the q path in the tests
int run_sweep(int mode, int num_iter, AxisParams ax, Pair &cur, Pair &spare, float *out32, double *out64,
              const unsigned long long *mm, double csf, long long n_outer, long long L, long long n_inner,
              bool has_w, cudaStream_t st, SweepCounters &ctr)
{
    FbSweep p{};
    p.n_outer = n_outer;
    p.L = L;
    p.n_inner = n_inner;
    p.n_groups = (n_inner + 15) / 16;
    p.T = ax.T;
    p.D = 2 * ax.T + 2;
    p.R = sweep_ring_depth(p.D);
    p.has_w = has_w ? 1 : 0;
    p.alpha = ax.alpha;
    p.csf = csf;
    p.mm = mm;
    p.out32 = out32;
    p.out64 = out64;
    if (L > 2147483647LL - 8 * (long long)(ax.T + 1) - 64) return fail(FB_EINVAL, "line too long: %lld", L);

    const int fmax = sweep_fmax(mode, p.D);
    const int nlaunch = (num_iter + fmax - 1) / fmax;
    int remaining = num_iter;
    for (int l = 0; l < nlaunch; ++l) {
        const int np = (remaining + (nlaunch - l) - 1) / (nlaunch - l);
        remaining -= np;
        const bool last = (l == nlaunch - 1);
        const int m = last ? mode : 0;
        const bool in_place = (m == 0 && np >= 2);
        p.work_counter = ctr.base + (ctr.next++ % kSweepCounterSlots);
        p.in_v = cur.v;
        p.in_w = cur.w;
        if (m == 2) {
            p.out_v = p.out_w = nullptr;
        } else if (in_place) {
            p.out_v = cur.v;
            p.out_w = cur.w;
        } else {
            if (!spare.v || (has_w && !spare.w)) return fail(FB_EINVAL, "internal: sweep needs a spare buffer pair");
            p.out_v = spare.v;
            p.out_w = spare.w;
        }
        int rc = FB_OK;
        {
            const size_t smem = sweep_smem_bytes(np, m, p.D);
            if (smem > kSmemLimit) return fail(FB_EKERNEL, "ring storage does not fit: T=%d passes=%d", ax.T, np);
            if (m == 0) rc = launch_sweep_m<0>(np, p, smem, st);
            else if (m == 1) rc = launch_sweep_m<1>(np, p, smem, st);
            else rc = launch_sweep_m<2>(np, p, smem, st);
        }
        if (rc != FB_OK) return rc;
        if (m != 2 && !in_place) { Pair t = cur; cur = spare; spare = t; }
    }
    return FB_OK;
}

// ---- problem validation / derived parameters -----------------------------------------------------
struct Derived {
    AxisParams ax[3];
    int32_t ks[3];
    double csf;
    long long W, H, Dz, total;
};

int derive(const fb_problem *pr, Derived &d)
{
    if (!pr) return fail(FB_EINVAL, "null problem");
    if (pr->dim < 1 || pr->dim > 3) return fail(FB_EINVAL, "dim must be 1, 2 or 3: %d", pr->dim);
    if (pr->num_iter < 1) return fail(FB_EINVAL, "num_iter must be >= 1: %d", pr->num_iter);
    if (pr->nfields < 1 || pr->nfields > 65535) return fail(FB_EINVAL, "nfields must be in [1, 65535]: %lld", (long long)pr->nfields);
    if (pr->method != FB_METHOD_OPTIMIZED_CONVOLUTION && pr->method != FB_METHOD_CONVOLUTION)
        return fail(FB_EINVAL, "unknown method id: %d", pr->method);
    d.W = pr->size[0];
    d.H = pr->dim > 1 ? pr->size[1] : 1;
    d.Dz = pr->dim > 2 ? pr->size[2] : 1;
    if (d.W < 1 || d.H < 1 || d.Dz < 1) return fail(FB_EINVAL, "grid size must be positive");
    d.total = d.W * d.H * d.Dz;
    double tv[3] = {0, 0, 0};
    for (int m = 0; m < pr->dim; ++m) {
        int T;
        if (pr->method == FB_METHOD_CONVOLUTION) {
            T = fb_half_kernel_size(pr->sigma[m], pr->step[m], pr->num_iter);
            tv[m] = 0.0;
        } else {
            T = fb_half_kernel_size_opt(pr->sigma[m], pr->step[m], pr->num_iter);
            tv[m] = fb_tail_value(pr->sigma[m], pr->step[m], pr->num_iter);
        }
        if (T < 0) return fail(FB_EINVAL, "negative half kernel size on axis %d", m);
        d.ax[m].T = T;
        d.ax[m].alpha = tv[m];
        d.ks[m] = 2 * T + 1;
    }
    d.csf = fb_conv_scale_factor(pr->dim, d.ks, tv, pr->sigma, pr->step, pr->num_iter, pr->max_dist_weight);
    return FB_OK;
}

// ---- 1D segmentation plan --------------------------------------------------------------------------
struct SegPlan { bool on; long long seg_len, halo, n_seg, Le; };

SegPlan seg_plan(const fb_problem *pr, const Derived &d)
{
    SegPlan s{false, 0, 0, 0, 0};
    if (pr->dim != 1 || pr->nfields != 1 || !(pr->flags & FB_FLAG_SEGMENTED_1D)) return s;
    s.halo = (long long)pr->num_iter * (d.ax[0].T + 1);
    long long seg = (d.W + 18943) / 18944;                 // about two rounds of 16-segment work items on 148 SMs
    if (seg < 4 * s.halo) seg = 4 * s.halo;                 // halo overhead <= 50 %
    if (seg < 256) seg = 256;
    seg = (seg + 63) / 64 * 64;
    s.seg_len = seg;
    s.n_seg = (d.W + seg - 1) / seg;
    s.Le = seg + 2 * s.halo;
    s.on = s.n_seg >= 2 && d.W >= s.Le;
    return s;
}

// ---- workspace -----------------------------------------------------------------------------------
struct Workspace {
    double *vA, *wA, *vB, *wB;
    unsigned long long *mm, *counters;
    long long *offsets;
    unsigned char *first_mask;
    int *rec_k;
    double *rec_w, *rec_wv;
    long long *seg_node;
    unsigned int *seg_base, *seg_n;
    double *seg[4];          // 1D segmented path: extended segments (v, w) x (in, out)
    unsigned int *link_next; // two-pass injection: successor of every record in its node's list
    FbRec *bin_rec;          // binned injection: the records (aliases rec_k .. seg_base, which the dense injection uses)
    unsigned int *bin_cnt, *bin_start, *bin_sums;   // binned injection (fb_sparse.cuh): records per bucket, first record, scan scratch
    long long bin_cap;       // buckets the bin arrays hold
    size_t bytes;
};

// carve (base may be nullptr to only compute the size)
void carve(Workspace &w, char *base, const fb_problem *pr, long long total, long long nsamples, const SegPlan *sp = nullptr)
{
    size_t off = 0;
    auto take = [&](size_t n) { char *p = base ? base + off : nullptr; off += align_up(n); return p; };
    const size_t g = (size_t)pr->nfields * (size_t)total * sizeof(double);
    const size_t R = (size_t)nsamples << pr->dim;
    w.vA = (double *)take(2 * g);                           // one block: also holds the interleaved (value, weight) form
    w.wA = base ? (double *)((char *)w.vA + g) : nullptr;
    w.vB = (double *)take(g);
    w.wB = (double *)take(g);
    w.mm = (unsigned long long *)take((size_t)pr->nfields * FB_MM_STRIDE * 8);
    w.counters = (unsigned long long *)take((4 + kSweepCounterSlots) * 8);
    w.offsets = (long long *)take((size_t)(pr->nfields + 1) * 8);
    w.first_mask = (unsigned char *)take((size_t)nsamples + 1);
    w.rec_k = (int *)take(R * 4 + 4);                       // rec_k .. seg_base: 4 + 8 + 8 + 8 + 4 = 32 bytes per record, contiguous
    w.bin_rec = (FbRec *)w.rec_k;
    w.rec_w = (double *)take(R * 8 + 8);
    w.rec_wv = (double *)take(R * 8 + 8);
    w.seg_node = (long long *)take(R * 8 + 8);
    w.seg_base = (unsigned int *)take(R * 4 + 4);
    w.seg_n = (unsigned int *)take(R * 4 + 4);
    for (int i = 0; i < 4; ++i)
        w.seg[i] = (sp && sp->on) ? (double *)take((size_t)sp->Le * sp->n_seg * sizeof(double)) : nullptr;
    w.link_next = (unsigned int *)take(R * 4 + 4);
    // buckets of the binned injection: (field x z plane) x 16-line groups along y x chunks of 8 rows along x (+ 2 for the
    // stream offset), dim >= 2 only
    w.bin_cap = pr->dim >= 2 ? (long long)pr->nfields * (pr->dim > 2 ? pr->size[2] : 1) * ((pr->size[1] + 15) / 16) * (pr->size[0] / 8 + 6) : 0;
    w.bin_cnt = (unsigned int *)take((size_t)(w.bin_cap + 1) * 4);
    w.bin_start = (unsigned int *)take((size_t)(w.bin_cap + 8) * 4);       // + slack: the producers copy whole 16-byte units
    w.bin_sums = (unsigned int *)take((size_t)(w.bin_cap / (FB_SCAN_BLOCK * FB_SCAN_PER_THREAD) + 4) * 4);
    w.bytes = off;
}

// ---- the pipeline: centre -> inject -> sweeps (+ finalize) -----------------------------------------
int check_offsets(const fb_problem *pr, long long nsamples, const int64_t *off, long long &max_n)
{
    if (nsamples < 0) return fail(FB_EINVAL, "negative sample count");
    if (!off) {
        if (nsamples % pr->nfields) return fail(FB_EINVAL, "nsamples (%lld) is not a multiple of nfields (%lld) and no sample_offsets given", nsamples, (long long)pr->nfields);
        max_n = nsamples / pr->nfields;
        return FB_OK;
    }
    max_n = 0;
    if (off[0] != 0 || off[pr->nfields] != nsamples) return fail(FB_EINVAL, "sample_offsets must start at 0 and end at nsamples");
    for (long long b = 0; b < pr->nfields; ++b) {
        long long n = off[b + 1] - off[b];
        if (n < 0) return fail(FB_EINVAL, "sample_offsets must be non-decreasing");
        if (n > max_n) max_n = n;
    }
    return FB_OK;
}

// z_off / z_cnt: z window of the injected buffers (z-slab runs); the whole grid is (0, d.Dz).
int run_inject(const fb_problem *pr, const Derived &d, long long nsamples, const int64_t *h_offsets,
               const double *d_pts, const double *d_val, Workspace &w, cudaStream_t st,
               long long z_off = 0, long long z_cnt = -1, bool f32 = false, bool il64 = false,
               const FbSamples *inject_from = nullptr)
{
    // inject_from: the samples the injection kernels read (z-slab runs: the compacted list; offsets on the device, at most
    // nsamples of them); min / max always run over d_pts / d_val
    // il64: the injected grid is ONE array of interleaved double2 (value, weight) nodes in w.vA (FORM 2)
    // f32: the injected grid is ONE array of interleaved float2 (value, weight) nodes in w.vA (fp32 path)
    if (z_cnt < 0) z_cnt = d.Dz;
    const long long slab_total = d.W * d.H * z_cnt;
    long long max_n = 0;
    int rc = check_offsets(pr, nsamples, h_offsets, max_n);
    if (rc != FB_OK) return rc;
    if (max_n > 2147483647LL) return fail(FB_EINVAL, "too many samples in one field");
    if ((nsamples << pr->dim) > (f32 ? 2147483647LL : 4294967295LL)) return fail(FB_EINVAL, "too many sample records");
    const size_t g = (size_t)pr->nfields * (size_t)slab_total * sizeof(double);
    CUDA_TRY(cudaMemsetAsync(w.vA, 0, g, st));
    if (!f32) CUDA_TRY(cudaMemsetAsync(w.wA, 0, g, st));
    fb_init_kernel<<<(unsigned)((pr->nfields + 255) / 256), 256, 0, st>>>(w.mm, pr->nfields, w.counters);
    LAUNCH_CHECK();
    {
        int prc = prof_mark(1, st);
        if (prc != FB_OK) return prc;
    }
    FbSamples s{};
    s.pts = d_pts;
    s.val = d_val;
    if (h_offsets) {
        CUDA_TRY(cudaMemcpyAsync(w.offsets, h_offsets, (size_t)(pr->nfields + 1) * 8, cudaMemcpyHostToDevice, st));
        s.offsets = w.offsets;
        s.n_uniform = 0;
    } else {
        s.offsets = nullptr;
        s.n_uniform = max_n;
    }
    if (max_n == 0) return FB_OK;
    FbGrid gr{};
    gr.dim = pr->dim;
    gr.W = d.W; gr.H = d.H; gr.Dz = d.Dz; gr.total = slab_total;
    gr.z_off = z_off; gr.z_cnt = z_cnt;
    for (int m = 0; m < 3; ++m) { gr.x0[m] = pr->x0[m]; gr.step[m] = pr->step[m]; }

    const unsigned nf = (unsigned)pr->nfields;
    long long mmb = (max_n + 256 * 8 - 1) / (256 * 8);
    if (mmb > 1024) mmb = 1024;
    fb_minmax_kernel<<<dim3((unsigned)mmb, nf), 256, 0, st>>>(s, w.mm);
    LAUNCH_CHECK();
    if (inject_from) s = *inject_from;
    const dim3 sg((unsigned)((max_n + 255) / 256), nf);
    const long long R = nsamples << pr->dim;
    long long rblocks = (R + 127) / 128;
    if (rblocks > 148 * 16) rblocks = 148 * 16;
    if (f32 && g_inject_lists.load() != 0) {
        fb_inject_link_kernel<1><<<sg, 256, 0, st>>>(s, gr, w.vA, w.wA, w.link_next);
        LAUNCH_CHECK();
        fb_inject_finish_kernel<1><<<sg, 256, 0, st>>>(s, gr, w.vA, w.wA, w.link_next, w.mm, w.counters, w.seg_node, w.seg_base,
                                                      w.seg_n, w.rec_k, w.rec_w, w.rec_wv);
        LAUNCH_CHECK();
        fb_inject_reduce_kernel<true><<<(unsigned)rblocks, 128, 0, st>>>(w.counters, w.seg_node, w.seg_base, w.seg_n,
                                                                               w.rec_k, w.rec_w, w.rec_wv, w.vA, w.wA);
    } else if (f32) {
        fb_inject_count_kernel<true><<<sg, 256, 0, st>>>(s, gr, w.vA, w.wA, w.first_mask);
        LAUNCH_CHECK();
        fb_inject_alloc_kernel<true><<<sg, 256, 0, st>>>(s, gr, w.vA, w.wA, w.first_mask, w.counters, w.seg_node, w.seg_base, w.seg_n);
        LAUNCH_CHECK();
        fb_inject_place_kernel<true><<<sg, 256, 0, st>>>(s, gr, w.vA, w.wA, w.mm, w.rec_k, w.rec_w, w.rec_wv);
        LAUNCH_CHECK();
        fb_inject_reduce_kernel<true><<<(unsigned)rblocks, 128, 0, st>>>(w.counters, w.seg_node, w.seg_base, w.seg_n,
                                                                               w.rec_k, w.rec_w, w.rec_wv, w.vA, w.wA);
    } else if (il64 && g_inject_lists.load() != 0) {
        // two passes over the samples: link the records of every node into a list, then finish the lists
        fb_inject_link_kernel<2><<<sg, 256, 0, st>>>(s, gr, w.vA, w.wA, w.link_next);
        LAUNCH_CHECK();
        fb_inject_finish_kernel<2><<<sg, 256, 0, st>>>(s, gr, w.vA, w.wA, w.link_next, w.mm, w.counters, w.seg_node, w.seg_base,
                                                      w.seg_n, w.rec_k, w.rec_w, w.rec_wv);
        LAUNCH_CHECK();
        fb_inject_reduce_kernel<2><<<(unsigned)rblocks, 128, 0, st>>>(w.counters, w.seg_node, w.seg_base, w.seg_n,
                                                                            w.rec_k, w.rec_w, w.rec_wv, w.vA, w.wA);
    } else if (il64) {
        fb_inject_count_kernel<2><<<sg, 256, 0, st>>>(s, gr, w.vA, w.wA, w.first_mask);
        LAUNCH_CHECK();
        fb_inject_alloc_kernel<2><<<sg, 256, 0, st>>>(s, gr, w.vA, w.wA, w.first_mask, w.counters, w.seg_node, w.seg_base, w.seg_n);
        LAUNCH_CHECK();
        fb_inject_place_kernel<2><<<sg, 256, 0, st>>>(s, gr, w.vA, w.wA, w.mm, w.rec_k, w.rec_w, w.rec_wv);
        LAUNCH_CHECK();
        fb_inject_reduce_kernel<2><<<(unsigned)rblocks, 128, 0, st>>>(w.counters, w.seg_node, w.seg_base, w.seg_n,
                                                                            w.rec_k, w.rec_w, w.rec_wv, w.vA, w.wA);
    } else {
        fb_inject_count_kernel<false><<<sg, 256, 0, st>>>(s, gr, w.vA, w.wA, w.first_mask);
        LAUNCH_CHECK();
        fb_inject_alloc_kernel<false><<<sg, 256, 0, st>>>(s, gr, w.vA, w.wA, w.first_mask, w.counters, w.seg_node, w.seg_base, w.seg_n);
        LAUNCH_CHECK();
        fb_inject_place_kernel<false><<<sg, 256, 0, st>>>(s, gr, w.vA, w.wA, w.mm, w.rec_k, w.rec_w, w.rec_wv);
        LAUNCH_CHECK();
        fb_inject_reduce_kernel<false><<<(unsigned)rblocks, 128, 0, st>>>(w.counters, w.seg_node, w.seg_base, w.seg_n,
                                                                                w.rec_k, w.rec_w, w.rec_wv, w.vA, w.wA);
    }
    LAUNCH_CHECK();
    return FB_OK;
}

// The fp64 injection writes interleaved (value, weight) nodes when the sweeps that consume them run on nodes (use_nodes).
// Per axis the grids of interleaved nodes are swept by the q kernels (fb_sweepq.cuh: every axis needs a kernel of at least 8
// elements, D = 2T+2 >= 8, and on-chip rings) or, when the axis offers the q kernels too few units of work to fill the GPU,
// by the pass-parallel kernels (fb_sweepp.cuh: any kernel whose rings fit in shared memory).
struct SweepPPlan { bool ok; int DL, RL, NL; size_t smem; };
SweepPPlan sweepp_plan(int npass, int T);
enum { FB_AXIS_NONE = 0, FB_AXIS_Q = 1, FB_AXIS_P = 2 };
inline int axis_mode(const fb_problem *pr, int m) { return m == 0 ? 1 : (m == pr->dim - 1 ? 2 : 0); }
// units of work (16-line groups) the q kernel of axis m has
inline long long axis_q_items(const fb_problem *pr, const Derived &d, int m)
{
    const long long nf = pr->nfields;
    if (m == 0) return nf * d.Dz * ((d.H + 15) / 16);
    if (m == 1) return pr->dim == 2 ? nf * ((d.W + 15) / 16) : nf * d.Dz * ((d.W + 15) / 16);
    return nf * ((d.H * d.W + 15) / 16);
}
int axis_path(const fb_problem *pr, const Derived &d, int m)
{
    const bool q_ok = g_sweepq.load() != 0 && sweepq_plan(pr->num_iter, axis_mode(pr, m), 2 * d.ax[m].T + 2).ok;
    const int pm = g_sweepp.load();
    const bool p_ok = pm != 0 && sweepp_plan(pr->num_iter, d.ax[m].T).ok;
    if (p_ok && pm >= 2) return FB_AXIS_P;
    // measured on the bench grid (2400 x 1200, T = 27, n = 4): the pass-parallel kernel wins up to 150 q units, ties at 225
    // (large batches of kernels the q path does not cover -- fewer than 8 elements -- stay with the first-generation
    // kernel, whose throughput is higher than the pass-parallel kernel's)
    if (p_ok && 2 * axis_q_items(pr, d, m) <= 3LL * sm_count_current()) return FB_AXIS_P;
    return q_ok ? FB_AXIS_Q : FB_AXIS_NONE;
}
// the grid runs on interleaved nodes when every axis has a kernel for them
bool use_nodes(const fb_problem *pr, const Derived &d)
{
    if (pr->dim < 2 || (pr->flags & FB_FLAG_FP32)) return false;
    // tensor-map coordinates and the kernels' row offsets are 32-bit
    if (d.W * d.H > (1LL << 28) || (long long)pr->nfields * d.Dz > (1LL << 30)) return false;
    for (int m = 0; m < pr->dim; ++m)
        if (axis_path(pr, d, m) == FB_AXIS_NONE) return false;
    return true;
}
// ... and on the q kernels alone (the z-slab calls, the binned injection)
bool use_sweepq(const fb_problem *pr, const Derived &d)
{
    if (pr->dim < 2 || (pr->flags & FB_FLAG_FP32) || g_sweepq.load() == 0) return false;
    if (d.W * d.H > (1LL << 28) || (long long)pr->nfields * d.Dz > (1LL << 30)) return false;
    for (int m = 0; m < pr->dim; ++m)
        if (!sweepq_plan(pr->num_iter, axis_mode(pr, m), 2 * d.ax[m].T + 2).ok) return false;
    return true;
}

// The q path can take its x-sweep input from the binned samples instead of a dense injected grid.
bool use_sparse(const fb_problem *pr, const Derived &d, long long nsamples, long long max_n)
{
    if (g_sparse.load() == 0 || !use_sweepq(pr, d) || axis_path(pr, d, 0) != FB_AXIS_Q) return false;
    if (max_n >= FB_BIN_MAX_SAMPLES || (nsamples << pr->dim) >= 0xfffffff0LL) return false;      // record key / index bits
    const FbBins bn = bins_geometry(pr->num_iter, d.ax[0].T, d.W, d.H, (long long)pr->nfields * d.Dz);
    if (!sweepq_plan(pr->num_iter, 1, 2 * d.ax[0].T + 2, 256, sweepqs_prod_bytes(bn.NB), 2).ok) return false;
    // the sparse x sweep runs with 8 pass + 8 producer warps per CTA (register redistribution between the warpgroups)
    const long long items = (long long)pr->nfields * d.Dz * bn.G;
    return items >= 8LL * sm_count_current() && bn.nbuckets < (1LL << 31) && d.W < (1LL << 27);
}

// min / max of the values, then sort by cell + segmented reduce (fb_sparse.cuh): no dense grid is touched
int run_inject_sparse(const fb_problem *pr, const Derived &d, long long nsamples, const int64_t *h_offsets, const double *d_pts,
                      const double *d_val, Workspace &w, cudaStream_t st)
{
    long long max_n = 0;
    int rc = check_offsets(pr, nsamples, h_offsets, max_n);
    if (rc != FB_OK) return rc;
    const FbBins bn = bins_geometry(pr->num_iter, d.ax[0].T, d.W, d.H, (long long)pr->nfields * d.Dz);
    if (bn.nbuckets > w.bin_cap) return fail(FB_EINVAL, "internal: bucket arrays too small (%lld > %lld)", bn.nbuckets, w.bin_cap);
    CUDA_TRY(cudaMemsetAsync(w.bin_cnt, 0, (size_t)(bn.nbuckets + 1) * 4, st));
    fb_init_kernel<<<(unsigned)((pr->nfields + 255) / 256), 256, 0, st>>>(w.mm, pr->nfields, w.counters);
    LAUNCH_CHECK();
    if ((rc = prof_mark(1, st)) != FB_OK) return rc;
    FbSamples s{};
    s.pts = d_pts;
    s.val = d_val;
    if (h_offsets) {
        CUDA_TRY(cudaMemcpyAsync(w.offsets, h_offsets, (size_t)(pr->nfields + 1) * 8, cudaMemcpyHostToDevice, st));
        s.offsets = w.offsets;
        s.n_uniform = 0;
    } else {
        s.offsets = nullptr;
        s.n_uniform = max_n;
    }
    const long long nscan = bn.nbuckets;
    const int nblocks = (int)((nscan + FB_SCAN_BLOCK * FB_SCAN_PER_THREAD - 1) / (FB_SCAN_BLOCK * FB_SCAN_PER_THREAD));
    if (max_n > 0) {
        FbGrid gr{};
        gr.dim = pr->dim;
        gr.W = d.W; gr.H = d.H; gr.Dz = d.Dz; gr.total = d.total;
        gr.z_off = 0; gr.z_cnt = d.Dz;
        for (int m = 0; m < 3; ++m) { gr.x0[m] = pr->x0[m]; gr.step[m] = pr->step[m]; }
        const unsigned nf = (unsigned)pr->nfields;
        long long mmb = (max_n + 256 * 8 - 1) / (256 * 8);
        if (mmb > 1024) mmb = 1024;
        fb_minmax_kernel<<<dim3((unsigned)mmb, nf), 256, 0, st>>>(s, w.mm);
        LAUNCH_CHECK();
        const dim3 sg((unsigned)((max_n + 255) / 256), nf);
        fb_bin_count_kernel<<<sg, 256, 0, st>>>(s, gr, bn, w.bin_cnt);
        LAUNCH_CHECK();
        fb_scan_local_kernel<<<(unsigned)nblocks, FB_SCAN_BLOCK, 0, st>>>(w.bin_cnt, w.bin_start, w.bin_sums, nscan);
        LAUNCH_CHECK();
        fb_scan_sums_kernel<<<1, 1024, 0, st>>>(w.bin_sums, nblocks, w.bin_sums + nblocks + 1);
        LAUNCH_CHECK();
        fb_scan_add_kernel<<<(unsigned)nblocks, FB_SCAN_BLOCK, 0, st>>>(w.bin_start, w.bin_sums, w.bin_sums + nblocks + 1, nscan);
        LAUNCH_CHECK();
        fb_bin_fill_kernel<<<sg, 256, 0, st>>>(s, gr, bn, w.mm, w.bin_start, w.bin_cnt, w.bin_rec);
        LAUNCH_CHECK();
        long long rb = (bn.nbuckets + 32 * FB_RED_WARPS - 1) / (32 * FB_RED_WARPS);
        if (rb > 148 * 16) rb = 148 * 16;
        fb_bin_reduce_kernel<<<(unsigned)rb, 32 * FB_RED_WARPS, 0, st>>>(bn, w.bin_start, w.bin_rec);
        LAUNCH_CHECK();
    } else {
        CUDA_TRY(cudaMemsetAsync(w.bin_start, 0, (size_t)(bn.nbuckets + 1) * 4, st));    // no samples: every bucket is empty
    }
    return FB_OK;
}

// the injection writes interleaved (value, weight) nodes exactly when the q path consumes them
bool inject_interleaved(const fb_problem *pr, const Derived &d) { return use_nodes(pr, d); }

// ---- 1D grids: the bit-exact walk of a long line (fb_line1d_kernel) ------------------------------------------------
struct Line1DPlan { bool ok; int DL, RL; size_t smem; };
Line1DPlan line1d_plan(int npass, int T)
{
    Line1DPlan q{false, 0, 0, 0};
    if (npass < 1 || npass > FB_MAX_FUSED_PASSES) return q;
    const int T1 = T + 1, D = 2 * T + 2;
    q.DL = (T1 + FBL_U - 1) / FBL_U + 1;
    const int ahead = q.DL > FBL_PD ? q.DL : FBL_PD;
    int need = D + FBL_U * (ahead + 3), rl = 64;
    while (rl < need) rl *= 2;
    q.RL = rl;
    q.smem = (size_t)2 * (npass + 1) * (rl + 2) * sizeof(double);       // rings are staggered by 2 doubles (bank conflicts)
    q.ok = q.smem <= kSmemLimit;
    return q;
}

template <int NPASS>
int launch_line1d_t(const FbLine1D &p, const Line1DPlan &q, long long nfields, cudaStream_t st)
{
    static thread_local bool configured[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 15]) {
        CUDA_TRY(cudaFuncSetAttribute(fb_line1d_kernel<NPASS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
        configured[dev & 15] = true;
    }
    fb_line1d_kernel<NPASS, 2><<<(unsigned)nfields, 64, q.smem, st>>>(p);
    LAUNCH_CHECK();
    return FB_OK;
}

int launch_line1d(int npass, const FbLine1D &p, const Line1DPlan &q, long long nfields, cudaStream_t st)
{
    switch (npass) {
    case 1: return launch_line1d_t<1>(p, q, nfields, st);
    case 2: return launch_line1d_t<2>(p, q, nfields, st);
    case 3: return launch_line1d_t<3>(p, q, nfields, st);
    case 4: return launch_line1d_t<4>(p, q, nfields, st);
    case 5: return launch_line1d_t<5>(p, q, nfields, st);
    case 6: return launch_line1d_t<6>(p, q, nfields, st);
    }
    return fail(FB_EINVAL, "unsupported number of fused passes: %d", npass);
}

// ---- small batches: pass-parallel sweeps (fb_sweepp_kernel) -----------------------------------------------------------
SweepPPlan sweepp_plan(int npass, int T)
{
    SweepPPlan q{false, 0, 0, 0, 0};
    if (npass < 1 || npass > FB_MAX_FUSED_PASSES) return q;
    const int T1 = T + 1, D = 2 * T + 2;
    q.DL = (T1 + FBP_U - 1) / FBP_U + 1;
    // a stream must hold what lies between its newest write and the oldest element still to be read:
    // the rows themselves U (PD + 1) + D, the output of a pass U (DL + 1) + T + 1
    int need = FBP_U * (FBP_PD + 1) + D;
    if (FBP_U * (q.DL + 1) + T1 > need) need = FBP_U * (q.DL + 1) + T1;
    q.RL = (need + FBP_U - 1) / FBP_U * FBP_U;
    const int lpl = npass <= 1 ? 2 : (npass <= 2 ? 4 : (npass <= 4 ? 8 : 16));
    q.NL = 32 / lpl;
    q.smem = (size_t)q.NL * (npass + 1) * 2 * (q.RL + 2) * sizeof(double);
    q.ok = q.smem <= kSmemLimit;
    return q;
}

template <int NPASS, int MODE>
int launch_sweepp_t(const FbSweepP &p, const SweepPPlan &q, cudaStream_t st)
{
    static thread_local bool configured[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 15]) {
        CUDA_TRY(cudaFuncSetAttribute(fb_sweepp_kernel<NPASS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
        CUDA_TRY(cudaFuncSetAttribute(fb_sweepp_kernel<NPASS, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
        configured[dev & 15] = true;
    }
    const long long items = p.n_outer * p.n_groups;
    if (items <= 0) return FB_OK;
    if (items > 2147483647LL) return fail(FB_EINVAL, "too many line groups for the small-batch sweep: %lld", items);
    fb_sweepp_kernel<NPASS, MODE><<<(unsigned)items, 64, q.smem, st>>>(p);
    LAUNCH_CHECK();
    return FB_OK;
}

template <int MODE>
int launch_sweepp_m(int npass, const FbSweepP &p, const SweepPPlan &q, cudaStream_t st)
{
    switch (npass) {
    case 1: return launch_sweepp_t<1, MODE>(p, q, st);
    case 2: return launch_sweepp_t<2, MODE>(p, q, st);
    case 3: return launch_sweepp_t<3, MODE>(p, q, st);
    case 4: return launch_sweepp_t<4, MODE>(p, q, st);
    case 5: return launch_sweepp_t<5, MODE>(p, q, st);
    case 6: return launch_sweepp_t<6, MODE>(p, q, st);
    }
    return fail(FB_EINVAL, "unsupported number of fused passes: %d", npass);
}

// one axis: src / dst are grids of interleaved (value, weight) nodes
int run_sweepp(int mode, int num_iter, const AxisParams &ax, const double *src, double *dst, float *out32, double *out64,
               const unsigned long long *mm, double csf, long long n_outer, long long L, long long n_inner, cudaStream_t st)
{
    const SweepPPlan q = sweepp_plan(num_iter, ax.T);
    if (!q.ok) return fail(FB_EKERNEL, "internal: the small-batch sweep does not cover T=%d, num_iter=%d", ax.T, num_iter);
    if (L > 2147483647LL - 64LL * (ax.T + 2) * (num_iter + 1)) return fail(FB_EINVAL, "line too long: %lld", L);
    FbSweepP p{};
    p.in = src; p.out = dst; p.out32 = out32; p.out64 = out64; p.mm = mm;
    p.n_outer = n_outer; p.L = L; p.n_inner = n_inner; p.n_groups = (n_inner + q.NL - 1) / q.NL;
    p.T = ax.T; p.D = 2 * ax.T + 2; p.DL = q.DL; p.RL = q.RL;
    p.alpha = ax.alpha; p.csf = csf;
    if (mode == 0) return launch_sweepp_m<0>(num_iter, p, q, st);
    if (mode == 1) return launch_sweepp_m<1>(num_iter, p, q, st);
    return launch_sweepp_m<2>(num_iter, p, q, st);
}

int run_sweeps(const fb_problem *pr, const Derived &d, Workspace &w, float *d_out, double *d_out64, cudaStream_t st,
               const SegPlan &sp, bool sparse = false)
{
    const long long nf = pr->nfields;
    const int n = pr->num_iter;
    Pair cur{w.vA, w.wA}, spare{w.vB, w.wB};
    SweepCounters ctr{w.counters + 4, 0};
    int rc;
    if (pr->dim == 1 && sp.on) {
        // segmented 1D: gather the extended segments side by side, sweep them as independent lines
        // (transposing output -> one contiguous run per segment), finalise
        dim3 gg((unsigned)((sp.Le + 31) / 32), (unsigned)((sp.n_seg + 31) / 32), 2);
        if (gg.y > 65535) return fail(FB_EINVAL, "too many segments");
        fb_seg_gather_kernel<<<gg, 256, 0, st>>>(w.vA, w.wA, w.seg[0], w.seg[1], d.W, sp.seg_len, sp.halo, sp.n_seg, sp.Le);
        LAUNCH_CHECK();
        Pair scur{w.seg[0], w.seg[1]}, sspare{w.seg[2], w.seg[3]};
        rc = run_sweep(1, n, d.ax[0], scur, sspare, nullptr, nullptr, w.mm, d.csf, 1, sp.Le, sp.n_seg, true, st, ctr);
        if (rc != FB_OK) return rc;
        fb_seg_finalize_kernel<<<(unsigned)((d.W + 255) / 256), 256, 0, st>>>(scur.v, scur.w, d_out, d_out64, d.W, sp.seg_len,
                                                                             sp.halo, sp.Le, w.mm, d.csf);
        LAUNCH_CHECK();
        return prof_mark(3, st);
    }
    if (pr->dim == 1) {
        const Line1DPlan lq = line1d_plan(n, d.ax[0].T);
        if (g_line1d.load() != 0 && lq.ok && d.W >= 1024 && nf <= 65535) {
            // long lines: the 2 n (field, pass) chains as lanes of one warp, a second warp feeds and finalises
            FbLine1D p{};
            p.in_v = cur.v; p.in_w = cur.w; p.out_v = p.out_w = nullptr; p.out32 = d_out; p.out64 = d_out64; p.mm = w.mm;
            p.L = d.W; p.T = d.ax[0].T; p.D = 2 * d.ax[0].T + 2; p.DL = lq.DL; p.RL = lq.RL;
            p.alpha = d.ax[0].alpha; p.csf = d.csf;
            rc = launch_line1d(n, p, lq, nf, st);
            if (rc != FB_OK) return rc;
            return prof_mark(3, st);
        }
        rc = run_sweep(2, n, d.ax[0], cur, spare, d_out, d_out64, w.mm, d.csf, nf, d.W, 1, true, st, ctr);
        if (rc != FB_OK) return rc;
        return prof_mark(3, st);
    }
    if (pr->flags & FB_FLAG_FP32) {
        // fp32 working precision: the B buffers hold interleaved float2 (value, weight) nodes
        const fb_f2 *a2 = reinterpret_cast<const fb_f2 *>(w.vA);     // injected float2 nodes, [..][x][y]
        fb_f2 *b2 = reinterpret_cast<fb_f2 *>(w.vB);
        rc = run_sweep32(1, n, d.ax[0], a2, b2, nullptr, w.mm, d.csf, nf * d.Dz, d.W, d.H, st, ctr);
        if (rc != FB_OK) return rc;
        if ((rc = prof_mark(3, st)) != FB_OK) return rc;
        if (pr->dim == 2) {
            rc = run_sweep32(2, n, d.ax[1], b2, nullptr, d_out, w.mm, d.csf, nf, d.H, d.W, st, ctr);
            if (rc != FB_OK) return rc;
            return prof_mark(4, st);
        }
        fb_f2 *c2 = reinterpret_cast<fb_f2 *>(w.wB);
        rc = run_sweep32(0, n, d.ax[1], b2, c2, nullptr, w.mm, d.csf, nf * d.Dz, d.H, d.W, st, ctr);
        if (rc != FB_OK) return rc;
        if ((rc = prof_mark(4, st)) != FB_OK) return rc;
        rc = run_sweep32(2, n, d.ax[2], c2, nullptr, d_out, w.mm, d.csf, nf, d.Dz, d.H * d.W, st, ctr);
        if (rc != FB_OK) return rc;
        return prof_mark(5, st);
    }
    if (use_nodes(pr, d)) {
        // interleaved nodes throughout: A (injected, [..][x][y]) -> B (natural order) -> float32 field; per axis the q kernel
        // or, for small batches, the pass-parallel kernel
        double *a2 = w.vA, *b2 = w.vB;                   // vB and wB are adjacent: one block of 2 g bytes
        const bool px = axis_path(pr, d, 0) == FB_AXIS_P, py = axis_path(pr, d, 1) == FB_AXIS_P;
        if (sparse) {
            // the x sweep synthesises its rows from the binned samples (run_inject_sparse)
            const int D = 2 * d.ax[0].T + 2;
            const FbBins bn = bins_geometry(n, d.ax[0].T, d.W, d.H, nf * d.Dz);
            const SweepQPlan q = sweepq_plan(n, 1, D, 256, sweepqs_prod_bytes(bn.NB), 2);
            FbSweepQ p{};
            p.in = nullptr; p.out = b2; p.mm = w.mm;
            p.n_outer = nf * d.Dz; p.L = d.W; p.n_inner = d.H; p.n_groups = (d.H + 15) / 16;
            p.T = d.ax[0].T; p.D = D; p.R = q.R; p.RP = q.RP;
            p.alpha = d.ax[0].alpha; p.csf = d.csf;
            p.work_counter = ctr.base + (ctr.next++ % kSweepCounterSlots);
            p.nst = q.nst; p.pf = 0;
            p.tmem_cols_per_warp = q.cols_per_warp;
            p.smem_per_warp = q.smem_per_warp; p.off_ring = q.off_ring;
            p.nb = bn.NB;
            p.bin_start = w.bin_start;
            p.nodes = w.bin_rec;
            rc = launch_sweepqs(n, p, q, st);
        } else if (px) {
            rc = run_sweepp(1, n, d.ax[0], a2, b2, nullptr, nullptr, w.mm, d.csf, nf * d.Dz, d.W, d.H, st);
        } else {
            rc = run_sweepq(1, n, d.ax[0], a2, b2, nullptr, nullptr, w.mm, d.csf, nf * d.Dz, d.W, d.H, st, ctr);
        }
        if (rc != FB_OK) return rc;
        if ((rc = prof_mark(3, st)) != FB_OK) return rc;
        if (pr->dim == 2) {
            if (py) rc = run_sweepp(2, n, d.ax[1], b2, nullptr, d_out, d_out64, w.mm, d.csf, nf, d.H, d.W, st);
            else rc = run_sweepq(2, n, d.ax[1], b2, nullptr, d_out, d_out64, w.mm, d.csf, nf, d.H, d.W, st, ctr);
            if (rc != FB_OK) return rc;
            return prof_mark(4, st);
        }
        // y sweep in place (the q kernel's pass 1 re-reads a row before the last pass overwrites it: it needs >= 2 passes;
        // the pass-parallel kernel reads every row once, ahead of its writes)
        double *c2 = (py || n >= 2) ? b2 : a2;
        if (py) rc = run_sweepp(0, n, d.ax[1], b2, c2, nullptr, nullptr, w.mm, d.csf, nf * d.Dz, d.H, d.W, st);
        else rc = run_sweepq(0, n, d.ax[1], b2, c2, nullptr, nullptr, w.mm, d.csf, nf * d.Dz, d.H, d.W, st, ctr);
        if (rc != FB_OK) return rc;
        if ((rc = prof_mark(4, st)) != FB_OK) return rc;
        if (axis_path(pr, d, 2) == FB_AXIS_P)
            rc = run_sweepp(2, n, d.ax[2], c2, nullptr, d_out, d_out64, w.mm, d.csf, nf, d.Dz, d.H * d.W, st);
        else
            rc = run_sweepq(2, n, d.ax[2], c2, nullptr, d_out, d_out64, w.mm, d.csf, nf, d.Dz, d.H * d.W, st, ctr);
        if (rc != FB_OK) return rc;
        return prof_mark(5, st);
    }
    // x sweep: A layout [..][x][y] -> natural layout [..][y][x]
    rc = run_sweep(1, n, d.ax[0], cur, spare, nullptr, nullptr, w.mm, d.csf, nf * d.Dz, d.W, d.H, true, st, ctr);
    if (rc != FB_OK) return rc;
    if ((rc = prof_mark(3, st)) != FB_OK) return rc;
    if (pr->dim == 2) {
        rc = run_sweep(2, n, d.ax[1], cur, spare, d_out, d_out64, w.mm, d.csf, nf, d.H, d.W, true, st, ctr);
        if (rc != FB_OK) return rc;
        return prof_mark(4, st);
    }
    rc = run_sweep(0, n, d.ax[1], cur, spare, nullptr, nullptr, w.mm, d.csf, nf * d.Dz, d.H, d.W, true, st, ctr);
    if (rc != FB_OK) return rc;
    if ((rc = prof_mark(4, st)) != FB_OK) return rc;
    rc = run_sweep(2, n, d.ax[2], cur, spare, d_out, d_out64, w.mm, d.csf, nf, d.Dz, d.H * d.W, true, st, ctr);
    if (rc != FB_OK) return rc;
    return prof_mark(5, st);
}

int check_kernel_vs_grid(const fb_problem *pr, const Derived &d)
{
    // interpolation.py:171-175 / :180-184: the rectangular kernel must be smaller than the grid
    for (int m = 0; m < pr->dim; ++m)
        if (d.ks[m] >= pr->size[m])
            return fail(FB_EKERNEL, "resulting rectangular kernel size should be smaller w.r.t. specified grid: axis %d kernel %d grid %lld",
                        m, d.ks[m], (long long)pr->size[m]);
    return FB_OK;
}

int pipeline(const fb_problem *pr, long long nsamples, const int64_t *h_offsets, const double *d_pts,
             const double *d_val, float *d_out, double *d_out64, void *d_ws, long long ws_bytes,
             cudaStream_t st, bool enforce_kernel_check)
{
    Derived d;
    int rc = derive(pr, d);
    if (rc != FB_OK) return rc;
    if (enforce_kernel_check) {
        rc = check_kernel_vs_grid(pr, d);
        if (rc != FB_OK) return rc;
    }
    if (pr->flags & FB_FLAG_FP32) {
        if (pr->dim < 2) return fail(FB_EINVAL, "the fp32 path covers 2D and 3D grids (a 1D grid is a single line: use the fp64 path)");
        if (d_out64) return fail(FB_EINVAL, "the fp32 path has no fp64 quotient output");
    }
    Workspace w;
    const SegPlan sp = seg_plan(pr, d);
    carve(w, (char *)d_ws, pr, d.total, nsamples, &sp);
    if ((long long)w.bytes > ws_bytes) return fail(FB_ENOMEM, "workspace too small: need %zu bytes, got %lld", w.bytes, ws_bytes);
    g_prof.launches_begin = g_launches.load();
    g_prof.marked = 0;
    if ((rc = prof_mark(0, st)) != FB_OK) return rc;
    long long max_n = 0;
    if ((rc = check_offsets(pr, nsamples, h_offsets, max_n)) != FB_OK) return rc;
    const bool sparse = use_sparse(pr, d, nsamples, max_n);
    if (sparse)
        rc = run_inject_sparse(pr, d, nsamples, h_offsets, d_pts, d_val, w, st);
    else
        rc = run_inject(pr, d, nsamples, h_offsets, d_pts, d_val, w, st, 0, -1, (pr->flags & FB_FLAG_FP32) != 0,
                        inject_interleaved(pr, d));
    if (rc != FB_OK) return rc;
    if ((rc = prof_mark(2, st)) != FB_OK) return rc;
    rc = run_sweeps(pr, d, w, d_out, d_out64, st, sp, sparse);
    if (rc != FB_OK) return rc;
    g_prof.launches_end = g_launches.load();
    g_prof.armed = g_profiling.load() != 0;
    return FB_OK;
}

// ---- device arena for the *_host entry points ------------------------------------------------------
struct Arena {
    void *ptr = nullptr;
    size_t bytes = 0;
    int device = -1;
};
std::mutex g_arena_mutex;
Arena g_arena[2];   // 0: workspace, 1: staging (inputs / outputs)

int arena_get(int which, size_t bytes, void **out)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    Arena &a = g_arena[which];
    if (a.device != dev || a.bytes < bytes) {
        if (a.ptr) {
            cudaSetDevice(a.device);
            cudaFree(a.ptr);
            cudaSetDevice(dev);
            a.ptr = nullptr;
            a.bytes = 0;
        }
        size_t want = align_up(bytes + bytes / 8, 1 << 20);
        cudaError_t e = cudaMalloc(&a.ptr, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = align_up(bytes, 1 << 20);
            e = cudaMalloc(&a.ptr, want);
        }
        if (e != cudaSuccess) {
            a.ptr = nullptr;
            return fail(FB_ENOMEM, "cudaMalloc of %zu bytes failed: %s", want, cudaGetErrorString(e));
        }
        a.bytes = want;
        a.device = dev;
    }
    *out = a.ptr;
    return FB_OK;
}

int require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1) {
        cudaGetLastError();
        return fail(FB_ECUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    return FB_OK;
}

struct Staging {
    char *base;
    size_t off = 0;
    explicit Staging(char *b) : base(b) {}
    template <typename T> T *take(size_t n) { T *p = (T *)(base + off); off += align_up(n * sizeof(T)); return p; }
};

}  // namespace

// ====================================================================================================
FB_EXPORT const char *fb_last_error(void) { return g_err.c_str(); }
FB_EXPORT int fb_version(void) { return 100; }

FB_EXPORT int fb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

FB_EXPORT int fb_set_device(int device)
{
    int rc = require_device();
    if (rc != FB_OK) return rc;
    CUDA_TRY(cudaSetDevice(device));
    return FB_OK;
}

// interpolation.py:549-552
FB_EXPORT int32_t fb_half_kernel_size_opt(double sigma, double step, int num_iter)
{
    const double s = sigma / step;
    return (int32_t)((std::sqrt(1.0 + 12 * s * s / num_iter) - 1.0) / 2.0);
}

// interpolation.py:783-785
FB_EXPORT int32_t fb_half_kernel_size(double sigma, double step, int num_iter)
{
    return (int32_t)(std::sqrt(3.0 / num_iter) * sigma / step + 0.5);
}

// interpolation.py:561-569
FB_EXPORT double fb_tail_value(double sigma, double step, int num_iter)
{
    const long long hks = fb_half_kernel_size_opt(sigma, step, num_iter);
    const long long ks = 2 * hks + 1;
    const double sigma_rect_sqr = (double)((hks + 1) * hks) / 3.0 * (step * step);
    const double hs = (double)(hks + 1) * step;
    return 0.5 * (double)ks * (sigma * sigma / num_iter - sigma_rect_sqr) / (hs * hs - sigma * sigma / num_iter);
}

// interpolation.py:424-425
FB_EXPORT double fb_conv_scale_factor(int dim, const int32_t *kernel_size, const double *tail_value,
                                      const double *sigma, const double *step, int num_iter,
                                      double max_dist_weight)
{
    double prod = 1.0;
    for (int m = 0; m < dim; ++m) {
        const double f = int_power((double)kernel_size[m] + 2 * tail_value[m], num_iter) / std::sqrt(2 * M_PI) /
                         (sigma[m] / step[m]);
        prod *= f;
    }
    return prod * max_dist_weight;
}

FB_EXPORT int64_t fb_workspace_bytes(const fb_problem *prob, int64_t nsamples)
{
    Derived d;
    if (derive(prob, d) != FB_OK) return FB_EINVAL;
    Workspace w;
    const SegPlan sp = seg_plan(prob, d);
    carve(w, nullptr, prob, d.total, nsamples, &sp);
    return (int64_t)w.bytes;
}

FB_EXPORT int fb_barnes_dev(const fb_problem *prob, int64_t nsamples, const int64_t *sample_offsets,
                            const double *d_pts, const double *d_val, float *d_out, double *d_out64,
                            void *d_workspace, int64_t workspace_bytes, void *stream)
{
    int rc = require_device();
    if (rc != FB_OK) return rc;
    if (!d_pts || !d_val || !d_out || !d_workspace) return fail(FB_EINVAL, "null device pointer");
    return pipeline(prob, nsamples, sample_offsets, d_pts, d_val, d_out, d_out64, d_workspace, workspace_bytes,
                    (cudaStream_t)stream, true);
}

// Host-buffer entry.  Large batches are cut into chunks of fields that run round-robin on a few
// streams, each with its own workspace slice, so that the H2D copy of chunk i+1, the kernels of
// chunk i and the D2H copy of chunk i-1 overlap (the float32 output dominates: 4 B per grid point
// over PCIe).  Chunking never changes results: fields are independent.
namespace {
constexpr int kHostStreams = 3;
cudaStream_t g_host_streams[kHostStreams] = {nullptr, nullptr, nullptr};
int g_host_streams_device = -1;
std::atomic<int> g_host_chunk_fields{16};   // 16 x 75 line groups >= 8 per SM: the chunks run the q / tensor-memory kernels

int host_streams_get()
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (g_host_streams_device != dev) {
        for (int i = 0; i < kHostStreams; ++i) {
            if (g_host_streams[i]) cudaStreamDestroy(g_host_streams[i]);
            CUDA_TRY(cudaStreamCreateWithFlags(&g_host_streams[i], cudaStreamNonBlocking));
        }
        g_host_streams_device = dev;
    }
    return FB_OK;
}
}  // namespace

FB_EXPORT int fb_barnes_host(const fb_problem *prob, int64_t nsamples, const int64_t *sample_offsets,
                             const double *pts, const double *val, float *out, double *out64)
{
    int rc = require_device();
    if (rc != FB_OK) return rc;
    if (!prob || !pts || !val || !out) return fail(FB_EINVAL, "null pointer");
    Derived d;
    if ((rc = derive(prob, d)) != FB_OK) return rc;
    if ((rc = check_kernel_vs_grid(prob, d)) != FB_OK) return rc;
    long long max_n = 0;
    if ((rc = check_offsets(prob, nsamples, sample_offsets, max_n)) != FB_OK) return rc;
    std::lock_guard<std::mutex> lock(g_arena_mutex);

    // chunking: cf fields per chunk, ns streams
    const long long nf = prob->nfields;
    long long cf = g_host_chunk_fields.load();
    if (cf < 1) cf = 1;
    const bool chunked = nf >= 2 * cf;
    if (!chunked) cf = nf;
    const int ns = chunked ? kHostStreams : 1;
    // chunk boundaries: the first chunk is a quarter of the others -- nothing overlaps its upload and its kernels, and the
    // device -> host copies, which bound the call, start that much earlier
    std::vector<long long> cb;
    cb.push_back(0);
    if (chunked && cf >= 4) cb.push_back(cf / 4);
    while (cb.back() < nf) cb.push_back(cb.back() + cf < nf ? cb.back() + cf : nf);
    const long long nchunks = (long long)cb.size() - 1;

    // per-slot sizes: worst-case chunk
    fb_problem cp = *prob;
    cp.nfields = cf;
    long long max_chunk_samples = 0;
    for (long long c = 0; c < nchunks; ++c) {
        const long long b0 = cb[c], b1 = cb[c + 1];
        const long long n = sample_offsets ? sample_offsets[b1] - sample_offsets[b0] : (b1 - b0) * max_n;
        if (n > max_chunk_samples) max_chunk_samples = n;
    }
    Workspace w;
    const SegPlan sp = seg_plan(&cp, d);
    carve(w, nullptr, &cp, d.total, max_chunk_samples, &sp);
    const size_t ws_slot = align_up(w.bytes);
    const size_t cgrid = (size_t)cf * d.total;
    const size_t stg_slot = align_up((size_t)max_chunk_samples * prob->dim * 8) + align_up((size_t)max_chunk_samples * 8) +
                            align_up(cgrid * 4) + (out64 ? align_up(cgrid * 8) : 0) + 1024;
    void *ws = nullptr, *stg = nullptr;
    if ((rc = arena_get(0, ws_slot * ns, &ws)) != FB_OK) return rc;
    if ((rc = arena_get(1, stg_slot * ns, &stg)) != FB_OK) return rc;
    if (chunked && (rc = host_streams_get()) != FB_OK) return rc;

    std::vector<int64_t> rebased;
    for (long long c = 0; c < nchunks; ++c) {
        const int slot = (int)(c % ns);
        cudaStream_t st = chunked ? g_host_streams[slot] : (cudaStream_t)0;
        const long long b0 = cb[c], b1 = cb[c + 1];
        const long long s0 = sample_offsets ? sample_offsets[b0] : b0 * max_n;
        const long long s1 = sample_offsets ? sample_offsets[b1] : b1 * max_n;
        const long long n = s1 - s0;
        cp.nfields = b1 - b0;
        const int64_t *offs = nullptr;
        if (sample_offsets) {
            // the pipeline copies the offsets to the device asynchronously: keep every chunk's copy alive
            const size_t base = rebased.size();
            if (c == 0) rebased.reserve((size_t)(nf + nchunks));
            for (long long b = b0; b <= b1; ++b) rebased.push_back(sample_offsets[b] - s0);
            offs = rebased.data() + base;
        }
        Staging s((char *)stg + (size_t)slot * stg_slot);
        double *d_pts = s.take<double>((size_t)max_chunk_samples * prob->dim);
        double *d_val = s.take<double>((size_t)max_chunk_samples);
        float *d_out = s.take<float>(cgrid);
        double *d_out64 = out64 ? s.take<double>(cgrid) : nullptr;
        const size_t g = (size_t)(b1 - b0) * d.total;
        // an error leaves the loop, not the function: work of earlier chunks is still in flight on the other streams and
        // the arena mutex must not be released before they have drained (below)
        auto copy = [&](void *dst, const void *src, size_t bytes, cudaMemcpyKind kind) {
            const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, st);
            if (e != cudaSuccess) rc = fail(FB_ECUDA, "cudaMemcpyAsync failed: %s", cudaGetErrorString(e));
            return e == cudaSuccess;
        };
        if (!copy(d_pts, pts + (size_t)s0 * prob->dim, (size_t)n * prob->dim * 8, cudaMemcpyHostToDevice)) break;
        if (!copy(d_val, val + s0, (size_t)n * 8, cudaMemcpyHostToDevice)) break;
        rc = pipeline(&cp, n, offs, d_pts, d_val, d_out, d_out64, (char *)ws + (size_t)slot * ws_slot, (long long)ws_slot, st, false);
        if (rc != FB_OK) break;
        if (!copy(out + (size_t)b0 * d.total, d_out, g * 4, cudaMemcpyDeviceToHost)) break;
        if (out64 && !copy(out64 + (size_t)b0 * d.total, d_out64, g * 8, cudaMemcpyDeviceToHost)) break;
    }
    if (chunked) {
        for (int i = 0; i < ns; ++i) {
            cudaError_t e = cudaStreamSynchronize(g_host_streams[i]);
            if (e != cudaSuccess && rc == FB_OK) rc = fail(FB_ECUDA, "stream sync failed: %s", cudaGetErrorString(e));
        }
    } else {
        cudaError_t e = cudaStreamSynchronize(0);
        if (e != cudaSuccess && rc == FB_OK) rc = fail(FB_ECUDA, "stream sync failed: %s", cudaGetErrorString(e));
    }
    return rc;
}

// ---- stage entry points -------------------------------------------------------------------------------
namespace {
// work counters for the stage entry points (they have no workspace; serialised by g_arena_mutex, stream 0)
int stage_counters(SweepCounters &ctr)
{
    static unsigned long long *buf[16] = {nullptr};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (!buf[dev & 15]) CUDA_TRY(cudaMalloc(&buf[dev & 15], kSweepCounterSlots * sizeof(unsigned long long)));
    ctr.base = buf[dev & 15];
    ctr.next = 0;
    return FB_OK;
}
}  // namespace

FB_EXPORT int fb_accumulate_lines_host(double *lines, int64_t n_outer, int64_t len, int64_t n_inner,
                                       int64_t rect_len, int num_iter, double alpha)
{
    int rc = require_device();
    if (rc != FB_OK) return rc;
    if (!lines || n_outer < 1 || len < 1 || n_inner < 1 || num_iter < 1 || rect_len < 1 || !(rect_len & 1))
        return fail(FB_EINVAL, "invalid line batch (rect_len must be odd and positive)");
    std::lock_guard<std::mutex> lock(g_arena_mutex);
    const size_t n = (size_t)n_outer * len * n_inner;
    void *stg = nullptr;
    if ((rc = arena_get(1, 2 * align_up(n * 8) + 1024, &stg)) != FB_OK) return rc;
    Staging s((char *)stg);
    Pair cur{s.take<double>(n), nullptr}, spare{s.take<double>(n), nullptr};
    cudaStream_t st = 0;
    CUDA_TRY(cudaMemcpyAsync(cur.v, lines, n * 8, cudaMemcpyHostToDevice, st));
    AxisParams ax{(int)((rect_len - 1) / 2), alpha};
    SweepCounters ctr{nullptr, 0};
    if ((rc = stage_counters(ctr)) != FB_OK) return rc;
    rc = run_sweep(0, num_iter, ax, cur, spare, nullptr, nullptr, nullptr, 0.0, n_outer, len, n_inner, false, st, ctr);
    if (rc != FB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(lines, cur.v, n * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return FB_OK;
}

namespace {
int launch_transpose(const double *in, double *out, long long n_outer, long long rows, long long cols, cudaStream_t st)
{
    // z dimension of the grid is limited to 65535 blocks: loop over slabs of outer indices
    for (long long o0 = 0; o0 < n_outer; o0 += 65535) {
        const long long no = n_outer - o0 < 65535 ? n_outer - o0 : 65535;
        dim3 g((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32), (unsigned)no);
        if (g.y > 65535) return fail(FB_EINVAL, "grid too tall for the transpose helper");
        fb_transpose_kernel<<<g, 256, 0, st>>>(in + o0 * rows * cols, out + o0 * rows * cols, no, rows, cols);
        LAUNCH_CHECK();
    }
    return FB_OK;
}
}  // namespace

FB_EXPORT int fb_convolve_host(int dim, double *vg, double *wg, const int64_t *size, const int32_t *kernel_size,
                               int num_iter, const double *tail_value, double conv_scale_factor)
{
    int rc = require_device();
    if (rc != FB_OK) return rc;
    if (dim < 1 || dim > 3 || !vg || !wg || !size || !kernel_size || num_iter < 1) return fail(FB_EINVAL, "invalid argument");
    const long long W = size[0], H = dim > 1 ? size[1] : 1, Dz = dim > 2 ? size[2] : 1;
    const size_t n = (size_t)W * H * Dz;
    AxisParams ax[3];
    for (int m = 0; m < dim; ++m) {
        if (kernel_size[m] < 1 || !(kernel_size[m] & 1)) return fail(FB_EINVAL, "kernel_size must be odd and positive");
        ax[m].T = (kernel_size[m] - 1) / 2;
        ax[m].alpha = tail_value ? tail_value[m] : 0.0;
    }
    std::lock_guard<std::mutex> lock(g_arena_mutex);
    void *stg = nullptr;
    if ((rc = arena_get(1, 4 * align_up(n * 8) + 1024, &stg)) != FB_OK) return rc;
    Staging s((char *)stg);
    double *v0 = s.take<double>(n), *w0 = s.take<double>(n), *v1 = s.take<double>(n), *w1 = s.take<double>(n);
    cudaStream_t st = 0;
    CUDA_TRY(cudaMemcpyAsync(v0, vg, n * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(w0, wg, n * 8, cudaMemcpyHostToDevice, st));
    Pair cur{v0, w0}, spare{v1, w1};
    SweepCounters ctr{nullptr, 0};
    if ((rc = stage_counters(ctr)) != FB_OK) return rc;
    if (dim == 1) {
        rc = run_sweep(0, num_iter, ax[0], cur, spare, nullptr, nullptr, nullptr, 0.0, 1, W, 1, true, st, ctr);
        if (rc != FB_OK) return rc;
    } else {
        // natural [z][y][x] -> A layout [z][x][y]
        if ((rc = launch_transpose(v0, v1, Dz, H, W, st)) != FB_OK) return rc;
        if ((rc = launch_transpose(w0, w1, Dz, H, W, st)) != FB_OK) return rc;
        cur = Pair{v1, w1};
        spare = Pair{v0, w0};
        rc = run_sweep(1, num_iter, ax[0], cur, spare, nullptr, nullptr, nullptr, 0.0, Dz, W, H, true, st, ctr);
        if (rc != FB_OK) return rc;
        rc = run_sweep(0, num_iter, ax[1], cur, spare, nullptr, nullptr, nullptr, 0.0, Dz, H, W, true, st, ctr);
        if (rc != FB_OK) return rc;
        if (dim == 3) {
            rc = run_sweep(0, num_iter, ax[2], cur, spare, nullptr, nullptr, nullptr, 0.0, 1, Dz, H * W, true, st, ctr);
            if (rc != FB_OK) return rc;
        }
    }
    double *rv = cur.v, *rw = cur.w;
    fb_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rw, (long long)n, conv_scale_factor);
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpyAsync(vg, rv, n * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(wg, rw, n * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return FB_OK;
}

FB_EXPORT int fb_inject_host(const fb_problem *prob, int64_t nsamples, const int64_t *sample_offsets,
                             const double *pts, const double *val, double *vg, double *wg, double *offsets)
{
    int rc = require_device();
    if (rc != FB_OK) return rc;
    if (!prob || !pts || !val || !vg || !wg) return fail(FB_EINVAL, "null pointer");
    Derived d;
    if ((rc = derive(prob, d)) != FB_OK) return rc;
    std::lock_guard<std::mutex> lock(g_arena_mutex);
    Workspace w;
    carve(w, nullptr, prob, d.total, nsamples);
    const size_t npts = (size_t)nsamples * prob->dim, ngrid = (size_t)prob->nfields * d.total;
    // dim 1 has no B buffers in the workspace: stage the transposed copies separately
    const size_t stage_bytes = align_up(npts * 8) + align_up((size_t)nsamples * 8) + 2 * align_up(ngrid * 8) + 1024;
    void *ws = nullptr, *stg = nullptr;
    if ((rc = arena_get(0, w.bytes, &ws)) != FB_OK) return rc;
    if ((rc = arena_get(1, stage_bytes, &stg)) != FB_OK) return rc;
    carve(w, (char *)ws, prob, d.total, nsamples);
    Staging s((char *)stg);
    double *d_pts = s.take<double>(npts);
    double *d_val = s.take<double>((size_t)nsamples);
    double *tv = s.take<double>(ngrid), *tw = s.take<double>(ngrid);
    cudaStream_t st = 0;
    CUDA_TRY(cudaMemcpyAsync(d_pts, pts, npts * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_val, val, (size_t)nsamples * 8, cudaMemcpyHostToDevice, st));
    if ((rc = run_inject(prob, d, nsamples, sample_offsets, d_pts, d_val, w, st)) != FB_OK) return rc;
    const double *rv = w.vA, *rw = w.wA;
    if (prob->dim > 1) {
        // A layout [..][x][y] -> natural [..][y][x]
        if ((rc = launch_transpose(w.vA, tv, prob->nfields * d.Dz, d.W, d.H, st)) != FB_OK) return rc;
        if ((rc = launch_transpose(w.wA, tw, prob->nfields * d.Dz, d.W, d.H, st)) != FB_OK) return rc;
        rv = tv;
        rw = tw;
    }
    CUDA_TRY(cudaMemcpyAsync(vg, rv, ngrid * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(wg, rw, ngrid * 8, cudaMemcpyDeviceToHost, st));
    std::vector<unsigned long long> mm((size_t)prob->nfields * FB_MM_STRIDE);
    CUDA_TRY(cudaMemcpyAsync(mm.data(), w.mm, mm.size() * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (offsets) {
        for (long long b = 0; b < prob->nfields; ++b) {
            const unsigned long long *m = &mm[(size_t)b * FB_MM_STRIDE];
            auto dec = [](unsigned long long k) {
                unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
                double v;
                memcpy(&v, &u, 8);
                return v;
            };
            offsets[b] = m[2] ? NAN : (dec(m[0]) + dec(m[1])) / 2.0;
        }
    }
    return FB_OK;
}

// ---- exact Gaussian sums: methods 'naive', 'radius', 'naive_S2' ("next" row N3) ---------------------------
namespace {
template <int KIND, int DIM>
int launch_exact_t(const FbExact &p, cudaStream_t st)
{
    const long long total = p.W * p.H * p.Dz;
    const long long blocks = (total + FB_EXACT_THREADS - 1) / FB_EXACT_THREADS;
    if (blocks > 2147483647LL) return fail(FB_EINVAL, "grid too large for the exact-sum kernel");
    fb_exact_kernel<KIND, DIM><<<(unsigned)blocks, FB_EXACT_THREADS, 0, st>>>(p);
    LAUNCH_CHECK();
    return FB_OK;
}

int run_exact(const fb_problem *pr, long long n, const double *d_pts, const double *d_val, double min_weight,
              double *d_out, unsigned long long *scratch, cudaStream_t st)
{
    if (!pr) return fail(FB_EINVAL, "null problem");
    const int dim = pr->dim, method = pr->method;
    if (method != FB_METHOD_NAIVE && method != FB_METHOD_RADIUS && method != FB_METHOD_NAIVE_S2)
        return fail(FB_EINVAL, "not an exact-sum method: %d", method);
    if (dim < 1 || dim > 3) return fail(FB_EINVAL, "Barnes interpolation supports only sample points in dimensions 1, 2 or 3");
    if (pr->nfields != 1) return fail(FB_EINVAL, "the exact-sum methods take one field per call");
    if (n < 1) return fail(FB_EINVAL, "no samples");
    if (method == FB_METHOD_RADIUS) {
        // interpolation.py:187-193
        if (dim != 2) return fail(FB_EINVAL, "radius algorithm works only in 2D but data is: %dD", dim);
        if (pr->sigma[0] != pr->sigma[1]) return fail(FB_EINVAL, "radius algorithm in 2D works only for scalar sigma value");
        if (!(min_weight > 0.0)) return fail(FB_EINVAL, "min_weight must be positive");
    }
    if (method == FB_METHOD_NAIVE_S2 && dim != 2) return fail(FB_EINVAL, "naive_S2 needs (lon, lat) sample points");
    FbExact p{};
    p.pts = d_pts;
    p.val = d_val;
    p.n = n;
    p.mm = scratch;
    p.W = pr->size[0];
    p.H = dim > 1 ? pr->size[1] : 1;
    p.Dz = dim > 2 ? pr->size[2] : 1;
    if (p.W < 1 || p.H < 1 || p.Dz < 1) return fail(FB_EINVAL, "grid size must be positive");
    for (int m = 0; m < 3; ++m) {
        p.x0[m] = m < dim ? pr->x0[m] : 0.0;
        p.step[m] = m < dim ? pr->step[m] : 1.0;
        p.scale[m] = m < dim ? 2 * (pr->sigma[m] * pr->sigma[m]) : 1.0;     // scale = 2*sigma**2  (:872)
    }
    if (method == FB_METHOD_RADIUS) {
        const double search_radius = std::sqrt(-2.0 * std::log(min_weight)) * pr->sigma[0];    // :823
        p.radius_sqr = search_radius * search_radius;                                          // kdtree.py:249
        p.max_dist_weight = pr->max_dist_weight;
    }
    p.out = d_out;
    fb_init_kernel<<<1, 32, 0, st>>>(scratch, 1, scratch + FB_MM_STRIDE);
    LAUNCH_CHECK();
    FbSamples sm{};
    sm.pts = d_pts;
    sm.val = d_val;
    sm.offsets = nullptr;
    sm.n_uniform = n;
    long long mmb = (n + 256 * 8 - 1) / (256 * 8);
    if (mmb > 1024) mmb = 1024;
    fb_minmax_kernel<<<dim3((unsigned)mmb, 1), 256, 0, st>>>(sm, scratch);
    LAUNCH_CHECK();
    if (method == FB_METHOD_NAIVE_S2) return launch_exact_t<1, 2>(p, st);
    if (method == FB_METHOD_RADIUS) return launch_exact_t<2, 2>(p, st);
    if (dim == 1) return launch_exact_t<0, 1>(p, st);
    if (dim == 2) return launch_exact_t<0, 2>(p, st);
    return launch_exact_t<0, 3>(p, st);
}
}  // namespace

FB_EXPORT int fb_barnes_exact_dev(const fb_problem *prob, int64_t nsamples, const double *d_pts, const double *d_val,
                                  double min_weight, double *d_out64, void *d_scratch, void *stream)
{
    int rc = require_device();
    if (rc != FB_OK) return rc;
    if (!d_pts || !d_val || !d_out64 || !d_scratch) return fail(FB_EINVAL, "null device pointer");
    return run_exact(prob, nsamples, d_pts, d_val, min_weight, d_out64, (unsigned long long *)d_scratch,
                     (cudaStream_t)stream);
}

FB_EXPORT int fb_barnes_exact_host(const fb_problem *prob, int64_t nsamples, const double *pts, const double *val,
                                   double min_weight, double *out64)
{
    int rc = require_device();
    if (rc != FB_OK) return rc;
    if (!prob || !pts || !val || !out64) return fail(FB_EINVAL, "null pointer");
    if (prob->dim < 1 || prob->dim > 3 || nsamples < 1) return fail(FB_EINVAL, "invalid argument");
    size_t total = 1;
    for (int m = 0; m < prob->dim; ++m) {
        if (prob->size[m] < 1) return fail(FB_EINVAL, "grid size must be positive");
        total *= (size_t)prob->size[m];
    }
    const size_t npts = (size_t)nsamples * prob->dim;
    std::lock_guard<std::mutex> lock(g_arena_mutex);
    void *stg = nullptr;
    if ((rc = arena_get(1, align_up(npts * 8) + align_up((size_t)nsamples * 8) + align_up(total * 8) + 2048, &stg)) != FB_OK)
        return rc;
    Staging s((char *)stg);
    double *d_pts = s.take<double>(npts), *d_val = s.take<double>((size_t)nsamples), *d_out = s.take<double>(total);
    unsigned long long *scratch = s.take<unsigned long long>(32);
    cudaStream_t st = 0;
    CUDA_TRY(cudaMemcpyAsync(d_pts, pts, npts * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_val, val, (size_t)nsamples * 8, cudaMemcpyHostToDevice, st));
    if ((rc = run_exact(prob, nsamples, d_pts, d_val, min_weight, d_out, scratch, st)) != FB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(out64, d_out, total * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return FB_OK;
}

// ---- 3D z-slab decomposition (one slab per GPU) ----------------------------------------------------------
// A rank owns the planes [z_begin, z_begin + z_count) of the volume and keeps an EXTENDED slab with
// halo_lo / halo_hi extra planes below / above.  Phase 1 injects into the own planes (records of
// nodes outside are dropped; every node's ordered sum is complete because all samples are visible)
// and runs the x and y sweeps, which are independent per plane and therefore bit-identical to the
// single-GPU run.  The caller then fills the halo planes of the extended B buffers with the
// neighbours' own planes (NVLink / NCCL send-recv) and phase 2 runs the fused z sweep + finalize
// over the extended lines.  Each of the n tailed passes reaches T+1 planes, so with
// halo >= n*(T_z+1) the own planes see every input they depend on; the only difference to the
// single-GPU run is where the sliding accumulator starts (rounding level).
namespace {
struct SlabLayout {
    double *vA, *wA, *vB, *wB;      // extended slabs: z_ext planes of H*W doubles each
    float *out_ext;                 // z_ext planes of float32
    double *out64_ext;
    Workspace inj;                  // injection scratch (record tables, min/max)
    size_t off_vB, off_wB;          // byte offsets of the extended B buffers
    size_t off_out, off_out64;      // ... and of the extended result buffers
    double *pts_c, *val_c;          // samples that touch the slab, in order
    unsigned int *cmp_cnt, *cmp_start, *cmp_sums;
    long long *cmp_offsets;         // {0, number of compacted samples} on the device
    size_t bytes;
};

int slab_carve(SlabLayout &s, char *base, const fb_problem *pr, const Derived &d, long long nsamples, long long z_ext,
               bool want64)
{
    size_t off = 0;
    auto take = [&](size_t n) { char *p = base ? base + off : nullptr; off += align_up(n); return p; };
    const size_t g = (size_t)d.W * d.H * (size_t)z_ext * sizeof(double);
    // each pair is one block of 2 g bytes: two arrays of planes, or (q path) one array of interleaved (value, weight) nodes
    s.vA = (double *)take(2 * g);
    s.wA = base ? (double *)((char *)s.vA + g) : nullptr;
    s.off_vB = off;
    s.vB = (double *)take(2 * g);
    s.off_wB = s.off_vB + g;
    s.wB = base ? (double *)((char *)s.vB + g) : nullptr;
    s.off_out = off;
    s.out_ext = (float *)take(g / 2);
    s.off_out64 = off;
    s.out64_ext = want64 ? (double *)take(g) : nullptr;
    const size_t R = (size_t)nsamples << pr->dim;
    memset(&s.inj, 0, sizeof s.inj);
    s.inj.mm = (unsigned long long *)take(FB_MM_STRIDE * 8);
    s.inj.counters = (unsigned long long *)take((4 + kSweepCounterSlots) * 8);
    s.inj.offsets = (long long *)take(2 * 8);
    s.inj.first_mask = (unsigned char *)take((size_t)nsamples + 1);
    s.inj.rec_k = (int *)take(R * 4 + 4);
    s.inj.rec_w = (double *)take(R * 8 + 8);
    s.inj.rec_wv = (double *)take(R * 8 + 8);
    s.inj.seg_node = (long long *)take(R * 8 + 8);
    s.inj.seg_base = (unsigned int *)take(R * 4 + 4);
    s.inj.seg_n = (unsigned int *)take(R * 4 + 4);
    s.inj.link_next = (unsigned int *)take(R * 4 + 4);
    // compaction of the samples to those that touch the slab (fb_sparse.cuh)
    const size_t nblk = ((size_t)nsamples + 255) / 256;
    s.pts_c = (double *)take((size_t)nsamples * pr->dim * 8 + 8);
    s.val_c = (double *)take((size_t)nsamples * 8 + 8);
    s.cmp_cnt = (unsigned int *)take((nblk + 1) * 4);
    s.cmp_start = (unsigned int *)take((nblk + 2) * 4);
    s.cmp_sums = (unsigned int *)take((nblk / (FB_SCAN_BLOCK * FB_SCAN_PER_THREAD) + 4) * 4);
    s.cmp_offsets = (long long *)take(2 * 8);
    s.bytes = off;
    return FB_OK;
}

// the slab calls run the q kernels on interleaved nodes when the whole-grid path would (the in-place y sweep needs two passes)
bool slab_interleaved(const fb_problem *pr, const Derived &d) { return use_sweepq(pr, d) && pr->num_iter >= 2; }

int slab_check(const fb_problem *pr, Derived &d, long long z_begin, long long z_count, long long halo_lo, long long halo_hi)
{
    int rc = derive(pr, d);
    if (rc != FB_OK) return rc;
    if (pr->dim != 3 || pr->nfields != 1) return fail(FB_EINVAL, "z-slab runs need dim == 3 and nfields == 1");
    if (pr->flags & FB_FLAG_FP32) return fail(FB_EINVAL, "z-slab runs are fp64 only");
    if (z_begin < 0 || z_count < 1 || z_begin + z_count > d.Dz || halo_lo < 0 || halo_hi < 0 ||
        z_begin - halo_lo < 0 || z_begin + z_count + halo_hi > d.Dz)
        return fail(FB_EINVAL, "invalid slab: planes [%lld, %lld) halo %lld / %lld of %lld", z_begin, z_begin + z_count,
                    halo_lo, halo_hi, d.Dz);
    return FB_OK;
}
}  // namespace

FB_EXPORT int64_t fb_slab_halo_planes(const fb_problem *prob)
{
    Derived d;
    if (derive(prob, d) != FB_OK) return FB_EINVAL;
    if (prob->dim != 3) return fail(FB_EINVAL, "z-slab runs need dim == 3");
    return (int64_t)prob->num_iter * (d.ax[2].T + 1);
}

FB_EXPORT int fb_slab_result_offsets(const fb_problem *prob, int64_t nsamples, int64_t z_count, int64_t halo_lo, int64_t halo_hi,
                                     int want_out64, int64_t *offset_out32, int64_t *offset_out64)
{
    Derived d;
    int rc = derive(prob, d);
    if (rc != FB_OK) return rc;
    if (prob->dim != 3 || prob->nfields != 1) return fail(FB_EINVAL, "z-slab runs need dim == 3 and nfields == 1");
    SlabLayout s;
    slab_carve(s, nullptr, prob, d, nsamples, z_count + halo_lo + halo_hi, want_out64 != 0);
    if (offset_out32) *offset_out32 = (int64_t)s.off_out;
    if (offset_out64) *offset_out64 = want_out64 ? (int64_t)s.off_out64 : -1;
    return FB_OK;
}

FB_EXPORT int fb_slab_interleaved(const fb_problem *prob)
{
    Derived d;
    if (derive(prob, d) != FB_OK) return FB_EINVAL;
    if (prob->dim != 3) return fail(FB_EINVAL, "z-slab runs need dim == 3");
    return slab_interleaved(prob, d) ? 1 : 0;
}

FB_EXPORT int fb_slab_layout(const fb_problem *prob, int64_t nsamples, int64_t z_count, int64_t halo_lo, int64_t halo_hi,
                             int want_out64, int64_t *workspace_bytes, int64_t *offset_vB, int64_t *offset_wB)
{
    Derived d;
    int rc = derive(prob, d);
    if (rc != FB_OK) return rc;
    if (prob->dim != 3 || prob->nfields != 1) return fail(FB_EINVAL, "z-slab runs need dim == 3 and nfields == 1");
    SlabLayout s;
    slab_carve(s, nullptr, prob, d, nsamples, z_count + halo_lo + halo_hi, want_out64 != 0);
    if (workspace_bytes) *workspace_bytes = (int64_t)s.bytes;
    if (offset_vB) *offset_vB = (int64_t)s.off_vB;
    if (offset_wB) *offset_wB = (int64_t)s.off_wB;
    return FB_OK;
}

namespace {
// injection (do_inject) and / or the x and y sweeps of the own planes [pb, pb + pc) of a slab
// the samples whose cell touches planes [z_begin, z_begin + z_count), in order (slabs that are the whole volume: nothing to do)
int slab_compact(const fb_problem *prob, const Derived &d, const SlabLayout &s, long long nsamples, const double *d_pts,
                 const double *d_val, long long z_begin, long long z_count, cudaStream_t st, FbSamples &out, bool &used)
{
    used = false;
    if (z_count >= d.Dz || nsamples < 1024 || nsamples >= (1LL << 31)) return FB_OK;
    FbGrid gr{};
    gr.dim = prob->dim;
    gr.W = d.W; gr.H = d.H; gr.Dz = d.Dz; gr.total = d.W * d.H * z_count;
    gr.z_off = z_begin; gr.z_cnt = z_count;
    for (int m = 0; m < 3; ++m) { gr.x0[m] = prob->x0[m]; gr.step[m] = prob->step[m]; }
    const long long nblk = (nsamples + 255) / 256;
    fb_slab_compact_count_kernel<<<(unsigned)nblk, 256, 0, st>>>(d_pts, nsamples, gr, s.cmp_cnt);
    LAUNCH_CHECK();
    const int nsb = (int)((nblk + FB_SCAN_BLOCK * FB_SCAN_PER_THREAD - 1) / (FB_SCAN_BLOCK * FB_SCAN_PER_THREAD));
    fb_scan_local_kernel<<<(unsigned)nsb, FB_SCAN_BLOCK, 0, st>>>(s.cmp_cnt, s.cmp_start, s.cmp_sums, nblk);
    LAUNCH_CHECK();
    fb_scan_sums_kernel<<<1, 1024, 0, st>>>(s.cmp_sums, nsb, s.cmp_sums + nsb + 1);
    LAUNCH_CHECK();
    fb_scan_add_kernel<<<(unsigned)nsb, FB_SCAN_BLOCK, 0, st>>>(s.cmp_start, s.cmp_sums, s.cmp_sums + nsb + 1, nblk);
    LAUNCH_CHECK();
    fb_slab_compact_scatter_kernel<<<(unsigned)nblk, 256, 0, st>>>(d_pts, d_val, nsamples, gr, s.cmp_start, nblk, s.pts_c, s.val_c,
                                                                  s.cmp_offsets);
    LAUNCH_CHECK();
    out.pts = s.pts_c;
    out.val = s.val_c;
    out.offsets = s.cmp_offsets;
    out.n_uniform = 0;
    used = true;
    return FB_OK;
}

int slab_phase1(const fb_problem *prob, int64_t z_begin, int64_t z_count, int64_t halo_lo, int64_t halo_hi, int64_t nsamples,
                const double *d_pts, const double *d_val, int want_out64, void *d_workspace, int64_t workspace_bytes,
                void *stream, bool do_inject, long long pb, long long pc)
{
    int rc = require_device();
    if (rc != FB_OK) return rc;
    Derived d;
    if ((rc = slab_check(prob, d, z_begin, z_count, halo_lo, halo_hi)) != FB_OK) return rc;
    if ((rc = check_kernel_vs_grid(prob, d)) != FB_OK) return rc;
    if (!d_workspace || (do_inject && (!d_pts || !d_val))) return fail(FB_EINVAL, "null device pointer");
    if (pb < 0 || pc < 0 || pb + pc > z_count) return fail(FB_EINVAL, "plane range [%lld, %lld) outside the slab", pb, pb + pc);
    const long long z_ext = z_count + halo_lo + halo_hi;
    SlabLayout s;
    slab_carve(s, (char *)d_workspace, prob, d, nsamples, z_ext, want_out64 != 0);
    if ((long long)s.bytes > workspace_bytes) return fail(FB_ENOMEM, "workspace too small: need %zu bytes", s.bytes);
    cudaStream_t st = (cudaStream_t)stream;
    const long long plane = d.W * d.H;
    // injection into the own planes, held in the middle of the extended A buffers
    Workspace w = s.inj;
    FbSamples near{};
    bool compacted = false;
    if (do_inject && (rc = slab_compact(prob, d, s, nsamples, d_pts, d_val, z_begin, z_count, st, near, compacted)) != FB_OK) return rc;
    const FbSamples *from = compacted ? &near : nullptr;
    if (slab_interleaved(prob, d)) {
        // interleaved nodes: A2 / B2 are the 2 g byte blocks; the own planes' nodes are one contiguous run
        double *a2 = s.vA + 2 * halo_lo * plane, *b2 = s.vB + 2 * halo_lo * plane;
        w.vA = a2;
        w.wA = a2 + z_count * plane;                  // second half of the own planes' block (run_inject zero-fills both halves)
        w.vB = b2;
        w.wB = b2 + z_count * plane;
        if (do_inject && (rc = run_inject(prob, d, nsamples, nullptr, d_pts, d_val, w, st, z_begin, z_count, false, true, from)) != FB_OK)
            return rc;
        if (pc == 0) return FB_OK;
        SweepCounters qctr{w.counters + 4, 0};
        // x sweep A2 -> B2 (transposing), y sweep in place on B2, planes [pb, pb + pc)
        rc = run_sweepq(1, prob->num_iter, d.ax[0], a2 + 2 * pb * plane, b2 + 2 * pb * plane, nullptr, nullptr, w.mm, d.csf, pc, d.W,
                        d.H, st, qctr);
        if (rc != FB_OK) return rc;
        return run_sweepq(0, prob->num_iter, d.ax[1], b2 + 2 * pb * plane, b2 + 2 * pb * plane, nullptr, nullptr, w.mm, d.csf, pc, d.H,
                          d.W, st, qctr);
    }
    w.vA = s.vA + halo_lo * plane;
    w.wA = s.wA + halo_lo * plane;
    w.vB = s.vB + halo_lo * plane;
    w.wB = s.wB + halo_lo * plane;
    if (do_inject && (rc = run_inject(prob, d, nsamples, nullptr, d_pts, d_val, w, st, z_begin, z_count, false, false, from)) != FB_OK)
        return rc;
    if (pc == 0) return FB_OK;
    // x sweep (A -> B, transposing) and y sweep (in place) on the planes [pb, pb + pc)
    Pair cur{w.vA + pb * plane, w.wA + pb * plane}, spare{w.vB + pb * plane, w.wB + pb * plane};
    double *home_v = spare.v, *home_w = spare.w;
    SweepCounters ctr{w.counters + 4, 0};
    rc = run_sweep(1, prob->num_iter, d.ax[0], cur, spare, nullptr, nullptr, w.mm, d.csf, pc, d.W, d.H, true, st, ctr);
    if (rc != FB_OK) return rc;
    rc = run_sweep(0, prob->num_iter, d.ax[1], cur, spare, nullptr, nullptr, w.mm, d.csf, pc, d.H, d.W, true, st, ctr);
    if (rc != FB_OK) return rc;
    if (cur.v != home_v) {   // per-pass ping-pong may end in the other pair: bring the result home
        CUDA_TRY(cudaMemcpyAsync(home_v, cur.v, (size_t)plane * pc * 8, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(home_w, cur.w, (size_t)plane * pc * 8, cudaMemcpyDeviceToDevice, st));
    }
    return FB_OK;
}
}  // namespace

FB_EXPORT int fb_slab_phase1_dev(const fb_problem *prob, int64_t z_begin, int64_t z_count, int64_t halo_lo, int64_t halo_hi,
                                 int64_t nsamples, const double *d_pts, const double *d_val, int want_out64,
                                 void *d_workspace, int64_t workspace_bytes, void *stream)
{
    return slab_phase1(prob, z_begin, z_count, halo_lo, halo_hi, nsamples, d_pts, d_val, want_out64, d_workspace, workspace_bytes,
                       stream, true, 0, z_count);
}

FB_EXPORT int fb_slab_inject_dev(const fb_problem *prob, int64_t z_begin, int64_t z_count, int64_t halo_lo, int64_t halo_hi,
                                 int64_t nsamples, const double *d_pts, const double *d_val, int want_out64,
                                 void *d_workspace, int64_t workspace_bytes, void *stream)
{
    return slab_phase1(prob, z_begin, z_count, halo_lo, halo_hi, nsamples, d_pts, d_val, want_out64, d_workspace, workspace_bytes,
                       stream, true, 0, 0);
}

FB_EXPORT int fb_slab_sweeps_dev(const fb_problem *prob, int64_t z_begin, int64_t z_count, int64_t halo_lo, int64_t halo_hi,
                                 int64_t nsamples, int want_out64, int64_t plane_begin, int64_t plane_count,
                                 void *d_workspace, int64_t workspace_bytes, void *stream)
{
    return slab_phase1(prob, z_begin, z_count, halo_lo, halo_hi, nsamples, nullptr, nullptr, want_out64, d_workspace,
                       workspace_bytes, stream, false, plane_begin, plane_count);
}

namespace {
int slab_phase2(const fb_problem *prob, int64_t z_begin, int64_t z_count, int64_t halo_lo, int64_t halo_hi, int64_t nsamples,
                bool want64, float *d_out, double *d_out64, void *d_workspace, int64_t workspace_bytes, void *stream)
{
    int rc = require_device();
    if (rc != FB_OK) return rc;
    Derived d;
    if ((rc = slab_check(prob, d, z_begin, z_count, halo_lo, halo_hi)) != FB_OK) return rc;
    if (!d_workspace) return fail(FB_EINVAL, "null device pointer");
    const long long z_ext = z_count + halo_lo + halo_hi;
    SlabLayout s;
    slab_carve(s, (char *)d_workspace, prob, d, nsamples, z_ext, want64);
    if ((long long)s.bytes > workspace_bytes) return fail(FB_ENOMEM, "workspace too small: need %zu bytes", s.bytes);
    cudaStream_t st = (cudaStream_t)stream;
    const long long plane = d.W * d.H;
    // fused z sweep + mask + divide + cast over the extended lines (A is free again: spare)
    SweepCounters ctr{s.inj.counters + 4, 6};
    if (slab_interleaved(prob, d)) {
        rc = run_sweepq(2, prob->num_iter, d.ax[2], s.vB, nullptr, s.out_ext, s.out64_ext, s.inj.mm, d.csf, 1, z_ext, plane, st, ctr);
    } else {
        Pair cur{s.vB, s.wB}, spare{s.vA, s.wA};
        rc = run_sweep(2, prob->num_iter, d.ax[2], cur, spare, s.out_ext, s.out64_ext, s.inj.mm, d.csf, 1, z_ext, plane, true, st, ctr);
    }
    if (rc != FB_OK) return rc;
    if (!d_out) return FB_OK;                        // the caller reads the own planes of the extended result in place
    CUDA_TRY(cudaMemcpyAsync(d_out, s.out_ext + halo_lo * plane, (size_t)plane * z_count * 4, cudaMemcpyDeviceToDevice, st));
    if (d_out64)
        CUDA_TRY(cudaMemcpyAsync(d_out64, s.out64_ext + halo_lo * plane, (size_t)plane * z_count * 8, cudaMemcpyDeviceToDevice, st));
    return FB_OK;
}
}  // namespace

FB_EXPORT int fb_slab_phase2_dev(const fb_problem *prob, int64_t z_begin, int64_t z_count, int64_t halo_lo, int64_t halo_hi,
                                 int64_t nsamples, float *d_out, double *d_out64, void *d_workspace,
                                 int64_t workspace_bytes, void *stream)
{
    if (!d_out) return fail(FB_EINVAL, "null device pointer");
    return slab_phase2(prob, z_begin, z_count, halo_lo, halo_hi, nsamples, d_out64 != nullptr, d_out, d_out64, d_workspace,
                       workspace_bytes, stream);
}

FB_EXPORT int fb_slab_phase2_inplace_dev(const fb_problem *prob, int64_t z_begin, int64_t z_count, int64_t halo_lo, int64_t halo_hi,
                                         int64_t nsamples, int want_out64, void *d_workspace, int64_t workspace_bytes, void *stream)
{
    return slab_phase2(prob, z_begin, z_count, halo_lo, halo_hi, nsamples, want_out64 != 0, nullptr, nullptr, d_workspace,
                       workspace_bytes, stream);
}

// ---- interprocess events (ordering of the peer-mapped halo exchange of the z-slab runs) -------------------------------
// One process per GPU: a rank records an event behind the sweeps of the planes its neighbours need; the neighbours, which
// have opened the event's IPC handle, let their copy stream wait for it before they pull those planes.
FB_EXPORT int fb_ipc_event_create(void **event, void *handle64)
{
    int rc = require_device();
    if (rc != FB_OK) return rc;
    if (!event || !handle64) return fail(FB_EINVAL, "null pointer");
    cudaEvent_t ev;
    CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | cudaEventInterprocess));
    cudaIpcEventHandle_t h;
    cudaError_t e = cudaIpcGetEventHandle(&h, ev);
    if (e != cudaSuccess) {
        cudaEventDestroy(ev);
        return fail(FB_ECUDA, "cudaIpcGetEventHandle failed: %s", cudaGetErrorString(e));
    }
    static_assert(sizeof(cudaIpcEventHandle_t) == 64, "IPC event handle size");
    memcpy(handle64, &h, 64);
    *event = (void *)ev;
    return FB_OK;
}

FB_EXPORT int fb_ipc_event_open(const void *handle64, void **event)
{
    int rc = require_device();
    if (rc != FB_OK) return rc;
    if (!event || !handle64) return fail(FB_EINVAL, "null pointer");
    cudaIpcEventHandle_t h;
    memcpy(&h, handle64, 64);
    cudaEvent_t ev;
    CUDA_TRY(cudaIpcOpenEventHandle(&ev, h));
    *event = (void *)ev;
    return FB_OK;
}

FB_EXPORT int fb_event_record(void *event, void *stream)
{
    if (!event) return fail(FB_EINVAL, "null event");
    CUDA_TRY(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
    return FB_OK;
}

FB_EXPORT int fb_stream_wait_event(void *stream, void *event)
{
    if (!event) return fail(FB_EINVAL, "null event");
    CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0));
    return FB_OK;
}

FB_EXPORT int fb_event_destroy(void *event)
{
    if (event) CUDA_TRY(cudaEventDestroy((cudaEvent_t)event));
    return FB_OK;
}

// ---- S2 -------------------------------------------------------------------------------------------------
// util/lambert_conformal.py:50-94 (host libm, like the Numba code)
FB_EXPORT int fb_lambert_create_proj(double center_lon, double center_lat, double lat1, double lat2, double *proj)
{
    if (!proj) return fail(FB_EINVAL, "null pointer");
    const double RAD = M_PI / 180.0, HALF = RAD / 2.0;
    double n;
    if (lat1 != lat2)
        n = std::log(std::cos(lat1 * RAD) / std::cos(lat2 * RAD)) /
            std::log(std::tan((90.0 + lat2) * HALF) / std::tan((90.0 + lat1) * HALF));
    else
        n = std::sin(lat1 * RAD);
    const double n_inv = 1.0 / n;
    const double F = std::cos(lat1 * RAD) * std::pow(std::tan((90.0 + lat1) * HALF), n) / n;
    const double rho0 = F / std::pow(std::tan((90.0 + center_lat) * HALF), n);
    proj[0] = center_lon; proj[1] = n; proj[2] = n_inv; proj[3] = F; proj[4] = rho0;
    return FB_OK;
}

namespace {
FbProj make_proj(const double *proj) { return FbProj{proj[0], proj[1], proj[2], proj[3], proj[4]}; }

int resample_dev(const float *d_lam, long long lamW, long long lamH, const double *lam_x0, const double *x0,
                 const double *step, const int64_t *size, const double *proj, double *d_tab, float *d_res, cudaStream_t st)
{
    const long long W = size[0], H = size[1];
    if (H > 65535) return fail(FB_EINVAL, "output grid too tall for the resampling launch: %lld", H);
    FbProj pr = make_proj(proj);
    fb_resample_tables_kernel<<<(unsigned)((W + H + 255) / 256), 256, 0, st>>>(d_tab, W, H, x0[0], x0[1], step[0], step[1], pr);
    LAUNCH_CHECK();
    fb_resample_kernel<<<dim3((unsigned)((W + 255) / 256), (unsigned)H), 256, 0, st>>>(
        d_lam, lamW, lamH, d_tab, d_res, W, H, lam_x0[0], lam_x0[1], step[0], step[1], pr);
    LAUNCH_CHECK();
    return FB_OK;
}

void s2_problem(fb_problem &p, const double *sigma, const double *step, int num_iter, double mdw, const fb_s2_map &map)
{
    memset(&p, 0, sizeof p);
    p.dim = 2;
    p.method = FB_METHOD_OPTIMIZED_CONVOLUTION;
    p.num_iter = num_iter;
    p.nfields = 1;
    // interpolationS2.py:187-188: the grid in Lambert space, lam_size = int(extent / step)
    p.size[0] = (int64_t)(map.lam_extent[0] / step[0]);
    p.size[1] = (int64_t)(map.lam_extent[1] / step[1]);
    p.size[2] = 1;
    p.x0[0] = map.lam_x0[0];
    p.x0[1] = map.lam_x0[1];
    for (int m = 0; m < 2; ++m) { p.sigma[m] = sigma[m]; p.step[m] = step[m]; }
    p.max_dist_weight = mdw;
}

int check_map(const fb_s2_map *map)
{
    if (!map) return fail(FB_EINVAL, "null map");
    if (!(map->lam_extent[0] > 0.0) || !(map->lam_extent[1] > 0.0)) return fail(FB_EINVAL, "map extent must be positive");
    if (!(map->proj[1] != 0.0) || !std::isfinite(map->proj[3]) || !std::isfinite(map->proj[4]))
        return fail(FB_EINVAL, "invalid Lambert projection constants");
    return FB_OK;
}

// part1 on the device: d_lam receives the Lambert field; returns the staging layout it used
int s2_part1_dev(const fb_problem &lp, long long nsamples, const double *d_pts, double *d_lam_pts, const double *d_val,
                 const double *proj, float *d_lam, void *ws, long long ws_bytes, cudaStream_t st)
{
    if (nsamples > 0) {
        fb_lambert_to_map_kernel<<<(unsigned)((nsamples + 255) / 256), 256, 0, st>>>(d_pts, d_lam_pts, nsamples, make_proj(proj));
        LAUNCH_CHECK();
    }
    // the S2 path has no kernel-size-vs-grid check (interpolationS2.py:131-132)
    return pipeline(&lp, nsamples, nullptr, d_lam_pts, d_val, d_lam, nullptr, ws, ws_bytes, st, false);
}
}  // namespace

FB_EXPORT int fb_lambert_to_map_host(const double *geoc, double *mapc, int64_t n, const double *proj)
{
    int rc = require_device();
    if (rc != FB_OK) return rc;
    if (!geoc || !mapc || !proj || n < 0) return fail(FB_EINVAL, "invalid argument");
    if (n == 0) return FB_OK;
    std::lock_guard<std::mutex> lock(g_arena_mutex);
    void *stg = nullptr;
    if ((rc = arena_get(1, 2 * align_up((size_t)n * 16) + 1024, &stg)) != FB_OK) return rc;
    Staging s((char *)stg);
    double *d_in = s.take<double>((size_t)n * 2), *d_out = s.take<double>((size_t)n * 2);
    cudaStream_t st = 0;
    CUDA_TRY(cudaMemcpyAsync(d_in, geoc, (size_t)n * 16, cudaMemcpyHostToDevice, st));
    fb_lambert_to_map_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_in, d_out, n, make_proj(proj));
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpyAsync(mapc, d_out, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return FB_OK;
}

// interpolationS2.py:187-188, :208: the reference's fixed European map
FB_EXPORT int fb_s2_default_map(fb_s2_map *map)
{
    if (!map) return fail(FB_EINVAL, "null pointer");
    map->lam_x0[0] = -32.0;
    map->lam_x0[1] = -2.0;
    map->lam_extent[0] = 64.0;
    map->lam_extent[1] = 44.0;
    return fb_lambert_create_proj(11.5, 34.5, 42.5, 65.5, map->proj);
}

FB_EXPORT int fb_s2_part1_map_host(int64_t nsamples, const double *pts, const double *val, const double *sigma,
                                   const double *step, int num_iter, double max_dist_weight, const fb_s2_map *map,
                                   float *lam_field)
{
    int rc = require_device();
    if (rc != FB_OK) return rc;
    if (!pts || !val || !sigma || !step || !lam_field) return fail(FB_EINVAL, "null pointer");
    if ((rc = check_map(map)) != FB_OK) return rc;
    const double *proj = map->proj;
    fb_problem lp;
    s2_problem(lp, sigma, step, num_iter, max_dist_weight, *map);
    Derived d;
    if ((rc = derive(&lp, d)) != FB_OK) return rc;
    std::lock_guard<std::mutex> lock(g_arena_mutex);
    Workspace w;
    carve(w, nullptr, &lp, d.total, nsamples);
    const size_t stage_bytes = 2 * align_up((size_t)nsamples * 16) + align_up((size_t)nsamples * 8) + align_up((size_t)d.total * 4) + 1024;
    void *ws = nullptr, *stg = nullptr;
    if ((rc = arena_get(0, w.bytes, &ws)) != FB_OK) return rc;
    if ((rc = arena_get(1, stage_bytes, &stg)) != FB_OK) return rc;
    Staging s((char *)stg);
    double *d_pts = s.take<double>((size_t)nsamples * 2), *d_lpts = s.take<double>((size_t)nsamples * 2);
    double *d_val = s.take<double>((size_t)nsamples);
    float *d_lam = s.take<float>((size_t)d.total);
    cudaStream_t st = 0;
    CUDA_TRY(cudaMemcpyAsync(d_pts, pts, (size_t)nsamples * 16, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_val, val, (size_t)nsamples * 8, cudaMemcpyHostToDevice, st));
    if ((rc = s2_part1_dev(lp, nsamples, d_pts, d_lpts, d_val, proj, d_lam, ws, (long long)g_arena[0].bytes, st)) != FB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(lam_field, d_lam, (size_t)d.total * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return FB_OK;
}

FB_EXPORT int fb_s2_part1_host(int64_t nsamples, const double *pts, const double *val, const double *sigma,
                               const double *step, int num_iter, double max_dist_weight, const double *proj,
                               float *lam_field)
{
    if (!proj) return fail(FB_EINVAL, "null pointer");
    fb_s2_map map;
    int rc = fb_s2_default_map(&map);
    if (rc != FB_OK) return rc;
    memcpy(map.proj, proj, sizeof map.proj);
    return fb_s2_part1_map_host(nsamples, pts, val, sigma, step, num_iter, max_dist_weight, &map, lam_field);
}

FB_EXPORT int fb_s2_resample_host(const float *lam_field, int64_t lam_w, int64_t lam_h, const double *lam_x0,
                                  const double *x0, const double *step, const int64_t *size, const double *proj,
                                  float *res)
{
    int rc = require_device();
    if (rc != FB_OK) return rc;
    if (!lam_field || !lam_x0 || !x0 || !step || !size || !proj || !res) return fail(FB_EINVAL, "null pointer");
    std::lock_guard<std::mutex> lock(g_arena_mutex);
    const size_t nl = (size_t)lam_w * lam_h, no = (size_t)size[0] * size[1];
    void *stg = nullptr;
    if ((rc = arena_get(1, align_up(nl * 4) + align_up(no * 4) + align_up((size_t)(2 * size[0] + size[1]) * 8) + 1024, &stg)) != FB_OK) return rc;
    Staging s((char *)stg);
    float *d_lam = s.take<float>(nl), *d_res = s.take<float>(no);
    double *d_tab = s.take<double>((size_t)(2 * size[0] + size[1]));
    cudaStream_t st = 0;
    CUDA_TRY(cudaMemcpyAsync(d_lam, lam_field, nl * 4, cudaMemcpyHostToDevice, st));
    if ((rc = resample_dev(d_lam, lam_w, lam_h, lam_x0, x0, step, size, proj, d_tab, d_res, st)) != FB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(res, d_res, no * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return FB_OK;
}

FB_EXPORT int fb_barnes_s2_map_host(int64_t nsamples, const double *pts, const double *val, const double *sigma,
                                    const double *x0, const double *step, const int64_t *size, int num_iter,
                                    double max_dist_weight, const fb_s2_map *map, float *res)
{
    int rc = require_device();
    if (rc != FB_OK) return rc;
    if (!pts || !val || !sigma || !x0 || !step || !size || !res) return fail(FB_EINVAL, "null pointer");
    if ((rc = check_map(map)) != FB_OK) return rc;
    const double *proj = map->proj;
    fb_problem lp;
    s2_problem(lp, sigma, step, num_iter, max_dist_weight, *map);
    Derived d;
    if ((rc = derive(&lp, d)) != FB_OK) return rc;
    std::lock_guard<std::mutex> lock(g_arena_mutex);
    Workspace w;
    carve(w, nullptr, &lp, d.total, nsamples);
    const size_t no = (size_t)size[0] * size[1];
    const size_t stage_bytes = 2 * align_up((size_t)nsamples * 16) + align_up((size_t)nsamples * 8) + align_up((size_t)d.total * 4) +
                               align_up(no * 4) + align_up((size_t)(2 * size[0] + size[1]) * 8) + 1024;
    void *ws = nullptr, *stg = nullptr;
    if ((rc = arena_get(0, w.bytes, &ws)) != FB_OK) return rc;
    if ((rc = arena_get(1, stage_bytes, &stg)) != FB_OK) return rc;
    Staging s((char *)stg);
    double *d_pts = s.take<double>((size_t)nsamples * 2), *d_lpts = s.take<double>((size_t)nsamples * 2);
    double *d_val = s.take<double>((size_t)nsamples);
    float *d_lam = s.take<float>((size_t)d.total), *d_res = s.take<float>(no);
    double *d_tab = s.take<double>((size_t)(2 * size[0] + size[1]));
    cudaStream_t st = 0;
    CUDA_TRY(cudaMemcpyAsync(d_pts, pts, (size_t)nsamples * 16, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_val, val, (size_t)nsamples * 8, cudaMemcpyHostToDevice, st));
    if ((rc = s2_part1_dev(lp, nsamples, d_pts, d_lpts, d_val, proj, d_lam, ws, (long long)g_arena[0].bytes, st)) != FB_OK) return rc;
    const double lam_x0[2] = {lp.x0[0], lp.x0[1]};
    if ((rc = resample_dev(d_lam, lp.size[0], lp.size[1], lam_x0, x0, step, size, proj, d_tab, d_res, st)) != FB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(res, d_res, no * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return FB_OK;
}

FB_EXPORT int fb_barnes_s2_host(int64_t nsamples, const double *pts, const double *val, const double *sigma,
                                const double *x0, const double *step, const int64_t *size, int num_iter,
                                double max_dist_weight, const double *proj, float *res)
{
    if (!proj) return fail(FB_EINVAL, "null pointer");
    fb_s2_map map;
    int rc = fb_s2_default_map(&map);
    if (rc != FB_OK) return rc;
    memcpy(map.proj, proj, sizeof map.proj);
    return fb_barnes_s2_map_host(nsamples, pts, val, sigma, x0, step, size, num_iter, max_dist_weight, &map, res);
}

// ---- introspection ------------------------------------------------------------------------------------------
FB_EXPORT int64_t fb_kernel_launch_count(void) { return g_launches.load(); }

FB_EXPORT int fb_set_option(const char *name, int value)
{
    if (!name) return fail(FB_EINVAL, "null option name");
    if (!strcmp(name, "inject_lists")) { g_inject_lists.store(value); return FB_OK; }
    if (!strcmp(name, "host_chunk_fields")) { g_host_chunk_fields.store(value); return FB_OK; }
    if (!strcmp(name, "sweepq")) { g_sweepq.store(value); return FB_OK; }
    if (!strcmp(name, "sweepq_stages")) { g_q_nst.store(value); return FB_OK; }
    if (!strcmp(name, "sweepq_deep_staging")) { g_q_deep.store(value); return FB_OK; }
    if (!strcmp(name, "sweepq_prefetch")) { g_q_pf.store(value); return FB_OK; }
    if (!strcmp(name, "sweepq_warps")) { g_q_warps.store(value); return FB_OK; }
    if (!strcmp(name, "sweepq_cap_warps_xy")) { g_q_cap_xy.store(value); return FB_OK; }
    if (!strcmp(name, "sweepq_cap_warps_final")) { g_q_cap_fin.store(value); return FB_OK; }
    if (!strcmp(name, "sparse_inject")) { g_sparse.store(value); return FB_OK; }
    if (!strcmp(name, "sweepp")) { g_sweepp.store(value); return FB_OK; }
    if (!strcmp(name, "sweepq_reserve_sms")) { g_q_reserve.store(value < 0 ? 0 : value); return FB_OK; }
    if (!strcmp(name, "line1d")) { g_line1d.store(value); return FB_OK; }
    return fail(FB_EINVAL, "unknown option: %s", name);
}

FB_EXPORT int fb_set_profiling(int enabled)
{
    g_profiling.store(enabled ? 1 : 0);
    return FB_OK;
}

FB_EXPORT int fb_last_profile(double *ms_segments, int nsegments, int64_t *launches)
{
    if (!g_prof.armed) return fail(FB_EINVAL, "no profiled call recorded on this thread (fb_set_profiling(1) first)");
    if (g_prof.marked < 2) return fail(FB_EINVAL, "profile incomplete");
    CUDA_TRY(cudaEventSynchronize(g_prof.ev[g_prof.marked - 1]));
    for (int i = 0; i < nsegments; ++i) {
        float ms = 0.f;
        if (i < kProfSegments && i + 1 < g_prof.marked) CUDA_TRY(cudaEventElapsedTime(&ms, g_prof.ev[i], g_prof.ev[i + 1]));
        if (ms_segments) ms_segments[i] = ms;
    }
    if (launches) *launches = g_prof.launches_end - g_prof.launches_begin;
    return FB_OK;
}
