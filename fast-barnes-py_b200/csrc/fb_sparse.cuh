// fb_sparse.cuh -- injection as a sort by cell + segmented reduce, feeding the x sweep directly.
//
// Replaces _inject_data_{2,3}d (fastbarnes/interpolation.py:241-322) for the q path.  The dense path scatters the
// samples into zero-filled fp64 grids (16 B per grid point written by the zero-fill, read again by the x sweep, although
// 93 % of the bench grid stays zero).  Here the 2^dim (node, w, w * val) records of every sample are BINNED by the unit
// of work of the x sweep -- (outer = field x z plane, 16-line group along y, chunk of 8 rows along x) -- with a counting
// sort (count, exclusive scan, fill: integer atomics only), and one thread per bucket orders its records by (node,
// sample index) and adds the records of each node up from 0.0 in ascending sample order, exactly the order of the
// reference's sequential `vg[..] += w * val[k]` (interpolation.py:232-233, :257-270, :293-322).  The result is, per line
// group, a stream of (row, line, vg, wg) node entries sorted by row; the producer warps of fb_sweepqs_kernel turn it into
// the rows the passes read.  No dense injection grid, no zero-fill, no random read-modify-write traffic.
//
//   records (FbRec, 32 bytes = one DRAM sector each, bucket after bucket):
//     after the fill:   key = (row in chunk << 4 | line in group) << 25 | sample index in the field,  w,  w * (val - offset)
//     after the reduce: key = (row << 4 | line in group) for the first entry of a node (w, wv = the node's sums),
//                       FB_BIN_HOLE for the slots the other records of the node occupied
#pragma once
#include "fb_kernels.cuh"

#define FB_BIN_HOLE 0xffffffffu
#define FB_BIN_MAX_SAMPLES (1 << 25)     // sample index bits of a record key

struct __align__(32) FbRec {
    unsigned int key, pad0;
    double w, wv;
    double pad1;
};

__device__ __forceinline__ void fb_rec_store(FbRec *r, unsigned int key, double w, double wv)
{
    // two 16-byte stores: one full sector, no read-modify-write
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(r), "r"(key), "r"(0u), "r"((unsigned)__double2loint(w)),
                 "r"((unsigned)__double2hiint(w)) : "memory");
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"((char *)r + 16), "r"((unsigned)__double2loint(wv)),
                 "r"((unsigned)__double2hiint(wv)), "r"(0u), "r"(0u) : "memory");
}

struct FbBins {
    int G;                  // 16-line groups along y
    int NB;                 // buckets (chunks of 8 rows of the x sweep's stream) per line group
    int t_begin;            // stream row of the first row of chunk 0 (-7 .. 0): rows of chunk m = t_begin + 8 m ..
    long long nbuckets;     // n_outer * G * NB   (n_outer = fields x z planes)
};

__device__ __forceinline__ long long fb_bucket_of(const FbBins &bn, long long outer, long long x, long long y)
{
    return (outer * bn.G + (y >> 4)) * bn.NB + ((x - bn.t_begin) >> 3);
}

// corner c of the cell: node coordinates and multilinear weight (corner and product order of interpolation.py:231-237,
// :256-270, :292-322 -- the same as fb_corner)
__device__ __forceinline__ void fb_corner_xyz(const FbGrid &g, int c, long long xi, long long yi, long long zi, double xw, double yw,
                                              double zw, long long &x, long long &y, long long &z, double &w)
{
    const int cx = ((c & 3) == 1 || (c & 3) == 2) ? 1 : 0;   // 0:(0,0) 1:(1,0) 2:(1,1) 3:(0,1)
    const int cy = ((c & 3) >= 2) ? 1 : 0;
    const int cz = c >> 2;
    const double wx = cx ? xw : __dsub_rn(1.0, xw);
    const double wy = cy ? yw : __dsub_rn(1.0, yw);
    w = __dmul_rn(wx, wy);
    if (g.dim == 3) {
        const double wz = cz ? zw : __dsub_rn(1.0, zw);
        w = __dmul_rn(w, wz);
    }
    x = xi + cx;
    y = yi + cy;
    z = zi + cz;
}

// K-a: records per bucket
__global__ void __launch_bounds__(256)
fb_bin_count_kernel(FbSamples s, FbGrid g, FbBins bn, unsigned int *cnt)
{
    const long long b = blockIdx.y;
    long long beg, n;
    fb_field_range(s, b, beg, n);
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    long long xi, yi, zi;
    double xw, yw, zw;
    if (!fb_sample_cell(g, s.pts, beg + k, xi, yi, zi, xw, yw, zw)) return;
    const int nc = 1 << g.dim;
    for (int c = 0; c < nc; ++c) {
        long long x, y, z;
        double w;
        fb_corner_xyz(g, c, xi, yi, zi, xw, yw, zw, x, y, z, w);
        atomicAdd(&cnt[fb_bucket_of(bn, b * g.Dz + z, x, y)], 1u);
    }
}

// K-b: exclusive scan of the bucket counts (three small launches: block-local scan, scan of the block sums, add).
// start[i] = sum of cnt[0 .. i-1]; start has n + 1 entries.
#define FB_SCAN_BLOCK 1024
#define FB_SCAN_PER_THREAD 4
__global__ void __launch_bounds__(FB_SCAN_BLOCK)
fb_scan_local_kernel(const unsigned int *cnt, unsigned int *start, unsigned int *block_sums, long long n)
{
    __shared__ unsigned int warp_tot[32];
    const long long base = ((long long)blockIdx.x * FB_SCAN_BLOCK + threadIdx.x) * FB_SCAN_PER_THREAD;
    unsigned int v[FB_SCAN_PER_THREAD], run = 0;
#pragma unroll
    for (int i = 0; i < FB_SCAN_PER_THREAD; ++i) {
        v[i] = (base + i < n) ? cnt[base + i] : 0u;
        run += v[i];
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned int incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        unsigned int t = warp_tot[lane], ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += up;
        }
        warp_tot[lane] = ti - t;                         // exclusive over the warps
        if (lane == 31) block_sums[blockIdx.x] = ti;
    }
    __syncthreads();
    unsigned int excl = warp_tot[wid] + incl - run;
#pragma unroll
    for (int i = 0; i < FB_SCAN_PER_THREAD; ++i) {
        if (base + i < n) start[base + i] = excl;
        excl += v[i];
    }
}

// one block: exclusive scan of the block sums in place; the grand total goes to total[0]
__global__ void __launch_bounds__(1024)
fb_scan_sums_kernel(unsigned int *block_sums, int nblocks, unsigned int *total)
{
    __shared__ unsigned int warp_tot[32];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + (int)threadIdx.x;
        const unsigned int v = i < nblocks ? block_sums[i] : 0u;
        unsigned int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            unsigned int t = warp_tot[lane], ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int up = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= o) ti += up;
            }
            warp_tot[lane] = ti - t;
        }
        __syncthreads();
        const unsigned int c0 = carry;
        if (i < nblocks) block_sums[i] = c0 + warp_tot[wid] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c0 + warp_tot[wid] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) total[0] = carry;
}

__global__ void __launch_bounds__(FB_SCAN_BLOCK)
fb_scan_add_kernel(unsigned int *start, const unsigned int *block_sums, const unsigned int *total, long long n)
{
    const long long base = ((long long)blockIdx.x * FB_SCAN_BLOCK + threadIdx.x) * FB_SCAN_PER_THREAD;
    const unsigned int add = block_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < FB_SCAN_PER_THREAD; ++i)
        if (base + i < n) start[base + i] += add;
    if (blockIdx.x == 0 && threadIdx.x == 0) start[n] = total[0];
}

// K-c: the records go to their buckets (slots are taken from the end of the bucket's range by counting `cnt` down to
// zero: the order inside a bucket is arbitrary, the reduce kernel sorts)
__global__ void __launch_bounds__(256)
fb_bin_fill_kernel(FbSamples s, FbGrid g, FbBins bn, const unsigned long long *mm, const unsigned int *start, unsigned int *cnt,
                   FbRec *rec)
{
    const long long b = blockIdx.y;
    long long beg, n;
    fb_field_range(s, b, beg, n);
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    long long xi, yi, zi;
    double xw, yw, zw;
    if (!fb_sample_cell(g, s.pts, beg + k, xi, yi, zi, xw, yw, zw)) return;
    const double valc = __dsub_rn(s.val[beg + k], fb_field_offset(mm, b));     // `val -= offset`, interpolation.py:210
    const int nc = 1 << g.dim;
    for (int c = 0; c < nc; ++c) {
        long long x, y, z;
        double w;
        fb_corner_xyz(g, c, xi, yi, zi, xw, yw, zw, x, y, z, w);
        const long long bk = fb_bucket_of(bn, b * g.Dz + z, x, y);
        const unsigned int slot = start[bk] + atomicSub(&cnt[bk], 1u) - 1u;
        const unsigned int pos7 = (unsigned int)(((x - bn.t_begin) & 7) << 4) | (unsigned int)(y & 15);
        fb_rec_store(rec + slot, (pos7 << 25) | (unsigned int)k, w, __dmul_rn(w, valc));
    }
}

// K-d: order the records of every bucket by (node, sample index), add the records of every node up from 0.0 in that
// order (`vg[..] += w * val[k]; wg[..] += w`), leave one entry per node at the front of its run and holes behind it.
// A warp takes 32 consecutive buckets: their records are one contiguous range, loaded into shared memory with
// coalesced 32-byte accesses; lane i then works on bucket i there (insertion sort: a bucket holds ~9 records in the
// bench workload; heap sort above 48), and the range is written back coalesced.  Ranges that do not fit
// (FB_RED_CAP records, clustered observations) are processed in global memory by the same code.
#define FB_RED_CAP 448
#define FB_RED_WARPS 4

template <typename KeyPtr, typename WPtr, typename VPtr>
__device__ __forceinline__ void fb_bin_sort_reduce(KeyPtr rk, WPtr rw, VPtr rv, unsigned int n, unsigned int row0)
{
    // fast path: every record of the bucket sits on its own node (the usual case away from clustered stations): the sums
    // are the records themselves (0.0 + x == x up to the sign of a zero) and no order matters -- only the keys change
    bool distinct = n <= 64;                             // (quadratic test: small buckets only)
    for (unsigned int i = 1; i < n && distinct; ++i) {
        const unsigned int pi = rk[i] >> 25;
        for (unsigned int j = 0; j < i; ++j) distinct = distinct && ((rk[j] >> 25) != pi);
    }
    if (distinct) {
        for (unsigned int i = 0; i < n; ++i) {
            const unsigned int pos7 = rk[i] >> 25;
            rk[i] = ((row0 + (pos7 >> 4)) << 4) | (pos7 & 15u);
            rw[i] = __dadd_rn(0.0, rw[i]);
            rv[i] = __dadd_rn(0.0, rv[i]);
        }
        return;
    }
    if (n > 1) {
        if (n <= 48) {
            for (unsigned int i = 1; i < n; ++i) {               // insertion sort
                const unsigned int kk = rk[i];
                const double ww = rw[i], vv = rv[i];
                unsigned int j = i;
                while (j > 0 && rk[j - 1] > kk) { rk[j] = rk[j - 1]; rw[j] = rw[j - 1]; rv[j] = rv[j - 1]; --j; }
                rk[j] = kk; rw[j] = ww; rv[j] = vv;
            }
        } else {                                                   // heap sort
            auto swp = [&](unsigned int i, unsigned int j) {
                const unsigned int tk = rk[i]; rk[i] = rk[j]; rk[j] = tk;
                const double tw = rw[i]; rw[i] = rw[j]; rw[j] = tw;
                const double tv = rv[i]; rv[i] = rv[j]; rv[j] = tv;
            };
            for (unsigned int st = n / 2; st-- > 0;) {
                unsigned int root = st;
                for (;;) {
                    unsigned int child = 2 * root + 1;
                    if (child >= n) break;
                    if (child + 1 < n && rk[child] < rk[child + 1]) ++child;
                    if (rk[root] >= rk[child]) break;
                    swp(root, child);
                    root = child;
                }
            }
            for (unsigned int end = n - 1; end > 0; --end) {
                swp(0, end);
                unsigned int root = 0;
                for (;;) {
                    unsigned int child = 2 * root + 1;
                    if (child >= end) break;
                    if (child + 1 < end && rk[child] < rk[child + 1]) ++child;
                    if (rk[root] >= rk[child]) break;
                    swp(root, child);
                    root = child;
                }
            }
        }
    }
    unsigned int i = 0;
    while (i < n) {
        const unsigned int pos7 = rk[i] >> 25;
        double sv = 0.0, sw = 0.0;
        unsigned int j = i;
        while (j < n && (rk[j] >> 25) == pos7) {
            sv = __dadd_rn(sv, rv[j]);
            sw = __dadd_rn(sw, rw[j]);
            ++j;
        }
        rk[i] = ((row0 + (pos7 >> 4)) << 4) | (pos7 & 15u);    // row0 may be "negative" (wrapped): rows < 0 hold no node
        rw[i] = sw;
        rv[i] = sv;
        for (unsigned int h = i + 1; h < j; ++h) rk[h] = FB_BIN_HOLE;
        i = j;
    }
}

// strided views of the key / weight / value members of an FbRec array (the global-memory fallback sorts in place)
struct FbRecKeys { FbRec *r; __device__ unsigned int &operator[](unsigned int i) const { return r[i].key; } };
struct FbRecW { FbRec *r; __device__ double &operator[](unsigned int i) const { return r[i].w; } };
struct FbRecV { FbRec *r; __device__ double &operator[](unsigned int i) const { return r[i].wv; } };

__global__ void __launch_bounds__(32 * FB_RED_WARPS)
fb_bin_reduce_kernel(FbBins bn, const unsigned int *start, FbRec *rec)
{
    __shared__ unsigned int s_key[FB_RED_WARPS][FB_RED_CAP];
    __shared__ double s_w[FB_RED_WARPS][FB_RED_CAP];
    __shared__ double s_v[FB_RED_WARPS][FB_RED_CAP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long nwarps = (long long)gridDim.x * FB_RED_WARPS;
    for (long long g0 = ((long long)blockIdx.x * FB_RED_WARPS + wid) * 32; g0 < bn.nbuckets; g0 += nwarps * 32) {
        const long long bk = g0 + lane;
        const bool have = bk < bn.nbuckets;
        const unsigned int s0 = have ? start[bk] : 0u, s1 = have ? start[bk + 1] : 0u;
        const unsigned int lo = __shfl_sync(0xffffffffu, s0, 0);
        const long long last = (g0 + 32 <= bn.nbuckets) ? 31 : (bn.nbuckets - 1 - g0);
        const unsigned int hi = __shfl_sync(0xffffffffu, s1, (int)last);
        const unsigned int total = hi - lo;
        const unsigned int n = s1 - s0;
        const unsigned int row0 = (unsigned int)(bn.t_begin + 8 * (int)(bk % bn.NB));
        if (total == 0) continue;
        if (total <= FB_RED_CAP) {
            for (unsigned int i = lane; i < total; i += 32) {
                unsigned int k0, k1, w0, w1, v0, v1, p0, p1;
                const FbRec *r = rec + lo + i;
                asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(k0), "=r"(k1), "=r"(w0), "=r"(w1) : "l"(r));
                asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(p0), "=r"(p1) : "l"((const char *)r + 16));
                s_key[wid][i] = k0;
                s_w[wid][i] = __hiloint2double((int)w1, (int)w0);
                s_v[wid][i] = __hiloint2double((int)v1, (int)v0);
            }
            __syncwarp();
            if (n > 0) fb_bin_sort_reduce(&s_key[wid][s0 - lo], &s_w[wid][s0 - lo], &s_v[wid][s0 - lo], n, row0);
            __syncwarp();
            for (unsigned int i = lane; i < total; i += 32) fb_rec_store(rec + lo + i, s_key[wid][i], s_w[wid][i], s_v[wid][i]);
            __syncwarp();
        } else if (n > 0) {
            fb_bin_sort_reduce(FbRecKeys{rec + s0}, FbRecW{rec + s0}, FbRecV{rec + s0}, n, row0);
        }
    }
}

// ------------------------------------------------------------------------------------------
// z-slab runs: every rank holds all samples, but only those whose cell touches the slab's planes leave records there.
// Injecting from the full list keeps one lane in eight busy (eight slabs); the samples are therefore compacted first --
// in order, so that the per-node sums keep the reference's sample order (a stable compaction: per-block counts, the
// exclusive scan above, scatter at block start + rank inside the block).  Min / max still run over ALL samples (the offset
// is the global one).
__device__ __forceinline__ bool fb_slab_keep(const FbGrid &g, const double *pts, long long gi)
{
    long long xi, yi, zi;
    double xw, yw, zw;
    if (!fb_sample_cell(g, pts, gi, xi, yi, zi, xw, yw, zw)) return false;      // outside the grid: no record anywhere
    return zi + 1 >= g.z_off && zi < g.z_off + g.z_cnt;                          // planes zi, zi + 1 against the window
}

__global__ void __launch_bounds__(256)
fb_slab_compact_count_kernel(const double *pts, long long n, FbGrid g, unsigned int *cnt)
{
    const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
    const int keep = (k < n && fb_slab_keep(g, pts, k)) ? 1 : 0;
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) cnt[blockIdx.x] = (unsigned int)c;
}

__global__ void __launch_bounds__(256)
fb_slab_compact_scatter_kernel(const double *pts, const double *val, long long n, FbGrid g, const unsigned int *start,
                               long long nblocks, double *pts_c, double *val_c, long long *offsets_c)
{
    __shared__ unsigned int warp_base[8];
    const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
    const bool keep = k < n && fb_slab_keep(g, pts, k);
    const unsigned int m = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_base[wid] = (unsigned int)__popc(m);
    __syncthreads();
    unsigned int base = start[blockIdx.x];
    for (int w = 0; w < wid; ++w) base += warp_base[w];
    if (keep) {
        const long long o = (long long)base + __popc(m & ((1u << lane) - 1u));
        for (int d = 0; d < g.dim; ++d) pts_c[o * g.dim + d] = pts[k * g.dim + d];
        val_c[o] = val[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        offsets_c[0] = 0;
        offsets_c[1] = (long long)start[nblocks];
    }
}
