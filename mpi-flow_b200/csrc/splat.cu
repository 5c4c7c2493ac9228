// splat.cu -- depth-aware forward splat, bit-exact with external/forward_warping/warping.c:6-33.
//
// The reference is a serial raster-order loop with a "last writer" z-LUT, so its result depends on
// the ORDER in which source pixels hit a target.  Closed form: for target T
// with writers k1 < ... < kn in raster order and z0 := 1000,
//     rgb(T)       = src[k_m],  m = max{ i : z_i < z_{i-1} }          (untouched if no such i)
//     valid(T)     = 1 iff n >= 1
//     collision(T) = 1 iff z_{n-1} == 1000   (n == 1, or the previous writer's z was exactly 1000)
// Parallel restatement: stable LSD radix sort (wide digits, two passes up to 4 M pixels) of the source pixels by
// target index (stability keeps
// raster order inside a segment) -> every writer sees its predecessor in the sorted array -> winners
// are marked independently, the last winner per target is an atomicMax over sorted positions.
// Everything is integer except the single fp32 '<', so the result is deterministic.
//
// Default path ("chains", round 2): real index maps send one, rarely a handful of source pixels to a target, so the sort
// is overkill.  splat_chain_build_kernel threads every source pixel onto a per-target list with one atomicExch
// (arbitrary order); splat_chain_resolve_kernel, one thread per target, pulls the list (up to CHAIN_MAX writers) into
// registers, orders it by raster index with an insertion sort and evaluates the closed form.  A target with more
// writers raises a flag, and the radix path below -- launched behind it, every kernel gated on that flag -- redoes the
// whole call (it writes the same bytes for the targets already resolved).  memset + 2 kernels + 1 gated launch.
//
// Radix path: the phases (keys + histogram, per pass a scan of the histograms and a stable scatter, resolve, colours)
// are device functions over "virtual blocks"; splat_fused_kernel is launched cooperatively with as many CTAs as are
// co-resident and walks them phase by phase with grid-wide barriers; when a cooperative launch is not possible (stream
// capture) every phase is its own launch.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace {

// LSD radix sort with wide digits: 2 passes of <= 11 bits sort up to 4 M pixels (3 passes beyond).  Per pass two
// launches: a per-digit scan of the per-block histograms, and the stable scatter, which also accumulates the
// histogram of the NEXT pass (by destination block) so that the keys are read once per pass.
constexpr int MAX_DIGIT_BITS = 11;
constexpr int MAX_RADIX = 1 << MAX_DIGIT_BITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ROUNDS = 8;                                   // items per thread
constexpr int SORT_TILE = SORT_THREADS * SORT_ROUNDS;            // items per block

// key = target raster index (y*w + x), narrowed through int exactly like warping.c:16-17;
// out-of-range targets (undefined behaviour in the reference) get the sentinel n and are dropped.
// Also: histogram of the first digit per block (digit-major, block-minor), winpos = 0, and zeroing of the
// histograms of the later passes (the scatter kernels add to them).
// vb / nvb: this virtual block and their number (= blockIdx.x / gridDim.x when every phase is its own launch)
__device__ __forceinline__ void
splat_keys_hist(uint32_t vb, uint32_t nvb, const int64_t *__restrict__ idx, const int64_t *__restrict__ idy, int h, int w, uint32_t n,
                int digit_bits, uint32_t nblocks, uint32_t *__restrict__ keys, uint32_t *__restrict__ winpos,
                uint32_t *__restrict__ hist0, uint32_t *__restrict__ hist_later, uint32_t hist_later_count)
{
    extern __shared__ uint32_t sh[];
    const uint32_t radix = 1u << digit_bits;
    for (uint32_t i = threadIdx.x; i < radix; i += SORT_THREADS) sh[i] = 0;
    for (uint32_t i = vb * SORT_THREADS + threadIdx.x; i < hist_later_count; i += nvb * SORT_THREADS)
        hist_later[i] = 0;
    __syncthreads();
    const uint32_t base = vb * SORT_TILE;
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const uint32_t k = base + r * SORT_THREADS + threadIdx.x;
        if (k < n) {
            const int x = (int)idx[k], y = (int)idy[k];
            const bool ok = (x >= 0) && (x < w) && (y >= 0) && (y < h);
            const uint32_t key = ok ? (uint32_t)y * (uint32_t)w + (uint32_t)x : n;
            keys[k] = key;
            winpos[k] = 0;
            atomicAdd(&sh[key & (radix - 1)], 1u);
        }
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < radix; d += SORT_THREADS) hist0[(size_t)d * nblocks + vb] = sh[d];
}

// one warp per digit: exclusive scan of that digit's per-block counts (in place) and the digit's total
__device__ __forceinline__ void
radix_offsets(uint32_t vb, uint32_t *__restrict__ hist, uint32_t nblocks, uint32_t radix, uint32_t *__restrict__ totals)
{
    const uint32_t d = vb * SORT_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (d >= radix) return;
    uint32_t *row = hist + (size_t)d * nblocks;
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < nblocks; b0 += 32) {
        const uint32_t b = b0 + lane;
        const uint32_t v = (b < nblocks) ? row[b] : 0;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (b < nblocks) row[b] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) totals[d] = carry;
}

// stable scatter: item order inside a tile is (warp, round, lane); ranks come from match_any.
// FIRST: values are the item indices themselves (nothing to read); hist_next != nullptr: count the next digit
// per destination block.
template <bool FIRST>
__device__ __forceinline__ void
radix_scatter(uint32_t vb, const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, uint32_t n,
              int shift, int digit_bits, uint32_t nblocks, const uint32_t *__restrict__ offsets,
              const uint32_t *__restrict__ totals, uint32_t *__restrict__ keys_out,
              uint32_t *__restrict__ vals_out, uint32_t *__restrict__ hist_next, int next_shift,
              int next_bits)
{
    extern __shared__ uint32_t sh[];
    const uint32_t radix = 1u << digit_bits, mask = radix - 1;
    uint32_t *cnt = sh;                        // [SORT_WARPS][radix]
    uint32_t *dbase = sh + SORT_WARPS * radix; // [radix] exclusive scan of the digit totals
    __shared__ uint32_t warp_sums[SORT_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < SORT_WARPS * radix; i += SORT_THREADS) cnt[i] = 0;
    {   // digit bases: block-wide exclusive scan of totals[0..radix), radix / 256 consecutive digits per thread
        const uint32_t per = radix / SORT_THREADS ? radix / SORT_THREADS : 1;
        const uint32_t d0 = threadIdx.x * per;
        uint32_t local = 0;
        for (uint32_t i = 0; i < per; i++)
            if (d0 + i < radix) local += totals[d0 + i];
        uint32_t inc = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        uint32_t before = inc - local;
        for (int wv = 0; wv < wid; wv++) before += warp_sums[wv];
        for (uint32_t i = 0; i < per; i++)
            if (d0 + i < radix) { dbase[d0 + i] = before; before += totals[d0 + i]; }
    }
    __syncthreads();
    const uint32_t wbase = vb * SORT_TILE + wid * (SORT_ROUNDS * 32);
    uint32_t key[SORT_ROUNDS], val[SORT_ROUNDS], peer[SORT_ROUNDS];   // peer: lanes of this round with the same digit
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool live = i < n;
        key[r] = live ? keys_in[i] : 0xffffffffu;
        val[r] = FIRST ? i : (live ? vals_in[i] : 0);
        peer[r] = 0;
        const unsigned act = __ballot_sync(0xffffffffu, live);
        if (live) {
            const uint32_t d = (key[r] >> shift) & mask;
            peer[r] = __match_any_sync(act, d);
            if (lane == __ffs(peer[r]) - 1) cnt[wid * radix + d] += __popc(peer[r]);
        }
        __syncwarp();
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < radix; d += SORT_THREADS) {
        uint32_t run = dbase[d] + offsets[(size_t)d * nblocks + vb];
#pragma unroll
        for (int wv = 0; wv < SORT_WARPS; wv++) {
            const uint32_t c = cnt[wv * radix + d];
            cnt[wv * radix + d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool live = i < n;
        const unsigned act = __ballot_sync(0xffffffffu, live);
        if (live) {
            const uint32_t d = (key[r] >> shift) & mask;
            const unsigned peers = peer[r];
            const uint32_t pos = cnt[wid * radix + d] + __popc(peers & ((1u << lane) - 1));
            keys_out[pos] = key[r];
            vals_out[pos] = val[r];
            if (hist_next)
                atomicAdd(&hist_next[(size_t)((key[r] >> next_shift) & ((1u << next_bits) - 1)) * nblocks + pos / SORT_TILE], 1u);
            __syncwarp(act);
            if (lane == __ffs(peers) - 1) cnt[wid * radix + d] += __popc(peers);
        }
        __syncwarp();
    }
}

// every sorted position i: compare with its raster-order predecessor in the same target
__device__ __forceinline__ void splat_resolve(uint32_t i, const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                              const float *__restrict__ z, uint32_t n, uint32_t *__restrict__ winpos,
                                              uint8_t *__restrict__ warped)
{
    if (i >= n) return;
    const uint32_t t = keys[i];
    if (t >= n) return;  // dropped (out-of-range target)
    const bool first = (i == 0) || (keys[i - 1] != t);
    const float zprev = first ? 1000.f : z[vals[i - 1]];  // dlut[t] when this writer arrives (:8-11, :28)
    if (z[vals[i]] < zprev) atomicMax(&winpos[t], i + 1);  // :19-21, the last such writer wins
    const bool last = (i + 1 == n) || (keys[i + 1] != t);
    if (last) {
        warped[(size_t)t * 5 + 3] = 1;                          // :23
        warped[(size_t)t * 5 + 4] = (zprev != 1000.f) ? 0 : 1;  // :24-27
    }
}

__device__ __forceinline__ void splat_rgb(uint32_t t, const uint32_t *__restrict__ vals, const uint32_t *__restrict__ winpos,
                                          const uint8_t *__restrict__ src, uint32_t n, uint8_t *__restrict__ warped)
{
    if (t >= n) return;
    const uint32_t wp = winpos[t];
    if (!wp) return;
    const uint32_t k = vals[wp - 1];
    warped[(size_t)t * 5 + 0] = src[(size_t)k * 3 + 0];
    warped[(size_t)t * 5 + 1] = src[(size_t)k * 3 + 1];
    warped[(size_t)t * 5 + 2] = src[(size_t)k * 3 + 2];
}

// ---- one launch per phase (MPIFLOW_SPLAT_FUSED=0, and whenever a cooperative launch is not possible) ----
__global__ void __launch_bounds__(SORT_THREADS)
splat_keys_hist_kernel(const int64_t *__restrict__ idx, const int64_t *__restrict__ idy, int h, int w, uint32_t n,
                       int digit_bits, uint32_t nblocks, uint32_t *__restrict__ keys, uint32_t *__restrict__ winpos,
                       uint32_t *__restrict__ hist0, uint32_t *__restrict__ hist_later, uint32_t hist_later_count,
                       const uint32_t *gate)
{
    if (gate && *gate == 0) return;      // the chain path resolved every target: nothing to redo
    splat_keys_hist(blockIdx.x, gridDim.x, idx, idy, h, w, n, digit_bits, nblocks, keys, winpos, hist0, hist_later, hist_later_count);
}
__global__ void __launch_bounds__(SORT_THREADS)
radix_offsets_kernel(uint32_t *__restrict__ hist, uint32_t nblocks, uint32_t radix, uint32_t *__restrict__ totals,
                     const uint32_t *gate)
{
    if (gate && *gate == 0) return;
    radix_offsets(blockIdx.x, hist, nblocks, radix, totals);
}
template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, uint32_t n,
                     int shift, int digit_bits, uint32_t nblocks, const uint32_t *__restrict__ offsets,
                     const uint32_t *__restrict__ totals, uint32_t *__restrict__ keys_out,
                     uint32_t *__restrict__ vals_out, uint32_t *__restrict__ hist_next, int next_shift,
                     int next_bits, const uint32_t *gate)
{
    if (gate && *gate == 0) return;
    radix_scatter<FIRST>(blockIdx.x, keys_in, vals_in, n, shift, digit_bits, nblocks, offsets, totals, keys_out, vals_out,
                         hist_next, next_shift, next_bits);
}
__global__ void splat_resolve_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                     const float *__restrict__ z, uint32_t n, uint32_t *__restrict__ winpos,
                                     uint8_t *__restrict__ warped, const uint32_t *gate)
{
    if (gate && *gate == 0) return;
    splat_resolve(blockIdx.x * blockDim.x + threadIdx.x, keys, vals, z, n, winpos, warped);
}
__global__ void splat_rgb_kernel(const uint32_t *__restrict__ vals, const uint32_t *__restrict__ winpos,
                                 const uint8_t *__restrict__ src, uint32_t n, uint8_t *__restrict__ warped,
                                 const uint32_t *gate)
{
    if (gate && *gate == 0) return;
    splat_rgb(blockIdx.x * blockDim.x + threadIdx.x, vals, winpos, src, n, warped);
}

// ---- chain path -------------------------------------------------------------------------------------------------
constexpr int MF_MAX_DEVICES_SPLAT = 64;
constexpr int CHAIN_MAX = 16;     // writers per target resolved in registers / local memory

// head[t] = -1 (memset) ; every source pixel k: next[k] = exchange(head[target(k)], k).  Thread 0 clears the overflow flag.
__global__ void __launch_bounds__(256)
splat_chain_build_kernel(const int64_t *__restrict__ idx, const int64_t *__restrict__ idy, int h, int w, uint32_t n,
                         int *__restrict__ head, int *__restrict__ next, uint32_t *__restrict__ overflow)
{
    const uint32_t k = blockIdx.x * 256u + threadIdx.x;
    if (k == 0) *overflow = 0;
    if (k >= n) return;
    const int x = (int)idx[k], y = (int)idy[k];               // narrowed through int exactly like warping.c:16-17
    if (x < 0 || x >= w || y < 0 || y >= h) return;           // out-of-range target: dropped (as the radix path does)
    next[k] = atomicExch(&head[(uint32_t)y * (uint32_t)w + (uint32_t)x], (int)k);
}

// one thread per target: its writers in raster order, then the serial loop's closed form (see the top of the file).
// A full block's 256 five-byte records are 320 aligned words: they are read, patched in shared memory and written
// back as words (bytes no writer touches keep the caller's content either way).
__global__ void __launch_bounds__(256)
splat_chain_resolve_kernel(const int *__restrict__ head, const int *__restrict__ next, const float *__restrict__ z,
                           const uint8_t *__restrict__ src, uint32_t n, uint8_t *__restrict__ warped,
                           uint32_t *__restrict__ overflow, int words_ok)
{
    __shared__ uint32_t stage[320];
    const uint32_t t0 = blockIdx.x * 256u, t = t0 + threadIdx.x;
    const bool staged = words_ok && t0 + 256u <= n;            // block-uniform
    uint32_t *gw = reinterpret_cast<uint32_t *>(warped + (size_t)t0 * 5);
    if (staged) {
        stage[threadIdx.x] = gw[threadIdx.x];
        if (threadIdx.x < 64) stage[256 + threadIdx.x] = gw[256 + threadIdx.x];
        __syncthreads();
    }
    int k = t < n ? head[t] : -1;
    int win = -1;
    float zprev = 1000.f;                                      // dlut[t] when the LAST writer arrives (:8-11, :28)
    bool touched = k >= 0;
    if (touched) {
        const int k2 = next[k];
        if (k2 < 0) {                                          // one writer: the common case
            if (z[k] < 1000.f) win = k;
        } else {
            // two to four writers (nearly all colliding targets): ordered in registers by a sorting network; the loads of
            // z are independent of each other
            const int k3 = next[k2], k4 = k3 >= 0 ? next[k3] : -1, k5 = k4 >= 0 ? next[k4] : -1;
            if (k5 < 0) {
                const int BIG = 0x7fffffff;
                int a = k, b = k2, c = k3 >= 0 ? k3 : BIG, d = k4 >= 0 ? k4 : BIG;
                auto cswap = [](int &x, int &y) { const int lo = min(x, y), hi = max(x, y); x = lo; y = hi; };
                cswap(a, b); cswap(c, d); cswap(a, c); cswap(b, d); cswap(b, c);
                const float za = z[a], zb = z[b], zc = c != BIG ? z[c] : 0.f, zd = d != BIG ? z[d] : 0.f;
                const int cnt = 2 + (c != BIG) + (d != BIG);
                if (za < zprev) win = a;                       // :19-21, then :28 unconditionally
                zprev = za;
                if (cnt == 2) { if (zb < zprev) win = b; }
                else {
                    if (zb < zprev) win = b;
                    zprev = zb;
                    if (cnt == 3) { if (zc < zprev) win = c; }
                    else {
                        if (zc < zprev) win = c;
                        zprev = zc;
                        if (zd < zprev) win = d;
                    }
                }
            } else {
                int ks[CHAIN_MAX];
                int cnt = 0;
                for (int c = k; c >= 0; c = next[c]) {
                    if (cnt == CHAIN_MAX) { *overflow = 1; touched = false; break; }   // left to the radix path
                    int j = cnt++;                             // insertion sort by raster index
                    while (j > 0 && ks[j - 1] > c) { ks[j] = ks[j - 1]; j--; }
                    ks[j] = c;
                }
                if (touched)
                    for (int i = 0; i < cnt; i++) {
                        const float zi = z[ks[i]];
                        if (zi < zprev) win = ks[i];           // :19-21 (for the last writer zprev is its dlut)
                        if (i + 1 < cnt) zprev = zi;           // :28 unconditionally
                    }
            }
        }
    }
    uint8_t *o = staged ? reinterpret_cast<uint8_t *>(stage) + threadIdx.x * 5 : warped + (size_t)t * 5;
    if (touched) {
        if (win >= 0) {
            const uint8_t *s = src + (size_t)win * 3;
            o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
        }
        o[3] = 1;                                              // :23
        o[4] = (zprev != 1000.f) ? 0 : 1;                      // :24-27
    }
    if (staged) {
        __syncthreads();
        gw[threadIdx.x] = stage[threadIdx.x];
        if (threadIdx.x < 64) gw[256 + threadIdx.x] = stage[256 + threadIdx.x];
    }
}

// ---- all phases in one cooperative launch ----
struct SplatArgs {
    const uint8_t *src; const int64_t *idx, *idy; const float *z; uint8_t *warped;
    int h, w; uint32_t n, nblocks; int passes, digit_bits; size_t hist_stride;
    uint32_t *kA, *vA, *kB, *vB, *winpos, *hist, *totals;
    const uint32_t *gate;      // non-null: run only if *gate != 0 (uniform over the grid)
};

__global__ void __launch_bounds__(SORT_THREADS)
splat_fused_kernel(const SplatArgs a)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    if (a.gate && *a.gate == 0) return;
    const uint32_t radix = 1u << a.digit_bits;
    for (uint32_t vb = blockIdx.x; vb < a.nblocks; vb += gridDim.x) {
        splat_keys_hist(vb, a.nblocks, a.idx, a.idy, a.h, a.w, a.n, a.digit_bits, a.nblocks, a.kA, a.winpos, a.hist,
                        a.hist + a.hist_stride, (uint32_t)(a.hist_stride * (a.passes - 1)));
        __syncthreads();      // the shared histogram is reused by the next virtual block
    }
    grid.sync();
    uint32_t *kA = a.kA, *vA = a.vA, *kB = a.kB, *vB = a.vB;
    for (int p = 0; p < a.passes; p++) {
        uint32_t *hp = a.hist + (size_t)p * a.hist_stride;
        const bool more = p + 1 < a.passes;
        for (uint32_t vb = blockIdx.x; vb < (radix + SORT_WARPS - 1) / SORT_WARPS; vb += gridDim.x)
            radix_offsets(vb, hp, a.nblocks, radix, a.totals);
        grid.sync();
        for (uint32_t vb = blockIdx.x; vb < a.nblocks; vb += gridDim.x) {
            if (p == 0)
                radix_scatter<true>(vb, kA, nullptr, a.n, 0, a.digit_bits, a.nblocks, hp, a.totals, kB, vB,
                                    more ? hp + a.hist_stride : nullptr, a.digit_bits, a.digit_bits);
            else
                radix_scatter<false>(vb, kA, vA, a.n, p * a.digit_bits, a.digit_bits, a.nblocks, hp, a.totals, kB, vB,
                                     more ? hp + a.hist_stride : nullptr, (p + 1) * a.digit_bits, a.digit_bits);
            __syncthreads();
        }
        grid.sync();
        uint32_t *t = kA; kA = kB; kB = t;
        t = vA; vA = vB; vB = t;
    }
    for (uint32_t i = blockIdx.x * SORT_THREADS + threadIdx.x; i < a.n; i += gridDim.x * SORT_THREADS)
        splat_resolve(i, kA, vA, a.z, a.n, a.winpos, a.warped);
    grid.sync();
    for (uint32_t t = blockIdx.x * SORT_THREADS + threadIdx.x; t < a.n; t += gridDim.x * SORT_THREADS)
        splat_rgb(t, vA, a.winpos, a.src, a.n, a.warped);
}

inline size_t al256(size_t v) { return (v + 255) / 256 * 256; }

struct SplatLayout {
    size_t keysA, valsA, keysB, valsB, winpos, hist, totals, flag, total;
    uint32_t nblocks;
    int passes, digit_bits;
    size_t hist_stride;   // uint32 per pass
    SplatLayout(size_t n)
    {
        nblocks = (uint32_t)((n + SORT_TILE - 1) / SORT_TILE);
        int bits = 0;
        while ((n >> bits) != 0) bits++;  // keys are in [0, n]
        passes = (bits + MAX_DIGIT_BITS - 1) / MAX_DIGIT_BITS;
        if (passes < 1) passes = 1;
        digit_bits = (bits + passes - 1) / passes;
        hist_stride = (size_t)(1u << digit_bits) * nblocks;
        size_t o = 0;
        keysA = o; o += al256(n * 4);
        valsA = o; o += al256(n * 4);
        keysB = o; o += al256(n * 4);
        valsB = o; o += al256(n * 4);
        winpos = o; o += al256(n * 4);
        hist = o; o += al256(hist_stride * passes * 4);
        totals = o; o += al256((size_t)MAX_RADIX * 4);
        flag = o; o += 256;          // chain path: 1 = some target has more than CHAIN_MAX writers
        total = o;
    }
};

}  // namespace

MPIFLOW_API size_t mpiflow_forward_warp_workspace_bytes(int h, int w)
{
    if (h <= 0 || w <= 0) return 0;
    return SplatLayout((size_t)h * w).total;
}

MPIFLOW_API int mpiflow_forward_warp_f32(const uint8_t *src, const int64_t *idx, const int64_t *idy,
                                         const float *z, uint8_t *warped, int h, int w, void *workspace,
                                         size_t workspace_bytes, void *stream)
{
    if (h == 0 || w == 0) return 0;  // the reference loops zero times
    MF_CHECK_ARG(src && idx && idy && z && warped && workspace && h > 0 && w > 0);
    MF_CHECK_ARG((int64_t)h * w < (1ll << 31) && ((uintptr_t)workspace & 255) == 0);
    const size_t n64 = (size_t)h * w;
    const SplatLayout L(n64);
    if (workspace_bytes < L.total) return MPIFLOW_E_WORKSPACE;
    const uint32_t n = (uint32_t)n64;
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = reinterpret_cast<char *>(workspace);
    uint32_t *kA = (uint32_t *)(ws + L.keysA), *vA = (uint32_t *)(ws + L.valsA);
    uint32_t *kB = (uint32_t *)(ws + L.keysB), *vB = (uint32_t *)(ws + L.valsB);
    uint32_t *winpos = (uint32_t *)(ws + L.winpos), *hist = (uint32_t *)(ws + L.hist);
    const unsigned g256 = (n + 255) / 256;

    uint32_t *totals = (uint32_t *)(ws + L.totals);
    const int db = L.digit_bits;
    const uint32_t radix = 1u << db;
    const size_t scatter_smem = (size_t)(SORT_WARPS + 1) * radix * sizeof(uint32_t);
    static std::once_flag attr_once;
    std::call_once(attr_once, []() {
        const int big = (SORT_WARPS + 1) * MAX_RADIX * (int)sizeof(uint32_t);
        cudaFuncSetAttribute(radix_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(radix_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(splat_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    });
    // chain path first (MPIFLOW_SPLAT_PATH=radix skips it); the radix path behind it runs only if it raised the flag
    static const bool chain_on = []() { const char *e = getenv("MPIFLOW_SPLAT_PATH"); return !(e && strcmp(e, "radix") == 0); }();
    uint32_t *gate = nullptr;
    if (chain_on) {
        int *head = (int *)kA, *next = (int *)vA;
        gate = (uint32_t *)(ws + L.flag);
        cudaError_t e = cudaMemsetAsync(head, 0xff, (size_t)n * 4, st);
        if (e != cudaSuccess) return (int)e;
        splat_chain_build_kernel<<<g256, 256, 0, st>>>(idx, idy, h, w, n, head, next, gate);
        splat_chain_resolve_kernel<<<g256, 256, 0, st>>>(head, next, z, src, n, warped, gate, ((uintptr_t)warped & 3) == 0);
        int rc = mf_launch_status();
        if (rc) return rc;
    }
    static const bool fused_on = []() { const char *e = getenv("MPIFLOW_SPLAT_FUSED"); return !(e && atoi(e) == 0); }();
    if (fused_on) {
        // as many CTAs as are co-resident (grid-wide barriers between the phases need all of them on the device);
        // the device queries are cached per device and shared-memory size (they cost microseconds of host time per call)
        static std::mutex mu;
        static int cache_grid[MF_MAX_DEVICES_SPLAT][2] = {};
        static size_t cache_smem[MF_MAX_DEVICES_SPLAT] = {};
        int dev = 0, sms = 0, per_sm = 0, coop = 0, full_grid = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < MF_MAX_DEVICES_SPLAT) {
            std::lock_guard<std::mutex> lock(mu);
            if (cache_grid[dev][0] && cache_smem[dev] == scatter_smem) full_grid = cache_grid[dev][1];
            else if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && coop &&
                     cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
                     cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, splat_fused_kernel, SORT_THREADS, scatter_smem) == cudaSuccess &&
                     per_sm > 0) {
                full_grid = sms * per_sm;
                cache_grid[dev][0] = 1; cache_grid[dev][1] = full_grid; cache_smem[dev] = scatter_smem;
            }
        }
        if (full_grid > 0) {
            SplatArgs a{src, idx, idy, z, warped, h, w, n, L.nblocks, L.passes, db, L.hist_stride, kA, vA, kB, vB, winpos, hist, totals, gate};
            unsigned grid = (unsigned)full_grid;
            if (grid > L.nblocks) grid = L.nblocks;
            void *args[] = {&a};
            cudaError_t e = cudaLaunchCooperativeKernel((void *)splat_fused_kernel, dim3(grid), dim3(SORT_THREADS), args,
                                                        scatter_smem, st);
            if (e == cudaSuccess) return mf_launch_status();
            (void)cudaGetLastError();      // not launchable here (e.g. inside a stream capture): one launch per phase
        }
    }
    splat_keys_hist_kernel<<<L.nblocks, SORT_THREADS, radix * sizeof(uint32_t), st>>>(
        idx, idy, h, w, n, db, L.nblocks, kA, winpos, hist, hist + L.hist_stride,
        (uint32_t)(L.hist_stride * (L.passes - 1)), gate);
    for (int p = 0; p < L.passes; p++) {
        uint32_t *hp = hist + (size_t)p * L.hist_stride;
        const bool more = p + 1 < L.passes;
        radix_offsets_kernel<<<(radix + SORT_WARPS - 1) / SORT_WARPS, SORT_THREADS, 0, st>>>(hp, L.nblocks, radix, totals, gate);
        if (p == 0)
            radix_scatter_kernel<true><<<L.nblocks, SORT_THREADS, scatter_smem, st>>>(
                kA, nullptr, n, 0, db, L.nblocks, hp, totals, kB, vB, more ? hp + L.hist_stride : nullptr, db, db, gate);
        else
            radix_scatter_kernel<false><<<L.nblocks, SORT_THREADS, scatter_smem, st>>>(
                kA, vA, n, p * db, db, L.nblocks, hp, totals, kB, vB, more ? hp + L.hist_stride : nullptr,
                (p + 1) * db, db, gate);
        uint32_t *t = kA; kA = kB; kB = t;
        t = vA; vA = vB; vB = t;
    }
    splat_resolve_kernel<<<g256, 256, 0, st>>>(kA, vA, z, n, winpos, warped, gate);
    splat_rgb_kernel<<<g256, 256, 0, st>>>(vA, winpos, src, n, warped, gate);
    return mf_launch_status();
}

// ---- host-pointer entry points ----------------------------------------------------------------
namespace {
// staging buffer of the host-pointer entry point, per device (the one current at the call); mpiflow_shutdown() frees them
constexpr int MF_MAX_DEVICES = 64;
std::mutex g_host_mu;
void *g_host_bufs[MF_MAX_DEVICES] = {nullptr};
size_t g_host_caps[MF_MAX_DEVICES] = {0};
}  // namespace

void mf_splat_shutdown()
{
    std::lock_guard<std::mutex> lock(g_host_mu);
    for (int d = 0; d < MF_MAX_DEVICES; d++) {
        if (g_host_bufs[d]) cudaFree(g_host_bufs[d]);
        g_host_bufs[d] = nullptr;
        g_host_caps[d] = 0;
    }
}

MPIFLOW_API int mpiflow_forward_warp_host(const void *src, const void *idx, const void *idy, const void *z,
                                          void *warped, int h, int w)
{
    if (h == 0 || w == 0) return 0;
    MF_CHECK_ARG(src && idx && idy && z && warped && h > 0 && w > 0 && (int64_t)h * w < (1ll << 31));
    std::lock_guard<std::mutex> lock(g_host_mu);
    const size_t n = (size_t)h * w;
    const size_t o_src = 0, o_idx = al256(n * 3), o_idy = o_idx + al256(n * 8), o_z = o_idy + al256(n * 8);
    const size_t o_out = o_z + al256(n * 4), o_ws = o_out + al256(n * 5);
    const size_t ws_bytes = mpiflow_forward_warp_workspace_bytes(h, w), need = o_ws + ws_bytes;
    cudaError_t e;
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e == cudaErrorNoDevice ? MPIFLOW_E_NODEVICE : (int)e;
    MF_CHECK_ARG(dev >= 0 && dev < MF_MAX_DEVICES);
    void *&g_host_buf = g_host_bufs[dev];
    size_t &g_host_cap = g_host_caps[dev];
    if (need > g_host_cap) {
        if (g_host_buf) cudaFree(g_host_buf);
        g_host_buf = nullptr; g_host_cap = 0;
        if ((e = cudaMalloc(&g_host_buf, need)) != cudaSuccess) return (int)e;
        g_host_cap = need;
    }
    char *d = reinterpret_cast<char *>(g_host_buf);
    cudaStream_t st = cudaStreamPerThread;
    if ((e = cudaMemcpyAsync(d + o_src, src, n * 3, cudaMemcpyHostToDevice, st)) != cudaSuccess) return (int)e;
    cudaMemcpyAsync(d + o_idx, idx, n * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d + o_idy, idy, n * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d + o_z, z, n * 4, cudaMemcpyHostToDevice, st);
    // bytes no writer touches keep the caller's content, exactly like the C loop
    cudaMemcpyAsync(d + o_out, warped, n * 5, cudaMemcpyHostToDevice, st);
    int rc = mpiflow_forward_warp_f32((const uint8_t *)(d + o_src), (const int64_t *)(d + o_idx),
                                      (const int64_t *)(d + o_idy), (const float *)(d + o_z),
                                      (uint8_t *)(d + o_out), h, w, d + o_ws, ws_bytes, st);
    if (rc) return rc;
    if ((e = cudaMemcpyAsync(warped, d + o_out, n * 5, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return (int)e;
    return (int)cudaStreamSynchronize(st);
}

// legacy symbol: external/forward_warping/warping.c:6 (void return, no error reporting)
MPIFLOW_API void forward_warping(const void *src, const void *idx, const void *idy, const void *z,
                                 void *warped, int h, int w)
{
    const int rc = mpiflow_forward_warp_host(src, idx, idy, z, warped, h, w);
    if (rc != 0) {
        // the reference cannot fail; a GPU failure must not pass silently as an all-zero image
        fprintf(stderr, "forward_warping (libmpiflow_sm100): %s (code %d)\n", mpiflow_error_string(rc), rc);
        abort();
    }
}
