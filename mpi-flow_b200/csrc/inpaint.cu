// inpaint.cu -- hole filling of the assembled sample on the device: cv2.inpaint(frame_mix,
// fill_mask, 3, cv2.INPAINT_NS) of the reference (utils/utils.py:284-286), bit for bit.
//
// OpenCV's Navier-Stokes inpainting is a fast-marching sweep: pixels leave a priority queue in order of their arrival
// time t (FIFO among equal t); when a pixel leaves, each of its hole neighbours (up, left, down, right) gets its time
// from the eikonal update, is inpainted AT ONCE from all non-hole pixels within the radius, and enters the queue.  A
// filled pixel is a source for every later one, so the result is defined by that global order (oracle/inpaint_oracle.c
// restates it and is pinned against cv2.inpaint).  What can run in parallel without changing a bit:
//   * images: a launch takes a batch (the generator's samples are independent);
//   * separate holes of one image: a filled pixel is read by later ones only within radius + 1 of it (window plus
//     gradient stencil), so hole pixels at least radius + 2 columns (or rows) apart never see each other's results.
//     inpaint_init_kernel cuts the image at hole-free runs of columns, then each column segment at hole-free runs of
//     rows, into up to MAXSEG rectangles; one warp marches each rectangle with its own queue;
//   * the window of one pixel: its (2r+1)^2 positions x 3 colours (flags, two gradient stencils, weight) go one per
//     lane; only the two running sums per colour are serial fp32 chains in row-major order, taken by six lanes from
//     shared memory over the order-preserving compaction of the valid positions;
//   * the queue: a new arrival time is at least 0.707 later than the pixel that caused it, so with buckets of width
//     0.5 nothing is ever inserted into the bucket being drained: the warp extracts the next non-empty bucket from the
//     pending list (stable ballot compaction), sorts it by (t, insertion number) with a bitonic network -- in
//     shared memory up to SORT_SMEM entries, in global memory beyond -- and pops it in order.
// All float steps use the explicit IEEE intrinsics (no contraction); the arithmetic types (float sqrt and quotient in
// the direction term, double quotient before the final rounding) are cv2's as pinned by the oracle tests.
#include "common.cuh"

namespace {

enum : uint8_t { KNOWN = 0, BAND = 1, INSIDE = 2 };
constexpr int MAX_RANGE = 5;                     // (2r+1)^2 = 121 window positions at most
constexpr int MAX_WIN = (2 * MAX_RANGE + 1) * (2 * MAX_RANGE + 1);
constexpr int SORT_SMEM = 512;                   // entries sorted in shared memory: the kernel's 9.6 KB fit next to the
                                                 // persistent renderer's 215 KB on the same SM, so it can run under it
constexpr int MAXC = 16;                         // column segments per image
constexpr int MAXSEG = 64;                       // rectangles (= marching warps) per image
constexpr unsigned FULL = 0xffffffffu;

struct Rect {
    int clo, chi, rlo, rhi;      // inclusive, in the frame-extended image; rlo > rhi: unused slot
    int offset;                  // its slice of the queue arrays
    int pad[3];
};

// Cuts the positions 1..n of an occupancy array at runs of >= gap - 1 free positions (occupied positions at least `gap`
// apart end up in different segments), at most `maxseg` segments; a segment spans its first occupied position - 1 to
// its last occupied position + 1 (the band around the hole).  Returns the number of segments.
__device__ int cut_segments(const uint8_t *occ, int n, int gap, int maxseg, int *lo, int *hi)
{
    int ns = 0, last = -(1 << 28);
    for (int j = 1; j <= n; j++) {
        if (!occ[j]) continue;
        if (ns == 0 || (j - last >= gap && ns < maxseg)) lo[ns++] = j - 1;
        hi[ns - 1] = j + 1;
        last = j;
    }
    return ns;
}

// One CTA per image.  Flags / arrival times of the image extended by a 1-pixel frame: hole INSIDE, its non-hole
// 4-neighbours BAND (t = 0), everything else (and the whole frame) KNOWN with t = 1e6 (icvInpaint: cross dilation minus
// mask, SET_BORDER1); the initial band as one row-major list (every warp counts the band pixels of its rows, the counts
// are prefix-summed, every warp compacts its rows); and the rectangles that can be marched independently.
__global__ void __launch_bounds__(1024)
inpaint_init_kernel(const uint8_t *__restrict__ mask_all, int rows, int cols, int range, int maxc,
                    uint8_t *__restrict__ f_all, float *__restrict__ t_all, uint32_t *__restrict__ bandp_all,
                    int *__restrict__ nband, Rect *__restrict__ rects_all, size_t cap)
{
    extern __shared__ int smi[];
    const int erows = rows + 2, ecols = cols + 2, im = blockIdx.x;
    int *rowcnt = smi;                                   // erows + 1
    int *clo = rowcnt + erows + 1, *chi = clo + MAXC;    // column segments
    int *ncs_p = chi + MAXC;
    Rect *rect = reinterpret_cast<Rect *>(ncs_p + 4);    // MAXSEG (32-byte records; the ints before keep it 16-byte aligned)
    uint8_t *colocc = reinterpret_cast<uint8_t *>(rect + MAXSEG);   // ecols
    uint8_t *colseg = colocc + ecols;                    // ecols
    uint8_t *rowocc = colseg + ecols;                    // maxc * erows
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const uint8_t *m = mask_all + (size_t)im * rows * cols;
    uint8_t *f = f_all + (size_t)im * erows * ecols;
    float *t = t_all + (size_t)im * erows * ecols;
    uint32_t *bandp = bandp_all + cap * im;
    for (int j = threadIdx.x; j < ecols; j += blockDim.x) { colocc[j] = 0; colseg[j] = 0; }
    for (int j = threadIdx.x; j < maxc * erows; j += blockDim.x) rowocc[j] = 0;
    __syncthreads();
    auto hole = [&](int y, int x) { return y >= 1 && x >= 1 && y <= rows && x <= cols && m[(y - 1) * cols + (x - 1)] != 0; };
    for (int i = warp; i < erows; i += nwarps) {
        int cnt = 0;
        for (int j0 = 0; j0 < ecols; j0 += 32) {
            const int j = j0 + lane;
            bool band = false;
            if (j < ecols) {
                uint8_t fv = KNOWN;
                float tv = 1.0e6f;
                if (hole(i, j)) { fv = INSIDE; colocc[j] = 1; }
                else if (i > 0 && j > 0 && i < erows - 1 && j < ecols - 1 &&
                         (hole(i - 1, j) || hole(i + 1, j) || hole(i, j - 1) || hole(i, j + 1))) {
                    fv = BAND; tv = 0.f; band = true;
                }
                f[i * ecols + j] = fv;
                t[i * ecols + j] = tv;
            }
            cnt += __popc(__ballot_sync(FULL, band));
        }
        if (lane == 0) rowcnt[i] = cnt;
    }
    __syncthreads();
    if (warp == 0) {            // exclusive prefix over the rows
        int carry = 0;
        for (int i0 = 0; i0 < erows; i0 += 32) {
            const int i = i0 + lane;
            const int v = i < erows ? rowcnt[i] : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += u;
            }
            if (i < erows) rowcnt[i] = carry + inc - v;
            carry += __shfl_sync(FULL, inc, 31);
        }
        if (lane == 0) nband[im] = carry;
    } else if (threadIdx.x == 32) {   // column segments
        const int ncs = cut_segments(colocc, cols, range + 2, maxc, clo, chi);
        *ncs_p = ncs;
        for (int c = 0; c < ncs; c++)
            for (int j = clo[c]; j <= chi[c]; j++) colseg[j] = (uint8_t)c;
    }
    __syncthreads();
    const int ncs = *ncs_p;
    const unsigned lt = (1u << lane) - 1u;
    for (int i = warp; i < erows; i += nwarps) {
        int at = rowcnt[i];
        for (int j0 = 0; j0 < ecols; j0 += 32) {
            const int j = j0 + lane;
            const uint8_t fv = j < ecols ? f[i * ecols + j] : (uint8_t)KNOWN;
            if (fv == INSIDE) rowocc[colseg[j] * erows + i] = 1;
            const unsigned b = __ballot_sync(FULL, fv == BAND);
            if (fv == BAND) bandp[at + __popc(b & lt)] = (uint32_t)(i * ecols + j);
            at += __popc(b);
        }
    }
    __syncthreads();
    // rows of every column segment -> rectangles in fixed slots: segment c owns slots c * per .. c * per + per - 1
    const int per = ncs > 0 ? MAXSEG / ncs : MAXSEG;
    if ((int)threadIdx.x < MAXSEG) rect[threadIdx.x] = Rect{0, 0, 1, 0, 0, {0, 0, 0}};
    __syncthreads();
    if ((int)threadIdx.x < ncs) {
        const int c = threadIdx.x;
        int rlo[MAXSEG], rhi[MAXSEG];
        const int nr = cut_segments(rowocc + c * erows, rows, range + 2, per, rlo, rhi);
        for (int r = 0; r < nr; r++) rect[c * per + r] = Rect{clo[c], chi[c], rlo[r], rhi[r], 0, {0, 0, 0}};
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int off = 0;
        for (int r = 0; r < MAXSEG; r++) {
            rect[r].offset = off;
            if (rect[r].rlo <= rect[r].rhi) off += (rect[r].chi - rect[r].clo + 1) * (rect[r].rhi - rect[r].rlo + 1);
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < MAXSEG) rects_all[(size_t)im * MAXSEG + threadIdx.x] = rect[threadIdx.x];
}

// FastMarching_solve: eikonal update of a pixel from two of its 4-neighbours, in double like OpenCV
__device__ __forceinline__ float fmm_solve(int p1, int p2, const uint8_t *f, const float *t)
{
    const double a11 = (double)t[p1], a22 = (double)t[p2], m12 = a11 < a22 ? a11 : a22;
    double sol;
    if (f[p1] != INSIDE) {
        if (f[p2] != INSIDE) {
            const double d = a11 - a22;
            if (fabs(d) >= 1.0) sol = 1.0 + m12;
            else sol = __dmul_rn(__dadd_rn(__dadd_rn(a11, a22), sqrt(__dsub_rn(2.0, __dmul_rn(d, d)))), 0.5);
        } else sol = 1.0 + a11;
    } else if (f[p2] != INSIDE) sol = 1.0 + a22;
    else sol = 1.0 + m12;
    return (float)sol;
}

// ascending bitonic sort of n (key, pos) pairs in shared or global memory by one warp.  Every merge runs in ascending
// direction (first step against the mirrored partner), so the padding up to the next power of two is virtual: a
// partner beyond n counts as +infinity and is never touched.
__device__ void bitonic_sort(unsigned long long *keys, uint32_t *pos, int n, int lane)
{
    for (int k = 2; (k >> 1) < n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < n; i += 32) {
                const int partner = (j == (k >> 1)) ? (i ^ (k - 1)) : (i ^ j);
                if (partner > i && partner < n) {
                    const unsigned long long a = keys[i], b = keys[partner];
                    if (a > b) {
                        keys[i] = b; keys[partner] = a;
                        const uint32_t pa = pos[i];
                        pos[i] = pos[partner]; pos[partner] = pa;
                    }
                }
            }
            __syncwarp();
        }
}

struct MarchSmem {
    float w[3][MAX_WIN];      // weights of the valid window positions, compacted in row-major order, per colour
    float wi[3][MAX_WIN];     // weight * pixel value
    int off[MAX_WIN];         // the window's in-circle offsets in row-major order: dk << 16 | (dl & 0xffff)
    unsigned long long skey[SORT_SMEM];
    uint32_t spos[SORT_SMEM];
};

// weight of one window pixel for one colour (icvNSInpaintFMM): gx / gy the two gradient sums, dst = 1 / (|r|^4 + 1)
__device__ __forceinline__ float ns_weight(float gx, float gy, float rx, float ry, float rl, float dst)
{
    gx = -gx;
    float dir = __fadd_rn(__fmul_rn(rx, gx), __fmul_rn(ry, gy));
    if (fabs((double)dir) <= 0.01) dir = 0.000001f;
    else dir = fabsf(__fdiv_rn(dir, __fsqrt_rn(__fmul_rn(rl, __fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy))))));
    return __fmul_rn(dst, dir);
}

// One warp per (rectangle, image).
__global__ void __launch_bounds__(32)
inpaint_ns_march_kernel(int rows, int cols, int range, uint8_t *f_all, float *t_all, uint8_t *out_all,
                        unsigned long long *curk_all, uint32_t *curp_all, unsigned long long *pendk_all,
                        uint32_t *pendp_all, const uint32_t *bandp_all, const int *nband, const Rect *rects_all, size_t cap)
{
    __shared__ MarchSmem sm;
    const int lane = threadIdx.x, im = blockIdx.y;
    const Rect rc = rects_all[(size_t)im * MAXSEG + blockIdx.x];
    if (rc.rlo > rc.rhi) return;
    const int erows = rows + 2, ecols = cols + 2;
    const size_t per = (size_t)erows * ecols;
    // The warp re-reads what one of its lanes wrote a moment ago.  Plain accesses (L1 hits) are enough: every write is
    // followed by a __syncwarp(), which orders memory among the lanes and keeps the compiler from caching values.
    uint8_t *f = f_all + per * im;
    float *t = t_all + per * im;
    uint8_t *out = out_all + (size_t)im * rows * cols * 3;
    unsigned long long *curk = curk_all + cap * im + rc.offset, *pendk = pendk_all + cap * im + rc.offset;
    uint32_t *curp = curp_all + cap * im + rc.offset, *pendp = pendp_all + cap * im + rc.offset;
    const uint32_t *bandp = bandp_all + cap * im;
    const unsigned lt = (1u << lane) - 1u;
    const int win = 2 * range + 1;

    // the window positions that can contribute: inside the circle, in OpenCV's row-major order (r = 3: 29 of the 49,
    // one lane each)
    int nwin = 0;
    for (int base = 0; base < win * win; base += 32) {
        const int p = base + lane, dk = p / win - range, dl = p - (p / win) * win - range;
        const bool in = p < win * win && dk * dk + dl * dl <= range * range;
        const unsigned b = __ballot_sync(FULL, in);
        if (in) sm.off[nwin + __popc(b & lt)] = (dk << 16) | (dl & 0xffff);
        nwin += __popc(b);
    }
    // ---- the rectangle's part of the initial band (row-major, t = 0): bucket 0, in pop order already ----
    int nc = 0;
    {
        const int nb = nband[im];
        for (int base = 0; base < nb; base += 32) {
            const int e = base + lane;
            uint32_t pos = 0;
            bool in = false;
            if (e < nb) {
                pos = bandp[e];
                const int i = pos / ecols, j = pos - i * ecols;
                in = i >= rc.rlo && i <= rc.rhi && j >= rc.clo && j <= rc.chi;
            }
            const unsigned b = __ballot_sync(FULL, in);
            if (in) curp[nc + __popc(b & lt)] = pos;
            nc += __popc(b);
        }
    }
    __syncwarp();
    // this lane's window position of the first 32 (r = (i - k, j - l), |r|^2, 1 / (|r|^4 + 1))
    const int od0 = sm.off[lane < nwin ? lane : 0], dk0 = od0 >> 16, dl0 = (int)(short)(od0 & 0xffff);
    const float ry0 = (float)(-dk0), rx0 = (float)(-dl0);
    const float rl0 = __fadd_rn(__fmul_rn(rx0, rx0), __fmul_rn(ry0, ry0));
    const float dst0 = __fdiv_rn(1.f, __fadd_rn(__fmul_rn(rl0, rl0), 1.f));
    uint32_t seq = (uint32_t)nc;
    int np = 0;

    for (;;) {
        // ---- drain the current bucket in order ----
        for (int e = 0; e < nc; e++) {
            const int pos = (int)curp[e];
            const int ii = pos / ecols, jj = pos - ii * ecols;
            if (lane == 0) f[pos] = KNOWN;
            // hole neighbours in OpenCV's order: up, left, down, right (frame rows / columns 0 are skipped: i <= 0, j <= 0)
            int qi = ii, qj = jj;
            if (lane == 0) qi = ii - 1; else if (lane == 1) qj = jj - 1; else if (lane == 2) qi = ii + 1; else if (lane == 3) qj = jj + 1;
            const bool cand = lane < 4 && qi > 0 && qj > 0 && qi <= erows - 1 && qj <= ecols - 1 && f[qi * ecols + qj] == INSIDE;
            unsigned todo = __ballot_sync(FULL, cand);
            while (todo) {
                const int q = __ffs(todo) - 1;
                todo &= todo - 1;
                const int i = q == 0 ? ii - 1 : q == 2 ? ii + 1 : ii, j = q == 1 ? jj - 1 : q == 3 ? jj + 1 : jj;
                const int ij = i * ecols + j;
                // arrival time: min4 of the four quadrant updates
                float sol = 3.0e38f;
                if (lane < 4) {
                    const int p1 = (lane & 1) ? ij + ecols : ij - ecols, p2 = (lane & 2) ? ij + 1 : ij - 1;
                    sol = fmm_solve(p1, p2, f, t);
                }
                sol = fminf(sol, __shfl_xor_sync(FULL, sol, 1));
                sol = fminf(sol, __shfl_xor_sync(FULL, sol, 2));
                const float dist = __shfl_sync(FULL, sol, 0);

                // ---- the window: one position per lane, all three colours ----
                // away from the image border the gradient stencils are the plain 4-neighbours of the window pixel
                const bool interior = i - range >= 2 && i + range <= erows - 3 && j - range >= 2 && j + range <= ecols - 3;
                int nvalid = 0;
                for (int base = 0; base < nwin; base += 32) {
                    const int p = base + lane;
                    int dk = dk0, dl = dl0;                       // radius <= 3: one position per lane, fixed for the launch
                    float ry = ry0, rx = rx0, rl = rl0, dst = dst0;
                    if (base) {
                        const int od = sm.off[p < nwin ? p : 0];
                        dk = od >> 16; dl = (int)(short)(od & 0xffff);
                        ry = (float)(-dk); rx = (float)(-dl);
                        rl = __fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry));
                        dst = __fdiv_rn(1.f, __fadd_rn(__fmul_rn(rl, rl), 1.f));
                    }
                    const int k = i + dk, l = j + dl, kl = k * ecols + l;
                    bool valid = p < nwin && (interior || (k > 0 && l > 0 && k < erows - 1 && l < ecols - 1));
                    if (valid) valid = f[kl] != INSIDE;
                    float wv[3] = {0.f, 0.f, 0.f}, wiv[3] = {0.f, 0.f, 0.f};
                    if (valid) {
                        const bool fdn = f[kl + ecols] != INSIDE, fup = f[kl - ecols] != INSIDE;
                        const bool frt = f[kl + 1] != INSIDE, flf = f[kl - 1] != INSIDE;
                        if (interior) {
                            const uint8_t *c0 = out + ((k - 1) * cols + (l - 1)) * 3;
                            const int rs = cols * 3;
#pragma unroll
                            for (int c = 0; c < 3; c++) {
                                const int ce = c0[c], dn = c0[rs + c], up = c0[c - rs], rt = c0[3 + c], lf = c0[c - 3];
                                float gx, gy;
                                if (fdn) gx = fup ? (float)(abs(dn - ce) + abs(ce - up)) : (float)abs(dn - ce) * 2.0f;
                                else gx = fup ? (float)abs(ce - up) * 2.0f : 0.f;
                                if (frt) gy = flf ? (float)(abs(rt - ce) + abs(ce - lf)) : (float)abs(rt - ce) * 2.0f;
                                else gy = flf ? (float)abs(ce - lf) * 2.0f : 0.f;
                                const float w = ns_weight(gx, gy, rx, ry, rl, dst);
                                wv[c] = w;
                                wiv[c] = __fmul_rn(w, (float)ce);
                            }
                        } else {
                            // at the border OpenCV shifts the stencil rows / columns inwards (km, kp, lm, lp); the value
                            // itself is read at the pixel
                            const int km = k - 1 + (k == 1), kp = k - 1 - (k == erows - 2);
                            const int lm = l - 1 + (l == 1), lp = l - 1 - (l == ecols - 2);
                            const int o_a = ((kp + 1) * cols + lm) * 3, o_b = (kp * cols + lm) * 3;
                            const int o_c = ((km - 1) * cols + lm) * 3, o_d = (km * cols + lp + 1) * 3;
                            const int o_e = (km * cols + lm) * 3, o_f = (km * cols + lm - 1) * 3;
                            const int o_v = ((k - 1) * cols + (l - 1)) * 3;
#pragma unroll
                            for (int c = 0; c < 3; c++) {
                                const int a = out[o_a + c], b = out[o_b + c], cc = out[o_c + c];
                                const int d = out[o_d + c], ee = out[o_e + c], ff = out[o_f + c];
                                float gx, gy;
                                if (fdn) gx = fup ? (float)(abs(a - b) + abs(b - cc)) : (float)abs(a - b) * 2.0f;
                                else gx = fup ? (float)abs(b - cc) * 2.0f : 0.f;
                                if (frt) gy = flf ? (float)(abs(d - ee) + abs(ee - ff)) : (float)abs(d - ee) * 2.0f;
                                else gy = flf ? (float)abs(ee - ff) * 2.0f : 0.f;
                                const float w = ns_weight(gx, gy, rx, ry, rl, dst);
                                wv[c] = w;
                                wiv[c] = __fmul_rn(w, (float)(int)out[o_v + c]);
                            }
                        }
                    }
                    const unsigned vb = __ballot_sync(FULL, valid);
                    if (valid) {
                        const int slot = nvalid + __popc(vb & lt);
#pragma unroll
                        for (int c = 0; c < 3; c++) { sm.w[c][slot] = wv[c]; sm.wi[c][slot] = wiv[c]; }
                    }
                    nvalid += __popc(vb);
                }
                __syncwarp();
                // ---- the six running sums, each a serial fp32 chain in window order ----
                float acc = lane < 3 ? 0.f : 1.0e-20f;
                if (lane < 6) {
                    const float *src = lane < 3 ? sm.wi[lane] : sm.w[lane - 3];
#pragma unroll 4
                    for (int v = 0; v < nvalid; v++) acc = __fadd_rn(acc, src[v]);
                }
                const float s = __shfl_sync(FULL, acc, (lane + 3) & 31);
                if (lane < 3) {
                    const double val = __ddiv_rn((double)acc, (double)s);
                    int r = __double2int_rn(val);          // cvRound: to nearest even
                    r = r < 0 ? 0 : r > 255 ? 255 : r;
                    out[((i - 1) * cols + (j - 1)) * 3 + lane] = (uint8_t)r;
                }
                if (lane == 0) {
                    t[ij] = dist;
                    f[ij] = BAND;
                    pendk[np] = ((unsigned long long)__float_as_uint(dist) << 32) | seq;
                    pendp[np] = (uint32_t)ij;
                }
                np++;
                seq++;
                __syncwarp();
            }
        }
        if (np == 0) break;
        // ---- next bucket: smallest floor(2 t) among the pending entries ----
        int minb = 0x7fffffff;
        for (int e = lane; e < np; e += 32) minb = min(minb, (int)(__uint_as_float((uint32_t)(pendk[e] >> 32)) * 2.0f));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) minb = min(minb, __shfl_xor_sync(FULL, minb, o));
        // stable split of the pending list: this bucket -> cur, the rest stays (compacted in place)
        int nkeep = 0;
        nc = 0;
        for (int base = 0; base < np; base += 32) {
            const int e = base + lane;
            unsigned long long key = 0;
            uint32_t ps = 0;
            bool take = false, keep = false;
            if (e < np) {
                key = pendk[e]; ps = pendp[e];
                take = (int)(__uint_as_float((uint32_t)(key >> 32)) * 2.0f) == minb;
                keep = !take;
            }
            const unsigned tb = __ballot_sync(FULL, take), kb = __ballot_sync(FULL, keep);
            __syncwarp();
            if (take) { const int d = nc + __popc(tb & lt); curk[d] = key; curp[d] = ps; }
            if (keep) { const int d = nkeep + __popc(kb & lt); pendk[d] = key; pendp[d] = ps; }
            nc += __popc(tb);
            nkeep += __popc(kb);
            __syncwarp();
        }
        np = nkeep;
        // ---- sort the bucket by (t, insertion number) ----
        if (nc > 1) {
            if (nc <= SORT_SMEM) {
                for (int e = lane; e < nc; e += 32) { sm.skey[e] = curk[e]; sm.spos[e] = curp[e]; }
                __syncwarp();
                bitonic_sort(sm.skey, sm.spos, nc, lane);
                for (int e = lane; e < nc; e += 32) curp[e] = sm.spos[e];
            } else {
                bitonic_sort(curk, curp, nc, lane);
            }
            __syncwarp();
        }
    }
}

}  // namespace

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
// shared memory of inpaint_init_kernel without the row occupancies, and the column segments whose row occupancies fit
// beside it (0: the image is too tall for the kernel)
static size_t inpaint_init_fixed_smem(int rows, int cols)
{
    return (size_t)(rows + 3 + 2 * MAXC + 4) * sizeof(int) + MAXSEG * sizeof(Rect) + 2 * (size_t)(cols + 2);
}
static int inpaint_maxc(int rows, int cols)
{
    const size_t budget = 200 * 1024, fixed = inpaint_init_fixed_smem(rows, cols);
    if (fixed + (size_t)(rows + 2) > budget) return 0;
    const size_t m = (budget - fixed) / (size_t)(rows + 2);
    return m > (size_t)MAXC ? MAXC : (int)m;
}

MPIFLOW_API size_t mpiflow_inpaint_ns_workspace_bytes(int n, int rows, int cols)
{
    if (n <= 0 || rows <= 0 || cols <= 0) return 0;
    const size_t per = (size_t)(rows + 2) * (cols + 2);      // every pixel enters a queue at most once
    return align256(per * n) + align256(per * n * 4) + 2 * align256(per * n * 8) + 3 * align256(per * n * 4) +
           align256((size_t)n * sizeof(int)) + align256((size_t)n * MAXSEG * sizeof(Rect));
}

MPIFLOW_API int mpiflow_inpaint_ns_u8(const uint8_t *img, const uint8_t *mask, int n, int rows, int cols, int radius,
                                      uint8_t *out, void *workspace, size_t workspace_bytes, void *stream)
{
    MF_CHECK_ARG(img && mask && out && workspace);
    MF_CHECK_ARG(n > 0 && n <= 65535 && rows > 0 && cols > 0 && (int64_t)(rows + 2) * (cols + 2) * 3 < (1ll << 31));
    MF_CHECK_ARG(rows <= 16000 && cols <= 16000);
    MF_CHECK_ARG(((uintptr_t)workspace & 255) == 0);
    int range = radius < 1 ? 1 : radius;        // icvInpaint: range = cvRound(radius) clamped to [1, 100]
    MF_CHECK_ARG(range <= MAX_RANGE);
    if (workspace_bytes < mpiflow_inpaint_ns_workspace_bytes(n, rows, cols)) return MPIFLOW_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t per = (size_t)(rows + 2) * (cols + 2);
    char *w = reinterpret_cast<char *>(workspace);
    uint8_t *f = reinterpret_cast<uint8_t *>(w); w += align256(per * n);
    float *t = reinterpret_cast<float *>(w); w += align256(per * n * 4);
    unsigned long long *curk = reinterpret_cast<unsigned long long *>(w); w += align256(per * n * 8);
    unsigned long long *pendk = reinterpret_cast<unsigned long long *>(w); w += align256(per * n * 8);
    uint32_t *curp = reinterpret_cast<uint32_t *>(w); w += align256(per * n * 4);
    uint32_t *pendp = reinterpret_cast<uint32_t *>(w); w += align256(per * n * 4);
    uint32_t *bandp = reinterpret_cast<uint32_t *>(w); w += align256(per * n * 4);
    int *nband = reinterpret_cast<int *>(w); w += align256((size_t)n * sizeof(int));
    Rect *rects = reinterpret_cast<Rect *>(w);
    if (out != img) {
        cudaError_t e = cudaMemcpyAsync(out, img, (size_t)n * rows * cols * 3, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return (int)e;
    }
    const int maxc = inpaint_maxc(rows, cols);
    MF_CHECK_ARG(maxc >= 1);                   // rows + cols beyond what the init kernel holds in shared memory
    const size_t smem = inpaint_init_fixed_smem(rows, cols) + (size_t)maxc * (rows + 2);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(inpaint_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    inpaint_init_kernel<<<n, 1024, smem, st>>>(mask, rows, cols, range, maxc, f, t, bandp, nband, rects, per);
    int rc = mf_launch_status();
    if (rc) return rc;
    inpaint_ns_march_kernel<<<dim3(MAXSEG, n), 32, 0, st>>>(rows, cols, range, f, t, out, curk, curp, pendk, pendp, bandp,
                                                          nband, rects, per);
    return mf_launch_status();
}
