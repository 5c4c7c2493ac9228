/* inpaint_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU checker; nothing under mpi-flow_b200/ links or loads it).
 *
 * Plain-C restatement of cv2.inpaint(img, mask, radius, cv2.INPAINT_NS) for 8-bit 3-channel images, the hole filling
 * of the reference's sample assembly (reference utils/utils.py:284-286; radius 3).  OpenCV is an un-vendored
 * dependency of the reference (README pins opencv-python; this image has cv2 4.13): the algorithm restated here is
 * OpenCV's published fast-marching Navier-Stokes inpainting (modules/photo/src/inpaint.cpp: icvInpaint,
 * icvNSInpaintFMM, FastMarching_solve, CvPriorityQueueFloat).  Pinned bit for bit against cv2.inpaint itself:
 * tests/test_inpaint_oracle.py (random images / masks incl. masks touching the image border) and the committed
 * fixtures tests/golden/inpaint.npz (made by tests/golden/make_inpaint_golden.py).
 *
 * Structure of the algorithm (all coordinates below are in the image extended by a 1-pixel frame, as OpenCV does):
 *   flags f: KNOWN 0, BAND 1, INSIDE 2; arrival times t (float), 1e6 initially, 0 on the initial band;
 *   initial band = 4-neighbours of the hole that are not hole (cross dilation minus mask), frame excluded;
 *   priority queue ordered by t, FIFO among equal t (OpenCV's sorted linked list inserts behind equal keys),
 *   filled row-major with the initial band;
 *   pop (ii,jj) -> KNOWN; each 4-neighbour (up, left, down, right) that is INSIDE gets its time from the eikonal
 *   update, is inpainted AT ONCE from all non-INSIDE pixels within the radius (weights 1/(|r|^4+1) * |r . grad_perp I|
 *   / sqrt(|r|^2 |grad I|^2)), becomes BAND and is pushed.  The result therefore depends on the global order, which is
 *   why the device version (csrc/inpaint.cu) keeps one image's march on one warp.
 *
 * Three details were settled against cv2 4.13 itself (3.2 M filled values, random images, masks and radii 1..5, holes
 * touching the image border; every other reading differs from cv2 on those):
 *   - a popped pixel's neighbour is skipped only on the frame (i <= 0 || j <= 0), so image row / column 0 is filled;
 *   - the value of a window pixel is read at the pixel itself, (k-1, l-1) in image coordinates; only the gradient
 *     stencils use the border-adjusted indices km / kp / lm / lp;
 *   - the direction term is float: fabsf(dir / sqrtf(|r|^2 |grad|^2)); the final quotient Ia / s is double.
 * `variant` != 0 switches single readings back (bit 0: skip i <= 1; bit 1: double sqrt and quotient; bit 2: value at
 * (km, lm)) -- used by the pinning test to show that the fixtures discriminate between them.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { KNOWN = 0, BAND = 1, INSIDE = 2 };

typedef struct { float t; uint32_t seq; int i, j; } HeapElem;
typedef struct { HeapElem *e; int n; uint32_t seq; } Heap;

static int heap_less(const HeapElem *a, const HeapElem *b) { return a->t < b->t || (a->t == b->t && a->seq < b->seq); }
static void heap_push(Heap *h, int i, int j, float t)
{
    int c = h->n++;
    HeapElem x = {t, h->seq++, i, j};
    while (c > 0) {
        int p = (c - 1) >> 1;
        if (!heap_less(&x, &h->e[p])) break;
        h->e[c] = h->e[p];
        c = p;
    }
    h->e[c] = x;
}
static int heap_pop(Heap *h, int *i, int *j)
{
    if (h->n == 0) return 0;
    *i = h->e[0].i; *j = h->e[0].j;
    HeapElem x = h->e[--h->n];
    int c = 0;
    for (;;) {
        int l = 2 * c + 1, r = l + 1, m = l;
        if (l >= h->n) break;
        if (r < h->n && heap_less(&h->e[r], &h->e[l])) m = r;
        if (!heap_less(&h->e[m], &x)) break;
        h->e[c] = h->e[m];
        c = m;
    }
    if (h->n) h->e[c] = x;
    return 1;
}

/* eikonal update from the neighbours (i1,j1), (i2,j2) (FastMarching_solve) */
static float fmm_solve(int i1, int j1, int i2, int j2, const uint8_t *f, const float *t, int ecols)
{
    double sol, a11 = t[i1 * ecols + j1], a22 = t[i2 * ecols + j2], m12 = a11 < a22 ? a11 : a22;
    if (f[i1 * ecols + j1] != INSIDE) {
        if (f[i2 * ecols + j2] != INSIDE) {
            if (fabs(a11 - a22) >= 1.0) sol = 1 + m12;
            else sol = (a11 + a22 + sqrt((double)(2 - (a11 - a22) * (a11 - a22)))) * 0.5;
        } else sol = 1 + a11;
    } else if (f[i2 * ecols + j2] != INSIDE) sol = 1 + a22;
    else sol = 1 + m12;
    return (float)sol;
}
static float min4f(float a, float b, float c, float d)
{
    a = a < b ? a : b;
    c = c < d ? c : d;
    return a < c ? a : c;
}

#define OUT(r, c, ch) ((int)out[((size_t)(r) * cols + (c)) * 3 + (ch)])

/* img: rows x cols x 3 u8 (copied to out first), mask: rows x cols u8 (non-zero = hole). */
int orc_inpaint_ns(const uint8_t *img, const uint8_t *mask, int rows, int cols, int range, int variant, uint8_t *out)
{
    const int erows = rows + 2, ecols = cols + 2;
    if (range < 1) range = 1;
    if (range > 100) range = 100;
    memcpy(out, img, (size_t)rows * cols * 3);
    uint8_t *f = (uint8_t *)calloc((size_t)erows * ecols, 1), *m = (uint8_t *)calloc((size_t)erows * ecols, 1);
    float *t = (float *)malloc(sizeof(float) * erows * ecols);
    size_t cap = 0;
    for (int i = 0; i < erows * ecols; i++) t[i] = 1.0e6f;
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++)
            if (mask[(size_t)i * cols + j]) m[(i + 1) * ecols + j + 1] = 1;
    Heap h = {NULL, 0, 0};
    /* band = cross-dilation of the hole minus the hole, frame excluded; row-major into the queue with t = 0 */
    for (int i = 0; i < erows; i++)
        for (int j = 0; j < ecols; j++) {
            if (m[i * ecols + j]) { f[i * ecols + j] = INSIDE; cap++; continue; }
            int nb = (i > 0 && m[(i - 1) * ecols + j]) || (i + 1 < erows && m[(i + 1) * ecols + j]) ||
                     (j > 0 && m[i * ecols + j - 1]) || (j + 1 < ecols && m[i * ecols + j + 1]);
            if (nb && i > 0 && j > 0 && i < erows - 1 && j < ecols - 1) { f[i * ecols + j] = BAND; cap++; }
        }
    h.e = (HeapElem *)malloc(sizeof(HeapElem) * (cap + 2));
    for (int i = 0; i < erows; i++)
        for (int j = 0; j < ecols; j++)
            if (f[i * ecols + j] == BAND) { t[i * ecols + j] = 0.f; heap_push(&h, i, j, 0.f); }

    int ii, jj;
    while (heap_pop(&h, &ii, &jj)) {
        f[ii * ecols + jj] = KNOWN;
        for (int q = 0; q < 4; q++) {
            int i, j;
            if (q == 0) { i = ii - 1; j = jj; }
            else if (q == 1) { i = ii; j = jj - 1; }
            else if (q == 2) { i = ii + 1; j = jj; }
            else { i = ii; j = jj + 1; }
            if (variant & 1) { if (i <= 1 || j <= 1 || i > erows - 1 || j > ecols - 1) continue; }
            else { if (i <= 0 || j <= 0 || i > erows - 1 || j > ecols - 1) continue; }
            if (f[i * ecols + j] != INSIDE) continue;
            const float dist = min4f(fmm_solve(i - 1, j, i, j - 1, f, t, ecols), fmm_solve(i + 1, j, i, j - 1, f, t, ecols),
                                     fmm_solve(i - 1, j, i, j + 1, f, t, ecols), fmm_solve(i + 1, j, i, j + 1, f, t, ecols));
            t[i * ecols + j] = dist;
            for (int color = 0; color <= 2; color++) {
                float Ia = 0, s = 1.0e-20f;
                for (int k = i - range; k <= i + range; k++) {
                    const int km = k - 1 + (k == 1), kp = k - 1 - (k == erows - 2);
                    for (int l = j - range; l <= j + range; l++) {
                        const int lm = l - 1 + (l == 1), lp = l - 1 - (l == ecols - 2);
                        const int kv = (variant & 4) ? km : k - 1, lv = (variant & 4) ? lm : l - 1;
                        const int kg = km, lg = lm;
                        if (!(k > 0 && l > 0 && k < erows - 1 && l < ecols - 1)) continue;
                        if (f[k * ecols + l] == INSIDE || (l - j) * (l - j) + (k - i) * (k - i) > range * range) continue;
                        const float ry = (float)(i - k), rx = (float)(j - l);
                        const float rl = rx * rx + ry * ry;
                        const float dst = 1 / (rl * rl + 1);
                        float gx, gy;
                        if (f[(k + 1) * ecols + l] != INSIDE) {
                            if (f[(k - 1) * ecols + l] != INSIDE)
                                gx = (float)(abs(OUT(kp + 1, lg, color) - OUT(kp, lg, color)) + abs(OUT(kp, lg, color) - OUT(km - 1, lg, color)));
                            else
                                gx = (float)(abs(OUT(kp + 1, lg, color) - OUT(kp, lg, color))) * 2.0f;
                        } else {
                            if (f[(k - 1) * ecols + l] != INSIDE)
                                gx = (float)(abs(OUT(kp, lg, color) - OUT(km - 1, lg, color))) * 2.0f;
                            else
                                gx = 0;
                        }
                        if (f[k * ecols + l + 1] != INSIDE) {
                            if (f[k * ecols + l - 1] != INSIDE)
                                gy = (float)(abs(OUT(kg, lp + 1, color) - OUT(kg, lm, color)) + abs(OUT(kg, lm, color) - OUT(kg, lm - 1, color)));
                            else
                                gy = (float)(abs(OUT(kg, lp + 1, color) - OUT(kg, lm, color))) * 2.0f;
                        } else {
                            if (f[k * ecols + l - 1] != INSIDE)
                                gy = (float)(abs(OUT(kg, lm, color) - OUT(kg, lm - 1, color))) * 2.0f;
                            else
                                gy = 0;
                        }
                        gx = -gx;
                        float dir = rx * gx + ry * gy;
                        if (fabs(dir) <= 0.01) dir = 0.000001f;
                        else if (variant & 2) dir = (float)fabs((double)dir / sqrt((double)(rl * (gx * gx + gy * gy))));
                        else dir = fabsf(dir / sqrtf(rl * (gx * gx + gy * gy)));
                        const float w = dst * dir;
                        Ia += w * (float)OUT(kv, lv, color);
                        s += w;
                    }
                }
                double v = (double)Ia / s;
                long r = lrint(v);
                out[((size_t)(i - 1) * cols + (j - 1)) * 3 + color] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
            }
            f[i * ecols + j] = BAND;
            heap_push(&h, i, j, dist);
        }
    }
    free(h.e); free(f); free(m); free(t);
    return 0;
}
