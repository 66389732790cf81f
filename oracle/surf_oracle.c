/* oracle/surf_oracle.c -- TEST INFRASTRUCTURE ONLY (see surf_oracle.h).
 *
 * Scalar CPU restatement of the reference hot path, one function per stage, each citing the
 * reference file:line it follows. Floating-point expressions use explicit fmaf() where the
 * sm_100a SASS of the reference build (oracle/_ref) shows a contracted FFMA, so that results are
 * bit-identical for the integer stages and the Hessian maps and within a few ulp elsewhere.
 * Built with -ffp-contract=off (no implicit contraction).
 *
 * Parity status: PINNED -- checked in tests/test_oracle_golden.py against tests/golden/*.npz,
 * which hold outputs of the unmodified reference run on a B200 (tests/golden/make_golden.py).
 */
#include "surf_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define R255 0.003921568627f
#define NBIN 72                        /* surfd.h:11 */
#define WINDOW 1.0471975511965976f     /* surfd.h:12 */
#define SEP_ANGLE 0.08726646259971647f /* surfd.h:13 */
#define HWN 6                          /* surfd.h:14 */
#define ORADIUS 9                      /* surfd.h:15 */
#define ORADIUSSQ 81.5f                /* surfd.h:16 */
#define H_PI_F 1.5707963267948966f     /* cuda_utils.h:8 */
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static int align_up(int a, int b) { return (a % b) ? a - a % b + b : a; } /* cuda_utils.h:160-163 */
static int f2i_rn(float x) { return (int)lrintf(x); }                     /* __float2int_rn */
static int f2i_rz(float x) { return (int)x; }                             /* __float2int_rz */

/* ------------------------------------------------------------------ parameters / schedule */

/* surf.cpp:60-80 */
void or_make_params(or_params* p, int noctaves, float thresh, int doubled, int init_mask_size, int sampling_step,
                    int upright, int extend, int desc_wsz) {
    p->doubled = doubled;
    p->noctaves = noctaves;
    p->divisor = doubled ? 0.5f : 1.f;
    p->init_lobe = init_mask_size / 3;
    p->max_scale = p->init_lobe + 2;
    p->sampling = sampling_step + (doubled ? sampling_step : 0);
    p->thresh = thresh;
    p->upright = upright;
    p->extend = extend;
    p->desc_wsz = desc_wsz;
    p->mag_factor = 12 / desc_wsz;
    p->orient_size = 4 + (extend ? 4 : 0);
    p->nfeatures = desc_wsz * desc_wsz * p->orient_size;
}

/* Geometry: surf.cpp:374-390. Octave loop: surf.cpp:240-294. Per-layer parameters:
 * surfd.cu:2844-2865. NMS borders: surfd.cu:3062-3073. Returns 0, or -1 for unsupported. */
int or_make_schedule(const or_params* p, int w, int h, or_octave* sched) {
    /* doubled: w, h are the dimensions of the 2x image (2*W-2, 2*H-2), see or_upsample2x */
    if (p->max_scale > OR_MAX_SCALE || p->noctaves > OR_MAX_OCTAVE) return -1;
    int iw = w + 1, ih = h + 1;
    int sw = (iw - 1) / p->sampling, sh = (ih - 1) / p->sampling;
    int mask = p->init_lobe - 2, octave = 1, s = 0, border1 = 0;
    int borders[OR_MAX_SCALE] = {0};
    for (int o = 0; o < p->noctaves; o++) {
        or_octave* q = &sched[o];
        memset(q, 0, sizeof(*q));
        q->octave = octave;
        q->sw = sw;
        q->sh = sh;
        if (o > 0) {
            border1 = ((3 * (mask + 4 * octave)) / 2) / (p->sampling * octave) + 1;
            borders[0] = border1;
            borders[1] = border1;
            s = 2;
        } else {
            border1 = ((3 * (mask + 6 * octave)) / 2) / (p->sampling * octave) + 1;
        }
        q->s0 = s;
        q->nl = p->max_scale - s;
        int init_mask = mask;
        for (int i = 0, ss = s; ss < p->max_scale; i++, ss++) {
            borders[ss] = border1; /* stored before the update: the lag of SURVEY.md 2.4-3 */
            q->delta[i] = p->sampling * octave;
            q->l[i] = init_mask + 2 * octave * (i + 1);
            if (ss > 2) border1 = 3 * q->l[i] / 2 / q->delta[i] + 1;
            q->b1[i] = border1;
            float n = 9.f / (float)(q->l[i] * q->l[i]);
            q->norm[i] = n * n;
            mask = q->l[i];
        }
        for (int k = 0; k < p->max_scale; k++) q->borders[k] = borders[k];
        q->nmb = 0;
        for (int k = 1; k < p->max_scale - 1; k += 2) q->mb[q->nmb++] = borders[k + 1] + 1;
        octave += octave;
        sw >>= 1;
        sh >>= 1;
    }
    return 0;
}

long long or_resp_floats(const or_params* p, const or_octave* sched) {
    long long t = 0;
    for (int o = 0; o < p->noctaves; o++) t += (long long)sched[o].sw * sched[o].sh * p->max_scale;
    return t;
}

/* ------------------------------------------------------------------ integral image */

/* surfd.cu:129-165 (integralRow then integralCol): out[y+1][x+1] = sum_{j<=y,i<=x} img[j][i],
 * row 0 and column 0 zero. */
void or_integral(const uint8_t* img, int w, int h, int pitch, int32_t* out) {
    const int iw = w + 1;
    memset(out, 0, sizeof(int32_t) * (size_t)iw);
    for (int y = 0; y < h; y++) {
        int32_t* row = out + (size_t)(y + 1) * iw;
        const int32_t* up = out + (size_t)y * iw;
        const uint8_t* src = img + (size_t)y * pitch;
        int32_t run = 0;
        row[0] = 0;
        for (int x = 0; x < w; x++) {
            run += src[x];
            row[x + 1] = run + up[x + 1];
        }
    }
}

/* doubled=true (surf.cpp:234-235): the integral image is that of a 2x bilinear up-sampling of the frame.
 * integralDoubleRow0U2 (surfd.cu:166-207) defines the up-sampled pixels -- even/even: the source pixel;
 * even row, odd column: rn((a+b)*0.5f) of the horizontal neighbours; odd row, even column: the same of
 * the vertical neighbours; odd/odd: rn((a+b+c+d)*0.25f) -- and the scan kernels (surfd.cu:209-318,
 * 2707-2772) sum columns 1..2W-2 and rows 1..2H-2 of the padded layout, i.e. the 2x image is
 * (2W-2) x (2H-2): its last row / column would need source pixels outside the frame (the reference reads
 * and writes out of bounds there, outside the region every later stage uses). */
void or_upsample2x(const uint8_t* img, int w, int h, int pitch, uint8_t* out) {
    const int W2 = 2 * w - 2, H2 = 2 * h - 2;
    for (int Y = 0; Y < H2; Y++) {
        const uint8_t* r0 = img + (size_t)(Y >> 1) * pitch;
        const uint8_t* r1 = r0 + pitch;
        uint8_t* d = out + (size_t)Y * W2;
        for (int X = 0; X < W2; X++) {
            const int x = X >> 1;
            int v;
            if (!(Y & 1)) v = (X & 1) ? f2i_rn((float)(r0[x] + r0[x + 1]) * 0.5f) : r0[x];
            else v = (X & 1) ? f2i_rn((float)(r0[x] + r0[x + 1] + r1[x] + r1[x + 1]) * 0.25f) : f2i_rn((float)(r0[x] + r1[x]) * 0.5f);
            d[X] = (uint8_t)v;
        }
    }
}

/* out: tight (2h-1) x (2w-1) */
void or_integral_doubled(const uint8_t* img, int w, int h, int pitch, int32_t* out) {
    const int W2 = 2 * w - 2, H2 = 2 * h - 2;
    uint8_t* up = (uint8_t*)malloc((size_t)W2 * H2);
    or_upsample2x(img, w, h, pitch, up);
    or_integral(up, W2, H2, W2, out);
    free(up);
}

/* The reference indexes the integral with pitch iAlignUp(w+1,128) and can stray a few elements
 * outside the image in getTrace (SURVEY.md section 7 "hard parts"): the padded view reproduces
 * what it reads there (zero padding / neighbouring rows / a zero guard row). */
typedef struct {
    int32_t* base;
    const int32_t* d; /* row 0 */
    int iw, ih, p;
} iview;

static iview iview_make(const int32_t* tight, int w, int h) {
    iview v;
    v.iw = w + 1;
    v.ih = h + 1;
    v.p = align_up(v.iw, 128);
    v.base = (int32_t*)calloc((size_t)v.p * (v.ih + 2), sizeof(int32_t));
    int32_t* d = v.base + v.p;
    for (int y = 0; y < v.ih; y++) memcpy(d + (size_t)y * v.p, tight + (size_t)y * v.iw, sizeof(int32_t) * v.iw);
    v.d = d;
    return v;
}
static void iview_free(iview* v) { free(v->base); }

/* surfd.cu:334-343: inclusive pixel-box sum over x in [x2,x1], y in [y2,y1]. */
static inline int get_sum(const iview* I, int x1, int y1, int x2, int y2) {
    const int p = I->p;
    const int yp1 = y1 * p + p, yp2 = y2 * p;
    return I->d[yp1 + x1 + 1] + I->d[yp2 + x2] - I->d[yp2 + x1 + 1] - I->d[yp1 + x2];
}

/* ------------------------------------------------------------------ Hessian */

/* surfd.cu:353-366 with the vas[] of surfd.cu:461-476. SASS order:
 * t=(float(Dxy)*0.6f)^2 ; det=fma(Dxx,Dyy,-t) ; det*=r*r ; out=det*norm. */
static inline float hessian_at(const iview* I, int cx, int cy, int l, float norm) {
    const int x2 = l / 2, x3 = x2 + x2, x4 = x2 + x3;
    const int dxx = get_sum(I, cx + l + x2, cy + x3, cx - l - x2, cy - x3) - 3 * get_sum(I, cx + x2, cy + x3, cx - x2, cy - x3);
    const int dyy = get_sum(I, cx + x3, cy + l + x2, cx - x3, cy - l - x2) - 3 * get_sum(I, cx + x3, cy + x2, cx - x3, cy - x2);
    const int dxy = get_sum(I, cx + x4, cy, cx, cy - x4) + get_sum(I, cx, cy + x4, cx - x4, cy) -
                    get_sum(I, cx + x4, cy + x4, cx, cy) - get_sum(I, cx, cy, cx - x4, cy - x4);
    const float fxy = 0.6f * (float)dxy;
    const float t = fxy * fxy;
    float det = fmaf((float)dxx, (float)dyy, -t);
    const float rr = R255 * R255;
    det = rr * det;
    return det * norm;
}

/* surfd.cu:445-481 (calcHessianMultiConst) per octave + surf.cpp:250-258 / surfd.cu:321-331
 * (halfImage of layers max_scale-3 and max_scale-1 into layers 0,1 of the next octave). */
void or_hessian(const or_params* p, const or_octave* sched, const int32_t* integral, int w, int h, float* resp) {
    iview I = iview_make(integral, w, h);
    const long long total = or_resp_floats(p, sched);
    memset(resp, 0, sizeof(float) * (size_t)total);
    float* oct = resp;
    const float* prev = NULL;
    for (int o = 0; o < p->noctaves; o++) {
        const or_octave* q = &sched[o];
        const size_t osz = (size_t)q->sw * q->sh;
        if (o > 0) {
            const or_octave* pq = &sched[o - 1];
            const size_t posz = (size_t)pq->sw * pq->sh;
            const float* s0 = prev + (size_t)(p->max_scale - 3) * posz;
            const float* s1 = prev + (size_t)(p->max_scale - 1) * posz;
            for (int y = 0; y < q->sh; y++)
                for (int x = 0; x < q->sw; x++) {
                    oct[(size_t)y * q->sw + x] = s0[(size_t)(2 * y) * pq->sw + 2 * x];
                    oct[osz + (size_t)y * q->sw + x] = s1[(size_t)(2 * y) * pq->sw + 2 * x];
                }
        }
        for (int i = 0; i < q->nl; i++) {
            float* dst = oct + (size_t)(q->s0 + i) * osz;
            const int b = q->b1[i], d = q->delta[i];
            for (int iy = b; iy < q->sh - b; iy++)
                for (int ix = b; ix < q->sw - b; ix++)
                    dst[(size_t)iy * q->sw + ix] = hessian_at(&I, d * ix, d * iy, q->l[i], q->norm[i]);
        }
        prev = oct;
        oct += (size_t)p->max_scale * osz;
    }
    iview_free(&I);
}

/* ------------------------------------------------------------------ NMS + refinement */

/* surfd.cu:835-887: 3x3 Gaussian elimination with partial pivoting; `a -= m*b` is an FFMA in the
 * reference SASS. */
static void solve3(float* sol, float sq[3][3]) {
    const int size = 3;
    int pivot = 0;
    for (int col = 0; col < size - 1; col++) {
        float maxc = -1.f;
        for (int row = col; row < size; row++) {
            float coef = sq[row][col];
            coef = (coef < 0.f ? -coef : coef);
            if (coef > maxc) { maxc = coef; pivot = row; }
        }
        if (pivot != col) {
            for (int i = 0; i < size; i++) { float t = sq[pivot][i]; sq[pivot][i] = sq[col][i]; sq[col][i] = t; }
            float t = sol[pivot]; sol[pivot] = sol[col]; sol[col] = t;
        }
        for (int row = col + 1; row < size; row++) {
            const float mult = sq[row][col] / sq[col][col];
            for (int c = col; c < size; c++) sq[row][c] = fmaf(-mult, sq[col][c], sq[row][c]);
            sol[row] = fmaf(-mult, sol[col], sol[row]);
        }
    }
    for (int row = size - 1; row >= 0; row--) {
        float val = sol[row];
        for (int col = size - 1; col > row; col--) val = fmaf(-sol[col], sq[row][col], val);
        sol[row] = val / sq[row][row];
    }
}

/* surfd.cu:942-988 */
static float fit_quadrat(const float* src, int sw, size_t osz, int s, int r, int c, float* off) {
    const float* cur = src + (size_t)s * osz;
    const float* prv = cur - osz;
    const float* nxt = cur + osz;
    const size_t idx = (size_t)r * sw + c, nr = idx + sw, pr = idx - sw, nc = idx + 1, pc = idx - 1;
    float g[3], H[3][3];
    g[0] = (nxt[idx] - prv[idx]) * 0.5f;
    g[1] = (cur[nr] - cur[pr]) * 0.5f;
    g[2] = (cur[nc] - cur[pc]) * 0.5f;
    const float temp = cur[idx] + cur[idx];
    H[0][0] = prv[idx] + nxt[idx] - temp;
    H[1][1] = cur[nr] + cur[pr] - temp;
    H[2][2] = cur[nc] + cur[pc] - temp;
    H[0][1] = ((nxt[nr] - nxt[pr]) - (prv[nr] - prv[pr])) * 0.25f;
    H[0][2] = ((nxt[nc] - nxt[pc]) - (prv[nc] - prv[pc])) * 0.25f;
    H[1][2] = ((cur[nr + 1] - cur[nr - 1]) - (cur[pr + 1] - cur[pr - 1])) * 0.25f;
    H[1][0] = H[0][1];
    H[2][0] = H[0][2];
    H[2][1] = H[1][2];
    off[0] = -g[0];
    off[1] = -g[1];
    off[2] = -g[2];
    solve3(off, H);
    /* SASS: t = o1*g1 ; t = fma(o0,g0,t) ; t = fma(o2,g2,t) ; strength = fma(t,0.5,c) */
    float t = off[1] * g[1];
    t = fmaf(off[0], g[0], t);
    t = fmaf(off[2], g[2], t);
    return fmaf(t, 0.5f, cur[idx]);
}

/* surfd.cu:369-377 */
static int get_trace(const iview* I, int cx, int cy, int l) {
    const int x2 = l / 2, x3 = x2 + x2;
    const int lxx = get_sum(I, cx + l + x2, cy + x3, cx - l - x2, cy - x3) - 3 * get_sum(I, cx + x2, cy + x3, cx - x2, cy - x3);
    const int lyy = get_sum(I, cx + x3, cy + l + x2, cx - x3, cy - l - x2) - 3 * get_sum(I, cx + x3, cy + x2, cx - x3, cy - x2);
    return (lxx + lyy > 0 ? 1 : -1);
}

/* surfd.cu:676-832 (findMaximumWithInterp) + surfd.cu:1001-1022 (makePoint). */
int or_find_keypoints(const or_params* p, const or_octave* sched, const int32_t* integral, int w, int h,
                      const float* resp, or_point* pts, int max_pts) {
    iview I = iview_make(integral, w, h);
    int n = 0;
    const float* src = resp;
    const int ms = p->max_scale;
    for (int o = 0; o < p->noctaves; o++) {
        const or_octave* q = &sched[o];
        const int sw = q->sw, sh = q->sh;
        const size_t osz = (size_t)sw * sh;
        for (int z = 0; z < q->nmb; z++) {
            const int k = 2 * z + 1;
            if (k >= ms - 1) continue;
            const int mb = q->mb[z];
            for (int i = mb; i < sh - mb; i += 2)
                for (int j = mb; j < sw - mb; j += 2) {
                    const float* cs = src + (size_t)k * osz;
                    const size_t iw_ = (size_t)i * sw + j, ix_ = iw_ + 1, iy_ = iw_ + sw, iz_ = iy_ + 1;
                    int cas = 0;
                    float best = cs[iw_];
                    if (cs[ix_] > best) { best = cs[ix_]; cas = 1; }
                    if (cs[iy_] > best) { best = cs[iy_]; cas = 2; }
                    if (cs[iz_] > best) { best = cs[iz_]; cas = 3; }
                    cs += osz;
                    if (cs[iw_] > best) { best = cs[iw_]; cas = 4; }
                    if (cs[ix_] > best) { best = cs[ix_]; cas = 5; }
                    if (cs[iy_] > best) { best = cs[iy_]; cas = 6; }
                    if (cs[iz_] > best) { best = cs[iz_]; cas = 7; }
                    if (best < p->thresh * 0.8f || (k + 1 == ms - 1 && cas > 3)) continue;
                    int s = k + (cas >> 2), r = i + ((cas >> 1) & 1), c = j + (cas & 1);
                    const int ds = (cas & 4) ? 1 : -1, dr = (cas & 2) ? 1 : -1, dc = (cas & 1) ? 1 : -1;
                    /* the 19 neighbours outside the cell (surfd.cu:757-792): reject if best < nb */
                    int ok = 1;
                    for (int a = -1; a <= 1 && ok; a++)
                        for (int b = -1; b <= 1 && ok; b++)
                            for (int e = -1; e <= 1 && ok; e++) {
                                /* inside the 2x2x2 cell <=> each of (a,b,e) is 0 or minus the outward direction */
                                const int in_cell = (a == 0 || a == -ds) && (b == 0 || b == -dr) && (e == 0 || e == -dc);
                                if (in_cell) continue;
                                const float nb = src[(size_t)(s + a) * osz + (size_t)(r + b) * sw + (c + e)];
                                if (best < nb) ok = 0;
                            }
                    if (!ok) continue;
                    float off[3] = {0, 0, 0};
                    float strength = 0;
                    int newr = r, newc = c;
                    for (int mv = 0; mv < 5; mv++) {
                        r = newr;
                        c = newc;
                        strength = fit_quadrat(src, sw, osz, s, r, c, off);
                        if (off[1] > 0.6f && r < sh - q->borders[s]) newr++;
                        if (off[1] < -0.6f && r > q->borders[s]) newr--;
                        if (off[2] > 0.6f && c < sw - q->borders[s]) newc++;
                        if (off[2] < -0.6f && c > q->borders[s]) newc--;
                        if (newr == r && newc == c) break;
                    }
                    if (isnan(off[0]) || isnan(off[1]) || isnan(off[2]) || fabsf(off[0]) > 1.5f || fabsf(off[1]) > 1.5f ||
                        fabsf(off[2]) > 1.5f || strength < p->thresh)
                        continue;
                    const int octave = q->octave;
                    /* SASS: ns = fma((s+off0)*2, octave, float(init_lobe+(octave-1)*max_scale)) / 3 */
                    float t = (float)s + off[0];
                    t = t + t;
                    const float ns = fmaf(t, (float)octave, (float)(p->init_lobe + (octave - 1) * ms)) / 3.f;
                    const float ny = (float)octave * ((float)r + off[1]);
                    const float nx = (float)octave * ((float)c + off[2]);
                    if (n < max_pts) {
                        or_point* P = &pts[n++];
                        memset(P, 0, sizeof(*P));
                        const float td = (float)p->sampling * p->divisor;
                        P->x = nx * td;
                        P->y = ny * td;
                        P->scale = 1.2f * ns * p->divisor;
                        P->o = o;
                        P->strength = strength;
                        P->ori = 0.f;
                        P->match = -1;
                        const int L = f2i_rz(fmaf(3.f, ns, 0.5f));
                        const int px = f2i_rz(fmaf(nx, (float)p->sampling, 0.5f));
                        const int py = f2i_rz(fmaf(ny, (float)p->sampling, 0.5f));
                        P->laplace = get_trace(&I, px, py, L);
                    }
                }
        }
        src += (size_t)ms * osz;
    }
    iview_free(&I);
    return n;
}

/* ------------------------------------------------------------------ Haar wavelets */

/* surfd.cu:1171-1175: upper half minus lower half */
static inline int wavelet1(const iview* I, int x, int y, int s) {
    return get_sum(I, x + s, y, x - s, y - s) - get_sum(I, x + s, y + s, x - s, y);
}
/* surfd.cu:1178-1182: right half minus left half */
static inline int wavelet2(const iview* I, int x, int y, int s) {
    return get_sum(I, x + s, y + s, x, y - s) - get_sum(I, x, y + s, x - s, y - s);
}

/* ------------------------------------------------------------------ orientation */

/* surfd.cu:114-126; H_PI is a float macro, M_PI the double one. */
static float fast_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float a = fminf(ax, ay) / fmaxf(ax, ay);
    const float s = a * a;
    float r = fmaf(fmaf(fmaf(-0.0464964749f, s, 0.15931422f), s, -0.327622764f), s * a, a);
    r = (ay > ax ? H_PI_F - r : r);
    r = (float)(x < 0 ? M_PI - (double)r : (double)r);
    r = (y < 0 ? -r : r);
    return r;
}

/* surfd.cu:1711-1960 (assignOrientationApprox); bins[] per surf.cpp:83-90. Shared-memory atomics
 * of the reference are replaced by the loop order below (float sums differ in the last ulps). */
void or_orientation(const or_params* p, const int32_t* integral, int w, int h, or_point* pts, int n) {
    iview I = iview_make(integral, w, h);
    float bins[NBIN];
    bins[0] = (float)(-M_PI);
    for (int i = 1; i < NBIN; i++) bins[i] = bins[i - 1] + SEP_ANGLE;
    const float lut1_den = 12.5f;
#define PASZ (NBIN + HWN + HWN)
    for (int pi = 0; pi < n; pi++) {
        or_point* P = &pts[pi];
        float x = P->x, y = P->y, scale = P->scale;
        if (p->doubled) { x = x + x; y = y + y; scale = scale + scale; }
        const int hs = f2i_rz(fmaf(2.f, scale, 1.6f));
        const int st = f2i_rz(scale + 0.8f);
        const int ix = f2i_rn(x), iy = f2i_rn(y);
        int hist[NBIN] = {0};
        float avg[NBIN] = {0}, ps[NBIN] = {0}, pas[PASZ] = {0}, ws[NBIN] = {0}, was[NBIN] = {0};
        for (int y1 = -ORADIUS; y1 <= ORADIUS; y1++)
            for (int x1 = -ORADIUS; x1 <= ORADIUS; x1++) {
                const int xx = ix + x1 * st, yy = iy + y1 * st;
                if (!(yy + hs + 2 < I.ih && yy - hs > -1 && xx + hs + 2 < I.iw && xx - hs > -1)) continue;
                const int distsq = y1 * y1 + x1 * x1;
                if (!((float)distsq < ORADIUSSQ)) continue;
                const float dx = (float)wavelet2(&I, xx, yy, hs) * R255;
                const float dy = (float)wavelet1(&I, xx, yy, hs) * R255;
                const float mag = sqrtf(fmaf(dx, dx, dy * dy));
                if (!(mag > 0.f)) continue;
                const float weight = expf(-((float)distsq + 0.5f) / lut1_den); /* lookup1, surf.cpp:360-363 */
                const float angle = fast_atan2(dy, dx);
                const int hid = f2i_rz((float)(((double)angle + M_PI) / (double)SEP_ANGLE)) % NBIN;
                const float psum = weight * mag;
                hist[hid] += 1;
                avg[hid] += angle;
                ps[hid] += psum;
                pas[hid + HWN] += angle * psum;
                if (hid - HWN < 0) pas[hid + HWN + NBIN] += (float)(((double)angle + 2 * M_PI) * (double)psum);
                else if (hid + HWN >= NBIN) pas[hid + HWN - NBIN] += (float)(((double)angle - 2 * M_PI) * (double)psum);
            }
        for (int i = 0; i < NBIN; i++) avg[i] = hist[i] > 0 ? avg[i] / (float)hist[i] : bins[i];
        for (int i = 0; i < NBIN; i++)
            for (int j = -HWN; j <= HWN; j++) {
                int k = i + j;
                if (j == -HWN) {
                    float res;
                    if (k < 0) {
                        k += NBIN;
                        const int k1 = (k + 1) % NBIN;
                        res = (float)((double)(bins[k1] + (WINDOW / 2) - avg[i]) - (bins[k1] < 0 ? 0.0 : 2 * M_PI));
                    } else {
                        res = bins[k + 1] + (WINDOW / 2) - avg[i];
                    }
                    const float er = res / SEP_ANGLE;
                    ws[i] += er * ps[k];
                    was[i] += er * pas[i];
                } else if (j == HWN) {
                    float res;
                    if (k >= NBIN) {
                        k -= NBIN;
                        res = (float)((double)(avg[i] + (WINDOW / 2)) - 2 * M_PI - (double)bins[k]);
                    } else {
                        res = avg[i] + (WINDOW / 2) - bins[k];
                    }
                    const float er = res / SEP_ANGLE;
                    ws[i] += er * ps[k];
                    was[i] += er * pas[i + HWN + HWN];
                } else {
                    was[i] += pas[k + HWN];
                    if (k < 0) k += NBIN;
                    else if (k >= NBIN) k -= NBIN;
                    ws[i] += ps[k];
                }
            }
        /* tournament max of surfd.cu:1921-1947: 64-wide tree, then 8-wide tree at offset 64 */
        int residual = NBIN, offset = 0;
        while (residual > 0) {
            int zn = 1;
            while (zn * 2 <= residual) zn *= 2;
            for (int stride = zn / 2; stride > 0; stride >>= 1)
                for (int tid = 0; tid < stride; tid++) {
                    const int id1 = tid + offset, id2 = id1 + stride;
                    if (ws[id1] < ws[id2]) { ws[id1] = ws[id2]; was[id1] = was[id2]; }
                }
            if (ws[0] < ws[offset]) { ws[0] = ws[offset]; was[0] = was[offset]; }
            residual -= zn;
            offset += zn;
        }
        P->ori = was[0] / ws[0];
    }
    iview_free(&I);
}

/* ------------------------------------------------------------------ descriptors */

/* surfd.cu:1199-1271 */
static void place_in_index(float* d, const or_params* p, float mag1, int ori1, float mag2, int ori2, float rx, float cx) {
    const int ri = f2i_rz(rx >= 0.f ? rx : rx - 1.f);
    const int ci = f2i_rz(cx >= 0.f ? cx : cx - 1.f);
    const float rfrac = rx - (float)ri, cfrac = cx - (float)ci, cfrac1 = 1.f - cfrac;
    const int W = p->desc_wsz, O = p->orient_size;
    for (int dr = 0; dr < 2; dr++) {
        const int r = ri + dr;
        if (dr == 0 ? (r < 0) : (r >= W)) continue;
        const float rw1 = dr == 0 ? mag1 * (1.f - rfrac) : mag1 * rfrac;
        const float rw2 = dr == 0 ? mag2 * (1.f - rfrac) : mag2 * rfrac;
        if (ci >= 0) {
            const int os = r * W * O + ci * O;
            d[os + ori1] += rw1 * cfrac1;
            d[os + ori2] += rw2 * cfrac1;
        }
        if (ci + 1 < W) {
            const int os = r * W * O + (ci + 1) * O;
            d[os + ori1] += rw1 * cfrac;
            d[os + ori2] += rw2 * cfrac;
        }
    }
}

/* surfd.cu:1566-1615 + 1288-1317 (upright) and surfd.cu:2391-2444 + 1984-2015 (rotated);
 * normalisation surfd.cu:2447-2493 (tree order of the 64-wide block). */
void or_describe(const or_params* p, const int32_t* integral, int w, int h, const or_point* pts, int n, float* desc,
                 int normalise) {
    iview I = iview_make(integral, w, h);
    const int W = p->desc_wsz, NF = p->nfeatures;
    float lut2[40];
    for (int k = 0; k < 40; k++) lut2[k] = expf(-((float)k + 0.5f) / 8.f); /* surf.cpp:366-369 */
    for (int pi = 0; pi < n; pi++) {
        const or_point* P = &pts[pi];
        float* d = desc + (size_t)pi * NF;
        memset(d, 0, sizeof(float) * NF);
        float x = P->x, y = P->y, scale;
        if (p->doubled) { x = x + x; y = y + y; scale = 3.3f * P->scale; }
        else scale = 1.65f * P->scale;
        const int step = f2i_rn(scale * 0.5f) > 1 ? f2i_rn(scale * 0.5f) : 1;
        const int ix = f2i_rn(x), iy = f2i_rn(y);
        const float fx = x - (float)ix, fy = y - (float)iy;
        const float spacing = scale * (float)p->mag_factor;
        const int S = f2i_rz(scale);
        const float wofs = fmaf((float)W, 0.5f, -0.5f);
        float sine = 0.f, cose = 1.f, fracr = fy, fracc = fx;
        int R;
        if (p->upright) {
            R = f2i_rn(spacing * (float)(W + 1) * 0.5f / (float)step);
        } else {
            sine = sinf(P->ori); /* __sinf/__cosf in the reference: ~1e-6 abs */
            cose = cosf(P->ori);
            fracc = fmaf(-sine, fy, cose * fx); /* nvcc: a*b + c*d -> fma(a,b,c*d) */
            fracr = fmaf(cose, fy, sine * fx);
            R = f2i_rn(1.4f * spacing * (float)(W + 1) * 0.5f / (float)step);
        }
        for (int i = -R; i <= R; i++)
            for (int j = -R; j <= R; j++) {
                float rpos, cpos;
                if (p->upright) {
                    rpos = ((float)(step * i) - fy) / spacing;
                    cpos = ((float)(step * j) - fx) / spacing;
                } else {
                    rpos = fmaf((float)step, fmaf(cose, (float)i, sine * (float)j), -fracr) / spacing;
                    cpos = fmaf((float)step, fmaf(-sine, (float)i, cose * (float)j), -fracc) / spacing;
                }
                const float rx = rpos + wofs, cx = cpos + wofs;
                if (!(rx > -1.f && rx < (float)W && cx > -1.f && cx < (float)W)) continue;
                const int r = iy + i * step, c = ix + j * step;
                if (!(r >= 1 + S && r < I.ih - 1 - S && c >= 1 + S && c < I.iw - 1 - S)) continue;
                const float weight = lut2[f2i_rz(fmaf(rpos, rpos, cpos * cpos))];
                const float a = weight * (float)wavelet2(&I, c, r, S) * R255;
                const float b = weight * (float)wavelet1(&I, c, r, S) * R255;
                float dx = a, dy = b;
                if (!p->upright) {
                    dx = fmaf(cose, a, sine * b);
                    dy = fmaf(sine, a, -(cose * b));
                }
                if (!p->extend) {
                    place_in_index(d, p, dx, (dx < 0 ? 0 : 1), dy, (dy < 0 ? 2 : 3), rx, cx);
                } else {
                    place_in_index(d, p, dx, (dy < 0 ? 0 : 1), fabsf(dx), (dy < 0 ? 2 : 3), rx, cx);
                    place_in_index(d, p, dy, (dx < 0 ? 4 : 5), fabsf(dy), (dx < 0 ? 6 : 7), rx, cx);
                }
            }
        if (normalise) {
            /* surfd.cu:2460-2492. The reference's tree (halve down to 32, then +32,+16,...,+1) is only defined for
             * nfeatures 64 and 128: for desc_wsz < 4 it reads shared memory past the nfeatures floats it allocated
             * (:2474-2480 with 16 / 36 / 32 / 72 elements). There the intent -- the plain sum of squares -- is
             * restated, in element order. */
            float sq[256], total;
            for (int t = 0; t < NF; t++) sq[t] = d[t] * d[t];
            if (NF == 64 || NF == 128) {
                for (int stride = NF / 2; stride > 0; stride >>= 1)
                    for (int t = 0; t < stride; t++) sq[t] += sq[t + stride];
                total = sq[0];
            } else {
                total = 0.f;
                for (int t = 0; t < NF; t++) total += sq[t];
            }
            const float f = 1.f / sqrtf(total);
            for (int t = 0; t < NF; t++) d[t] *= f;
        }
    }
    iview_free(&I);
}

/* ------------------------------------------------------------------ matching */

/* surfd.cu:2535-2671 (findMaxCorr). Candidates: p2 < n2 - n2%32 (tail ignored, :2569). Dot
 * product is a sequential FFMA chain over d=0..nf-1 (:2591-2609). Group g=(p2%32)/4 keeps an
 * exact running top-2 in increasing p2 (:2610-2625); the merge (:2646-2664) starts from group 0
 * and only sees the maxima of the other groups. */
void or_match(or_point* pts1, int n1, const float* f1, const or_point* pts2, int n2, const float* f2, int nf) {
    const int ncand = n2 - n2 % 32;
#pragma omp parallel for schedule(static)
    for (int p1 = 0; p1 < n1; p1++) {
        float mx[8], sc[8];
        int id[8];
        for (int g = 0; g < 8; g++) { mx[g] = 0.f; sc[g] = 0.f; id[g] = -1; }
        const float* a = f1 + (size_t)p1 * nf;
        for (int p2 = 0; p2 < ncand; p2++) {
            const float* b = f2 + (size_t)p2 * nf;
            float s = 0.f;
            for (int d = 0; d < nf; d++) s = fmaf(a[d], b[d], s);
            const int g = (p2 & 31) >> 2;
            if (s > mx[g]) { sc[g] = mx[g]; mx[g] = s; id[g] = p2; }
            else if (s > sc[g]) sc[g] = s;
        }
        float m = mx[0], s2 = sc[0];
        int idx = id[0];
        for (int g = 0; g < 8; g++) {
            if (idx != id[g]) {
                if (mx[g] > m) { s2 = fmaxf(m, s2); m = mx[g]; idx = id[g]; }
                else if (mx[g] > s2) s2 = mx[g];
            }
        }
        pts1[p1].score = m;
        pts1[p1].match = idx;
        if (idx >= 0) { pts1[p1].match_x = pts2[idx].x; pts1[p1].match_y = pts2[idx].y; }
        pts1[p1].ambiguity = s2 / (m + 1e-6f);
    }
}

/* ------------------------------------------------------------------ whole frame */

int or_detect_and_compute(const or_params* p, const uint8_t* img, int w, int h, int pitch, or_point* pts, int max_pts,
                          float* desc) {
    or_octave sched[OR_MAX_OCTAVE];
    uint8_t* up = NULL;
    if (p->doubled) { /* everything downstream runs on the 2x image (surf.cpp:234-235) */
        if (w < 2 || h < 2) return -1;
        up = (uint8_t*)malloc((size_t)(2 * w - 2) * (2 * h - 2));
        or_upsample2x(img, w, h, pitch, up);
        img = up; w = 2 * w - 2; h = 2 * h - 2; pitch = w;
    }
    if (or_make_schedule(p, w, h, sched) != 0) { free(up); return -1; }
    int32_t* I = (int32_t*)malloc(sizeof(int32_t) * (size_t)(w + 1) * (h + 1));
    float* resp = (float*)malloc(sizeof(float) * (size_t)or_resp_floats(p, sched));
    or_integral(img, w, h, pitch, I);
    free(up);
    or_hessian(p, sched, I, w, h, resp);
    const int n = or_find_keypoints(p, sched, I, w, h, resp, pts, max_pts);
    if (desc) {
        if (!p->upright) or_orientation(p, I, w, h, pts, n);
        or_describe(p, I, w, h, pts, n, desc, 1);
    }
    free(resp);
    free(I);
    return n;
}

double or_time_frames(const or_params* p, const uint8_t* imgs, long long frame_stride, int nframes, int w, int h,
                      int pitch, int max_pts, int threads, long long* total_pts) {
    long long tot = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : tot)
    for (int f = 0; f < nframes; f++) {
        or_point* pts = (or_point*)malloc(sizeof(or_point) * (size_t)max_pts);
        float* desc = (float*)malloc(sizeof(float) * (size_t)max_pts * p->nfeatures);
        const int n = or_detect_and_compute(p, imgs + (size_t)f * frame_stride, w, h, pitch, pts, max_pts, desc);
        tot += n > 0 ? n : 0;
        free(pts);
        free(desc);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (total_pts) *total_pts = tot;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
