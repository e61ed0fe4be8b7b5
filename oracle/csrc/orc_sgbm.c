/*
 * orc_sgbm.c -- CPU restatement of cv2.StereoSGBM.compute (modes SGBM, HH, SGBM_3WAY, HH4) as the
 * reference calls it (camera/single_usb_stereo_camera.py:252-274 parameters, :324-325 calls,
 * test_improved_laser.py:151, test_depth.py:68).
 *
 * TEST INFRASTRUCTURE ONLY (see l3d_oracle.h).  The algorithm lives in the third-party
 * dependency OpenCV (calib3d StereoSGBM; reference pins only opencv-python>=4.5.0,
 * requirements.txt:7).  This file restates the published algorithm (Hirschmueller SGM with
 * Birchfield-Tomasi cost, OpenCV's fixed-point conventions) and is pinned bit-exactly against the
 * cv2 4.13.0 binary of this image by tests/test_oracle_sgbm.py.
 */
#include "l3d_oracle.h"
#include <stdlib.h>
#include <string.h>

#define MAX_COST 32767
#define DISP_SHIFT 4
#define DISP_SCALE 16

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int sat16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }

/* Pre-filter (x-Sobel clipped to [0,2*ftzero]) + raw intensity, two channels per image row.
 * pre[0][x] = clipped gradient, pre[1][x] = intensity; both = ftzero at x=0 and x=W-1. */
static void prefilter_row(const uint8_t* img, int W, int H, int y, int ftzero, uint8_t* c0,
                          uint8_t* c1) {
    const uint8_t* r = img + (long)y * W;
    const uint8_t* rn = img + (long)(y > 0 ? y - 1 : y) * W;
    const uint8_t* rs = img + (long)(y < H - 1 ? y + 1 : y) * W;
    for (int x = 1; x < W - 1; x++) {
        int g = (r[x + 1] - r[x - 1]) * 2 + rn[x + 1] - rn[x - 1] + rs[x + 1] - rs[x - 1];
        g = imin(imax(g, -ftzero), ftzero) + ftzero;
        c0[x] = (uint8_t)g;
        c1[x] = r[x];
    }
    c0[0] = c0[W - 1] = (uint8_t)ftzero;
    c1[0] = c1[W - 1] = (uint8_t)ftzero;
}

static void halfpix_minmax(const uint8_t* p, int W, uint8_t* lo, uint8_t* hi) {
    for (int x = 0; x < W; x++) {
        int v = p[x];
        int vl = x > 0 ? (v + p[x - 1]) / 2 : v;
        int vr = x < W - 1 ? (v + p[x + 1]) / 2 : v;
        lo[x] = (uint8_t)imin(imin(vl, vr), v);
        hi[x] = (uint8_t)imax(imax(vl, vr), v);
    }
}

typedef struct {
    int W, H, minD, maxD, D, minX1, maxX1, width1, ftzero;
    uint8_t* buf; /* 12*W bytes scratch */
} cost_ctx;

/* Birchfield-Tomasi pixel cost of image row y into cost[width1*D] (int16). */
static void pixel_cost_row(const cost_ctx* c, const uint8_t* img1, const uint8_t* img2, int y,
                           int16_t* cost) {
    int W = c->W, D = c->D;
    uint8_t* l[2] = {c->buf, c->buf + W};
    uint8_t* r[2] = {c->buf + 2 * W, c->buf + 3 * W};
    uint8_t *llo = c->buf + 4 * W, *lhi = c->buf + 5 * W, *rlo = c->buf + 6 * W,
            *rhi = c->buf + 7 * W;
    prefilter_row(img1, W, c->H, y, c->ftzero, l[0], l[1]);
    prefilter_row(img2, W, c->H, y, c->ftzero, r[0], r[1]);
    memset(cost, 0, sizeof(int16_t) * (size_t)c->width1 * D);
    for (int ch = 0; ch < 2; ch++) {
        int shift = ch == 0 ? 0 : 2;
        halfpix_minmax(l[ch], W, llo, lhi);
        halfpix_minmax(r[ch], W, rlo, rhi);
        for (int x = c->minX1; x < c->maxX1; x++) {
            int u = l[ch][x], u0 = llo[x], u1 = lhi[x];
            int16_t* cp = cost + (long)(x - c->minX1) * D;
            for (int d = c->minD; d < c->maxD; d++) {
                int xr = x - d;
                int v = r[ch][xr], v0 = rlo[xr], v1 = rhi[xr];
                int c0 = imax(imax(0, u - v1), v0 - u);
                int c1 = imax(imax(0, v - u1), u0 - v);
                cp[d - c->minD] = (int16_t)(cp[d - c->minD] + (imin(c0, c1) >> shift));
            }
        }
    }
}

/* horizontal box sum (replicate clamp in width1 coordinates), int16 wrap */
static void hsum_row(const int16_t* pd, int width1, int D, int SW2, int16_t* h) {
    for (int d = 0; d < D; d++) {
        int s = pd[d] * (SW2 + 1);
        for (int k = 1; k <= SW2; k++) s += pd[(long)imin(k, width1 - 1) * D + d];
        h[d] = (int16_t)s;
    }
    for (int x = 1; x < width1; x++) {
        const int16_t* add = pd + (long)imin(x + SW2, width1 - 1) * D;
        const int16_t* sub = pd + (long)imax(x - SW2 - 1, 0) * D;
        for (int d = 0; d < D; d++)
            h[(long)x * D + d] = (int16_t)(h[(long)(x - 1) * D + d] + add[d] - sub[d]);
    }
}

/* Full cost volume C[rows y0..y1)[width1][D] with the vertical box sum restarted at y0 (y0=0,
 * y1=H for modes 0/1; stripe bounds for 3WAY).  C carries +P2. */
static void cost_volume(const cost_ctx* c, const uint8_t* img1, const uint8_t* img2, int y0,
                        int y1, int SW2, int SH2, int P2, int16_t* C) {
    int width1 = c->width1, D = c->D, H = c->H;
    long row = (long)width1 * D;
    int nrows = y1 - y0;
    /* hsum for image rows y0 .. min(y1-1+SH2, H-1) */
    int hlast = imin(y1 - 1 + SH2, H - 1);
    int nh = hlast - y0 + 1;
    int16_t* hs = (int16_t*)malloc(sizeof(int16_t) * row * (size_t)nh);
    int16_t* pd = (int16_t*)malloc(sizeof(int16_t) * row);
    for (int k = 0; k < nh; k++) {
        pixel_cost_row(c, img1, img2, y0 + k, pd);
        hsum_row(pd, width1, D, SW2, hs + row * k);
    }
#define HS(yy) (hs + row * (long)(imin(imax((yy), y0), H - 1) - y0))
    for (long i = 0; i < row; i++) {
        /* OpenCV accumulates C += hsum*scale per k with int16 stores */
        int16_t acc = (int16_t)(P2);
        acc = (int16_t)(acc + HS(y0)[i] * (SH2 + 1));
        for (int k = 1; k <= SH2; k++) acc = (int16_t)(acc + HS(y0 + k)[i]);
        C[i] = acc;
    }
    for (int y = y0 + 1; y < y1; y++) {
        const int16_t* add = HS(y + SH2);
        const int16_t* sub = HS(y - SH2 - 1);
        int16_t* Cy = C + row * (long)(y - y0);
        const int16_t* Cp = Cy - row;
        for (long i = 0; i < row; i++) Cy[i] = (int16_t)(Cp[i] + add[i] - sub[i]);
    }
#undef HS
    (void)nrows;
    free(hs);
    free(pd);
}

/* one SGM path update (SIMD int16 semantics of OpenCV: saturating +P1, -delta, +C) */
static inline int path_update(const int16_t* Lp, int minLp, const int16_t* Cp, int D, int P1,
                              int P2, int16_t* L) {
    int16_t delta = (int16_t)(P2 + minLp);
    int mn = MAX_COST;
    for (int d = 0; d < D; d++) {
        int a = Lp[d];
        int b = d > 0 ? sat16(Lp[d - 1] + P1) : MAX_COST;
        int e = d < D - 1 ? sat16(Lp[d + 1] + P1) : MAX_COST;
        int m = imin(imin(imin(a, b), e), delta);
        int v = sat16(sat16(m - delta) + Cp[d]);
        L[d] = (int16_t)v;
        if (v < mn) mn = v;
    }
    return mn;
}

static void lr_check_row(int16_t* disp1, const int16_t* disp2, int W, int minX1, int maxX1,
                         int minD, int disp12MaxDiff, int INVALID_SCALED) {
    for (int x = minX1; x < maxX1; x++) {
        int d1 = disp1[x];
        if (d1 == INVALID_SCALED) continue;
        int _d = d1 >> DISP_SHIFT;
        int d_ = (d1 + DISP_SCALE - 1) >> DISP_SHIFT;
        int _x = x - _d, x_ = x - d_;
        if (0 <= _x && _x < W && disp2[_x] >= minD && abs(disp2[_x] - _d) > disp12MaxDiff &&
            0 <= x_ && x_ < W && disp2[x_] >= minD && abs(disp2[x_] - d_) > disp12MaxDiff)
            disp1[x] = (int16_t)INVALID_SCALED;
    }
}

static inline int subpixel(const int16_t* Sp, int d, int D) {
    if (0 < d && d < D - 1) {
        int denom2 = imax(Sp[d - 1] + Sp[d + 1] - 2 * Sp[d], 1);
        return d * DISP_SCALE + ((Sp[d - 1] - Sp[d + 1]) * DISP_SCALE + denom2) / (denom2 * 2);
    }
    return d * DISP_SCALE;
}

/* modes 0 (5 paths), 1 (8 paths) and 3 (HH4: the horizontal and the vertical path of each of the two passes) */
static int sgbm_full(const uint8_t* img1, const uint8_t* img2, int W, int H,
                     const orc_sgbm_params* p, int16_t* disp, int16_t* C_out, int16_t* S_out) {
    int minD = p->minDisparity, D = p->numDisparities, maxD = minD + D;
    int uniq = p->uniquenessRatio >= 0 ? p->uniquenessRatio : 10;
    int d12 = p->disp12MaxDiff > 0 ? p->disp12MaxDiff : 1;
    int P1 = p->P1 > 0 ? p->P1 : 2, P2 = imax(p->P2 > 0 ? p->P2 : 5, P1 + 1);
    int minX1 = imax(maxD, 0), maxX1 = W + imin(minD, 0), width1 = maxX1 - minX1;
    int INVALID = minD - 1, INVALID_SCALED = INVALID * DISP_SCALE;
    int SW2 = p->blockSize / 2, SH2 = p->blockSize / 2;
    int npasses = (p->mode == 1 || p->mode == 3) ? 2 : 1;
    int hh4 = p->mode == 3; /* MODE_HH4: only the horizontal (r = 0) and the vertical (r = 2) path of each pass */
    for (long i = 0; i < (long)W * H; i++) disp[i] = (int16_t)INVALID_SCALED;
    if (minX1 >= maxX1) return 0;

    long row = (long)width1 * D;
    cost_ctx cc = {W, H, minD, maxD, D, minX1, maxX1, width1, imax(p->preFilterCap, 15) | 1, NULL};
    cc.buf = (uint8_t*)malloc((size_t)12 * W);
    int16_t* C = (int16_t*)malloc(sizeof(int16_t) * row * (size_t)H);
    int16_t* S = (int16_t*)calloc((size_t)row * H, sizeof(int16_t));
    cost_volume(&cc, img1, img2, 0, H, SW2, SH2, P2, C);
    if (hh4) {
        /* OpenCV's HH4 cost loop (CalcVerticalSums) has no branch for window rows below the image: C(y) of a row
         * y >= 1 is only written while y + SH2 < H, so the last SH2 rows keep their initial value P2 for every
         * disparity.  Found by differential testing against cv2 4.13 (mode = 3); reproduced here bit for bit. */
        for (int y = imax(H - SH2, 1); y < H; y++)
            for (long i = 0; i < row; i++) C[row * y + i] = (int16_t)P2;
    }

    /* Lr ring: [2 rows][width1+2][4 paths][D], minLr [2][width1+2][4] */
    long lrrow = (long)(width1 + 2) * 4 * D;
    int16_t* Lr = (int16_t*)malloc(sizeof(int16_t) * 2 * lrrow);
    int16_t* mLr = (int16_t*)malloc(sizeof(int16_t) * 2 * (width1 + 2) * 4);
    int16_t* disp2 = (int16_t*)malloc(sizeof(int16_t) * (W + 2));
    int16_t* disp2cost = (int16_t*)malloc(sizeof(int16_t) * (W + 2));
    int16_t* Ltmp = (int16_t*)malloc(sizeof(int16_t) * D);
#define LR(id, x, r) (Lr + (id) * lrrow + ((long)((x) + 1) * 4 + (r)) * D)
#define MLR(id, x, r) (mLr[(id) * (width1 + 2) * 4 + ((x) + 1) * 4 + (r)])
    for (int pass = 1; pass <= npasses; pass++) {
        int x1, y1, x2, y2, dx, dy;
        if (pass == 1) { y1 = 0; y2 = H; dy = 1; x1 = 0; x2 = width1; dx = 1; }
        else { y1 = H - 1; y2 = -1; dy = -1; x1 = width1 - 1; x2 = -1; dx = -1; }
        int id = 0;
        memset(Lr, 0, sizeof(int16_t) * 2 * lrrow);
        memset(mLr, 0, sizeof(int16_t) * 2 * (width1 + 2) * 4);
        for (int y = y1; y != y2; y += dy) {
            const int16_t* Cy = C + row * y;
            int16_t* Sy = S + row * y;
            for (int x = x1; x != x2; x += dx) {
                const int16_t* Cp = Cy + (long)x * D;
                int16_t* Sp = Sy + (long)x * D;
                const int16_t* Lp[4] = {LR(id, x - dx, 0), LR(1 - id, x - 1, 1),
                                        LR(1 - id, x, 2), LR(1 - id, x + 1, 3)};
                int mp[4] = {MLR(id, x - dx, 0), MLR(1 - id, x - 1, 1), MLR(1 - id, x, 2),
                             MLR(1 - id, x + 1, 3)};
                for (int r = 0; r < 4; r++) {
                    if (hh4 && (r & 1)) continue;
                    int16_t* L = LR(id, x, r);
                    int mn = path_update(Lp[r], mp[r], Cp, D, P1, P2, L);
                    MLR(id, x, r) = (int16_t)mn;
                }
                for (int d = 0; d < D; d++) {
                    int s = Sp[d];
                    for (int r = 0; r < 4; r++) {
                        if (hh4 && (r & 1)) continue;
                        s = sat16(s + LR(id, x, r)[d]);
                    }
                    Sp[d] = (int16_t)s;
                }
            }
            if (pass == npasses) {
                int16_t* d1 = disp + (long)y * W;
                for (int x = 0; x < W + 2; x++) { disp2[x] = (int16_t)INVALID_SCALED; disp2cost[x] = MAX_COST; }
                /* zero state beyond the right border for the 5th path of mode 0 */
                if (npasses == 1) { memset(LR(id, width1, 0), 0, sizeof(int16_t) * D); MLR(id, width1, 0) = 0; }
                for (int x = width1 - 1; x >= 0; x--) {
                    int16_t* Sp = Sy + (long)x * D;
                    int minS = MAX_COST, best = -1;
                    if (npasses == 1) {
                        const int16_t* Cp = Cy + (long)x * D;
                        int mn = path_update(LR(id, x + 1, 0), MLR(id, x + 1, 0), Cp, D, P1, P2, Ltmp);
                        memcpy(LR(id, x, 0), Ltmp, sizeof(int16_t) * D);
                        MLR(id, x, 0) = (int16_t)mn;
                        for (int d = 0; d < D; d++) Sp[d] = (int16_t)sat16(Sp[d] + Ltmp[d]);
                    }
                    for (int d = 0; d < D; d++)
                        if (Sp[d] < minS) { minS = Sp[d]; best = d; }
                    int d;
                    for (d = 0; d < D; d++)
                        if (Sp[d] * (100 - uniq) < minS * 100 && abs(best - d) > 1) break;
                    if (d < D) continue;
                    d = best;
                    int _x2 = x + minX1 - d - minD;
                    if (_x2 >= 0 && _x2 < W + 2 && disp2cost[_x2] > minS) {
                        disp2cost[_x2] = (int16_t)minS;
                        disp2[_x2] = (int16_t)(d + minD);
                    }
                    d = subpixel(Sp, d, D);
                    d1[x + minX1] = (int16_t)(d + minD * DISP_SCALE);
                }
                lr_check_row(d1, disp2, W, minX1, maxX1, minD, d12, INVALID_SCALED);
            }
            id = 1 - id;
        }
    }
#undef LR
#undef MLR
    if (C_out) memcpy(C_out, C, sizeof(int16_t) * row * (size_t)H);
    if (S_out) memcpy(S_out, S, sizeof(int16_t) * row * (size_t)H);
    free(cc.buf); free(C); free(S); free(Lr); free(mLr); free(disp2); free(disp2cost); free(Ltmp);
    return 0;
}

/* mode 2: SGBM_3WAY, fixed 4 stripes with overlap, three saturating paths */
static int sgbm_3way(const uint8_t* img1, const uint8_t* img2, int W, int H,
                     const orc_sgbm_params* p, int16_t* disp) {
    int minD = p->minDisparity, D = p->numDisparities, maxD = minD + D;
    int uniq = p->uniquenessRatio >= 0 ? p->uniquenessRatio : 10;
    int d12 = p->disp12MaxDiff > 0 ? p->disp12MaxDiff : 1;
    int P1 = p->P1 > 0 ? p->P1 : 2, P2 = imax(p->P2 > 0 ? p->P2 : 5, P1 + 1);
    int minX1 = imax(maxD, 0), maxX1 = W + imin(minD, 0), width1 = maxX1 - minX1;
    int INVALID_SCALED = (minD - 1) * DISP_SCALE;
    int SW2 = p->blockSize / 2, SH2 = p->blockSize / 2;
    for (long i = 0; i < (long)W * H; i++) disp[i] = (int16_t)INVALID_SCALED;
    if (minX1 >= maxX1) return 0;

    const int nstripes = 4;
    int stripe_sz = (H + nstripes - 1) / nstripes;
    int overlap;
    { /* (blockSize/2 + 1) + ceil(0.1 * stripe_sz) */
        double t = 0.1 * stripe_sz; int ci = (int)t; if ((double)ci < t) ci++;
        overlap = (p->blockSize / 2 + 1) + ci;
    }
    long row = (long)width1 * D;
    cost_ctx cc = {W, H, minD, maxD, D, minX1, maxX1, width1, imax(p->preFilterCap, 15) | 1, NULL};
    cc.buf = (uint8_t*)malloc((size_t)12 * W);
    int16_t* top = (int16_t*)malloc(sizeof(int16_t) * row);
    int16_t* topMin = (int16_t*)malloc(sizeof(int16_t) * width1);
    int16_t* hor = (int16_t*)malloc(sizeof(int16_t) * row); /* left + top */
    int16_t* La = (int16_t*)malloc(sizeof(int16_t) * D);
    int16_t* Lb = (int16_t*)malloc(sizeof(int16_t) * D);
    int16_t* disp2 = (int16_t*)malloc(sizeof(int16_t) * W);
    int16_t* disp2cost = (int16_t*)malloc(sizeof(int16_t) * W);

    for (int s = 0; s < nstripes; s++) {
        int y0 = imax(imin(s * stripe_sz - overlap, H), 0);
        int y1 = imin((s + 1) * stripe_sz, H);
        int emit0 = s * stripe_sz;
        /* OpenCV writes stripe s into a buffer of its own, image row y at buffer row (s == 0 ? overlap : 0) + (y - y0),
         * and assembles output row i from buffer row overlap + i % stripe_sz of stripe i / stripe_sz.  When the stripe
         * start is clamped at the image top (s >= 1 and s * stripe_sz < overlap: images of a few rows only) the two do
         * not meet: output row i then shows the result of image row i + shift, rows past the stripe end stay invalid
         * (found by differential fuzzing against cv2 4.13; stereosgbm.cpp SGBM3WayMainLoop / computeDisparity3WaySGBM). */
        int shift = s >= 1 ? imax(overlap - s * stripe_sz, 0) : 0;
        if (y1 <= y0) continue;
        int16_t* C = (int16_t*)malloc(sizeof(int16_t) * row * (size_t)(y1 - y0));
        cost_volume(&cc, img1, img2, y0, y1, SW2, SH2, P2, C);
        memset(top, 0, sizeof(int16_t) * row);
        memset(topMin, 0, sizeof(int16_t) * width1);
        for (int y = y0; y < y1; y++) {
            const int16_t* Cy = C + row * (long)(y - y0);
            /* left->right + top->down */
            memset(La, 0, sizeof(int16_t) * D);
            int leftMin = 0;
            for (int x = 0; x < width1; x++) {
                const int16_t* Cp = Cy + (long)x * D;
                int mn = path_update(La, leftMin, Cp, D, P1, P2, Lb);
                leftMin = mn;
                memcpy(La, Lb, sizeof(int16_t) * D);
                int16_t* T = top + (long)x * D;
                int tm = path_update(T, topMin[x], Cp, D, P1, P2, Lb);
                topMin[x] = (int16_t)tm;
                memcpy(T, Lb, sizeof(int16_t) * D);
                int16_t* Hp = hor + (long)x * D;
                for (int d = 0; d < D; d++) Hp[d] = (int16_t)sat16(La[d] + T[d]);
            }
            if (y - shift < emit0) {
                /* overlap rows: the right pass result is discarded, but nothing else depends on it */
                continue;
            }
            int16_t* d1 = disp + (long)(y - shift) * W;
            for (int x = 0; x < W; x++) { disp2[x] = (int16_t)INVALID_SCALED; disp2cost[x] = MAX_COST; }
            memset(La, 0, sizeof(int16_t) * D);
            int rightMin = 0;
            for (int x = width1 - 1; x >= 0; x--) {
                const int16_t* Cp = Cy + (long)x * D;
                int16_t* Hp = hor + (long)x * D;
                int mn = path_update(La, rightMin, Cp, D, P1, P2, Lb);
                rightMin = mn;
                memcpy(La, Lb, sizeof(int16_t) * D);
                /* S = hor + right; arg-min with the 8-lane rule: per lane (d mod 8) the LAST
                 * minimum, across lanes the smallest d among lanes that hold the global min */
                int lane_min[8], lane_pos[8];
                for (int l = 0; l < 8; l++) { lane_min[l] = MAX_COST; lane_pos[l] = 0; }
                for (int d = 0; d < D; d++) {
                    int v = sat16(Hp[d] + La[d]);
                    Hp[d] = (int16_t)v;
                    int l = d & 7;
                    if (v <= lane_min[l]) { lane_min[l] = v; lane_pos[l] = d; }
                }
                int minS = MAX_COST, best = 0;
                for (int l = 0; l < 8; l++) if (lane_min[l] < minS) minS = lane_min[l];
                { int bp = 32767; for (int l = 0; l < 8; l++) if (lane_min[l] == minS && lane_pos[l] < bp) bp = lane_pos[l]; best = bp; }
                if (uniq > 0) {
                    int thresh = (100 * minS) / (100 - uniq);
                    int16_t tr = (int16_t)(thresh + 1);
                    int d;
                    for (d = 0; d < D; d++)
                        if (Hp[d] < tr && (d < best - 1 || d > best + 1)) break;
                    if (d < D) continue;
                }
                int d = best;
                int _x2 = x + minX1 - d - minD;
                if (_x2 >= 0 && _x2 < W && disp2cost[_x2] > minS) {
                    disp2cost[_x2] = (int16_t)minS;
                    disp2[_x2] = (int16_t)(d + minD);
                }
                d = subpixel(Hp, d, D);
                d1[x + minX1] = (int16_t)(d + minD * DISP_SCALE);
            }
            lr_check_row(d1, disp2, W, minX1, maxX1, minD, d12, INVALID_SCALED);
        }
        free(C);
    }
    free(cc.buf); free(top); free(topMin); free(hor); free(La); free(Lb); free(disp2); free(disp2cost);
    return 0;
}

void orc_median3_s16(const int16_t* src, int W, int H, int16_t* dst) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int16_t v[9];
            int n = 0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int yy = imin(imax(y + dy, 0), H - 1), xx = imin(imax(x + dx, 0), W - 1);
                    v[n++] = src[(long)yy * W + xx];
                }
            for (int i = 1; i < 9; i++) { /* insertion sort */
                int16_t k = v[i]; int j = i - 1;
                while (j >= 0 && v[j] > k) { v[j + 1] = v[j]; j--; }
                v[j + 1] = k;
            }
            dst[(long)y * W + x] = v[4];
        }
}

void orc_filter_speckles(int16_t* img, int W, int H, int newVal, int maxSize, int maxDiff) {
    long n = (long)W * H;
    int* labels = (int*)calloc((size_t)n, sizeof(int));
    int* stack = (int*)malloc(sizeof(int) * (size_t)n);
    uint8_t* small = (uint8_t*)malloc((size_t)n + 1);
    int cur = 0;
    for (long i = 0; i < n; i++) {
        if (img[i] == newVal) continue;
        if (labels[i]) { if (small[labels[i]]) img[i] = (int16_t)newVal; continue; }
        int sp = 0, count = 0;
        cur++;
        labels[i] = cur;
        stack[sp++] = (int)i;
        while (sp) {
            int q = stack[--sp];
            count++;
            int y = q / W, x = q % W;
            int dq = img[q];
            const int nb[4] = {x > 0 ? q - 1 : -1, x < W - 1 ? q + 1 : -1, y > 0 ? q - W : -1, y < H - 1 ? q + W : -1};
            for (int k = 0; k < 4; k++) {
                int r = nb[k];
                if (r < 0 || labels[r] || img[r] == newVal) continue;
                if (abs(img[r] - dq) <= maxDiff) { labels[r] = cur; stack[sp++] = r; }
            }
        }
        small[cur] = count <= maxSize;
        if (small[cur]) img[i] = (int16_t)newVal;
    }
    free(labels); free(stack); free(small);
}

int orc_sgbm_compute(const uint8_t* left, const uint8_t* right, int W, int H,
                     const orc_sgbm_params* p, int16_t* disp, int16_t* raw_out, int16_t* C_out,
                     int16_t* S_out) {
    int rc;
    int16_t* raw = (int16_t*)malloc(sizeof(int16_t) * (size_t)W * H);
    if (p->mode == 2) rc = sgbm_3way(left, right, W, H, p, raw);
    else if (p->mode == 0 || p->mode == 1 || p->mode == 3) rc = sgbm_full(left, right, W, H, p, raw, C_out, S_out);
    else { free(raw); return -1; }
    if (raw_out) memcpy(raw_out, raw, sizeof(int16_t) * (size_t)W * H);
    orc_median3_s16(raw, W, H, disp);
    if (p->speckleWindowSize > 0)
        orc_filter_speckles(disp, W, H, (p->minDisparity - 1) * DISP_SCALE, p->speckleWindowSize,
                            DISP_SCALE * p->speckleRange);
    free(raw);
    return rc;
}
