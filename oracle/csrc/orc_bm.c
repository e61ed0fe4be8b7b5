/*
 * orc_bm.c -- CPU restatement of cv2.StereoBM.compute (XSOBEL prefilter, SAD block matching, texture / uniqueness
 * tests, sub-pixel interpolation, valid-ROI masking, filterSpeckles), the matcher the reference's README suggests as a
 * drop-in for StereoSGBM (readme.md:392-397).  TEST INFRASTRUCTURE ONLY (see l3d_oracle.h).
 * OpenCV's source is not in the image: the arithmetic is restated from the published algorithm and pinned by
 * differential tests against the cv2 4.13 binary (tests/test_oracle_cv2.py::test_stereobm_vs_cv2).  Supported:
 * minDisparity <= 0 (OpenCV itself writes past the row end for positive values), PREFILTER_XSOBEL; disp12MaxDiff >= 0
 * runs cv2.validateDisparity between matching and masking (pinned for preFilterCap <= 31 and blockSize <= 21, OpenCV's
 * 16-bit SIMD path; its scalar path combines the check with a differently laid out cost buffer and is not restated).
 */
#include "l3d_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <limits.h>
typedef orc_bm_params bm_params;
static int imin(int a,int b){return a<b?a:b;} static int imax(int a,int b){return a>b?a:b;}

void orc_bm_prefilter_xsobel(const uint8_t* src, int W, int H, int ftzero, uint8_t* dst) {
    enum { OFS = 256*4, TABSZ = OFS*2 + 256 };
    uint8_t tab[TABSZ];
    for (int x = 0; x < TABSZ; x++) tab[x] = (uint8_t)(x - OFS < -ftzero ? 0 : x - OFS > ftzero ? ftzero*2 : x - OFS + ftzero);
    uint8_t val0 = tab[0 + OFS];
    int y;
    for (y = 0; y < H - 1; y += 2) {
        const uint8_t* srow1 = src + (long)y*W;
        const uint8_t* srow0 = y > 0 ? srow1 - W : H > 1 ? srow1 + W : srow1;
        const uint8_t* srow2 = y < H-1 ? srow1 + W : H > 1 ? srow1 - W : srow1;
        const uint8_t* srow3 = y < H-2 ? srow1 + W*2 : srow1;
        uint8_t* d0p = dst + (long)y*W; uint8_t* d1p = d0p + W;
        d0p[0] = d0p[W-1] = d1p[0] = d1p[W-1] = val0;
        for (int x = 1; x < W-1; x++) {
            int d0 = srow0[x+1] - srow0[x-1], d1 = srow1[x+1] - srow1[x-1], d2 = srow2[x+1] - srow2[x-1], d3 = srow3[x+1] - srow3[x-1];
            d0p[x] = tab[d0 + d1*2 + d2 + OFS];
            d1p[x] = tab[d1 + d2*2 + d3 + OFS];
        }
    }
    for (; y < H; y++) { uint8_t* dp = dst + (long)y*W; for (int x = 0; x < W; x++) dp[x] = val0; }
}

/* findStereoCorrespondenceBM, whole image (dy0 = dy1 = 0) */
static void bm_correspond(const uint8_t* left, const uint8_t* right, int width, int height, const bm_params* st, int16_t* disp,
                          int* cost) {
    int wsz = st->blockSize, wsz2 = wsz/2;
    int dy0 = 0, dy1 = 0;
    int ndisp = st->numDisparities, mindisp = st->minDisparity;
    int lofs = imax(ndisp - 1 + mindisp, 0), rofs = -imin(ndisp - 1 + mindisp, 0);
    int width1 = width - rofs - ndisp + 1;
    int ftzero = st->preFilterCap, textureThreshold = st->textureThreshold, uniquenessRatio = st->uniquenessRatio;
    int16_t FILTERED = (int16_t)((mindisp - 1) << 4);
    for (long i = 0; i < (long)width*height; i++) disp[i] = FILTERED;
    if (lofs >= width || rofs >= width || width1 < 1) return;
    int sstep = width, dstep = width;
    int cstep = (height+dy0+dy1)*ndisp;
    int* sadb = (int*)calloc(ndisp + 2, sizeof(int)); int* sad = sadb + 1;
    int* hsad0 = (int*)calloc((size_t)(height+dy0+dy1)*ndisp, sizeof(int));
    int* htextb = (int*)calloc(height + wsz + 2, sizeof(int)); int* htext = htextb + wsz2 + 1;
    uint8_t* cbuf0 = (uint8_t*)calloc((size_t)(wsz+1)*cstep, 1);
    uint8_t tab[256];
    for (int x = 0; x < 256; x++) tab[x] = (uint8_t)abs(x - ftzero);
    const uint8_t* lptr0 = left + lofs; const uint8_t* rptr0 = right + rofs;
    for (int x = -wsz2-1; x < wsz2; x++) {
        int* hsad = hsad0 - dy0*ndisp; uint8_t* cbuf = cbuf0 + (x + wsz2 + 1)*cstep - dy0*ndisp;
        const uint8_t* lptr = lptr0 + imin(imax(x, -lofs), width-lofs-1) - dy0*sstep;
        const uint8_t* rptr = rptr0 + imin(imax(x, -rofs), width-rofs-ndisp) - dy0*sstep;
        for (int y = -dy0; y < height + dy1; y++, hsad += ndisp, cbuf += ndisp, lptr += sstep, rptr += sstep) {
            int lval = lptr[0];
            for (int d = 0; d < ndisp; d++) { int diff = abs(lval - rptr[d]); cbuf[d] = (uint8_t)diff; hsad[d] += diff; }
            htext[y] += tab[lval];
        }
    }
    int16_t* dptr = disp + lofs;
    for (int x = 0; x < width1; x++, dptr++) {
        int x0 = x - wsz2 - 1, x1 = x + wsz2;
        const uint8_t* cbuf_sub = cbuf0 + ((x0 + wsz2 + 1) % (wsz + 1))*cstep - dy0*ndisp;
        uint8_t* cbuf = cbuf0 + ((x1 + wsz2 + 1) % (wsz + 1))*cstep - dy0*ndisp;
        int* hsad = hsad0 - dy0*ndisp;
        const uint8_t* lptr_sub = lptr0 + imin(imax(x0, -lofs), width-1-lofs) - dy0*sstep;
        const uint8_t* lptr = lptr0 + imin(imax(x1, -lofs), width-1-lofs) - dy0*sstep;
        const uint8_t* rptr = rptr0 + imin(imax(x1, -rofs), width-ndisp-rofs) - dy0*sstep;
        for (int y = -dy0; y < height + dy1; y++, cbuf += ndisp, cbuf_sub += ndisp, hsad += ndisp, lptr += sstep, lptr_sub += sstep, rptr += sstep) {
            int lval = lptr[0];
            for (int d = 0; d < ndisp; d++) { int diff = abs(lval - rptr[d]); cbuf[d] = (uint8_t)diff; hsad[d] = hsad[d] + diff - cbuf_sub[d]; }
            htext[y] += tab[lval] - tab[lptr_sub[0]];
        }
        for (int y = dy1; y <= wsz2; y++) htext[height+y] = htext[height+dy1-1];
        for (int y = -wsz2-1; y < -dy0; y++) htext[y] = htext[-dy0];
        for (int d = 0; d < ndisp; d++) sad[d] = (int)(hsad0[d-ndisp*dy0]*(wsz2 + 2 - dy0));
        hsad = hsad0 + (1 - dy0)*ndisp;
        for (int y = 1 - dy0; y < wsz2; y++, hsad += ndisp) for (int d = 0; d < ndisp; d++) sad[d] += hsad[d];
        int tsum = 0;
        for (int y = -wsz2-1; y < wsz2; y++) tsum += htext[y];
        for (int y = 0; y < height; y++) {
            int minsad = INT_MAX, mind = -1;
            hsad = hsad0 + imin(y + wsz2, height+dy1-1)*ndisp;
            int* hsad_sub = hsad0 + imax(y - wsz2 - 1, -dy0)*ndisp;
            for (int d = 0; d < ndisp; d++) {
                int currsad = sad[d] + hsad[d] - hsad_sub[d];
                sad[d] = currsad;
                if (currsad < minsad) { minsad = currsad; mind = d; }
            }
            tsum += htext[y + wsz2] - htext[y - wsz2 - 1];
            if (tsum < textureThreshold) { dptr[y*dstep] = FILTERED; continue; }
            if (uniquenessRatio > 0) {
                int thresh = minsad + (minsad * uniquenessRatio/100);
                int d;
                for (d = 0; d < ndisp; d++) if ((d < mind-1 || d > mind+1) && sad[d] <= thresh) break;
                if (d < ndisp) { dptr[y*dstep] = FILTERED; continue; }
            }
            sad[-1] = sad[1]; sad[ndisp] = sad[ndisp-2];
            int p = sad[mind+1], n = sad[mind-1];
            int d = p + n - 2*sad[mind] + abs(p - n);
            dptr[y*dstep] = (int16_t)(((ndisp - mind - 1 + mindisp)*256 + (d != 0 ? (p-n)*256/d : 0) + 15) >> 4);
            if (cost) cost[(long)y*width + lofs + x] = sad[mind];
        }
    }
    free(sadb); free(hsad0); free(htextb); free(cbuf0);
}

/* cv2.validateDisparity as StereoBM applies it when disp12MaxDiff >= 0 (per row: the best-cost left pixel claims its right
 * pixel, then a disparity survives if either of its two integer roundings agrees with the claim within the tolerance) */
static void bm_validate(int16_t* disp, const int* cost, int cols, int rows, int minD, int ndisp, int disp12MaxDiff) {
    int maxD = minD + ndisp;
    int minX1 = imax(maxD, 0), maxX1 = cols + imin(minD, 0);
    int INVALID_SCALED = (minD - 1) * 16;
    int* disp2buf = (int*)malloc(sizeof(int) * cols * 2);
    int* disp2cost = disp2buf + cols;
    disp12MaxDiff *= 16;
    for (int y = 0; y < rows; y++) {
        int16_t* dptr = disp + (long)y * cols;
        const int* cptr = cost + (long)y * cols;
        for (int x = 0; x < cols; x++) { disp2buf[x] = INVALID_SCALED; disp2cost[x] = INT_MAX; }
        for (int x = minX1; x < maxX1; x++) {
            int d = dptr[x], c = cptr[x];
            if (d == INVALID_SCALED) continue;
            int x2 = x - ((d + 8) >> 4);
            if (disp2cost[x2] > c) { disp2cost[x2] = c; disp2buf[x2] = d; }
        }
        for (int x = minX1; x < maxX1; x++) {
            int d = dptr[x];
            if (d == INVALID_SCALED) continue;
            int d0 = d >> 4, d1 = (d + 15) >> 4;
            int x0 = x - d0, x1 = x - d1;
            if ((0 <= x0 && x0 < cols && disp2buf[x0] > INVALID_SCALED && abs(disp2buf[x0] - d) > disp12MaxDiff) &&
                (0 <= x1 && x1 < cols && disp2buf[x1] > INVALID_SCALED && abs(disp2buf[x1] - d) > disp12MaxDiff))
                dptr[x] = (int16_t)INVALID_SCALED;
        }
    }
    free(disp2buf);
}

int orc_bm_compute(const uint8_t* l, const uint8_t* r, int W, int H, const orc_bm_params* st, int16_t* disp) {
    if (st->minDisparity > 0 || st->numDisparities < 16 || (st->numDisparities % 16) || st->blockSize < 5 || !(st->blockSize & 1) ||
        st->blockSize > W || st->blockSize > H || st->preFilterCap < 1 || st->preFilterCap > 63 ||
        (st->disp12MaxDiff >= 0 && (st->preFilterCap > 31 || st->blockSize > 21)))
        return -1;
    uint8_t* lf = malloc((size_t)W*H); uint8_t* rf = malloc((size_t)W*H);
    orc_bm_prefilter_xsobel(l, W, H, st->preFilterCap, lf);
    orc_bm_prefilter_xsobel(r, W, H, st->preFilterCap, rf);
    int* cost = st->disp12MaxDiff >= 0 ? (int*)calloc((size_t)W * H, sizeof(int)) : NULL;
    bm_correspond(lf, rf, W, H, st, disp, cost);
    if (cost) { bm_validate(disp, cost, W, H, st->minDisparity, st->numDisparities, st->disp12MaxDiff); free(cost); }
    /* valid ROI */
    int SW2 = st->blockSize/2, minD = st->minDisparity, maxD = minD + st->numDisparities - 1;
    int xmin = imax(0, 0 + maxD) + SW2, xmax = W - SW2, ymin = SW2, ymax = H - SW2;
    int16_t FILTERED = (int16_t)((minD - 1) << 4);
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++)
        if (x < xmin || x >= xmax || y < ymin || y >= ymax) disp[(long)y*W + x] = FILTERED;
    free(lf); free(rf);
    if (st->speckleRange >= 0 && st->speckleWindowSize > 0)
        orc_filter_speckles(disp, W, H, FILTERED, st->speckleWindowSize, st->speckleRange);
    return 0;
}
