/*
 * orc_image.c -- CPU restatement of the cv2 image primitives on the reference hot path:
 * colour conversion, fixed-point remap, 3x3 morphology, contour-area mask model, f32 Gaussian /
 * Sobel, disparity->depth.
 *
 * TEST INFRASTRUCTURE ONLY (see l3d_oracle.h).  Pinned against cv2 4.13.0 by
 * tests/test_oracle_image.py (bit-exact for the integer ops, tolerance for f32 filters).
 */
#include "l3d_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* core/laser_extractor.py:60 / camera/...:320 -- 15-bit fixed point BT.601 */
void orc_bgr2gray(const uint8_t* bgr, long n, uint8_t* gray) {
    for (long i = 0; i < n; i++) {
        int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
        gray[i] = (uint8_t)((3735 * b + 19235 * g + 9798 * r + 16384) >> 15);
    }
}

/* core/laser_extractor.py:56 -- 8-bit HSV, H in [0,180), 12-bit reciprocal tables */
void orc_bgr2hsv(const uint8_t* bgr, long n, uint8_t* hsv) {
    static int sdiv[256], hdiv[256], init = 0;
    if (!init) {
        sdiv[0] = hdiv[0] = 0;
        for (int i = 1; i < 256; i++) {
            sdiv[i] = (int)lrint((255 << 12) / (1.0 * i));
            hdiv[i] = (int)lrint((180 << 12) / (6.0 * i));
        }
        init = 1;
    }
    for (long i = 0; i < n; i++) {
        int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
        int v = imax(imax(b, g), r), m = imin(imin(b, g), r), diff = v - m;
        int s = (diff * sdiv[v] + (1 << 11)) >> 12;
        int h;
        if (v == r) h = g - b;
        else if (v == g) h = b - r + 2 * diff;
        else h = r - g + 4 * diff;
        h = (h * hdiv[diff] + (1 << 11)) >> 12;
        if (h < 0) h += 180;
        hsv[3 * i] = (uint8_t)h;
        hsv[3 * i + 1] = (uint8_t)s;
        hsv[3 * i + 2] = (uint8_t)v;
    }
}

/* camera/single_usb_stereo_camera.py:313-314.  5 fractional bits, weights sum 2^15, round at 2^14,
 * taps outside the source contribute the border value 0. */
void orc_remap_bilinear(const uint8_t* src, int sw, int sh, int cn, const float* mapx,
                        const float* mapy, int dw, int dh, uint8_t* dst) {
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++) {
            long i = (long)y * dw + x;
            float fx = mapx[i] * 32.0f, fy = mapy[i] * 32.0f;
            int sx = (int)lrintf(fx), sy = (int)lrintf(fy); /* round half to even */
            int ix = sx >> 5, iy = sy >> 5, ax = sx & 31, ay = sy & 31;
            if (ix > 32767) ix = 32767; if (ix < -32768) ix = -32768;
            if (iy > 32767) iy = 32767; if (iy < -32768) iy = -32768;
            int w00 = (32 - ay) * (32 - ax) * 32, w01 = (32 - ay) * ax * 32,
                w10 = ay * (32 - ax) * 32, w11 = ay * ax * 32;
            for (int c = 0; c < cn; c++) {
                int t00 = 0, t01 = 0, t10 = 0, t11 = 0;
                if (iy >= 0 && iy < sh) {
                    if (ix >= 0 && ix < sw) t00 = src[((long)iy * sw + ix) * cn + c];
                    if (ix + 1 >= 0 && ix + 1 < sw) t01 = src[((long)iy * sw + ix + 1) * cn + c];
                }
                if (iy + 1 >= 0 && iy + 1 < sh) {
                    if (ix >= 0 && ix < sw) t10 = src[((long)(iy + 1) * sw + ix) * cn + c];
                    if (ix + 1 >= 0 && ix + 1 < sw) t11 = src[((long)(iy + 1) * sw + ix + 1) * cn + c];
                }
                int v = (t00 * w00 + t01 * w01 + t10 * w10 + t11 * w11 + 16384) >> 15;
                dst[i * cn + c] = (uint8_t)(v > 255 ? 255 : v);
            }
        }
}

/* 3x3 rectangular dilate (outside = 0) / erode (outside = 255) on 0/255 masks */
static void dilate3(const uint8_t* s, int W, int H, uint8_t* d) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int v = 0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int yy = y + dy, xx = x + dx;
                    if (yy >= 0 && yy < H && xx >= 0 && xx < W && s[(long)yy * W + xx]) v = 255;
                }
            d[(long)y * W + x] = (uint8_t)v;
        }
}
static void erode3(const uint8_t* s, int W, int H, uint8_t* d) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int v = 255;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int yy = y + dy, xx = x + dx;
                    if (yy >= 0 && yy < H && xx >= 0 && xx < W && !s[(long)yy * W + xx]) v = 0;
                }
            d[(long)y * W + x] = (uint8_t)v;
        }
}
/* core/laser_extractor.py:67-69: MORPH_CLOSE then MORPH_OPEN with np.ones((3,3)) */
void orc_close_open3(const uint8_t* src, int W, int H, uint8_t* dst) {
    uint8_t* a = (uint8_t*)malloc((size_t)W * H);
    uint8_t* b = (uint8_t*)malloc((size_t)W * H);
    dilate3(src, W, H, a); erode3(a, W, H, b);  /* close */
    erode3(b, W, H, a);    dilate3(a, W, H, dst); /* open */
    free(a); free(b);
}

/* core/laser_extractor.py:56-82 */
void orc_simple_masks(const uint8_t* bgr, int W, int H, const int* lo, const int* hi,
                      int bright_thr, double min_area, uint8_t* mask_morph, uint8_t* mask_final) {
    long n = (long)W * H;
    uint8_t* hsv = (uint8_t*)malloc((size_t)n * 3);
    uint8_t* gray = (uint8_t*)malloc((size_t)n);
    uint8_t* comb = (uint8_t*)malloc((size_t)n);
    orc_bgr2hsv(bgr, n, hsv);
    orc_bgr2gray(bgr, n, gray);
    for (long i = 0; i < n; i++) {
        int ok = hsv[3 * i] >= lo[0] && hsv[3 * i] <= hi[0] && hsv[3 * i + 1] >= lo[1] &&
                 hsv[3 * i + 1] <= hi[1] && hsv[3 * i + 2] >= lo[2] && hsv[3 * i + 2] <= hi[2];
        comb[i] = (ok && gray[i] > bright_thr) ? 255 : 0;
    }
    orc_close_open3(comb, W, H, mask_morph);
    /* findContours(RETR_EXTERNAL) + contourArea > min_area + drawContours(filled):
     * filled = not(4-connected background reachable from the image border); 8-connected labels;
     * area = #full 2x2 windows + 0.5 * #2x2 windows with exactly three set pixels. */
    uint8_t* filled = (uint8_t*)malloc((size_t)n);
    int* stack = (int*)malloc(sizeof(int) * (size_t)n);
    int* label = (int*)calloc((size_t)n, sizeof(int));
    memset(filled, 255, (size_t)n);
    int sp = 0;
#define PUSH_BG(q) do { if (!mask_morph[q] && filled[q]) { filled[q] = 0; stack[sp++] = (int)(q); } } while (0)
    for (int x = 0; x < W; x++) { PUSH_BG((long)x); PUSH_BG((long)(H - 1) * W + x); }
    for (int y = 0; y < H; y++) { PUSH_BG((long)y * W); PUSH_BG((long)y * W + W - 1); }
    while (sp) {
        int q = stack[--sp]; int y = q / W, x = q % W;
        if (x > 0) PUSH_BG(q - 1);
        if (x < W - 1) PUSH_BG(q + 1);
        if (y > 0) PUSH_BG(q - W);
        if (y < H - 1) PUSH_BG(q + W);
    }
#undef PUSH_BG
    int nl = 0;
    for (long i = 0; i < n; i++) {
        if (!filled[i] || label[i]) continue;
        nl++; label[i] = nl; stack[sp++] = (int)i;
        while (sp) {
            int q = stack[--sp]; int y = q / W, x = q % W;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int yy = y + dy, xx = x + dx;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    long r = (long)yy * W + xx;
                    if (filled[r] && !label[r]) { label[r] = nl; stack[sp++] = (int)r; }
                }
        }
    }
    long* area2 = (long*)calloc((size_t)nl + 1, sizeof(long)); /* twice the area */
    for (int y = 0; y + 1 < H; y++)
        for (int x = 0; x + 1 < W; x++) {
            long q = (long)y * W + x;
            int c = (filled[q] != 0) + (filled[q + 1] != 0) + (filled[q + W] != 0) + (filled[q + W + 1] != 0);
            if (c < 3) continue;
            int l = label[q] ? label[q] : label[q + 1];
            area2[l] += c == 4 ? 2 : 1;
        }
    for (long i = 0; i < n; i++)
        mask_final[i] = (filled[i] && 0.5 * (double)area2[label[i]] > min_area) ? 255 : 0;
    free(hsv); free(gray); free(comb); free(filled); free(stack); free(label); free(area2);
}

static inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; }
    return p;
}

/* core/laser_extractor.py:193, improved_steger.py:59,148,279 */
void orc_gaussian_blur_f32(const float* src, int W, int H, double sigma, float* dst) {
    int ks = (int)lrint(sigma * 8 + 1) | 1;
    int r = ks / 2;
    float* k = (float*)malloc(sizeof(float) * ks);
    double sum = 0, s2 = -0.5 / (sigma * sigma);
    double* kd = (double*)malloc(sizeof(double) * ks);
    for (int i = 0; i < ks; i++) { double x = i - r; kd[i] = exp(s2 * x * x); sum += kd[i]; }
    for (int i = 0; i < ks; i++) k[i] = (float)(kd[i] / sum);
    float* tmp = (float*)malloc(sizeof(float) * (size_t)W * H);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            float s = k[r] * src[(long)y * W + x];
            for (int j = 1; j <= r; j++)
                s += k[r + j] * (src[(long)y * W + reflect101(x - j, W)] + src[(long)y * W + reflect101(x + j, W)]);
            tmp[(long)y * W + x] = s;
        }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            float s = k[r] * tmp[(long)y * W + x];
            for (int j = 1; j <= r; j++)
                s += k[r + j] * (tmp[(long)reflect101(y - j, H) * W + x] + tmp[(long)reflect101(y + j, H) * W + x]);
            dst[(long)y * W + x] = s;
        }
    free(k); free(kd); free(tmp);
}

/* improved_steger.py:63-69 -- 3x3 Sobel, (dx,dy) in {(1,0),(0,1)}, unnormalised, REFLECT_101 */
void orc_sobel3_f32(const float* src, int W, int H, int dx, int dy, float* dst) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
            int ym = reflect101(y - 1, H), yp = reflect101(y + 1, H);
#define S(yy, xx) src[(long)(yy) * W + (xx)]
            float v;
            if (dx == 1 && dy == 0)
                v = ((S(ym, xp) - S(ym, xm)) + (S(yp, xp) - S(yp, xm))) + 2.0f * (S(y, xp) - S(y, xm));
            else /* cv2's symmetric [1,2,1] order: (a + c) + 2b, rows first */
                v = ((S(yp, xm) + S(yp, xp)) + 2.0f * S(yp, x)) - ((S(ym, xm) + S(ym, xp)) + 2.0f * S(ym, x));
#undef S
            dst[(long)y * W + x] = v;
        }
}

/* camera/single_usb_stereo_camera.py:335-346 */
void orc_disp_to_depth_q(const int16_t* disp16, int W, int H, const double* Q, float* depth) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            float disp = (float)disp16[(long)y * W + x] / 16.0f;
            double d = disp;
            double Z = Q[8] * x + Q[9] * y + Q[10] * d + Q[11];
            double Wh = Q[12] * x + Q[13] * y + Q[14] * d + Q[15];
            float zf = (float)Z; /* Vec3f store before the division (probed against cv2 4.13) */
            float z = (float)((double)zf / Wh);
            if (z < 0) z = 0;
            if (z > 10) z = 0;
            if (disp <= 0) z = 0;
            if (z != z) z = z; /* NaN (0/0) survives the three masks in numpy too */
            depth[(long)y * W + x] = z;
        }
}

/* camera/single_usb_stereo_camera.py:347-357 */
void orc_disp_to_depth_default(const int16_t* disp16, int W, int H, float* depth) {
    for (long i = 0; i < (long)W * H; i++) {
        float disp = (float)disp16[i] / 16.0f;
        float z = 0.0f;
        if (disp > 0) z = (float)(0.06 * 350) / disp; /* python float (weak) / np.float32 array -> f32 divide */
        if (z > 10) z = 0;
        depth[i] = z;
    }
}
