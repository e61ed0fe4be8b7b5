/*
 * orc_wls.c -- CPU restatement of cv2.ximgproc.DisparityWLSFilter.filter(dl, guide,
 * disparity_map_right=dr) as the reference sets it up (camera/single_usb_stereo_camera.py:277-282:
 * createRightMatcher + createDisparityWLSFilter(left_matcher), lambda 8000, sigma_color 1.5) and
 * calls it (:328-332).
 *
 * TEST INFRASTRUCTURE ONLY (see l3d_oracle.h).
 * PARITY UNPINNED: the algorithm lives in opencv_contrib/ximgproc (disparity_filters.cpp,
 * fgs_filter.cpp), a third-party module that is neither under /root/reference nor installed in
 * this image (cv2.ximgproc missing, no wheel, no network), and the reference pins no version for
 * it.  This file restates the published algorithm: confidence = LR-consistency x (1 - 0.001 *
 * local disparity variance), then confidence-weighted Fast Global Smoother (Min et al. 2014;
 * 3 iterations, lambda attenuation 0.25, horizontal + vertical Thomas solves per iteration).
 */
#include "l3d_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; }
    return p;
}

/* normalised (2r+1)^2 box filter, BORDER_REFLECT_101, on a w x h f32 image */
static void box_f32(const float* s, int w, int h, int r, float* d) {
    float* t = (float*)malloc(sizeof(float) * (size_t)w * h);
    float inv = 1.0f / (float)((2 * r + 1) * (2 * r + 1));
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float a = 0;
            for (int k = -r; k <= r; k++) a += s[(long)y * w + reflect101(x + k, w)];
            t[(long)y * w + x] = a;
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float a = 0;
            for (int k = -r; k <= r; k++) a += t[(long)reflect101(y + k, h) * w + x];
            d[(long)y * w + x] = a * inv;
        }
    free(t);
}

/* (2r+1)^2 box mean whose horizontal taps come from the FULL image row (columns x0+x+k reflected at the image border):
 * variant ORC_WLS_BOX_FULL_IMAGE */
static void box_full(const int16_t* disp, int W, int x0, int w, int h, int r, int squared, float* d) {
    float* t = (float*)malloc(sizeof(float) * (size_t)w * h);
    float inv = 1.0f / (float)((2 * r + 1) * (2 * r + 1));
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float a = 0;
            for (int k = -r; k <= r; k++) {
                float v = (float)disp[(long)y * W + reflect101(x0 + x + k, W)];
                a += squared ? v * v : v;
            }
            t[(long)y * w + x] = a;
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float a = 0;
            for (int k = -r; k <= r; k++) a += t[(long)reflect101(y + k, h) * w + x];
            d[(long)y * w + x] = a * inv;
        }
    free(t);
}

/* 1 - 0.001*var(disparity) clamped below at 0 (variant: to [0,1]), on the ROI copy (variant: across the ROI edge) */
static void discontinuity_map(const int16_t* disp, int W, int x0, int w, int h, int r, int variant, float* dst) {
    size_t n = (size_t)w * h;
    float* a = (float*)malloc(sizeof(float) * n);
    float* b = (float*)malloc(sizeof(float) * n);
    float* ma = (float*)malloc(sizeof(float) * n);
    float* mb = (float*)malloc(sizeof(float) * n);
    if (variant & ORC_WLS_BOX_FULL_IMAGE) {
        box_full(disp, W, x0, w, h, r, 0, ma);
        box_full(disp, W, x0, w, h, r, 1, mb);
    } else {
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                float v = (float)disp[(long)y * W + x0 + x];
                a[(long)y * w + x] = v;
                b[(long)y * w + x] = v * v;
            }
        box_f32(a, w, h, r, ma);
        box_f32(b, w, h, r, mb);
    }
    for (size_t i = 0; i < n; i++) {
        float var = mb[i] - ma[i] * ma[i];
        float c = 1.0f - 0.001f * var;
        if ((variant & ORC_WLS_CONF_CLAMP_1) && c > 1.0f) c = 1.0f;
        dst[i] = c < 0.0f ? 0.0f : c;
    }
    free(a); free(b); free(ma); free(mb);
}

/* one Fast-Global-Smoother solve over the w x h image u (in place); ch/cv = -exp(-|dg|/sigma) */
static void fgs_solve(float* u, int w, int h, const float* ch, const float* cv, double lambda0, int variant) {
    float* D = (float*)malloc(sizeof(float) * (size_t)(w > h ? w : h));
    float lam = (float)lambda0;
    for (int it = 0; it < 3; it++) {
        for (int y = 0; y < h; y++) { /* horizontal pass */
            float* r = u + (long)y * w;
            const float* c = ch + (long)y * w;
            float den = 1.0f - lam * c[0];
            D[0] = (lam * c[0]) / den;
            r[0] = r[0] / den;
            for (int j = 1; j < w; j++) {
                den = (1.0f - lam * (c[j - 1] + c[j])) - (lam * c[j - 1]) * D[j - 1];
                D[j] = (lam * c[j]) / den;
                r[j] = (r[j] - (lam * c[j - 1]) * r[j - 1]) / den;
            }
            for (int j = w - 2; j >= 0; j--) r[j] = r[j] - D[j] * r[j + 1];
        }
        if (variant & ORC_WLS_LAMBDA_PER_PASS) lam *= 0.25f;
        for (int x = 0; x < w; x++) { /* vertical pass */
            float den = 1.0f - lam * cv[x];
            D[0] = (lam * cv[x]) / den;
            u[x] = u[x] / den;
            for (int j = 1; j < h; j++) {
                float cm = cv[(long)(j - 1) * w + x], cc = cv[(long)j * w + x];
                den = (1.0f - lam * (cm + cc)) - (lam * cm) * D[j - 1];
                D[j] = (lam * cc) / den;
                u[(long)j * w + x] = (u[(long)j * w + x] - (lam * cm) * u[(long)(j - 1) * w + x]) / den;
            }
            for (int j = h - 2; j >= 0; j--)
                u[(long)j * w + x] = u[(long)j * w + x] - D[j] * u[(long)(j + 1) * w + x];
        }
        lam *= 0.25f;
    }
    free(D);
}

void orc_wls_filter(const int16_t* dl, const int16_t* dr, const uint8_t* guide, int W, int H,
                    int min_disp, int num_disp, int dd_radius, double lambda, double sigma_color,
                    int lrc_thresh, int16_t* out, float* conf_out) {
    orc_wls_filter_v(dl, dr, guide, W, H, min_disp, num_disp, dd_radius, lambda, sigma_color, lrc_thresh, 0, out, conf_out);
}

/* variant: ORC_WLS_* bits -- the points of opencv_contrib's implementation this restatement cannot pin;
 * 0 = the documented reading, each bit switches one point to the alternative reading */
void orc_wls_filter_v(const int16_t* dl, const int16_t* dr, const uint8_t* guide, int W, int H,
                      int min_disp, int num_disp, int dd_radius, double lambda, double sigma_color,
                      int lrc_thresh, int variant, int16_t* out, float* conf_out) {
    int x0 = min_disp + num_disp; if (x0 < 0) x0 = 0;
    int w = W - x0, h = H;
    int16_t outside = (int16_t)(16 * (min_disp - 1));
    for (long i = 0; i < (long)W * H; i++) out[i] = outside;
    if (conf_out) memset(conf_out, 0, sizeof(float) * (size_t)W * H);
    if (w <= 0) return;
    size_t n = (size_t)w * h;
    float* cl = (float*)malloc(sizeof(float) * n);
    float* cr = (float*)malloc(sizeof(float) * n);
    discontinuity_map(dl, W, x0, w, h, dd_radius, variant, cl); /* left ROI  = columns [x0, W)   */
    discontinuity_map(dr, W, 0, w, h, dd_radius, variant, cr);  /* right ROI = columns [0, W-x0) */
    float* conf = (float*)malloc(sizeof(float) * n);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int j = x0 + x;
            int l = dl[(long)y * W + j];
            int ridx = j - (l >> 4);
            float c = cl[(long)y * w + x]; /* destination aliases the left map */
            if (ridx >= 0 && ridx < w) {
                int rr = dr[(long)y * W + ridx];
                if (abs(l + rr) < lrc_thresh) {
                    float c2 = cr[(long)y * w + ridx];
                    c = c < c2 ? c : c2;
                } else c = 0.0f;
            } else if (variant & ORC_WLS_LRC_OUTSIDE_ZERO) c = 0.0f;
            conf[(long)y * w + x] = 255.0f * c;
        }
    /* guide weights */
    float* lut = (float*)malloc(sizeof(float) * 65536);
    for (int i = 0; i < 65536; i++) lut[i] = (float)(-exp(-sqrt((double)(float)i) / sigma_color));
    float* ch = (float*)malloc(sizeof(float) * n);
    float* cv = (float*)malloc(sizeof(float) * n);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int g = guide[(long)y * W + x0 + x];
            int gx = x < w - 1 ? guide[(long)y * W + x0 + x + 1] : g;
            int gy = y < h - 1 ? guide[(long)(y + 1) * W + x0 + x] : g;
            ch[(long)y * w + x] = x < w - 1 ? lut[(g - gx) * (g - gx)] : 0.0f;
            cv[(long)y * w + x] = y < h - 1 ? lut[(g - gy) * (g - gy)] : 0.0f;
        }
    float* num = (float*)malloc(sizeof(float) * n);
    float* den = (float*)malloc(sizeof(float) * n);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            long i = (long)y * w + x;
            num[i] = conf[i] * (float)dl[(long)y * W + x0 + x];
            den[i] = conf[i];
        }
    fgs_solve(num, w, h, ch, cv, lambda, variant);
    fgs_solve(den, w, h, ch, cv, lambda, variant);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            long i = (long)y * w + x;
            float v = num[i] * (1.0f / (den[i] + 1e-43f));
            long q = (fabsf(v) < 2147483648.0f) ? lrintf(v) : -32768; /* half-even, saturate; NaN or |v|>=2^31 -> INT_MIN saturated (x86 cvRound) */
            if (q > 32767) q = 32767; if (q < -32768) q = -32768;
            out[(long)y * W + x0 + x] = (int16_t)q;
            if (conf_out) conf_out[(long)y * W + x0 + x] = conf[i];
        }
    free(cl); free(cr); free(conf); free(lut); free(ch); free(cv); free(num); free(den);
}
