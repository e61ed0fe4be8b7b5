// laser.cu -- K4: laser centre-line extraction.
//   Simple : SimpleLaserExtractor.extract_centerline        core/laser_extractor.py:45-100
//   Steger : FastStegerExtractor.extract_centerline         core/laser_extractor.py:160-261
//            ImprovedStegerExtractor.extract_centerline     improved_steger.py:39-126
//            ImprovedStegerExtractor.extract_centerline_optimized  improved_steger.py:128-223
//            HybridLaserExtractor.extract_centerline        improved_steger.py:250-344
// Integer stages (HSV, gray, masks, morphology, contour-area model, centroid sums) are bit-exact
// with cv2; the Steger stages are f32 (separable Gaussian with shared-memory halos, closed-form
// 2x2 Hessian eigen-analysis) and agree with the reference within the stated tolerance.
#include "ccl.cuh"
#include "common.cuh"

namespace l3d {

__device__ __forceinline__ int refl101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

__device__ __forceinline__ int gray_bgr(int b, int g, int r) { return (3735 * b + 19235 * g + 9798 * r + 16384) >> 15; }

// cv2 8-bit HSV (H in [0,180)): 12-bit reciprocal tables, computed on the fly with the same rounding
__device__ __forceinline__ void hsv_bgr(int b, int g, int r, int& h, int& s, int& v) {
    v = max(max(b, g), r);
    int m = min(min(b, g), r), diff = v - m;
    // sdiv = round((255<<12)/v), hdiv = round((180<<12)/(6*diff)) -- half-to-even never hit (odd numerators)
    int sdiv = v ? __double2int_rn((double)(255 << 12) / (double)v) : 0;
    int hdiv = diff ? __double2int_rn((double)(180 << 12) / (6.0 * (double)diff)) : 0;
    s = (diff * sdiv + (1 << 11)) >> 12;
    int hh;
    if (v == r) hh = g - b;
    else if (v == g) hh = b - r + 2 * diff;
    else hh = r - g + 4 * diff;
    hh = (hh * hdiv + (1 << 11)) >> 12;
    if (hh < 0) hh += 180;
    h = hh;
}

struct HsvRange { int lo[3], hi[3]; };

// mask0 = inRange(HSV) & (gray > thr)   (0/1); optionally also the u8 gray image
__global__ void colour_mask_kernel(const uint8_t* __restrict__ bgr, int n, HsvRange rg, int thr,
                                   uint8_t* __restrict__ mask, uint8_t* __restrict__ gray) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
    int h, s, v;
    hsv_bgr(b, g, r, h, s, v);
    int gy = gray_bgr(b, g, r);
    bool ok = h >= rg.lo[0] && h <= rg.hi[0] && s >= rg.lo[1] && s <= rg.hi[1] && v >= rg.lo[2] && v <= rg.hi[2] && gy > thr;
    mask[i] = ok ? 1 : 0;
    if (gray) gray[i] = (uint8_t)gy;
}

// 3x3 rectangular dilate (outside = 0) / erode (outside ignored) on 0/1 masks
template <bool DILATE>
__global__ void morph3_kernel(const uint8_t* __restrict__ src, int W, int H, uint8_t* __restrict__ dst) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    int v = DILATE ? 0 : 1;
    for (int dy = -1; dy <= 1; dy++) {
        int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = -1; dx <= 1; dx++) {
            int xx = x + dx;
            if (xx < 0 || xx >= W) continue;
            int s = src[(size_t)yy * W + xx];
            if (DILATE) v |= s; else v &= s;
        }
    }
    dst[(size_t)y * W + x] = (uint8_t)v;
}

// MORPH_CLOSE then MORPH_OPEN (core/laser_extractor.py:67-69); result in `a`, `b` is scratch
static int close_open(Lane& L, uint8_t* a, uint8_t* b, int W, int H) {
    dim3 g(cdiv(W, 128), H);
    L3D_LAUNCH(L, morph3_kernel<true>, g, 128, 0, a, W, H, b);
    L3D_LAUNCH(L, morph3_kernel<false>, g, 128, 0, b, W, H, a);
    L3D_LAUNCH(L, morph3_kernel<false>, g, 128, 0, a, W, H, b);
    L3D_LAUNCH(L, morph3_kernel<true>, g, 128, 0, b, W, H, a);
    return L3D_OK;
}

// ---- contour model: findContours(RETR_EXTERNAL) + contourArea > min_area + filled drawContours
struct BgRule {  // 4-connected background
    const uint8_t* m;
    __device__ bool node(int i) const { return m[i] == 0; }
    __device__ bool edge(int, int) const { return true; }
};
struct FgRule {  // 8-connected filled foreground
    const uint8_t* m;
    __device__ bool node(int i) const { return m[i] != 0; }
    __device__ bool edge(int, int) const { return true; }
};

__global__ void border_flag_kernel(const int* __restrict__ label, int W, int H, int* __restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = 2 * (W + H);
    if (i >= n) return;
    int x, y;
    if (i < W) { x = i; y = 0; }
    else if (i < 2 * W) { x = i - W; y = H - 1; }
    else if (i < 2 * W + H) { x = 0; y = i - 2 * W; }
    else { x = W - 1; y = i - 2 * W - H; }
    int l = label[y * W + x];
    if (l >= 0) flag[l] = 1;
}

__global__ void fill_holes_kernel(const uint8_t* __restrict__ mask, const int* __restrict__ label,
                                  const int* __restrict__ flag, int n, uint8_t* __restrict__ filled) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    filled[i] = (mask[i] || !flag[label[i]]) ? 1 : 0;
}

// twice the contour area per component: 2 per full 2x2 window, 1 per window with three set pixels
__global__ void quad_area_kernel(const uint8_t* __restrict__ f, const int* __restrict__ label, int W, int H,
                                 int* __restrict__ area2) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W - 1 || y >= H - 1) return;
    int q = y * W + x;
    int c = f[q] + f[q + 1] + f[q + W] + f[q + W + 1];
    if (c < 3) return;
    int l = f[q] ? label[q] : label[q + 1];
    atomicAdd(&area2[l], c == 4 ? 2 : 1);
}

__global__ void final_mask_kernel(const uint8_t* __restrict__ f, const int* __restrict__ label,
                                  const int* __restrict__ area2, int n, double min_area, uint8_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = (f[i] && 0.5 * (double)area2[label[i]] > min_area) ? 1 : 0;
}

// per row: sum x*gray / sum gray over mask pixels (exact integers, one f64 divide)
__global__ void row_centroid_kernel(const uint8_t* __restrict__ mask, const uint8_t* __restrict__ gray, int W,
                                    double* __restrict__ rowx, int* __restrict__ rowcnt) {
    int y = blockIdx.x;
    unsigned long long sxg = 0, sg = 0;
    int cnt = 0;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        size_t i = (size_t)y * W + x;
        if (mask[i]) { unsigned g = gray[i]; sxg += (unsigned long long)x * g; sg += g; cnt++; }
    }
    __shared__ unsigned long long s1[128], s2[128];
    __shared__ int s3[128];
    s1[threadIdx.x] = sxg; s2[threadIdx.x] = sg; s3[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; s3[threadIdx.x] += s3[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        bool ok = s3[0] > 0 && s2[0] > 0;
        rowcnt[y] = ok ? 1 : 0;
        rowx[y] = ok ? (double)s1[0] / (double)s2[0] : 0.0;
    }
}

// exclusive scan of per-row counts (single block); total -> *n
__global__ void row_scan_kernel(const int* __restrict__ cnt, int H, int* __restrict__ off, int* __restrict__ n) {
    __shared__ int part[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < H; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < H ? cnt[i] : 0;
        part[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int t = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
            __syncthreads();
            part[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < H) off[i] = carry + part[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += part[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *n = carry;
}

__global__ void simple_emit_kernel(const double* __restrict__ rowx, const int* __restrict__ cnt,
                                   const int* __restrict__ off, int H, double* __restrict__ xy) {
    int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= H || !cnt[y]) return;
    xy[2 * off[y]] = rowx[y];
    xy[2 * off[y] + 1] = (double)y;
}

__global__ void mask_to_255_kernel(const uint8_t* __restrict__ m, int n, uint8_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = m[i] ? 255 : 0;
}

int dev_simple(Lane& L, const uint8_t* bgr, int W, int H, const int* lo, const int* hi, int thr,
               double min_area, uint8_t* mask_morph, uint8_t* mask_final, double* xy, int* n_dev) {
    int n = W * H;
    uint8_t* m0 = L.get<uint8_t>(S_SM_MASK0, (size_t)n * 3);  // mask | scratch | gray
    uint8_t* m1 = m0 + n;
    uint8_t* gray = m0 + 2 * (size_t)n;
    uint8_t* filled = L.get<uint8_t>(S_SM_MASK1, (size_t)n * 2);
    uint8_t* fin = filled + n;
    int* label = L.get<int>(S_SM_LABEL, n);
    int* aux = L.get<int>(S_SM_AREA, (size_t)n + 2 * H + 8);  // flag / area2, then row counts, offsets
    int* rowcnt = aux + n;
    int* rowoff = rowcnt + H;
    double* rowx = L.get<double>(S_SM_ROW, H);
    HsvRange rg;
    for (int k = 0; k < 3; k++) { rg.lo[k] = lo[k]; rg.hi[k] = hi[k]; }
    L3D_LAUNCH(L, colour_mask_kernel, cdiv(n, 256), 256, 0, bgr, n, rg, thr, m0, gray);
    int rc = close_open(L, m0, m1, W, H);
    if (rc != L3D_OK) return rc;
    if (mask_morph) L3D_LAUNCH(L, mask_to_255_kernel, cdiv(n, 256), 256, 0, m0, n, mask_morph);
    // background reachable from the border
    rc = ccl_label<BgRule, false>(L, BgRule{m0}, W, H, label);
    if (rc != L3D_OK) return rc;
    L3D_CHECK(L, cudaMemsetAsync(aux, 0, sizeof(int) * n, L.stream));
    L3D_LAUNCH(L, border_flag_kernel, cdiv(2 * (W + H), 256), 256, 0, label, W, H, aux);
    L3D_LAUNCH(L, fill_holes_kernel, cdiv(n, 256), 256, 0, m0, label, aux, n, filled);
    // 8-connected components of the filled mask, contour area by 2x2 windows
    rc = ccl_label<FgRule, true>(L, FgRule{filled}, W, H, label);
    if (rc != L3D_OK) return rc;
    L3D_CHECK(L, cudaMemsetAsync(aux, 0, sizeof(int) * n, L.stream));
    L3D_LAUNCH(L, quad_area_kernel, dim3(cdiv(W, 128), H), 128, 0, filled, label, W, H, aux);
    L3D_LAUNCH(L, final_mask_kernel, cdiv(n, 256), 256, 0, filled, label, aux, n, min_area, fin);
    if (mask_final) L3D_LAUNCH(L, mask_to_255_kernel, cdiv(n, 256), 256, 0, fin, n, mask_final);
    L3D_LAUNCH(L, row_centroid_kernel, H, 128, 0, fin, gray, W, rowx, rowcnt);
    L3D_LAUNCH(L, row_scan_kernel, 1, 1024, 0, rowcnt, H, rowoff, n_dev);
    L3D_LAUNCH(L, simple_emit_kernel, cdiv(H, 128), 128, 0, rowx, rowcnt, rowoff, H, xy);
    return L3D_OK;
}

// steps (1)-(3) of SimpleLaserExtractor.extract_centerline alone (core/laser_extractor.py:56-64): inRange(HSV) & (gray > thr)
// as a 0/255 mask -- the exhaustive colour-cube tests go through this
int dev_colour_mask(Lane& L, const uint8_t* bgr, int W, int H, const int* lo, const int* hi, int thr, uint8_t* mask255) {
    const size_t n = (size_t)W * H;
    uint8_t* m0 = L.get<uint8_t>(S_SM_MASK0, n);
    HsvRange rg;
    for (int k = 0; k < 3; k++) { rg.lo[k] = lo[k]; rg.hi[k] = hi[k]; }
    L3D_LAUNCH(L, colour_mask_kernel, cdiv(n, 256), 256, 0, bgr, (int)n, rg, thr, m0, (uint8_t*)nullptr);
    L3D_LAUNCH(L, mask_to_255_kernel, cdiv(n, 256), 256, 0, m0, (int)n, mask255);
    return L3D_OK;
}

// ============================================================================================
// Steger
// ============================================================================================
constexpr int MAX_TAPS = 129;
struct Taps { int r; float k[MAX_TAPS]; };

// u8 gray (+ f32 copy) of the working image.  channels 3: BGR->gray; 1: copy.  `mask` (optional,
// Hybrid): pixels outside the mask become 0 before the conversion (roi_image[clean_mask==0] = 0).
__global__ void steger_gray_kernel(const uint8_t* __restrict__ img, int channels, long stride, int w, int h,
                                   const uint8_t* __restrict__ mask, long mstride, uint8_t* __restrict__ g8,
                                   float* __restrict__ gf) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    const uint8_t* p = img + (long)y * stride + (long)x * channels;
    int g = channels == 3 ? gray_bgr(p[0], p[1], p[2]) : p[0];
    if (mask && !mask[(long)y * mstride + x]) g = 0;
    size_t i = (size_t)y * w + x;
    g8[i] = (uint8_t)g;
    gf[i] = (float)g;
}

// separable Gaussian, BORDER_REFLECT_101: row pass.  One block = one 256-pixel row segment; the
// segment and its r-pixel halos are staged in shared memory.
__global__ void gauss_row_kernel(const float* __restrict__ src, int w, int h, Taps t, float* __restrict__ dst) {
    extern __shared__ float sm[];
    const int r = t.r;
    int x0 = blockIdx.x * blockDim.x, y = blockIdx.y;
    const float* row = src + (size_t)y * w;
    for (int i = threadIdx.x; i < (int)blockDim.x + 2 * r; i += blockDim.x) sm[i] = row[refl101(x0 - r + i, w)];
    __syncthreads();
    int x = x0 + threadIdx.x;
    if (x >= w) return;
    const float* c = sm + threadIdx.x + r;
    float s = __fmul_rn(t.k[r], c[0]);
    for (int j = 1; j <= r; j++) s = __fadd_rn(s, __fmul_rn(t.k[r + j], __fadd_rn(c[-j], c[j])));
    dst[(size_t)y * w + x] = s;
}

// column pass: block = 32 columns x 32 rows, halo rows staged in shared memory
__global__ void gauss_col_kernel(const float* __restrict__ src, int w, int h, Taps t, float* __restrict__ dst) {
    extern __shared__ float sm[];  // [(32 + 2r)][32]
    const int r = t.r;
    int x = blockIdx.x * 32 + threadIdx.x;
    int y0 = blockIdx.y * 32;
    int xc = min(x, w - 1);
    for (int i = threadIdx.y; i < 32 + 2 * r; i += blockDim.y) sm[i * 32 + threadIdx.x] = src[(size_t)refl101(y0 - r + i, h) * w + xc];
    __syncthreads();
    for (int yy = threadIdx.y; yy < 32; yy += blockDim.y) {
        int y = y0 + yy;
        if (x >= w || y >= h) continue;
        const float* c = sm + (yy + r) * 32 + threadIdx.x;
        float s = __fmul_rn(t.k[r], c[0]);
        for (int j = 1; j <= r; j++) s = __fadd_rn(s, __fmul_rn(t.k[r + j], __fadd_rn(c[-j * 32], c[j * 32])));
        dst[(size_t)y * w + x] = s;
    }
}

struct Img { const float* p; int w, h; };
__device__ __forceinline__ float px(const Img& I, int y, int x) { return I.p[(size_t)refl101(y, I.h) * I.w + refl101(x, I.w)]; }

// cv2's symmetric [1,2,1] tap order in f32: (a + c) + 2b  (pinned by tests/test_oracle_image.py)
__device__ __forceinline__ float smooth121(float a, float b, float c) { return __fadd_rn(__fadd_rn(a, c), __fmul_rn(2.0f, b)); }
// cv2.Sobel(ksize=3) first derivatives of the smoothed image (REFLECT_101 at every stage)
__device__ __forceinline__ float sobel_dx(const Img& I, int y, int x) {
    float a = __fsub_rn(px(I, y - 1, x + 1), px(I, y - 1, x - 1));
    float b = __fsub_rn(px(I, y, x + 1), px(I, y, x - 1));
    float c = __fsub_rn(px(I, y + 1, x + 1), px(I, y + 1, x - 1));
    return smooth121(a, b, c);
}
__device__ __forceinline__ float sobel_dy(const Img& I, int y, int x) {
    float lo = smooth121(px(I, y - 1, x - 1), px(I, y - 1, x), px(I, y - 1, x + 1));
    float hi = smooth121(px(I, y + 1, x - 1), px(I, y + 1, x), px(I, y + 1, x + 1));
    return __fsub_rn(hi, lo);
}
// the derivative images are themselves border-reflected when differentiated again
__device__ __forceinline__ float dxr(const Img& I, int y, int x) { return sobel_dx(I, refl101(y, I.h), refl101(x, I.w)); }
__device__ __forceinline__ float dyr(const Img& I, int y, int x) { return sobel_dy(I, refl101(y, I.h), refl101(x, I.w)); }

struct Hess { float dx, dy, dxx, dyy, dxy; };

__device__ __forceinline__ Hess hess_sobel(const Img& I, int y, int x) {
    Hess H;
    H.dx = sobel_dx(I, y, x);
    H.dy = sobel_dy(I, y, x);
    {   // dxx = Sobel_x(dx)
        float a = __fsub_rn(dxr(I, y - 1, x + 1), dxr(I, y - 1, x - 1));
        float b = __fsub_rn(dxr(I, y, x + 1), dxr(I, y, x - 1));
        float c = __fsub_rn(dxr(I, y + 1, x + 1), dxr(I, y + 1, x - 1));
        H.dxx = smooth121(a, b, c);
    }
    {   // dyy = Sobel_y(dy)
        float lo = smooth121(dyr(I, y - 1, x - 1), dyr(I, y - 1, x), dyr(I, y - 1, x + 1));
        float hi = smooth121(dyr(I, y + 1, x - 1), dyr(I, y + 1, x), dyr(I, y + 1, x + 1));
        H.dyy = __fsub_rn(hi, lo);
    }
    {   // dxy = Sobel_y(dx)
        float lo = smooth121(dxr(I, y - 1, x - 1), dxr(I, y - 1, x), dxr(I, y - 1, x + 1));
        float hi = smooth121(dxr(I, y + 1, x - 1), dxr(I, y + 1, x), dxr(I, y + 1, x + 1));
        H.dxy = __fsub_rn(hi, lo);
    }
    return H;
}

// FastSteger: cv2.filter2D differences (correlation, anchor ksize/2, REFLECT_101)
__device__ __forceinline__ Hess hess_diff(const Img& I, int y, int x) {
    Hess H;
    float c = px(I, y, x), l = px(I, y, x - 1), r = px(I, y, x + 1), u = px(I, y - 1, x), d = px(I, y + 1, x);
    H.dx = __fsub_rn(l, c);
    H.dy = __fsub_rn(u, c);
    H.dxx = __fadd_rn(__fadd_rn(l, __fmul_rn(-2.0f, c)), r);
    H.dyy = __fadd_rn(__fadd_rn(u, __fmul_rn(-2.0f, c)), d);
    H.dxy = __fadd_rn(__fsub_rn(__fsub_rn(px(I, y - 1, x - 1), u), l), c);
    return H;
}

// eigen-decomposition of [[a,b],[b,c]]: eigenvalue of larger magnitude and its unit eigenvector
__device__ __forceinline__ void eig2(float a, float b, float c, float& lam, float& nx, float& ny) {
    float hm = 0.5f * (a + c), hd = 0.5f * (a - c);
    float rad = sqrtf(hd * hd + b * b);
    float l1 = hm + rad, l2 = hm - rad;
    lam = fabsf(l1) >= fabsf(l2) ? l1 : l2;
    float v1x = b, v1y = lam - a, v2x = lam - c, v2y = b;
    float n1 = v1x * v1x + v1y * v1y, n2 = v2x * v2x + v2y * v2y;
    float vx = n1 >= n2 ? v1x : v2x, vy = n1 >= n2 ? v1y : v2y, nn = fmaxf(n1, n2);
    if (nn > 0.f) { float inv = rsqrtf(nn); nx = vx * inv; ny = vy * inv; }
    else if (fabsf(a) >= fabsf(c)) { nx = 1.f; ny = 0.f; }
    else { nx = 0.f; ny = 1.f; }
}

struct StegerEval {
    int variant, thr;
    float resp;
    int w, h, offx, offy;
};

// candidate test of one pixel; returns true and the sub-pixel point / response
__device__ __forceinline__ bool steger_pixel(const StegerEval& e, const Img& I, const uint8_t* __restrict__ g8,
                                             const uint8_t* __restrict__ mask, int y, int x, float& ox_out,
                                             float& oy_out, float& response) {
    size_t i = (size_t)y * e.w + x;
    if (e.variant == L3D_STEGER_HYBRID) { if (!mask[i]) return false; }
    else if ((int)g8[i] <= e.thr) return false;
    if (e.variant != L3D_STEGER_FAST && (x < 1 || x > e.w - 2 || y < 1 || y > e.h - 2)) return false;
    Hess H = e.variant == L3D_STEGER_FAST ? hess_diff(I, y, x) : hess_sobel(I, y, x);
    float lam, nx, ny;
    eig2(H.dxx, H.dxy, H.dyy, lam, nx, ny);
    if (e.variant != L3D_STEGER_FAST && lam >= 0.f) return false;
    float den = nx * nx * H.dxx + 2.f * nx * ny * H.dxy + ny * ny * H.dyy;
    if (e.variant == L3D_STEGER_FAST) { if (!(fabsf(den) > 1e-10f)) return false; }
    else if (fabsf(den) < 1e-6f) return false;
    float t = -(nx * H.dx + ny * H.dy) / den;
    float ox = t * nx, oy = t * ny;
    if (e.variant == L3D_STEGER_HYBRID) { if (!(fabsf(ox) <= 0.5f)) return false; }
    else if (!(fabsf(ox) <= e.resp && fabsf(oy) <= e.resp)) return false;
    ox_out = ox; oy_out = oy; response = fabsf(lam);
    return true;
}

// one block per image row; points appended in x order (raster order overall)
__global__ void steger_rows_kernel(StegerEval e, Img I, const uint8_t* __restrict__ g8, const uint8_t* __restrict__ mask,
                                   float2* __restrict__ rowbuf, int* __restrict__ rowcnt) {
    const int y = blockIdx.x;
    __shared__ int warp_cnt[8];
    __shared__ int base;
    __shared__ unsigned long long best_key;
    __shared__ float best_val;
    if (threadIdx.x == 0) { base = 0; best_key = 0ull; }
    __syncthreads();
    const bool per_row_best = e.variant == L3D_STEGER_OPTIMIZED || e.variant == L3D_STEGER_HYBRID;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int x0 = 0; x0 < e.w; x0 += blockDim.x) {
        int x = x0 + threadIdx.x;
        float ox = 0.f, oy = 0.f, resp = 0.f;
        bool ok = x < e.w && steger_pixel(e, I, g8, mask, y, x, ox, oy, resp);
        float cx = __fadd_rn((float)x, ox), cy = __fadd_rn((float)y, oy);
        if (e.variant == L3D_STEGER_IMPROVED) ok = ok && cx >= 0.f && cx < (float)e.w && cy >= 0.f && cy < (float)e.h;
        if (per_row_best) {
            // max response, first x wins ties: key = (response bits << 32) | ~x
            if (ok) atomicMax(&best_key, ((unsigned long long)__float_as_uint(resp) << 32) | (unsigned)(0xffffffffu - (unsigned)x));
            __syncthreads();
            if (ok && (unsigned)(0xffffffffu - (unsigned)x) == (unsigned)(best_key & 0xffffffffu) &&
                __float_as_uint(resp) == (unsigned)(best_key >> 32)) best_val = cx;
            __syncthreads();
        } else {
            unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (lane == 0) warp_cnt[warp] = __popc(bal);
            __syncthreads();
            int off = base;
            for (int k = 0; k < warp; k++) off += warp_cnt[k];
            off += __popc(bal & ((1u << lane) - 1));
            if (ok) rowbuf[(size_t)y * e.w + off] = make_float2(__fadd_rn(cx, (float)e.offx), __fadd_rn(cy, (float)e.offy));
            __syncthreads();
            if (threadIdx.x == 0) { int s = 0; for (int k = 0; k < nw; k++) s += warp_cnt[k]; base += s; }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        if (per_row_best) {
            // the winner of the LAST chunk that improved the key wrote best_val; re-check bounds
            bool have = best_key != 0ull && y >= 1 && y <= e.h - 2;
            if (have && best_val >= 0.f && best_val < (float)e.w) {
                rowbuf[(size_t)y * e.w] = make_float2(best_val, (float)y);
                rowcnt[y] = 1;
            } else rowcnt[y] = 0;
        } else rowcnt[y] = base;
    }
}

__global__ void steger_emit_kernel(const float2* __restrict__ rowbuf, const int* __restrict__ cnt,
                                   const int* __restrict__ off, int w, int cap, float* __restrict__ xy) {
    int y = blockIdx.x;
    int c = cnt[y], o = off[y];
    for (int k = threadIdx.x; k < c; k += blockDim.x) {
        if (o + k < cap) { float2 p = rowbuf[(size_t)y * w + k]; xy[2 * (o + k)] = p.x; xy[2 * (o + k) + 1] = p.y; }
    }
}

int dev_steger(Lane& L, const l3d_steger_params& p, const uint8_t* img, int channels, int W, int H,
               float* xy, int cap, int* n_dev) {
    L3D_ARG(L, channels == 1 || channels == 3, "steger: channels must be 1 or 3");
    L3D_ARG(L, p.variant >= 0 && p.variant <= 3, "steger: variant");
    L3D_ARG(L, p.sigma > 0, "steger: sigma");
    // working window (FastSteger ROI = numpy slice semantics: clipped to the image)
    int rx = 0, ry = 0, w = W, h = H;
    if (p.variant == L3D_STEGER_FAST && p.roi[2] > 0 && p.roi[3] > 0) {
        rx = std::min(std::max(p.roi[0], 0), W); ry = std::min(std::max(p.roi[1], 0), H);
        w = std::min(p.roi[0] + p.roi[2], W) - rx; h = std::min(p.roi[1] + p.roi[3], H) - ry;
    }
    if (w <= 0 || h <= 0) { L3D_CHECK(L, cudaMemsetAsync(n_dev, 0, sizeof(int), L.stream)); return L3D_OK; }
    int ks = ((int)lrint(p.sigma * 8 + 1)) | 1;
    L3D_ARG(L, ks <= MAX_TAPS, "steger: sigma too large");
    Taps t;
    t.r = ks / 2;
    {
        std::vector<double> kd(ks);
        double sum = 0, s2 = -0.5 / (p.sigma * p.sigma);
        for (int i = 0; i < ks; i++) { double x = i - t.r; kd[i] = exp(s2 * x * x); sum += kd[i]; }
        for (int i = 0; i < ks; i++) t.k[i] = (float)(kd[i] / sum);
    }
    size_t n = (size_t)w * h;
    uint8_t* g8 = L.get<uint8_t>(S_ST_GRAY, n * 2 + (size_t)W * H * 2);
    uint8_t* mask = nullptr;
    float* gf = L.get<float>(S_ST_TMP, n * 2);
    float* tmp = gf + n;
    float* sm = L.get<float>(S_ST_SM, n);
    const long stride = (long)W * channels;
    const uint8_t* src = img + (long)ry * stride + (long)rx * channels;
    if (p.variant == L3D_STEGER_HYBRID) {
        L3D_ARG(L, channels == 3, "hybrid extractor needs a BGR image");
        uint8_t* m0 = g8 + n * 2;
        uint8_t* m1 = m0 + (size_t)W * H;
        HsvRange rg;
        for (int k = 0; k < 3; k++) { rg.lo[k] = p.hsv_lo[k]; rg.hi[k] = p.hsv_hi[k]; }
        L3D_LAUNCH(L, colour_mask_kernel, cdiv(W * H, 256), 256, 0, img, W * H, rg, p.bright_thr, m0, (uint8_t*)nullptr);
        int rc = close_open(L, m0, m1, W, H);
        if (rc != L3D_OK) return rc;
        mask = m0;
    }
    L3D_LAUNCH(L, steger_gray_kernel, dim3(cdiv(w, 128), h), 128, 0, src, channels, stride, w, h, mask, (long)W, g8, gf);
    L3D_LAUNCH(L, gauss_row_kernel, dim3(cdiv(w, 256), h), 256, (256 + 2 * t.r) * sizeof(float), gf, w, h, t, tmp);
    L3D_LAUNCH(L, gauss_col_kernel, dim3(cdiv(w, 32), cdiv(h, 32)), dim3(32, 8), (32 + 2 * t.r) * 32 * sizeof(float), tmp, w, h, t, sm);
    StegerEval e;
    e.variant = p.variant; e.thr = p.bright_thr; e.resp = (float)p.resp_thr; e.w = w; e.h = h; e.offx = rx; e.offy = ry;
    float2* rowbuf = L.get<float2>(S_ST_ROW, n);
    int* rowcnt = L.get<int>(S_ST_CNT, (size_t)2 * h);
    int* rowoff = rowcnt + h;
    Img I{sm, w, h};
    L3D_LAUNCH(L, steger_rows_kernel, h, 128, 0, e, I, g8, mask, rowbuf, rowcnt);
    L3D_LAUNCH(L, row_scan_kernel, 1, 1024, 0, rowcnt, h, rowoff, n_dev);
    L3D_LAUNCH(L, steger_emit_kernel, h, 64, 0, rowbuf, rowcnt, rowoff, w, cap, xy);
    return L3D_OK;
}

}  // namespace l3d
