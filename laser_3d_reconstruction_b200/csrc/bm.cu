// bm.cu -- cv2.StereoBM.compute on sm_100a, bit-exact (the matcher the reference's README offers as a drop-in for
// StereoSGBM, readme.md:392-397).  Arithmetic as restated and pinned in oracle/csrc/orc_bm.c:
//   prefilter  XSOBEL: clip(3x3 x-Sobel, +-cap) + cap with rows reflected at the image border; columns 0 and W-1, the
//              last row of an odd-height image and every row of a one-row image hold `cap`
//   matching   SAD over blockSize^2 between left(x) and right(x - d); OpenCV indexes the right row by
//              d' = numDisparities - 1 - (d - minDisparity) ascending and keeps the FIRST minimum in that order
//   tests      texture (sum of |pixel - cap| over the window < textureThreshold), uniqueness (another d' more than one
//              step away with SAD <= minSAD + minSAD * ratio / 100)
//   sub-pixel  ((ndisp - d' - 1 + minD) * 256 + (p - n) * 256 / (p + n - 2 minSAD + |p - n|) + 15) >> 4, neighbours
//              mirrored at the ends of the range
//   outside the valid rectangle [maxD + r, W - r) x [r, H - r): (minDisparity - 1) * 16; then filterSpeckles.
//   disp12MaxDiff >= 0: cv2.validateDisparity between matching and masking (bm_claim_kernel / bm_validate_kernel).
// Only pixels inside the valid rectangle are computed (everything else is the constant), so window rows never clamp;
// the column clamps of OpenCV's sliding sums (left: [0, W-1], right: [0, W - ndisp]) are kept.
#include "common.cuh"

namespace l3d {

__global__ void bm_prefilter_kernel(const uint8_t* __restrict__ src, int W, int H, int cap, uint8_t* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    int v = cap;
    const bool flat_row = H == 1 || ((H & 1) && y == H - 1);
    if (!flat_row && x > 0 && x < W - 1) {
        const uint8_t* r1 = src + (size_t)y * W;
        const uint8_t* r0 = y > 0 ? r1 - W : r1 + W;
        const uint8_t* r2 = y < H - 1 ? r1 + W : r1 - W;
        const int g = ((int)r0[x + 1] - (int)r0[x - 1]) + 2 * ((int)r1[x + 1] - (int)r1[x - 1]) + ((int)r2[x + 1] - (int)r2[x - 1]);
        v = min(max(g, -cap), cap) + cap;
    }
    dst[(size_t)y * W + x] = (uint8_t)v;
}

struct BmArgs {
    const uint8_t* L; const uint8_t* R; int16_t* disp;
    int* cost;  // min SAD per computed pixel (only for the left-right check), else nullptr
    int W, H, ndisp, mindisp, wsz, lofs, rofs, cap, tex, uniq;
    int xmin, xmax, ymin, ymax;  // valid rectangle (image coordinates)
    int filtered;
};

constexpr int BM_RUN = 64;  // output columns per warp

// a warp owns BM_RUN consecutive output pixels of one row; lane l holds the SADs of d' = l, l + 32, ... (NK of them)
template <int NK>
__global__ void __launch_bounds__(128) bm_match_kernel(const BmArgs a) {
    const int lane = threadIdx.x & 31;
    const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int ncols = a.xmax - a.xmin, nruns = (ncols + BM_RUN - 1) / BM_RUN;
    const int nrows = a.ymax - a.ymin;
    if (wid >= nruns * nrows) return;
    const int y = a.ymin + wid / nruns;
    const int xo0 = a.xmin + (wid % nruns) * BM_RUN, xo1 = min(xo0 + BM_RUN, a.xmax);
    const int W = a.W, wsz2 = a.wsz >> 1, ndisp = a.ndisp;
    const uint8_t* Lb = a.L + (size_t)(y - wsz2) * W;  // top row of the window
    const uint8_t* Rb = a.R + (size_t)(y - wsz2) * W;
    bool dv[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) dv[k] = lane + 32 * k < ndisp;
    // column sums of window column xw (x coordinates relative to lofs, as in OpenCV's loop): added with sign s
    int sad[NK], tsum = 0;
#pragma unroll
    for (int k = 0; k < NK; k++) sad[k] = 0;
    auto add_col = [&](int xw, int s) {
        const int cl = min(max(xw + a.lofs, 0), W - 1);
        const int cr = min(max(xw + a.rofs, 0), W - ndisp);
        const uint8_t* lp = Lb + cl;
        const uint8_t* rp = Rb + cr + lane;
        for (int r = 0; r < a.wsz; r++, lp += W, rp += W) {
            const int lv = __ldg(lp);
            tsum += s * abs(lv - a.cap);
#pragma unroll
            for (int k = 0; k < NK; k++)
                if (dv[k]) sad[k] += s * abs(lv - (int)__ldg(rp + 32 * k));
        }
    };
    const int x0 = xo0 - a.lofs;
    for (int dx = -wsz2; dx <= wsz2; dx++) add_col(x0 + dx, +1);
    for (int xo = xo0; xo < xo1; xo++) {
        const int x = xo - a.lofs;
        if (xo > xo0) { add_col(x + wsz2, +1); add_col(x - wsz2 - 1, -1); }
        // first minimum in ascending d': key = SAD << 9 | d'
        unsigned key = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < NK; k++)
            if (dv[k]) key = min(key, ((unsigned)sad[k] << 9) | (unsigned)(lane + 32 * k));
        key = __reduce_min_sync(0xffffffffu, key);
        const int minsad = (int)(key >> 9), mind = (int)(key & 511u);
        bool reject = tsum < a.tex;
        if (a.uniq > 0) {
            const int thresh = minsad + (minsad * a.uniq / 100);
            bool viol = false;
#pragma unroll
            for (int k = 0; k < NK; k++) {
                const int d = lane + 32 * k;
                if (dv[k] && (d < mind - 1 || d > mind + 1) && sad[k] <= thresh) viol = true;
            }
            reject = reject || __any_sync(0xffffffffu, viol);
        }
        // neighbours of the minimum (mirrored at the ends of the range)
        const int dp = mind + 1 < ndisp ? mind + 1 : ndisp - 2, dn = mind - 1 >= 0 ? mind - 1 : 1;
        int p = 0, n = 0;
#pragma unroll
        for (int k = 0; k < NK; k++) {
            const int tp = __shfl_sync(0xffffffffu, sad[k], dp & 31), tn = __shfl_sync(0xffffffffu, sad[k], dn & 31);
            if ((dp >> 5) == k) p = tp;
            if ((dn >> 5) == k) n = tn;
        }
        if (lane == 0) {
            int out = a.filtered;
            if (!reject) {
                const int d = p + n - 2 * minsad + abs(p - n);
                out = ((ndisp - mind - 1 + a.mindisp) * 256 + (d != 0 ? (p - n) * 256 / d : 0) + 15) >> 4;
            }
            a.disp[(size_t)y * W + xo] = (int16_t)out;
            if (a.cost) a.cost[(size_t)y * W + xo] = minsad;
        }
    }
}

// cv2.validateDisparity as StereoBM applies it for disp12MaxDiff >= 0 (restated in oracle/csrc/orc_bm.c::bm_validate).
// Pass 1: every valid left pixel claims its right pixel x2 = x - round(d); the smallest SAD wins, the smaller x on a
// tie (OpenCV scans x upwards with a strict '>'), carried in one 64-bit atomicMin key (cost, x, d).
__global__ void bm_claim_kernel(const int16_t* __restrict__ disp, const int* __restrict__ cost, int W, int minX1, int maxX1,
                                int invalid, unsigned long long* __restrict__ claim) {
    const int x = minX1 + blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= maxX1) return;
    const int d = disp[(size_t)y * W + x];
    if (d == invalid) return;
    const int x2 = x - ((d + 8) >> 4);
    const unsigned long long key = ((unsigned long long)(unsigned)cost[(size_t)y * W + x] << 32) | ((unsigned long long)x << 16) |
                                   (unsigned long long)(unsigned short)d;
    atomicMin(claim + (size_t)y * W + x2, key);
}
// Pass 2: a disparity is dropped when BOTH of its integer roundings point at a claimed right pixel whose claimant's
// disparity differs by more than the tolerance.
__global__ void bm_validate_kernel(int16_t* __restrict__ disp, const unsigned long long* __restrict__ claim, int W, int minX1,
                                   int maxX1, int invalid, int tol16) {
    const int x = minX1 + blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= maxX1) return;
    const int d = disp[(size_t)y * W + x];
    if (d == invalid) return;
    const int d0 = d >> 4, d1 = (d + 15) >> 4;
    const int x0 = x - d0, x1 = x - d1;
    bool bad = true;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int xx = k ? x1 : x0;
        bool c = false;
        if (0 <= xx && xx < W) {
            const unsigned long long key = claim[(size_t)y * W + xx];
            if (key != ~0ull) {
                const int d2 = (int)(short)(key & 0xffffull);
                c = d2 > invalid && abs(d2 - d) > tol16;
            }
        }
        bad = bad && c;
    }
    if (bad) disp[(size_t)y * W + x] = (int16_t)invalid;  // claims carry their own copy of d: no read-after-write hazard
}
__global__ void bm_roi_kernel(int16_t* __restrict__ disp, int W, int H, int xmin, int xmax, int ymin, int ymax, int16_t v) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < W && (x < xmin || x >= xmax || y < ymin || y >= ymax)) disp[(size_t)y * W + x] = v;
}

__global__ void bm_fill_kernel(int16_t* p, size_t n, int16_t v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

int dev_bm(Lane& L, const l3d_bm_params& p, const uint8_t* left, const uint8_t* right, int W, int H, int16_t* disp) {
    L3D_ARG(L, p.numDisparities >= 16 && p.numDisparities <= 256 && p.numDisparities % 16 == 0, "StereoBM numDisparities");
    L3D_ARG(L, p.blockSize >= 5 && p.blockSize <= 255 && (p.blockSize & 1) && p.blockSize <= W && p.blockSize <= H, "StereoBM blockSize");
    L3D_ARG(L, p.preFilterCap >= 1 && p.preFilterCap <= 63, "StereoBM preFilterCap");
    if (p.minDisparity > 0) {
        set_err(L.err, "StereoBM: minDisparity > 0 is not supported");
        return L3D_ERR_UNSUPPORTED;
    }
    const bool lrc = p.disp12MaxDiff >= 0;
    if (lrc && (p.preFilterCap > 31 || p.blockSize > 21)) {
        set_err(L.err, "StereoBM: disp12MaxDiff >= 0 is supported for preFilterCap <= 31 and blockSize <= 21 (OpenCV's 16-bit path)");
        return L3D_ERR_UNSUPPORTED;
    }
    const size_t n = (size_t)W * H;
    const int filtered = (p.minDisparity - 1) * 16;
    L3D_LAUNCH(L, bm_fill_kernel, cdiv(n, 256), 256, 0, disp, n, (int16_t)filtered);
    BmArgs a;
    a.ndisp = p.numDisparities; a.mindisp = p.minDisparity; a.wsz = p.blockSize;
    a.lofs = std::max(a.ndisp - 1 + a.mindisp, 0); a.rofs = -std::min(a.ndisp - 1 + a.mindisp, 0);
    const int width1 = W - a.rofs - a.ndisp + 1;
    const int r = p.blockSize / 2, maxD = p.minDisparity + p.numDisparities - 1;
    a.xmin = std::max(0, maxD) + r; a.xmax = W - r; a.ymin = r; a.ymax = H - r;
    const int rxmin = a.xmin, rxmax = a.xmax, rymin = a.ymin, rymax = a.ymax;  // the valid rectangle proper
    // OpenCV only computes columns [lofs, lofs + width1); with the left-right check the columns of that range outside
    // the valid rectangle matter too (their claims count), so they are computed and masked afterwards
    if (lrc) { a.xmin = a.lofs; a.xmax = a.lofs + width1; }
    a.xmin = std::max(a.xmin, a.lofs); a.xmax = std::min(a.xmax, a.lofs + width1);
    if (a.lofs >= W || a.rofs >= W || width1 < 1 || a.xmax <= a.xmin || a.ymax <= a.ymin) return L3D_OK;
    uint8_t* lf = L.get<uint8_t>(S_IO_D, n);
    uint8_t* rf = L.get<uint8_t>(S_IO_E, n);
    dim3 pg(cdiv(W, 128), H);
    L3D_LAUNCH(L, bm_prefilter_kernel, pg, 128, 0, left, W, H, p.preFilterCap, lf);
    L3D_LAUNCH(L, bm_prefilter_kernel, pg, 128, 0, right, W, H, p.preFilterCap, rf);
    a.L = lf; a.R = rf; a.disp = disp; a.W = W; a.H = H; a.cap = p.preFilterCap; a.tex = p.textureThreshold; a.uniq = p.uniquenessRatio;
    a.filtered = filtered;
    a.cost = lrc ? L.get<int>(S_LABEL, n) : nullptr;
    const int nwarps = cdiv(a.xmax - a.xmin, BM_RUN) * (a.ymax - a.ymin);
    const int nk = cdiv(a.ndisp, 32);
    const int grid = cdiv(nwarps, 4);
    switch (nk) {
        case 1: L3D_LAUNCH(L, bm_match_kernel<1>, grid, 128, 0, a); break;
        case 2: L3D_LAUNCH(L, bm_match_kernel<2>, grid, 128, 0, a); break;
        case 3: L3D_LAUNCH(L, bm_match_kernel<3>, grid, 128, 0, a); break;
        case 4: L3D_LAUNCH(L, bm_match_kernel<4>, grid, 128, 0, a); break;
        case 5: L3D_LAUNCH(L, bm_match_kernel<5>, grid, 128, 0, a); break;
        case 6: L3D_LAUNCH(L, bm_match_kernel<6>, grid, 128, 0, a); break;
        case 7: L3D_LAUNCH(L, bm_match_kernel<7>, grid, 128, 0, a); break;
        default: L3D_LAUNCH(L, bm_match_kernel<8>, grid, 128, 0, a); break;
    }
    if (lrc) {
        const int maxD = p.minDisparity + p.numDisparities;
        const int minX1 = std::max(maxD, 0), maxX1 = W + std::min(p.minDisparity, 0);
        if (maxX1 > minX1) {
            unsigned long long* claim = L.get<unsigned long long>(S_IO_A, n);
            L3D_CHECK(L, cudaMemsetAsync(claim, 0xff, n * sizeof(unsigned long long), L.stream));
            dim3 vg(cdiv(maxX1 - minX1, 128), H);
            L3D_LAUNCH(L, bm_claim_kernel, vg, 128, 0, disp, a.cost, W, minX1, maxX1, filtered, claim);
            L3D_LAUNCH(L, bm_validate_kernel, vg, 128, 0, disp, claim, W, minX1, maxX1, filtered, p.disp12MaxDiff * 16);
        }
        L3D_LAUNCH(L, bm_roi_kernel, dim3(cdiv(W, 128), H), 128, 0, disp, W, H, rxmin, rxmax, rymin, rymax, (int16_t)filtered);
    }
    if (p.speckleRange >= 0 && p.speckleWindowSize > 0)
        return dev_speckles(L, disp, W, H, filtered, p.speckleWindowSize, p.speckleRange);
    return L3D_OK;
}

}  // namespace l3d
