// sgbm.cu -- cv2.StereoSGBM.compute on sm_100a (modes SGBM / HH / SGBM_3WAY), bit-exact.
//
// Replaces stereo_matcher.compute / right_matcher.compute of the reference
// (camera/single_usb_stereo_camera.py:252-274 parameters, :324-325 calls).
//
// HBM layout: cost volume C and aggregated volume S are int16 [vrow][x][d] (d fastest), x in
// width1 = maxX1-minX1 coordinates.  A warp owns one pixel's disparity range: lane l holds DPL
// consecutive disparities packed as u16x2 words (DPL = 2/4/8 for D <= 64/128/256), so one warp
// access is one contiguous 64..512 B segment.
//
// Kernels:
//   sgbm_prefilter_kernel   x-Sobel clip + half-pixel min/max descriptors (8 B / pixel)
//   sgbm_cost_kernel        Birchfield-Tomasi pixel cost -> blockSize^2 box sum -> C (+P2), fused;
//                           pixel costs and the row ring of horizontal sums live in shared memory
//   sgbm_scan_kernel        one SGM path direction per launch, one warp per scan line, path
//                           state in registers, DPX u16x2 min/add, warp-wide min via CREDUX,
//                           C/S streamed through a per-lane cp.async ring
//   sgbm_wta_kernel         WTA + uniqueness + disp2 (atomicMax key) + sub-pixel
//   sgbm_lrcheck_kernel     left-right consistency
// Value domain: 0 <= L,S <= 32767 and C >= P2, which holds whenever the block sum
// does not wrap int16 (always for blockSize <= 9; for 11 unless every pixel of a block mismatches
// by more than 92 % of the maximum cost).  Inside that domain OpenCV's saturating int16 SIMD and
// the unsigned 16-bit arithmetic used here give identical bits.
#include <cuda_pipeline.h>

#include "common.cuh"

namespace l3d {

constexpr int MAXSEG = 4;
constexpr int MAXBAND = 64;
constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t INF2 = 0x7fff7fffu;

struct Geom {
    int W, H, minD, D, maxD, minX1, maxX1, width1, bs, SW2, P1, P2, uniq, d12, ftzero, mode;
    int DPL, NP, nact;
    int HV, nseg;
    int seg_vr0[MAXSEG], seg_y0[MAXSEG], seg_rows[MAXSEG], seg_emit[MAXSEG];
};

static int make_geom(const l3d_sgbm_params& p, int W, int H, Geom& g, std::string* err) {
    g.W = W; g.H = H; g.minD = p.minDisparity; g.D = p.numDisparities; g.maxD = g.minD + g.D;
    g.mode = p.mode;
    if (W < 2 || H < 1) { set_err(err, "sgbm: image too small"); return L3D_ERR_ARG; }
    if (g.D < 16 || g.D > 256 || (g.D % 16)) {
        set_err(err, "sgbm: numDisparities must be a multiple of 16 in [16,256], got %d", g.D);
        return L3D_ERR_UNSUPPORTED;
    }
    if (p.blockSize < 1 || !(p.blockSize & 1) || p.blockSize > 21) {
        set_err(err, "sgbm: blockSize must be odd in [1,21], got %d", p.blockSize);
        return L3D_ERR_UNSUPPORTED;
    }
    if (p.mode < 0 || p.mode > 2) { set_err(err, "sgbm: mode %d unsupported (0,1,2)", p.mode); return L3D_ERR_UNSUPPORTED; }
    g.bs = p.blockSize; g.SW2 = p.blockSize / 2;
    g.uniq = p.uniquenessRatio >= 0 ? p.uniquenessRatio : 10;
    g.d12 = p.disp12MaxDiff > 0 ? p.disp12MaxDiff : 1;
    g.P1 = p.P1 > 0 ? p.P1 : 2;
    g.P2 = std::max(p.P2 > 0 ? p.P2 : 5, g.P1 + 1);
    if (g.P2 > 16000) { set_err(err, "sgbm: P2=%d exceeds the int16 value domain", g.P2); return L3D_ERR_UNSUPPORTED; }
    g.ftzero = std::max(p.preFilterCap, 15) | 1;
    if (g.ftzero > 127) { set_err(err, "sgbm: preFilterCap too large"); return L3D_ERR_UNSUPPORTED; }
    g.minX1 = std::max(g.maxD, 0); g.maxX1 = W + std::min(g.minD, 0); g.width1 = g.maxX1 - g.minX1;
    g.DPL = g.D <= 64 ? 2 : (g.D <= 128 ? 4 : 8);
    g.NP = g.DPL / 2;
    g.nact = g.D / g.DPL;
    if (g.mode == 2) {
        const int nstripes = 4;
        int stripe_sz = (H + nstripes - 1) / nstripes;
        double t = 0.1 * stripe_sz; int ci = (int)t; if ((double)ci < t) ci++;
        int overlap = (g.bs / 2 + 1) + ci;
        g.nseg = 0; g.HV = 0;
        for (int s = 0; s < nstripes; s++) {
            int y0 = std::max(std::min(s * stripe_sz - overlap, H), 0);
            int y1 = std::min((s + 1) * stripe_sz, H);
            if (y1 <= y0) continue;
            int i = g.nseg++;
            g.seg_vr0[i] = g.HV; g.seg_y0[i] = y0; g.seg_rows[i] = y1 - y0; g.seg_emit[i] = s * stripe_sz;
            g.HV += y1 - y0;
        }
    } else {
        g.nseg = 1; g.HV = H;
        g.seg_vr0[0] = 0; g.seg_y0[0] = 0; g.seg_rows[0] = H; g.seg_emit[0] = 0;
    }
    return L3D_OK;
}

int sgbm_volume_rows(const l3d_sgbm_params& p, int W, int H) {
    Geom g; std::string e;
    if (make_geom(p, W, H, g, &e) != L3D_OK) return -1;
    return g.HV;
}

// ------------------------------------------------------------------------------------------
// prefilter: per pixel {c0.v, c0.lo, c0.hi, c1.v | c1.lo, c1.hi, 0, 0}
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void pre_pixel(const uint8_t* __restrict__ img, int W, int H, int y, int x,
                                          int ftzero, int& c0, int& c1) {
    if (x <= 0 || x >= W - 1) { c0 = ftzero; c1 = ftzero; return; }
    const uint8_t* r = img + (size_t)y * W;
    const uint8_t* rn = img + (size_t)(y > 0 ? y - 1 : y) * W;
    const uint8_t* rs = img + (size_t)(y < H - 1 ? y + 1 : y) * W;
    int g = ((int)r[x + 1] - (int)r[x - 1]) * 2 + (int)rn[x + 1] - (int)rn[x - 1] + (int)rs[x + 1] - (int)rs[x - 1];
    c0 = min(max(g, -ftzero), ftzero) + ftzero;
    c1 = r[x];
}

__global__ void sgbm_prefilter_kernel(const uint8_t* __restrict__ img, int W, int H, int ftzero,
                                      uint2* __restrict__ desc) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    int a0, a1, b0, b1, c0, c1;
    pre_pixel(img, W, H, y, x, ftzero, b0, b1);
    int lo0 = b0, hi0 = b0, lo1 = b1, hi1 = b1;
    if (x > 0) {
        pre_pixel(img, W, H, y, x - 1, ftzero, a0, a1);
        int t0 = (b0 + a0) >> 1, t1 = (b1 + a1) >> 1;
        lo0 = min(lo0, t0); hi0 = max(hi0, t0); lo1 = min(lo1, t1); hi1 = max(hi1, t1);
    }
    if (x < W - 1) {
        pre_pixel(img, W, H, y, x + 1, ftzero, c0, c1);
        int t0 = (b0 + c0) >> 1, t1 = (b1 + c1) >> 1;
        lo0 = min(lo0, t0); hi0 = max(hi0, t0); lo1 = min(lo1, t1); hi1 = max(hi1, t1);
    }
    uint2 d;
    d.x = (uint32_t)b0 | ((uint32_t)lo0 << 8) | ((uint32_t)hi0 << 16) | ((uint32_t)b1 << 24);
    d.y = (uint32_t)lo1 | ((uint32_t)hi1 << 8);
    desc[(size_t)y * W + x] = d;
}

// ------------------------------------------------------------------------------------------
// cost volume
// ------------------------------------------------------------------------------------------
struct CostArgs {
    const uint2* Ldesc; const uint2* Rdesc; int16_t* C;
    int W, minD, D, minX1, width1, SW2, bs, P2, TX;
    int nbands;
    int band_vr0[MAXBAND], band_y0[MAXBAND], band_rows[MAXBAND], band_clo[MAXBAND], band_chi[MAXBAND];
};

__device__ __forceinline__ uint32_t bt_cost(uint2 l, uint2 r) {
    int u = l.x & 255, u0 = (l.x >> 8) & 255, u1 = (l.x >> 16) & 255;
    int v = r.x & 255, v0 = (r.x >> 8) & 255, v1 = (r.x >> 16) & 255;
    int c0 = max(max(0, u - v1), v0 - u);
    int c1 = max(max(0, v - u1), u0 - v);
    int a = min(c0, c1);
    u = l.x >> 24; u0 = l.y & 255; u1 = (l.y >> 8) & 255;
    v = r.x >> 24; v0 = r.y & 255; v1 = (r.y >> 8) & 255;
    c0 = max(max(0, u - v1), v0 - u);
    c1 = max(max(0, v - u1), u0 - v);
    return (uint32_t)(a + (min(c0, c1) >> 2));
}

constexpr int COST_THREADS = 256;
constexpr int COST_NB = 8;  // phase-B items per thread (TX*D/2 <= 2048)

__global__ void __launch_bounds__(COST_THREADS) sgbm_cost_kernel(const CostArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int TX = a.TX, SW2 = a.SW2, bs = a.bs, D = a.D, D2 = D >> 1, TXH = TX + 2 * SW2;
    const int width1 = a.width1, W = a.W;
    uint2* sL = (uint2*)smem_raw;
    uint2* sR = sL + TXH;
    uint32_t* pd = (uint32_t*)(sR + TXH + D);
    uint32_t* ring = pd + TXH * D2;
    const int tid = threadIdx.x;
    const int b = blockIdx.y, x0 = blockIdx.x * TX;
    const int y0 = a.band_y0[b], rows = a.band_rows[b], clo = a.band_clo[b], chi = a.band_chi[b], vr0 = a.band_vr0[b];
    const int xa = min(max(x0 - SW2, 0), width1 - 1);
    const int xb = min(max(x0 + TX - 1 + SW2, 0), width1 - 1);
    const int xr_base = xa + a.minX1 - (a.minD + D - 1);
    const int nR = (xb - xa) + D;
    const int nitemsA = TXH * D2, nitemsB = TX * D2;
    const uint32_t p2x2 = (uint32_t)a.P2 * 0x10001u;
    uint32_t crun[COST_NB];
#pragma unroll
    for (int j = 0; j < COST_NB; j++) crun[j] = p2x2;

    for (int k = 0; k < rows + bs - 1; k++) {
        const int ky = min(max(y0 - SW2 + k, clo), chi);
        const uint2* Lrow = a.Ldesc + (size_t)ky * W;
        const uint2* Rrow = a.Rdesc + (size_t)ky * W;
        for (int i = tid; i < TXH; i += COST_THREADS) {
            int xc = min(max(x0 - SW2 + i, 0), width1 - 1);
            sL[i] = Lrow[xc + a.minX1];
        }
        for (int i = tid; i < nR; i += COST_THREADS) sR[i] = Rrow[xr_base + i];
        __syncthreads();
        for (int item = tid; item < nitemsA; item += COST_THREADS) {
            int c = item / D2, dp = item - c * D2;
            int xc = min(max(x0 - SW2 + c, 0), width1 - 1);
            uint2 l = sL[c];
            int ri = xc - xa + (D - 1) - 2 * dp;
            pd[item] = bt_cost(l, sR[ri]) | (bt_cost(l, sR[ri - 1]) << 16);
        }
        __syncthreads();
        const int slot = k % bs;
#pragma unroll
        for (int j = 0; j < COST_NB; j++) {
            int item = tid + COST_THREADS * j;
            if (item < nitemsB) {
                int xx = item / D2, dp = item - xx * D2;
                uint32_t h = 0;
                const uint32_t* pp = pd + xx * D2 + dp;
                for (int jj = 0; jj < bs; jj++) h = __vadd2(h, pp[jj * D2]);
                uint32_t* rp = ring + (slot * TX + xx) * D2 + dp;
                uint32_t c = __vadd2(crun[j], h);
                if (k >= bs) c = __vsub2(c, *rp);
                *rp = h;
                crun[j] = c;
                int x = x0 + xx;
                if (k >= bs - 1 && x < width1) {
                    size_t off = ((size_t)(vr0 + k - (bs - 1)) * width1 + x) * D + 2 * dp;
                    *(uint32_t*)(a.C + off) = c;
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// directional aggregation scan
// ------------------------------------------------------------------------------------------
struct ScanArgs {
    const int16_t* C; int16_t* S;
    int width1, D, nact, P1, P2;
    int kind;   // 0 ->, 1 <-, 2 down, 3 down-right, 4 down-left, 5 up, 6 up-left, 7 up-right
    int store;  // 1: S = L ; 0: S = min(S + L, 32767)
    int HV, nseg;
    int seg_vr0[MAXSEG], seg_rows[MAXSEG];
};

template <int NP> struct VecOf;
template <> struct VecOf<1> { typedef uint32_t T; };
template <> struct VecOf<2> { typedef uint2 T; };
template <> struct VecOf<4> { typedef uint4 T; };

template <int NP> __device__ __forceinline__ void vec_unpack(const typename VecOf<NP>::T& v, uint32_t (&o)[NP]);
template <> __device__ __forceinline__ void vec_unpack<1>(const uint32_t& v, uint32_t (&o)[1]) { o[0] = v; }
template <> __device__ __forceinline__ void vec_unpack<2>(const uint2& v, uint32_t (&o)[2]) { o[0] = v.x; o[1] = v.y; }
template <> __device__ __forceinline__ void vec_unpack<4>(const uint4& v, uint32_t (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
template <int NP> __device__ __forceinline__ typename VecOf<NP>::T vec_pack(const uint32_t (&o)[NP]);
template <> __device__ __forceinline__ uint32_t vec_pack<1>(const uint32_t (&o)[1]) { return o[0]; }
template <> __device__ __forceinline__ uint2 vec_pack<2>(const uint32_t (&o)[2]) { return make_uint2(o[0], o[1]); }
template <> __device__ __forceinline__ uint4 vec_pack<4>(const uint32_t (&o)[4]) { return make_uint4(o[0], o[1], o[2], o[3]); }

constexpr int SCAN_WARPS = 4;
constexpr int SCAN_DEPTH = 16;

// one SGM step for the disparities held by this lane; returns the warp-wide min of the new L
template <int NP>
__device__ __forceinline__ int sgm_step(uint32_t (&L)[NP], int minL, const uint32_t (&Cv)[NP], uint32_t p1x2,
                                        int P2, int lane, int nact) {
    uint32_t up = __shfl_up_sync(FULL, L[NP - 1], 1);
    uint32_t dn = __shfl_down_sync(FULL, L[0], 1);
    if (lane == 0) up = INF2;
    if (lane >= nact - 1) dn = INF2;
    const uint32_t delta = (uint32_t)(minL + P2) & 0xffffu;
    const uint32_t delta2 = delta * 0x10001u;
    const uint32_t ndelta2 = ((0x10000u - delta) & 0xffffu) * 0x10001u;
    uint32_t Ln[NP];
    uint32_t mn = INF2;
#pragma unroll
    for (int k = 0; k < NP; k++) {
        uint32_t prev = k ? L[k - 1] : up;
        uint32_t next = (k < NP - 1) ? L[k + 1] : dn;
        uint32_t dm1 = __byte_perm(prev, L[k], 0x5432);
        uint32_t dp1 = __byte_perm(L[k], next, 0x5432);
        uint32_t m = __vminu2(L[k], delta2);
        m = __viaddmin_u16x2(dm1, p1x2, m);
        m = __viaddmin_u16x2(dp1, p1x2, m);
        Ln[k] = __vadd2(__vadd2(Cv[k], m), ndelta2);
        mn = __vminu2(mn, Ln[k]);
    }
    int m16 = (int)min(mn & 0xffffu, mn >> 16);
    if (lane >= nact) m16 = 0x7fff;
#pragma unroll
    for (int k = 0; k < NP; k++) L[k] = Ln[k];
    return __reduce_min_sync(FULL, m16);
}

template <int NP>
__global__ void __launch_bounds__(SCAN_WARPS * 32) sgbm_scan_kernel(const ScanArgs a) {
    typedef typename VecOf<NP>::T vec;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    vec* ringC = (vec*)smem_raw + (size_t)warp * SCAN_DEPTH * 32;
    vec* ringS = (vec*)smem_raw + (size_t)(SCAN_WARPS + warp) * SCAN_DEPTH * 32;
    const int line = blockIdx.x * SCAN_WARPS + warp;
    const int width1 = a.width1, kind = a.kind;
    int n, vr, x, dvr, dx;
    if (kind <= 1) {
        if (line >= a.HV) return;
        n = width1; vr = line; dvr = 0;
        if (kind == 0) { x = 0; dx = 1; } else { x = width1 - 1; dx = -1; }
    } else {
        int seg = line / width1, xs = line - seg * width1;
        if (seg >= a.nseg) return;
        n = a.seg_rows[seg];
        if (kind <= 4) { vr = a.seg_vr0[seg]; dvr = 1; } else { vr = a.seg_vr0[seg] + n - 1; dvr = -1; }
        x = xs;
        dx = (kind == 2 || kind == 5) ? 0 : ((kind == 3 || kind == 7) ? 1 : -1);
    }
    const bool active = lane < a.nact;
    const bool store = a.store != 0;
    const size_t lane_off = (size_t)lane * (NP * 2);  // int16 elements
    const int D = a.D;
    const uint32_t p1x2 = (uint32_t)a.P1 * 0x10001u;
    const int P2 = a.P2;

    // look-ahead iterator for the cp.async ring
    int pvr = vr, px = x;
    auto issue = [&](int stage) {
        if (active) {
            size_t off = ((size_t)pvr * width1 + px) * D + lane_off;
            __pipeline_memcpy_async(&ringC[stage * 32 + lane], a.C + off, sizeof(vec));
            if (!store) __pipeline_memcpy_async(&ringS[stage * 32 + lane], a.S + off, sizeof(vec));
        }
        pvr += dvr; px += dx;
        if (px >= width1) px = 0;
        if (px < 0) px = width1 - 1;
    };
    for (int s = 0; s < SCAN_DEPTH; s++) {
        if (s < n) issue(s);
        __pipeline_commit();
    }
    uint32_t L[NP];
#pragma unroll
    for (int k = 0; k < NP; k++) L[k] = 0;
    int minL = 0;
    for (int i = 0; i < n; i++) {
        const int stage = i % SCAN_DEPTH;
        __pipeline_wait_prior(SCAN_DEPTH - 1);
        uint32_t Cv[NP], Sv[NP];
        if (active) {
            vec_unpack<NP>(ringC[stage * 32 + lane], Cv);
            if (!store) vec_unpack<NP>(ringS[stage * 32 + lane], Sv);
        } else {
#pragma unroll
            for (int k = 0; k < NP; k++) { Cv[k] = 0; Sv[k] = 0; }
        }
        // diagonal lines wrap around the image; the predecessor of the re-entry pixel is outside
        if (dx != 0 && dvr != 0 && i > 0 && ((dx > 0 && x == 0) || (dx < 0 && x == width1 - 1))) {
#pragma unroll
            for (int k = 0; k < NP; k++) L[k] = 0;
            minL = 0;
        }
        minL = sgm_step<NP>(L, minL, Cv, p1x2, P2, lane, a.nact);
        if (active) {
            uint32_t out[NP];
#pragma unroll
            for (int k = 0; k < NP; k++) out[k] = store ? L[k] : __vminu2(__vadd2(Sv[k], L[k]), INF2);
            size_t off = ((size_t)vr * width1 + x) * D + lane_off;
            *(vec*)(a.S + off) = vec_pack<NP>(out);
        }
        if (i + SCAN_DEPTH < n) issue(stage);
        __pipeline_commit();
        vr += dvr; x += dx;
        if (x >= width1) x = 0;
        if (x < 0) x = width1 - 1;
    }
}

// ------------------------------------------------------------------------------------------
// WTA + uniqueness + disp2 + sub-pixel
// ------------------------------------------------------------------------------------------
struct WtaArgs {
    const int16_t* S; int16_t* raw; unsigned* disp2key;
    int W, width1, D, nact, DPL, minD, minX1, uniq, mode;
    int HV, nseg;
    int seg_vr0[MAXSEG], seg_y0[MAXSEG], seg_rows[MAXSEG], seg_emit[MAXSEG];
};
constexpr int WTA_WARPS = 8;

template <int NP>
__global__ void __launch_bounds__(WTA_WARPS * 32) sgbm_wta_kernel(const WtaArgs a) {
    typedef typename VecOf<NP>::T vec;
    constexpr int DPL = NP * 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long pix = (long)blockIdx.x * WTA_WARPS + warp;
    const int width1 = a.width1;
    if (pix >= (long)a.HV * width1) return;
    const int vr = (int)(pix / width1), x = (int)(pix - (long)vr * width1);
    int seg = 0;
    for (int s = 1; s < a.nseg; s++) if (vr >= a.seg_vr0[s]) seg = s;
    const int y = a.seg_y0[seg] + (vr - a.seg_vr0[seg]);
    if (y < a.seg_emit[seg]) return;
    const int16_t* Sp = a.S + ((size_t)vr * width1 + x) * a.D;
    uint32_t w[NP];
    if (lane < a.nact) vec_unpack<NP>(*(const vec*)(Sp + lane * DPL), w);
    else {
#pragma unroll
        for (int k = 0; k < NP; k++) w[k] = INF2;
    }
    int s[DPL];
#pragma unroll
    for (int k = 0; k < NP; k++) { s[2 * k] = (int)(w[k] & 0xffffu); s[2 * k + 1] = (int)(w[k] >> 16); }
    const int d0 = lane * DPL;
    int minS, best;
    if (a.mode != 2) {
        unsigned key = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < DPL; j++) key = min(key, ((unsigned)s[j] << 8) | (unsigned)((d0 + j) & 255));
        key = __reduce_min_sync(FULL, key);
        minS = (int)(key >> 8); best = (int)(key & 255);
        if (minS >= 32767) return;  // nothing beats MAX_COST: pixel stays invalid, disp2 untouched
        bool rej = false;
#pragma unroll
        for (int j = 0; j < DPL; j++) {
            int d = d0 + j;
            if (lane < a.nact && s[j] * (100 - a.uniq) < minS * 100 && abs(best - d) > 1) rej = true;
        }
        if (__any_sync(FULL, rej)) return;
    } else {
        int m = 32767;
#pragma unroll
        for (int j = 0; j < DPL; j++) m = min(m, s[j]);
        minS = __reduce_min_sync(FULL, m);
        best = 0x7fffffff;
        for (int c = 0; c < 8; c++) {
            int v = -1;
#pragma unroll
            for (int j = 0; j < DPL; j++) {
                int d = d0 + j;
                if (lane < a.nact && (d & 7) == c && s[j] == minS) v = max(v, d);
            }
            v = __reduce_max_sync(FULL, v);
            if (v >= 0) best = min(best, v);
        }
        if (a.uniq > 0) {
            int thresh = (100 * minS) / (100 - a.uniq);
            int tr = (int)(short)(thresh + 1);
            bool rej = false;
#pragma unroll
            for (int j = 0; j < DPL; j++) {
                int d = d0 + j;
                if (lane < a.nact && s[j] < tr && (d < best - 1 || d > best + 1)) rej = true;
            }
            if (__any_sync(FULL, rej)) return;
        }
    }
    if (lane == 0) {
        int d = best;
        int x2 = x + a.minX1 - d - a.minD;
        bool ok2 = (a.mode == 2) ? (x2 >= 0 && x2 < a.W) : (x2 >= 0 && x2 < a.W + 2);
        if (ok2 && minS < 32767)
            atomicMax(a.disp2key + (size_t)y * (a.W + 2) + x2, ((unsigned)(0x7fff - minS) << 16) | (unsigned)x);
        int dd;
        if (0 < d && d < a.D - 1) {
            int sm = Sp[d - 1], sp = Sp[d + 1], sc = Sp[d];
            int denom2 = max(sm + sp - 2 * sc, 1);
            dd = d * 16 + ((sm - sp) * 16 + denom2) / (denom2 * 2);
        } else dd = d * 16;
        a.raw[(size_t)y * a.W + x + a.minX1] = (int16_t)(dd + a.minD * 16);
    }
}

__global__ void sgbm_lrcheck_kernel(int16_t* __restrict__ raw, const unsigned* __restrict__ disp2key, int W, int H,
                                    int minX1, int maxX1, int minD, int d12) {
    int x = minX1 + blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= maxX1) return;
    const int INVALID = (minD - 1) * 16;
    int d1 = raw[(size_t)y * W + x];
    if (d1 == INVALID) return;
    int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
    int _x = x - _d, x_ = x - d_;
    const unsigned* k2 = disp2key + (size_t)y * (W + 2);
    bool c = true;
    if (0 <= _x && _x < W) {
        unsigned k = k2[_x];
        int d2 = (k >> 16) ? (int)(k & 0xffffu) + minX1 - _x : INVALID;  // unset entries hold the SCALED invalid value, as in OpenCV
        c = c && d2 >= minD && abs(d2 - _d) > d12;
    } else c = false;
    if (0 <= x_ && x_ < W) {
        unsigned k = k2[x_];
        int d2 = (k >> 16) ? (int)(k & 0xffffu) + minX1 - x_ : INVALID;
        c = c && d2 >= minD && abs(d2 - d_) > d12;
    } else c = false;
    if (c) raw[(size_t)y * W + x] = (int16_t)INVALID;
}

__global__ void fill_s16_kernel(int16_t* p, size_t n, int16_t v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------
template <int NP>
static int launch_scan(Lane& L, const ScanArgs& sa, int lines) {
    size_t smem = (size_t)2 * SCAN_WARPS * SCAN_DEPTH * 32 * sizeof(typename VecOf<NP>::T);
    static bool attr_done = false;
    if (!attr_done) {
        L3D_CHECK(L, cudaFuncSetAttribute(sgbm_scan_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
        attr_done = true;
    }
    L3D_LAUNCH(L, sgbm_scan_kernel<NP>, cdiv(lines, SCAN_WARPS), SCAN_WARPS * 32, smem, sa);
    return L3D_OK;
}
static int launch_scan_np(Lane& L, int NP, const ScanArgs& sa, int lines) {
    if (NP == 1) return launch_scan<1>(L, sa, lines);
    if (NP == 2) return launch_scan<2>(L, sa, lines);
    return launch_scan<4>(L, sa, lines);
}

int dev_sgbm(Lane& L, const l3d_sgbm_params& p, const uint8_t* left, const uint8_t* right, int W, int H,
             int16_t* disp, SgbmDebug* dbg) {
    Geom g;
    int rc = make_geom(p, W, H, g, L.err);
    if (rc != L3D_OK) return rc;
    const size_t npix = (size_t)W * H;
    const int16_t INVALID = (int16_t)((g.minD - 1) * 16);
    int16_t* raw = L.get<int16_t>(S_RAW, npix);
    L3D_LAUNCH(L, fill_s16_kernel, cdiv(npix, 256), 256, 0, raw, npix, INVALID);
    if (g.width1 > 0) {
        // --- descriptors
        uint2* dL = L.get<uint2>(S_DESC_L, npix);
        uint2* dR = L.get<uint2>(S_DESC_R, npix);
        dim3 pg(cdiv(W, 128), H);
        L3D_LAUNCH(L, sgbm_prefilter_kernel, pg, 128, 0, left, W, H, g.ftzero, dL);
        L3D_LAUNCH(L, sgbm_prefilter_kernel, pg, 128, 0, right, W, H, g.ftzero, dR);
        // --- cost volume
        const size_t nvol = (size_t)g.HV * g.width1 * g.D;
        int16_t* C = L.get<int16_t>(S_COST, nvol);
        int16_t* S = L.get<int16_t>(S_AGGR, nvol);
        CostArgs ca;
        ca.Ldesc = dL; ca.Rdesc = dR; ca.C = C;
        ca.W = W; ca.minD = g.minD; ca.D = g.D; ca.minX1 = g.minX1; ca.width1 = g.width1;
        ca.SW2 = g.SW2; ca.bs = g.bs; ca.P2 = g.P2;
        int TX = std::min(32, 4096 / g.D);
        auto cost_smem = [&](int tx) {
            int TXH = tx + 2 * g.SW2;
            return (size_t)TXH * 8 + (size_t)(TXH + g.D) * 8 + (size_t)TXH * (g.D / 2) * 4 + (size_t)g.bs * tx * (g.D / 2) * 4;
        };
        while (TX > 1 && cost_smem(TX) > 200 * 1024) TX /= 2;
        ca.TX = TX;
        // bands: split each segment so the grid fills the SMs about twice
        int xtiles = cdiv(g.width1, TX);
        int want = std::max(1, (2 * NUM_SMS * 2 + xtiles - 1) / xtiles);
        int per_seg = std::max(1, std::min(want / g.nseg, MAXBAND / g.nseg));
        ca.nbands = 0;
        for (int s = 0; s < g.nseg; s++) {
            int rows = g.seg_rows[s];
            int nb = std::max(1, std::min(per_seg, rows / (2 * g.bs) > 0 ? rows / (2 * g.bs) : 1));
            int br = cdiv(rows, nb);
            for (int r0 = 0; r0 < rows; r0 += br) {
                int i = ca.nbands++;
                ca.band_vr0[i] = g.seg_vr0[s] + r0;
                ca.band_y0[i] = g.seg_y0[s] + r0;
                ca.band_rows[i] = std::min(br, rows - r0);
                ca.band_clo[i] = g.seg_y0[s];  // the vertical box sum restarts at the segment top
                ca.band_chi[i] = H - 1;
            }
        }
        size_t smem = cost_smem(TX);
        L3D_CHECK(L, cudaFuncSetAttribute(sgbm_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        L.t_begin("sgbm_cost");
        L3D_LAUNCH(L, sgbm_cost_kernel, dim3(xtiles, ca.nbands), COST_THREADS, smem, ca);
        L.t_end("sgbm_cost");
        // --- aggregation
        ScanArgs sa;
        sa.C = C; sa.S = S; sa.width1 = g.width1; sa.D = g.D; sa.nact = g.nact; sa.P1 = g.P1; sa.P2 = g.P2;
        sa.HV = g.HV; sa.nseg = g.nseg;
        for (int s = 0; s < MAXSEG; s++) { sa.seg_vr0[s] = g.seg_vr0[s]; sa.seg_rows[s] = g.seg_rows[s]; }
        int kinds[8], nk = 0;
        kinds[nk++] = 0; kinds[nk++] = 1; kinds[nk++] = 2;  // ->, <-, down: all modes
        if (g.mode != 2) { kinds[nk++] = 3; kinds[nk++] = 4; }
        if (g.mode == 1) { kinds[nk++] = 5; kinds[nk++] = 6; kinds[nk++] = 7; }
        for (int i = 0; i < nk; i++) {
            sa.kind = kinds[i]; sa.store = (i == 0);
            int lines = kinds[i] <= 1 ? g.HV : g.nseg * g.width1;
            L.t_begin("sgbm_scan");
            rc = launch_scan_np(L, g.NP, sa, lines);
            L.t_end("sgbm_scan");
            if (rc != L3D_OK) return rc;
        }
        // --- WTA, LR check
        unsigned* d2 = L.get<unsigned>(S_DISP2, (size_t)H * (W + 2));
        L3D_CHECK(L, cudaMemsetAsync(d2, 0, (size_t)H * (W + 2) * sizeof(unsigned), L.stream));
        WtaArgs wa;
        wa.S = S; wa.raw = raw; wa.disp2key = d2; wa.W = W; wa.width1 = g.width1; wa.D = g.D; wa.nact = g.nact;
        wa.DPL = g.DPL; wa.minD = g.minD; wa.minX1 = g.minX1; wa.uniq = g.uniq; wa.mode = g.mode;
        wa.HV = g.HV; wa.nseg = g.nseg;
        for (int s = 0; s < MAXSEG; s++) {
            wa.seg_vr0[s] = g.seg_vr0[s]; wa.seg_y0[s] = g.seg_y0[s]; wa.seg_rows[s] = g.seg_rows[s]; wa.seg_emit[s] = g.seg_emit[s];
        }
        int wgrid = cdiv((long)g.HV * g.width1, WTA_WARPS);
        L.t_begin("sgbm_wta");
        if (g.NP == 1) L3D_LAUNCH(L, sgbm_wta_kernel<1>, wgrid, WTA_WARPS * 32, 0, wa);
        else if (g.NP == 2) L3D_LAUNCH(L, sgbm_wta_kernel<2>, wgrid, WTA_WARPS * 32, 0, wa);
        else L3D_LAUNCH(L, sgbm_wta_kernel<4>, wgrid, WTA_WARPS * 32, 0, wa);
        L.t_end("sgbm_wta");
        L3D_LAUNCH(L, sgbm_lrcheck_kernel, dim3(cdiv(g.width1, 128), H), 128, 0, raw, d2, W, H, g.minX1, g.maxX1, g.minD, g.d12);
        if (dbg) {
            if (dbg->C) L3D_CHECK(L, cudaMemcpyAsync(dbg->C, C, nvol * 2, cudaMemcpyDeviceToDevice, L.stream));
            if (dbg->S) L3D_CHECK(L, cudaMemcpyAsync(dbg->S, S, nvol * 2, cudaMemcpyDeviceToDevice, L.stream));
        }
    }
    if (dbg && dbg->raw) L3D_CHECK(L, cudaMemcpyAsync(dbg->raw, raw, npix * 2, cudaMemcpyDeviceToDevice, L.stream));
    // --- medianBlur(3) then filterSpeckles, as StereoSGBM::compute does
    rc = dev_median3(L, raw, W, H, disp);
    if (rc != L3D_OK) return rc;
    if (p.speckleWindowSize > 0) rc = dev_speckles(L, disp, W, H, INVALID, p.speckleWindowSize, 16 * p.speckleRange);
    return rc;
}

}  // namespace l3d
