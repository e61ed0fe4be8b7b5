// post.cu -- disparity post-filters applied inside cv2.StereoSGBM.compute
// (camera/single_usb_stereo_camera.py:324-325): medianBlur(disp, 3) then filterSpeckles.
// Also hosts the union-find connected-component labelling shared with the Simple extractor.
#include "common.cuh"
#include "ccl.cuh"

namespace l3d {

__device__ __forceinline__ void cswap(int& a, int& b) { int t = min(a, b); b = max(a, b); a = t; }

// 3x3 median, BORDER_REPLICATE, int16.  HBM-bound: 2 B in + 2 B out per pixel.
__global__ void median3_s16_kernel(const int16_t* __restrict__ src, int W, int H, int16_t* __restrict__ dst) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    int xm = max(x - 1, 0), xp = min(x + 1, W - 1);
    const int16_t* r0 = src + (size_t)max(y - 1, 0) * W;
    const int16_t* r1 = src + (size_t)y * W;
    const int16_t* r2 = src + (size_t)min(y + 1, H - 1) * W;
    int p0 = r0[xm], p1 = r0[x], p2 = r0[xp], p3 = r1[xm], p4 = r1[x], p5 = r1[xp], p6 = r2[xm], p7 = r2[x], p8 = r2[xp];
    // 19-exchange median-of-9 network
    cswap(p1, p2); cswap(p4, p5); cswap(p7, p8); cswap(p0, p1); cswap(p3, p4); cswap(p6, p7);
    cswap(p1, p2); cswap(p4, p5); cswap(p7, p8); cswap(p0, p3); cswap(p5, p8); cswap(p4, p7);
    cswap(p3, p6); cswap(p1, p4); cswap(p2, p5); cswap(p4, p7); cswap(p4, p2); cswap(p6, p4);
    cswap(p4, p2);
    dst[(size_t)y * W + x] = (int16_t)p4;
}

int dev_median3(Lane& L, const int16_t* src, int W, int H, int16_t* dst) {
    L3D_LAUNCH(L, median3_s16_kernel, dim3(cdiv(W, 128), H), 128, 0, src, W, H, dst);
    return L3D_OK;
}

// ---- filterSpeckles: 4-connected components over pixels != newVal, edge iff |a-b| <= maxDiff;
//      components with size <= maxSize are set to newVal.
struct SpeckleRule {
    const int16_t* img; int newVal, maxDiff;
    __device__ bool node(int i) const { return img[i] != newVal; }
    __device__ bool edge(int i, int j) const { return abs((int)img[i] - (int)img[j]) <= maxDiff; }
};

__global__ void speckle_apply_kernel(int16_t* img, const int* label, const int* count, int n, int newVal, int maxSize) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int l = label[i];
    if (l >= 0 && count[l] <= maxSize) img[i] = (int16_t)newVal;
}

int dev_speckles(Lane& L, int16_t* img, int W, int H, int newVal, int maxSize, int maxDiff) {
    int n = W * H;
    int* label = L.get<int>(S_LABEL, n);
    int* count = L.get<int>(S_CNT, n);
    SpeckleRule rule{img, newVal, maxDiff};
    int rc = ccl_label<SpeckleRule, false>(L, rule, W, H, label);
    if (rc != L3D_OK) return rc;
    L3D_CHECK(L, cudaMemsetAsync(count, 0, sizeof(int) * n, L.stream));
    L3D_LAUNCH(L, ccl_count_kernel, cdiv(n, 256), 256, 0, label, count, n);
    L3D_LAUNCH(L, speckle_apply_kernel, cdiv(n, 256), 256, 0, img, label, count, n, newVal, maxSize);
    return L3D_OK;
}

}  // namespace l3d
