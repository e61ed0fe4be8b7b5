// ccl.cuh -- union-find connected-component labelling on a W x H grid (label-equivalence with
// atomicMin linking).  Rule supplies node(i) (pixel takes part) and edge(i,j) (adjacent pixels are
// connected).  Result: label[i] = smallest pixel index of i's component, -1 for non-nodes.
// Used by filterSpeckles (4-connectivity) and the Simple extractor's contour model (4 and 8).
#pragma once
#include "common.cuh"

namespace l3d {

__device__ __forceinline__ int ccl_find(const int* label, int i) {
    const volatile int* L = label;
    int n = L[i];
    while (n != i) { i = n; n = L[i]; }
    return i;
}

__device__ __forceinline__ void ccl_unite(int* label, int a, int b) {
    bool done;
    do {
        a = ccl_find(label, a);
        b = ccl_find(label, b);
        if (a < b) { int old = atomicMin(&label[b], a); done = (old == b); b = old; }
        else if (b < a) { int old = atomicMin(&label[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

template <class Rule>
__global__ void ccl_init_kernel(Rule r, int n, int* label) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) label[i] = r.node(i) ? i : -1;
}

template <class Rule, bool EIGHT>
__global__ void ccl_merge_kernel(Rule r, int W, int H, int* label) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    int i = y * W + x;
    if (label[i] < 0) return;
    if (x > 0 && label[i - 1] >= 0 && r.edge(i, i - 1)) ccl_unite(label, i, i - 1);
    if (y > 0) {
        if (label[i - W] >= 0 && r.edge(i, i - W)) ccl_unite(label, i, i - W);
        if (EIGHT) {
            if (x > 0 && label[i - W - 1] >= 0 && r.edge(i, i - W - 1)) ccl_unite(label, i, i - W - 1);
            if (x < W - 1 && label[i - W + 1] >= 0 && r.edge(i, i - W + 1)) ccl_unite(label, i, i - W + 1);
        }
    }
}

static __global__ void ccl_flatten_kernel(int* label, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && label[i] >= 0) label[i] = ccl_find(label, i);
}

static __global__ void ccl_count_kernel(const int* label, int* count, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && label[i] >= 0) atomicAdd(&count[label[i]], 1);
}

template <class Rule, bool EIGHT>
int ccl_label(Lane& L, Rule r, int W, int H, int* label) {
    int n = W * H;
    L3D_LAUNCH(L, (ccl_init_kernel<Rule>), cdiv(n, 256), 256, 0, r, n, label);
    L3D_LAUNCH(L, (ccl_merge_kernel<Rule, EIGHT>), dim3(cdiv(W, 128), H), 128, 0, r, W, H, label);
    L3D_LAUNCH(L, ccl_flatten_kernel, cdiv(n, 256), 256, 0, label, n);
    return L3D_OK;
}

}  // namespace l3d
