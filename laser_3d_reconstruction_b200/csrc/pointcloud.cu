// pointcloud.cu -- the point-cloud sink right after the hot path: PointCloudProcessor.voxel_downsample
// and .statistical_outlier_removal of utils/point_cloud.py, in the form that file runs without Open3D
// (_simple_voxel_downsample :54-78, _simple_outlier_removal :108-131), with its exact f64 arithmetic:
//   voxel_downsample   key = floor(p / voxel) per axis; one output point per occupied voxel = np.mean of its points
//                      (sequential sum in input order, / count), voxels in order of first appearance (dict order)
//   outlier removal    per point the nb_neighbors + 1 smallest Euclidean distances (cKDTree.query incl. the point
//                      itself), drop the first, mean / std as numpy computes them (8-lane pairwise sum), keep when
//                      mean < mean + std_ratio * std  -- the reference's predicate as written
#include "common.cuh"

namespace l3d {

// ---- block-wise exclusive prefix sum of ints (n up to ~2^31): scan_blocks -> scan of the block totals -> add
constexpr int PS_THREADS = 256;
constexpr int PS_ITEMS = 8;
constexpr int PS_TILE = PS_THREADS * PS_ITEMS;

__global__ void __launch_bounds__(PS_THREADS) ps_tile_kernel(const int* __restrict__ in, int* __restrict__ out, int n, int* __restrict__ tile_sum) {
    __shared__ int warp_tot[PS_THREADS / 32];
    const int base = blockIdx.x * PS_TILE + threadIdx.x * PS_ITEMS;
    int v[PS_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < PS_ITEMS; i++) { v[i] = base + i < n ? in[base + i] : 0; s += v[i]; }
    int incl = s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; w++) woff += warp_tot[w];
    int run = woff + incl - s;
#pragma unroll
    for (int i = 0; i < PS_ITEMS; i++) { if (base + i < n) out[base + i] = run; run += v[i]; }
    if (threadIdx.x == PS_THREADS - 1) tile_sum[blockIdx.x] = woff + incl;
}
__global__ void ps_add_kernel(int* __restrict__ out, int n, const int* __restrict__ tile_off) {
    const int i = blockIdx.x * PS_TILE + threadIdx.x;
    const int off = tile_off[blockIdx.x];
    for (int k = threadIdx.x; k < PS_TILE && blockIdx.x * PS_TILE + k < n; k += blockDim.x) out[blockIdx.x * PS_TILE + k] += off;
    (void)i;
}
// exclusive scan of in[0..n) into out; *total (device) = sum.  tmp: scratch of >= 2 * (ntiles + PS_TILE) ints
static int dev_exclusive_scan(Lane& L, const int* in, int* out, int n, int* tmp, int* total_dev) {
    const int ntiles = cdiv(n, PS_TILE);
    int* tsum = tmp;
    int* toff = tmp + ntiles + 1;
    L3D_LAUNCH(L, ps_tile_kernel, ntiles, PS_THREADS, 0, in, out, n, tsum);
    if (ntiles > 1) {
        const int rc = dev_exclusive_scan(L, tsum, toff, ntiles, toff + ntiles + 1, total_dev);
        if (rc != L3D_OK) return rc;
        L3D_LAUNCH(L, ps_add_kernel, ntiles, PS_THREADS, 0, out, n, toff);
    } else {
        L3D_CHECK(L, cudaMemcpyAsync(total_dev, tsum, sizeof(int), cudaMemcpyDeviceToDevice, L.stream));
    }
    return L3D_OK;
}

// ---- voxel down-sampling
constexpr unsigned long long VX_EMPTY = ~0ull;
constexpr long long VX_LIM = 1ll << 20;  // |index| < 2^20 per axis: three 21-bit fields in one 64-bit key

__global__ void vx_insert_kernel(const double* __restrict__ pts, int n, double voxel, int f32, unsigned long long* __restrict__ keys,
                                 int* __restrict__ first, int* __restrict__ count, int cap_mask, int* __restrict__ slot_of,
                                 int* __restrict__ bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long q[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        // f32: the reference hands over a float32 cloud (main.py:208), numpy then divides and floors in float32
        const double f = f32 ? (double)floorf(__fdiv_rn((float)pts[(size_t)i * 3 + a], (float)voxel))
                             : floor(__ddiv_rn(pts[(size_t)i * 3 + a], voxel));
        if (!(f > -(double)VX_LIM && f < (double)VX_LIM)) { atomicExch(bad, 1); slot_of[i] = -1; return; }  // also NaN
        q[a] = (long long)f + VX_LIM;
    }
    const unsigned long long key = ((unsigned long long)q[0] << 42) | ((unsigned long long)q[1] << 21) | (unsigned long long)q[2];
    unsigned long long h = key * 0x9E3779B97F4A7C15ull;
    int s = (int)(h >> 40) & cap_mask;
    for (;;) {
        const unsigned long long prev = atomicCAS(keys + s, VX_EMPTY, key);
        if (prev == VX_EMPTY || prev == key) break;
        s = (s + 1) & cap_mask;
    }
    atomicMin(first + s, i);
    atomicAdd(count + s, 1);
    slot_of[i] = s;
}
__global__ void vx_flag_kernel(const int* __restrict__ slot_of, const int* __restrict__ first, int n, int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (slot_of[i] >= 0 && first[slot_of[i]] == i) ? 1 : 0;
}
// rank[i] = position of voxel in first-appearance order for the voxel's first point i
__global__ void vx_rank_kernel(const int* __restrict__ slot_of, const int* __restrict__ first, const int* __restrict__ count,
                               const int* __restrict__ rank_of_point, int n, int* __restrict__ rank_of_slot,
                               int* __restrict__ cnt_ranked) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = slot_of[i];
    if (s >= 0 && first[s] == i) { rank_of_slot[s] = rank_of_point[i]; cnt_ranked[rank_of_point[i]] = count[s]; }
}
__global__ void vx_scatter_kernel(const int* __restrict__ slot_of, const int* __restrict__ rank_of_slot, const int* __restrict__ offset,
                                  int* __restrict__ fill, int n, int* __restrict__ bucket) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || slot_of[i] < 0) return;
    const int r = rank_of_slot[slot_of[i]];
    bucket[offset[r] + atomicAdd(fill + r, 1)] = i;
}
// The mean of a voxel is numpy's sequential sum over its members IN INPUT ORDER, so the member indices (scattered in
// arbitrary order above) have to be sorted first.  Voxels with up to VX_SMALL members: one thread per voxel, insertion
// sort.  Larger ones (a static laser line accumulated over many frames puts 1e3 .. 1e5 points into one 2 mm voxel; the
// quadratic sort would run for seconds in one thread): one CTA per voxel, bitonic sort in place, then the sequential sum.
constexpr int VX_SMALL = 48;
constexpr int VX_BIG_THREADS = 256;

__device__ void vx_sum_store(const double* __restrict__ pts, const int* __restrict__ b, int m, int f32, double* __restrict__ o);

__global__ void __launch_bounds__(VX_BIG_THREADS) vx_mean_big_kernel(const double* __restrict__ pts, int* __restrict__ bucket,
                                                                      const int* __restrict__ offset, const int* __restrict__ cnt,
                                                                      int nvox, int f32, double* __restrict__ out) {
    const int r = blockIdx.x;
    if (r >= nvox) return;
    const int m = cnt[r];
    if (m <= VX_SMALL) return;  // handled by vx_mean_kernel
    int* b = bucket + offset[r];
    int p2 = 1;
    while (p2 < m) p2 <<= 1;
    // bitonic network over p2 virtual slots in its all-ascending form (every merge starts with the "flip" partner
    // i ^ (k - 1)): slots >= m hold +infinity, and as every comparator puts the smaller value into the lower slot, a
    // comparator with a virtual partner never moves anything
    auto cmpswap = [&](int i, int l) {
        if (l > i && l < m) {
            const int a = b[i], c = b[l];
            if (a > c) { b[i] = c; b[l] = a; }
        }
    };
    for (int k = 2; k <= p2; k <<= 1) {
        for (int i = threadIdx.x; i < m; i += VX_BIG_THREADS) cmpswap(i, i ^ (k - 1));
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < m; i += VX_BIG_THREADS) cmpswap(i, i ^ j);
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) vx_sum_store(pts, b, m, f32, out + (size_t)r * 3);
}

// one thread per voxel: put its member indices into input order, sum sequentially in f64, divide by the count
__global__ void vx_mean_kernel(const double* __restrict__ pts, int* __restrict__ bucket, const int* __restrict__ offset,
                               const int* __restrict__ cnt, int nvox, int f32, double* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nvox) return;
    int* b = bucket + offset[r];
    const int m = cnt[r];
    if (m > VX_SMALL) return;  // vx_mean_big_kernel
    for (int i = 1; i < m; i++) {  // insertion sort (voxels hold a handful of points)
        const int v = b[i];
        int j = i - 1;
        while (j >= 0 && b[j] > v) { b[j + 1] = b[j]; j--; }
        b[j + 1] = v;
    }
    vx_sum_store(pts, b, m, f32, out + (size_t)r * 3);
}

// numpy's mean of the member rows in input order (b sorted): sequential sum, then the division
__device__ void vx_sum_store(const double* __restrict__ pts, const int* __restrict__ b, int m, int f32, double* __restrict__ out) {
    const int r = 0;
    if (f32) {  // np.mean of float32 rows: float32 running sum, float32 division
        float fx = 0.f, fy = 0.f, fz = 0.f;
        for (int i = 0; i < m; i++) {
            const double* p = pts + (size_t)b[i] * 3;
            fx = __fadd_rn(fx, (float)p[0]); fy = __fadd_rn(fy, (float)p[1]); fz = __fadd_rn(fz, (float)p[2]);
        }
        const float fm = (float)m;
        out[(size_t)r * 3 + 0] = (double)__fdiv_rn(fx, fm); out[(size_t)r * 3 + 1] = (double)__fdiv_rn(fy, fm);
        out[(size_t)r * 3 + 2] = (double)__fdiv_rn(fz, fm);
        return;
    }
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int i = 0; i < m; i++) {
        const double* p = pts + (size_t)b[i] * 3;
        sx = __dadd_rn(sx, p[0]); sy = __dadd_rn(sy, p[1]); sz = __dadd_rn(sz, p[2]);
    }
    const double dm = (double)m;
    out[(size_t)r * 3 + 0] = __ddiv_rn(sx, dm); out[(size_t)r * 3 + 1] = __ddiv_rn(sy, dm); out[(size_t)r * 3 + 2] = __ddiv_rn(sz, dm);
}
__global__ void fill_i32_kernel(int* p, size_t n, int v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// pts, out: device, n x 3 f64 (f32 != 0: the values are float32 numbers and the arithmetic is float32 like numpy's);
// *nvox_host = number of output points
int dev_voxel_downsample(Lane& L, const double* pts, int n, double voxel, int f32, double* out, int* nvox_host) {
    *nvox_host = 0;
    if (n <= 0) return L3D_OK;
    L3D_ARG(L, voxel > 0.0, "voxel_size must be positive");
    int cap = 1024;
    while (cap < 2 * n) cap <<= 1;
    const int ntiles = cdiv(n, PS_TILE);
    unsigned long long* keys = L.get<unsigned long long>(S_IO_A, cap);
    int* ib = L.get<int>(S_IO_B, (size_t)2 * cap + (size_t)8 * n + 4 * (ntiles + PS_TILE) + 64);
    int *first = ib, *count = first + cap, *slot_of = count + cap, *flag = slot_of + n, *rankp = flag + n;
    int *cnt_ranked = rankp + n, *offset = cnt_ranked + n, *fill = offset + n, *bucket = fill + n, *scan_tmp = bucket + n;
    int* misc = scan_tmp + 4 * (ntiles + PS_TILE);  // [0] bad flag, [1] number of voxels, [2] scratch total
    int* rank_of_slot = L.get<int>(S_IO_C, cap);
    L3D_CHECK(L, cudaMemsetAsync(keys, 0xff, sizeof(unsigned long long) * cap, L.stream));
    L3D_LAUNCH(L, fill_i32_kernel, cdiv(cap, 256), 256, 0, first, (size_t)cap, 0x7fffffff);
    L3D_CHECK(L, cudaMemsetAsync(count, 0, sizeof(int) * cap, L.stream));
    L3D_CHECK(L, cudaMemsetAsync(fill, 0, sizeof(int) * n, L.stream));
    L3D_CHECK(L, cudaMemsetAsync(misc, 0, sizeof(int) * 4, L.stream));
    const int g = cdiv(n, 256);
    L3D_LAUNCH(L, vx_insert_kernel, g, 256, 0, pts, n, voxel, f32, keys, first, count, cap - 1, slot_of, misc);
    L3D_LAUNCH(L, vx_flag_kernel, g, 256, 0, slot_of, first, n, flag);
    int rc = dev_exclusive_scan(L, flag, rankp, n, scan_tmp, misc + 1);
    if (rc != L3D_OK) return rc;
    int h[2] = {0, 0};
    L3D_CHECK(L, cudaMemcpyAsync(h, misc, sizeof(int) * 2, cudaMemcpyDeviceToHost, L.stream));
    L3D_CHECK(L, cudaStreamSynchronize(L.stream));
    if (h[0]) { set_err(L.err, "voxel_downsample: a voxel index is not finite or outside +-2^20"); return L3D_ERR_UNSUPPORTED; }
    const int nvox = h[1];
    L3D_LAUNCH(L, vx_rank_kernel, g, 256, 0, slot_of, first, count, rankp, n, rank_of_slot, cnt_ranked);
    rc = dev_exclusive_scan(L, cnt_ranked, offset, nvox, scan_tmp, misc + 2);
    if (rc != L3D_OK) return rc;
    L3D_LAUNCH(L, vx_scatter_kernel, g, 256, 0, slot_of, rank_of_slot, offset, fill, n, bucket);
    L3D_LAUNCH(L, vx_mean_kernel, cdiv(nvox, 128), 128, 0, pts, bucket, offset, cnt_ranked, nvox, f32, out);
    L3D_LAUNCH(L, vx_mean_big_kernel, nvox, VX_BIG_THREADS, 0, pts, bucket, offset, cnt_ranked, nvox, f32, out);
    *nvox_host = nvox;
    return L3D_OK;
}

// ---- statistical outlier removal (the reference's predicate as written, utils/point_cloud.py:108-131)
constexpr int SOR_MAXK = 64;     // nb_neighbors + 1 <= SOR_MAXK
constexpr int SOR_THREADS = 128;
constexpr int SOR_TILE = 512;

__device__ __forceinline__ double np_pairwise_sum(const double* a, int n) {  // numpy's add.reduce for n <= 128
    if (n < 8) { double r = 0.0; for (int i = 0; i < n; i++) r = __dadd_rn(r, a[i]); return r; }
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) for (int j = 0; j < 8; j++) r[j] = __dadd_rn(r[j], a[i + j]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; i++) res = __dadd_rn(res, a[i]);
    return res;
}

__global__ void __launch_bounds__(SOR_THREADS) sor_kernel(const double* __restrict__ pts, int n, int k, double std_ratio,
                                                          int* __restrict__ keep) {
    __shared__ double tile[SOR_TILE * 3];
    const int i = blockIdx.x * SOR_THREADS + threadIdx.x;
    const bool act = i < n;
    double px = 0, py = 0, pz = 0;
    if (act) { px = pts[(size_t)i * 3]; py = pts[(size_t)i * 3 + 1]; pz = pts[(size_t)i * 3 + 2]; }
    double best[SOR_MAXK];  // ascending squared distances, k + 1 entries in use
    const int kk = k + 1;
    for (int q = 0; q < kk; q++) best[q] = __longlong_as_double(0x7ff0000000000000ll);  // +inf
    for (int t0 = 0; t0 < n; t0 += SOR_TILE) {
        const int tn = min(SOR_TILE, n - t0);
        __syncthreads();
        for (int q = threadIdx.x; q < tn * 3; q += SOR_THREADS) tile[q] = pts[(size_t)t0 * 3 + q];
        __syncthreads();
        if (!act) continue;
        for (int j = 0; j < tn; j++) {
            const double dx = __dsub_rn(px, tile[j * 3]), dy = __dsub_rn(py, tile[j * 3 + 1]), dz = __dsub_rn(pz, tile[j * 3 + 2]);
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (d2 < best[kk - 1]) {
                int q = kk - 1;
                while (q > 0 && best[q - 1] > d2) { best[q] = best[q - 1]; q--; }
                best[q] = d2;
            }
        }
    }
    if (!act) return;
    double d[SOR_MAXK];
    for (int q = 1; q < kk; q++) d[q - 1] = sqrt(best[q]);  // drop the first (the point itself); missing neighbours stay inf
    const double dk = (double)k;
    const double mean = __ddiv_rn(np_pairwise_sum(d, k), dk);
    for (int q = 0; q < k; q++) { const double e = __dsub_rn(d[q], mean); d[q] = __dmul_rn(e, e); }
    const double sd = sqrt(__ddiv_rn(np_pairwise_sum(d, k), dk));
    keep[i] = (mean < __dadd_rn(mean, __dmul_rn(std_ratio, sd))) ? 1 : 0;
}
__global__ void sor_compact_kernel(const double* __restrict__ pts, const int* __restrict__ keep, const int* __restrict__ pos, int n,
                                   double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    out[(size_t)pos[i] * 3] = pts[(size_t)i * 3]; out[(size_t)pos[i] * 3 + 1] = pts[(size_t)i * 3 + 1]; out[(size_t)pos[i] * 3 + 2] = pts[(size_t)i * 3 + 2];
}

int dev_outlier_removal(Lane& L, const double* pts, int n, int nb_neighbors, double std_ratio, double* out, int* nout_host) {
    *nout_host = 0;
    if (n <= 0) return L3D_OK;
    L3D_ARG(L, nb_neighbors >= 1 && nb_neighbors + 1 <= SOR_MAXK, "statistical_outlier_removal: 1 <= nb_neighbors <= 63");
    const int ntiles = cdiv(n, PS_TILE);
    int* ib = L.get<int>(S_IO_B, (size_t)2 * n + 4 * (ntiles + PS_TILE) + 16);
    int *keep = ib, *pos = keep + n, *scan_tmp = pos + n, *total = scan_tmp + 4 * (ntiles + PS_TILE);
    L3D_LAUNCH(L, sor_kernel, cdiv(n, SOR_THREADS), SOR_THREADS, 0, pts, n, nb_neighbors, std_ratio, keep);
    int rc = dev_exclusive_scan(L, keep, pos, n, scan_tmp, total);
    if (rc != L3D_OK) return rc;
    L3D_LAUNCH(L, sor_compact_kernel, cdiv(n, 256), 256, 0, pts, keep, pos, n, out);
    L3D_CHECK(L, cudaMemcpyAsync(nout_host, total, sizeof(int), cudaMemcpyDeviceToHost, L.stream));
    L3D_CHECK(L, cudaStreamSynchronize(L.stream));
    return L3D_OK;
}

}  // namespace l3d
