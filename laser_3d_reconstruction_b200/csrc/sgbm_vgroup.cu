// sgbm_vgroup.cu -- the three "from the previous row" SGM paths of one pass, fused, in row lock-step
// on a thread-block cluster (modes SGBM / HH of cv2.StereoSGBM, camera/single_usb_stereo_camera.py:324-325).
//
// Pass 1 (top-down):  r1 from (x-1, y-1), r2 from (x, y-1), r3 from (x+1, y-1)
// Pass 2 (bottom-up): r1 from (x+1, y+1), r2 from (x, y+1), r3 from (x-1, y+1)   (MODE_HH only)
// Processing rows in travel order, all three predecessors live in the previously processed row, so a
// row of the cost volume is read ONCE and the row of S is read-modified-written ONCE for three paths
// (the direction-split kernels in sgbm.cu read C three times and read-modify-write S three times).
//
// One cluster of VG_CLUSTER CTAs owns one matcher run ("job"); the grid is (VG_CLUSTER, jobs), so the
// kernel wants many jobs per launch (frames x {left, right} matcher) to fill the GPU.  CTA r of a
// cluster owns a strip of VG_WARPS * cpw columns; warp w owns cpw adjacent columns whose path state
// (3 paths x cpw vectors + their minima) lives in registers for the whole pass.  A diagonal path's
// state moves one column per row: inside a warp that is a register rename done by updating in place in
// the right order; between warps of a CTA it goes through a double-buffered shared-memory slot; between
// CTAs the edge warps write the slot in the NEIGHBOUR's shared memory (DSMEM), and a split cluster
// barrier (arrive after the export, wait before the next row's import) orders both.
// Per row the C strip and the S strip (contiguous wc * 2D bytes each) arrive by one bulk async copy
// each (TMA, mbarrier completion), two rows ahead; the updated S strip leaves by one bulk store.
#include <cooperative_groups.h>

#include "common.cuh"
#include "sgm_step.cuh"

namespace l3d {
namespace cg = cooperative_groups;

constexpr int VG_CLUSTER_MAX = 16;  // cluster sizes used: 8 (portable) and 16 (non-portable)
constexpr int VG_WARPS = 16;     // warps per CTA (D <= 128); D = 256 runs 8 warps per CTA with up to 255 registers each
constexpr int VG_MAXCPW = 9;     // columns per warp at 16 warps: the cluster spans 8 * 16 * 9 = 1152 columns
constexpr int VG_MAXCPW8 = 13;   // columns per warp at 8 warps (D = 256): 16 * 8 * 13 = 1664 columns
constexpr int VG_MAXJOBS = 64;
// 16 warps x 120 registers leave 4096 registers of the SM free: the one-warp CTAs of the back-half kernels (FGS
// solver, ...) can then share an SM with an aggregation CTA instead of blocking a whole cluster from launching
#define VG_MAXREG 120
// interior neighbour words of the SGM step on the FMA pipe (sgm_step.cuh); only D = 256 has any that would pay
constexpr bool VG_FMAFUNNEL = false;

struct VGroupArgs {
    const int16_t* C[VG_MAXJOBS];
    int16_t* S[VG_MAXJOBS];
    int width1, H, D, P1, P2, cpw, dir;  // dir = +1 top-down (pass 1), -1 bottom-up (pass 2)
    int cluster;                         // CTAs per cluster (= per job)
    // final != 0: this is the last pass -- the winner-takes-all stage runs on the finished S rows while they
    // are still in shared memory and S is NOT written back (per job: raw disparity image, disp2 vote buffer,
    // minDisparity, minX1, uniquenessRatio; W = image width)
    int final, W;
    uint32_t zero;                       // 0, as a value the compiler cannot see (sgm_step.cuh)
    int16_t* raw[VG_MAXJOBS];
    unsigned* d2[VG_MAXJOBS];
    int minD[VG_MAXJOBS], minX1[VG_MAXJOBS], uniq[VG_MAXJOBS];
};

static size_t vgroup_smem_bytes(int D, int cpw, int nwarps) {
    const size_t strip = (size_t)nwarps * cpw * D * 2;
    const size_t halo = (size_t)2 * 2 * nwarps * (D * 2 + 16);
    return 2 * 2 * strip + halo + 64;  // + 6 mbarriers (2 row stages, 2 x 2 remote-halo parities)
}

__device__ __forceinline__ void vg_mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void vg_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void vg_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "VG_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra VG_DONE_%=;\n\t"
        "bra VG_WAIT_%=;\n\t"
        "VG_DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void vg_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void vg_bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

// asynchronous store into a NEIGHBOUR CTA's shared memory that signals the neighbour's mbarrier (DSMEM hand-off
// without any fence on the producer side)
__device__ __forceinline__ uint32_t vg_mapa(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void vg_st_async(uint32_t raddr, uint32_t v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(raddr), "r"(v), "r"(rbar) : "memory");
}
template <int NP> __device__ __forceinline__ void vg_st_async_vec(uint32_t raddr, const uint32_t (&v)[NP], uint32_t rbar);
template <> __device__ __forceinline__ void vg_st_async_vec<1>(uint32_t raddr, const uint32_t (&v)[1], uint32_t rbar) {
    vg_st_async(raddr, v[0], rbar);
}
template <> __device__ __forceinline__ void vg_st_async_vec<2>(uint32_t raddr, const uint32_t (&v)[2], uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
                 ::"r"(raddr), "r"(v[0]), "r"(v[1]), "r"(rbar) : "memory");
}
template <> __device__ __forceinline__ void vg_st_async_vec<4>(uint32_t raddr, const uint32_t (&v)[4], uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(raddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(rbar) : "memory");
}

template <int NP> struct VgVec;
template <> struct VgVec<1> { typedef uint32_t T; };
template <> struct VgVec<2> { typedef uint2 T; };
template <> struct VgVec<4> { typedef uint4 T; };
template <int NP> __device__ __forceinline__ void vg_unpack(const typename VgVec<NP>::T& v, uint32_t (&o)[NP]);
template <> __device__ __forceinline__ void vg_unpack<1>(const uint32_t& v, uint32_t (&o)[1]) { o[0] = v; }
template <> __device__ __forceinline__ void vg_unpack<2>(const uint2& v, uint32_t (&o)[2]) { o[0] = v.x; o[1] = v.y; }
template <> __device__ __forceinline__ void vg_unpack<4>(const uint4& v, uint32_t (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
template <int NP> __device__ __forceinline__ typename VgVec<NP>::T vg_pack(const uint32_t (&o)[NP]);
template <> __device__ __forceinline__ uint32_t vg_pack<1>(const uint32_t (&o)[1]) { return o[0]; }
template <> __device__ __forceinline__ uint2 vg_pack<2>(const uint32_t (&o)[2]) { return make_uint2(o[0], o[1]); }
template <> __device__ __forceinline__ uint4 vg_pack<4>(const uint32_t (&o)[4]) { return make_uint4(o[0], o[1], o[2], o[3]); }

// OpenCV's uniqueness test (modes SGBM / HH), see wta_not_unique in sgbm.cu; out of line, uniquenessRatio > 0 only
template <int NP>
__device__ __noinline__ bool vg_not_unique(typename VgVec<NP>::T wv, unsigned key, int uniq, unsigned dkey) {
    uint32_t w[NP];  // by value: a by-reference array would force the caller's S words through local memory every column
    vg_unpack<NP>(wv, w);
    const int minS = (int)(key >> 8), best = (int)(key & 255u);
    bool rej = false;
#pragma unroll
    for (int q = 0; q < NP; q++) {
        const int s0 = (int)(w[q] & 0xffffu), s1 = (int)(w[q] >> 16);
        const int d0 = (int)dkey + 2 * q;
        if (s0 * (100 - uniq) < minS * 100 && abs(best - d0) > 1) rej = true;
        if (s1 * (100 - uniq) < minS * 100 && abs(best - d0 - 1) > 1) rej = true;
    }
    return __any_sync(0xffffffffu, rej) && minS < 32767;
}

// NP = D / 64 words per lane (D in {64, 128, 256}: all 32 lanes hold disparities); CPW = columns per warp; NW = warps per
// CTA.  D = 256 at 1920 columns: a column's state is 15 registers, the 1664 columns of width1 spread over the largest
// cluster (16 CTAs) leave 104 per CTA -- 8 warps x 13 columns (195 state registers of the 255 a 256-thread CTA may
// use; each warp carries 39 independent recurrences per row, so two warps per sub-partition still fill the ALU pipe).
template <int NP, int CPW, bool FINAL, int NW>
__global__ void __launch_bounds__(NW * 32, 1) __maxnreg__(NW == 16 ? VG_MAXREG : 255) sgbm_vgroup_kernel(const VGroupArgs a) {
    constexpr int VG_WARPS = NW, VG_THREADS = NW * 32;
    typedef typename VgVec<NP>::T vec;
    constexpr uint32_t INF = 0x7fff7fffu;
    extern __shared__ __align__(128) unsigned char vg_smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int VG_CLUSTER = a.cluster;
    const int job = blockIdx.y;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    constexpr int cpw = CPW;
    const int width1 = a.width1, H = a.H;
    const uint32_t B = (uint32_t)a.D * 2u;                       // bytes per pixel vector
    const int wstrip = VG_WARPS * cpw;                           // columns per CTA
    const int x0 = rank * wstrip;                                // first column of this CTA
    const int wc = max(0, min(wstrip, width1 - x0));             // valid columns of this CTA
    const uint32_t strip_bytes = (uint32_t)wstrip * B;
    // smem: Cbuf[2] | Sbuf[2] | haloA[2][VG_WARPS] | haloB[2][VG_WARPS] | mbarriers[2]
    unsigned char* Cbuf = vg_smem;
    unsigned char* Sbuf = vg_smem + 2 * strip_bytes;
    const uint32_t slot = B + 16;                                // vector + packed minimum
    unsigned char* haloA = Sbuf + 2 * strip_bytes;               // state entering a warp from its LEFT neighbour
    unsigned char* haloB = haloA + 2 * VG_WARPS * slot;          // state entering a warp from its RIGHT neighbour
    const uint32_t bars = (uint32_t)__cvta_generic_to_shared(haloB + 2 * VG_WARPS * slot);
    const char* Cg = (const char*)a.C[job];
    char* Sg = (char*)a.S[job];
    const int dir = a.dir;
    constexpr bool final = FINAL;  // last pass: WTA on the finished S rows, S not written back
    const int uniq = a.uniq[job], minD = a.minD[job], minX1 = a.minX1[job];
    int16_t* rawg = a.raw[job];
    unsigned* d2g = a.d2[job];
    const unsigned dkey = (unsigned)(lane * 2 * NP);
    const unsigned dpair = dkey | ((dkey + 1u) << 8);

    // zero both parities of every halo slot: "no predecessor" = (L = 0, min = 0), OpenCV's out-of-image rule
    for (int i = threadIdx.x; i < (int)(4 * VG_WARPS * slot / 4); i += VG_THREADS) ((uint32_t*)haloA)[i] = 0u;
    // bars[0..1]: row stage full (expect_tx + the two bulk loads)
    // bars[2..3]: A state arrived from the left neighbour CTA (parity of the exporting row), bars[4..5]: B state from the right
    if (threadIdx.x == 0) {
        for (int i = 0; i < 6; i++) vg_mbar_init(bars + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto load_row = [&](int it) {  // thread 0: C and S strips of the it-th processed row into stage it & 1
        const int row = dir > 0 ? it : H - 1 - it;
        const uint32_t bytes = (uint32_t)wc * B;
        const uint32_t bar = bars + 8 * (it & 1);
        const size_t off = ((size_t)row * width1 + x0) * B;
        vg_mbar_expect_tx(bar, 2 * bytes);
        vg_bulk_g2s((uint32_t)__cvta_generic_to_shared(Cbuf + (it & 1) * strip_bytes), Cg + off, bytes, bar);
        vg_bulk_g2s((uint32_t)__cvta_generic_to_shared(Sbuf + (it & 1) * strip_bytes), Sg + off, bytes, bar);
    };
    if (threadIdx.x == 0 && wc > 0) {
        load_row(0);
        if (H > 1) load_row(1);
    }
    cluster.sync();  // halo zeroing of every CTA is complete before any neighbour writes into it

    // path state: A = diagonal fed from the left neighbour column, V = vertical, Bp = diagonal fed from the right
    uint32_t LA[CPW][NP], LV[CPW][NP], LB[CPW][NP];
    uint32_t mA[CPW], mV[CPW], mB[CPW];
#pragma unroll
    for (int j = 0; j < CPW; j++) {
#pragma unroll
        for (int k = 0; k < NP; k++) { LA[j][k] = 0; LV[j][k] = 0; LB[j][k] = 0; }
        mA[j] = 0; mV[j] = 0; mB[j] = 0;
    }
    const uint32_t p1x2 = (uint32_t)a.P1 * 0x10001u;
    const uint32_t k2 = (0x10000u - (uint32_t)a.P2) * 0x10001u;
    const SgmLane sl = sgm_lane_init(lane, a.zero);
    const int c0 = warp * cpw;  // first column of this warp inside the strip
    // where this warp's exports go: A state leaves to the right (warp + 1 or the next CTA's warp 0),
    // B state leaves to the left (warp - 1 or the previous CTA's last warp)
    unsigned char* expA = nullptr;   // local slot (same CTA)
    unsigned char* expB = nullptr;
    uint32_t rexpA = 0, rexpB = 0, rbarA = 0, rbarB = 0;  // remote slot + the neighbour's mbarrier (shared::cluster addresses)
    if (warp < VG_WARPS - 1) expA = haloA + (size_t)(warp + 1) * slot;
    else if (rank < VG_CLUSTER - 1) {
        rexpA = vg_mapa((uint32_t)__cvta_generic_to_shared(haloA), (uint32_t)(rank + 1));
        rbarA = vg_mapa(bars + 16, (uint32_t)(rank + 1));
    }
    if (warp > 0) expB = haloB + (size_t)(warp - 1) * slot;
    else if (rank > 0) {
        rexpB = vg_mapa((uint32_t)__cvta_generic_to_shared(haloB + (size_t)(VG_WARPS - 1) * slot), (uint32_t)(rank - 1));
        rbarB = vg_mapa(bars + 32, (uint32_t)(rank - 1));
    }
    const bool remote_inA = warp == 0 && rank > 0, remote_inB = warp == VG_WARPS - 1 && rank < VG_CLUSTER - 1;
    const uint32_t par_stride = VG_WARPS * slot;  // second parity of a halo array

    for (int it = 0; it < H; it++) {
        const int st = it & 1, par = it & 1;
        if (wc > 0) vg_mbar_wait(bars + 8 * st, (uint32_t)((it >> 1) & 1));
        if (it > 0) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // row it-1 exports visible
        // ---- import the diagonal states entering this warp (exported during row it-1 into parity (it-1)&1).
        // State from a neighbour CTA arrives by st.async and completes this CTA's mbarrier for that parity.
        if (it > 0 && (remote_inA || remote_inB)) {
            const uint32_t hb = bars + (remote_inA ? 16u : 32u) + 8u * (uint32_t)(par ^ 1);
            if (lane == 0) vg_mbar_expect_tx(hb, B + 4u);
            __syncwarp();
            vg_mbar_wait(hb, (uint32_t)(((it - 1) >> 1) & 1));
        }
        uint32_t inA[NP], inB[NP], inAm, inBm;
        {
            const unsigned char* pa = haloA + (size_t)(par ^ 1) * par_stride + (size_t)warp * slot;
            const unsigned char* pb = haloB + (size_t)(par ^ 1) * par_stride + (size_t)warp * slot;
            vg_unpack<NP>(*(const vec*)(pa + lane * sizeof(vec)), inA);
            vg_unpack<NP>(*(const vec*)(pb + lane * sizeof(vec)), inB);
            inAm = *(const uint32_t*)(pa + B);
            inBm = *(const uint32_t*)(pb + B);
        }
        const vec* cs = (const vec*)(Cbuf + st * strip_bytes) + (size_t)c0 * 32 + lane;  // cs[j * 32]: column j of this warp
        vec* ss = (vec*)(Sbuf + st * strip_bytes) + (size_t)c0 * 32 + lane;
        uint32_t Cw[NP];
        // ---- diagonal A (fed from the left): in place, right to left (column j reads column j-1's old state)
#pragma unroll
        for (int j = CPW - 1; j >= 1; j--) {
            vg_unpack<NP>(cs[j * 32], Cw);
            mA[j] = sgm_step<NP, VG_FMAFUNNEL>(LA[j], LA[j - 1], mA[j - 1], Cw, p1x2, k2, sl);
        }
        vg_unpack<NP>(cs[0], Cw);
        mA[0] = sgm_step<NP, VG_FMAFUNNEL>(LA[0], inA, inAm, Cw, p1x2, k2, sl);
        // ---- diagonal B (fed from the right): in place, left to right
#pragma unroll
        for (int j = 0; j < CPW - 1; j++) {
            vg_unpack<NP>(cs[j * 32], Cw);
            mB[j] = sgm_step<NP, VG_FMAFUNNEL>(LB[j], LB[j + 1], mB[j + 1], Cw, p1x2, k2, sl);
        }
        vg_unpack<NP>(cs[(CPW - 1) * 32], Cw);
        mB[CPW - 1] = sgm_step<NP, VG_FMAFUNNEL>(LB[CPW - 1], inB, inBm, Cw, p1x2, k2, sl);
        if (x0 + c0 + CPW > width1) {
            // columns outside the image: their state must read as "no predecessor" (0) for the last valid column
#pragma unroll
            for (int j = 0; j < CPW; j++) {
                if (x0 + c0 + j >= width1) {
#pragma unroll
                    for (int k = 0; k < NP; k++) LB[j][k] = 0;
                    mB[j] = 0;
                }
            }
        }
        // ---- export the new edge states (parity of this row).  Inside the CTA: plain shared-memory stores, made
        // visible by a CTA-scope fence before the (relaxed) cluster-barrier arrive -- the barrier then only orders
        // the reuse of the slots.  To a neighbour CTA: st.async + the neighbour's mbarrier, no fence at all.  (An
        // arrive.release would cost every warp a GPU-scope MEMBAR + ERRBAR per row: 18 % of the stall samples.)
        if (expA) {
            unsigned char* q = expA + (size_t)par * par_stride;
            *(vec*)(q + lane * sizeof(vec)) = vg_pack<NP>(LA[CPW - 1]);
            if (lane == 0) *(uint32_t*)(q + B) = mA[CPW - 1];
        } else if (rexpA && it + 1 < H) {
            const uint32_t q = rexpA + (uint32_t)par * par_stride, rb = rbarA + 8u * (uint32_t)par;
            vg_st_async_vec<NP>(q + lane * (uint32_t)sizeof(vec), LA[CPW - 1], rb);
            if (lane == 0) vg_st_async(q + B, mA[CPW - 1], rb);
        }
        if (expB) {
            unsigned char* q = expB + (size_t)par * par_stride;
            *(vec*)(q + lane * sizeof(vec)) = vg_pack<NP>(LB[0]);
            if (lane == 0) *(uint32_t*)(q + B) = mB[0];
        } else if (rexpB && it + 1 < H) {
            const uint32_t q = rexpB + (uint32_t)par * par_stride, rb = rbarB + 8u * (uint32_t)par;
            vg_st_async_vec<NP>(q + lane * (uint32_t)sizeof(vec), LB[0], rb);
            if (lane == 0) vg_st_async(q + B, mB[0], rb);
        }
        asm volatile("fence.acq_rel.cta;" ::: "memory");
        asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
        // ---- vertical path + S update (the cluster barrier's latency hides behind this)
        unsigned wkey = 0xffffffffu;  // final pass: arg-min key of column `lane` of this warp
        bool wrej = false;
#pragma unroll
        for (int j = 0; j < CPW; j++) {
            vg_unpack<NP>(cs[j * 32], Cw);
            mV[j] = sgm_step<NP, VG_FMAFUNNEL>(LV[j], LV[j], mV[j], Cw, p1x2, k2, sl);
            uint32_t Sw[NP];
            vg_unpack<NP>(ss[j * 32], Sw);
#pragma unroll
            for (int k = 0; k < NP; k++) {
                uint32_t s3 = __viaddmin_u16x2(LA[j][k], LB[j][k], INF);
                s3 = __viaddmin_u16x2(s3, LV[j][k], INF);
                Sw[k] = __viaddmin_u16x2(Sw[k], s3, INF);
            }
            ss[j * 32] = vg_pack<NP>(Sw);
            if (final) {  // first minimum wins (OpenCV's strict '<' scan over d): packed (cost << 8 | d) key
                // key = cost << 8 | d: one PRMT per disparity (cost bytes from the S word, d byte and the zero top byte
                // from a per-lane constant), the minima pair up into 3-input VIMNMX
                unsigned key = 0xffffffffu;
#pragma unroll
                for (int k = 0; k < NP; k++) {
                    const unsigned dk = dpair + 0x0202u * (unsigned)k;  // byte 0 = d, byte 1 = d + 1, bytes 2-3 = 0
                    key = min(key, min(__byte_perm(Sw[k], dk, 0x7104), __byte_perm(Sw[k], dk, 0x7325)));
                }
                key = __reduce_min_sync(0xffffffffu, key);
                bool rej = false;
                if (uniq > 0) rej = vg_not_unique<NP>(vg_pack<NP>(Sw), key, uniq, dkey);
                if (lane == j) { wkey = key; wrej = rej; }
            }
        }
        if (final) {
            // lane j finishes pixel j of this warp: disp2 vote, sub-pixel interpolation (neighbour costs from
            // the S row still in shared memory), store -- once per row for CPW pixels in parallel
            __syncwarp();
            const int x = x0 + c0 + lane;
            const int minS = (int)(wkey >> 8), d = (int)(wkey & 255u);
            if (lane < CPW && x < width1 && minS < 32767 && !wrej) {
                const int y = dir > 0 ? it : H - 1 - it;
                const int x2 = x + minX1 - d - minD;
                if (x2 >= 0 && x2 < a.W + 2)
                    atomicMax(d2g + (size_t)y * (a.W + 2) + x2, ((unsigned)(0x7fff - minS) << 16) | (unsigned)x);
                int dd = d * 16;
                if (0 < d && d < a.D - 1) {
                    const uint16_t* Sp = (const uint16_t*)(Sbuf + st * strip_bytes + (size_t)(c0 + lane) * B);
                    const int sm = Sp[d - 1], sp = Sp[d + 1];
                    const int denom2 = max(sm + sp - 2 * minS, 1);
                    dd += ((sm - sp) * 16 + denom2) / (denom2 * 2);
                }
                rawg[(size_t)y * a.W + x + minX1] = (int16_t)(dd + minD * 16);
            }
        }
        // ---- S strip back to HBM, next-but-one row in.  The bulk copies of row `it` are owned by warp it % 16
        // (a 17th producer warp was tried and lost more to the tighter register budget of a 544-thread CTA than
        // it gained).  Split named barrier: the other 15 warps only ARRIVE ("my reads and writes of this stage are
        // done") and run on into the next row; the owner waits for them, then stores / reloads the stage.  Its
        // lateness (the wait for the store's shared-memory read) is absorbed by the slack before the next cluster
        // barrier wait, and rotating the role keeps any one warp from falling behind row after row.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // named barriers 1 / 2 alternate by row (immediate ids: a register id would make ptxas reserve all 16 barriers)
        if (warp == (it & (VG_WARPS - 1))) {
            if (it & 1) asm volatile("bar.sync 2, %0;" ::"r"(VG_THREADS) : "memory");
            else asm volatile("bar.sync 1, %0;" ::"r"(VG_THREADS) : "memory");
            if (lane == 0 && wc > 0) {
                if (!final) {
                    const int row = dir > 0 ? it : H - 1 - it;
                    vg_bulk_s2g(Sg + ((size_t)row * width1 + x0) * B, (uint32_t)__cvta_generic_to_shared(Sbuf + st * strip_bytes),
                                (uint32_t)wc * B);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the stage's smem may be overwritten
                }
                if (it + 2 < H) load_row(it + 2);
            }
            __syncwarp();
        } else {
            if (it & 1) asm volatile("bar.arrive 2, %0;" ::"r"(VG_THREADS) : "memory");
            else asm volatile("bar.arrive 1, %0;" ::"r"(VG_THREADS) : "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every warp owned some rows' stores
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // pair the last arrive
    cluster.sync();  // no CTA exits while a neighbour may still write into its shared memory
}

// Host entry: aggregate the three previous-row paths of pass `dir` for `njobs` volumes at once.
// Returns L3D_ERR_UNSUPPORTED when the geometry does not fit (caller falls back to the scan kernels).
// launch shape for a geometry: CTAs per cluster (= per job), warps per CTA, columns per warp; false = not covered
struct VGShape { int cluster, nwarps, cpw; };
static bool vgroup_shape(int width1, int D, VGShape& s) {
    // 8 CTAs per job when the per-warp state fits (measured: more jobs resident, better throughput than 16);
    // 16 (non-portable) only for volumes wider than that
    const int cand[4][3] = {{8, 16, D == 256 ? 4 : VG_MAXCPW}, {16, 16, D == 256 ? 4 : VG_MAXCPW},
                            {8, 8, D == 256 ? VG_MAXCPW8 : 0}, {16, 8, D == 256 ? VG_MAXCPW8 : 0}};
    for (int i = 0; i < 4; i++) {
        const int cl = cand[i][0], nw = cand[i][1], maxcpw = cand[i][2];
        if (!maxcpw) continue;
        const int cpw = cdiv(width1, cl * nw);
        if (cpw <= maxcpw && vgroup_smem_bytes(D, cpw, nw) <= 227 * 1024) { s.cluster = cl; s.nwarps = nw; s.cpw = cpw; return true; }
    }
    return false;
}

bool vgroup_supported(int width1, int H, int D) {
    if (!(D == 64 || D == 128 || D == 256) || width1 < 1 || H < 1) return false;
    VGShape s;
    return vgroup_shape(width1, D, s);
}

int dev_sgbm_vgroup(Lane& L, const int16_t* const* C, int16_t* const* S, int njobs, int width1, int H, int D, int P1,
                    int P2, int dir, const VGroupWta* wta) {
    L3D_ARG(L, vgroup_supported(width1, H, D), "vgroup geometry");
    for (int j0 = 0; j0 < njobs; j0 += VG_MAXJOBS) {
        VGroupArgs a;
        const int nj = std::min(VG_MAXJOBS, njobs - j0);
        for (int j = 0; j < nj; j++) {
            a.C[j] = C[j0 + j]; a.S[j] = S[j0 + j];
            a.raw[j] = nullptr; a.d2[j] = nullptr; a.minD[j] = 0; a.minX1[j] = 0; a.uniq[j] = 0;
            if (wta) {
                a.raw[j] = wta[j0 + j].raw; a.d2[j] = wta[j0 + j].d2; a.minD[j] = wta[j0 + j].minD;
                a.minX1[j] = wta[j0 + j].minX1; a.uniq[j] = wta[j0 + j].uniq;
            }
        }
        a.final = wta ? 1 : 0; a.W = wta ? wta[0].W : 0;
        a.zero = 0;
        a.width1 = width1; a.H = H; a.D = D; a.P1 = P1; a.P2 = P2; a.dir = dir;
        VGShape shp;
        vgroup_shape(width1, D, shp);
        a.cluster = shp.cluster;
        a.cpw = shp.cpw;
        const size_t smem = vgroup_smem_bytes(D, a.cpw, shp.nwarps);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(a.cluster, nj);
        cfg.blockDim = dim3(shp.nwarps * 32);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = L.stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = a.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int rc = L3D_ERR_UNSUPPORTED;
#define VG_LAUNCH(KERN)                                                                                               \
    {                                                                                                                 \
        L3D_CHECK(L, cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
        L3D_CHECK(L, cudaFuncSetAttribute(KERN, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));                  \
        if (getenv("L3D_DEBUG_CLUSTERS")) {                                                                           \
            int ncl = -1;                                                                                             \
            cudaOccupancyMaxActiveClusters(&ncl, KERN, &cfg);                                                         \
            fprintf(stderr, "[l3d] %s cluster %d: max active clusters %d\n", #KERN, a.cluster, ncl);                 \
        }                                                                                                             \
        if (!dbg_skip("sgbm_vgroup_kernel")) L3D_CHECK(L, cudaLaunchKernelEx(&cfg, KERN, a));                         \
        L.launches++;                                                                                                 \
        rc = L3D_OK;                                                                                                  \
    }
#define VG_CASE_W(NPV, CPWV, NWV)                                                                                     \
    if (D == 64 * NPV && a.cpw == CPWV && shp.nwarps == NWV) {                                                        \
        if (a.final) VG_LAUNCH((sgbm_vgroup_kernel<NPV, CPWV, true, NWV>))                                            \
        else VG_LAUNCH((sgbm_vgroup_kernel<NPV, CPWV, false, NWV>))                                                   \
    }
#define VG_CASE(NPV, CPWV) VG_CASE_W(NPV, CPWV, 16)
#define VG_NP(NPV) VG_CASE(NPV, 1) VG_CASE(NPV, 2) VG_CASE(NPV, 3) VG_CASE(NPV, 4) VG_CASE(NPV, 5) VG_CASE(NPV, 6) \
                   VG_CASE(NPV, 7) VG_CASE(NPV, 8) VG_CASE(NPV, 9)
        VG_NP(1) VG_NP(2)
        VG_CASE(4, 1) VG_CASE(4, 2) VG_CASE(4, 3) VG_CASE(4, 4)
        VG_CASE_W(4, 5, 8) VG_CASE_W(4, 6, 8) VG_CASE_W(4, 7, 8) VG_CASE_W(4, 8, 8) VG_CASE_W(4, 9, 8) VG_CASE_W(4, 10, 8)
        VG_CASE_W(4, 11, 8) VG_CASE_W(4, 12, 8) VG_CASE_W(4, 13, 8)
#undef VG_CASE_W
#undef VG_NP
#undef VG_CASE
#undef VG_LAUNCH
        if (rc != L3D_OK) return rc;
    }
    return L3D_OK;
}

}  // namespace l3d
