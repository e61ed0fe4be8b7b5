// sgbm_vwave256.cu -- the wavefront aggregation (sgbm_vwave.cuh) at D = 256: 8 warps per CTA with 255 registers per thread,
// 16-CTA clusters, 10..13 columns per warp (BASELINE config 4: 1664 columns = 128 strips of 13)
#include "sgbm_vwave.cuh"

namespace l3d {

int vwave_launch_256(Lane& L, const VWaveArgs& a, int cpw, int nj, bool last, bool full, size_t smem) {
    switch (cpw) {
    case 10: return vwave_launch_shape<4, 10, 8>(L, a, nj, last, full, smem);
    case 11: return vwave_launch_shape<4, 11, 8>(L, a, nj, last, full, smem);
    case 12: return vwave_launch_shape<4, 12, 8>(L, a, nj, last, full, smem);
    case 13: return vwave_launch_shape<4, 13, 8>(L, a, nj, last, full, smem);
    default: return L3D_ERR_UNSUPPORTED;
    }
}

}  // namespace l3d
