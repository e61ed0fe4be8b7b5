// l3d_all.cu -- single translation unit of libl3d.so (kernels launch each other's helpers, so
// everything is compiled together without relocatable device code).
#include "common.cuh"
#include "ccl.cuh"
#include "remap.cu"
#include "sgbm_vgroup.cu"
#include "sgbm.cu"
#include "post.cu"
#include "wls.cu"
#include "bm.cu"
#include "laser.cu"
#include "recon.cu"
#include "pointcloud.cu"
#include "api.cu"
