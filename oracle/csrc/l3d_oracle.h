/*
 * l3d_oracle.h -- CPU restatement (plain C) of the per-frame dense vision hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in laser_3d_reconstruction_b200/ may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg use it, and only as the checker.
 *
 * The reference (alo-i-sia/laser_3d_reconstruction) is pure Python; its arithmetic lives in the
 * third-party dependency OpenCV (requirements.txt:7 "opencv-python>=4.5.0", unpinned; the oracle
 * is pinned against the binary in this image: opencv-python-headless 4.13.0.92) and, for WLS, in
 * opencv_contrib/ximgproc (absent from the image: parity unpinned for that stage).
 * Each function cites the reference call site it restates.
 */
#ifndef L3D_ORACLE_H
#define L3D_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cv2.cvtColor(BGR2GRAY): camera/single_usb_stereo_camera.py:320-321, core/laser_extractor.py:60 */
void orc_bgr2gray(const uint8_t* bgr, long n, uint8_t* gray);
/* cv2.cvtColor(BGR2HSV): core/laser_extractor.py:56, improved_steger.py:255 */
void orc_bgr2hsv(const uint8_t* bgr, long n, uint8_t* hsv);
/* cv2.remap(img, mapx, mapy, INTER_LINEAR) with CV_32FC1 maps, BORDER_CONSTANT(0):
 * camera/single_usb_stereo_camera.py:313-314 */
void orc_remap_bilinear(const uint8_t* src, int sw, int sh, int cn, const float* mapx,
                        const float* mapy, int dw, int dh, uint8_t* dst);

typedef struct {
    int minDisparity, numDisparities, blockSize, P1, P2, disp12MaxDiff, preFilterCap,
        uniquenessRatio, speckleWindowSize, speckleRange, mode; /* 0 SGBM, 1 HH, 2 3WAY, 3 HH4 */
} orc_sgbm_params;

/* cv2.StereoSGBM.compute(left, right): camera/single_usb_stereo_camera.py:252-274 (params),
 * :324-325 (calls).  disp = int16 HxW, x16 fixed point.  C_out/S_out (optional, may be NULL) get
 * the H*width1*D cost volume / aggregated volume (modes 0,1 only; S as seen by the WTA).
 * raw_out (optional) gets the disparity before medianBlur/filterSpeckles.  Returns 0 on success. */
int orc_sgbm_compute(const uint8_t* left, const uint8_t* right, int W, int H,
                     const orc_sgbm_params* p, int16_t* disp, int16_t* raw_out, int16_t* C_out,
                     int16_t* S_out);
/* cv2.medianBlur(disp,3) on int16 (inside StereoSGBM.compute) */
void orc_median3_s16(const int16_t* src, int W, int H, int16_t* dst);
/* cv2.filterSpeckles (inside StereoSGBM.compute) */
void orc_filter_speckles(int16_t* img, int W, int H, int newVal, int maxSize, int maxDiff);

/* cv2.ximgproc DisparityWLSFilter.filter(dl, guide, disparity_map_right=dr):
 * camera/single_usb_stereo_camera.py:277-282 (setup), :328-332 (call).  PARITY UNPINNED (module
 * not installed; restated from the published opencv_contrib algorithm).  conf_out optional. */
void orc_wls_filter(const int16_t* dl, const int16_t* dr, const uint8_t* guide, int W, int H,
                    int min_disp, int num_disp, int dd_radius, double lambda, double sigma_color,
                    int lrc_thresh, int16_t* out, float* conf_out);
/* the same with the unpinned points of the restatement switchable (bits as L3D_WLS_* of include/l3d.h) */
enum { ORC_WLS_LAMBDA_PER_PASS = 1, ORC_WLS_LRC_OUTSIDE_ZERO = 2, ORC_WLS_BOX_FULL_IMAGE = 4, ORC_WLS_CONF_CLAMP_1 = 8 };
void orc_wls_filter_v(const int16_t* dl, const int16_t* dr, const uint8_t* guide, int W, int H,
                      int min_disp, int num_disp, int dd_radius, double lambda, double sigma_color,
                      int lrc_thresh, int variant, int16_t* out, float* conf_out);

/* disparity(int16 x16) -> depth: camera/single_usb_stereo_camera.py:335-346 (Q branch) */
void orc_disp_to_depth_q(const int16_t* disp16, int W, int H, const double* Q, float* depth);
/* camera/single_usb_stereo_camera.py:347-357 (no-calibration branch) */
void orc_disp_to_depth_default(const int16_t* disp16, int W, int H, float* depth);

/* Simple extractor mask chain, core/laser_extractor.py:56-82.  mask_morph = after CLOSE/OPEN (:69),
 * mask_final = after contour-area filter + filled drawContours (:81-82). */
void orc_simple_masks(const uint8_t* bgr, int W, int H, const int* hsv_lo, const int* hsv_hi,
                      int bright_thr, double min_area, uint8_t* mask_morph, uint8_t* mask_final);
/* 3x3 rectangular CLOSE then OPEN on a 0/255 mask (cv2.morphologyEx pair) */
void orc_close_open3(const uint8_t* src, int W, int H, uint8_t* dst);

/* cv2.GaussianBlur(f32,(0,0),sigma), BORDER_REFLECT_101: core/laser_extractor.py:193, improved_steger.py:59 */
void orc_gaussian_blur_f32(const float* src, int W, int H, double sigma, float* dst);
/* cv2.Sobel(f32, CV_32F, dx, dy, ksize=3), REFLECT_101: improved_steger.py:63-69 */
void orc_sobel3_f32(const float* src, int W, int H, int dx, int dy, float* dst);

/* cv2.StereoBM (readme.md:392-397 suggests it as a drop-in matcher); see orc_bm.c for the supported subset */
typedef struct {
    int minDisparity, numDisparities, blockSize, preFilterCap, textureThreshold, uniquenessRatio, speckleWindowSize,
        speckleRange, disp12MaxDiff;
} orc_bm_params;
void orc_bm_prefilter_xsobel(const uint8_t* src, int W, int H, int ftzero, uint8_t* dst);
int orc_bm_compute(const uint8_t* left, const uint8_t* right, int W, int H, const orc_bm_params* p, int16_t* disp);

#ifdef __cplusplus
}
#endif
#endif
