/* kmc_b200.h — C ABI of the B200-native LiDAR deskew (motion compensation) path.
 *
 * This is the drop-in boundary for ONE path of fracgawd/kitti_motion_compensation: the per-point deskew
 *   kmc::MotionCompensateFrame        include/kitti_motion_compensation/motion_compensation.hpp:13
 *   kmc::MotionCompensatePoint        include/kitti_motion_compensation/motion_compensation.hpp:10-11
 *   TrajectoryInterpolator            include/kitti_motion_compensation/trajectory_interpolation.hpp:11-31
 *   kmc::lie::{Exp,Log,...}           include/kitti_motion_compensation/lie_algebra.hpp:12-26
 *   kmc::GetPseudoTimeStamps          include/kitti_motion_compensation/timestamp_mocking.hpp:7-11
 * and the rows either side of it: KITTI .bin files as the kernel's I/O (data_io.hpp:19-83, data_io.cpp:287-313),
 * OxtsToPose / MakeFrame (data_io.hpp:13, 89-90), the whole-run handler (handlers.hpp:11) and the projection onto the
 * cameras (camera_model.hpp:7-11).
 * (citations are relative to the reference repository root).  The reference has no FFI of its own — its boundary is
 * a C++ shared library with Eigen types in the signatures — so this header is what a binding of that path would
 * bind: plain pointers, sizes and doubles, no C++/Eigen/torch types.  The headers under include/kitti_motion_compensation/ hold
 * the C++ mirror of the reference API implemented on top of these entry points.
 *
 * Conventions
 *   - A scan is the KITTI on-disk format (reference data_io.hpp:24-33): n points x 4 float32, "x y z i" interleaved,
 *     16 bytes per point.  Device pointers must be 16-byte aligned.
 *   - Poses are 4x4 doubles, COLUMN-major, i.e. exactly Eigen::Affine3d::matrix().data().
 *   - Times are doubles in seconds (kmc::Time, data_types.hpp:20).
 *   - Every function returns a kmc_b200_status: 0 == KMC_B200_OK, negative = error (nothing usable was produced unless
 *     the entry point says otherwise), positive = warning (all outputs valid).  Nothing here aborts or throws — C++
 *     exceptions raised inside the library (std::bad_alloc, thread creation) are caught at the boundary and returned as
 *     KMC_B200_ERR_INTERNAL; the C++ mirror re-creates the reference's assert-abort behaviour on top
 *     (trajectory_interpolation.cpp:9,32).
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Device entry points are
 *     asynchronous with respect to the host; *_host entry points return after the result is in the caller's memory.
 *   - There is NO CPU fallback: without a usable CUDA device the compute entry points return
 *     KMC_B200_ERR_CUDA / KMC_B200_ERR_NO_DEVICE.
 */
#ifndef KMC_B200_H_
#define KMC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMC_B200_VERSION 200 /* 0.2.0: round 2 — warnings (positive statuses), batch forms, checksums, file callback */

#if defined(__GNUC__)
#define KMC_B200_API __attribute__((visibility("default")))
#else
#define KMC_B200_API
#endif

typedef enum kmc_b200_status {
  /* Positive values are WARNINGS: the call did its work and every output is valid. */
  KMC_B200_WARN_ACCURACY = 1,           /* per-frame constants built, but the frame moves so fast that max |dxyz| < 1e-5 m is
                                           not guaranteed for points out to 120 m (see kmc_b200_frame_accuracy_bound) */
  KMC_B200_OK = 0,
  KMC_B200_ERR_NULL_POINTER = -1,
  KMC_B200_ERR_BAD_SIZE = -2,           /* negative n, n_frames, offsets not non-decreasing, misaligned pointer */
  KMC_B200_ERR_TIME_OUT_OF_RANGE = -3,  /* requested/point time outside [t_start, t_end]: the reference asserts here */
  KMC_B200_ERR_EMPTY_INTERVAL = -4,     /* t_end <= t_start (the reference divides by zero, :49-51) */
  KMC_B200_ERR_NOT_RIGID = -5,          /* pose with non-finite entries or det(linear) <= 0 */
  KMC_B200_ERR_CUDA = -6,               /* a CUDA runtime call failed; text in kmc_b200_last_error() */
  KMC_B200_ERR_NO_DEVICE = -7,
  KMC_B200_ERR_BAD_MODE = -8,
  KMC_B200_ERR_CAPACITY = -9,           /* scan larger than the handle's capacity */
  KMC_B200_ERR_IO = -10,                /* file could not be opened, read or written / size is not a multiple of 4 bytes */
  KMC_B200_ERR_INTERNAL = -11           /* a C++ exception (out of memory, thread creation) was caught at the C boundary */
} kmc_b200_status;

/* Where a point's position on the trajectory comes from. */
typedef enum kmc_b200_time_mode {
  /* w holds the intensity and passes through bit-exactly.  The fraction of the scan completed is computed in the
   * kernel from the azimuth: (pi - atan2(y, x)) / 2pi  (timestamp_mocking.cpp:46) — the fusion of
   * GetPseudoTimeStamps into the deskew kernel.  32 B/point of traffic, no per-point stamps in memory. */
  KMC_B200_TIME_FROM_AZIMUTH = 0,
  /* w holds the point's fraction of the trajectory x_i = (t_i - t_start)/(t_end - t_start) in [0, 1]
   * (trajectory_interpolation.cpp:49-51) and passes through unchanged.  Used by the C++ mirror of
   * MotionCompensateFrame, which must honour whatever LidarScan::timestamps holds (data_types.hpp:58). */
  KMC_B200_TIME_FROM_W = 1
} kmc_b200_time_mode;

/* Per-frame constants of the fused kernel: 64 bytes, computed once per frame on the host in double and rounded to
 * float.  With xi = [rho; phi] = Log(T_start^-1 T_end), theta = |phi|, a = phi/theta (0 if theta < 1e-12):
 *   phi[3], theta2 = theta^2
 *   rho_perp[3] = rho - a (a.rho),  c0 = 0.5 - x_req          (FROM_AZIMUTH: s = c0 - atan2(y,x)/2pi)
 *   rho_par[3]  = a (a.rho),        x_req = (t_req - t_start)/(t_end - t_start)   (FROM_W: s = w - x_req)
 *   phi_x_rho[3] = phi x rho,       wide = 1 when theta^2 > KMC_B200_SERIES_THETA2_MAX (kernel then uses the
 *                                   half-angle polynomials valid up to theta = pi), else 0
 * The correction applied to a point p captured at trajectory fraction x is  Exp((x - x_req) xi) p.
 *
 * ACCURACY DOMAIN of the fp32 kernels (float4 xyzi in, float4 out).  The displacement delta = Exp(s xi) p - p is computed
 * in fp32 and added to p once, so against the reference's double result
 *     |dxyz|  <=  ulp32(|p'|)/2  +  2.5e-7 |delta|  +  5e-8 (|rho| + theta |p|)
 * (output rounding at the magnitude of the coordinate: 3.8e-6 m for 64-128 m, 7.6e-6 m for 128-256 m; fp32 arithmetic on
 * the displacement; the azimuth polynomial and the rounding of s).  The 1e-5 m contract therefore holds for |delta| up to
 * ~20 m per scan with coordinates below 128 m (KITTI's HDL-64E reaches 120 m; a car at 30 m/s turning at 1 rad/s moves a
 * point at 120 m by 15 m) and for |delta| up to ~6 m with coordinates in 128-256 m; beyond 256 m a float32 coordinate itself cannot hold 1e-5 m.  kmc_b200_frame_accuracy_bound evaluates the
 * bound for a frame's constants, and kmc_b200_frame_params_from_* return KMC_B200_WARN_ACCURACY (constants still valid)
 * when the bound at 120 m exceeds 1e-5 m.  The reference-layout entry points (kmc_b200_deskew_cloud_f64_*) add the
 * displacement to the caller's doubles and have no output-rounding term.
 * FROM_W mode: w must lie in [0, 1] (the reference asserts on every point stamp, trajectory_interpolation.cpp:32); the
 * fp32 kernels do not check it — values outside extrapolate the motion, NaN gives NaN — use
 * kmc_b200_check_fractions_device to validate a buffer, or the f64 entry points, which flag out-of-range stamps. */
typedef struct kmc_b200_frame_params {
  float phi[3];
  float theta2;
  float rho_perp[3];
  float c0;
  float rho_par[3];
  float x_req;
  float phi_x_rho[3];
  float wide;
} kmc_b200_frame_params;

/* Largest theta^2 (rad^2 per scan) handled by the short in-kernel power series. */
#define KMC_B200_SERIES_THETA2_MAX 1.0f

/* Per-camera constants of the projection kernels (reference camera_model.cpp:5-36,38-95):
 *   rect = R_rect_00 * (R|T)_velo_to_cam            3x4 row-major, X_rect = rect * (x y z 1)
 *   pix  = P_rect_xx * [rect; 0 0 0 1]              3x4 row-major, (u v) = (pix*X)[0,1] / (pix*X)[2]
 * A point is culled when z_rect < min_depth (0.01), z_rect > max_range (15) or y_rect > max_below (1.25);
 * color_gain = 255 / (max_range - 0.01) is the reference's range colouring. */
typedef struct kmc_b200_camera_params {
  float rect[12];
  float pix[12];
  float min_depth;
  float max_range;
  float max_below;
  float color_gain;
} kmc_b200_camera_params;

typedef struct kmc_b200_handle kmc_b200_handle; /* opaque: device + stream + pinned/device staging (the "DataHandle") */

/* ---- library ------------------------------------------------------------------------------------------------ */
KMC_B200_API int kmc_b200_version(void);
KMC_B200_API const char* kmc_b200_status_string(int status);
/* Text of the last error raised on the calling thread (CUDA error strings included); never NULL. */
KMC_B200_API const char* kmc_b200_last_error(void);
/* Number of CUDA devices visible, or a negative status. */
KMC_B200_API int kmc_b200_device_count(void);
/* Kernels this library has launched in the calling process so far (every compute entry point ends in one). */
KMC_B200_API uint64_t kmc_b200_launch_count(void);

/* ---- host, double precision: the once-per-frame part of the path ----------------------------------------------
 * Replaces, per frame, what the reference recomputes for every point: Log(P1^-1 P2) incl. the polar projection of
 * T.rotation() (trajectory_interpolation.cpp:35, lie_algebra.cpp:94-103). */
KMC_B200_API int kmc_b200_frame_params_from_poses(const double T_start_colmajor[16], const double T_end_colmajor[16],
                                     double t_start, double t_end, double t_req, kmc_b200_frame_params* out);
/* Same from an explicit twist xi = [rho; phi] of the whole scan and the requested fraction x_req in [0,1]. */
KMC_B200_API int kmc_b200_frame_params_from_twist(const double xi[6], double x_req, kmc_b200_frame_params* out);

/* Upper bound (metres) of max |dxyz| of the fp32 kernels against the reference's double result for a frame with these
 * constants and points within max_range_m of the sensor (formula above).  Returns KMC_B200_WARN_ACCURACY when the bound
 * exceeds 1e-5 m, else KMC_B200_OK; *bound_m is written in both cases. */
KMC_B200_API int kmc_b200_frame_accuracy_bound(const kmc_b200_frame_params* params, double max_range_m, double* bound_m);

/* Lie algebra on the host (lie_algebra.hpp:12-26); 3x3 / 4x4 matrices column-major. */
KMC_B200_API int kmc_b200_so3_hat(const double phi[3], double out3x3[9]);
KMC_B200_API int kmc_b200_so3_vee(const double m3x3[9], double out[3]);
KMC_B200_API int kmc_b200_so3_exp(const double phi[3], double out3x3[9]);
KMC_B200_API int kmc_b200_so3_log(const double R3x3[9], double out[3]);
KMC_B200_API int kmc_b200_so3_left_jacobian(const double phi[3], double out3x3[9]);
KMC_B200_API int kmc_b200_so3_inverse_left_jacobian(const double phi[3], double out3x3[9]);
KMC_B200_API int kmc_b200_se3_exp(const double xi[6], double T_colmajor[16]);
KMC_B200_API int kmc_b200_se3_log(const double T_colmajor[16], double xi[6]);
/* TrajectoryInterpolator::GetPoseAtTime / RelativePoseBetweenTimes (trajectory_interpolation.hpp:17-19).
 * KMC_B200_ERR_TIME_OUT_OF_RANGE where the reference would abort. */
KMC_B200_API int kmc_b200_pose_at_time(double t1, const double P1[16], double t2, const double P2[16], double t, double out[16]);
KMC_B200_API int kmc_b200_relative_pose_between_times(double t1, const double P1[16], double t2, const double P2[16],
                                         double anchor_time, double query_time, double out[16]);
/* FractionOfScanCompleted / GetPseudoTimeStamp for one point (timestamp_mocking.hpp:7-9). */
KMC_B200_API double kmc_b200_fraction_of_scan_completed(double x, double y);
KMC_B200_API double kmc_b200_pseudo_time_stamp(double x, double y, double scan_start, double scan_end);

/* OxtsToPose (data_io.cpp:68-88): Mercator position (scale * R_earth * ...) and Rz(yaw) Ry(pitch) Rx(roll). */
KMC_B200_API int kmc_b200_oxts_to_pose(double lat, double lon, double alt, double roll, double pitch, double yaw, double scale,
                                       double T_colmajor[16]);

/* Camera constants from the KITTI calibration: P_rect_xx (3x4), R_rect_00 (3x3) and the velodyne -> camera_00 transform
 * (4x4), all column-major doubles (Eigen's .data()); max_range as in ProjectPointcloudOnImage (default 15 m). */
KMC_B200_API int kmc_b200_camera_params_from_calibration(const double P_rect_colmajor[12], const double R_rect_00_colmajor[9],
                                                         const double T_velo_to_cam_colmajor[16], double max_range,
                                                         kmc_b200_camera_params* out);

/* Contiguous split of n_items over n_parts (frame sharding across GPUs): part `index` owns [*begin, *end). */
KMC_B200_API int kmc_b200_shard_range(int64_t n_items, int32_t n_parts, int32_t index, int64_t* begin, int64_t* end);

/* ---- device entry points (pointers are DEVICE pointers on the current device) ---------------------------------- */
/* One frame.  Per-frame constants travel as a __grid_constant__ kernel parameter (constant bank).  in == out is
 * allowed (each thread reads then writes its own point). */
KMC_B200_API int kmc_b200_deskew_frame_device(const float* xyzi_in, float* xyzi_out, int64_t n_points,
                                 const kmc_b200_frame_params* params_host, int time_mode, void* stream);
/* A batch of n_frames frames stored back to back.  frame_offsets_dev has n_frames+1 non-decreasing int64 entries
 * (points, not bytes), frame_offsets[0] == 0 and frame_offsets[n_frames] == n_points_total; params_dev has n_frames
 * records.  Both tables live in device memory, so the host cannot verify them: PRECONDITION n_points_total ==
 * frame_offsets[n_frames] (points beyond the last frame's end are left untouched, the tables are never read out of
 * bounds). */
KMC_B200_API int kmc_b200_deskew_batch_device(const float* xyzi_in, float* xyzi_out, const int64_t* frame_offsets_dev,
                                 const kmc_b200_frame_params* params_dev, int32_t n_frames, int64_t n_points_total,
                                 int time_mode, void* stream);
/* MotionCompensateFrame on the reference's OWN data layout (motion_compensation.cpp:16-28; data_types.hpp:14,58): the
 * cloud and the result are COLUMN-major N x 4 doubles (Eigen::MatrixX4d::data()), stamps is the per-point time vector
 * (LidarScan::timestamps).  The displacement is computed in fp32 and added to the double coordinate, so the result
 * carries no float32 output rounding.  The 4th column is honoured as the reference does (Affine3d * Vector4d,
 * motion_compensation.cpp:13): p' = R p + t w, w passed through.  flags_dev (device int, caller-zeroed) receives bit 0 if
 * any stamp lies outside [t_start, t_end] (where the reference asserts) and bit 1 if a 4th-column entry is not the
 * homogeneous 1 (informational).  x_req = (t_req - t_start)/(t_end - t_start) must be the value params was built with.
 * out_colmajor may be cloud_colmajor itself (in place); any other overlap is not allowed.  Pointers need 8-byte alignment only
 * (16-byte aligned columns are moved with 128-bit accesses). */
KMC_B200_API int kmc_b200_deskew_cloud_f64_device(const double* cloud_colmajor, const double* stamps, double* out_colmajor,
                                                  int64_t n_points, double t_start, double t_end, double t_req,
                                                  const kmc_b200_frame_params* params_host, int* flags_dev, void* stream);
/* GetPseudoTimeStamps (timestamp_mocking.cpp:56-63) on the device, double precision: stamps[i] =
 * start + frac(x_i, y_i) * (end - start). */
KMC_B200_API int kmc_b200_pseudo_time_stamps_device(const float* xyzi_in, double* stamps_out, int64_t n_points, double scan_start,
                                       double scan_end, void* stream);
/* The same on the reference's own cloud layout: x and y are the first two COLUMNS of the column-major N x 4 double
 * matrix (Pointcloud::data() and data() + N), device pointers. */
KMC_B200_API int kmc_b200_pseudo_time_stamps_xy_device(const double* x, const double* y, double* stamps_out, int64_t n_points,
                                                       double scan_start, double scan_end, void* stream);
/* Projection of a scan onto one rectified camera — the per-point part of viz::ProjectPointcloudOnFrame/-OnImage
 * (camera_model.cpp:5-36,38-95); the cv::circle drawing stays with the caller.  For every point the kernel writes
 * one float4 (u, v, z_rect, c): pixel coordinates before the reference's int truncation, depth in front of the camera,
 * and c = the reference's colour scale 255 z/(max_range - 0.01) for kept points or -1 for culled ones. */
KMC_B200_API int kmc_b200_project_frame_device(const float* xyzi_in, float* uvzc_out, int64_t n_points,
                                               const kmc_b200_camera_params* camera_host, void* stream);
/* Deskew and projection of the DESKEWED point in one pass (48 B/point instead of 32 + 32): what
 * GenerateProjectionVisualizationOfRun does in two steps (handlers.cpp:83-87).  xyzi_out may be NULL to skip writing
 * the deskewed cloud (32 B/point). */
KMC_B200_API int kmc_b200_deskew_project_frame_device(const float* xyzi_in, float* xyzi_out, float* uvzc_out, int64_t n_points,
                                                      const kmc_b200_frame_params* params_host,
                                                      const kmc_b200_camera_params* camera_host, int time_mode, void* stream);
/* All four cameras of the KITTI rig in one pass (camera_model.cpp:85-92 projects the same cloud four times): the cloud is
 * read once and uvzc_out[c] receives camera c's records.  params_host may be NULL (projection of the input cloud as it
 * is); with params the points are deskewed first and, if xyzi_out is not NULL, the deskewed cloud is written too. */
KMC_B200_API int kmc_b200_deskew_project_frame4_device(const float* xyzi_in, float* xyzi_out, float* const uvzc_out[4],
                                                       int64_t n_points, const kmc_b200_frame_params* params_host,
                                                       const kmc_b200_camera_params cameras_host[4], int time_mode, void* stream);
/* Batch forms of the secondary entry points (the viz handler loops frames: handlers.cpp:67-92, one MotionCompensateFrame and
 * four projections per frame).  Frames back to back, frame_offsets_dev / params_dev as for kmc_b200_deskew_batch_device.
 *
 * Deskew + projection of the deskewed points onto n_cameras (1 or 4) cameras for every frame of the batch in one pass;
 * the cameras are the same for all frames (one calibration per run).  uvzc_out[c] receives n_points_total records of
 * camera c (see kmc_b200_project_frame_device); xyzi_out (optional) the deskewed cloud.  Bit-identical to the per-frame
 * calls. */
KMC_B200_API int kmc_b200_deskew_project_batch_device(const float* xyzi_in, float* xyzi_out, float* const uvzc_out[], int32_t n_cameras,
                                                      const int64_t* frame_offsets_dev, const kmc_b200_frame_params* params_dev,
                                                      int32_t n_frames, int64_t n_points_total,
                                                      const kmc_b200_camera_params* cameras_host, int time_mode, void* stream);
/* MotionCompensateFrame on the reference's layout for a batch: frame f owns the 4 N_f doubles at cloud + 4 offsets[f] (its
 * column-major N_f x 4 matrix), the N_f stamps at stamps + offsets[f], and the same block of out.  times_dev holds
 * (t_start, t_end, t_req) per frame (3 doubles each; the host must have validated t_start < t_end and t_req within, as
 * kmc_b200_frame_params_from_poses does); flags_dev receives one int per frame (zeroed by the call) with the bits of
 * kmc_b200_deskew_cloud_f64_device.  Bit-identical to the per-frame call. */
KMC_B200_API int kmc_b200_deskew_cloud_f64_batch_device(const double* cloud, const double* stamps, double* out,
                                                        const int64_t* frame_offsets_dev, const kmc_b200_frame_params* params_dev,
                                                        const double* times_dev, int32_t n_frames, int64_t n_points_total,
                                                        int* flags_dev, void* stream);
/* Validation pass for KMC_B200_TIME_FROM_W buffers: *flags_dev (device int, zeroed by the call) receives bit 0 if any
 * point's w lies outside [0, 1] or is NaN — the condition on which the reference asserts for every point stamp
 * (trajectory_interpolation.cpp:32,47).  Read-only, 16 B/point. */
KMC_B200_API int kmc_b200_check_fractions_device(const float* xyzi, int64_t n_points, int* flags_dev, void* stream);
/* Verification aid for sharded runs ("outputs for G = 8 bit-equal G = 1"): one 64-bit checksum per
 * frame of a batch stored back to back (frame_offsets_dev as for kmc_b200_deskew_batch_device).  Word j of a frame — the
 * bit patterns of its floats, counted from the frame's first point — contributes (bits + 0x9E3779B9) * (2 j + 1) mod 2^64;
 * the checksum is the wrapping sum, so it depends on every bit and position but not on how the work was split.
 * checksums_dev receives n_frames values (zeroed by the call). */
KMC_B200_API int kmc_b200_frame_checksums_device(const float* xyzi, const int64_t* frame_offsets_dev, int32_t n_frames,
                                                 int64_t n_points_total, uint64_t* checksums_dev, void* stream);
/* Seeded synthetic HDL-64E style scans written straight into device memory (benchmark input):
 * n_scans scans of points_per_scan points, scan k uses seed + first_scan_index + k, so a scan's content does not
 * depend on which GPU generates it.  n_rings x azimuth steps, ring-major, log-uniform range in [2, 120) m. */
KMC_B200_API int kmc_b200_synth_scans_device(float* xyzi_out, int64_t points_per_scan, int32_t n_scans, int32_t n_rings,
                                uint64_t seed, int64_t first_scan_index, void* stream);
/* Per-frame constants for synthetic frame `first_scan_index + k` (host, double -> float), the twist drawn from the
 * distribution in SURVEY 8d config 2.  xi_out (optional, 6 doubles per frame) receives the twist. */
KMC_B200_API int kmc_b200_synth_frame_params(int32_t n_frames, uint64_t seed, int64_t first_scan_index, double x_req,
                                kmc_b200_frame_params* params_out_host, double* xi_out);

/* ---- handle: device + stream + staging buffers (the KittiPclLoader-style owner, data_io.hpp:19-83) -------------- */
/* capacity_points: largest single transfer chunk the handle can stage (reference loader: 250 000, data_io.hpp:17). */
KMC_B200_API int kmc_b200_handle_create(int device, int64_t capacity_points, kmc_b200_handle** out);
KMC_B200_API int kmc_b200_handle_destroy(kmc_b200_handle* h);
/* Process-wide handle of a device, created on first use with the reference loader's capacity (250 000 points) and owned
 * by the library (do not destroy).  Used by the C++ mirror, whose reference signatures carry no handle. */
KMC_B200_API int kmc_b200_default_handle(int device, kmc_b200_handle** out);
/* Progress callback of the file pipelines (kmc_b200_deskew_bin_files, kmc_b200_motion_compensate_run): called once per
 * output file after it has been written, in file order, from the pipeline's retiring thread — the reference prints one
 * line per frame as its loop advances (handlers.cpp:63).  file_index counts the call's paths (for a run: frame id - 1).
 * NULL switches it off. */
typedef void (*kmc_b200_file_done_fn)(int32_t file_index, int64_t n_points, void* user);
KMC_B200_API int kmc_b200_handle_set_file_callback(kmc_b200_handle* h, kmc_b200_file_done_fn fn, void* user);
KMC_B200_API int kmc_b200_handle_device(const kmc_b200_handle* h);
KMC_B200_API int64_t kmc_b200_handle_capacity(const kmc_b200_handle* h);

/* ---- host entry points: HOST pointers, H2D + kernel + D2H inside the call ---------------------------------------- */
/* One frame (any n_points; processed in capacity-sized chunks, copies and kernels overlapped on the handle's
 * streams).  Pageable or pinned host memory. */
KMC_B200_API int kmc_b200_deskew_frame_host(kmc_b200_handle* h, const float* xyzi_in, float* xyzi_out, int64_t n_points,
                               const kmc_b200_frame_params* params, int time_mode);
/* A batch of frames stored back to back in host memory; frame_offsets/params are HOST tables (n_frames+1 / n_frames). */
KMC_B200_API int kmc_b200_deskew_batch_host(kmc_b200_handle* h, const float* xyzi_in, float* xyzi_out, const int64_t* frame_offsets,
                               const kmc_b200_frame_params* params, int32_t n_frames, int time_mode);
/* The same batch split into contiguous frame ranges over several devices, one host thread + handle per device, no
 * collective (frames are independent).  handles[i] must live on distinct devices. */
KMC_B200_API int kmc_b200_deskew_batch_multi_gpu(kmc_b200_handle* const* handles, int32_t n_handles, const float* xyzi_in,
                                    float* xyzi_out, const int64_t* frame_offsets, const kmc_b200_frame_params* params,
                                    int32_t n_frames, int time_mode);
/* The same from HOST memory (pageable or pinned), result into host memory.  The double cloud itself never crosses the
 * link: the handle's host threads round x, y, z to float and form every point's signed trajectory fraction
 * (t_i - t_start)/(t_end - t_start) - x_req in double; four float columns go up (16 B/point; a fifth only when some
 * w != 1), three displacement columns come back (12 B/point) and are added to the caller's doubles — bit-identical to
 * kmc_b200_deskew_cloud_f64_device.  Chunked over the handle's three streams so conversion, copies, kernel and the final
 * add overlap.  out == cloud (in place) is allowed.  Returns KMC_B200_ERR_TIME_OUT_OF_RANGE when a stamp was outside
 * [t_start, t_end] (the result is still written); *flags_out (optional) receives the raw bits described above. */
KMC_B200_API int kmc_b200_deskew_cloud_f64_host(kmc_b200_handle* h, const double* cloud_colmajor, const double* stamps,
                                                double* out_colmajor, int64_t n_points, double t_start, double t_end, double t_req,
                                                const kmc_b200_frame_params* params, int* flags_out);
/* The same for n_frames frames that live in separate host allocations (each reference Frame owns its Eigen matrices):
 * clouds[f] / stamps[f] / outs[f] as for the single-frame call with n_points[f] points, times holds (t_start, t_end, t_req)
 * per frame, params one record per frame.  One pipeline over all frames: the host passes, copies and kernels of
 * neighbouring frames overlap (the single-frame call drains its streams at the end of every frame).  flags_out
 * (optional, n_frames ints) receives each frame's bits; the status is the first frame's failure, if any, else OK. */
KMC_B200_API int kmc_b200_deskew_cloud_f64_batch_host(kmc_b200_handle* h, const double* const* clouds, const double* const* stamps,
                                                      double* const* outs, const int64_t* n_points, const double* times,
                                                      const kmc_b200_frame_params* params, int32_t n_frames, int* flags_out);
/* GetPseudoTimeStamps on host columns x, y (length n each): H2D + kernel + D2H. */
KMC_B200_API int kmc_b200_pseudo_time_stamps_xy_host(kmc_b200_handle* h, const double* x, const double* y, int64_t n_points,
                                                     double scan_start, double scan_end, double* stamps_out);
/* Projection of a host scan (n x 4 float32 xyzi) onto one camera: H2D + kernel + D2H, chunked like the deskew calls.
 * uvzc_out receives n x 4 float32 (u, v, z_rect, colour | -1), see kmc_b200_project_frame_device. */
KMC_B200_API int kmc_b200_project_frame_host(kmc_b200_handle* h, const float* xyzi_in, float* uvzc_out, int64_t n_points,
                                             const kmc_b200_camera_params* camera);
/* KITTI .bin in, deskewed .bin out (KittiPclLoader::LoadPointcloud + MotionCompensateFrame + WritePointcloud,
 * data_io.cpp:101-138, 287-313) without the float->double->float round trip, streamed through the handle's pinned slots
 * (any file size).  As the reference's loader, a size that is a multiple of 4 bytes is accepted and a trailing partial
 * point is dropped (data_io.cpp:107-112).  n_points_out may be NULL. */
KMC_B200_API int kmc_b200_deskew_bin_file(kmc_b200_handle* h, const char* path_in, const char* path_out,
                             const kmc_b200_frame_params* params, int64_t* n_points_out);

/* Many KITTI .bin files in one overlapped pass: files are packed in order into groups that fit one staging slot of the
 * handle (so create the handle with a capacity of several scans), and three slots rotate through
 * read -> H2D -> batched kernel -> D2H -> write with io_threads readers and writers (<= 0: min(16, host threads)).
 * This is the loop body of MotionCompensateRun (handlers.cpp:55-64: LoadSingleFrame + MotionCompensateFrame +
 * WritePointcloud per frame) for n_files frames at once; params[f] belongs to paths_in[f].  points_out (optional)
 * receives the number of points of every file.  A file larger than a staging slot is streamed through the slots in chunks on
 * its own (the pipeline drains around it). */
KMC_B200_API int kmc_b200_deskew_bin_files(kmc_b200_handle* h, int32_t n_files, const char* const* paths_in,
                                           const char* const* paths_out, const kmc_b200_frame_params* params, int time_mode,
                                           int32_t io_threads, int64_t* points_out);

typedef struct kmc_b200_run_stats {
  int64_t frames;           /* files found in velodyne_points/data */
  int64_t frames_deskewed;  /* frames 1 .. n-2 */
  int64_t points_deskewed;
  double seconds_prepare;   /* listing, time stamps, OxTS packets, poses, per-frame records */
  double seconds_pipeline;  /* kmc_b200_deskew_bin_files */
  double seconds_total;
} kmc_b200_run_stats;

/* The host half of MotionCompensateRun, no GPU involved: counts the frames of a KITTI raw run folder (every entry of
 * velodyne_points/data, handlers.cpp:15-17) and builds the per-frame kernel records of frames 1 .. n-2 — LoadTimeStamp /
 * LoadOxts / OxtsToPose / MakeFrame (data_io.cpp:18-88, 253-269) with every text file read once, then
 * kmc_b200_frame_params_from_poses with the camera-trigger time as the requested time (handlers.cpp:59).
 * params_out (capacity_frames records, may be NULL to only count) receives n_frames - 2 records; record k belongs to frame k + 1.
 * Statuses as kmc_b200_motion_compensate_run. */
KMC_B200_API int kmc_b200_run_prepare(const char* run_folder, int64_t capacity_frames, kmc_b200_frame_params* params_out,
                                      int64_t* n_frames_out);

/* MotionCompensateRun (handlers.cpp:41-65) on a KITTI raw run folder: reads oxts/timestamps.txt, every oxts/data packet and
 * velodyne_points/{timestamps_start,timestamps,timestamps_end}.txt once, builds every frame's start/end pose as MakeFrame
 * does (data_io.cpp:253-269), deskews frames 1 .. n-2 to their camera-trigger time through kmc_b200_deskew_bin_files into
 * velodyne_points/data_motion_compensated/, and copies frames 0 and n-1 through unchanged (the reference writes frame 0's
 * cloud under the last id, handlers.cpp:36-38; here the last file is the last frame's own cloud).  Missing or malformed
 * files are KMC_B200_ERR_IO; a scan stamp outside its OxTS interval — where the reference asserts — is
 * KMC_B200_ERR_TIME_OUT_OF_RANGE.  stats may be NULL. */
KMC_B200_API int kmc_b200_motion_compensate_run(kmc_b200_handle* h, const char* run_folder, int32_t io_threads,
                                                kmc_b200_run_stats* stats);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* KMC_B200_H_ */
