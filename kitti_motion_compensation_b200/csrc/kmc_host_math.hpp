// kmc_host_math.hpp — double-precision, once-per-frame part of the deskew path (host side, no CUDA).
//
// The reference recomputes Log(P1^-1 P2) — including a 3x3 SVD polar projection and three general inverses — for every
// point (trajectory_interpolation.cpp:31-45, lie_algebra.cpp:94-103).  Here it is done once per frame and folded into
// the 64-byte kmc_b200_frame_params record the kernel consumes.
#pragma once

#include "kmc_b200.h"

namespace kmc_b200::host {

// All matrices column-major (Eigen layout): m3[c*3 + r], m4[c*4 + r].

void So3Hat(const double phi[3], double out[9]);                   // lie_algebra.cpp:7-18
void So3Vee(const double m[9], double out[3]);                     // lie_algebra.cpp:20
void So3Exp(const double phi[3], double out[9]);                   // lie_algebra.cpp:22-35
void So3Log(const double R[9], double out[3]);                     // lie_algebra.cpp:37-49
void So3LeftJacobian(const double phi[3], double out[9]);          // lie_algebra.cpp:51-65
void So3InverseLeftJacobian(const double phi[3], double out[9]);   // lie_algebra.cpp:67-81
void Se3Exp(const double xi[6], double T[16]);                     // lie_algebra.cpp:83-92
// lie_algebra.cpp:94-103 incl. the polar projection done by Eigen's Affine-mode rotation().
// Returns false when the linear block has no proper-rotation polar factor (det <= 0 or non-finite).
bool Se3Log(const double T[16], double xi[6]);

// General (Affine-mode) inverse and product of 4x4 affine transforms, as Eigen does them.
bool AffineInverse(const double T[16], double out[16]);
void AffineMul(const double A[16], const double B[16], double out[16]);

// Log(P1^-1 P2) with the translation difference formed before the inverse is applied (keeps the ~6e6 m Mercator
// magnitudes of KITTI poses out of the cancellation).
bool RelativeTwist(const double P1[16], const double P2[16], double xi[6]);

// TrajectoryInterpolator::GetPoseAtTime (trajectory_interpolation.cpp:31-41) without the abort.
int PoseAtTime(double t1, const double P1[16], double t2, const double P2[16], double t, double out[16]);

// Per-frame kernel constants from the scan twist and the requested fraction.
void FrameParamsFromTwist(const double xi[6], double x_req, kmc_b200_frame_params* out);

double FractionOfScanCompleted(double x, double y);

// OxtsToPose (data_io.cpp:68-88): Mercator projection for the position, Rz(yaw) Ry(pitch) Rx(roll) for the attitude.
void OxtsToPose(double lat, double lon, double alt, double roll, double pitch, double yaw, double scale, double T[16]);                // timestamp_mocking.cpp:46

// rect = R_rect_00 * T_velo_to_cam (3x4), pix = P_rect * [rect; 0 0 0 1] (3x4), composed in double
// (camera_model.cpp:9,75,78-81); inputs column-major.
void CameraParamsFromCalibration(const double P_rect[12], const double R_rect_00[9], const double T_velo_to_cam[16],
                                 double max_range, kmc_b200_camera_params* out);

}  // namespace kmc_b200::host
