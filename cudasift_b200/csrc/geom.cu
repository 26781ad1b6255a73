// geom.cu -- ImproveHomography: host-side refinement of a RANSAC homography.
//
// Reference: geomFuncs.cpp:6-72.  The reference runs this on the host copy
// of the records (data.h_data) with OpenCV's cv::Mat / cv::solve(DECOMP_CHOLESKY); it is host
// code there and host code here -- the only change is that the 8x8 normal equations are
// solved by the small Cholesky routine below instead of OpenCV.
//
// Per iteration: every match that passes the score/ambiguity gates and whose reprojection
// error under the current estimate is below thresh contributes its two DLT rows
//   [x y 1 0 0 0 -x*u -y*u | u]   and   [0 0 0 x y 1 -x*v -y*v | v]
// to M = sum r r^T, b = sum r*rhs (double accumulation; the gate arithmetic is float, as in
// the reference, geomFuncs.cpp:29-33).  After the loops match_error of every record is set.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace {

// In-place Cholesky solve of the symmetric 8x8 system M a = b.  Returns false when M is not
// positive definite (cv::solve then leaves a zero solution, OpenCV lapack.cpp).
bool solve_spd8(double M[8][8], const double b[8], double a[8])
{
  double L[8][8];
  for (int j = 0; j < 8; j++) {
    double d = M[j][j];
    for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    L[j][j] = d;
    for (int i = j + 1; i < 8; i++) {
      double s = M[i][j];
      for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
      L[i][j] = s / d;
    }
  }
  double y[8];
  for (int i = 0; i < 8; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[i][k] * y[k];
    y[i] = s / L[i][i];
  }
  for (int i = 7; i >= 0; i--) {
    double s = y[i];
    for (int k = i + 1; k < 8; k++) s -= L[k][i] * a[k];
    a[i] = s / L[i][i];
  }
  return true;
}

// Squared reprojection error with the reference's mixed precision: products in double, the
// denominator, the two residuals and their squared sum rounded to float (geomFuncs.cpp:29-32).
inline float reproj_err2(const double a[8], const SiftPoint &pt)
{
  float den = (float)(a[6] * pt.xpos + a[7] * pt.ypos + 1.0);
  float dx = (float)((a[0] * pt.xpos + a[1] * pt.ypos + a[2]) / den - pt.match_xpos);
  float dy = (float)((a[3] * pt.xpos + a[4] * pt.ypos + a[5]) / den - pt.match_ypos);
  return dx * dx + dy * dy;
}

inline void add_row(double M[8][8], double b[8], const double r[8], double rhs)
{
  for (int i = 0; i < 8; i++) {
    if (r[i] == 0.0) continue;
    for (int j = 0; j < 8; j++) M[i][j] += r[i] * r[j];
    b[i] += r[i] * rhs;
  }
}

int improve_homography(SiftPoint *pts, int numPts, float *homography, int numLoops, float minScore,
                       float maxAmbiguity, float thresh)
{
  const float limit = thresh * thresh;
  double a[8];
  for (int i = 0; i < 8; i++) a[i] = homography[i] / homography[8];   // float division, geomFuncs.cpp:21
  for (int loop = 0; loop < numLoops; loop++) {
    double M[8][8], b[8];
    std::memset(M, 0, sizeof(M));
    std::memset(b, 0, sizeof(b));
    for (int i = 0; i < numPts; i++) {
      const SiftPoint &pt = pts[i];
      if (pt.score < minScore || pt.ambiguity > maxAmbiguity) continue;
      if (!(reproj_err2(a, pt) < limit)) continue;               // weight 0
      const double x = pt.xpos, y = pt.ypos, u = pt.match_xpos, v = pt.match_ypos;
      const double ru[8] = {x, y, 1.0, 0.0, 0.0, 0.0, -(pt.xpos * pt.match_xpos), -(pt.ypos * pt.match_xpos)};
      const double rv[8] = {0.0, 0.0, 0.0, x, y, 1.0, -(pt.xpos * pt.match_ypos), -(pt.ypos * pt.match_ypos)};
      add_row(M, b, ru, u);
      add_row(M, b, rv, v);
    }
    double sol[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    solve_spd8(M, b, sol);
    for (int i = 0; i < 8; i++) a[i] = sol[i];
  }
  int numfit = 0;
  for (int i = 0; i < numPts; i++) {
    float err = reproj_err2(a, pts[i]);
    if (err < limit) numfit++;
    pts[i].match_error = std::sqrt(err);
  }
  for (int i = 0; i < 8; i++) homography[i] = (float)a[i];
  homography[8] = 1.0f;
  return numfit;
}

}  // namespace

// geomFuncs.cpp:6 (declared by the caller in the reference, mainSift.cpp:16).
int ImproveHomography(SiftData &data, float *homography, int numLoops, float minScore, float maxAmbiguity,
                      float thresh)
{
#ifdef MANAGEDMEM
  SiftPoint *pts = data.m_data;
#else
  SiftPoint *pts = data.h_data;
#endif
  if (pts == NULL) return 0;
  return improve_homography(pts, data.numPts, homography, numLoops, minScore, maxAmbiguity, thresh);
}

extern "C" int cs_improve_homography(void *h_pts, int numPts, float *homography, int numLoops, float minScore,
                                     float maxAmbiguity, float thresh, int *numFit)
{
  if (!h_pts || !homography) { cs::set_error("cs_improve_homography: null argument"); return CS_E_ARG; }
  int n = improve_homography((SiftPoint *)h_pts, numPts, homography, numLoops, minScore, maxAmbiguity, thresh);
  if (numFit) *numFit = n;
  return 0;
}
