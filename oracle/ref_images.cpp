// ref_images.cpp — TEST INFRASTRUCTURE ONLY. The reference's own FrameHessian::makeImages
// (src/FullSystem/HessianBlocks.cpp:127-190, row a1) compiled VERBATIM: HessianBlocks.cpp as a whole needs Sophus, PCL and
// the residual machinery, so `make ref` copies exactly this definition (and CalibHessian::getBGradOnly out of the header)
// into git-ignored intermediates under oracle/_ref/ and this file includes it inside namespace dso (see ref_extract.py,
// ref_standin/FullSystem/HessianBlocks.h).
// The reference leaves what it does not write uninitialised (`new Eigen::Vector3f[...]`: dx, dy and absSquaredGrad of the
// first and last image row of every level); this repository defines those as 0, so the driver
// below zeroes exactly those entries after the call and nothing else.
#include <cmath>
#include <cstring>

#include "FullSystem/HessianBlocks.h"  // stub (see there)
#include "util/globalCalib.h"
#include "util/settings.h"

namespace dso {
#include "make_images_extract.inc"
}  // namespace dso

using namespace dso;

extern "C" {
// color: [w*h]; B256: 256 floats or null (HCalib == 0); outputs concatenated over levels: dIp [tot][3], absgrad [tot]
void ref_pin_make_images(int w, int h, int levels, const float* color, const float* B256, float* dIp_out, float* ag_out) {
  Eigen::Matrix3f K;
  K << 500.0, 0.0, 0.5 * w, 0.0, 500.0, 0.5 * h, 0.0, 0.0, 1.0;
  setGlobalCalib(w, h, K);
  pyrLevelsUsed = levels;  // the caller's level count (setGlobalCalib derives its own from the size)
  for (int l = 0; l < levels; l++) { wG[l] = w >> l; hG[l] = h >> l; }
  CalibHessian calib;
  if (B256) { std::memcpy(calib.B, B256, sizeof(calib.B)); std::memset(calib.Binv, 0, sizeof(calib.Binv)); }
  FrameHessian fh;
  fh.makeImages(const_cast<float*>(color), B256 ? &calib : nullptr);
  size_t off = 0;
  for (int l = 0; l < levels; l++) {
    const int wl = wG[l], hl = hG[l];
    for (int i = 0; i < wl * hl; i++) {
      const bool written = i >= wl && i < wl * (hl - 1);  // the index range of the gradient loop
      dIp_out[3 * (off + i) + 0] = fh.dIp[l][i][0];
      dIp_out[3 * (off + i) + 1] = written ? fh.dIp[l][i][1] : 0.f;
      dIp_out[3 * (off + i) + 2] = written ? fh.dIp[l][i][2] : 0.f;
      ag_out[off + i] = written ? fh.absSquaredGrad[l][i] : 0.f;
    }
    off += (size_t)wl * hl;
    delete[] fh.dIp[l];
    delete[] fh.absSquaredGrad[l];
  }
}
}  // extern "C"
