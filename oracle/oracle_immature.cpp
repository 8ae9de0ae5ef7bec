// oracle_immature.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into or called by the product).
//
// Restatement of SURVEY.md §8(f) row f4:
//   ImmaturePoint::ImmaturePoint   src/FullSystem/ImmaturePoint.cpp:32-66
//   ImmaturePoint::traceOn         src/FullSystem/ImmaturePoint.cpp:81-436
//   getInterpolatedElement33BiLin / 31 / 33   src/util/globalFuncs.h:166-188, 126-140, 75-89
//   pattern 8, settings            src/util/settings.cpp:99-100,146,165-174,297; settings.h:232-234
// Eigen expressions are restated with a fixed evaluation order (coefficient-wise, left to right):
//   M*Vec3f(u,v,1) = (m0*u + m1*v) + m2 ; a^T G b = (a0*G00 + a1*G10)*b0 + (a0*G01 + a1*G11)*b1 ; Vec3f*0.01 uses 0.01f.
// Parity: PINNED bit for bit to the reference's own ImmaturePoint constructor and traceOn copied verbatim at build time and
// compiled by `make ref` (oracle/ref_immature.cpp, tests/test_ref_pin.py, fixture tests/golden/ref_pin.npz); analytic KATs on top
// (tests/test_oracle_immature.py).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>

namespace {

const int kPattern[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};
enum { IPS_GOOD = 0, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED };

inline void interp33(const float* mat, float x, float y, int width, float* out3) {  // globalFuncs.h:75-89
  int ix = (int)x;
  int iy = (int)y;
  float dx = x - ix;
  float dy = y - iy;
  float dxdy = dx * dy;
  const float* bp = mat + 3 * (ix + iy * width);
  const float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
  for (int k = 0; k < 3; k++)
    out3[k] = ((w11 * bp[3 * (1 + width) + k] + w01 * bp[3 * width + k]) + w10 * bp[3 + k]) + w00 * bp[k];
}
inline float interp31(const float* mat, float x, float y, int width) {  // globalFuncs.h:126-140
  int ix = (int)x;
  int iy = (int)y;
  float dx = x - ix;
  float dy = y - iy;
  float dxdy = dx * dy;
  const float* bp = mat + 3 * (ix + iy * width);
  return ((dxdy * bp[3 * (1 + width)] + (dy - dxdy) * bp[3 * width]) + (dx - dxdy) * bp[3]) + (1 - dx - dy + dxdy) * bp[0];
}
inline void interp33BiLin(const float* mat, float x, float y, int width, float* out3) {  // globalFuncs.h:166-188
  int ix = (int)x;
  int iy = (int)y;
  const float* bp = mat + 3 * (ix + iy * width);
  float tl = bp[0], tr = bp[3], bl = bp[3 * width], br = bp[3 * (width + 1)];
  float dx = x - ix;
  float dy = y - iy;
  float topInt = dx * tr + (1 - dx) * tl;
  float botInt = dx * br + (1 - dx) * bl;
  float leftInt = dy * bl + (1 - dy) * tl;
  float rightInt = dy * br + (1 - dy) * tr;
  out3[0] = dx * rightInt + (1 - dx) * leftInt;
  out3[1] = rightInt - leftInt;
  out3[2] = botInt - topInt;
}

}  // namespace

extern "C" {

struct OracleTraceSettings {
  float maxPixSearch, stepsize, GNThreshold, extraSlackOnTH, slackInterval, minImprovementFactor, huberTH, outlierTH,
      outlierTHSumComponent, overallEnergyTHWeight;
  int GNIterations, minTraceTestRadius;
};

// ImmaturePoint constructor for n points of one host frame. dI = level-0 AoS {I,dx,dy}. u, v are integer pixel
// coordinates (stored as float, as in the reference). Outputs: color [n][8], weights [n][8], gradH [n][4] (row-major 2x2),
// energyTH [n] (NaN = the constructor bailed out on a non-finite colour; color/weights/gradH then hold the partial state).
void oracle_immature_init(int w, const float* dI, int n, const float* u, const float* v, const OracleTraceSettings* S, float* color,
                          float* weights, float* gradH, float* energyTH) {
  for (int i = 0; i < n; i++) {
    float* c = color + 8 * (size_t)i;
    float* wt = weights + 8 * (size_t)i;
    float* g = gradH + 4 * (size_t)i;
    g[0] = g[1] = g[2] = g[3] = 0;
    bool bail = false;
    for (int idx = 0; idx < 8; idx++) {
      float ptc[3];
      interp33BiLin(dI, u[i] + kPattern[idx][0], v[i] + kPattern[idx][1], w, ptc);
      c[idx] = ptc[0];
      if (!std::isfinite(c[idx])) {
        energyTH[i] = NAN;
        bail = true;
        break;
      }
      g[0] += ptc[1] * ptc[1];
      g[1] += ptc[1] * ptc[2];
      g[2] += ptc[2] * ptc[1];
      g[3] += ptc[2] * ptc[2];
      wt[idx] = sqrtf(S->outlierTHSumComponent / (S->outlierTHSumComponent + (ptc[1] * ptc[1] + ptc[2] * ptc[2])));
    }
    if (bail) continue;
    float e = 8 * S->outlierTH;
    e *= S->overallEnergyTHWeight * S->overallEnergyTHWeight;
    energyTH[i] = e;
  }
}

// FullSystem::makeNewTraces after makeMaps (FullSystem.cpp:1677-1687): raster scan of the selection map over
// x in [patternPadding+1, w-patternPadding-2), y likewise (patternPadding = 2, util/settings.h), one ImmaturePoint per
// non-zero entry, points whose constructor ends with a non-finite energyTH deleted. Outputs have capacity `cap` points;
// returns the number of points kept (may exceed cap: then only the first cap are written).
int oracle_make_new_traces(int w, int h, const float* dI, const float* map, int cap, const OracleTraceSettings* S, float* u, float* v,
                           float* type, float* color, float* weights, float* gradH, float* energyTH) {
  const int patternPadding = 2;
  int n = 0;
  for (int y = patternPadding + 1; y < h - patternPadding - 2; y++)
    for (int x = patternPadding + 1; x < w - patternPadding - 2; x++) {
      const int i = x + y * w;
      if (map[i] == 0) continue;
      const float fx = (float)x, fy = (float)y;
      float c[8] = {0, 0, 0, 0, 0, 0, 0, 0}, wt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, g[4], e;
      oracle_immature_init(w, dI, 1, &fx, &fy, S, c, wt, g, &e);
      if (!std::isfinite(e)) continue;
      if (n < cap) {
        u[n] = fx;
        v[n] = fy;
        type[n] = map[i];
        for (int k = 0; k < 8; k++) {
          color[8 * (size_t)n + k] = c[k];
          weights[8 * (size_t)n + k] = wt[k];
        }
        for (int k = 0; k < 4; k++) gradH[4 * (size_t)n + k] = g[k];
        energyTH[n] = e;
      }
      n++;
    }
  return n;
}

// traceOn for n points of one host frame into `frame` (level-0 AoS {I,dx,dy}, size w x h).
// In/out per point: idepth_min, idepth_max, quality, status, lastTraceUV [n][2], lastTracePixelInterval.
void oracle_immature_trace(int w, int h, const float* dI, int n, const float* pu, const float* pv, const float* color, const float* weights,
                           const float* gradH, const float* energyTH, const float* KRKi, const float* Kt, const float* aff,
                           const OracleTraceSettings* S, float* idepth_min_, float* idepth_max_, float* quality_, int* status_, float* lastTraceUV,
                           float* lastTracePixelInterval) {
  const float maxPixSearch = (w + h) * S->maxPixSearch;
  for (int p = 0; p < n; p++) {
    int& lastTraceStatus = status_[p];
    if (lastTraceStatus == IPS_OOB) continue;
    const float u = pu[p], v = pv[p];
    float& idepth_min = idepth_min_[p];
    float& idepth_max = idepth_max_[p];
    float& quality = quality_[p];
    float* UV = lastTraceUV + 2 * (size_t)p;
    float& interval = lastTracePixelInterval[p];
    const float* col = color + 8 * (size_t)p;
    const float* wts = weights + 8 * (size_t)p;
    const float* G = gradH + 4 * (size_t)p;
    auto oob = [&]() { UV[0] = -1; UV[1] = -1; interval = 0; lastTraceStatus = IPS_OOB; };

    float pr[3];
    for (int k = 0; k < 3; k++) pr[k] = (KRKi[3 * k] * u + KRKi[3 * k + 1] * v) + KRKi[3 * k + 2];
    float ptpMin[3];
    for (int k = 0; k < 3; k++) ptpMin[k] = pr[k] + Kt[k] * idepth_min;
    float uMin = ptpMin[0] / ptpMin[2];
    float vMin = ptpMin[1] / ptpMin[2];
    if (!(uMin > 4 && vMin > 4 && uMin < w - 5 && vMin < h - 5)) { oob(); continue; }

    float dist, uMax, vMax;
    float ptpMax[3];
    if (std::isfinite(idepth_max)) {
      for (int k = 0; k < 3; k++) ptpMax[k] = pr[k] + Kt[k] * idepth_max;
      uMax = ptpMax[0] / ptpMax[2];
      vMax = ptpMax[1] / ptpMax[2];
      if (!(uMax > 4 && vMax > 4 && uMax < w - 5 && vMax < h - 5)) { oob(); continue; }
      dist = (uMin - uMax) * (uMin - uMax) + (vMin - vMax) * (vMin - vMax);
      dist = sqrtf(dist);
      if (dist < S->slackInterval) {
        UV[0] = (uMax + uMin) * 0.5f;
        UV[1] = (vMax + vMin) * 0.5f;
        interval = dist;
        lastTraceStatus = IPS_SKIPPED;
        continue;
      }
    } else {
      dist = maxPixSearch;
      for (int k = 0; k < 3; k++) ptpMax[k] = pr[k] + Kt[k] * 0.01f;
      uMax = ptpMax[0] / ptpMax[2];
      vMax = ptpMax[1] / ptpMax[2];
      float dx = uMax - uMin;
      float dy = vMax - vMin;
      float d = 1.0f / sqrtf(dx * dx + dy * dy);
      uMax = uMin + dist * dx * d;
      vMax = vMin + dist * dy * d;
      if (!(uMax > 4 && vMax > 4 && uMax < w - 5 && vMax < h - 5)) { oob(); continue; }
    }
    if (!(idepth_min < 0 || (ptpMin[2] > 0.75f && ptpMin[2] < 1.5f))) { oob(); continue; }

    float dx = S->stepsize * (uMax - uMin);
    float dy = S->stepsize * (vMax - vMin);
    float a = (dx * G[0] + dy * G[2]) * dx + (dx * G[1] + dy * G[3]) * dy;
    float b = (dy * G[0] + (-dx) * G[2]) * dy + (dy * G[1] + (-dx) * G[3]) * (-dx);
    float errorInPixel = 0.2f + 0.2f * (a + b) / a;
    if (errorInPixel * S->minImprovementFactor > dist && std::isfinite(idepth_max)) {
      UV[0] = (uMax + uMin) * 0.5f;
      UV[1] = (vMax + vMin) * 0.5f;
      interval = dist;
      lastTraceStatus = IPS_BADCONDITION;
      continue;
    }
    if (errorInPixel > 10) errorInPixel = 10;

    dx /= dist;
    dy /= dist;
    if (dist > maxPixSearch) {
      uMax = uMin + maxPixSearch * dx;
      vMax = vMin + maxPixSearch * dy;
      dist = maxPixSearch;
    }
    int numSteps = 1.9999f + dist / S->stepsize;
    float randShift = uMin * 1000 - floorf(uMin * 1000);
    float ptx = uMin - randShift * dx;
    float pty = vMin - randShift * dy;
    float rot[8][2];
    for (int idx = 0; idx < 8; idx++) {
      rot[idx][0] = KRKi[0] * kPattern[idx][0] + KRKi[1] * kPattern[idx][1];
      rot[idx][1] = KRKi[3] * kPattern[idx][0] + KRKi[4] * kPattern[idx][1];
    }
    if (!std::isfinite(dx) || !std::isfinite(dy)) { interval = 0; UV[0] = -1; UV[1] = -1; lastTraceStatus = IPS_OOB; continue; }

    float errors[100];
    float bestU = 0, bestV = 0, bestEnergy = 1e10;
    int bestIdx = -1;
    if (numSteps >= 100) numSteps = 99;
    for (int i = 0; i < numSteps; i++) {
      float energy = 0;
      for (int idx = 0; idx < 8; idx++) {
        float hitColor = interp31(dI, (float)(ptx + rot[idx][0]), (float)(pty + rot[idx][1]), w);
        if (!std::isfinite(hitColor)) { energy += 1e5; continue; }
        float residual = hitColor - (float)(aff[0] * col[idx] + aff[1]);
        float hw = fabs(residual) < S->huberTH ? 1 : S->huberTH / fabs(residual);
        energy += hw * residual * residual * (2 - hw);
      }
      errors[i] = energy;
      if (energy < bestEnergy) { bestU = ptx; bestV = pty; bestEnergy = energy; bestIdx = i; }
      ptx += dx;
      pty += dy;
    }
    float secondBest = 1e10;
    for (int i = 0; i < numSteps; i++)
      if ((i < bestIdx - S->minTraceTestRadius || i > bestIdx + S->minTraceTestRadius) && errors[i] < secondBest) secondBest = errors[i];
    float newQuality = secondBest / bestEnergy;
    if (newQuality < quality || numSteps > 10) quality = newQuality;

    float uBak = bestU, vBak = bestV, gnstepsize = 1, stepBack = 0;
    if (S->GNIterations > 0) bestEnergy = 1e5;
    for (int it = 0; it < S->GNIterations; it++) {
      float H = 1, bb = 0, energy = 0;
      for (int idx = 0; idx < 8; idx++) {
        float hit[3];
        interp33(dI, (float)(bestU + rot[idx][0]), (float)(bestV + rot[idx][1]), w, hit);
        if (!std::isfinite((float)hit[0])) { energy += 1e5; continue; }
        float residual = hit[0] - (aff[0] * col[idx] + aff[1]);
        float dResdDist = dx * hit[1] + dy * hit[2];
        float hw = fabs(residual) < S->huberTH ? 1 : S->huberTH / fabs(residual);
        H += hw * dResdDist * dResdDist;
        bb += hw * residual * dResdDist;
        energy += wts[idx] * wts[idx] * hw * residual * residual * (2 - hw);
      }
      if (energy > bestEnergy) {
        stepBack *= 0.5f;
        bestU = uBak + stepBack * dx;
        bestV = vBak + stepBack * dy;
      } else {
        float step = -gnstepsize * bb / H;
        if (step < -0.5f) step = -0.5f;
        else if (step > 0.5f) step = 0.5f;
        if (!std::isfinite(step)) step = 0;
        uBak = bestU;
        vBak = bestV;
        stepBack = step;
        bestU += step * dx;
        bestV += step * dy;
        bestEnergy = energy;
      }
      if (fabsf(stepBack) < S->GNThreshold) break;
    }

    if (!(bestEnergy < energyTH[p] * S->extraSlackOnTH)) {
      interval = 0;
      UV[0] = -1; UV[1] = -1;
      lastTraceStatus = (lastTraceStatus == IPS_OUTLIER) ? IPS_OOB : IPS_OUTLIER;
      continue;
    }
    if (dx * dx > dy * dy) {
      idepth_min = (pr[2] * (bestU - errorInPixel * dx) - pr[0]) / (Kt[0] - Kt[2] * (bestU - errorInPixel * dx));
      idepth_max = (pr[2] * (bestU + errorInPixel * dx) - pr[0]) / (Kt[0] - Kt[2] * (bestU + errorInPixel * dx));
    } else {
      idepth_min = (pr[2] * (bestV - errorInPixel * dy) - pr[1]) / (Kt[1] - Kt[2] * (bestV - errorInPixel * dy));
      idepth_max = (pr[2] * (bestV + errorInPixel * dy) - pr[1]) / (Kt[1] - Kt[2] * (bestV + errorInPixel * dy));
    }
    if (idepth_min > idepth_max) std::swap(idepth_min, idepth_max);
    if (!std::isfinite(idepth_min) || !std::isfinite(idepth_max) || (idepth_max < 0)) {
      interval = 0;
      UV[0] = -1; UV[1] = -1;
      lastTraceStatus = IPS_OUTLIER;
      continue;
    }
    interval = 2 * errorInPixel;
    UV[0] = bestU;
    UV[1] = bestV;
    lastTraceStatus = IPS_GOOD;
  }
}

// ---- pin hooks (tests/test_ref_pin.py): the interpolation restatements above vs. the reference's util/globalFuncs.h
// compiled by `make ref`. mat3 is the reference's Eigen::Vector3f image ({I, dx, dy} per pixel), xy n pairs.
void oracle_pin_interp33(const float* mat3, int width, int n, const float* xy, float* out3) {
  for (int i = 0; i < n; i++) interp33(mat3, xy[2 * i], xy[2 * i + 1], width, out3 + 3 * i);
}
void oracle_pin_interp31(const float* mat3, int width, int n, const float* xy, float* out) {
  for (int i = 0; i < n; i++) out[i] = interp31(mat3, xy[2 * i], xy[2 * i + 1], width);
}
void oracle_pin_interp33bilin(const float* mat3, int width, int n, const float* xy, float* out3) {
  for (int i = 0; i < n; i++) interp33BiLin(mat3, xy[2 * i], xy[2 * i + 1], width, out3 + 3 * i);
}
}  // extern "C"
