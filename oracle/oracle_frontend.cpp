// oracle/oracle_frontend.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product path).
//
// CPU restatement of the reference's per-frame front end:
//   a1  FrameHessian::makeImages            src/FullSystem/HessianBlocks.cpp:127-190
//       CalibHessian::getBGradOnly          src/FullSystem/HessianBlocks.h:402-408
//   a2  PixelSelector::makeHists            src/FullSystem/PixelSelector2.cpp:78-143
//       computeHistQuantil                  src/FullSystem/PixelSelector2.cpp:66-75
//   a3  PixelSelector::select               src/FullSystem/PixelSelector2.cpp:564-707
//   a4  PixelSelector::makeMaps             src/FullSystem/PixelSelector2.cpp:144-291
//       PixelSelector ctor (randomPattern)  src/FullSystem/PixelSelector2.cpp:41-56
// Pyramid sizes follow src/util/globalCalib.cpp:89-95 (w>>lvl, h>>lvl) with the level count FORCED by the
// caller (the reference's own setGlobalCalib would stop at one level for odd 1241).
//
// Defined behaviour where the reference reads uninitialised memory:
//   * dIp[l][.][1..2] and absSquaredGrad on the first/last image row are 0.
//   * ths / thsSmoothed are zero-initialised, so the ragged bottom strip (y >= 32*h32) and the wrapped
//     column (x >= 32*w32) read whatever the reference's index formula lands on, with never-written
//     slots being 0.
// Parity: the PixelSelector part (constructor / randomPattern, makeHists, select, makeMaps) is PINNED bit for bit to the
// reference's own FullSystem/PixelSelector2.cpp compiled by `make ref` (oracle/_ref; tests/test_ref_pin.py, fixture
// tests/golden/ref_pin.npz), and so is makeImages (the reference's FrameHessian::makeImages copied verbatim at build time,
// oracle/ref_images.cpp). Analytic KATs on top: tests/test_oracle_frontend.py. Build with -ffp-contract=off.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

inline bool is_fin(float v) { return std::isfinite(v); }

}  // namespace

extern "C" {

// (w>>l, h>>l) — src/util/globalCalib.cpp:89-95, src/FullSystem/CoarseTracker.cpp:126-129
void oracle_pyr_sizes(int w0, int h0, int levels, int* w, int* h) {
  for (int l = 0; l < levels; l++) { w[l] = w0 >> l; h[l] = h0 >> l; }
}

// a1. dIp: concatenated per-level AoS {I,dx,dy} (3 floats / px); absgrad: concatenated per-level.
// B256 may be null (== HCalib==0 or setting_gammaWeightsPixelSelect!=1).
void oracle_make_images(int w0, int h0, int levels, const float* color, const float* B256, float* dIp,
                        float* absgrad) {
  int wl[8], hl[8];
  oracle_pyr_sizes(w0, h0, levels, wl, hl);
  size_t off[8];
  size_t tot = 0;
  for (int l = 0; l < levels; l++) { off[l] = tot; tot += (size_t)wl[l] * hl[l]; }
  memset(dIp, 0, sizeof(float) * 3 * tot);
  memset(absgrad, 0, sizeof(float) * tot);

  float* dI0 = dIp;
  for (int i = 0; i < w0 * h0; i++) dI0[3 * i] = color[i];

  for (int lvl = 0; lvl < levels; lvl++) {
    const int w = wl[lvl], h = hl[lvl];
    float* dI_l = dIp + 3 * off[lvl];
    float* dabs_l = absgrad + off[lvl];
    if (lvl > 0) {
      const int wlm1 = wl[lvl - 1];
      const float* dI_lm = dIp + 3 * off[lvl - 1];
      for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
          dI_l[3 * (x + y * w)] = 0.25f * (dI_lm[3 * (2 * x + 2 * y * wlm1)] + dI_lm[3 * (2 * x + 1 + 2 * y * wlm1)] +
                                            dI_lm[3 * (2 * x + 2 * y * wlm1 + wlm1)] +
                                            dI_lm[3 * (2 * x + 1 + 2 * y * wlm1 + wlm1)]);
        }
    }
    for (int idx = w; idx < w * (h - 1); idx++) {
      float dx = 0.5f * (dI_l[3 * (idx + 1)] - dI_l[3 * (idx - 1)]);
      float dy = 0.5f * (dI_l[3 * (idx + w)] - dI_l[3 * (idx - w)]);
      if (!is_fin(dx)) dx = 0;
      if (!is_fin(dy)) dy = 0;
      dI_l[3 * idx + 1] = dx;
      dI_l[3 * idx + 2] = dy;
      dabs_l[idx] = dx * dx + dy * dy;
      if (B256 != nullptr) {
        int c = (int)(dI_l[3 * idx] + 0.5f);
        if (c < 5) c = 5;
        if (c > 250) c = 250;
        float gw = B256[c + 1] - B256[c];
        dabs_l[idx] *= gw * gw;
      }
    }
  }
}

// PixelSelector ctor: srand(3141592); rand() & 0xFF  (glibc rand()).  PixelSelector2.cpp:43-45
void oracle_random_pattern(int n, unsigned char* out) {
  std::srand(3141592);
  for (int i = 0; i < n; i++) out[i] = (unsigned char)(std::rand() & 0xFF);
}

static int compute_hist_quantil(const int* hist, float below) {  // PixelSelector2.cpp:66-75
  int th = (int)(hist[0] * below + 0.5f);
  for (int i = 0; i < 90; i++) {
    th -= hist[i + 1];
    if (th < 0) return i;
  }
  return 90;
}

struct OSelector {
  int w, h;
  std::vector<unsigned char> randomPattern;
  int currentPotential;
  std::vector<float> ths, thsSmoothed;
  int thsStep;
  // settings (src/util/settings.cpp:154-157)
  float minGradHistCut = 0.5f, minGradHistAdd = 7.f, gradDownweightPerLevel = 0.75f;
  int selectDirectionDistribution = 1;
};

void* oracle_selector_create(int w, int h) {
  OSelector* s = new OSelector();
  s->w = w; s->h = h;
  s->randomPattern.resize((size_t)w * h);
  oracle_random_pattern(w * h, s->randomPattern.data());
  s->currentPotential = 3;
  // reference allocates (w/32)*(h/32)+100 floats (uninitialised); defined here as zero-initialised and
  // large enough for every index select() can form: (w>>5) + (h>>5)*w32.
  size_t cap = (size_t)(w / 32) * (h / 32) + 100;
  size_t need = (size_t)((w - 1) >> 5) + (size_t)((h - 1) >> 5) * (w / 32) + 1;
  s->ths.assign(std::max(cap, need), 0.f);
  s->thsSmoothed.assign(std::max(cap, need), 0.f);
  s->thsStep = w / 32;
  return s;
}
void oracle_selector_destroy(void* p) { delete (OSelector*)p; }
int oracle_selector_get_potential(void* p) { return ((OSelector*)p)->currentPotential; }
void oracle_selector_set_potential(void* p, int v) { ((OSelector*)p)->currentPotential = v; }
void oracle_selector_set_settings(void* p, float cut, float add, float dw, int dirDist) {
  OSelector* s = (OSelector*)p;
  s->minGradHistCut = cut; s->minGradHistAdd = add; s->gradDownweightPerLevel = dw;
  s->selectDirectionDistribution = dirDist;
}
int oracle_selector_ths_size(void* p) { return (int)((OSelector*)p)->thsSmoothed.size(); }
void oracle_selector_get_ths(void* p, float* ths, float* thsSmoothed) {
  OSelector* s = (OSelector*)p;
  memcpy(ths, s->ths.data(), sizeof(float) * s->ths.size());
  memcpy(thsSmoothed, s->thsSmoothed.data(), sizeof(float) * s->thsSmoothed.size());
}
const unsigned char* oracle_selector_random_pattern(void* p) { return ((OSelector*)p)->randomPattern.data(); }

// a2
void oracle_selector_make_hists(void* p, const float* absgrad0) {
  OSelector* s = (OSelector*)p;
  const int w = s->w, h = s->h;
  const int w32 = w / 32, h32 = h / 32;
  s->thsStep = w32;
  int hist0[100];
  for (int y = 0; y < h32; y++)
    for (int x = 0; x < w32; x++) {
      const float* map0 = absgrad0 + 32 * x + 32 * y * w;
      memset(hist0, 0, sizeof(int) * 50);
      for (int j = 0; j < 32; j++)
        for (int i = 0; i < 32; i++) {
          int it = i + 32 * x;
          int jt = j + 32 * y;
          if (it > w - 2 || jt > h - 2 || it < 1 || jt < 1) continue;
          int g = (int)sqrtf(map0[i + j * w]);
          if (g > 48) g = 48;
          hist0[g + 1]++;
          hist0[0]++;
        }
      // hist0[50..90] stay zero in the reference too (gradHist is only memset for 50 ints per block, but
      // nothing ever writes beyond index 49): computeHistQuantil scanning up to 90 sees zeros.
      for (int k = 50; k < 100; k++) hist0[k] = 0;
      s->ths[x + y * w32] = compute_hist_quantil(hist0, s->minGradHistCut) + s->minGradHistAdd;
    }
  for (int y = 0; y < h32; y++)
    for (int x = 0; x < w32; x++) {
      float sum = 0, num = 0;
      if (x > 0) {
        if (y > 0) { num++; sum += s->ths[x - 1 + (y - 1) * w32]; }
        if (y < h32 - 1) { num++; sum += s->ths[x - 1 + (y + 1) * w32]; }
        num++; sum += s->ths[x - 1 + (y)*w32];
      }
      if (x < w32 - 1) {
        if (y > 0) { num++; sum += s->ths[x + 1 + (y - 1) * w32]; }
        if (y < h32 - 1) { num++; sum += s->ths[x + 1 + (y + 1) * w32]; }
        num++; sum += s->ths[x + 1 + (y)*w32];
      }
      if (y > 0) { num++; sum += s->ths[x + (y - 1) * w32]; }
      if (y < h32 - 1) { num++; sum += s->ths[x + (y + 1) * w32]; }
      num++; sum += s->ths[x + y * w32];
      s->thsSmoothed[x + y * w32] = (sum / num) * (sum / num);
    }
}

static const float kDirections[16][2] = {  // PixelSelector2.cpp:581-597 (double literals -> float)
    {(float)0, (float)1.0000},      {(float)0.3827, (float)0.9239},  {(float)0.1951, (float)0.9808},
    {(float)0.9239, (float)0.3827}, {(float)0.7071, (float)0.7071},  {(float)0.3827, (float)-0.9239},
    {(float)0.8315, (float)0.5556}, {(float)0.8315, (float)-0.5556}, {(float)0.5556, (float)-0.8315},
    {(float)0.9808, (float)0.1951}, {(float)0.9239, (float)-0.3827}, {(float)0.7071, (float)-0.7071},
    {(float)0.5556, (float)0.8315}, {(float)0.9808, (float)-0.1951}, {(float)1.0000, (float)0.0000},
    {(float)0.1951, (float)-0.9808}};

// a3. dI0: level-0 AoS {I,dx,dy}; ag0/ag1/ag2: absSquaredGrad of levels 0,1,2. n_out = (n2,n3,n4).
void oracle_selector_select(void* p, const float* dI0, const float* ag0p, const float* ag1p, const float* ag2p,
                            float* map_out, int pot, float thFactor, int* n_out) {
  OSelector* s = (OSelector*)p;
  const int w = s->w, h = s->h;
  const int w1 = w >> 1, w2 = w >> 2;
  const unsigned char* randomPattern = s->randomPattern.data();
  const float* thsSmoothed = s->thsSmoothed.data();
  const int thsStep = s->thsStep;
  memset(map_out, 0, sizeof(float) * (size_t)w * h);
  const float dw1 = s->gradDownweightPerLevel;
  const float dw2 = dw1 * dw1;
  int n3 = 0, n2 = 0, n4 = 0;
  for (int y4 = 0; y4 < h; y4 += (4 * pot))
    for (int x4 = 0; x4 < w; x4 += (4 * pot)) {
      int my3 = std::min((4 * pot), h - y4);
      int mx3 = std::min((4 * pot), w - x4);
      int bestIdx4 = -1;
      float bestVal4 = 0;
      const float* dir4 = kDirections[randomPattern[n2] & 0xF];
      for (int y3 = 0; y3 < my3; y3 += (2 * pot))
        for (int x3 = 0; x3 < mx3; x3 += (2 * pot)) {
          int x34 = x3 + x4;
          int y34 = y3 + y4;
          int my2 = std::min((2 * pot), h - y34);
          int mx2 = std::min((2 * pot), w - x34);
          int bestIdx3 = -1;
          float bestVal3 = 0;
          const float* dir3 = kDirections[randomPattern[n2] & 0xF];
          for (int y2 = 0; y2 < my2; y2 += pot)
            for (int x2 = 0; x2 < mx2; x2 += pot) {
              int x234 = x2 + x34;
              int y234 = y2 + y34;
              int my1 = std::min(pot, h - y234);
              int mx1 = std::min(pot, w - x234);
              int bestIdx2 = -1;
              float bestVal2 = 0;
              const float* dir2 = kDirections[randomPattern[n2] & 0xF];
              for (int y1 = 0; y1 < my1; y1 += 1)
                for (int x1 = 0; x1 < mx1; x1 += 1) {
                  int idx = x1 + x234 + w * (y1 + y234);
                  int xf = x1 + x234;
                  int yf = y1 + y234;
                  if (xf < 4 || xf >= w - 5 || yf < 4 || yf > h - 4) continue;
                  float pixelTH0 = thsSmoothed[(xf >> 5) + (yf >> 5) * thsStep];
                  float pixelTH1 = pixelTH0 * dw1;
                  float pixelTH2 = pixelTH1 * dw2;
                  float ag0 = ag0p[idx];
                  if (ag0 > pixelTH0 * thFactor) {
                    float gx = dI0[3 * idx + 1], gy = dI0[3 * idx + 2];
                    float dirNorm = fabsf(gx * dir2[0] + gy * dir2[1]);
                    if (!s->selectDirectionDistribution) dirNorm = ag0;
                    if (dirNorm > bestVal2) { bestVal2 = dirNorm; bestIdx2 = idx; bestIdx3 = -2; bestIdx4 = -2; }
                  }
                  if (bestIdx3 == -2) continue;
                  float ag1 = ag1p[(int)(xf * 0.5f + 0.25f) + (int)(yf * 0.5f + 0.25f) * w1];
                  if (ag1 > pixelTH1 * thFactor) {
                    float gx = dI0[3 * idx + 1], gy = dI0[3 * idx + 2];
                    float dirNorm = fabsf(gx * dir3[0] + gy * dir3[1]);
                    if (!s->selectDirectionDistribution) dirNorm = ag1;
                    if (dirNorm > bestVal3) { bestVal3 = dirNorm; bestIdx3 = idx; bestIdx4 = -2; }
                  }
                  if (bestIdx4 == -2) continue;
                  float ag2 = ag2p[(int)(xf * 0.25f + 0.125) + (int)(yf * 0.25f + 0.125) * w2];
                  if (ag2 > pixelTH2 * thFactor) {
                    float gx = dI0[3 * idx + 1], gy = dI0[3 * idx + 2];
                    float dirNorm = fabsf(gx * dir4[0] + gy * dir4[1]);
                    if (!s->selectDirectionDistribution) dirNorm = ag2;
                    if (dirNorm > bestVal4) { bestVal4 = dirNorm; bestIdx4 = idx; }
                  }
                }
              if (bestIdx2 > 0) { map_out[bestIdx2] = 1; bestVal3 = 1e10; n2++; }
            }
          if (bestIdx3 > 0) { map_out[bestIdx3] = 2; bestVal4 = 1e10; n3++; }
        }
      if (bestIdx4 > 0) { map_out[bestIdx4] = 4; n4++; }
    }
  n_out[0] = n2; n_out[1] = n3; n_out[2] = n4;
}

// a4. `hists_valid` plays the role of `fh == gradHistFrame` (PixelSelector2.cpp:184).
int oracle_selector_make_maps(void* p, const float* dI0, const float* ag0, const float* ag1, const float* ag2,
                              float* map_out, float density, int recursionsLeft, float thFactor, int hists_valid) {
  OSelector* s = (OSelector*)p;
  float numHave = 0;
  float numWant = density;
  float quotia;
  int idealPotential = s->currentPotential;
  {
    if (!hists_valid) oracle_selector_make_hists(p, ag0);
    int n[3];
    oracle_selector_select(p, dI0, ag0, ag1, ag2, map_out, s->currentPotential, thFactor, n);
    numHave = n[0] + n[1] + n[2];
    quotia = numWant / numHave;
    float K = numHave * (s->currentPotential + 1) * (s->currentPotential + 1);
    idealPotential = sqrtf(K / numWant) - 1;
    if (idealPotential < 1) idealPotential = 1;
    if (recursionsLeft > 0 && quotia > 1.25 && s->currentPotential > 1) {
      if (idealPotential >= s->currentPotential) idealPotential = s->currentPotential - 1;
      s->currentPotential = idealPotential;
      return oracle_selector_make_maps(p, dI0, ag0, ag1, ag2, map_out, density, recursionsLeft - 1, thFactor, 1);
    } else if (recursionsLeft > 0 && quotia < 0.25) {
      if (idealPotential <= s->currentPotential) idealPotential = s->currentPotential + 1;
      s->currentPotential = idealPotential;
      return oracle_selector_make_maps(p, dI0, ag0, ag1, ag2, map_out, density, recursionsLeft - 1, thFactor, 1);
    }
  }
  int numHaveSub = numHave;
  if (quotia < 0.95) {
    int wh = s->w * s->h;
    int rn = 0;
    unsigned char charTH = 255 * quotia;
    for (int i = 0; i < wh; i++) {
      if (map_out[i] != 0) {
        if (s->randomPattern[rn] > charTH) { map_out[i] = 0; numHaveSub--; }
        rn++;
      }
    }
  }
  s->currentPotential = idealPotential;
  return numHaveSub;
}

}  // extern "C"
