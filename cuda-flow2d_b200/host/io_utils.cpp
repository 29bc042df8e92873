#include "io_utils.h"

#include <cmath>
#include <cstdio>
#include <vector>

namespace IOUtils {

namespace {
// The colour wheel as a table: segment k covers half-angles [lo, hi) (in units of pi) and blends
// linearly from colour c0 to colour c1 (src/utils/io_utils.cpp:130-190 spells the same six segments
// out one by one).
struct Segment {
  double lo, hi;
  double c0[3], c1[3];
};
const Segment kWheel[6] = {
    {0.0, 0.125, {255, 0, 0}, {255, 0, 255}},      // red -> magenta
    {0.125, 0.25, {255, 0, 255}, {64, 64, 255}},   // magenta -> blue
    {0.25, 0.375, {64, 64, 255}, {0, 255, 255}},   // blue -> cyan
    {0.375, 0.5, {0, 255, 255}, {0, 255, 0}},      // cyan -> green
    {0.5, 0.75, {0, 255, 0}, {255, 255, 0}},       // green -> yellow
    {0.75, 1.0, {255, 255, 0}, {255, 0, 0}},       // yellow -> red (upper end inclusive)
};
int to_byte(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
}  // namespace

void FlowToRGB(float x, float y, int rgb[3]) {
  // fp32 for amplitude and angle, double for the blend: the same mixed precision as upstream, so
  // that the floor() lands on the same byte
  const float Pi = (float)(2.0 * std::acos(0.0));
  float amp = std::sqrt(x * x + y * y);
  if (amp > 1) amp = 1;
  float phi;
  if (x == 0.0f) phi = (y >= 0.0f) ? (float)(0.5 * Pi) : (float)(1.5 * Pi);
  else if (x > 0.0f) phi = (y >= 0.0f) ? std::atan(y / x) : (float)(2.0 * Pi + std::atan(y / x));
  else phi = Pi + std::atan(y / x);
  phi = (float)(phi / 2.0);
  rgb[0] = rgb[1] = rgb[2] = 0;
  for (int k = 0; k < 6; k++) {
    const Segment& s = kWheel[k];
    const bool in = (phi >= s.lo * Pi) && (k == 5 ? phi <= s.hi * Pi : phi < s.hi * Pi);
    if (!in) continue;
    const float beta = (float)((phi - s.lo * Pi) / ((s.hi - s.lo) * Pi));
    const float alpha = (float)(1.0 - beta);
    for (int c = 0; c < 3; c++) rgb[c] = (int)std::floor(amp * (alpha * s.c0[c] + beta * s.c1[c]));
  }
  for (int c = 0; c < 3; c++) rgb[c] = to_byte(rgb[c]);
}

bool WriteFlowToImageRGB(Data2D& u, Data2D& v, float flowMaxScale, const std::string& fileName) {
  std::FILE* f = std::fopen(fileName.c_str(), "wb");
  if (!f) {
    std::fprintf(stderr, "Error: cannot save file %s\n", fileName.c_str());
    return false;
  }
  const int nx = (int)u.Width(), ny = (int)u.Height();
  const float factor = (float)(1.0 / flowMaxScale);
  std::fprintf(f, "P6 \n%d %d \n255\n", nx, ny);
  std::vector<unsigned char> row(3 * (size_t)nx);
  for (int i = 0; i < ny; i++) {
    for (int j = 0; j < nx; j++) {
      int rgb[3];
      FlowToRGB(u.Data(j, i) * factor, v.Data(j, i) * factor, rgb);
      row[3 * j] = (unsigned char)rgb[0];
      row[3 * j + 1] = (unsigned char)rgb[1];
      row[3 * j + 2] = (unsigned char)rgb[2];
    }
    std::fwrite(row.data(), 1, row.size(), f);
  }
  std::fclose(f);
  return true;
}

bool WriteMagnitudeToFileF32(Data2D& u, Data2D& v, const std::string& fileName) {
  Data2D amp(u.Width(), u.Height());
  for (size_t y = 0; y < u.Height(); ++y)
    for (size_t x = 0; x < u.Width(); ++x) {
      const float a = u.Data(x, y), b = v.Data(x, y);
      amp.Data(x, y) = std::sqrt(a * a + b * b);
    }
  return amp.WriteRAWToFileF32(fileName.c_str());
}

}  // namespace IOUtils
