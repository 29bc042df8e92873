// settings.h -- reader of the reference's settings.xml (schema: settings.xml:1-27, parsed upstream
// by src/utils/settings.cpp:53-144 through TinyXML).  Only attribute lookup on a handful of
// elements is needed, so this is a small self-contained scanner, not an XML library.
//
// Accepted unchanged: Input/Path@inputPath, Input/Mode@Nx,@Ny,@imageType, Mode/Files@file1,@file2,
// Parameters/Method@key, Solver/Iterations@inner,@outer, Solver/Warping@levels,@scaling,
// @medianRadius, Solver/Model@sigma,@alpha,@e_smooth,@e_data, Output/Path@outputPath.
// Optional additions with reference-preserving defaults: Solver/Model@constancy = grey|gradient.
#pragma once
#include <string>

namespace OpticFlow {

class Settings {
 public:
  std::string inputPath, outputPath, fileName1, fileName2;
  std::string imageType = "32-bit";  // "8-bit" selects the U8 reader (upstream parses but ignores it)
  std::string constancy = "grey";
  int width = 0, height = 0;
  float sigma = 0.f;
  int medianRadius = 0;
  int iterInner = 0, iterOuter = 0;
  float alpha = 0.f, e_smooth = 0.f, e_data = 0.f;
  int levels = 0;
  float warpScale = 0.f;
  bool press_key = false;

  // 0 on success; -1 if the file cannot be read or a required element/attribute is missing
  // (upstream dereferences a null pointer in that case).
  int LoadSettings(const std::string& fileName);
  std::string error;
};

}  // namespace OpticFlow
