// io_utils.h -- the two post-processing writers of the reference CLI (src/utils/io_utils.{h,cpp}):
// the colour-coded flow image (binary PPM written under the name res.pgm, src/main.cpp:212) and the
// flow magnitude as raw float32 (amp-W-H.raw).  CPU code, off the timed path.
#pragma once
#include <string>

#include "data2d.h"

namespace IOUtils {

// Colour code of one flow vector scaled to the unit disc (after Bruhn): hue = direction on a
// red - blue - green - yellow - red wheel, brightness = min(|flow|, 1).  Returns r, g, b in 0..255.
void FlowToRGB(float x, float y, int rgb[3]);
// P6 PPM, header "P6 \n<w> <h> \n255\n", one RGB byte triple per pixel; flow is divided by flowMaxScale.
bool WriteFlowToImageRGB(Data2D& u, Data2D& v, float flowMaxScale, const std::string& fileName);
bool WriteMagnitudeToFileF32(Data2D& u, Data2D& v, const std::string& fileName);

}  // namespace IOUtils
