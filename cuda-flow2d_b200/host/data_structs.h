// data_structs.h -- value types of the solver surface, mirroring the reference's
// src/data_types/data_structs.h:27-35 (names and meaning kept so that caller code compiles unchanged).
#pragma once
#include <cstddef>

enum class DataConstancy { Grey, Gradient, LogDerivatives };

struct DataSize3 {
  size_t width;
  size_t height;
  size_t pitch;  // ignored on input, like upstream (Initialize resets it)
};
