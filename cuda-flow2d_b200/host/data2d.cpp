#include "data2d.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "flow2d.h"

Data2D::Data2D(size_t width, size_t height) { Allocate(width, height); }
Data2D::~Data2D() { Release(); }

bool Data2D::Allocate(size_t width, size_t height) {
  Release();
  const size_t bytes = width * height * sizeof(float);
  if (bytes == 0) return false;
  data_ = static_cast<float*>(flow2d_host_alloc(bytes));
  pinned_ = data_ != nullptr;
  if (!data_) data_ = static_cast<float*>(std::malloc(bytes));
  if (!data_) {
    std::printf("Error. Cannot allocate memory on the host.\n");
    return false;
  }
  width_ = width;
  height_ = height;
  return true;
}

void Data2D::Release() {
  if (data_) {
    if (pinned_) flow2d_host_free(data_);
    else std::free(data_);
  }
  data_ = nullptr;
  width_ = height_ = 0;
  pinned_ = false;
}

void Data2D::Swap(Data2D& other) {
  if (width_ == other.width_ && height_ == other.height_) {
    std::swap(data_, other.data_);
    std::swap(pinned_, other.pinned_);
  } else {
    std::printf("Error. Cannot swap two Data2D objects (wrong dimensions).\n");
  }
}

void Data2D::ZeroData() {
  if (data_) std::memset(data_, 0, width_ * height_ * sizeof(float));
}

namespace {
struct File {
  std::FILE* f;
  File(const char* name, const char* mode) : f(std::fopen(name, mode)) {}
  ~File() { if (f) std::fclose(f); }
};
}  // namespace

bool Data2D::ReadRAWFromFileU8(const char* filename, size_t width, size_t height) {
  File file(filename, "rb");
  if (!file.f) {
    std::printf("Cannot open file '%s'.\n", filename);
    return false;
  }
  if (!Allocate(width, height)) return false;
  std::vector<unsigned char> row(width);
  bool ok = true;
  for (size_t y = 0; y < height && ok; ++y) {
    ok = std::fread(row.data(), 1, width, file.f) == width;
    if (ok)
      for (size_t x = 0; x < width; ++x) data_[y * width + x] = static_cast<float>(row[x]);
  }
  if (ok) ok = std::fread(row.data(), 1, width, file.f) == 0;  // the file must end exactly here
  if (!ok) {
    std::printf("Error reading RAW data from file '%s': wrong dimensions.", filename);
    Release();
  }
  return ok;
}

bool Data2D::ReadRAWFromFileF32(const char* filename, size_t width, size_t height) {
  File file(filename, "rb");
  if (!file.f) {
    std::printf("Cannot open file '%s'.\n", filename);
    return false;
  }
  if (!Allocate(width, height)) return false;
  bool ok = std::fread(data_, sizeof(float), width * height, file.f) == width * height;
  unsigned char extra;
  if (ok) ok = std::fread(&extra, 1, 1, file.f) == 0;
  if (!ok) {
    std::printf("Error reading RAW data from file '%s': wrong dimensions.", filename);
    Release();
  }
  return ok;
}

bool Data2D::WriteRAWToFileU8(const char* filename) const {
  File file(filename, "wb");
  if (!file.f) {
    std::printf("Cannot open file '%s'.\n", filename);
    return false;
  }
  std::vector<unsigned char> row(width_);
  for (size_t y = 0; y < height_; ++y) {
    for (size_t x = 0; x < width_; ++x)
      row[x] = static_cast<unsigned char>(std::min(255.f, std::max(0.f, data_[y * width_ + x])));
    if (std::fwrite(row.data(), 1, width_, file.f) != width_) {
      std::printf("Error writing RAW data to file '%s'.", filename);
      return false;
    }
  }
  return true;
}

bool Data2D::WriteRAWToFileF32(const char* filename) const {
  File file(filename, "wb");
  if (!file.f) {
    std::printf("Cannot open file '%s'.\n", filename);
    return false;
  }
  if (std::fwrite(data_, sizeof(float), width_ * height_, file.f) != width_ * height_) {
    std::printf("Error writing RAW data to file '%s'.", filename);
    return false;
  }
  return true;
}
