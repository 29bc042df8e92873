// data2d.h -- dense row-major host image with headerless raw I/O; same surface as the reference's
// Data2D (src/data_types/data2d.h:30-66).  Storage is page-locked when a CUDA device is present so
// that the H2D/D2H copies of ComputeFlow are true DMA transfers.
#pragma once
#include <cstddef>

class Data2D {
 public:
  Data2D() = default;
  Data2D(size_t width, size_t height);
  Data2D(const Data2D&) = delete;
  Data2D& operator=(const Data2D&) = delete;
  ~Data2D();

  size_t Width() const { return width_; }
  size_t Height() const { return height_; }
  float* DataPtr() { return data_; }
  const float* DataPtr() const { return data_; }
  float& Data(size_t x, size_t y) { return data_[y * width_ + x]; }

  void Swap(Data2D& other);
  void ZeroData();

  // Exact-size check like upstream (data2d.cpp:98-178): one more fread after the last row must hit EOF.
  bool ReadRAWFromFileU8(const char* filename, size_t width, size_t height);
  bool ReadRAWFromFileF32(const char* filename, size_t width, size_t height);
  bool WriteRAWToFileU8(const char* filename) const;
  bool WriteRAWToFileF32(const char* filename) const;

 private:
  bool Allocate(size_t width, size_t height);
  void Release();
  float* data_ = nullptr;
  size_t width_ = 0, height_ = 0;
  bool pinned_ = false;
};
