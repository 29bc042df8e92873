// sequence.h -- flow between every pair of consecutive frames of an image sequence: the
// X-ray-radiography use case the reference README names (README.md:28) and SURVEY.md §8(f) ranks
// first after the hot path.  The reference CLI handles one pair per process (src/main.cpp:99-213);
// this driver keeps the GPU busy with several pairs at once (flow2d_compute_async on K handles)
// while one thread reads frames ahead and another writes finished flow fields.
//
// Frames come from (a) a list of files, one frame each, (b) a directory (every regular file in it,
// sorted by name) or (c) one headerless multi-frame stack.  8-bit and float32 frames are told apart
// by file size (the reference has a reader for each, src/data_types/data2d.cpp:98-141 and :143-178,
// but its CLI only ever calls the float32 one).
#pragma once
#include <cstddef>
#include <cstdio>
#include <string>
#include <vector>

#include "flow2d.h"

namespace FlowSequence {

enum class PixelType { Auto, U8, F32 };

class FrameSource {
 public:
  ~FrameSource();
  // false + error on: nothing readable, a file whose size matches neither pixel type, fewer than two frames
  bool Open(const std::vector<std::string>& inputs, size_t width, size_t height, PixelType type = PixelType::Auto);
  int Count() const { return count_; }
  size_t Width() const { return width_; }
  size_t Height() const { return height_; }
  PixelType Type() const { return type_; }     // U8 or F32 after Open
  const char* Kind() const { return kind_; }   // "files" | "directory" | "stack"
  std::string Name(int i) const;               // where frame i comes from (for messages)
  bool Read(int i, float* dst);                // whole frame, converted to float32
  std::string error;

 private:
  std::vector<std::string> files_;
  std::FILE* stack_ = nullptr;
  std::vector<unsigned char> bytes_;
  size_t width_ = 0, height_ = 0;
  int count_ = 0;
  PixelType type_ = PixelType::Auto;
  const char* kind_ = "files";
};

struct Options {
  int handles = 4;            // frame pairs in flight PER GPU
  int device = 0;
  std::vector<int> devices;   // several GPUs: pair i -> GPU devices[i mod N] (empty = {device}); no data moves between GPUs
  int io_threads = 0;         // reader threads and writer threads each; 0 = one per two GPUs
  int constancy = FLOW2D_GREY;
  bool write_color = false;   // NNNN_res.pgm  (src/utils/io_utils.cpp:140-225)
  bool write_amp = false;     // NNNN_amp-W-H.raw
  bool write_flow = true;     // NNNN_flow-u-W-H.raw, NNNN_flow-v-W-H.raw
  flow2d_params params;       // flow2d_default_params = src/main.cpp:70-80
};

struct Stats {
  int pairs = 0, handles = 0, gpus = 1, io_threads = 1;
  double seconds = 0;         // first read issued -> last file closed
  double read_seconds = 0, write_seconds = 0;  // busy time of the reader / writer threads (summed over the threads)
  double wait_frames_seconds = 0, wait_gpu_seconds = 0, wait_writer_seconds = 0;  // where the scheduler thread stalled
};

// 0 ok, 1 GPU-side failure (message on stderr), 2 a frame could not be read, 4 an output file could not be written
int Run(FrameSource& source, const std::string& out_prefix, const Options& opt, Stats* stats);

}  // namespace FlowSequence
