#include "sequence.h"

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <filesystem>
#include <memory>
#include <mutex>
#include <thread>

#include "data2d.h"
#include "io_utils.h"

namespace fs = std::filesystem;

namespace FlowSequence {

namespace {
double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

FrameSource::~FrameSource() {
  if (stack_) std::fclose(stack_);
}

std::string FrameSource::Name(int i) const {
  if (stack_) return files_[0] + "[" + std::to_string(i) + "]";
  return (i >= 0 && i < (int)files_.size()) ? files_[i] : std::string();
}

bool FrameSource::Open(const std::vector<std::string>& inputs, size_t width, size_t height, PixelType type) {
  width_ = width;
  height_ = height;
  files_.clear();
  count_ = 0;
  const size_t px = width * height;
  if (px == 0 || inputs.empty()) {
    error = "no frames";
    return false;
  }
  std::error_code ec;
  if (inputs.size() == 1 && fs::is_directory(inputs[0], ec)) {
    kind_ = "directory";
    for (const auto& e : fs::directory_iterator(inputs[0], ec))
      if (e.is_regular_file(ec)) files_.push_back(e.path().string());
    std::sort(files_.begin(), files_.end());
  } else {
    files_ = inputs;
  }
  if (files_.empty()) {
    error = "no frame files in '" + inputs[0] + "'";
    return false;
  }
  std::vector<size_t> sizes;
  for (const auto& f : files_) {
    const auto s = fs::file_size(f, ec);
    if (ec) {
      error = "Cannot open file '" + f + "'.";
      return false;
    }
    sizes.push_back((size_t)s);
  }
  if (files_.size() == 1) {
    // one file with more than one frame in it: a stack
    kind_ = "stack";
    const size_t s = sizes[0];
    if (type == PixelType::Auto) {
      // a stack of 4k 8-bit frames is as long as one of k float32 frames: float32 (the reference's
      // native type) wins the tie, PixelType::U8 overrides
      if (s % (4 * px) == 0) type = PixelType::F32;
      else if (s % px == 0) type = PixelType::U8;
    }
    const size_t frame_bytes = (type == PixelType::U8) ? px : 4 * px;
    if (type == PixelType::Auto || s % frame_bytes != 0 || s / frame_bytes < 2) {
      error = "'" + files_[0] + "': " + std::to_string(s) + " bytes is not a stack of two or more " + std::to_string(width) +
              "x" + std::to_string(height) + " frames";
      return false;
    }
    count_ = (int)(s / frame_bytes);
    stack_ = std::fopen(files_[0].c_str(), "rb");
    if (!stack_) {
      error = "Cannot open file '" + files_[0] + "'.";
      return false;
    }
  } else {
    for (size_t i = 0; i < files_.size(); i++) {
      PixelType t = sizes[i] == 4 * px ? PixelType::F32 : (sizes[i] == px ? PixelType::U8 : PixelType::Auto);
      if (type != PixelType::Auto && t != type) t = PixelType::Auto;
      if (t == PixelType::Auto || (i > 0 && t != type_)) {
        error = "Error reading RAW data from file '" + files_[i] + "': wrong dimensions.";
        return false;
      }
      type_ = t;
    }
    type = type_;
    count_ = (int)files_.size();
  }
  type_ = type;
  if (type_ == PixelType::U8) bytes_.resize(px);
  return true;
}

bool FrameSource::Read(int i, float* dst) {
  if (i < 0 || i >= count_) return false;
  const size_t px = width_ * height_;
  std::FILE* f = stack_;
  if (stack_) {
    const long long off = (long long)i * (long long)(type_ == PixelType::U8 ? px : 4 * px);
    if (fseeko(stack_, (off_t)off, SEEK_SET) != 0) return false;
  } else {
    f = std::fopen(files_[i].c_str(), "rb");
    if (!f) return false;
  }
  bool ok;
  if (type_ == PixelType::U8) {
    ok = std::fread(bytes_.data(), 1, px, f) == px;
    if (ok)
      for (size_t k = 0; k < px; k++) dst[k] = static_cast<float>(bytes_[k]);  // data2d.cpp:118-121
  } else {
    ok = std::fread(dst, sizeof(float), px, f) == px;
  }
  if (!stack_) std::fclose(f);
  if (!ok) error = "Error reading RAW data from '" + Name(i) + "'.";
  return ok;
}

// ---------------------------------------------------------------------------------------------
// Pipeline.  Reader threads, a scheduler, writer threads and two rings of page-locked buffers:
//   readers    frame j -> frame slot j % R          (may run R-T-1 frames ahead of the GPUs)
//   scheduler  pair i = (frame i, frame i+1) -> handle i % T, flow -> output slot i % M
//   writers    output slot -> files
// T = K handles on each of the N GPUs, handle t on GPU t % N, so pair i goes to GPU i mod N (independent units: nothing
// moves between GPUs).  A frame slot is recycled when both pairs that read it are complete, an output slot when its
// files are closed.  Everything is ordered by monotone counters (contiguous prefixes of per-item flags) under one mutex.
// ---------------------------------------------------------------------------------------------
int Run(FrameSource& source, const std::string& out_prefix, const Options& opt, Stats* stats) {
  const int n_frames = source.Count(), n_pairs = n_frames - 1;
  if (n_pairs < 1) return 2;
  const size_t width = source.Width(), height = source.Height();
  std::vector<int> devices = opt.devices;
  if (devices.empty()) devices.push_back(opt.device);
  const int N = (int)devices.size();
  const int T = std::max(1, std::min(std::max(1, opt.handles) * N, n_pairs));
  const int R = T + 3 + N, M = 2 * T;
  const int n_io = std::max(1, opt.io_threads > 0 ? opt.io_threads : (N + 1) / 2);

  // (throughput_mode stays as the caller set it: measured on B200 with 4-8 handles, the latency schedule of the
  // one-pixel kernels is also the better throughput schedule)
  flow2d_params p = opt.params;

  std::vector<flow2d_handle*> handles(T, nullptr);
  auto destroy_handles = [&]() {
    for (auto* h : handles)
      if (h) flow2d_destroy(h);
  };
  for (int k = 0; k < T; k++)
    if (flow2d_create(&handles[k], devices[k % N], (int)width, (int)height, opt.constancy) != FLOW2D_OK) {
      std::fprintf(stderr, "Error: cannot create a flow2d handle on device %d (no sm_100 GPU?)\n", devices[k % N]);
      destroy_handles();
      return 1;
    }

  // the first call of a handle captures its schedule into a CUDA graph (tens of milliseconds of host time): with K
  // handles on each of N GPUs that is done here, from one host thread per GPU, instead of T times in a row by the
  // scheduler while the GPUs wait
  {
    std::vector<std::thread> prep;
    for (int d = 0; d < N; d++)
      prep.emplace_back([&, d]() {
        for (int k = d; k < T; k += N) flow2d_prepare(handles[k], &p);
      });
    for (auto& t : prep) t.join();
  }

  std::vector<std::unique_ptr<Data2D>> frames, out_u, out_v;
  for (int i = 0; i < R; i++) frames.emplace_back(new Data2D(width, height));
  for (int i = 0; i < M; i++) {
    out_u.emplace_back(new Data2D(width, height));
    out_v.emplace_back(new Data2D(width, height));
  }

  std::mutex mu;
  std::condition_variable cv;
  std::vector<char> frame_ready(n_frames, 0), pair_written(n_pairs, 0);
  int loaded = 0;     // frames 0..loaded-1 are in their slots
  int completed = 0;  // pairs 0..completed-1 are off the GPU
  int written = 0;    // pairs 0..written-1 are on disk
  int failure = 0;    // first error code; stops all threads
  std::deque<int> to_write;
  bool no_more_output = false;
  Stats st;
  st.pairs = n_pairs;
  st.handles = T;
  st.gpus = N;
  st.io_threads = n_io;
  const double t_begin = now();

  // FrameSource::Read of a stack seeks in one FILE*: serialise the readers on it (files and directories read in parallel)
  std::mutex stack_mu;
  const bool serial_source = std::string(source.Kind()) == "stack";
  std::vector<std::thread> readers, writers;
  for (int rt = 0; rt < n_io; rt++)
    readers.emplace_back([&, rt]() {
      for (int j = rt; j < n_frames; j += n_io) {
        {
          std::unique_lock<std::mutex> lk(mu);
          // slot j % R held frame j-R, last read by pair j-R
          cv.wait(lk, [&] { return failure || completed >= j - R + 1; });
          if (failure) return;
        }
        const double t0 = now();
        bool ok;
        if (serial_source) {
          std::lock_guard<std::mutex> g(stack_mu);
          ok = source.Read(j, frames[j % R]->DataPtr());
        } else {
          ok = source.Read(j, frames[j % R]->DataPtr());
        }
        const double dt = now() - t0;
        std::lock_guard<std::mutex> lk(mu);
        st.read_seconds += dt;
        if (!ok) {
          std::fprintf(stderr, "%s\n", source.error.c_str());
          if (!failure) failure = 2;
        } else {
          frame_ready[j] = 1;
          while (loaded < n_frames && frame_ready[loaded]) ++loaded;
        }
        cv.notify_all();
        if (!ok) return;
      }
    });

  const std::string suffix = "-" + std::to_string(width) + "-" + std::to_string(height) + ".raw";
  for (int wt = 0; wt < n_io; wt++)
    writers.emplace_back([&]() {
      for (;;) {
        int pair;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return failure || !to_write.empty() || no_more_output; });
          if (failure || to_write.empty()) return;
          pair = to_write.front();
          to_write.pop_front();
        }
        const double t0 = now();
        char tag[16];
        std::snprintf(tag, sizeof tag, "%04d_", pair);
        Data2D &u = *out_u[pair % M], &v = *out_v[pair % M];
        bool ok = true;
        if (opt.write_flow)
          ok = u.WriteRAWToFileF32((out_prefix + tag + "flow-u" + suffix).c_str()) &&
               v.WriteRAWToFileF32((out_prefix + tag + "flow-v" + suffix).c_str());
        if (ok && opt.write_color) ok = IOUtils::WriteFlowToImageRGB(u, v, 10, out_prefix + tag + "res.pgm");  // src/main.cpp:212
        if (ok && opt.write_amp) ok = IOUtils::WriteMagnitudeToFileF32(u, v, out_prefix + tag + "amp" + suffix);
        const double dt = now() - t0;
        std::lock_guard<std::mutex> lk(mu);
        st.write_seconds += dt;
        if (!ok && !failure) failure = 4;
        pair_written[pair] = 1;
        while (written < n_pairs && pair_written[written]) ++written;
        cv.notify_all();
      }
    });

  // scheduler (this thread)
  auto retire = [&](int pair) {  // pair is the oldest one still on the GPUs
    const double t0 = now();
    const int rc = flow2d_synchronize(handles[pair % T]);
    st.wait_gpu_seconds += now() - t0;
    std::lock_guard<std::mutex> lk(mu);
    if (rc != FLOW2D_OK) {
      std::fprintf(stderr, "Error: %s\n", flow2d_last_error(handles[pair % T]));
      if (!failure) failure = 1;
    } else {
      completed = pair + 1;
      to_write.push_back(pair);
    }
    cv.notify_all();
  };
  int issued = 0;
  for (int i = 0; i < n_pairs; i++) {
    if (i >= T) retire(i - T);
    {
      std::unique_lock<std::mutex> lk(mu);
      double t0 = now();
      cv.wait(lk, [&] { return failure || loaded >= i + 2; });
      st.wait_frames_seconds += now() - t0;
      t0 = now();
      cv.wait(lk, [&] { return failure || written >= i - M + 1; });
      st.wait_writer_seconds += now() - t0;
      if (failure) break;
    }
    const int k = i % T;
    if (flow2d_compute_async(handles[k], frames[i % R]->DataPtr(), frames[(i + 1) % R]->DataPtr(), out_u[i % M]->DataPtr(),
                             out_v[i % M]->DataPtr(), &p) != FLOW2D_OK) {
      std::fprintf(stderr, "Error: %s\n", flow2d_last_error(handles[k]));
      std::lock_guard<std::mutex> lk(mu);
      if (!failure) failure = 1;
      cv.notify_all();
      break;
    }
    issued = i + 1;
  }
  for (int i = std::max(0, issued - T); i < issued; i++) {
    bool stop;
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = failure != 0;
    }
    if (stop) {  // still drain the GPUs before the buffers go away
      flow2d_synchronize(handles[i % T]);
      continue;
    }
    if (i >= completed) retire(i);
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    no_more_output = true;
    cv.notify_all();
  }
  for (auto& t : readers) t.join();
  for (auto& t : writers) t.join();
  st.seconds = now() - t_begin;
  destroy_handles();
  if (stats) *stats = st;
  return failure;
}

}  // namespace FlowSequence
