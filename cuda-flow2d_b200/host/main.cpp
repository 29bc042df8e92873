// main.cpp -- the cuda-flow2d command line, same three argv forms, output files and exit codes as
// the reference (src/main.cpp:46-229):
//   cuda-flow2d                      settings.xml in the current directory
//   cuda-flow2d <settings file>
//   cuda-flow2d <file1> <file2> <width> <height> [<counter>] <output path> [<alpha> <sigma>]
// Outputs: <out><counter>flow-u-W-H.raw, flow-v-W-H.raw (float32), res.pgm (colour-coded flow, binary PPM),
// amp-W-H.raw (float32 magnitude).
// Exit codes: 0 ok / usage, 1 no CUDA device, 2 input files unreadable, 3 settings unreadable.
//   cuda-flow2d --sequence <width> <height> <output path> <frames...> | <directory> | <stack file> [options]   (new)
//       flow between every pair of consecutive frames of a sequence, several pairs in flight at once
//       (host/sequence.h); 8-bit / float32 frames told apart by file size; writes
//       <out>NNNN_flow-u-W-H.raw / NNNN_flow-v-W-H.raw (and NNNN_res.pgm / NNNN_amp-W-H.raw on request)
//   cuda-flow2d --slab <devices> <file1> <file2> <width> <height> <output path> [--gradient] [--settings file]   (new)
//       one large frame pair on several GPUs (rows of the large levels split across them); bit-identical output
//   any pair form + trailing --residuals                                                                   (new)
//       prints the RMS residual of every level's last linear system (flow2d_level_residuals)
// Differences: no getchar() at exit; 8-bit and float32 input files are told apart by their size (Mode@imageType only breaks ties);
// files are also looked up under Input/Path@inputPath when they are not found in the CWD.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "flow2d.h"
#include "io_utils.h"
#include "optical_flow_2d.h"
#include "sequence.h"
#include "settings.h"

using std::string;

// "0-7", "0,2,5", "0-3,6" -> device indices
static std::vector<int> parse_devices(const string& spec) {
  std::vector<int> out;
  size_t pos = 0;
  while (pos < spec.size()) {
    size_t end = spec.find(',', pos);
    if (end == string::npos) end = spec.size();
    const string part = spec.substr(pos, end - pos);
    const size_t dash = part.find('-');
    const int lo = std::atoi(part.substr(0, dash).c_str());
    const int hi = dash == string::npos ? lo : std::atoi(part.substr(dash + 1).c_str());
    for (int d = lo; d <= hi && d - lo < 64; d++) out.push_back(d);
    pos = end + 1;
  }
  return out;
}

static long long file_size(const string& p);

// Slab mode: ONE large frame pair on several GPUs, rows of the large pyramid levels split across them
// (flow2d_slab_group_*, include/flow2d.h; BASELINE.json configs[4]).  Same inputs, defaults and output files as the
// <file1> <file2> <width> <height> <output path> form; the result is bit-identical to one GPU.
static int run_slab(int argc, char** argv) {
  if (argc < 8) {
    std::cout << "Usage: " << argv[0] << " --slab <devices, e.g. 0-7> <file1> <file2> <width> <height> <output path> [--gradient] [--settings file]"
              << std::endl;
    return 0;
  }
  const std::vector<int> devices = parse_devices(argv[2]);
  const string file1 = argv[3], file2 = argv[4], out = argv[7];
  const size_t width = std::atoi(argv[5]), height = std::atoi(argv[6]);
  flow2d_params p;
  flow2d_default_params(&p);  // src/main.cpp:70-80
  int constancy = FLOW2D_GREY;
  for (int i = 8; i < argc; i++) {
    const string a = argv[i];
    if (a == "--gradient") constancy = FLOW2D_GRADIENT;
    else if (a == "--settings" && i + 1 < argc) {
      OpticFlow::Settings settings;
      if (settings.LoadSettings(argv[++i]) != 0) {
        std::cout << settings.error << std::endl;
        return 3;
      }
      p.warp_levels_count = settings.levels; p.warp_scale_factor = settings.warpScale;
      p.outer_iterations_count = settings.iterOuter; p.inner_iterations_count = settings.iterInner;
      p.equation_alpha = settings.alpha; p.equation_data = settings.e_data; p.equation_smoothness = settings.e_smooth;
      p.median_radius = settings.medianRadius; p.gaussian_sigma = settings.sigma;
      if (settings.constancy == "gradient") constancy = FLOW2D_GRADIENT;
    }
  }
  if (devices.empty() || width < 4 || height < 4) return 0;
  Data2D frame_0, frame_1;
  auto read_frame = [&](Data2D& d, const string& name) {
    const bool u8 = file_size(name) == (long long)width * (long long)height;
    return u8 ? d.ReadRAWFromFileU8(name.c_str(), width, height) : d.ReadRAWFromFileF32(name.c_str(), width, height);
  };
  if (!read_frame(frame_0, file1) || !read_frame(frame_1, file2)) return 2;
  flow2d_slab_group* g = nullptr;
  if (flow2d_slab_group_create(&g, devices.data(), (int)devices.size(), width, height, constancy) != FLOW2D_OK) {
    std::cerr << "Error: cannot set up " << devices.size() << " GPUs for slab mode (sm_100 devices with peer access are needed)" << std::endl;
    return 1;
  }
  Data2D flow_u(width, height), flow_v(width, height);
  flow_u.ZeroData();
  flow_v.ZeroData();
  float ms = 0.f;
  const int rc = flow2d_slab_group_compute(g, frame_0.DataPtr(), frame_1.DataPtr(), flow_u.DataPtr(), flow_v.DataPtr(), &p, &ms);
  if (rc != FLOW2D_OK) {
    std::cerr << "TERMINATING. " << flow2d_slab_group_last_error(g) << " (" << rc << ")" << std::endl;
    flow2d_slab_group_destroy(g);
    return 4;
  }
  flow2d_slab_group_destroy(g);
  std::printf("Slab mode: %zu GPUs\nTotal GPU computation time: % 4.4fs\n", devices.size(), ms / 1000.);
  const string suffix = "-" + std::to_string(width) + "-" + std::to_string(height) + ".raw";
  bool written = flow_u.WriteRAWToFileF32((out + "flow-u" + suffix).c_str());
  written = flow_v.WriteRAWToFileF32((out + "flow-v" + suffix).c_str()) && written;
  IOUtils::WriteFlowToImageRGB(flow_u, flow_v, 10, out + "res.pgm");
  IOUtils::WriteMagnitudeToFileF32(flow_u, flow_v, out + "amp" + suffix);
  return written ? 0 : 5;
}

// Sequence mode: the X-ray-radiography use case of the reference README (one flow per consecutive
// frame pair); host/sequence.{h,cpp} does the work.
static int run_sequence(int argc, char** argv) {
  FlowSequence::Options opt;
  flow2d_default_params(&opt.params);  // src/main.cpp:70-80
  FlowSequence::PixelType type = FlowSequence::PixelType::Auto;
  bool list_only = false;
  std::vector<string> pos;
  for (int i = 2; i < argc; i++) {
    const string a = argv[i];
    if (a == "--handles" && i + 1 < argc) opt.handles = std::atoi(argv[++i]);
    else if (a == "--device" && i + 1 < argc) opt.device = std::atoi(argv[++i]);
    else if (a == "--devices" && i + 1 < argc) {
      opt.devices = parse_devices(argv[++i]);
    } else if (a == "--io-threads" && i + 1 < argc) opt.io_threads = std::atoi(argv[++i]);
    else if (a == "--u8") type = FlowSequence::PixelType::U8;
    else if (a == "--f32") type = FlowSequence::PixelType::F32;
    else if (a == "--color") opt.write_color = true;
    else if (a == "--amp") opt.write_amp = true;
    else if (a == "--no-flow") opt.write_flow = false;
    else if (a == "--gradient") opt.constancy = FLOW2D_GRADIENT;
    else if (a == "--list") list_only = true;
    else if (a == "--settings" && i + 1 < argc) {
      OpticFlow::Settings settings;  // solver values only; sizes and files come from the command line
      if (settings.LoadSettings(argv[++i]) != 0) {
        std::cout << settings.error << std::endl;
        return 3;
      }
      opt.params.warp_levels_count = settings.levels;
      opt.params.warp_scale_factor = settings.warpScale;
      opt.params.outer_iterations_count = settings.iterOuter;
      opt.params.inner_iterations_count = settings.iterInner;
      opt.params.equation_alpha = settings.alpha;
      opt.params.equation_data = settings.e_data;
      opt.params.equation_smoothness = settings.e_smooth;
      opt.params.median_radius = settings.medianRadius;
      opt.params.gaussian_sigma = settings.sigma;
      if (settings.constancy == "gradient") opt.constancy = FLOW2D_GRADIENT;
    } else if (a.rfind("--", 0) == 0) {
      std::cout << "Unknown option " << a << std::endl;
      return 0;
    } else {
      pos.push_back(a);
    }
  }
  if (pos.size() < 4) {
    std::cout << "Usage: " << argv[0] << " --sequence <width> <height> <output path> <frame 0> <frame 1> [...] | <directory> | <stack file>\n"
              << "       [--handles K] [--u8|--f32] [--settings file] [--gradient] [--color] [--amp] [--no-flow] [--device D | --devices 0-7]\n"
              << "       [--io-threads T] [--list]     (K pairs in flight per GPU; pair i runs on GPU i mod N)"
              << std::endl;
    return 0;
  }
  const size_t width = std::atoi(pos[0].c_str()), height = std::atoi(pos[1].c_str());
  const string out = pos[2];
  FlowSequence::FrameSource source;
  if (!source.Open(std::vector<string>(pos.begin() + 3, pos.end()), width, height, type) || source.Count() < 2) {
    std::cout << (source.error.empty() ? string("a sequence needs two or more frames") : source.error) << std::endl;
    return 2;
  }
  std::printf("Sequence: %d frames of %zux%zu (%s, %s)\n", source.Count(), width, height, source.Kind(),
              source.Type() == FlowSequence::PixelType::U8 ? "8-bit" : "float32");
  if (list_only) {
    for (int i = 0; i < source.Count(); i++) std::printf("  %04d %s\n", i, source.Name(i).c_str());
    return 0;
  }
  FlowSequence::Stats st;
  const int rc = FlowSequence::Run(source, out, opt, &st);
  if (rc == 0) {
    std::printf("Sequence: %d frame pairs on %d concurrent handles in %.3f s: %.2f pairs/s, %.2f Mpix/s end to end\n", st.pairs,
                st.handles, st.seconds, st.pairs / st.seconds, st.pairs * (double)width * height / st.seconds * 1e-6);
    std::printf("Sequence: %d GPU(s), %d reader and %d writer thread(s)\n", st.gpus, st.io_threads, st.io_threads);
    std::printf("Sequence: reader busy %.3f s, writer busy %.3f s; scheduler waited %.3f s for frames, %.3f s for the GPU, %.3f s for the writer\n",
                st.read_seconds, st.write_seconds, st.wait_frames_seconds, st.wait_gpu_seconds, st.wait_writer_seconds);
  }
  return rc;
}

static bool exists(const string& p) {
  std::FILE* f = std::fopen(p.c_str(), "rb");
  if (f) std::fclose(f);
  return f != nullptr;
}

static long long file_size(const string& p) {  // -1 if the file cannot be opened
  std::FILE* f = std::fopen(p.c_str(), "rb");
  if (!f) return -1;
  long long n = -1;
  if (std::fseek(f, 0, SEEK_END) == 0) n = std::ftell(f);
  std::fclose(f);
  return n;
}

int main(int argc, char** argv) {
  std::printf("//----------------------------------------------------------------------//\n");
  std::printf("//   2D optical flow, Blackwell-native (%s)   //\n", flow2d_version());
  std::printf("//----------------------------------------------------------------------//\n");

  if (argc >= 2 && string(argv[1]) == "--sequence") return run_sequence(argc, argv);
  if (argc >= 2 && string(argv[1]) == "--slab") return run_slab(argc, argv);
  // Optional flags of the pair forms (none of them exists upstream; without them the program behaves like the
  // reference's): --residuals prints per-level residual norms; the rest switch on the opt-in solver extensions of
  // include/flow2d.h (flow2d_params.scheme ... cascaded_restriction).  Flags may stand anywhere and are removed from argv.
  bool report_residuals = false, cascaded = false, level_times = false;
  int scheme = 0, data_term = 0, check_every = 0;
  float omega = 0.f, gamma = 0.f, tolerance = 0.f;
  {
    int kept = 1;
    for (int i = 1; i < argc; i++) {
      const string a = argv[i];
      auto val = [&](const char* key) -> const char* {
        const size_t n = std::strlen(key);
        return a.compare(0, n, key) == 0 ? a.c_str() + n : nullptr;
      };
      const char* v;
      if (a == "--residuals") report_residuals = true;
      else if (a == "--cascade") cascaded = true;
      else if (a == "--level-times") level_times = true;
      else if ((v = val("--scheme="))) scheme = (string(v) == "rb" || string(v) == "red-black") ? 1 : 0;
      else if ((v = val("--omega="))) omega = (float)std::atof(v);
      else if ((v = val("--term="))) data_term = string(v) == "gradient" ? 1 : string(v) == "log" ? 2 : string(v) == "combined" ? 3 : 0;
      else if ((v = val("--gamma="))) gamma = (float)std::atof(v);
      else if ((v = val("--tol="))) tolerance = (float)std::atof(v);
      else if ((v = val("--check-every="))) check_every = std::atoi(v);
      else argv[kept++] = argv[i];
    }
    argc = kept;
  }

  size_t width = 584, height = 388;
  size_t warp_levels_count = 50;  // src/main.cpp:70-80
  float warp_scale_factor = 0.9f;
  size_t outer_iterations_count = 40;
  size_t inner_iterations_count = 5;
  float equation_alpha = 35.0f;
  float equation_smoothness = 0.001f;
  float equation_data = 0.001f;
  size_t median_radius = 5;
  float gaussian_sigma = 1.5f;
  DataConstancy data_constancy = DataConstancy::Grey;
  string file_name1 = "rub1.raw", file_name2 = "rub2.raw";
  string input_path = "./data/", output_path = "./data/output/", counter = "";
  bool eight_bit = false;

  if (argc == 6 || argc == 7 || argc == 9) {
    file_name1 = argv[1];
    file_name2 = argv[2];
    width = std::atoi(argv[3]);
    height = std::atoi(argv[4]);
    // upstream reads argv[6] for every form (NULL for argc == 6); the 6-argument form means
    // <file1> <file2> <W> <H> <output path> here
    output_path = string(argc == 6 ? argv[5] : argv[6]);
    if (argc == 7) counter = argv[5];
    if (argc == 9) {
      equation_alpha = std::atof(argv[7]);
      gaussian_sigma = std::atof(argv[8]);
      counter = "alpha" + string(argv[7]) + "_sigma" + string(argv[8]) + "_";
    }
  } else if (argc < 3) {
    const string settingsFile = (argc == 1) ? "settings.xml" : string(argv[1]);
    std::cout << "Reading settings: " << settingsFile << std::endl;
    OpticFlow::Settings settings;
    if (settings.LoadSettings(settingsFile) != 0) {
      std::cout << settings.error << std::endl;
      std::cout << "TERMINATING. Error reading settings: " << settingsFile << std::endl;
      return 3;
    }
    std::cout << "OK" << std::endl << std::endl;
    width = settings.width;
    height = settings.height;
    input_path = settings.inputPath;
    output_path = settings.outputPath;
    file_name1 = settings.fileName1;
    file_name2 = settings.fileName2;
    warp_levels_count = settings.levels;
    warp_scale_factor = settings.warpScale;
    outer_iterations_count = settings.iterOuter;
    inner_iterations_count = settings.iterInner;
    equation_alpha = settings.alpha;
    equation_data = settings.e_data;
    equation_smoothness = settings.e_smooth;
    median_radius = settings.medianRadius;
    gaussian_sigma = settings.sigma;
    eight_bit = settings.imageType == "8-bit";
    if (settings.constancy == "gradient") data_constancy = DataConstancy::Gradient;
    if (!exists(file_name1) && exists(input_path + file_name1)) file_name1 = input_path + file_name1;
    if (!exists(file_name2) && exists(input_path + file_name2)) file_name2 = input_path + file_name2;
  } else {
    std::cout << "Usage: " << argv[0] << " <settings file>. Otherwise settings.xml in the current directory is used\n"
              << "       " << argv[0] << " <file1> <file2> <width> <height> [<counter>] <output path> [<alpha> <sigma>]" << std::endl;
    return 0;
  }

  OpticalFlow2D optical_flow;
  DataSize3 data_size = {width, height, 1};
  if (!optical_flow.Initialize(data_size, data_constancy)) return 1;

  // The reference always calls ReadRAWFromFileF32 and ignores Mode@imageType (src/main.cpp:175-183); its stock
  // settings.xml nevertheless says imageType="8-bit".  So the reader follows the FILE: W*H bytes = 8-bit, 4*W*H bytes =
  // float32 (the same rule as the sequence driver); imageType only breaks the tie of a size that is neither.
  Data2D frame_0, frame_1;
  auto read_frame = [&](Data2D& d, const string& name) {
    const long long bytes = file_size(name);
    const long long px = (long long)width * (long long)height;
    const bool u8 = bytes == px ? true : bytes == 4 * px ? false : eight_bit;
    return u8 ? d.ReadRAWFromFileU8(name.c_str(), width, height) : d.ReadRAWFromFileF32(name.c_str(), width, height);
  };
  if (!read_frame(frame_0, file_name1) || !read_frame(frame_1, file_name2)) return 2;

  Data2D flow_u(width, height), flow_v(width, height);
  optical_flow.silent = true;

  OperationParameters params;
  params.PushValuePtr("warp_levels_count", &warp_levels_count);
  params.PushValuePtr("warp_scale_factor", &warp_scale_factor);
  params.PushValuePtr("outer_iterations_count", &outer_iterations_count);
  params.PushValuePtr("inner_iterations_count", &inner_iterations_count);
  params.PushValuePtr("equation_alpha", &equation_alpha);
  params.PushValuePtr("equation_smoothness", &equation_smoothness);
  params.PushValuePtr("equation_data", &equation_data);
  params.PushValuePtr("median_radius", &median_radius);
  params.PushValuePtr("gaussian_sigma", &gaussian_sigma);
  if (report_residuals) params.PushValuePtr("report_residuals", &report_residuals);
  if (scheme) params.PushValuePtr("solver_scheme", &scheme);
  if (omega != 0.f) params.PushValuePtr("solver_omega", &omega);
  if (data_term) params.PushValuePtr("data_term", &data_term);
  if (gamma != 0.f) params.PushValuePtr("data_gamma", &gamma);
  if (tolerance > 0.f) params.PushValuePtr("residual_tolerance", &tolerance);
  if (check_every > 0) params.PushValuePtr("residual_check_every", &check_every);
  if (cascaded) params.PushValuePtr("cascaded_restriction", &cascaded);
  if (level_times) params.PushValuePtr("report_level_times", &level_times);

  flow_u.ZeroData();
  flow_v.ZeroData();
  optical_flow.ComputeFlow(frame_0, frame_1, flow_u, flow_v, params);
  if (optical_flow.last_status() != 0) {
    // upstream's ComputeFlow is void and main() writes whatever the buffers hold; a failed solve must not
    // leave plausible-looking files behind
    std::cerr << "TERMINATING. Optical flow computation failed (status " << optical_flow.last_status() << ")." << std::endl;
    return 4;
  }

  const string suffix = "-" + std::to_string(width) + "-" + std::to_string(height) + ".raw";
  bool written = flow_u.WriteRAWToFileF32((output_path + counter + "flow-u" + suffix).c_str());
  written = flow_v.WriteRAWToFileF32((output_path + counter + "flow-v" + suffix).c_str()) && written;
  IOUtils::WriteFlowToImageRGB(flow_u, flow_v, 10, output_path + counter + "res.pgm");  // src/main.cpp:212
  IOUtils::WriteMagnitudeToFileF32(flow_u, flow_v, output_path + counter + "amp" + suffix);
  optical_flow.Destroy();
  if (!written) {
    std::cerr << "TERMINATING. Cannot write the flow files to " << output_path << std::endl;
    return 5;
  }
  return 0;
}
