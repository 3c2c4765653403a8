// mosaic-b200 — a small C++ host over the C ABI with the reference CLI's flag names (struct Args, main.rs:20-88) and
// main()'s sequence (main.rs:741-779): target aspect -> MosaicGenerator::new -> generate_mosaic -> save.
//
// Images are PNG (include/mosaic_png.hpp: every colour type / bit depth, as `image::open(path)?.to_rgb8()` reads them), JPEG
// (include/mosaic_jpeg.hpp: baseline + progressive, libjpeg-turbo's default arithmetic), WebP (include/mosaic_webp.hpp: VP8 lossy
// with libwebp's default conversion to RGB, VP8L lossless) or binary PPM (P6, maxval 255), told apart by the extension,
// case-insensitively like main.rs:175-184 — the four extensions main.rs:181 lists; --material-src is a directory of such files
// of any sizes; the output is PNG, JPEG, WebP (lossless) or PPM by the extension of --output (main.rs:668-672).
// The material selection follows load_tiles (main.rs:166-261): aspect
// filter (lib.rs:121-123), no-match fallback to the first min(len, 2*max) files, truncate to --max-materials; files are
// taken in sorted order (read_dir order is unspecified in the reference).
//
//   g++ -std=c++17 -O2 -I../../include mosaic_cli.cpp -L.. -lmosaic_b200 -Wl,-rpath,'$ORIGIN/..' -o mosaic-b200
#include <dirent.h>
#include <sys/stat.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <thread>

#include "mosaic_b200.hpp"
#include "mosaic_jpeg.hpp"
#include "mosaic_png.hpp"
#include "mosaic_webp.hpp"

using namespace mosaic_b200;

struct Args {  // main.rs:23-88, same defaults
  std::string target, material_src, output;
  unsigned grid_w = 50, grid_h = 28;
  size_t max_materials = 500;
  float aspect_tolerance = 0.1f;
  size_t max_usage_per_image = 3;
  float adjacency_penalty_weight = 0.3f;
  bool enable_optimization = true;
  size_t optimization_iterations = 1000;
  std::string similarity_db = "similarity_db.json";
  bool rebuild_similarity_db = false;
  float color_adjustment_strength = 0.3f;
  bool show_time = true, show_grid = true;
  unsigned long long seed = 0;  // extension: the reference's annealer is OS-seeded (optimizer.rs:48)
  bool list_tiles = false;      // extension: print what load_tiles selects and stop (no device needed)
};

static bool parse_bool(const std::string& v) { return v == "true" || v == "1"; }

// clap's value parsers (main.rs:23-88): `usize` / `u32` take decimal digits only, `f32` what str::parse::<f32> takes; anything
// else is a usage error (exit code 2), not an exception
struct BadValue : std::runtime_error {
  using std::runtime_error::runtime_error;
};
static unsigned long long parse_unsigned(const std::string& flag, const std::string& v, unsigned long long max) {
  unsigned long long x = 0;
  bool ok = !v.empty() && v.size() <= 20;
  for (size_t i = 0; ok && i < v.size(); ++i) {
    const char c = v[i];
    if (c == '+' && i == 0 && v.size() > 1) continue;  // (Rust's integer parser accepts a leading '+')
    if (c < '0' || c > '9' || x > (max - (unsigned)(c - '0')) / 10) ok = false;
    else x = x * 10 + (unsigned)(c - '0');
  }
  if (!ok) throw BadValue("error: invalid value '" + v + "' for '" + flag + "': not an unsigned integer in range");
  return x;
}
static float parse_float(const std::string& flag, const std::string& v) {
  char* end = nullptr;
  const float x = v.empty() || std::isspace((unsigned char)v[0]) ? 0.0f : std::strtof(v.c_str(), &end);
  if (!end || end == v.c_str() || *end != 0 || (v.size() > 1 && (v[1] == 'x' || v[1] == 'X')))  // (no hex floats in Rust)
    throw BadValue("error: invalid value '" + v + "' for '" + flag + "': invalid float literal");
  return x;
}

static bool parse(int argc, char** argv, Args* a) {
  for (int i = 1; i < argc; ++i) {
    std::string k = argv[i], v;
    const size_t eq = k.find('=');
    const bool flag_only = k == "--rebuild-similarity-db" || k == "--list-tiles";
    if (eq != std::string::npos) {
      v = k.substr(eq + 1);
      k = k.substr(0, eq);
    } else if (!flag_only) {
      if (i + 1 >= argc) {
        std::fprintf(stderr, "error: a value is required for '%s'\n", k.c_str());
        return false;
      }
      v = argv[++i];
    }
    if (k == "-t" || k == "--target") a->target = v;
    else if (k == "-m" || k == "--material-src") a->material_src = v;
    else if (k == "-o" || k == "--output") a->output = v;
    else if (k == "--grid-w") a->grid_w = (unsigned)parse_unsigned(k, v, 0xFFFFFFFFull);
    else if (k == "--grid-h") a->grid_h = (unsigned)parse_unsigned(k, v, 0xFFFFFFFFull);
    else if (k == "--max-materials") a->max_materials = (size_t)parse_unsigned(k, v, SIZE_MAX);
    else if (k == "--aspect-tolerance") a->aspect_tolerance = parse_float(k, v);
    else if (k == "--max-usage-per-image") a->max_usage_per_image = (size_t)parse_unsigned(k, v, SIZE_MAX);
    else if (k == "--adjacency-penalty-weight") a->adjacency_penalty_weight = parse_float(k, v);
    else if (k == "--enable-optimization") a->enable_optimization = parse_bool(v);
    else if (k == "--optimization-iterations") a->optimization_iterations = (size_t)parse_unsigned(k, v, SIZE_MAX);
    else if (k == "--similarity-db") a->similarity_db = v;
    else if (k == "--rebuild-similarity-db") a->rebuild_similarity_db = true;
    else if (k == "--color-adjustment-strength") a->color_adjustment_strength = parse_float(k, v);
    else if (k == "--show-time") a->show_time = parse_bool(v);
    else if (k == "--show-grid") a->show_grid = parse_bool(v);
    else if (k == "--seed") a->seed = parse_unsigned(k, v, ~0ull);
    else if (k == "--list-tiles") a->list_tiles = true;
    else {
      std::fprintf(stderr, "error: unexpected argument '%s'\n", k.c_str());
      return false;
    }
  }
  if (a->list_tiles && a->output.empty()) a->output = "-";
  if (a->target.empty() || a->material_src.empty() || a->output.empty()) {
    std::fprintf(stderr, "error: the following required arguments were not provided: --target --material-src --output\n");
    return false;
  }
  return true;
}

static bool read_ppm(const std::string& file, RgbImage* img, bool header_only = false) {
  std::ifstream in(file, std::ios::binary);
  std::string magic;
  int w = 0, h = 0, maxv = 0;
  auto token = [&](std::string* t) {
    t->clear();
    int c;
    while ((c = in.get()) != EOF) {
      if (c == '#') {
        while ((c = in.get()) != EOF && c != '\n') {
        }
      } else if (!std::isspace(c)) {
        t->push_back((char)c);
        break;
      }
    }
    while ((c = in.peek()) != EOF && !std::isspace(c)) t->push_back((char)in.get());
  };
  auto number = [](const std::string& t) -> int {  // decimal digits only, at most 2^24 (no atoi: overflow is undefined)
    if (t.empty() || t.size() > 8) return -1;
    int v = 0;
    for (char c : t) {
      if (c < '0' || c > '9') return -1;
      v = v * 10 + (c - '0');
    }
    return v <= (1 << 24) ? v : -1;
  };
  std::string t;
  token(&magic);
  token(&t), w = number(t);
  token(&t), h = number(t);
  token(&t), maxv = number(t);
  if (magic != "P6" || w <= 0 || h <= 0 || maxv != 255) return false;
  in.get();  // the single whitespace after maxval
  // a header may claim any size: the pixels have to BE in the file before anything of that size is allocated
  const std::streamoff at = in.tellg();
  in.seekg(0, std::ios::end);
  const std::streamoff end = in.tellg();
  const uint64_t need = (uint64_t)w * (uint64_t)h * 3u;
  if (at < 0 || end < at || (uint64_t)(end - at) < need) return false;
  in.seekg(at);
  img->width = w;
  img->height = h;
  if (header_only) return true;
  img->data.resize((size_t)need);
  in.read((char*)img->data.data(), (std::streamsize)img->data.size());
  return (size_t)in.gcount() == img->data.size();
}

static bool write_ppm(const std::string& file, const RgbImage& img) {
  std::ofstream out(file, std::ios::binary);
  out << "P6\n" << img.width << " " << img.height << "\n255\n";
  out.write((const char*)img.data.data(), (std::streamsize)img.data.size());
  return (bool)out;
}

// fn(i) for i in [0, n) on the host's threads (MOSAIC_HOST_THREADS overrides the count); every i writes its own slot
template <typename F>
static void parallel_for(size_t n, F fn) {
  unsigned nt = std::thread::hardware_concurrency();
  if (const char* e = std::getenv("MOSAIC_HOST_THREADS")) nt = (unsigned)std::max(1, std::atoi(e));
  nt = (unsigned)std::min<size_t>(std::max(1u, nt), n);
  if (nt <= 1) {
    for (size_t i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<size_t> next{0};
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nt; ++t)
    pool.emplace_back([&] {
      for (size_t i; (i = next.fetch_add(1)) < n;) fn(i);
    });
  for (std::thread& t : pool) t.join();
}

static std::string lower_ext(const std::string& file) {  // main.rs:177-182: extension, lower-cased
  const size_t dot = file.find_last_of('.'), slash = file.find_last_of('/');
  if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return "";
  std::string e = file.substr(dot + 1);
  for (char& c : e) c = (char)std::tolower((unsigned char)c);
  return e;
}

// image::open(path)?.to_rgb8() (main.rs:268, 484) for the formats this host has
static bool read_image(const std::string& file, RgbImage* img, bool header_only = false) {
  const std::string ext = lower_ext(file);
  if (ext == "jpg" || ext == "jpeg") {
    try {
      if (header_only) return codec::jpeg_dimensions(file, &img->width, &img->height);
      codec::JpegImage d = codec::decode_jpeg(file);
      img->width = d.width, img->height = d.height;
      img->data = std::move(d.rgb);
      return true;
    } catch (const std::exception& e) {
      std::fprintf(stderr, "%s\n", e.what());
      return false;
    }
  }
  if (ext == "webp") {
    try {
      if (header_only) return codec::webp_dimensions(file, &img->width, &img->height);
      codec::Image d = codec::decode_webp(file);
      img->width = d.width, img->height = d.height;
      img->data = std::move(d.rgb);
      return true;
    } catch (const std::exception& e) {
      std::fprintf(stderr, "%s\n", e.what());
      return false;
    }
  }
  if (ext != "png") return read_ppm(file, img, header_only);
  try {
    if (header_only) return codec::png_dimensions(file, &img->width, &img->height);
    codec::Image d = codec::decode_png(file);
    img->width = d.width, img->height = d.height;
    img->data = std::move(d.rgb);
    return true;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return false;
  }
}

static bool write_image(const std::string& file, const RgbImage& img) {  // output_img.save(path): by extension
  const std::string ext = lower_ext(file);
  if (ext == "ppm") return write_ppm(file, img);
  if (ext != "png" && ext != "jpg" && ext != "jpeg" && ext != "webp") {  // `save` fails on an extension it has no encoder for
    std::fprintf(stderr, "unsupported output extension \"%s\" (png, jpg, jpeg, webp or ppm)\n", ext.c_str());
    return false;
  }
  try {
    if (ext == "png")
      codec::encode_png(file, img.data.data(), img.width, img.height);
    else if (ext == "webp")
      codec::encode_webp(file, img.data.data(), img.width, img.height);  // (lossless VP8L, as the `image` crate writes WebP)
    else
      codec::encode_jpeg(file, img.data.data(), img.width, img.height, 75);  // (the `image` crate's default quality)
    return true;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return false;
  }
}

// load_tiles (main.rs:166-261).  The reference decodes EVERY file of the directory in its first pass (`process_tile` opens
// the image, main.rs:268), drops the ones that fail with a message and the ones off the target's aspect (lib.rs:121-123), falls
// back to the first min(len, 2 * max) files without the aspect filter when NOTHING is left (main.rs:211-253), and keeps the
// first `max_materials` of what it has (main.rs:255-257).  The same tiles come out of this function with fewer decodes: the
// aspect test needs only the header, and the survivors are decoded in order, a batch at a time on the host's threads (the
// reference's par_iter, main.rs:194), until `max_materials` of them have succeeded — a file whose body is damaged does not
// take a slot, exactly as in the reference, where truncation comes after the failures are gone.  Files are taken in sorted
// order (read_dir's order is unspecified).
static void decode_in_order(const std::vector<std::string>& files, size_t want, std::vector<std::string>* paths, std::vector<RgbImage>* images) {
  size_t next = 0;
  while (paths->size() < want && next < files.size()) {
    const size_t batch = std::min(files.size() - next, want - paths->size());
    std::vector<RgbImage> decoded(batch);
    std::vector<char> ok(batch, 0);
    parallel_for(batch, [&](size_t i) { ok[i] = read_image(files[next + i], &decoded[i]) ? 1 : 0; });
    for (size_t i = 0; i < batch; ++i) {
      if (!ok[i]) {
        std::fprintf(stderr, "Error processing \"%s\": cannot decode\n", files[next + i].c_str());
        continue;
      }
      images->push_back(std::move(decoded[i]));
      paths->push_back(files[next + i]);
    }
    next += batch;
  }
}

static void load_tiles(const std::string& material_dir, float target_aspect, float aspect_tolerance, size_t max_materials,
                       std::vector<std::string>* paths, std::vector<RgbImage>* images) {
  std::printf("Collecting material images...\n");
  std::vector<std::string> entries;
  if (DIR* d = opendir(material_dir.c_str())) {
    while (dirent* e = readdir(d)) {
      const std::string name = e->d_name, path = material_dir + "/" + name;
      struct stat st;
      const std::string ext = lower_ext(name);  // main.rs:175-184: is_file() && the extension, lower-cased, is one of the four
      if ((ext == "png" || ext == "jpg" || ext == "jpeg" || ext == "webp" || ext == "ppm") && stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode))
        entries.push_back(path);
    }
    closedir(d);
  } else {
    throw std::runtime_error("cannot read directory " + material_dir);
  }
  std::sort(entries.begin(), entries.end());
  std::vector<RgbImage> headers(entries.size());
  std::vector<char> header_ok(entries.size(), 0);
  parallel_for(entries.size(), [&](size_t i) { header_ok[i] = read_image(entries[i], &headers[i], true) ? 1 : 0; });
  std::vector<std::string> matching;
  for (size_t i = 0; i < entries.size(); ++i) {
    if (!header_ok[i]) {
      std::fprintf(stderr, "Error processing \"%s\": not a PNG / JPEG / WebP / binary PPM\n", entries[i].c_str());
      continue;
    }
    const float aspect = (float)headers[i].width / (float)headers[i].height;  // main.rs:270
    if (MosaicGeneratorImpl::is_aspect_ratio_match(aspect, target_aspect, aspect_tolerance)) matching.push_back(entries[i]);
  }
  // (at least one, whatever the limit: whether the fallback fires depends on `tiles.is_empty()` BEFORE the truncation,
  // main.rs:211 vs 255 — `--max-materials 0` keeps nothing but does not fall back when a matching tile exists)
  decode_in_order(matching, std::max<size_t>(max_materials, 1), paths, images);
  if (paths->empty()) {  // main.rs:211-253
    std::printf("No tiles matched target aspect ratio %.3f, loading tiles without aspect filtering...\n", target_aspect);
    const size_t take = std::min(entries.size(), max_materials * 2);
    std::printf("Processing %zu material images (sampled from %zu total)...\n", take, entries.size());
    const std::vector<std::string> sampled(entries.begin(), entries.begin() + (std::ptrdiff_t)take);
    decode_in_order(sampled, max_materials, paths, images);  // main.rs:255-257: the first max_materials that decoded
  }
  if (paths->size() > max_materials) {  // main.rs:255-257
    paths->resize(max_materials);
    images->resize(max_materials);
  }
  std::printf("Loaded %zu tiles (aspect ratio matched)\n", paths->size());
}

int main(int argc, char** argv) {
  Args args;
  try {
    if (!parse(argc, argv, &args)) return 2;
  } catch (const BadValue& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 2;
  }
  try {
    const auto t0 = std::chrono::steady_clock::now();
    RgbImage target;
    if (!read_image(args.target, &target)) throw std::runtime_error("cannot read target (PNG, JPEG or binary PPM expected): " + args.target);
    const float target_aspect = (float)target.width / (float)target.height;  // main.rs:745-747
    std::printf("Target aspect ratio: %.3f\n", target_aspect);

    std::vector<RgbImage> images;
    std::vector<std::string> paths;
    load_tiles(args.material_src, target_aspect, args.aspect_tolerance, args.max_materials, &paths, &images);
    if (args.list_tiles) {  // the selection alone (no device needed): what load_tiles would hand to MosaicGenerator::new
      for (size_t i = 0; i < paths.size(); ++i) std::printf("tile: %s %dx%d\n", paths[i].c_str(), images[i].width, images[i].height);
      return 0;
    }
    if (paths.empty()) throw std::runtime_error("no material images found in " + args.material_src);
    const auto t1 = std::chrono::steady_clock::now();

    MosaicGenerator generator(paths, images, args.max_usage_per_image, args.similarity_db, args.rebuild_similarity_db,
                              args.adjacency_penalty_weight, args.color_adjustment_strength);
    if (!generator.save_warning().empty()) std::fprintf(stderr, "%s\n", generator.save_warning().c_str());
    const auto t2 = std::chrono::steady_clock::now();

    if (args.grid_w == 0 || args.grid_h == 0 || args.grid_w > (unsigned)target.width || args.grid_h > (unsigned)target.height)
      throw std::runtime_error("the grid (" + std::to_string(args.grid_w) + "x" + std::to_string(args.grid_h) +
                               ") needs at least one pixel of the target per tile (the reference divides by zero here, main.rs:495-496)");
    std::printf("Target image: %dx%d\n", target.width, target.height);
    std::printf("Grid: %ux%u, Tile size: %ux%u\n", args.grid_w, args.grid_h, target.width / args.grid_w, target.height / args.grid_h);
    mosaic_report rep{};
    const RgbImage out = generator.generate_mosaic(target, (int)args.grid_w, (int)args.grid_h, args.enable_optimization,
                                                   args.optimization_iterations, args.seed, nullptr, &rep);
    const auto t3 = std::chrono::steady_clock::now();
    if (rep.annealed)
      std::printf("Optimization complete: cost %.3f -> %.3f, improved %llu, accepted %llu of %llu\n", rep.opt.initial_cost,
                  rep.opt.final_cost, (unsigned long long)rep.opt.improved_count, (unsigned long long)rep.opt.accepted_count,
                  (unsigned long long)rep.opt.iterations);
    if (rep.n_fallbacks) std::printf("Fallback selections: %u\n", rep.n_fallbacks);
    if (!write_image(args.output, out)) throw std::runtime_error("cannot write " + args.output);
    const auto t4 = std::chrono::steady_clock::now();
    if (args.show_time) {
      auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
      std::printf("Time: load %.1f ms, materials + similarity database %.1f ms, mosaic %.1f ms, save %.1f ms\n", ms(t0, t1), ms(t1, t2),
                  ms(t2, t3), ms(t3, t4));
    }
    std::printf("Mosaic saved to \"%s\"\n", args.output.c_str());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "Error: %s\n", e.what());
    return 1;
  }
  return 0;
}
