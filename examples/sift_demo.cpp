// sift_demo.cpp -- two-view demo written against the drop-in headers only.
//
// Plays the role of the reference's mainSift.cpp (main at :25-93, PrintMatchData at :150-200)
// without OpenCV: 8-bit PGM (P5) or PNG (8-bit grey / RGB / RGBA, decoded here with zlib) in, optional PGM with the
// matches drawn out.  It is a caller of the public API (cudaSift.h / cudaImage.h), nothing here is on the hot path.
//
//   sift_demo left.{pgm,png} right.{pgm,png} [--thresh T] [--octaves N] [--repeat R] [--device D]
//             [--ransac LOOPS] [--out marked.pgm] [--style marks|reference] [--print K]
//   sift_demo --decode-only in.{pgm,png} out.pgm          (image decoding only, no GPU needed)
//
// --style reference draws what the reference's PrintMatchData draws (mainSift.cpp:150-200): a line to the matched
// feature for matches within 5 px of the homography, and a black-and-white cross of half-length 1.41*scale per feature.
// Compiled with -DMANAGEDMEM (against libcudasift_b200_managed.so) it exercises the unified-memory flavour of the
// API (cudaSift.h:35-40, SiftData::m_data).
//
// Build (cudasift_b200/build.py: build_demo):
//   g++ -O2 -Iinclude examples/sift_demo.cpp -Lcudasift_b200/lib -lcudasift_b200 -lz -o sift_demo
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <zlib.h>

#include "cudaSift.h"

#ifdef MANAGEDMEM
#define HOST_POINTS(d) ((d).m_data)
#else
#define HOST_POINTS(d) ((d).h_data)
#endif

namespace {

struct GrayImage {
  int w = 0, h = 0;
  std::vector<float> px;   // row-major, 0..255
};

// Skips whitespace and '#' comment lines of a netpbm header, then reads one integer.
bool pnm_int(FILE *f, int *v)
{
  int c = fgetc(f);
  while (c != EOF) {
    if (c == '#') { while (c != '\n' && c != EOF) c = fgetc(f); }
    else if (c == ' ' || c == '\t' || c == '\n' || c == '\r') c = fgetc(f);
    else break;
  }
  if (c == EOF) return false;
  ungetc(c, f);
  return fscanf(f, "%d", v) == 1;
}

bool read_pgm(const char *path, GrayImage *img)
{
  FILE *f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); return false; }
  char magic[3] = {0, 0, 0};
  int maxval = 0;
  bool ok = fread(magic, 1, 2, f) == 2 && magic[0] == 'P' && magic[1] == '5' && pnm_int(f, &img->w) &&
            pnm_int(f, &img->h) && pnm_int(f, &maxval) && maxval > 0 && maxval < 65536;
  if (ok) {
    fgetc(f);   // the single whitespace byte that ends the header
    size_t n = (size_t)img->w * img->h, bpp = maxval > 255 ? 2 : 1;
    std::vector<unsigned char> raw(n * bpp);
    ok = fread(raw.data(), 1, raw.size(), f) == raw.size();
    img->px.resize(n);
    for (size_t i = 0; ok && i < n; i++)
      img->px[i] = bpp == 1 ? (float)raw[i] : (float)((raw[2 * i] << 8) | raw[2 * i + 1]) * 255.0f / (float)maxval;
  }
  fclose(f);
  if (!ok) fprintf(stderr, "%s: not a binary PGM (P5)\n", path);
  return ok;
}

// Minimal PNG reader: 8 bits per channel, colour types 0 (grey), 2 (RGB), 4 (grey+alpha), 6 (RGBA), non-interlaced.
// Colour goes to grey with the weights of cv::cvtColor(BGR2GRAY): (9798 R + 19235 G + 3735 B + 16384) >> 15
// (mainSift.cpp:37 reads its PNGs with cv::imread(.., 0)).
bool read_png(const char *path, GrayImage *img)
{
  FILE *f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); return false; }
  std::vector<unsigned char> file;
  unsigned char buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) file.insert(file.end(), buf, buf + n);
  fclose(f);
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  if (file.size() < 33 || memcmp(file.data(), sig, 8) != 0) { fprintf(stderr, "%s: not a PNG\n", path); return false; }
  auto be32 = [&](size_t o) { return ((unsigned)file[o] << 24) | ((unsigned)file[o + 1] << 16) | ((unsigned)file[o + 2] << 8) | file[o + 3]; };
  int depth = 0, ctype = 0, interlace = 0;
  std::vector<unsigned char> idat;
  for (size_t o = 8; o + 12 <= file.size();) {
    const size_t len = be32(o);
    if (o + 12 + len > file.size()) break;
    const char *type = (const char *)&file[o + 4];
    if (!memcmp(type, "IHDR", 4) && len >= 13) {
      img->w = (int)be32(o + 8); img->h = (int)be32(o + 12);
      depth = file[o + 16]; ctype = file[o + 17]; interlace = file[o + 20];
    } else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), file.begin() + o + 8, file.begin() + o + 8 + len);
    else if (!memcmp(type, "IEND", 4)) break;
    o += 12 + len;
  }
  const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
  if (depth != 8 || ch == 0 || interlace != 0 || img->w < 1 || img->h < 1) {
    fprintf(stderr, "%s: unsupported PNG (need 8-bit grey/RGB/RGBA, non-interlaced)\n", path);
    return false;
  }
  const size_t stride = (size_t)img->w * ch;
  std::vector<unsigned char> raw((stride + 1) * img->h);
  uLongf rawLen = (uLongf)raw.size();
  if (uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size()) {
    fprintf(stderr, "%s: corrupt PNG data\n", path);
    return false;
  }
  std::vector<unsigned char> prev(stride, 0), cur(stride);
  img->px.resize((size_t)img->w * img->h);
  for (int y = 0; y < img->h; y++) {
    const unsigned char *line = &raw[(stride + 1) * y];
    const int filter = line[0];
    for (size_t i = 0; i < stride; i++) {
      const int a = i >= (size_t)ch ? cur[i - ch] : 0, b = prev[i], c = i >= (size_t)ch ? prev[i - ch] : 0;
      int pred = 0;
      if (filter == 1) pred = a;
      else if (filter == 2) pred = b;
      else if (filter == 3) pred = (a + b) >> 1;
      else if (filter == 4) { const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }
      cur[i] = (unsigned char)(line[1 + i] + pred);
    }
    for (int x = 0; x < img->w; x++) {
      const unsigned char *q = &cur[(size_t)x * ch];
      img->px[(size_t)y * img->w + x] = ch <= 2 ? (float)q[0] : (float)((9798 * q[0] + 19235 * q[1] + 3735 * q[2] + 16384) >> 15);
    }
    prev.swap(cur);
  }
  return true;
}

bool read_image(const char *path, GrayImage *img)
{
  const size_t n = strlen(path);
  if (n > 4 && (!strcmp(path + n - 4, ".png") || !strcmp(path + n - 4, ".PNG"))) return read_png(path, img);
  return read_pgm(path, img);
}

bool write_pgm(const char *path, const GrayImage &img)
{
  FILE *f = fopen(path, "wb");
  if (!f) return false;
  fprintf(f, "P5\n%d %d\n255\n", img.w, img.h);
  std::vector<unsigned char> raw(img.px.size());
  for (size_t i = 0; i < raw.size(); i++) raw[i] = (unsigned char)std::min(255.0f, std::max(0.0f, img.px[i]));
  bool ok = fwrite(raw.data(), 1, raw.size(), f) == raw.size();
  fclose(f);
  return ok;
}

void draw_segment(GrayImage *img, float x0, float y0, float x1, float y1, float value)
{
  int steps = (int)std::ceil(std::max(std::fabs(x1 - x0), std::fabs(y1 - y0))) + 1;
  for (int s = 0; s <= steps; s++) {
    float t = (float)s / (float)steps;
    int x = (int)std::lround(x0 + t * (x1 - x0)), y = (int)std::lround(y0 + t * (y1 - y0));
    if (x >= 0 && y >= 0 && x < img->w && y < img->h) img->px[(size_t)y * img->w + x] = value;
  }
}

// Marks every feature with a scale-sized cross in its orientation and, for matches that agree with
// the homography (match_error < 5 px), the displacement to the matched position.
void mark_features(GrayImage *img, const SiftData &d)
{
  for (int i = 0; i < d.numPts; i++) {
    const SiftPoint &p = HOST_POINTS(d)[i];
    float r = 2.0f * p.scale, a = p.orientation * 3.14159265f / 180.0f;
    float cx = r * std::cos(a), cy = r * std::sin(a);
    draw_segment(img, p.xpos - cx, p.ypos - cy, p.xpos + cx, p.ypos + cy, 255.0f);
    draw_segment(img, p.xpos + cy, p.ypos - cx, p.xpos - cy, p.ypos + cx, 0.0f);
    if (p.match >= 0 && p.match_error < 5.0f) draw_segment(img, p.xpos, p.ypos, p.match_xpos, p.match_ypos, 255.0f);
  }
}

// What the reference's PrintMatchData draws (mainSift.cpp:150-200), with every write bounds-checked: for matches whose
// match_error is below 5 px a white line from the feature towards the matched feature's position in the other image,
// and for every feature a black cross offset by (+1,+1) under a white cross, arms of min(distance to the border,
// (int)(1.41 * scale)) pixels.
void draw_like_reference(GrayImage *img, const SiftData &d1, const SiftData &d2)
{
  const SiftPoint *a = HOST_POINTS(d1), *b = HOST_POINTS(d2);
  const int w = img->w, h = img->h;
  auto put = [&](int x, int y, float v) { if (x >= 0 && y >= 0 && x < w && y < h) img->px[(size_t)y * w + x] = v; };
  for (int j = 0; j < d1.numPts; j++) {
    const int k = a[j].match;
    if (k >= 0 && k < d2.numPts && a[j].match_error < 5.0f) {
      const float dx = b[k].xpos - a[j].xpos, dy = b[k].ypos - a[j].ypos;
      const int len = (int)std::max(std::fabs(dx), std::fabs(dy));
      for (int l = 0; l < len; l++) put((int)(a[j].xpos + dx * l / len), (int)(a[j].ypos + dy * l / len), 255.0f);
    }
    const int x = (int)(a[j].xpos + 0.5f), y = (int)(a[j].ypos + 0.5f);
    const int s = std::min(std::min(x, y), std::min(std::min(w - x - 2, h - y - 2), (int)(1.41f * a[j].scale)));
    for (int t = 0; t < s; t++) { put(x + 1 - t, y + 1, 0.0f); put(x + 1 + t, y + 1, 0.0f); put(x + 1, y + 1 - t, 0.0f); put(x + 1, y + 1 + t, 0.0f); }
    for (int t = 0; t < s; t++) { put(x - t, y, 255.0f); put(x + t, y, 255.0f); put(x, y - t, 255.0f); put(x, y + t, 255.0f); }
  }
}

const char *arg_value(int argc, char **argv, const char *name, const char *dflt)
{
  for (int i = 3; i + 1 < argc; i++)
    if (!strcmp(argv[i], name)) return argv[i + 1];
  return dflt;
}

double now_ms()
{
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

int main(int argc, char **argv)
{
  if (argc == 4 && !strcmp(argv[1], "--decode-only")) {        // image I/O check without a GPU: in.{png,pgm} -> out.pgm
    GrayImage img;
    if (!read_image(argv[2], &img) || !write_pgm(argv[3], img)) return 1;
    printf("Image size = (%d,%d)\n", img.w, img.h);
    return 0;
  }
  if (argc < 3) {
    fprintf(stderr, "usage: %s left.{pgm,png} right.{pgm,png} [--thresh T] [--octaves N] [--repeat R] [--device D] "
                    "[--ransac LOOPS] [--out marked.pgm] [--style marks|reference] [--print K]\n", argv[0]);
    return 2;
  }
  const float thresh = (float)atof(arg_value(argc, argv, "--thresh", "3.0"));
  const int octaves = atoi(arg_value(argc, argv, "--octaves", "5"));
  const int repeat = std::max(1, atoi(arg_value(argc, argv, "--repeat", "100")));
  const int device = atoi(arg_value(argc, argv, "--device", "0"));
  const int ransac = atoi(arg_value(argc, argv, "--ransac", "10000"));
  const int nprint = atoi(arg_value(argc, argv, "--print", "0"));
  const char *out = arg_value(argc, argv, "--out", "");
  const char *style = arg_value(argc, argv, "--style", "marks");
  const float initBlur = 1.0f;

  GrayImage left, right;
  if (!read_image(argv[1], &left) || !read_image(argv[2], &right)) return 1;
  if (left.w != right.w || left.h != right.h) { fprintf(stderr, "image sizes differ\n"); return 1; }
  printf("Image size = (%d,%d)\n", left.w, left.h);

  InitCuda(device);
  CudaImage img1, img2;
  img1.Allocate(left.w, left.h, iAlignUp(left.w, 128), false, NULL, left.px.data());
  img2.Allocate(right.w, right.h, iAlignUp(right.w, 128), false, NULL, right.px.data());
  double up = img1.Download() + img2.Download();

  SiftData sift1, sift2;
  InitSiftData(sift1, 32768, true, true);
  InitSiftData(sift2, 32768, true, true);
  float *scratch = AllocSiftTempMemory(left.w, left.h, octaves, false);
  ExtractSift(sift1, img1, octaves, initBlur, thresh, 0.0f, false, scratch);   // warm-up
  double t0 = now_ms();
  for (int r = 0; r < repeat; r++) {
    ExtractSift(sift1, img1, octaves, initBlur, thresh, 0.0f, false, scratch);
    ExtractSift(sift2, img2, octaves, initBlur, thresh, 0.0f, false, scratch);
  }
  double extract_ms = (now_ms() - t0) / (2.0 * repeat);
  FreeSiftTempMemory(scratch);

  double match_ms = MatchSiftData(sift1, sift2);
  float homography[9];
  int numMatches = 0;
  FindHomography(sift1, homography, &numMatches, ransac, 0.00f, 0.80f, 5.0f);
  int numFit = ImproveHomography(sift1, homography, 5, 0.00f, 0.80f, 3.0f);

  printf("Number of original features: %d %d\n", sift1.numPts, sift2.numPts);
  printf("Number of matching features: %d %d %g%% %g %g\n", numFit, numMatches,
         100.0f * numFit / std::max(1, std::min(sift1.numPts, sift2.numPts)), initBlur, thresh);
  printf("Upload %.3f ms, ExtractSift %.3f ms/image (blocking calls), MatchSiftData %.3f ms\n", up, extract_ms, match_ms);
  printf("Homography:");
  for (int i = 0; i < 9; i++) printf("%s%.6g", i % 3 == 0 ? "\n  " : " ", homography[i]);
  printf("\n");
  for (int i = 0; i < std::min(nprint, sift1.numPts); i++) {
    const SiftPoint &p = HOST_POINTS(sift1)[i];
    printf("%5d: (%7.2f,%7.2f) scale %5.2f ori %6.1f -> %5d (%7.2f,%7.2f) score %.4f ambiguity %.4f error %.2f\n", i,
           p.xpos, p.ypos, p.scale, p.orientation, p.match, p.match_xpos, p.match_ypos, p.score, p.ambiguity, p.match_error);
  }
  if (out[0]) {
    if (!strcmp(style, "reference")) draw_like_reference(&left, sift1, sift2);
    else mark_features(&left, sift1);
    if (!write_pgm(out, left)) { fprintf(stderr, "cannot write %s\n", out); return 1; }
    printf("Wrote %s\n", out);
  }
  FreeSiftData(sift1);
  FreeSiftData(sift2);
  return 0;
}
