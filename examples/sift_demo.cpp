// sift_demo.cpp -- two-view demo written against the drop-in headers only.
//
// Plays the role of the reference's mainSift.cpp (main at :25-93, PrintMatchData at :150-200)
// without OpenCV: 8-bit PGM (P5) in, optional PGM with the match vectors drawn out.  It is a
// caller of the public API (cudaSift.h / cudaImage.h), nothing here is on the hot path.
//
//   sift_demo left.pgm right.pgm [--thresh T] [--octaves N] [--repeat R] [--device D]
//             [--ransac LOOPS] [--out marked.pgm] [--print K]
//
// Build (cudasift_b200/build.py: build_demo):
//   g++ -O2 -Iinclude examples/sift_demo.cpp -Lcudasift_b200/lib -lcudasift_b200 -o sift_demo
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cudaSift.h"

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
    const SiftPoint &p = d.h_data[i];
    float r = 2.0f * p.scale, a = p.orientation * 3.14159265f / 180.0f;
    float cx = r * std::cos(a), cy = r * std::sin(a);
    draw_segment(img, p.xpos - cx, p.ypos - cy, p.xpos + cx, p.ypos + cy, 255.0f);
    draw_segment(img, p.xpos + cy, p.ypos - cx, p.xpos - cy, p.ypos + cx, 0.0f);
    if (p.match >= 0 && p.match_error < 5.0f) draw_segment(img, p.xpos, p.ypos, p.match_xpos, p.match_ypos, 255.0f);
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
  if (argc < 3) {
    fprintf(stderr, "usage: %s left.pgm right.pgm [--thresh T] [--octaves N] [--repeat R] [--device D] "
                    "[--ransac LOOPS] [--out marked.pgm] [--print K]\n", argv[0]);
    return 2;
  }
  const float thresh = (float)atof(arg_value(argc, argv, "--thresh", "3.0"));
  const int octaves = atoi(arg_value(argc, argv, "--octaves", "5"));
  const int repeat = std::max(1, atoi(arg_value(argc, argv, "--repeat", "100")));
  const int device = atoi(arg_value(argc, argv, "--device", "0"));
  const int ransac = atoi(arg_value(argc, argv, "--ransac", "10000"));
  const int nprint = atoi(arg_value(argc, argv, "--print", "0"));
  const char *out = arg_value(argc, argv, "--out", "");
  const float initBlur = 1.0f;

  GrayImage left, right;
  if (!read_pgm(argv[1], &left) || !read_pgm(argv[2], &right)) return 1;
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
    const SiftPoint &p = sift1.h_data[i];
    printf("%5d: (%7.2f,%7.2f) scale %5.2f ori %6.1f -> %5d (%7.2f,%7.2f) score %.4f ambiguity %.4f error %.2f\n", i,
           p.xpos, p.ypos, p.scale, p.orientation, p.match, p.match_xpos, p.match_ypos, p.score, p.ambiguity, p.match_error);
  }
  if (out[0]) {
    mark_features(&left, sift1);
    if (!write_pgm(out, left)) { fprintf(stderr, "cannot write %s\n", out); return 1; }
    printf("Wrote %s\n", out);
  }
  FreeSiftData(sift1);
  FreeSiftData(sift2);
  return 0;
}
