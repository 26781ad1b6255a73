// cudaImage.h -- drop-in image container of the cudasift_b200 library.
//
// Source-compatible with the reference's cudaImage.h (Celebrandil/CudaSift,
// cudaImage.h:8-32): a caller written against the reference recompiles unchanged.
// Members are public and accessed directly by callers (mainSift.cpp:51-54), so the
// member order and types below are ABI: sizeof(CudaImage) == 48.
//
// Semantics kept from the reference (cudaImage.cu:15-78):
//   * Allocate() adopts caller-provided device/host pointers; memory it allocates itself
//     (cudaMallocPitch / malloc) is released by the destructor, adopted memory is not.
//   * pitch is in floats.  ExtractSift expects pitch == iAlignUp(width, 128) (reference
//     quirk Q5, cudaSiftH.cu:422-428); Allocate() with devMem == NULL guarantees it for
//     the usual widths because cudaMallocPitch pads rows to 512 bytes.
//   * Download() = blocking host->device copy, Readback() = blocking device->host copy;
//     both return the elapsed milliseconds.
#ifndef CUDAIMAGE_H
#define CUDAIMAGE_H

class CudaImage {
public:
  int width, height;
  int pitch;            // row stride in floats
  float *h_data;        // host pixels (width*height, densely packed rows)
  float *d_data;        // device pixels (pitch*height)
  float *t_data;        // legacy cudaArray handle (InitTexture / CopyToTexture)
  bool d_internalAlloc;
  bool h_internalAlloc;
public:
  CudaImage();
  ~CudaImage();
  void Allocate(int width, int height, int pitch, bool withHost, float *devMem = 0, float *hostMem = 0);
  double Download();
  double Readback();
  double InitTexture();
  double CopyToTexture(CudaImage &dst, bool host);
};

// Integer helpers exported by the reference (cudaImage.cu:10-13).
int iDivUp(int a, int b);
int iDivDown(int a, int b);
int iAlignUp(int a, int b);
int iAlignDown(int a, int b);
// Declared (never defined) by the reference, cudaImage.h:31-32; defined here as a
// wall-clock stopwatch so that callers that reference them link.
void StartTimer(unsigned int *hTimer);
double StopTimer(unsigned int hTimer);

#endif // CUDAIMAGE_H
