// cudaSift.h -- drop-in public API of the cudasift_b200 library.
//
// Source-compatible with the reference's cudaSift.h (Celebrandil/CudaSift,
// cudaSift.h:6-43).  The same C++-mangled symbols are exported by
// libcudasift_b200.so, so mainSift.cpp-style callers re-link without edits.
// Struct layouts are ABI (callers index the arrays directly, mainSift.cpp:101-102,
// geomFuncs.cpp:13-69):  sizeof(SiftPoint) == 576 with data[] at byte 64,
// sizeof(SiftData) == 24.
#ifndef CUDASIFT_H
#define CUDASIFT_H

#include "cudaImage.h"

typedef struct {
  float xpos;          // full-resolution pixel coordinates after ExtractSift
  float ypos;
  float scale;
  float sharpness;     // refined DoG value
  float edgeness;      // tr(H)^2/det(H) of the 2x2 spatial Hessian
  float orientation;   // degrees, [0,360)
  float score;         // best correlation (MatchSiftData)
  float ambiguity;     // second best / best
  int match;           // index into the other set, -1 if none
  float match_xpos;
  float match_ypos;
  float match_error;   // filled by the CPU homography refinement only
  float subsampling;   // 1,2,4,... = octave the point was found in
  float empty[3];
  float data[128];     // descriptor, [ycell][xcell][angle]
} SiftPoint;

typedef struct {
  int numPts;          // valid points
  int maxPts;          // capacity
#ifdef MANAGEDMEM
  SiftPoint *m_data;
#else
  SiftPoint *h_data;   // host copy (may be NULL)
  SiftPoint *d_data;   // device array (may be NULL)
#endif
} SiftData;

// cudaSiftH.cu:19-37.  Selects the device; prints the device banner only when
// CUDASIFT_VERBOSE is set in the environment.
void InitCuda(int devNum = 0);
// cudaSiftH.cu:39-70.  Scratch arena for ExtractSift; same sizing rule as the reference.
float *AllocSiftTempMemory(int width, int height, int numOctaves, bool scaleUp = false);
void FreeSiftTempMemory(float *memoryTmp);
// cudaSiftH.cu:72-144.  Synchronous: on return numPts, d_data and (if non-NULL) h_data
// are valid.
void ExtractSift(SiftData &siftData, CudaImage &img, int numOctaves, double initBlur, float thresh,
                 float lowestScale = 0.0f, bool scaleUp = false, float *tempMemory = 0);
// cudaSiftH.cu:234-264
void InitSiftData(SiftData &data, int num = 1024, bool host = false, bool dev = true);
void FreeSiftData(SiftData &data);
void PrintSiftData(SiftData &data);
// matching.cu:1090-1206.  Writes score/ambiguity/match/match_xpos/match_ypos of data1;
// returns elapsed milliseconds.
double MatchSiftData(SiftData &data1, SiftData &data2);
// matching.cu:1000-1087 (RANSAC homography; outside the round-1 hot path).
double FindHomography(SiftData &data, float *homography, int *numMatches, int numLoops = 1000,
                      float minScore = 0.85f, float maxAmbiguity = 0.95f, float thresh = 5.0f);
// geomFuncs.cpp:6-72 (declared by the caller in the reference, mainSift.cpp:16).  Host-side
// refinement over data.h_data; fills match_error, returns the number of matches within thresh.
int ImproveHomography(SiftData &data, float *homography, int numLoops, float minScore, float maxAmbiguity,
                      float thresh);

#endif
