// homography.cu -- FindHomography entry point (matching.cu:1000-1087 in the reference).
// Outside the round-1 hot path; defined so that callers link.
#include "common.cuh"
double FindHomography(SiftData &data, float *homography, int *numMatches, int numLoops, float minScore,
                      float maxAmbiguity, float thresh)
{
  (void)data; (void)numLoops; (void)minScore; (void)maxAmbiguity; (void)thresh;
  fprintf(stderr, "cudasift_b200: FindHomography is not part of the round-1 hot path (see DESIGN.md)\n");
  if (homography) { for (int i = 0; i < 9; i++) homography[i] = (i % 4 == 0) ? 1.0f : 0.0f; }
  if (numMatches) *numMatches = 0;
  return 0.0;
}
