// match_tc.cu -- tcgen05 tensor-core matcher (placeholder until the kernel lands).
#include "common.cuh"
namespace cs {
bool match_tensor_supported() { return false; }
int match_tensor(SiftPoint *s1, int n1, const SiftPoint *s2, int n2, cudaStream_t st, unsigned long long stats[4])
{
  stats[0] = stats[1] = stats[2] = 0;
  return match_exact(s1, n1, s2, n2, st);
}
}  // namespace cs
