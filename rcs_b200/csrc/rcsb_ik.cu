#include "rcsb_kernels.h"
extern "C" cudaError_t rcsb_launch_ik(const rcsb::IkArgs* a, int nSlots, cudaStream_t stream)
{
  (void) a; (void) nSlots; (void) stream;
  return cudaErrorNotSupported;
}
