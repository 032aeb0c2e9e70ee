// Tensor-core-mode build of the backward interpreter (see program_fwd_fast.cu).
#define DFOL_PROGRAM_FAST 1
#include "program_bwd.cu"
