// Tensor-core-mode build of the forward interpreter: MUFU approximations, relation tiles streamed through a
// bulk-async shared-memory ring, relate hops evaluated in probability space (see program_common.cuh).
#define DFOL_PROGRAM_FAST 1
#include "program_fwd.cu"
