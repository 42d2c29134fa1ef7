/* stand-in for <cuda_gl_interop.h>, absent on this machine: see refshim.h (test infrastructure only) */
#include "refshim.h"
