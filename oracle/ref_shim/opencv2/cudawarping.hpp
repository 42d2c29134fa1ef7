/* stand-in for <opencv2/cudawarping.hpp>, absent on this machine: see refshim.h (test infrastructure only) */
#include "../refshim.h"
