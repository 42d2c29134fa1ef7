/* stand-in for "opencv2/cudawarping.hpp", absent on this machine: see refshim_host.h (test infrastructure only) */
#include "../refshim_host.h"
