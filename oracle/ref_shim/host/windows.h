/* stand-in for "windows.h", absent on this machine: see refshim_host.h (test infrastructure only) */
#include "./refshim_host.h"
