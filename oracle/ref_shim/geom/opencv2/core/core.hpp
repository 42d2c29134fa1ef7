/* stand-in, absent on this machine: see cvshim.h (test infrastructure only) */
#include "cvshim.h"
