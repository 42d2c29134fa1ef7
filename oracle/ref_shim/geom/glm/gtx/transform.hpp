/* stand-in, absent on this machine: see glmshim.h (test infrastructure only) */
#include "glmshim.h"
