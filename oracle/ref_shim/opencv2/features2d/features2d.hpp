/* stand-in for <opencv2/features2d/features2d.hpp>, absent on this machine: see refshim.h (test infrastructure only) */
#include "../../refshim.h"
