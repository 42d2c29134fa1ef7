/* stand-in for <opencv2/highgui/highgui.hpp>, absent on this machine: see refshim.h (test infrastructure only) */
#include "../../refshim.h"
