/* stand-in for <GL/glew.h>, absent on this machine: see refshim.h (test infrastructure only) */
#include "../refshim.h"
