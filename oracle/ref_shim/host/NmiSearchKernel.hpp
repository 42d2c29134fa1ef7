/* the reference includes its own header with this capitalisation (nmiSearchKernel.cpp:21) */
#include "nmiSearchKernel.hpp"
