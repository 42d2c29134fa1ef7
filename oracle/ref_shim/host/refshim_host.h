/* refshim_host.h -- TEST INFRASTRUCTURE ONLY (oracle/README_ref.md).
 * Stand-ins for the include names Thirdparty/Localization/{nmiSearchKernel,helperFunctions}.cpp
 * ask for that do not exist on this machine: OpenCV headers (nothing of OpenCV is used by the
 * two functions compiled -- find_max_elements, log), <windows.h>/<Shellapi.h>, and the two
 * project headers spelled with the wrong case for a case-sensitive file system
 * ("NmiSearchKernel.hpp" -> nmiSearchKernel.hpp; "iodata.hpp" -> not needed by these functions).
 * kernel.cuh (included by helperFunctions.cpp) only needs the name cv::cuda::PtrStep.       */
#ifndef NMI_REFSHIM_HOST_H_
#define NMI_REFSHIM_HOST_H_
#include <cstddef>
#include <sstream>
#include <string>
namespace cv { namespace cuda {
template <typename T> struct PtrStep { T* data; size_t step; };
} }
#endif
