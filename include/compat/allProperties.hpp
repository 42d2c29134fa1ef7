// allProperties.hpp -- compile-time knobs of the localization path, same names and
// values as the reference (Thirdparty/Localization/allProperties.hpp:27-50).
// In the B200 build the knobs that matter to the kernels (BG rule, score variant, bin
// count) are ALSO run-time fields of nmi_flags; these macros give the defaults.
#pragma once

#define nmi_prop_MAX_ITERATION_COUNT 4     // allProperties.hpp:27
#define orb_prop_log false                 // :29
#define nmi_prop_RELOC_FREQUENCY 2         // :31
#define nmi_prop_STEPFACTOR 0.5f           // :33
#define nmi_prop_NUMBER_OF_THREADS 12      // :35 (host allocation only)
#define nmi_prop_BG true                   // :39
#ifndef nmi_prop_RENDER
#define nmi_prop_RENDER 4                  // :42  4 = xyz point cloud, 1 = obj mesh
#endif
#define nmi_prop_OUTPUT_LOC "/ORBSLAM_NMI_results"  // :47
#define nmi_prop_MIN_KERNEL_ROTATION 0.001     // rad, :49
#define nmi_prop_MIN_KERNEL_TRANSLATION 0.005  // m,   :50

#define RENDER_TEXTURE 1       // rendering.hpp render modes
#define RENDER_POINT_CLOUD 4
