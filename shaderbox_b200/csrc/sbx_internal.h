// sbx_internal.h -- shared declarations of the host library (not part of the ABI).
#ifndef SBX_INTERNAL_H_
#define SBX_INTERNAL_H_
#include <string>
#include <vector>

#include "../../include/sbx.h"

namespace sbx {
std::string suffix_float_literals(const std::string& src);
std::string library_dir();
// NVRTC-compile an unchanged app header to an sm_100a cubin.  0 or SBX_ERR_*; log gets the compiler output.
int compile_app_header(const std::string& header_path, const std::string& app_name,
                       const std::vector<std::string>& extra_defines, std::string* cubin, std::string* log);
}  // namespace sbx
#endif
