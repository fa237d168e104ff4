// sbx_internal.h -- shared declarations of the host library (not part of the ABI).
#ifndef SBX_INTERNAL_H_
#define SBX_INTERNAL_H_
#include <string>
#include <vector>

#include <cuda.h>

#include "../../include/sbx.h"

namespace sbx {
// ---- lazily bound driver API (dlopen: the library loads, and exports its symbols, on a machine without a driver)
struct driver_api {
    void* lib = nullptr;
    CUresult (*Init)(unsigned);
    CUresult (*DeviceGet)(CUdevice*, int);
    CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice);
    CUresult (*DevicePrimaryCtxRetain)(CUcontext*, CUdevice);
    CUresult (*DevicePrimaryCtxRelease)(CUdevice);
    CUresult (*CtxPushCurrent)(CUcontext);
    CUresult (*CtxPopCurrent)(CUcontext*);
    CUresult (*CtxSynchronize)(void);
    CUresult (*ModuleLoadData)(CUmodule*, const void*);
    CUresult (*ModuleUnload)(CUmodule);
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*);
    CUresult (*ModuleGetGlobal)(CUdeviceptr*, size_t*, CUmodule, const char*);
    CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction);
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int);
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t);
    CUresult (*MemAlloc)(CUdeviceptr*, size_t);
    CUresult (*MemFree)(CUdeviceptr);
    CUresult (*MemcpyHtoD)(CUdeviceptr, const void*, size_t);
    CUresult (*MemcpyDtoH)(void*, CUdeviceptr, size_t);
    CUresult (*MemcpyDtoHAsync)(void*, CUdeviceptr, size_t, CUstream);
    CUresult (*MemcpyHtoDAsync)(CUdeviceptr, const void*, size_t, CUstream);
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             CUstream, void**, void**);
    CUresult (*StreamSynchronize)(CUstream);
    CUresult (*EventCreate)(CUevent*, unsigned);
    CUresult (*EventRecord)(CUevent, CUstream);
    CUresult (*EventSynchronize)(CUevent);
    CUresult (*EventElapsedTime)(float*, CUevent, CUevent);
    CUresult (*EventDestroy)(CUevent);
    CUresult (*GetErrorString)(CUresult, const char**);
    CUresult (*PointerGetAttribute)(void*, CUpointer_attribute, CUdeviceptr);
    CUresult (*MemHostRegister)(void*, size_t, unsigned);
    CUresult (*MemHostUnregister)(void*);
    CUresult (*MemHostGetDevicePointer)(CUdeviceptr*, void*, unsigned);
    CUresult (*IpcGetMemHandle)(CUipcMemHandle*, CUdeviceptr);
    CUresult (*IpcOpenMemHandle)(CUdeviceptr*, CUipcMemHandle, unsigned);
    CUresult (*IpcCloseMemHandle)(CUdeviceptr);
    CUresult (*MemsetD32)(CUdeviceptr, unsigned, size_t);
    CUresult (*MemHostAlloc)(void**, size_t, unsigned);
    CUresult (*MemFreeHost)(void*);
    CUresult (*StreamWaitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned);
    CUresult (*StreamWriteValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned);
    CUresult (*StreamBatchMemOp)(CUstream, unsigned, CUstreamBatchMemOpParams*, unsigned);
    CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUresult (*StreamWaitEvent)(CUstream, CUevent, unsigned);
    CUresult (*StreamCreate)(CUstream*, unsigned);
    CUresult (*StreamDestroy)(CUstream);
    CUresult (*MemAllocAsync)(CUdeviceptr*, size_t, CUstream);
    CUresult (*MemFreeAsync)(CUdeviceptr, CUstream);
    CUresult (*DeviceGetCount)(int*);
    CUresult (*DeviceCanAccessPeer)(int*, CUdevice, CUdevice);
    CUresult (*CtxEnablePeerAccess)(CUcontext, unsigned);
};

driver_api* load_driver(std::string* err);   // NULL + *err when libcuda is missing
CUcontext context_of(sbx_ctx* ctx);          // the primary context a sbx_ctx is bound to

std::string suffix_float_literals(const std::string& src);
std::string bind_hlsl_registers(const std::string& src);
std::string library_dir();
// NVRTC-compile an unchanged app header to an sm_100a cubin.  0 or SBX_ERR_*; log gets the compiler output.
int compile_app_header(const std::string& header_path, const std::string& app_name,
                       const std::vector<std::string>& extra_defines, std::string* cubin, std::string* log);
}  // namespace sbx
#endif
