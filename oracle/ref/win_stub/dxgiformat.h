// dxgiformat.h -- TEST INFRASTRUCTURE.  The four Windows SDK names that the reference's vendored
// lib/DirectXTex/DirectXTex/DDS.h needs in order to compile on Linux (the SDK is a third-party
// dependency of the reference that is not in /root/reference).  Values: the published DXGI_FORMAT
// enumeration and the mmsystem.h MAKEFOURCC macro.
#pragma once
#include <stdint.h>
typedef enum DXGI_FORMAT {
    DXGI_FORMAT_UNKNOWN = 0,
    DXGI_FORMAT_R32G32B32A32_FLOAT = 2,
    DXGI_FORMAT_R32G32B32_FLOAT = 6,
    DXGI_FORMAT_R32G32B32_UINT = 7,
    DXGI_FORMAT_R32G32_FLOAT = 16,
    DXGI_FORMAT_R32_FLOAT = 41,
} DXGI_FORMAT;
#ifndef MAKEFOURCC
#define MAKEFOURCC(ch0, ch1, ch2, ch3) \
    ((uint32_t)(uint8_t)(ch0) | ((uint32_t)(uint8_t)(ch1) << 8) | ((uint32_t)(uint8_t)(ch2) << 16) | ((uint32_t)(uint8_t)(ch3) << 24))
#endif
#define __declspec(x)
