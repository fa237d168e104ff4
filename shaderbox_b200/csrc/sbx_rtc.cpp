// sbx_rtc.cpp -- run-time compilation of an UNCHANGED shaderbox app header into an sm_100a
// kernel image.  This is the CUDA counterpart of hlsltoy compiling the same headers with
// D3DCompileFromFile and the HLSL/HLSLTOY macros (util/hlsltoy/src/hlsltoy.cpp:382-388).
//
// Two things the reference's C++ build gets from its compiler flags have to be done here:
//   * -fsingle-precision-constant (src/Makefile:12): every unsuffixed floating literal is fp32.
//     NVRTC has no such switch, so the header TEXT handed to NVRTC is a copy with `f` appended to
//     those literals (sbx_suffix_float_literals).  The file on disk is never modified.
//   * free functions callable per pixel: NVRTC's -default-device makes the app's un-annotated
//     functions __device__.
#include "sbx_internal.h"

#include <dlfcn.h>
#include <nvrtc.h>

#include <cctype>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace sbx {

static bool is_ident(char c) { return std::isalnum((unsigned char)c) || c == '_'; }

// Append 'f' to floating literals that have no suffix.  Comments, string/char literals and
// identifiers are copied through untouched; integer and hex literals are left alone.
std::string suffix_float_literals(const std::string& src) {
    std::string out;
    out.reserve(src.size() + src.size() / 16);
    const size_t n = src.size();
    size_t i = 0;
    while (i < n) {
        const char c = src[i];
        if (c == '/' && i + 1 < n && src[i + 1] == '/') {            // line comment
            while (i < n && src[i] != '\n') out += src[i++];
            continue;
        }
        if (c == '/' && i + 1 < n && src[i + 1] == '*') {            // block comment
            out += src[i++]; out += src[i++];
            while (i < n && !(src[i] == '*' && i + 1 < n && src[i + 1] == '/')) out += src[i++];
            if (i < n) { out += src[i++]; out += src[i++]; }
            continue;
        }
        if (c == '"' || c == '\'') {                                 // string / char literal
            const char q = c;
            out += src[i++];
            while (i < n && src[i] != q && src[i] != '\n') {
                if (src[i] == '\\' && i + 1 < n) out += src[i++];
                out += src[i++];
            }
            if (i < n) out += src[i++];
            continue;
        }
        if (is_ident(c) && !std::isdigit((unsigned char)c)) {        // identifier / keyword
            while (i < n && is_ident(src[i])) out += src[i++];
            continue;
        }
        const bool starts_number =
            std::isdigit((unsigned char)c) || (c == '.' && i + 1 < n && std::isdigit((unsigned char)src[i + 1]));
        if (starts_number) {
            size_t j = i;
            bool hex = (c == '0' && j + 1 < n && (src[j + 1] == 'x' || src[j + 1] == 'X'));
            bool has_dot = false, has_exp = false;
            if (hex) {
                j += 2;
                while (j < n && (std::isxdigit((unsigned char)src[j]) || src[j] == '.')) ++j;
            } else {
                while (j < n && (std::isdigit((unsigned char)src[j]) || src[j] == '.')) {
                    if (src[j] == '.') has_dot = true;
                    ++j;
                }
                if (j < n && (src[j] == 'e' || src[j] == 'E')) {
                    size_t k = j + 1;
                    if (k < n && (src[k] == '+' || src[k] == '-')) ++k;
                    if (k < n && std::isdigit((unsigned char)src[k])) {
                        has_exp = true;
                        j = k;
                        while (j < n && std::isdigit((unsigned char)src[j])) ++j;
                    }
                }
            }
            out.append(src, i, j - i);
            // existing suffix (f, F, l, L, u, U ...) is copied as is
            const bool has_suffix = j < n && is_ident(src[j]);
            while (j < n && is_ident(src[j])) out += src[j++];
            if (!hex && (has_dot || has_exp) && !has_suffix) out += 'f';
            i = j;
            continue;
        }
        out += src[i++];
    }
    return out;
}

// HLSL register bindings: `Type name : register(t1);` becomes `Type name = sbx_hlsl_bind(sbx_T, 1, (const Type*)0);`
// (hlsl_tex.h).  Only the USE_NOISE_TEX branch of src/app_clouds.h (:51-55) has them; for every other header this is
// the identity.  Works on the declaration as a whole, so comments and other uses of ':' are left alone.
std::string bind_hlsl_registers(const std::string& src) {
    std::string out;
    out.reserve(src.size() + 256);
    size_t i = 0;
    const size_t n = src.size();
    while (i < n) {
        // try to match at i:  <ident> <ws> <ident> <ws>? ':' <ws>? "register" <ws>? '(' <letter><digits> ')' <ws>? ';'
        size_t j = i;
        auto skip_ws = [&](size_t k) { while (k < n && (src[k] == ' ' || src[k] == '\t')) ++k; return k; };
        auto ident = [&](size_t k) { size_t e = k; while (e < n && is_ident(src[e])) ++e; return e; };
        const bool at_ident_start = is_ident(src[i]) && !std::isdigit((unsigned char)src[i]) && (i == 0 || !is_ident(src[i - 1]));
        if (at_ident_start) {
            const size_t t_end = ident(j);
            size_t k = skip_ws(t_end);
            const size_t n_beg = k, n_end = ident(k);
            k = skip_ws(n_end);
            if (k > t_end && n_end > n_beg && k < n && src[k] == ':') {
                k = skip_ws(k + 1);
                if (src.compare(k, 8, "register") == 0) {
                    k = skip_ws(k + 8);
                    if (k < n && src[k] == '(') {
                        k = skip_ws(k + 1);
                        if (k < n && std::isalpha((unsigned char)src[k])) {
                            size_t d = k + 1, d_end = d;
                            while (d_end < n && std::isdigit((unsigned char)src[d_end])) ++d_end;
                            size_t c = skip_ws(d_end);
                            if (d_end > d && c < n && src[c] == ')') {
                                c = skip_ws(c + 1);
                                if (c < n && src[c] == ';') {
                                    const std::string type = src.substr(i, t_end - i), name = src.substr(n_beg, n_end - n_beg);
                                    out += type + " " + name + " = sbx_hlsl_bind(sbx_T, " + src.substr(d, d_end - d) + ", (const " + type + "*)0);";
                                    i = c + 1;
                                    continue;
                                }
                            }
                        }
                    }
                }
            }
            out.append(src, i, t_end - i);      // not a binding: copy the identifier and go on
            i = t_end;
            continue;
        }
        out += src[i++];
    }
    return out;
}

std::string library_dir() {
    Dl_info info;
    if (dladdr((void*)&library_dir, &info) && info.dli_fname) {
        std::string p(info.dli_fname);
        const size_t s = p.rfind('/');
        return s == std::string::npos ? std::string(".") : p.substr(0, s);
    }
    return ".";
}

static bool read_file(const std::string& path, std::string* out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    *out = ss.str();
    return true;
}

// NVRTC is loaded lazily so that the library itself loads on machines without the toolkit libs.
struct nvrtc_api {
    void* lib = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    nvrtcResult (*DestroyProgram)(nvrtcProgram*);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
    const char* (*GetErrorString)(nvrtcResult);
};

static nvrtc_api* load_nvrtc(std::string* err) {
    static nvrtc_api api;
    if (api.lib) return &api;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    void* lib = nullptr;
    for (const char* nm : names) {
        lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (lib) break;
    }
    if (!lib) { *err = "cannot load libnvrtc"; return nullptr; }
#define SBX_SYM(field, name)                                                  \
    *(void**)(&api.field) = dlsym(lib, name);                                 \
    if (!api.field) { *err = std::string("libnvrtc lacks ") + name; return nullptr; }
    SBX_SYM(CreateProgram, "nvrtcCreateProgram")
    SBX_SYM(DestroyProgram, "nvrtcDestroyProgram")
    SBX_SYM(CompileProgram, "nvrtcCompileProgram")
    SBX_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    SBX_SYM(GetProgramLog, "nvrtcGetProgramLog")
    SBX_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    SBX_SYM(GetCUBIN, "nvrtcGetCUBIN")
    SBX_SYM(GetErrorString, "nvrtcGetErrorString")
#undef SBX_SYM
    api.lib = lib;
    return &api;
}

int compile_app_header(const std::string& header_path, const std::string& app_name,
                       const std::vector<std::string>& extra_defines, std::string* cubin, std::string* log) {
    std::string text;
    if (!read_file(header_path, &text)) {
        *log = "cannot read " + header_path;
        return SBX_ERR_INVALID;
    }
    std::string err;
    nvrtc_api* rtc = load_nvrtc(&err);
    if (!rtc) { *log = err; return SBX_ERR_COMPILE; }

    const std::string app_text = bind_hlsl_registers(suffix_float_literals(text));
    const std::string dir = library_dir();
    // translation unit: just the kernel header; the app text is an in-memory header
    const char* tu = "#include \"sbx/sbx_kernel.cuh\"\n";
    const char* hdr_src[] = {app_text.c_str()};
    const char* hdr_names[] = {"sbx_app_source.h"};
    nvrtcProgram prog;
    nvrtcResult r = rtc->CreateProgram(&prog, tu, (app_name + ".cu").c_str(), 1, hdr_src, hdr_names);
    if (r != NVRTC_SUCCESS) { *log = rtc->GetErrorString(r); return SBX_ERR_COMPILE; }

    std::vector<std::string> opts = {
        "--gpu-architecture=sm_100a", "-std=c++17", "-default-device", "-lineinfo",
        // strict IEEE: the parity contract of sbx_vec.cuh
        "--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
        "-I" + dir + "/include", "-I" + dir + "/include/sbx", "-I" + dir + "/../include",
        "-D" + app_name + "=1", "-DSBX_APP_HEADER=\"sbx_app_source.h\""};
    for (const auto& d : extra_defines) opts.push_back("-D" + d);
    std::vector<const char*> copts;
    for (const auto& o : opts) copts.push_back(o.c_str());
    r = rtc->CompileProgram(prog, (int)copts.size(), copts.data());
    size_t log_size = 0;
    rtc->GetProgramLogSize(prog, &log_size);
    if (log_size > 1) {
        log->resize(log_size);
        rtc->GetProgramLog(prog, &(*log)[0]);
    }
    if (r != NVRTC_SUCCESS) {
        if (log->empty()) *log = rtc->GetErrorString(r);
        rtc->DestroyProgram(&prog);
        return SBX_ERR_COMPILE;
    }
    size_t sz = 0;
    rtc->GetCUBINSize(prog, &sz);
    cubin->resize(sz);
    rtc->GetCUBIN(prog, &(*cubin)[0]);
    rtc->DestroyProgram(&prog);
    return SBX_OK;
}

}  // namespace sbx
