/* fake_cuda.c -- a MOCK of the CUDA driver API for the CPU tests of the host library (tests/test_host_mock_cpu.py).
 *
 * TEST INFRASTRUCTURE ONLY.  It renders nothing: a kernel launch is recorded (name, grid, block, a copy of its
 * parameters) and otherwise ignored.  It exists so that the host logic of libsbx.so -- image loading, launch planning
 * (regions, grids, issue order), the N-GPU worker threads, the flag protocol, error paths -- can be exercised (also
 * under ThreadSanitizer) in a container without a GPU.  The tests build it into a TEMPORARY directory as
 * libcuda.so.1 and point LD_LIBRARY_PATH of a child process at it; nothing in the repository ships or loads it.
 *
 * Model: "device" memory is host memory; streams execute at enqueue time, in order; events hold a wall-clock stamp.
 * Kernel images are the real sm_100a cubins: the ELF is parsed for symbols, initialised globals (sbx_image_info /
 * sbx_image_hints), the register count, __launch_bounds__ and the parameter sizes (.nv.info), so the host sees the
 * same image properties as on a B200.
 *
 * Environment: SBX_FAKE_GPUS (default 1), SBX_FAKE_CC_MAJOR (default 10), SBX_FAKE_FAIL_ALLOC_AT=n (the n-th
 * device or pinned-host allocation fails, 1-based), SBX_FAKE_NO_PEER=1 (cuCtxEnablePeerAccess fails).
 */
#define _GNU_SOURCE
#include <elf.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <cuda.h>

#include "sbx/sbx_launch.h"   /* the parameter block of sbx_render: its completion flag is honoured below */

#define FAKE_MAX_GPUS 64
#define FAKE_MAX_PARAMS 4
#define FAKE_PARAM_BYTES 1024

typedef struct fake_launch {
    char name[64];
    unsigned grid[3], block[3], smem;
    int device;
    void* stream;
    int n_params;
    unsigned param_size[FAKE_MAX_PARAMS];
    unsigned char param[FAKE_MAX_PARAMS][FAKE_PARAM_BYTES];
} fake_launch;

typedef struct fake_module {
    unsigned char* image;
    size_t bytes;
    const Elf64_Ehdr* eh;
    const Elf64_Shdr* sh;
    const char* shstr;
    const Elf64_Sym* sym;
    size_t n_sym;
    const char* str;
    struct fake_global { char name[64]; void* mem; size_t bytes; struct fake_global* next; } * globals;
    struct fake_function* functions;
} fake_module;

typedef struct fake_function {
    fake_module* mod;
    char name[64];
    int regs, max_threads, n_params;
    unsigned param_size[FAKE_MAX_PARAMS];
    struct fake_function* next;
} fake_function;

typedef struct fake_context { int device; int retained; } fake_context;
typedef struct fake_stream { int device; } fake_stream;
typedef struct fake_event { double ms; int recorded; } fake_event;
typedef struct fake_range { uintptr_t lo, hi; int kind; /* 1 device, 2 pinned host */ int device; struct fake_range* next; } fake_range;

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static fake_context g_ctx[FAKE_MAX_GPUS];
static __thread fake_context* t_stack[16];
static __thread int t_depth;
static fake_range* g_ranges;
static fake_launch* g_log;
static size_t g_log_len, g_log_cap;
static long g_allocs, g_alloc_calls, g_fail_alloc_at;
static int g_gpus = 1, g_cc_major = 10, g_no_peer, g_init;

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void read_env(void) {
    const char* s;
    if ((s = getenv("SBX_FAKE_GPUS"))) g_gpus = atoi(s);
    if (g_gpus < 0) g_gpus = 0;
    if (g_gpus > FAKE_MAX_GPUS) g_gpus = FAKE_MAX_GPUS;
    if ((s = getenv("SBX_FAKE_CC_MAJOR"))) g_cc_major = atoi(s);
    if ((s = getenv("SBX_FAKE_FAIL_ALLOC_AT"))) g_fail_alloc_at = atol(s);
    if ((s = getenv("SBX_FAKE_NO_PEER"))) g_no_peer = atoi(s);
}

static int current_device(void) { return t_depth > 0 ? t_stack[t_depth - 1]->device : -1; }

static void add_range(void* p, size_t bytes, int kind) {
    fake_range* r = (fake_range*)calloc(1, sizeof *r);
    r->lo = (uintptr_t)p; r->hi = r->lo + bytes; r->kind = kind; r->device = current_device();
    pthread_mutex_lock(&g_lock);
    r->next = g_ranges; g_ranges = r;
    pthread_mutex_unlock(&g_lock);
}
static int drop_range(void* p, int kind) {
    int found = 0;
    pthread_mutex_lock(&g_lock);
    for (fake_range** pp = &g_ranges; *pp; pp = &(*pp)->next)
        if ((*pp)->lo == (uintptr_t)p && (*pp)->kind == kind) { fake_range* r = *pp; *pp = r->next; free(r); found = 1; break; }
    pthread_mutex_unlock(&g_lock);
    return found;
}
static int range_kind(uintptr_t p) {
    int kind = 0;
    pthread_mutex_lock(&g_lock);
    for (fake_range* r = g_ranges; r; r = r->next)
        if (p >= r->lo && p < r->hi) { kind = r->kind; break; }
    pthread_mutex_unlock(&g_lock);
    return kind;
}

/* ---- inspection hooks for the tests -------------------------------------------------------------------------- */
size_t fake_cuda_launch_count(void) { pthread_mutex_lock(&g_lock); size_t n = g_log_len; pthread_mutex_unlock(&g_lock); return n; }
int fake_cuda_get_launch(size_t i, fake_launch* out) {
    int ok = 0;
    pthread_mutex_lock(&g_lock);
    if (i < g_log_len) { *out = g_log[i]; ok = 1; }
    pthread_mutex_unlock(&g_lock);
    return ok;
}
void fake_cuda_reset_log(void) { pthread_mutex_lock(&g_lock); g_log_len = 0; pthread_mutex_unlock(&g_lock); }
long fake_cuda_live_allocs(void) { pthread_mutex_lock(&g_lock); long n = g_allocs; pthread_mutex_unlock(&g_lock); return n; }
size_t fake_cuda_sizeof_launch(void) { return sizeof(fake_launch); }
static int g_fail_launch_device = -1, g_fail_launch_count;
/* the next `count` sbx_render launches on `device` fail with CUDA_ERROR_LAUNCH_FAILED */
void fake_cuda_fail_launches(int device, int count) { pthread_mutex_lock(&g_lock); g_fail_launch_device = device; g_fail_launch_count = count; pthread_mutex_unlock(&g_lock); }
void fake_cuda_fail_alloc_at(long n) { pthread_mutex_lock(&g_lock); g_alloc_calls = 0; g_fail_alloc_at = n; pthread_mutex_unlock(&g_lock); }

/* ---- init / devices / contexts ------------------------------------------------------------------------------- */
CUresult cuInit(unsigned flags) {
    (void)flags;
    pthread_mutex_lock(&g_lock);
    if (!g_init) { read_env(); for (int i = 0; i < FAKE_MAX_GPUS; ++i) g_ctx[i].device = i; g_init = 1; }
    pthread_mutex_unlock(&g_lock);
    return g_gpus > 0 ? CUDA_SUCCESS : CUDA_ERROR_NO_DEVICE;
}
CUresult cuDeviceGetCount(int* n) { *n = g_gpus; return CUDA_SUCCESS; }
CUresult cuDeviceGet(CUdevice* dev, int ordinal) {
    if (ordinal < 0 || ordinal >= g_gpus) return CUDA_ERROR_INVALID_DEVICE;
    *dev = ordinal;
    return CUDA_SUCCESS;
}
CUresult cuDeviceGetAttribute(int* v, CUdevice_attribute a, CUdevice dev) {
    if (dev < 0 || dev >= g_gpus) return CUDA_ERROR_INVALID_DEVICE;
    switch (a) {
        case CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR: *v = g_cc_major; break;
        case CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR: *v = 0; break;
        case CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT: *v = 148; break;
        default: *v = 0; break;
    }
    return CUDA_SUCCESS;
}
CUresult cuDevicePrimaryCtxRetain(CUcontext* c, CUdevice dev) {
    if (dev < 0 || dev >= g_gpus) return CUDA_ERROR_INVALID_DEVICE;
    pthread_mutex_lock(&g_lock);
    g_ctx[dev].retained++;
    pthread_mutex_unlock(&g_lock);
    *c = (CUcontext)&g_ctx[dev];
    return CUDA_SUCCESS;
}
CUresult cuDevicePrimaryCtxRelease_v2(CUdevice dev) {
    if (dev < 0 || dev >= g_gpus) return CUDA_ERROR_INVALID_DEVICE;
    pthread_mutex_lock(&g_lock);
    g_ctx[dev].retained--;
    pthread_mutex_unlock(&g_lock);
    return CUDA_SUCCESS;
}
CUresult cuCtxPushCurrent_v2(CUcontext c) {
    if (!c || t_depth >= 16) return CUDA_ERROR_INVALID_CONTEXT;
    t_stack[t_depth++] = (fake_context*)c;
    return CUDA_SUCCESS;
}
CUresult cuCtxPopCurrent_v2(CUcontext* c) {
    if (t_depth <= 0) return CUDA_ERROR_INVALID_CONTEXT;
    *c = (CUcontext)t_stack[--t_depth];
    return CUDA_SUCCESS;
}
CUresult cuCtxSynchronize(void) { return t_depth > 0 ? CUDA_SUCCESS : CUDA_ERROR_INVALID_CONTEXT; }
CUresult cuDeviceCanAccessPeer(int* can, CUdevice a, CUdevice b) { *can = (a != b) && !g_no_peer; return CUDA_SUCCESS; }
CUresult cuCtxEnablePeerAccess(CUcontext peer, unsigned flags) {
    (void)flags;
    if (!peer || t_depth <= 0) return CUDA_ERROR_INVALID_CONTEXT;
    if (g_no_peer) return CUDA_ERROR_PEER_ACCESS_UNSUPPORTED;
    if (((fake_context*)peer)->device == current_device()) return CUDA_ERROR_INVALID_DEVICE;
    return CUDA_SUCCESS;
}
CUresult cuGetErrorString(CUresult r, const char** s) {
    static __thread char buf[48];
    snprintf(buf, sizeof buf, "fake CUDA error %d", (int)r);
    *s = buf;
    return CUDA_SUCCESS;
}

/* ---- modules: the real cubins, parsed ------------------------------------------------------------------------ */
static const Elf64_Shdr* find_section(const fake_module* m, const char* name) {
    for (int i = 0; i < m->eh->e_shnum; ++i)
        if (!strcmp(m->shstr + m->sh[i].sh_name, name)) return &m->sh[i];
    return NULL;
}
static const Elf64_Sym* find_symbol(const fake_module* m, const char* name, int* index) {
    for (size_t i = 0; i < m->n_sym; ++i)
        if (!strcmp(m->str + m->sym[i].st_name, name)) { if (index) *index = (int)i; return &m->sym[i]; }
    return NULL;
}
CUresult cuModuleLoadData(CUmodule* out, const void* image) {
    if (t_depth <= 0) return CUDA_ERROR_INVALID_CONTEXT;
    const Elf64_Ehdr* eh = (const Elf64_Ehdr*)image;
    if (memcmp(eh->e_ident, ELFMAG, SELFMAG) != 0 || eh->e_ident[EI_CLASS] != ELFCLASS64 || eh->e_machine != 190 /* EM_CUDA */)
        return CUDA_ERROR_INVALID_IMAGE;
    const Elf64_Shdr* sh = (const Elf64_Shdr*)((const unsigned char*)image + eh->e_shoff);
    size_t bytes = eh->e_shoff + (size_t)eh->e_shnum * eh->e_shentsize;
    for (int i = 0; i < eh->e_shnum; ++i)
        if (sh[i].sh_type != SHT_NOBITS && sh[i].sh_offset + sh[i].sh_size > bytes) bytes = sh[i].sh_offset + sh[i].sh_size;
    fake_module* m = (fake_module*)calloc(1, sizeof *m);
    m->image = (unsigned char*)malloc(bytes);
    memcpy(m->image, image, bytes);
    m->bytes = bytes;
    m->eh = (const Elf64_Ehdr*)m->image;
    m->sh = (const Elf64_Shdr*)(m->image + m->eh->e_shoff);
    m->shstr = (const char*)(m->image + m->sh[m->eh->e_shstrndx].sh_offset);
    const Elf64_Shdr* st = find_section(m, ".symtab");
    if (!st) { free(m->image); free(m); return CUDA_ERROR_INVALID_IMAGE; }
    m->sym = (const Elf64_Sym*)(m->image + st->sh_offset);
    m->n_sym = st->sh_size / sizeof(Elf64_Sym);
    m->str = (const char*)(m->image + m->sh[st->sh_link].sh_offset);
    *out = (CUmodule)m;
    return CUDA_SUCCESS;
}
CUresult cuModuleUnload(CUmodule mod) {
    fake_module* m = (fake_module*)mod;
    if (!m) return CUDA_ERROR_INVALID_HANDLE;
    for (struct fake_global* g = m->globals; g;) { struct fake_global* n = g->next; drop_range(g->mem, 1); free(g->mem); free(g); g = n; }
    for (fake_function* f = m->functions; f;) { fake_function* n = f->next; free(f); f = n; }
    free(m->image);
    free(m);
    return CUDA_SUCCESS;
}
/* walk .nv.info-style records: { u8 format, u8 attribute, u16 size-or-value } [+ payload when format == 4] */
static void scan_info(const fake_module* m, const Elf64_Shdr* s, int sym_index, fake_function* f) {
    if (!s) return;
    const unsigned char* d = m->image + s->sh_offset;
    for (size_t i = 0; i + 4 <= s->sh_size;) {
        const unsigned fmt = d[i], attr = d[i + 1], val = d[i + 2] | (d[i + 3] << 8);
        const unsigned char* pay = d + i + 4;
        if (fmt == 4) {
            uint32_t w[3] = {0, 0, 0};
            memcpy(w, pay, val < 12 ? val : 12);
            if (attr == 0x2f && (int)w[0] == sym_index) f->regs = (int)w[1];              /* EIATTR_REGCOUNT */
            if (attr == 0x05) f->max_threads = (int)w[0];                                   /* EIATTR_MAX_THREADS */
            if (attr == 0x17) {                                                             /* EIATTR_KPARAM_INFO */
                const unsigned ordinal = w[1] & 0xffffu, size = (w[2] >> 18) & 0x3fffu;
                if (ordinal < FAKE_MAX_PARAMS) { f->param_size[ordinal] = size; if ((int)ordinal + 1 > f->n_params) f->n_params = ordinal + 1; }
            }
            i += 4 + val;
        } else {
            i += 4;
        }
    }
}
CUresult cuModuleGetFunction(CUfunction* out, CUmodule mod, const char* name) {
    fake_module* m = (fake_module*)mod;
    int index = 0;
    const Elf64_Sym* s = find_symbol(m, name, &index);
    if (!s || ELF64_ST_TYPE(s->st_info) != STT_FUNC) return CUDA_ERROR_NOT_FOUND;
    fake_function* f = (fake_function*)calloc(1, sizeof *f);
    f->mod = m;
    snprintf(f->name, sizeof f->name, "%s", name);
    f->max_threads = 1024;
    char sec[96];
    snprintf(sec, sizeof sec, ".nv.info.%s", name);
    scan_info(m, find_section(m, ".nv.info"), index, f);
    scan_info(m, find_section(m, sec), index, f);
    f->next = m->functions;
    m->functions = f;
    *out = (CUfunction)f;
    return CUDA_SUCCESS;
}
CUresult cuModuleGetGlobal_v2(CUdeviceptr* ptr, size_t* bytes, CUmodule mod, const char* name) {
    fake_module* m = (fake_module*)mod;
    for (struct fake_global* g = m->globals; g; g = g->next)
        if (!strcmp(g->name, name)) { if (ptr) *ptr = (CUdeviceptr)(uintptr_t)g->mem; if (bytes) *bytes = g->bytes; return CUDA_SUCCESS; }
    const Elf64_Sym* s = find_symbol(m, name, NULL);
    if (!s || ELF64_ST_TYPE(s->st_info) != STT_OBJECT || s->st_shndx == SHN_UNDEF || s->st_shndx >= m->eh->e_shnum) return CUDA_ERROR_NOT_FOUND;
    const Elf64_Shdr* sec = &m->sh[s->st_shndx];
    struct fake_global* g = (struct fake_global*)calloc(1, sizeof *g);
    snprintf(g->name, sizeof g->name, "%s", name);
    g->bytes = s->st_size;
    g->mem = calloc(1, s->st_size ? s->st_size : 1);
    if (sec->sh_type != SHT_NOBITS && s->st_value + s->st_size <= sec->sh_size) memcpy(g->mem, m->image + sec->sh_offset + s->st_value, s->st_size);
    g->next = m->globals;
    m->globals = g;
    add_range(g->mem, g->bytes ? g->bytes : 1, 1);
    if (ptr) *ptr = (CUdeviceptr)(uintptr_t)g->mem;
    if (bytes) *bytes = g->bytes;
    return CUDA_SUCCESS;
}
CUresult cuFuncGetAttribute(int* v, CUfunction_attribute a, CUfunction fn) {
    fake_function* f = (fake_function*)fn;
    if (!f) return CUDA_ERROR_INVALID_HANDLE;
    switch (a) {
        case CU_FUNC_ATTRIBUTE_NUM_REGS: *v = f->regs; break;
        case CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK: *v = f->max_threads; break;
        default: *v = 0; break;
    }
    return CUDA_SUCCESS;
}
CUresult cuFuncSetAttribute(CUfunction fn, CUfunction_attribute a, int v) { (void)fn; (void)a; (void)v; return CUDA_SUCCESS; }
CUresult cuOccupancyMaxActiveBlocksPerMultiprocessor(int* blocks, CUfunction fn, int block_threads, size_t smem) {
    /* sm_100: 64 K registers per SM allocated per warp in units of 256, at most 64 warps and 32 blocks, 227 KB shared */
    fake_function* f = (fake_function*)fn;
    const int warps = (block_threads + 31) / 32;
    const int regs_per_warp = ((f->regs > 0 ? f->regs : 32) * 32 + 255) / 256 * 256;
    int by_regs = 65536 / regs_per_warp / warps;
    int by_warps = 64 / warps;
    int by_smem = smem > 0 ? (int)((227u * 1024u) / (smem + 1024u)) : 32;
    int b = by_regs < by_warps ? by_regs : by_warps;
    if (by_smem < b) b = by_smem;
    if (b > 32) b = 32;
    *blocks = b;
    return CUDA_SUCCESS;
}

/* ---- memory ---------------------------------------------------------------------------------------------------- */
static CUresult alloc_device(CUdeviceptr* p, size_t bytes) {
    if (t_depth <= 0) return CUDA_ERROR_INVALID_CONTEXT;
    if (bytes == 0) return CUDA_ERROR_INVALID_VALUE;
    pthread_mutex_lock(&g_lock);
    const long call = ++g_alloc_calls;
    const int fail = g_fail_alloc_at > 0 && call == g_fail_alloc_at;
    pthread_mutex_unlock(&g_lock);
    if (fail) return CUDA_ERROR_OUT_OF_MEMORY;
    void* m = NULL;
    if (posix_memalign(&m, 512, bytes) != 0) return CUDA_ERROR_OUT_OF_MEMORY;
    memset(m, 0, bytes);   /* deterministic "frames": nothing renders here */
    add_range(m, bytes, 1);
    pthread_mutex_lock(&g_lock);
    g_allocs++;
    pthread_mutex_unlock(&g_lock);
    *p = (CUdeviceptr)(uintptr_t)m;
    return CUDA_SUCCESS;
}
static CUresult free_device(CUdeviceptr p) {
    if (!drop_range((void*)(uintptr_t)p, 1)) return CUDA_ERROR_INVALID_VALUE;
    free((void*)(uintptr_t)p);
    pthread_mutex_lock(&g_lock);
    g_allocs--;
    pthread_mutex_unlock(&g_lock);
    return CUDA_SUCCESS;
}
CUresult cuMemAlloc_v2(CUdeviceptr* p, size_t bytes) { return alloc_device(p, bytes); }
CUresult cuMemFree_v2(CUdeviceptr p) { return free_device(p); }
CUresult cuMemAllocAsync(CUdeviceptr* p, size_t bytes, CUstream s) { (void)s; return alloc_device(p, bytes); }
CUresult cuMemFreeAsync(CUdeviceptr p, CUstream s) { (void)s; return free_device(p); }
CUresult cuMemcpyHtoD_v2(CUdeviceptr d, const void* h, size_t n) { memcpy((void*)(uintptr_t)d, h, n); return CUDA_SUCCESS; }
CUresult cuMemcpyDtoH_v2(void* h, CUdeviceptr d, size_t n) { memcpy(h, (const void*)(uintptr_t)d, n); return CUDA_SUCCESS; }
CUresult cuMemcpyDtoHAsync_v2(void* h, CUdeviceptr d, size_t n, CUstream s) { (void)s; memcpy(h, (const void*)(uintptr_t)d, n); return CUDA_SUCCESS; }
CUresult cuMemcpyHtoDAsync_v2(CUdeviceptr d, const void* h, size_t n, CUstream s) { (void)s; memcpy((void*)(uintptr_t)d, h, n); return CUDA_SUCCESS; }
CUresult cuMemsetD32_v2(CUdeviceptr d, unsigned v, size_t n) {
    unsigned* p = (unsigned*)(uintptr_t)d;
    for (size_t i = 0; i < n; ++i) p[i] = v;
    return CUDA_SUCCESS;
}
CUresult cuMemHostAlloc(void** p, size_t bytes, unsigned flags) {
    (void)flags;
    if (t_depth <= 0) return CUDA_ERROR_INVALID_CONTEXT;
    pthread_mutex_lock(&g_lock);
    const long call = ++g_alloc_calls;
    const int fail = g_fail_alloc_at > 0 && call == g_fail_alloc_at;
    pthread_mutex_unlock(&g_lock);
    if (fail || bytes == 0 || posix_memalign(p, 4096, bytes) != 0) return CUDA_ERROR_OUT_OF_MEMORY;
    memset(*p, 0, bytes);
    add_range(*p, bytes, 2);
    pthread_mutex_lock(&g_lock);
    g_allocs++;
    pthread_mutex_unlock(&g_lock);
    return CUDA_SUCCESS;
}
CUresult cuMemFreeHost(void* p) {
    if (!drop_range(p, 2)) return CUDA_ERROR_INVALID_VALUE;
    free(p);
    pthread_mutex_lock(&g_lock);
    g_allocs--;
    pthread_mutex_unlock(&g_lock);
    return CUDA_SUCCESS;
}
CUresult cuMemHostRegister_v2(void* p, size_t bytes, unsigned flags) {
    (void)flags;
    if (!p || bytes == 0) return CUDA_ERROR_INVALID_VALUE;
    if (range_kind((uintptr_t)p)) return CUDA_ERROR_HOST_MEMORY_ALREADY_REGISTERED;
    add_range(p, bytes, 2);
    return CUDA_SUCCESS;
}
CUresult cuMemHostUnregister(void* p) { return drop_range(p, 2) ? CUDA_SUCCESS : CUDA_ERROR_HOST_MEMORY_NOT_REGISTERED; }
CUresult cuMemHostGetDevicePointer_v2(CUdeviceptr* d, void* p, unsigned flags) {
    (void)flags;
    if (range_kind((uintptr_t)p) != 2) return CUDA_ERROR_INVALID_VALUE;
    *d = (CUdeviceptr)(uintptr_t)p;
    return CUDA_SUCCESS;
}
CUresult cuPointerGetAttribute(void* out, CUpointer_attribute a, CUdeviceptr p) {
    const int kind = range_kind((uintptr_t)p);
    if (!kind) return CUDA_ERROR_INVALID_VALUE;   /* pageable memory is unknown to the driver */
    if (a == CU_POINTER_ATTRIBUTE_MEMORY_TYPE) { *(unsigned*)out = kind == 1 ? CU_MEMORYTYPE_DEVICE : CU_MEMORYTYPE_HOST; return CUDA_SUCCESS; }
    if (a == CU_POINTER_ATTRIBUTE_DEVICE_POINTER) { *(CUdeviceptr*)out = p; return CUDA_SUCCESS; }
    return CUDA_ERROR_INVALID_VALUE;
}
CUresult cuIpcGetMemHandle(CUipcMemHandle* h, CUdeviceptr p) {
    if (range_kind((uintptr_t)p) != 1) return CUDA_ERROR_INVALID_VALUE;
    memset(h, 0, sizeof *h);
    memcpy(h->reserved, &p, sizeof p);
    return CUDA_SUCCESS;
}
CUresult cuIpcOpenMemHandle_v2(CUdeviceptr* p, CUipcMemHandle h, unsigned flags) {
    (void)flags;
    memcpy(p, h.reserved, sizeof *p);   /* same process only */
    return range_kind((uintptr_t)*p) == 1 ? CUDA_SUCCESS : CUDA_ERROR_INVALID_VALUE;
}
CUresult cuIpcCloseMemHandle(CUdeviceptr p) { (void)p; return CUDA_SUCCESS; }

/* ---- streams / events: everything executes at enqueue time ---------------------------------------------------- */
CUresult cuStreamCreate(CUstream* s, unsigned flags) {
    (void)flags;
    if (t_depth <= 0) return CUDA_ERROR_INVALID_CONTEXT;
    fake_stream* st = (fake_stream*)calloc(1, sizeof *st);
    st->device = current_device();
    *s = (CUstream)st;
    return CUDA_SUCCESS;
}
CUresult cuStreamDestroy_v2(CUstream s) { free(s); return CUDA_SUCCESS; }
CUresult cuStreamSynchronize(CUstream s) { (void)s; return t_depth > 0 ? CUDA_SUCCESS : CUDA_ERROR_INVALID_CONTEXT; }
CUresult cuEventCreate(CUevent* e, unsigned flags) { (void)flags; *e = (CUevent)calloc(1, sizeof(fake_event)); return CUDA_SUCCESS; }
CUresult cuEventDestroy_v2(CUevent e) { free(e); return CUDA_SUCCESS; }
CUresult cuEventRecord(CUevent e, CUstream s) {
    (void)s;
    fake_event* ev = (fake_event*)e;
    if (!ev) return CUDA_ERROR_INVALID_HANDLE;
    /* other threads may wait on / re-record the same event concurrently: as in the driver, that is allowed */
    pthread_mutex_lock(&g_lock);
    ev->ms = now_ms(); ev->recorded = 1;
    pthread_mutex_unlock(&g_lock);
    return CUDA_SUCCESS;
}
CUresult cuEventSynchronize(CUevent e) { return e ? CUDA_SUCCESS : CUDA_ERROR_INVALID_HANDLE; }
CUresult cuEventElapsedTime(float* ms, CUevent a, CUevent b) {
    fake_event *x = (fake_event*)a, *y = (fake_event*)b;
    if (!x || !y) return CUDA_ERROR_INVALID_HANDLE;
    pthread_mutex_lock(&g_lock);
    const int ok = x->recorded && y->recorded;
    *ms = (float)(y->ms - x->ms);
    pthread_mutex_unlock(&g_lock);
    return ok ? CUDA_SUCCESS : CUDA_ERROR_INVALID_HANDLE;
}
CUresult cuStreamWaitEvent(CUstream s, CUevent e, unsigned flags) { (void)s; (void)flags; return e ? CUDA_SUCCESS : CUDA_ERROR_INVALID_HANDLE; }

static CUresult write32(CUdeviceptr addr, uint32_t v) {
    if (!range_kind((uintptr_t)addr) || (addr & 3u)) return CUDA_ERROR_INVALID_VALUE;
    __atomic_store_n((uint32_t*)(uintptr_t)addr, v, __ATOMIC_RELEASE);
    return CUDA_SUCCESS;
}
static CUresult wait32(CUdeviceptr addr, uint32_t v, unsigned flags) {
    if (!range_kind((uintptr_t)addr) || (addr & 3u)) return CUDA_ERROR_INVALID_VALUE;
    const double t0 = now_ms();
    for (;;) {   /* another thread's stream may still have to write the word */
        const uint32_t have = __atomic_load_n((uint32_t*)(uintptr_t)addr, __ATOMIC_ACQUIRE);
        const int ok = (flags & 3u) == CU_STREAM_WAIT_VALUE_GEQ ? (int32_t)(have - v) >= 0 : (flags & 3u) == CU_STREAM_WAIT_VALUE_EQ ? have == v : (have & v) != 0;
        if (ok) return CUDA_SUCCESS;
        if (now_ms() - t0 > 5000.0) return CUDA_ERROR_LAUNCH_TIMEOUT;
    }
}
CUresult cuStreamWriteValue32_v2(CUstream s, CUdeviceptr addr, cuuint32_t v, unsigned flags) { (void)s; (void)flags; return write32(addr, v); }
CUresult cuStreamWaitValue32_v2(CUstream s, CUdeviceptr addr, cuuint32_t v, unsigned flags) { (void)s; return wait32(addr, v, flags); }
CUresult cuStreamBatchMemOp_v2(CUstream s, unsigned n, CUstreamBatchMemOpParams* ops, unsigned flags) {
    (void)s; (void)flags;
    for (unsigned i = 0; i < n; ++i) {
        CUresult r = CUDA_ERROR_INVALID_VALUE;
        if (ops[i].operation == CU_STREAM_MEM_OP_WAIT_VALUE_32) r = wait32(ops[i].waitValue.address, ops[i].waitValue.value, ops[i].waitValue.flags);
        else if (ops[i].operation == CU_STREAM_MEM_OP_WRITE_VALUE_32) r = write32(ops[i].writeValue.address, ops[i].writeValue.value);
        if (r != CUDA_SUCCESS) return r;
    }
    return CUDA_SUCCESS;
}

/* ---- launches: recorded, not executed -------------------------------------------------------------------------- */
CUresult cuLaunchKernel(CUfunction fn, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz, unsigned smem,
                        CUstream s, void** params, void** extra) {
    (void)extra;
    fake_function* f = (fake_function*)fn;
    if (!f || t_depth <= 0) return CUDA_ERROR_INVALID_HANDLE;
    if (gx == 0 || gy == 0 || gz == 0 || bx * by * bz == 0 || bx * by * bz > (unsigned)f->max_threads || gy > 65535u || gz > 65535u)
        return CUDA_ERROR_INVALID_VALUE;
    pthread_mutex_lock(&g_lock);
    const int refuse = g_fail_launch_count > 0 && g_fail_launch_device == current_device() && !strcmp(f->name, "sbx_render");
    if (refuse) g_fail_launch_count--;
    pthread_mutex_unlock(&g_lock);
    if (refuse) return CUDA_ERROR_LAUNCH_FAILED;
    fake_launch L;
    memset(&L, 0, sizeof L);
    snprintf(L.name, sizeof L.name, "%s", f->name);
    L.grid[0] = gx; L.grid[1] = gy; L.grid[2] = gz;
    L.block[0] = bx; L.block[1] = by; L.block[2] = bz;
    L.smem = smem;
    L.device = current_device();
    L.stream = s;
    L.n_params = f->n_params;
    for (int i = 0; i < f->n_params; ++i) {
        L.param_size[i] = f->param_size[i];
        memcpy(L.param[i], params[i], f->param_size[i] < FAKE_PARAM_BYTES ? f->param_size[i] : FAKE_PARAM_BYTES);
    }
    if (!strcmp(f->name, "sbx_render") && f->n_params >= 1 && f->param_size[0] == sizeof(sbx_launch)) {
        /* the one effect of a render launch the host protocol depends on: its last thread block publishes done_value */
        const sbx_launch* R = (const sbx_launch*)params[0];
        if (R->done_flag) __atomic_store_n(R->done_flag, R->done_value, __ATOMIC_RELEASE);
    }
    pthread_mutex_lock(&g_lock);
    if (g_log_len == g_log_cap) { g_log_cap = g_log_cap ? g_log_cap * 2 : 64; g_log = (fake_launch*)realloc(g_log, g_log_cap * sizeof *g_log); }
    g_log[g_log_len++] = L;
    pthread_mutex_unlock(&g_lock);
    return CUDA_SUCCESS;
}

/* ---- TMA descriptors: the driver's argument checks ------------------------------------------------------------- */
CUresult cuTensorMapEncodeTiled(CUtensorMap* map, CUtensorMapDataType type, cuuint32_t rank, void* base, const cuuint64_t* dims,
                                const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* elem_strides,
                                CUtensorMapInterleave il, CUtensorMapSwizzle sw, CUtensorMapL2promotion l2, CUtensorMapFloatOOBfill oob) {
    (void)l2; (void)oob;
    if (!map || ((uintptr_t)map & 63u) || rank < 1 || rank > 5 || !base || ((uintptr_t)base & 15u)) return CUDA_ERROR_INVALID_VALUE;
    if (type != CU_TENSOR_MAP_DATA_TYPE_FLOAT32 || il != CU_TENSOR_MAP_INTERLEAVE_NONE || sw != CU_TENSOR_MAP_SWIZZLE_NONE) return CUDA_ERROR_INVALID_VALUE;
    for (unsigned i = 0; i < rank; ++i) {
        if (dims[i] == 0 || dims[i] > (1ull << 32) || box[i] == 0 || box[i] > 256 || elem_strides[i] == 0 || elem_strides[i] > 8) return CUDA_ERROR_INVALID_VALUE;
        if (i + 1 < rank && (strides[i] == 0 || (strides[i] & 15u) || strides[i] >= (1ull << 40))) return CUDA_ERROR_INVALID_VALUE;
    }
    if ((box[0] * 4u) & 15u) return CUDA_ERROR_INVALID_VALUE;   /* the innermost box extent is a multiple of 16 bytes */
    memset(map, 0, sizeof *map);
    memcpy(map, &base, sizeof base);
    for (unsigned i = 0; i < rank; ++i) map->opaque[1 + i] = dims[i] | ((cuuint64_t)box[i] << 40);
    return CUDA_SUCCESS;
}
