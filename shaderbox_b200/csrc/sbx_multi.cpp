// sbx_multi.cpp -- one frame over the GPUs of a box, from ONE process, behind the C ABI (include/sbx.h sbx_multi_*).
//
// The host a shaderbox maintainer writes is a plain C++ program: it cannot be asked to start one
// Python process per GPU.  A group holds one sbx_ctx per GPU (primary contexts, peer access enabled towards the GPU
// that owns the frame), one stream per GPU and one worker thread per GPU beyond the first, so the N launches of a
// frame are issued concurrently instead of one after the other (a launch is ~4 us of host time: 8 of them in a row
// would delay the last GPU by a tenth of a 0.3 ms frame).  Every GPU renders its part of the frame -- 4-row stripes
// dealt round-robin, sbx_frame_part -- straight into the destination:
//   sbx_multi_render_device  the group's frame in the first GPU's HBM: peers store over NVLink, the first GPU's
//                            stream waits for every part with cross-device events (no collective)
//   sbx_multi_render_host    a pinned host frame (sbx_host_alloc): every GPU stores over its own PCIe link and its
//                            stream publishes a completion flag in pinned memory (cuStreamWriteValue32) that the
//                            caller polls; a pageable frame is rendered in HBM as above and copied
// Pixels are independent (src/main.h:6-53), so the assembled frame is bit-identical to a 1-GPU render.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "sbx_internal.h"

struct sbx_multi {
    sbx::driver_api* cu = nullptr;
    int n = 0;
    std::vector<int> devices;
    std::vector<sbx_ctx*> ctx;
    std::vector<CUstream> stream;
    std::vector<CUevent> done;            // recorded on stream[i] after GPU i's part
    CUevent start = nullptr;              // recorded on stream[0]: the other streams start a frame after it
    CUdeviceptr frame = 0;                // on devices[0]
    size_t frame_bytes = 0;
    unsigned* flags = nullptr;            // pinned + mapped: one completion word per GPU (host frames)
    unsigned epoch = 0;
    std::string last_error;
    std::vector<float> kernel_ms;
    float step_ms = 0.0f;

    // workers: GPU i >= 1 is driven by its own thread
    struct task { const sbx_params* p = nullptr; float* dest = nullptr; unsigned* flag = nullptr; unsigned value = 0; int to_host = 0; };
    task cur;
    std::vector<std::thread> workers;
    std::vector<int> status;
    std::atomic<unsigned> go{0}, finished{0};
    std::atomic<bool> quit{false};
    std::mutex m;
    std::condition_variable cv;
};

namespace {

struct scope {
    sbx::driver_api* cu;
    scope(sbx::driver_api* cu_, CUcontext c) : cu(cu_) { cu->CtxPushCurrent(c); }
    ~scope() { CUcontext old; cu->CtxPopCurrent(&old); }
};

int fail(sbx_multi* m, int st, const std::string& what) {
    m->last_error = what;
    return st;
}

// GPU i's share of the current task: wait for the frame to be free, render the part, mark it done
int run_part(sbx_multi* m, int i) {
    const sbx_multi::task& t = m->cur;
    sbx_frame_part part;
    std::memset(&part, 0, sizeof part);
    // 4-row stripes dealt round-robin (measured on B200: balanced to 1 %, and faster per GPU than a checkerboard of
    // tiles, which spreads a GPU's concurrent warps 8 tiles apart -- tools/part_time.py)
    part.rows.stripe_rows = 4; part.rows.n_parts = m->n; part.rows.part = i;
    part.tile_parts = 1;
    part.tile_part = 0;
    part.done_flag = nullptr;
    part.done_value = 0;
    scope s(m->cu, sbx::context_of(m->ctx[i]));
    if (i > 0 && m->cu->StreamWaitEvent(m->stream[i], m->start, 0) != CUDA_SUCCESS) return SBX_ERR_CUDA;
    int st = sbx_render_frame_part(m->ctx[i], t.p, &part, t.dest, m->stream[i]);
    if (st != SBX_OK) return st;
    // host frames: the stream publishes "this part has landed" in pinned memory behind the launch
    if (t.flag && (st = sbx_stream_write_flag(m->ctx[i], t.flag + i, t.value, m->stream[i])) != SBX_OK) return st;
    if (m->cu->EventRecord(m->done[i], m->stream[i]) != CUDA_SUCCESS) return SBX_ERR_CUDA;
    return SBX_OK;
}

void worker_main(sbx_multi* m, int i) {
    unsigned seen = 0;
    for (;;) {
        // spin briefly (frames arrive back to back in a render loop), then sleep
        const auto t0 = std::chrono::steady_clock::now();
        while (m->go.load(std::memory_order_acquire) == seen && !m->quit.load(std::memory_order_acquire)) {
            if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(500)) {
                std::unique_lock<std::mutex> lk(m->m);
                m->cv.wait_for(lk, std::chrono::milliseconds(50), [&] { return m->go.load() != seen || m->quit.load(); });
            }
        }
        if (m->quit.load(std::memory_order_acquire)) return;
        seen = m->go.load(std::memory_order_acquire);
        m->status[i] = run_part(m, i);
        m->finished.fetch_add(1, std::memory_order_release);
    }
}

// issue every GPU's part of the task; returns when all launches have been enqueued
int dispatch(sbx_multi* m, const sbx_multi::task& t) {
    m->cur = t;
    {
        scope s(m->cu, sbx::context_of(m->ctx[0]));
        if (m->cu->EventRecord(m->start, m->stream[0]) != CUDA_SUCCESS) return fail(m, SBX_ERR_CUDA, "cuEventRecord(start)");
    }
    m->finished.store(0, std::memory_order_relaxed);
    if (m->n > 1) {
        m->go.fetch_add(1, std::memory_order_release);
        { std::lock_guard<std::mutex> lk(m->m); }   // a worker between its predicate check and its wait holds m: no lost wake-up
        m->cv.notify_all();
    }
    m->status[0] = run_part(m, 0);
    while (m->finished.load(std::memory_order_acquire) != (unsigned)(m->n - 1)) std::this_thread::yield();
    for (int i = 0; i < m->n; ++i)
        if (m->status[i] != SBX_OK) return fail(m, m->status[i], std::string("GPU part ") + std::to_string(i) + ": " + sbx_last_error(m->ctx[i]));
    return SBX_OK;
}

int ensure_frame(sbx_multi* m, size_t bytes) {
    if (bytes <= m->frame_bytes) return SBX_OK;
    scope s(m->cu, sbx::context_of(m->ctx[0]));
    if (m->frame) { m->cu->StreamSynchronize(m->stream[0]); m->cu->MemFree(m->frame); m->frame = 0; m->frame_bytes = 0; }
    if (m->cu->MemAlloc(&m->frame, bytes) != CUDA_SUCCESS) return fail(m, SBX_ERR_NOMEM, "cuMemAlloc(frame) failed");
    m->frame_bytes = bytes;
    return SBX_OK;
}

}  // namespace

extern "C" {

int sbx_multi_create(const int* devices, int n_gpus, sbx_multi** out) {
    if (!out || n_gpus < 1 || n_gpus > 64) return SBX_ERR_INVALID;
    *out = nullptr;
    sbx_multi* m = new sbx_multi;
    m->n = n_gpus;
    for (int i = 0; i < n_gpus; ++i) m->devices.push_back(devices ? devices[i] : i);
    std::string err;
    m->cu = sbx::load_driver(&err);
    int st = m->cu ? SBX_OK : SBX_ERR_NO_DEVICE;
    for (int i = 0; st == SBX_OK && i < n_gpus; ++i) {
        sbx_ctx* c = nullptr;
        st = sbx_create(m->devices[i], &c);
        if (st == SBX_OK) m->ctx.push_back(c);
    }
    if (st != SBX_OK) {
        for (sbx_ctx* c : m->ctx) sbx_destroy(c);
        delete m;
        return st;
    }
    m->stream.assign(n_gpus, nullptr);
    m->done.assign(n_gpus, nullptr);
    m->status.assign(n_gpus, SBX_OK);
    m->kernel_ms.assign(n_gpus, 0.0f);
    for (int i = 0; i < n_gpus; ++i) {
        scope s(m->cu, sbx::context_of(m->ctx[i]));
        m->cu->StreamCreate(&m->stream[i], CU_STREAM_NON_BLOCKING);
        m->cu->EventCreate(&m->done[i], CU_EVENT_DISABLE_TIMING);
        if (i == 0) m->cu->EventCreate(&m->start, CU_EVENT_DISABLE_TIMING);
        // peers store into the frame that lives on devices[0]
        if (i > 0 && m->devices[i] != m->devices[0]) {
            const CUresult r = m->cu->CtxEnablePeerAccess(sbx::context_of(m->ctx[0]), 0);
            if (r != CUDA_SUCCESS && r != CUDA_ERROR_PEER_ACCESS_ALREADY_ENABLED) {
                sbx_multi_destroy(m);
                return SBX_ERR_UNSUPPORTED;   // no peer path between the two GPUs
            }
        }
    }
    {
        void* h = nullptr;
        if (sbx_host_alloc(m->ctx[0], 4096, &h) != SBX_OK) { sbx_multi_destroy(m); return SBX_ERR_NOMEM; }
        m->flags = (unsigned*)h;
        std::memset(m->flags, 0, 4096);
    }
    for (int i = 1; i < n_gpus; ++i) m->workers.emplace_back(worker_main, m, i);
    *out = m;
    return SBX_OK;
}

void sbx_multi_destroy(sbx_multi* m) {
    if (!m) return;
    m->quit.store(true);
    { std::lock_guard<std::mutex> lk(m->m); }
    m->cv.notify_all();
    for (auto& t : m->workers) t.join();
    for (int i = 0; i < (int)m->ctx.size(); ++i) {
        scope s(m->cu, sbx::context_of(m->ctx[i]));
        if (m->stream[i]) { m->cu->StreamSynchronize(m->stream[i]); m->cu->StreamDestroy(m->stream[i]); }
        if (m->done[i]) m->cu->EventDestroy(m->done[i]);
        if (i == 0) {
            if (m->start) m->cu->EventDestroy(m->start);
            if (m->frame) m->cu->MemFree(m->frame);
        }
    }
    if (m->flags) sbx_host_free(m->ctx[0], m->flags);
    for (sbx_ctx* c : m->ctx) sbx_destroy(c);
    delete m;
}

int sbx_multi_gpus(sbx_multi* m) { return m ? m->n : SBX_ERR_INVALID; }
sbx_ctx* sbx_multi_ctx(sbx_multi* m, int i) { return (m && i >= 0 && i < m->n) ? m->ctx[i] : nullptr; }
const char* sbx_multi_last_error(sbx_multi* m) { return m ? m->last_error.c_str() : ""; }

int sbx_multi_load_app(sbx_multi* m, const char* app_name, const char* variant) {
    if (!m) return SBX_ERR_INVALID;
    for (int i = 0; i < m->n; ++i) {
        const int st = sbx_load_app(m->ctx[i], app_name, variant);
        if (st != SBX_OK) return fail(m, st, sbx_last_error(m->ctx[i]));
    }
    return SBX_OK;
}

int sbx_multi_set_option(sbx_multi* m, const char* key, int value) {
    if (!m) return SBX_ERR_INVALID;
    for (int i = 0; i < m->n; ++i) {
        const int st = sbx_set_option(m->ctx[i], key, value);
        if (st != SBX_OK) return st;
    }
    return SBX_OK;
}

int sbx_multi_render_device(sbx_multi* m, const sbx_params* p, float** dev_frame_out) {
    if (!m || !p || p->width <= 0 || p->height <= 0) return SBX_ERR_INVALID;
    int st = ensure_frame(m, (size_t)p->width * p->height * 4 * sizeof(float));
    if (st != SBX_OK) return st;
    sbx_multi::task t;
    t.p = p;
    t.dest = (float*)(uintptr_t)m->frame;
    if ((st = dispatch(m, t)) != SBX_OK) return st;
    scope s(m->cu, sbx::context_of(m->ctx[0]));
    for (int i = 1; i < m->n; ++i)   // the first GPU's stream continues when every part has landed
        if (m->cu->StreamWaitEvent(m->stream[0], m->done[i], 0) != CUDA_SUCCESS) return fail(m, SBX_ERR_CUDA, "cuStreamWaitEvent");
    if (dev_frame_out) *dev_frame_out = (float*)(uintptr_t)m->frame;
    return SBX_OK;
}

void* sbx_multi_stream(sbx_multi* m) { return m ? (void*)m->stream[0] : nullptr; }

int sbx_multi_sync(sbx_multi* m) {
    if (!m) return SBX_ERR_INVALID;
    scope s(m->cu, sbx::context_of(m->ctx[0]));
    if (m->cu->StreamSynchronize(m->stream[0]) != CUDA_SUCCESS) return fail(m, SBX_ERR_CUDA, "cuStreamSynchronize");
    for (int i = 0; i < m->n; ++i) {
        sbx_timing tm;
        if (sbx_last_timing(m->ctx[i], &tm) == SBX_OK) m->kernel_ms[i] = tm.kernel_ms;
    }
    return SBX_OK;
}

int sbx_multi_render_host(sbx_multi* m, const sbx_params* p, float* host_rgba) {
    if (!m || !p || !host_rgba || p->width <= 0 || p->height <= 0) return SBX_ERR_INVALID;
    const size_t bytes = (size_t)p->width * p->height * 4 * sizeof(float);
    CUdeviceptr alias = 0;
    {
        scope s(m->cu, sbx::context_of(m->ctx[0]));
        unsigned mem_type = 0;
        if (((uintptr_t)host_rgba & 15u) == 0 &&
            m->cu->PointerGetAttribute(&mem_type, CU_POINTER_ATTRIBUTE_MEMORY_TYPE, (CUdeviceptr)(uintptr_t)host_rgba) == CUDA_SUCCESS &&
            mem_type == CU_MEMORYTYPE_HOST)
            m->cu->PointerGetAttribute(&alias, CU_POINTER_ATTRIBUTE_DEVICE_POINTER, (CUdeviceptr)(uintptr_t)host_rgba);
    }
    if (alias) {
        // pinned + mapped frame: every GPU stores its part over its own PCIe link, then its stream publishes the
        // frame number in pinned memory; the caller polls the N words
        sbx_multi::task t;
        t.p = p;
        t.dest = (float*)(uintptr_t)alias;
        t.flag = m->flags;     // portable pinned memory with unified addressing: the host pointer is the device pointer
        t.value = ++m->epoch;
        t.to_host = 1;
        const int st = dispatch(m, t);
        if (st != SBX_OK) return st;
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < m->n; ++i) {
            volatile unsigned* f = m->flags + i;
            unsigned spins = 0;
            while ((int)(*f - t.value) < 0) {
                if ((++spins & 0xfffu) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30))
                    return fail(m, SBX_ERR_CUDA, "GPU part " + std::to_string(i) + " did not complete within 30 s");
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        return SBX_OK;
    }
    float* dev = nullptr;
    int st = sbx_multi_render_device(m, p, &dev);
    if (st != SBX_OK) return st;
    scope s(m->cu, sbx::context_of(m->ctx[0]));
    if (m->cu->MemcpyDtoHAsync(host_rgba, (CUdeviceptr)(uintptr_t)dev, bytes, m->stream[0]) != CUDA_SUCCESS ||
        m->cu->StreamSynchronize(m->stream[0]) != CUDA_SUCCESS)
        return fail(m, SBX_ERR_CUDA, "frame read-back failed");
    return SBX_OK;
}

int sbx_multi_last_timing(sbx_multi* m, float* kernel_ms_per_gpu, int capacity) {
    if (!m || !kernel_ms_per_gpu || capacity < m->n) return SBX_ERR_INVALID;
    for (int i = 0; i < m->n; ++i) {
        sbx_timing tm;
        kernel_ms_per_gpu[i] = sbx_last_timing(m->ctx[i], &tm) == SBX_OK ? tm.kernel_ms : -1.0f;
    }
    return SBX_OK;
}

}  // extern "C"
