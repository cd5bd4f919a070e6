// common.cuh -- error plumbing and small helpers shared by the translation units of libastsph_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/astro_sph_b200.h"

namespace ast {

void set_error(const char *fmt, ...);          // defined in capi_common.cu (thread-local message)

#define AST_CUDA_TRY(expr)                                                                             \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            ast::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return AST_ECUDA;                                                                          \
        }                                                                                              \
    } while (0)

#define AST_REQUIRE(cond, ...)          \
    do {                                \
        if (!(cond)) {                  \
            ast::set_error(__VA_ARGS__); \
            return AST_EINVAL;          \
        }                               \
    } while (0)

// environment switches are read once; callers keep the result in a function-local `static const` (thread-safe in C++11)
inline bool env_flag(const char *name, bool dflt)
{
    const char *e = getenv(name);
    if (!e || !e[0]) return dflt;
    return e[0] != '0';
}
inline int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
}

// AST_DEBUG_SYNC=1 in the environment: synchronise after every kernel and name the one that faulted
inline bool debug_sync_enabled()
{
    static const bool v = env_flag("AST_DEBUG_SYNC", false);
    return v;
}
#define AST_KERNEL_CHECK(stream, name)                                                              \
    do {                                                                                            \
        cudaError_t _e = cudaGetLastError();                                                        \
        if (_e == cudaSuccess && ast::debug_sync_enabled()) _e = cudaStreamSynchronize(stream);     \
        if (_e != cudaSuccess) {                                                                    \
            ast::set_error("%s:%d: kernel %s failed: %s", __FILE__, __LINE__, name, cudaGetErrorString(_e)); \
            return AST_ECUDA;                                                                       \
        }                                                                                           \
    } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct Carver {
    char *base;
    size_t off;
    explicit Carver(void *p) : base((char *)p), off(0) {}
    template <class T>
    T *take(size_t count)
    {
        off = align_up(off, 256);
        T *r = (T *)(base ? base + off : nullptr);
        off += count * sizeof(T);
        return r;
    }
    size_t bytes() const { return align_up(off, 256); }
};

inline int ceil_log2_u64(uint64_t v)
{
    int b = 0;
    while (b < 63 && (1ull << b) < v) ++b;
    return b;
}

// optional per-stage timing with CUDA events on the launching stream
struct StageTimer {
    bool on;
    cudaStream_t s;
    static constexpr int kMax = 256;
    cudaEvent_t ev[kMax][2];
    int stage[kMax];
    int n;
    StageTimer(bool enable, cudaStream_t st) : on(enable), s(st), n(0) {}
    void begin(int st)
    {
        if (!on || n >= kMax) return;
        cudaEventCreate(&ev[n][0]);
        cudaEventCreate(&ev[n][1]);
        stage[n] = st;
        cudaEventRecord(ev[n][0], s);
    }
    void end()
    {
        if (!on || n >= kMax) return;
        cudaEventRecord(ev[n][1], s);
        ++n;
    }
    void collect(float *stage_ms, int n_stage)
    {
        for (int i = 0; i < n_stage; ++i) stage_ms[i] = 0.f;
        if (!on) return;
        cudaStreamSynchronize(s);
        for (int i = 0; i < n; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i][0], ev[i][1]);
            if (stage[i] < n_stage) stage_ms[stage[i]] += ms;
            cudaEventDestroy(ev[i][0]);
            cudaEventDestroy(ev[i][1]);
        }
        n = 0;
    }
};

}  // namespace ast
