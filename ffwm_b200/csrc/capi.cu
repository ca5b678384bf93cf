// C-ABI plumbing shared by every entry point: argument validation, error text.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace ffwm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

template <typename T>
int make_view(const ffwm_tensor4* t, const char* name, View<T>* out, bool allow_null_data) {
    if (!t) {
        set_error("%s: null tensor descriptor", name);
        return FFWM_ERR_NULL;
    }
    int64_t numel = 1;
    for (int i = 0; i < 4; ++i) {
        if (t->size[i] < 0 || t->size[i] > 0x7fffffffLL) {
            set_error("%s: size[%d]=%lld out of range", name, i, (long long)t->size[i]);
            return FFWM_ERR_SHAPE;
        }
        numel *= t->size[i];
    }
    if (!t->data && numel > 0 && !allow_null_data) {
        set_error("%s: null data pointer", name);
        return FFWM_ERR_NULL;
    }
    // in-plane offsets are 32-bit in the kernels
    int64_t span = 0;
    if (numel > 0) {
        int64_t sh = t->stride[2] < 0 ? -t->stride[2] : t->stride[2];
        int64_t sw = t->stride[3] < 0 ? -t->stride[3] : t->stride[3];
        span = (t->size[2] - 1) * sh + (t->size[3] - 1) * sw;
    }
    if (span >= 0x7fffffffLL) {
        set_error("%s: one (H,W) plane spans %lld elements (limit 2^31-1)", name, (long long)span);
        return FFWM_ERR_TOO_LARGE;
    }
    out->p = static_cast<T*>(t->data);
    out->sb = t->stride[0];
    out->sc = t->stride[1];
    out->sh = int(t->stride[2]);
    out->sw = int(t->stride[3]);
    out->n = int(t->size[0]);
    out->c = int(t->size[1]);
    out->h = int(t->size[2]);
    out->w = int(t->size[3]);
    return FFWM_OK;
}

template int make_view<float>(const ffwm_tensor4*, const char*, View<float>*, bool);
template int make_view<double>(const ffwm_tensor4*, const char*, View<double>*, bool);
template int make_view<const float>(const ffwm_tensor4*, const char*, View<const float>*, bool);
template int make_view<const double>(const ffwm_tensor4*, const char*, View<const double>*, bool);

static unsigned long long g_launches = 0;   // kernels enqueued by this process (diagnostic, not synchronised)

int check_launch(const char* what) {
    __atomic_add_fetch(&g_launches, 1ULL, __ATOMIC_RELAXED);
    // Peek, do not consume: a sticky/asynchronous error raised earlier by someone else's kernel on this context stays
    // visible to its owner (PyTorch) instead of being swallowed and mis-attributed here.
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed or the context already carried an error: %s", what, cudaGetErrorString(e));
        return int(e);
    }
    return FFWM_OK;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

static const char* const g_opt_names[OPT_COUNT] = {
    "DISABLE_TILED", "FORCE_TILED", "DISABLE_ROLL", "FORCE_ROLL", "SCATTER_TILED", "DISABLE_TILED_GFLOW",
    "DISABLE_QUAD", "GQ_SCALAR_FILL", "SCATTER_SCALAR_FLUSH", "WGRAD_NO_ROWS", "BN_NO_SMALL", "CONV_OCC2", "WGRAD_CHAIN"};
static int g_opts[OPT_COUNT];
static int g_opts_ready = 0;

static void init_opts() {
    if (__atomic_load_n(&g_opts_ready, __ATOMIC_ACQUIRE)) return;
    for (int i = 0; i < OPT_COUNT; ++i) {
        char env[64];
        snprintf(env, sizeof(env), "FFWM_%s", g_opt_names[i]);
        const char* v = getenv(env);
        int val = 0;                                                   // defaults: everything off
        if (v && *v) val = (v[0] >= '0' && v[0] <= '9') ? atoi(v) : 1; // FFWM_X=1, FFWM_X=anything -> 1, FFWM_X=0 -> 0
        g_opts[i] = val;
    }
    __atomic_store_n(&g_opts_ready, 1, __ATOMIC_RELEASE);
}

int opt(int id) {
    init_opts();
    return g_opts[id];
}

static int opt_index(const char* name) {
    if (!name) return -1;
    if (!strncmp(name, "FFWM_", 5)) name += 5;
    for (int i = 0; i < OPT_COUNT; ++i)
        if (!strcmp(name, g_opt_names[i])) return i;
    return -1;
}

}  // namespace ffwm

extern "C" int ffwm_set_option(const char* name, int value) {
    const int i = ffwm::opt_index(name);
    if (i < 0) { ffwm::set_error("ffwm_set_option: unknown option '%s'", name ? name : "(null)"); return FFWM_ERR_ARG; }
    ffwm::init_opts();
    ffwm::g_opts[i] = value;
    return FFWM_OK;
}
extern "C" int ffwm_get_option(const char* name) {
    const int i = ffwm::opt_index(name);
    return i < 0 ? -1 : ffwm::opt(i);
}

extern "C" int ffwm_abi_version(void) { return FFWM_ABI_VERSION; }
extern "C" const char* ffwm_last_error(void) { return ffwm::g_err; }
extern "C" unsigned long long ffwm_kernel_launches(void) { return __atomic_load_n(&ffwm::g_launches, __ATOMIC_RELAXED); }
