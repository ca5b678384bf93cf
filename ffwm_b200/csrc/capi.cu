// C-ABI plumbing shared by every entry point: argument validation, error text.
#include <stdarg.h>
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
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
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

}  // namespace ffwm

extern "C" int ffwm_abi_version(void) { return FFWM_ABI_VERSION; }
extern "C" const char* ffwm_last_error(void) { return ffwm::g_err; }
extern "C" unsigned long long ffwm_kernel_launches(void) { return __atomic_load_n(&ffwm::g_launches, __ATOMIC_RELAXED); }
