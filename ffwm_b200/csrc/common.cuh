// Shared device/host helpers for the ffwm_b200 warp kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ffwm_b200.h"

namespace ffwm {

// A (N,C,H,W) view as the kernels see it: 64-bit batch/channel strides for
// the plane base, 32-bit row/column strides inside one plane (checked on the
// host: a plane never spans 2^31 elements).
template <typename T>
struct View {
    T* p;
    int64_t sb, sc;   // batch, channel stride (elements)
    int sh, sw;       // row, column stride (elements)
    int n, c, h, w;
    __device__ __forceinline__ T* plane(int b, int ch) const { return p + b * sb + ch * sc; }
};

// Host-side error text for ffwm_last_error().
void set_error(const char* fmt, ...);

// Validates an ffwm_tensor4 and converts it; returns FFWM_OK or an error code.
template <typename T>
int make_view(const ffwm_tensor4* t, const char* name, View<T>* out, bool allow_null_data = false);

int check_launch(const char* what);

// float -> int exactly like the reference's int(floor(v)): cvt.rzi (saturating, NaN -> 0).
__device__ __forceinline__ int f2i(float v) { return __float2int_rz(v); }
__device__ __forceinline__ int f2i(double v) { return __double2int_rz(v); }

__device__ __forceinline__ int clampi(int v, int hi) { return max(min(v, hi), 0); }

// The reference's SAFE_DIV (resample2d_kernel.cu:14-15): 1e-8 is a double
// literal, so the whole expression is a double whatever T is.
template <typename T>
__device__ __forceinline__ double safe_div(T a, T b) {
    return (b == T(0)) ? (double(a) / 1e-8) : double(T(a / b));
}

// Gaussian tap weight, exp evaluated in double and rounded to T.
template <typename T>
__device__ __forceinline__ T gauss(T d, T sigma) {
    return T(exp(safe_div<T>(-d * d, T(2) * sigma * sigma)));
}

// Streaming accesses: operands that are touched once should not linger in L1.
template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) { return __ldcs(p); }
template <typename T>
__device__ __forceinline__ void st_stream(T* p, T v) { __stcs(p, v); }

// Fire-and-forget scatter add (RED.E.ADD): no return value is consumed.
__device__ __forceinline__ void red_add(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(double* p, double v) { atomicAdd(p, v); }

inline int ceil_div(int64_t a, int64_t b) { return int((a + b - 1) / b); }

// Number of SMs of the current device (cached); grids are sized against it.
int sm_count();

// Runtime options (diagnostics / A-B switches).  Each is read ONCE from the environment variable FFWM_<NAME> when
// the library is first used and can be changed afterwards through ffwm_set_option() (tests do that); launch paths
// only read a cached int.
enum Opt {
    OPT_DISABLE_TILED = 0,       // every warp op through the direct kernels
    OPT_FORCE_TILED,             // small / ragged shapes through the tiled kernels (tests)
    OPT_DISABLE_ROLL,
    OPT_FORCE_ROLL,
    OPT_SCATTER_TILED,           // previous-generation destination-sorted scatter
    OPT_DISABLE_TILED_GFLOW,
    OPT_DISABLE_QUAD,
    OPT_GQ_SCALAR_FILL,
    OPT_SCATTER_SCALAR_FLUSH,
    OPT_WGRAD_NO_ROWS,           // conv_wgrad: never use the row mode (one tile per kernel row shared by its taps)
    OPT_BN_NO_SMALL,             // batch_norm: never the one-block-per-channel kernels for small maps
    OPT_CONV_OCC2,               // conv_forward: 0 two CTAs per SM for small grids (automatic), 1 never, 2 always
    OPT_WGRAD_CHAIN,             // conv_wgrad: longest accumulation chain per CTA in 64-pixel stages (0 = the default, 64)
    OPT_COUNT
};
int opt(int id);

}  // namespace ffwm
