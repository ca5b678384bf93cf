// Guided filter (models/external_function.py:164-195,239-277: BoxFilter + GuidedFilter.forward) and its gradient
// with respect to the filtered image x, as four small kernels per direction.
//
// STATUS: default on since round 2 (FFWM_FUSED_GF=0 in ffwm_b200/external_function.py for the A/B).  The hand-derived
// backward is checked on the CPU in float64 against autograd of the reference formula (tests/test_guided_filter_math.py)
// and the kernels against the same on a B200 (tests/test_guided_filter_gpu.py).
//
// Why: the reference builds every box filter from two cumsums, three slices and a cat per axis, seven box filters
// per call, three calls per train step plus their autograd mirror — several hundred tiny launches per step
// (r01m launch list: scan_outer_dim 1.3 ms, scan_innermost, index_select/gather 0.7 ms, plus the elementwise glue).
// The maps are tiny (8 x 3 x 128 x 128 at most: 1.5 MB, L2 resident), so the box sums are taken directly — a
// separable truncated window sum, one thread per element walking its 2r+1 window — which is also more accurate
// than the difference of two running sums.
//
//   forward   N = window size; mx = box(x)/N, my = box(y)/N, cov = box(xy)/N - mx my, var = box(xx)/N - mx^2,
//             A = cov / (var + eps), b = my - A mx, q = (box(A)/N) x + box(b)/N
//   backward  (x only; y is data)  with g = dL/dq and box self-adjoint:
//             pA = box(g x / N), pb = box(g / N); gA = pA - pb mx; gcov = gA / ve; gvar = -gA A / ve  (ve = var + eps)
//             gmx = -pb A - gcov my - 2 mx gvar
//             dL/dx = g mA + y box(gcov / N) + 2 x box(gvar / N) + box(gmx / N)
//
// Layout: contiguous fp32 (planes, H, W), planes = B*C; x and y with the same channel count.
#include <stdint.h>

#include "common.cuh"

namespace ffwm {

struct GfGeo {
    int64_t total;   // planes * h * w
    int h, w, r;
};

__device__ __forceinline__ int gf_cnt(int i, int n, int r) { return min(i + r, n - 1) - max(i - r, 0) + 1; }

// window sums along W of K derived planes; F::load(e) returns the K values of source element e
template <int K, typename F>
__global__ void __launch_bounds__(256) gf_row_kernel(F f, float* __restrict__ out, GfGeo g) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= g.total) return;
    const int j = (int)(e % g.w);
    const int lo = max(j - g.r, 0), hi = min(j + g.r, g.w - 1);
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
    for (int jj = lo; jj <= hi; ++jj) {
        float v[K];
        f.load(e - j + jj, v);
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += v[k];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) out[k * g.total + e] = acc[k];
}

// window sums along H of the K planes in `in`, handed to F::store(e, sums, 1/N-able count)
template <int K, typename F>
__global__ void __launch_bounds__(256) gf_col_kernel(const float* __restrict__ in, F f, GfGeo g) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= g.total) return;
    const int j = (int)(e % g.w), i = (int)((e / g.w) % g.h);
    const int lo = max(i - g.r, 0), hi = min(i + g.r, g.h - 1);
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
    for (int ii = lo; ii <= hi; ++ii) {
        const int64_t s = e + (int64_t)(ii - i) * g.w;
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += in[k * g.total + s];
    }
    f.store(e, acc, float(gf_cnt(i, g.h, g.r) * gf_cnt(j, g.w, g.r)));
}

// ---- forward functors
struct GfLoadXY {       // x, y, x*y, x*x
    const float *x, *y;
    __device__ void load(int64_t e, float* v) const {
        const float a = x[e], b = y[e];
        v[0] = a, v[1] = b, v[2] = a * b, v[3] = a * a;
    }
};
struct GfStoreCoef {    // -> A, b; saves mx, my, var + eps
    float *A, *b, *mx, *my, *ve;
    float eps;
    __device__ void store(int64_t e, const float* s, float n) const {
        const float m_x = s[0] / n, m_y = s[1] / n;
        const float cov = s[2] / n - m_x * m_y, var = s[3] / n - m_x * m_x;
        const float a = cov / (var + eps);
        A[e] = a, b[e] = m_y - a * m_x, mx[e] = m_x, my[e] = m_y, ve[e] = var + eps;
    }
};
struct GfLoad2 {        // two planes as they are
    const float *p0, *p1;
    __device__ void load(int64_t e, float* v) const { v[0] = p0[e], v[1] = p1[e]; }
};
struct GfStoreOut {     // q = mA x + mb; saves mA
    const float* x;
    float *q, *mA;
    __device__ void store(int64_t e, const float* s, float n) const {
        const float a = s[0] / n;
        mA[e] = a;
        q[e] = a * x[e] + s[1] / n;
    }
};

// ---- backward functors
struct GfLoadG {        // g x / N, g / N   (N of the SOURCE element)
    const float *g, *x;
    int h, w, r;
    __device__ void load(int64_t e, float* v) const {
        const int j = (int)(e % w), i = (int)((e / w) % h);
        const float n = float(gf_cnt(i, h, r) * gf_cnt(j, w, r));
        const float gg = g[e];
        v[0] = gg * x[e] / n, v[1] = gg / n;
    }
};
struct GfStoreGCoef {   // pA, pb -> gcov / N, gvar / N, gmx / N
    const float *A, *mx, *my, *ve;
    float *u0, *u1, *u2;
    __device__ void store(int64_t e, const float* s, float n) const {
        const float a = A[e], m_x = mx[e], v = ve[e];
        const float gA = s[0] - s[1] * m_x;
        const float gcov = gA / v, gvar = -gA * a / v;
        const float gmx = -s[1] * a - gcov * my[e] - 2.f * m_x * gvar;
        u0[e] = gcov / n, u1[e] = gvar / n, u2[e] = gmx / n;
    }
};
struct GfLoad3 {
    const float *p0, *p1, *p2;
    __device__ void load(int64_t e, float* v) const { v[0] = p0[e], v[1] = p1[e], v[2] = p2[e]; }
};
struct GfStoreGx {      // dL/dx = g mA + y box(gcov/N) + 2 x box(gvar/N) + box(gmx/N)
    const float *g, *x, *y, *mA;
    float* gx;
    __device__ void store(int64_t e, const float* s, float) const {
        gx[e] = g[e] * mA[e] + y[e] * s[0] + 2.f * x[e] * s[1] + s[2];
    }
};

static int gf_check(const char* what, int64_t planes, int h, int w, int r, GfGeo* g) {
    if (planes < 0 || h < 0 || w < 0 || r < 0) { set_error("%s: negative size", what); return FFWM_ERR_SHAPE; }
    // the reference asserts h > 2r+1 and w > 2r+1 (external_function.py:253)
    if (planes > 0 && (h <= 2 * r + 1 || w <= 2 * r + 1)) { set_error("%s: needs H, W > 2r+1 (H %d, W %d, r %d)", what, h, w, r); return FFWM_ERR_ARG; }
    g->total = planes * (int64_t)h * w;
    g->h = h, g->w = w, g->r = r;
    if (g->total > (int64_t)0x7fffffff * 256) { set_error("%s: too large", what); return FFWM_ERR_TOO_LARGE; }
    return FFWM_OK;
}

}  // namespace ffwm

// q = GuidedFilter(r, eps)(x, y) (models/external_function.py:239-277); x, y, q: (planes, H, W) contiguous fp32.
// save: 5 * planes*H*W floats (mean_x, mean_y, var_x + eps, A, mean_A) for the backward pass.
// scratch: 5 * planes*H*W floats.
extern "C" int ffwm_guided_filter_forward(const float* x, const float* y, float* q, float* save, float* scratch,
                                          int64_t planes, int h, int w, int r, float eps, void* stream) {
    using namespace ffwm;
    GfGeo g;
    int rc;
    if ((rc = gf_check("guided_filter_forward", planes, h, w, r, &g))) return rc;
    if (g.total == 0) return FFWM_OK;
    if (!x || !y || !q || !save || !scratch) { set_error("guided_filter_forward: null pointer"); return FFWM_ERR_NULL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned blocks = (unsigned)((g.total + 255) / 256);
    const int64_t n = g.total;
    float *mx = save, *my = save + n, *ve = save + 2 * n, *A = save + 3 * n, *mA = save + 4 * n;
    float *T = scratch, *b = scratch + 4 * n;
    gf_row_kernel<4><<<blocks, 256, 0, st>>>(GfLoadXY{x, y}, T, g);
    if ((rc = check_launch("guided_filter_forward(rows xy)"))) return rc;
    gf_col_kernel<4><<<blocks, 256, 0, st>>>(T, GfStoreCoef{A, b, mx, my, ve, eps}, g);
    if ((rc = check_launch("guided_filter_forward(coefficients)"))) return rc;
    gf_row_kernel<2><<<blocks, 256, 0, st>>>(GfLoad2{A, b}, T, g);
    if ((rc = check_launch("guided_filter_forward(rows Ab)"))) return rc;
    gf_col_kernel<2><<<blocks, 256, 0, st>>>(T, GfStoreOut{x, q, mA}, g);
    return check_launch("guided_filter_forward(output)");
}

// grad_x = dL/dx for grad_q = dL/dq; save as written by the forward call; scratch: 6 * planes*H*W floats.
extern "C" int ffwm_guided_filter_backward(const float* x, const float* y, const float* grad_q, const float* save,
                                           float* grad_x, float* scratch, int64_t planes, int h, int w, int r, void* stream) {
    using namespace ffwm;
    GfGeo g;
    int rc;
    if ((rc = gf_check("guided_filter_backward", planes, h, w, r, &g))) return rc;
    if (g.total == 0) return FFWM_OK;
    if (!x || !y || !grad_q || !save || !grad_x || !scratch) { set_error("guided_filter_backward: null pointer"); return FFWM_ERR_NULL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned blocks = (unsigned)((g.total + 255) / 256);
    const int64_t n = g.total;
    const float *mx = save, *my = save + n, *ve = save + 2 * n, *A = save + 3 * n, *mA = save + 4 * n;
    float *S = scratch, *U = scratch + 3 * n;
    gf_row_kernel<2><<<blocks, 256, 0, st>>>(GfLoadG{grad_q, x, h, w, r}, S, g);
    if ((rc = check_launch("guided_filter_backward(rows g)"))) return rc;
    gf_col_kernel<2><<<blocks, 256, 0, st>>>(S, GfStoreGCoef{A, mx, my, ve, U, U + n, U + 2 * n}, g);
    if ((rc = check_launch("guided_filter_backward(coefficients)"))) return rc;
    gf_row_kernel<3><<<blocks, 256, 0, st>>>(GfLoad3{U, U + n, U + 2 * n}, S, g);
    if ((rc = check_launch("guided_filter_backward(rows u)"))) return rc;
    gf_col_kernel<3><<<blocks, 256, 0, st>>>(S, GfStoreGx{grad_q, x, y, mA, grad_x}, g);
    return check_launch("guided_filter_backward(grad x)");
}
