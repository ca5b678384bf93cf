/*
 * ORACLE (test infrastructure, not product code; see oracle/README.md).
 *
 * Plain-C restatement of the reference's CUDA kernels K1-K7 and of the ATen
 * grid sampler behind WarpNet, instantiated for float and double.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load the library built from this file.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -ffp-contract=off -shared)
 */
#include <math.h>
#include <stdint.h>

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* float->int the way the device does it (cvt.rzi.s32: NaN -> 0, saturating),
 * so that wild flows hit the same clamped taps as the reference kernels. */
static inline int f2i(double v) {
    if (v != v) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (-2147483647 - 1);
    return (int)v;
}

#define REAL float
#define SUFFIX _f32
#define FLOOR floorf
#include "warp_ops.inc"
#undef REAL
#undef SUFFIX
#undef FLOOR

#define REAL double
#define SUFFIX _f64
#define FLOOR floor
#include "warp_ops.inc"
#undef REAL
#undef SUFFIX
#undef FLOOR

int oracle_abi_version(void) { return 1; }
