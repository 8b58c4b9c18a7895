// scipy.ndimage.zoom(order=1, mode='constant', grid_mode=False) for one output pixel, operation by
// operation (scipy ni_interpolation.c NI_ZoomShift + ni_splines.c; oracle/crf_oracle.py
// zoom_order1_restated is the numpy twin, checked bit-for-bit against scipy in tests/):
//   * output index o -> input coordinate c = o * ((in-1)/(out-1)), float64
//   * c > in-1 (one rounding error at the last row / column of some size pairs) is out of bounds and the
//     output pixel is cval = 0
//   * weights w0 = 1 - (c - floor(c)), w1 = 1 - w0
//   * value = ((v00*wy0)*wx0 + (v01*wy0)*wx1) + (v10*wy1)*wx0 + (v11*wy1)*wx1, summed left to right in
//     float64 without contraction, then cast to float32 (the array dtype)
#pragma once
#include <cuda_runtime.h>

namespace dsrg {

struct ZoomTap {
    int y0, y1, x0, x1;
    double wy0, wy1, wx0, wx1;
    bool oob;
};

__device__ __forceinline__ ZoomTap zoom_tap(int oy, int ox, int Hi, int Wi, int Ho, int Wo) {
    ZoomTap z;
    const double zy = Ho > 1 ? (double)(Hi - 1) / (double)(Ho - 1) : 0.0;
    const double zx = Wo > 1 ? (double)(Wi - 1) / (double)(Wo - 1) : 0.0;
    const double ys = __dmul_rn((double)oy, zy), xs = __dmul_rn((double)ox, zx);
    z.oob = ys > (double)(Hi - 1) || xs > (double)(Wi - 1);
    z.y0 = min((int)floor(ys), Hi - 1);
    z.x0 = min((int)floor(xs), Wi - 1);
    z.y1 = min(z.y0 + 1, Hi - 1);
    z.x1 = min(z.x0 + 1, Wi - 1);
    z.wy0 = __dsub_rn(1.0, __dsub_rn(ys, (double)z.y0));
    z.wx0 = __dsub_rn(1.0, __dsub_rn(xs, (double)z.x0));
    z.wy1 = __dsub_rn(1.0, z.wy0);
    z.wx1 = __dsub_rn(1.0, z.wx0);
    return z;
}

// p: one [Hi][Wi] float32 plane
__device__ __forceinline__ float zoom_apply(const float *__restrict__ p, int Wi, const ZoomTap &z) {
    if (z.oob) return 0.0f;
    double t = __dmul_rn(__dmul_rn((double)p[(size_t)z.y0 * Wi + z.x0], z.wy0), z.wx0);
    t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)p[(size_t)z.y0 * Wi + z.x1], z.wy0), z.wx1));
    t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)p[(size_t)z.y1 * Wi + z.x0], z.wy1), z.wx0));
    t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)p[(size_t)z.y1 * Wi + z.x1], z.wy1), z.wx1));
    return (float)t;
}

}  // namespace dsrg
