// Image preprocessing of the Python layers on the device:
//   im = zoom(im, (1, 1, h/H, w/W), order=1); im = im.transpose(0,2,3,1) + mean_pixel; np.round(im)
// (pylayers/pylayers/pylayers.py:70-75, :315-319), followed by the ubyte cast CRF() applies
// (CRF/krahenbuhl2013/CRF.py:32).  At the training shape scipy's zoom alone costs ~13 ms per batch on
// the host, an order of magnitude more than the whole GPU pass.
//
// scipy.ndimage.zoom(order=1, grid_mode=False) maps output index o to input coordinate
// o * (in-1)/(out-1) and interpolates linearly in float64; the restatement below (validated
// bit-for-bit against scipy in tests/test_oracle_golden.py) keeps scipy's operation order:
//   t = v00*(wy0*wx0) + v01*(wy0*wx1) + v10*(wy1*wx0) + v11*(wy1*wx1), all float64, no contraction,
// cast to float32 (the blob dtype), + mean in float64, round half to even, C cast to unsigned char.
#include "common.cuh"

namespace dsrg {

__global__ void __launch_bounds__(kThreads)
k_prepare_image(const float *in, uint8_t *out, int Hi, int Wi, int Ho, int Wo, double m0, double m1,
                double m2) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ho * Wo) return;
    const int oy = i / Wo, ox = i - oy * Wo;
    const double zy = Ho > 1 ? (double)(Hi - 1) / (double)(Ho - 1) : 0.0;
    const double zx = Wo > 1 ? (double)(Wi - 1) / (double)(Wo - 1) : 0.0;
    const double ys = __dmul_rn((double)oy, zy), xs = __dmul_rn((double)ox, zx);
    const int y0 = (int)floor(ys), x0 = (int)floor(xs);
    const double fy = __dsub_rn(ys, (double)y0), fx = __dsub_rn(xs, (double)x0);
    const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
    const double wy0 = __dsub_rn(1.0, fy), wx0 = __dsub_rn(1.0, fx);
    const double w00 = __dmul_rn(wy0, wx0), w01 = __dmul_rn(wy0, fx), w10 = __dmul_rn(fy, wx0),
                 w11 = __dmul_rn(fy, fx);
    const double mean[3] = {m0, m1, m2};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float *p = in + ((size_t)b * 3 + c) * Hi * Wi;
        double t = __dmul_rn((double)p[(size_t)y0 * Wi + x0], w00);
        t = __dadd_rn(t, __dmul_rn((double)p[(size_t)y0 * Wi + x1], w01));
        t = __dadd_rn(t, __dmul_rn((double)p[(size_t)y1 * Wi + x0], w10));
        t = __dadd_rn(t, __dmul_rn((double)p[(size_t)y1 * Wi + x1], w11));
        const float z = (float)t;                              // zoom returns the input dtype (float32)
        const double r = rint(__dadd_rn((double)z, mean[c]));  // + mean_pixel (float64), np.round
        out[((size_t)b * Ho * Wo + i) * 3 + c] = (uint8_t)(long long)r;  // .astype('ubyte')
    }
}

int prepare_image(Engine *e, int B, int Hi, int Wi, const float *in, const double *mean, uint8_t *out,
                  cudaStream_t s) {
    dim3 g(cdiv((long long)e->H * e->W, kThreads), B);
    DSRG_LAUNCH(e, T_PREP, s,
                k_prepare_image<<<g, kThreads, 0, s>>>(in, out, Hi, Wi, e->H, e->W, mean[0], mean[1], mean[2]));
    DSRG_CUDA_TRY(cudaGetLastError());
    return DSRG_OK;
}

}  // namespace dsrg

using namespace dsrg;

extern "C" int dsrg_prepare_image_dev(dsrg_engine *h, int B, int Hi, int Wi, const float *images_dev,
                                      const double *mean_pixel, uint8_t *image_out_dev, void *stream) {
    Engine *e = (Engine *)h;
    int rc = check_batch(e, B);
    if (rc) return rc;
    if (!images_dev || !mean_pixel || !image_out_dev || Hi < 1 || Wi < 1) {
        set_error("bad argument");
        return DSRG_E_INVALID;
    }
    return prepare_image(e, B, Hi, Wi, images_dev, mean_pixel, image_out_dev, (cudaStream_t)stream);
}

extern "C" int dsrg_prepare_image_host(dsrg_engine *h, int B, int Hi, int Wi, const float *images,
                                       const double *mean_pixel, uint8_t *image_out) {
    Engine *e = (Engine *)h;
    int rc = check_batch(e, B);
    if (rc) return rc;
    if (!images || !mean_pixel || !image_out || Hi < 1 || Wi < 1) {
        set_error("bad argument");
        return DSRG_E_INVALID;
    }
    if ((rc = ensure_staging(e))) return rc;
    const size_t need = (size_t)B * 3 * Hi * Wi;
    if (need > e->st_raw_cap) {  // raw images come in any size: grow on demand (not on the hot CRF path)
        cudaFree(e->st_raw);
        e->st_raw = nullptr;
        e->st_raw_cap = 0;
        if ((rc = dalloc(e, &e->st_raw, need))) return rc;
        e->st_raw_cap = need;
    }
    cudaStream_t s = e->own_stream;
    DSRG_CUDA_TRY(cudaMemcpyAsync(e->st_raw, images, need * sizeof(float), cudaMemcpyHostToDevice, s));
    if ((rc = prepare_image(e, B, Hi, Wi, e->st_raw, mean_pixel, e->st_image, s))) return rc;
    DSRG_CUDA_TRY(cudaMemcpyAsync(image_out, e->st_image, (size_t)B * e->N * 3, cudaMemcpyDeviceToHost, s));
    DSRG_CUDA_TRY(cudaStreamSynchronize(s));
    return DSRG_OK;
}
