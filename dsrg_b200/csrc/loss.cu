// Balanced seeding loss (sm_100a); replaces BalancedSeedLossLayer's Theano graphs
// (pylayers/pylayers/pylayers.py:126-152).
//   S_bg(n) = sum lab_bg log p_bg, cnt_bg(n) = sum lab_bg   (channel 0)
//   S_fg(n), cnt_fg(n) likewise over channels 1..M-1
//   loss = -mean_n S_bg/max(cnt_bg,1e-4) - mean_n S_fg/max(cnt_fg,1e-4)
// Per-image sums are accumulated in float64 (Theano's float32 reduction order is unspecified).
#include "common.cuh"

namespace dsrg {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// grid (chunks, B); acc[b][0..3] = S_bg, cnt_bg, S_fg, cnt_fg
__global__ void __launch_bounds__(kThreads)
k_seedloss_partial(const float *probs, const float *seeds, double *acc, int M, int N) {
    const int b = blockIdx.y;
    double s_bg = 0, c_bg = 0, s_fg = 0, c_fg = 0;
    const size_t base = (size_t)b * M * N;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (long long)M * N;
         t += (long long)gridDim.x * blockDim.x) {
        const float lab = seeds[base + t];
        if (lab != 0.0f) {
            const double term = (double)lab * (double)logf(probs[base + t]);
            if (t < N) {
                s_bg += term;
                c_bg += lab;
            } else {
                s_fg += term;
                c_fg += lab;
            }
        }
    }
    __shared__ double sh[4][kThreads / 32];
    s_bg = warp_sum(s_bg);
    c_bg = warp_sum(c_bg);
    s_fg = warp_sum(s_fg);
    c_fg = warp_sum(c_fg);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        sh[0][w] = s_bg;
        sh[1][w] = c_bg;
        sh[2][w] = s_fg;
        sh[3][w] = c_fg;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double v = 0;
        for (int k = 0; k < kThreads / 32; k++) v += sh[threadIdx.x][k];
        if (v != 0.0) atomicAdd(acc + (size_t)b * 4 + threadIdx.x, v);
    }
}

__global__ void k_seedloss_final(const double *acc, float *terms, int B) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double t0 = 0, t1 = 0;
        for (int b = 0; b < B; b++) {
            t0 += acc[b * 4 + 0] / fmax(acc[b * 4 + 1], (double)kMinProb);
            t1 += acc[b * 4 + 2] / fmax(acc[b * 4 + 3], (double)kMinProb);
        }
        terms[0] = (float)t0;
        terms[1] = (float)t1;
    }
}

// counts only (for backward)
__global__ void __launch_bounds__(kThreads)
k_seedloss_counts(const float *seeds, double *acc, int M, int N) {
    const int b = blockIdx.y;
    double c_bg = 0, c_fg = 0;
    const size_t base = (size_t)b * M * N;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (long long)M * N;
         t += (long long)gridDim.x * blockDim.x) {
        const float lab = seeds[base + t];
        if (t < N) c_bg += lab; else c_fg += lab;
    }
    c_bg = warp_sum(c_bg);
    c_fg = warp_sum(c_fg);
    if ((threadIdx.x & 31) == 0) {
        if (c_bg != 0.0) atomicAdd(acc + (size_t)b * 4 + 1, c_bg);
        if (c_fg != 0.0) atomicAdd(acc + (size_t)b * 4 + 3, c_fg);
    }
}

__global__ void __launch_bounds__(kThreads)
k_seedloss_grad(const float *probs, const float *seeds, const double *acc, float scale, float *grad,
                int M, int N) {
    const int b = blockIdx.y;
    const size_t base = (size_t)b * M * N;
    const float cnt_bg = fmaxf((float)acc[b * 4 + 1], kMinProb);
    const float cnt_fg = fmaxf((float)acc[b * 4 + 3], kMinProb);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (long long)M * N;
         t += (long long)gridDim.x * blockDim.x) {
        const float lab = seeds[base + t];
        const float cnt = t < N ? cnt_bg : cnt_fg;
        grad[base + t] = (lab != 0.0f) ? -scale * lab / (probs[base + t] * cnt) : 0.0f;
    }
}

int seedloss_forward(Engine *e, int B, const float *probs, const float *seeds, float *terms_out,
                     cudaStream_t s) {
    DSRG_CUDA_TRY(cudaMemsetAsync(e->loss_acc, 0, sizeof(double) * 4 * B, s));
    dim3 g(cdiv((long long)e->M * e->N, kThreads * 8), B);
    DSRG_LAUNCH(e, T_LOSS, s, k_seedloss_partial<<<g, kThreads, 0, s>>>(probs, seeds, e->loss_acc, e->M, e->N));
    DSRG_LAUNCH(e, T_LOSS, s, k_seedloss_final<<<1, 32, 0, s>>>(e->loss_acc, terms_out, B));
    DSRG_CUDA_TRY(cudaGetLastError());
    return DSRG_OK;
}

int seedloss_backward(Engine *e, int B, int n_global, const float *probs, const float *seeds,
                      float top_diff, float *grad, cudaStream_t s) {
    DSRG_CUDA_TRY(cudaMemsetAsync(e->loss_acc, 0, sizeof(double) * 4 * B, s));
    dim3 g(cdiv((long long)e->M * e->N, kThreads * 8), B);
    DSRG_LAUNCH(e, T_LOSS, s, k_seedloss_counts<<<g, kThreads, 0, s>>>(seeds, e->loss_acc, e->M, e->N));
    DSRG_LAUNCH(e, T_LOSS, s,
                k_seedloss_grad<<<g, kThreads, 0, s>>>(probs, seeds, e->loss_acc, top_diff / (float)n_global, grad,
                                                       e->M, e->N));
    DSRG_CUDA_TRY(cudaGetLastError());
    return DSRG_OK;
}

}  // namespace dsrg
