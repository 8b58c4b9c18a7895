// Mean-field inference of the fully connected CRF (sm_100a).
//
// Replaces DenseCRF::inference (CRF/src/densecrf.cpp:115-131) with its per-iteration calls
// PairwisePotential::apply -> DenseKernel::filter -> Permutohedral::sseCompute ->
// PottsCompatibility::apply (CRF/src/pairwise.cpp:173-178, :63-80; permutohedral.cpp:529-589;
// labelcompatibility.cpp:46-48) and expAndNormalize (densecrf.cpp:98-106), for the whole batch.
//
// State is planar: U, Q are [B][M][N] float32 (one coalesced load per label per thread);
// lattice value rows are [rows][MP] float32 with MP = M rounded up to a multiple of 4
// (the reference pads 21 -> 24 the same way, permutohedral.cpp:531).
#include "common.cuh"

namespace dsrg {

// ---------------------------------------------------------------------------------------------
// init: unary (any layout) -> U planar; Q = softmax(U)   (densecrf.cpp:120 with energy = -unary)
// ---------------------------------------------------------------------------------------------
template <int MP>
__global__ void __launch_bounds__(kThreads)
k_mf_init(const float *unary, float *unary_rw, int layout, int clamp, float *U, float *Q, int M,
          int N) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float u[MP];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < MP; k++) {
        if (k < M) {
            size_t at = (layout == DSRG_LAYOUT_NCHW) ? ((size_t)b * M + k) * N + i
                                                     : ((size_t)b * N + i) * M + k;
            float v = unary[at];
            if (clamp && v < kMinProb) {  // probs[probs < min_prob] = min_prob, pylayers.py:312
                v = kMinProb;
                unary_rw[at] = v;
            }
            u[k] = v;
            mx = fmaxf(mx, v);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < MP; k++)
        if (k < M) {
            U[((size_t)b * M + k) * N + i] = u[k];
            u[k] = expf(u[k] - mx);
            s += u[k];
        }
#pragma unroll
    for (int k = 0; k < MP; k++)
        if (k < M) Q[((size_t)b * M + k) * N + i] = u[k] / s;
}

// ---------------------------------------------------------------------------------------------
// splat: values[row] += bary * (Q * norm) for both lattices (pairwise.cpp:66, permutohedral.cpp
// :545-553).  One thread per pixel, scattered float atomics (first correct version).
// ---------------------------------------------------------------------------------------------
struct LatView {
    const int32_t *off;   // [nimg][dp1][N]
    const float *bary;    // [nimg][dp1][N]
    const float *norm;    // [nimg][N]
    const int32_t *rowbase;
    float *val;           // values being splatted into / sliced from
    int dp1;
    int shared;
};

template <int MP>
__device__ __forceinline__ void splat_one(const LatView &L, int b, int i, int N, int M,
                                          const float *q) {
    const int sb = L.shared ? 0 : b;
    const float nrm = L.norm[(size_t)sb * N + i];
    const int base = L.rowbase[b];
    for (int r = 0; r < L.dp1; r++) {
        size_t at = ((size_t)sb * L.dp1 + r) * N + i;
        float *row = L.val + (size_t)(base + L.off[at]) * MP;
        const float w = L.bary[at];
#pragma unroll
        for (int k = 0; k < MP; k++)
            if (k < M) atomicAdd(row + k, __fmul_rn(w, __fmul_rn(q[k], nrm)));
    }
}

template <int MP>
__global__ void __launch_bounds__(kThreads)
k_mf_splat(const float *Q, LatView sp, LatView bi, int M, int N) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float q[MP];
#pragma unroll
    for (int k = 0; k < MP; k++)
        if (k < M) q[k] = Q[((size_t)b * M + k) * N + i];
    splat_one<MP>(sp, b, i, N, M, q);
    splat_one<MP>(bi, b, i, N, M, q);
}

// zero the splat targets of both lattices (row counts are device-resident)
template <int MP>
__global__ void __launch_bounds__(kThreads)
k_mf_zero(float4 *a, const int32_t *rowbase_a, float4 *c, const int32_t *rowbase_c, int B) {
    constexpr int CH = MP / 4;
    const long long na = (long long)rowbase_a[B] * CH, nc = (long long)rowbase_c[B] * CH;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < na + nc;
         t += (long long)gridDim.x * blockDim.x) {
        if (t < na) a[t] = z; else c[t - na] = z;
    }
}

// ---------------------------------------------------------------------------------------------
// blur along one lattice axis: new = old + 0.5 (old[n1] + old[n2])  (permutohedral.cpp:556-569)
// one thread per (row, float4 chunk)
// ---------------------------------------------------------------------------------------------
template <int MP>
__global__ void __launch_bounds__(kThreads)
k_mf_blur(const float4 *in, float4 *out, const int2 *nbr, const int32_t *rowbase, int B, int shared) {
    constexpr int CH = MP / 4;
    const long long rows = rowbase[B];
    const int rows_img = shared ? rowbase[1] : 0;
    const long long total = rows * CH;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long g = t / CH;
        const int c = (int)(t - g * CH);
        long long n1, n2;
        if (shared) {
            const long long img0 = (g / rows_img) * rows_img;
            int2 n = nbr[g - img0];
            n1 = img0 + n.x;
            n2 = img0 + n.y;
        } else {
            int2 n = nbr[g];
            n1 = n.x;
            n2 = n.y;
        }
        const float4 o = in[g * CH + c], a = in[n1 * CH + c], d = in[n2 * CH + c];
        float4 r;
        r.x = o.x + 0.5f * (a.x + d.x);
        r.y = o.y + 0.5f * (a.y + d.y);
        r.z = o.z + 0.5f * (a.z + d.z);
        r.w = o.w + 0.5f * (a.w + d.w);
        out[g * CH + c] = r;
    }
}

// ---------------------------------------------------------------------------------------------
// slice both lattices, apply norm + Potts, add to the unary and renormalise:
//   t = U + w2 * norm_sp * slice_sp + w1 * norm_bi * slice_bi ;  Q = softmax(t)
// (permutohedral.cpp:574-584, pairwise.cpp:79, labelcompatibility.cpp:46-48,
//  densecrf.cpp:123-128: tmp1 = -unary; tmp1 -= (-w * K Q) for the Gaussian, then the bilateral).
// ---------------------------------------------------------------------------------------------
template <int MP>
__device__ __forceinline__ void slice_one(const LatView &L, int b, int i, int N, float alpha,
                                          float *acc) {
    const int sb = L.shared ? 0 : b;
    const int base = L.rowbase[b];
#pragma unroll
    for (int k = 0; k < MP; k++) acc[k] = 0.0f;
    for (int r = 0; r < L.dp1; r++) {
        size_t at = ((size_t)sb * L.dp1 + r) * N + i;
        const float4 *row = reinterpret_cast<const float4 *>(L.val + (size_t)(base + L.off[at]) * MP);
        const float w = __fmul_rn(L.bary[at], alpha);  // bary * alpha (:579)
#pragma unroll
        for (int c = 0; c < MP / 4; c++) {
            float4 v = row[c];
            acc[4 * c + 0] += w * v.x;
            acc[4 * c + 1] += w * v.y;
            acc[4 * c + 2] += w * v.z;
            acc[4 * c + 3] += w * v.w;
        }
    }
}

template <int MP>
__global__ void __launch_bounds__(kThreads)
k_mf_slice_update(const float *U, float *Qout, LatView sp, LatView bi, float w_sp, float w_bi,
                  float alpha_sp, float alpha_bi, int M, int N) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float t[MP], acc[MP];
#pragma unroll
    for (int k = 0; k < MP; k++) t[k] = (k < M) ? U[((size_t)b * M + k) * N + i] : 0.0f;
    {
        slice_one<MP>(sp, b, i, N, alpha_sp, acc);
        const float nrm = sp.norm[i];
#pragma unroll
        for (int k = 0; k < MP; k++) t[k] += w_sp * (acc[k] * nrm);
    }
    {
        slice_one<MP>(bi, b, i, N, alpha_bi, acc);
        const float nrm = bi.norm[(size_t)b * N + i];
#pragma unroll
        for (int k = 0; k < MP; k++) t[k] += w_bi * (acc[k] * nrm);
    }
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < MP; k++)
        if (k < M) mx = fmaxf(mx, t[k]);
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < MP; k++)
        if (k < M) {
            t[k] = expf(t[k] - mx);
            s += t[k];
        }
#pragma unroll
    for (int k = 0; k < MP; k++)
        if (k < M) Qout[((size_t)b * M + k) * N + i] = t[k] / s;
}

// ---------------------------------------------------------------------------------------------
// exports
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_mf_export(const float *Q, float *out, int layout, int M, int N) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    for (int k = 0; k < M; k++) {
        float v = Q[((size_t)b * M + k) * N + i];
        size_t at = (layout == DSRG_LAYOUT_NCHW) ? ((size_t)b * M + k) * N + i
                                                 : ((size_t)b * N + i) * M + k;
        out[at] = v;
    }
}

// DenseCRF::currentMap (densecrf.cpp:202-211): first maximum wins
__global__ void __launch_bounds__(kThreads)
k_mf_export_map(const float *Q, int32_t *labels, int M, int N) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int m = 0;
    float best = Q[((size_t)b * M) * N + i];
    for (int k = 1; k < M; k++) {
        float v = Q[((size_t)b * M + k) * N + i];
        if (v > best) {
            best = v;
            m = k;
        }
    }
    labels[(size_t)b * N + i] = m;
}

// result[result < min_prob] = min_prob; result /= sum (float64), pylayers.py:328-330 / :85-86;
// optional log (CRFLayer top, pylayers.py:88)
__global__ void __launch_bounds__(kThreads)
k_mf_export_renorm(const float *Q, float *result_out, float *log_out, int M, int N) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double s = 0.0;
    for (int k = 0; k < M; k++) {
        double v = (double)Q[((size_t)b * M + k) * N + i];
        if (v < 0.0001) v = 0.0001;
        s += v;
    }
    for (int k = 0; k < M; k++) {
        size_t at = ((size_t)b * M + k) * N + i;
        double v = (double)Q[at];
        if (v < 0.0001) v = 0.0001;
        v = v / s;
        if (result_out) result_out[at] = (float)v;
        if (log_out) log_out[at] = (float)log(v);
    }
}

static LatView make_view(const Lattice &L, float *val) {
    LatView v;
    v.off = L.off;
    v.bary = L.bary;
    v.norm = L.norm;
    v.rowbase = L.rowbase;
    v.val = val;
    v.dp1 = L.d + 1;
    v.shared = L.shared;
    return v;
}

template <int MP>
static int run_impl(Engine *e, int B, const float *unary, int layout, bool clamp, float *unary_rw,
                    const dsrg_crf_params &p, cudaStream_t s) {
    const int M = e->M, N = e->N;
    dim3 gp(cdiv(N, kThreads), B);
    DSRG_LAUNCH(e, T_MF_INIT, s, k_mf_init<MP><<<gp, kThreads, 0, s>>>(unary, unary_rw, layout, clamp ? 1 : 0, e->U, e->Q0, M, N));
    float *Qc = e->Q0, *Qn = e->Q1;
    const float alpha_sp = 1.0f / (1 + powf(2, -e->sp.d));  // permutohedral.cpp:571
    const float alpha_bi = 1.0f / (1 + powf(2, -e->bi.d));
    const int blur_grid = 8 * e->sm_count;
    for (int it = 0; it < p.n_iters; it++) {
        // the row counts live on the device only (no host sync on the path)
        DSRG_LAUNCH(e, T_MF_ZERO, s,
                    k_mf_zero<MP><<<blur_grid, kThreads, 0, s>>>((float4 *)e->spA, e->sp.rowbase, (float4 *)e->biA,
                                                                 e->bi.rowbase, B));
        DSRG_LAUNCH(e, T_MF_SPLAT, s,
                    k_mf_splat<MP><<<gp, kThreads, 0, s>>>(Qc, make_view(e->sp, e->spA), make_view(e->bi, e->biA), M, N));
        float *src = e->spA, *dst = e->spB;
        for (int j = 0; j <= e->sp.d; j++) {
            DSRG_LAUNCH(e, T_MF_BLUR_SP, s,
                        k_mf_blur<MP><<<blur_grid, kThreads, 0, s>>>((const float4 *)src, (float4 *)dst,
                                                                     e->sp.nbr + (size_t)j * e->sp.nbr_stride,
                                                                     e->sp.rowbase, B, 1));
            float *t = src; src = dst; dst = t;
        }
        float *sp_final = src;
        src = e->biA; dst = e->biB;
        for (int j = 0; j <= e->bi.d; j++) {
            DSRG_LAUNCH(e, T_MF_BLUR_BI, s,
                        k_mf_blur<MP><<<blur_grid, kThreads, 0, s>>>((const float4 *)src, (float4 *)dst,
                                                                     e->bi.nbr + (size_t)j * e->bi.nbr_stride,
                                                                     e->bi.rowbase, B, 0));
            float *t = src; src = dst; dst = t;
        }
        float *bi_final = src;
        DSRG_LAUNCH(e, T_MF_SLICE, s,
                    k_mf_slice_update<MP><<<gp, kThreads, 0, s>>>(e->U, Qn, make_view(e->sp, sp_final),
                                                                  make_view(e->bi, bi_final), p.w2, p.w1,
                                                                  alpha_sp, alpha_bi, M, N));
        float *t = Qc; Qc = Qn; Qn = t;
    }
    e->Qcur = Qc;
    DSRG_CUDA_TRY(cudaGetLastError());
    return DSRG_OK;
}

int meanfield_run(Engine *e, int B, const float *unary, int unary_layout, bool clamp_inplace,
                  float *unary_rw, const dsrg_crf_params &p, cudaStream_t s) {
    switch (e->MP) {
        case 4: return run_impl<4>(e, B, unary, unary_layout, clamp_inplace, unary_rw, p, s);
        case 8: return run_impl<8>(e, B, unary, unary_layout, clamp_inplace, unary_rw, p, s);
        case 12: return run_impl<12>(e, B, unary, unary_layout, clamp_inplace, unary_rw, p, s);
        case 16: return run_impl<16>(e, B, unary, unary_layout, clamp_inplace, unary_rw, p, s);
        case 20: return run_impl<20>(e, B, unary, unary_layout, clamp_inplace, unary_rw, p, s);
        case 24: return run_impl<24>(e, B, unary, unary_layout, clamp_inplace, unary_rw, p, s);
        case 28: return run_impl<28>(e, B, unary, unary_layout, clamp_inplace, unary_rw, p, s);
        case 32: return run_impl<32>(e, B, unary, unary_layout, clamp_inplace, unary_rw, p, s);
    }
    set_error("unsupported label count %d", e->M);
    return DSRG_E_INVALID;
}

int meanfield_export(Engine *e, int B, float *out, int layout, cudaStream_t s) {
    dim3 gp(cdiv(e->N, kThreads), B);
    DSRG_LAUNCH(e, T_MF_EXPORT, s, k_mf_export<<<gp, kThreads, 0, s>>>(e->Qcur, out, layout, e->M, e->N));
    DSRG_CUDA_TRY(cudaGetLastError());
    return DSRG_OK;
}

int meanfield_export_map(Engine *e, int B, int32_t *labels, cudaStream_t s) {
    dim3 gp(cdiv(e->N, kThreads), B);
    DSRG_LAUNCH(e, T_MF_EXPORT, s, k_mf_export_map<<<gp, kThreads, 0, s>>>(e->Qcur, labels, e->M, e->N));
    DSRG_CUDA_TRY(cudaGetLastError());
    return DSRG_OK;
}

int meanfield_export_renorm(Engine *e, int B, float *result_out, float *log_out, cudaStream_t s) {
    dim3 gp(cdiv(e->N, kThreads), B);
    DSRG_LAUNCH(e, T_MF_EXPORT, s, k_mf_export_renorm<<<gp, kThreads, 0, s>>>(e->Qcur, result_out, log_out, e->M, e->N));
    DSRG_CUDA_TRY(cudaGetLastError());
    return DSRG_OK;
}

}  // namespace dsrg
