// Tile-local lattice views (sm_100a).
//
// The permutohedral splat / slice of the reference touch, per pixel, d+1 lattice rows of 24
// floats each (CRF/src/permutohedral.cpp:545-553, :574-584).  Neighbouring pixels share most of
// their vertices, so a 32x8-pixel tile only touches a few dozen DISTINCT rows.  This kernel
// finds them once per lattice build: per tile the list of distinct rows (tl_rows), per pixel the
// index into that list (tl_loc), and the transposed incidence as a CSR grouped by local vertex
// (tl_ptr / tl_ent), which lets the mean-field kernel splat without shared-memory float atomics
// (those are CAS loops on this architecture).  It also folds the symmetric normalisation into
// the barycentric weights: wn = bary * norm (pairwise.cpp:66,79).
#include "common.cuh"

namespace dsrg {

template <int DP1, int MAXLOC>
__global__ void __launch_bounds__(256)
k_tile_build(const int32_t *off, const float *bary, const float *norm, int32_t *tl_nloc,
             int32_t *tl_rows, uint16_t *tl_ptr, uint16_t *tl_ent, uint16_t *tl_loc, float *wn, int N,
             int W, int H, int tiles_x, int ntiles) {
    constexpr int HS = 2048;  // >= 256*DP1 distinct rows in the worst case, power of two
    static_assert(MAXLOC <= 256, "one scan element per thread");
    __shared__ int hkey[HS];
    __shared__ int hlv[HS];
    __shared__ int rows_s[MAXLOC];
    __shared__ int cnt[MAXLOC];
    __shared__ int ptr[MAXLOC + 1];
    __shared__ int wsum[8];
    __shared__ int count;

    const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int x = tx * kTileW + (tid & 31), y = ty * kTileH + (tid >> 5);
    const bool in = x < W && y < H;
    const int pix = y * W + x;
    for (int i = tid; i < HS; i += 256) hkey[i] = -1;
    if (tid < MAXLOC) cnt[tid] = 0;
    if (tid == 0) count = 0;
    __syncthreads();

    int slot[DP1];
    if (in) {
        const float nrm = norm[(size_t)b * N + pix];
#pragma unroll
        for (int r = 0; r < DP1; r++) {
            const size_t at = ((size_t)b * DP1 + r) * N + pix;
            const int row = off[at];
            wn[at] = __fmul_rn(bary[at], nrm);
            unsigned s = ((unsigned)row * 2654435761u) >> 21;  // 11 bits
            while (true) {
                int old = atomicCAS(&hkey[s], -1, row);
                if (old == -1) {
                    int lv = atomicAdd(&count, 1);
                    hlv[s] = lv;
                    if (lv < MAXLOC) rows_s[lv] = row;
                    break;
                }
                if (old == row) break;
                s = (s + 1) & (HS - 1);
            }
            slot[r] = (int)s;
        }
    }
    __syncthreads();
    const int nloc = count;
    tl_nloc += (size_t)b * ntiles;
    if (nloc > MAXLOC) {  // too many distinct vertices for the shared-memory path: fallback tile
        if (tid == 0) tl_nloc[tile] = -1;
        return;
    }
    int lv[DP1];
    if (in) {
#pragma unroll
        for (int r = 0; r < DP1; r++) {
            lv[r] = hlv[slot[r]];
            tl_loc[((size_t)b * DP1 + r) * N + pix] = (uint16_t)lv[r];
            atomicAdd(&cnt[lv[r]], 1);
        }
    }
    __syncthreads();
    // exclusive scan of cnt[0..MAXLOC) -> ptr
    {
        const int v = tid < MAXLOC ? cnt[tid] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += n;
        }
        if ((tid & 31) == 31) wsum[tid >> 5] = incl;
        __syncthreads();
        int base = 0;
        for (int w = 0; w < (tid >> 5); w++) base += wsum[w];
        if (tid < MAXLOC) {
            ptr[tid + 1] = base + incl;
            cnt[tid] = 0;  // reused as the fill cursor
        }
        if (tid == 0) ptr[0] = 0;
    }
    __syncthreads();
    uint16_t *ent = tl_ent + ((size_t)b * ntiles + tile) * (256 * DP1);
    if (in) {
#pragma unroll
        for (int r = 0; r < DP1; r++) {
            int pos = ptr[lv[r]] + atomicAdd(&cnt[lv[r]], 1);
            ent[pos] = (uint16_t)((tid << 3) | r);
        }
    }
    uint16_t *optr = tl_ptr + ((size_t)b * ntiles + tile) * (MAXLOC + 1);
    int32_t *orow = tl_rows + ((size_t)b * ntiles + tile) * MAXLOC;
    for (int i = tid; i <= nloc; i += 256) optr[i] = (uint16_t)ptr[i];
    for (int i = tid; i < nloc; i += 256) orow[i] = rows_s[i];
    if (tid == 0) tl_nloc[tile] = nloc;
}

int tiles_build(Engine *e, Lattice &L, int nb, cudaStream_t s) {
    dim3 g(e->ntiles, nb);
    if (L.d == 2) {
        DSRG_LAUNCH(e, T_LAT_MISC, s,
                    k_tile_build<3, kMaxLocSp><<<g, 256, 0, s>>>(L.off, L.bary, L.norm, L.tl_nloc, L.tl_rows,
                                                                 L.tl_ptr, L.tl_ent, L.tl_loc, L.wn, L.N, e->W,
                                                                 e->H, e->tiles_x, e->ntiles));
    } else {
        DSRG_LAUNCH(e, T_LAT_MISC, s,
                    k_tile_build<6, kMaxLocBi><<<g, 256, 0, s>>>(L.off, L.bary, L.norm, L.tl_nloc, L.tl_rows,
                                                                 L.tl_ptr, L.tl_ent, L.tl_loc, L.wn, L.N, e->W,
                                                                 e->H, e->tiles_x, e->ntiles));
    }
    DSRG_CUDA_TRY(cudaGetLastError());
    return DSRG_OK;
}

}  // namespace dsrg
