// Track B, 2-D: leapfrog acoustic step with fused source injection / receiver sampling, forward-field
// snapshotting and adjoint imaging, for sm_100a.
//
//   w_n     = lap(u_n) + f_n
//   u_{n+1} = g * (2 u_n - g * u_{n-1} + m * w_n)           (oracle/fd_oracle.py, spec B1-B3)
//
// Data layout: every field is [nz][px] fp32, x contiguous, px = nx rounded up to 32 floats so each row
// starts on a 128-byte line; no ghost cells in memory.  The (BX+8) x (BZ+8) input tile of u_n is staged
// into shared memory by ONE TMA tiled load per CTA (cp.async.bulk.tensor.2d); the hardware zero-fills
// the part of the box that falls outside the grid, which is the Dirichlet halo.  Each lane owns four
// consecutive x points (one float4), reads its x-neighbours as two more float4 from the tile, and keeps
// the nine z-neighbours in a register window that rotates down the column, so shared memory is read
// ~4 x 16 B per float4 of output.  u_{n-1} and m are touched once per point straight from global
// (coalesced float4) and u_{n+1} overwrites u_{n-1} in place: 16 B of algorithmic traffic per update.
// The forward pass can stream w_n to an HBM snapshot (evict-first stores); the adjoint pass reads it
// back and accumulates the zero-lag cross-correlation in the same kernel.
#include "fd_common.cuh"
#include <vector>
#include <algorithm>
#include <cmath>
#include <cstring>

namespace fwi {

enum { STEP_FWD = 0, STEP_FWD_SAVE = 1, STEP_ADJ = 2, STEP_ADJ2 = 3 };   // ADJ2: image this step AND the previous one (deferred)
}  // namespace fwi
#include "fd3d.cuh"
#include "fd2d_tb2.cuh"
namespace fwi {

struct Step2DArgs {
    float* oldnew;        // u_{n-1} in, u_{n+1} out (in place)
    const float* m;
    const float* gx;
    const float* gz;
    float* snap;          // w_n: written (FWD_SAVE) or read (ADJ)
    const float* snap_prev; // ADJ2: snapshot paired with the field of the previous adjoint step (= u_n of this launch)
    float* acc;           // imaging accumulator (ADJ)
    int nx, nz, px;
    PointListDev inj;     // sources (forward) / receivers (adjoint)
    const float* inj_vals;
    PointListDev rec;     // receivers to sample (forward only; tile_ptr == nullptr otherwise)
    float* rec_out;
};

#ifndef FWI_PREFETCH_ROW0
// rows of u_{n-1} / m requested before the dependency wait.  ms per 1000-step gradient at 1000x3000: 0 rows 17.0, 1 row
// 16.2, 2 rows 16.0 (80 registers, still 6 CTAs per SM); prefetching the adjoint's snapshot row as well: 16.5 - 19.0
#define FWI_PREFETCH_ROW0 2
#endif
#ifndef FWI_PDL_TRIGGER
// Where a step lets its successor's CTAs become resident: 0 = implicit, when its own CTAs exit (default), 1 = top,
// 2 = after the dense part, 3 = after the TMA wait.  Measured on the bench workload (ms per 5000-step gradient):
// no PDL 85.5, 0: 84.5, 2: 88.6, 1: 101 - an early trigger parks the whole next grid at its dependency wait and releases
// all CTAs in the same instant, which lines up their load / compute / store phases and costs L2 bandwidth; with the
// implicit trigger only the launch latency and the prologue overlap and the CTAs stay staggered.
#define FWI_PDL_TRIGGER 0
#endif
#ifndef FWI_ROW_PIPE
// rows of u_{n-1} / m requested AHEAD of the row being computed (rolling, on top of FWI_PREFETCH_ROW0 = depth).  Without it the
// loads of row r + 1 sit behind the store of row r in program order (the compiler cannot prove the rows disjoint), so a warp
// pays one L2 round trip per row.  0 = off
#define FWI_ROW_PIPE 0
#endif
#ifndef FWI_L1_PREFETCH
// rows of u_{n-1} / m (beyond the FWI_PREFETCH_ROW0 rows held in registers) requested into L1 with prefetch.global.L1 before the
// dependency wait: no registers, and the row loop's loads then hit L1 instead of paying an L2 round trip each.  0 = off
#define FWI_L1_PREFETCH 0
#endif
#ifndef FWI_STEP_MINB
#define FWI_STEP_MINB 0           // > 0: __launch_bounds__ min CTAs per SM of the step kernel (caps the registers)
#endif
#if FWI_STEP_MINB > 0
#define FWI_STEP_BOUNDS(nthreads) __launch_bounds__(nthreads, FWI_STEP_MINB)
#else
// no minimum given: ptxas settles at 80 registers (3 CTAs of 256 threads per SM); with an explicit minimum of 1 it takes 110
#define FWI_STEP_BOUNDS(nthreads) __launch_bounds__(nthreads)
#endif
#ifndef FWI_M_SMEM
// 1: the tile's rows of m (static inside a sweep) are staged in shared memory by a TMA load issued BEFORE the dependency wait
#define FWI_M_SMEM 0
#endif
constexpr size_t step_tile_floats(int bz) { return ((size_t)(128 + 2 * kHalo) * (bz + 2 * kHalo) + 31) / 32 * 32; }    // [SZ][SX] rounded up to a 128-byte multiple
template <int BZ, int NW, int MODE>
__global__ void FWI_STEP_BOUNDS(NW * 32) fd2d_step_kernel(const __grid_constant__ CUtensorMap tm_cur,
#if FWI_M_SMEM
                                                             const __grid_constant__ CUtensorMap tm_m,
#endif
                                                             Step2DArgs a) {
    constexpr int BX = 128, SX = BX + 2 * kHalo, SZ = BZ + 2 * kHalo, RPW = BZ / NW;
    static_assert(BZ % NW == 0, "rows must split evenly over warps");
    extern __shared__ __align__(128) float tile[];          // [SZ][SX] (+ [BZ][BX] of m)
    __shared__ __align__(8) uint64_t bar;
#if FWI_M_SMEM
    __shared__ __align__(8) uint64_t bar_m;
    float* sM = tile + step_tile_floats(BZ);
#endif

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx0 = blockIdx.x * BX, tz0 = blockIdx.y * BZ;
    const int x = tx0 + 4 * lane;
    const int zw = tz0 + warp * RPW;
    const bool col_ok = x < a.px;
    // ---- prologue that does not depend on the previous time step: with programmatic dependent launch this part of
    //      step n+1 runs while step n is still draining (the launch latency and the tail of a 6 us kernel overlap)
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
#if FWI_M_SMEM
        mbar_init(&bar_m, 1);
#endif
        fence_mbar_init();
        fence_proxy_async();
#if FWI_M_SMEM
        mbar_expect_tx(&bar_m, BZ * BX * (uint32_t)sizeof(float));
        tma_load_2d(sM, &tm_m, tx0, tz0, &bar_m);
#endif
    }
    float4 gx4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (col_ok) gx4 = ld4(a.gx + x);
    // (prefetching the adjoint's snapshot rows into L2 from here was measured: 17.3 vs 17.0 ms per 1000-step gradient, not kept;
    //  so was a bulk L2 prefetch, cp.async.bulk.prefetch.L2, of the two snapshot tiles of the NEXT deferred-imaging launch
    //  issued by the propagate-only step before it: 16.5 vs 15.6 us per forward+adjoint step pair at 1000 x 3000)
#if FWI_PREFETCH_ROW0
    // u_{n-1} and m of this warp's first rows: neither is written by the step this launch is chained to (u_{n-1} is the
    // output of step n-2, which completed before step n-1 passed its own wait), and requesting them here takes one L2
    // round trip off the serial chain  TMA wait -> row loads -> compute
    float4 o4p[FWI_PREFETCH_ROW0], m4p[FWI_PREFETCH_ROW0];
#pragma unroll
    for (int r = 0; r < FWI_PREFETCH_ROW0; ++r) {
        o4p[r] = make_float4(0.f, 0.f, 0.f, 0.f); m4p[r] = o4p[r];
        if (col_ok && zw + r < a.nz) {
            o4p[r] = ld4(a.oldnew + (size_t)(zw + r) * a.px + x);
#if !FWI_M_SMEM
            m4p[r] = ld4(a.m + (size_t)(zw + r) * a.px + x);
#endif
        }
    }
#endif
#if FWI_L1_PREFETCH
#pragma unroll
    for (int r = FWI_PREFETCH_ROW0; r < RPW && r < FWI_PREFETCH_ROW0 + FWI_L1_PREFETCH; ++r)
        if (col_ok && zw + r < a.nz) {
            prefetch_l1(a.oldnew + (size_t)(zw + r) * a.px + x);
            prefetch_l1(a.m + (size_t)(zw + r) * a.px + x);
        }
#endif
    griddep_wait();                 // step n complete and visible (no-op for a plain launch)
#if FWI_PDL_TRIGGER == 1
    griddep_launch_dependents();    // every CTA of this grid is resident once all have passed here: step n+2's may queue up
#endif
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, SX * SZ * (uint32_t)sizeof(float));
        tma_load_2d(tile, &tm_cur, tx0 - kHalo, tz0 - kHalo, &bar);
    }
    __syncthreads();        // barrier init visible to every waiter

    mbar_wait(&bar, 0);
#if FWI_M_SMEM
    mbar_wait(&bar_m, 0);
#endif
#if FWI_PDL_TRIGGER == 3
    griddep_launch_dependents();
#endif

    if (col_ok && zw < a.nz) {
        // register window over z: win[k] holds row (z - 4 + k) of this lane's float4 column
        float4 win[9];
        const float* tcol = tile + (warp * RPW) * SX + kHalo + 4 * lane;     // row z-4 of the first output row
#pragma unroll
        for (int k = 0; k < 8; ++k) win[k + 1] = ld4(tcol + k * SX);
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int z = zw + r;
#pragma unroll
            for (int k = 0; k < 8; ++k) win[k] = win[k + 1];
            win[8] = ld4(tcol + (r + 8) * SX);
            if (z < a.nz) {
                const float* trow = tile + (warp * RPW + r + kHalo) * SX + 4 * lane;
                const float4 L = ld4(trow), R = ld4(trow + 8);
                const float4 C = win[4];
                const float ax[12] = {L.x, L.y, L.z, L.w, C.x, C.y, C.z, C.w, R.x, R.y, R.z, R.w};
                float zc[9][4];
#pragma unroll
                for (int k = 0; k < 9; ++k) { zc[k][0] = win[k].x; zc[k][1] = win[k].y; zc[k][2] = win[k].z; zc[k][3] = win[k].w; }
                const size_t off = (size_t)z * a.px + x;
#if FWI_ROW_PIPE
                // queue: o4p[k] / m4p[k] hold row r + k; row r + FWI_PREFETCH_ROW0 is requested now, before row r's store
                const float4 o4 = o4p[0], m4 = m4p[0];
                float4 o4n = make_float4(0.f, 0.f, 0.f, 0.f), m4n = o4n;
                if (r + FWI_PREFETCH_ROW0 < RPW && z + FWI_PREFETCH_ROW0 < a.nz) {
                    o4n = ld4(a.oldnew + off + (size_t)FWI_PREFETCH_ROW0 * a.px);
                    m4n = ld4(a.m + off + (size_t)FWI_PREFETCH_ROW0 * a.px);
                }
#pragma unroll
                for (int k = 0; k + 1 < FWI_PREFETCH_ROW0; ++k) { o4p[k] = o4p[k + 1]; m4p[k] = m4p[k + 1]; }
                o4p[FWI_PREFETCH_ROW0 - 1] = o4n; m4p[FWI_PREFETCH_ROW0 - 1] = m4n;
#elif FWI_M_SMEM
                const float4 o4 = (r < FWI_PREFETCH_ROW0) ? o4p[r < FWI_PREFETCH_ROW0 ? r : 0] : ld4(a.oldnew + off);
                const float4 m4 = ld4(sM + (warp * RPW + r) * BX + 4 * lane);
#elif FWI_PREFETCH_ROW0
                const float4 o4 = (r < FWI_PREFETCH_ROW0) ? o4p[r < FWI_PREFETCH_ROW0 ? r : 0] : ld4(a.oldnew + off);
                const float4 m4 = (r < FWI_PREFETCH_ROW0) ? m4p[r < FWI_PREFETCH_ROW0 ? r : 0] : ld4(a.m + off);
#else
                const float4 o4 = ld4(a.oldnew + off);
                const float4 m4 = ld4(a.m + off);
#endif
                const float gzv = __ldg(a.gz + z);
                const float ov[4] = {o4.x, o4.y, o4.z, o4.w}, mv[4] = {m4.x, m4.y, m4.z, m4.w};
                const float gv[4] = {gx4.x * gzv, gx4.y * gzv, gx4.z * gzv, gx4.w * gzv};
                float wv[4], nv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float c = ax[4 + j];
                    float lap = (2.0f * kC0) * c;
                    lap = fmaf(kC1, (ax[3 + j] + ax[5 + j]) + (zc[3][j] + zc[5][j]), lap);
                    lap = fmaf(kC2, (ax[2 + j] + ax[6 + j]) + (zc[2][j] + zc[6][j]), lap);
                    lap = fmaf(kC3, (ax[1 + j] + ax[7 + j]) + (zc[1][j] + zc[7][j]), lap);
                    lap = fmaf(kC4, (ax[0 + j] + ax[8 + j]) + (zc[0][j] + zc[8][j]), lap);
                    wv[j] = lap;
                    nv[j] = gv[j] * fmaf(mv[j], lap, fmaf(-gv[j], ov[j], 2.0f * c));
                }
                st4(a.oldnew + off, make_float4(nv[0], nv[1], nv[2], nv[3]));
                if (MODE == STEP_FWD_SAVE) st4_stream(a.snap + off, make_float4(wv[0], wv[1], wv[2], wv[3]));
                if (MODE == STEP_ADJ || MODE == STEP_ADJ2) {
                    const float4 s4 = ld4_stream(a.snap + off);
                    float4 c4 = ld4(a.acc + off);
                    if (MODE == STEP_ADJ2) {
                        // deferred imaging of the previous adjoint step: its field is this launch's u_n (centre value C);
                        // same fma order as two consecutive ADJ launches, so the accumulator is bit-identical
                        const float4 sp = ld4_stream(a.snap_prev + off);
                        c4.x = fmaf(C.x, sp.x, c4.x); c4.y = fmaf(C.y, sp.y, c4.y);
                        c4.z = fmaf(C.z, sp.z, c4.z); c4.w = fmaf(C.w, sp.w, c4.w);
                    }
                    c4.x = fmaf(nv[0], s4.x, c4.x); c4.y = fmaf(nv[1], s4.y, c4.y);
                    c4.z = fmaf(nv[2], s4.z, c4.z); c4.w = fmaf(nv[3], s4.w, c4.w);
                    st4(a.acc + off, c4);
                }
            }
        }
    }

#if FWI_PDL_TRIGGER == 2
    griddep_launch_dependents();
#endif
    // ---- sparse fix-ups for the points this tile owns: injection, then receiver sampling ----------
    const int tid = blockIdx.y * gridDim.x + blockIdx.x;
    const int i0 = a.inj.tile_ptr ? a.inj.tile_ptr[tid] : 0, i1 = a.inj.tile_ptr ? a.inj.tile_ptr[tid + 1] : 0;
    const int r0 = a.rec.tile_ptr ? a.rec.tile_ptr[tid] : 0, r1 = a.rec.tile_ptr ? a.rec.tile_ptr[tid + 1] : 0;
    if (i1 > i0 || r1 > r0) {             // uniform per CTA
        __syncthreads();
        for (int e = i0 + threadIdx.x; e < i1; e += blockDim.x) {
            const int off = a.inj.off[e];
            const int z = off / a.px, xx = off - z * a.px;
            const float val = a.inj_vals[a.inj.id[e]];
            const float gm = a.gx[xx] * a.gz[z] * a.m[off];
            atomicAdd(a.oldnew + off, gm * val);                       // u_{n+1} += g m f
            if (MODE == STEP_FWD_SAVE) atomicAdd(a.snap + off, val);   // w_n includes f_n
            if (MODE == STEP_ADJ || MODE == STEP_ADJ2) atomicAdd(a.acc + off, gm * val * a.snap[off]);
        }
        if (r1 > r0) {
            __syncthreads();
            for (int e = r0 + threadIdx.x; e < r1; e += blockDim.x)
                a.rec_out[a.rec.id[e]] = __ldcg(a.oldnew + a.rec.off[e]);
        }
    }
}

// ------------------------------------------------------------------------------------------ small kernels
__global__ void fd_model_kernel(const float* __restrict__ v, int nz, int nx, int px, float dt_over_h,
                                float* __restrict__ m, float* __restrict__ vp) {
    const int x = blockIdx.y * blockDim.x + threadIdx.x, z = blockIdx.x;   // rows on grid.x (no 65535 limit)
    if (x >= px || z >= nz) return;
    float val = 0.f, vv = 0.f;
    if (x < nx) { vv = v[(size_t)z * nx + x]; const float c = vv * dt_over_h; val = c * c; }
    m[(size_t)z * px + x] = val;
    vp[(size_t)z * px + x] = vv;
}

__global__ void fd_grad_finalize_kernel(const float* __restrict__ acc, const float* __restrict__ vp, int nz, int nx,
                                        int px, float* __restrict__ grad) {
    const int x = blockIdx.y * blockDim.x + threadIdx.x, z = blockIdx.x;   // rows on grid.x (no 65535 limit)
    if (x >= nx || z >= nz) return;
    const size_t o = (size_t)z * px + x;
    grad[(size_t)z * nx + x] += 2.0f * acc[o] / vp[o];                 // dJ/dv = (2/v) I
}

__global__ void fd_unpitch_kernel(const float* __restrict__ src, int nz, int nx, int px, float* __restrict__ dst) {
    const int x = blockIdx.y * blockDim.x + threadIdx.x, z = blockIdx.x;   // rows on grid.x (no 65535 limit)
    if (x < nx && z < nz) dst[(size_t)z * nx + x] = src[(size_t)z * px + x];
}

__global__ void fd_residual_kernel(const float* __restrict__ syn, const float* __restrict__ obs, int64_t n,
                                   float* __restrict__ res, double* __restrict__ J) {
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float r = syn[i] - obs[i];
        res[i] = r;
        s += (double)r * (double)r;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double ws[32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += ws[w];
        atomicAdd(J, 0.5 * t);
    }
}

__global__ void fd_update_kernel(float* __restrict__ v, const float* __restrict__ g, float step, float vmin, float vmax,
                                 int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = fminf(vmax, fmaxf(vmin, v[i] - step * g[i]));
}

__global__ void fd_absmax_kernel(const float* __restrict__ x, int64_t n, unsigned int* __restrict__ out) {
    float mx = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        mx = fmaxf(mx, fabsf(x[i]));
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(mx));     // non-negative floats order like uints
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_tiled_f32(CUtensorMap* out, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        FWI_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled not available from the driver"); return FWI_ECUDA; }
        fn = (EncodeTiledFn)p;
    }
    cuuint64_t d[5]; cuuint64_t s[5]; cuuint32_t b[5]; cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
    for (int i = 0; i < rank - 1; ++i) s[i] = strides_bytes[i];
    static int promo = -1;
    if (promo < 0) { const char* e = getenv("FWI_TMA_L2PROMO"); promo = e ? atoi(e) : 2; }      // 0 none, 1 64B, 2 128B, 3 256B
    const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B :
                                      (promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B));
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, base, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return FWI_ECUDA; }
    return FWI_OK;
}

}  // namespace fwi

using namespace fwi;

// =============================================================================================== host plan
//
// Time loops are replayed as CUDA graphs: a 1000 x 3000 step kernel runs for ~6 us, about what one
// cudaLaunchKernel costs the host, so a stream of individual launches is launch-bound.  The whole forward
// pass (nt step nodes) and the whole gradient (forward with snapshots, residual, adjoint with imaging, and the
// checkpoint copies when they are needed) are each captured ONCE per (nt, geometry size) on the plan's private
// stream and re-launched for every shot.  That requires every pointer inside the graph to be stable, so the
// graph only touches plan-owned buffers: wavelet / observed data are copied into staging buffers before the
// launch, synthetics are copied out after it, and the sparse source / receiver lists keep their allocations
// (set_geometry rewrites their contents in place).
namespace {

struct PointList {
    int n = 0, cap = 0, nbins = 0;
    int* d_tile_ptr = nullptr; int* d_off = nullptr; int* d_id = nullptr;
    void release() {
        if (d_tile_ptr) cudaFree(d_tile_ptr);
        if (d_off) cudaFree(d_off);
        if (d_id) cudaFree(d_id);
        d_tile_ptr = d_off = d_id = nullptr; n = cap = nbins = 0;
    }
    PointListDev dev() const { return PointListDev{d_tile_ptr, d_off, d_id}; }
};

constexpr int kBX = 128;

struct GraphEntry {
    int kind, nt, nsrc, nrec, seg, nseg;
    cudaGraphExec_t exec;
    int64_t kernels;
};

}  // namespace

struct fwi_fd2d {
    int device = 0, nz = 0, ny = 1, nx = 0, px = 0, nabs = 0;   // ny == 1: 2-D plan; rows() = nz * ny
    float h = 0, dt = 0, alpha = 0;
    int tiles_y = 1, zchunk = 0, nzch = 1, by = 16;             // 3-D tiling: 128 x by columns (by = 16 or 14), z chunks
    float* gy = nullptr;
    CUtensorMap tm3[8], tm3_old[8], tm3_m;
    // slab decomposition over peer memory (3-D only)
    int z_own0 = 0, z_own1 = 0;                 // owned planes of the local grid; 0,0 = whole grid
    void* peer_arena[2] = {nullptr, nullptr};   // IPC-opened arenas of the upper / lower neighbour
    size_t peer_fld_off[2][8] = {};             // byte offsets of the neighbours' wavefield buffers in their arenas
    size_t peer_flags_off[2] = {0, 0};
    int peer_up_z = 0;
    bool peer_ipc[2] = {false, false};          // opened with cudaIpcOpenMemHandle (else: another plan of this process)
    int* sync_area = nullptr;                   // SlabSync words (fd3d.cuh) inside the arena; the step id lives there
    int last_desc = 0;                          // the last z chunk marches downwards (slab with a lower neighbour)
    long long slab_timeout_cycles = 4000000000LL;   // bounded spin on a neighbour's flag (~2 s; FWI_SLAB_TIMEOUT_MS)
    bool peers() const { return peer_arena[0] || peer_arena[1]; }
    int bz = 32, nw = 4;              // tiled variant: tile rows / warps per CTA (tunable)
    int tiles_x = 0, tiles_z = 0;
    int variant = 0;                  // 0 = one-tile-per-CTA kernel, 2 = two-steps-per-pass kernel (temporal blocking)
    int sm_count = 148;
    CUtensorMap tb_cur[8], tb_old[8], tb_m;     // temporally blocked kernel (variant 2)
    int cz = 32, tb_nw = 8, tiles_x2 = 0, tiles_z2 = 0;
    float *m = nullptr, *vp = nullptr, *gx = nullptr, *gz = nullptr;
    void* arena = nullptr; size_t l2_persist_bytes = 0;
    float* fld[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // forward 0..3, adjoint 4..7 (one-step kernels use 0/1 and 4/5)
    CUtensorMap tmap[8];
    CUtensorMap tmap_m;               // FWI_M_SMEM: 128 x bz box of m
    float* acc = nullptr;
    float* snap = nullptr; size_t snap_steps = 0;
    float* ckpt = nullptr; size_t ckpt_slots = 0;
    float* resid = nullptr; size_t resid_cap = 0;
    float* syn = nullptr; size_t syn_cap = 0;
    float* obs = nullptr; size_t obs_cap = 0;
    float* wav = nullptr; size_t wav_cap = 0;
    double* d_J = nullptr;
    PointList src, rec;
    PointList src_ext2, src_own2, rec_ext2, rec_own2;   // variant 2: binned by the 120 x cz core tiles (+-4 for *_ext2)
    int nsrc = 0, nrec = 0;
    size_t mem_limit = 0;             // 0 = automatic (fraction of free memory)
    bool pdl = true;                  // chain 2-D steps with programmatic dependent launch (FWI_PDL=0 turns it off)
    bool pdl_chain = false;           // false right after a kernel that rewrites what a step reads BEFORE its dependency wait (m)
    int split_nt = -1, split_seg = 0, split_nseg = 0; size_t split_limit = 0;   // cached storage decision of the last gradient
    int fwd_c = 0, fwd_o = 1;         // fld[] indices of u_n and u_{n-1} after the last forward
    bool model_set = false;
    bool use_graphs = true;
    bool defer_imaging = true;        // tile variant: image two adjoint steps per accumulator update
    int64_t launches = 0;
    cudaStream_t work = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr, ev_geom = nullptr;
    bool geom_pending = false;        // set_geometry queued uploads on `work` that a caller-driven step has not yet waited for
    // pinned host staging for the point lists: three slots used in turn, so set_geometry never waits for the shot that
    // is still running (its uploads are queued behind that shot's graph) and never reads freed pageable memory
    struct Stage { int* host = nullptr; size_t cap = 0, used = 0; cudaEvent_t done = nullptr; bool busy = false; } stage[3];
    int stage_cur = 0;
    std::vector<GraphEntry> graphs;
    int rows() const { return nz * ny; }
    size_t plane() const { return (size_t)nz * ny * px; }
};

static void drop_graphs(fwi_fd2d* p) {
    for (auto& g : p->graphs) cudaGraphExecDestroy(g.exec);
    p->graphs.clear();
}

static int make_tmaps3(fwi_fd2d* p) {
    p->tiles_x = (p->nx + k3BX - 1) / k3BX;
    // Tile height and z chunks.  One CTA per SM is resident (170 KB of plane rings) and every chunk re-reads 8 halo planes;
    // for a tile height BY and a chunk count the model cost is  waves x (planes per chunk + 8) x BY.  Single-GPU plans keep
    // BY = 16 (0.96 of the HBM roofline at 512^3); peer-memory slabs, whose few chunks make the last wave's tail expensive,
    // take the cheaper of 16 and 14 (512-wide planes: 4 x 37 = 148 tiles of 14 rows, one per SM).  FWI_FD3D_BY overrides.
    {
        const int nzo = (p->z_own1 > 0 ? p->z_own1 : p->nz) - p->z_own0;
        // peer-memory slabs: with a lower neighbour and two or more chunks the last chunk marches downwards, so that both
        // boundaries are pushed early; a 4-plane boundary must not straddle two chunks
        // (measured on 8 GPUs, 64 planes per rank, us per launch: one chunk 93.0 / 95.7 with 14- / 16-row tiles, two chunks
        // 97.9 / 100.1 - a single chunk re-reads 8 instead of 16 halo planes, which outweighs pushing the lower boundary
        // at the end of the launch; the model charges that late push 3 %; 4 GPUs, 128 planes per rank: 152.4 vs 157.2 us)
        int min_ch = 1;
        int force_by = 0, force_ch = 0;
        if (const char* e = getenv("FWI_FD3D_BY")) force_by = atoi(e);
        if (const char* e = getenv("FWI_FD3D_NZCH")) force_ch = atoi(e);       // tuning aid: fixed chunk count (1 = the lower boundary is pushed last)
        if (force_ch >= 1) min_ch = std::min(force_ch, std::max(1, nzo / 8));
        int best_by = 16, best = 1;
        double best_cost = 1e300;
        for (int by : {16, 14}) {
            if (force_by ? by != force_by : (by != 16 && !p->peers())) continue;
            const int txy = p->tiles_x * ((p->ny + by - 1) / by);
            for (int nzch = min_ch; nzch <= (force_ch >= 1 ? min_ch : std::max(min_ch, nzo / 8)); ++nzch) {
                const int zc = (nzo + nzch - 1) / nzch;
                const int real = (nzo + zc - 1) / zc;
                if (real < min_ch) continue;
                if (p->peers() && nzo - (real - 1) * zc < kHalo) continue;
                const double waves = std::ceil((double)txy * real / p->sm_count);
                const double late_push = (p->peer_arena[1] && real == 1) ? 1.03 : 1.0;
                const double cost = waves * (zc + 2 * kHalo) * by * late_push;
                if (cost < best_cost - 1e-9) { best_cost = cost; best = nzch; best_by = by; }
            }
        }
        if (force_by == 14 || force_by == 16) best_by = force_by;
        p->by = best_by;
        p->tiles_y = (p->ny + p->by - 1) / p->by;
        p->zchunk = (nzo + best - 1) / best;
        p->nzch = (nzo + p->zchunk - 1) / p->zchunk;
        p->last_desc = (p->peer_arena[1] && p->nzch >= 2) ? 1 : 0;
    }
    for (int i = 0; i < 8; ++i) {
        const uint64_t dims[3] = {(uint64_t)p->nx, (uint64_t)p->ny, (uint64_t)p->nz};
        const uint64_t strides[2] = {(uint64_t)p->px * sizeof(float), (uint64_t)p->px * p->ny * sizeof(float)};
        const uint32_t box[3] = {(uint32_t)k3SX, (uint32_t)(p->by + 2 * kHalo), 1u};
        int rc = encode_tiled_f32(&p->tm3[i], p->fld[i], 3, dims, strides, box);
        if (rc) return rc;
        const uint64_t dims_p[3] = {(uint64_t)p->px, (uint64_t)p->ny, (uint64_t)p->nz};
        const uint32_t box_o[3] = {(uint32_t)k3BX, (uint32_t)p->by, 1u};
        rc = encode_tiled_f32(&p->tm3_old[i], p->fld[i], 3, dims_p, strides, box_o);
        if (rc) return rc;
        if (i == 0 && (rc = encode_tiled_f32(&p->tm3_m, p->m, 3, dims_p, strides, box_o))) return rc;
    }
    return FWI_OK;
}

static int make_tmaps(fwi_fd2d* p) {
    if (p->ny > 1) return make_tmaps3(p);
    p->tiles_x2 = (p->nx + kT2CX - 1) / kT2CX;
    p->tiles_z2 = (p->nz + p->cz - 1) / p->cz;
    {
        const uint64_t dims[2] = {(uint64_t)p->nx, (uint64_t)p->nz};
        const uint64_t dims_p[2] = {(uint64_t)p->px, (uint64_t)p->nz};
        const uint64_t strides[1] = {(uint64_t)p->px * sizeof(float)};
        const uint32_t box_c[2] = {(uint32_t)kT2W0, (uint32_t)(p->cz + 16)}, box_o[2] = {(uint32_t)kT2W1, (uint32_t)(p->cz + 8)};
        for (int i = 0; i < 8; ++i) {
            int rc = encode_tiled_f32(&p->tb_cur[i], p->fld[i], 2, dims, strides, box_c);
            if (rc) return rc;
            rc = encode_tiled_f32(&p->tb_old[i], p->fld[i], 2, dims_p, strides, box_o);
            if (rc) return rc;
        }
        int rc = encode_tiled_f32(&p->tb_m, p->m, 2, dims_p, strides, box_o);
        if (rc) return rc;
    }
    for (int i = 0; i < 8; ++i) {
        const uint64_t dims[2] = {(uint64_t)p->nx, (uint64_t)p->nz};
        const uint64_t strides[1] = {(uint64_t)p->px * sizeof(float)};
        const uint32_t box[2] = {(uint32_t)(kBX + 2 * kHalo), (uint32_t)(p->bz + 2 * kHalo)};
        int rc = encode_tiled_f32(&p->tmap[i], p->fld[i], 2, dims, strides, box);
        if (rc) return rc;
    }
    {
        const uint64_t dims_p[2] = {(uint64_t)p->px, (uint64_t)p->nz};
        const uint64_t strides[1] = {(uint64_t)p->px * sizeof(float)};
        const uint32_t box_m[2] = {(uint32_t)kBX, (uint32_t)p->bz};
        int rc = encode_tiled_f32(&p->tmap_m, p->m, 2, dims_p, strides, box_m);
        if (rc) return rc;
    }
    return FWI_OK;
}

// bin that owns grid point (z, x): the CTA tile (tiled variant) or the warp whose row-unit range holds it
static int owner_bin(const fwi_fd2d* p, int z, int y, int x) {
    if (p->ny > 1) return ((std::max(0, z - p->z_own0) / p->zchunk) * p->tiles_y + y / p->by) * p->tiles_x + x / k3BX;
    return (z / p->bz) * p->tiles_x + x / kBX;
}

// ---- pinned staging of the point lists ---------------------------------------------------------------------------
static int stage_begin(fwi_fd2d* p, size_t need_ints) {
    p->stage_cur = (p->stage_cur + 1) % 3;
    auto& sl = p->stage[p->stage_cur];
    if (!sl.done) FWI_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    if (sl.busy) { FWI_CUDA(cudaEventSynchronize(sl.done)); sl.busy = false; }      // uploads of three set_geometry calls ago
    if (sl.cap < need_ints) {
        if (sl.host) cudaFreeHost(sl.host);
        sl.host = nullptr; sl.cap = 0;
        const size_t cap = need_ints + need_ints / 2 + 1024;
        FWI_CUDA(cudaMallocHost(&sl.host, cap * sizeof(int)));
        sl.cap = cap;
    }
    sl.used = 0;
    return FWI_OK;
}
static int stage_upload(fwi_fd2d* p, int* dst_dev, const int* src, size_t n) {
    auto& sl = p->stage[p->stage_cur];
    if (sl.used + n > sl.cap) {          // (bound computed by set_geometry was too small: fall back to a synchronous copy)
        FWI_CUDA(cudaStreamSynchronize(p->work));
        FWI_CUDA(cudaMemcpy(dst_dev, src, n * sizeof(int), cudaMemcpyHostToDevice));
        return FWI_OK;
    }
    int* h = sl.host + sl.used;
    memcpy(h, src, n * sizeof(int));
    sl.used += n;
    FWI_CUDA(cudaMemcpyAsync(dst_dev, h, n * sizeof(int), cudaMemcpyHostToDevice, p->work));
    return FWI_OK;
}
static int stage_end(fwi_fd2d* p) {
    auto& sl = p->stage[p->stage_cur];
    FWI_CUDA(cudaEventRecord(sl.done, p->work));
    sl.busy = true;
    FWI_CUDA(cudaEventRecord(p->ev_geom, p->work));
    p->geom_pending = true;
    return FWI_OK;
}

static int build_point_list(fwi_fd2d* p, PointList& pl, int n, const int* iz, const int* iy, const int* ix, const char* what) {
    const int nbins = (p->ny > 1) ? p->tiles_x * p->tiles_y * p->nzch : p->tiles_x * p->tiles_z;
    std::vector<int> tile_ptr(nbins + 1, 0), off(std::max(n, 1)), id(std::max(n, 1));
    for (int i = 0; i < n; ++i) {
        const int yy = iy ? iy[i] : 0;
        FWI_REQUIRE(iz[i] >= 0 && iz[i] < p->nz && ix[i] >= 0 && ix[i] < p->nx && yy >= 0 && yy < p->ny, "%s %d at (z=%d, y=%d, x=%d) is outside the %d x %d x %d grid", what, i, iz[i], yy, ix[i], p->nz, p->ny, p->nx);
        tile_ptr[owner_bin(p, iz[i], yy, ix[i]) + 1]++;
    }
    for (int t = 0; t < nbins; ++t) tile_ptr[t + 1] += tile_ptr[t];
    std::vector<int> fill(tile_ptr.begin(), tile_ptr.end() - 1);
    for (int i = 0; i < n; ++i) {
        const int yy = iy ? iy[i] : 0;
        const int e = fill[owner_bin(p, iz[i], yy, ix[i])]++;
        off[e] = (iz[i] * p->ny + yy) * p->px + ix[i];
        id[e] = i;
    }
    if (pl.nbins != nbins || pl.cap < std::max(n, 1)) {     // allocations change -> cached graphs hold stale pointers
        pl.release();
        drop_graphs(p);
        const int cap = std::max(n, 1);
        FWI_CUDA(cudaMalloc(&pl.d_tile_ptr, (nbins + 1) * sizeof(int)));
        FWI_CUDA(cudaMalloc(&pl.d_off, cap * sizeof(int)));
        FWI_CUDA(cudaMalloc(&pl.d_id, cap * sizeof(int)));
        pl.cap = cap; pl.nbins = nbins;
    }
    // contents are rewritten in place on the work stream, ordered after the previous shot's graph
    int rc = stage_upload(p, pl.d_tile_ptr, tile_ptr.data(), nbins + 1);
    if (!rc) rc = stage_upload(p, pl.d_off, off.data(), std::max(n, 1));
    if (!rc) rc = stage_upload(p, pl.d_id, id.data(), std::max(n, 1));
    if (rc) return rc;
    pl.n = n;
    return FWI_OK;
}

// variant 2: bin points by the 120 x cz core tiles; with `ext` a point is listed in every tile whose core +- 4 holds it
static int build_point_list_tb2(fwi_fd2d* p, PointList& pl, int n, const int* iz, const int* ix, bool ext) {
    const int nbins = p->tiles_x2 * p->tiles_z2;
    std::vector<std::vector<std::pair<int, int>>> bins(nbins);
    const int h = ext ? kHalo : 0;
    for (int i = 0; i < n; ++i) {
        const int tz0 = std::max(0, (iz[i] - h) / p->cz), tz1 = std::min(p->tiles_z2 - 1, (iz[i] + h) / p->cz);
        const int tx0 = std::max(0, (ix[i] - h) / kT2CX), tx1 = std::min(p->tiles_x2 - 1, (ix[i] + h) / kT2CX);
        for (int tz = tz0; tz <= tz1; ++tz)
            for (int tx = tx0; tx <= tx1; ++tx) bins[tz * p->tiles_x2 + tx].push_back({iz[i] * p->px + ix[i], i});
    }
    std::vector<int> tile_ptr(nbins + 1, 0), off, id;
    for (int t = 0; t < nbins; ++t) {
        tile_ptr[t + 1] = tile_ptr[t] + (int)bins[t].size();
        for (auto& e : bins[t]) { off.push_back(e.first); id.push_back(e.second); }
    }
    const int tot = std::max<int>(1, (int)off.size());
    off.resize(tot); id.resize(tot);
    if (pl.nbins != nbins || pl.cap < tot) {
        pl.release();
        drop_graphs(p);
        const int cap = tot + tot / 2 + 16;          // extended lists vary a little from shot to shot
        FWI_CUDA(cudaMalloc(&pl.d_tile_ptr, (nbins + 1) * sizeof(int)));
        FWI_CUDA(cudaMalloc(&pl.d_off, cap * sizeof(int)));
        FWI_CUDA(cudaMalloc(&pl.d_id, cap * sizeof(int)));
        pl.cap = cap; pl.nbins = nbins;
    }
    int rc = stage_upload(p, pl.d_tile_ptr, tile_ptr.data(), nbins + 1);
    if (!rc) rc = stage_upload(p, pl.d_off, off.data(), tot);
    if (!rc) rc = stage_upload(p, pl.d_id, id.data(), tot);
    if (rc) return rc;
    pl.n = tile_ptr[nbins];
    return FWI_OK;
}

template <int BZ, int NW>
static int launch_step_cfg(fwi_fd2d* p, int mode, int cur, float* oldnew, const PointList* inj, const float* inj_vals,
                           const PointList* rec, float* rec_out, float* snap, cudaStream_t st, const float* snap_prev = nullptr) {
    Step2DArgs a{};
    a.oldnew = oldnew; a.m = p->m; a.gx = p->gx; a.gz = p->gz; a.snap = snap; a.acc = p->acc; a.snap_prev = snap_prev;
    a.nx = p->nx; a.nz = p->nz; a.px = p->px;
    a.inj = (inj && inj->n) ? inj->dev() : PointListDev{nullptr, nullptr, nullptr};
    a.inj_vals = inj_vals;
    a.rec = (rec && rec->n) ? rec->dev() : PointListDev{nullptr, nullptr, nullptr};
    a.rec_out = rec_out;
    const dim3 grid(p->tiles_x, p->tiles_z), block(NW * 32);
#if FWI_M_SMEM
    const size_t smem = (step_tile_floats(BZ) + (size_t)BZ * kBX) * sizeof(float);
#define STEP_TMAPS p->tmap[cur], p->tmap_m
    static bool smem_attr_set = false;          // per <BZ, NW> instantiation: tile + m rows exceed the 48 KB default
    if (!smem_attr_set) {
        FWI_CUDA(cudaFuncSetAttribute(fd2d_step_kernel<BZ, NW, STEP_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FWI_CUDA(cudaFuncSetAttribute(fd2d_step_kernel<BZ, NW, STEP_FWD_SAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FWI_CUDA(cudaFuncSetAttribute(fd2d_step_kernel<BZ, NW, STEP_ADJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FWI_CUDA(cudaFuncSetAttribute(fd2d_step_kernel<BZ, NW, STEP_ADJ2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_attr_set = true;
    }
#else
    const size_t smem = (size_t)(kBX + 2 * kHalo) * (BZ + 2 * kHalo) * sizeof(float);
#define STEP_TMAPS p->tmap[cur]
#endif
    // Programmatic dependent launch: consecutive steps are chained with a programmatic edge (also inside captured graphs),
    // the kernel orders itself behind its predecessor with griddepcontrol.wait.
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (p->pdl && p->pdl_chain) ? 1 : 0;
    p->pdl_chain = true;
    if (mode == STEP_FWD) FWI_CUDA(cudaLaunchKernelEx(&cfg, fd2d_step_kernel<BZ, NW, STEP_FWD>, STEP_TMAPS, a));
    else if (mode == STEP_FWD_SAVE) FWI_CUDA(cudaLaunchKernelEx(&cfg, fd2d_step_kernel<BZ, NW, STEP_FWD_SAVE>, STEP_TMAPS, a));
    else if (mode == STEP_ADJ2) FWI_CUDA(cudaLaunchKernelEx(&cfg, fd2d_step_kernel<BZ, NW, STEP_ADJ2>, STEP_TMAPS, a));
    else FWI_CUDA(cudaLaunchKernelEx(&cfg, fd2d_step_kernel<BZ, NW, STEP_ADJ>, STEP_TMAPS, a));
#undef STEP_TMAPS
    return FWI_OK;
}

static int launch_step3(fwi_fd2d* p, int mode, int cur, float* oldnew, const PointList* inj, const float* inj_vals,
                        const PointList* rec, float* rec_out, float* snap, cudaStream_t st, const float* snap_prev) {
    Step3DArgs a{};
    a.oldnew = oldnew; a.m = p->m; a.gx = p->gx; a.gy = p->gy; a.gz = p->gz; a.snap = snap; a.acc = p->acc; a.snap_prev = snap_prev;
    a.nx = p->nx; a.ny = p->ny; a.nz = p->nz; a.px = p->px; a.zchunk = p->zchunk;
    a.inj = (inj && inj->n) ? inj->dev() : PointListDev{nullptr, nullptr, nullptr};
    a.inj_vals = inj_vals;
    a.rec = (rec && rec->n) ? rec->dev() : PointListDev{nullptr, nullptr, nullptr};
    a.rec_out = rec_out;
    a.z_own0 = p->z_own0; a.z_own1 = p->z_own1 > 0 ? p->z_own1 : p->nz;
    int oi = -1;
    for (int i = 0; i < 8; ++i) if (p->fld[i] == oldnew) oi = i;
    FWI_REQUIRE(oi >= 0, "fd3d: output buffer is not one of the plan's wavefields");
    a.last_desc = p->last_desc;
    a.peer_up = p->peer_arena[0] ? (float*)((char*)p->peer_arena[0] + p->peer_fld_off[0][oi]) : nullptr;
    a.peer_dn = p->peer_arena[1] ? (float*)((char*)p->peer_arena[1] + p->peer_fld_off[1][oi]) : nullptr;
    a.peer_up_z = p->peer_up_z;
    a.sync = p->sync_area;
    a.flag_peer_up = p->peer_arena[0] ? (int*)((char*)p->peer_arena[0] + p->peer_flags_off[0]) + kSyncFlagDn : nullptr;   // I am its lower neighbour
    a.flag_peer_dn = p->peer_arena[1] ? (int*)((char*)p->peer_arena[1] + p->peer_flags_off[1]) + kSyncFlagUp : nullptr;   // I am its upper neighbour
    a.timeout_cycles = p->slab_timeout_cycles;
    const dim3 grid(p->tiles_x, p->tiles_y, p->nzch);
    // (programmatic dependent launch was measured on this kernel too: 2 % at 128^3, nothing from 256^3 up - not used)
#define FD3_LAUNCH(BYV)                                                                                                         \
    do {                                                                                                                        \
        const dim3 block(T3<BYV>::Threads);                                                                                     \
        const size_t smem = T3<BYV>::Smem;                                                                                      \
        if (mode == STEP_FWD) fd3d_step_kernel<STEP_FWD, BYV><<<grid, block, smem, st>>>(p->tm3[cur], p->tm3_old[oi], p->tm3_m, a);               \
        else if (mode == STEP_FWD_SAVE) fd3d_step_kernel<STEP_FWD_SAVE, BYV><<<grid, block, smem, st>>>(p->tm3[cur], p->tm3_old[oi], p->tm3_m, a); \
        else if (mode == STEP_ADJ2) fd3d_step_kernel<STEP_ADJ2, BYV><<<grid, block, smem, st>>>(p->tm3[cur], p->tm3_old[oi], p->tm3_m, a);        \
        else fd3d_step_kernel<STEP_ADJ, BYV><<<grid, block, smem, st>>>(p->tm3[cur], p->tm3_old[oi], p->tm3_m, a);                               \
    } while (0)
    if (p->by == 14) FD3_LAUNCH(14); else FD3_LAUNCH(16);
#undef FD3_LAUNCH
    return FWI_OK;
}

static int launch_step(fwi_fd2d* p, int mode, int cur, float* oldnew, const PointList* inj, const float* inj_vals,
                       const PointList* rec, float* rec_out, float* snap, cudaStream_t st, const float* snap_prev = nullptr) {
    p->launches++;
    if (p->ny > 1) return launch_step3(p, mode, cur, oldnew, inj, inj_vals, rec, rec_out, snap, st, snap_prev);
#define CFG(BZV, NWV) if (p->bz == BZV && p->nw == NWV) return launch_step_cfg<BZV, NWV>(p, mode, cur, oldnew, inj, inj_vals, rec, rec_out, snap, st, snap_prev)
    CFG(32, 4); CFG(64, 8); CFG(16, 2);      // the round-1 sweep's winner, and one taller / one flatter tile for tests
    CFG(28, 4); CFG(42, 6); CFG(56, 8);      // 7 rows per warp: heights among which pick_tile() finds one whose tiles fill whole waves
#undef CFG
    set_error("fd2d: unsupported tile configuration bz=%d nw=%d", p->bz, p->nw);
    return FWI_EINVAL;
}

static int ensure_floats(fwi_fd2d* p, float** ptr, size_t* cap, size_t need_floats) {
    if (*cap >= need_floats && *ptr) return FWI_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr; *cap = 0;
    drop_graphs(p);
    FWI_CUDA(cudaMalloc(ptr, std::max<size_t>(need_floats, 1) * sizeof(float)));
    *cap = need_floats;
    return FWI_OK;
}

struct State { int c, o; };      // fld[] indices of u_n and u_{n-1}

static void other_two(const State& s, int base, int& a, int& b) {
    int f[2], k = 0;
    for (int i = base; i < base + 4 && k < 2; ++i) if (i != s.c && i != s.o) f[k++] = i;
    a = f[0]; b = f[1];
}
static State advance_state(const fwi_fd2d* p, State s, int base, int nsteps) {
    if (p->variant == 2 && p->ny == 1) {
        for (int k = 0; k + 1 < nsteps; k += 2) { int a, b; other_two(s, base, a, b); s = State{a, b}; }
        if (nsteps & 1) s = State{s.o, s.c};
        return s;
    }
    if (nsteps & 1) s = State{s.o, s.c};
    return s;
}

template <int CZ, int NW>
static int launch_tb2_cfg(fwi_fd2d* p, int mode, const State& s, int jc, int jd, const PointList* inj_ext, const PointList* inj_own,
                          const float* inj1, const float* inj2, const PointList* rec, float* rec1, float* rec2, float* snap1,
                          float* snap2, cudaStream_t st) {
    Tb2Args a{};
    a.out_new = p->fld[jc]; a.out_mid = p->fld[jd]; a.gx = p->gx; a.gz = p->gz; a.m = p->m; a.snap1 = snap1; a.snap2 = snap2; a.acc = p->acc;
    a.nx = p->nx; a.nz = p->nz; a.px = p->px;
    const PointListDev none{nullptr, nullptr, nullptr};
    a.inj_ext = (inj_ext && inj_ext->n) ? inj_ext->dev() : none;
    a.inj_own = (inj_own && inj_own->n) ? inj_own->dev() : none;
    a.inj1 = inj1; a.inj2 = inj2;
    a.rec = (rec && rec->n) ? rec->dev() : none;
    a.rec1 = rec1; a.rec2 = rec2;
    const dim3 grid(p->tiles_x2, p->tiles_z2), block(NW * 32);
    const size_t smem = ((size_t)(CZ + 16) * kT2W0 + 2 * (size_t)(CZ + 8) * kT2W1) * sizeof(float);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (p->pdl && p->pdl_chain) ? 1 : 0;      // the pass loads m before its dependency wait
    p->pdl_chain = true;
    if (mode == STEP_FWD) FWI_CUDA(cudaLaunchKernelEx(&cfg, fd2d_tb2_kernel<CZ, NW, STEP_FWD>, p->tb_cur[s.c], p->tb_old[s.o], p->tb_m, a));
    else if (mode == STEP_FWD_SAVE) FWI_CUDA(cudaLaunchKernelEx(&cfg, fd2d_tb2_kernel<CZ, NW, STEP_FWD_SAVE>, p->tb_cur[s.c], p->tb_old[s.o], p->tb_m, a));
    else FWI_CUDA(cudaLaunchKernelEx(&cfg, fd2d_tb2_kernel<CZ, NW, STEP_ADJ>, p->tb_cur[s.c], p->tb_old[s.o], p->tb_m, a));
    return FWI_OK;
}
template <int CZ, int NW>
static int tb2_attrs() {
    const int smem = (int)(((size_t)(CZ + 16) * kT2W0 + 2 * (size_t)(CZ + 8) * kT2W1) * sizeof(float));
    FWI_CUDA(cudaFuncSetAttribute(fd2d_tb2_kernel<CZ, NW, STEP_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    FWI_CUDA(cudaFuncSetAttribute(fd2d_tb2_kernel<CZ, NW, STEP_FWD_SAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    FWI_CUDA(cudaFuncSetAttribute(fd2d_tb2_kernel<CZ, NW, STEP_ADJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return FWI_OK;
}
static int launch_tb2(fwi_fd2d* p, int mode, const State& s, int jc, int jd, const PointList* inj_ext, const PointList* inj_own,
                      const float* inj1, const float* inj2, const PointList* rec, float* rec1, float* rec2, float* snap1, float* snap2,
                      cudaStream_t st) {
    p->launches++;
#define TB(CZV, NWV) if (p->cz == CZV && p->tb_nw == NWV) return launch_tb2_cfg<CZV, NWV>(p, mode, s, jc, jd, inj_ext, inj_own, inj1, inj2, rec, rec1, rec2, snap1, snap2, st)
    TB(32, 8); TB(24, 8); TB(16, 8);
#undef TB
    set_error("fd2d: unsupported temporal-blocking tile cz=%d", p->cz);
    return FWI_EINVAL;
}

// forward time loop over steps [n0, n1); `s` holds the fld[] indices of (u_n, u_{n-1}) on entry and on return
static int run_forward(fwi_fd2d* p, const float* wavelet, int n0, int n1, float* traces, bool save, size_t snap_base,
                       State& s, cudaStream_t st) {
    const size_t pl = p->plane();
    p->pdl_chain = false;             // (segment of) a sweep starts behind memsets / checkpoint copies: full dependency
    int n = n0;
    if (p->variant == 2 && p->ny == 1) {
        for (; n + 1 < n1; n += 2) {
            int jc, jd;
            other_two(s, 0, jc, jd);
            float* s1 = save ? p->snap + (size_t)(n - snap_base) * pl : nullptr;
            int rc = launch_tb2(p, save ? STEP_FWD_SAVE : STEP_FWD, s, jc, jd, &p->src_ext2, &p->src_own2, wavelet + (size_t)n * p->nsrc,
                                wavelet + (size_t)(n + 1) * p->nsrc, traces ? &p->rec_own2 : nullptr,
                                traces ? traces + (size_t)n * p->nrec : nullptr, traces ? traces + (size_t)(n + 1) * p->nrec : nullptr,
                                s1, save ? s1 + pl : nullptr, st);
            if (rc) return rc;
            s = State{jc, jd};
        }
    }
    // a leftover single step behind a two-steps-per-pass launch reads u_{n-1} written by that very pass: no chaining
    if (p->variant == 2 && p->ny == 1) p->pdl_chain = false;
    for (; n < n1; ++n) {
        float* snap = save ? p->snap + (size_t)(n - snap_base) * pl : nullptr;
        int rc = launch_step(p, save ? STEP_FWD_SAVE : STEP_FWD, s.c, p->fld[s.o], &p->src, wavelet + (size_t)n * p->nsrc,
                             traces ? &p->rec : nullptr, traces ? traces + (size_t)n * p->nrec : nullptr, snap, st);
        if (rc) return rc;
        s = State{s.o, s.c};
    }
    FWI_CUDA(cudaGetLastError());
    return FWI_OK;
}

// adjoint steps for trace rows n1-1 down to n0 (fields fld[4..7])
static int run_adjoint(fwi_fd2d* p, const float* resid, int n0, int n1, size_t snap_base, State& s, cudaStream_t st) {
    const size_t pl = p->plane();
    p->pdl_chain = false;             // first adjoint step follows the residual kernel and the field memsets
    int n = n1 - 1;
    if (p->variant == 2 && p->ny == 1) {
        for (; n - 1 >= n0; n -= 2) {
            int jc, jd;
            other_two(s, 4, jc, jd);
            int rc = launch_tb2(p, STEP_ADJ, s, jc, jd, &p->rec_ext2, &p->rec_own2, resid + (size_t)n * p->nrec, resid + (size_t)(n - 1) * p->nrec,
                                nullptr, nullptr, nullptr, p->snap + (size_t)(n - snap_base) * pl, p->snap + (size_t)(n - 1 - snap_base) * pl, st);
            if (rc) return rc;
            s = State{jc, jd};
        }
    }
    if ((p->ny > 1 || p->variant == 0) && p->defer_imaging) {
        // deferred imaging: the first step of a pair only propagates (17 B/pt of traffic instead of 29), the second one
        // images both fields, touching the accumulator once per two steps
        for (; n - 1 >= n0; n -= 2) {
            int rc = launch_step(p, STEP_FWD, s.c, p->fld[s.o], &p->rec, resid + (size_t)n * p->nrec, nullptr, nullptr, nullptr, st);
            if (rc) return rc;
            s = State{s.o, s.c};
            rc = launch_step(p, STEP_ADJ2, s.c, p->fld[s.o], &p->rec, resid + (size_t)(n - 1) * p->nrec, nullptr, nullptr,
                             p->snap + (size_t)(n - 1 - snap_base) * pl, st, p->snap + (size_t)(n - snap_base) * pl);
            if (rc) return rc;
            s = State{s.o, s.c};
        }
    }
    if (p->variant == 2 && p->ny == 1) p->pdl_chain = false;          // see run_forward: leftover step behind a two-step pass
    for (; n >= n0; --n) {
        int rc = launch_step(p, STEP_ADJ, s.c, p->fld[s.o], &p->rec, resid + (size_t)n * p->nrec, nullptr, nullptr,
                             p->snap + (size_t)(n - snap_base) * pl, st);
        if (rc) return rc;
        s = State{s.o, s.c};
    }
    FWI_CUDA(cudaGetLastError());
    return FWI_OK;
}

// ---- peer-memory slabs: what may be overwritten when ---------------------------------------------------------------
// A neighbour stores its boundary planes of launch k into THIS GPU's ghost planes of the buffer launch k writes.  Before a
// run (or a checkpointed segment) rewrites wavefields, a one-thread kernel waits until the neighbours have pushed
// everything up to the current step id, so no late push of the previous sweep lands on freshly written planes; and the
// ghost planes of the buffer that holds u_{n-1} are never written locally, because the neighbours' FIRST launch of the
// new sweep pushes into exactly those planes and may run ahead of this GPU's memset / restore.
static int slab_fence(fwi_fd2d* p, cudaStream_t st) {
    if (!p->peers()) return FWI_OK;
    fd3d_slab_sync_kernel<<<1, 32, 0, st>>>(p->sync_area, p->peer_arena[0] != nullptr, p->peer_arena[1] != nullptr, p->slab_timeout_cycles);
    FWI_CUDA(cudaGetLastError());
    return FWI_OK;
}
static void own_range(const fwi_fd2d* p, size_t& off, size_t& n) {      // floats of the planes a rank may write itself
    if (p->peers()) { off = (size_t)p->z_own0 * p->ny * p->px; n = (size_t)(p->z_own1 - p->z_own0) * p->ny * p->px; }
    else { off = 0; n = p->plane(); }
}
static int zero_state(fwi_fd2d* p, int cur, int old, cudaStream_t st) {
    int rc = slab_fence(p, st);
    if (rc) return rc;
    size_t off, n;
    own_range(p, off, n);
    FWI_CUDA(cudaMemsetAsync(p->fld[cur], 0, p->plane() * sizeof(float), st));
    FWI_CUDA(cudaMemsetAsync(p->fld[old] + off, 0, n * sizeof(float), st));
    return FWI_OK;
}
static int restore_state(fwi_fd2d* p, const State& c, const float* ck_cur, const float* ck_old, cudaStream_t st) {
    int rc = slab_fence(p, st);
    if (rc) return rc;
    size_t off, n;
    own_range(p, off, n);
    FWI_CUDA(cudaMemcpyAsync(p->fld[c.c], ck_cur, p->plane() * sizeof(float), cudaMemcpyDeviceToDevice, st));
    FWI_CUDA(cudaMemcpyAsync(p->fld[c.o] + off, ck_old + off, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return FWI_OK;
}

// The forward pass as a sequence of stream operations on `st` (captured into a graph or run directly).
static int record_forward(fwi_fd2d* p, int nt, cudaStream_t st) {
    int rc = zero_state(p, 0, 1, st);
    if (rc) return rc;
    State s{0, 1};
    return run_forward(p, p->wav, 0, nt, p->nrec ? p->syn : nullptr, false, 0, s, st);
}

// The gradient of one shot as a sequence of stream operations on `st`: forward (+ snapshots or checkpoints),
// residual, adjoint with imaging.  Reads p->wav / p->obs, leaves synthetics in p->syn, I in p->acc, J in p->d_J.
static int record_gradient(fwi_fd2d* p, int nt, int seg, int nseg, cudaStream_t st) {
    const size_t pl = p->plane();
    const size_t ntr = (size_t)nt * p->nrec;
    int rc = zero_state(p, 0, 1, st);
    if (rc) return rc;
    State fs{0, 1};
    std::vector<State> seg_state(nseg, fs);
    if (nseg == 1) {
        rc = run_forward(p, p->wav, 0, nt, p->nrec ? p->syn : nullptr, true, 0, fs, st);
        if (rc) return rc;
    } else {
        for (int s = 0; s < nseg; ++s) {
            if ((rc = slab_fence(p, st))) return rc;       // the ghost planes of u_n must have arrived before they are saved
            FWI_CUDA(cudaMemcpyAsync(p->ckpt + (size_t)(2 * s) * pl, p->fld[fs.c], pl * sizeof(float), cudaMemcpyDeviceToDevice, st));
            FWI_CUDA(cudaMemcpyAsync(p->ckpt + (size_t)(2 * s + 1) * pl, p->fld[fs.o], pl * sizeof(float), cudaMemcpyDeviceToDevice, st));
            seg_state[s] = fs;
            rc = run_forward(p, p->wav, s * seg, std::min(nt, (s + 1) * seg), p->nrec ? p->syn : nullptr, false, 0, fs, st);
            if (rc) return rc;
        }
    }
    FWI_CUDA(cudaMemsetAsync(p->d_J, 0, sizeof(double), st));
    if (ntr) {
        fd_residual_kernel<<<(unsigned)std::min<size_t>(1024, (ntr + 255) / 256), 256, 0, st>>>(p->syn, p->obs, (int64_t)ntr, p->resid, p->d_J);
        FWI_CUDA(cudaGetLastError());
        p->launches += 1;
    }
    if ((rc = zero_state(p, 4, 5, st))) return rc;
    FWI_CUDA(cudaMemsetAsync(p->acc, 0, pl * sizeof(float), st));
    State as{4, 5};
    if (nseg == 1) {
        rc = run_adjoint(p, p->resid, 0, nt, 0, as, st);
        if (rc) return rc;
    } else {
        for (int s = nseg - 1; s >= 0; --s) {
            const int n0 = s * seg, n1 = std::min(nt, (s + 1) * seg);
            State c = seg_state[s];
            if ((rc = restore_state(p, c, p->ckpt + (size_t)(2 * s) * pl, p->ckpt + (size_t)(2 * s + 1) * pl, st))) return rc;
            rc = run_forward(p, p->wav, n0, n1, nullptr, true, n0, c, st);     // recompute w_n for this segment
            if (rc) return rc;
            rc = run_adjoint(p, p->resid, n0, n1, n0, as, st);
            if (rc) return rc;
        }
    }
    return FWI_OK;
}

// Cached graph of a time loop, keyed by (kind, nt, nsrc, nrec, seg, nseg); captured and instantiated on first use.
template <typename F>
static int get_graph(fwi_fd2d* p, int kind, int nt, int seg, int nseg, F&& record, GraphEntry** out) {
    for (auto& g : p->graphs)
        if (g.kind == kind && g.nt == nt && g.nsrc == p->nsrc && g.nrec == p->nrec && g.seg == seg && g.nseg == nseg) { *out = &g; return FWI_OK; }
    cudaGraph_t graph = nullptr;
    FWI_CUDA(cudaStreamBeginCapture(p->work, cudaStreamCaptureModeThreadLocal));
    const int64_t l0 = p->launches;
    int rc = record(p->work);
    cudaError_t e = cudaStreamEndCapture(p->work, &graph);
    const int64_t kernels = p->launches - l0;
    p->launches = l0;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) { set_error("stream capture failed: %s", cudaGetErrorString(e)); return FWI_ECUDA; }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); return FWI_ECUDA; }
    if (p->graphs.size() >= 6) { cudaGraphExecDestroy(p->graphs.front().exec); p->graphs.erase(p->graphs.begin()); }
    p->graphs.push_back(GraphEntry{kind, nt, p->nsrc, p->nrec, seg, nseg, exec, kernels});
    *out = &p->graphs.back();
    return FWI_OK;
}

// Run `record` either directly on the work stream or as a cached graph.
template <typename F>
static int run_cached(fwi_fd2d* p, int kind, int nt, int seg, int nseg, F&& record) {
    p->pdl_chain = false;             // the first step of a sweep follows memsets / copies / foreign kernels: full dependency
    if (!p->use_graphs) return record(p->work);
    GraphEntry* g = nullptr;
    int rc = get_graph(p, kind, nt, seg, nseg, record, &g);
    if (rc) return rc;
    FWI_CUDA(cudaGraphLaunch(g->exec, p->work));
    p->launches += g->kernels;
    return FWI_OK;
}

static int enter(fwi_fd2d* p, cudaStream_t user) {       // work stream waits for everything queued on the caller's
    FWI_CUDA(cudaEventRecord(p->ev_in, user));
    FWI_CUDA(cudaStreamWaitEvent(p->work, p->ev_in, 0));
    return FWI_OK;
}
static int leave(fwi_fd2d* p, cudaStream_t user) {       // ... and the caller's stream waits for the work stream
    FWI_CUDA(cudaEventRecord(p->ev_out, p->work));
    FWI_CUDA(cudaStreamWaitEvent(user, p->ev_out, 0));
    return FWI_OK;
}

extern "C" {

static int init_plan(fwi_fd2d* p, int device, int nz, int ny, int nx, float h, float dt, int nabs, float alpha);
extern "C" int fwi_fd2d_destroy(fwi_fd2d* p);

static int create_plan(int device, int nz, int ny, int nx, float h, float dt, int nabs, float alpha, fwi_fd2d** out) {
    FWI_REQUIRE(out != nullptr, "fwi_fd_create: out is NULL");
    FWI_REQUIRE(nz >= 1 && nx >= 1 && ny >= 1, "fwi_fd_create: grid must be at least 1 x 1 (got %d x %d x %d)", nz, ny, nx);
    FWI_REQUIRE((int64_t)nz * ny * ((nx + 31) & ~31) < (int64_t)2147483647, "fwi_fd_create: grid too large for 32-bit point offsets");
    FWI_REQUIRE(h > 0.f && dt > 0.f, "fwi_fd2d_create: h and dt must be positive");
    FWI_REQUIRE(nabs >= 0 && alpha >= 0.f, "fwi_fd2d_create: nabs and alpha must be non-negative");
    int ndev = 0;
    FWI_CUDA(cudaGetDeviceCount(&ndev));
    FWI_REQUIRE(device >= 0 && device < ndev, "fwi_fd2d_create: device %d out of range (%d visible)", device, ndev);
    DeviceGuard g(device);
    auto* p = new fwi_fd2d();
    const int rc = init_plan(p, device, nz, ny, nx, h, dt, nabs, alpha);
    if (rc) { fwi_fd2d_destroy(p); return rc; }          // nothing half-built leaks
    *out = p;
    return FWI_OK;
}

static int init_plan(fwi_fd2d* p, int device, int nz, int ny, int nx, float h, float dt, int nabs, float alpha) {
    p->device = device; p->nz = nz; p->ny = ny; p->nx = nx; p->h = h; p->dt = dt; p->nabs = nabs; p->alpha = alpha;
    p->px = (nx + 31) & ~31;
    p->tiles_x = (nx + kBX - 1) / kBX;
    p->tiles_z = (nz + p->bz - 1) / p->bz;
    const size_t pl = p->plane();
    FWI_CUDA(cudaStreamCreateWithFlags(&p->work, cudaStreamNonBlocking));
    FWI_CUDA(cudaEventCreateWithFlags(&p->ev_in, cudaEventDisableTiming));
    FWI_CUDA(cudaEventCreateWithFlags(&p->ev_out, cudaEventDisableTiming));
    FWI_CUDA(cudaEventCreateWithFlags(&p->ev_geom, cudaEventDisableTiming));
    // one arena for everything the step kernels re-read every step, hottest first, so that a single L2
    // access-policy window can keep it resident while the snapshots stream past:
    //   [m, acc, fld0, fld1, fld4, fld5 | fld2, fld3, fld6, fld7, vp]
    const size_t plb = (pl * sizeof(float) + 255) & ~(size_t)255;
    FWI_CUDA(cudaMalloc(&p->arena, 11 * plb + 256));
    FWI_CUDA(cudaMemset(p->arena, 0, 11 * plb + 256));
    p->sync_area = (int*)((char*)p->arena + 11 * plb);
    {
        char* a = (char*)p->arena;
        const int order[8] = {0, 1, 4, 5, 2, 3, 6, 7};
        p->m = (float*)a; p->acc = (float*)(a + plb);
        for (int k = 0; k < 8; ++k) p->fld[order[k]] = (float*)(a + (2 + k) * plb);
        p->vp = (float*)(a + 10 * plb);
    }
    {
        cudaDeviceProp prop;
        FWI_CUDA(cudaGetDeviceProperties(&prop, device));
        { const char* e = getenv("FWI_PDL"); if (e && e[0] == '0') p->pdl = false; }
        const char* env = getenv("FWI_L2_PERSIST");
        const bool want = env && env[0] == '1';       // opt-in: measured no gain on B200 (profiles/), the .cs stores suffice
        if (want && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
            const size_t hot = 6 * plb;                                  // m, acc and the four one-step buffers
            const size_t setaside = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, hot);
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setaside) == cudaSuccess) {
                cudaStreamAttrValue at{};
                at.accessPolicyWindow.base_ptr = p->arena;
                at.accessPolicyWindow.num_bytes = std::min<size_t>(hot, (size_t)prop.accessPolicyMaxWindowSize);
                at.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)setaside / (double)at.accessPolicyWindow.num_bytes);
                at.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                at.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                if (cudaStreamSetAttribute(p->work, cudaStreamAttributeAccessPolicyWindow, &at) != cudaSuccess) cudaGetLastError();
                p->l2_persist_bytes = setaside;
            } else cudaGetLastError();
        }
    }
    FWI_CUDA(cudaMalloc(&p->gx, p->px * sizeof(float)));
    FWI_CUDA(cudaMalloc(&p->gz, nz * sizeof(float)));
    FWI_CUDA(cudaMalloc(&p->gy, ny * sizeof(float)));
    FWI_CUDA(cudaMalloc(&p->d_J, sizeof(double)));
    // sponge profiles (oracle/fd_oracle.py sponge_profile), float64 on the host then rounded once
    auto profile = [&](int n, int padded) {
        std::vector<float> prof(padded, 1.0f);
        for (int i = 0; i < std::min(nabs, n); ++i) {
            const double t = (double)alpha * (nabs - i) / nabs;
            const float val = (float)std::exp(-t * t);
            prof[i] = std::min(prof[i], val);
            prof[n - 1 - i] = std::min(prof[n - 1 - i], val);
        }
        return prof;
    };
    std::vector<float> gxh = profile(nx, p->px), gzh = profile(nz, nz), gyh = (ny > 1) ? profile(ny, ny) : std::vector<float>(1, 1.0f);
    FWI_CUDA(cudaMemcpy(p->gy, gyh.data(), ny * sizeof(float), cudaMemcpyHostToDevice));
    FWI_CUDA(cudaMemcpy(p->gx, gxh.data(), p->px * sizeof(float), cudaMemcpyHostToDevice));
    FWI_CUDA(cudaMemcpy(p->gz, gzh.data(), nz * sizeof(float), cudaMemcpyHostToDevice));
    cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
    int rc = make_tmaps(p);
    if (rc) return rc;
    if ((rc = tb2_attrs<32, 8>()) || (rc = tb2_attrs<24, 8>()) || (rc = tb2_attrs<16, 8>())) return rc;
#define FD3_ATTR(BYV)                                                                                                                        \
    FWI_CUDA(cudaFuncSetAttribute(fd3d_step_kernel<STEP_FWD, BYV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T3<BYV>::Smem));        \
    FWI_CUDA(cudaFuncSetAttribute(fd3d_step_kernel<STEP_FWD_SAVE, BYV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T3<BYV>::Smem));   \
    FWI_CUDA(cudaFuncSetAttribute(fd3d_step_kernel<STEP_ADJ, BYV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T3<BYV>::Smem));        \
    FWI_CUDA(cudaFuncSetAttribute(fd3d_step_kernel<STEP_ADJ2, BYV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T3<BYV>::Smem))
    FD3_ATTR(16); FD3_ATTR(14);
#undef FD3_ATTR
    return FWI_OK;
}

// Tile height of the one-step kernel.  All tiles of a step are resident at once (3-6 CTAs per SM), so a step lasts as long as
// the SM with the most tiles: 1000 x 3000 in 128 x 32 tiles is 768 tiles = 5.2 per SM, i.e. 28 SMs carry 6 and set the pace
// (0.865 balance), while 42-row tiles are 576 = 3.9 per SM (0.97) with less halo (50 rows read for 42 instead of 40 for 32).
// Model cost: ceil(tiles / SMs) x (rows + 4) - the 8 halo rows are loaded but not computed.  Measured at 1000 x 3000, us per
// forward+adjoint step pair: 32 rows 16.67, 28: 15.34, 42: 15.50, 56: 15.56 (forward step 6.77 / 5.83 / 5.94 / 5.85); at
// 1500 x 3000: 27.9 / 25.9 / 26.1 / 26.8.  The kernels are bit-identical, only the decomposition differs.
static void pick_tile(int nz, int nx, int sms, int* bz, int* nw) {
    const int cand[][2] = {{32, 4}, {28, 4}, {42, 6}, {56, 8}};          // (5 and 7 warps per CTA measured 1.5x slower at 1000 x 3000)
    const long tx = (nx + kBX - 1) / kBX;
    double best = 1e300;
    for (const auto& c : cand) {
        const long tiles = tx * ((nz + c[0] - 1) / c[0]);
        const double cost = (double)((tiles + sms - 1) / sms) * (c[0] + 4);
        if (cost < best - 1e-9) { best = cost; *bz = c[0]; *nw = c[1]; }
    }
}

int fwi_fd2d_create(int device, int nz, int nx, float h, float dt, int nabs, float alpha, fwi_fd2d** out) {
    int rc = create_plan(device, nz, 1, nx, h, dt, nabs, alpha, out);
    if (rc) return rc;
    // Default step kernel by grid size (tools/variant_bench.py, us per forward+adjoint step pair, tile vs two-steps-per-pass):
    //   250x3000 12.8 vs 7.8 | 500x3000 14.6 vs 11.5 | 1000x3000 18.0 vs 19.6 | 2000x3000 41.9 vs 42.7 | 3000x3000 79.7 vs 60.4
    // Small grids are launch-bound (half the launches wins), grids whose fields leave the L2 are HBM-bound (12.7 instead
    // of 17 B per step wins); in between the one-step tile kernel streams from L2 faster than tb2 can compute.
    // fwi_fd2d_set_tile / set_tb2 override this.
    const char* e = getenv("FWI_FD2D_VARIANT");
    const bool force_tile = e && !strcmp(e, "tile");
    const double pts = (double)nz * nx;
    if (!force_tile && nz >= 32 && nx >= 128) {
        if (pts <= 2.0e6) rc = fwi_fd2d_set_tb2(*out, 24);
        else if (pts > 7.5e6) rc = fwi_fd2d_set_tb2(*out, 32);
        if (rc) { fwi_fd2d_destroy(*out); *out = nullptr; return rc; }
    }
    if ((*out)->variant == 0 && nz >= 32) {
        int bz = 32, nw = 4;
        pick_tile(nz, nx, (*out)->sm_count, &bz, &nw);
        if (const char* t = getenv("FWI_FD2D_BZ")) { const int v = atoi(t); if (v == 28 || v == 42 || v == 56) { bz = v; nw = v / 7; } else if (v == 32) { bz = 32; nw = 4; } }   // tuning aid
        rc = fwi_fd2d_set_tile(*out, bz, nw);
        if (rc) { fwi_fd2d_destroy(*out); *out = nullptr; }
    }
    return rc;
}

int fwi_fd3d_create(int device, int nz, int ny, int nx, float h, float dt, int nabs, float alpha, fwi_fd2d** out) {
    FWI_REQUIRE(ny >= 2, "fwi_fd3d_create: ny must be >= 2 (use fwi_fd2d_create for 2-D grids)");
    return create_plan(device, nz, ny, nx, h, dt, nabs, alpha, out);
}

int fwi_fd2d_destroy(fwi_fd2d* p) {
    if (!p) return FWI_OK;
    DeviceGuard g(p->device);
    if (p->work) cudaStreamSynchronize(p->work);
    drop_graphs(p);
    cudaFree(p->arena); cudaFree(p->gx); cudaFree(p->gz); cudaFree(p->gy); cudaFree(p->d_J);
    if (p->snap) cudaFree(p->snap);
    if (p->ckpt) cudaFree(p->ckpt);
    if (p->resid) cudaFree(p->resid);
    if (p->syn) cudaFree(p->syn);
    if (p->obs) cudaFree(p->obs);
    if (p->wav) cudaFree(p->wav);
    p->src.release(); p->rec.release();
    p->src_ext2.release(); p->src_own2.release(); p->rec_ext2.release(); p->rec_own2.release();
    for (int s = 0; s < 2; ++s) if (p->peer_arena[s] && p->peer_ipc[s]) cudaIpcCloseMemHandle(p->peer_arena[s]);
    if (p->ev_in) cudaEventDestroy(p->ev_in);
    if (p->ev_out) cudaEventDestroy(p->ev_out);
    if (p->ev_geom) cudaEventDestroy(p->ev_geom);
    for (auto& sl : p->stage) { if (sl.host) cudaFreeHost(sl.host); if (sl.done) cudaEventDestroy(sl.done); }
    if (p->work) cudaStreamDestroy(p->work);
    delete p;
    return FWI_OK;
}

int fwi_fd2d_set_tile(fwi_fd2d* p, int bz, int nw) {
    FWI_REQUIRE(p, "fwi_fd2d_set_tile: NULL plan");
    const bool ok = (bz == 32 && nw == 4) || (bz == 64 && nw == 8) || (bz == 16 && nw == 2) || (bz == 28 && nw == 4) || (bz == 42 && nw == 6) || (bz == 56 && nw == 8);
    FWI_REQUIRE(ok, "fwi_fd2d_set_tile: unsupported (bz=%d, nw=%d)", bz, nw);
    FWI_REQUIRE(p->ny == 1, "fwi_fd2d_set_tile: 2-D plans only");
    DeviceGuard g(p->device);
    FWI_CUDA(cudaStreamSynchronize(p->work));
    drop_graphs(p);
    p->variant = 0; p->bz = bz; p->nw = nw;
    p->tiles_z = (p->nz + bz - 1) / bz;
    p->src.release(); p->rec.release(); p->nsrc = p->nrec = 0;      // tile binning changed
    return make_tmaps(p);
}

int fwi_fd2d_set_tb2(fwi_fd2d* p, int cz) {
    FWI_REQUIRE(p, "fwi_fd2d_set_tb2: NULL plan");
    FWI_REQUIRE(cz == 16 || cz == 24 || cz == 32, "fwi_fd2d_set_tb2: unsupported core rows %d (16, 24, 32)", cz);
    FWI_REQUIRE(p->ny == 1, "fwi_fd2d_set_tb2: 2-D plans only");
    DeviceGuard g(p->device);
    FWI_CUDA(cudaStreamSynchronize(p->work));
    drop_graphs(p);
    p->variant = 2; p->cz = cz;
    p->tb_nw = 8;                                                      // warps per CTA (4 warps at 24 rows measured 8 % slower)
    p->src.release(); p->rec.release(); p->nsrc = p->nrec = 0;
    p->src_ext2.release(); p->src_own2.release(); p->rec_ext2.release(); p->rec_own2.release();
    return make_tmaps(p);
}

int fwi_fd2d_set_graphs(fwi_fd2d* p, int enable) {
    FWI_REQUIRE(p, "fwi_fd2d_set_graphs: NULL plan");
    p->use_graphs = enable != 0;
    return FWI_OK;
}

int fwi_fd2d_set_memory_limit(fwi_fd2d* p, uint64_t bytes) {
    FWI_REQUIRE(p, "fwi_fd2d_set_memory_limit: NULL plan");
    p->mem_limit = (size_t)bytes;
    return FWI_OK;
}

int fwi_fd2d_set_model(fwi_fd2d* p, const float* v_dev, void* stream) {
    FWI_REQUIRE(p && v_dev, "fwi_fd2d_set_model: NULL argument");
    DeviceGuard g(p->device);
    cudaStream_t user = (cudaStream_t)stream;
    int rc = enter(p, user);
    if (rc) return rc;
    dim3 grid(p->rows(), (p->px + 127) / 128);
    fd_model_kernel<<<grid, 128, 0, p->work>>>(v_dev, p->rows(), p->nx, p->px, p->dt / p->h, p->m, p->vp);
    FWI_CUDA(cudaGetLastError());
    p->model_set = true;
    p->pdl_chain = false;             // the next step must not start reading m while this kernel is still writing it
    return leave(p, user);
}

static int set_geometry(fwi_fd2d* p, int nsrc, const int* src_z, const int* src_y, const int* src_x, int nrec,
                        const int* rec_z, const int* rec_y, const int* rec_x) {
    FWI_REQUIRE(p, "fwi_fd_set_geometry: NULL plan");
    FWI_REQUIRE(nsrc >= 0 && nrec >= 0 && (nsrc == 0 || (src_z && src_x)) && (nrec == 0 || (rec_z && rec_x)), "fwi_fd_set_geometry: bad arguments");
    DeviceGuard g(p->device);
    if (nsrc != p->nsrc || nrec != p->nrec) drop_graphs(p);
    const size_t bins = (size_t)std::max({p->tiles_x * p->tiles_y * p->nzch, p->tiles_x * p->tiles_z, p->tiles_x2 * p->tiles_z2}) + 1;
    int rc = stage_begin(p, 6 * bins + 2 * 10 * ((size_t)nsrc + nrec + 16));      // 2 + 4 lists; a tb2 point sits in <= 4 tiles
    if (rc) return rc;
    rc = build_point_list(p, p->src, nsrc, src_z, src_y, src_x, "source");
    if (rc) return rc;
    rc = build_point_list(p, p->rec, nrec, rec_z, rec_y, rec_x, "receiver");
    if (rc) return rc;
    if (p->variant == 2 && p->ny == 1) {
        if ((rc = build_point_list_tb2(p, p->src_ext2, nsrc, src_z, src_x, true))) return rc;
        if ((rc = build_point_list_tb2(p, p->src_own2, nsrc, src_z, src_x, false))) return rc;
        if ((rc = build_point_list_tb2(p, p->rec_ext2, nrec, rec_z, rec_x, true))) return rc;
        if ((rc = build_point_list_tb2(p, p->rec_own2, nrec, rec_z, rec_x, false))) return rc;
    }
    p->nsrc = nsrc; p->nrec = nrec;
    return stage_end(p);
}

int fwi_fd2d_set_geometry(fwi_fd2d* p, int nsrc, const int* src_z, const int* src_x, int nrec, const int* rec_z,
                          const int* rec_x) {
    FWI_REQUIRE(p && p->ny == 1, "fwi_fd2d_set_geometry: needs a 2-D plan");
    return set_geometry(p, nsrc, src_z, nullptr, src_x, nrec, rec_z, nullptr, rec_x);
}

int fwi_fd3d_set_geometry(fwi_fd2d* p, int nsrc, const int* src_z, const int* src_y, const int* src_x, int nrec,
                          const int* rec_z, const int* rec_y, const int* rec_x) {
    FWI_REQUIRE(p && p->ny > 1, "fwi_fd3d_set_geometry: needs a 3-D plan");
    FWI_REQUIRE((nsrc == 0 || src_y) && (nrec == 0 || rec_y), "fwi_fd3d_set_geometry: y indices missing");
    return set_geometry(p, nsrc, src_z, src_y, src_x, nrec, rec_z, rec_y, rec_x);
}

int fwi_fd2d_forward(fwi_fd2d* p, const float* wavelet_dev, int nt, float* traces_dev, void* stream) {
    FWI_REQUIRE(p && p->model_set, "fwi_fd2d_forward: set the model first");
    FWI_REQUIRE(nt >= 0 && (p->nsrc == 0 || wavelet_dev), "fwi_fd2d_forward: bad wavelet / nt");
    FWI_REQUIRE(p->nrec == 0 || traces_dev, "fwi_fd2d_forward: traces_dev is NULL but receivers are set");
    DeviceGuard g(p->device);
    cudaStream_t user = (cudaStream_t)stream;
    int rc;
    if ((rc = ensure_floats(p, &p->wav, &p->wav_cap, (size_t)nt * p->nsrc))) return rc;
    if ((rc = ensure_floats(p, &p->syn, &p->syn_cap, (size_t)nt * p->nrec))) return rc;
    if ((rc = enter(p, user))) return rc;
    if (nt * p->nsrc) FWI_CUDA(cudaMemcpyAsync(p->wav, wavelet_dev, (size_t)nt * p->nsrc * sizeof(float), cudaMemcpyDeviceToDevice, p->work));
    rc = run_cached(p, 0, nt, 0, 0, [&](cudaStream_t st) { return record_forward(p, nt, st); });
    if (rc) return rc;
    if (nt * p->nrec) FWI_CUDA(cudaMemcpyAsync(traces_dev, p->syn, (size_t)nt * p->nrec * sizeof(float), cudaMemcpyDeviceToDevice, p->work));
    { State fs = advance_state(p, State{0, 1}, 0, nt); p->fwd_c = fs.c; p->fwd_o = fs.o; }
    return leave(p, user);
}

int fwi_fd2d_wavefield(fwi_fd2d* p, int which, float* out_dev, void* stream) {
    FWI_REQUIRE(p && out_dev && which >= 0 && which <= 2, "fwi_fd2d_wavefield: bad arguments");
    DeviceGuard g(p->device);
    cudaStream_t user = (cudaStream_t)stream;
    int rc = enter(p, user);
    if (rc) return rc;
    const float* src = (which == 0) ? p->fld[p->fwd_c] : (which == 1 ? p->fld[p->fwd_o] : p->acc);
    dim3 grid(p->rows(), (p->nx + 127) / 128);
    fd_unpitch_kernel<<<grid, 128, 0, p->work>>>(src, p->rows(), p->nx, p->px, out_dev);
    FWI_CUDA(cudaGetLastError());
    return leave(p, user);
}

// Storage decision and buffers of a gradient over nt steps (no launches).  Every w_n in HBM if it fits, else two-level
// checkpointing.  The decision is cached per (nt, limit): cudaMemGetInfo takes a driver-wide lock and was measured
// stalling the host for 5 - 90 ms every few calls, which showed up as sporadic slow shots.
static int prepare_gradient(fwi_fd2d* p, int nt, int& seg, int& nseg) {
    const size_t pl = p->plane();
    seg = nt; nseg = 1;
    if (p->split_nt == nt && p->split_limit == p->mem_limit) {
        seg = p->split_seg; nseg = p->split_nseg;
    } else {
        size_t budget = p->mem_limit;
        if (!budget) {
            size_t fr = 0, tot = 0;
            FWI_CUDA(cudaMemGetInfo(&fr, &tot));
            budget = (size_t)((fr + (p->snap_steps + p->ckpt_slots) * sizeof(float)) * 0.85);   // both capacities are in floats
        }
        const size_t max_planes = budget / (pl * sizeof(float));
        if ((size_t)nt > max_planes) {
            // segment length S with S + 2*ceil(nt/S) planes minimal-ish: start from sqrt(2 nt)
            seg = std::max(1, (int)std::ceil(std::sqrt(2.0 * nt)));
            nseg = (nt + seg - 1) / seg;
            FWI_REQUIRE((size_t)seg + 2 * (size_t)nseg <= max_planes, "fwi_fd2d_gradient: %zu bytes are not enough even with checkpointing (need %zu planes of %zu bytes)", budget, (size_t)seg + 2 * (size_t)nseg, pl * sizeof(float));
        }
    }
    int rc;
    if ((rc = ensure_floats(p, &p->snap, &p->snap_steps, (size_t)seg * pl))) return rc;     // snap_steps counts floats here
    if (nseg > 1 && (rc = ensure_floats(p, &p->ckpt, &p->ckpt_slots, (size_t)nseg * 2 * pl))) return rc;
    const size_t ntr = (size_t)nt * p->nrec;
    if ((rc = ensure_floats(p, &p->resid, &p->resid_cap, ntr))) return rc;
    if ((rc = ensure_floats(p, &p->syn, &p->syn_cap, ntr))) return rc;
    if ((rc = ensure_floats(p, &p->obs, &p->obs_cap, ntr))) return rc;
    if ((rc = ensure_floats(p, &p->wav, &p->wav_cap, (size_t)nt * p->nsrc))) return rc;
    p->split_nt = nt; p->split_limit = p->mem_limit; p->split_seg = seg; p->split_nseg = nseg;      // buffers exist now
    return FWI_OK;
}

// Allocate everything fwi_fd2d_forward (gradient = 0) or fwi_fd2d_gradient (1) over nt steps with the current geometry
// will need, without launching anything: keeps allocations (which synchronise the device) out of timed regions and out
// of the way of kernels that wait for another plan.
int fwi_fd_reserve(fwi_fd2d* p, int nt, int gradient) {
    FWI_REQUIRE(p && nt >= 1, "fwi_fd_reserve: bad arguments");
    DeviceGuard g(p->device);
    int rc, seg = 0, nseg = 0;
    if (gradient) { if ((rc = prepare_gradient(p, nt, seg, nseg))) return rc; }
    else {
        if ((rc = ensure_floats(p, &p->wav, &p->wav_cap, (size_t)nt * p->nsrc))) return rc;
        if ((rc = ensure_floats(p, &p->syn, &p->syn_cap, (size_t)nt * p->nrec))) return rc;
    }
    if (p->use_graphs && p->model_set) {          // capture, instantiate and upload the time loop now as well
        GraphEntry* g = nullptr;
        p->pdl_chain = false;
        if (gradient) rc = get_graph(p, 1, nt, seg, nseg, [&](cudaStream_t st) { return record_gradient(p, nt, seg, nseg, st); }, &g);
        else rc = get_graph(p, 0, nt, 0, 0, [&](cudaStream_t st) { return record_forward(p, nt, st); }, &g);
        if (rc) return rc;
        FWI_CUDA(cudaGraphUpload(g->exec, p->work));
        FWI_CUDA(cudaStreamSynchronize(p->work));
    }
    return FWI_OK;
}

int fwi_fd2d_gradient(fwi_fd2d* p, const float* wavelet_dev, const float* obs_dev, int nt, float* grad_dev,
                      float* traces_dev, double* misfit_host, void* stream) {
    FWI_REQUIRE(p && p->model_set, "fwi_fd2d_gradient: set the model first");
    FWI_REQUIRE((wavelet_dev || p->nsrc == 0) && (obs_dev || p->nrec == 0) && grad_dev && nt >= 1, "fwi_fd2d_gradient: NULL argument or nt < 1");
    FWI_REQUIRE((p->nsrc >= 1 && p->nrec >= 1) || p->z_own1 > 0, "fwi_fd2d_gradient: geometry needs at least one source and one receiver");
    DeviceGuard g(p->device);
    cudaStream_t user = (cudaStream_t)stream;
    int seg = nt, nseg = 1, rc;
    if ((rc = prepare_gradient(p, nt, seg, nseg))) return rc;
    const size_t ntr = (size_t)nt * p->nrec;

    if ((rc = enter(p, user))) return rc;
    if (nt * p->nsrc) FWI_CUDA(cudaMemcpyAsync(p->wav, wavelet_dev, (size_t)nt * p->nsrc * sizeof(float), cudaMemcpyDeviceToDevice, p->work));
    if (ntr) FWI_CUDA(cudaMemcpyAsync(p->obs, obs_dev, ntr * sizeof(float), cudaMemcpyDeviceToDevice, p->work));
    rc = run_cached(p, 1, nt, seg, nseg, [&](cudaStream_t st) { return record_gradient(p, nt, seg, nseg, st); });
    if (rc) return rc;
    { State fs = advance_state(p, State{0, 1}, 0, nt); p->fwd_c = fs.c; p->fwd_o = fs.o; }
    dim3 grid(p->rows(), (p->nx + 127) / 128);
    fd_grad_finalize_kernel<<<grid, 128, 0, p->work>>>(p->acc, p->vp, p->rows(), p->nx, p->px, grad_dev);
    FWI_CUDA(cudaGetLastError());
    p->launches += 1;
    if (traces_dev && ntr) FWI_CUDA(cudaMemcpyAsync(traces_dev, p->syn, ntr * sizeof(float), cudaMemcpyDeviceToDevice, p->work));
    if (misfit_host) FWI_CUDA(cudaMemcpyAsync(misfit_host, p->d_J, sizeof(double), cudaMemcpyDeviceToHost, p->work));
    if ((rc = leave(p, user))) return rc;
    if (misfit_host) FWI_CUDA(cudaStreamSynchronize(p->work));
    return FWI_OK;
}

int64_t fwi_fd2d_launch_count(fwi_fd2d* p) { return p ? p->launches : 0; }

// ---- low-level stepping API: the caller drives the time loop (used by the slab-decomposed multi-GPU path, which
// interleaves halo exchanges with single steps on the caller's stream) -----------------------------------------
int fwi_fd_set_profiles(fwi_fd2d* p, const float* gz_host, const float* gy_host, const float* gx_host) {
    FWI_REQUIRE(p, "fwi_fd_set_profiles: NULL plan");
    DeviceGuard g(p->device);
    FWI_CUDA(cudaStreamSynchronize(p->work));
    if (gz_host) FWI_CUDA(cudaMemcpy(p->gz, gz_host, p->nz * sizeof(float), cudaMemcpyHostToDevice));
    if (gy_host) FWI_CUDA(cudaMemcpy(p->gy, gy_host, p->ny * sizeof(float), cudaMemcpyHostToDevice));
    if (gx_host) FWI_CUDA(cudaMemcpy(p->gx, gx_host, p->nx * sizeof(float), cudaMemcpyHostToDevice));
    return FWI_OK;
}

void* fwi_fd_field_ptr(fwi_fd2d* p, int idx) { return (p && idx >= 0 && idx < 8) ? (void*)p->fld[idx] : nullptr; }
int fwi_fd_pitch(fwi_fd2d* p) { return p ? p->px : 0; }

int fwi_fd_reserve_snapshots(fwi_fd2d* p, int nsteps) {
    FWI_REQUIRE(p && nsteps >= 0, "fwi_fd_reserve_snapshots: bad arguments");
    DeviceGuard g(p->device);
    FWI_CUDA(cudaStreamSynchronize(p->work));
    return ensure_floats(p, &p->snap, &p->snap_steps, (size_t)nsteps * p->plane());
}

// Zero the wavefield pair (0 = forward, 1 = adjoint; the adjoint pair also clears the imaging accumulator).
int fwi_fd_reset(fwi_fd2d* p, int pair, void* stream) {
    FWI_REQUIRE(p && (pair == 0 || pair == 1), "fwi_fd_reset: bad arguments");
    DeviceGuard g(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t pl = p->plane();
    FWI_CUDA(cudaMemsetAsync(p->fld[4 * pair], 0, pl * sizeof(float), st));
    FWI_CUDA(cudaMemsetAsync(p->fld[4 * pair + 1], 0, pl * sizeof(float), st));
    if (pair == 1) FWI_CUDA(cudaMemsetAsync(p->acc, 0, pl * sizeof(float), st));
    return FWI_OK;
}

// One leapfrog step on `stream`.  mode: 0 forward, 1 forward + snapshot[snap_index], 2 adjoint + imaging with
// snapshot[snap_index].  `cur` (0/1) says which buffer of the pair holds u_n; u_{n+1} lands in the other one.
// inj_vals_dev: this step's injected values (sources for modes 0/1, receiver residuals for mode 2);
// rec_out_dev: this step's trace row (modes 0/1, nullable).
int fwi_fd_step(fwi_fd2d* p, int mode, int cur, const float* inj_vals_dev, float* rec_out_dev, int64_t snap_index, void* stream) {
    FWI_REQUIRE(p && p->model_set, "fwi_fd_step: set the model first");
    FWI_REQUIRE(mode >= 0 && mode <= 2 && (cur == 0 || cur == 1), "fwi_fd_step: bad mode / cur");
    p->pdl_chain = false;             // caller-driven loop: other kernels (halo exchange, resets) sit between the steps
    FWI_REQUIRE(mode == 0 || (snap_index >= 0 && (size_t)(snap_index + 1) * p->plane() <= p->snap_steps), "fwi_fd_step: snapshot %lld not reserved", (long long)snap_index);
    DeviceGuard g(p->device);
    if (p->geom_pending) {
        // (not while the caller captures a graph: it synchronised before starting the capture, and an event recorded
        // outside a capture must not be waited for inside it)
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        FWI_CUDA(cudaStreamIsCapturing((cudaStream_t)stream, &cs));
        if (cs == cudaStreamCaptureStatusNone) { FWI_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, p->ev_geom, 0)); p->geom_pending = false; }
    }
    const int base = (mode == STEP_ADJ) ? 4 : 0;
    float* snap = (mode == 0) ? nullptr : p->snap + (size_t)snap_index * p->plane();
    int rc = launch_step(p, mode, base + cur, p->fld[base + (cur ^ 1)], mode == STEP_ADJ ? &p->rec : &p->src, inj_vals_dev,
                         (mode != STEP_ADJ && rec_out_dev) ? &p->rec : nullptr, rec_out_dev, snap, (cudaStream_t)stream);
    if (rc) return rc;
    FWI_CUDA(cudaGetLastError());
    if (mode != STEP_ADJ) { p->fwd_c = cur ^ 1; p->fwd_o = cur; }
    return FWI_OK;
}

// ---- slab decomposition over NVLink peer memory (3-D): the step kernel itself stores its boundary planes into the
// neighbours' ghost planes and publishes a step id; the next launch of the neighbour waits for it. -------------------
int fwi_fd_slab_info(fwi_fd2d* p, void* ipc_handle_out /*64 bytes*/, uint64_t* offsets_out /*9: 8 wavefield buffers + sync area*/) {
    FWI_REQUIRE(p && ipc_handle_out && offsets_out, "fwi_fd_slab_info: bad arguments");
    DeviceGuard g(p->device);
    cudaIpcMemHandle_t h;
    FWI_CUDA(cudaIpcGetMemHandle(&h, p->arena));
    memcpy(ipc_handle_out, &h, sizeof(h));
    for (int i = 0; i < 8; ++i) offsets_out[i] = (uint64_t)((char*)p->fld[i] - (char*)p->arena);
    offsets_out[8] = (uint64_t)((char*)p->sync_area - (char*)p->arena);
    return FWI_OK;
}

static int slab_finish_connect(fwi_fd2d* p, int z_own0, int z_own1, int up_ghost_z) {
    p->z_own0 = z_own0; p->z_own1 = z_own1; p->peer_up_z = up_ghost_z;
    if (const char* e = getenv("FWI_SLAB_TIMEOUT_MS")) { const double ms = atof(e); if (ms > 0) p->slab_timeout_cycles = (long long)(ms * 1.9e6); }
    p->src.release(); p->rec.release(); p->nsrc = p->nrec = 0;         // z chunking (and so the point binning) changes
    return make_tmaps(p);
}

// own planes [z_own0, z_own1) of the local grid; neighbour handles (null = no neighbour on that side); up_ghost_z = first
// ghost plane index, in the upper neighbour's local grid, that receives this rank's first four owned planes.
int fwi_fd_slab_connect(fwi_fd2d* p, int z_own0, int z_own1, const void* up_handle, const uint64_t* up_offsets, int up_ghost_z,
                        const void* dn_handle, const uint64_t* dn_offsets) {
    FWI_REQUIRE(p && p->ny > 1, "fwi_fd_slab_connect: needs a 3-D plan");
    FWI_REQUIRE(z_own0 >= 0 && z_own1 <= p->nz && z_own1 - z_own0 >= 2 * kHalo, "fwi_fd_slab_connect: owned range [%d, %d) invalid for nz=%d", z_own0, z_own1, p->nz);
    DeviceGuard g(p->device);
    FWI_CUDA(cudaStreamSynchronize(p->work));
    drop_graphs(p);
    const void* hs[2] = {up_handle, dn_handle};
    const uint64_t* offs[2] = {up_offsets, dn_offsets};
    for (int s = 0; s < 2; ++s) {
        if (!hs[s]) continue;
        FWI_REQUIRE(offs[s], "fwi_fd_slab_connect: offsets missing");
        cudaIpcMemHandle_t h;
        memcpy(&h, hs[s], sizeof(h));
        FWI_CUDA(cudaIpcOpenMemHandle(&p->peer_arena[s], h, cudaIpcMemLazyEnablePeerAccess));
        p->peer_ipc[s] = true;
        for (int i = 0; i < 8; ++i) p->peer_fld_off[s][i] = (size_t)offs[s][i];
        p->peer_flags_off[s] = (size_t)offs[s][8];
    }
    return slab_finish_connect(p, z_own0, z_own1, up_ghost_z);
}

// Same protocol between plans of ONE process (peers are other plans on this or a peer-accessible device): used by the
// single-GPU protocol test, where two small plans run on two streams of the same GPU.
int fwi_fd_slab_connect_local(fwi_fd2d* p, int z_own0, int z_own1, fwi_fd2d* up, int up_ghost_z, fwi_fd2d* dn) {
    FWI_REQUIRE(p && p->ny > 1, "fwi_fd_slab_connect_local: needs a 3-D plan");
    FWI_REQUIRE(z_own0 >= 0 && z_own1 <= p->nz && z_own1 - z_own0 >= 2 * kHalo, "fwi_fd_slab_connect_local: owned range [%d, %d) invalid for nz=%d", z_own0, z_own1, p->nz);
    DeviceGuard g(p->device);
    FWI_CUDA(cudaStreamSynchronize(p->work));
    drop_graphs(p);
    fwi_fd2d* nb[2] = {up, dn};
    for (int s = 0; s < 2; ++s) {
        if (!nb[s]) continue;
        FWI_REQUIRE(nb[s]->ny == p->ny && nb[s]->px == p->px, "fwi_fd_slab_connect_local: neighbour has a different plane shape");
        p->peer_arena[s] = nb[s]->arena;
        p->peer_ipc[s] = false;
        for (int i = 0; i < 8; ++i) p->peer_fld_off[s][i] = (size_t)((char*)nb[s]->fld[i] - (char*)nb[s]->arena);
        p->peer_flags_off[s] = (size_t)((char*)nb[s]->sync_area - (char*)nb[s]->arena);
    }
    return slab_finish_connect(p, z_own0, z_own1, up_ghost_z);
}

int fwi_fd_slab_set_timeout(fwi_fd2d* p, double milliseconds) {
    FWI_REQUIRE(p && milliseconds > 0, "fwi_fd_slab_set_timeout: bad arguments");
    p->slab_timeout_cycles = (long long)(milliseconds * 1.9e6);      // clock64 ticks at ~1.9 GHz
    drop_graphs(p);                                                   // the timeout is a kernel argument
    return FWI_OK;
}

// Clears the error flag after the caller has dealt with a timeout (all ranks together, with the device idle).
int fwi_fd_slab_clear_error(fwi_fd2d* p) {
    FWI_REQUIRE(p, "fwi_fd_slab_clear_error: NULL plan");
    DeviceGuard g(p->device);
    FWI_CUDA(cudaDeviceSynchronize());
    FWI_CUDA(cudaMemset(p->sync_area + kSyncError, 0, sizeof(int)));
    return FWI_OK;
}

int fwi_fd_slab_error(fwi_fd2d* p, int* error_out) {
    FWI_REQUIRE(p && error_out, "fwi_fd_slab_error: bad arguments");
    DeviceGuard g(p->device);
    FWI_CUDA(cudaMemcpy(error_out, p->sync_area + kSyncError, sizeof(int), cudaMemcpyDeviceToHost));
    return FWI_OK;
}

// grad_dev (dense grid) += (2 / v) * imaging accumulator
int fwi_fd_finalize_gradient(fwi_fd2d* p, float* grad_dev, void* stream) {
    FWI_REQUIRE(p && grad_dev, "fwi_fd_finalize_gradient: bad arguments");
    DeviceGuard g(p->device);
    dim3 grid(p->rows(), (p->nx + 127) / 128);
    fd_grad_finalize_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(p->acc, p->vp, p->rows(), p->nx, p->px, grad_dev);
    FWI_CUDA(cudaGetLastError());
    return FWI_OK;
}


// Small persistent device scratch per GPU for the plan-less reductions below.  (cudaMallocAsync / cudaFreeAsync around a
// synchronising call hands the block back to the driver every time: measured 8 ms per call with 25 GB allocated, which
// was 40 % of a config-3 line search.)  One caller per device at a time: both users synchronise before returning.
static void* device_scratch() {
    static void* scratch[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!scratch[dev] && cudaMalloc(&scratch[dev], 256) != cudaSuccess) { cudaGetLastError(); scratch[dev] = nullptr; }
    return scratch[dev];
}

int fwi_fd_misfit(const float* syn_dev, const float* obs_dev, int64_t n, float* resid_dev, double* misfit_host, void* stream) {
    FWI_REQUIRE(syn_dev && obs_dev && resid_dev && misfit_host && n >= 0, "fwi_fd_misfit: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    double* dJ = (double*)device_scratch();
    FWI_REQUIRE(dJ, "fwi_fd_misfit: no device scratch");
    FWI_CUDA(cudaMemsetAsync(dJ, 0, sizeof(double), st));
    if (n > 0) fd_residual_kernel<<<(unsigned)std::min<int64_t>(1024, (n + 255) / 256), 256, 0, st>>>(syn_dev, obs_dev, n, resid_dev, dJ);
    FWI_CUDA(cudaGetLastError());
    FWI_CUDA(cudaMemcpyAsync(misfit_host, dJ, sizeof(double), cudaMemcpyDeviceToHost, st));
    FWI_CUDA(cudaStreamSynchronize(st));
    return FWI_OK;
}

int fwi_fd_model_update(float* v_dev, const float* grad_dev, int64_t n, float step, float vmin, float vmax, void* stream) {
    FWI_REQUIRE(v_dev && grad_dev && n >= 0 && vmin <= vmax, "fwi_fd_model_update: bad arguments");
    if (n > 0) fd_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(v_dev, grad_dev, step, vmin, vmax, n);
    FWI_CUDA(cudaGetLastError());
    return FWI_OK;
}

int fwi_fd_absmax(const float* x_dev, int64_t n, float* out_host, void* stream) {
    FWI_REQUIRE(x_dev && out_host && n >= 1, "fwi_fd_absmax: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int* d = (unsigned int*)device_scratch();
    FWI_REQUIRE(d, "fwi_fd_absmax: no device scratch");
    d += 16;                                             // its own slot (the misfit uses the first 8 bytes)
    FWI_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned int), st));
    fd_absmax_kernel<<<(unsigned)std::min<int64_t>(1024, (n + 255) / 256), 256, 0, st>>>(x_dev, n, d);
    FWI_CUDA(cudaGetLastError());
    unsigned int h = 0;
    FWI_CUDA(cudaMemcpyAsync(&h, d, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    FWI_CUDA(cudaStreamSynchronize(st));
    memcpy(out_host, &h, sizeof(float));
    return FWI_OK;
}

}  // extern "C"
