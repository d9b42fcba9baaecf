// Track B, 2-D streaming step kernel: persistent, warp-specialised-by-column, TMA-pipelined.
//
// The grid is cut into 128-column strips; the (strip, row) pairs are linearised strip-major into
// U = nstrips * nz "row units" and every warp of the persistent grid (one CTA per SM) owns one contiguous
// range of units, so all 148 x NW warps carry the same number of rows (+-1) whatever the grid shape.
// A warp marches down its rows as an independent worker:
//   * lane 0 keeps NC TMA transactions in flight; a transaction brings a 136 x 4 box of u_n (4 more rows of the
//     column incl. the x halo; out-of-grid parts zero-filled = Dirichlet) and the 128 x 4 boxes of u_{n-1} and m
//     that the block two stages behind needs, and completes on that stage's mbarrier;
//   * the nine z-neighbours live in a register window that rotates down the column (one new LDS.128 per output
//     float4), the x-neighbours come from two more LDS.128 of the same smem row;
//   * u_{n+1} is stored over u_{n-1} with coalesced 512-byte warp rows; the forward field w_n streams to the HBM
//     snapshot with evict-first stores; the adjoint variant prefetches its snapshot / accumulator rows one block
//     ahead in registers and fuses the zero-lag cross-correlation.
// No __syncthreads anywhere: warps never share data, so a slow warp never stalls the others.
#pragma once
#include "fd_common.cuh"

namespace fwi {

constexpr int kSR = 4;                      // rows per pipeline stage
constexpr int kSCW = 128 + 2 * kHalo;       // columns of a u_n box
constexpr int kCurBytes = kSR * kSCW * 4;   // 2176
constexpr int kRowBytes = kSR * 128 * 4;    // 2048 (u_{n-1} or m block)
constexpr int kStageBytes = kCurBytes + 2 * kRowBytes;   // 6272 = 49 * 128

struct Stream2DArgs {
    float* oldnew;
    const float* gx;
    const float* gz;
    const float* m;            // for the sparse fix-ups only (the dense path reads m through TMA)
    float* snap;
    float* acc;
    int nx, nz, px, nstrips;
    const int* warp_u0;        // [W + 1] first row-unit of every warp
    PointListDev inj;          // binned by owning warp
    const float* inj_vals;
    PointListDev rec;
    float* rec_out;
};

template <int NW, int NC, int MODE>
__global__ void __launch_bounds__(NW * 32, 1)
fd2d_stream_kernel(const __grid_constant__ CUtensorMap tm_cur, const __grid_constant__ CUtensorMap tm_old,
                   const __grid_constant__ CUtensorMap tm_m, Stream2DArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = smem_raw + (size_t)warp * NC * kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NW * NC * kStageBytes) + warp * NC;

    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NC; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncwarp();

    const int gw = blockIdx.x * NW + warp;
    int u = a.warp_u0[gw];
    const int u_end = a.warp_u0[gw + 1];
    uint32_t gstage = 0;                      // running stage counter (slot = g % NC, parity = (g / NC) & 1)

    while (u < u_end) {
        const int strip = u / a.nz;
        const int z0 = u - strip * a.nz;
        const int z1 = min(a.nz, z0 + (u_end - u));
        u += z1 - z0;
        const int x0 = strip * 128;
        const int x = x0 + 4 * lane;
        const bool col_ok = x < a.px;
        const int nblocks = (z1 - z0 + kSR - 1) / kSR;
        const int nstages = nblocks + 2;

        auto issue = [&](int j) {            // lane 0 only
            const uint32_t g = gstage + j;
            const int slot = g % NC;
            uint64_t* bar = &bars[slot];
            unsigned char* cur_dst = wbase + (size_t)slot * kStageBytes;
            mbar_expect_tx(bar, j >= 2 ? kStageBytes : kCurBytes);
            tma_load_2d(cur_dst, &tm_cur, x0 - kHalo, z0 - kHalo + kSR * j, bar);
            if (j >= 2) {
                // u_{n-1} and m rows of block j-2 ride in the slot of stage j
                tma_load_2d(cur_dst + kCurBytes, &tm_old, x0, z0 + kSR * (j - 2), bar);
                tma_load_2d(cur_dst + kCurBytes + kRowBytes, &tm_m, x0, z0 + kSR * (j - 2), bar);
            }
        };
        auto wait = [&](int j) {
            const uint32_t g = gstage + j;
            mbar_wait(&bars[g % NC], (g / NC) & 1);
        };
        auto cur_ptr = [&](int j) { return reinterpret_cast<const float*>(wbase + (size_t)((gstage + j) % NC) * kStageBytes); };

        if (lane == 0) {
            const int pre = min(NC, nstages);
            for (int j = 0; j < pre; ++j) issue(j);
        }
        float4 gx4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (col_ok) gx4 = ld4(a.gx + x);

        // adjoint: snapshot / accumulator rows, prefetched one block ahead
        float4 sn[kSR], ac[kSR];
        auto prefetch_adj = [&](int b) {
#pragma unroll
            for (int r = 0; r < kSR; ++r) {
                const int z = z0 + kSR * b + r;
                if (col_ok && z < z1) {
                    const size_t off = (size_t)z * a.px + x;
                    sn[r] = ld4_stream(a.snap + off);
                    ac[r] = ld4(a.acc + off);
                }
            }
        };
        if (MODE == STEP_ADJ) prefetch_adj(0);

        // prime the register window with the 8 rows above the first output row
        float4 win[9];
        wait(0);
        wait(1);
        {
            const float* c0 = cur_ptr(0) + kHalo + 4 * lane;
            const float* c1 = cur_ptr(1) + kHalo + 4 * lane;
#pragma unroll
            for (int k = 0; k < 4; ++k) { win[k + 1] = ld4(c0 + k * kSCW); win[k + 5] = ld4(c1 + k * kSCW); }
        }
        __syncwarp();
        if (lane == 0 && NC < nstages) issue(NC);      // chunk 0 lives in registers now: its slot takes stage NC

        for (int b = 0; b < nblocks; ++b) {
            const int j = b + 2;
            wait(j);
            const float* cnew = cur_ptr(j) + kHalo + 4 * lane;          // rows entering the window
            const float* cmid = cur_ptr(j - 1) + 4 * lane;              // centre rows (for the x neighbours)
            const float* ob = cur_ptr(j) + (kCurBytes / 4) + 4 * lane;  // u_{n-1} rows of this block
            const float* mb = ob + (kRowBytes / 4);
            float4 snn[kSR], acn[kSR];
            if (MODE == STEP_ADJ && b + 1 < nblocks) {
#pragma unroll
                for (int r = 0; r < kSR; ++r) {
                    const int z = z0 + kSR * (b + 1) + r;
                    if (col_ok && z < z1) {
                        const size_t off = (size_t)z * a.px + x;
                        snn[r] = ld4_stream(a.snap + off);
                        acn[r] = ld4(a.acc + off);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < kSR; ++r) {
                const int z = z0 + kSR * b + r;
#pragma unroll
                for (int k = 0; k < 8; ++k) win[k] = win[k + 1];
                win[8] = ld4(cnew + r * kSCW);
                if (col_ok && z < z1) {
                    const float4 L = ld4(cmid + r * kSCW), R = ld4(cmid + r * kSCW + 8);
                    const float4 C = win[4];
                    const float ax[12] = {L.x, L.y, L.z, L.w, C.x, C.y, C.z, C.w, R.x, R.y, R.z, R.w};
                    float zc[9][4];
#pragma unroll
                    for (int k = 0; k < 9; ++k) { zc[k][0] = win[k].x; zc[k][1] = win[k].y; zc[k][2] = win[k].z; zc[k][3] = win[k].w; }
                    const float4 o4 = ld4(ob + r * 128), m4 = ld4(mb + r * 128);
                    const float gzv = __ldg(a.gz + z);
                    const float ov[4] = {o4.x, o4.y, o4.z, o4.w}, mv[4] = {m4.x, m4.y, m4.z, m4.w};
                    const float gv[4] = {gx4.x * gzv, gx4.y * gzv, gx4.z * gzv, gx4.w * gzv};
                    float wv[4], nv[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float c = ax[4 + q];
                        float lap = (2.0f * kC0) * c;
                        lap = fmaf(kC1, (ax[3 + q] + ax[5 + q]) + (zc[3][q] + zc[5][q]), lap);
                        lap = fmaf(kC2, (ax[2 + q] + ax[6 + q]) + (zc[2][q] + zc[6][q]), lap);
                        lap = fmaf(kC3, (ax[1 + q] + ax[7 + q]) + (zc[1][q] + zc[7][q]), lap);
                        lap = fmaf(kC4, (ax[0 + q] + ax[8 + q]) + (zc[0][q] + zc[8][q]), lap);
                        wv[q] = lap;
                        nv[q] = gv[q] * fmaf(mv[q], lap, fmaf(-gv[q], ov[q], 2.0f * c));
                    }
                    const size_t off = (size_t)z * a.px + x;
                    st4(a.oldnew + off, make_float4(nv[0], nv[1], nv[2], nv[3]));
                    if (MODE == STEP_FWD_SAVE) st4_stream(a.snap + off, make_float4(wv[0], wv[1], wv[2], wv[3]));
                    if (MODE == STEP_ADJ) {
                        float4 c4 = ac[r];
                        c4.x = fmaf(nv[0], sn[r].x, c4.x); c4.y = fmaf(nv[1], sn[r].y, c4.y);
                        c4.z = fmaf(nv[2], sn[r].z, c4.z); c4.w = fmaf(nv[3], sn[r].w, c4.w);
                        st4(a.acc + off, c4);
                    }
                }
            }
            if (MODE == STEP_ADJ) {
#pragma unroll
                for (int r = 0; r < kSR; ++r) { sn[r] = snn[r]; ac[r] = acn[r]; }
            }
            // the slot of stage j-1 (centre rows) is dead now: refill it with stage j-1+NC
            __syncwarp();
            if (lane == 0 && j - 1 + NC < nstages) issue(j - 1 + NC);
        }
        gstage += nstages;
        __syncwarp();
    }

    // ---- sparse fix-ups for the points this warp owns: injection, then receiver sampling -------------------
    const int i0 = a.inj.tile_ptr ? a.inj.tile_ptr[gw] : 0, i1 = a.inj.tile_ptr ? a.inj.tile_ptr[gw + 1] : 0;
    const int r0 = a.rec.tile_ptr ? a.rec.tile_ptr[gw] : 0, r1 = a.rec.tile_ptr ? a.rec.tile_ptr[gw + 1] : 0;
    if (i1 > i0 || r1 > r0) {
        __syncwarp();
        for (int e = i0 + lane; e < i1; e += 32) {
            const int off = a.inj.off[e];
            const int z = off / a.px, xx = off - z * a.px;
            const float val = a.inj_vals[a.inj.id[e]];
            const float gm = a.gx[xx] * a.gz[z] * a.m[off];
            atomicAdd(a.oldnew + off, gm * val);                       // u_{n+1} += g m f
            if (MODE == STEP_FWD_SAVE) atomicAdd(a.snap + off, val);   // w_n includes f_n
            if (MODE == STEP_ADJ) atomicAdd(a.acc + off, gm * val * a.snap[off]);
        }
        if (r1 > r0) {
            __syncwarp();
            for (int e = r0 + lane; e < r1; e += 32) a.rec_out[a.rec.id[e]] = __ldcg(a.oldnew + a.rec.off[e]);
        }
    }
}

}  // namespace fwi
