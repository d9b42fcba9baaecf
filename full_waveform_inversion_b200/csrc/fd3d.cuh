// Track B, 3-D: 25-point (8th-order) leapfrog step, marching along z with a TMA plane pipeline.
//
// Layout [nz][ny][px] fp32, x contiguous.  A CTA owns a 128 (x) x 16 (y) column of the grid over a chunk of z and
// marches through it plane by plane:
//   * two dedicated producer warps (one elected lane each) feed shared memory by TMA with full/empty mbarrier pairs:
//     the first keeps a ring of NP = 8 u_n planes filled (one 136 x 24 x 1 box per z plane, halo included,
//     out-of-grid zero-filled = Dirichlet), the second a 4-deep ring of u_{n-1} / m planes (128 x 16 x 1 boxes).
//     They are separate warps on purpose: a single producer that blocks on one ring's empty barrier also holds back
//     the other ring's loads, which cost 25 % of the step (290 -> 390 Gpt/s at 512^3);
//   * 8 consumer warps own two y rows each, a lane owns a float4 of x.  The nine z-neighbours of every output live
//     in a register window that rotates as the march advances (one LDS.128 per new plane); the x and y neighbours
//     come from the centre plane, which is still in the ring four planes behind the newest one;
//   * u_{n+1} overwrites u_{n-1} in place (plain per-warp global loads of u_{n-1} / m instead of the second ring cost
//     40 % of the step); snapshot write / snapshot read + imaging are fused exactly as in 2-D; the CTA applies its own
//     source / receiver points at the end.
// Algorithmic traffic: 16 B per point update (+ halo re-reads that hit L2).
#pragma once
#include "fd_common.cuh"

namespace fwi {

#ifndef FD3_RPW
#define FD3_RPW 2
#endif
#ifndef FD3_NP
#define FD3_NP 8
#endif
#ifndef FD3_NO
#define FD3_NO 4
#endif
#ifndef FD3_OMLEAD
#define FD3_OMLEAD 5
#endif
#ifndef FD3_SPLIT
#define FD3_SPLIT 1
#endif
constexpr int k3BX = 128, k3BY = 16, k3NP = FD3_NP;                  // tile, ring depth
constexpr int k3RPW = FD3_RPW, k3CW = k3BY / k3RPW;                  // y rows per consumer warp, consumer warps
constexpr int k3SX = k3BX + 2 * kHalo, k3SY = k3BY + 2 * kHalo;     // 136 x 24
constexpr int k3PlaneFloats = k3SX * k3SY;                          // 3264 floats = 13056 B (102 * 128)
constexpr int k3NO = FD3_NO, k3OmLead = FD3_OMLEAD, k3Prod = 1 + FD3_SPLIT;                               // u_{n-1}/m plane ring depth; issued 3 planes before use
constexpr int k3OmFloats = k3BX * k3BY;                             // 2048 floats = 8 KB per array per plane

struct Step3DArgs {
    float* oldnew;
    const float* m;
    const float* gx;
    const float* gy;
    const float* gz;
    float* snap;
    const float* snap_prev;   // ADJ2: snapshot paired with the previous adjoint step's field (= u_n of this launch)
    float* acc;
    int nx, ny, nz, px, zchunk;
    PointListDev inj;
    const float* inj_vals;
    PointListDev rec;
    float* rec_out;
    // ---- slab decomposition over NVLink peer memory (all null / zero for a single-GPU plan) -----------------------
    int z_own0, z_own1;        // owned plane range of the local grid (ghost planes lie outside and are never computed)
    float* peer_up;            // the upper / lower neighbour's copy of `oldnew` (peer-mapped), or null
    float* peer_dn;
    int peer_up_z;             // first ghost plane (in the upper neighbour's local grid) that receives my first 4 owned planes
    int* flags_local;          // [0] written by the upper neighbour, [1] by the lower one: "my step k is complete"
    int* flag_peer_up;         // where I announce completion to the upper / lower neighbour (their flags_local slots)
    int* flag_peer_dn;
    int wait_id, signal_id;    // this launch needs neighbours' flags >= wait_id and publishes signal_id
    unsigned int* done_counter;
    int* error_flag;
};

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__((k3CW + k3Prod) * 32, 1) fd3d_step_kernel(const __grid_constant__ CUtensorMap tm_cur,
                                                                       const __grid_constant__ CUtensorMap tm_old,
                                                                       const __grid_constant__ CUtensorMap tm_m, Step3DArgs a) {
    extern __shared__ __align__(128) float ring[];                   // [NP][SY][SX] u_n planes, then [NO][2][BY][BX] u_{n-1} / m planes
    float* om_ring = ring + (size_t)k3NP * k3PlaneFloats;
    __shared__ __align__(8) uint64_t full_bar[k3NP], empty_bar[k3NP], om_full[k3NO], om_empty[k3NO];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * k3BX, y0 = blockIdx.y * k3BY;
    const int zc0 = a.z_own0 + blockIdx.z * a.zchunk, zc1 = min(a.z_own1, zc0 + a.zchunk);
    const int nout = zc1 - zc0;                   // output planes of this CTA
    const int nplanes = nout + 2 * kHalo;         // planes zc0-4 .. zc1+3

    if (threadIdx.x == 0) {
        // slab mode: the ghost planes of u_n are written by the neighbours' previous launch straight into this GPU's
        // memory; wait until both have announced it (bounded spin: a dead neighbour raises error_flag instead of hanging)
        // Only the first / last z chunk reads (and later pushes to) the upper / lower ghost planes, so only those CTAs wait.
        if (a.wait_id > 0) {
            for (int side = 0; side < 2; ++side) {
                if ((side == 0 && !a.peer_up) || (side == 1 && !a.peer_dn)) continue;
                if ((side == 0 && blockIdx.z != 0) || (side == 1 && blockIdx.z != gridDim.z - 1)) continue;
                const long long t0 = clock64();
                while (ld_acquire_sys(a.flags_local + side) < a.wait_id) {
                    if (clock64() - t0 > 6000000000LL) { atomicExch(a.error_flag, 1); break; }      // ~3 s
                    __nanosleep(200);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < k3NP; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], k3CW); }
#pragma unroll
        for (int i = 0; i < k3NO; ++i) { mbar_init(&om_full[i], 1); mbar_init(&om_empty[i], k3CW); }
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();

    if (warp >= k3CW) {
        // ---------------- producer warp(s): one lane feeds the u_n plane ring, one the u_{n-1} / m ring
        const bool do_u = warp == k3CW, do_om = warp == k3CW + k3Prod - 1;
        if (lane == 0) {
            // tick t: u_n plane t, then the u_{n-1} / m planes of output row t - k3OmLead (3 planes before they are used)
            for (int t = 0; t < nplanes + k3OmLead; ++t) {
                if (do_u && t < nplanes) {
                    const int slot = t % k3NP;
                    if (t >= k3NP) mbar_wait(&empty_bar[slot], ((t / k3NP) - 1) & 1);
                    mbar_expect_tx(&full_bar[slot], k3PlaneFloats * (uint32_t)sizeof(float));
                    tma_load_3d(ring + (size_t)slot * k3PlaneFloats, &tm_cur, x0 - kHalo, y0 - kHalo, zc0 - kHalo + t, &full_bar[slot]);
                }
                const int j = t - k3OmLead;
                if (do_om && j >= 0 && j < nout) {
                    const int slot = j % k3NO;
                    if (j >= k3NO) mbar_wait(&om_empty[slot], ((j / k3NO) - 1) & 1);
                    float* dst = om_ring + (size_t)slot * 2 * k3OmFloats;
                    mbar_expect_tx(&om_full[slot], 2 * k3OmFloats * (uint32_t)sizeof(float));
                    tma_load_3d(dst, &tm_old, x0, y0, zc0 + j, &om_full[slot]);
                    tma_load_3d(dst + k3OmFloats, &tm_m, x0, y0, zc0 + j, &om_full[slot]);
                }
            }
        }
    } else {
        // ---------------- consumer warps
        const int x = x0 + 4 * lane;
        const int yl = k3RPW * warp;                   // first of this warp's rows inside the tile
        const bool col_ok = x < a.px;
        float4 gx4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (col_ok) gx4 = ld4(a.gx + x);
        float gyv[k3RPW];
#pragma unroll
        for (int r = 0; r < k3RPW; ++r) gyv[r] = (y0 + yl + r < a.ny) ? __ldg(a.gy + y0 + yl + r) : 1.f;

        float4 win[k3RPW][9];
        auto own = [&](int p, int r) {                 // this lane's element of row r in ring plane p
            return ld4(ring + (size_t)(p % k3NP) * k3PlaneFloats + (yl + r + kHalo) * k3SX + kHalo + 4 * lane);
        };
        // prime with planes 0..7 (z = zc0-4 .. zc0+3); planes 0..3 are never centre planes -> release them
        for (int p = 0; p < 2 * kHalo; ++p) {
            mbar_wait(&full_bar[p % k3NP], (p / k3NP) & 1);
#pragma unroll
            for (int r = 0; r < k3RPW; ++r) win[r][p + 1] = own(p, r);
            if (p < kHalo) {
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[p % k3NP])) : "memory");
            }
        }
        for (int iz = 0; iz < nout; ++iz) {
            const int z = zc0 + iz;
            const int ptop = iz + 2 * kHalo, pmid = iz + kHalo;
            // global operands first: their latency overlaps the barrier wait and the shared-memory reads
            float4 o4[k3RPW], m4[k3RPW], s4[k3RPW], c4[k3RPW], sp4[k3RPW];
            size_t off[k3RPW];
            bool ok[k3RPW];
#pragma unroll
            for (int r = 0; r < k3RPW; ++r) {
                const int y = y0 + yl + r;
                ok[r] = col_ok && y < a.ny;
                off[r] = ((size_t)z * a.ny + y) * a.px + x;
                if (ok[r]) {
                    if (MODE == STEP_ADJ || MODE == STEP_ADJ2) { s4[r] = ld4_stream(a.snap + off[r]); c4[r] = ld4(a.acc + off[r]); }
                    if (MODE == STEP_ADJ2) sp4[r] = ld4_stream(a.snap_prev + off[r]);
                }
            }
            const float gzv = __ldg(a.gz + z);
            mbar_wait(&full_bar[ptop % k3NP], (ptop / k3NP) & 1);
            mbar_wait(&om_full[iz % k3NO], (iz / k3NO) & 1);
            {
                const float* om = om_ring + (size_t)(iz % k3NO) * 2 * k3OmFloats + (yl * k3BX) + 4 * lane;
#pragma unroll
                for (int r = 0; r < k3RPW; ++r) { o4[r] = ld4(om + r * k3BX); m4[r] = ld4(om + k3OmFloats + r * k3BX); }
            }
#pragma unroll
            for (int r = 0; r < k3RPW; ++r) {
#pragma unroll
                for (int k = 0; k < 8; ++k) win[r][k] = win[r][k + 1];
                win[r][8] = own(ptop, r);
            }
            const float* mid = ring + (size_t)(pmid % k3NP) * k3PlaneFloats;
            float4 yc[8 + k3RPW];                        // rows yl-4 .. yl+3+RPW of the centre plane, this lane's float4
#pragma unroll
            for (int k = 0; k < 8 + k3RPW; ++k) yc[k] = ld4(mid + (yl + k) * k3SX + kHalo + 4 * lane);
#pragma unroll
            for (int r = 0; r < k3RPW; ++r) {
                const float* row = mid + (yl + r + kHalo) * k3SX + 4 * lane;
                const float4 L = ld4(row), R = ld4(row + 8);
                const float4 C = yc[r + 4];
                const float ax[12] = {L.x, L.y, L.z, L.w, C.x, C.y, C.z, C.w, R.x, R.y, R.z, R.w};
                const float gg = gyv[r] * gzv;
                const float gv[4] = {gx4.x * gg, gx4.y * gg, gx4.z * gg, gx4.w * gg};
                const float ov[4] = {o4[r].x, o4[r].y, o4[r].z, o4[r].w}, mv[4] = {m4[r].x, m4[r].y, m4[r].z, m4[r].w};
                float wv[4], nv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float c = ax[4 + q];
                    auto comp = [&](const float4& v) { return q == 0 ? v.x : (q == 1 ? v.y : (q == 2 ? v.z : v.w)); };
                    float lap = (3.0f * kC0) * c;
                    lap = fmaf(kC1, ((ax[3 + q] + ax[5 + q]) + (comp(yc[r + 3]) + comp(yc[r + 5]))) + (comp(win[r][3]) + comp(win[r][5])), lap);
                    lap = fmaf(kC2, ((ax[2 + q] + ax[6 + q]) + (comp(yc[r + 2]) + comp(yc[r + 6]))) + (comp(win[r][2]) + comp(win[r][6])), lap);
                    lap = fmaf(kC3, ((ax[1 + q] + ax[7 + q]) + (comp(yc[r + 1]) + comp(yc[r + 7]))) + (comp(win[r][1]) + comp(win[r][7])), lap);
                    lap = fmaf(kC4, ((ax[0 + q] + ax[8 + q]) + (comp(yc[r + 0]) + comp(yc[r + 8]))) + (comp(win[r][0]) + comp(win[r][8])), lap);
                    wv[q] = lap;
                    nv[q] = gv[q] * fmaf(mv[q], lap, fmaf(-gv[q], ov[q], 2.0f * c));
                }
                if (ok[r]) {
                    st4(a.oldnew + off[r], make_float4(nv[0], nv[1], nv[2], nv[3]));
                    if (MODE == STEP_FWD_SAVE) st4_stream(a.snap + off[r], make_float4(wv[0], wv[1], wv[2], wv[3]));
                    if (MODE == STEP_ADJ || MODE == STEP_ADJ2) {
                        float4 c = c4[r];
                        if (MODE == STEP_ADJ2) {      // deferred imaging of the previous adjoint step (its field is this launch's u_n)
                            c.x = fmaf(C.x, sp4[r].x, c.x); c.y = fmaf(C.y, sp4[r].y, c.y);
                            c.z = fmaf(C.z, sp4[r].z, c.z); c.w = fmaf(C.w, sp4[r].w, c.w);
                        }
                        c.x = fmaf(nv[0], s4[r].x, c.x); c.y = fmaf(nv[1], s4[r].y, c.y);
                        c.z = fmaf(nv[2], s4[r].z, c.z); c.w = fmaf(nv[3], s4[r].w, c.w);
                        st4(a.acc + off[r], c);
                    }
                }
            }
            // the centre plane is dead now (later outputs see it only through the register windows)
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[pmid % k3NP])) : "memory");
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&om_empty[iz % k3NO])) : "memory");
            }
        }
    }

    // ---- sparse fix-ups for the points this CTA owns: injection, then receiver sampling -------------------
    const int tid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const int i0 = a.inj.tile_ptr ? a.inj.tile_ptr[tid] : 0, i1 = a.inj.tile_ptr ? a.inj.tile_ptr[tid + 1] : 0;
    const int r0 = a.rec.tile_ptr ? a.rec.tile_ptr[tid] : 0, r1 = a.rec.tile_ptr ? a.rec.tile_ptr[tid + 1] : 0;
    if (i1 > i0 || r1 > r0) {
        __syncthreads();
        for (int e = i0 + threadIdx.x; e < i1; e += blockDim.x) {
            const int off = a.inj.off[e];
            const int zy = off / a.px, xx = off - zy * a.px;
            const int z = zy / a.ny, y = zy - z * a.ny;
            const float val = a.inj_vals[a.inj.id[e]];
            const float gm = a.gx[xx] * a.gy[y] * a.gz[z] * a.m[off];
            atomicAdd(a.oldnew + off, gm * val);
            if (MODE == STEP_FWD_SAVE) atomicAdd(a.snap + off, val);
            if (MODE == STEP_ADJ || MODE == STEP_ADJ2) atomicAdd(a.acc + off, gm * val * a.snap[off]);
        }
        if (r1 > r0) {
            __syncthreads();
            for (int e = r0 + threadIdx.x; e < r1; e += blockDim.x) a.rec_out[a.rec.id[e]] = __ldcg(a.oldnew + a.rec.off[e]);
        }
    }

    // ---- slab mode: push this CTA's share of the 4 boundary planes into the neighbours' ghost planes over NVLink
    //      (plain stores to peer-mapped pointers), then the last CTA of the grid publishes the step id -------------
    if (a.peer_up || a.peer_dn) {
        __syncthreads();                                   // dense stores and fix-ups of this CTA are done
        bool pushed = false;
        for (int side = 0; side < 2; ++side) {
            float* peer = side == 0 ? a.peer_up : a.peer_dn;
            if (!peer) continue;
            const int zb0 = side == 0 ? a.z_own0 : a.z_own1 - kHalo;            // my 4 boundary planes
            const int zdst0 = side == 0 ? a.peer_up_z : 0;                      // where they land in the neighbour's grid
            for (int z = max(zb0, zc0); z < min(zb0 + kHalo, zc1); ++z) {
                pushed = true;
                for (int i = threadIdx.x; i < k3BY * (k3BX / 4); i += blockDim.x) {
                    const int yy = y0 + i / (k3BX / 4), xx = x0 + 4 * (i % (k3BX / 4));
                    if (yy < a.ny && xx < a.px) {
                        const float4 v = __ldcg(reinterpret_cast<const float4*>(a.oldnew + ((size_t)z * a.ny + yy) * a.px + xx));
                        *reinterpret_cast<float4*>(peer + ((size_t)(zdst0 + z - zb0) * a.ny + yy) * a.px + xx) = v;
                    }
                }
            }
        }
        if (pushed) __threadfence_system();                // block-uniform: make the peer stores visible before counting in
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned total = gridDim.x * gridDim.y * gridDim.z;
            if (atomicAdd(a.done_counter, 1u) == total - 1) {
                *a.done_counter = 0;
                __threadfence_system();
                if (a.flag_peer_up) st_release_sys(a.flag_peer_up, a.signal_id);
                if (a.flag_peer_dn) st_release_sys(a.flag_peer_dn, a.signal_id);
            }
        }
    }
}

}  // namespace fwi
