/* C restatement of oracle/fd_oracle.py (Track B self-oracle), 2-D and 3-D, float32 or float64.
 *
 * TEST INFRASTRUCTURE ONLY (tests/, tools/, bench.py's CPU legs).  The reference repository has no propagator
 * (SURVEY 0), so this is a port of the SELF-oracle, not of reference code: PARITY UNPINNED BY THE REFERENCE.
 * tests/test_fd_oracle.py pins it to the NumPy oracle (float64 build: 1e-12) and the NumPy oracle to analytic
 * solutions (tests/test_fd_analytic.py).
 *
 * One source, two builds (oracle/Makefile): `real` = float  -> _build/libfd_oracle.so    (the CPU baseline of bench.py)
 *                                           `real` = double -> _build/libfd_oracle64.so  (the arbiter for the GPU parity
 *                                                              tests at benchmark sizes; -DFDC_DOUBLE)
 * Specification (fd_oracle.py B1-B3):  w_n = lap8(u_n) + f_n ;  u_{n+1} = g (2 u_n - g u_{n-1} + m w_n) ;
 * trace[n] = u_{n+1}[rec] ;  adjoint = the same step on the time-reversed residual ;  I = sum_n q_n w_{n-1}.
 * POSIX threads over z rows / z planes (the image has no libgomp).  The gradient can hold every w_n or use
 * two-level checkpointing (segments of `seg` steps: state pairs at the segment starts, w_n recomputed per segment),
 * which gives the same numbers and lets a 5000-step shot on 1000 x 3000 run in a few GB of host memory.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#ifdef FDC_DOUBLE
typedef double real;
#else
typedef float real;
#endif

#define H4 4
static const real C0 = (real)(-205.0 / 72.0), C1 = (real)(8.0 / 5.0), C2 = (real)(-1.0 / 5.0), C3 = (real)(8.0 / 315.0),
                  C4 = (real)(-1.0 / 560.0);

static int g_threads = 0;
int fdc_num_threads(void) {
    if (g_threads <= 0) {
        const char* e = getenv("FDC_THREADS");
        long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
        g_threads = (int)(n < 1 ? 1 : (n > 256 ? 256 : n));
    }
    return g_threads;
}
void fdc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int fdc_real_size(void) { return (int)sizeof(real); }

/* ---- grid: fields carry a zero ghost ring of 4 (no y ring for 2-D grids, ny == 1) ------------------------- */
typedef struct {
    int nz, ny, nx, hy;          /* hy = 4 for 3-D, 0 for 2-D */
    size_t px, pxy, nf;          /* padded row length, padded plane size, padded field size */
    size_t n;                    /* nz * ny * nx */
    real *m, *gz, *gy, *gx;
} grid_t;

static size_t pidx(const grid_t* g, int z, int y, int x) { return (size_t)(z + H4) * g->pxy + (size_t)(y + g->hy) * g->px + (size_t)(x + H4); }

static void profile(real* prof, int n, int nabs, double alpha) {
    for (int i = 0; i < n; ++i) prof[i] = 1;
    for (int i = 0; i < nabs && i < n; ++i) {
        const double t = alpha * (nabs - i) / nabs;
        const real v = (real)exp(-t * t);
        if (v < prof[i]) prof[i] = v;
        if (v < prof[n - 1 - i]) prof[n - 1 - i] = v;
    }
}

static int grid_init(grid_t* g, const real* v, int nz, int ny, int nx, double h, double dt, int nabs, double alpha) {
    g->nz = nz; g->ny = ny; g->nx = nx; g->hy = ny > 1 ? H4 : 0;
    g->px = (size_t)nx + 2 * H4; g->pxy = g->px * (size_t)(ny + 2 * g->hy); g->nf = g->pxy * (size_t)(nz + 2 * H4);
    g->n = (size_t)nz * ny * nx;
    g->m = malloc(g->n * sizeof(real)); g->gz = malloc(nz * sizeof(real)); g->gy = malloc(ny * sizeof(real)); g->gx = malloc(nx * sizeof(real));
    if (!g->m || !g->gz || !g->gy || !g->gx) return -1;
    const real r = (real)dt / (real)h;                      /* same rounding as the device: c = v * (dt/h); m = c * c */
    for (size_t i = 0; i < g->n; ++i) { const real c = v[i] * r; g->m[i] = c * c; }
    profile(g->gz, nz, nabs, alpha); profile(g->gx, nx, nabs, alpha);
    if (ny > 1) profile(g->gy, ny, nabs, alpha); else g->gy[0] = 1;
    return 0;
}
static void grid_free(grid_t* g) { free(g->m); free(g->gz); free(g->gy); free(g->gx); }

/* dense part of one step on z rows/planes [za, zb): oldnew <- u_{n+1}; optionally w_out <- lap, img += u_{n+1} * w_in */
static void step_dense(const grid_t* g, const real* cur, real* oldnew, int za, int zb, real* w_out, const real* w_in, real* img) {
    const int nx = g->nx, ny = g->ny;
    const long p = (long)g->px, q = (long)g->pxy;
    const int three = ny > 1;
    for (int z = za; z < zb; ++z)
        for (int y = 0; y < ny; ++y) {
            const real* c = cur + pidx(g, z, y, 0);
            real* o = oldnew + pidx(g, z, y, 0);
            const size_t d0 = ((size_t)z * ny + y) * nx;
            const real* mm = g->m + d0;
            const real gzy = g->gz[z] * g->gy[y];
            for (int x = 0; x < nx; ++x) {
                real lap;
                if (three)
                    lap = 3 * C0 * c[x]
                        + C1 * (c[x - 1] + c[x + 1] + c[x - p] + c[x + p] + c[x - q] + c[x + q])
                        + C2 * (c[x - 2] + c[x + 2] + c[x - 2 * p] + c[x + 2 * p] + c[x - 2 * q] + c[x + 2 * q])
                        + C3 * (c[x - 3] + c[x + 3] + c[x - 3 * p] + c[x + 3 * p] + c[x - 3 * q] + c[x + 3 * q])
                        + C4 * (c[x - 4] + c[x + 4] + c[x - 4 * p] + c[x + 4 * p] + c[x - 4 * q] + c[x + 4 * q]);
                else
                    lap = 2 * C0 * c[x]
                        + C1 * (c[x - 1] + c[x + 1] + c[x - q] + c[x + q])
                        + C2 * (c[x - 2] + c[x + 2] + c[x - 2 * q] + c[x + 2 * q])
                        + C3 * (c[x - 3] + c[x + 3] + c[x - 3 * q] + c[x + 3 * q])
                        + C4 * (c[x - 4] + c[x + 4] + c[x - 4 * q] + c[x + 4 * q]);
                const real gg = gzy * g->gx[x];
                const real nv = gg * (2 * c[x] - gg * o[x] + mm[x] * lap);
                o[x] = nv;
                if (w_out) w_out[d0 + x] = lap;
                if (w_in) img[d0 + x] += nv * w_in[d0 + x];
            }
        }
}

/* ---- a threaded run of consecutive steps ------------------------------------------------------------------ */
typedef struct {
    const grid_t* g;
    int n0, n1, adjoint;           /* forward: n = n0 .. n1-1; adjoint: n = n1-1 .. n0 */
    real **cur, **old;             /* state pair (swapped in place; on return *cur = newest field) */
    int ninj; const int *iz, *iy, *ix; const real* inj; int inj_stride;     /* values inj[n * stride + q] */
    int nrec; const int *rz, *ry, *rx; real* traces;                        /* forward only: traces[n * nrec + r] */
    real* ws; int ws_base;         /* snapshot of step n lives at ws + (n - ws_base) * g->n; NULL = none */
    real* img;
} run_t;

typedef struct { int tid, nthreads; run_t* r; pthread_barrier_t* bar; } job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    run_t* r = j->r;
    const grid_t* g = r->g;
    const int za = (int)((long)g->nz * j->tid / j->nthreads), zb = (int)((long)g->nz * (j->tid + 1) / j->nthreads);
    real *cur = *r->cur, *old = *r->old;
    const int steps = r->n1 - r->n0;
    for (int s = 0; s < steps; ++s) {
        const int n = r->adjoint ? r->n1 - 1 - s : r->n0 + s;
        real* w = r->ws ? r->ws + (size_t)(n - r->ws_base) * g->n : NULL;
        step_dense(g, cur, old, za, zb, r->adjoint ? NULL : w, r->adjoint ? w : NULL, r->img);
        pthread_barrier_wait(j->bar);
        if (j->tid == 0) {                  /* sparse part: injection (f_n enters w_n), then receiver sampling */
            for (int q = 0; q < r->ninj; ++q) {
                const int z = r->iz[q], y = r->iy ? r->iy[q] : 0, x = r->ix[q];
                const size_t i = ((size_t)z * g->ny + y) * g->nx + x;
                const real f = r->inj[(size_t)n * r->inj_stride + q];
                const real add = g->gz[z] * g->gy[y] * g->gx[x] * g->m[i] * f;
                old[pidx(g, z, y, x)] += add;
                if (!r->adjoint && w) w[i] += f;
                if (r->adjoint) r->img[i] += add * w[i];
            }
            if (r->traces)
                for (int k = 0; k < r->nrec; ++k)
                    r->traces[(size_t)n * r->nrec + k] = old[pidx(g, r->rz[k], r->ry ? r->ry[k] : 0, r->rx[k])];
        }
        pthread_barrier_wait(j->bar);
        real* t = cur; cur = old; old = t;
    }
    if (j->tid == 0) { *r->cur = cur; *r->old = old; }
    return NULL;
}

static int run_steps(run_t* r) {
    if (r->n1 <= r->n0) return 0;
    int T = fdc_num_threads();
    if (T > r->g->nz) T = r->g->nz;
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, T);
    pthread_t* th = malloc(sizeof(pthread_t) * T);
    job_t* jobs = malloc(sizeof(job_t) * T);
    if (!th || !jobs) return -1;
    for (int t = 0; t < T; ++t) {
        jobs[t].tid = t; jobs[t].nthreads = T; jobs[t].r = r; jobs[t].bar = &bar;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
    pthread_barrier_destroy(&bar);
    free(th); free(jobs);
    return 0;
}

/* forward: traces[nt][nrec]; ws (nullable) receives every w_n ([nt][nz*ny*nx]); state_out (nullable) receives the
 * final (u_nt, u_{nt-1}) as two dense grids */
int fdc_forward(const real* v, int nz, int ny, int nx, double h, double dt, int nabs, double alpha, int nsrc, const int* sz,
                const int* sy, const int* sx, int nrec, const int* rz, const int* ry, const int* rx, const real* wavelet, int nt,
                real* traces, real* ws, real* state_out) {
    grid_t g;
    if (grid_init(&g, v, nz, ny, nx, h, dt, nabs, alpha)) return -1;
    real* a = calloc(g.nf, sizeof(real)); real* b = calloc(g.nf, sizeof(real));
    if (!a || !b) return -1;
    real *cur = a, *old = b;
    run_t r = {&g, 0, nt, 0, &cur, &old, nsrc, sz, ny > 1 ? sy : NULL, sx, wavelet, nsrc, nrec, rz, ny > 1 ? ry : NULL, rx, traces, ws, 0, NULL};
    int rc = run_steps(&r);
    if (state_out && !rc)
        for (int z = 0; z < nz; ++z)
            for (int y = 0; y < ny; ++y) {
                memcpy(state_out + ((size_t)z * ny + y) * nx, cur + pidx(&g, z, y, 0), nx * sizeof(real));
                memcpy(state_out + g.n + ((size_t)z * ny + y) * nx, old + pidx(&g, z, y, 0), nx * sizeof(real));
            }
    free(a); free(b); grid_free(&g);
    return rc;
}

/* misfit J = 1/2 sum (traces - obs)^2 (float64 accumulation), imaging sum img = sum_n q_n w_{n-1}, traces.
 * seg <= 0 or seg >= nt: every w_n is held; otherwise two-level checkpointing with segments of `seg` steps. */
int fdc_gradient(const real* v, int nz, int ny, int nx, double h, double dt, int nabs, double alpha, int nsrc, const int* sz,
                 const int* sy, const int* sx, int nrec, const int* rz, const int* ry, const int* rx, const real* wavelet,
                 const real* obs, int nt, int seg, real* traces, real* img, double* J_out) {
    grid_t g;
    if (grid_init(&g, v, nz, ny, nx, h, dt, nabs, alpha)) return -1;
    if (seg <= 0 || seg > nt) seg = nt;
    const int nseg = (nt + seg - 1) / seg;
    real* fa = calloc(g.nf, sizeof(real)); real* fb = calloc(g.nf, sizeof(real));
    real* qa = calloc(g.nf, sizeof(real)); real* qb = calloc(g.nf, sizeof(real));
    real* ws = malloc((size_t)seg * g.n * sizeof(real));
    real* ck = nseg > 1 ? malloc((size_t)nseg * 2 * g.nf * sizeof(real)) : NULL;
    real* res = malloc((size_t)nt * nrec * sizeof(real));
    if (!fa || !fb || !qa || !qb || !ws || !res || (nseg > 1 && !ck)) return -1;
    memset(img, 0, g.n * sizeof(real));
    const int* sy_ = ny > 1 ? sy : NULL; const int* ry_ = ny > 1 ? ry : NULL;
    real *cur = fa, *old = fb;
    int rc = 0;
    for (int s = 0; s < nseg && !rc; ++s) {
        if (nseg > 1) { memcpy(ck + (size_t)(2 * s) * g.nf, cur, g.nf * sizeof(real)); memcpy(ck + (size_t)(2 * s + 1) * g.nf, old, g.nf * sizeof(real)); }
        const int n0 = s * seg, n1 = (s + 1) * seg < nt ? (s + 1) * seg : nt;
        run_t r = {&g, n0, n1, 0, &cur, &old, nsrc, sz, sy_, sx, wavelet, nsrc, nrec, rz, ry_, rx, traces, nseg == 1 ? ws : NULL, 0, NULL};
        rc = run_steps(&r);
    }
    double J = 0.0;
    for (size_t i = 0; i < (size_t)nt * nrec; ++i) { res[i] = traces[i] - obs[i]; J += (double)res[i] * (double)res[i]; }
    *J_out = 0.5 * J;
    real *qc = qa, *qo = qb;
    for (int s = nseg - 1; s >= 0 && !rc; --s) {
        const int n0 = s * seg, n1 = (s + 1) * seg < nt ? (s + 1) * seg : nt;
        if (nseg > 1) {
            memcpy(fa, ck + (size_t)(2 * s) * g.nf, g.nf * sizeof(real)); memcpy(fb, ck + (size_t)(2 * s + 1) * g.nf, g.nf * sizeof(real));
            cur = fa; old = fb;
            run_t r = {&g, n0, n1, 0, &cur, &old, nsrc, sz, sy_, sx, wavelet, nsrc, 0, NULL, NULL, NULL, NULL, ws, n0, NULL};
            rc = run_steps(&r);
            if (rc) break;
        }
        run_t r = {&g, n0, n1, 1, &qc, &qo, nrec, rz, ry_, rx, res, nrec, 0, NULL, NULL, NULL, NULL, ws, nseg > 1 ? n0 : 0, img};
        rc = run_steps(&r);
    }
    free(fa); free(fb); free(qa); free(qb); free(ws); free(ck); free(res); grid_free(&g);
    return rc;
}
