/* C restatement of oracle/fd_oracle.py (Track B self-oracle), for the CPU baseline of bench.py.
 *
 * TEST INFRASTRUCTURE ONLY.  The reference repository has no propagator, so this is a port of the
 * SELF-oracle, not of reference code; tests/test_fd_oracle.py checks it against the NumPy oracle.
 * float32 fields, POSIX threads over row slabs (the image has no libgomp).  Same specification: u+ = g (2u - g u- + m (lap8(u) + f)), spec B1-B3.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

static const float C0 = -205.0f / 72.0f, C1 = 8.0f / 5.0f, C2 = -1.0f / 5.0f, C3 = 8.0f / 315.0f, C4 = -1.0f / 560.0f;

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

/* fields are (nz+8) x (nx+8) with a zero ghost ring of 4; pitch = nx + 8 */
static void step2d(const float* cur, float* oldnew, const float* m, const float* gz, const float* gx, int za, int zb, int nx,
                   float* w_out, const float* w_in, float* img) {
    const int p = nx + 8;
    for (int z = za; z < zb; ++z) {
        const float* c = cur + (size_t)(z + 4) * p + 4;
        float* o = oldnew + (size_t)(z + 4) * p + 4;
        const float* mm = m + (size_t)z * nx;
        for (int x = 0; x < nx; ++x) {
            float lap = 2.0f * C0 * c[x]
                + C1 * (c[x - 1] + c[x + 1] + c[x - p] + c[x + p])
                + C2 * (c[x - 2] + c[x + 2] + c[x - 2 * p] + c[x + 2 * p])
                + C3 * (c[x - 3] + c[x + 3] + c[x - 3 * p] + c[x + 3 * p])
                + C4 * (c[x - 4] + c[x + 4] + c[x - 4 * p] + c[x + 4 * p]);
            const float g = gz[z] * gx[x];
            const float nv = g * (2.0f * c[x] - g * o[x] + mm[x] * lap);
            o[x] = nv;
            if (w_out) w_out[(size_t)z * nx + x] = lap;
            if (w_in) img[(size_t)z * nx + x] += nv * w_in[(size_t)z * nx + x];
        }
    }
}

static void profile(float* prof, int n, int nabs, double alpha) {
    for (int i = 0; i < n; ++i) prof[i] = 1.0f;
    for (int i = 0; i < nabs && i < n; ++i) {
        const double t = alpha * (nabs - i) / nabs;
        const float v = (float)exp(-t * t);
        if (v < prof[i]) prof[i] = v;
        if (v < prof[n - 1 - i]) prof[n - 1 - i] = v;
    }
}


typedef struct {
    int tid, nthreads, nz, nx, nt, npts_inj, npts_rec, adjoint;
    float *a, *b;
    const float *m, *gz, *gx;
    const int *iz, *ix, *rz, *rx;
    const float* inj;       /* [nt][npts_inj] */
    float* traces;          /* [nt][npts_rec] or NULL */
    float* ws;              /* snapshots: written (forward) or read (adjoint); may be NULL for forward */
    float* img;
    pthread_barrier_t* bar;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    const int nz = j->nz, nx = j->nx, p = nx + 8;
    const int za = (int)((long)nz * j->tid / j->nthreads), zb = (int)((long)nz * (j->tid + 1) / j->nthreads);
    float *cur = j->a, *old = j->b;
    for (int s = 0; s < j->nt; ++s) {
        const int n = j->adjoint ? j->nt - 1 - s : s;
        float* w = j->ws ? j->ws + (size_t)n * nz * nx : NULL;
        step2d(cur, old, j->m, j->gz, j->gx, za, zb, nx, j->adjoint ? NULL : w, j->adjoint ? w : NULL, j->img);
        pthread_barrier_wait(j->bar);
        if (j->tid == 0) {
            for (int q = 0; q < j->npts_inj; ++q) {
                const int z = j->iz[q], x = j->ix[q];
                const size_t i = (size_t)z * nx + x;
                const float f = j->inj[(size_t)n * j->npts_inj + q];
                const float add = j->gz[z] * j->gx[x] * j->m[i] * f;
                old[(size_t)(z + 4) * p + x + 4] += add;
                if (!j->adjoint && w) w[i] += f;
                if (j->adjoint) j->img[i] += add * w[i];
            }
            if (j->traces)
                for (int r = 0; r < j->npts_rec; ++r) j->traces[(size_t)n * j->npts_rec + r] = old[(size_t)(j->rz[r] + 4) * p + j->rx[r] + 4];
        }
        pthread_barrier_wait(j->bar);
        float* t = cur; cur = old; old = t;
    }
    return NULL;
}

static int run_loop(const float* v, int nz, int nx, float h, float dt, int nabs, float alpha, int adjoint, int ninj,
                    const int* iz, const int* ix, const float* inj, int nrec, const int* rz, const int* rx, int nt,
                    float* traces, float* ws, float* img) {
    const int p = nx + 8;
    const size_t nf = (size_t)(nz + 8) * p;
    float* a = calloc(nf, 4); float* b = calloc(nf, 4);
    float* m = malloc((size_t)nz * nx * 4); float* gz = malloc(nz * 4); float* gx = malloc(nx * 4);
    if (!a || !b || !m || !gz || !gx) return -1;
    for (size_t i = 0; i < (size_t)nz * nx; ++i) { const float c = v[i] * (dt / h); m[i] = c * c; }
    profile(gz, nz, nabs, alpha); profile(gx, nx, nabs, alpha);
    int T = fdc_num_threads();
    if (T > nz) T = nz;
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, T);
    pthread_t* th = malloc(sizeof(pthread_t) * T);
    job_t* jobs = malloc(sizeof(job_t) * T);
    for (int t = 0; t < T; ++t) {
        job_t j = {t, T, nz, nx, nt, ninj, nrec, adjoint, a, b, m, gz, gx, iz, ix, rz, rx, inj, traces, ws, img, &bar};
        jobs[t] = j;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
    pthread_barrier_destroy(&bar);
    free(th); free(jobs); free(a); free(b); free(m); free(gz); free(gx);
    return 0;
}

/* forward (optionally saving w_n into ws[nt][nz*nx]) and recording traces[nt][nrec] */
int fdc_forward(const float* v, int nz, int nx, float h, float dt, int nabs, float alpha, int nsrc, const int* sz,
                const int* sx, int nrec, const int* rz, const int* rx, const float* wavelet, int nt, float* traces, float* ws) {
    return run_loop(v, nz, nx, h, dt, nabs, alpha, 0, nsrc, sz, sx, wavelet, nrec, rz, rx, nt, traces, ws, NULL);
}

/* adjoint with imaging: img[nz*nx] = sum_n q_n * w_{n-1} */
int fdc_adjoint(const float* v, int nz, int nx, float h, float dt, int nabs, float alpha, int nrec, const int* rz,
                const int* rx, const float* resid, int nt, const float* ws, float* img) {
    memset(img, 0, (size_t)nz * nx * 4);
    return run_loop(v, nz, nx, h, dt, nabs, alpha, 1, nrec, rz, rx, resid, 0, NULL, NULL, nt, NULL, (float*)ws, img);
}
