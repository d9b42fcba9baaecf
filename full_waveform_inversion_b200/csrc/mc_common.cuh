// Shared between the CUDA-core (mc_kernels.cu) and tensor-core (mc_umma.cu) likelihood kernels of Track A.
#pragma once
#include "common.cuh"

namespace fwi {

struct TraceConst {
    double mean_d;   // mean of the trace's data (the centring constant used for d')
    double ssd;      // sum (d - mean_d)^2
    double sumd2;    // sum d^2
    double maxd;     // max |d| of the ORIGINAL trace (normalisation, FWI:598)
    double sigma;    // mean |d[-60:-10]| (FWI:580), NaN when T < 60
    double d_first, d_last;
};

struct FlatConst {   // constants of the flattened (K*Tv) data array; index = normalised?
    double n;
    double D1[2], D2[2];   // sum d, sum d^2
    double sigma[2];       // gaussian noise level of the flattened array
};

enum { MODE_SSE = 0, MODE_MOM = 1, MODE_MOM_MAX = 2 };

// The tensor-core evaluation path (mc_umma.cu), attached to a Monte-Carlo context at upload when the shapes allow it.
struct UmmaPath;
int umma_build(UmmaPath** out, int device, const double* G, const double* d, int K, int C, int T, const TraceConst* tc_dev,
               const FlatConst& fc);
void umma_free(UmmaPath* u);
bool umma_supports(const UmmaPath* u, int metric, int flags);
// same contract as fwi_mc_eval for one medium: M (rows, N) with leading dimension ldm, similarity and likelihood out
int umma_eval(UmmaPath* u, const float* M, int64_t ldm, int64_t N, int metric, int flags, float* sim, float* like, cudaStream_t st);

}  // namespace fwi
