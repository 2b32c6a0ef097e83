/* Oracle: exact FP64 brute-force 1-nearest-neighbour search.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference (LCJebe/PCReg) contains no nearest-neighbour correspondence code (SURVEY.md
 * section 0 / 8 a0: `knnsearch` appears nowhere; the only literal 1-NN loop is
 * ColorCodeModel.m:15-18 via the closed Computer Vision Toolbox).  The semantics restated here
 * are those of MATLAB's documented knnsearch(X,Y,'K',1): Euclidean distance in double precision,
 * ties resolved to the smallest index.  PARITY UNPINNED (no reference code, test or vector).
 *
 * d2 = ((mx-qx)^2 + (my-qy)^2) + (mz-qz)^2 evaluated in exactly this order with no FMA
 * contraction (build with -ffp-contract=off), which the CUDA re-check mirrors with
 * __dmul_rn/__dadd_rn so that d2 is bit-identical on both sides.
 *
 * Build: see oracle/Makefile  ->  oracle/_build/liboracle_nn.so
 */
#include <stdint.h>
#include <math.h>

void oracle_nn_brute_f64(const double* mx, const double* my, const double* mz, int64_t nm,
                         const double* qx, const double* qy, const double* qz, int64_t nq,
                         int32_t* idx, double* d2out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; ++i) {
        const double x = qx[i], y = qy[i], z = qz[i];
        double best = INFINITY;
        int64_t bj = -1;
        for (int64_t j = 0; j < nm; ++j) {
            const double dx = mx[j] - x, dy = my[j] - y, dz = mz[j] - z;
            const double d = (dx * dx + dy * dy) + dz * dz;
            if (d < best) { best = d; bj = j; }     /* strict <  => smallest index on ties */
        }
        idx[i] = (int32_t)bj;
        d2out[i] = best;
    }
}

/* Second-best distance (excluding index `skip`), used by tests to decide where index parity is
 * REQUIRED (north_star: wherever best and second-best differ by more than 1e-9 relative). */
void oracle_nn_second_f64(const double* mx, const double* my, const double* mz, int64_t nm,
                          const double* qx, const double* qy, const double* qz, int64_t nq,
                          const int32_t* skip, double* d2second)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; ++i) {
        const double x = qx[i], y = qy[i], z = qz[i];
        double best = INFINITY;
        for (int64_t j = 0; j < nm; ++j) {
            if (j == skip[i]) continue;
            const double dx = mx[j] - x, dy = my[j] - y, dz = mz[j] - z;
            const double d = (dx * dx + dy * dy) + dz * dz;
            if (d < best) best = d;
        }
        d2second[i] = best;
    }
}
