/*
 * tests/integration/b200_harness.c -- TEST INFRASTRUCTURE.
 *
 * The drop-in, exercised from C with the REAL reference on both sides of the boundary:
 *   - matrix parsing (MMtoCSR / MMtoELL), the serial oracle (sgemvSerial), the comparator
 *     (doubleVectorsDiff) and the statistics (statsAvgVar) are the unmodified reference functions,
 *     linked from oracle/_ref/libspmv_ref.so;
 *   - the implementations under test are the SPMV-typed adapters of include/spmv_b200.h
 *     (b200SpMVRowsCSR, ...), i.e. exactly what a maintainer appends to SpmvCSRFuncs / SpmvELLFuncs
 *     (src/include/SpMV.h:144-159).
 * The loop restates the protocol of testSpMVImplOMP (test/SpMV_test.cu:67-101): AVG_TIMES_ITERATION timed
 * calls, result compared after EVERY call, mean/variance of wall and internal (= kernel) time, one
 * stdout line per implementation in the format scripts/parseLog.py expects (test/SpMV_test.cu:93-96).
 *
 * Build (here, where /root/reference exists):  make -C tests/integration
 * Run   (GPU box):  tests/integration/_build/b200_harness <matrix.mtx>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#include "sparseMatrix.h"
#include "SpMV.h"
#include "parser.h"
#include "utils.h"
#include "ompGetICV.h"
#include "spmv_b200.h" /* after the reference headers: defines the b200SpMV* adapters for THIS spmat layout */

#ifndef AVG_TIMES_ITERATION
#define AVG_TIMES_ITERATION 5
#endif

static CONFIG Conf = {.gridRows = 8, .gridCols = 8};

static int run_impl(const char* label, SPMV_INTERF f, spmat* mat, double* x, double* y, double* oracle_y, ulong rows) {
    double wall[AVG_TIMES_ITERATION], internal[AVG_TIMES_ITERATION], s_wall[2], s_int[2];
    for (unsigned i = 0; i < AVG_TIMES_ITERATION; i++) {
        memset(y, 0xFF, rows * sizeof(*y)); /* stale output can never pass */
        double t0 = omp_get_wtime();
        if (f(mat, x, &Conf, y)) {
            fprintf(stderr, "%s failed: %s\n", label, spmvb200_last_error());
            return EXIT_FAILURE;
        }
        wall[i] = omp_get_wtime() - t0;
        internal[i] = ElapsedInternal;
        ElapsedInternal = 0;
        for (ulong r = 0; r < rows; r++)
            if (y[r] != y[r]) { fprintf(stderr, "%s: NaN at row %lu\n", label, r); return EXIT_FAILURE; }
        if (doubleVectorsDiff(oracle_y, y, rows, NULL)) return EXIT_FAILURE;
    }
    statsAvgVar(wall, AVG_TIMES_ITERATION, s_wall);
    statsAvgVar(internal, AVG_TIMES_ITERATION, s_int);
    printf("@computing SpMV   with func: %s at:%p\n", label, (void*) f); /* "func:<id> at:<ptr>": scripts/parseLog.py:102 */
    printf("threadNum: %d\tompGridSize: %ux%u\ttimeAvg:%le timeVar:%le\ttimeInternalAvg:%le timeInternalVar:%le \n",
           omp_get_max_threads(), Conf.gridRows, Conf.gridCols, s_wall[0], s_wall[1], s_int[0], s_int[1]);
    return EXIT_SUCCESS;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s matrix.mtx\n", argv[0]); return EXIT_FAILURE; }
    int out = EXIT_FAILURE;
    spmat *csr = MMtoCSR(argv[1]), *ell = MMtoELL(argv[1]);
    if (!csr || !ell) { fprintf(stderr, "cannot parse %s\n", argv[1]); return EXIT_FAILURE; }
    double *x = malloc(csr->N * sizeof(*x)), *y = malloc(csr->M * sizeof(*y)), *oracle_y = malloc(csr->M * sizeof(*y));
    if (!x || !y || !oracle_y) return EXIT_FAILURE;
    spmvb200_synth_vector_host(0x5EED0077ull, 0, csr->N, 3e-5, x); /* finite, seeded; |x| < MAXRND (config.h:115) */
    sgemvSerial(csr, x, &Conf, oracle_y);
    /* the three header lines scripts/parseLog.py splits a matrix group into (test/SpMV_test.cu:255-257, src/commons/ompGetICV.c:43) */
    printf("#%s\n", argv[1]);
    printf("SpMV_OMP_test.c\tAVG_TIMES_ITERATION:%d\tsparse matrix: %lux%lu-%luNNZ-%ld=MAX_ROW_NZ\n", AVG_TIMES_ITERATION, csr->M,
           csr->N, csr->NZ, (long) ell->MAX_ROW_NZ);
    ompGetRuntimeSchedule(NULL);

    static const SPMV_INTERF csr_funcs[] = {&b200SpMVRowsCSR, &b200SpMVWarpPerRowCSR, &b200SpMVAdaptiveCSR, &b200SpMVRowsSELL};
    static const char* csr_names[] = {"B200 CSR 0 (b200SpMVRowsCSR)", "B200 CSR 1 (b200SpMVWarpPerRowCSR)", "B200 CSR 2 (b200SpMVAdaptiveCSR)",
                                      "B200 CSR 3 (b200SpMVRowsSELL)"};
    static const SPMV_INTERF ell_funcs[] = {&b200SpMVRowsELL, &b200SpMVRowsELLNNTransposed, &b200SpMVWarpsPerRowELLNTrasposed};
    static const char* ell_names[] = {"B200 ELL 0 (b200SpMVRowsELL)", "B200 ELL 1 (b200SpMVRowsELLNNTransposed)",
                                      "B200 ELL 2 (b200SpMVWarpsPerRowELLNTrasposed)"};
    for (unsigned f = 0; f < 4; f++)
        if (run_impl(csr_names[f], csr_funcs[f], csr, x, y, oracle_y, csr->M)) goto _free;
    for (unsigned f = 0; f < 3; f++)
        if (run_impl(ell_names[f], ell_funcs[f], ell, x, y, oracle_y, csr->M)) goto _free;
    /* the exact kinds must agree with the serial oracle to the last bit */
    b200SpMVRowsCSR(csr, x, &Conf, y);
    if (memcmp(y, oracle_y, csr->M * sizeof(*y))) { fprintf(stderr, "b200SpMVRowsCSR is not bit-identical to sgemvSerial\n"); goto _free; }
    b200SpMVRowsELL(ell, x, &Conf, y);
    if (memcmp(y, oracle_y, csr->M * sizeof(*y))) { fprintf(stderr, "b200SpMVRowsELL is not bit-identical to sgemvSerial\n"); goto _free; }
    printf("B200_HARNESS_OK\n");
    out = EXIT_SUCCESS;
_free:
    spmvb200_cache_drop(NULL);
    free(x); free(y); free(oracle_y);
    return out;
}
