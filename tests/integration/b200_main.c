/*
 * tests/integration/b200_main.c -- the reference's command-line driver (src/main.cu:69-283) with the B200 engine behind it.
 *
 *     b200_main <MatrixMarket_sparse_matrix_COO> <vectorFile | RNDVECT> [COMPUTE MODE] [--check]
 *
 * Same arguments, same mode-string matching (strEqual is a PREFIX match, src/include/macros.h:43: longer names are tested
 * before their prefixes, src/main.cu:103-121), same final stdout line `cmode:%d\telapsed:\t %le elapsedInternal %le`
 * (src/main.cu:268-269) and the same result dumps (OUTVECTORDUMP / OUTVECTORDUMPRAW, src/main.cu:264-267).
 *   OMP modes    CSR_ROWS, CSR_ROWS_GROUPS, CSR_TILES, CSR_TILES_ALLOCD, ELL_ROWS, ELL_ROWS_GROUPS, ELL_TILES: the unmodified
 *                reference functions (oracle/_ref/libspmv_ref.so), exactly as src/main.cu:123-131 maps them;
 *   B200 modes   appended after the reference's enum ("new modes must be appended"); each is the SPMV-typed
 *                adapter of include/spmv_b200.h, i.e. upload once + kernel + download through the C ABI:
 *                  B200_CSR_ROWS  B200_CSR_ROWS_WARP  B200_CSR_ADAPTIVE  B200_CSR_SELL  B200_CSR_XWINDOW
 *                  B200_ELL_ROWS  B200_ELL_ROWS_NN_TRANSPOSED  B200_ELL_ROWS_WARP_NN_TRANSPOSED
 *                The reference's own CUDA_* strings select the same engine kinds (CUDA_CSR_ROWS -> B200_CSR_ROWS, ...), so an
 *                existing script that passes CUDA_CSR_ROWS keeps working.
 * RNDVECT here is the seeded finite generator (the reference's /dev/urandom + sin(bits) vector contains NaNs).
 * --check: also run the reference's sgemvSerial and compare with the strict NaN-aware comparator of the C ABI
 *          (|dy| <= 1e-12 * sum|a x| per row) and with the reference's doubleVectorsDiff; non-zero exit on mismatch.
 * TEST INFRASTRUCTURE around the product's C ABI: built only where /root/reference exists (make -C tests/integration cli).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#include "sparseMatrix.h"
#include "SpMV.h"
#include "ompChunksDivide.h"
#include "parser.h"
#include "utils.h"
#include "spmv_b200.h" /* after the reference headers: defines the b200SpMV* adapters for THIS spmat layout */

#define RNDVECT "RNDVECT"
/* appended after the reference's COMPUTE_MODE enum (src/include/SpMV.h:43-59) */
enum {
    _B200_CSR_ROWS = _CUDA_ELL_ROWS_WARP_NT + 1,
    _B200_CSR_ROWS_WARP,
    _B200_CSR_ADAPTIVE,
    _B200_CSR_SELL,
    _B200_CSR_XWINDOW,
    _B200_ELL_ROWS,
    _B200_ELL_ROWS_NT,
    _B200_ELL_ROWS_WARP_NT,
};
static const struct {
    const char* name;
    int mode;
} MODES[] = {
    /* longer names before their prefixes */
    {"B200_CSR_ROWS_WARP", _B200_CSR_ROWS_WARP}, {"B200_CSR_ROWS", _B200_CSR_ROWS}, {"B200_CSR_ADAPTIVE", _B200_CSR_ADAPTIVE},
    {"B200_CSR_SELL", _B200_CSR_SELL}, {"B200_CSR_XWINDOW", _B200_CSR_XWINDOW},
    {"B200_ELL_ROWS_WARP_NN_TRANSPOSED", _B200_ELL_ROWS_WARP_NT}, {"B200_ELL_ROWS_NN_TRANSPOSED", _B200_ELL_ROWS_NT}, {"B200_ELL_ROWS", _B200_ELL_ROWS},
    /* the reference's CUDA strings (src/include/SpMV.h:37-41), same order of tests as src/main.cu:114-118 */
    {CUDA_CSR_ROWS_WARP, _B200_CSR_ROWS_WARP}, {CUDA_CSR_ROWS, _B200_CSR_ROWS}, {CUDA_ELL_ROWS_WARP_NT, _B200_ELL_ROWS_WARP_NT},
    {CUDA_ELL_ROWS_WARP, _B200_ELL_ROWS_WARP_NT}, {CUDA_ELL_ROWS, _B200_ELL_ROWS},
    /* OMP modes, src/main.cu:106-112 */
    {CSR_ROWS_GROUPS, _CSR_ROWS_GROUPS}, {CSR_ROWS, _CSR_ROWS}, {CSR_TILES_ALLOCD, _CSR_TILES_ALLOCD}, {CSR_TILES, _CSR_TILES},
    {ELL_ROWS_GROUPS, _ELL_ROWS_GROUPS}, {ELL_ROWS, _ELL_ROWS}, {ELL_TILES, _ELL_TILES},
};
#define HELP                                                                                                                     \
    "usage: MatrixMarket_sparse_matrix_COO, vectorFile || " RNDVECT ", [COMPUTE MODE] [--check]\n"                                \
    "\tOMP:\tCSR_ROWS,CSR_ROWS_GROUPS,CSR_TILES,CSR_TILES_ALLOCD,ELL_ROWS,ELL_ROWS_GROUPS,ELL_TILES\n"                             \
    "\tB200:\tB200_CSR_ROWS,B200_CSR_ROWS_WARP,B200_CSR_ADAPTIVE,B200_CSR_SELL,B200_CSR_XWINDOW,B200_ELL_ROWS,"                    \
    "B200_ELL_ROWS_NN_TRANSPOSED,B200_ELL_ROWS_WARP_NN_TRANSPOSED\n"                                                               \
    "\t(the reference's CUDA_CSR_ROWS, CUDA_CSR_ROWS_WARP, CUDA_ELL_ROWS, CUDA_ELL_ROWS_WARP_NN_TRANSPOSED select the same B200 kinds)\n"

static CONFIG Conf = {.gridRows = 8, .gridCols = 8};

int main(int argc, char** argv) {
    int out = EXIT_FAILURE, check = 0;
    if (argc < 3) { ERRPRINT(HELP); return out; }
    for (int i = 3; i < argc; i++)
        if (!strcmp(argv[i], "--check")) check = 1;
    Conf.threadNum = (uint) omp_get_max_threads();
    Conf.chunkDistrbFunc = (void*) chunksNOOP;
    int cmode = _CSR_ROWS, found = argc <= 3 || !strcmp(argv[3], "--check");
    for (unsigned i = 0; !found && i < sizeof(MODES) / sizeof(*MODES); i++)
        if (strEqual(argv[3], MODES[i].name)) { cmode = MODES[i].mode; found = 1; }
    if (!found) { ERRPRINT("INVALID COMPUTE_MODE ARGV[3] GIVEN\n" HELP); return out; }
    SPMV_INTERF func = NULL;
    int toCSR = 1;
    switch (cmode) { /* src/main.cu:123-137 */
        case _CSR_ROWS_GROUPS: func = &spmvRowsBlocksCSR; break;
        case _CSR_TILES: func = &spmvTilesCSR; break;
        case _CSR_TILES_ALLOCD: func = &spmvTilesAllocdCSR; break;
        case _CSR_ROWS: func = &spmvRowsBasicCSR; break;
        case _ELL_ROWS: func = &spmvRowsBasicELL; toCSR = 0; break;
        case _ELL_ROWS_GROUPS: func = &spmvRowsBlocksELL; toCSR = 0; break;
        case _ELL_TILES: func = &spmvTilesELL; toCSR = 0; break;
        case _B200_CSR_ROWS: func = &b200SpMVRowsCSR; break;
        case _B200_CSR_ROWS_WARP: func = &b200SpMVWarpPerRowCSR; break;
        case _B200_CSR_ADAPTIVE: func = &b200SpMVAdaptiveCSR; break;
        case _B200_CSR_SELL: func = &b200SpMVRowsSELL; break;
        case _B200_CSR_XWINDOW: func = &b200SpMVRowsXWIN; break;
        case _B200_ELL_ROWS: func = &b200SpMVRowsELL; toCSR = 0; break;
        case _B200_ELL_ROWS_NT: func = &b200SpMVRowsELLNNTransposed; toCSR = 0; break;
        case _B200_ELL_ROWS_WARP_NT: func = &b200SpMVWarpsPerRowELLNTrasposed; toCSR = 0; break;
    }
    spmat* mat = toCSR ? MMtoCSR(argv[1]) : MMtoELL(argv[1]);
    if (!mat) { ERRPRINT("err during parsing MatrixMarket\n"); return out; }
    ulong vectSize = mat->N;
    double *vector = NULL, *outV = NULL, *oracleV = NULL;
    spmat* csr = NULL;
    if (!strncmp(argv[2], RNDVECT, strlen(RNDVECT))) {
        if (!(vector = malloc(vectSize * sizeof(*vector)))) { ERRPRINT("rnd vector malloc failed\n"); goto _free; }
        spmvb200_synth_vector_host(0x5EED0077ull, 0, vectSize, 3e-5, vector); /* finite, seeded; |x| < MAXRND (config.h:115) */
    } else {
        if (!(vector = readDoubleVector(argv[2], &vectSize))) { fprintf(stderr, "err during readDoubleVector at:%s\n", argv[2]); goto _free; }
        if (vectSize != mat->N) { ERRPRINT("vector not compatible with sparse matrix\n"); goto _free; }
    }
    if (!(outV = malloc(mat->M * sizeof(*outV)))) { ERRPRINT("outV malloc errd\n"); goto _free; }
    memset(outV, 0xFF, mat->M * sizeof(*outV));
    if (cmode >= _B200_CSR_ROWS) { /* page-lock the caller's vectors in place (optional) -- released at _free before free() */
        spmvb200_host_register(vector, vectSize * sizeof(*vector));
        spmvb200_host_register(outV, mat->M * sizeof(*outV));
    }
    double start = omp_get_wtime();
    if ((out = func(mat, vector, &Conf, outV))) { ERRPRINT("compute function selected failed...\n"); goto _free; }
    double elapsed = omp_get_wtime() - start;
    if (writeDoubleVector(OUTVECTORDUMPRAW, outV, mat->M) || writeDoubleVectorAsStr(OUTVECTORDUMP, outV, mat->M)) ERRPRINT("outV dump err\n");
    printf("cmode:%d\telapsed:\t %le elapsedInternal %le\n", cmode, elapsed, ElapsedInternal);
    if (check) {
        out = EXIT_FAILURE;
        csr = toCSR ? mat : MMtoCSR(argv[1]);
        if (!csr || !(oracleV = malloc(csr->M * sizeof(*oracleV)))) goto _free;
        sgemvSerial(csr, vector, &Conf, oracleV);
        uint64_t bad = 1;
        double worst = 0;
        int failed = 1;
        double dmax = 0;
        if (spmvb200_compare_strict_csr(csr->M, (const uint64_t*) csr->IRP, (const uint64_t*) csr->JA, csr->AS, vector, oracleV, outV, 1e-12, &bad, &worst) ||
            spmvb200_compare_abs(csr->M, oracleV, outV, DOUBLE_DIFF_THREASH, &failed, &dmax))
            goto _free;
        const int ref_failed = doubleVectorsDiff(oracleV, outV, csr->M, NULL);
        printf("check: strict rows failing %lu (worst |dy|/sum|ax| %.3e)  max|dy| %.3e  doubleVectorsDiff %s\n", (unsigned long) bad, worst, dmax,
               ref_failed ? "FAILED" : "ok");
        out = (bad || failed || ref_failed) ? EXIT_FAILURE : EXIT_SUCCESS;
    }
_free:
    spmvb200_cache_drop(NULL); /* also releases the page-lock registrations */
    if (csr && csr != mat) freeSpmat(csr);
    if (mat) freeSpmat(mat);
    free(vector);
    free(outV);
    free(oracleV);
    return out;
}
