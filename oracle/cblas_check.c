/*
 * oracle/cblas_check.c -- TEST INFRASTRUCTURE (not product code).
 * The reference pins its serial oracle sgemvSerial with a dense CBLAS GEMV on small cases (test/SpMV_CBLAS.c:32-57,
 * test/SpMV_test.cu:221-236: densify the CSR matrix, cblas_dgemv, compare).  This little program makes that one call with the
 * reference's own prebuilt reference-BLAS archives (test/CBLAS_LAPACK/lib_deb/libcblas.a + librefblas.a, linked from where they
 * lie; they are not position independent, hence a program and not a shared object) so that tests/test_oracle.py can cross-check
 * the restatement -- and the live reference -- the same way.
 *   stdin : uint64 M, uint64 N, M*N doubles (row-major dense matrix), N doubles (x)        stdout: M doubles (y = A x)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <cblas.h>

int main(void) {
    uint64_t dims[2];
    if (fread(dims, 8, 2, stdin) != 2) return 2;
    const size_t M = dims[0], N = dims[1];
    double *a = malloc((M * N + 1) * 8), *x = malloc((N + 1) * 8), *y = malloc((M + 1) * 8);
    if (!a || !x || !y) return 3;
    if (fread(a, 8, M * N, stdin) != M * N || fread(x, 8, N, stdin) != N) return 4;
    cblas_dgemv(CblasRowMajor, CblasNoTrans, (int) M, (int) N, 1.0, a, (int) N, x, 1, 0.0, y, 1);
    if (fwrite(y, 8, M, stdout) != M) return 5;
    return 0;
}
