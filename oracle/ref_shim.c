/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE (not product code).
 *
 * Glue that lets the UNMODIFIED reference C sources under /root/reference/src be
 * linked into a shared object (oracle/_ref/libspmv_ref.so) without the reference's
 * own driver (src/main.cu / test/SpMV_test.cu).  The driver normally provides:
 *   - the audit globals `Start,End,Elapsed,ElapsedInternal` (extern in
 *     src/include/config.h:112, defined in src/main.cu:56 / test/SpMV_test.cu:59);
 *   - the external definitions of the C99 `inline` chunk functions
 *     (src/include/ompChunksDivide.h:33-91, exported by src/main.cu:40); the spmat
 *     alloc/free inlines are already exported by src/commons/sparseUtils.c:28-34.
 * Nothing here restates any reference algorithm; it only includes the reference
 * headers at build time (from where they lie, never copied into this repo) and adds
 * a few ctypes-friendly accessors.
 */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <omp.h>

#include "sparseMatrix.h"
#include "SpMV.h"
#include "ompChunksDivide.h"
#include "parser.h"
#include "utils.h"
#include "macros.h"

/* external definitions of the reference's inline functions (C99 semantics) */
CHUNKS_DISTR chunksFair, chunksFairFolded, chunksNOOP;
extern inline int BISECT_ARRAY(ulong target, ulong* arr, ulong len);
extern inline int IS_NNZ(spmat* smat, ulong i, ulong j);
extern inline int IS_NNZ_linear(spmat* smat, ulong i, ulong j);
extern inline void freeSpAcc(SPACC* r);

double Start, End, Elapsed, ElapsedInternal;

/* ---- accessors used by the python side (tests / cpu baseline) ---- */
size_t refshim_sizeof_spmat(void)  { return sizeof(spmat); }
size_t refshim_sizeof_config(void) { return sizeof(CONFIG); }
int    refshim_rowlens(void) {
#ifdef ROWLENS
    return 1;
#else
    return 0;
#endif
}
double refshim_elapsed_internal(void) { return ElapsedInternal; }
void*  refshim_chunks_fn(int which) {
    switch (which) {
        case 0:  return (void*) chunksNOOP;
        case 1:  return (void*) chunksFair;
        default: return (void*) chunksFairFolded;
    }
}
void refshim_free_spmat(spmat* m) { if (m) freeSpmat(m); }
