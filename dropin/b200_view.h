// dropin/b200_view.h -- the narrow device-side VIEW the strict drop-ins keep next to the reference's device layout.
//
// The reference's kernels get `spmat* m` with 64-bit JA / IRP / RL (16 B per non-zero) and whatever launch geometry the driver
// picked (src/main.cu:221-226, test/SpMV_test.cu:284-302).  The drop-in uploaders (dropin/b200_cudaUtils.cu) build, next to those
// arrays, the layout this engine's kernels want -- 32-bit ids, sliced / column-major storage, row lengths -- and the drop-in
// kernels (dropin/b200_SpMV_CUDA.cu) find it through the struct itself:
//
//   CSR struct : the view lives BEHIND the row pointer, in the same allocation (IRP[M+1] | header | arrays); the two pitch fields,
//                unused for CSR and zeroed by the reference's own uploader, carry  pitchJA = B200_VIEW_MAGIC,  pitchAS = device
//                address of the header.
//   ELL structs: IRP is unused for ELL (NULL after the reference's uploader); here it points at the view's allocation, whose
//                first word is the magic.
//
// Either way the view is released by the reference's own cudaFreeSpmat (it cudaFree()s AS, JA, IRP, RL:
// src/include/cudaUtils.h:70-78) -- no registry, no extra free call, nothing the unchanged drivers need to know.  A struct
// uploaded by the REFERENCE's spMatCpy* has neither marker: the kernels then take the plain 64-bit walk.
#ifndef B200_VIEW_H
#define B200_VIEW_H

#define B200_VIEW_MAGIC 0xB200B200C5E11ull
#define B200_LONG_ROW 256u  // CSR rows longer than this stay out of the slices: a warp each

struct B200ViewCSR {
    unsigned long long magic;
    unsigned Mpad, nslices, nlong, lanes;  // lanes: sub-warp width of the warp-per-row entry point, from the mean row length
    // narrow CSR: values stay in the struct's AS (CSR order)
    const unsigned* irp32;     // [M+1]
    const unsigned* ja32;      // [NZ]
    // SELL-32-sigma of the rows of at most B200_LONG_ROW entries: rows sorted by length inside windows of 16384 rows (stable:
    // equal lengths keep their order), slices of 32 sorted rows stored column-major and padded to the slice's longest row
    const unsigned* slice_ptr;  // [nslices+1]
    const unsigned* perm;       // [Mpad] sorted position -> row, 0xffffffff = none
    const unsigned* rl_sorted;  // [Mpad]
    const unsigned* sja;        // [slots]
    const double* sas;          // [slots]
    const unsigned* long_rows;  // [nlong]
};

struct B200ViewELL {
    unsigned long long magic;
    unsigned rows, K;
    unsigned long long pitch;  // elements between slot k and slot k+1 of one row (column-major)
    const unsigned* ja32;      // [K * pitch] column-major
    const double* as_cm;       // [K * pitch] column-major copy of the values; NULL: use the struct's AS (already column-major)
    const unsigned* rl32;      // [rows] effective row lengths
};

#endif
