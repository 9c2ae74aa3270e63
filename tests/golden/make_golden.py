#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref/libspmv_ref.so,
compiled from /root/reference by `make -C oracle ref`).

The reference ships no golden vectors, so the pin for this repo's oracle and
CUDA path is the reference's own output on small seeded inputs:
  * a Matrix Market text file (stored verbatim in the fixture) parsed by the reference's
    MMtoCSR / MMtoELL (src/lib/parser.c:298-376) and ellTranspose (src/commons/sparseUtils.c:145),
  * y from sgemvSerial (src/SpMV_CSR_OMP.c:229) and from every OpenMP implementation in
    SpmvCSRFuncs / SpmvELLFuncs (src/include/SpMV.h:144-159) on a seeded finite x.
Run here (where /root/reference exists):  python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def mm_text(M, N, rows, cols, vals, symmetric=False, pattern=False):
    kind = "pattern" if pattern else "real"
    sym = "symmetric" if symmetric else "general"
    lines = ["%%%%MatrixMarket matrix coordinate %s %s" % (kind, sym), "%% golden fixture", "%d %d %d" % (M, N, len(rows))]
    for r, c, v in zip(rows, cols, vals):
        lines.append("%d %d" % (r + 1, c + 1) if pattern else "%d %d %.17g" % (r + 1, c + 1, v))
    return "\n".join(lines) + "\n"


def case_lap2d(n):
    rows, cols, vals = [], [], []
    for i in range(n * n):
        iy, ix = divmod(i, n)
        for dc, ok, v in ((-n, iy > 0, -1.0), (-1, ix > 0, -1.0), (0, True, 4.0), (1, ix < n - 1, -1.0), (n, iy < n - 1, -1.0)):
            if ok:
                rows.append(i); cols.append(i + dc); vals.append(v)
    return n * n, n * n, rows, cols, vals, {}


def case_random(M, N, density, seed, empty_rows=(), long_row=None):
    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    for r in range(M):
        if r in empty_rows:
            continue
        k = rng.binomial(N, density)
        if long_row is not None and r == long_row[0]:
            k = long_row[1]
        cs = np.sort(rng.choice(N, size=min(k, N), replace=False))
        for c in cs:
            rows.append(r); cols.append(int(c)); vals.append(float(rng.uniform(-1, 1)))
    return M, N, rows, cols, vals, {}


def case_symmetric(n, seed):
    """lower triangle, column-major file order (the usual MM layout) -> reference mirrors entries."""
    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    for c in range(n):
        for r in range(c, n):
            if r == c or rng.random() < 0.15:
                rows.append(r); cols.append(c); vals.append(float(rng.uniform(-2, 2)))
    return n, n, rows, cols, vals, dict(symmetric=True)


def case_pattern(M, N, seed):
    Mx, Nx, rows, cols, vals, _ = case_random(M, N, 0.08, seed)
    return Mx, Nx, rows, cols, [1.0] * len(rows), dict(pattern=True)


CASES = {
    "lap2d_8": lambda: case_lap2d(8),
    "lap2d_33": lambda: case_lap2d(33),
    "rect_empty_rows": lambda: case_random(61, 37, 0.12, 11, empty_rows=(0, 5, 6, 60)),
    "skew_long_row": lambda: case_random(300, 700, 0.01, 12, long_row=(17, 650)),
    "sym_40": lambda: case_symmetric(40, 13),
    "pattern_50x45": lambda: case_pattern(50, 45, 14),
    "single_row": lambda: case_random(1, 129, 0.5, 15),
    "single_col": lambda: case_random(97, 1, 0.6, 16),
}


def finite_x(n, seed):
    """|x| < 3e-5 like the reference's MAXRND (src/include/config.h:115) but seeded and finite."""
    return np.random.default_rng(seed).uniform(-1, 1, n) * 3e-5


def main():
    if not oracle.ref_available():
        oracle.build(ref=True)
    os.environ.setdefault("OMP_SCHEDULE", "nonmonotonic:static")
    for name, mk in CASES.items():
        M, N, rows, cols, vals, kw = mk()
        txt = mm_text(M, N, rows, cols, vals, **kw)
        with tempfile.NamedTemporaryFile("w", suffix=".mtx", delete=False) as f:
            f.write(txt)
            path = f.name
        csr = oracle.ref_mm_to_csr(path)
        ell = oracle.ref_mm_to_ell(path, transpose=True)
        os.unlink(path)
        x = finite_x(csr["N"], 1000 + len(name))
        mat = oracle.ref_spmat(csr["M"], csr["N"], csr["NZ"], csr["ja"], csr["as_"], irp=csr["irp"], rl=csr["rl"])
        emat = oracle.ref_spmat(ell["M"], ell["N"], ell["NZ"], ell["ja"], ell["as_"], rl=ell["rl"], max_row_nz=ell["K"])
        cfg = oracle.ref_config(grid_rows=min(8, max(1, csr["M"])), grid_cols=min(8, max(1, csr["N"])), chunks=0)
        out = dict(mtx=np.array(txt), M=csr["M"], N=csr["N"], NZ=csr["NZ"], K=ell["K"],
                   irp=csr["irp"], ja=csr["ja"], as_=csr["as_"], rl=csr["rl"],
                   ell_ja=ell["ja"], ell_as=ell["as_"], ell_ja_t=ell["ja_t"], ell_as_t=ell["as_t"],
                   ell_t_dims=np.array([ell["t_M"], ell["t_N"], ell["t_MAX_ROW_NZ"]], dtype=np.uint64), x=x)
        M_ = csr["M"]
        out["y_sgemvSerial"] = oracle.ref_call("sgemvSerial", mat, x, cfg, M_)
        for fn in ("spmvRowsBasicCSR", "spmvRowsBlocksCSR", "spmvTilesCSR", "spmvTilesAllocdCSR"):
            out["y_" + fn] = oracle.ref_call(fn, mat, x, cfg, M_)
        for fn in ("spmvRowsBasicELL", "spmvRowsBlocksELL", "spmvTilesELL"):
            out["y_" + fn] = oracle.ref_call(fn, emat, x, cfg, M_)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print("%-18s M=%d N=%d NZ=%d K=%d  max|y|=%.3e" % (name, csr["M"], csr["N"], csr["NZ"], ell["K"],
                                                          np.abs(out["y_sgemvSerial"]).max() if M_ else 0))


if __name__ == "__main__":
    main()
