#!/usr/bin/env python
"""Per-kernel SASS census of the built libraries (cuobjdump -sass): the instructions that show what each kernel is built from --
bulk-TMA copies (UBLKCP), mbarrier ops (SYNCS), streaming / wide / narrow global loads (LDG.E...), shared-memory traffic, shuffles,
warp reductions, fp64 math.  Writes profiles/sass_summary.txt (committed, regenerated whenever a kernel changes).
    python tools/sass_summary.py [lib ...]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBS = sys.argv[1:] or [os.path.join(ROOT, "spmv_openmp_cuda_b200", "lib", "libspmv_b200.so"),
                        os.path.join(ROOT, "tests", "integration", "_build", "ref_harness_b200")]
BUCKETS = [("UBLKCP (bulk TMA g->s)", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("LDG.*128", r"\bLDG\.[A-Z0-9.]*128"),
           ("LDG.*64", r"\bLDG\.[A-Z0-9.]*\b64"), ("LDG.*U16", r"\bLDG\.[A-Z0-9.]*U16"), ("LDG_all", r"\bLDG\b|\bLDG\."),
           ("LDG_EF (evict-first)", r"\bLDG\.[A-Z0-9.]*EF"), ("STG", r"\bSTG"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"), ("ATOMS/ATOMG/RED", r"\bATOMS|\bATOMG|\bRED\b|\bRED\."),
           ("SHFL", r"\bSHFL"), ("REDUX/VOTE", r"\bREDUX|\bVOTE"), ("DFMA", r"\bDFMA"), ("DMUL", r"\bDMUL"), ("DADD", r"\bDADD"),
           ("BAR", r"\bBAR\."), ("instructions", r"^\s+/\*[0-9a-f]{4}\*/")]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:  # noqa: BLE001
        return {n: n for n in names}


def main():
    lines_out = ["# SASS census per kernel (cuobjdump -sass, sm_100a); regenerate with tools/sass_summary.py", ""]
    for lib in LIBS:
        if not os.path.exists(lib):
            continue
        sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
        arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
        kernels = collections.OrderedDict()
        cur = None
        for ln in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", ln)
            if m:
                cur = m.group(1)
                kernels[cur] = collections.Counter()
                continue
            if cur is None:
                continue
            for name, pat in BUCKETS:
                if re.search(pat, ln):
                    kernels[cur][name] += 1
        dm = demangle(list(kernels))
        lines_out.append("## %s   (%s, %d kernels)" % (os.path.relpath(lib, ROOT), ", ".join(arch), len(kernels)))
        for k, c in kernels.items():
            short = re.sub(r"\(.*", "", dm[k]).replace("spmvb200::", "").replace("void ", "")
            if c["instructions"] < 40 and not any(c[b] for b in ("UBLKCP (bulk TMA g->s)", "SYNCS (mbarrier)")):
                continue  # tiny helpers
            lines_out.append("%-78s %s" % (short[:78], "  ".join("%s=%d" % (b.split(" ")[0], c[b]) for b, _ in BUCKETS if c[b])))
        lines_out.append("")
    text = "\n".join(lines_out)
    path = os.path.join(ROOT, "profiles", "sass_summary.txt")
    with open(path, "w") as f:
        f.write(text)
    print(text[:6000])


if __name__ == "__main__":
    main()
