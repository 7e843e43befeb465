"""Builds oracle/_ref/: the reference's own Python modules for the hot path, COMPILED to sourceless bytecode.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/mvd_oracle.py's header for who may use oracle/).

The reference (zhizdev/mvdfusion) is pure Python, so "building the reference" = byte-compiling the source files where they
lie under /root/reference; only the compiled bytecode is written, as ONE archive oracle/_ref/ref_bytecode.zip (git-ignored, NOT
gpurun-ignored: it travels to the GPU box like the in-tree `.so`; loose `*.pyc` files do not survive the snapshot).  No reference
SOURCE is copied into the repository.  The interpreter on the GPU box is the same image's CPython, so the bytecode loads there
(`zipimport` of sourceless modules).

    python oracle/build_ref.py            # no-op (exit 0) when /root/reference is absent

What is compiled: mvdfusion/*, utils/*, external/sd1/ldm/** — the files `ViewFusion`, `DDIMSampler`, `GridAttn`, `UNetModel`
and `utils.load_model` import; third-party imports resolve to oracle/ref_shims at run time.
"""
import os
import py_compile
import shutil
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MVD_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(OUT, "ref_bytecode.zip")
TREES = ("mvdfusion", "utils", os.path.join("external", "sd1", "ldm"))
EXTRA = (os.path.join("external", "sd1", "__init__.py"),)


def build(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print(f"oracle/build_ref: {REF} not present — keeping whatever oracle/_ref/ already holds")
        return False
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    n = 0
    files = [os.path.join(REF, e) for e in EXTRA if os.path.exists(os.path.join(REF, e))]
    for tree in TREES:
        for d, _, fs in os.walk(os.path.join(REF, tree)):
            files += [os.path.join(d, f) for f in fs if f.endswith(".py")]
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as zf:
        members, pkg_dirs = set(), set()
        for src in files:
            rel = os.path.relpath(src, REF)
            dst = os.path.join(tmp, rel[:-3] + ".pyc")
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            try:
                py_compile.compile(src, cfile=dst, dfile=rel, doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            except py_compile.PyCompileError as e:  # python-2 era files in the vendored tree that nothing on the path imports
                if verbose:
                    print(f"oracle/build_ref: skipped {rel}: {type(e.exc_value).__name__}")
                continue
            arc = rel[:-3] + ".pyc"
            zf.write(dst, arc)
            members.add(arc)
            d = os.path.dirname(arc)
            while d:
                pkg_dirs.add(d)
                d = os.path.dirname(d)
            n += 1
        # the reference's top-level packages are namespace packages (no __init__.py): give every directory an empty __init__ so
        # that zipimport treats them as regular packages
        empty_src = os.path.join(tmp, "_empty.py")
        open(empty_src, "w").close()
        empty_pyc = os.path.join(tmp, "_empty.pyc")
        py_compile.compile(empty_src, cfile=empty_pyc, dfile="__init__.py", doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        for d in sorted(pkg_dirs):
            if os.path.join(d, "__init__.pyc") not in members:
                zf.write(empty_pyc, os.path.join(d, "__init__.pyc"))
    with open(os.path.join(OUT, "MANIFEST.txt"), "w") as f:
        f.write(f"ref_bytecode.zip: sourceless bytecode of {n} reference modules, compiled by oracle/build_ref.py with CPython {sys.version.split()[0]}\n")
    if verbose:
        print(f"oracle/build_ref: {n} modules -> {ARCHIVE}")
    return True


if __name__ == "__main__":
    build()
