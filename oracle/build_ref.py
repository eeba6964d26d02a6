"""TEST INFRASTRUCTURE ONLY — compile the reference's own native k-mer counter.

The reference's single native component is idelucs/kmers.pyx (Cython).  When
/root/reference is mounted (build container) this recipe cythonizes it FROM WHERE IT LIES
and compiles the generated C with gcc into oracle/_ref/ (git-ignored, travels to the GPU
box with the snapshot).  No reference source is copied into the repository: the generated
.c file is written to a temporary directory and deleted.  On the GPU box (no mount) the
prebuilt oracle/_ref/*.so is used as is.

Used by: tests (cross-check of the oracle), bench.py's cpu_baseline / --impl reference legs
(the counting part of the reference CPU path).  Never imported by the product package.
"""
import glob
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("IDELUCS_REFERENCE_ROOT", "/root/reference")


def main():
    pyx = os.path.join(REFERENCE_ROOT, "idelucs", "kmers.pyx")
    os.makedirs(OUT, exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    target = os.path.join(OUT, "refkmers" + ext)
    if not os.path.exists(pyx):
        have = glob.glob(os.path.join(OUT, "refkmers*.so"))
        print("reference not mounted;", "using prebuilt " + have[0] if have else "oracle/_ref is empty")
        return
    if os.path.exists(target) and os.path.getmtime(target) >= os.path.getmtime(pyx):
        return
    tmp = tempfile.mkdtemp(prefix="refkmers_")
    try:
        c_file = os.path.join(tmp, "refkmers.c")
        # module name must match the file name: cythonize under the name "refkmers"
        subprocess.check_call([sys.executable, "-m", "cython", "-3", "--module-name", "refkmers", pyx, "-o", c_file])
        inc = sysconfig.get_paths()["include"]
        subprocess.check_call(["gcc", "-O3", "-shared", "-fPIC", "-I", inc, c_file, "-o", target])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print("built", target)


if __name__ == "__main__":
    main()
