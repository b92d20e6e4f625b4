"""Installs the UNMODIFIED reference (nicococo/scRNA, read-only at /root/reference) into oracle/_ref/ so that
it can travel to the GPU box, where /root/reference does not exist.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/refshim.py).  oracle/_ref/ is git-ignored -- reference sources never
enter the history -- but not gpurun-ignored, like the built .so files.  The install is the reference's own
`setup.py` through pip, from a scratch copy (the build writes egg-info into the source tree and /root/reference is
read-only), with no dependency resolution: the third-party packages it needs are the ones already in the image,
adapted by the 7-item shim in oracle/refshim.py.

    python oracle/make_ref.py            # idempotent; prints what it did

Used by: bench.py --impl reference (the reference's stock solver timed on the host cores), bench.py's
`small_data` leg (configs[0]/[1]: the reference's classes beside scrna_b200 on the default simulated data) and
the live-reference CPU tests when /root/reference itself is absent.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("SCRNA_REFERENCE_SOURCE", "/root/reference")


def installed():
    return os.path.isfile(os.path.join(TARGET, "scRNA", "nmf_clustering.py"))


def main():
    if not os.path.isdir(os.path.join(SOURCE, "scRNA")):
        print("make_ref: no reference tree at %s; %s" % (SOURCE, "keeping the existing install" if installed()
                                                          else "nothing installed"))
        return 0 if installed() else 1
    if installed():
        # re-install only when a source file changed
        same = True
        for name in os.listdir(os.path.join(SOURCE, "scRNA")):
            if name.endswith(".py"):
                a, b = os.path.join(SOURCE, "scRNA", name), os.path.join(TARGET, "scRNA", name)
                if not os.path.isfile(b) or open(a, "rb").read() != open(b, "rb").read():
                    same = False
        if same:
            print("make_ref: oracle/_ref is up to date")
            return 0
        shutil.rmtree(TARGET)
    scratch = tempfile.mkdtemp(prefix="scrna_ref_")
    try:
        src = os.path.join(scratch, "reference")
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns(".git", "notebooks", "doc", "R", "matlab"))
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        subprocess.run(cmd, check=True)
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    if not installed():
        print("make_ref: pip reported success but oracle/_ref/scRNA is missing")
        return 1
    print("make_ref: installed the reference package into %s" % TARGET)
    return 0


if __name__ == "__main__":
    sys.exit(main())
