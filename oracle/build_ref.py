"""Recipe: compile the reference's own Cython metric (src/helpers/miou_utils.pyx) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Runs only where /root/reference exists (this container).  The source is
read where it lies; the two numpy aliases Cython >= 3.1 removed (np.int_t / np.float_t ->
np.int64_t / np.float64_t, nothing else) are substituted in a scratch copy under a temp dir; only
the built extension lands in oracle/_ref/ (git-ignored, travels with gpurun).  No reference
source is copied into the repo."""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NASB_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def build(force=False):
    pyx = os.path.join(REF, "src", "helpers", "miou_utils.pyx")
    if not os.path.exists(pyx):
        return None
    have = glob.glob(os.path.join(OUT, "miou_utils*.so"))
    if have and not force:
        return have[0]
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="nasb_ref_")
    try:
        src = open(pyx).read().replace("np.int_t", "np.int64_t").replace("np.float_t", "np.float64_t")
        with open(os.path.join(tmp, "miou_utils.pyx"), "w") as f:
            f.write(src)
        with open(os.path.join(tmp, "setup.py"), "w") as f:
            f.write("from setuptools import setup, Extension\nfrom Cython.Build import cythonize\nimport numpy\n"
                    "setup(ext_modules=cythonize([Extension('miou_utils', ['miou_utils.pyx'],"
                    " include_dirs=[numpy.get_include()])], language_level=2))\n")
        subprocess.check_call([sys.executable, "setup.py", "-q", "build_ext", "--inplace"], cwd=tmp,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        so = glob.glob(os.path.join(tmp, "miou_utils*.so"))[0]
        dst = os.path.join(OUT, os.path.basename(so))
        shutil.copy(so, dst)
        return dst
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def load():
    """Import the built reference extension, or None if it was never built."""
    import importlib.util
    have = glob.glob(os.path.join(OUT, "miou_utils*.so"))
    if not have:
        return None
    spec = importlib.util.spec_from_file_location("miou_utils", have[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
