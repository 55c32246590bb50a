"""A/B builds: compile libpfasr with some csrc files taken from another git revision, next to the product library.

    python scripts/build_variant.py NAME file.cu@REV [file2.cuh@REV ...]

writes aliparaformerasr_b200/_ab/libpfasr_NAME.so (git-ignored, travels to the GPU box); select it with
PFASR_LIB=aliparaformerasr_b200/_ab/libpfasr_NAME.so.  scripts/quick_bench.py times several libraries in one gpurun call,
which is the only way to compare two kernels on the same box.
"""
import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aliparaformerasr_b200 import build as B      # noqa: E402


def main():
    name, specs = sys.argv[1], sys.argv[2:]
    tmp = tempfile.mkdtemp(prefix="pfasr_variant_")
    pkg = os.path.join(tmp, "aliparaformerasr_b200")
    shutil.copytree(os.path.join(ROOT, "aliparaformerasr_b200", "csrc"), os.path.join(pkg, "csrc"))
    shutil.copytree(os.path.join(ROOT, "include"), os.path.join(tmp, "include"))
    for spec in specs:
        f, rev = spec.split("@")
        data = subprocess.run(["git", "show", f"{rev}:aliparaformerasr_b200/csrc/{f}"], cwd=ROOT, capture_output=True, check=True).stdout
        open(os.path.join(pkg, "csrc", f), "wb").write(data)
    objdir = os.path.join(tmp, "obj")
    os.makedirs(objdir)

    def cc(src):
        obj = os.path.join(objdir, src + ".o")
        r = subprocess.run([B._nvcc(), *B.NVCC_FLAGS, "-c", os.path.join(pkg, "csrc", src), "-o", obj], capture_output=True, text=True)
        if r.returncode:
            raise SystemExit(r.stderr)
        return obj

    with ThreadPoolExecutor(8) as ex:
        objs = list(ex.map(cc, B.SOURCES))
    out_dir = os.path.join(ROOT, "aliparaformerasr_b200", "_ab")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libpfasr_{name}.so")
    subprocess.run([B._nvcc(), "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
    shutil.rmtree(tmp)
    print(out)


if __name__ == "__main__":
    main()
