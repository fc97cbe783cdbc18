"""Build a variant of libsparray_b200.so with extra -D flags (kernel A/B experiments; select it at run
time with SPB_LIB_PATH=<path>).  Usage: python profiles/build_variant.py out.so -DSPB_SORT_IPT=16 ..."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sparray_b200"))
import build as b  # noqa: E402


def main():
    out, flags = sys.argv[1], sys.argv[2:]
    tag = os.path.basename(out).replace(".so", "")
    objdir = os.path.join(b.OBJ, "variant_" + tag)
    os.makedirs(objdir, exist_ok=True)

    def one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        subprocess.run([b.NVCC] + b.FLAGS + flags + ["-c", src, "-o", obj], check=True)
        return obj
    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, b.sources()))
    subprocess.run([b.NVCC, "-shared", "-o", out] + objs + ["-lcudart"], check=True)
    print(out)


if __name__ == "__main__":
    main()
