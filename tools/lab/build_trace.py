"""Developer build: libsmot with the correlation kernel's phase stamps compiled in (-DSMOT_XCORR_TRACE) -> tools/lab/libsmot_trace.so.
Only emm.cu is recompiled; the other objects are those of the product build (python -m siammot_b200.build)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from siammot_b200 import build as B  # noqa: E402


def build():
    B.build()
    out = os.path.join(HERE, "libsmot_trace.so")
    obj = os.path.join(HERE, "emm_trace.o")
    src = os.path.join(B.CSRC, "emm.cu")
    subprocess.run([B.NVCC] + B.FLAGS + ["-DSMOT_XCORR_TRACE", "-c", src, "-o", obj], check=True)
    objs = [os.path.join(B.HERE, "build", os.path.basename(s)[:-3] + ".o") for s in B.sources() if not s.endswith("emm.cu")] + [obj]
    subprocess.run([B.NVCC, "-shared", "-cudart", "shared", "-o", out] + objs, check=True)
    os.remove(obj)
    return out


if __name__ == "__main__":
    print(build())
