"""Builds compat/libhns_compat.so: the reference's seven launcher symbols on top of libhns_b200.so, compiled against the
reference's own headers (REF, default /root/reference) and, when OpenVDB is not installed, the POD stand-in under oracle/shim.
Also builds oracle/_ref/libcompat_driver.so: the SAME plain-C driver the tests use for the unmodified reference
(oracle/ref_shim.cu), linked against libhns_compat.so instead -- so identical caller code exercises both implementations."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("REF", "/root/reference")
NVCC = os.environ.get("NVCC", "nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
INC = [f"-I{ROOT}/oracle/shim", f"-I{REF}/externals", f"-I{REF}/src", f"-I{REF}/src/Cuda"]
OUT = os.path.join(HERE, "libhns_compat.so")
DRV = os.path.join(ROOT, "oracle", "_ref", "libcompat_driver.so")


def build() -> None:
    if not os.path.isdir(os.path.join(REF, "src", "Utils")):
        print("compat: reference headers not found, skipping (prebuilt library is used)")
        return
    lib = os.path.join(ROOT, "hnanosolver_b200")
    src = os.path.join(HERE, "hns_compat.cu")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(lib, "libhns_b200.so"))):
        subprocess.check_call([NVCC, "-std=c++17", "--extended-lambda", *ARCH, "-O2", "-lineinfo", "-w", "-Xcompiler", "-fPIC", "-shared", *INC, src,
                               "-o", OUT, f"-L{lib}", "-lhns_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../hnanosolver_b200", "-lcudart"])
    shim = os.path.join(ROOT, "oracle", "ref_shim.cu")
    os.makedirs(os.path.dirname(DRV), exist_ok=True)
    if not os.path.exists(DRV) or os.path.getmtime(DRV) < max(os.path.getmtime(shim), os.path.getmtime(OUT)):
        subprocess.check_call([NVCC, "-std=c++17", "--extended-lambda", *ARCH, "-O2", "-w", "-Xcompiler", "-fPIC", "-shared", "-DHNS_SHIM_LAUNCHERS_ONLY", *INC,
                               shim, "-o", DRV, f"-L{HERE}", "-lhns_compat", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../compat", "-lcudart"])
    print(OUT)


if __name__ == "__main__":
    build()
