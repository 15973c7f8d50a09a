"""Builds compat/libhns_compat.so: the reference's seven launcher symbols on top of libhns_b200.so, compiled against the
reference's own headers (REF, default /root/reference) and, when OpenVDB is not installed, the POD stand-in under compat/shim
(openvdb::Coord / openvdb::Vec3f as the 12-byte PODs the launchers use them as)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("REF", "/root/reference")
NVCC = os.environ.get("NVCC", "nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
INC = [f"-I{HERE}/shim", f"-I{REF}/externals", f"-I{REF}/src", f"-I{REF}/src/Cuda"]
OUT = os.path.join(HERE, "libhns_compat.so")


def build() -> None:
    if not os.path.isdir(os.path.join(REF, "src", "Utils")):
        print("compat: reference headers not found, skipping (prebuilt library is used)")
        return
    lib = os.path.join(ROOT, "hnanosolver_b200")
    src = os.path.join(HERE, "hns_compat.cu")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(lib, "libhns_b200.so"))):
        subprocess.check_call([NVCC, "-std=c++17", "--extended-lambda", *ARCH, "-O2", "-lineinfo", "-w", "-Xcompiler", "-fPIC", "-shared", *INC, src,
                               "-o", OUT, f"-L{lib}", "-lhns_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../hnanosolver_b200", "-lcudart"])
    print(OUT)


if __name__ == "__main__":
    build()
