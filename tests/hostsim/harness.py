"""TEST-ONLY: builds the sources of gimm-vfi_b200/csrc with g++ -DGV_HOSTSIM
(every thread-per-element kernel body runs as an OpenMP loop; conv / corr GEMM use
naive host loops; the tensor-core kernels' operand rounding is emulated by tests/hostsim/tc_hostsim.cu) so the host orchestration and kernel arithmetic can be checked
against the oracle in the GPU-less build container.  Never loaded by the product."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "gimm-vfi_b200", "csrc")
SOURCES = ["ops_pointwise.cu", "corr.cu", "conv.cu", "ops_tokens.cu", "flowformer.cu", "engine.cu", "c_api.cu"]
SIM_SOURCES = ["tc_hostsim.cu"]   # the emulation of the tcgen05 kernels' arithmetic lives with the tests, not in the product tree
OUT = os.path.join(HERE, "libgimmvfi_hostsim.so")


def build_hostsim(force=False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(HERE, s) for s in SIM_SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")] + [os.path.join(ROOT, "include", "gimmvfi_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cmd = ["g++", "-O2", "-fopenmp", "-DGV_HOSTSIM", "-std=c++17", "-fPIC", "-shared", "-I", CSRC, "-x", "c++"] + srcs + ["-o", OUT]
    subprocess.check_call(cmd)
    return OUT


def hostsim_engine():
    from gimmvfi_b200._lib import Lib
    from gimmvfi_b200.engine import EngineHandle

    return EngineHandle("cpu", lib=Lib(build_hostsim()), allow_hostsim=True)
