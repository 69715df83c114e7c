"""Build tests/emu/libplf_emu.so: the product kernels compiled for the CPU with the TEST-ONLY
CUDA emulation (tests/emu/cuda_emu.h).  Never used by the product path."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "spl_slam_b200", "csrc")
LIB = os.path.join(HERE, "libplf_emu.so")


def build_emu(force=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.inc")) + \
        glob.glob(os.path.join(HERE, "cuda_emu.*")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) > os.path.getmtime(d) for d in deps):
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = ["-std=c++17", "-O2", "-g", "-fPIC", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-DPLF_EMU", "-DLSD_BIG_BUCKET=5", "-DLSD_GIANT_BUCKET=7", "-DGC_MAXWARPS=4",   # tiny threshold: the emulated tests stress the speculative LSD path
             
             "-Wno-unknown-pragmas", "-Wno-unused-function", "-I", HERE, "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    extra = os.environ.get("PLF_EMU_EXTRA_FLAGS", "").split()   # e.g. -fsanitize=address (profiles/emu_asan.sh)
    flags = flags + extra
    procs, objs = [], []
    for s in srcs + [os.path.join(HERE, "cuda_emu.cc")]:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = ["g++"] + flags + ["-x", "c++", "-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("emu compile failed for %s:\n%s" % (s, out))
    subprocess.check_call(["g++", "-shared", "-pthread"] + extra + ["-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build_emu(force=True))
