"""Build tests/emu/libdmb200_emu.so (CPU emulation of the kernels; test infrastructure only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "libdmb200_emu.so")
SRCS = [os.path.join(HERE, "emu_api.cpp")]
DEPS = SRCS + [os.path.join(ROOT, "qiskit-aakash_b200", "csrc", "dm_device.h"),
               os.path.join(ROOT, "qiskit-aakash_b200", "csrc", "dm_schedule.h"),
               os.path.join(ROOT, "include", "dmb200.h")]


def build(force=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "qiskit-aakash_b200", "csrc")] + SRCS + ["-o", OUT, "-pthread"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
