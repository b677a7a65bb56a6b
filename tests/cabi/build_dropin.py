"""Build tests/cabi/_bin/dropin: the C drop-in test, compiled with gcc against the reference's own
headers and linked with the reference library (oracle/_ref) and libmjb.so. Runs where
/root/reference exists (this container); the binary travels to the GPU box with the snapshot."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_bin", "dropin")


def build(reference="/root/reference"):
    src = os.path.join(HERE, "dropin.c")
    deps = [src, os.path.join(ROOT, "include", "mjb.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(p) for p in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["gcc", "-O1", "-std=c11", "-Wall", "-I" + os.path.join(reference, "include"),
           "-I" + os.path.join(ROOT, "include"), src, "-o", OUT,
           "-L" + os.path.join(ROOT, "oracle", "_ref"), "-lmujoco_ref",
           "-L" + os.path.join(ROOT, "mujoco_inversedynamicstest_b200", "lib"), "-lmjb", "-lm",
           "-Wl,-rpath,$ORIGIN/../../../oracle/_ref",
           "-Wl,-rpath,$ORIGIN/../../../mujoco_inversedynamicstest_b200/lib"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build())
