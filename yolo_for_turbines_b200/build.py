"""Builds libyolo_b200.so in-tree with nvcc for sm_100a (`python -m yolo_for_turbines_b200.build`)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose: bool = False, jobs: int = 8) -> str:
    csrc = os.path.join(HERE, "csrc")
    cmd = ["make", "-C", csrc, f"-j{jobs}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libyolo_b200.so failed (see nvcc output above)")
    return os.path.join(HERE, "libyolo_b200.so")


if __name__ == "__main__":
    print(build(verbose=True))
