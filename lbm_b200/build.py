"""Build lbm_b200/libblbm.so in-tree with nvcc for sm_100a (see csrc/Makefile)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=False, force=False):
    csrc = os.path.join(HERE, "csrc")
    cmd = ["make", "-C", csrc, "--no-print-directory", "-j4"]
    if force:
        subprocess.run(["make", "-C", csrc, "--no-print-directory", "clean"], check=True,
                       stdout=subprocess.DEVNULL)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        sys.stdout.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("building libblbm.so failed")
    return os.path.join(HERE, "libblbm.so")


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
