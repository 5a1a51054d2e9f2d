"""Compiles the C restatement of the CSR builder (oracle/csr_oracle.c) into oracle/libgnnfd_csr_oracle.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csr_oracle.c")
OUT = os.path.join(HERE, "libgnnfd_csr_oracle.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", OUT, SRC])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
