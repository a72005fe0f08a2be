"""Compile the reference's own CUDA source for the voxel query into oracle/_ref/ (git-ignored).

TEST INFRASTRUCTURE ONLY.  The reference keeps its six query kernels as a CUDA-C string inside
models/neural_points/query_point_indices_worldcoords.py:108-524 and JIT-compiles it with pycuda
(not installable here).  This recipe reads that string FROM /root/reference at build time, resolves
the `#define KN <K>` concatenation the same way the reference does, and runs nvcc on it.  Only the
compiled cubin (a build output) lands in oracle/_ref/; no reference source is written into the repo
(the temporary .cu lives under /tmp and is deleted).  The cubin travels to the GPU box with the
snapshot, where tests/test_gpu_query_vs_reference.py launches the reference kernels through
cuda.bindings to validate both the numpy restatement (oracle/query_oracle.py) and the CUDA product.
"""
import os
import subprocess
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HNR_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF, "models", "neural_points", "query_point_indices_worldcoords.py")


def extract_source(K: int) -> str:
    text = open(SRC).read()
    start = text.index("mod = SourceModule(") + len("mod = SourceModule(")
    end = text.index(", no_extern_c=True)", start)
    expr = text[start:end]
    holder = types.SimpleNamespace(opt=types.SimpleNamespace(K=K))
    return eval("(" + expr + ")", {"self": holder, "str": str})


def build(K: int = 8, out_dir: str = os.path.join(HERE, "_ref")) -> str:
    if not os.path.isfile(SRC):
        raise FileNotFoundError(SRC)
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"ref_query_k{K}.cubin")
    with tempfile.TemporaryDirectory() as td:
        cu = os.path.join(td, "wq.cu")
        with open(cu, "w") as f:
            f.write(extract_source(K))
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-cubin", "-w", "-o", out, cu]
        subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    print(build(k))
