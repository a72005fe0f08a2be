"""Stage the UNMODIFIED torch half of the reference where it can travel to the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY.  `/root/reference` exists in the build container only; `bench.py --impl reference`, the
`cpu_baseline` leg and the `reference_gpu` leg want to time the reference's own code on the GPU box.  build() therefore copies the
13 source files the aggregation + compositing path imports (found by importing it and listing sys.modules), byte for byte, into
oracle/_ref/pyref/ -- git-ignored like the reference query cubin next to it, so no reference source ever enters the history, but
not gpurun-ignored, so it is there when the box runs bench.py.  oracle/ref_import.py looks for /root/reference first, then here."""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "pyref")
FILES = ["models/__init__.py", "models/aggregators/__init__.py", "models/aggregators/attention.py", "models/aggregators/point_aggregators.py",
         "models/base_model.py", "models/helpers/__init__.py", "models/helpers/geometrics.py", "models/helpers/networks.py",
         "models/rendering/__init__.py", "models/rendering/diff_ray_marching.py", "models/rendering/diff_render_func.py",
         "utils/format.py", "utils/spherical.py"]


def stage(src_root: str = "/root/reference") -> str:
    for f in FILES:
        s, d = os.path.join(src_root, f), os.path.join(DST, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not os.path.exists(d) or os.path.getmtime(s) > os.path.getmtime(d) or os.path.getsize(s) != os.path.getsize(d):
            shutil.copyfile(s, d)
    return DST


if __name__ == "__main__":
    print(stage())
