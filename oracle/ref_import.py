"""Import the UNMODIFIED reference from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` to generate the committed
fixtures and by optional ``-m "not gpu"`` cross-checks that skip when /root/reference is absent
(it does not exist on the GPU box).  Nothing is copied from the reference; it is imported where it
lies.  Shims: ``scipy.special.sph_harm/lpmn`` placeholders (imported by
utils/spherical.py:2 but unused on this path) and ``opt.agg_axis_weight=None`` (avoids a hard-coded
device="cuda" at point_aggregators.py:454; arithmetic identical, :828-829).
"""
import argparse
import os
import sys
import types

# /root/reference where it exists (build container), else the byte-for-byte copy staged by oracle/stage_reference.py (GPU box)
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "pyref")
REF_ROOT = os.environ.get("HNR_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REF_ROOT, "models", "aggregators")) and os.path.isdir(os.path.join(_STAGED, "models", "aggregators")):
    REF_ROOT = _STAGED


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models", "aggregators"))


def _prepare():
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import scipy.special
    for n in ("sph_harm", "lpmn"):
        if not hasattr(scipy.special, n):
            setattr(scipy.special, n, None)


class _Stub(types.ModuleType):
    """placeholder for an uninstallable import of the reference (pycuda, matplotlib, ...): any
    attribute is an empty class, so `class Holder(pycuda.driver.PointerHolderBase)` still parses.
    None of these modules does arithmetic on the paths the goldens exercise."""
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = sys.modules.get(self.__name__ + "." + name)
        if sub is not None:
            return sub
        return type(name, (object,), {"__init__": lambda self, *a, **k: None})


def stub_missing(names=("matplotlib", "matplotlib.pyplot", "pycuda", "pycuda.compiler", "pycuda.driver",
                        "pycuda.gpuarray", "pycuda.autoinit", "pytorch_msssim", "imageio", "kornia",
                        "imutils", "h5py", "plyfile", "open3d", "inplace_abn", "torch_scatter", "lpips",
                        "skimage", "skimage.metrics", "tensorboardX", "warmup_scheduler", "data.load_blender")):
    import importlib
    for n in names:
        if n in sys.modules:
            continue
        try:
            importlib.import_module(n)
        except Exception:
            sys.modules[n] = _Stub(n)


def import_with_stubs(modname, max_tries=40):
    """import a reference module, stubbing every module that is not installable here."""
    import importlib
    _prepare()
    stub_missing()
    for _ in range(max_tries):
        try:
            return importlib.import_module(modname)
        except ModuleNotFoundError as e:
            if e.name is None or e.name.startswith("models"):
                raise
            sys.modules[e.name] = _Stub(e.name)
    raise RuntimeError("too many missing modules importing " + modname)


def shipped_opt(**over):
    """argparse namespace with the values every shipped dev_script uses."""
    _prepare()
    from models.aggregators.point_aggregators import PointAggregator
    p = argparse.ArgumentParser()
    PointAggregator.modify_commandline_options(p)
    opt = p.parse_args([])
    vals = dict(agg_dist_pers=20, agg_intrp_order=2, agg_distance_kernel="linear", act_type="LeakyReLU",
                shading_color_mlp_layer=4, num_feat_freqs=3, dist_xyz_freq=5, point_features_dim=32,
                num_pos_freqs=10, num_viewdir_freqs=4, point_color_mode="1", point_dir_mode="1",
                point_conf_mode="1", use_nearest=4, is_train=False, dynamic_nearest=0,
                zero_one_loss_items="conf_coefficient", sparse_loss_weight=0, prob=0,
                dilation_setup="7_8_1_8", agg_axis_weight=None, drop_ratio=0.0, drop_patch=1,
                raydist_mode_unit=1)
    vals.update(over)
    for k, v in vals.items():
        setattr(opt, k, v)
    return opt


def aggregator(opt):
    _prepare()
    import contextlib, io
    from models.aggregators.point_aggregators import PointAggregator
    with contextlib.redirect_stdout(io.StringIO()):
        return PointAggregator(opt)


def rendering():
    _prepare()
    from models.rendering import diff_ray_marching, diff_render_func
    return diff_ray_marching, diff_render_func


def networks():
    _prepare()
    from models.helpers import networks
    return networks
