"""Run the reference's OWN source files from /root/reference (never copied).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Works only where
/root/reference exists (the build container); the GPU box never calls this.

The reference is Python 2.7 / torch 0.4.1.  What is shimmed, and nothing else:
  * ``transforms3d`` (absent here; imported at common/pose_utils.py:13-14 but
    not used by the torch section :21-304) -> empty stub modules;
  * ``xrange`` (common/pose_utils.py:242,256) -> ``range``;
  * common/pose_utils.py is exec'd up to the first Python-2 ``print`` statement
    (numpy/PGO code past that line is out of scope);
  * common/criterion.py:150 ``T = s[1] / 2`` relies on py2 integer division;
    the source text is loaded verbatim and that one expression is rewritten to
    ``//`` at load time (py2 semantics under py3).
models/posenet.py imports unmodified.
"""
import builtins
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("GEOMAPNET_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "common", "criterion.py"))


_cache = {}


def _stub(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace with the reference's PoseNet, MapNet, criteria and
    torch pose utils, executed from REF_ROOT."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)

    if "transforms3d" not in sys.modules:
        t3d = _stub("transforms3d")
        t3d.quaternions = _stub("transforms3d.quaternions")
        t3d.euler = _stub("transforms3d.euler")
    if not hasattr(builtins, "xrange"):
        builtins.xrange = range

    # --- common.pose_utils: torch section + numpy helpers, up to first py2 print
    pu_path = os.path.join(REF_ROOT, "common", "pose_utils.py")
    with open(pu_path) as f:
        lines = f.readlines()
    stop = len(lines)
    for i, ln in enumerate(lines):
        s = ln.strip()
        if s.startswith("print ") or s.startswith("print'") or s.startswith('print"'):
            stop = i
            break
    # cut back to the start of the enclosing top-level def
    while stop > 0 and not lines[stop].startswith("def ") and not lines[stop].startswith("class "):
        stop -= 1
    src = "".join(lines[:stop])
    common_pkg = types.ModuleType("common")
    common_pkg.__path__ = []
    pose_utils = types.ModuleType("common.pose_utils")
    pose_utils.__file__ = pu_path
    exec(compile(src, pu_path, "exec"), pose_utils.__dict__)
    common_pkg.pose_utils = pose_utils

    saved = {k: sys.modules.get(k) for k in ("common", "common.pose_utils", "common.criterion")}
    sys.modules["common"] = common_pkg
    sys.modules["common.pose_utils"] = pose_utils
    try:
        # --- common.criterion: verbatim, with the py2 integer division shim
        cr_path = os.path.join(REF_ROOT, "common", "criterion.py")
        with open(cr_path) as f:
            cr_src = f.read()
        assert "T = s[1] / 2" in cr_src
        cr_src = cr_src.replace("T = s[1] / 2", "T = s[1] // 2")
        criterion = types.ModuleType("common.criterion")
        criterion.__file__ = cr_path
        exec(compile(cr_src, cr_path, "exec"), criterion.__dict__)

        # --- models.posenet: unmodified
        pn_path = os.path.join(REF_ROOT, "models", "posenet.py")
        spec = importlib.util.spec_from_file_location("_ref_models_posenet", pn_path)
        posenet = importlib.util.module_from_spec(spec)
        env_before = os.environ.get("TORCH_MODEL_ZOO")
        path_before = list(sys.path)
        spec.loader.exec_module(posenet)
        sys.path[:] = path_before  # posenet.py inserts '../' (models/posenet.py:19)
        if env_before is None:
            os.environ.pop("TORCH_MODEL_ZOO", None)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    ns = types.SimpleNamespace(
        pose_utils=pose_utils, criterion=criterion, posenet=posenet,
        PoseNet=posenet.PoseNet, MapNet=posenet.MapNet,
        PoseNetCriterion=criterion.PoseNetCriterion,
        MapNetCriterion=criterion.MapNetCriterion,
        MapNetOnlineCriterion=criterion.MapNetOnlineCriterion)
    _cache["ns"] = ns
    return ns


def build_reference_model(state, kind="posenet", droprate=0.0, filter_nans=False):
    """Reference PoseNet/MapNet around torchvision resnet34 (scripts/train.py:76-84)
    with weights from ``state`` (oracle.weights.make_state)."""
    import torch
    import torchvision
    ns = load()
    fe = torchvision.models.resnet34(weights=None)
    net = ns.PoseNet(fe, droprate=droprate, pretrained=False, filter_nans=filter_nans)
    missing = net.load_state_dict({k: v.clone() for k, v in state.items()}, strict=True)
    del missing
    if kind == "posenet":
        return net
    return ns.MapNet(mapnet=net)


def reference_step(model, criterion, x, targ, lr=1e-4, weight_decay=5e-4,
                   max_grad_norm=0.0, do_step=True):
    """Restates the 10 lines of common/train.py:339-361 (step_feedfwd) around
    the reference model/criterion (the function itself does not parse on py3:
    ``async=`` kwarg at :341)."""
    import torch
    params = [{"params": list(model.parameters())}]
    crit_params = [p for p in criterion.parameters() if p.requires_grad]
    if crit_params:
        params.append({"params": crit_params})
    opt = torch.optim.Adam(params, lr=lr, weight_decay=weight_decay)
    x = x.clone().requires_grad_(True)            # train.py:339
    out = model(x)                                 # :343
    loss = criterion(out, targ)                    # :351
    opt.zero_grad()                                # :355
    loss.backward()                                # :356
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    cgrads = {n: (p.grad.clone() if p.grad is not None else None)
              for n, p in criterion.named_parameters()}
    if max_grad_norm > 0.0:                        # :357-358
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)
    if do_step:
        opt.step()                                 # :359
    return float(loss.item()), out.detach(), grads, cgrads
