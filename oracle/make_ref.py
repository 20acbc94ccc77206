"""Recipe for oracle/_ref/: a run-time copy of the reference's hot-path Python files so that `bench.py --impl
reference` and the `cpu_baseline` leg can time the REFERENCE'S OWN modules on the GPU box's host cores, where /root/reference
does not exist.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (no reference source enters the repository history)
but not gpurun-ignored, so it travels to the GPU box like the built .so files.  `__graft_entry__.build()` runs this when
/root/reference is present; nothing in the product package imports oracle/.

    python oracle/make_ref.py            # (re)create oracle/_ref from /root/reference
"""
import os
import shutil
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("JMT_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")

# the ten files of SURVEY 8(a) (+ the package marker of models/)
FILES = ["models/__init__.py", "models/two_transformers.py", "models/mm_multi_transformers.py", "models/mm_transformers.py",
         "models/intra_modal_transformer_fusion.py", "models/fc_layer.py", "models/temporal_convolutional_model.py",
         "losses/loss.py", "losses/CCCLoss.py", "EvaluationMetrics/cccmetric.py", "padSequence.py"]


def make(force: bool = False) -> str:
    """Copy FILES from the reference checkout into oracle/_ref/ (no-op when the checkout is absent)."""
    if not os.path.isdir(os.path.join(REF, "models")):
        return DST
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            raise FileNotFoundError(src)
        if force or not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
    return DST


def available() -> bool:
    return all(os.path.exists(os.path.join(DST, rel)) for rel in FILES)


def import_reference():
    """The reference's own modules from oracle/_ref (stubs for the comet_ml / matplotlib imports at the top of
    models/mm_transformers.py:2-6).  Returns a dict of classes / functions."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
    for name in ["comet_ml", "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["comet_ml"].Experiment = object
    sys.modules["mpl_toolkits.axes_grid1"].ImageGrid = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if DST not in sys.path:
        sys.path.insert(0, DST)
    warnings.filterwarnings("ignore")
    import torch
    from models.two_transformers import Two_transformers
    from models.fc_layer import FcLayer
    from models.temporal_convolutional_model import TemporalConvNet
    from losses.loss import CCCLoss
    orig = torch.Tensor.cuda                   # losses/loss.py:16 calls .cuda() in __init__ (bins unused when digitize_num == 1)
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        live = CCCLoss(digitize_num=1)
    finally:
        torch.Tensor.cuda = orig
    return dict(Two_transformers=Two_transformers, FcLayer=FcLayer, TemporalConvNet=TemporalConvNet, live_loss=live)


if __name__ == "__main__":
    print(make(force=True), "available" if available() else "NOT available")
