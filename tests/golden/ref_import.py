"""Import the UNMODIFIED reference modules from /root/reference with import-time stubs.

Only usable in the build container (where /root/reference is mounted).  Used by
``tests/golden/make_golden.py`` to generate the committed golden vectors and by the
``not gpu`` tests that pin the oracle against the live reference when it is present.
Nothing here is product code; nothing is copied from the reference.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("EEG_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "Retrieval", "ATMS_retrieval.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns the reference's ``ATMS_retrieval`` module (ATMS, ClipLoss, train_model ...)."""
    if "ATMS_retrieval" in sys.modules:
        return sys.modules["ATMS_retrieval"]
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    os.environ.setdefault("WANDB_MODE", "disabled")

    class _Dummy:  # placeholder for third-party model classes that ATM-S never instantiates
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed third-party class")

    # packages that the reference imports at module top but that contribute no arithmetic to ATM-S
    if "reformer_pytorch" not in sys.modules:
        _stub("reformer_pytorch", LSHSelfAttention=_Dummy)
    if "clip" not in sys.modules:
        _stub("clip")
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
    if "braindecode" not in sys.modules:
        bd = _stub("braindecode")
        bd.models = _stub("braindecode.models", EEGNetv4=_Dummy, ATCNet=_Dummy, EEGConformer=_Dummy,
                          EEGITNet=_Dummy, ShallowFBCSPNet=_Dummy)
    if "torchvision" not in sys.modules:
        try:
            import torchvision  # noqa: F401
        except Exception:
            tv = _stub("torchvision")
            tv.transforms = _stub("torchvision.transforms")
    if "tqdm" not in sys.modules:
        try:
            import tqdm  # noqa: F401
        except Exception:
            _stub("tqdm")
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:
            _stub("wandb")
    _stub("eegdatasets_leaveone", EEGDataset=_Dummy)
    for p in (os.path.join(REF_ROOT, "Retrieval"), REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ATMS_retrieval  # noqa: E402
    return ATMS_retrieval


def _import_file(mod_name: str, path: str):
    """import one reference script under its own module name (two of them are both called ATMS_*.py with a class ATMS)"""
    import importlib.util
    if mod_name in sys.modules:
        return sys.modules[mod_name]
    spec = importlib.util.spec_from_file_location(mod_name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[mod_name] = mod
    spec.loader.exec_module(mod)
    return mod


def import_reference_joint():
    """Retrieval/ATMS_retrieval_joint_train.py (per-subject value embeddings), unmodified"""
    import_reference()          # installs the stubs and sys.path entries
    class _Dummy:
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed third-party class")
    _stub("eegdatasets_joint_subjects", EEGDataset=_Dummy)
    return _import_file("ATMS_retrieval_joint_train", os.path.join(REF_ROOT, "Retrieval", "ATMS_retrieval_joint_train.py"))


def import_reference_reconstruction():
    """Generation/ATMS_reconstruction.py (MSE + InfoNCE loss mix), unmodified"""
    import_reference()
    return _import_file("ATMS_reconstruction", os.path.join(REF_ROOT, "Generation", "ATMS_reconstruction.py"))
