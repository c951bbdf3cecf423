"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference layer modules.

Works only where /root/reference exists (the build container, never the GPU box).
The reference's package __init__ files pull coqpit/librosa/trainer, which are not
installed; registering two empty namespace packages lets the
hot-path layer modules import untouched.  Used by tests/golden/make_golden.py and by
the `not gpu` tests that pin oracle/vits_oracle.py against the real reference.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("TTS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "TTS", "tts", "layers"))


def _namespace(name: str, path: str) -> None:
    if name in sys.modules:
        return
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    sys.modules[name] = mod


def load_ref_mas_core():
    """The reference's Cython MAS kernel compiled by oracle/Makefile into oracle/_ref/
    (travels to the GPU box as a prebuilt .so).  Returns None when it was never built."""
    import glob
    import importlib.util

    ref_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    hits = sorted(glob.glob(os.path.join(ref_dir, "core*.so")))
    if not hits:
        return None
    name = "TTS.tts.utils.monotonic_align.core"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns a dict of the reference modules on the hot path."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _namespace("TTS.tts.layers", os.path.join(REF_ROOT, "TTS", "tts", "layers"))
    _namespace("TTS.vocoder.models", os.path.join(REF_ROOT, "TTS", "vocoder", "models"))
    # the compiled Cython MAS (oracle/_ref) is injected so helpers.CYTHON is True
    core = load_ref_mas_core()
    if core is not None:
        _namespace("TTS.tts.utils.monotonic_align", os.path.join(REF_ROOT, "TTS", "tts", "utils", "monotonic_align"))
        sys.modules["TTS.tts.utils.monotonic_align.core"] = core
    names = {
        "networks": "TTS.tts.layers.vits.networks",
        "sdp": "TTS.tts.layers.vits.stochastic_duration_predictor",
        "transforms": "TTS.tts.layers.vits.transforms",
        "transformer": "TTS.tts.layers.glow_tts.transformer",
        "wavenet": "TTS.tts.layers.generic.wavenet",
        "normalization": "TTS.tts.layers.generic.normalization",
        "duration_predictor": "TTS.tts.layers.glow_tts.duration_predictor",
        "hifigan": "TTS.vocoder.models.hifigan_generator",
        "helpers": "TTS.tts.utils.helpers",
    }
    out = {k: importlib.import_module(v) for k, v in names.items()}
    out["mas_core"] = core
    return out
