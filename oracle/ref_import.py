"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference layer modules.

Works only where /root/reference exists (the build container, never the GPU box).
The reference's package __init__ files pull coqpit/librosa/trainer, which are not
installed; registering two empty namespace packages lets the
hot-path layer modules import untouched.  Used by tests/golden/make_golden.py and by
the `not gpu` tests that pin oracle/vits_oracle.py against the real reference.
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
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


# ----------------------------------------------------------------------------- third-party placeholders
# The reference imports ~20 third-party packages that are not installed here (no network): coqpit, trainer, librosa,
# matplotlib, soundfile, pysbd, anyascii, inflect, gruut, ... .  None of them does arithmetic on the path under test
# (the one exception, librosa.filters.mel, is documented as "parity unpinned").  Registering inert placeholders for
# THOSE packages -- never for anything under TTS/ except the phonemizer sub-package, which is out of scope -- lets
# the UNMODIFIED reference modules import: TTS.tts.models.vits.Vits (the real inference glue), TTSTokenizer,
# AudioProcessor, HifiganConfig / setup_generator / GAN, Synthesizer.  No reference source is copied or patched.
_STUB_ROOTS = {"anyascii", "bangla", "bnnumerizer", "bnunicodenormalizer", "gruut", "gruut_ipa", "inflect", "jamo", "jieba",
               "librosa", "matplotlib", "mutagen", "pypinyin", "pysbd", "soundfile", "g2pkk", "hangul_romanize",
               "num2words", "unidecode", "nltk", "umap", "encodec", "spacy", "pandas_stub_never"}
_STUB_PREFIXES = ("TTS.tts.utils.text.phonemizers",)
_stubs_installed = False


class _PlaceholderMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _placeholder(name)

    def __iter__(cls):
        return iter(())


def _placeholder(name):
    def _getattr(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _placeholder(n)()

    return _PlaceholderMeta(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: self,
                                       "__iter__": lambda self: iter(()), "__len__": lambda self: 0,
                                       "__getattr__": _getattr})


class _PlaceholderModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _placeholder(name)


class _PlaceholderFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS or fullname.startswith(_STUB_PREFIXES):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _PlaceholderModule(spec.name)
        m.__path__ = []
        if spec.name in _STUB_PREFIXES:
            m.DEF_LANG_TO_PHONEMIZER = {}
        return m

    def exec_module(self, module):
        pass


def install_third_party_placeholders():
    """Idempotent.  coqpit.Coqpit and trainer.TrainerModel / TrainerConfig get *functional* minimal stand-ins (attribute
    + item access on a dataclass; an nn.Module base), because the reference's config and model classes derive from them."""
    global _stubs_installed
    if _stubs_installed:
        return
    import dataclasses

    import torch

    def _have(name):
        try:
            return importlib.util.find_spec(name) is not None
        except Exception:
            return False

    sys.meta_path.insert(0, _PlaceholderFinder())
    if not _have("coqpit"):
        class Coqpit:  # the subset of coqpit.Coqpit the reference touches on the inference path
            def __getitem__(self, k):
                return getattr(self, k)

            def __setitem__(self, k, v):
                setattr(self, k, v)

            def __contains__(self, k):
                return hasattr(self, k)

            def has(self, k):
                return hasattr(self, k)

            def get(self, k, d=None):
                return getattr(self, k, d)

            def check_values(self):
                pass

            def keys(self):
                return [f.name for f in dataclasses.fields(self)]

            def items(self):
                return [(f.name, getattr(self, f.name)) for f in dataclasses.fields(self)]

            def to_dict(self):
                return dataclasses.asdict(self)

            def update(self, d, allow_new=False):
                for k, v in d.items():
                    setattr(self, k, v)

        cm = types.ModuleType("coqpit")
        cm.Coqpit, cm.check_argument, cm.MISSING = Coqpit, (lambda *a, **k: None), dataclasses.MISSING
        sys.modules["coqpit"] = cm
    if not _have("trainer"):
        Coqpit = sys.modules["coqpit"].Coqpit

        @dataclasses.dataclass
        class TrainerConfig(Coqpit):
            output_path: str = "output"
            epochs: int = 1
            batch_size: int = 1
            eval_batch_size: int = 1
            mixed_precision: bool = False
            lr: float = 1e-3
            optimizer: str = None
            optimizer_params: dict = None
            lr_scheduler: str = None
            lr_scheduler_params: dict = None
            grad_clip: float = 0.0
            scheduler_after_epoch: bool = False

        class TrainerModel(torch.nn.Module):
            pass

        tm = _PlaceholderModule("trainer")
        tm.__path__ = []
        tm.TrainerConfig, tm.TrainerModel = TrainerConfig, TrainerModel
        sys.modules["trainer"] = tm
        _STUB_ROOTS.add("trainer")    # trainer.* sub-modules resolve to placeholders
    _stubs_installed = True


def load_full():
    """The unmodified reference with its real package __init__ files (third-party placeholders installed): returns a
    dict with the model / tokenizer / audio / vocoder-config classes the parity tests pin the oracle against."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    install_third_party_placeholders()
    mods = load()
    names = {
        "vits_model": "TTS.tts.models.vits",
        "vits_config": "TTS.tts.configs.vits_config",
        "tokenizer": "TTS.tts.utils.text.tokenizer",
        "characters": "TTS.tts.utils.text.characters",
        "cleaners": "TTS.tts.utils.text.cleaners",
        "fairseq": "TTS.tts.utils.fairseq",
        "processor": "TTS.utils.audio.processor",
        "numpy_transforms": "TTS.utils.audio.numpy_transforms",
        "vocoder_generic_utils": "TTS.vocoder.utils.generic_utils",
        "hifigan_config": "TTS.vocoder.configs.hifigan_config",
        "vocoder_models": "TTS.vocoder.models",
        "gan": "TTS.vocoder.models.gan",
        "synthesizer": "TTS.utils.synthesizer",
        "synthesis": "TTS.tts.utils.synthesis",
    }
    mods.update({k: importlib.import_module(v) for k, v in names.items()})
    return mods


def load():
    """Returns a dict of the reference modules on the hot path."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # the package __init__ files pull coqpit / librosa / trainer: inert placeholders stand in for those third-party
    # packages (see install_third_party_placeholders), the reference's own modules are imported unmodified
    install_third_party_placeholders()
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
