"""The tokeniser hand-off into ``Vits.inference`` and the fairseq (MMS) checkpoint import.

Mirrors, with the reference's names and argument meaning:
  BaseVocabulary / BaseCharacters  <- /root/reference/TTS/tts/utils/text/characters.py:38-131, 134-330 (vocabulary
                                      order ``[PAD, EOS, BOS, BLANK, CHARACTERS, PUNCTUATIONS]``, the *_id fallbacks)
  TTSTokenizer                     <- TTS/tts/utils/text/tokenizer.py:10-133 (encode / decode / text_to_ids /
                                      intersperse_blank_char / pad_with_bos_eos; unknown characters are discarded)
  basic_cleaners                   <- TTS/tts/utils/text/cleaners.py:79-83
  FairseqVocab                     <- TTS/tts/models/vits.py:1982-1999
  rehash_fairseq_vits_checkpoint   <- TTS/tts/utils/fairseq.py:4-48
  (``Vits.load_fairseq_checkpoint`` in tts_b200/vits.py uses the last two.)

What is new here is the batch form: the reference tokenises one sentence per ``Synthesizer.tts`` iteration with Python
lists; ``TTSTokenizer.batch_text_to_ids`` maps a list of sentences through a code-point lookup table with numpy,
writes the interspersed blanks with one strided store per sentence into a PINNED ``[B, T_max]`` int64 buffer and
returns it with the lengths, ready for a single non-blocking host->device copy.  Phonemisation is out of scope
(SURVEY 2): ``phonemizer`` may be any callable ``(text, language) -> str``.
"""
import re
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

_whitespace_re = re.compile(r"\s+")


def basic_cleaners(text: str) -> str:
    """Lower-case and collapse whitespace (cleaners.py:79-83)."""
    return re.sub(_whitespace_re, " ", text.lower()).strip()


class BaseVocabulary:
    """A vocabulary given as an ordered list of symbols (characters.py:38-131)."""

    def __init__(self, vocab, pad: str = None, blank: str = None, bos: str = None, eos: str = None):
        self.vocab = vocab
        self.pad, self.blank, self.bos, self.eos = pad, blank, bos, eos

    @property
    def vocab(self):
        return self._vocab

    @vocab.setter
    def vocab(self, vocab):
        self._vocab, self._char_to_id, self._id_to_char = None, None, None
        if vocab is not None:
            self._vocab = list(vocab)
            self._char_to_id = {ch: i for i, ch in enumerate(self._vocab)}
            self._id_to_char = dict(enumerate(self._vocab))

    def _special(self, ch):
        return self.char_to_id(ch) if ch else len(self.vocab)   # unspecified specials sit one past the vocabulary

    pad_id = property(lambda self: self._special(self.pad))
    blank_id = property(lambda self: self._special(self.blank))
    bos_id = property(lambda self: self._special(self.bos))
    eos_id = property(lambda self: self._special(self.eos))

    @property
    def num_chars(self):
        return len(self._vocab)

    def char_to_id(self, char: str) -> int:
        try:
            return self._char_to_id[char]
        except KeyError as e:
            raise KeyError(f" [!] {repr(char)} is not in the vocabulary.") from e

    def id_to_char(self, idx: int) -> str:
        return self._id_to_char[idx]


class BaseCharacters(BaseVocabulary):
    """Vocabulary built from a character set (characters.py:134-330): ``[PAD, EOS, BOS, BLANK] + characters``
    (sorted unless ``is_sorted=False``) ``+ punctuations``."""

    def __init__(self, characters: str = None, punctuations: str = None, pad: str = None, eos: str = None,
                 bos: str = None, blank: str = None, is_unique: bool = False, is_sorted: bool = True):
        chars = list(characters or "")
        if is_unique:
            chars = list(set(chars))
        if is_sorted:
            chars = sorted(chars)
        head = [s for s in (pad, eos, bos, blank) if s is not None and len(s) > 0]
        super().__init__(head + chars + list(punctuations or ""), pad=pad, blank=blank, bos=bos, eos=eos)
        self.characters, self.punctuations = characters, punctuations
        if is_unique:
            assert len(self.vocab) == len(self._char_to_id), " [!] There are duplicate characters in the character set."


class FairseqVocab(BaseVocabulary):
    """``vocab.txt`` of a fairseq MMS checkpoint: one symbol per line, blank = first line, pad = space
    (vits.py:1982-1999)."""

    def __init__(self, vocab_file: str):
        with open(vocab_file, encoding="utf-8") as f:
            symbols = [line.replace("\n", "") for line in f.readlines()]
        super().__init__(symbols, pad=" ", blank=symbols[0])


class TTSTokenizer:
    """Same constructor and methods as TTS.tts.utils.text.tokenizer.TTSTokenizer, plus ``batch_text_to_ids``."""

    def __init__(self, use_phonemes=False, text_cleaner: Callable = None, characters: BaseVocabulary = None,
                 phonemizer: Optional[Callable] = None, add_blank: bool = False, use_eos_bos=False):
        self.text_cleaner = text_cleaner
        self.use_phonemes = use_phonemes
        self.add_blank = add_blank
        self.use_eos_bos = use_eos_bos
        self.characters = characters
        self.not_found_characters: List[str] = []
        self.phonemizer = phonemizer

    @property
    def characters(self):
        return self._characters

    @characters.setter
    def characters(self, new_characters):
        self._characters = new_characters
        self.pad_id = new_characters.char_to_id(new_characters.pad) if new_characters.pad else None
        self.blank_id = new_characters.char_to_id(new_characters.blank) if new_characters.blank else None
        self._lut = None

    # ------------------------------------------------------------------ the reference's per-sentence API
    def encode(self, text: str) -> List[int]:
        """Characters -> ids; characters outside the vocabulary are discarded and remembered (tokenizer.py:65-78)."""
        ids = []
        for ch in text:
            try:
                ids.append(self.characters.char_to_id(ch))
            except KeyError:
                if ch not in self.not_found_characters:
                    self.not_found_characters.append(ch)
                    print(text)
                    print(f" [!] Character {repr(ch)} not found in the vocabulary. Discarding it.")
        return ids

    def decode(self, token_ids: Sequence[int]) -> str:
        return "".join(self.characters.id_to_char(int(i)) for i in token_ids)

    def _normalise(self, text: str, language: str = None) -> str:
        if self.text_cleaner is not None:
            text = self.text_cleaner(text)
        if self.use_phonemes:
            if self.phonemizer is None:
                raise RuntimeError("tts_b200.TTSTokenizer: use_phonemes=True needs a phonemizer callable "
                                   "(the reference's phonemizer back ends are out of scope)")
            ph = self.phonemizer
            text = ph.phonemize(text, separator="", language=language) if hasattr(ph, "phonemize") else ph(text, language)
        return text

    def text_to_ids(self, text: str, language: str = None) -> List[int]:
        """cleaner -> (phonemizer) -> ids -> blanks -> BOS/EOS (tokenizer.py:87-116)."""
        ids = self.encode(self._normalise(text, language))
        if self.add_blank:
            ids = self.intersperse_blank_char(ids, True)
        if self.use_eos_bos:
            ids = self.pad_with_bos_eos(ids)
        return ids

    def ids_to_text(self, id_sequence: Sequence[int]) -> str:
        return self.decode(id_sequence)

    def pad_with_bos_eos(self, char_sequence: Sequence[int]):
        return [self.characters.bos_id] + list(char_sequence) + [self.characters.eos_id]

    def intersperse_blank_char(self, char_sequence: Sequence[int], use_blank_char: bool = False):
        """``[b, c0, b, c1, ..., b]`` with the blank (or, like the reference, the pad *character*) (tokenizer.py:125-133)."""
        fill = self.characters.blank_id if use_blank_char else self.characters.pad
        out = [fill] * (len(char_sequence) * 2 + 1)
        out[1::2] = char_sequence
        return out

    # ------------------------------------------------------------------ the batch hand-off
    def _lookup_table(self):
        """Code point -> id table (-1 = not in the vocabulary).  Symbols longer than one code point cannot be produced
        by a per-character ``encode`` either, so they are simply unreachable here as they are there."""
        if self._lut is None:
            single = [s for s in self.characters.vocab if len(s) == 1]
            lut = np.full(max((ord(s) for s in single), default=0) + 1, -1, dtype=np.int64)
            for s in single:
                lut[ord(s)] = self.characters.char_to_id(s)    # a duplicated symbol maps like the reference's dict does
            self._lut = lut
        return self._lut

    def batch_text_to_ids(self, texts: Sequence[str], language: str = None, pin_memory: bool = True):
        """``[text_to_ids(t) for t in texts]`` as one padded block: returns (tokens int64 [B, T_max] -- pinned host
        memory when CUDA is present -- and lengths int64 [B]).  Padding uses ``pad_id`` (0 when the vocabulary has
        no pad symbol, like the zero-padded batches the reference's datasets build)."""
        lut = self._lookup_table()
        ids_list = []
        for t in texts:
            t = self._normalise(t, language)
            cp = np.frombuffer(t.encode("utf-32-le"), dtype=np.uint32).astype(np.int64)
            ids = np.where(cp < lut.shape[0], lut[np.minimum(cp, lut.shape[0] - 1)], -1) if cp.size else cp
            if cp.size and (ids < 0).any():
                for ch in {chr(c) for c in cp[ids < 0].tolist()}:
                    if ch not in self.not_found_characters:
                        self.not_found_characters.append(ch)
                        print(t)
                        print(f" [!] Character {repr(ch)} not found in the vocabulary. Discarding it.")
                ids = ids[ids >= 0]
            ids_list.append(ids)
        extra = 2 if self.use_eos_bos else 0
        lens = np.array([(2 * len(i) + 1 if self.add_blank else len(i)) + extra for i in ids_list], dtype=np.int64)
        tmax = int(lens.max()) if len(lens) else 0
        pad = self.pad_id if self.pad_id is not None else 0
        tokens = torch.full((len(texts), tmax), int(pad), dtype=torch.int64)
        if pin_memory and torch.cuda.is_available():
            tokens = tokens.pin_memory()
        view = tokens.numpy()
        for r, ids in enumerate(ids_list):
            off = 1 if self.use_eos_bos else 0
            n = int(lens[r]) - extra
            if self.add_blank:
                view[r, off:off + n] = self.characters.blank_id
                view[r, off + 1:off + n:2] = ids
            else:
                view[r, off:off + n] = ids
            if self.use_eos_bos:
                view[r, 0], view[r, off + n] = self.characters.bos_id, self.characters.eos_id
        return tokens, torch.from_numpy(lens)


# ----------------------------------------------------------------------------- fairseq (MMS) checkpoints
_FAIRSEQ_PREFIXES = (("enc_p.", "text_encoder."), ("dec.", "waveform_decoder."), ("enc_q.", "posterior_encoder."))


def _rename_fairseq_key(k: str) -> str:
    """The key translation of TTS/tts/utils/fairseq.py:4-48: fairseq's VITS interleaves a Flip module after every
    coupling layer (flow.flows.{0,2,4,6} / dp.flows.{1,3,5,7} are the learnable ones) and names the element-wise
    affine's parameters m / logs."""
    for old, new in _FAIRSEQ_PREFIXES:
        if old in k:
            return k.replace(old, new)
    m = re.search(r"flow\.flows\.([246])\.", k)
    if m:
        return k.replace(m.group(0), f"flow.flows.{int(m.group(1)) // 2}.")
    for stack in ("flows", "post_flows"):
        if f"dp.{stack}.0.m" in k:
            return k.replace(f"dp.{stack}.0.m", f"duration_predictor.{stack}.0.translation")
        if f"dp.{stack}.0.logs" in k:
            return k.replace(f"dp.{stack}.0.logs", f"duration_predictor.{stack}.0.log_scale")
        m = re.search(rf"dp\.{stack}\.([1357])", k)
        if m:
            n = int(m.group(1))
            return k.replace(m.group(0), f"duration_predictor.{stack}.{(n + 1) // 2}")
    if "dp." in k:
        return k.replace("dp.", "duration_predictor.")
    return k


def rehash_fairseq_vits_checkpoint(checkpoint_file):
    """fairseq ``G_*.pth`` -> a state dict with this package's (= the reference's) key names."""
    chk = torch.load(checkpoint_file, map_location=torch.device("cpu"), weights_only=False)["model"]
    return {_rename_fairseq_key(k): v for k, v in chk.items()}
