"""text -> symbol ids (reference: models/synthesizer/utils/text.py:13-40 with the ``basic_cleaners``
pipeline of utils/cleaners.py:53-70: lowercase + whitespace collapse; EOS id appended)."""
from __future__ import annotations

import re

from .symbols import symbols

_symbol_to_id = {s: i for i, s in enumerate(symbols)}
_whitespace_re = re.compile(r"\s+")


def basic_cleaners(text: str) -> str:
    return re.sub(_whitespace_re, " ", text.lower())


_CLEANERS = {"basic_cleaners": basic_cleaners}


def text_to_sequence(text, cleaner_names):
    for name in cleaner_names:
        if name not in _CLEANERS:
            raise Exception("Unknown cleaner: %s" % name)
        text = _CLEANERS[name](text)
    seq = [_symbol_to_id[s] for s in text if s in _symbol_to_id and s not in ("_", "~")]
    seq.append(_symbol_to_id["~"])
    return seq
