"""Symbol table of the text front-end (reference: models/synthesizer/utils/symbols.py:8-18)."""
_pad = "_"
_eos = "~"
_characters = 'ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz1234567890!\'(),-.:;? '

symbols = [_pad, _eos] + list(_characters)
