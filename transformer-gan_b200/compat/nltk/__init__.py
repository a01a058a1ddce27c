"""Import-time stand-in for `nltk` (absent from this image; the reference's utils/bleu.py imports it at module level,
utils/bleu.py:37-40).  Only the BLEU self-evaluation metric needs it; using that metric without the real package raises.
Lives under ``compat/`` (last on sys.path): a real install always wins."""
